"""tf.contrib.layers stand-in  [TF-ext]: restated from the TF 1.x contrib definitions (variable names, defaults, arithmetic).

  conv2d / fully_connected : variables `<scope>/weights` (HWIO / [in,out]) and `<scope>/biases`; default activation ReLU;
                             with normalizer_fn=batch_norm there are no biases and BN lives under `<scope>/BatchNorm`.
  batch_norm               : `<scope>/{beta,gamma,moving_mean,moving_variance}`, inference form
                             y = (x - moving_mean) * rsqrt(moving_variance + eps) * gamma + beta, then activation_fn.
  group_norm               : `<scope>/{beta,gamma}`, groups=32, epsilon=1e-6; moments over reduction_axes and the
                             within-group channel axis, gain = rsqrt(var + eps) * gamma, offset = beta - mean * gain.
"""
import numpy as np

import tensorflow as tf
from tensorflow import nn


def _pair(v):
    return (int(v), int(v)) if isinstance(v, (int, np.integer)) else (int(v[0]), int(v[1]))


def batch_norm(inputs, decay=0.999, center=True, scale=False, epsilon=0.001, activation_fn=None, is_training=True, reuse=None,
               scope=None, **kw):
    if is_training:
        raise NotImplementedError('stand-in is inference only')
    with tf.variable_scope(scope, 'BatchNorm', [inputs], reuse=reuse):
        C = inputs.shape[-1].value
        beta = tf.get_variable('beta', [C]) if center else None
        gamma = tf.get_variable('gamma', [C]) if scale else None
        mean = tf.get_variable('moving_mean', [C], trainable=False)
        var = tf.get_variable('moving_variance', [C], trainable=False)
    eps = np.float32(epsilon)
    inv = tf.rsqrt(var + eps)
    if gamma is not None:
        inv = inv * gamma
    out = inputs * inv + ((beta - mean * inv) if beta is not None else (-mean * inv))
    return activation_fn(out) if activation_fn is not None else out


def group_norm(inputs, groups=32, channels_axis=-1, reduction_axes=(-3, -2), center=True, scale=True, epsilon=1e-6,
               activation_fn=None, reuse=None, scope=None, **kw):
    shp = inputs.shape.as_list()
    nd = len(shp)
    if channels_axis not in (-1, nd - 1):
        raise NotImplementedError('channels_axis')
    red = sorted(a % nd for a in reduction_axes)
    C = shp[-1]
    if C % groups:
        raise ValueError('channels %d not divisible by groups %d' % (C, groups))
    with tf.variable_scope(scope, 'GroupNorm', [inputs], reuse=reuse):
        beta = tf.get_variable('beta', [C]) if center else None
        gamma = tf.get_variable('gamma', [C]) if scale else None
    x = tf.reshape(inputs, shp[:-1] + [groups, C // groups])
    mean, var = nn.moments(x, red + [nd], keep_dims=True)          # reduction axes + the within-group channel axis
    bshape = [1] * (nd - 1) + [groups, C // groups]
    gain = tf.rsqrt(var + np.float32(epsilon))
    offset = -mean * gain
    if gamma is not None:
        g = tf.reshape(gamma, bshape)
        gain = gain * g
        offset = offset * g
    if beta is not None:
        offset = offset + tf.reshape(beta, bshape)
    out = tf.reshape(x * gain + offset, shp)
    return activation_fn(out) if activation_fn is not None else out


def conv2d(inputs, num_outputs, kernel_size, stride=1, padding='SAME', data_format=None, rate=1, activation_fn=nn.relu,
           normalizer_fn=None, normalizer_params=None, weights_initializer=None, weights_regularizer=None, biases_initializer='zeros',
           reuse=None, scope=None, **kw):
    if data_format not in (None, 'NHWC') or rate != 1:
        raise NotImplementedError
    kh, kw_ = _pair(kernel_size)
    s = _pair(stride)
    if s[0] != s[1]:
        raise NotImplementedError
    with tf.variable_scope(scope, 'Conv', [inputs], reuse=reuse):
        cin = inputs.shape[-1].value
        w = tf.get_variable('weights', [kh, kw_, cin, num_outputs])
        out = nn.conv2d(inputs, w, [1, s[0], s[1], 1], padding)
        if normalizer_fn is not None:
            out = normalizer_fn(out, **(normalizer_params or {}))
        elif biases_initializer is not None:
            out = nn.bias_add(out, tf.get_variable('biases', [num_outputs]))
        if activation_fn is not None:
            out = activation_fn(out)
    return out


convolution2d = conv2d


def fully_connected(inputs, num_outputs, activation_fn=nn.relu, normalizer_fn=None, normalizer_params=None, weights_initializer=None,
                    weights_regularizer=None, biases_initializer='zeros', reuse=None, scope=None, **kw):
    inputs = tf.convert_to_tensor(inputs)
    with tf.variable_scope(scope, 'fully_connected', [inputs], reuse=reuse):
        cin = inputs.shape[-1].value
        w = tf.get_variable('weights', [cin, num_outputs])
        out = tf.matmul(inputs, w)                              # rank > 2: contraction over the last axis (tensordot in TF)
        if normalizer_fn is not None:
            out = normalizer_fn(out, **(normalizer_params or {}))
        elif biases_initializer is not None:
            out = nn.bias_add(out, tf.get_variable('biases', [num_outputs]))
        if activation_fn is not None:
            out = activation_fn(out)
    return out


def max_pool2d(inputs, kernel_size, stride=2, padding='VALID', scope=None, **kw):
    k, s = _pair(kernel_size), _pair(stride)
    return nn.max_pool(inputs, [1, k[0], k[1], 1], [1, s[0], s[1], 1], padding)


def dropout(inputs, keep_prob=0.5, is_training=True, scope=None, **kw):
    if is_training:
        raise NotImplementedError('stand-in is inference only')
    return inputs


def l2_regularizer(scale, scope=None):
    return None


def variance_scaling_initializer(factor=2.0, mode='FAN_IN', uniform=False, seed=None, dtype=None):
    return ('variance_scaling', factor, mode, uniform)


def xavier_initializer(*a, **kw):
    return ('xavier',)
