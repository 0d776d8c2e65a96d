"""tf.nn stand-in."""
import numpy as np

import tensorflow as tf
from . import _kernels


def relu(x, name=None):
    return tf._op(lambda a: np.maximum(a, np.zeros((), a.dtype)), [tf.convert_to_tensor(x)], 'relu')


def conv2d(input, filter, strides, padding, name=None, **kw):     # noqa: A002
    s = strides[1]
    return tf._op(lambda a, w: _kernels.conv2d_nhwc(a, w, s, padding), [tf.convert_to_tensor(input), tf.convert_to_tensor(filter)], 'conv2d')


def bias_add(value, bias, name=None):
    return tf.add(value, bias)


def max_pool(value, ksize, strides, padding, name=None, **kw):
    return tf._op(lambda a: _kernels.max_pool_nhwc(a, (ksize[1], ksize[2]), strides[1], padding), [tf.convert_to_tensor(value)], 'max_pool')


def moments(x, axes, keep_dims=False, name=None):
    """mean, then variance = mean(squared_difference(x, mean)) (two passes, as tf.nn.moments)."""
    x = tf.convert_to_tensor(x)
    ax = tuple(axes)
    mean = tf._op(lambda a: np.mean(a, axis=ax, keepdims=True).astype(a.dtype), [x], 'mean')
    var = tf._op(lambda a, m: np.mean(np.square(a - m), axis=ax, keepdims=True).astype(a.dtype), [x, mean], 'variance')
    if not keep_dims:
        mean, var = tf.squeeze(mean, list(ax)), tf.squeeze(var, list(ax))
    return mean, var


def dropout(x, keep_prob, **kw):
    raise NotImplementedError('inference only')
