"""tf.contrib.slim.nets.resnet_v2 stand-in [TF-ext]: resnet_v2_50 as published in TF 1.x contrib/slim (He et al. 2016 pre-activation
bottlenecks), restated -- the reference only calls it (models.py:67-74); slim itself is not part of /root/reference.

  root      : conv2d_same(64, 7x7, stride 2) with biases, no BN / activation  ->  max_pool2d 3x3 stride 2, SAME
  blocks    : (64,3,2) (128,4,2) (256,6,2) (512,3,1); units 1..n-1 stride 1, the LAST unit of a block carries the stride
  bottleneck: preact = relu(BN(x));  shortcut = x subsampled (max_pool 1x1, stride) if depth_in == depth
              else conv1x1(preact, depth, stride) + bias;  residual = conv1x1(preact, base)+BN+relu ->
              conv2d_same(3x3, base, stride)+BN+relu -> conv1x1(depth) + bias;  out = shortcut + residual
  tail      : postnorm BN + relu, global mean over H, W (keep_dims) when global_pool
  conv2d_same: stride 1 -> SAME; else explicit zero pad (k-1)//2 before, (k-1) - (k-1)//2 after, then VALID
  resnet_arg_scope: BN epsilon 1e-5, scale=True; max_pool2d padding SAME
Variable names: resnet_v2_50/conv1/{weights,biases}, resnet_v2_50/block{b}/unit_{u}/bottleneck_v2/{preact/*, shortcut/*, conv1/*,
conv1/BatchNorm/*, conv2/*, conv2/BatchNorm/*, conv3/*}, resnet_v2_50/postnorm/*.
"""
import tensorflow as tf
from tensorflow import nn
from tensorflow.contrib import layers

_BN = dict(epsilon=1e-5, scale=True, center=True)


def resnet_arg_scope(weight_decay=0.0001, batch_norm_decay=0.997, batch_norm_epsilon=1e-5, batch_norm_scale=True, **kw):
    if batch_norm_epsilon != 1e-5 or not batch_norm_scale:
        raise NotImplementedError
    return {'resnet_arg_scope': True}


def _conv2d_same(x, cout, k, stride, is_training, bn, scope):
    kw = dict(normalizer_fn=layers.batch_norm, normalizer_params=dict(is_training=is_training, **_BN), activation_fn=nn.relu) if bn \
        else dict(normalizer_fn=None, activation_fn=None)
    if stride == 1:
        return layers.conv2d(x, cout, k, stride=1, padding='SAME', scope=scope, **kw)
    total = k - 1
    beg = total // 2
    x = tf.pad(x, [[0, 0], [beg, total - beg], [beg, total - beg], [0, 0]])
    return layers.conv2d(x, cout, k, stride=stride, padding='VALID', scope=scope, **kw)


def _subsample(x, factor):
    return x if factor == 1 else layers.max_pool2d(x, [1, 1], stride=factor, padding='SAME')


def bottleneck(x, depth, depth_bottleneck, stride, is_training, scope=None):
    with tf.variable_scope(scope, 'bottleneck_v2', [x]):
        depth_in = x.shape[-1].value
        bn = dict(is_training=is_training, **_BN)
        preact = layers.batch_norm(x, activation_fn=nn.relu, scope='preact', **bn)
        if depth == depth_in:
            shortcut = _subsample(x, stride)
        else:
            shortcut = layers.conv2d(preact, depth, [1, 1], stride=stride, normalizer_fn=None, activation_fn=None, scope='shortcut')
        r = layers.conv2d(preact, depth_bottleneck, [1, 1], stride=1, normalizer_fn=layers.batch_norm, normalizer_params=bn,
                          activation_fn=nn.relu, scope='conv1')
        r = _conv2d_same(r, depth_bottleneck, 3, stride, is_training, True, 'conv2')
        r = layers.conv2d(r, depth, [1, 1], stride=1, normalizer_fn=None, activation_fn=None, scope='conv3')
        return shortcut + r


def resnet_v2_50(inputs, num_classes=None, is_training=True, global_pool=True, output_stride=None, spatial_squeeze=True, reuse=None,
                 scope='resnet_v2_50'):
    if num_classes is not None or output_stride is not None:
        raise NotImplementedError
    blocks = (('block1', 64, 3, 2), ('block2', 128, 4, 2), ('block3', 256, 6, 2), ('block4', 512, 3, 1))
    end_points = {}
    with tf.variable_scope(scope, 'resnet_v2', [inputs], reuse=reuse):
        net = _conv2d_same(inputs, 64, 7, 2, is_training, False, 'conv1')
        net = layers.max_pool2d(net, [3, 3], stride=2, padding='SAME', scope='pool1')
        for name, base, units, stride in blocks:
            with tf.variable_scope(name, 'block', [net]):
                for u in range(1, units + 1):
                    with tf.variable_scope('unit_%d' % u, values=[net]):
                        net = bottleneck(net, 4 * base, base, stride if u == units else 1, is_training)
            end_points[scope + '/' + name] = net
        net = layers.batch_norm(net, activation_fn=nn.relu, scope='postnorm', is_training=is_training, **_BN)
        if global_pool:
            net = tf.reduce_mean(net, [1, 2], name='pool5', keep_dims=True)
            end_points['global_pool'] = net
    return net, end_points
