"""tf.contrib.slim stand-in [TF-ext]: the layer functions are tf.contrib.layers'; arg_scope is only used by the reference around
resnet_v2_50 (models.py:67-74), whose stand-in (nets/resnet_v2.py) applies resnet_arg_scope's settings explicitly."""
import contextlib

from tensorflow.contrib.layers import (batch_norm, conv2d, dropout, fully_connected, l2_regularizer, max_pool2d,     # noqa: F401
                                       variance_scaling_initializer)


@contextlib.contextmanager
def arg_scope(list_ops_or_scope, **kwargs):
    yield list_ops_or_scope
