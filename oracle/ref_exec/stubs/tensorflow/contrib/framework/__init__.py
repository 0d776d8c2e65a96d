"""tf.contrib.framework stand-in."""
import tensorflow as tf


def get_variables(scope=None, suffix=None, collection=None):
    """Variables of the GLOBAL_VARIABLES collection whose name starts with `scope` (a string or a VariableScope)."""
    vs = tf.global_variables()
    if scope is not None:
        prefix = scope if isinstance(scope, str) else scope.name
        if prefix:
            vs = [v for v in vs if v.op_name == prefix or v.op_name.startswith(prefix.rstrip('/') + '/')]
    if suffix is not None:
        vs = [v for v in vs if v.op_name.endswith(suffix)]
    return vs


get_variables_to_restore = get_variables


def get_trainable_variables(scope=None, suffix=None):
    return [v for v in get_variables(scope, suffix) if v.trainable]
