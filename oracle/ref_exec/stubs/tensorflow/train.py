"""tf.train stand-in: Saver.restore assigns variables BY NAME from a checkpoint and fails on a missing key, like TF."""
import importlib.util
import os

import numpy as np

_CKPT_READER = None


def _reader():
    """The repo's TensorFlow-V2-bundle reader, loaded by file path (no package import, no CUDA library)."""
    global _CKPT_READER
    if _CKPT_READER is None:
        here = os.path.dirname(os.path.abspath(__file__))
        root = os.path.abspath(os.path.join(here, '..', '..', '..', '..'))
        spec = importlib.util.spec_from_file_location('_hd_tf_checkpoint', os.path.join(root, 'human_dynamics_b200', 'tf_checkpoint.py'))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _CKPT_READER = mod
    return _CKPT_READER


class Saver(object):
    def __init__(self, var_list=None, **kw):
        import tensorflow as tf
        self.var_list = list(var_list) if var_list is not None else tf.global_variables()

    def restore(self, sess, save_path):
        if os.path.exists(save_path + '.index'):
            tensors = _reader().load_checkpoint(save_path)
        elif os.path.exists(save_path) and save_path.endswith('.npz'):
            tensors = dict(np.load(save_path))
        else:
            raise IOError('checkpoint %s not found' % save_path)
        for v in self.var_list:
            key = v.op_name
            if key not in tensors:
                raise KeyError('NotFoundError: Key %s not found in checkpoint' % key)
            v.load(tensors[key])
        self.restored = [v.op_name for v in self.var_list]

    def save(self, *a, **kw):
        raise NotImplementedError
