"""TensorFlow V2 checkpoint -> .npz of TF-named variables (no TensorFlow needed).

  python tools/ckpt_to_npz.py /path/hmmr_model.ckpt-1119816 hmmr.npz [--resnet /path/hmr_noS5.ckpt-642561] [--list]

Keeps what the inference graph restores (src/evaluation/tester.py:92-116,163-167): everything except discriminator
(`D_*`) variables, optimizer slots and step counters; `--resnet` overlays the `resnet_v2_50/*` variables of a second
checkpoint the way `Tester(config, pretrained_resnet_path=...)` does (:99-109).
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from human_dynamics_b200 import tf_checkpoint  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('prefix')
    ap.add_argument('out', nargs='?')
    ap.add_argument('--resnet', default='')
    ap.add_argument('--list', action='store_true')
    a = ap.parse_args()
    if a.list:
        _, entries = tf_checkpoint.read_index(a.prefix + '.index')
        for n in sorted(entries):
            print('%-90s %s dtype=%d' % (n, entries[n].shape, entries[n].dtype))
        return 0
    w = tf_checkpoint.load_checkpoint(a.prefix)
    if a.resnet:
        rw = tf_checkpoint.load_checkpoint(a.resnet, names=lambda n: n.startswith('resnet_v2_50') and not n.endswith(('Adam', 'Adam_1')))
        w.update(rw)
    if not a.out:
        ap.error('output .npz path required')
    np.savez(a.out, **w)
    print('wrote %d variables (%.1f MB) to %s' % (len(w), sum(v.nbytes for v in w.values()) / 1e6, a.out))
    return 0


if __name__ == '__main__':
    sys.exit(main())
