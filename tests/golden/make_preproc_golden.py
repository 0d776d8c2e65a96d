"""Generates tests/golden/process_image_golden_v1.npz from the cv2 restatement of the reference's process_image
(oracle/preproc_ref.py, run_video.py:56-107).  Inputs are seeded; crops are stored sub-sampled (every 7th pixel) to keep the
fixture small.  Run from the repo root:  python tests/golden/make_preproc_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import preproc_ref  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from preproc_cases import CASES, frame  # noqa: E402


def main():
    out = {'cases': np.array(CASES, np.float64)}
    for i, (H, W, cx, cy, s) in enumerate(CASES):
        r = preproc_ref.process_image(frame(i, H, W), [cx, cy, s])
        assert r['image'].shape == (224, 224, 3), r['image'].shape
        out['img_%d' % i] = r['image'][::7, ::7].astype(np.float32)
        out['meta_%d' % i] = np.array(list(r['center']) + list(r['start_pt']) + list(r['im_shape']), np.int64)
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'process_image_golden_v1.npz'), **out)
    print('wrote', len(CASES), 'cases')


if __name__ == '__main__':
    main()
