"""Generates tests/golden/process_image_golden_v1.npz from the cv2 restatement of the reference's process_image
(oracle/preproc_ref.py, run_video.py:56-107).  Inputs are seeded; crops are stored sub-sampled (every 7th pixel) to keep the
fixture small.  Run from the repo root:  python tests/golden/make_preproc_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import preproc_ref  # noqa: E402

CASES = [  # H, W, cx, cy, scale
    (240, 320, 160.0, 120.0, 1.0),
    (240, 320, 40.3, 200.7, 0.62),        # downscale, crop hangs over the left / bottom edge
    (180, 260, 250.0, 10.0, 1.7),         # upscale, top-right corner
    (360, 200, 100.5, 180.5, 0.5),        # exact 2x downscale (cv2 switches INTER_LINEAR to its area fast path)
    (224, 224, 112.0, 112.0, 1.0),        # identity crop
    (300, 500, 499.0, 299.0, 0.933),
]


def frame(i, H, W):
    rng = np.random.RandomState(1000 + i)
    yy, xx = np.mgrid[0:H, 0:W]
    base = (127 + 80 * np.sin(xx / 9.0 + i) * np.cos(yy / 7.0))[..., None] + rng.randint(-40, 40, size=(H, W, 3))
    return np.clip(base, 0, 255).astype(np.uint8)


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
