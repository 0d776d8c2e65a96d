"""Seeded process_image cases shared by make_preproc_golden.py (oracle) and make_ref_exec_golden.py (reference source)."""
import numpy as np

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
