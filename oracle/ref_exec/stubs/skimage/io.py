"""skimage.io.imread stand-in (run_video.py:69): RGB uint8 array of an image file, via cv2."""
import cv2


def imread(path):
    img = cv2.imread(path, cv2.IMREAD_COLOR)
    if img is None:
        raise IOError(path)
    return cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
