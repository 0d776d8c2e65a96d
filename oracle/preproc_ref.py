"""CPU restatement of the reference's crop pre-processing (TEST INFRASTRUCTURE, like everything under oracle/).

Follows src/evaluation/run_video.py:56-107 (`process_image`) and src/util/common.py:7-14 (`resize_img`) line by line,
with the image handed in as an array instead of read from `im_path` (skimage's imread is not installed here).  cv2 is the
third-party dependency that does the arithmetic (cv2.resize, default INTER_LINEAR, on the float64 image).
"""
import cv2
import numpy as np

IMG_SIZE = 224                                                                      # run_video.py:27


def resize_img(img, scale_factor):
    """src/util/common.py:7-14."""
    new_size = (np.floor(np.array(img.shape[0:2]) * scale_factor)).astype(int)
    new_img = cv2.resize(img, (new_size[1], new_size[0]))
    actual_factor = [new_size[0] / float(img.shape[0]), new_size[1] / float(img.shape[1])]      # [y, x]
    return new_img, actual_factor


def process_image(image, bbox_param):
    """run_video.py:56-107.  image: HxWx3 uint8; bbox_param (3,) = [cx, cy, scale]."""
    bbox_param = np.asarray(bbox_param, np.float64)
    center = bbox_param[:2]
    scale = bbox_param[2]
    image = ((image / 255.) - 0.5) * 2                                              # :72-73
    image_scaled, scale_factors = resize_img(image, scale)                          # :74
    center_scaled = np.round(center * scale_factors).astype(int)                    # :75  (x*fy, y*fx: the reference's own mix-up)
    image_padded = np.pad(array=image_scaled, pad_width=((IMG_SIZE,), (IMG_SIZE,), (0,)), mode='edge')      # :78-82
    height, width = image_padded.shape[:2]
    center_scaled += IMG_SIZE
    margin = IMG_SIZE // 2
    start_pt = (center_scaled - margin).astype(int)
    end_pt = (center_scaled + margin).astype(int)
    end_pt[0] = min(end_pt[0], width)
    end_pt[1] = min(end_pt[1], height)
    image_scaled = image_padded[start_pt[1]:end_pt[1], start_pt[0]:end_pt[0], :]    # :93-94
    center_scaled -= start_pt
    height, width = image_scaled.shape[:2]
    return {'image': image_scaled, 'im_shape': [height, width], 'center': center_scaled, 'scale': scale, 'start_pt': start_pt}
