"""Crop pre-processing in front of the Tester -- drop-in for `process_image` of the reference's
src/evaluation/run_video.py:56-107.  The arithmetic (scale to [-1,1], cv2-convention bilinear resize, edge pad, crop) runs
in one CUDA kernel on the uint8 frame (human_dynamics_b200.preprocess / hd_process_image); rendering is out of scope."""
import numpy as np
import torch

from human_dynamics_b200.preprocess import IMG_SIZE, crop_geometry, process_images


def _read_rgb(im_path):
    import cv2                                  # decoding only (the reference uses skimage.io.imread, RGB order)
    bgr = cv2.imread(im_path, cv2.IMREAD_COLOR)
    if bgr is None:
        raise IOError('cannot read image %s' % im_path)
    return np.ascontiguousarray(bgr[:, :, ::-1])


def process_image(im_path, bbox_param):
    """Processes an image, producing 224x224 crop.

    Args:
        im_path (str | HxWx3 uint8 array).
        bbox_param (3,): [cx, cy, scale].

    Returns:
        dict: image, im_path, im_shape, center, scale, start_pt   (run_video.py:99-107).
    """
    image = im_path if isinstance(im_path, np.ndarray) else _read_rgb(im_path)
    crops, geoms = process_images(image[None], np.asarray(bbox_param, np.float64).reshape(1, 3))
    torch.cuda.current_stream().synchronize()
    g = geoms[0]
    return {'image': crops[0].cpu().numpy(), 'im_path': im_path if isinstance(im_path, str) else None, 'im_shape': g['im_shape'],
            'center': g['center'], 'scale': g['scale'], 'start_pt': g['start_pt']}


def process_video_frames(frames, bbox_params):
    """Batched form for a whole track: frames (N,H,W,3) uint8 (host or CUDA), bbox_params (N,3) ->
    (crops (N,224,224,3) float32 CUDA -- feed `Tester.predict_all_images` / `HMMREngine.encode_images` directly --, infos)."""
    return process_images(frames, bbox_params, IMG_SIZE)


__all__ = ['process_image', 'process_video_frames', 'crop_geometry', 'IMG_SIZE']
