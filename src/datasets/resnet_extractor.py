"""Extracts image features from a sequence of images -- drop-in for src/datasets/resnet_extractor.py:13-98."""
import numpy as np
import torch

from human_dynamics_b200 import runtime as _rt
from human_dynamics_b200.config import HMMRConfig
from human_dynamics_b200.engine import load_weights
from human_dynamics_b200.nets import PackedResNet, ResNetPlan


class FeatureExtractor(object):
    def __init__(self, model_path, img_size=224, batch_size=64, sess=None, impl='auto'):
        self.model_path = model_path
        self.img_size = img_size
        self.batch_size = batch_size
        w = load_weights(model_path)
        self.device = torch.device('cuda', torch.cuda.current_device())
        self.packed = PackedResNet(w, self.device, tc=(impl if impl != 'simt' else False))
        self.plan = ResNetPlan(self.packed, batch_size, img_size, impl)
        self.phis = torch.empty((batch_size, self.packed.out_dim), dtype=torch.float32, device=self.device)

    def compute_phis(self, images):
        """images (B x H x W x 3) -> phis (B x 2048)   (resnet_extractor.py:58-72)."""
        if tuple(images.shape) != (self.batch_size, self.img_size, self.img_size, 3):
            raise ValueError('images must be %s' % ((self.batch_size, self.img_size, self.img_size, 3),))
        x = torch.as_tensor(np.ascontiguousarray(images, dtype=np.float32)).to(self.device)
        self.plan.run(x, self.phis)
        return self.phis.cpu().numpy()

    def compute_all_phis(self, all_images):
        """all_images (T x H x W x 3) -> (T x 2048); the last partial batch is zero-padded (:88-92)."""
        all_phis = []
        T = len(all_images)
        for i in range(0, T, self.batch_size):
            images = np.asarray(all_images[i:i + self.batch_size], np.float32)
            if len(images) < self.batch_size:
                pad = np.zeros((self.batch_size - len(images), self.img_size, self.img_size, 3), np.float32)
                images = np.vstack((images, pad))
            all_phis.append(self.compute_phis(images))
        return np.vstack(all_phis)[:T]
