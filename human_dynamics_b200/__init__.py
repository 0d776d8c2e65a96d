"""B200-native HMMR video->SMPL hot path (ResNet-v2-50 -> f_movie -> IEF -> SMPL -> projection).

The product is libhd_b200.so (hand-written sm_100a kernels behind the C-ABI of include/hd_b200.h);
this package is its Python host side.  Importing it needs the built library; running it needs a B200.
"""
from .config import HMMRConfig  # noqa: F401

__all__ = ['HMMRConfig']
