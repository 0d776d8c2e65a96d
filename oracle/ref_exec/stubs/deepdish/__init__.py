"""deepdish stand-in: dd.io.load of neutral_smpl_meanwjoints.h5 (tester.py:120-121).  h5 files cannot be read here (no h5py):
the generator writes the same two arrays ('pose' 72, 'shape' 10) as a .npz next to the path the reference asks for."""
from . import io      # noqa: F401
