"""N-GPU == 1-GPU, bit for bit (needs >= 2 GPUs: run with `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('world', [2, 4, 8])
def test_sharded_run_is_bit_identical_to_single_gpu(world):
    if torch.cuda.device_count() < world:
        pytest.skip('needs %d GPUs, have %d' % (world, torch.cuda.device_count()))
    # ragged shards, and few enough poses (< 256) that every rank and the 1-GPU run take the same SMPL kernel (the tensor-core
    # blend path for large batches differs from the fused small-batch kernel in the last bits)
    env = dict(os.environ, HD_MGPU_CLIPS=str(min(2 * world + 1, 12)))
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world), '--master-addr', '127.0.0.1',
           '--master-port', str(29500 + world), os.path.join(ROOT, 'tests', '_mgpu_worker.py')]
    r = subprocess.run(cmd, env=env, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    print(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-4000:]
    assert 'bit-identical to the 1-GPU run' in r.stdout
