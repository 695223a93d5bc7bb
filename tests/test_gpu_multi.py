"""Slab decomposition over >= 2 GPUs against the single-domain oracle, under `pytest -m gpu`.

The check itself is tests/multi_gpu_check.py (one process per GPU under torchrun: drum, periodic
box with migration and wrap-around, polydisperse hopper with a floating wall and an outlet, a solid
surface sweeping through the cuts, and the reference's own 2-rank golden
tests/dem/particle_particle_contact_on_two_processors); this wrapper launches it on two GPUs of
the box and is skipped on a single-GPU box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_decomposition_matches_oracle_on_two_gpus():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs on the box")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=1500)
    sys.stdout.write(r.stdout[-4000:])
    assert r.returncode == 0 and "MULTI_GPU_CHECK PASS" in r.stdout, r.stdout[-4000:] + r.stderr[-2000:]
