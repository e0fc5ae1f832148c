"""GPU (>= 2 devices): feature-sharded update with one NCCL all-reduce (tests/dist_sharded_update.py under torchrun)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_update_nccl():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (covered on one GPU by test_updater_feature_sharding_matches_unsharded)")
    world = 2 if n < 4 else 4
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", os.path.join(ROOT, "tests", "dist_sharded_update.py")],
                         capture_output=True, text=True, timeout=600, env=dict(os.environ, RVIO_TEST_FEATS="192"))
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "sharded update ok" in out.stdout


def test_sharded_tracker_nccl():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                          "--master-port", "29534", os.path.join(ROOT, "tests", "dist_sharded_tracker.py")],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "sharded tracker ok" in out.stdout



def test_sharded_vio_end_to_end_nccl():
    """The whole frame feature-sharded through rvio_vio_step with the library's own in-stream NCCL collectives."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                          "--master-port", "29535", os.path.join(ROOT, "tests", "dist_sharded_vio.py")],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "sharded vio ok" in out.stdout
    print(out.stdout[-600:])
