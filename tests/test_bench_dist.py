"""CPU: the N>1 plumbing of bench.py (one process per GPU, max-over-ranks timing, rank 0 reports) exercised with the
gloo backend at world_size 2."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # same reduction bench.py performs on [t_dev, t_e2e]
    t = torch.tensor([0.10 + 0.05 * rank, 0.20 - 0.03 * rank], dtype=torch.float64)
    dist.barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    frames = 100 * world / float(t[0])
    q.put((rank, float(t[0]), float(t[1]), frames))
    dist.destroy_process_group()


def test_max_over_ranks_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, 29631, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in ps:
        p.join(60)
    for rank, tdev, te2e, frames in res:
        assert abs(tdev - 0.15) < 1e-12 and abs(te2e - 0.20) < 1e-12
        assert abs(frames - 200 / 0.15) < 1e-9


def test_reference_arm_nonzero_ranks_exit_quietly():
    import subprocess
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "4", "--warmup", "3"],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == ""
