"""Feature-sharded tracking across GPUs (BASELINE configs[4]): every rank holds the same tracker state and sees
the same frame, runs the pyramidal LK for its share of the feature indices, ONE all-gather per per-feature array (pixels,
normalized coordinates, status: <= 17 B / feature) over NCCL completes the arrays, RANSAC + bookkeeping run replicated.
Rank 0 compares every frame, bit for bit, against an unsharded tracker on the same GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/dist_sharded_tracker.py
"""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rvio_b200  # noqa: E402,F401
from rvio_b200 import synth, host, capi  # noqa: E402
from oracle import oracle as orc  # noqa: E402  (detector = cv2 goodFeaturesToTrack + cornerSubPix, identical on every rank)


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = synth.baseline_config(1)
    cfg.n_features = int(os.environ.get("RVIO_TEST_FEATS", "300"))
    cfg.min_dist = 8
    n_frames = int(os.environ.get("RVIO_TEST_FRAMES", "24"))
    st = synth.Stream(cfg, n_frames, 20260926, t_static=0.25)
    det = lambda img, n, s: orc.detect_with_subpix(img, n, s, cfg)
    trk = host.Tracker(cfg, local, det)
    ref = host.Tracker(cfg, local, det) if rank == 0 else None
    cudart = C.cdll.LoadLibrary("libcudart.so.12")
    lo, hi, S = host.shard_range(cfg.n_features, rank, world)
    bufs = {}

    def exchange(lk, un, stat, shard):
        assert shard == S
        for name, ptr, item in (("lk", lk, 8), ("un", un, 8), ("st", stat, 1)):
            mine = bufs.setdefault(name, torch.empty(S * item, dtype=torch.uint8, device="cuda"))
            full = bufs.setdefault(name + "_all", torch.empty(world * S * item, dtype=torch.uint8, device="cuda"))
            cudart.cudaMemcpy(C.c_void_p(mine.data_ptr()), C.c_void_p(ptr + lo * item), C.c_size_t(S * item), 3)
            dist.all_gather_into_tensor(full, mine)
            torch.cuda.synchronize()
            cudart.cudaMemcpy(C.c_void_p(ptr), C.c_void_p(full.data_ptr()), C.c_size_t(world * S * item), 3)

    consumed = 0
    checked = emitted = 0
    for i in range(n_frames):
        imu, consumed = st.imu_for_frame(i, consumed)
        if len(imu) < 2:
            continue
        rc = trk.track(st.frames[i], imu, shard=(rank, world, exchange))
        t, o, xy = trk.update_lists() if rc == capi.OK else (np.zeros(0, np.uint8), np.zeros(1, np.int32), np.zeros((0, 2), np.float32))
        # all ranks must agree on the emitted lists
        sig = torch.tensor([float(len(t)), float(o[-1]), float(np.asarray(xy, np.float64).sum())], dtype=torch.float64, device="cuda")
        a = sig.clone(); b = sig.clone()
        dist.all_reduce(a, op=dist.ReduceOp.MIN); dist.all_reduce(b, op=dist.ReduceOp.MAX)
        assert torch.equal(a, b), f"frame {i}: ranks disagree"
        if rank == 0:
            rc0 = ref.track(st.frames[i], imu)
            assert rc0 == rc, (i, rc0, rc)
            if rc == capi.OK:
                d0, d1 = ref.debug(), trk.debug()
                for k in ("status", "flags"):
                    assert np.array_equal(d0[k], d1[k]), (i, k)
                for k in ("lk", "un"):
                    assert np.array_equal(np.asarray(d0[k], np.float32).view(np.uint32), np.asarray(d1[k], np.float32).view(np.uint32)), (i, k)
                t0, o0, xy0 = ref.update_lists()
                assert np.array_equal(t0, t) and np.array_equal(o0, o)
                assert np.array_equal(np.asarray(xy0, np.float32).view(np.uint32), np.asarray(xy, np.float32).view(np.uint32))
                checked += 1; emitted += len(t)
    if rank == 0:
        assert checked >= n_frames // 2
        print(f"sharded tracker ok: world={world} feats={cfg.n_features} shard={S} frames_checked={checked} emitted={emitted}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
