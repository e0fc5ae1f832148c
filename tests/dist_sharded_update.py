"""Feature-sharded MSCKF update across GPUs (BASELINE configs[4]): every rank forms the normal terms of its
share of the features (f % world == rank) on its own GPU, ONE ncclAllReduce sums [G | z | counters] over NVLink, every rank
finishes the (replicated) solve.  Launched by torchrun; rank 0 compares against the unsharded update and the CPU oracle.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/dist_sharded_update.py
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


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = synth.baseline_config(4)                       # 2048 features, 30 clones
    n_feat = int(os.environ.get("RVIO_TEST_FEATS", "256"))
    x, P, types, off, xy = synth.make_update_case(cfg, n_feat, 31)
    L = capi.lib()
    upd = host.Updater(cfg, local)
    d = P.shape[0]; n = d - 24
    Pc = np.ascontiguousarray(P.T)
    xyf = np.ascontiguousarray(xy).reshape(-1)
    cudart = C.cdll.LoadLibrary("libcudart.so.12")
    stream = torch.cuda.ExternalStream(L.rvio_updater_stream(upd.h), device=torch.device("cuda", local))
    count = n * n + n + 8 + n + 1          # [G | z | counters | per-class information]
    buf = torch.empty(count, dtype=torch.float64, device="cuda")
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    times = []
    for it in range(4):
        dist.barrier(); torch.cuda.synchronize()
        ev0.record()
        capi.check(L.rvio_updater_update_begin(upd.h, x, len(x), Pc, d, types, off, xyf, len(types), rank, world))
        ptr, cnt = C.c_void_p(), C.c_int()
        capi.check(L.rvio_updater_reduce_buffer(upd.h, C.byref(ptr), C.byref(cnt)))
        assert cnt.value == count
        stream.synchronize()
        cudart.cudaMemcpy(C.c_void_p(buf.data_ptr()), ptr, C.c_size_t(8 * count), 3)
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)        # the single collective of the sharded update (<= 261 KB)
        torch.cuda.synchronize()
        cudart.cudaMemcpy(ptr, C.c_void_p(buf.data_ptr()), C.c_size_t(8 * count), 3)
        xo = np.empty_like(x); Po = np.empty_like(Pc); info = capi.UpdateInfo()
        capi.check(L.rvio_updater_update_finish(upd.h, xo, Po, C.byref(info)))
        ev1.record(); torch.cuda.synchronize()
        times.append(ev0.elapsed_time(ev1))
    # every rank must hold the same posterior
    chk = torch.tensor([float(np.abs(xo).sum()), float(np.abs(Po).sum())], dtype=torch.float64, device="cuda")
    lo = chk.clone(); hi = chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert torch.equal(lo, hi), "ranks disagree on the posterior"
    if rank == 0:
        xr, Pr = upd.update(x, P, types, (off, xy))      # unsharded on one GPU
        assert info.n_good == upd.info.n_good and info.updated == 1
        np.testing.assert_allclose(xo, xr, rtol=0, atol=1e-10)
        np.testing.assert_allclose(Po.T, Pr, rtol=0, atol=1e-10 * np.abs(Pr).max())
        msg = f"sharded update ok: world={world} feats={n_feat} N={cfg.window} good={info.n_good} rows={info.rows_stacked} ms/update={min(times):.3f}"
        if os.environ.get("RVIO_TEST_ORACLE", "1") == "1":
            from oracle import oracle as orc
            xc, Pcpu, oi = orc.updater_update(cfg, x, P, types, off, xy)       # the reference's rule
            assert (info.rank == oi.rank or info.rank_flags & 16) and not (info.rank_flags & 4), (info.rank, oi.rank, info.rank_flags)
            np.testing.assert_allclose(xo, xc, rtol=0, atol=1e-9)
            msg += " (matches CPU oracle)"
        print(msg)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
