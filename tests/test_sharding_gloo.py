"""CPU, world_size 2 over gloo: the arithmetic of the two feature-sharded forms, with the CPU oracle standing in
for the device kernels (the NCCL versions with the real kernels are tests/dist_sharded_*.py, run by test_gpu_multi.py):

  * update  -- each rank forms the normal terms [G | z | n_good] of the features f % world == rank; one all-reduce(sum)
               must reproduce the unsharded normal terms (then the replicated solve is the unsharded solve);
  * tracker -- each rank runs LK + undistortion for the feature indices host.shard_range gives it; one all-gather of the
               fixed-size shards must reproduce the unsharded per-feature arrays bit for bit."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import rvio_b200  # noqa: F401
    from rvio_b200 import synth, host
    from oracle import np_updater, oracle as orc
    # ---- update
    cfg = synth.baseline_config(0)
    x, P, types, off, xy = synth.make_update_case(cfg, 24, 77, n_clones=6)
    sigma = float(max(np.float32(cfg.sigma_px), np.float32(cfg.sigma_py)))
    mine = [f for f in range(len(types)) if f % world == rank]
    t_m = np.array([types[f] for f in mine], np.uint8)
    xy_m = np.concatenate([xy[off[f]:off[f + 1]] for f in mine])
    off_m = np.concatenate([[0], np.cumsum([off[f + 1] - off[f] for f in mine])]).astype(np.int32)
    _, _, info = np_updater.update(sigma, cfg.T_BC0, x, P, t_m, off_m, xy_m)
    n = info["G"].shape[0]
    buf = torch.from_numpy(np.concatenate([info["G"].reshape(-1), info["z"], [float(info["n_good"]), float(info["rows"])]]))
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)                       # the single collective of the sharded update
    _, _, full = np_updater.update(sigma, cfg.T_BC0, x, P, types, off, xy)
    G = buf[:n * n].numpy().reshape(n, n); z = buf[n * n:n * n + n].numpy()
    ok_upd = (np.allclose(G, full["G"], rtol=0, atol=1e-9 * max(1.0, np.abs(full["G"]).max())) and
              np.allclose(z, full["z"], rtol=0, atol=1e-9 * max(1.0, np.abs(full["z"]).max())) and
              int(round(float(buf[-2]))) == full["n_good"] and int(round(float(buf[-1]))) == full["rows"])
    # ---- tracker
    cfg2 = synth.baseline_config(0); cfg2.width, cfg2.height, cfg2.n_features = 320, 240, 60
    st = synth.Stream(cfg2, 3, 5, t_static=0.05)
    e0, e1 = orc.clahe(st.frames[1]), orc.clahe(st.frames[2])
    r = np.random.default_rng(3)
    F = 53                                                           # not a multiple of the world size
    pts = np.stack([r.uniform(10, 310, F), r.uniform(10, 230, F)], 1).astype(np.float32)
    lo, hi, S = host.shard_range(F, rank, world)
    px_m, st_m = orc.lk(e0, e1, pts[lo:hi]) if hi > lo else (np.zeros((0, 2), np.float32), np.zeros(0, np.uint8))
    un_m = orc.undistort(px_m, cfg2) if hi > lo else np.zeros((0, 2), np.float32)
    pad = S - (hi - lo)
    mine_t = torch.from_numpy(np.concatenate([np.concatenate([px_m, un_m], 1).view(np.uint8).reshape(hi - lo, 16),
                                              np.zeros((pad, 16), np.uint8)]).reshape(-1))
    mine_s = torch.from_numpy(np.concatenate([st_m, np.zeros(pad, np.uint8)]))
    all_t = torch.empty(world * S * 16, dtype=torch.uint8); all_s = torch.empty(world * S, dtype=torch.uint8)
    dist.all_gather_into_tensor(all_t, mine_t); dist.all_gather_into_tensor(all_s, mine_s)   # equal-sized shards
    got = all_t.numpy().reshape(world * S, 16)[:F].copy().view(np.float32).reshape(F, 4)
    px_f, st_f = orc.lk(e0, e1, pts)
    un_f = orc.undistort(px_f, cfg2)
    ok_trk = (np.array_equal(got[:, :2].view(np.uint32), px_f.view(np.uint32)) and
              np.array_equal(got[:, 2:].view(np.uint32), un_f.view(np.uint32)) and
              np.array_equal(all_s.numpy()[:F], st_f))
    q.put((rank, bool(ok_upd), bool(ok_trk), int(full["n_good"])))
    dist.destroy_process_group()


def test_sharded_forms_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, 29641, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(2))
    for p in ps:
        p.join(60)
    for rank, ok_upd, ok_trk, n_good in res:
        assert ok_upd, f"rank {rank}: reduced normal terms differ from the unsharded ones"
        assert ok_trk, f"rank {rank}: gathered LK arrays differ from the unsharded ones"
        assert n_good >= 3
