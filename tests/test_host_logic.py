"""CPU: host-side logic above the C ABI and the ABI surface itself (no compute calls without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

import rvio_b200  # noqa: F401
from rvio_b200 import synth, host, capi
from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    if not os.path.exists(capi.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    hdr = open(os.path.join(ROOT, "include", "rvio_b200.h")).read()
    declared = set(re.findall(r"\b(rvio_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
    lib = ctypes.CDLL(capi.LIB_PATH)
    for s in capi.SYMBOLS:
        assert hasattr(lib, s), s
    lib.rvio_b200_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.rvio_b200_version()


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cfg = synth.baseline_config(1)
    with pytest.raises(capi.RvioError):
        host.Tracker(cfg)
    with pytest.raises(capi.RvioError):
        host.Updater(cfg)
    with pytest.raises(capi.RvioError):
        host.Vio(cfg)


def test_find_newer_host_matches_oracle():
    r = np.random.default_rng(2)
    for idx in (0, 1, 2, 4):
        cfg = synth.baseline_config(idx)
        for _ in range(5):
            nref, nc = r.integers(0, cfg.n_features), r.integers(0, cfg.n_features)
            ref = np.stack([r.uniform(0, cfg.width, nref), r.uniform(0, cfg.height, nref)], 1).astype(np.float32)
            cand = np.stack([r.uniform(0, cfg.width, nc), r.uniform(0, cfg.height, nc)], 1).astype(np.float32)
            a = host.find_newer(cfg, cand, ref)
            b = orc.find_newer(cfg, cand, ref)
            assert a.shape == b.shape and np.array_equal(a, b), (idx, a.shape, b.shape)


def test_configs_match_baseline_json():
    # N = Lmax-1 clones, F_u = ceil(F/2)
    expect = {0: (150, 10), 1: (200, 11), 2: (600, 25), 4: (2048, 30)}
    for idx, (F, N) in expect.items():
        c = synth.baseline_config(idx)
        assert c.n_features == F and c.window == N
        u = capi.updater_cfg(c)
        assert u.max_clones == N and u.max_features == (F + 1) // 2
    c = synth.baseline_config(2)
    assert (c.width, c.height) == (1280, 720)


def _build_selfcheck(tmp_path):
    import subprocess
    exe = str(tmp_path / "host_selfcheck")
    libdir = os.path.join(ROOT, "r-vio_b200")
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-Wall", "-o", exe, os.path.join(libdir, "host", "host_selfcheck.cpp"),
                           "-L" + libdir, "-lrvio_b200", "-Wl,-rpath," + libdir])
    return exe


def test_cpp_host_adaptor_compiles_and_fails_loudly_without_gpu(tmp_path):
    """r-vio_b200/host/rvio_host.hpp (the C++ twin of host.py: RVIO::Tracker / RVIO::Updater with the reference's member
    names) builds against the C ABI; without a device its constructors throw -- no CPU fallback."""
    import subprocess
    import torch
    if not os.path.exists(capi.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    exe = _build_selfcheck(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    if torch.cuda.is_available():
        assert out.returncode == 0, out.stdout
    else:
        assert out.returncode == 3 and "no device path" in out.stdout


@pytest.mark.gpu
def test_cpp_host_adaptor_runs_on_gpu(tmp_path):
    import subprocess
    exe = _build_selfcheck(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "GPU path ok" in out.stdout, out.stdout + out.stderr
