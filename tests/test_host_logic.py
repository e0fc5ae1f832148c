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


def test_io_formats_roundtrip(tmp_path):
    """EuRoC ASL layout in, stamped_pose_ests.dat / time_cost.dat out.  A synthetic stream written as ASL
    and read back must reproduce the frames and, image by image, the IMU rows InputBuffer::GetMeasurements hands out
    (InputBuffer.cc:53-81), with dt = stamp difference to the previous message (rvio_mono.cc:103-108)."""
    from rvio_b200 import io_formats
    cfg = synth.baseline_config(0)
    cfg.width, cfg.height = 160, 120
    st = synth.Stream(cfg, 6, 5, t_static=0.1)
    io_formats.write_asl(str(tmp_path), st.frame_t, st.frames, st.imu)
    rd = io_formats.EurocAslReader(str(tmp_path), cfg.time_offset)
    assert len(rd) == 6
    consumed = 0
    seen = 0
    frames = {round(t, 6): (im, rows) for t, im, rows in rd}
    for i in range(6):
        imu, consumed = st.imu_for_frame(i, consumed)
        if len(imu) < 2:
            continue
        im, rows = frames[round(float(st.frame_t[i]), 6)]
        assert np.array_equal(im, st.frames[i])
        assert rows.shape == (len(imu), 8)
        np.testing.assert_allclose(rows[:, :6], imu[:, :6], rtol=0, atol=0)      # repr() round-trips doubles exactly
        np.testing.assert_allclose(rows[:, 6], imu[:, 6], rtol=0, atol=1e-9)     # stamps are integer nanoseconds
        seen += 1
    assert seen >= 4
    w = io_formats.PoseWriter(str(tmp_path / "out"))
    w.write(1403715273.262142976, [0.1, -0.2, 0.3, 0, 0, 0.6, 0.8], 1, 0.25, 0.125)
    w.close()
    p = io_formats.read_pose_file(str(tmp_path / "out" / "stamped_pose_ests.dat"))
    assert p.shape == (1, 8) and p[0, 0] == 1403715273.262142976 and p[0, 7] == 0.8
    assert open(tmp_path / "out" / "time_cost.dat").read().split() == ["1", "0.25", "0.125"]


def test_config_from_reference_yaml(tmp_path):
    """The YAML keys are the reference's (config/rvio_euroc.yaml); written here from the defaults, not read from it."""
    y = tmp_path / "cfg.yaml"
    y.write_text("%YAML:1.0\nIMU.dps: 100\nCamera.Fisheye: 1\nCamera.fx: 300.5\nTracker.nFeatures: 321\n"
                 "Tracker.nMaxTrackingLength: 9\nCamera.T_BC0: !!opencv-matrix\n    rows: 4\n    cols: 4\n    dt: d\n"
                 "    data: [1, 0, 0, 0.1, 0, 1, 0, 0.2, 0, 0, 1, 0.3, 0, 0, 0, 1]\n")
    c = synth.Config.from_yaml(str(y))
    assert (c.imu_rate, c.fisheye, c.fx, c.n_features, c.max_track_len) == (100.0, 1, 300.5, 321, 9)
    assert c.T_BC0[3] == 0.1 and len(c.T_BC0) == 16 and c.window == 8


def test_shard_range_covers_all_features():
    for F, world in [(200, 2), (200, 3), (2048, 8), (150, 4), (7, 8)]:
        seen = []
        for r in range(world):
            lo, hi, S = host.shard_range(F, r, world)
            assert S == -(-F // world)
            seen += list(range(lo, hi))
        assert seen == list(range(F))


def _build_transcript(tmp_path):
    import subprocess
    exe = str(tmp_path / "system_transcript")
    libdir = os.path.join(ROOT, "r-vio_b200")
    stubs = os.path.join(ROOT, "tests", "stubs")
    subprocess.check_call(["g++", "-std=c++11", "-O1", "-Wall", "-Werror", "-I" + stubs, "-I" + os.path.join(libdir, "host"), "-o", exe,
                           os.path.join(stubs, "system_transcript.cpp"), "-L" + libdir, "-lrvio_b200", "-Wl,-rpath," + libdir])
    return exe


def test_reference_call_sites_compile_unchanged_against_the_literal_adaptor(tmp_path):
    """-DRVIO_B200_WITH_OPENCV_EIGEN flavour (r-vio_b200/host/rvio_ref_api.hpp): RefTracker / RefUpdater carry the reference's
    literal member signatures (cv::Mat, std::list<ImuData*>, Eigen::VectorXd / MatrixXd, cv::FileStorage constructors,
    std::vector<std::list<cv::Point2f>> results); a transcript of System.cc:96-98,258,268-271 compiles against them and
    stand-in OpenCV / Eigen headers (tests/stubs/; C++11 like the reference, CMakeLists.txt:15).  Without a GPU the
    constructors throw (no CPU fallback); with one the transcript runs two frames."""
    import subprocess
    import torch
    if not os.path.exists(capi.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    exe = _build_transcript(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    if torch.cuda.is_available():
        assert out.returncode == 0 and "GPU path ok" in out.stdout, out.stdout + out.stderr
    else:
        assert out.returncode == 3 and "no device path" in out.stdout


@pytest.mark.gpu
def test_reference_call_sites_run_on_gpu(tmp_path):
    import subprocess
    exe = _build_transcript(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "GPU path ok" in out.stdout, out.stdout + out.stderr
