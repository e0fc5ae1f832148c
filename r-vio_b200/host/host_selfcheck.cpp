// host_selfcheck.cpp -- compiles rvio_host.hpp (dependency-free flavour) against librvio_b200.so and exercises the
// adaptor's error path: without a usable sm_100 device the constructors must throw (there is no CPU fallback);
// with one, a synthetic two-frame track + a pass-through update must run.  Used by tests/test_host_logic.py.
#include <cstdio>
#include <cstdlib>
#include "rvio_host.hpp"

struct NoDetector : RVIO::Detector {
    int DetectWithSubPix(const uint8_t*, int w, int h, int n, int, std::vector<RVIO::Point2f>& v) override
    {
        v.clear();
        for (int i = 0; i < n && i < 64; ++i) v.push_back(RVIO::Point2f{40.f + 9.f * (i % 8) * (w / 100.f), 40.f + 9.f * (i / 8) * (h / 100.f)});
        return (int)v.size();
    }
    int FindNewer(const std::vector<RVIO::Point2f>&, const std::vector<RVIO::Point2f>&, std::vector<RVIO::Point2f>& q) override { q.clear(); return 0; }
};

int main()
{
    rvio_tracker_cfg tc;
    std::memset(&tc, 0, sizeof tc);
    tc.width = 320; tc.height = 240; tc.fx = 195.f; tc.fy = 228.f; tc.cx = 156.f; tc.cy = 124.f;
    tc.k1 = -0.28f; tc.k2 = 0.07f; tc.n_features = 64; tc.max_track_len = 8; tc.min_track_len = 3; tc.enable_equalizer = 1;
    tc.use_sampson = 1; tc.inlier_thr = 1e-5; tc.small_angle = 0.001745329;
    for (int i = 0; i < 4; ++i) tc.T_BC0[5 * i] = 1.0;
    rvio_updater_cfg uc;
    std::memset(&uc, 0, sizeof uc);
    uc.sigma_px = 0.00218f; uc.sigma_py = 0.00219f; uc.max_clones = 7; uc.max_features = 32; uc.max_track_len = 8;
    for (int i = 0; i < 4; ++i) uc.T_BC0[5 * i] = 1.0;
    NoDetector det;
    try {
        RVIO::Tracker trk(tc, &det, 0);
        RVIO::Updater upd(uc, 0);
        std::vector<uint8_t> img((size_t)tc.width * tc.height);
        for (size_t i = 0; i < img.size(); ++i) img[i] = (uint8_t)((i * 2654435761u) >> 24);
        RVIO::ImuData a{}, b{};
        a.LinearAccel[2] = b.LinearAccel[2] = 9.8; b.TimeInterval = 0.005;
        std::list<RVIO::ImuData*> imu{&a, &b};
        trk.track(img.data(), tc.width, tc.height, tc.width, 1, imu);
        trk.track(img.data(), tc.width, tc.height, tc.width, 1, imu);
        std::vector<double> x(26 + 7 * 3, 0.0), P((24 + 18) * (24 + 18), 0.0);
        x[3] = x[13] = 1; x[9] = 1;
        for (int c = 0; c < 3; ++c) x[26 + 7 * c + 3] = 1;
        for (int i = 0; i < 42; ++i) P[i * 42 + i] = 1e-4;
        upd.update(x, P, trk.mvFeatTypesForUpdate, trk.mvlFeatMeasForUpdate);
        std::printf("selfcheck: GPU path ok (track status %d, %zu features for update, update status %d)\n", trk.last_status(),
                    trk.mvFeatTypesForUpdate.size(), upd.last_status());
        return 0;
    } catch (const std::exception& e) {
        std::printf("selfcheck: no device path: %s\n", e.what());
        return 3;
    }
}
