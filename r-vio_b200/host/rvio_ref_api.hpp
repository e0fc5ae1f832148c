// rvio_ref_api.hpp -- the reference's LITERAL call surface on top of the C ABI (needs OpenCV >= 2.4.3 and Eigen >= 3.1
// like the reference, CMakeLists.txt:43-51).
//
//   RVIO::b200::RefTracker<FeatureDetector, ImuData>
//        explicit RefTracker(const cv::FileStorage&)                                   Tracker.h:46,  System.cc:97
//        void track(const cv::Mat& im, std::list<ImuData*>& lImuData)                  Tracker.h:50,  System.cc:258
//        std::vector<unsigned char> mvFeatTypesForUpdate                               Tracker.h:70
//        std::vector<std::list<cv::Point2f> > mvlFeatMeasForUpdate                     Tracker.h:74
//   RVIO::b200::RefUpdater
//        explicit RefUpdater(const cv::FileStorage&)                                   Updater.h:41,  System.cc:98
//        void update(Eigen::VectorXd&, Eigen::MatrixXd&, std::vector<unsigned char>&,
//                    std::vector<std::list<cv::Point2f> >&)                            Updater.h:43-44, System.cc:268
//        Eigen::VectorXd xk1k1;  Eigen::MatrixXd Pk1k1                                 Updater.h:51-52, System.cc:270-271
//
// With the three-line Tracker.h / Updater.h of INTEGRATION.md ("using Tracker = b200::RefTracker<FeatureDetector, ImuData>;
// using Updater = b200::RefUpdater;") src/rvio/System.cc compiles UNCHANGED.  tests/test_host_logic.py compiles this header
// together with a transcript of those System.cc lines against minimal stand-in OpenCV / Eigen headers (tests/stubs/), since
// this image has neither library.
#pragma once

#include <deque>
#include <Eigen/Core>
#include <opencv2/core/core.hpp>

#include "rvio_host.hpp"

namespace RVIO {
namespace b200 {

namespace pod = ::RVIO_B200_HOST_NS;       // the dependency-free layer of rvio_host.hpp

inline rvio_tracker_cfg tracker_cfg_from(const cv::FileStorage& fs)                 // keys of Tracker.cc:39-79, Ransac.cc:34-46
{
    rvio_tracker_cfg c;
    std::memset(&c, 0, sizeof c);
    c.width = (int)fs["Camera.width"]; c.height = (int)fs["Camera.height"];
    c.fx = (float)fs["Camera.fx"]; c.fy = (float)fs["Camera.fy"]; c.cx = (float)fs["Camera.cx"]; c.cy = (float)fs["Camera.cy"];
    c.k1 = (float)fs["Camera.k1"]; c.k2 = (float)fs["Camera.k2"]; c.p1 = (float)fs["Camera.p1"]; c.p2 = (float)fs["Camera.p2"];
    c.k3 = (float)fs["Camera.k3"];
    c.is_rgb = (int)fs["Camera.RGB"]; c.is_fisheye = (int)fs["Camera.Fisheye"];
    c.enable_equalizer = (int)fs["Tracker.EnableEqualizer"];
    c.n_features = (int)fs["Tracker.nFeatures"];
    c.max_track_len = (int)fs["Tracker.nMaxTrackingLength"]; c.min_track_len = (int)fs["Tracker.nMinTrackingLength"];
    c.use_sampson = (int)fs["Tracker.UseSampson"]; c.inlier_thr = (double)fs["Tracker.nInlierThrd"];
    c.small_angle = (double)fs["IMU.nSmallAngle"];
    cv::Mat T; fs["Camera.T_BC0"] >> T;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) c.T_BC0[4 * i + j] = T.at<double>(i, j);
    return c;
}

inline rvio_updater_cfg updater_cfg_from(const cv::FileStorage& fs)                 // keys of Updater.cc:40-53; capacities System.cc:71-72, Tracker.cc:74
{
    rvio_updater_cfg c;
    std::memset(&c, 0, sizeof c);
    c.sigma_px = (float)fs["Camera.sigma_px"]; c.sigma_py = (float)fs["Camera.sigma_py"];
    cv::Mat T; fs["Camera.T_BC0"] >> T;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) c.T_BC0[4 * i + j] = T.at<double>(i, j);
    c.max_track_len = (int)fs["Tracker.nMaxTrackingLength"];
    c.max_clones = c.max_track_len - 1;
    c.max_features = ((int)fs["Tracker.nFeatures"] + 1) / 2;
    return c;
}

// FeatureDetectorT: the reference's own class (FeatureDetector.h:33-76, stays on the host, compiled unchanged);
// ImuDataT: the reference's struct (InputBuffer.h:35-51; AngularVel / LinearAccel are Eigen::Vector3d there).
template <class FeatureDetectorT, class ImuDataT>
class RefTracker {
    struct Bridge : pod::Detector {
        explicit Bridge(const cv::FileStorage& fs) : det(fs) {}
        int DetectWithSubPix(const uint8_t* im, int w, int h, int n, int s, std::vector<pod::Point2f>& out) override
        {
            cv::Mat m(h, w, CV_8UC1, const_cast<uint8_t*>(im));
            std::vector<cv::Point2f> c;
            const int k = det.DetectWithSubPix(m, n, s, c);                           // FeatureDetector.cc:55-75
            out.clear();
            for (const cv::Point2f& p : c) out.push_back(pod::Point2f{p.x, p.y});
            return k;
        }
        int FindNewer(const std::vector<pod::Point2f>& a, const std::vector<pod::Point2f>& b, std::vector<pod::Point2f>& q) override
        {
            std::vector<cv::Point2f> va, vb;
            std::deque<cv::Point2f> dq;                                               // FeatureDetector.cc:97-150
            for (const pod::Point2f& p : a) va.push_back(cv::Point2f(p.x, p.y));
            for (const pod::Point2f& p : b) vb.push_back(cv::Point2f(p.x, p.y));
            const int k = det.FindNewer(va, vb, dq);
            q.clear();
            for (const cv::Point2f& p : dq) q.push_back(pod::Point2f{p.x, p.y});
            return k;
        }
        FeatureDetectorT det;
    };

public:
    explicit RefTracker(const cv::FileStorage& fsSettings, int device = 0)
        : mBridge(fsSettings), mImpl(tracker_cfg_from(fsSettings), &mBridge, device)
    {
        mvlFeatMeasForUpdate.resize((size_t)std::ceil(.5 * (int)fsSettings["Tracker.nFeatures"]));      // Tracker.cc:74
    }
    // cv::goodFeaturesToTrack + cv::cornerSubPix on the GPU instead of through FeatureDetectorT (FindNewer stays on the host)
    void UseDeviceDetector(const cv::FileStorage& fs) { mImpl.UseDeviceDetector((float)fs["Tracker.nMinDist"], (float)fs["Tracker.nQualLvl"]); }

    void track(const cv::Mat& im, std::list<ImuDataT*>& lImuData)
    {
        std::list<pod::ImuData> flat;
        std::list<pod::ImuData*> ptrs;
        for (const ImuDataT* d : lImuData) {
            pod::ImuData r;
            for (int k = 0; k < 3; ++k) { r.AngularVel[k] = d->AngularVel[k]; r.LinearAccel[k] = d->LinearAccel[k]; }
            r.Timestamp = d->Timestamp; r.TimeInterval = d->TimeInterval;
            flat.push_back(r);
            ptrs.push_back(&flat.back());
        }
        mImpl.track(im.data, im.cols, im.rows, (int)im.step, im.channels(), ptrs);
        if (mImpl.last_status() < 0 || mImpl.last_status() == RVIO_NO_FEATURES) return;   // outputs stay stale (Tracker.cc:246-250)
        mvFeatTypesForUpdate = mImpl.mvFeatTypesForUpdate;
        mvlFeatMeasForUpdate.assign(mImpl.mvlFeatMeasForUpdate.size(), std::list<cv::Point2f>());
        for (size_t f = 0; f < mImpl.mvlFeatMeasForUpdate.size(); ++f)
            for (const pod::Point2f& p : mImpl.mvlFeatMeasForUpdate[f]) mvlFeatMeasForUpdate[f].push_back(cv::Point2f(p.x, p.y));
    }

    int last_status() const { return mImpl.last_status(); }

    std::vector<unsigned char> mvFeatTypesForUpdate;
    std::vector<std::list<cv::Point2f> > mvlFeatMeasForUpdate;

private:
    Bridge mBridge;
    pod::Tracker mImpl;
};

class RefUpdater {
public:
    explicit RefUpdater(const cv::FileStorage& fsSettings, int device = 0) : mImpl(updater_cfg_from(fsSettings), device)
    {
        xk1k1.setZero(26, 1);                                                          // Updater.cc:55-56
        Pk1k1.setZero(24, 24);
    }

    void update(Eigen::VectorXd& xk1k, Eigen::MatrixXd& Pk1k, std::vector<unsigned char>& vFeatTypesForUpdate,
                std::vector<std::list<cv::Point2f> >& vlFeatMeasForUpdate)
    {
        std::vector<double> x(xk1k.data(), xk1k.data() + xk1k.size()), P(Pk1k.data(), Pk1k.data() + Pk1k.size());
        std::vector<std::list<pod::Point2f> > m(vlFeatMeasForUpdate.size());
        for (size_t f = 0; f < vlFeatMeasForUpdate.size(); ++f)
            for (const cv::Point2f& p : vlFeatMeasForUpdate[f]) m[f].push_back(pod::Point2f{p.x, p.y});
        mImpl.update(x, P, vFeatTypesForUpdate, m);
        const Eigen::Index d = Pk1k.rows();
        xk1k1 = Eigen::Map<Eigen::VectorXd>(mImpl.xk1k1.data(), (Eigen::Index)mImpl.xk1k1.size());
        Pk1k1 = Eigen::Map<Eigen::MatrixXd>(mImpl.Pk1k1.data(), d, d);
    }

    int last_status() const { return mImpl.last_status(); }
    const rvio_update_info& info() const { return mImpl.info(); }

    Eigen::VectorXd xk1k1;
    Eigen::MatrixXd Pk1k1;

private:
    pod::Updater mImpl;
};

}  // namespace b200
}  // namespace RVIO
