// Transcript of the lines of the reference's System.cc that touch the hot path, compiled against
// r-vio_b200/host/rvio_ref_api.hpp (through the three-line Tracker.h / Updater.h of INTEGRATION.md) and the stand-in
// OpenCV / Eigen headers in this directory: proves that the reference's call sites type-check UNCHANGED.
//   System.cc:96-98   mpTracker = new Tracker(fsSettings); mpUpdater = new Updater(fsSettings);
//   System.cc:258     mpTracker->track(pMeasurements.first->Image, pMeasurements.second);
//   System.cc:268     mpUpdater->update(mpPreIntegrator->xk1k, mpPreIntegrator->Pk1k, mpTracker->mvFeatTypesForUpdate, mpTracker->mvlFeatMeasForUpdate);
//   System.cc:270-271 xkk = mpUpdater->xk1k1; Pkk = mpUpdater->Pk1k1;
// With a GPU it also runs two frames + one pass-through update; without one the constructors throw (no CPU fallback).
#include <cstdio>
#include <deque>
#include <list>
#include <utility>
#define RVIO_B200_WITH_OPENCV_EIGEN
#include "rvio_host.hpp"

namespace RVIO {
// ---- reference InputBuffer.h:35-66 (types only)
struct ImuData { Eigen::Vector3d AngularVel; Eigen::Vector3d LinearAccel; double Timestamp; double TimeInterval; ImuData() : Timestamp(0), TimeInterval(0) {} };
struct ImageData { cv::Mat Image; double Timestamp; };
// ---- reference FeatureDetector.h:33-76 (signatures only; the real class stays in the reference tree)
class FeatureDetector {
public:
    FeatureDetector(const cv::FileStorage&) {}
    int DetectWithSubPix(const cv::Mat& im, const int nCorners, const int, std::vector<cv::Point2f>& vCorners)
    {
        vCorners.clear();
        for (int i = 0; i < nCorners && i < 64; ++i) vCorners.push_back(cv::Point2f(40.f + 9.f * (i % 8) * (im.cols / 100.f), 40.f + 9.f * (i / 8) * (im.rows / 100.f)));
        return (int)vCorners.size();
    }
    int FindNewer(const std::vector<cv::Point2f>&, const std::vector<cv::Point2f>&, std::deque<cv::Point2f>& q) { q.clear(); return 0; }
};
// ---- the replacement src/rvio/Tracker.h and Updater.h: three lines
typedef b200::RefTracker<FeatureDetector, ImuData> Tracker;
typedef b200::RefUpdater Updater;
struct PreIntegrator { Eigen::VectorXd xk1k; Eigen::MatrixXd Pk1k; };
}  // namespace RVIO

using namespace RVIO;

int main()
{
    cv::FileStorage fsSettings;
    const char* keys[] = {"Camera.width", "Camera.height", "Camera.fx", "Camera.fy", "Camera.cx", "Camera.cy", "Camera.k1", "Camera.k2",
                          "Tracker.nFeatures", "Tracker.nMaxTrackingLength", "Tracker.nMinTrackingLength", "Tracker.EnableEqualizer",
                          "Tracker.UseSampson", "Tracker.nInlierThrd", "IMU.nSmallAngle", "Camera.sigma_px", "Camera.sigma_py"};
    const double vals[] = {320, 240, 195, 228, 156, 124, -0.28, 0.07, 64, 8, 3, 1, 1, 1e-5, 0.001745329, 0.00218, 0.00219};
    for (size_t i = 0; i < sizeof vals / sizeof vals[0]; ++i) fsSettings.scalars[keys[i]] = vals[i];
    fsSettings.mats["Camera.T_BC0"] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    try {
        Tracker* mpTracker = new Tracker(fsSettings);                                  // System.cc:97
        Updater* mpUpdater = new Updater(fsSettings);                                  // System.cc:98
        PreIntegrator pre; PreIntegrator* mpPreIntegrator = &pre;
        std::vector<unsigned char> pix((size_t)320 * 240);
        for (size_t i = 0; i < pix.size(); ++i) pix[i] = (unsigned char)((i * 2654435761u) >> 24);
        ImageData img; img.Image = cv::Mat(240, 320, CV_8UC1, pix.data()); img.Timestamp = 0;
        ImuData a, b; a.LinearAccel[2] = b.LinearAccel[2] = 9.8; b.TimeInterval = 0.005;
        std::pair<ImageData*, std::list<ImuData*> > pMeasurements(&img, std::list<ImuData*>{&a, &b});
        Eigen::VectorXd xkk; Eigen::MatrixXd Pkk;
        pre.xk1k.setZero(26 + 7 * 3); pre.Pk1k.setZero(42, 42);
        pre.xk1k(3) = pre.xk1k(13) = 1; pre.xk1k(9) = 1;
        for (int c = 0; c < 3; ++c) pre.xk1k(26 + 7 * c + 3) = 1;
        for (int i = 0; i < 42; ++i) pre.Pk1k(i, i) = 1e-4;
        for (int it = 0; it < 2; ++it) {
            mpTracker->track(pMeasurements.first->Image, pMeasurements.second);        // System.cc:258
            mpUpdater->update(mpPreIntegrator->xk1k, mpPreIntegrator->Pk1k, mpTracker->mvFeatTypesForUpdate, mpTracker->mvlFeatMeasForUpdate);   // System.cc:268
            xkk = mpUpdater->xk1k1;                                                    // System.cc:270
            Pkk = mpUpdater->Pk1k1;                                                    // System.cc:271
        }
        std::printf("transcript: GPU path ok (track status %d, update status %d, x %ld, P %ldx%ld)\n", mpTracker->last_status(),
                    mpUpdater->last_status(), (long)xkk.size(), (long)Pkk.rows(), (long)Pkk.cols());
        delete mpTracker; delete mpUpdater;
        return 0;
    } catch (const std::exception& e) {
        std::printf("transcript: no device path: %s\n", e.what());
        return 3;
    }
}
