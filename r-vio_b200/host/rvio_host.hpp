// rvio_host.hpp -- C++ host adaptor above the C ABI (include/rvio_b200.h).
//
// Mirrors the reference's call surface for the hot path so that System::MonoVIO (reference src/rvio/System.cc:173-437)
// compiles unchanged against it:
//     class RVIO::Tracker  { void track(im, lImuData); mvFeatTypesForUpdate; mvlFeatMeasForUpdate; }   Tracker.h:43-127
//     class RVIO::Updater  { void update(xk1k, Pk1k, types, meas); xk1k1; Pk1k1; }                      Updater.h:36-71
// Same member names, argument meaning and "errors are not signalled" behaviour (both return void; failures of the device
// path are kept in last_status() and leave the outputs stale / pass-through exactly like the reference's early returns).
//
// Two flavours:
//   * default: dependency-free (std containers + the small POD types below).  This is what is compiled and tested here
//     (this image has neither OpenCV C++ headers nor Eigen).
//   * -DRVIO_B200_WITH_OPENCV_EIGEN (rvio_ref_api.hpp): classes with the reference's literal member signatures (cv::Mat,
//     std::list<ImuData*>, Eigen::VectorXd / MatrixXd, cv::FileStorage constructors, cv::Point2f result lists) so that
//     System.cc compiles unchanged; compiled in tests/ against stand-in headers (tests/stubs/), see INTEGRATION.md.
//
// The corner detector: by default, exactly as in the reference, a host object
// (FeatureDetector::DetectWithSubPix / FindNewer) that the adaptor calls through the `Detector` interface below between
// rvio_tracker_track() and rvio_tracker_commit(); Tracker::UseDeviceDetector() moves DetectWithSubPix onto the GPU.
//
//     class RVIO::FusedVio { bool MonoVIO(im, lImuData, pose) }  -- one System::MonoVIO iteration (System.cc:173-365) as
//     ONE device call (rvio_vio_step): track, propagate, update, augment and compose with x, P, pyramids and feature lists
//     resident on the GPU, detector included; only the pose [pGk, qkG] of System.cc:369-374 comes back.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <list>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/rvio_b200.h"

// The dependency-free classes below carry the reference's names (RVIO::Tracker, RVIO::Updater, RVIO::ImuData).  In the
// literal flavour those names belong to the reference tree itself (its ImuData, and the Tracker / Updater aliases of
// INTEGRATION.md), so the dependency-free layer moves to its own namespace.
#ifndef RVIO_B200_HOST_NS
#ifdef RVIO_B200_WITH_OPENCV_EIGEN
#define RVIO_B200_HOST_NS rvio_b200_pod
#else
#define RVIO_B200_HOST_NS RVIO
#endif
#endif

namespace RVIO_B200_HOST_NS {

struct Point2f { float x, y; };                       // stands in for cv::Point2f

// struct ImuData, reference src/rvio/InputBuffer.h:35-51 (same field order; Eigen::Vector3d -> double[3])
struct ImuData {
    double AngularVel[3];
    double LinearAccel[3];
    double Timestamp;
    double TimeInterval;
};

// What FeatureDetector offers Tracker::track (reference src/rvio/FeatureDetector.h:37-52).
struct Detector {
    virtual ~Detector() {}
    // DetectWithSubPix(im, nCorners, s, vCorners): im is the equalised 8-bit image, row stride = width
    virtual int DetectWithSubPix(const uint8_t* im, int width, int height, int nCorners, int s, std::vector<Point2f>& vCorners) = 0;
    // FindNewer(vCorners, vRefCorners, qNewCorners)
    virtual int FindNewer(const std::vector<Point2f>& vCorners, const std::vector<Point2f>& vRefCorners, std::vector<Point2f>& qNewCorners) = 0;
};

class Tracker {
public:
    // Tracker(const cv::FileStorage&) in the reference (Tracker.cc:37-90); here the parsed keys arrive as rvio_tracker_cfg.
    Tracker(const rvio_tracker_cfg& cfg, Detector* detector, int device = 0)
        : mCfg(cfg), mpFeatureDetector(detector), mHandle(nullptr), mLastStatus(RVIO_OK), mbDeviceDetector(false),
          mnMinDist(0.f), mnQualLvl(0.f)
    {
        mLastStatus = rvio_tracker_create(&mCfg, device, &mHandle);
        if (mLastStatus != RVIO_OK) throw std::runtime_error(std::string("rvio_tracker_create: ") + rvio_b200_last_error());
        mEq.resize((size_t)cfg.width * cfg.height);
        mvlFeatMeasForUpdate.resize((size_t)std::ceil(.5 * cfg.n_features));
    }
    ~Tracker() { rvio_tracker_destroy(mHandle); }
    // FeatureDetector::DetectWithSubPix on the GPU (rvio_tracker_detect) instead of through the host detector object;
    // nMinDist / nQualLvl are Tracker.nMinDist / Tracker.nQualLvl (FeatureDetector.cc:31-32).  FindNewer stays with the
    // Detector object passed to the constructor.
    void UseDeviceDetector(float nMinDist, float nQualLvl) { mbDeviceDetector = true; mnMinDist = nMinDist; mnQualLvl = nQualLvl; }
    Tracker(const Tracker&) = delete;
    Tracker& operator=(const Tracker&) = delete;

    // void Tracker::track(const cv::Mat& im, std::list<ImuData*>& lImuData)   -- Tracker.h:50, called at System.cc:258
    void track(const uint8_t* im, int width, int height, int stride_bytes, int channels, std::list<ImuData*>& lImuData)
    {
        std::vector<double> imu;
        imu.reserve(lImuData.size() * 8);
        for (const ImuData* d : lImuData) {
            imu.insert(imu.end(), d->AngularVel, d->AngularVel + 3);
            imu.insert(imu.end(), d->LinearAccel, d->LinearAccel + 3);
            imu.push_back(d->Timestamp);
            imu.push_back(d->TimeInterval);
        }
        const int rc = rvio_tracker_track(mHandle, im, width, height, stride_bytes, channels, imu.data(), (int)lImuData.size());
        mLastStatus = rc;
        if (rc < 0 || rc == RVIO_NO_FEATURES) return;              // Tracker.cc:246-250: early return, lists stay stale
        if (rc == RVIO_FIRST_IMAGE) {
            // Tracker.cc:204-234
            std::vector<Point2f> corners;
            const int n = DetectWithSubPix(1, corners);
            if (n > 0) mLastStatus = rvio_tracker_seed(mHandle, &corners[0].x, n);
            rvio_tracker_commit(mHandle);
            return;
        }
        // results: Tracker.cc:271-342
        int nFeat = 0, nMeas = 0;
        rvio_tracker_get_update_count(mHandle, &nFeat, &nMeas);
        std::vector<uint8_t> types((size_t)nFeat + 1);
        std::vector<int32_t> off((size_t)nFeat + 1);
        std::vector<float> xy((size_t)2 * nMeas + 2);
        rvio_tracker_get_update_lists(mHandle, types.data(), off.data(), xy.data());
        mvFeatTypesForUpdate.assign(types.begin(), types.begin() + nFeat);
        mvlFeatMeasForUpdate.clear();
        mvlFeatMeasForUpdate.resize((size_t)std::ceil(.5 * mCfg.n_features));           // Tracker.cc:272-274
        for (int f = 0; f < nFeat; ++f)
            for (int k = off[f]; k < off[f + 1]; ++k) mvlFeatMeasForUpdate[f].push_back(Point2f{xy[2 * k], xy[2 * k + 1]});
        // refill: Tracker.cc:344-387
        int nFree = 0;
        rvio_tracker_n_free(mHandle, &nFree);
        if (nFree > 0 && mpFeatureDetector) {
            {
                std::vector<Point2f> vTempFeats, qNewFeats, vRef((size_t)mCfg.n_features);
                DetectWithSubPix(2, vTempFeats);
                int nRef = 0;
                rvio_tracker_get_tracked_px(mHandle, &vRef[0].x, &nRef);
                vRef.resize((size_t)nRef);
                const int nNew = mpFeatureDetector->FindNewer(vTempFeats, vRef, qNewFeats);
                if (nNew > 0) rvio_tracker_refill(mHandle, &qNewFeats[0].x, nNew, nullptr);
            }
        }
        mLastStatus = rvio_tracker_commit(mHandle);                                     // Tracker.cc:389-395
    }

    int last_status() const { return mLastStatus; }
    rvio_tracker* handle() { return mHandle; }

private:
    // mpFeatureDetector->DetectWithSubPix(im, mnMaxFeatsPerImage, s, corners) (Tracker.cc:207,350) on the equalised image
    int DetectWithSubPix(int s, std::vector<Point2f>& vCorners)
    {
        vCorners.clear();
        if (mbDeviceDetector) {
            vCorners.resize((size_t)mCfg.n_features);
            int n = 0;
            const int rc = rvio_tracker_detect(mHandle, s, mnMinDist, mnQualLvl, &vCorners[0].x, &n);
            if (rc != RVIO_OK) { mLastStatus = rc; n = 0; }
            vCorners.resize((size_t)n);
            return n;
        }
        if (!mpFeatureDetector || rvio_tracker_get_image(mHandle, mEq.data(), mCfg.width) != RVIO_OK) return 0;
        return mpFeatureDetector->DetectWithSubPix(mEq.data(), mCfg.width, mCfg.height, mCfg.n_features, s, vCorners);
    }

public:
    // Feature types for update: '1' lose track, '2' reach the max. tracking length   (Tracker.h:66-70)
    std::vector<unsigned char> mvFeatTypesForUpdate;
    // Feature measurements for update (Tracker.h:72-74)
    std::vector<std::list<Point2f> > mvlFeatMeasForUpdate;

private:
    rvio_tracker_cfg mCfg;
    Detector* mpFeatureDetector;
    rvio_tracker* mHandle;
    int mLastStatus;
    bool mbDeviceDetector;
    float mnMinDist, mnQualLvl;
    std::vector<uint8_t> mEq;
};

class Updater {
public:
    // Updater(const cv::FileStorage&) in the reference (Updater.cc:38-69)
    Updater(const rvio_updater_cfg& cfg, int device = 0) : mHandle(nullptr), mLastStatus(RVIO_OK)
    {
        mLastStatus = rvio_updater_create(&cfg, device, &mHandle);
        if (mLastStatus != RVIO_OK) throw std::runtime_error(std::string("rvio_updater_create: ") + rvio_b200_last_error());
        xk1k1.assign(26, 0.0);                                                         // Updater.cc:55-56
        Pk1k1.assign(24 * 24, 0.0);
    }
    ~Updater() { rvio_updater_destroy(mHandle); }
    Updater(const Updater&) = delete;
    Updater& operator=(const Updater&) = delete;

    // void Updater::update(Eigen::VectorXd& xk1k, Eigen::MatrixXd& Pk1k, std::vector<unsigned char>&,
    //                      std::vector<std::list<cv::Point2f> >&)   -- Updater.h:43-44, called at System.cc:268.
    // xk1k: 26+7N doubles; Pk1k: column-major d x d (== Eigen::MatrixXd::data()).
    void update(const std::vector<double>& xk1k, const std::vector<double>& Pk1k,
                const std::vector<unsigned char>& vFeatTypesForUpdate,
                const std::vector<std::list<Point2f> >& vlFeatMeasForUpdate)
    {
        const int xdim = (int)xk1k.size();
        const int d = 24 + 6 * ((xdim - 26) / 7);
        const int nFeat = (int)vFeatTypesForUpdate.size();                             // Updater.cc:90
        std::vector<int32_t> off((size_t)nFeat + 1, 0);
        std::vector<float> xy;
        for (int f = 0; f < nFeat; ++f) {
            for (const Point2f& p : vlFeatMeasForUpdate[(size_t)f]) { xy.push_back(p.x); xy.push_back(p.y); }
            off[(size_t)f + 1] = (int32_t)(xy.size() / 2);
        }
        xk1k1.resize((size_t)xdim);
        Pk1k1.resize((size_t)d * d);
        mLastStatus = rvio_updater_update(mHandle, xk1k.data(), xdim, Pk1k.data(), d,
                                          nFeat ? vFeatTypesForUpdate.data() : nullptr, off.data(), xy.empty() ? nullptr : xy.data(), nFeat,
                                          xk1k1.data(), Pk1k1.data(), &mInfo);
        if (mLastStatus != RVIO_OK) { xk1k1 = xk1k; Pk1k1 = Pk1k; }                   // keep the filter alive: posterior = prior
    }

    int last_status() const { return mLastStatus; }
    const rvio_update_info& info() const { return mInfo; }
    rvio_updater* handle() { return mHandle; }

public:
    // Outputs (Updater.h:50-52)
    std::vector<double> xk1k1;
    std::vector<double> Pk1k1;      // column-major d x d

private:
    rvio_updater* mHandle;
    int mLastStatus;
    rvio_update_info mInfo{};
};

// One System::MonoVIO iteration per call (System.cc:173-365) on the device.  cfg carries every key System, Tracker, Updater,
// PreIntegrator and FeatureDetector read from the YAML (rvio_vio_cfg).  The motion-detection / initialisation statics of
// System.cc:175-249 live in the handle, so several sequences can run in one process.
class FusedVio {
public:
    explicit FusedVio(const rvio_vio_cfg& cfg, int device = 0) : mHandle(nullptr), mLastStatus(RVIO_OK)
    {
        mLastStatus = rvio_vio_create(&cfg, device, &mHandle);
        if (mLastStatus != RVIO_OK) throw std::runtime_error(std::string("rvio_vio_create: ") + rvio_b200_last_error());
    }
    ~FusedVio() { rvio_vio_destroy(mHandle); }
    FusedVio(const FusedVio&) = delete;
    FusedVio& operator=(const FusedVio&) = delete;

    // pMeasurements = {image, imus} of System.cc:179-181.  Returns true when a pose was produced (false while the filter is
    // still initialising, System.cc:183-249); pose = [pGk(3), qkG(4)] as written to stamped_pose_ests.dat (System.cc:371-373).
    bool MonoVIO(const uint8_t* im, int width, int height, int stride_bytes, int channels, const std::list<ImuData*>& lImuData,
                 double pose[7])
    {
        std::vector<double> imu;
        imu.reserve(lImuData.size() * 8);
        for (const ImuData* d : lImuData) {
            imu.insert(imu.end(), d->AngularVel, d->AngularVel + 3);
            imu.insert(imu.end(), d->LinearAccel, d->LinearAccel + 3);
            imu.push_back(d->Timestamp);
            imu.push_back(d->TimeInterval);
        }
        int valid = 0;
        mLastStatus = rvio_vio_step(mHandle, im, width, height, stride_bytes, channels, imu.data(), (int)lImuData.size(),
                                    nullptr, /*n_cand: device detector*/ -1, 0, pose, &valid);
        return mLastStatus >= 0 && valid != 0;
    }

    // System::PushImageData (System.h:50, called from the image callback rvio_mono.cc:76): the frame is known to the host one
    // or two iterations before MonoVIO pops it.  Announcing it here lets the library upload a pinned mono8 frame beside the
    // frame in flight (rvio_vio_prefetch); MonoVIO on the same buffer then starts without its own upload.  Optional.
    void PushImageData(const uint8_t* im, int width, int height, int stride_bytes, int channels)
    {
        mLastStatus = rvio_vio_prefetch(mHandle, im, width, height, stride_bytes, channels);
    }

    int last_status() const { return mLastStatus; }
    rvio_vio* handle() { return mHandle; }

private:
    rvio_vio* mHandle;
    int mLastStatus;
};

}  // namespace RVIO_B200_HOST_NS

#ifdef RVIO_B200_WITH_OPENCV_EIGEN
// The reference's literal signatures (cv::Mat, std::list<ImuData*>, Eigen::VectorXd / MatrixXd, cv::FileStorage
// constructors, public result members of the reference's types): rvio_ref_api.hpp.
#include "rvio_ref_api.hpp"
#endif  // RVIO_B200_WITH_OPENCV_EIGEN
