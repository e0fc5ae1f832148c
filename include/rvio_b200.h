/*
 * rvio_b200.h -- C ABI of the B200-native R-VIO hot path (librvio_b200.so).
 *
 * Drop-in boundary for the two calls System::MonoVIO makes into the hot path
 *   mpTracker->track(image, imuList)                      src/rvio/System.cc:258   (Tracker.h:50)
 *   mpUpdater->update(xk1k, Pk1k, types, measurements)     src/rvio/System.cc:268   (Updater.h:43-44)
 * and for the results the host reads back
 *   Tracker::mvFeatTypesForUpdate / mvlFeatMeasForUpdate   Tracker.h:70,74
 *   Updater::xk1k1 / Pk1k1                                 Updater.h:51-52, read at System.cc:270-271
 *
 * Conventions: POD only, opaque handles, int status returns (0 = RVIO_OK), no exceptions cross the
 * boundary, one handle per stream, caller-thread-affine (the reference is single threaded:
 * src/rvio_mono.cc:127).  All pointers are HOST pointers unless the name says `_dev`.
 * Doubles at the boundary keep Eigen::VectorXd / MatrixXd (column-major) unchanged on the host.
 * There is no CPU fallback: every entry point fails with RVIO_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef RVIO_B200_H
#define RVIO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RVIO_OK               0
#define RVIO_FIRST_IMAGE      1   /* track(): first image equalised; caller must seed (Tracker.cc:204-234) */
#define RVIO_NO_FEATURES      2   /* track(): nothing to track, state untouched (Tracker.cc:246-250) */
#define RVIO_DETECTOR_TRUNCATED 3  /* rvio_vio_step with the device detector: the frame was processed (outputs valid), but the detector
                                   * kept only part of its corner candidates (capacity); rvio_b200_last_error() has the text */
#define RVIO_ERR_ARG         -1
#define RVIO_ERR_CUDA        -2
#define RVIO_ERR_STATE       -3
#define RVIO_ERR_CAPACITY    -4

/* ------------------------------------------------------------------------------------------------
 * Tracker  (replaces class RVIO::Tracker, src/rvio/Tracker.h:43-127, and RVIO::Ransac, Ransac.h:61-131)
 * ---------------------------------------------------------------------------------------------- */
typedef struct rvio_tracker rvio_tracker;

/* Every key the Tracker / Ransac constructors read (Tracker.cc:39-79, Ransac.cc:34-46). */
typedef struct rvio_tracker_cfg {
    int32_t width, height;            /* Camera.width / Camera.height */
    float   fx, fy, cx, cy;           /* Camera.fx.. (read into float: Tracker.cc:39-42) */
    float   k1, k2, p1, p2, k3;       /* Camera.k1.. ; k3 == 0 -> 4-coefficient model (Tracker.cc:56-61) */
    int32_t is_rgb;                   /* Camera.RGB (channel order for 3/4-channel input, Tracker.cc:183-196) */
    int32_t is_fisheye;               /* Camera.Fisheye: 0 radtan (cv::undistortPoints), 1 equidistant (cv::fisheye::undistortPoints
                                       * with (k1,k2,p1,p2) as its four coefficients, Tracker.cc:119; k3 must then be 0) */
    int32_t enable_equalizer;         /* Tracker.EnableEqualizer */
    int32_t n_features;               /* Tracker.nFeatures */
    int32_t max_track_len;            /* Tracker.nMaxTrackingLength */
    int32_t min_track_len;            /* Tracker.nMinTrackingLength */
    int32_t use_sampson;              /* Tracker.UseSampson */
    double  inlier_thr;               /* Tracker.nInlierThrd */
    double  small_angle;              /* IMU.nSmallAngle */
    double  T_BC0[16];                /* Camera.T_BC0, row-major 4x4 */
} rvio_tracker_cfg;

/* Tracker::Tracker(const cv::FileStorage&)  -- System.cc:97.  device = CUDA ordinal. */
int rvio_tracker_create(const rvio_tracker_cfg* cfg, int device, rvio_tracker** out);
/* Tracker::~Tracker  -- System.cc:108 */
void rvio_tracker_destroy(rvio_tracker* trk);

/* Tracker::track(im, lImuData) minus the corner detector  -- Tracker.cc:179-342.
 *   img: width x height x channels u8 (channels 1, 3 or 4), row stride in bytes.
 *   imu: n_imu rows of 8 doubles {w[3], a[3], Timestamp, TimeInterval}  (struct ImuData, InputBuffer.h:35-51).
 * Does gray -> CLAHE -> pyramids -> LK -> undistort -> RANSAC -> bookkeeping on the device.
 * Returns RVIO_OK, RVIO_FIRST_IMAGE or RVIO_NO_FEATURES (see above).  After RVIO_OK / RVIO_FIRST_IMAGE the caller
 * runs its detector if it wants to (rvio_tracker_get_image, rvio_tracker_get_tracked_px), hands new corners in with
 * rvio_tracker_seed / rvio_tracker_refill, and finishes the frame with rvio_tracker_commit (Tracker.cc:389-395). */
int rvio_tracker_track(rvio_tracker* trk, const uint8_t* img, int width, int height, int stride_bytes,
                       int channels, const double* imu, int n_imu);
/* Same, image already resident in device memory (single channel, pitch in bytes). */
int rvio_tracker_track_dev(rvio_tracker* trk, const uint8_t* img_dev, int pitch_bytes, const double* imu, int n_imu);

/* Feature-sharded form of rvio_tracker_track (one stream, features split over `world` GPUs, every rank holds
 * the same tracker state and sees the same frame):
 *   _track_begin   image pipeline (redundant on every rank: cheaper than exchanging the pyramid) + pyramidal LK and
 *                  undistortion for the feature indices [rank*S, (rank+1)*S), S = ceil(nFeatures / world); same return
 *                  codes as rvio_tracker_track
 *   _lk_results    device pointers of the per-feature LK outputs (float2 pixels, float2 normalized, u8 status; capacity
 *                  nFeatures + 64 entries each) and S: the caller all-gathers the shards in place (NCCL, <= 17 B/feature)
 *   _track_finish  2-point RANSAC + bookkeeping (Tracker.cc:264-342) on the complete arrays, identical on every rank.
 * Seeding / refill / commit are the ordinary calls, replicated. */
int rvio_tracker_track_begin(rvio_tracker* trk, const uint8_t* img, int width, int height, int stride_bytes,
                             int channels, const double* imu, int n_imu, int rank, int world);
int rvio_tracker_lk_results(rvio_tracker* trk, int world, void** lk_px_dev, void** undist_dev, void** status_dev, int* shard);
int rvio_tracker_track_finish(rvio_tracker* trk);

/* FeatureDetector::DetectWithSubPix(equalised current image, nFeatures, s) ON THE DEVICE (FeatureDetector.cc:55-75:
 * cv::goodFeaturesToTrack(quality = Tracker.nQualLvl, minDistance = s * Tracker.nMinDist) + cv::cornerSubPix(half window
 * floor(.5 nMinDist), 30 iterations / 0.01)); s = 1 on the first image (Tracker.cc:207), 2 for the refill (:350).
 * xy_out: up to nFeatures float2 pixels, strongest first; valid between rvio_tracker_track and rvio_tracker_commit. */
int rvio_tracker_detect(rvio_tracker* trk, int s, float min_dist, float quality, float* xy_out, int* n_out);
/* Equalised current image (what the reference's detector sees: Tracker.cc:207,350). out: width*height bytes. */
int rvio_tracker_get_image(rvio_tracker* trk, uint8_t* out, int out_stride_bytes);
/* mlFreeIndices.size() after bookkeeping (Tracker.cc:344) */
int rvio_tracker_n_free(rvio_tracker* trk, int* n_free);
/* mvFeatsToTrack after loop 2 (Tracker.cc:313-314): pixel coords of the surviving features. */
int rvio_tracker_get_tracked_px(rvio_tracker* trk, float* xy /* 2*n_features */, int* n);
/* First image: Tracker.cc:215-233.  px = detector output (float2 pixels). */
int rvio_tracker_seed(rvio_tracker* trk, const float* px, int n);
/* Refill: Tracker.cc:358-386.  px = corners that passed FeatureDetector::FindNewer.  *n_used = how many were taken. */
int rvio_tracker_refill(rvio_tracker* trk, const float* px, int n, int* n_used);
/* Tracker.cc:389-395: commit feature set, current image becomes mLastImage. */
int rvio_tracker_commit(rvio_tracker* trk);

/* Results == public members Tracker.h:70,74 in CSR form: feature f has type types[f] ('1' lost, '2' max length)
 * and measurements xy[2*offsets[f] .. 2*offsets[f+1]) (normalized, undistorted float2, oldest first). */
int rvio_tracker_get_update_count(rvio_tracker* trk, int* n_feat, int* n_meas);
int rvio_tracker_get_update_lists(rvio_tracker* trk, uint8_t* types, int32_t* offsets /* n_feat+1 */, float* xy);

/* Debug / parity observables of the last track() call (all sized n = features fed to LK). */
int rvio_tracker_get_debug(rvio_tracker* trk, int* n, uint8_t* lk_status, uint8_t* inlier_flags,
                           float* lk_px /* 2n */, float* undist /* 2n */, int32_t* slots /* n */);
/* RANSAC internals of the last call: 32 sampled indices, 16 inlier counts, winner, candidates (Ransac.h:43-59). */
int rvio_tracker_get_ransac_debug(rvio_tracker* trk, int32_t* two_points /* 32 */, int32_t* n_inliers /* 16 */,
                                  int32_t* winner, int32_t* n_candidates, double* hypotheses /* 16*9 row-major */);
/* Pyramid level l (0..3) of the current (0) or previous (1) image, for parity tests. out: lw*lh bytes. */
int rvio_tracker_get_pyramid(rvio_tracker* trk, int which, int level, uint8_t* out, int* lw, int* lh);

/* ------------------------------------------------------------------------------------------------
 * Updater  (replaces class RVIO::Updater, src/rvio/Updater.h:36-71)
 * ---------------------------------------------------------------------------------------------- */
typedef struct rvio_updater rvio_updater;

typedef struct rvio_updater_cfg {
    float   sigma_px, sigma_py;       /* Camera.sigma_px / sigma_py (read into float: Updater.cc:42-44) */
    double  T_BC0[16];                /* Camera.T_BC0, row-major 4x4 (Updater.cc:46-53) */
    int32_t max_clones;               /* capacity: Tracker.nMaxTrackingLength - 1 (System.cc:71-72) */
    int32_t max_features;             /* capacity: ceil(Tracker.nFeatures / 2) (Tracker.cc:74) */
    int32_t max_track_len;            /* capacity: Tracker.nMaxTrackingLength */
} rvio_updater_cfg;

typedef struct rvio_update_info {
    int32_t n_feat, n_good, rows_stacked, updated;
    int32_t n_reject_init, n_reject_lm, n_reject_gate;
    int32_t rank;                     /* rows of the compressed system that entered the EKF step (nRank, Updater.cc:515-524) */
    int32_t rank_flags;               /* RVIO_RANK_* bits below */
} rvio_update_info;

/* Model compression (Updater.cc:474-536).  The reference keeps the rows of its Givens trapezoid up to the FIRST row with
 * norm < 1e-4; when a dependent column sits in the middle of the stacked Jacobian that cut also discards later,
 * informative rows.  RVIO_RANK_RULE_REFERENCE (default) reproduces exactly that; RVIO_RANK_RULE_FULL_INFORMATION keeps
 * every row (the normal terms of the whole stack).  rank_flags of rvio_update_info reports per frame what happened. */
#define RVIO_RANK_RULE_REFERENCE         0
#define RVIO_RANK_RULE_FULL_INFORMATION  1
#define RVIO_RANK_CUT_DISCARDED  1    /* the reference's cut discarded information on this frame */
#define RVIO_RANK_BY_SWEEP       2    /* decided by replaying the reference's Givens sweep (otherwise by the certificate) */
#define RVIO_RANK_UNDECIDED      4    /* feature-sharded call with a dependent column in the middle: full information used */
#define RVIO_RANK_REBUILT        8    /* normal terms rebuilt from the kept rows */
#define RVIO_RANK_DEPENDENT_ROWS 16   /* dependent columns inside the stacked Jacobian, nothing discarded: `rank` counts independent
                                      * directions, the reference's nRank additionally counts its linearly dependent rows */

/* Updater::Updater(const cv::FileStorage&)  -- System.cc:98 */
int rvio_updater_create(const rvio_updater_cfg* cfg, int device, rvio_updater** out);
void rvio_updater_destroy(rvio_updater* upd);

/* Updater::update(xk1k, Pk1k, types, meas) + outputs xk1k1 / Pk1k1  -- Updater.cc:72-628.
 *   x: xdim = 26+7N doubles (layout SURVEY Appendix B), P: d x d column-major, d = 24+6N.
 *   types/offsets/xy: CSR feature lists as produced by rvio_tracker_get_update_lists.
 *   x_out/P_out: same shapes.  info may be NULL. */
int rvio_updater_update(rvio_updater* upd, const double* x, int xdim, const double* P, int d,
                        const uint8_t* types, const int32_t* offsets, const float* xy, int n_feat,
                        double* x_out, double* P_out, rvio_update_info* info);

/* Fused form: consume the feature lists of `trk`'s last track() directly from device memory
 * (no D2H/H2D of the lists); both handles must live on the same device. */
int rvio_updater_update_from_tracker(rvio_updater* upd, rvio_tracker* trk, const double* x, int xdim,
                                     const double* P, int d, double* x_out, double* P_out, rvio_update_info* info);

/* Selects the compression rule (see above).  May be called between frames at any time (also on the updater a fused
 * pipeline owns: rvio_vio_updater). */
int rvio_updater_set_rank_rule(rvio_updater* upd, int mode);

/* Per-feature parity observables of the last update: status (0 accepted, 1 init reject, 2 LM reject, 3 gate reject),
 * inverse-depth estimate [phi psi rho], Mahalanobis distance, dof. */
int rvio_updater_get_debug(rvio_updater* upd, int n_feat, uint8_t* status, double* pfinv /* 3n */,
                           double* gamma, int32_t* dof);
/* Compressed normal terms of the last update: G = H^T H (n x n, row-major), z = H^T r (n), n = 6N. */
int rvio_updater_get_normal_terms(rvio_updater* upd, double* G, double* z, int n);

/* Multi-GPU (feature-sharded) form: rank `rank` of `world` processes only the features f with
 * f % world == rank and leaves its partial normal terms on the device; the host adaptor sums them across ranks
 * (one ncclAllReduce over [G | z | counters], see r-vio_b200/host) between _begin and _finish. */
int rvio_updater_update_begin(rvio_updater* upd, const double* x, int xdim, const double* P, int d,
                              const uint8_t* types, const int32_t* offsets, const float* xy, int n_feat,
                              int rank, int world);
/* Device pointer + element count of the contiguous fp64 reduce buffer [G (n*n) | z (n) | counters (8) | per-class
 * information (n+1)] -- everything one ncclAllReduce(sum) has to carry. */
int rvio_updater_reduce_buffer(rvio_updater* upd, double** buf_dev, int* count);
int rvio_updater_update_finish(rvio_updater* upd, double* x_out, double* P_out, rvio_update_info* info);

/* ------------------------------------------------------------------------------------------------
 * Fused per-frame pipeline (SURVEY 8b "optional fused form rvio_frame"): one call == one System::MonoVIO
 * iteration (src/rvio/System.cc:173-365) with the filter state (x, P), the feature lists and the pyramids resident
 * on the device; per frame only the image + IMU samples go up and the pose (7 doubles) comes back, with a single
 * stream synchronisation.  The corner detector stays outside: its output for this frame is passed in.
 * ---------------------------------------------------------------------------------------------- */
typedef struct rvio_vio rvio_vio;

typedef struct rvio_vio_cfg {
    rvio_tracker_cfg tracker;
    rvio_updater_cfg updater;
    /* IMU.* (System.cc:60-67, PreIntegrator.cc:32-38) */
    double imu_rate, sigma_g, sigma_wg, sigma_a, sigma_wa, gravity;
    /* INI.* (System.cc:77-91) */
    double thr_angle, thr_displ;
    int32_t enable_alignment;
    /* FeatureDetector grid filter (FeatureDetector.cc:31-46) */
    float   min_dist;
    int32_t block_x, block_y;
    float   qual_lvl;                 /* Tracker.nQualLvl (device detector only) */
} rvio_vio_cfg;

int  rvio_vio_create(const rvio_vio_cfg* cfg, int device, rvio_vio** out);
void rvio_vio_destroy(rvio_vio* vio);

/* One frame.  cand_px: n_cand detector corners (float2 pixels) for THIS image: used as seeds on the first tracked
 * image (Tracker.cc:204-234) and passed through FindNewer on the device afterwards (cand_filtered = 0), or taken as
 * already FindNewer-filtered (cand_filtered = 1).  n_cand < 0: the corners come from the device detector
 * (FeatureDetector::DetectWithSubPix on the GPU, beside LK and the update; cand_px is ignored) -- the whole
 * Tracker::track then runs without the host.  pose_out = [pGk(3), qkG(4)] (System.cc:369-374); *pose_valid = 0
 * while the filter is still initialising (System.cc:183-249). */
int rvio_vio_step(rvio_vio* vio, const uint8_t* img, int width, int height, int stride_bytes, int channels,
                  const double* imu, int n_imu, const float* cand_px, int n_cand, int cand_filtered,
                  double* pose_out, int* pose_valid);
/* Same with the (single-channel) image and the candidate list already in device memory. */
int rvio_vio_step_dev(rvio_vio* vio, const uint8_t* img_dev, int pitch_bytes, const double* imu, int n_imu,
                      const float* cand_px_dev, int n_cand, int cand_filtered, double* pose_out, int* pose_valid);
/* Optional: announce a frame ahead of its step -- the moment the reference's host receives it (System::PushImageData,
 * src/rvio/System.h:50, InputBuffer.cc:42), one or two frames before System::MonoVIO pops it (System.cc:176-181).  A
 * single-channel frame in PINNED host memory is uploaded on a copy stream beside the frame being processed (the copy is enqueued
 * as soon as the next rvio_vio_step has launched its own frame, so that it never sits in front of that step's image copy); the
 * rvio_vio_step that is later handed the same buffer (within the next two steps, contents unchanged) takes the uploaded copy
 * instead of uploading inside the step.  Any other frame (colour, pageable) is left to the step: the call does nothing.
 * Results are identical with or without the announcement. */
int rvio_vio_prefetch(rvio_vio* vio, const uint8_t* img, int width, int height, int stride_bytes, int channels);
/* Measurement aid: orders the pipeline's stream after the most recent rvio_vio_prefetch upload (no host blocking), so that an
 * event recorded on rvio_tracker_stream afterwards bounds the upload; *hits (optional) = steps that consumed a prefetched frame. */
int rvio_vio_prefetch_fence(rvio_vio* vio, uint64_t* hits);
/* Filter state after the last step: x (26+7N), P (d x d column-major). */
int rvio_vio_get_state(rvio_vio* vio, double* x, int* xdim, double* P, int* d);
/* Update counters of the last step (n_feat = 0 when no update ran). */
int rvio_vio_get_update_info(rvio_vio* vio, rvio_update_info* info);
/* Debug: per-stage CUDA-event times (ms) of the last step on the main stream: [tracker, per-feature + normal terms,
 * wait for propagation, solve, augment + compose, tail], then two host wall-clock times of the same step: [enqueue of
 * the whole frame, blocked in the one synchronisation]; enable != 0 switches the event instrumentation on. */
int rvio_vio_timeline(rvio_vio* vio, int enable, float* ms8);
/* Steady-state frames are replayed as CUDA graphs (one graph per ping-pong parity / input mode). enable: 1 on (default),
 * 0 off (every frame is enqueued operation by operation), negative: leave unchanged; *graph_launches (optional) receives
 * the number of frames replayed so far. */
int rvio_vio_graphs(rvio_vio* vio, int enable, uint64_t* graph_launches);
/* Feature-sharded single stream (BASELINE configs[4]): `world` processes, one GPU each, every one fed the same
 * frames.  Rank 0 obtains a 128-byte NCCL id (rvio_b200_nccl_unique_id), the host distributes it (MPI, sockets,
 * torch.distributed ...), every rank calls rvio_vio_shard_init before its first frame.  From then on rvio_vio_step runs
 * pyramidal LK for the feature indices [rank*S, (rank+1)*S), S = ceil(nFeatures / world), all-gathers the per-feature
 * results (ncclAllGather, <= 17 B / feature), runs RANSAC + bookkeeping replicated, builds the Jacobian blocks / gate /
 * normal terms of the features f % world == rank, sums [G | z | counters] with ONE ncclAllReduce and solves replicated --
 * both collectives on the pipeline's own stream, inside the captured frame graph.  Poses are identical on every rank.
 * libnccl.so.2 is loaded at run time; nothing else in this library needs it. */
int rvio_b200_nccl_unique_id(void* id128);
int rvio_vio_shard_init(rvio_vio* vio, int rank, int world, const void* nccl_unique_id);
/* Measurement aid: enqueues `iters` x (the frame's all-gather + all-reduce, at this configuration's sizes) on the pipeline's
 * stream and returns the mean time of each in microseconds (us2[0] all-gather, us2[1] all-reduce); collective on all ranks. */
int rvio_vio_shard_probe(rvio_vio* vio, int iters, float* us2);
/* The tracker / updater handles the pipeline owns (for the debug getters above). */
rvio_tracker* rvio_vio_tracker(rvio_vio* vio);
rvio_updater* rvio_vio_updater(rvio_vio* vio);

/* ------------------------------------------------------------------------------------------------
 * Misc
 * ---------------------------------------------------------------------------------------------- */
const char* rvio_b200_version(void);
/* Last CUDA / argument error text for this thread. */
const char* rvio_b200_last_error(void);
/* Number of kernels launched by this library since load (for bench accounting), and a reset. */
uint64_t rvio_b200_kernel_launches(void);
/* Per-kernel CUDA-event timing (bench.py roofline leg): enable, run some steps, then read "kernel count total_ms" lines. */
void rvio_b200_profile(int enable);
int rvio_b200_profile_report(char* buf, int cap);
/* Programmatic dependent launch between the short dependent kernels of a frame (CLAHE -> pyramid -> LK -> RANSAC -> per-feature
 * -> normal terms): each of them becomes resident behind its predecessor and blocks in griddepcontrol.wait until that has
 * completed; results are unchanged, the launch gaps go.  enable: 1 / 0, negative = query; returns the previous setting.
 * Process-wide; on by default (measured +3.4 % frames/s at configs[1]), RVIO_B200_PDL=0 in the environment switches it off. */
int rvio_b200_pdl(int enable);
/* Raw CUDA stream used by a handle (so a host can order its own work / events against it). */
void* rvio_tracker_stream(rvio_tracker* trk);
void* rvio_updater_stream(rvio_updater* upd);

#ifdef __cplusplus
}
#endif
#endif /* RVIO_B200_H */
