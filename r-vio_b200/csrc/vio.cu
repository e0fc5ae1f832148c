// vio.cu -- fused per-frame pipeline: one rvio_vio_step == one System::MonoVIO iteration
// (reference src/rvio/System.cc:173-365) with x, P, pyramids and feature lists resident on the device.
// Host side keeps only what is inherently serial and tiny: the motion-detection / initialisation logic
// (System.cc:183-249, System::initialize :115-170) and the clone counters.
#include <time.h>
#include "common.cuh"
#include "shard.cuh"
#include "tracker_kernels.cuh"
#include "filter_kernels.cuh"
#include "detector_kernels.cuh"

#include <math.h>
#include <new>
#include <unordered_map>
#include <vector>

namespace rvio {
// tracker.cu
int tracker_enqueue_frame_host(rvio_tracker* t, const uint8_t* img, int w, int h, int stride, int ch, const double* imu, int n_imu);
int tracker_enqueue_frame_dev(rvio_tracker* t, const uint8_t* img_dev, int pitch, const double* imu, int n_imu);
int tracker_enqueue_seed_dev(rvio_tracker* t, const float2* px_dev, int n, const int* n_dev);
Detector* tracker_detector(rvio_tracker* t);
const PyrLevel* tracker_level0(const rvio_tracker* t);
cudaEvent_t tracker_level0_event(const rvio_tracker* t);
void tracker_set_first(rvio_tracker* t, bool first);
int tracker_enqueue_frame_staged(rvio_tracker* t, const double* imu, int n_imu);
uint8_t* tracker_gray(rvio_tracker* t, size_t* pitch);
int tracker_sync(rvio_tracker* t);
int tracker_enqueue_scalars(rvio_tracker* t);
int tracker_wait(rvio_tracker* t);
int tracker_parity(const rvio_tracker* t);
int tracker_n_track(const rvio_tracker* t);
bool tracker_is_first(const rvio_tracker* t);
const CamParams* tracker_cam(const rvio_tracker* t);
const TrackerBuffers* tracker_buffers(const rvio_tracker* t);
cudaStream_t tracker_stream(const rvio_tracker* t);
const TrackerScalars* tracker_host_scalars(const rvio_tracker* t);
// updater.cu
void updater_hint_kernel_predecessor(rvio_updater* u, bool yes);
int updater_enqueue_normal_terms(rvio_updater* u, cudaStream_t s, const double* x_dev, int xdim, const double* P_dev, int d,
                                 const uint8_t* types_dev, const int32_t* off_dev, const float2* xy_dev,
                                 int n_feat_cap, const int* n_feat_dev, int rank, int world);
int updater_enqueue_solve(rvio_updater* u, cudaStream_t s, double* x_out_dev, double* P_out_dev);
int updater_enqueue_solve_on(rvio_updater* u, cudaStream_t s, const double* x_dev, const double* P_dev, double* x_out_dev, double* P_out_dev);
const double* updater_counters_dev(const rvio_updater* u);
double* updater_reduce_dev(rvio_updater* u, int* count);
int tracker_enqueue_frame_sharded(rvio_tracker* t, const uint8_t* img_host, int w, int h, int stride, int ch, const uint8_t* img_dev, int pitch,
                                  bool staged, const double* imu, int n_imu, int rank, int world);
int tracker_enqueue_ransac(rvio_tracker* t);
int tracker_shard_size(const rvio_tracker* t, int world);
}  // namespace rvio

using namespace rvio;

struct rvio_vio {
    rvio_vio_cfg cfg;
    int device;
    rvio_tracker* trk;
    rvio_updater* upd;
    cudaStream_t stream;          // main: tracker -> per-feature -> normal terms -> solve -> augment
    cudaStream_t side;            // side: propagate, FindNewer+refill (off the critical path)
    cudaStream_t dets;            // device detector (DetectWithSubPix), beside LK / RANSAC / propagate
    cudaEvent_t ev_det_done;
    cudaEvent_t ev_frame_in, ev_prop_done, ev_bookkeep_done, ev_side_done;
    // frame graphs: the steady-state frame is a fixed sequence of ~20 stream operations whose parameters only depend on a
    // few host-known bits (ping-pong parities, input mode, IMU block size); it is captured once per signature and replayed
    bool use_graphs;
    std::unordered_map<uint64_t, cudaGraphExec_t> graphs;
    uint64_t graph_launches;
    bool timeline; cudaEvent_t tl[8]; float tl_ms[8];   // optional per-stage stamps on the main stream
    unsigned long long* d_stamps; unsigned long long h_stamps[8];
    std::unordered_map<const void*, bool> pin_cache;   // is this host frame buffer pinned? (cudaPointerGetAttributes, asked once per address)
    // rvio_vio_prefetch: a frame announced ahead of its step (System::PushImageData time) is uploaded on a copy stream into one
    // of two device slots while the previous frame is still being processed; the step that is later handed the same host
    // buffer takes the slot instead of uploading
    cudaStream_t copys; cudaEvent_t ev_pref[2]; uint8_t* d_pref[2]; const uint8_t* pref_src[2]; int pref_next, pref_last;
    uint64_t pref_hits, pref_step[2], n_steps;    // a slot is honoured by the next two steps only (a forgotten announcement expires)
    // An announcement is only NOTED by rvio_vio_prefetch; its H2D copy is enqueued right after the next step has launched its frame
    // (or at once when that step is the announced frame's own): a copy enqueued BEFORE the step would sit on the copy engine in front
    // of the step's own image copy, and one isolated H2D transfer has ~50 us of latency on the measured boxes.
    const uint8_t* pend_src; int pend_stride; uint64_t pend_step;
    int window, min_clones, Fu, F;
    // device state (ping-pong)
    double* d_x[2]; double* d_P[2]; int xi, pi;
    double* d_pose; double* d_imu; float2* d_cand;      // d_imu: [frame header {n_imu, n_cand} (16 B)][n_imu x 8 doubles]
    // pinned
    double* h_pose; double* h_imu; float* h_cand; double* h_cnt; double* h_state; int* h_detctrl;
    // System.cc statics, per instance
    bool moving, ready;
    double wm[3], am[3];
    int n_imu_count, n_clones, n_img_after_init;
    rvio_update_info last_info;
    ShardComm shard;              // feature-sharded single stream (rvio_vio_shard_init): world > 1
    // FindNewer geometry
    int gc, gr, offx, offy, max_per_block;
    std::vector<void*> allocs, hallocs;
};

namespace {

int xdim_of(int N) { return 26 + 7 * N; }
int d_of(int N) { return 24 + 6 * N; }

void h_quat_from_rot(const double* R, double* q)      // Numerics.h:126-167
{
    const double T = R[0] + R[4] + R[8];
    if (R[0] > T && R[0] > R[4] && R[0] > R[8]) {
        q[0] = sqrt((1 + 2 * R[0] - T) / 4);
        q[1] = (1 / (4 * q[0])) * (R[1] + R[3]); q[2] = (1 / (4 * q[0])) * (R[2] + R[6]); q[3] = (1 / (4 * q[0])) * (R[5] - R[7]);
    } else if (R[4] > T && R[4] > R[0] && R[4] > R[8]) {
        q[1] = sqrt((1 + 2 * R[4] - T) / 4);
        q[0] = (1 / (4 * q[1])) * (R[1] + R[3]); q[2] = (1 / (4 * q[1])) * (R[5] + R[7]); q[3] = (1 / (4 * q[1])) * (R[6] - R[2]);
    } else if (R[8] > T && R[8] > R[0] && R[8] > R[4]) {
        q[2] = sqrt((1 + 2 * R[8] - T) / 4);
        q[0] = (1 / (4 * q[2])) * (R[2] + R[6]); q[1] = (1 / (4 * q[2])) * (R[5] + R[7]); q[3] = (1 / (4 * q[2])) * (R[1] - R[3]);
    } else {
        q[3] = sqrt((1 + T) / 4);
        q[0] = (1 / (4 * q[3])) * (R[5] - R[7]); q[1] = (1 / (4 * q[3])) * (R[6] - R[2]); q[2] = (1 / (4 * q[3])) * (R[1] - R[3]);
    }
    const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int i = 0; i < 4; ++i) q[i] /= n;
    if (q[3] < 0) for (int i = 0; i < 4; ++i) q[i] = -q[i];
}

double nrm3(const double* v) { return sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }

// System::initialize (System.cc:115-170): x (26), P (24x24)
void h_initialize(const rvio_vio_cfg& c, const double* w, const double* a, int n_imu, double* x, double* P)
{
    double g[3] = {a[0], a[1], a[2]};
    const double gn = nrm3(g);
    for (int k = 0; k < 3; ++k) g[k] /= gn;
    double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if (c.enable_alignment) {
        const double* zv = g;
        const double ex[3] = {1, 0, 0};
        double t[3], xv[3], yv[3];
        for (int i = 0; i < 3; ++i) t[i] = zv[i] * zv[0] * ex[0] + zv[i] * zv[1] * ex[1] + zv[i] * zv[2] * ex[2];
        for (int k = 0; k < 3; ++k) xv[k] = ex[k] - t[k];
        const double n1 = nrm3(xv);
        for (int k = 0; k < 3; ++k) xv[k] /= n1;
        yv[0] = -zv[2] * xv[1] + zv[1] * xv[2];
        yv[1] = zv[2] * xv[0] - zv[0] * xv[2];
        yv[2] = -zv[1] * xv[0] + zv[0] * xv[1];
        const double n2 = nrm3(yv);
        for (int k = 0; k < 3; ++k) yv[k] /= n2;
        for (int i = 0; i < 3; ++i) { R[3 * i] = xv[i]; R[3 * i + 1] = yv[i]; R[3 * i + 2] = zv[i]; }
    }
    memset(x, 0, 26 * sizeof(double));
    h_quat_from_rot(R, x);
    for (int k = 0; k < 3; ++k) x[7 + k] = g[k];
    if (n_imu > 1)
        for (int k = 0; k < 3; ++k) { x[20 + k] = w[k]; x[23 + k] = a[k] - c.gravity * g[k]; }
    const double dt = 1. / c.imu_rate;
    memset(P, 0, 24 * 24 * sizeof(double));
    for (int k = 0; k < 6; ++k) P[k * 24 + k] = 1e-3 * 1e-3;
    for (int k = 6; k < 9; ++k) P[k * 24 + k] = n_imu * dt * (c.sigma_a * c.sigma_a);
    for (int k = 18; k < 21; ++k) P[k * 24 + k] = n_imu * dt * (c.sigma_wg * c.sigma_wg);
    for (int k = 21; k < 24; ++k) P[k * 24 + k] = n_imu * dt * (c.sigma_wa * c.sigma_wa);
}

template <typename T>
int valloc(rvio_vio* v, T** p, size_t count)
{
    void* q = nullptr;
    RVIO_CUDA_TRY(cudaMalloc(&q, count * sizeof(T) + 16));
    RVIO_CUDA_TRY(cudaMemset(q, 0, count * sizeof(T) + 16));
    v->allocs.push_back(q);
    *p = (T*)q;
    return RVIO_OK;
}
template <typename T>
int vhalloc(rvio_vio* v, T** p, size_t count)
{
    void* q = nullptr;
    RVIO_CUDA_TRY(cudaMallocHost(&q, count * sizeof(T) + 16));
    memset(q, 0, count * sizeof(T) + 16);
    v->hallocs.push_back(q);
    *p = (T*)q;
    return RVIO_OK;
}

// Motion detection + initialisation, System.cc:183-249.  Returns the number of leading IMU rows consumed
// (the rest of the list is what tracking / propagation see), or -1 while still waiting.
int init_step(rvio_vio* v, const double* imu, int n_imu)
{
    const rvio_vio_cfg& c = v->cfg;
    if (!v->moving) {
        double ang[3] = {0, 0, 0}, vel[3] = {0, 0, 0}, displ[3] = {0, 0, 0};
        for (int s = 0; s < n_imu; ++s) {
            const double* w = imu + 8 * s;
            double a[3] = {w[3], w[4], w[5]};
            const double dt = w[7];
            const double an = nrm3(a);
            const double a0[3] = {a[0], a[1], a[2]};
            for (int k = 0; k < 3; ++k) a[k] -= c.gravity * a0[k] / an;
            for (int k = 0; k < 3; ++k) {
                ang[k] += dt * w[k];
                vel[k] += dt * a[k];
                displ[k] += dt * vel[k] + .5 * (dt * dt) * a[k];
            }
        }
        if (nrm3(ang) > c.thr_angle || nrm3(displ) > c.thr_displ) v->moving = true;
    }
    int consumed = 0;
    while (consumed < n_imu) {
        const double* r = imu + 8 * consumed;
        if (!v->moving) {
            for (int k = 0; k < 3; ++k) { v->wm[k] += r[k]; v->am[k] += r[3 + k]; }
            consumed++;
            v->n_imu_count++;
        } else {
            if (v->n_imu_count == 0) {
                for (int k = 0; k < 3; ++k) { v->wm[k] = r[k]; v->am[k] = r[3 + k]; }
                v->n_imu_count = 1;
            } else {
                for (int k = 0; k < 3; ++k) { v->wm[k] /= v->n_imu_count; v->am[k] /= v->n_imu_count; }
            }
            v->ready = true;
            return consumed;
        }
    }
    return -1;
}

}  // namespace

extern "C" int rvio_vio_create(const rvio_vio_cfg* cfg, int device, rvio_vio** out)
{
    RVIO_ARG_CHECK(cfg && out);
    rvio_vio* v = new (std::nothrow) rvio_vio();
    if (!v) return RVIO_ERR_CUDA;
    v->cfg = *cfg; v->device = device;
    int rc = rvio_tracker_create(&cfg->tracker, device, &v->trk);
    if (rc != RVIO_OK) { delete v; return rc; }
    rc = rvio_updater_create(&cfg->updater, device, &v->upd);
    if (rc != RVIO_OK) { rvio_tracker_destroy(v->trk); delete v; return rc; }
    v->stream = tracker_stream(v->trk);
    RVIO_CUDA_TRY(cudaStreamCreateWithFlags(&v->side, cudaStreamNonBlocking));
    RVIO_CUDA_TRY(cudaStreamCreateWithFlags(&v->dets, cudaStreamNonBlocking));
    RVIO_CUDA_TRY(cudaEventCreateWithFlags(&v->ev_det_done, cudaEventDisableTiming));
    RVIO_CUDA_TRY(cudaEventCreateWithFlags(&v->ev_frame_in, cudaEventDisableTiming));
    RVIO_CUDA_TRY(cudaEventCreateWithFlags(&v->ev_prop_done, cudaEventDisableTiming));
    RVIO_CUDA_TRY(cudaEventCreateWithFlags(&v->ev_bookkeep_done, cudaEventDisableTiming));
    RVIO_CUDA_TRY(cudaEventCreateWithFlags(&v->ev_side_done, cudaEventDisableTiming));
    v->timeline = false;
    v->use_graphs = true; v->graph_launches = 0;
    v->copys = nullptr; v->pref_next = 0; v->pref_last = -1; v->pref_hits = 0; v->n_steps = 0; v->pref_step[0] = v->pref_step[1] = 0;
    v->pend_src = nullptr; v->pend_stride = 0; v->pend_step = 0;
    for (int k = 0; k < 2; ++k) { v->ev_pref[k] = nullptr; v->d_pref[k] = nullptr; v->pref_src[k] = nullptr; }
    for (int k = 0; k < 8; ++k) { RVIO_CUDA_TRY(cudaEventCreate(&v->tl[k])); v->tl_ms[k] = 0.f; }
    RVIO_CUDA_TRY(cudaMalloc((void**)&v->d_stamps, sizeof(unsigned long long) * 8));
    RVIO_CUDA_TRY(cudaMemset(v->d_stamps, 0, sizeof(unsigned long long) * 8));
    v->window = cfg->tracker.max_track_len - 1;             // System.cc:71-72
    v->min_clones = cfg->tracker.min_track_len - 1;         // System.cc:74-75
    v->F = cfg->tracker.n_features; v->Fu = (v->F + 1) / 2;
    RVIO_ARG_CHECK(cfg->updater.max_clones >= v->window && cfg->updater.max_features >= v->Fu &&
                   cfg->updater.max_track_len >= cfg->tracker.max_track_len);
    const size_t xmax = xdim_of(v->window), dmax = d_of(v->window);
    for (int k = 0; k < 2; ++k) {
        if ((rc = valloc(v, &v->d_x[k], xmax)) != RVIO_OK) return rc;
        if ((rc = valloc(v, &v->d_P[k], dmax * dmax)) != RVIO_OK) return rc;
    }
    if ((rc = valloc(v, &v->d_pose, 8)) != RVIO_OK) return rc;
    if ((rc = valloc(v, &v->d_imu, 512 * 8 + 2)) != RVIO_OK) return rc;
    if ((rc = valloc(v, &v->d_cand, (size_t)v->F + 1)) != RVIO_OK) return rc;
    if ((rc = vhalloc(v, &v->h_pose, 8)) != RVIO_OK) return rc;
    if ((rc = vhalloc(v, &v->h_imu, 512 * 8 + 2)) != RVIO_OK) return rc;
    if ((rc = vhalloc(v, &v->h_cand, 2 * ((size_t)v->F + 1))) != RVIO_OK) return rc;
    if ((rc = vhalloc(v, &v->h_cnt, 8)) != RVIO_OK) return rc;
    if ((rc = vhalloc(v, &v->h_detctrl, 4)) != RVIO_OK) return rc;
    memset(v->h_detctrl, 0, 4 * sizeof(int));
    if ((rc = vhalloc(v, &v->h_state, xmax + dmax * dmax)) != RVIO_OK) return rc;
    v->xi = v->pi = 0;
    v->moving = v->ready = false;
    for (int k = 0; k < 3; ++k) { v->wm[k] = 0; v->am[k] = 0; }
    v->n_imu_count = 0; v->n_clones = 0; v->n_img_after_init = 0;
    memset(&v->last_info, 0, sizeof v->last_info);
    // FeatureDetector.cc:40-46 (int / float member types as in FeatureDetector.h:61-74)
    const int W = cfg->tracker.width, H = cfg->tracker.height;
    v->gc = (int)floor((double)W / (float)cfg->block_x);
    v->gr = (int)floor((double)H / (float)cfg->block_y);
    RVIO_ARG_CHECK(v->gc > 0 && v->gr > 0);
    v->offx = (int)(.5 * (W - v->gc * (float)cfg->block_x));
    v->offy = (int)(.5 * (H - v->gr * (float)cfg->block_y));
    v->max_per_block = (int)((float)v->F / (v->gc * v->gr));
    if (.75 * v->max_per_block >= kFindNewerCellCap) { set_error("rvio_vio_create", "grid cell capacity exceeded"); return RVIO_ERR_CAPACITY; }
    *out = v;
    return RVIO_OK;
}

extern "C" void rvio_vio_destroy(rvio_vio* v)
{
    if (!v) return;
    cudaSetDevice(v->device);
    cudaStreamSynchronize(v->stream);
    cudaStreamSynchronize(v->side);
    if (v->copys) {
        cudaStreamSynchronize(v->copys);
        for (int k = 0; k < 2; ++k) { cudaEventDestroy(v->ev_pref[k]); cudaFree(v->d_pref[k]); }
        cudaStreamDestroy(v->copys);
    }
    for (auto& kv : v->graphs) cudaGraphExecDestroy(kv.second);
    cudaEventDestroy(v->ev_frame_in); cudaEventDestroy(v->ev_prop_done); cudaEventDestroy(v->ev_bookkeep_done); cudaEventDestroy(v->ev_side_done);
    cudaStreamDestroy(v->side); cudaStreamDestroy(v->dets); cudaEventDestroy(v->ev_det_done);
    for (void* p : v->allocs) cudaFree(p);
    for (void* p : v->hallocs) cudaFreeHost(p);
    shard_comm_destroy(&v->shard);
    rvio_updater_destroy(v->upd);
    rvio_tracker_destroy(v->trk);
    delete v;
}

// (stage timeline) one-thread kernel that writes the GPU's global timer: unlike event-record nodes, it can be read back
// when the frame is replayed as a graph
__global__ void k_stamp(unsigned long long* out) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); *out = t; }

struct FrameOutcome { bool committable, ran_update; };

// The frame's results (pose, update counters, detector control block, tracker scalars) go to the host in ONE step: a small
// kernel stores them into the pinned host buffers directly (pinned memory is device-addressable under unified addressing),
// instead of four dependent D2H copy nodes at the end of the frame's critical path.
__global__ void k_results(const double* pose, double* h_pose, const double* cnt, double* h_cnt, const int* det, int* h_det,
                          const int* sc, int* h_sc, int sc_words)
{
    const int t = threadIdx.x;
    if (t < 7) h_pose[t] = pose[t];
    if (cnt && t >= 8 && t < 16) h_cnt[t - 8] = cnt[t - 8];
    if (det && t >= 16 && t < 20) h_det[t - 16] = det[t - 16];
    for (int i = t; i < sc_words; i += blockDim.x) h_sc[i] = sc[i];
    __threadfence_system();
}

// Everything one frame puts on the two streams, from the IMU upload to the device->host copies of the results; no
// synchronisation and no host-visible result is touched here, so the sequence can be captured into a CUDA graph and
// replayed (with rvio::t_replay set the stream operations are skipped and only the host-side bookkeeping advances).
static int enqueue_frame(rvio_vio* v, bool staged, const uint8_t* img_host, int width, int height, int stride, int channels,
                         const uint8_t* img_dev, int pitch, const double* imu, int n_imu, size_t imu_bytes,
                         const float2* cand_dev, bool cand_upload, int n_cand, int cand_cap, int cand_filtered, bool use_det,
                         FrameOutcome* out)
{
    cudaStream_t s = v->stream, side = v->side;
    const int N = v->n_clones;
    const int xdim = xdim_of(N), d = d_of(N);
    const int* hdr = reinterpret_cast<const int*>(v->d_imu);
    out->committable = false; out->ran_update = false;
    // ---- propagation (System.cc:263) on the side stream: it only needs last frame's x, P and the IMU samples, so it
    //      overlaps the whole tracker; the solve waits for it.
    RVIO_ENQ(cudaEventRecord(v->ev_frame_in, s));
    RVIO_ENQ(cudaStreamWaitEvent(side, v->ev_frame_in, 0));
    const int xi0 = v->xi, pi0 = v->pi;             // prior (pre-propagation) buffers: also what k_feature reads
    RVIO_ENQ(cudaMemcpyAsync(v->d_imu, v->h_imu, imu_bytes, cudaMemcpyHostToDevice, side));
    {
        PropagateParams pp;
        pp.x_in = v->d_x[xi0]; pp.P_in = v->d_P[pi0]; pp.xdim = xdim; pp.d = d;
        pp.imu = v->d_imu + 2; pp.n_imu = n_imu; pp.hdr = hdr; pp.x_out = v->d_x[1 - xi0]; pp.P_out = v->d_P[1 - pi0];
        pp.c.gravity = v->cfg.gravity; pp.c.small_angle = v->cfg.tracker.small_angle;
        pp.c.sigma_g = v->cfg.sigma_g; pp.c.sigma_wg = v->cfg.sigma_wg; pp.c.sigma_a = v->cfg.sigma_a; pp.c.sigma_wa = v->cfg.sigma_wa;
        int r2 = launch_propagate(side, pp);
        if (r2 != RVIO_OK) return r2;
        RVIO_ENQ(cudaEventRecord(v->ev_prop_done, side));
        v->xi = 1 - xi0; v->pi = 1 - pi0;
    }

    // ---- visual tracking (System.cc:258) on the main stream
    if (v->timeline) RVIO_LAUNCH(k_stamp, 1, 1, 0, s, v->d_stamps + 0);
    int rc;
    const int world = v->shard.world, srank = v->shard.rank;
    if (world > 1) {
        // feature-sharded: LK for this rank's share of the feature indices, one all-gather of the per-feature results on
        // this stream, RANSAC + bookkeeping replicated (bit-identical on every rank)
        rc = tracker_enqueue_frame_sharded(v->trk, img_host, width, height, stride, channels, img_dev, pitch, staged, imu, n_imu, srank, world);
        if (rc < 0) return rc;
        if (rc == RVIO_OK) {
            const TrackerBuffers* Bs = tracker_buffers(v->trk);
            int r2 = shard_allgather_lk(&v->shard, s, Bs->lk, Bs->un, Bs->status, tracker_shard_size(v->trk, world));
            if (r2 != RVIO_OK) return r2;
            r2 = tracker_enqueue_ransac(v->trk);
            if (r2 != RVIO_OK) return r2;
        }
    } else if (staged) rc = tracker_enqueue_frame_staged(v->trk, imu, n_imu);
    else if (img_dev) rc = tracker_enqueue_frame_dev(v->trk, img_dev, pitch, imu, n_imu);
    else rc = tracker_enqueue_frame_host(v->trk, img_host, width, height, stride, channels, imu, n_imu);
    if (rc < 0) return rc;
    if (v->timeline) RVIO_LAUNCH(k_stamp, 1, 1, 0, s, v->d_stamps + 1);          // tracker kernels enqueued (upload .. bookkeeping)
    const int* n_cand_dev = hdr + 1;
    if (use_det && rc != RVIO_NO_FEATURES) {
        // FeatureDetector::DetectWithSubPix on its own stream: it only needs the equalised level 0, so it runs beside
        // the pyramid, LK, RANSAC and the propagation; FindNewer (or the seeding) waits for it
        Detector* D = tracker_detector(v->trk);
        RVIO_ENQ(cudaStreamWaitEvent(v->dets, tracker_level0_event(v->trk), 0));
        int r2 = detector_enqueue(D, v->dets, *tracker_level0(v->trk), rc == RVIO_FIRST_IMAGE ? 1 : 2, v->cfg.min_dist, v->cfg.qual_lvl);
        if (r2 != RVIO_OK) return r2;
        RVIO_ENQ(cudaEventRecord(v->ev_det_done, v->dets));
        cand_dev = D->out; n_cand_dev = &D->ctrl->n_out; n_cand = cand_cap = v->F; cand_filtered = 0;
    }
    if (n_cand > 0 && cand_upload && !use_det)
        RVIO_ENQ(cudaMemcpyAsync(v->d_cand, v->h_cand, sizeof(float) * 2 * cand_cap, cudaMemcpyHostToDevice, s));
    bool side_refill = false;
    if (rc == RVIO_FIRST_IMAGE) {
        if (use_det) RVIO_ENQ(cudaStreamWaitEvent(s, v->ev_det_done, 0));
        if (n_cand > 0) { int r2 = tracker_enqueue_seed_dev(v->trk, cand_dev, n_cand, use_det ? n_cand_dev : nullptr); if (r2 != RVIO_OK) return r2; }
        out->committable = true;
    } else if (rc == RVIO_OK) {
        if (n_cand > 0) {
            // FindNewer + refill only prepare the NEXT frame's feature set: run them beside the update
            RVIO_ENQ(cudaEventRecord(v->ev_bookkeep_done, s));
            RVIO_ENQ(cudaStreamWaitEvent(side, v->ev_bookkeep_done, 0));
            if (use_det) RVIO_ENQ(cudaStreamWaitEvent(side, v->ev_det_done, 0));
            FindNewerParams fp;
            fp.B = *tracker_buffers(v->trk); fp.cand = cand_dev; fp.n_cand = cand_cap; fp.n_cand_dev = n_cand_dev; fp.raw = cand_filtered ? 1 : 0;
            fp.W = v->cfg.tracker.width; fp.H = v->cfg.tracker.height; fp.gc = v->gc; fp.gr = v->gr;
            fp.offx = v->offx; fp.offy = v->offy; fp.max_per_block = v->max_per_block;
            fp.bx = (float)v->cfg.block_x; fp.by = (float)v->cfg.block_y; fp.min_dist = v->cfg.min_dist;
            fp.cam = *tracker_cam(v->trk);
            int r2 = launch_find_newer_refill(side, fp);
            if (r2 != RVIO_OK) return r2;
            side_refill = true;
        }
        out->committable = true;
    }

    // ---- update (System.cc:266-277).  The per-feature kernel and the normal terms read only the clone states and the
    //      clone-clone covariance block, which propagation leaves untouched (PreIntegrator.cc:186-192 rewrites the
    //      IMU block and the cross terms only): they run on the prior buffers, concurrently with k_propagate.
    if (N > v->min_clones) {
        const TrackerBuffers* B = tracker_buffers(v->trk);
        // k_feature directly follows k_ransac_bookkeep on the main stream unless a candidate upload or a stage stamp sits between
        updater_hint_kernel_predecessor(v->upd, rc == RVIO_OK && world == 1 && !(n_cand > 0 && cand_upload && !use_det));
        int r2 = updater_enqueue_normal_terms(v->upd, s, v->d_x[xi0], xdim, v->d_P[pi0], d, B->up_types, B->up_off, B->up_xy,
                                              v->Fu, &B->sc->n_up, srank, world);
        if (r2 != RVIO_OK) return r2;
        if (world > 1) {                                               // the single reduce of the sharded update, in stream
            int cnt = 0;
            double* red = updater_reduce_dev(v->upd, &cnt);
            r2 = shard_allreduce_terms(&v->shard, s, red, cnt);
            if (r2 != RVIO_OK) return r2;
        }
        if (v->timeline) RVIO_LAUNCH(k_stamp, 1, 1, 0, s, v->d_stamps + 2);      // per-feature + normal terms done
        RVIO_ENQ(cudaStreamWaitEvent(s, v->ev_prop_done, 0));
        if (v->timeline) RVIO_LAUNCH(k_stamp, 1, 1, 0, s, v->d_stamps + 3);      // (waited for propagation)
        r2 = updater_enqueue_solve_on(v->upd, s, v->d_x[v->xi], v->d_P[v->pi], v->d_x[1 - v->xi], v->d_P[1 - v->pi]);
        if (r2 != RVIO_OK) return r2;
        v->xi = 1 - v->xi; v->pi = 1 - v->pi;
        out->ran_update = true;
        if (v->timeline) RVIO_LAUNCH(k_stamp, 1, 1, 0, s, v->d_stamps + 4);      // solve done
    } else {
        RVIO_ENQ(cudaStreamWaitEvent(s, v->ev_prop_done, 0));
    }
    // ---- augmentation + composition (System.cc:280-365)
    {
        AugmentParams ap;
        ap.x = v->d_x[v->xi]; ap.P_in = v->d_P[v->pi]; ap.P_out = v->d_P[1 - v->pi];
        ap.d = d; ap.N = N; ap.window = v->window; ap.do_augment = v->n_img_after_init > 1 ? 1 : 0; ap.pose_out = v->d_pose;
        int r2 = launch_augment_compose(s, ap);
        if (r2 != RVIO_OK) return r2;
        v->pi = 1 - v->pi;
        if (ap.do_augment && N < v->window) v->n_clones = N + 1;
    }
    if (v->timeline) RVIO_LAUNCH(k_stamp, 1, 1, 0, s, v->d_stamps + 5);          // augmentation + composition done
    if (side_refill) {
        RVIO_ENQ(cudaEventRecord(v->ev_side_done, side));
        RVIO_ENQ(cudaStreamWaitEvent(s, v->ev_side_done, 0));
    }
    {
        const bool det_out = use_det && rc != RVIO_NO_FEATURES && (rc == RVIO_FIRST_IMAGE || side_refill);      // main stream is ordered after the detector here
        RVIO_LAUNCH(k_results, 1, 64, 0, s, v->d_pose, v->h_pose, out->ran_update ? updater_counters_dev(v->upd) : nullptr, v->h_cnt,
                    det_out ? reinterpret_cast<const int*>(tracker_detector(v->trk)->ctrl) : nullptr, reinterpret_cast<int*>(v->h_detctrl),
                    reinterpret_cast<const int*>(tracker_buffers(v->trk)->sc), reinterpret_cast<int*>(const_cast<TrackerScalars*>(tracker_host_scalars(v->trk))),
                    (int)(sizeof(TrackerScalars) / sizeof(int)));
    }
    const int rsc = RVIO_OK;
    if (v->timeline) RVIO_LAUNCH(k_stamp, 1, 1, 0, s, v->d_stamps + 6);          // end of the frame's device work (inside the frame graph when one is replayed)
    return rsc;
}

// a driver call: its answer is remembered per buffer address
static bool host_buffer_is_pinned(rvio_vio* v, const uint8_t* p)
{
    auto pc = v->pin_cache.find(p);
    if (pc != v->pin_cache.end()) return pc->second;
    cudaPointerAttributes pa;
    const bool pinned = cudaPointerGetAttributes(&pa, p) == cudaSuccess && pa.type == cudaMemoryTypeHost;
    if (!pinned) cudaGetLastError();
    if (v->pin_cache.size() < 4096) v->pin_cache.emplace(p, pinned);
    return pinned;
}

// Enqueues the H2D copy of the noted announcement on the copy stream (rvio_vio_prefetch has created the stream, events and slots).
static int issue_pending_prefetch(rvio_vio* v)
{
    const uint8_t* img = v->pend_src;
    if (!img) return RVIO_OK;
    v->pend_src = nullptr;
    const size_t W = (size_t)v->cfg.tracker.width, H = (size_t)v->cfg.tracker.height;
    int slot = v->pref_next;
    for (int k = 0; k < 2; ++k) if (v->pref_src[k] == img) slot = k;      // announced twice: refresh the same slot
    v->pref_next = 1 - slot;
    // (the slot's previous reader, a step's device-to-device copy on the main stream, has completed: steps are synchronous, and the
    // step in flight reads the OTHER slot)
    RVIO_CUDA_TRY(cudaMemcpy2DAsync(v->d_pref[slot], W, img, (size_t)v->pend_stride, W, H, cudaMemcpyHostToDevice, v->copys));
    RVIO_CUDA_TRY(cudaEventRecord(v->ev_pref[slot], v->copys));
    v->pref_src[slot] = img; v->pref_last = slot; v->pref_step[slot] = v->pend_step;
    return RVIO_OK;
}

static int vio_step_impl(rvio_vio* v, const uint8_t* img_host, int width, int height, int stride, int channels,
                         const uint8_t* img_dev, int pitch, const double* imu, int n_imu,
                         const float* cand_host, const float* cand_dev_in, int n_cand, int cand_filtered,
                         double* pose_out, int* pose_valid)
{
    RVIO_ARG_CHECK(v && pose_out && pose_valid && (n_imu == 0 || imu) && n_imu <= 512);
    const bool use_det = n_cand < 0;                      // corners from the device detector
    if (use_det) { n_cand = 0; cand_host = nullptr; cand_dev_in = nullptr; }
    *pose_valid = 0;
    RVIO_CUDA_TRY(cudaSetDevice(v->device));
    if (n_imu < 2) return RVIO_OK;                          // InputBuffer.cc:76-77
    cudaStream_t s = v->stream;
    timespec h0, h1, h2;
    clock_gettime(CLOCK_MONOTONIC, &h0);
    if (!v->ready) {
        const int used = init_step(v, imu, n_imu);
        if (used < 0) return RVIO_OK;
        double x0[26], P0[576];
        h_initialize(v->cfg, v->wm, v->am, v->n_imu_count, x0, P0);
        memcpy(v->h_state, x0, sizeof x0);
        memcpy(v->h_state + 26, P0, sizeof P0);
        v->xi = v->pi = 0;
        RVIO_CUDA_TRY(cudaMemcpyAsync(v->d_x[0], v->h_state, sizeof x0, cudaMemcpyHostToDevice, s));
        RVIO_CUDA_TRY(cudaMemcpyAsync(v->d_P[0], v->h_state + 26, sizeof P0, cudaMemcpyHostToDevice, s));
        RVIO_CUDA_TRY(cudaStreamSynchronize(s));
        imu += 8 * used; n_imu -= used;
    }
    v->n_img_after_init++;
    if (n_cand > v->F) n_cand = v->F;

    // ---- per-frame host staging (pinned): frame header + IMU rows, corner candidates
    int* hh = reinterpret_cast<int*>(v->h_imu);
    hh[0] = n_imu; hh[1] = n_cand; hh[2] = hh[3] = 0;
    memcpy(v->h_imu + 2, imu, sizeof(double) * 8 * n_imu);
    if (n_cand > 0 && !cand_dev_in) memcpy(v->h_cand, cand_host, sizeof(float) * 2 * n_cand);

    const bool was_first = tracker_is_first(v->trk);
    // ---- a steady-state frame (window full, features being tracked) is replayed as a CUDA graph
    const bool steady = v->use_graphs && g_profile_on.load(std::memory_order_relaxed) == 0 &&
                        !tracker_is_first(v->trk) && tracker_n_track(v->trk) > 0 && v->n_clones == v->window &&
                        v->n_clones > v->min_clones && v->n_img_after_init > 1;
    FrameOutcome fo;
    int rc;
    bool host_pinned = false;
    if (steady) {
        const int imu16 = (n_imu + 15) / 16;
        const size_t imu_bytes = 16 + sizeof(double) * 8 * 16 * (size_t)imu16;
        // inputs that arrive in caller-owned device memory are copied into the pipeline's own buffers first (the graph's
        // kernel arguments are fixed addresses)
        if (img_dev) {
            size_t gp; uint8_t* g = tracker_gray(v->trk, &gp);
            RVIO_CUDA_TRY(cudaMemcpy2DAsync(g, gp, img_dev, pitch, v->cfg.tracker.width, v->cfg.tracker.height, cudaMemcpyDeviceToDevice, s));
        } else if (channels == 1) {
            // a single-channel frame in PINNED host memory (cudaHostAlloc / cudaHostRegister by the caller) is DMA'd straight
            // into the pipeline's gray buffer: no staging copy on the host; the frame graph then runs its "staged" variant
            // (the attribute query is a driver call: its answer is remembered per buffer address)
            const bool pinned = host_buffer_is_pinned(v, img_host);
            if (pinned) {
                size_t gp; uint8_t* g = tracker_gray(v->trk, &gp);
                const size_t wbytes = (size_t)v->cfg.tracker.width;
                if ((size_t)stride == wbytes && gp == wbytes)
                    RVIO_CUDA_TRY(cudaMemcpyAsync(g, img_host, wbytes * v->cfg.tracker.height, cudaMemcpyHostToDevice, s));
                else
                    RVIO_CUDA_TRY(cudaMemcpy2DAsync(g, gp, img_host, stride, wbytes, v->cfg.tracker.height, cudaMemcpyHostToDevice, s));
                host_pinned = true;
            }
        }
        if (n_cand > 0 && cand_dev_in)
            RVIO_CUDA_TRY(cudaMemcpyAsync(v->d_cand, cand_dev_in, sizeof(float) * 2 * n_cand, cudaMemcpyDeviceToDevice, s));
        const bool staged_in = img_dev != nullptr || host_pinned;
        const uint64_t key = (uint64_t)use_det << 20 | (uint64_t)(v->timeline ? 1 : 0) << 40 | (uint64_t)(g_pdl_on.load(std::memory_order_relaxed) != 0) << 41 |     // (the stage events are graph nodes of their own variant)
                            
                             (uint64_t)tracker_parity(v->trk) | (uint64_t)v->xi << 1 | (uint64_t)v->pi << 2 | (uint64_t)(n_cand > 0) << 3 |
                             (uint64_t)(cand_filtered != 0) << 4 | (uint64_t)staged_in << 5 | (uint64_t)(channels & 7) << 6 |
                             (uint64_t)(cand_dev_in != nullptr) << 9 | (uint64_t)imu16 << 10;
        auto it = v->graphs.find(key);
        const bool cand_upload = n_cand > 0 && !cand_dev_in;
        if (it != v->graphs.end()) {
            t_replay = true;
            rc = enqueue_frame(v, staged_in, img_host, width, height, stride, channels, nullptr, 0, imu, n_imu, imu_bytes,
                               v->d_cand, cand_upload, n_cand, v->F, cand_filtered, use_det, &fo);
            t_replay = false;
            if (rc != RVIO_OK) return rc;
            RVIO_CUDA_TRY(cudaGraphLaunch(it->second, s));
        } else {
            RVIO_CUDA_TRY(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
            rc = enqueue_frame(v, staged_in, img_host, width, height, stride, channels, nullptr, 0, imu, n_imu, imu_bytes,
                               v->d_cand, cand_upload, n_cand, v->F, cand_filtered, use_det, &fo);
            cudaGraph_t g = nullptr;
            const cudaError_t ce = cudaStreamEndCapture(s, &g);
            if (rc != RVIO_OK) { if (g) cudaGraphDestroy(g); return rc; }
            if (ce != cudaSuccess) { set_error("cudaStreamEndCapture", cudaGetErrorString(ce)); return RVIO_ERR_CUDA; }
            cudaGraphExec_t ex = nullptr;
            const cudaError_t ci = cudaGraphInstantiate(&ex, g, 0);
            cudaGraphDestroy(g);
            if (ci != cudaSuccess) { set_error("cudaGraphInstantiate", cudaGetErrorString(ci)); return RVIO_ERR_CUDA; }
            v->graphs.emplace(key, ex);
            RVIO_CUDA_TRY(cudaGraphLaunch(ex, s));
        }
        v->graph_launches++;
    } else {
        const float2* cand_dev = cand_dev_in ? reinterpret_cast<const float2*>(cand_dev_in) : v->d_cand;
        rc = enqueue_frame(v, false, img_host, width, height, stride, channels, img_dev, pitch, imu, n_imu,
                           16 + sizeof(double) * 8 * (size_t)n_imu, cand_dev, n_cand > 0 && !cand_dev_in, n_cand, n_cand, cand_filtered, use_det, &fo);
        if (rc != RVIO_OK) return rc;
    }
    if (v->pend_src) {                                      // an announced frame: its upload starts now, beside this frame's work
        const int rp = issue_pending_prefetch(v);
        if (rp != RVIO_OK) return rp;
    }
    clock_gettime(CLOCK_MONOTONIC, &h1);
    int r3 = tracker_wait(v->trk);                          // the one synchronisation of the frame
    if (r3 != RVIO_OK) return r3;
    clock_gettime(CLOCK_MONOTONIC, &h2);
    v->tl_ms[6] = (float)((h1.tv_sec - h0.tv_sec) * 1e3 + (h1.tv_nsec - h0.tv_nsec) * 1e-6);   // host: enqueue
    v->tl_ms[7] = (float)((h2.tv_sec - h1.tv_sec) * 1e3 + (h2.tv_nsec - h1.tv_nsec) * 1e-6);   // host: blocked in the sync
    if (v->timeline) {
        if (cudaMemcpy(v->h_stamps, v->d_stamps, sizeof(unsigned long long) * 8, cudaMemcpyDeviceToHost) == cudaSuccess)
            for (int k = 0; k < 6; ++k) v->tl_ms[k] = (float)((double)(long long)(v->h_stamps[k + 1] - v->h_stamps[k]) * 1e-6);
        else cudaGetLastError();
    }
    if (use_det && was_first && tracker_host_scalars(v->trk)->n_new == 0) {     // Tracker.cc:209-213: nothing detected, still "first image"
        tracker_set_first(v->trk, true);
    }
    if (fo.committable) { r3 = rvio_tracker_commit(v->trk); if (r3 != RVIO_OK) return r3; }
    memcpy(pose_out, v->h_pose, sizeof(double) * 7);
    *pose_valid = 1;
    rvio_update_info inf;
    memset(&inf, 0, sizeof inf);
    if (fo.ran_update) {
        inf.n_feat = tracker_host_scalars(v->trk)->n_up;
        inf.n_good = (int)v->h_cnt[0]; inf.rows_stacked = (int)v->h_cnt[1];
        inf.n_reject_init = (int)v->h_cnt[2]; inf.n_reject_lm = (int)v->h_cnt[3]; inf.n_reject_gate = (int)v->h_cnt[4];
        inf.updated = inf.n_good > 2 ? 1 : 0;
        inf.rank = (int)v->h_cnt[6]; inf.rank_flags = (int)v->h_cnt[7];
    }
    v->last_info = inf;
    if (use_det && v->h_detctrl[3]) {
        // DetCtrl.overflow: more local maxima than the detector keeps, one histogram bin larger than a selection block, or a
        // grid cell with too many corners -- the frame was processed, but from a truncated corner set
        v->h_detctrl[3] = 0;
        set_error("rvio_vio_step", "device detector overflow: the frame was seeded / refilled from a truncated corner set");
        return RVIO_DETECTOR_TRUNCATED;
    }
    return RVIO_OK;
}

// Upload of a frame ahead of its step (the host's System::PushImageData moment, System.h:50 / InputBuffer.cc:42): the H2D copy
// runs on a copy stream beside the frame that is being processed.  Only for single-channel frames in pinned host memory
// (anything else is left to the step itself and the call is a no-op); at most two frames are kept.
extern "C" int rvio_vio_prefetch(rvio_vio* v, const uint8_t* img, int width, int height, int stride_bytes, int channels)
{
    RVIO_ARG_CHECK(v && img && width == v->cfg.tracker.width && height == v->cfg.tracker.height && stride_bytes >= width);
    if (channels != 1) return RVIO_OK;
    RVIO_CUDA_TRY(cudaSetDevice(v->device));
    if (!host_buffer_is_pinned(v, img)) return RVIO_OK;
    const size_t W = (size_t)width, H = (size_t)height;
    if (!v->copys) {
        RVIO_CUDA_TRY(cudaStreamCreateWithFlags(&v->copys, cudaStreamNonBlocking));
        for (int k = 0; k < 2; ++k) {
            RVIO_CUDA_TRY(cudaEventCreateWithFlags(&v->ev_pref[k], cudaEventDisableTiming));
            RVIO_CUDA_TRY(cudaMalloc((void**)&v->d_pref[k], W * H));
        }
    }
    if (v->pend_src && v->pend_src != img) {               // two announcements without a step in between: the older one goes up now
        const int rc = issue_pending_prefetch(v);
        if (rc != RVIO_OK) return rc;
    }
    v->pend_src = img; v->pend_stride = stride_bytes; v->pend_step = v->n_steps;
    return RVIO_OK;
}

// Orders the pipeline's stream after the most recent prefetch (no host blocking): measurement code records its end-of-step
// event after this so that the upload provably lies inside the timed region.
extern "C" int rvio_vio_prefetch_fence(rvio_vio* v, uint64_t* hits)
{
    RVIO_ARG_CHECK(v);
    if (hits) *hits = v->pref_hits;
    if (v->pend_src) { const int rp = issue_pending_prefetch(v); if (rp != RVIO_OK) return rp; }
    if (v->copys && v->pref_last >= 0) {
        RVIO_CUDA_TRY(cudaSetDevice(v->device));
        RVIO_CUDA_TRY(cudaStreamWaitEvent(v->stream, v->ev_pref[v->pref_last], 0));
    }
    return RVIO_OK;
}

extern "C" int rvio_vio_step(rvio_vio* v, const uint8_t* img, int width, int height, int stride_bytes, int channels,
                             const double* imu, int n_imu, const float* cand_px, int n_cand, int cand_filtered,
                             double* pose_out, int* pose_valid)
{
    RVIO_ARG_CHECK(v && img && (n_cand <= 0 || cand_px));
    const uint64_t step_no = v->n_steps++;
    if (v->pend_src == img && channels == 1 && n_imu >= 2) {      // announced and stepped back to back: the upload cannot wait
        RVIO_CUDA_TRY(cudaSetDevice(v->device));
        const int rp = issue_pending_prefetch(v);
        if (rp != RVIO_OK) return rp;
    }
    if (v->copys && channels == 1 && n_imu >= 2)
        for (int k = 0; k < 2; ++k)
            if (v->pref_src[k] && step_no - v->pref_step[k] > 1) v->pref_src[k] = nullptr;      // expired
            else if (v->pref_src[k] == img) {             // uploaded ahead of time by rvio_vio_prefetch: take the device slot
                v->pref_src[k] = nullptr;
                RVIO_ARG_CHECK(width == v->cfg.tracker.width && height == v->cfg.tracker.height);
                RVIO_CUDA_TRY(cudaSetDevice(v->device));
                RVIO_CUDA_TRY(cudaStreamWaitEvent(v->stream, v->ev_pref[k], 0));
                v->pref_hits++;
                return vio_step_impl(v, nullptr, width, height, stride_bytes, 1, v->d_pref[k], width, imu, n_imu, cand_px, nullptr,
                                     n_cand, cand_filtered, pose_out, pose_valid);
            }
    return vio_step_impl(v, img, width, height, stride_bytes, channels, nullptr, 0, imu, n_imu, cand_px, nullptr, n_cand,
                         cand_filtered, pose_out, pose_valid);
}

extern "C" int rvio_vio_step_dev(rvio_vio* v, const uint8_t* img_dev, int pitch_bytes, const double* imu, int n_imu,
                                 const float* cand_px_dev, int n_cand, int cand_filtered, double* pose_out, int* pose_valid)
{
    RVIO_ARG_CHECK(v && img_dev && (n_cand <= 0 || cand_px_dev));
    return vio_step_impl(v, nullptr, 0, 0, 0, 1, img_dev, pitch_bytes, imu, n_imu, nullptr, cand_px_dev, n_cand,
                         cand_filtered, pose_out, pose_valid);
}

extern "C" int rvio_vio_get_state(rvio_vio* v, double* x, int* xdim, double* P, int* d)
{
    RVIO_ARG_CHECK(v && xdim && d);
    RVIO_CUDA_TRY(cudaSetDevice(v->device));
    const int xd = xdim_of(v->n_clones), dd = d_of(v->n_clones);
    *xdim = xd; *d = dd;
    if (x) RVIO_CUDA_TRY(cudaMemcpyAsync(v->h_state, v->d_x[v->xi], sizeof(double) * xd, cudaMemcpyDeviceToHost, v->stream));
    if (P) RVIO_CUDA_TRY(cudaMemcpyAsync(v->h_state + xd, v->d_P[v->pi], sizeof(double) * (size_t)dd * dd, cudaMemcpyDeviceToHost, v->stream));
    RVIO_CUDA_TRY(cudaStreamSynchronize(v->stream));
    if (x) memcpy(x, v->h_state, sizeof(double) * xd);
    if (P) memcpy(P, v->h_state + xd, sizeof(double) * (size_t)dd * dd);
    return RVIO_OK;
}

extern "C" int rvio_vio_get_update_info(rvio_vio* v, rvio_update_info* info)
{
    RVIO_ARG_CHECK(v && info);
    *info = v->last_info;
    return RVIO_OK;
}

extern "C" int rvio_vio_graphs(rvio_vio* v, int enable, uint64_t* graph_launches)
{
    RVIO_ARG_CHECK(v);
    if (enable >= 0) v->use_graphs = enable != 0;
    if (graph_launches) *graph_launches = v->graph_launches;
    return RVIO_OK;
}

extern "C" int rvio_vio_shard_init(rvio_vio* v, int rank, int world, const void* nccl_unique_id)
{
    RVIO_ARG_CHECK(v && nccl_unique_id && world >= 1 && world <= 64 && rank >= 0 && rank < world);
    if (v->n_img_after_init > 0 || v->ready) { set_error("rvio_vio_shard_init", "must be called before the first frame"); return RVIO_ERR_STATE; }
    if (world == 1) return RVIO_OK;
    return shard_comm_create(&v->shard, rank, world, nccl_unique_id, v->device);
}

extern "C" int rvio_vio_shard_probe(rvio_vio* v, int iters, float* us2)
{
    RVIO_ARG_CHECK(v && us2 && iters > 0);
    if (v->shard.world <= 1) { us2[0] = us2[1] = 0.f; return RVIO_OK; }
    RVIO_CUDA_TRY(cudaSetDevice(v->device));
    cudaStream_t s = v->stream;
    const TrackerBuffers* B = tracker_buffers(v->trk);
    const int S = tracker_shard_size(v->trk, v->shard.world);
    const int n = 6 * v->window;
    const int cnt = n * n + n + 8 + n + 1;
    double* scratch = nullptr;
    RVIO_CUDA_TRY(cudaMalloc(&scratch, sizeof(double) * cnt));
    RVIO_CUDA_TRY(cudaMemsetAsync(scratch, 0, sizeof(double) * cnt, s));
    cudaEvent_t e0, e1, e2;
    RVIO_CUDA_TRY(cudaEventCreate(&e0)); RVIO_CUDA_TRY(cudaEventCreate(&e1)); RVIO_CUDA_TRY(cudaEventCreate(&e2));
    int rc = RVIO_OK;
    for (int k = 0; k < 3 && rc == RVIO_OK; ++k) { rc = shard_allgather_lk(&v->shard, s, B->lk, B->un, B->status, S); if (rc == RVIO_OK) rc = shard_allreduce_terms(&v->shard, s, scratch, cnt); }
    cudaEventRecord(e0, s);
    for (int k = 0; k < iters && rc == RVIO_OK; ++k) rc = shard_allgather_lk(&v->shard, s, B->lk, B->un, B->status, S);
    cudaEventRecord(e1, s);
    for (int k = 0; k < iters && rc == RVIO_OK; ++k) rc = shard_allreduce_terms(&v->shard, s, scratch, cnt);
    cudaEventRecord(e2, s);
    cudaStreamSynchronize(s);
    float a = 0.f, b = 0.f;
    cudaEventElapsedTime(&a, e0, e1); cudaEventElapsedTime(&b, e1, e2);
    us2[0] = 1e3f * a / iters; us2[1] = 1e3f * b / iters;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
    cudaFree(scratch);
    return rc;
}

extern "C" rvio_tracker* rvio_vio_tracker(rvio_vio* v) { return v ? v->trk : nullptr; }
extern "C" rvio_updater* rvio_vio_updater(rvio_vio* v) { return v ? v->upd : nullptr; }

// Debug: per-stage CUDA-event times of the main stream for the last step (ms): [tracker, per-feature+normal terms,
// wait for propagation, solve, augment+compose, tail].  enable = 1 switches the instrumentation on.
extern "C" int rvio_vio_timeline(rvio_vio* v, int enable, float* ms8)
{
    RVIO_ARG_CHECK(v);
    v->timeline = enable != 0;
    if (ms8) for (int k = 0; k < 8; ++k) ms8[k] = v->tl_ms[k];
    return RVIO_OK;
}
