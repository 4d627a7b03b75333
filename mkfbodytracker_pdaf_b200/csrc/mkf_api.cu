// mkf_api.cu -- C-ABI entry points of libmkf_b200.so that touch the device.
// See include/mkf_b200.h for the contract and the reference interfaces each call replaces.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <utility>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "mkf_internal.h"
#include "mkf_kernels.cuh"
#include "mkf_runs.cuh"
#include "mkf_heads_tma.cuh"
#include "mkf_frame_small.cuh"
#include "../../include/mkf_expf.h"

// events per profiled update: start | bounds | share keys | slot kernel | repair | resample
#define MKF_PROF_EV 6
static std::atomic<uint64_t> g_launches{0};
extern "C" uint64_t mkf_launch_count(void) { return g_launches.load(); }

#define MKF_LAUNCHED() g_launches.fetch_add(1, std::memory_order_relaxed)

// Launch with the programmatic-stream-serialization attribute (PDL, see mkf_device.cuh): kernels of the per-frame
// chain only.  MKF_PDL=0 in the environment turns the attribute off (plain stream order) for A/B measurements.
static bool pdl_enabled()
{
    static const bool on = [] {
        const char* e = getenv("MKF_PDL");
        return !(e && e[0] == '0');
    }();
    return on;
}
static int sm_count(int device)
{
    static std::atomic<int> cached[64];
    if (device < 0 || device >= 64) return 148;
    int v = cached[device].load(std::memory_order_relaxed);
    if (v == 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || v <= 0) v = 148;
        cached[device].store(v, std::memory_order_relaxed);
    }
    return v;
}
// 0: the next mkf_launch calls of this thread go without the programmatic-launch attribute (update_device_runs, per kernel)
static thread_local int g_pdl_override = -1;

template <class... P, class... A>
static void mkf_launch(void (*kern)(P...), unsigned grid, unsigned block, size_t smem, cudaStream_t stream, A&&... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (pdl_enabled() && g_pdl_override != 0) ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, P(std::forward<A>(args))...); // errors surface through cudaGetLastError()
}

#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) {                                                                          \
            mkf_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);    \
            return MKF_E_CUDA;                                                                            \
        }                                                                                                 \
    } while (0)

extern "C" int mkf_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// ------------------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return MKF_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        if (cudaMalloc(&p, bytes) != cudaSuccess) {
            cudaGetLastError();
            mkf_set_error("cudaMalloc(%zu) failed", bytes);
            return MKF_E_NOMEM;
        }
        cap = bytes;
        return MKF_OK;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

// MKF_MEM_HOST_ASYNC pipeline: host->device copies of an update run on their own stream into double-buffered
// staging, so the copies of frame f+1 overlap the kernels of frame f; the device->host copy of an estimate runs on a
// third stream and overlaps the next frame's kernels.  Events order the three streams; mkf_batch_sync drains all.
struct AsyncIo {
    bool ready = false;
    cudaStream_t s_in = nullptr, s_out = nullptr;
    int in_slot = 0, out_slot = 0;
    DevBuf in_meas[2], in_u0[2], in_u1[2], in_seed[2], out_a[2], out_b[2];
    cudaEvent_t in_done[2] = {nullptr, nullptr};   // copies into slot finished            (recorded on s_in)
    cudaEvent_t in_free[2] = {nullptr, nullptr};   // kernels that read slot finished      (recorded on the batch stream)
    cudaEvent_t out_ready[2] = {nullptr, nullptr}; // estimate wrote slot                  (recorded on the batch stream)
    cudaEvent_t out_done[2] = {nullptr, nullptr};  // copy of slot to the host finished    (recorded on s_out)
    bool in_used[2] = {false, false}, out_used[2] = {false, false};
    int init()
    {
        if (ready) return MKF_OK;
        if (cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking) != cudaSuccess) {
            mkf_set_error("cudaStreamCreate failed");
            return MKF_E_CUDA;
        }
        cudaEvent_t* ev[] = {in_done, in_free, out_ready, out_done};
        for (cudaEvent_t* e : ev)
            for (int i = 0; i < 2; i++)
                if (cudaEventCreateWithFlags(&e[i], cudaEventDisableTiming) != cudaSuccess) {
                    mkf_set_error("cudaEventCreate failed");
                    return MKF_E_CUDA;
                }
        ready = true;
        return MKF_OK;
    }
    void release()
    {
        if (s_in) cudaStreamSynchronize(s_in);
        if (s_out) cudaStreamSynchronize(s_out);
        DevBuf* bufs[] = {in_meas, in_u0, in_u1, in_seed, out_a, out_b};
        for (DevBuf* d : bufs)
            for (int i = 0; i < 2; i++) d[i].release();
        cudaEvent_t* ev[] = {in_done, in_free, out_ready, out_done};
        for (cudaEvent_t* e : ev)
            for (int i = 0; i < 2; i++)
                if (e[i]) cudaEventDestroy(e[i]);
        if (s_in) cudaStreamDestroy(s_in);
        if (s_out) cudaStreamDestroy(s_out);
        s_in = s_out = nullptr;
        ready = false;
    }
};

struct mkf_batch {
    const mkf_model* m = nullptr;
    long long T = 0;
    int N = 0, device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    mkf_layout lay;
    long long total = 0, n_tiles = 0;
    // device state
    double2* st[2] = {nullptr, nullptr};
    int cur = 0; // st[cur] holds the children of the last update (read through `parent`)
    int32_t* parent = nullptr;
    // record sharing (k_slot_update<.., DEDUP>): rep[slot] = record of st[cur] holding the slot's state, written by
    // the slot kernel; src[slot] = rep[parent[slot]], written by the resampler.  shared == false: src is `parent`.
    int32_t *rep = nullptr, *src = nullptr;
    // two-kernel record sharing (k_share_keys -> k_slot_update_heads_direct): list of heads, its
    // length (two counters used alternately: the heads kernel of one frame clears the one the next frame appends
    // with), weight per record
    int4* hd16 = nullptr;
    int* head_count = nullptr;
    int head_flip = 0;
    double* w_rec = nullptr;
    bool share_split = true; // MKF_SHARE_SPLIT=0 at creation: the single-launch variant k_slot_update_shared (A/B runs)
    bool shared = false; // the children of the last update share records (read state through src, not parent)
    // run-length particle sets (mkf_runs.cuh): the current set as a list of (record of st[cur], multiplicity) per track,
    // the head table of the last frame, and the draw / seed of its resample (for the on-demand per-slot replay)
    int2* runs = nullptr;
    int* nruns = nullptr;
    int4* hmeta = nullptr;
    int* nheads = nullptr;
    int* lbase = nullptr; // 2 x T: first list position of every track's heads (contiguous records), ping-pong by lb_flip
    int lb_flip = 0;
    char heads_kernel[64] = ""; // name of the heads' slot kernel the last run-length frame launched (mkf_batch_heads_kernel)
    double2* xs = nullptr; // the heads' means tiled by list position (allocated with the first contiguous-record frame)
    double* u_keep = nullptr;
    uint64_t* seed_keep = nullptr;
    // the estimate of the current set as k_resample_runs left it: est[slot] = [xbar T x d | pose T x D]; two slots so that
    // a device->host copy of frame f (MKF_MEM_HOST_ASYNC) may still be in flight while frame f+1 is computed
    double* est[2] = {nullptr, nullptr};
    int est_slot = 0;
    bool est_valid = false;
    cudaEvent_t est_ready = nullptr, est_done[2] = {nullptr, nullptr};
    bool est_used[2] = {false, false};
    bool aos = false;         // st[cur] holds contiguous records (mkf_heads_tma.cuh) instead of tiles: only between
                              // run-length frames; relayout() converts before anything else reads the state
    bool run_mode = false;    // `runs` describes the current particle set
    bool slots_valid = true;  // parent / src / rep / w_raw describe it too (false after a run-level frame until
                              // ensure_slots() replays the resample per slot)
    bool dedup_ok = true; // MKF_DEDUP=0 in the environment turns the sharing off (A/B measurements)
    bool small_fused = false; // short tracks (9 <= N <= 16) in batches of <= 16384: the frame is k_frame_small
                              // (mkf_frame_small.cuh); MKF_SMALL_FUSED=0 / 1 force the five-launch frame / this one
    const int32_t* gather_index() const { return shared ? src : parent; }
    int32_t* bounds = nullptr;
    uint8_t* ind_tail = nullptr; // T x N: per-slot components behind a wrapped indicator draw (SlotArgs::ind_tail)
    double* w_raw = nullptr;
    double* wsum = nullptr;
    uint32_t* status = nullptr;
    uint32_t* unsorted = nullptr;  // per track: last posterior resample left unsorted parents
    int32_t* chain_last = nullptr; // T x N scratch of the literal alias mode
    // literal alias mode with dynamic run assignment (k_alias_runs + k_slot_update_chain_dyn): the run starts, and four
    // counters {runs listed, runs taken} x 2 used alternately (a frame's walker clears the pair the next frame uses)
    int* alias_list = nullptr;
    int* alias_cnt = nullptr;
    int alias_flip = 0;
    // model constants on this device
    double *d_comp = nullptr, *d_init = nullptr, *d_cw_hi = nullptr, *d_cw_lo = nullptr, *d_wprior = nullptr,
           *d_recon = nullptr, *d_pmean = nullptr, *d_tm = nullptr, *d_tinv = nullptr;
    double* d_small_const = nullptr; // k_frame_small: [comp_const | reconstruction coefficients [c][r]] (SmallTailArgs)
    unsigned small_const_bytes = 0;
    // staging
    DevBuf in_meas, in_u0, in_u1, in_seed, out_a, out_b, in_x, in_p;
    AsyncIo aio;
    // association scratch (arm0 owns)
    // the pose of every track as the last estimate computed it, kept for the next association step (which needs the
    // posterior hand position, src/pf2DRao.cpp:111-116); switched on by the first mkf_batch_associate
    DevBuf pose_cache;
    bool pose_cache_on = false, pose_valid = false;
    bool clear_status_next = false; // the next update_device's first kernel zeroes the status words
    // MKF_MEAS_CAND source of the next update_device call (set and cleared by mkf_batch_associate)
    const double* cm_cand = nullptr;
    const int32_t* cm_bins = nullptr;
    const double* cm_roi = nullptr;
    int cm_C = 0, cm_hand = 0;
    const int32_t* cm_cuts = nullptr; // (T x 2) x 32 run boundaries of the candidate bins (k_resample_warp), or null
    DevBuf as_cand, as_L, as_roi, as_u, as_w, as_gate, as_bins, as_meas, as_wsum, as_hand, as_status, as_seed,
        as_ui, as_up, as_cuts;
    int as_C = 0;
    // optional per-kernel timing (mkf_batch_profile): 4 events per update
    int stage = 3; // see SlotArgs::stage (mkf_kf_apply runs single stages)
    bool prof_on = false;
    unsigned long long* prof_ts = nullptr; // per sampled update: {min CTA start, max CTA end} of the slot kernel
    int prof_ts_cap = 0;
    std::vector<cudaEvent_t> prof_ev;
    int prof_n = 0, prof_every = 1;
    uint64_t prof_tick = 0;
};

static int ensure_slots(mkf_batch* b); // per-slot views of a run-level particle set (defined with update_device)

static bool is_device_ptr(const void* p, int mem)
{
    if (mem == MKF_MEM_DEVICE) return true;
    if (mem == MKF_MEM_HOST || mem == MKF_MEM_HOST_ASYNC) return false;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// input pointer -> device pointer (copying through `stage` when it is host memory)
template <class Tp>
static int in_ptr(mkf_batch* b, const Tp* p, size_t count, int mem, DevBuf& stage, const Tp** out)
{
    if (!p) {
        *out = nullptr;
        return MKF_OK;
    }
    if (is_device_ptr(p, mem)) {
        *out = p;
        return MKF_OK;
    }
    int rc = stage.ensure(count * sizeof(Tp));
    if (rc) return rc;
    CK(cudaMemcpyAsync(stage.p, p, count * sizeof(Tp), cudaMemcpyHostToDevice, b->stream));
    *out = (const Tp*)stage.p;
    return MKF_OK;
}

// output helper: kernels write to dev(); finish() copies back when the user pointer is host memory
template <class Tp>
struct OutPtr {
    Tp* user = nullptr;
    Tp* devp = nullptr;
    size_t count = 0;
    bool host = false;
    int init(mkf_batch* b, Tp* p, size_t n, int mem, DevBuf& stage)
    {
        (void)b;
        user = p;
        count = n;
        if (!p) return MKF_OK;
        if (is_device_ptr(p, mem)) {
            devp = p;
            return MKF_OK;
        }
        host = true;
        int rc = stage.ensure(n * sizeof(Tp));
        if (rc) return rc;
        devp = (Tp*)stage.p;
        return MKF_OK;
    }
    int finish(mkf_batch* b)
    {
        if (user && host) CK(cudaMemcpyAsync(user, devp, count * sizeof(Tp), cudaMemcpyDeviceToHost, b->stream));
        return MKF_OK;
    }
};

static int upload_const(const std::vector<double>& v, double** d)
{
    CK(cudaMalloc((void**)d, v.size() * sizeof(double)));
    CK(cudaMemcpy(*d, v.data(), v.size() * sizeof(double), cudaMemcpyHostToDevice));
    return MKF_OK;
}

extern "C" void mkf_batch_destroy(mkf_batch* b)
{
    if (!b) return;
    cudaSetDevice(b->device);
    if (b->stream) cudaStreamSynchronize(b->stream);
    for (int i = 0; i < 2; i++)
        if (b->st[i]) cudaFree(b->st[i]);
    void* ptrs[] = {b->parent, b->rep, b->src, b->runs, b->nruns, b->hmeta, b->nheads, b->lbase, b->xs, b->u_keep, b->seed_keep, b->est[0], b->est[1], b->hd16, b->head_count, b->w_rec, b->bounds, b->ind_tail, b->w_raw, b->wsum,   b->status, b->unsorted, b->chain_last, b->alias_list, b->alias_cnt, b->prof_ts, b->d_comp, b->d_init,
                    b->d_cw_hi, b->d_cw_lo, b->d_wprior, b->d_recon, b->d_pmean, b->d_tm,   b->d_tinv, b->d_small_const};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    DevBuf* bufs[] = {&b->in_meas, &b->in_u0, &b->in_u1, &b->in_seed, &b->out_a,   &b->out_b,  &b->in_x,   &b->in_p,
                      &b->as_cand, &b->as_L,  &b->as_roi, &b->as_u,   &b->as_w,    &b->as_gate, &b->as_bins,
                      &b->as_meas, &b->as_wsum, &b->as_hand, &b->as_status, &b->as_seed, &b->pose_cache,
                      &b->as_ui,   &b->as_up,  &b->as_cuts};
    for (DevBuf* d : bufs) d->release();
    b->aio.release();
    if (b->est_ready) cudaEventDestroy(b->est_ready);
    for (int i = 0; i < 2; i++)
        if (b->est_done[i]) cudaEventDestroy(b->est_done[i]);
    for (cudaEvent_t e : b->prof_ev) cudaEventDestroy(e);
    if (b->own_stream && b->stream) cudaStreamDestroy(b->stream);
    delete b;
}

extern "C" int mkf_batch_create(mkf_batch** out, const mkf_model* m, int64_t T, int N, int device, void* stream)
{
    if (!out || !m || T <= 0 || N <= 0) {
        mkf_set_error("mkf_batch_create: invalid argument (T=%lld N=%d)", (long long)T, N);
        return MKF_E_INVALID;
    }
    *out = nullptr;
    // slot, record and head-list indices are 32-bit on the device (src / rep / parent arrays, head entries)
    if ((long long)T > 0x7fffffffll / N) {
        mkf_set_error("mkf_batch_create: T*N = %lld exceeds 2^31 - 1 slots per batch (split the tracks over batches)",
                      (long long)T * N);
        return MKF_E_INVALID;
    }
    int ndev = mkf_device_count();
    if (ndev <= 0) {
        mkf_set_error("no CUDA device available: libmkf_b200 has no CPU fallback");
        return MKF_E_CUDA;
    }
    if (device < 0 || device >= ndev) {
        mkf_set_error("mkf_batch_create: device %d out of range (%d visible)", device, ndev);
        return MKF_E_INVALID;
    }
    CK(cudaSetDevice(device));
    mkf_batch* b = new (std::nothrow) mkf_batch;
    if (!b) return MKF_E_NOMEM;
    b->m = m;
    b->T = T;
    b->N = N;
    b->device = device;
    b->lay = m->lay;
    b->total = (long long)T * N;
    b->n_tiles = (b->total + 31) / 32;
    int rc = MKF_OK;
    auto fail = [&](int code) {
        mkf_batch_destroy(b);
        return code;
    };
    if (stream) {
        b->stream = (cudaStream_t)stream;
    } else {
        if (cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking) != cudaSuccess) {
            mkf_set_error("cudaStreamCreate failed");
            return fail(MKF_E_CUDA);
        }
        b->own_stream = true;
    }
    constexpr int CH2 = MKF_CW / 2; // double2 per storage chunk (SlotLay::H)
    const size_t tile_bytes = (size_t)((b->lay.np + CH2 - 1) / CH2 * CH2) * 32 * sizeof(double2);
    auto dmalloc = [&](void** p, size_t bytes) {
        if (cudaMalloc(p, bytes) != cudaSuccess) {
            cudaGetLastError();
            mkf_set_error("cudaMalloc(%zu bytes) failed for a batch of T=%lld N=%d", bytes, (long long)T, N);
            return MKF_E_NOMEM;
        }
        return MKF_OK;
    };
    for (int i = 0; i < 2; i++)
        if ((rc = dmalloc((void**)&b->st[i], (size_t)b->n_tiles * tile_bytes))) return fail(rc);
    if ((rc = dmalloc((void**)&b->parent, (size_t)b->total * sizeof(int32_t)))) return fail(rc);
    {
        const char* e = getenv("MKF_DEDUP");
        b->dedup_ok = !(e && e[0] == '0') && m->prm.alias_mode == MKF_ALIAS_INDEPENDENT;
        const char* e2 = getenv("MKF_SHARE_SPLIT");
        b->share_split = !(e2 && e2[0] == '0');
    }
    {
        // Worth it while the frame is launch-bound: at 4096 x 15 the five-launch frame takes 49.5 us, this one 35.2.  A
        // warp carries its tracks' indicator draw, resample and estimate around the slot arithmetic -- 2 390 instructions
        // where k_slot_update executes 1 580, at two warps per scheduler -- so past ~30 k tracks, where the slot kernel
        // alone runs at the HBM roofline, the separate kernels win (1 M x 15: 4.51 ms against 4.99).
        // MKF_SMALL_FUSED=0 / =1 force either (A/B runs and tests).
        const char* e = getenv("MKF_SMALL_FUSED");
        const bool forced_on = e && e[0] == '1';
        b->small_fused = !(e && e[0] == '0') && N >= 9 && N <= 16 && m->K <= 32 &&
                         m->prm.alias_mode == MKF_ALIAS_INDEPENDENT && (forced_on || T <= 16384);
        if (b->small_fused) {
            if ((rc = dmalloc((void**)&b->est[0], (size_t)T * (m->d + m->D) * sizeof(double))) ||
                (rc = dmalloc((void**)&b->est[1], (size_t)T * (m->d + m->D) * sizeof(double))))
                return fail(rc);
            if (cudaEventCreateWithFlags(&b->est_ready, cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&b->est_done[0], cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&b->est_done[1], cudaEventDisableTiming) != cudaSuccess) {
                mkf_set_error("cudaEventCreate failed");
                return fail(MKF_E_CUDA);
            }
        }
    }
    if (b->dedup_ok && ((rc = dmalloc((void**)&b->rep, (size_t)b->total * sizeof(int32_t))) ||
                        (rc = dmalloc((void**)&b->src, (size_t)b->total * sizeof(int32_t)))))
        return fail(rc);
    if (b->dedup_ok) {
        if ((rc = dmalloc((void**)&b->hd16, (size_t)b->total * sizeof(int4))) ||
            (rc = dmalloc((void**)&b->head_count, 2 * sizeof(int))) ||
            (rc = dmalloc((void**)&b->w_rec, (size_t)b->total * sizeof(double))))
            return fail(rc);
        if (cudaMemset(b->head_count, 0, 2 * sizeof(int)) != cudaSuccess) return fail(MKF_E_CUDA);
        // run-length pipeline (MKF_RUNS=0 keeps the per-slot record-sharing kernels, for A/B runs and tests)
        const char* e3 = getenv("MKF_RUNS");
        if (!(e3 && e3[0] == '0') && b->share_split && N > 64 && N >= 4 * m->K) {
            if ((rc = dmalloc((void**)&b->runs, (size_t)b->total * sizeof(int2))) ||
                (rc = dmalloc((void**)&b->nruns, (size_t)T * sizeof(int))) ||
                (rc = dmalloc((void**)&b->hmeta, (size_t)b->total * sizeof(int4))) ||
                (rc = dmalloc((void**)&b->nheads, (size_t)T * sizeof(int))) ||
                (rc = dmalloc((void**)&b->lbase, (size_t)T * 2 * sizeof(int))) ||
                (rc = dmalloc((void**)&b->u_keep, (size_t)T * sizeof(double))) ||
                (rc = dmalloc((void**)&b->seed_keep, (size_t)T * sizeof(uint64_t))) ||
                (rc = dmalloc((void**)&b->est[0], (size_t)T * (m->d + m->D) * sizeof(double))) ||
                (rc = dmalloc((void**)&b->est[1], (size_t)T * (m->d + m->D) * sizeof(double))))
                return fail(rc);
            if (cudaEventCreateWithFlags(&b->est_ready, cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&b->est_done[0], cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&b->est_done[1], cudaEventDisableTiming) != cudaSuccess) {
                mkf_set_error("cudaEventCreate failed");
                return fail(MKF_E_CUDA);
            }
        }
    }
    if ((rc = dmalloc((void**)&b->bounds, (size_t)T * (m->K + 2) * sizeof(int32_t)))) return fail(rc);
    if (m->K <= 256 && (rc = dmalloc((void**)&b->ind_tail, (size_t)b->total))) return fail(rc);
    if ((rc = dmalloc((void**)&b->w_raw, (size_t)b->total * sizeof(double)))) return fail(rc);
    if ((rc = dmalloc((void**)&b->wsum, (size_t)T * sizeof(double)))) return fail(rc);
    if ((rc = dmalloc((void**)&b->status, (size_t)T * sizeof(uint32_t)))) return fail(rc);
    if ((rc = dmalloc((void**)&b->unsorted, (size_t)T * sizeof(uint32_t)))) return fail(rc);
    if (m->prm.alias_mode == MKF_ALIAS_CV_SHALLOW_LITERAL &&
        (rc = dmalloc((void**)&b->chain_last, (size_t)b->total * sizeof(int32_t))))
        return fail(rc);
    {
        // MKF_ALIAS_DYN=1: the dynamic chain walker (k_alias_runs + k_slot_update_chain_dyn) instead of the
        // thread-per-run-start kernel.  An experiment kept for A/B runs: every lane stays busy, but lanes walking
        // different runs touch 16-byte pieces two or three slots apart and out of step, and DRAM moves 6.8 GB per frame
        // where 2.2 GB are needed (ncu, profiles/r02_ncu_chain_dyn.csv): 1.70 ms per frame against 0.91.
        const char* e4 = getenv("MKF_ALIAS_DYN");
        if (m->prm.alias_mode == MKF_ALIAS_CV_SHALLOW_LITERAL && e4 && e4[0] == '1') {
            if ((rc = dmalloc((void**)&b->alias_list, (size_t)b->total * sizeof(int))) ||
                (rc = dmalloc((void**)&b->alias_cnt, 4 * sizeof(int))))
                return fail(rc);
            if (cudaMemset(b->alias_cnt, 0, 4 * sizeof(int)) != cudaSuccess) return fail(MKF_E_CUDA);
        }
    }
    // the tail lanes of the last tile are read by nobody but keep them defined
    if (cudaMemset(b->st[0], 0, (size_t)b->n_tiles * tile_bytes) != cudaSuccess ||
        cudaMemset(b->st[1], 0, (size_t)b->n_tiles * tile_bytes) != cudaSuccess ||
        cudaMemset(b->status, 0, (size_t)T * sizeof(uint32_t)) != cudaSuccess ||
        cudaMemset(b->unsorted, 0, (size_t)T * sizeof(uint32_t)) != cudaSuccess ||
        cudaMemset(b->w_raw, 0, (size_t)b->total * sizeof(double)) != cudaSuccess ||
        cudaMemset(b->wsum, 0, (size_t)T * sizeof(double)) != cudaSuccess ||
        cudaMemset(b->parent, 0, (size_t)b->total * sizeof(int32_t)) != cudaSuccess ||
        cudaMemset(b->bounds, 0, (size_t)T * (m->K + 2) * sizeof(int32_t)) != cudaSuccess) {
        mkf_set_error("cudaMemset failed");
        return fail(MKF_E_CUDA);
    }
    if ((rc = upload_const(m->comp_const, &b->d_comp)) || (rc = upload_const(m->init_const, &b->d_init)) ||
        (rc = upload_const(m->cw_hi, &b->d_cw_hi)) || (rc = upload_const(m->cw_lo, &b->d_cw_lo)) ||
        (rc = upload_const(m->weights, &b->d_wprior)) || (rc = upload_const(m->recon, &b->d_recon)) ||
        (rc = upload_const(m->pmean, &b->d_pmean)) || (rc = upload_const(m->Tm, &b->d_tm)) ||
        (rc = upload_const(m->Tinv, &b->d_tinv)))
        return fail(rc);
    if (b->small_fused) {
        const int R = m->D + m->d;
        std::vector<double> sc(m->comp_const);
        sc.resize(sc.size() + (size_t)R * m->d + 1, 0.0);
        double* coef = sc.data() + m->comp_const.size();
        for (int r = 0; r < R; r++)
            for (int c = 0; c < m->d; c++)
                coef[(size_t)c * R + r] = r < m->D ? m->recon[(size_t)r * m->d + c] : m->Tinv[(size_t)(r - m->D) * m->d + c];
        sc.resize((m->comp_const.size() + (size_t)R * m->d + 1) / 2 * 2); // (bulk copies move multiples of 16 bytes)
        b->small_const_bytes = (unsigned)(sc.size() * sizeof(double));
        if ((rc = upload_const(sc, &b->d_small_const))) return fail(rc);
    }
    *out = b;
    return MKF_OK;
}

extern "C" int mkf_batch_sync(mkf_batch* b)
{
    if (!b) {
        mkf_set_error("null batch");
        return MKF_E_INVALID;
    }
    CK(cudaSetDevice(b->device));
    CK(cudaStreamSynchronize(b->stream));
    if (b->aio.ready) {
        CK(cudaStreamSynchronize(b->aio.s_in));
        CK(cudaStreamSynchronize(b->aio.s_out));
    }
    return MKF_OK;
}

__global__ void k_count_records(const int32_t* __restrict__ rep, long long total, int N, unsigned long long* out)
{
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool first = false;
    if (s < total) first = (s % N == 0) || rep[s] != rep[s - 1]; // rep is non-decreasing within a track
    const unsigned m = __ballot_sync(0xffffffffu, first);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(out, (unsigned long long)__popc(m));
}

extern "C" int mkf_batch_heads_kernel(mkf_batch* b, char* name, int len)
{
    if (!b || !name || len <= 0) {
        mkf_set_error("mkf_batch_heads_kernel: null argument");
        return MKF_E_INVALID;
    }
    snprintf(name, (size_t)len, "%s", b->heads_kernel);
    return MKF_OK;
}

extern "C" int mkf_batch_shared_records(mkf_batch* b, int64_t* records, int64_t* slots)
{
    if (!b || !records || !slots) {
        mkf_set_error("mkf_batch_shared_records: null argument");
        return MKF_E_INVALID;
    }
    CK(cudaSetDevice(b->device));
    *slots = b->total;
    *records = b->total;
    if (!b->shared) return MKF_OK;
    {
        int rc0 = ensure_slots(b);
        if (rc0) return rc0;
    }
    unsigned long long* d = nullptr;
    CK(cudaMalloc((void**)&d, 8));
    CK(cudaMemsetAsync(d, 0, 8, b->stream));
    k_count_records<<<(unsigned)((b->total + 255) / 256), 256, 0, b->stream>>>(b->rep, b->total, b->N, d);
    MKF_LAUNCHED();
    unsigned long long h = 0;
    cudaError_t e = cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, b->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(b->stream);
    cudaFree(d);
    if (e != cudaSuccess) {
        mkf_set_error("mkf_batch_shared_records: %s", cudaGetErrorString(e));
        return MKF_E_CUDA;
    }
    *records = (int64_t)h;
    return MKF_OK;
}

extern "C" int mkf_batch_join(mkf_batch* b)
{
    if (!b) {
        mkf_set_error("null batch");
        return MKF_E_INVALID;
    }
    CK(cudaSetDevice(b->device));
    AsyncIo& io = b->aio;
    if (!io.ready) return MKF_OK;
    for (int i = 0; i < 2; i++) {
        if (io.in_used[i]) CK(cudaStreamWaitEvent(b->stream, io.in_done[i], 0));
        if (io.out_used[i]) CK(cudaStreamWaitEvent(b->stream, io.out_done[i], 0));
        if (b->est_used[i]) CK(cudaStreamWaitEvent(b->stream, b->est_done[i], 0));
    }
    return MKF_OK;
}

static inline unsigned grid_for(long long n, int bt) { return (unsigned)((n + bt - 1) / bt); }

static int launch_bounds_kernel(mkf_batch* b, const double* d_u, int clear_status = 0)
{
    const mkf_model* m = b->m;
#define LAUNCH_BOUNDS(G)                                                                                       \
    mkf_launch(k_indicator_bounds<G>, grid_for(b->T * G, 128), 128, 0, b->stream, d_u, b->T, b->N, m->K, b->d_cw_hi,       \
                                                                           b->d_cw_lo, b->d_wprior, m->prior_wmax, \
                                                                           b->bounds, b->status, clear_status, b->ind_tail)
    if (m->K <= 16)
        LAUNCH_BOUNDS(16);
    else
        LAUNCH_BOUNDS(32); // K <= 64: two components per lane
#undef LAUNCH_BOUNDS
    MKF_LAUNCHED();
    CK(cudaGetLastError());
    return MKF_OK;
}

extern "C" int mkf_batch_reset(mkf_batch* b, const double* u_init, int mem)
{
    if (!b || !u_init) {
        mkf_set_error("mkf_batch_reset: null argument");
        return MKF_E_INVALID;
    }
    CK(cudaSetDevice(b->device));
    const double* d_u;
    int rc = in_ptr(b, u_init, (size_t)b->T, mem, b->in_u0, &d_u);
    if (rc) return rc;
    CK(cudaMemsetAsync(b->status, 0, (size_t)b->T * sizeof(uint32_t), b->stream));
    CK(cudaMemsetAsync(b->unsorted, 0, (size_t)b->T * sizeof(uint32_t), b->stream));
    if ((rc = launch_bounds_kernel(b, d_u))) return rc;
    b->cur = 0;
    b->aos = false;
    b->shared = false;
    b->run_mode = false;
    b->slots_valid = true;
    b->est_valid = false;
    b->pose_valid = false;
    if (b->m->d == 12)
        k_reset<12><<<grid_for(b->total, 256), 256, 0, b->stream>>>(b->st[0], b->parent, b->bounds, b->d_init,
                                                                    b->total, b->N, b->m->K, b->ind_tail);
    else
        k_reset<10><<<grid_for(b->total, 256), 256, 0, b->stream>>>(b->st[0], b->parent, b->bounds, b->d_init,
                                                                    b->total, b->N, b->m->K, b->ind_tail);
    MKF_LAUNCHED();
    CK(cudaGetLastError());
    return MKF_OK;
}

// function attributes are per device: true the first time a call site runs on the current device
static bool first_on_this_device(std::atomic<uint64_t>& seen)
{
    int dev = 0;
    cudaGetDevice(&dev);
    const uint64_t bit = 1ull << (dev & 63);
    return (seen.fetch_or(bit) & bit) == 0;
}

// weight normalisation + systematic resampling of `nt` tracks: w (nt x L) -> out (nt x N)
static int run_resample(cudaStream_t stream, long long nt, const double* d_w, int L, int N, const double* d_u,
                        int u_stride, int normalise, double* d_wsum, int32_t* d_out, uint32_t* d_status,
                        const uint64_t* d_seeds, int seed_stride, int seed_off, uint32_t bit_fb, uint32_t bit_deg,
                        uint32_t* d_unsorted = nullptr, const int32_t* d_rep = nullptr, int32_t* d_src = nullptr,
                        double* d_w_slot_out = nullptr, const double* d_wsum_in = nullptr, int32_t* d_cut_out = nullptr,
                        bool* cuts_written = nullptr)
{
    if (cuts_written) *cuts_written = false;
    if (L <= 64 && N <= 64) {
        if (d_w_slot_out || d_wsum_in) {
            mkf_set_error("internal: per-record weights are not supported by the small-track resampler");
            return MKF_E_INVALID;
        }
        const size_t smem = (size_t)129 * ((size_t)L * 8 + (size_t)N * 4); // <= 99 KB at L = N = 64
        static std::atomic<uint64_t> seen{0};
        if (first_on_this_device(seen))
            CK(cudaFuncSetAttribute(k_resample_small, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        mkf_launch(k_resample_small, grid_for(nt, 128), 128, smem, stream, d_w, nt, L, N, d_u, u_stride, normalise, d_wsum,
                   d_out, d_status, 1, bit_deg, d_seeds, seed_stride, seed_off, d_unsorted, d_rep, d_src);
    } else if (L <= 32 && !d_rep && !d_src && !d_w_slot_out && !d_wsum_in) {
        // few weights, many outputs (candidate resample): a warp per track
        mkf_launch(k_resample_warp, grid_for(nt, 4), 128, 0, stream, d_w, nt, L, N, d_u, u_stride, normalise, d_wsum, d_out,
                   d_status, 1, bit_fb, bit_deg, d_seeds, seed_stride, seed_off, d_unsorted, d_cut_out);
        if (cuts_written) *cuts_written = d_cut_out != nullptr;
    } else {
        // one CTA per track; wider CTAs for long weight / index vectors so a track's tiles are few
        const int span = L > N ? L : N;
#define MKF_RS_BLOCK(BT)                                                                                             \
    mkf_launch(k_resample_block<BT, 4>, (unsigned)nt, BT, 0, stream, d_w, L, N, d_u, u_stride, normalise, d_wsum, d_out, \
               d_status, 1, bit_fb, bit_deg, d_seeds, seed_stride, seed_off, d_unsorted, d_rep, d_src, d_w_slot_out,     \
               d_wsum_in)
        if (span <= 1024)
            MKF_RS_BLOCK(128);
        else if (span <= 8192)
            MKF_RS_BLOCK(512);
        else
            MKF_RS_BLOCK(1024);
#undef MKF_RS_BLOCK
    }
    MKF_LAUNCHED();
    CK(cudaGetLastError());
    return MKF_OK;
}

// per-slot views of a run-level particle set (mkf_runs.cuh): rep[] from the head table, then the exact per-slot resampler
// replays the last resample from the same head weights, normaliser, draw and seed -> parent[], src[], w_raw[]
// st[cur] between the tile layout and contiguous records, out of place into the idle ping-pong buffer.  In run-length
// mode only the records the last frame wrote are live (head i of track t, i < nheads[t]: at t*N + i in tiles, at
// lbase[t] + i as contiguous records).
static int relayout(mkf_batch* b, bool to_aos)
{
    if (b->aos == to_aos) return MKF_OK;
    const int* cnt = b->run_mode ? b->nheads : nullptr;
    int* lb = b->lbase + (size_t)b->lb_flip * b->T;
    if (b->m->d == 12)
        k_relayout<12><<<(unsigned)b->T, 128, 0, b->stream>>>(b->st[b->cur], b->st[b->cur ^ 1], cnt, b->N, to_aos ? 1 : 0, lb);
    else
        k_relayout<10><<<(unsigned)b->T, 128, 0, b->stream>>>(b->st[b->cur], b->st[b->cur ^ 1], cnt, b->N, to_aos ? 1 : 0, lb);
    MKF_LAUNCHED();
    CK(cudaGetLastError());
    b->cur ^= 1;
    b->aos = to_aos;
    return MKF_OK;
}

static int ensure_slots(mkf_batch* b)
{
    if (b->aos) {
        int rc0 = relayout(b, false);
        if (rc0) return rc0;
    }
    if (!b->run_mode || b->slots_valid) return MKF_OK;
    k_expand_rep<<<grid_for(b->T, 4), 128, 0, b->stream>>>(b->hmeta, b->nheads, b->T, b->N, b->rep);
    MKF_LAUNCHED();
    CK(cudaGetLastError());
    int rc = run_resample(b->stream, b->T, b->w_rec, b->N, b->N, b->u_keep, 1, 1, nullptr, b->parent, b->status,
                          b->seed_keep, 1, 0, MKF_ST_POST_FALLBACK, MKF_ST_POST_DEGENERATE, b->unsorted, b->rep, b->src,
                          b->w_raw, b->wsum);
    if (rc) return rc;
    b->slots_valid = true;
    return MKF_OK;
}

static void fill_slot_args(mkf_batch* b, SlotArgs& a, const double* d_meas, int meas_layout)
{
    const mkf_model* m = b->m;
    a.st_in = b->st[b->cur];
    a.st_out = b->st[b->cur ^ 1];
    a.parent = b->parent;
    a.src = b->gather_index();
    a.cand = b->cm_cand;
    a.bins = b->cm_bins;
    a.roi = b->cm_roi;
    a.cand_C = b->cm_C;
    a.hand = b->cm_hand;
    a.neck = m->prm.neck_offset;
    a.hd16 = b->hd16;
    a.head_count = b->head_count ? b->head_count + b->head_flip : nullptr;
    a.w_rec = b->w_rec;
    a.bounds = b->bounds;
    a.ind_tail = b->ind_tail;
    a.meas = d_meas;
    a.comp_const = b->d_comp;
    a.w_raw = b->w_raw;
    a.status = b->status;
    a.total = b->total;
    a.N = b->N;
    a.K = m->K;
    a.meas_layout = meas_layout;
    a.chol_mode = m->prm.chol_mode;
    a.stage = b->stage;
    a.alias_chain = (m->prm.alias_mode == MKF_ALIAS_CV_SHALLOW_LITERAL) ? 1 : 0;
    a.unsorted = b->unsorted;
    for (int r = 0; r < MKF_M; r++) a.bh[r] = m->BH[r];
    a.r = m->prm.meas_noise_var;
}

// CTAs of k_slot_update_heads_direct per SM: 2 are resident; with more, the later ones start as earlier ones finish and
// the gather / arithmetic / store phases of the resident CTAs drift apart (measured at 4096 x 500: 77.7 us with 2, 74.5
// with 4, 72.1 with 12-24).  The kernel strides over the list, so any grid is correct.  MKF_HEADS_CTAS_PER_SM overrides.
// The head count is only known on the device; ~1/8 of the slots is what a shared-measurement frame keeps in steady state,
// so the default grid is one CTA per 1024 slots, at least 12 and at most 128 per SM: large batches then also run one
// step per CTA (32768 x 500: 0.765 -> 0.80 of the HBM peak; profiles/r02_heads_sweep.jsonl).
static unsigned heads_grid(const mkf_batch* b, int sms)
{
    static const int forced = [] {
        const char* e = getenv("MKF_HEADS_CTAS_PER_SM");
        const int x = e ? atoi(e) : 0;
        return (x >= 1 && x <= 128) ? x : 0;
    }();
    if (forced) return (unsigned)(forced * sms);
    long long g = b->total / 1024;
    if (g < 12ll * sms) g = 12ll * sms;
    if (g > 128ll * sms) g = 128ll * sms;
    return (unsigned)g;
}

static bool heads_tma_enabled()
{
    static const bool on = [] {
        const char* e = getenv("MKF_HEADS_TMA");
        return !(e && e[0] == '0');
    }();
    return on;
}

// shape of k_slot_update_heads_tma, MKF_HEADS_TMA_CFG = consumer warps * 100 + output stages * 10 + producer warps
static int heads_tma_cfg()
{
    static const int v = [] {
        const char* e = getenv("MKF_HEADS_TMA_CFG");
        return e ? atoi(e) : 442;
    }();
    return v;
}

static int heads_block()
{
    static const int v = [] {
        const char* e = getenv("MKF_HEADS_BLOCK");
        const int x = e ? atoi(e) : 128;
        return (x == 64 || x == 128) ? x : 128;
    }();
    return v;
}

// Tracks whose records the heads kernel keeps in L2 (SlotArgs::l2_tracks).  A run-length track holds ~N/8 records of
// 720 bytes in each of the two ping-pong buffers; MKF_L2_RESIDENT_MB (default below) is the L2 budget those may take,
// MKF_L2_TRACKS sets the track count directly (0 switches the hints off).
static int l2_resident_tracks(const mkf_batch* b)
{
    static const long long forced = [] {
        const char* e = getenv("MKF_L2_TRACKS");
        return e ? atoll(e) : -1ll;
    }();
    static const long long budget_mb = [] {
        const char* e = getenv("MKF_L2_RESIDENT_MB");
        return e ? atoll(e) : 0ll;
    }();
    long long t;
    if (forced >= 0)
        t = forced;
    else {
        const long long per_track = 2ll * (b->N / 8 + b->m->K) * b->lay.np * 16;
        t = (budget_mb << 20) / (per_track > 0 ? per_track : 1);
    }
    if (t > b->T) t = b->T;
    return (int)t;
}

// one frame on a run-length particle set (mkf_runs.cuh): frame heads -> slot update of the heads -> repair -> resample
static int update_device_runs(mkf_batch* b, const double* d_meas, int meas_layout, const double* d_uind,
                              const double* d_upost, const uint64_t* d_seeds, int seed_stride, int seed_off,
                              cudaEvent_t* pe)
{
    const mkf_model* m = b->m;
    // which of the frame's four launches carry the programmatic-launch attribute (bit 0: k_frame_heads, 1: the slot
    // kernel, 2: k_runs_repair, 3: k_resample_runs); MKF_PDL_MASK overrides for A/B runs.  What matters is WHEN a kernel
    // releases its dependents: the slot kernel released at the start of k_frame_heads starts 1.0 us after it instead of
    // 3.5 but then lasts 62 us instead of 54 (device timeline of the pipelined loop at 4096 x 500, tools/tma_timeline.py:
    // period 100.7 us; 94.9 without the attribute on the slot kernel); released while k_frame_heads writes its work list it
    // starts as early and runs at its normal speed (93.9 us; profiles/r02_pdl_masks.txt)
    static const int pdl_mask = [] {
        const char* e = getenv("MKF_PDL_MASK");
        return e ? atoi(e) : 15;
    }();
    static const int heads_late = [] { // MKF_PDL_HEADS_LATE=0: k_frame_heads releases the slot kernel at its start
        const char* e = getenv("MKF_PDL_HEADS_LATE");
        return (e && e[0] == '0') ? 0 : 1;
    }();
    // MKF_FUSED=1: the single-launch frame kernel k_frame_fused instead of the three grid-wide kernels (k_frame_heads,
    // k_slot_update_heads_direct, k_resample_runs).  Measured at 4096 x 500: 0.147 ms per frame against 0.104 -- a warp
    // that owns its tracks walks their bookkeeping latency chains one after the other with only 8 warps per SM to hide
    // them -- so it is an experiment kept for A/B runs, not the default.
    const bool fused = [] {
        const char* e = getenv("MKF_FUSED");
        return e && e[0] == '1';
    }();
    // the heads' slot update with TMA-staged contiguous records (mkf_heads_tma.cuh) unless MKF_HEADS_TMA=0 or the model
    // constants leave no room for the stages (K > ~50)
    const int tma_cfg = heads_tma_cfg(); // consumer warps * 100 + output stages * 10 + producer warps
    size_t tma_smem = 0;
#define MKF_TMA_CASES(X) X(4, 4, 2) X(4, 4, 0) X(4, 4, 4) X(4, 4, 1) X(6, 3, 2)
#define MKF_TMA_SMEM(W_, O_, P_)                                                                                     \
    if (tma_cfg == W_ * 100 + O_ * 10 + P_)                                                                           \
        tma_smem = m->d == 12 ? HeadsTmaLay<12, W_, O_>::smem_bytes(m->K) : HeadsTmaLay<10, W_, O_>::smem_bytes(m->K);
    MKF_TMA_CASES(MKF_TMA_SMEM)
#undef MKF_TMA_SMEM
    const bool use_tma = !fused && heads_tma_enabled() && tma_smem && tma_smem <= 226 * 1024;
    {
        int rc0 = relayout(b, use_tma);
        if (rc0) return rc0;
    }
    if (!b->run_mode) { // entering from a per-slot set (reset, upload, a per-slot frame): its run list
        mkf_launch(k_runs_from_slots, grid_for(b->T, 4), 128, 0, b->stream, b->gather_index(), b->T, b->N, b->runs,
                   b->nruns);
        MKF_LAUNCHED();
        CK(cudaGetLastError());
        b->run_mode = true; // (slots_valid keeps its value: the per-slot arrays still describe this set)
    }
    if (pe) {
        cudaEventRecord(pe[0], b->stream);
        cudaEventRecord(pe[1], b->stream); // no separate indicator kernel: the draw is part of k_frame_heads
    }
    FrameArgs f{};
    f.u_ind = d_uind;
    f.T = b->T;
    f.N = b->N;
    f.K = m->K;
    f.cw_hi = b->d_cw_hi;
    f.cw_lo = b->d_cw_lo;
    f.wprior = b->d_wprior;
    f.wmax = m->prior_wmax;
    f.bounds = b->bounds;
    f.status = b->status;
    f.clear_status = b->clear_status_next ? 1 : 0;
    f.ind_tail = b->ind_tail;
    f.runs = b->runs;
    f.nruns = b->nruns;
    f.hmeta = b->hmeta;
    f.nheads = b->nheads;
    f.hd16 = b->hd16;
    f.head_count = b->head_count + b->head_flip;
    if (meas_layout == MKF_MEAS_CAND) {
        f.bin_cuts = b->cm_cuts;
        f.bins = b->cm_bins;
        f.cand_C = b->cm_C;
        f.hand = b->cm_hand;
    }
    b->clear_status_next = false;
    SlotArgs a{};
    fill_slot_args(b, a, d_meas, meas_layout);
    a.dedup = 1;
    a.split = 1;
    a.rep = nullptr;
    a.l2_tracks = l2_resident_tracks(b);
    a.aos = use_tma ? 1 : 0;
    a.st_in_alias = a.st_in;
    {
        static std::atomic<int> frame_no{0};
        f.dbg_frame = a.dbg_frame = frame_no.fetch_add(1, std::memory_order_relaxed);
    }
    static const bool no_xs = [] {
        const char* e = getenv("MKF_NO_XS");
        return e && e[0] == '1';
    }();
    if (use_tma && !b->xs && !no_xs) {
        const size_t n32 = ((size_t)b->total + 31) / 32;
        CK(cudaMalloc((void**)&b->xs, n32 * 32 * (size_t)(m->d / 2) * sizeof(double2)));
    }
    if (use_tma) { // records in list order: this frame's parents at lbase[lb_flip], its heads at lbase[lb_flip ^ 1]
        a.xs = b->xs;
        f.lbase_prev = a.lbase_prev = b->lbase + (size_t)b->lb_flip * b->T;
        f.lbase_cur = b->lbase + (size_t)(b->lb_flip ^ 1) * b->T;
        a.lbase_cur = f.lbase_cur;
    }
    if (pe && b->prof_ts) a.ts = b->prof_ts + 2 * (size_t)b->prof_n;
    ResampleRunsArgs ra{};
    ra.T = b->T;
    ra.N = b->N;
    ra.hmeta = b->hmeta;
    ra.nheads = b->nheads;
    ra.w_rec = b->w_rec;
    ra.u = d_upost;
    ra.seeds = d_seeds;
    ra.seed_stride = seed_stride;
    ra.seed_off = seed_off;
    ra.wsum_out = b->wsum;
    ra.status = b->status;
    ra.runs = b->runs;
    ra.nruns = b->nruns;
    ra.u_keep = b->u_keep;
    ra.seed_keep = b->seed_keep;
    // the estimate of the new set goes to the other slot (a host copy of the previous one may be in flight)
    b->est_slot ^= 1;
    if (b->est_used[b->est_slot]) CK(cudaStreamWaitEvent(b->stream, b->est_done[b->est_slot], 0));
    ra.st_new = b->st[b->cur ^ 1];
    ra.aos = use_tma ? 1 : 0;
    ra.dbg_frame = a.dbg_frame;
    ra.xs = use_tma ? b->xs : nullptr; // (null with MKF_NO_XS=1: the estimator then gathers from the records)
    ra.lbase = use_tma ? b->lbase + (size_t)(b->lb_flip ^ 1) * b->T : nullptr;
    ra.Dpose = m->D;
    ra.recon = b->d_recon;
    ra.pmean = b->d_pmean;
    ra.tinv = b->d_tinv;
    ra.est_xbar = b->est[b->est_slot];
    ra.est_pose = b->est[b->est_slot] + (size_t)b->T * m->d;
    ra.est_pose2 = b->pose_cache_on ? (double*)b->pose_cache.p : nullptr;
    const size_t smem = (size_t)m->K * b->lay.cs * sizeof(double);
    const size_t coef_bytes = (size_t)(m->D + m->d) * m->d * sizeof(double);
    {
        static std::atomic<uint64_t> seen{0};
        if (first_on_this_device(seen)) {
            CK(cudaFuncSetAttribute(k_slot_update_heads_direct<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            CK(cudaFuncSetAttribute(k_slot_update_heads_direct<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
#define MKF_TMA_ATTR(W_, O_, P_)                                                                                      \
    CK(cudaFuncSetAttribute(k_slot_update_heads_tma<12, W_, O_, P_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024)); \
    CK(cudaFuncSetAttribute(k_slot_update_heads_tma<10, W_, O_, P_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
            MKF_TMA_CASES(MKF_TMA_ATTR)
#undef MKF_TMA_ATTR
            CK(cudaFuncSetAttribute(k_frame_fused<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
            CK(cudaFuncSetAttribute(k_frame_fused<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
        }
    }
    if (fused) {
        // persistent grid: 2 CTAs of 4 warps per SM, but no more warps than tracks
        long long ctas = 2ll * sm_count(b->device);
        if (ctas * 4 > b->T) ctas = (b->T + 3) / 4;
        if (m->d == 12)
            mkf_launch(k_frame_fused<12>, (unsigned)ctas, 128, smem + coef_bytes, b->stream, f, a, ra);
        else
            mkf_launch(k_frame_fused<10>, (unsigned)ctas, 128, smem + coef_bytes, b->stream, f, a, ra);
        MKF_LAUNCHED();
        CK(cudaGetLastError());
        if (pe) {
            cudaEventRecord(pe[2], b->stream); // (one kernel: reported as the slot-kernel stage)
            cudaEventRecord(pe[3], b->stream);
        }
    } else {
        g_pdl_override = (pdl_mask & 1) ? -1 : 0;
        f.pdl_late = heads_late;
        mkf_launch(k_frame_heads, grid_for(b->T, MKF_FH_WARPS), 32 * MKF_FH_WARPS, 0, b->stream, f);
        g_pdl_override = (pdl_mask & 2) ? -1 : 0; // (the slot kernel)
        MKF_LAUNCHED();
        CK(cudaGetLastError());
        if (pe) cudaEventRecord(pe[2], b->stream);
        const unsigned hblock = (unsigned)heads_block();
        const unsigned hgrid = heads_grid(b, sm_count(b->device)) * (128 / hblock);
        {
            static const int late = [] {
                const char* e = getenv("MKF_PDL_SLOT_LATE");
                return (e && e[0] == '0') ? 0 : 1;
            }();
            a.pdl_late = late;
        }
        snprintf(b->heads_kernel, sizeof b->heads_kernel,
                 use_tma ? "k_slot_update_heads_tma<%d, %d, %d, %d>" : "k_slot_update_heads_direct<%d>", m->d,
                 tma_cfg / 100, tma_cfg / 10 % 10, tma_cfg % 10);
        if (use_tma) {
            const unsigned tgrid = (unsigned)sm_count(b->device);
            int* const clr = b->head_count + (b->head_flip ^ 1);
#define MKF_TMA_LAUNCH(W_, O_, P_)                                                                                   \
    if (tma_cfg == W_ * 100 + O_ * 10 + P_) {                                                                         \
        if (m->d == 12)                                                                                               \
            mkf_launch(k_slot_update_heads_tma<12, W_, O_, P_>, tgrid, 32 * (W_ + P_), tma_smem, b->stream, a, clr);  \
        else                                                                                                          \
            mkf_launch(k_slot_update_heads_tma<10, W_, O_, P_>, tgrid, 32 * (W_ + P_), tma_smem, b->stream, a, clr);  \
    }
            MKF_TMA_CASES(MKF_TMA_LAUNCH)
#undef MKF_TMA_LAUNCH
        } else if (m->d == 12)
            mkf_launch(k_slot_update_heads_direct<12>, hgrid, hblock, smem, b->stream, a, b->head_count + (b->head_flip ^ 1));
        else
            mkf_launch(k_slot_update_heads_direct<10>, hgrid, hblock, smem, b->stream, a, b->head_count + (b->head_flip ^ 1));
        b->head_flip ^= 1;
        MKF_LAUNCHED();
        CK(cudaGetLastError());
        if (pe) cudaEventRecord(pe[3], b->stream);
    }
    g_pdl_override = (pdl_mask & 4) ? -1 : 0;
    // flagged tracks (cv::Cholesky failure) are redone with the literal failure semantics; after the fused kernel, which
    // leaves them unresampled, their resample happens here too
    if (m->d == 12)
        mkf_launch(k_runs_repair<12>, grid_for(b->T, 128), 128, coef_bytes, b->stream, a, (const int4*)b->hmeta,
                   (const int*)b->nheads, ra, fused ? 1 : 0);
    else
        mkf_launch(k_runs_repair<10>, grid_for(b->T, 128), 128, coef_bytes, b->stream, a, (const int4*)b->hmeta,
                   (const int*)b->nheads, ra, fused ? 1 : 0);
    MKF_LAUNCHED();
    CK(cudaGetLastError());
    if (pe) cudaEventRecord(pe[4], b->stream);
    b->cur ^= 1;
    if (use_tma) b->lb_flip ^= 1;
    g_pdl_override = (pdl_mask & 8) ? -1 : 0;
    if (!fused) {
        if (m->d == 12)
            mkf_launch(k_resample_runs<12>, grid_for(b->T, 4), 128, coef_bytes, b->stream, ra);
        else
            mkf_launch(k_resample_runs<10>, grid_for(b->T, 4), 128, coef_bytes, b->stream, ra);
        MKF_LAUNCHED();
        CK(cudaGetLastError());
    }
    g_pdl_override = -1;
    if (pe) cudaEventRecord(pe[5], b->stream);
    b->shared = true;
    b->slots_valid = false;
    b->est_valid = true;
    if (b->pose_cache_on) b->pose_valid = true;
    return MKF_OK;
}

// the frame of a batch of short tracks: k_frame_small (indicator draw, slot update, resample, estimate) + the repair
// kernel for flagged tracks (mkf_frame_small.cuh)
static int update_device_small(mkf_batch* b, const double* d_meas, int meas_layout, const double* d_uind,
                               const double* d_upost, const uint64_t* d_seeds, int seed_stride, int seed_off,
                               cudaEvent_t* pe)
{
    const mkf_model* m = b->m;
    if (pe) { // (no indicator / keys kernel: two empty intervals)
        cudaEventRecord(pe[0], b->stream);
        cudaEventRecord(pe[1], b->stream);
        cudaEventRecord(pe[2], b->stream);
    }
    SlotArgs a{};
    fill_slot_args(b, a, d_meas, meas_layout);
    if (pe && b->prof_ts) a.ts = b->prof_ts + 2 * (size_t)b->prof_n;
    SmallTailArgs s{};
    s.u_ind = d_uind;
    s.cw_hi = b->d_cw_hi;
    s.cw_lo = b->d_cw_lo;
    s.wprior = b->d_wprior;
    s.wmax = m->prior_wmax;
    s.bounds_out = b->bounds;
    s.ind_tail_out = b->ind_tail;
    s.clear_status = b->clear_status_next ? 1 : 0;
    b->clear_status_next = false;
    s.u_post = d_upost;
    s.seeds = d_seeds;
    s.seed_stride = seed_stride;
    s.seed_off = seed_off;
    s.parent_out = b->parent;
    s.wsum = b->wsum;
    s.unsorted_out = b->unsorted;
    s.Dpose = m->D;
    s.recon = b->d_recon;
    s.pmean = b->d_pmean;
    s.tinv = b->d_tinv;
    // the estimate of the new set goes to the other slot (a host copy of the previous one may be in flight)
    b->est_slot ^= 1;
    if (b->est_used[b->est_slot]) CK(cudaStreamWaitEvent(b->stream, b->est_done[b->est_slot], 0));
    s.est_xbar = b->est[b->est_slot];
    s.est_pose = b->est[b->est_slot] + (size_t)b->T * m->d;
    s.est_pose2 = b->pose_cache_on ? (double*)b->pose_cache.p : nullptr;
    s.small_const = b->d_small_const;
    s.small_const_bytes = b->small_const_bytes;
    s.step = 1.0 / (double)b->N;
    const size_t smem = b->small_const_bytes;
    static std::atomic<uint64_t> seen{0};
    if (smem > 48 * 1024 && first_on_this_device(seen)) {
        CK(cudaFuncSetAttribute(k_frame_small<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        CK(cudaFuncSetAttribute(k_frame_small<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    }
    const unsigned g = grid_for(b->T, 8); // 4 warps x 2 tracks
    if (m->d == 12)
        mkf_launch(k_frame_small<12>, g, 128, smem, b->stream, a, s);
    else
        mkf_launch(k_frame_small<10>, g, 128, smem, b->stream, a, s);
    MKF_LAUNCHED();
    CK(cudaGetLastError());
    if (pe) cudaEventRecord(pe[3], b->stream);
    if (m->d == 12)
        mkf_launch(k_slot_update_repair<12, true>, grid_for(b->T, 128), 128, 0, b->stream, a, b->chain_last, s);
    else
        mkf_launch(k_slot_update_repair<10, true>, grid_for(b->T, 128), 128, 0, b->stream, a, b->chain_last, s);
    MKF_LAUNCHED();
    CK(cudaGetLastError());
    if (pe) {
        cudaEventRecord(pe[4], b->stream);
        cudaEventRecord(pe[5], b->stream);
    }
    b->cur ^= 1;
    b->shared = false;
    b->est_valid = true;
    if (b->pose_cache_on) b->pose_valid = true;
    return MKF_OK;
}

// the frame pipeline on device pointers
static int update_device(mkf_batch* b, const double* d_meas, int meas_layout, const double* d_uind,
                         const double* d_upost, int u_stride, const uint64_t* d_seeds, int seed_stride, int seed_off)
{
    const mkf_model* m = b->m;
    int rc;
    g_pdl_override = -1; // (an error return inside update_device_runs may have left it cleared)
    if (u_stride != 1) {
        mkf_set_error("internal: strided u_ind unsupported");
        return MKF_E_INVALID;
    }
    const bool prof = b->prof_on && (b->prof_tick++ % (uint64_t)b->prof_every) == 0 &&
                      (size_t)(b->prof_n + 1) * MKF_PROF_EV <= b->prof_ev.size();
    cudaEvent_t* pe = prof ? &b->prof_ev[(size_t)b->prof_n * MKF_PROF_EV] : nullptr;
    b->pose_valid = false;
    // identical children exist only when the slots of a track see one measurement (and never in the literal alias
    // mode, where duplicates are filtered one after the other)
    // Worth it when a component owns several slots of a track (N >= 4 K); at N = K = 15 nearly every slot is its own
    // (parent, component) pair and the bookkeeping costs more than it saves (measured: 3.65 -> 4.15 ms at 1 M x 15).
    // ... or one of a few candidate columns (MKF_MEAS_CAND): the candidate bin is then part of the key, which only the
    // two-launch path carries
    const bool dedup = b->dedup_ok && b->stage == 3 && b->N >= 4 * m->K &&
                       (meas_layout == MKF_MEAS_SHARED ||
                        (meas_layout == MKF_MEAS_CAND && b->share_split && b->N > 64));
    // (tracks of <= 64 slots go to k_resample_small, which reads per-slot weights)
    const bool use_split = dedup && b->share_split && b->N > 64;
    // run-length pipeline: one measurement per track, or a few candidate columns whose bins came as run boundaries
    if (use_split && b->runs &&
        (meas_layout == MKF_MEAS_SHARED || (meas_layout == MKF_MEAS_CAND && b->cm_cuts && b->cm_C <= 32))) {
        if (meas_layout == MKF_MEAS_CAND && (!b->cm_cand || !b->cm_bins || !b->cm_roi)) {
            mkf_set_error("internal: MKF_MEAS_CAND without a candidate source");
            return MKF_E_INVALID;
        }
        rc = update_device_runs(b, d_meas, meas_layout, d_uind, d_upost, d_seeds, seed_stride, seed_off, pe);
        if (prof && rc == MKF_OK) b->prof_n++;
        return rc;
    }
    if (b->run_mode) { // leaving the run-level representation: this frame reads per-slot record indices
        if ((rc = ensure_slots(b))) return rc;
        b->run_mode = false;
    }
    b->est_valid = false;
    if (b->small_fused && b->stage == 3 && !dedup && (meas_layout != MKF_MEAS_CAND || (b->cm_cand && b->cm_bins && b->cm_roi))) {
        rc = update_device_small(b, d_meas, meas_layout, d_uind, d_upost, d_seeds, seed_stride, seed_off, pe);
        if (prof && rc == MKF_OK) b->prof_n++;
        return rc;
    }
    if (prof) cudaEventRecord(pe[0], b->stream);
    if ((rc = launch_bounds_kernel(b, d_uind, b->clear_status_next ? 1 : 0))) return rc;
    b->clear_status_next = false;
    if (prof) cudaEventRecord(pe[1], b->stream);
    SlotArgs a{};
    fill_slot_args(b, a, d_meas, meas_layout);
    if (meas_layout == MKF_MEAS_CAND && (!a.cand || !a.bins || !a.roi)) {
        mkf_set_error("internal: MKF_MEAS_CAND without a candidate source");
        return MKF_E_INVALID;
    }
    a.dedup = dedup ? 1 : 0;
    a.rep = dedup ? b->rep : nullptr;
    if (prof && b->prof_ts) a.ts = b->prof_ts + 2 * (size_t)b->prof_n;
    const size_t smem = (size_t)m->K * b->lay.cs * sizeof(double);
    {
        static std::atomic<uint64_t> seen{0};
        if (smem > 48 * 1024 && first_on_this_device(seen)) {
#define MKF_SLOT_ATTR(DD, CH)                                                                                     \
    CK(cudaFuncSetAttribute(k_slot_update<DD, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024))
            MKF_SLOT_ATTR(12, false);
            MKF_SLOT_ATTR(12, true);
            MKF_SLOT_ATTR(10, false);
            MKF_SLOT_ATTR(10, true);
#undef MKF_SLOT_ATTR
            CK(cudaFuncSetAttribute(k_slot_update_chain_dyn<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            CK(cudaFuncSetAttribute(k_slot_update_chain_dyn<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        }
        const unsigned g = grid_for(b->total, 128);
        // slots per CTA of the record-sharing kernel = 128 * share_g (MKF_SHARE_G overrides for experiments)
        static const int share_g = [] {
            const char* e = getenv("MKF_SHARE_G");
            const int v = e ? atoi(e) : 8;
            return (v == 4 || v == 8 || v == 16) ? v : 8;
        }();
        a.split = use_split ? 1 : 0;
        const bool alias_dyn = a.alias_chain && b->alias_list && b->stage == 3;
        if (prof && !use_split && !alias_dyn) cudaEventRecord(pe[2], b->stream); // no keys kernel: an empty interval
        const size_t smem_shared = smem + (size_t)128 * share_g * (8 + 4 * 4 + 1);
        static std::atomic<uint64_t> seen_shared{0};
        if (dedup && first_on_this_device(seen_shared)) {
#define MKF_SHARED_ATTR(DD, GG)                                                                                        \
    CK(cudaFuncSetAttribute(k_slot_update_shared<DD, GG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024))
            MKF_SHARED_ATTR(12, 4);
            MKF_SHARED_ATTR(12, 8);
            MKF_SHARED_ATTR(12, 16);
            MKF_SHARED_ATTR(10, 4);
            MKF_SHARED_ATTR(10, 8);
            MKF_SHARED_ATTR(10, 16);
#undef MKF_SHARED_ATTR
            CK(cudaFuncSetAttribute(k_slot_update_heads_direct<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            CK(cudaFuncSetAttribute(k_slot_update_heads_direct<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        }
#define MKF_SHARED_LAUNCH(DD, GG)                                                                                      \
    mkf_launch(k_slot_update_shared<DD, GG>, grid_for(b->total, 128 * GG), 128, smem_shared, b->stream, a)
#define MKF_SLOT_LAUNCH(DD)                                                            \
    do {                                                                               \
        if (a.alias_chain && b->alias_list && b->stage == 3) {                         \
            int* cnt = b->alias_cnt + 2 * b->alias_flip;                               \
            mkf_launch(k_alias_runs, grid_for(b->total, 1024), 256, 0, b->stream, a.src, (const uint32_t*)b->unsorted, \
                       b->total, b->N, b->alias_list, cnt);                            \
            MKF_LAUNCHED();                                                            \
            if (prof) cudaEventRecord(pe[2], b->stream);                               \
            mkf_launch(k_slot_update_chain_dyn<DD>, (unsigned)(2 * sm_count(b->device)), 128, smem, b->stream, a, \
                       (const int*)b->alias_list, (const int*)cnt, cnt + 1, b->alias_cnt + 2 * (b->alias_flip ^ 1)); \
            b->alias_flip ^= 1;                                                        \
        } else if (a.alias_chain)                                                      \
            mkf_launch(k_slot_update<DD, true>, g, 128, smem, b->stream, a);    \
        else if (use_split) {                                                          \
            if (meas_layout == MKF_MEAS_CAND)                                                                          \
                mkf_launch(k_share_keys<true>, grid_for(b->total, MKF_SHARE_CHUNK), 256, 0, b->stream, a);            \
            else                                                                                                       \
                mkf_launch(k_share_keys<false>, grid_for(b->total, MKF_SHARE_CHUNK), 256, 0, b->stream, a);           \
            MKF_LAUNCHED();                                                            \
            if (prof) cudaEventRecord(pe[2], b->stream);                               \
            mkf_launch(k_slot_update_heads_direct<DD>, heads_grid(b, sm_count(b->device)), 128, smem, b->stream, a,   \
                       b->head_count + (b->head_flip ^ 1));                            \
            b->head_flip ^= 1;                                                         \
        } else if (dedup && share_g == 4)                                                \
            MKF_SHARED_LAUNCH(DD, 4);                                                  \
        else if (dedup && share_g == 8)                                                \
            MKF_SHARED_LAUNCH(DD, 8);                                                  \
        else if (dedup)                                                                \
            MKF_SHARED_LAUNCH(DD, 16);                                                 \
        else                                                                           \
            mkf_launch(k_slot_update<DD, false>, g, 128, smem, b->stream, a);   \
    } while (0)
        if (m->d == 12)
            MKF_SLOT_LAUNCH(12);
        else
            MKF_SLOT_LAUNCH(10);
#undef MKF_SLOT_LAUNCH
#undef MKF_SHARED_LAUNCH
    }
    MKF_LAUNCHED();
    CK(cudaGetLastError());
    if (prof) cudaEventRecord(pe[3], b->stream);
    // rare tracks (cv::Cholesky failure flagged, or unsorted parents in the literal alias mode) are redone
    if (m->d == 12)
        mkf_launch(k_slot_update_repair<12, false>, grid_for(b->T, 128), 128, 0, b->stream, a, b->chain_last, SmallTailArgs{});
    else
        mkf_launch(k_slot_update_repair<10, false>, grid_for(b->T, 128), 128, 0, b->stream, a, b->chain_last, SmallTailArgs{});
    MKF_LAUNCHED();
    CK(cudaGetLastError());
    if (prof) cudaEventRecord(pe[4], b->stream);
    b->cur ^= 1;
    rc = run_resample(b->stream, b->T, use_split ? b->w_rec : b->w_raw, b->N, b->N, d_upost, 1, 1, b->wsum, b->parent,
                      b->status, d_seeds, seed_stride, seed_off, MKF_ST_POST_FALLBACK, MKF_ST_POST_DEGENERATE,
                      b->unsorted, dedup ? b->rep : nullptr, dedup ? b->src : nullptr,
                      use_split ? b->w_raw : nullptr);
    b->shared = dedup;
    if (prof) {
        cudaEventRecord(pe[5], b->stream);
        b->prof_n++;
    }
    return rc;
}

// per-kernel device timing of mkf_batch_update with CUDA events on the batch's stream
extern "C" int mkf_batch_profile_every(mkf_batch* b, int max_samples, int every);
extern "C" int mkf_batch_profile(mkf_batch* b, int max_updates) { return mkf_batch_profile_every(b, max_updates, 1); }

extern "C" int mkf_batch_profile_every(mkf_batch* b, int max_updates, int every)
{
    if (!b || max_updates < 0 || every < 1) {
        mkf_set_error("mkf_batch_profile: invalid argument");
        return MKF_E_INVALID;
    }
    b->prof_every = every;
    b->prof_tick = 0;
    CK(cudaSetDevice(b->device));
    CK(cudaStreamSynchronize(b->stream));
    for (cudaEvent_t e : b->prof_ev) cudaEventDestroy(e);
    b->prof_ev.clear();
    b->prof_n = 0;
    b->prof_on = max_updates > 0;
    for (int i = 0; i < max_updates * MKF_PROF_EV; i++) {
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        b->prof_ev.push_back(e);
    }
    if (max_updates > b->prof_ts_cap) {
        if (b->prof_ts) cudaFree(b->prof_ts);
        b->prof_ts = nullptr;
        b->prof_ts_cap = 0;
        CK(cudaMalloc((void**)&b->prof_ts, (size_t)max_updates * 16));
        b->prof_ts_cap = max_updates;
    }
    if (max_updates > 0) { // start stamps take atomicMin, end stamps atomicMax
        std::vector<unsigned long long> init((size_t)max_updates * 2);
        for (int i = 0; i < max_updates; i++) {
            init[2 * i] = ~0ull;
            init[2 * i + 1] = 0ull;
        }
        CK(cudaMemcpy(b->prof_ts, init.data(), init.size() * 8, cudaMemcpyHostToDevice));
    }
    return MKF_OK;
}

extern "C" int mkf_batch_profile_read_stages(mkf_batch* b, double* ms /* 5 */, int* n_updates)
{
    if (!b) {
        mkf_set_error("null batch");
        return MKF_E_INVALID;
    }
    CK(cudaSetDevice(b->device));
    CK(cudaStreamSynchronize(b->stream));
    double acc[MKF_PROF_EV - 1] = {0, 0, 0, 0, 0};
    for (int i = 0; i < b->prof_n; i++) {
        for (int k = 0; k < MKF_PROF_EV - 1; k++) {
            float t = 0;
            CK(cudaEventElapsedTime(&t, b->prof_ev[(size_t)i * MKF_PROF_EV + k], b->prof_ev[(size_t)i * MKF_PROF_EV + k + 1]));
            acc[k] += t;
        }
    }
    if (ms)
        for (int k = 0; k < MKF_PROF_EV - 1; k++) ms[k] = acc[k];
    if (n_updates) *n_updates = b->prof_n;
    b->prof_n = 0;
    if (b->prof_ts && b->prof_ts_cap > 0) { // rearm the device-span stamps as well
        std::vector<unsigned long long> init((size_t)b->prof_ts_cap * 2);
        for (int i = 0; i < b->prof_ts_cap; i++) {
            init[2 * i] = ~0ull;
            init[2 * i + 1] = 0ull;
        }
        CK(cudaMemcpy(b->prof_ts, init.data(), init.size() * 8, cudaMemcpyHostToDevice));
    }
    return MKF_OK;
}

// the slot kernel's own span on the device (earliest CTA start to latest CTA end, %globaltimer) summed over the sampled
// updates; call BEFORE mkf_batch_profile_read_stages (which rearms the sample counter)
extern "C" int mkf_batch_profile_read_slot_span(mkf_batch* b, double* ms, int* n_updates)
{
    if (!b || !ms) {
        mkf_set_error("mkf_batch_profile_read_slot_span: null argument");
        return MKF_E_INVALID;
    }
    CK(cudaSetDevice(b->device));
    CK(cudaStreamSynchronize(b->stream));
    *ms = 0.0;
    int n = 0;
    if (b->prof_ts && b->prof_n > 0) {
        std::vector<unsigned long long> h((size_t)b->prof_n * 2);
        CK(cudaMemcpy(h.data(), b->prof_ts, h.size() * 8, cudaMemcpyDeviceToHost));
        for (int i = 0; i < b->prof_n; i++)
            if (h[2 * i] != ~0ull && h[2 * i + 1] > h[2 * i]) {
                *ms += (double)(h[2 * i + 1] - h[2 * i]) * 1e-6;
                n++;
            }
    }
    if (n_updates) *n_updates = n;
    return MKF_OK;
}

extern "C" int mkf_batch_profile_read(mkf_batch* b, double* ms_bounds, double* ms_slot_update, double* ms_resample,
                                      int* n_updates)
{
    double ms[MKF_PROF_EV - 1];
    const int rc = mkf_batch_profile_read_stages(b, ms, n_updates);
    if (rc) return rc;
    if (ms_bounds) *ms_bounds = ms[0];
    if (ms_slot_update) *ms_slot_update = ms[1] + ms[2] + ms[3]; // keys + slot kernel + repair
    if (ms_resample) *ms_resample = ms[4];
    return MKF_OK;
}

extern "C" int mkf_batch_update(mkf_batch* b, const double* meas, int meas_layout, const double* u_ind,
                                const double* u_post, const uint64_t* seeds, int mem)
{
    if (!b || !meas || !u_ind || !u_post) {
        mkf_set_error("mkf_batch_update: null argument");
        return MKF_E_INVALID;
    }
    if (meas_layout != MKF_MEAS_SHARED && meas_layout != MKF_MEAS_PER_SLOT) {
        mkf_set_error("mkf_batch_update: unknown measurement layout %d", meas_layout);
        return MKF_E_INVALID;
    }
    CK(cudaSetDevice(b->device));
    const size_t nmeas = (size_t)b->T * MKF_M * (meas_layout == MKF_MEAS_PER_SLOT ? (size_t)b->N : 1);
    const double *d_meas, *d_ui, *d_up;
    const uint64_t* d_seeds;
    int rc;
    if (mem == MKF_MEM_HOST_ASYNC) {
        // host inputs (ready at call time) travel on the copy stream into the staging slot of this frame
        AsyncIo& io = b->aio;
        if ((rc = io.init())) return rc;
        const int sl = io.in_slot;
        io.in_slot ^= 1;
        if (io.in_used[sl]) CK(cudaStreamWaitEvent(io.s_in, io.in_free[sl], 0));
        auto stage = [&](const void* src, size_t bytes, DevBuf& buf, const void** out) -> int {
            *out = nullptr;
            if (!src) return MKF_OK;
            int r = buf.ensure(bytes);
            if (r) return r;
            CK(cudaMemcpyAsync(buf.p, src, bytes, cudaMemcpyHostToDevice, io.s_in));
            *out = buf.p;
            return MKF_OK;
        };
        if ((rc = stage(meas, nmeas * 8, io.in_meas[sl], (const void**)&d_meas)) ||
            (rc = stage(u_ind, (size_t)b->T * 8, io.in_u0[sl], (const void**)&d_ui)) ||
            (rc = stage(u_post, (size_t)b->T * 8, io.in_u1[sl], (const void**)&d_up)) ||
            (rc = stage(seeds, (size_t)b->T * 16, io.in_seed[sl], (const void**)&d_seeds)))
            return rc;
        CK(cudaEventRecord(io.in_done[sl], io.s_in));
        CK(cudaStreamWaitEvent(b->stream, io.in_done[sl], 0));
        b->clear_status_next = true; // k_indicator_bounds zeroes the status words (no memset node in the chain)
        rc = update_device(b, d_meas, meas_layout, d_ui, d_up, 1, d_seeds, 2, 1);
        CK(cudaEventRecord(io.in_free[sl], b->stream));
        io.in_used[sl] = true;
        return rc;
    }
    if ((rc = in_ptr(b, meas, nmeas, mem, b->in_meas, &d_meas))) return rc;
    if ((rc = in_ptr(b, u_ind, (size_t)b->T, mem, b->in_u0, &d_ui))) return rc;
    if ((rc = in_ptr(b, u_post, (size_t)b->T, mem, b->in_u1, &d_up))) return rc;
    if ((rc = in_ptr(b, seeds, (size_t)b->T * 2, mem, b->in_seed, &d_seeds))) return rc;
    b->clear_status_next = true;
    return update_device(b, d_meas, meas_layout, d_ui, d_up, 1, d_seeds, 2, 1);
}

// getEstimator + reconstruction of every track of the batch (device pointers; either output may be null)
template <int DD>
static bool launch_estimate_d(mkf_batch* b, double* d_xbar, double* d_pose) // false: served by copies, no kernel
{
    const mkf_model* m = b->m;
    const double2* st = b->st[b->cur];
    double* d_pose2 = b->pose_cache_on ? (double*)b->pose_cache.p : nullptr;
    const size_t coef_bytes = (size_t)(m->D + DD) * DD * sizeof(double);
    if (b->est_valid) { // k_resample_runs / k_frame_small left the estimate of this set in the batch: copies only
        const double* ex = b->est[b->est_slot];
        const double* ep = ex + (size_t)b->T * DD;
        if (d_xbar) cudaMemcpyAsync(d_xbar, ex, (size_t)b->T * DD * 8, cudaMemcpyDeviceToDevice, b->stream);
        if (d_pose) cudaMemcpyAsync(d_pose, ep, (size_t)b->T * m->D * 8, cudaMemcpyDeviceToDevice, b->stream);
        if (d_pose2 && !b->pose_valid)
            cudaMemcpyAsync(d_pose2, ep, (size_t)b->T * m->D * 8, cudaMemcpyDeviceToDevice, b->stream);
        return false;
    }
    if (b->run_mode) { // the current set is a run list: sum of multiplicity x mean
        mkf_launch(k_estimate_runs<DD>, grid_for(b->T, 4), 128, coef_bytes, b->stream, st, (const int2*)b->runs,
                   (const int*)b->nruns, b->T, b->N, m->D, b->d_recon, b->d_pmean, b->d_tinv, d_xbar, d_pose, d_pose2);
        return true;
    }
    // tracks per CTA = TRIPS * 128 / GROUP: 8 trips amortise staging the reconstruction matrices once there are
    // enough tracks to fill the GPU several times over; small batches keep one trip so that they still spread out
#define MKF_EST_SMALL(G, TR)                                                                                  \
    mkf_launch(k_estimate_small<DD, G, TR>, grid_for(b->T, TR * 128 / G), 128, coef_bytes, b->stream,                 \
        st, b->gather_index(), b->T, b->N, m->D, b->d_recon, b->d_pmean, b->d_tinv, d_xbar, d_pose, d_pose2)
    if (b->N <= 16) {
        if (b->T >= 8 * 8 * 4 * 148)
            MKF_EST_SMALL(16, 8);
        else
            MKF_EST_SMALL(16, 1);
    } else if (b->N <= 96) {
        if (b->T >= 8 * 4 * 4 * 148)
            MKF_EST_SMALL(32, 8);
        else
            MKF_EST_SMALL(32, 1);
    }
#undef MKF_EST_SMALL
    else if (b->N <= 2048)
        mkf_launch(k_estimate<DD, 128>, (unsigned)b->T, 128, 0, b->stream, st, b->gather_index(), b->N, m->D, b->d_recon, b->d_pmean,
                                                                   b->d_tinv, d_xbar, d_pose, d_pose2);
    else // long tracks: more loads in flight per track (BT = 512 is slower at N = 500: 64 vs 41 us at 4096 tracks)
        mkf_launch(k_estimate<DD, 512>, (unsigned)b->T, 512, 0, b->stream, st, b->gather_index(), b->N, m->D, b->d_recon, b->d_pmean,
                                                                   b->d_tinv, d_xbar, d_pose, d_pose2);
    return true;
}
static int launch_estimate(mkf_batch* b, double* d_xbar, double* d_pose)
{
    if (b->aos && !b->est_valid) { // the estimate kernels read tiles
        int rc0 = relayout(b, false);
        if (rc0) return rc0;
    }
    const bool kernel = (b->m->d == 12) ? launch_estimate_d<12>(b, d_xbar, d_pose) : launch_estimate_d<10>(b, d_xbar, d_pose);
    if (kernel) MKF_LAUNCHED();
    CK(cudaGetLastError());
    if (b->pose_cache_on) b->pose_valid = true;
    return MKF_OK;
}

extern "C" int mkf_batch_estimate(mkf_batch* b, double* xbar, double* pose, int mem)
{
    if (!b) {
        mkf_set_error("null batch");
        return MKF_E_INVALID;
    }
    CK(cudaSetDevice(b->device));
    const mkf_model* m = b->m;
    OutPtr<double> ox, op;
    int rc;
    if (b->est_valid) {
        // run-length pipeline / short tracks: the frame already left xbar / pose of the current set in the batch -- copies only
        const int sl = b->est_slot;
        const double* ex = b->est[sl];
        const double* ep = ex + (size_t)b->T * m->d;
        const size_t nx = (size_t)b->T * m->d * 8, np = (size_t)b->T * m->D * 8;
        if (mem == MKF_MEM_HOST_ASYNC) {
            AsyncIo& io = b->aio;
            if ((rc = io.init())) return rc;
            CK(cudaEventRecord(b->est_ready, b->stream));
            CK(cudaStreamWaitEvent(io.s_out, b->est_ready, 0));
            if (xbar) CK(cudaMemcpyAsync(xbar, ex, nx, cudaMemcpyDeviceToHost, io.s_out));
            if (pose) CK(cudaMemcpyAsync(pose, ep, np, cudaMemcpyDeviceToHost, io.s_out));
            CK(cudaEventRecord(b->est_done[sl], io.s_out));
            b->est_used[sl] = true;
            return MKF_OK;
        }
        bool any_host = false;
        auto put = [&](double* dst, const double* src, size_t bytes) -> int {
            if (!dst) return MKF_OK;
            const bool dev = is_device_ptr(dst, mem);
            any_host |= !dev;
            CK(cudaMemcpyAsync(dst, src, bytes, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, b->stream));
            return MKF_OK;
        };
        if ((rc = put(xbar, ex, nx)) || (rc = put(pose, ep, np))) return rc;
        if (b->pose_cache_on && !b->pose_valid) {
            CK(cudaMemcpyAsync(b->pose_cache.p, ep, np, cudaMemcpyDeviceToDevice, b->stream));
            b->pose_valid = true;
        }
        if (any_host) CK(cudaStreamSynchronize(b->stream));
        return MKF_OK;
    }
    if (mem == MKF_MEM_HOST_ASYNC) {
        // the kernel writes this call's staging slot; the copy to the host runs on the output stream
        AsyncIo& io = b->aio;
        if ((rc = io.init())) return rc;
        const int sl = io.out_slot;
        io.out_slot ^= 1;
        const size_t nx = (size_t)b->T * m->d * 8, np = (size_t)b->T * m->D * 8;
        if (xbar && (rc = io.out_a[sl].ensure(nx))) return rc;
        if (pose && (rc = io.out_b[sl].ensure(np))) return rc;
        if (io.out_used[sl]) CK(cudaStreamWaitEvent(b->stream, io.out_done[sl], 0));
        if ((rc = launch_estimate(b, xbar ? (double*)io.out_a[sl].p : nullptr, pose ? (double*)io.out_b[sl].p : nullptr)))
            return rc;
        CK(cudaEventRecord(io.out_ready[sl], b->stream));
        CK(cudaStreamWaitEvent(io.s_out, io.out_ready[sl], 0));
        if (xbar) CK(cudaMemcpyAsync(xbar, io.out_a[sl].p, nx, cudaMemcpyDeviceToHost, io.s_out));
        if (pose) CK(cudaMemcpyAsync(pose, io.out_b[sl].p, np, cudaMemcpyDeviceToHost, io.s_out));
        CK(cudaEventRecord(io.out_done[sl], io.s_out));
        io.out_used[sl] = true;
        return MKF_OK;
    }
    if ((rc = ox.init(b, xbar, (size_t)b->T * m->d, mem, b->out_a))) return rc;
    if ((rc = op.init(b, pose, (size_t)b->T * m->D, mem, b->out_b))) return rc;
    if ((rc = launch_estimate(b, ox.devp, op.devp))) return rc;
    if ((rc = ox.finish(b)) || (rc = op.finish(b))) return rc;
    if ((ox.host || op.host) && mem != MKF_MEM_HOST_ASYNC) CK(cudaStreamSynchronize(b->stream));
    return MKF_OK;
}

extern "C" int mkf_batch_download(mkf_batch* b, double* x, double* P, double* w_raw, double* w_norm,
                                  int32_t* indicators, int32_t* parents, double* wsum, uint32_t* status, int mem)
{
    if (!b) {
        mkf_set_error("null batch");
        return MKF_E_INVALID;
    }
    CK(cudaSetDevice(b->device));
    const mkf_model* m = b->m;
    const int d = m->d;
    const size_t tot = (size_t)b->total;
    if (x || P || w_raw || w_norm || parents) { // a run-level particle set: replay the last resample per slot first
        int rc0 = ensure_slots(b);
        if (rc0) return rc0;
    }
    DevBuf sx, sp, sw, si; // temporaries for host-bound outputs (freed on return)
    OutPtr<double> ox, op, own;
    OutPtr<int32_t> oi;
    int rc = MKF_OK;
    auto done = [&](int code) {
        sx.release();
        sp.release();
        sw.release();
        si.release();
        return code;
    };
    if ((rc = ox.init(b, x, tot * d, mem, sx)) || (rc = op.init(b, P, tot * d * d, mem, sp)) ||
        (rc = own.init(b, w_norm, tot, mem, sw)) || (rc = oi.init(b, indicators, tot, mem, si)))
        return done(rc);
    if (x || P) {
        if (d == 12)
            k_download<12><<<grid_for(b->total, 64), 64, 0, b->stream>>>(b->st[b->cur], b->gather_index(), b->d_tinv, ox.devp,
                                                                          op.devp, b->total, b->N);
        else
            k_download<10><<<grid_for(b->total, 64), 64, 0, b->stream>>>(b->st[b->cur], b->gather_index(), b->d_tinv, ox.devp,
                                                                          op.devp, b->total, b->N);
        MKF_LAUNCHED();
        if (cudaGetLastError() != cudaSuccess) {
            mkf_set_error("k_download launch failed");
            return done(MKF_E_CUDA);
        }
    }
    if (w_norm || indicators) {
        k_aux_outputs<<<grid_for(b->total, 256), 256, 0, b->stream>>>(b->w_raw, b->wsum, b->bounds, b->total, b->N,
                                                                       m->K, own.devp, oi.devp, b->ind_tail);
        MKF_LAUNCHED();
        if (cudaGetLastError() != cudaSuccess) {
            mkf_set_error("k_aux_outputs launch failed");
            return done(MKF_E_CUDA);
        }
    }
    if ((rc = ox.finish(b)) || (rc = op.finish(b)) || (rc = own.finish(b)) || (rc = oi.finish(b))) return done(rc);
    auto copy_out = [&](void* dst, const void* src, size_t bytes) -> int {
        if (!dst) return MKF_OK;
        cudaMemcpyKind kind = is_device_ptr(dst, mem) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
        if (cudaMemcpyAsync(dst, src, bytes, kind, b->stream) != cudaSuccess) {
            mkf_set_error("cudaMemcpyAsync failed in mkf_batch_download");
            return MKF_E_CUDA;
        }
        return MKF_OK;
    };
    if ((rc = copy_out(w_raw, b->w_raw, tot * sizeof(double))) ||
        (rc = copy_out(parents, b->parent, tot * sizeof(int32_t))) ||
        (rc = copy_out(wsum, b->wsum, (size_t)b->T * sizeof(double))) ||
        (rc = copy_out(status, b->status, (size_t)b->T * sizeof(uint32_t))))
        return done(rc);
    if (cudaStreamSynchronize(b->stream) != cudaSuccess) {
        mkf_set_error("cudaStreamSynchronize failed: %s", cudaGetErrorString(cudaGetLastError()));
        return done(MKF_E_CUDA);
    }
    return done(MKF_OK);
}

extern "C" int mkf_batch_upload(mkf_batch* b, const double* x, const double* P, int mem)
{
    if (!b || !x || !P) {
        mkf_set_error("mkf_batch_upload: null argument");
        return MKF_E_INVALID;
    }
    CK(cudaSetDevice(b->device));
    const int d = b->m->d;
    const double *dx, *dP;
    int rc;
    if ((rc = in_ptr(b, x, (size_t)b->total * d, mem, b->in_x, &dx))) return rc;
    if ((rc = in_ptr(b, P, (size_t)b->total * d * d, mem, b->in_p, &dP))) return rc;
    b->cur = 0;
    b->aos = false;
    b->shared = false;
    b->run_mode = false;
    b->slots_valid = true;
    b->est_valid = false;
    b->pose_valid = false;
    CK(cudaMemsetAsync(b->unsorted, 0, (size_t)b->T * sizeof(uint32_t), b->stream));
    if (d == 12)
        k_upload<12><<<grid_for(b->total, 64), 64, 0, b->stream>>>(b->st[0], b->parent, dx, dP, b->d_tm, b->total, b->N);
    else
        k_upload<10><<<grid_for(b->total, 64), 64, 0, b->stream>>>(b->st[0], b->parent, dx, dP, b->d_tm, b->total, b->N);
    MKF_LAUNCHED();
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(b->stream));
    b->in_x.release();
    b->in_p.release();
    return MKF_OK;
}

// single weight vector, through the same kernels (one "track")
extern "C" int mkf_resample(const double* w, int L, int N, double u, uint64_t seed, int32_t* out, int device)
{
    if (!w || !out || L <= 0 || N <= 0) {
        mkf_set_error("mkf_resample: invalid argument");
        return MKF_E_INVALID;
    }
    int ndev = mkf_device_count();
    if (ndev <= 0) {
        mkf_set_error("no CUDA device available: libmkf_b200 has no CPU fallback");
        return MKF_E_CUDA;
    }
    if (device < 0 || device >= ndev) {
        mkf_set_error("mkf_resample: bad device");
        return MKF_E_INVALID;
    }
    CK(cudaSetDevice(device));
    double uu = u;
    if (u < 0.0) { // the reference's own draw order: one discarded int, then uniform(0.0, 1.0)
        uint64_t st = seed ? seed : 0xffffffffull;
        auto next = [&]() {
            st = (uint64_t)(unsigned)st * 4164903690u + (unsigned)(st >> 32);
            return (unsigned)st;
        };
        (void)next();
        unsigned t = next();
        uu = (double)(((uint64_t)t << 32) | next()) * 5.4210108624275221700372640043497e-20;
    }
    // one grow-only device scratch per device (the call is synchronous; a mutex serialises concurrent callers):
    //   [ w : L f64 ][ u, wsum : f64 ][ seed : u64 ][ status : u32, pad ][ out : N i32 ]
    // one packed upload of everything up to `out`, one download of status..out
    static std::mutex mtx;
    static void* scratch[64] = {nullptr};
    static size_t scratch_cap[64] = {0};
    std::lock_guard<std::mutex> lock(mtx);
    const size_t off_hdr = (size_t)L * 8, off_st = off_hdr + 24, off_out = off_st + 8, total = off_out + (size_t)N * 4;
    if (device >= 64) {
        mkf_set_error("mkf_resample: bad device");
        return MKF_E_INVALID;
    }
    if (scratch_cap[device] < total) {
        if (scratch[device]) cudaFree(scratch[device]);
        scratch[device] = nullptr;
        scratch_cap[device] = 0;
        if (cudaMalloc(&scratch[device], total + total / 2) != cudaSuccess) {
            mkf_set_error("mkf_resample: cudaMalloc failed");
            return MKF_E_NOMEM;
        }
        scratch_cap[device] = total + total / 2;
    }
    char* base = (char*)scratch[device];
    std::vector<char> host(off_out > (size_t)N * 4 + 8 ? off_out : (size_t)N * 4 + 8, 0);
    std::memcpy(host.data(), w, (size_t)L * 8);
    std::memcpy(host.data() + off_hdr, &uu, 8);
    std::memcpy(host.data() + off_hdr + 16, &seed, 8);
    CK(cudaMemcpy(base, host.data(), off_out, cudaMemcpyHostToDevice));
    const double* d_w = (const double*)base;
    const double* d_u = (const double*)(base + off_hdr);
    double* d_ws = (double*)(base + off_hdr + 8);
    const uint64_t* d_seed = (const uint64_t*)(base + off_hdr + 16);
    uint32_t* d_st = (uint32_t*)(base + off_st);
    int32_t* d_out = (int32_t*)(base + off_out);
    // the reference applies resample() to already-normalised weights: no division here
    int rc = run_resample(0, 1, d_w, L, N, d_u, 1, 0, d_ws, d_out, d_st, d_seed, 1, 0, MKF_ST_POST_FALLBACK,
                          MKF_ST_POST_DEGENERATE);
    if (rc) return rc;
    CK(cudaMemcpy(host.data(), base + off_st, 8 + (size_t)N * 4, cudaMemcpyDeviceToHost));
    std::memcpy(out, host.data() + 8, (size_t)N * 4);
    uint32_t st = 0;
    std::memcpy(&st, host.data(), 4);
    return (st & MKF_ST_POST_DEGENERATE) ? 1 : 0; // like orc_resample: 1 = degenerate fallback taken
}

extern "C" int mkf_synth_fill(mkf_batch* b, uint64_t seed, int64_t track0, uint64_t frame, int jitter,
                              int meas_layout, double* meas_dev, double* u_ind_dev, double* u_post_dev)
{
    if (!b || !meas_dev) {
        mkf_set_error("mkf_synth_fill: null argument");
        return MKF_E_INVALID;
    }
    CK(cudaSetDevice(b->device));
    const long long n = meas_layout == MKF_MEAS_SHARED ? b->T : b->total;
    k_synth_fill<<<grid_for(n, 256), 256, 0, b->stream>>>(seed, track0, frame, jitter, meas_layout, b->T, b->N,
                                                          meas_dev, u_ind_dev, u_post_dev);
    MKF_LAUNCHED();
    CK(cudaGetLastError());
    return MKF_OK;
}

#include "mkf_assoc.cuh"
#include "mkf_extra.cuh"
#include "mkf_pf2d.cuh"
#include "mkf_comm.cuh"

// debugging aid of MKF_TIMELINE builds: {start, end} of the four frame kernels for the last 64 frames (reset arms it)
extern "C" int mkf_debug_timeline(unsigned long long* out512, int reset)
{
#ifdef MKF_TIMELINE
    if (out512 && cudaMemcpyFromSymbol(out512, g_timeline, 64 * 4 * 2 * 8) != cudaSuccess) return MKF_E_CUDA;
    if (reset) {
        std::vector<unsigned long long> z(64 * 4 * 2);
        for (size_t i = 0; i < z.size(); i += 2) {
            z[i] = ~0ull;
            z[i + 1] = 0;
        }
        if (cudaMemcpyToSymbol(g_timeline, z.data(), z.size() * 8) != cudaSuccess) return MKF_E_CUDA;
    }
    return MKF_OK;
#else
    (void)out512;
    (void)reset;
    return MKF_E_INVALID;
#endif
}

// debugging aid of MKF_TMA_PROF builds (not part of include/mkf_b200.h): per-phase cycle counters of k_slot_update_heads_tma
extern "C" int mkf_debug_tma_prof(unsigned long long* out8, int reset)
{
#ifdef MKF_TMA_PROF
    if (out8 && cudaMemcpyFromSymbol(out8, g_tma_prof, 64) != cudaSuccess) return MKF_E_CUDA;
    if (reset) {
        unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (cudaMemcpyToSymbol(g_tma_prof, z, 64) != cudaSuccess) return MKF_E_CUDA;
    }
    return MKF_OK;
#else
    (void)out8;
    (void)reset;
    return MKF_E_INVALID;
#endif
}
