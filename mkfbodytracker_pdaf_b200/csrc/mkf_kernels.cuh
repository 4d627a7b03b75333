// mkf_kernels.cuh -- sm_100a kernels of the RBPF hot path (included by mkf_api.cu).
//
//   k_indicator_bounds  K->N systematic resample of the GMM prior weights     src/pf2DRao.cpp:128
//   k_slot_update       fused gather-by-parent + KF predict + innovation      src/pf2DRao.cpp:134-142
//                       likelihood + KF update, one thread per slot           src/KF_model.cpp:11-25
//                                                                              src/pf2DRao.cpp:34-67
//   k_resample_block    weight sum/normalise + N->N (or C->N) resample        src/pf2DRao.cpp:145-152,175-210
//   k_resample_small    same, one thread per track, literal loop (small N)
//                       (flagged tracks: literal loop / cv::RNG branch in the same kernels, src/pf2DRao.cpp:184-207)
//   k_estimate          getEstimator + PCA reconstruction                      src/pf2DRao.cpp:23-31, src/pfPose.cpp:347-348
//   k_reset / k_upload / k_download / k_aux_outputs : state I/O in reference coordinates
//
// Data layout: slot state lives in "tiles" of 32 consecutive global slots,
// tile[pair p][lane] as double2, so a warp reading pair p of 32 neighbouring slots touches
// 512 contiguous bytes.  Global slot index s = track * N + j.
#ifndef MKF_KERNELS_CUH
#define MKF_KERNELS_CUH

#include <float.h>

#include <type_traits>

#include "mkf_device.cuh"
#include "../../include/mkf_synth.h"

#define MKF_M 6
#ifndef MKF_CW
#define MKF_CW 2 // doubles per storage chunk (2: 16-byte chunks, 4: 32-byte chunks); see SlotLay
#endif
static_assert(MKF_CW == 2 || MKF_CW == 4 || MKF_CW == 8, "chunk width");

template <int D>
struct SlotLay {
    static constexpr int M = MKF_M;
    static constexpr int D2 = D - M;
    static constexpr int NA = M * (M + 1) / 2;
    static constexpr int NB = D2 * M;
    static constexpr int NC = D2 * (D2 + 1) / 2;
    static constexpr int NE = D + NA + NB + NC;
    static constexpr int NP = (NE + 1) / 2;
    // a lane's record is stored in chunks of H consecutive double2 (MKF_CW = 2 H doubles): pair p sits in chunk p / H.
    // H = 1 is the plain tile[pair][lane] arrangement (16-byte chunks); H = 2 gives 32-byte chunks, so a gathered
    // record uses whole 32-byte sectors even when the neighbouring lane's record is not wanted
    static constexpr int H = MKF_CW / 2;
    static constexpr int NCH = (NP + H - 1) / H;   // chunks per record (the last one may hold padding pairs)
    static constexpr int TILE2 = NCH * H * 32;     // double2 per tile of 32 records
    __host__ __device__ static constexpr int po(int p) { return (p / H) * (32 * H) + (p % H); } // double2 offset of pair p
    static constexpr int OA = D;
    static constexpr int OB = D + NA;
    static constexpr int OC = D + NA + NB;
    static constexpr int CS = (2 + NE + 1) / 2 * 2;
};

__host__ __device__ __forceinline__ constexpr int tri(int i, int j) { return i * (i + 1) / 2 + j; } // i >= j

struct SlotArgs {
    const double2* __restrict__ st_in;
    double2* __restrict__ st_out;
    const int32_t* __restrict__ parent; // T*N, local index of the parent slot
    const int32_t* __restrict__ src;    // T*N, local index of the RECORD in st_in that holds the parent's state
                                        // (== parent unless the previous frame stored identical children once)
    int32_t* __restrict__ rep;          // T*N out (dedup only): local index of the record in st_out holding this slot
    int dedup;                          // 1: identical children (same parent record, component, track measurement)
                                        //    are computed and stored once per warp
    // three-kernel record sharing (k_share_keys -> k_slot_update_heads_flat -> k_share_expand): the batch-wide list of
    // heads (16 bytes each: source record, destination record, track, component), its length (appended to with
    // atomicAdd), and the weight of every record at the record's position
    int4* __restrict__ hd16;
    int* __restrict__ head_count;
    double* __restrict__ w_rec;
    int split; // 1: this frame runs k_share_keys -> k_slot_update_heads_direct (weights per record in w_rec)
    const int32_t* __restrict__ bounds; // T x (K+2): e_0..e_{K-1}, wrap_from, wrap_k (-1: read ind_tail)
    const uint8_t* __restrict__ ind_tail; // T x N: component of the slots at or after wrap_from (written by the
                                          // literal loop of k_indicator_bounds when the draw wrapped past component K-1)
    const double* __restrict__ meas;
    const double* __restrict__ comp_const; // K x CS (global; staged to shared memory by TMA)
    double* __restrict__ w_raw;
    uint32_t* __restrict__ status;
    long long total; // T*N
    int N, K, meas_layout, chol_mode;
    int stage; // bit 0: KF_model::predict, bit 1: likelihood + KF_model::update (3 = the fused frame step)
    int alias_chain;                      // 1 = MKF_ALIAS_CV_SHALLOW_LITERAL
    const uint32_t* __restrict__ unsorted; // per track: parents are not sorted (random-index fallback ran)
    double bh[MKF_M];
    double r; // measurement noise variance (R = r * I)
    // MKF_MEAS_CAND (internal, mkf_batch_associate): a slot's column is assembled on the fly from the person's face
    // ROI and the candidate its bin selects -- [roi.x + w/2, roi.y + h/2, cand_x(bin), cand_y(bin), roi.x + w/2,
    // roi.y + neck h] (src/pfPose.cpp:303-323) -- instead of being materialised as T x 6 x N doubles
    const double* __restrict__ cand;   // T x 2 hands x 2 (x row, y row) x cand_C
    const int32_t* __restrict__ bins;  // T x 2 hands x N
    const double* __restrict__ roi;    // T x 4
    double neck;
    int cand_C, hand;
    // profiling only (null otherwise): {earliest CTA start, latest CTA end} of the slot kernel in %globaltimer
    // nanoseconds -- the kernel's own span on the device, free of the launch gap an event-bracketed kernel pays
    unsigned long long* ts;
    // L2-resident tracks (run-length pipeline): the records of tracks [0, l2_tracks) are read and written with the
    // L2 evict_last policy -- frame f+1 overwrites the lines frame f-1 wrote (ping-pong buffers, same addresses), so in
    // steady state those tracks cause no DRAM traffic at all -- every other track's with evict_first (read once, written
    // once per frame: nothing of theirs is worth a line).  0: no hints (plain __ldg / __stcs).
    int l2_tracks;
    // run-length pipeline with contiguous records (mkf_heads_tma.cuh): st_in / st_out hold record r at r * NP double2
    // instead of the tile layout; st_in_alias == st_in (see k_slot_update_heads_tma)
    int aos;
    const double2* __restrict__ st_in_alias;
    const int* __restrict__ lbase_prev; // (aos) list-order records: the parents of track t start at lbase_prev[t] ...
    const int* __restrict__ lbase_cur;  // ... its heads' records at lbase_cur[t]
    double2* __restrict__ xs;           // (aos) the heads' means again, tiled by list position (ResampleRunsArgs::xs)
    int dbg_frame;
    // k_slot_update_heads_tma: 1 = let the dependent kernels' CTAs be scheduled when this kernel's warps END instead of
    // at its start (see the kernel)
    int pdl_late;
};

// record `sp` of a state buffer in either layout: pair p lives at mkf_rec<D>(st, sp, aos)[mkf_rec_off<D>(p, aos)]
template <int D>
__device__ __forceinline__ long long mkf_rec_base(long long sp, int aos)
{
    using L = SlotLay<D>;
    return aos ? sp * L::NP : (sp >> 5) * (long long)L::TILE2 + (sp & 31) * L::H;
}
template <int D>
__device__ __forceinline__ int mkf_rec_off(int p, int aos)
{
    return aos ? p : SlotLay<D>::po(p);
}
#define MKF_MEAS_CAND 2

// MKF_TIMELINE builds: first-CTA-start / last-thread-end of the four kernels of a run-length frame in %globaltimer ns,
// for the last 64 frames (mkf_debug_timeline); dbg_frame travels in the kernels' argument blocks
#ifdef MKF_TIMELINE
__device__ unsigned long long g_timeline[64][4][2];
// (MKF_TL_LATEST_START: the slot kernel's entry holds the complement of its LATEST CTA start instead of the earliest)
#ifdef MKF_TL_LATEST_START
#define MKF_TL_START(k, fr)                                                                                            \
    do {                                                                                                               \
        if (threadIdx.x == 0)                                                                                          \
            atomicMin(&g_timeline[(fr) & 63][k][0], (k) == 1 ? ~mkf_globaltimer() : mkf_globaltimer());               \
    } while (0)
#else
#define MKF_TL_START(k, fr)                                                                                            \
    do {                                                                                                               \
        if (threadIdx.x == 0) atomicMin(&g_timeline[(fr) & 63][k][0], mkf_globaltimer());                             \
    } while (0)
#endif
#define MKF_TL_END(k, fr)                                                                                              \
    do {                                                                                                               \
        if ((threadIdx.x & 31) == 0) atomicMax(&g_timeline[(fr) & 63][k][1], mkf_globaltimer());                      \
    } while (0)
#else
#define MKF_TL_START(k, fr)
#define MKF_TL_END(k, fr)
#endif

// L2 cache-policy descriptors (createpolicy) and 16-byte accesses that carry one
// The policy operand travels in a uniform register: it has to be a constant the compiler can see (the value
// `createpolicy.fractional.L2::evict_last / evict_first ..., 1.0` produces; routed through a vector register by an asm
// output it costs two R2UR per access).
#define MKF_L2_EVICT_LAST 0x14F0000000000000ull
#define MKF_L2_EVICT_FIRST 0x12F0000000000000ull
template <bool KEEP>
__device__ __forceinline__ constexpr uint64_t mkf_l2_policy()
{
    return KEEP ? MKF_L2_EVICT_LAST : MKF_L2_EVICT_FIRST;
}
__device__ __forceinline__ double2 mkf_ldg_policy(const double2* p, uint64_t pol)
{
    double2 v;
    asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void mkf_stg_policy(double2* p, double2 v, uint64_t pol)
{
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
}

__device__ __forceinline__ unsigned long long mkf_globaltimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
} // internal third value of meas_layout (the public ones: include/mkf_b200.h)

// -----------------------------------------------------------------------------------------
// cv::Cholesky failure branch (a pivot < DBL_EPSILON, src/pf2DRao.cpp:37,52), literal:
//   * chol() hands back the PARTIALLY factored clone (rows above the failing one factored with 1/L_ii on
//     their diagonal, the failing row's off-diagonals replaced, everything else untouched) and mvnpdf uses
//     it as R: v = y^T R^-1 through cv::invert(DECOMP_LU), log of its diagonal (possibly NaN);
//   * KF_model::update inverts S by LU regardless (src/KF_model.cpp:21).
// Rare, so it lives out of line on local-memory arrays.  W = S^-1 (lower packed) feeds the common update.
// -----------------------------------------------------------------------------------------
__device__ __noinline__ bool mkf_lu_invert6(double* A /* 36, destroyed */, double* b /* 36: out */)
{
    const int m = 6;
    for (int i = 0; i < 36; i++) b[i] = ((i / 6) == (i % 6)) ? 1.0 : 0.0;
    for (int i = 0; i < m; i++) {
        int k = i;
        for (int j = i + 1; j < m; j++)
            if (fabs(A[j * m + i]) > fabs(A[k * m + i])) k = j;
        if (fabs(A[k * m + i]) < DBL_EPSILON) {
            for (int q = 0; q < 36; q++) b[q] = 0.0; // cv::invert: dst = Scalar(0)
            return false;
        }
        if (k != i) {
            for (int j = i; j < m; j++) {
                double tmp = A[i * m + j];
                A[i * m + j] = A[k * m + j];
                A[k * m + j] = tmp;
            }
            for (int j = 0; j < m; j++) {
                double tmp = b[i * m + j];
                b[i * m + j] = b[k * m + j];
                b[k * m + j] = tmp;
            }
        }
        const double d = -1.0 / A[i * m + i];
        for (int j = i + 1; j < m; j++) {
            const double alpha = A[j * m + i] * d;
            for (k = i + 1; k < m; k++) A[j * m + k] = fma(alpha, A[i * m + k], A[j * m + k]);
            for (k = 0; k < m; k++) b[j * m + k] = fma(alpha, b[i * m + k], b[j * m + k]);
        }
        A[i * m + i] = -d;
    }
    for (int i = m - 1; i >= 0; i--)
        for (int j = 0; j < m; j++) {
            double sacc = b[i * m + j];
            for (int k = i + 1; k < m; k++) sacc = fma(-A[i * m + k], b[k * m + j], sacc);
            b[i * m + j] = sacc * A[i * m + i];
        }
    return true;
}

__device__ __noinline__ void mkf_chol_fail_path(const double* Sfull /* 36 */, const double* y /* 6 */, int chol_mode,
                                                double* w_out, double* Wpacked /* 21 */)
{
    const int m = 6;
    double R[36], Ri[36], T[36];
    for (int i = 0; i < 36; i++) R[i] = Sfull[i];
    if (chol_mode != MKF_CHOL_EXACT) { // CholImpl in place until the failing pivot
        for (int i = 0; i < m; i++) {
            int j;
            double sacc;
            for (j = 0; j < i; j++) {
                sacc = R[i * m + j];
                for (int k = 0; k < j; k++) sacc -= R[i * m + k] * R[j * m + k];
                R[i * m + j] = sacc * R[j * m + j];
            }
            sacc = R[i * m + i];
            for (int k = 0; k < j; k++) sacc -= R[i * m + k] * R[i * m + k];
            if (sacc < DBL_EPSILON) break;
            R[i * m + i] = 1.0 / sqrt(sacc);
        }
    }
    for (int i = 0; i < 36; i++) T[i] = R[i];
    mkf_lu_invert6(T, Ri);
    double q0 = 0.0, q1 = 0.0, lsd = 0.0;
    for (int j = 0; j < m; j++) {
        double v = 0.0;
        for (int k = 0; k < m; k++) v = fma(y[k], Ri[k * m + j], v);
        if (j & 1)
            q1 = fma(v, v, q1);
        else
            q0 = fma(v, v, q0);
        lsd += log(R[j * m + j]);
    }
    *w_out = exp(-0.5 * (q0 + q1) - lsd - 5.5136311992280356);
    for (int i = 0; i < 36; i++) T[i] = Sfull[i];
    mkf_lu_invert6(T, Ri); // S.inv() of KF_model::update
    for (int a = 0; a < m; a++)
        for (int b = 0; b <= a; b++) Wpacked[a * (a + 1) / 2 + b] = 0.5 * (Ri[a * m + b] + Ri[b * m + a]);
}

// -----------------------------------------------------------------------------------------
// the per-slot arithmetic in the measurement-aligned basis (H' = [I 0]):
//   predict      x <- g x + b',  P <- g^2 P + Q'
//   innovation   y = (z - BH) - x1,  S = A + r I   (A = P11)
//   likelihood   pseudo-Cholesky of src/pf2DRao.cpp:34-67 (chol_mode)
//   update       W = S^-1; x1 += y - r W y; x2 += B W y;
//                A <- r I - r^2 W;  B <- r (W B^T)^T;  C <- C - B W B^T
// returns false when cv::Cholesky would have failed (pivot < DBL_EPSILON)
// -----------------------------------------------------------------------------------------
template <int D, bool SLOW>
__device__ __forceinline__ bool slot_math(double (&v)[SlotLay<D>::NE], const double* __restrict__ c,
                                          const double (&zc)[MKF_M], const double r, const int chol_mode,
                                          const int stage, double& w_out)
{
    using L = SlotLay<D>;
    constexpr int M = MKF_M;
#define A_(i, j) v[L::OA + tri((i), (j))]
#define B_(i, a) v[L::OB + (i) * M + (a)]
#define C_(i, j) v[L::OC + tri((i), (j))]
    if (stage & 1) {
        const double g = c[0], g2 = c[1];
#pragma unroll
        for (int e = 0; e < D; e++) v[e] = fma(g, v[e], c[2 + e]);
#pragma unroll
        for (int e = D; e < L::NE; e++) v[e] = fma(g2, v[e], c[2 + e]);
    }
    w_out = 0.0;
    if (!(stage & 2)) return true;

    double y[M];
#pragma unroll
    for (int a = 0; a < M; a++) y[a] = zc[a] - v[a];

    // Cholesky of S = A + r I: Lm lower (strict), inv[i] = 1/L_ii  (cv::Cholesky, CholImpl)
    double Lm[L::NA], inv[M];
    bool ok = true;
#pragma unroll
    for (int i = 0; i < M; i++) {
#pragma unroll
        for (int j = 0; j < i; j++) {
            double s = A_(i, j);
#pragma unroll
            for (int k = 0; k < j; k++) s = fma(-Lm[tri(i, k)], Lm[tri(j, k)], s);
            Lm[tri(i, j)] = s * inv[j];
        }
        double s = A_(i, i) + r;
#pragma unroll
        for (int k = 0; k < i; k++) s = fma(-Lm[tri(i, k)], Lm[tri(i, k)], s);
        if (s < DBL_EPSILON) ok = false;
        inv[i] = rsqrt(s);
        Lm[tri(i, i)] = s * inv[i]; // L_ii
    }

    // likelihood: v = y^T Rt^-1 by forward substitution, w = exp(-q/2 - sum log Rt_ee - 3 log 2pi)
    {
        double vv[M], ve[M], q = 0.0, pinv = 1.0;
        if (chol_mode == MKF_CHOL_EXACT) {
#pragma unroll
            for (int j = 0; j < M; j++) {
                double acc = y[j];
#pragma unroll
                for (int e = 0; e < j; e++) acc = fma(-vv[e], Lm[tri(j, e)], acc);
                vv[j] = acc * inv[j];
                q = fma(vv[j], vv[j], q);
                pinv *= inv[j];
            }
        } else {
            // Rt_ee = 1/elem_e, Rt_ej = S_ej * elem_e (j > e); elem = 1/L_ee (OpenCV 2.4) or L_ee (>= 3.0)
#pragma unroll
            for (int j = 0; j < M; j++) {
                const double elem = (chol_mode == MKF_CHOL_CV24_LITERAL) ? inv[j] : Lm[tri(j, j)];
                double acc = y[j];
#pragma unroll
                for (int e = 0; e < j; e++) acc = fma(-ve[e], A_(j, e), acc);
                vv[j] = acc * elem;
                ve[j] = vv[j] * elem;
                q = fma(vv[j], vv[j], q);
                pinv *= elem;
            }
        }
        // sum_e log Rt_ee = -log(prod elem)
        w_out = exp(fma(-0.5, q, log(pinv)) - 5.5136311992280356); // 3*log(2*pi)
    }

    // W = S^-1 = Linv^T Linv
    double Li[L::NA], W[L::NA];
#pragma unroll
    for (int i = 0; i < M; i++) {
        Li[tri(i, i)] = inv[i];
#pragma unroll
        for (int j = 0; j < i; j++) {
            double s = 0.0;
#pragma unroll
            for (int k = j; k < i; k++) s = fma(Lm[tri(i, k)], Li[tri(k, j)], s);
            Li[tri(i, j)] = -s * inv[i];
        }
    }
#pragma unroll
    for (int a = 0; a < M; a++)
#pragma unroll
        for (int b = 0; b <= a; b++) {
            double s = 0.0;
#pragma unroll
            for (int k = a; k < M; k++) s = fma(Li[tri(k, a)], Li[tri(k, b)], s);
            W[tri(a, b)] = s;
        }
    if (SLOW && !ok) { // literal cv::Cholesky-failure semantics (rare): likelihood from the unfactored matrix, LU inverse
        double Sf[36];
#pragma unroll
        for (int a = 0; a < M; a++)
#pragma unroll
            for (int b = 0; b < M; b++) Sf[a * M + b] = ((a >= b) ? A_(a, b) : A_(b, a)) + ((a == b) ? r : 0.0);
        mkf_chol_fail_path(Sf, y, chol_mode, &w_out, W);
    }
#define W_(a, b) W[((a) >= (b)) ? tri((a), (b)) : tri((b), (a))]

    // state mean
    double t[M];
#pragma unroll
    for (int a = 0; a < M; a++) {
        double s = 0.0;
#pragma unroll
        for (int b = 0; b < M; b++) s = fma(W_(a, b), y[b], s);
        t[a] = s;
    }
#pragma unroll
    for (int a = 0; a < M; a++) v[a] += fma(-r, t[a], y[a]);
#pragma unroll
    for (int i = 0; i < L::D2; i++) {
        double s = v[M + i];
#pragma unroll
        for (int a = 0; a < M; a++) s = fma(B_(i, a), t[a], s);
        v[M + i] = s;
    }

    // covariance
#pragma unroll
    for (int j = 0; j < L::D2; j++) {
        double G[M];
#pragma unroll
        for (int a = 0; a < M; a++) {
            double s = 0.0;
#pragma unroll
            for (int b = 0; b < M; b++) s = fma(W_(a, b), B_(j, b), s);
            G[a] = s;
        }
#pragma unroll
        for (int i = j; i < L::D2; i++) {
            double s = C_(i, j);
#pragma unroll
            for (int a = 0; a < M; a++) s = fma(-B_(i, a), G[a], s);
            C_(i, j) = s;
        }
#pragma unroll
        for (int a = 0; a < M; a++) B_(j, a) = r * G[a];
    }
    const double r2 = r * r;
#pragma unroll
    for (int a = 0; a < M; a++)
#pragma unroll
        for (int b = 0; b <= a; b++) A_(a, b) = fma(-r2, W[tri(a, b)], (a == b) ? r : 0.0);
#undef A_
#undef B_
#undef C_
#undef W_
    return ok;
}

// component of local slot j from the run boundaries written by k_indicator_bounds.  tail: the track's row of
// ind_tail -- after a draw that wrapped past the last component (prior weights summing to less than the thresholds
// reach, src/pf2DRao.cpp:198-207 with idx = (idx+1) % L) the components are no longer monotone in j and are read per slot
__device__ __forceinline__ int mkf_component_of(const int32_t* __restrict__ bt, int K, int j,
                                                const uint8_t* __restrict__ tail)
{
    int k = 0;
    for (int q = 0; q < K - 1; q++) k += (j >= __ldg(bt + q)) ? 1 : 0;
    if (j >= __ldg(bt + K)) {
        const int wk = __ldg(bt + K + 1);
        k = (wk >= 0 || !tail) ? max(wk, 0) : (int)tail[j];
    }
    return k;
}

// measurement column of local slot j of track t, minus BH
// the column of a slot whose candidate bin is bsel (MKF_MEAS_CAND); same operations as k_assoc_meas
__device__ __forceinline__ void mkf_load_meas_cand(const SlotArgs& a, long long t, int bsel, double (&zc)[MKF_M])
{
    const double rx = __ldg(a.roi + t * 4 + 0), ry = __ldg(a.roi + t * 4 + 1), rw = __ldg(a.roi + t * 4 + 2),
                 rh = __ldg(a.roi + t * 4 + 3);
    const double* __restrict__ px = a.cand + (t * 2 + a.hand) * 2 * (long long)a.cand_C;
    const double cxv = __dadd_rn(rx, __ddiv_rn(rw, 2.0));
    zc[0] = cxv - a.bh[0];
    zc[1] = __dadd_rn(ry, __dmul_rn(0.5, rh)) - a.bh[1];
    zc[2] = __ldg(px + bsel) - a.bh[2];
    zc[3] = __ldg(px + a.cand_C + bsel) - a.bh[3];
    zc[4] = cxv - a.bh[4];
    zc[5] = __dadd_rn(ry, __dmul_rn(a.neck, rh)) - a.bh[5];
}

__device__ __forceinline__ void mkf_load_meas(const SlotArgs& a, long long t, int j, double (&zc)[MKF_M])
{
    if (a.meas_layout == MKF_MEAS_SHARED) {
#pragma unroll
        for (int r = 0; r < MKF_M; r++) zc[r] = __ldg(a.meas + t * MKF_M + r) - a.bh[r];
    } else if (a.meas_layout == MKF_MEAS_CAND) {
        mkf_load_meas_cand(a, t, __ldg(a.bins + (t * 2 + a.hand) * (long long)a.N + j), zc);
    } else {
#pragma unroll
        for (int r = 0; r < MKF_M; r++) zc[r] = __ldg(a.meas + (t * MKF_M + r) * a.N + j) - a.bh[r];
    }
}

// One thread per slot.  CHAIN == false (MKF_ALIAS_INDEPENDENT): every slot starts from its parent's
// Gaussian.  CHAIN == true (MKF_ALIAS_CV_SHALLOW_LITERAL, quirk B3 of src/pf2DRao.cpp:153-156): the
// slots that drew the same parent share one cv::Mat buffer in the reference and are predicted/updated
// sequentially in place, so slot j starts from the snapshot slot j-1 left behind.  Parents are sorted
// after systematic resampling, so such slots form a run of consecutive slots: the thread of the run's
// first slot walks the whole run with the Gaussian held in registers and stores a snapshot per slot;
// the other threads of the run retire at once.
// A cv::Cholesky failure only raises MKF_ST_CHOL_FAIL here; k_slot_update_repair then redoes the track
// with the literal failure semantics, which keeps that (never taken in practice) branch out of this kernel.
template <int D, bool CHAIN>
__global__ void __launch_bounds__(128, 2) k_slot_update(const SlotArgs a)
{
    using L = SlotLay<D>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* cst = reinterpret_cast<double*>(smem_raw);
    __shared__ __align__(8) uint64_t mbar;

    const uint32_t cbytes = (uint32_t)(a.K * L::CS * sizeof(double));
    if (threadIdx.x == 0) mkf_mbar_init(&mbar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        mkf_mbar_expect_tx(&mbar, cbytes);
        mkf_tma_load_1d(cst, a.comp_const, cbytes, &mbar); // model constants: never written by the frame chain
    }
    mkf_pdl_launch_dependents();
    mkf_pdl_wait();

    if (a.ts && threadIdx.x == 0) atomicMin(a.ts, mkf_globaltimer());
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool active = s < a.total;
    double v[L::NE];
    double zc[MKF_M];
    int k = 0, j = 0, len = 1;
    long long t = 0;
    const int32_t* __restrict__ bt = nullptr;

    if (active) {
        t = s / a.N;
        j = (int)(s - t * a.N);
        const int par = __ldg(a.src + s);
        if (CHAIN) {
            if (a.unsorted[t]) {
                active = false; // unsorted parents (random-index fallback): k_slot_update_repair walks the track
            } else if (j > 0 && __ldg(a.src + s - 1) == par) {
                active = false; // not the first slot of its run
            } else {
                while (j + len < a.N && __ldg(a.src + s + len) == par) len++;
            }
        }
        if (active) {
            const long long sp = t * a.N + par;
            const double2* __restrict__ src = a.st_in + (sp >> 5) * (long long)L::TILE2 + (sp & 31) * L::H;
#pragma unroll
            for (int p = 0; p < L::NP; p++) {
                const double2 q = __ldg(src + L::po(p));
                v[2 * p] = q.x;
                if (2 * p + 1 < L::NE) v[2 * p + 1] = q.y;
            }
            mkf_load_meas(a, t, j, zc);
            bt = a.bounds + t * (a.K + 2);
            k = mkf_component_of(bt, a.K, j, a.ind_tail ? a.ind_tail + t * a.N : nullptr);
        }
    }
    mkf_mbar_wait(&mbar, 0);
    if (!active) return;

    if (!CHAIN) {
        double w;
        const bool ok = slot_math<D, false>(v, cst + k * L::CS, zc, a.r, a.chol_mode, a.stage, w);
        if (!ok) atomicOr(a.status + t, MKF_ST_CHOL_FAIL);
        double2* __restrict__ dst = a.st_out + (s >> 5) * (long long)L::TILE2 + (s & 31) * L::H;
#pragma unroll
        for (int p = 0; p < L::NP; p++) {
            double2 q;
            q.x = v[2 * p];
            q.y = (2 * p + 1 < L::NE) ? v[2 * p + 1] : 0.0;
            __stcs(dst + L::po(p), q);
        }
        a.w_raw[s] = w;
        if (a.ts && (threadIdx.x & 31) == 0) atomicMax(a.ts + 1, mkf_globaltimer());
    } else {
        for (int i = 0;;) {
            double w;
            const bool ok = slot_math<D, false>(v, cst + k * L::CS, zc, a.r, a.chol_mode, a.stage, w);
            if (!ok) atomicOr(a.status + t, MKF_ST_CHOL_FAIL);
            const long long so = s + i;
            double2* __restrict__ dst = a.st_out + (so >> 5) * (long long)L::TILE2 + (so & 31) * L::H;
#pragma unroll
            for (int p = 0; p < L::NP; p++) {
                double2 q;
                q.x = v[2 * p];
                q.y = (2 * p + 1 < L::NE) ? v[2 * p + 1] : 0.0;
                __stcs(dst + L::po(p), q);
            }
            a.w_raw[so] = w;
            if (++i >= len) break;
            mkf_load_meas(a, t, j + i, zc);
            k = mkf_component_of(bt, a.K, j + i, a.ind_tail ? a.ind_tail + t * a.N : nullptr);
        }
    }
}

// -----------------------------------------------------------------------------------------
// MKF_ALIAS_CV_SHALLOW_LITERAL with dynamic run assignment.  k_slot_update<D, true> gives every run of duplicates to
// the thread of its first slot: half the lanes of a warp retire at once (they are not run starts) and the rest wait for
// the longest run among them (mean run length 2, the longest of 32 is 6-8: ncu showed ~10 of 32 lanes active).  Here
//   k_alias_runs             lists the run starts (one atomicAdd per 1024-slot chunk, like k_share_keys);
//   k_slot_update_chain_dyn  persistent warps whose lanes each walk one run at a time -- predict / likelihood / update,
//                            snapshot stored per slot, exactly the in-place sequence of src/pf2DRao.cpp:134-142 on a
//                            shared cv::Mat -- and, when a lane's run ends, take the next run of the warp's stretch
//                            of the list.  Every iteration of the warp's loop is then the
//                            gather / arithmetic / store step of the independent kernel with all lanes busy; a parent
//                            is read once per run, so the kernel moves N x 720 B of snapshots + runs x 720 B of parents.
// Tracks with unsorted parents (after the cv::RNG fallback) are left to k_slot_update_repair, as before.
// -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_alias_runs(const int32_t* __restrict__ src, const uint32_t* __restrict__ unsorted,
                                                    long long total, int N, int* __restrict__ list,
                                                    int* __restrict__ count)
{
    constexpr int G = 4;
    __shared__ int warp_tot[8];
    __shared__ int list_base;
    mkf_pdl_launch_dependents();
    mkf_pdl_wait();
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const long long s0 = ((long long)blockIdx.x * 256 + tid) * G;
    unsigned flags = 0;
    if (s0 < total) {
        int prev = s0 > 0 ? __ldg(src + s0 - 1) : -1;
        long long t = s0 / N;
        int j = (int)(s0 - t * N);
        bool skip = unsorted[t] != 0u;
#pragma unroll
        for (int g = 0; g < G; g++) {
            if (s0 + g >= total) break;
            const int cur = __ldg(src + s0 + g);
            if (!skip && (j == 0 || cur != prev)) flags |= 1u << g;
            prev = cur;
            if (++j == N) {
                j = 0;
                t++;
                skip = (s0 + g + 1 < total) ? unsorted[t] != 0u : false;
            }
        }
    }
    const int cnt = __popc(flags);
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    int before = inc - cnt;
    for (int w = 0; w < wid; w++) before += warp_tot[w];
    if (tid == 255) list_base = atomicAdd(count, before + cnt);
    __syncthreads();
    int r = list_base + before;
#pragma unroll
    for (int g = 0; g < G; g++)
        if (flags & (1u << g)) list[r++] = (int)(s0 + g);
}

template <int D>
__global__ void __launch_bounds__(128, 2) k_slot_update_chain_dyn(const SlotArgs a, const int* __restrict__ list,
                                                                  const int* __restrict__ count, int* __restrict__ next,
                                                                  int* __restrict__ to_clear)
{
    using L = SlotLay<D>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* cst = reinterpret_cast<double*>(smem_raw);
    __shared__ __align__(8) uint64_t mbar;
    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t cbytes = (uint32_t)(a.K * L::CS * sizeof(double));
    if (tid == 0) mkf_mbar_init(&mbar, 1);
    __syncthreads();
    if (tid == 0) {
        mkf_mbar_expect_tx(&mbar, cbytes);
        mkf_tma_load_1d(cst, a.comp_const, cbytes, &mbar);
    }
    mkf_pdl_launch_dependents();
    mkf_pdl_wait();
    const int n = *reinterpret_cast<const volatile int*>(count);
    if (blockIdx.x == 0 && tid == 0) { // the counters the NEXT frame's k_alias_runs / this kernel start from
        to_clear[0] = 0;
        to_clear[1] = 0;
    }
    mkf_mbar_wait(&mbar, 0);

    // A lane holds the run it is walking and, claimed one step ahead, the run it will walk next: the claim (atomicAdd,
    // list entry, parent index -- three dependent round trips) is issued while the current step computes, so that a
    // lane whose run ends only has the gather of the new parent in front of its next step.
    double v[L::NE];
    bool have = false, has_next = false, drained = false;
    long long s = 0, t = 0;
    int j = 0, par = 0, ns = 0, npar = 0;
    // the list is split evenly over the warps of the grid (run lengths average out over the ~900 runs of a stretch)
    const long long nwarps = (long long)gridDim.x * 4, gw = (long long)blockIdx.x * 4 + (tid >> 5);
    long long wcur = (long long)n * gw / nwarps;
    const long long wend = (long long)n * (gw + 1) / nwarps;
    for (;;) {
        if (!have && has_next) { // start the run claimed earlier: gather its parent
            s = ns;
            t = s / a.N;
            j = (int)(s - t * a.N);
            par = npar;
            const long long sp = t * a.N + par;
            const double2* __restrict__ srcp = a.st_in + (sp >> 5) * (long long)L::TILE2 + (sp & 31) * L::H;
#pragma unroll
            for (int p = 0; p < L::NP; p++) {
                const double2 q = __ldg(srcp + L::po(p));
                v[2 * p] = q.x;
                if (2 * p + 1 < L::NE) v[2 * p + 1] = q.y;
            }
            have = true;
            has_next = false;
        }
        // claim ahead: lanes without a next run take the following entries of this warp's stretch of the list.  (A
        // batch-wide cursor advanced with one atomicAdd per refill was the first version: 64 k same-address atomics per
        // frame serialise in the L2 -- 1.13 ms per frame against 0.91 for the static kernel.)
        const bool need = !has_next && !drained;
        const unsigned mask = __ballot_sync(0xffffffffu, need);
        if (mask) {
            if (need) {
                const long long my = wcur + __popc(mask & ((1u << lane) - 1u));
                if (my < wend) {
                    ns = __ldg(list + my);
                    npar = __ldg(a.src + ns);
                    has_next = true;
                } else {
                    drained = true;
                }
            }
            wcur += __popc(mask);
        }
        if (!__any_sync(0xffffffffu, have || has_next)) break;
        if (have) {
            // does the next slot continue this run?  (it drew the same parent: it shares the cv::Mat updated below)
            const bool cont = (j + 1 < a.N) && (__ldg(a.src + s + 1) == par);
            double zc[MKF_M], w;
            mkf_load_meas(a, t, j, zc);
            const int k = mkf_component_of(a.bounds + t * (a.K + 2), a.K, j, a.ind_tail ? a.ind_tail + t * a.N : nullptr);
            const bool ok = slot_math<D, false>(v, cst + k * L::CS, zc, a.r, a.chol_mode, a.stage, w);
            if (!ok) atomicOr(a.status + t, MKF_ST_CHOL_FAIL);
            double2* __restrict__ dst = a.st_out + (s >> 5) * (long long)L::TILE2 + (s & 31) * L::H;
#pragma unroll
            for (int p = 0; p < L::NP; p++) {
                double2 q;
                q.x = v[2 * p];
                q.y = (2 * p + 1 < L::NE) ? v[2 * p + 1] : 0.0;
                dst[L::po(p)] = q; // (not streaming: the neighbouring slot's half of the sector arrives a step or two
                                   // later, from another lane -- evict-first stores turned those into read-modify-writes)
            }
            a.w_raw[s] = w;
            s++;
            j++;
            have = cont;
        }
    }
}

// -----------------------------------------------------------------------------------------
// Slot update with record sharing (MKF_ALIAS_INDEPENDENT, one measurement per track).
// Children of one parent RECORD that drew the same component are bit-identical Gaussians, and because resampled parents
// are sorted they are consecutive slots.  A CTA takes CHUNK = 128 G consecutive slots:
//   A  every slot's key (track, parent record, component) goes to shared memory; a slot whose key differs from its
//      predecessor's is a HEAD; a block scan numbers the heads;
//   B  the heads -- and only they, packed into full warps -- gather the parent record, run the slot arithmetic and
//      store the child record.  Records of one track are packed from the track's first slot in the chunk onwards, so
//      the next frame gathers from dense sectors;
//   C  every slot takes its head's weight and notes in rep[] which record holds its state.
// In steady state the 4096 x 500 benchmark keeps ~11 % distinct records.  (A first version shared per warp -- head
// lanes compute, the others take the weight by shuffle: it saved the DRAM traffic but still issued every warp's
// arithmetic for a few live lanes, ncu: 10 of 32 lanes active, 0.306 ms; packing the heads of a whole chunk into
// full warps makes the arithmetic shrink with the number of distinct Gaussians as well.)
// -----------------------------------------------------------------------------------------
template <int D, int G>
__global__ void __launch_bounds__(128, 2) k_slot_update_shared(const SlotArgs a)
{
    using L = SlotLay<D>;
    constexpr int CHUNK = 128 * G;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* cst = reinterpret_cast<double*>(smem_raw);
    double* h_w = cst + a.K * L::CS;                          // [CHUNK] weight of head h
    int* sm_par = reinterpret_cast<int*>(h_w + CHUNK);        // [CHUNK] per slot: parent record
    int* sm_t = sm_par + CHUNK;                               // [CHUNK] per slot: track (-1 beyond the end)
    int* sm_rank = sm_t + CHUNK;                              // [CHUNK] per slot: index of its head
    int* h_slot = sm_rank + CHUNK;                            // [CHUNK] head h -> its slot offset in the chunk
    unsigned char* sm_k = reinterpret_cast<unsigned char*>(h_slot + CHUNK); // [CHUNK] per slot: component
    __shared__ __align__(8) uint64_t mbar;
    __shared__ int warp_tot[4];

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t cbytes = (uint32_t)(a.K * L::CS * sizeof(double));
    if (tid == 0) mkf_mbar_init(&mbar, 1);
    __syncthreads();
    if (tid == 0) {
        mkf_mbar_expect_tx(&mbar, cbytes);
        mkf_tma_load_1d(cst, a.comp_const, cbytes, &mbar); // model constants: never written by the frame chain
    }
    mkf_pdl_launch_dependents();
    mkf_pdl_wait();

    const long long base = (long long)blockIdx.x * CHUNK;
    // ---- A: keys.  A thread owns G consecutive slots (vector loads / stores; one division, then increments)
    static_assert(G % 4 == 0, "vector accesses below assume groups of four slots");
    const int so0 = tid * G;
    const long long s0 = base + so0;
    int H;
    {
        int par[G], tl[G], kk[G];
        int t_run = 0, j_run = 0;
        if (s0 < a.total) {
            t_run = (int)((unsigned)s0 / (unsigned)a.N); // T*N < 2^32: 180 GB hold at most 1.25e8 slots
            j_run = (int)((unsigned)s0 - (unsigned)t_run * (unsigned)a.N);
        }
        if (s0 + G <= a.total) {
            const int4* __restrict__ sv = reinterpret_cast<const int4*>(a.src + s0);
#pragma unroll
            for (int q = 0; q < G / 4; q++) {
                const int4 v4 = __ldg(sv + q);
                par[4 * q] = v4.x;
                par[4 * q + 1] = v4.y;
                par[4 * q + 2] = v4.z;
                par[4 * q + 3] = v4.w;
            }
        } else {
#pragma unroll
            for (int g = 0; g < G; g++) par[g] = (s0 + g < a.total) ? __ldg(a.src + s0 + g) : -1;
        }
#pragma unroll
        for (int g = 0; g < G; g++) {
            if (s0 + g < a.total) {
                tl[g] = t_run;
                kk[g] = mkf_component_of(a.bounds + (long long)t_run * (a.K + 2), a.K, j_run,
                                         a.ind_tail ? a.ind_tail + (long long)t_run * a.N : nullptr);
                if (++j_run == a.N) {
                    j_run = 0;
                    t_run++;
                }
            } else {
                tl[g] = -1;
                kk[g] = 0;
                par[g] = -1;
            }
            sm_par[so0 + g] = par[g];
            sm_t[so0 + g] = tl[g];
            sm_k[so0 + g] = (unsigned char)kk[g];
        }
        __syncthreads();
        // heads, numbered in slot order: flags of my G slots, one block scan of the per-thread counts
        unsigned flags = 0;
        int p_par = -2, p_t = -2, p_k = -1; // the slot before my first one (none for the chunk's first slot)
        if (so0 > 0) {
            p_par = sm_par[so0 - 1];
            p_t = sm_t[so0 - 1];
            p_k = sm_k[so0 - 1];
        }
#pragma unroll
        for (int g = 0; g < G; g++) {
            if (tl[g] >= 0 && (par[g] != p_par || kk[g] != p_k || tl[g] != p_t)) flags |= 1u << g;
            p_par = par[g];
            p_t = tl[g];
            p_k = kk[g];
        }
        const int cnt = __popc(flags);
        int inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += n;
        }
        if (lane == 31) warp_tot[wid] = inc;
        __syncthreads();
        int before = inc - cnt;
        for (int w = 0; w < wid; w++) before += warp_tot[w];
        H = warp_tot[0] + warp_tot[1] + warp_tot[2] + warp_tot[3];
        int r = before - 1; // index of the head governing the slot before my first one
#pragma unroll
        for (int g = 0; g < G; g++) {
            if (flags & (1u << g)) h_slot[++r] = so0 + g;
            sm_rank[so0 + g] = r;
        }
    }
    __syncthreads();

    // ---- B: one thread per head
    mkf_mbar_wait(&mbar, 0);
    for (int h = tid; h < H; h += 128) {
        const int so = h_slot[h];
        const long long s = base + so;
        const long long t = sm_t[so];
        const int j = (int)(s - t * a.N);
        const int k = sm_k[so];
        const long long sp = t * a.N + sm_par[so];
        const double2* __restrict__ src = a.st_in + (sp >> 5) * (long long)L::TILE2 + (sp & 31) * L::H;
        double v[L::NE];
#pragma unroll
        for (int p = 0; p < L::NP; p++) {
            const double2 q = __ldg(src + L::po(p));
            v[2 * p] = q.x;
            if (2 * p + 1 < L::NE) v[2 * p + 1] = q.y;
        }
        double zc[MKF_M];
        mkf_load_meas(a, t, j, zc);
        double w;
        const bool ok = slot_math<D, false>(v, cst + k * L::CS, zc, a.r, a.chol_mode, a.stage, w);
        if (!ok) atomicOr(a.status + t, MKF_ST_CHOL_FAIL);
        // record position: the track's first slot in this chunk + the head's number within the track
        const long long seg = t * a.N > base ? t * a.N - base : 0;
        const long long so_rec = base + seg + (h - sm_rank[(int)seg]);
        double2* __restrict__ dst = a.st_out + (so_rec >> 5) * (long long)L::TILE2 + (so_rec & 31) * L::H;
#pragma unroll
        for (int p = 0; p < L::NP; p++) {
            double2 q;
            q.x = v[2 * p];
            q.y = (2 * p + 1 < L::NE) ? v[2 * p + 1] : 0.0;
            __stcs(dst + L::po(p), q);
        }
        h_w[h] = w;
    }
    __syncthreads();

    // ---- C: per-slot outputs (my G consecutive slots again)
    {
        double wv[G];
        int rv[G];
#pragma unroll
        for (int g = 0; g < G; g++) {
            const int h = sm_rank[so0 + g];
            const long long t = sm_t[so0 + g];
            const long long seg = (t >= 0 && t * a.N > base) ? t * a.N - base : 0;
            wv[g] = t >= 0 ? h_w[h] : 0.0;
            rv[g] = (int)(base + seg + (h - sm_rank[(int)seg]) - t * a.N);
        }
        if (s0 + G <= a.total) {
            double2* __restrict__ wo = reinterpret_cast<double2*>(a.w_raw + s0);
            int4* __restrict__ ro = reinterpret_cast<int4*>(a.rep + s0);
#pragma unroll
            for (int q = 0; q < G / 2; q++) wo[q] = make_double2(wv[2 * q], wv[2 * q + 1]);
#pragma unroll
            for (int q = 0; q < G / 4; q++) ro[q] = make_int4(rv[4 * q], rv[4 * q + 1], rv[4 * q + 2], rv[4 * q + 3]);
        } else {
#pragma unroll
            for (int g = 0; g < G; g++)
                if (s0 + g < a.total) {
                    a.w_raw[s0 + g] = wv[g];
                    a.rep[s0 + g] = rv[g];
                }
        }
    }
}

// -----------------------------------------------------------------------------------------
// Record sharing in two launches (the default; k_slot_update_shared above remains for models whose constants leave no
// room and as the A/B reference, MKF_SHARE_SPLIT=0).  ncu on k_slot_update_shared showed the slot arithmetic to be
// 20-25 % of the stall samples: the rest was a chain of dependent global loads (parent index -> record gather ... per
// slot head index) and the bookkeeping instructions, all executed at the 8 warps per SM that 255 registers allow, and
// the CTA barrier in front of the per-slot outputs.  So:
//   k_share_keys                 256 threads x 4 consecutive slots = one 1024-slot chunk at full occupancy: keys, head
//                                flags, block scan; writes rep[] (final) and appends one 16-byte record per head
//                                {source record, destination record, track, component} to a batch-wide list (one
//                                atomicAdd per chunk; the order of the chunks in the list is immaterial);
//   k_slot_update_heads_direct   grid-stride over the list (12 CTAs per SM, 2 resident), one thread per list entry, no barrier: gather, slot
//                                arithmetic, store; the weight goes to w_rec[] at the record's position.  One
//                                dependent hop (the gather) in front of the arithmetic; the next step's list entry is
//                                already in flight.  The kernel moves 1 440 B per distinct Gaussian and runs at ~65 %
//                                of the HBM copy rate on them.
// The resampler reads slot i's weight as w_rec[rep[i]] (k_resample_block, w_slot_out) and materialises w_raw on the way.
// Same chunking and record placement as k_slot_update_shared, so everything downstream is unchanged.
// Tried and dropped: staging the next step's parent record in shared memory with cp.async (per-thread slices, no
// registers), CTA- and warp-level variants: the extra LSU traffic (48 LDGSTS + 48 LDS per head) and the spills it
// caused cost more than the hidden latency gained (73-88 us against 67 us); prefetch.global.L2 of the next step's
// record: 69 us.
// -----------------------------------------------------------------------------------------
constexpr int MKF_SHARE_CHUNK = 1024;

template <bool WITH_BIN> // WITH_BIN: MKF_MEAS_CAND, the candidate bin is part of the key
__global__ void __launch_bounds__(256) k_share_keys(const SlotArgs a)
{
    using key_t = typename std::conditional<WITH_BIN, int, unsigned char>::type;
    constexpr int CHUNK = MKF_SHARE_CHUNK, G = 4;
    __shared__ int sm_par[CHUNK], sm_t[CHUNK], sm_rank[CHUNK];
    __shared__ key_t sm_k[CHUNK]; // component (| candidate bin << 8 with MKF_MEAS_CAND)
    __shared__ int warp_tot[8];
    __shared__ int list_base;
    mkf_pdl_launch_dependents();
    mkf_pdl_wait();

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const long long base = (long long)blockIdx.x * CHUNK;
    const int so0 = tid * G;
    const long long s0 = base + so0;
    int par[G], tl[G], kk[G];
    int t_run = 0, j_run = 0;
    if (s0 < a.total) {
        t_run = (int)((unsigned)s0 / (unsigned)a.N);
        j_run = (int)((unsigned)s0 - (unsigned)t_run * (unsigned)a.N);
    }
    if (s0 + G <= a.total) {
        const int4 v4 = __ldg(reinterpret_cast<const int4*>(a.src + s0));
        par[0] = v4.x;
        par[1] = v4.y;
        par[2] = v4.z;
        par[3] = v4.w;
    } else {
#pragma unroll
        for (int g = 0; g < G; g++) par[g] = (s0 + g < a.total) ? __ldg(a.src + s0 + g) : -1;
    }
    {
        // component of a slot = number of run boundaries e_0..e_{K-2} (non-decreasing) that are <= j, unless j lies in
        // the wrapped tail (mkf_component_of).  Counted once for my first slot of a track, then carried forward.
        const int K = a.K;
        const int32_t* bt = a.bounds;
        int k_lin = 0, nb = 0, wf = 0, wk = 0;
        bool fresh = true;
#pragma unroll
        for (int g = 0; g < G; g++) {
            if (s0 + g < a.total) {
                if (fresh) {
                    bt = a.bounds + (long long)t_run * (K + 2);
                    k_lin = 0;
                    for (int q = 0; q < K - 1; q++) k_lin += (j_run >= __ldg(bt + q)) ? 1 : 0;
                    wf = __ldg(bt + K);
                    wk = __ldg(bt + K + 1);
                    fresh = false;
                    nb = (k_lin < K - 1) ? __ldg(bt + k_lin) : 0x7fffffff;
                } else {
                    while (j_run >= nb) {
                        k_lin++;
                        nb = (k_lin < K - 1) ? __ldg(bt + k_lin) : 0x7fffffff;
                    }
                }
                tl[g] = t_run;
                kk[g] = (j_run >= wf) ? ((wk >= 0 || !a.ind_tail) ? max(wk, 0) : (int)a.ind_tail[(long long)t_run * a.N + j_run])
                                      : k_lin;
                // slots of a track that drew the same candidate see the same measurement column
                if (WITH_BIN) kk[g] |= __ldg(a.bins + ((long long)t_run * 2 + a.hand) * a.N + j_run) << 8;
                if (++j_run == a.N) {
                    j_run = 0;
                    t_run++;
                    fresh = true;
                }
            } else {
                tl[g] = -1;
                kk[g] = 0;
                par[g] = -1;
            }
            sm_par[so0 + g] = par[g];
            sm_t[so0 + g] = tl[g];
            sm_k[so0 + g] = (key_t)kk[g];
        }
    }
    __syncthreads();
    unsigned flags = 0;
    int p_par = -2, p_t = -2, p_k = -1; // the slot before my first one (none for the chunk's first slot)
    if (so0 > 0) {
        p_par = sm_par[so0 - 1];
        p_t = sm_t[so0 - 1];
        p_k = sm_k[so0 - 1];
    }
#pragma unroll
    for (int g = 0; g < G; g++) {
        if (tl[g] >= 0 && (par[g] != p_par || kk[g] != p_k || tl[g] != p_t)) flags |= 1u << g;
        p_par = par[g];
        p_t = tl[g];
        p_k = kk[g];
    }
    const int cnt = __popc(flags);
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    int before = inc - cnt;
    for (int w = 0; w < wid; w++) before += warp_tot[w];
    int rk[G];
    {
        int r = before - 1; // index of the head governing the slot before my first one
#pragma unroll
        for (int g = 0; g < G; g++) {
            if (flags & (1u << g)) ++r;
            rk[g] = r;
            sm_rank[so0 + g] = r;
        }
    }
    if (tid == 255) list_base = atomicAdd(a.head_count, before + cnt); // this chunk's stretch of the head list
    __syncthreads();
    // record position: the track's first slot in this chunk + the head's number within the track
    int rv[G];
    const long long lb = list_base;
#pragma unroll
    for (int g = 0; g < G; g++) {
        const long long tN = (long long)tl[g] * a.N;
        const long long seg = (tl[g] >= 0 && tN > base) ? tN - base : 0;
        rv[g] = (int)(base + seg + (rk[g] - sm_rank[(int)seg]) - tN);
        if (flags & (1u << g)) a.hd16[lb + rk[g]] = make_int4((int)(tN + par[g]), (int)(tN + rv[g]), tl[g], kk[g]);
    }
    if (s0 + G <= a.total) {
        *reinterpret_cast<int4*>(a.rep + s0) = make_int4(rv[0], rv[1], rv[2], rv[3]);
    } else {
#pragma unroll
        for (int g = 0; g < G; g++)
            if (s0 + g < a.total) a.rep[s0 + g] = rv[g];
    }
}

template <int D>
__global__ void __launch_bounds__(128, 2) k_slot_update_heads_direct(const SlotArgs a, int* __restrict__ count_to_clear)
{
    using L = SlotLay<D>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* cst = reinterpret_cast<double*>(smem_raw); // K x CS model constants (TMA)
    __shared__ __align__(8) uint64_t mbar;

    const int tid = threadIdx.x;
    const uint32_t cbytes = (uint32_t)(a.K * L::CS * sizeof(double));
    if (tid == 0) mkf_mbar_init(&mbar, 1);
    __syncthreads();
    if (tid == 0) {
        mkf_mbar_expect_tx(&mbar, cbytes);
        mkf_tma_load_1d(cst, a.comp_const, cbytes, &mbar); // model constants: never written by the frame chain
    }
    mkf_pdl_launch_dependents();
    mkf_pdl_wait();

    const int n = *reinterpret_cast<const volatile int*>(a.head_count);
    if (blockIdx.x == 0 && tid == 0) *count_to_clear = 0; // the counter the next frame's k_share_keys appends with
    if (a.ts && tid == 0) atomicMin(a.ts, mkf_globaltimer());
    MKF_TL_START(1, a.dbg_frame);
    const int step = (int)(gridDim.x * blockDim.x); // (the block size is a launch parameter: 128, or 64 for A/B runs)
    int h = (int)(blockIdx.x * blockDim.x) + tid;
    int4 rec = make_int4(-1, 0, 0, 0);
    if (h < n) rec = __ldg(a.hd16 + h);
    mkf_mbar_wait(&mbar, 0);
    while (h < n) {
        const long long sp = (unsigned)rec.x, so_rec = (unsigned)rec.y, t = rec.z;
        const int k = rec.w & 0xff;
        const double2* __restrict__ src = a.st_in + (sp >> 5) * (long long)L::TILE2 + (sp & 31) * L::H;
        double v[L::NE];
        if (t < a.l2_tracks) {
            const uint64_t pol = mkf_l2_policy<true>();
#pragma unroll
            for (int p = 0; p < L::NP; p++) {
                const double2 q = mkf_ldg_policy(src + L::po(p), pol);
                v[2 * p] = q.x;
                if (2 * p + 1 < L::NE) v[2 * p + 1] = q.y;
            }
        } else if (a.l2_tracks) {
            const uint64_t pol = mkf_l2_policy<false>();
#pragma unroll
            for (int p = 0; p < L::NP; p++) {
                const double2 q = mkf_ldg_policy(src + L::po(p), pol);
                v[2 * p] = q.x;
                if (2 * p + 1 < L::NE) v[2 * p + 1] = q.y;
            }
        } else {
#pragma unroll
            for (int p = 0; p < L::NP; p++) {
                const double2 q = __ldg(src + L::po(p));
                v[2 * p] = q.x;
                if (2 * p + 1 < L::NE) v[2 * p + 1] = q.y;
            }
        }
        double zc[MKF_M];
        if (a.meas_layout == MKF_MEAS_CAND)
            mkf_load_meas_cand(a, t, rec.w >> 8, zc); // the column of the head's candidate bin
        else
            mkf_load_meas(a, t, 0, zc); // shared layout: the track's column
        h += step;
        if (h < n) rec = __ldg(a.hd16 + h); // next step's head record: in flight during the arithmetic
        double w;
        const bool ok = slot_math<D, false>(v, cst + k * L::CS, zc, a.r, a.chol_mode, a.stage, w);
        if (!ok) atomicOr(a.status + t, MKF_ST_CHOL_FAIL);
        double2* __restrict__ dst = a.st_out + (so_rec >> 5) * (long long)L::TILE2 + (so_rec & 31) * L::H;
        if (t < a.l2_tracks) {
            const uint64_t pol = mkf_l2_policy<true>();
#pragma unroll
            for (int p = 0; p < L::NP; p++) {
                double2 q;
                q.x = v[2 * p];
                q.y = (2 * p + 1 < L::NE) ? v[2 * p + 1] : 0.0;
                mkf_stg_policy(dst + L::po(p), q, pol);
            }
        } else {
#pragma unroll
            for (int p = 0; p < L::NP; p++) {
                double2 q;
                q.x = v[2 * p];
                q.y = (2 * p + 1 < L::NE) ? v[2 * p + 1] : 0.0;
                __stcs(dst + L::po(p), q);
            }
        }
        a.w_rec[so_rec] = w;
    }
    if (a.ts && (tid & 31) == 0) atomicMax(a.ts + 1, mkf_globaltimer());
    MKF_TL_END(1, a.dbg_frame);
}

// what the tail of a short track's frame needs besides SlotArgs
struct SmallTailArgs {
    // indicator draw
    const double* __restrict__ u_ind;
    const double* __restrict__ cw_hi;
    const double* __restrict__ cw_lo;
    const double* __restrict__ wprior;
    double wmax;
    int32_t* __restrict__ bounds_out;
    uint8_t* __restrict__ ind_tail_out;
    int clear_status;
    // posterior resample
    const double* __restrict__ u_post;
    const uint64_t* __restrict__ seeds;
    int seed_stride, seed_off;
    int32_t* __restrict__ parent_out; // == SlotArgs::parent (a track's row is read at the start and written at the
                                      // end by the same half warp)
    double* __restrict__ wsum;
    uint32_t* __restrict__ unsorted_out;
    // estimator
    int Dpose;
    const double* __restrict__ recon;
    const double* __restrict__ pmean;
    const double* __restrict__ tinv;
    double* __restrict__ est_xbar;  // T x d
    double* __restrict__ est_pose;  // T x Dpose
    double* __restrict__ est_pose2; // the batch's pose cache for the next association step, or null
    // k_frame_small only: [per-component constants K x CS | reconstruction coefficients [c][r], r < Dpose + d (rows of
    // recon, then rows of tinv)] as one block for one TMA bulk copy, its size in bytes, and fl(1 / N)
    const double* __restrict__ small_const;
    unsigned small_const_bytes;
    double step;
};

// The tail of ONE track by ONE thread, from global memory (k_slot_update_repair, after it has redone a flagged track):
// same operations in the same order as the half-warp version below, except the estimate's summation order.
template <int D>
__device__ __noinline__ void mkf_small_tail_serial(const SlotArgs& a, const SmallTailArgs& s, const long long t)
{
    using L = SlotLay<D>;
    const int N = a.N;
    const double* __restrict__ w = a.w_raw + t * N;
    double wsum = 0.0;
    for (int i = 0; i < N; i++) wsum = __dadd_rn(wsum, w[i]);
    s.wsum[t] = wsum;
    double mw = 0.0;
    for (int i = 0; i < N; i++) {
        const double x = __ddiv_rn(w[i], wsum);
        if (x > mw) mw = x;
    }
    int32_t* __restrict__ out = s.parent_out + t * N;
    if (!(mw > 0.0)) {
        atomicOr(a.status + t, MKF_ST_POST_DEGENERATE);
        mkf_cvrng rng(s.seeds ? s.seeds[t * s.seed_stride + s.seed_off] : 1ull);
        (void)rng.uniform_int(0, N);
        for (int i = 0; i < N; i++) out[i] = rng.uniform_int(0, N);
    } else {
        mkf_resample_sequential([&](int i) { return __ddiv_rn(w[i], wsum); }, N, N, s.u_post[t],
                                [&](int i, int idx) { out[i] = idx; });
    }
    if (s.unsorted_out) s.unsorted_out[t] = (mw > 0.0) ? 0u : 1u;
    double acc[D];
    for (int e = 0; e < D; e++) acc[e] = 0.0;
    for (int i = 0; i < N; i++) {
        const long long sp = t * N + out[i];
        const double2* src = a.st_out + (sp >> 5) * (long long)L::TILE2 + (sp & 31) * L::H;
        for (int p = 0; p < D / 2; p++) {
            const double2 q = src[L::po(p)];
            acc[2 * p] += q.x;
            acc[2 * p + 1] += q.y;
        }
    }
    const double inv_n = 1.0 / (double)N;
    for (int e = 0; e < D; e++) acc[e] *= inv_n;
    for (int r = 0; r < s.Dpose; r++) {
        double sacc = 0.0;
        for (int c = 0; c < D; c++) sacc = fma(s.recon[r * D + c], acc[c], sacc);
        s.est_pose[t * s.Dpose + r] = sacc + s.pmean[r];
        if (s.est_pose2) s.est_pose2[t * s.Dpose + r] = sacc + s.pmean[r];
    }
    for (int r = 0; r < D; r++) {
        double sacc = 0.0;
        for (int c = 0; c < D; c++) sacc = fma(s.tinv[r * D + c], acc[c], sacc);
        s.est_xbar[t * D + r] = sacc;
    }
}

// Rare tracks redone after k_slot_update: (i) a cv::Cholesky failure was flagged (literal failure semantics
// through slot_math<SLOW>), (ii) literal alias mode with UNSORTED parents (after the degenerate random-index
// fallback of src/pf2DRao.cpp:184-192 the slots sharing a parent are not adjacent).  One CTA scans 128
// tracks' flags.  Independent mode: the CTA's threads stride over the track's slots.  Literal alias mode:
// one thread walks the slots in order, each starting from the latest snapshot taken for its parent
// (last: T x N scratch, -1 = none yet).  st_in is intact (ping-pong), so everything is recomputed from it.
// TAIL (after k_frame_small, mkf_frame_small.cuh): the redone track's resample and estimate follow, by one thread.
template <int D, bool TAIL>
__global__ void __launch_bounds__(128, 2) k_slot_update_repair(const SlotArgs a, int32_t* __restrict__ last,
                                                               const SmallTailArgs tl)
{
    using L = SlotLay<D>;
    __shared__ uint32_t flags[128];
    mkf_pdl_launch_dependents();
    mkf_pdl_wait();
    const long long T = a.total / a.N;
    const long long base = (long long)blockIdx.x * 128;
    {
        const long long t = base + threadIdx.x;
        uint32_t f = 0;
        if (t < T) f = (a.status[t] & MKF_ST_CHOL_FAIL) || (a.alias_chain && a.unsorted[t]);
        flags[threadIdx.x] = f;
        if (!__syncthreads_or((int)f)) return;
    }
    for (int q = 0; q < 128; q++) {
        if (!flags[q]) continue;
        const long long t = base + q;
        const int32_t* __restrict__ bt = a.bounds + t * (a.K + 2);
        int32_t* __restrict__ lt = last ? last + t * a.N : nullptr;
        if (a.alias_chain && threadIdx.x == 0)
            for (int j = 0; j < a.N; j++) lt[j] = -1;
        const int j0 = a.alias_chain ? 0 : (int)threadIdx.x;
        const int jstep = a.alias_chain ? 1 : 128;
        if (!a.alias_chain || threadIdx.x == 0) {
            for (int j = j0; j < a.N; j += jstep) {
                const long long s = t * a.N + j;
                const int par = a.src[s];
                const int snap = a.alias_chain ? lt[par] : -1;
                const double2* base_p = snap >= 0 ? (const double2*)a.st_out : a.st_in;
                const long long sp = t * a.N + (snap >= 0 ? snap : par);
                const double2* src = base_p + (sp >> 5) * (long long)L::TILE2 + (sp & 31) * L::H;
                double v[L::NE];
                for (int p = 0; p < L::NP; p++) {
                    const double2 qq = src[L::po(p)]; // plain load: may be a snapshot this thread stored earlier
                    v[2 * p] = qq.x;
                    if (2 * p + 1 < L::NE) v[2 * p + 1] = qq.y;
                }
                double zc[MKF_M], w;
                mkf_load_meas(a, t, j, zc);
                const int k = mkf_component_of(bt, a.K, j, a.ind_tail ? a.ind_tail + t * a.N : nullptr);
                slot_math<D, true>(v, a.comp_const + (long long)k * L::CS, zc, a.r, a.chol_mode, a.stage, w);
                double2* dst = a.st_out + (s >> 5) * (long long)L::TILE2 + (s & 31) * L::H;
                for (int p = 0; p < L::NP; p++) {
                    double2 qq;
                    qq.x = v[2 * p];
                    qq.y = (2 * p + 1 < L::NE) ? v[2 * p + 1] : 0.0;
                    dst[L::po(p)] = qq;
                }
                a.w_raw[s] = w;
                if (a.dedup) a.rep[s] = j; // the redone track stores every slot at its own position
                if (a.split) a.w_rec[s] = w; // ... and the resampler reads the weights through rep[]
                if (a.alias_chain) lt[par] = j;
            }
        }
        __syncthreads();
        if (TAIL && threadIdx.x == 0) mkf_small_tail_serial<D>(a, tl, t);
    }
}

// -----------------------------------------------------------------------------------------
// K -> N indicator resample: per-track run boundaries e_k = #{j : indicator_j <= k}.
// GROUP lanes (a power of two >= K, <= 32) cooperate on one track: lane k evaluates e_k in closed form
// against the double-double prefix sum of the prior weights; if any lane is ambiguous (or the thresholds
// run past the total prior mass) lane 0 replays the literal loop.  bounds[t] = { e_0..e_{K-1}, wrap_from, wrap_k }
// -----------------------------------------------------------------------------------------
// (a device function so that k_frame_heads, whose warps are such groups, runs the same code; every lane of the calling
// warp must enter it)
// Returns true when the closed form decided (then e_lo / e_hi hold this lane's e_k and e_{k+GROUP}, no wrap).
template <int GROUP>
__device__ __forceinline__ bool mkf_indicator_bounds_group(const long long t, const int k, const bool live,
                                                           const double* __restrict__ u, int N, int K,
                                                           const double* __restrict__ cw_hi,
                                                           const double* __restrict__ cw_lo,
                                                           const double* __restrict__ wprior, double wmax,
                                                           int32_t* __restrict__ bounds, uint32_t* __restrict__ status,
                                                           int clear_status, uint8_t* __restrict__ ind_tail, int& e_lo,
                                                           int& e_hi, const double step_in = 0.0)
{
    // first kernel of a frame update: the track's status word starts from zero (no memset node in front of the
    // chain); the only writer of status in this kernel is this same thread, below
    if (clear_status && live && k == 0) status[t] = 0u;
    const double step = step_in != 0.0 ? step_in : __ddiv_rn(1.0, (double)N); // (step_in: the same quotient, from the host)
    const double beta0 = live ? __dmul_rn(u[t], step) : 0.0;
    const double tol = mkf_resample_tol(N, K, wmax, step);
    bool amb = false;
    int e[2] = {0, 0}; // K <= 2 * GROUP: components k and k + GROUP
    if (live) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int kk = k + h * GROUP;
            if (kk >= K) break;
            dd C;
            C.hi = cw_hi[kk];
            C.lo = cw_lo[kk];
            e[h] = mkf_count_le(C, beta0, step, N, tol, amb);
            if (kk == K - 1 && e[h] < N) amb = true; // thresholds beyond the total prior mass: the loop wraps around
        }
    }
    // combine the ambiguity flags of the GROUP lanes that share a track
    const unsigned lane = threadIdx.x & 31u;
    const unsigned gmask = (GROUP == 32) ? 0xffffffffu : (((1u << GROUP) - 1u) << (lane & ~(unsigned)(GROUP - 1)));
    const unsigned votes = __ballot_sync(0xffffffffu, amb);
    const bool any_amb = (votes & gmask) != 0u;
    e_lo = e[0];
    e_hi = e[1];
    if (!live) return false;
    int32_t* bt = bounds + t * (K + 2);
    if (!any_amb) {
        if (k < K) bt[k] = e[0];
        if (k + GROUP < K) bt[k + GROUP] = e[1];
        if (k == 0) {
            bt[K] = N;
            bt[K + 1] = 0;
        }
        return true;
    }
    if (k != 0) return false;
    // literal loop (src/pf2DRao.cpp:195-207) -> counts per component
    uint32_t st = MKF_ST_IND_FALLBACK;
    for (int q = 0; q < K; q++) bt[q] = 0;
    int idx = 0, wraps = 0, wrap_from = N, wrap_k = 0;
    double beta = beta0;
    double wi = wprior[0];
    for (int i = 0; i < N; i++) {
        while (beta > wi) {
            beta = __dsub_rn(beta, wi);
            idx++;
            if (idx == K) {
                idx = 0;
                wraps++;
            }
            wi = wprior[idx];
        }
        beta = __dadd_rn(beta, step);
        if (wraps == 0) {
            bt[idx] = i + 1; // last output index + 1 holding component idx (made cumulative below)
        } else {
            // past the last component the loop starts over at component 0 and keeps walking: from here on the
            // components are stored per slot (one value when they all agree, which is what a prior that sums to
            // 1 - 3e-14 produces for its last output; a tail row otherwise)
            if (wrap_from == N) {
                wrap_from = i;
                wrap_k = idx;
                st |= MKF_ST_IND_WRAP;
            } else if (idx != wrap_k) {
                wrap_k = -1;
            }
            if (ind_tail) ind_tail[t * N + i] = (uint8_t)idx;
        }
    }
    if (wrap_k < 0 && !ind_tail) wrap_k = 0; // (no tail storage: never the case for a batch)
    int run = 0;
    for (int q = 0; q < K; q++) {
        if (bt[q] > run) run = bt[q];
        bt[q] = run;
    }
    for (int q = 0; q < K; q++)
        if (bt[q] > wrap_from) bt[q] = wrap_from;
    bt[K - 1] = (wrap_from < N) ? wrap_from : N;
    bt[K] = wrap_from;
    bt[K + 1] = wrap_k;
    atomicOr(status + t, st);
    return false;
}

template <int GROUP>
__global__ void k_indicator_bounds(const double* __restrict__ u, long long T, int N, int K,
                                   const double* __restrict__ cw_hi, const double* __restrict__ cw_lo,
                                   const double* __restrict__ wprior, double wmax, int32_t* __restrict__ bounds,
                                   uint32_t* __restrict__ status, int clear_status, uint8_t* __restrict__ ind_tail)
{
    mkf_pdl_launch_dependents();
    mkf_pdl_wait();
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long t = gid / GROUP;
    int e_lo, e_hi;
    mkf_indicator_bounds_group<GROUP>(t, (int)(gid % GROUP), t < T, u, N, K, cw_hi, cw_lo, wprior, wmax, bounds, status,
                                      clear_status, ind_tail, e_lo, e_hi);
}

// the prefix-sum / head-marker pass of k_resample_block carried in double-double throughout; kept out of line so that
// its registers do not weigh on the common path.  Returns this thread's ambiguity flag.
template <int BT, int ITEMS>
__device__ __noinline__ bool mkf_resample_precise_pass(const double* __restrict__ w, int L, int N, int normalise,
                                                       double wsum, double beta0, double step, double tol2,
                                                       int32_t* __restrict__ out, double* sc_d, double* sc_d2,
                                                       const int32_t* __restrict__ rep, int32_t* __restrict__ src,
                                                       const bool by_record)
{
    constexpr bool DIRECT = (BT == 128); // see k_resample_block
    const int tid = threadIdx.x;
    dd carry2 = dd_make(0.0);
    bool amb2 = false;
    for (int base = 0; base < L; base += BT * ITEMS) {
        const int i0 = base + tid * ITEMS;
        dd pre[ITEMS];
        dd run = dd_make(0.0);
#pragma unroll
        for (int q = 0; q < ITEMS; q++) {
            double x = (i0 + q < L) ? (by_record ? w[rep[i0 + q]] : w[i0 + q]) : 0.0;
            if (normalise) x = __ddiv_rn(x, wsum);
            run = dd_add_d(run, x);
            pre[q] = run;
        }
        dd tile_tot;
        const dd excl = mkf_block_excl_scan_dd<BT>(run, sc_d, sc_d2, tile_tot);
        const dd start = dd_add(carry2, excl);
        int e_prev = 0;
        if (i0 > 0 && i0 < L) e_prev = mkf_count_le(start, beta0, step, N, tol2, amb2);
#pragma unroll
        for (int q = 0; q < ITEMS; q++) {
            if (i0 + q < L) {
                const int e = mkf_count_le(dd_add(start, pre[q]), beta0, step, N, tol2, amb2);
                if (e > e_prev) {
                    if (DIRECT) {
                        const int rv = rep ? __ldg(rep + i0 + q) : 0;
                        for (int i = e_prev; i < e; i++) {
                            out[i] = i0 + q;
                            if (rep) src[i] = rv;
                        }
                    } else {
                        out[e_prev] = i0 + q;
                    }
                }
                if (i0 + q == L - 1 && e < N) amb2 = true;
                e_prev = e;
            }
        }
        carry2 = dd_add(carry2, tile_tot);
    }
    return amb2;
}

// -----------------------------------------------------------------------------------------
// per-track weight normalisation + systematic resampling, one CTA per track
//   w_raw : T x L raw weights; out: T x N parents; wsum: T
// -----------------------------------------------------------------------------------------
template <int BT, int ITEMS>
__global__ void __launch_bounds__(BT, 1024 / BT) k_resample_block(const double* __restrict__ w_all, int L, int N,
                                                        const double* __restrict__ u, int u_stride, int normalise,
                                                        double* __restrict__ wsum_out, int32_t* __restrict__ out_all,
                                                        uint32_t* __restrict__ status, int status_stride,
                                                        uint32_t bit_fb, uint32_t bit_deg,
                                                        const uint64_t* __restrict__ seeds, int seed_stride,
                                                        int seed_off, uint32_t* __restrict__ unsorted,
                                                        const int32_t* __restrict__ rep_all,
                                                        int32_t* __restrict__ src_all,
                                                        double* __restrict__ w_slot_out,
                                                        const double* __restrict__ wsum_in)
{
    // wsum_in: the normaliser is given (the per-slot replay of a run-level resample, mkf_runs.cuh) instead of summed here
    // rep_all / src_all (both or neither): besides the parent SLOT of every output, also write the RECORD that holds
    // that slot's state, src = rep[parent] (k_slot_update with dedup stores identical children once).
    // w_slot_out (with rep_all): w_all holds one weight per RECORD (k_slot_update_heads_direct), slot i's weight is
    // w[rep[i]]; pass 1 also writes the per-slot weights out (mkf_batch_download reads them)
    constexpr int CH = 512;
    __shared__ int sc_i[BT / 32];
    __shared__ double sc_d[BT / 32], sc_d2[BT / 32], sc_d3[BT / 32], sc_d4[BT / 32];
    __shared__ double chunk[CH]; // weights staged for the literal loop (rare)
    __shared__ int sh_flag;
    mkf_pdl_launch_dependents();
    mkf_pdl_wait();
    const long long t = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const double* __restrict__ w = w_all + t * L;
    int32_t* __restrict__ out = out_all + t * N;
    const int32_t* __restrict__ rep = rep_all ? rep_all + t * L : nullptr;
    int32_t* __restrict__ src = rep_all ? src_all + t * N : nullptr;
    const bool by_record = w_slot_out != nullptr;
    if (tid == 0 && unsorted) unsorted[t] = 0u;

    // pass 1: sum and NaN-ignoring max (src/pf2DRao.cpp:139,161-172).  The sum only has to be an accurate
    // normaliser (the reference's own sequential sum is no more exact); the literal loop below divides by the
    // same value, so both paths see identical normalised weights.
    // A track that fits one tile (L <= BT * ITEMS) is read once: each thread keeps its ITEMS consecutive weights for
    // pass 2.
    // The sum is accumulated in double-double and rounded once, so that its value does not depend on how the slots
    // are spread over threads -- k_resample_runs (sum over runs of multiplicity x weight) arrives at the same double.
    const bool single = L <= BT * ITEMS;
    double xs[ITEMS];
    dd accd = dd_make(0.0);
    double mx = 0.0, sq = 0.0;
    if (single) {
#pragma unroll
        for (int q = 0; q < ITEMS; q++) {
            const int i = tid * ITEMS + q;
            double x = 0.0;
            if (i < L) {
                if (by_record) {
                    x = w[rep[i]];
                    w_slot_out[t * L + i] = x;
                } else {
                    x = w[i];
                }
                accd = dd_add_d(accd, x);
                sq = fma(x, x, sq);
                if (x > mx) mx = x;
            }
            xs[q] = x;
        }
    } else {
        for (int i = tid; i < L; i += BT) {
            double x;
            if (by_record) {
                x = w[rep[i]];
                w_slot_out[t * L + i] = x;
            } else {
                x = w[i];
            }
            accd = dd_add_d(accd, x);
            sq = fma(x, x, sq);
            if (x > mx) mx = x;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        dd other;
        other.hi = __shfl_xor_sync(0xffffffffu, accd.hi, o);
        other.lo = __shfl_xor_sync(0xffffffffu, accd.lo, o);
        accd = dd_add(accd, other);
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) {
        sc_d[wid] = accd.hi;
        sc_d4[wid] = accd.lo;
        sc_d2[wid] = mx;
        sc_d3[wid] = sq;
    }
    if (tid == 0) sh_flag = 0;
    __syncthreads();
    accd = dd_make(0.0);
    mx = 0.0;
    sq = 0.0;
#pragma unroll
    for (int q = 0; q < BT / 32; q++) {
        dd part;
        part.hi = sc_d[q];
        part.lo = sc_d4[q];
        accd = dd_add(accd, part);
        mx = fmax(mx, sc_d2[q]);
        sq += sc_d3[q];
    }
    __syncthreads();
    const double acc = wsum_in ? wsum_in[t] : accd.hi;
    const double wsum = normalise ? acc : 1.0;
    if (tid == 0 && wsum_out) wsum_out[t] = acc;
    const double wmax_n = normalise ? __ddiv_rn(mx, wsum) : mx;
    if (!(wmax_n > 0.0)) { // max weight 0 / NaN -> random indices from cv::RNG (src/pf2DRao.cpp:184-192)
        if (tid == 0) {
            atomicOr(status + t * status_stride, bit_deg);
            mkf_cvrng rng(seeds ? seeds[t * seed_stride + seed_off] : 1ull);
            (void)rng.uniform_int(0, L); // `int idx = rng.uniform(0, L);` drawn and discarded
            for (int i = 0; i < N; i++) {
                const int idx = rng.uniform_int(0, L);
                out[i] = idx;
                if (rep) src[i] = rep[idx];
            }
            if (unsorted) unsorted[t] = 1u; // random indices are not sorted (matters for the literal alias mode)
        }
        return;
    }
    const double step = __ddiv_rn(1.0, (double)N);
    const double beta0 = __dmul_rn(u[t * u_stride], step);
    // ambiguity band: rounding the sequential loop may have accumulated (mkf_resample_tol) plus the error of the
    // prefix sums below -- at most ITEMS - 1 + 11 (mkf_block_excl_scan_d) <= 14 roundings on the way to any C_k,
    // each <= 2^-53 x (a partial sum <= S): bounded by 16 * 2^-53 * S with S the total mass (1 after normalisation)
    static_assert(ITEMS <= 5, "prefix-sum rounding depth exceeds the 16-ulp tolerance term");
    const double mass = normalise ? 1.0 : acc;
    // sum of squared (normalised) weights, inflated so that it is an upper bound whatever the summation order; a
    // non-finite value drops out of the fmin and leaves the first bound
    const double s2 = (normalise ? __ddiv_rn(sq, __dmul_rn(wsum, wsum)) : sq) * (1.0 + 1e-9);
    const double tol_loop = fmin(mkf_resample_tol(N, L, wmax_n, step), mkf_resample_tol_s2(N, L, s2, mass));
    const double tol = tol_loop + 16.0 * 1.1102230246251565e-16 * (1.0 + mass);

    // Short tracks (the 128-thread variant, N <= 1024): a thread that finds parent k owning the outputs [e_{k-1}, e_k)
    // writes that range on the spot, together with the parent's record index.  Long tracks: the thread drops a head
    // marker at e_{k-1} and a max-scan fills the ranges afterwards (a single parent may own tens of thousands).
    constexpr bool DIRECT = (BT == 128);
    if (!DIRECT) {
        for (int i = tid; i < N; i += BT) out[i] = -1;
        __syncthreads();
    }

    // pass 2: prefix sums of the normalised weights -> child ranges.  Within a tile the scan is plain double;
    // the carry across tiles is kept in double-double so the error does not grow with the number of tiles.
    dd carry = dd_make(0.0);
    bool amb = false;
    for (int base = 0; base < L; base += BT * ITEMS) {
        const int i0 = base + tid * ITEMS;
        double pre[ITEMS];
        double run = 0.0;
#pragma unroll
        for (int q = 0; q < ITEMS; q++) {
            double x = single ? xs[q] : ((i0 + q < L) ? (by_record ? w[rep[i0 + q]] : w[i0 + q]) : 0.0);
            if (normalise) x = __ddiv_rn(x, wsum);
            run += x;
            pre[q] = run;
        }
        double tile_tot;
        const double excl = mkf_block_excl_scan_d<BT>(run, sc_d, tile_tot);
        const dd start_b = dd_add_d(dd_add_d(carry, excl), -beta0); // prefix sum before my first weight, minus beta0
        int e_prev = 0; // e_{-1} = 0 by definition: no output precedes the first weight
        if (i0 > 0 && i0 < L) e_prev = mkf_count_le_df(start_b, step, N, tol, amb);
#pragma unroll
        for (int q = 0; q < ITEMS; q++) {
            if (i0 + q < L) {
                const int e = mkf_count_le_df(dd_add_d(start_b, pre[q]), step, N, tol, amb);
                if (e > e_prev) {
                    if (DIRECT) {
                        const int rv = rep ? __ldg(rep + i0 + q) : 0;
                        for (int i = e_prev; i < e; i++) {
                            out[i] = i0 + q;
                            if (rep) src[i] = rv;
                        }
                    } else {
                        out[e_prev] = i0 + q;
                    }
                }
                if (i0 + q == L - 1 && e < N) amb = true; // literal loop would wrap past the last weight
                e_prev = e;
            }
        }
        carry = dd_add_d(carry, tile_tot);
    }
    if (amb) sh_flag = 1;
    __syncthreads();
    if (sh_flag) {
        // second opinion before giving the track to the literal loop: the same pass with the prefix sums carried in
        // double-double throughout (error ~1e-30), so only the loop's own rounding bound (plus the rounding of
        // mkf_count_le's remainder, <= 2 ulp of step) remains in the band -- 8x narrower at N = 65 536, where one
        // literal re-run costs milliseconds.
        __syncthreads();
        if (tid == 0) sh_flag = 0;
        if (!DIRECT)
            for (int i = tid; i < N; i += BT) out[i] = -1;
        __syncthreads();
        const double tol2 = tol_loop + 8.0 * 1.1102230246251565e-16 * step + 8.0e-28 * (1.0 + mass);
        const bool amb2 = mkf_resample_precise_pass<BT, ITEMS>(w, L, N, normalise, wsum, beta0, step, tol2, out, sc_d,
                                                               sc_d2, rep, src, by_record);
        if (amb2) sh_flag = 1;
        __syncthreads();
        if (sh_flag) {
            // still undecidable in closed form: the first warp runs the reference's loop itself on the same
            // normalised weights (staged through shared memory, lane 0 walking them)
            if (tid == 0) atomicOr(status + t * status_stride, bit_fb);
            if (wid == 0) {
                auto wf = [&](int i) {
                    const double x = by_record ? w[rep[i]] : w[i];
                    return normalise ? __ddiv_rn(x, wsum) : x;
                };
                mkf_resample_sequential_warp<CH>(wf, L, N, u[t * u_stride],
                                                 [&](int i, int idx) {
                                                     out[i] = idx;
                                                     if (rep) src[i] = rep[idx];
                                                 },
                                                 chunk);
            }
            return;
        }
    }
    if (DIRECT) return;
    // fill: inclusive max-scan of the head markers
    int carry_max = -1;
    for (int base = 0; base < N; base += BT * ITEMS) {
        const int i0 = base + tid * ITEMS;
        int loc[ITEMS];
        int run = -1;
#pragma unroll
        for (int q = 0; q < ITEMS; q++) {
            const int h = (i0 + q < N) ? out[i0 + q] : -1;
            run = max(run, h);
            loc[q] = run;
        }
        int tile_max;
        const int excl = mkf_block_excl_scan_max<BT>(run, sc_i, tile_max);
        const int pre = max(carry_max, excl);
#pragma unroll
        for (int q = 0; q < ITEMS; q++)
            if (i0 + q < N) {
                const int idx = max(pre, loc[q]);
                out[i0 + q] = idx;
                if (rep) src[i0 + q] = __ldg(rep + idx);
            }
        carry_max = max(carry_max, tile_max);
    }
}

// Few weights, many outputs (the C -> N candidate resample of the association step with C <= 32 candidates per hand,
// src/pfPose.cpp:300-301): one WARP per track, lane k owns weight k.  Same closed form, same ambiguity band and
// the same normaliser semantics as k_resample_block (whose CTA of 128 threads would idle on 17 weights: 61 us for
// 8192 tracks against 9 us here); an ambiguous track is handed to the literal loop at once (lane 0, weights in shared
// memory).  max weight == 0 / NaN -> cv::RNG indices, as everywhere.
__global__ void __launch_bounds__(128) k_resample_warp(const double* __restrict__ w_all, long long T, int L, int N,
                                                        const double* __restrict__ u, int u_stride, int normalise,
                                                        double* __restrict__ wsum_out, int32_t* __restrict__ out_all,
                                                        uint32_t* __restrict__ status, int status_stride,
                                                        uint32_t bit_fb, uint32_t bit_deg,
                                                        const uint64_t* __restrict__ seeds, int seed_stride,
                                                        int seed_off, uint32_t* __restrict__ unsorted,
                                                        int32_t* __restrict__ cut_out)
{
    // cut_out (T x 32, optional): e_k = number of outputs whose index is <= k, for k < L -- the outputs as L sorted runs
    // (what k_frame_heads cuts the particle runs with); cut_out[t * 32] = -1 when the indices came from the literal
    // loop or from cv::RNG and have to be read per output
    __shared__ double xn_s[4][32];
    mkf_pdl_launch_dependents();
    mkf_pdl_wait();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long t = (long long)blockIdx.x * 4 + wid;
    if (t >= T) return;
    int32_t* __restrict__ out = out_all + t * N;
    const double x = (lane < L) ? w_all[t * L + lane] : 0.0;
    double acc = x, mx = (x > 0.0) ? x : 0.0, sq = x * x; // NaN-ignoring max starting at 0 (src/pf2DRao.cpp:161-172)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    const double wsum = normalise ? acc : 1.0;
    if (lane == 0 && wsum_out) wsum_out[t] = acc;
    if (lane == 0 && unsorted) unsorted[t] = 0u;
    const double wmax_n = normalise ? __ddiv_rn(mx, wsum) : mx;
    if (!(wmax_n > 0.0)) { // src/pf2DRao.cpp:184-192
        if (lane == 0) {
            atomicOr(status + t * status_stride, bit_deg);
            mkf_cvrng rng(seeds ? seeds[t * seed_stride + seed_off] : 1ull);
            (void)rng.uniform_int(0, L); // `int idx = rng.uniform(0, L);` drawn and discarded
            for (int i = 0; i < N; i++) out[i] = rng.uniform_int(0, L);
            if (unsorted) unsorted[t] = 1u;
            if (cut_out) cut_out[t * 32] = -1;
        }
        return;
    }
    const double step = __ddiv_rn(1.0, (double)N);
    const double beta0 = __dmul_rn(u[t * u_stride], step);
    const double mass = normalise ? 1.0 : acc;
    const double s2 = (normalise ? __ddiv_rn(sq, __dmul_rn(wsum, wsum)) : sq) * (1.0 + 1e-9);
    const double tol_loop = fmin(mkf_resample_tol(N, L, wmax_n, step), mkf_resample_tol_s2(N, L, s2, mass));
    // the inclusive scan below reaches any prefix sum through <= 5 roundings (<= 14 is what this term allows)
    const double tol = tol_loop + 16.0 * 1.1102230246251565e-16 * (1.0 + mass);
    const double xn = (lane < L) ? (normalise ? __ddiv_rn(x, wsum) : x) : 0.0;
    double incl = xn;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    bool amb = false;
    int e = N;
    if (lane < L) {
        e = mkf_count_le_df(dd_add_d(dd_make(incl), -beta0), step, N, tol, amb);
        if (lane == L - 1 && e < N) amb = true; // the literal loop would wrap past the last weight
    }
    int e_prev = __shfl_up_sync(0xffffffffu, e, 1);
    if (lane == 0) e_prev = 0;
    if (__any_sync(0xffffffffu, amb)) {
        xn_s[wid][lane] = xn;
        __syncwarp();
        if (lane == 0) {
            atomicOr(status + t * status_stride, bit_fb);
            const double* xs = xn_s[wid];
            mkf_resample_sequential([&](int i) { return xs[i]; }, L, N, u[t * u_stride],
                                    [&](int i, int idx) { out[i] = idx; });
            if (cut_out) cut_out[t * 32] = -1;
        }
        return;
    }
    if (cut_out && lane < L) cut_out[t * 32 + lane] = e;
    for (int k = 0; k < L; k++) {
        const int lo = __shfl_sync(0xffffffffu, e_prev, k), hi = __shfl_sync(0xffffffffu, e, k);
        for (int i = lo + lane; i < hi; i += 32) out[i] = k;
    }
}

// one thread per track, literal sequential semantics throughout (sequential wsum as the
// reference, src/pf2DRao.cpp:139; then the loop of :195-207).  Used when N and L are small.
// The CTA's 128 weight rows are staged transposed in shared memory (coalesced loads, one division per weight instead
// of one per visit) and the indices leave through shared memory as well (coalesced stores).  LD = 129 keeps both the
// transposing accesses and the per-thread walks free of bank conflicts.
__global__ void __launch_bounds__(128) k_resample_small(const double* __restrict__ w_all, long long T, int L, int N,
                                                         const double* __restrict__ u, int u_stride, int normalise,
                                                         double* __restrict__ wsum_out, int32_t* __restrict__ out_all,
                                                         uint32_t* __restrict__ status, int status_stride,
                                                         uint32_t bit_deg, const uint64_t* __restrict__ seeds,
                                                         int seed_stride, int seed_off, uint32_t* __restrict__ unsorted,
                                                         const int32_t* __restrict__ rep_all,
                                                         int32_t* __restrict__ src_all)
{
    constexpr int LD = 129;
    extern __shared__ double sm_w[];                                       // [L][LD]
    int32_t* sm_out = reinterpret_cast<int32_t*>(sm_w + (size_t)L * LD);   // [N][LD]
    mkf_pdl_launch_dependents();
    mkf_pdl_wait();
    const int tid = threadIdx.x;
    const long long t0 = (long long)blockIdx.x * 128;
    const int nt = (int)((T - t0) < 128 ? (T - t0) : 128);
    for (int i = tid; i < nt * L; i += 128) {
        const int tr = i / L, k = i - tr * L;
        sm_w[k * LD + tr] = w_all[t0 * L + i];
    }
    __syncthreads();
    if (tid < nt) {
        const long long t = t0 + tid;
        double* __restrict__ w = sm_w + tid;
        double wsum = 0.0;
        for (int i = 0; i < L; i++) wsum = __dadd_rn(wsum, w[i * LD]);
        if (wsum_out) wsum_out[t] = wsum;
        double mw = 0.0;
        for (int i = 0; i < L; i++) {
            const double x = normalise ? __ddiv_rn(w[i * LD], wsum) : w[i * LD];
            w[i * LD] = x;
            if (x > mw) mw = x;
        }
        int32_t* o = sm_out + tid;
        if (!(mw > 0.0)) { // max weight 0 / NaN -> random indices from cv::RNG (src/pf2DRao.cpp:184-192)
            atomicOr(status + t * status_stride, bit_deg);
            mkf_cvrng rng(seeds ? seeds[t * seed_stride + seed_off] : 1ull);
            (void)rng.uniform_int(0, L); // `int idx = rng.uniform(0, L);` drawn and discarded
            for (int i = 0; i < N; i++) o[i * LD] = rng.uniform_int(0, L);
        } else {
            mkf_resample_sequential([&](int i) { return w[i * LD]; }, L, N, u[t * u_stride],
                                    [&](int i, int idx) { o[i * LD] = idx; });
        }
        if (unsorted) unsorted[t] = (mw > 0.0) ? 0u : 1u; // random indices are not sorted (literal alias mode)
    }
    __syncthreads();
    for (int i = tid; i < nt * N; i += 128) {
        const int tr = i / N, k = i - tr * N;
        const int idx = sm_out[k * LD + tr];
        out_all[t0 * N + i] = idx;
        if (rep_all) src_all[t0 * N + i] = __ldg(rep_all + (t0 + tr) * L + idx);
    }
}

// -----------------------------------------------------------------------------------------
// getEstimator + reconstruction: one CTA per track
// -----------------------------------------------------------------------------------------
template <int D, int BT>
__global__ void __launch_bounds__(BT, 1024 / BT) k_estimate(const double2* __restrict__ st, const int32_t* __restrict__ parent,
                                                  int N, int Dpose, const double* __restrict__ recon,
                                                  const double* __restrict__ pmean, const double* __restrict__ tinv,
                                                  double* __restrict__ xbar_out, double* __restrict__ pose_out,
                                                  double* __restrict__ pose_out2)
{
    using L = SlotLay<D>;
    __shared__ double red[BT / 32][D];
    __shared__ double xb[D];
    mkf_pdl_launch_dependents();
    mkf_pdl_wait();
    const long long t = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    double acc[D];
#pragma unroll
    for (int e = 0; e < D; e++) acc[e] = 0.0;
    // a thread takes RUN consecutive children: their record indices are loaded together (one latency), sorted parents
    // make neighbours share a record more often than not (then the gather is skipped and the registers reused), and
    // the gathers of one thread are independent of each other
    constexpr int RUN = 4;
    for (int j0 = tid * RUN; j0 < N; j0 += BT * RUN) {
        int idx[RUN];
#pragma unroll
        for (int u = 0; u < RUN; u++) idx[u] = (j0 + u < N) ? __ldg(parent + t * N + j0 + u) : -1;
        double2 v[D / 2];
        int have = -2;
#pragma unroll
        for (int u = 0; u < RUN; u++) {
            if (idx[u] < 0) continue;
            if (idx[u] != have) {
                const long long sp = t * N + idx[u];
                const double2* __restrict__ src = st + (sp >> 5) * (long long)L::TILE2 + (sp & 31) * L::H;
#pragma unroll
                for (int p = 0; p < D / 2; p++) v[p] = __ldg(src + L::po(p));
                have = idx[u];
            }
#pragma unroll
            for (int p = 0; p < D / 2; p++) {
                acc[2 * p] += v[p].x;
                acc[2 * p + 1] += v[p].y;
            }
        }
    }
    // warp reduction that halves the vector at every step: at offset o the lanes with bit o clear keep the lower half
    // of their n partial sums and receive the partner's lower half, the others keep the upper half -- 6+3+2+1+1 = 13
    // shuffles for D = 12 instead of 5 per element (60).  A lane ends with the warp total of ONE element in acc[0].
    int e_mine = 0;   // which element my acc[0] ends up holding
    bool valid = true; // false: my acc[0] is padding
    {
        int n = D;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const int h = (n + 1) / 2;
            const bool upper = (lane & o) != 0;
#pragma unroll
            for (int e = 0; e < h; e++) {
                const double lo = acc[e];
                const double hi = (e + h < n) ? acc[e + h] : 0.0;
                const double got = __shfl_xor_sync(0xffffffffu, upper ? lo : hi, o);
                acc[e] = (upper ? hi : lo) + got;
            }
            n = h;
        }
        // element index: walk the steps backwards; taking the upper half at a step with (n, h) maps local e to e + h
        int ns[5], nn = D;
#pragma unroll
        for (int q = 0; q < 5; q++) {
            ns[q] = nn;
            nn = (nn + 1) / 2;
        }
#pragma unroll
        for (int q = 4; q >= 0; q--) {
            const int o = 16 >> q, hq = (ns[q] + 1) / 2;
            if (lane & o) {
                if (e_mine + hq >= ns[q]) valid = false;
                e_mine += hq;
            }
        }
    }
    if (valid) red[wid][e_mine] = acc[0];
    __syncthreads();
    if (tid < D) {
        double s = 0.0;
        for (int q = 0; q < BT / 32; q++) s += red[q][tid];
        xb[tid] = s * (1.0 / (double)N);
    }
    __syncthreads();
    if ((pose_out || pose_out2) && tid < Dpose) { // pose_out2: the batch's copy for the next association step
        double s = 0.0;
        for (int c = 0; c < D; c++) s = fma(recon[tid * D + c], xb[c], s);
        if (pose_out) pose_out[t * Dpose + tid] = s + pmean[tid];
        if (pose_out2) pose_out2[t * Dpose + tid] = s + pmean[tid];
    }
    if (xbar_out && tid < D) {
        double s = 0.0;
        for (int c = 0; c < D; c++) s = fma(tinv[tid * D + c], xb[c], s);
        xbar_out[t * D + tid] = s;
    }
}

// getEstimator + reconstruction for small N: GROUP (16 or 32) lanes per track, 128 / GROUP tracks per CTA and trip,
// TRIPS trips per CTA.  The reconstruction matrices are staged once per CTA (transposed, so that lane r reading
// row r is conflict-free); within a trip there is no block barrier: the butterfly reduction leaves the mean in every
// lane of the group, which then computes its own output rows.
template <int D, int GROUP, int TRIPS>
__global__ void __launch_bounds__(128) k_estimate_small(const double2* __restrict__ st, const int32_t* __restrict__ parent,
                                                         long long T, int N, int Dpose, const double* __restrict__ recon,
                                                         const double* __restrict__ pmean, const double* __restrict__ tinv,
                                                         double* __restrict__ xbar_out, double* __restrict__ pose_out,
                                                         double* __restrict__ pose_out2)
{
    using L = SlotLay<D>;
    constexpr int TPB = 128 / GROUP; // tracks per CTA and trip
    extern __shared__ double coef[]; // [c][r], r < Dpose + D: rows of recon (pose) then rows of tinv (xbar)
    mkf_pdl_launch_dependents();
    const int R = Dpose + D;
    for (int i = threadIdx.x; i < R * D; i += 128) { // model constants: safe before the dependency wait
        const int r = i / D, c = i - r * D;
        coef[c * R + r] = r < Dpose ? recon[r * D + c] : tinv[(r - Dpose) * D + c];
    }
    mkf_pdl_wait();
    __syncthreads();
    const int g = threadIdx.x / GROUP, l = threadIdx.x % GROUP;
    const double inv_n = 1.0 / (double)N;
    for (int trip = 0; trip < TRIPS; trip++) {
        const long long t = ((long long)blockIdx.x * TRIPS + trip) * TPB + g;
        if (t >= T) break; // uniform within the group; groups only meet at warp shuffles of their own lanes
        double acc[D];
#pragma unroll
        for (int e = 0; e < D; e++) acc[e] = 0.0;
        for (int j = l; j < N; j += GROUP) {
            const long long sp = t * N + __ldg(parent + t * N + j);
            const double2* __restrict__ src = st + (sp >> 5) * (long long)L::TILE2 + (sp & 31) * L::H;
#pragma unroll
            for (int p = 0; p < D / 2; p++) {
                const double2 q = __ldg(src + L::po(p));
                acc[2 * p] += q.x;
                acc[2 * p + 1] += q.y;
            }
        }
        // lanes of one group are contiguous in a warp and GROUP divides 32: xor offsets < GROUP stay inside it
        const unsigned gmask = (GROUP == 32) ? 0xffffffffu : (((1u << GROUP) - 1u) << ((threadIdx.x & 31) / GROUP * GROUP));
#pragma unroll
        for (int e = 0; e < D; e++) {
#pragma unroll
            for (int o = GROUP / 2; o > 0; o >>= 1) acc[e] += __shfl_xor_sync(gmask, acc[e], o);
            acc[e] *= inv_n;
        }
        for (int r = l; r < R; r += GROUP) {
            double sacc = 0.0;
#pragma unroll
            for (int c = 0; c < D; c++) sacc = fma(coef[c * R + r], acc[c], sacc);
            if (r < Dpose) {
                if (pose_out) pose_out[t * Dpose + r] = sacc + pmean[r];
                if (pose_out2) pose_out2[t * Dpose + r] = sacc + pmean[r];
            } else if (xbar_out) {
                xbar_out[t * D + (r - Dpose)] = sacc;
            }
        }
    }
}

// -----------------------------------------------------------------------------------------
// state I/O
// -----------------------------------------------------------------------------------------
// resetTracker (src/my_gmm.cpp:30-42): slot j <- (mu_k, Sigma_k) of its drawn component
template <int D>
__global__ void k_reset(double2* __restrict__ st, int32_t* __restrict__ parent, const int32_t* __restrict__ bounds,
                        const double* __restrict__ init_const, long long total, int N, int K,
                        const uint8_t* __restrict__ ind_tail)
{
    using L = SlotLay<D>;
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= total) return;
    const long long t = s / N;
    const int j = (int)(s - t * N);
    const int k = mkf_component_of(bounds + t * (K + 2), K, j, ind_tail ? ind_tail + t * N : nullptr);
    const double* __restrict__ ic = init_const + (long long)k * L::NE;
    double2* __restrict__ dst = st + (s >> 5) * (long long)L::TILE2 + (s & 31) * L::H;
    for (int p = 0; p < L::NP; p++) {
        double2 q;
        q.x = ic[2 * p];
        q.y = (2 * p + 1 < L::NE) ? ic[2 * p + 1] : 0.0;
        dst[L::po(p)] = q;
    }
    parent[s] = j;
}

template <int D>
__device__ __forceinline__ int packed_index_dev(int r, int c)
{
    using L = SlotLay<D>;
    if (r < c) {
        int tmp = r;
        r = c;
        c = tmp;
    }
    if (r < MKF_M) return tri(r, c);
    if (c < MKF_M) return L::NA + (r - MKF_M) * MKF_M + c;
    return L::NA + L::NB + tri(r - MKF_M, c - MKF_M);
}

// reference coordinates -> device layout:  x' = T x,  P' = T sym(P) T^T
template <int D>
__global__ void k_upload(double2* __restrict__ st, int32_t* __restrict__ parent, const double* __restrict__ x,
                         const double* __restrict__ P, const double* __restrict__ Tm, long long total, int N)
{
    using L = SlotLay<D>;
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= total) return;
    double v[L::NP * 2];
    v[L::NP * 2 - 1] = 0.0;
    const double* xs = x + s * D;
    const double* Ps = P + s * D * D;
    for (int r = 0; r < D; r++) {
        double acc = 0.0;
        for (int c = 0; c < D; c++) acc = fma(Tm[r * D + c], xs[c], acc);
        v[r] = acc;
    }
    double TP[D * D];
    for (int r = 0; r < D; r++)
        for (int c = 0; c < D; c++) {
            double acc = 0.0;
            for (int k = 0; k < D; k++) acc = fma(Tm[r * D + k], 0.5 * (Ps[k * D + c] + Ps[c * D + k]), acc);
            TP[r * D + c] = acc;
        }
    for (int r = 0; r < D; r++)
        for (int c = 0; c <= r; c++) {
            double a1 = 0.0, a2 = 0.0;
            for (int k = 0; k < D; k++) {
                a1 = fma(TP[r * D + k], Tm[c * D + k], a1);
                a2 = fma(TP[c * D + k], Tm[r * D + k], a2);
            }
            v[D + packed_index_dev<D>(r, c)] = 0.5 * (a1 + a2);
        }
    double2* __restrict__ dst = st + (s >> 5) * (long long)L::TILE2 + (s & 31) * L::H;
    for (int p = 0; p < L::NP; p++) {
        double2 q;
        q.x = v[2 * p];
        q.y = v[2 * p + 1];
        dst[L::po(p)] = q;
    }
    parent[s] = (int)(s % N);
}

// device layout (through the parent gather) -> reference coordinates
template <int D>
__global__ void k_download(const double2* __restrict__ st, const int32_t* __restrict__ parent,
                           const double* __restrict__ Ti, double* __restrict__ x, double* __restrict__ P,
                           long long total, int N)
{
    using L = SlotLay<D>;
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= total) return;
    const long long t = s / N;
    const long long sp = t * N + parent[s];
    const double2* __restrict__ src = st + (sp >> 5) * (long long)L::TILE2 + (sp & 31) * L::H;
    double v[L::NP * 2];
    for (int p = 0; p < L::NP; p++) {
        const double2 q = src[L::po(p)];
        v[2 * p] = q.x;
        v[2 * p + 1] = q.y;
    }
    if (x) {
        for (int r = 0; r < D; r++) {
            double acc = 0.0;
            for (int c = 0; c < D; c++) acc = fma(Ti[r * D + c], v[c], acc);
            x[s * D + r] = acc;
        }
    }
    if (P) {
        double TP[D * D];
        for (int r = 0; r < D; r++)
            for (int c = 0; c < D; c++) {
                double acc = 0.0;
                for (int k = 0; k < D; k++) acc = fma(Ti[r * D + k], v[D + packed_index_dev<D>(k, c)], acc);
                TP[r * D + c] = acc;
            }
        for (int r = 0; r < D; r++)
            for (int c = 0; c < D; c++) {
                double acc = 0.0;
                for (int k = 0; k < D; k++) acc = fma(TP[r * D + k], Ti[c * D + k], acc);
                P[(s * D + r) * D + c] = acc;
            }
    }
}

// w_norm = w_raw / wsum (src/pf2DRao.cpp:145-148) and the per-slot component indicators
__global__ void k_aux_outputs(const double* __restrict__ w_raw, const double* __restrict__ wsum,
                              const int32_t* __restrict__ bounds, long long total, int N, int K,
                              double* __restrict__ w_norm, int32_t* __restrict__ indicators,
                              const uint8_t* __restrict__ ind_tail)
{
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= total) return;
    const long long t = s / N;
    if (w_norm) w_norm[s] = __ddiv_rn(w_raw[s], wsum[t]);
    if (indicators)
        indicators[s] = mkf_component_of(bounds + t * (K + 2), K, (int)(s - t * N), ind_tail ? ind_tail + t * N : nullptr);
}

// synthetic workload of include/mkf_synth.h generated in place (bench / large-batch tests)
__global__ void k_synth_fill(uint64_t seed, long long track0, uint64_t frame, int jitter, int meas_layout, long long T,
                             int N, double* __restrict__ meas, double* __restrict__ u_ind, double* __restrict__ u_post)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (meas_layout == MKF_MEAS_SHARED) {
        if (i >= T) return;
        double z[6];
        mkf_synth_meas(seed, (uint64_t)(track0 + i), frame, -1, jitter, z);
        for (int r = 0; r < 6; r++) meas[i * 6 + r] = z[r];
        if (u_ind) u_ind[i] = mkf_synth_u(seed, (uint64_t)(track0 + i), frame, MKF_SYNTH_LANE_U_IND);
        if (u_post) u_post[i] = mkf_synth_u(seed, (uint64_t)(track0 + i), frame, MKF_SYNTH_LANE_U_POST);
    } else {
        if (i >= T * N) return;
        const long long t = i / N;
        const int j = (int)(i - t * N);
        double z[6];
        mkf_synth_meas(seed, (uint64_t)(track0 + t), frame, j, jitter, z);
        for (int r = 0; r < 6; r++) meas[(t * 6 + r) * N + j] = z[r];
        if (j == 0) {
            if (u_ind) u_ind[t] = mkf_synth_u(seed, (uint64_t)(track0 + t), frame, MKF_SYNTH_LANE_U_IND);
            if (u_post) u_post[t] = mkf_synth_u(seed, (uint64_t)(track0 + t), frame, MKF_SYNTH_LANE_U_POST);
        }
    }
}

#endif
