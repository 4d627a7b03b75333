// mkf_frame_small.cuh -- the whole frame of a SHORT track (9 <= N <= 16 slots: the GMM-KF bank with one slot per
// component, BASELINE configs 5 and "bank mode") in one launch:
//     component indicators  (K -> N systematic resample of the prior weights,           src/pf2DRao.cpp:128)
//     per-slot predict + innovation + Cholesky + mvnpdf + update                        (src/pf2DRao.cpp:134-142)
//     weight sum, normalisation, N -> N systematic resample / cv::RNG fallback          (src/pf2DRao.cpp:139-210)
//     getEstimator + PCA reconstruction                                                 (src/pf2DRao.cpp:23-31, src/pfPose.cpp:347-348)
//
// Why: at N = 15 the per-slot frame is five launches -- k_indicator_bounds, k_slot_update, k_slot_update_repair,
// k_resample_small, k_estimate_small -- and a batch of a few thousand tracks is launch-bound (4096 x 15: 49.5 us per
// frame, of which the slot kernel needs 9).  Here a track belongs to HALF A WARP (16 lanes, lane j = slot
// j): the lanes draw the track's component indicators themselves (the code of k_indicator_bounds<16>) while their
// parents' records are in flight, run the same slot_math as every other slot kernel, and then -- the weights and means
// still in registers -- sum and normalise the weights in the reference's sequential order, resample in closed form
// (lane j owns weight j; the reference's loop itself when a threshold is too close to call), sum the children's means
// in the order of k_estimate_small's butterfly (bit-identical estimates) and reconstruct the pose.  Everything a lane
// needs from its 15 neighbours travels through a few hundred bytes of shared memory private to the half warp: the first
// version exchanged it by shuffles and a warp then executed 2 850 instructions where k_slot_update executes 1 580; this
// one executes 2 390 (ncu).  Measured: 4096 x 15 in 35.2 us per frame (49.5 with the five launches); 1 M x 15 in 4.99 ms
// against 4.51 -- at two warps per scheduler the extra instructions are serial latency the slot kernel does not have,
// so mkf_batch_create picks this kernel for batches of <= 16 384 tracks only (MKF_SMALL_FUSED=1 forces it).  The state stays in the dense tile layout: warp w owns slots [2 N w, 2 N (w + 1)), 480 bytes per
// record piece at N = 15, sector-aligned, so the DRAM traffic is that of k_slot_update.
//
// A track with a flagged cv::Cholesky failure (never seen in practice) skips the tail; k_slot_update_repair redoes its
// slots with the literal failure semantics and then runs mkf_small_tail_serial for it.
#ifndef MKF_FRAME_SMALL_CUH
#define MKF_FRAME_SMALL_CUH

#include "mkf_kernels.cuh"

template <int D>
__global__ void __launch_bounds__(128, 2) k_frame_small(const SlotArgs a, const SmallTailArgs s)
{
    using L = SlotLay<D>;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ __align__(16) double sm_w[4][2][16];     // raw weights of the warp's two tracks (0 behind slot N - 1)
    __shared__ __align__(16) int sm_par[4][2][16];      // parent of every output (j behind N - 1: a row of zeros)
    __shared__ __align__(16) int sm_cmp[4][2][16];      // component of every slot
    __shared__ __align__(16) double sm_x[4][2][16][D];  // the children's updated means
    __shared__ __align__(16) double sm_mean[4][2][D];
    const int R = s.Dpose + D;
    double* cst = reinterpret_cast<double*>(smem_raw);
    const double* coef = cst + (size_t)a.K * L::CS; // [c][r]: rows of recon (pose), then rows of tinv

    if (threadIdx.x == 0) mkf_mbar_init(&mbar, 1);
    __syncthreads();
    if (threadIdx.x == 0) { // model constants (never written by the frame chain): safe before the dependency wait
        mkf_mbar_expect_tx(&mbar, s.small_const_bytes);
        mkf_tma_load_1d(cst, s.small_const, s.small_const_bytes, &mbar);
    }
    mkf_pdl_launch_dependents();
    mkf_pdl_wait();
    if (a.ts && threadIdx.x == 0) atomicMin(a.ts, mkf_globaltimer());

    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int g = lane >> 4, j = lane & 15;
    const int N = a.N, K = a.K;
    const long long T = a.total / N;
    const long long t = ((long long)blockIdx.x * 4 + wid) * 2 + g;
    const bool live_t = t < T;           // the half warp has a track
    const bool live = live_t && j < N;   // the lane has a slot
    const long long s_ = live_t ? t * N + j : 0;
    const unsigned gmask = 0xffffu << (g * 16);
    const int gbase = g * 16;
    const double step = s.step;

    // the parent's record and the measurement column: in flight during the indicator draw
    double v[L::NE];
    double zc[MKF_M];
    if (live) {
        const int par = __ldg(a.src + s_);
        const long long sp = t * N + par;
        const double2* __restrict__ src = a.st_in + (sp >> 5) * (long long)L::TILE2 + (sp & 31) * L::H;
#pragma unroll
        for (int p = 0; p < L::NP; p++) {
            const double2 q = __ldg(src + L::po(p));
            v[2 * p] = q.x;
            if (2 * p + 1 < L::NE) v[2 * p + 1] = q.y;
        }
        mkf_load_meas(a, t, j, zc);
    }

    // component indicators: the 16 lanes are the group of k_indicator_bounds<16> (lane q evaluates e_q and e_{q+16});
    // component q then marks its slots [e_{q-1}, e_q) in shared memory
    int k = 0;
    {
        int e_lo, e_hi;
        const bool closed = mkf_indicator_bounds_group<16>(live_t ? t : 0, j, live_t, s.u_ind, N, K, s.cw_hi, s.cw_lo,
                                                           s.wprior, s.wmax, s.bounds_out, a.status, s.clear_status,
                                                           s.ind_tail_out, e_lo, e_hi, step);
        const bool fast = ((__ballot_sync(FULL, closed || !live_t) & gmask) == gmask);
        int p_lo = __shfl_up_sync(FULL, e_lo, 1, 16), p_hi = __shfl_up_sync(FULL, e_hi, 1, 16);
        const int e15 = __shfl_sync(FULL, e_lo, gbase + 15);
        if (j == 0) {
            p_lo = 0;
            p_hi = e15;
        }
        int* cmp = sm_cmp[wid][g];
        if (fast && live_t) { // (closed form: e is non-decreasing and e_{K-1} = N)
            if (j < K)
                for (int i = p_lo; i < min(e_lo, N); i++) cmp[i] = j;
            if (j + 16 < K)
                for (int i = p_hi; i < min(e_hi, N); i++) cmp[i] = j + 16;
        }
        __syncwarp(); // the marks; the status word's reset; (slow path) the boundaries lane 0 wrote to global memory
        if (live) {
            if (fast) {
                k = cmp[j];
            } else { // the literal loop wrote the boundaries (and maybe a per-slot tail)
                const int32_t* bt = s.bounds_out + t * (K + 2);
                for (int q = 0; q < K - 1; q++) k += (j >= __ldcg(bt + q)) ? 1 : 0;
                if (j >= __ldcg(bt + K)) {
                    const int wk = __ldcg(bt + K + 1);
                    k = (wk >= 0 || !s.ind_tail_out) ? max(wk, 0) : (int)__ldcg(s.ind_tail_out + t * N + j);
                }
            }
        }
    }
    mkf_mbar_wait(&mbar, 0);

    double w = 0.0;
    bool ok = true;
    if (live) {
        ok = slot_math<D, false>(v, cst + k * L::CS, zc, a.r, a.chol_mode, a.stage, w);
        double2* __restrict__ dst = a.st_out + (s_ >> 5) * (long long)L::TILE2 + (s_ & 31) * L::H;
#pragma unroll
        for (int p = 0; p < L::NP; p++) {
            double2 q;
            q.x = v[2 * p];
            q.y = (2 * p + 1 < L::NE) ? v[2 * p + 1] : 0.0;
            __stcs(dst + L::po(p), q);
        }
        a.w_raw[s_] = w;
    }
    // the tail's inputs: raw weight and updated mean of every slot (zeros behind slot N - 1)
    sm_w[wid][g][j] = w;
    {
        double2* xr = reinterpret_cast<double2*>(sm_x[wid][g][j]);
#pragma unroll
        for (int p = 0; p < D / 2; p++) xr[p] = live ? make_double2(v[2 * p], v[2 * p + 1]) : make_double2(0.0, 0.0);
    }
    const unsigned fail_votes = __ballot_sync(FULL, !ok);
    const bool failed = (fail_votes & gmask) != 0u;
    if (failed && live && j == 0) atomicOr(a.status + t, MKF_ST_CHOL_FAIL);
    const bool tail = live_t && !failed;
    __syncwarp();

    // weight sum in the reference's order (src/pf2DRao.cpp:139), every lane of the group (w + 0.0 = w behind slot N - 1)
    double wsum = 0.0;
    {
        const double2* wr = reinterpret_cast<const double2*>(sm_w[wid][g]);
#pragma unroll
        for (int p = 0; p < 8; p++) {
            const double2 q = wr[p];
            wsum = __dadd_rn(__dadd_rn(wsum, q.x), q.y);
        }
    }
    const double x = live ? __ddiv_rn(w, wsum) : 0.0;
    double mw = (x > 0.0) ? x : 0.0; // (NaN compares false, as in `if (x > mw) mw = x`)
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
        const double y = __shfl_xor_sync(FULL, mw, o);
        mw = (y > mw) ? y : mw;
    }
    // N -> N systematic resample (src/pf2DRao.cpp:195-207) in closed form, as k_resample_warp: lane j owns weight j, e_j =
    // #{outputs whose threshold is <= the prefix sum C_j}, and marks its children [e_{j-1}, e_j); a threshold within
    // the rounding band of a prefix sum hands the track to the reference's loop below (as does max weight == 0 / NaN,
    // for the cv::RNG branch)
    const double beta0 = __dmul_rn(live_t ? s.u_post[t] : 0.0, step);
    // (the 16-lane scan reaches any prefix sum through <= 4 roundings; the second term allows 14)
    const double tol = mkf_resample_tol(N, N, mw, step) + 16.0 * 1.1102230246251565e-16 * 2.0;
    double incl = x;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
        const double y = __shfl_up_sync(FULL, incl, o, 16);
        if (j >= o) incl += y;
    }
    bool amb = false;
    int e = N;
    if (live) {
        e = mkf_count_le_df(dd_add_d(dd_make(incl), -beta0), step, N, tol, amb);
        if (j == N - 1 && e < N) amb = true; // the literal loop would wrap past the last weight
    }
    int e_prev = __shfl_up_sync(FULL, e, 1, 16);
    if (j == 0) e_prev = 0;
    const bool degenerate = !(mw > 0.0);
    const unsigned amb_votes = __ballot_sync(FULL, amb); // (every lane votes: not inside the short-circuit below)
    const bool by_loop = tail && (degenerate || (amb_votes & gmask) != 0u);
    int* par = sm_par[wid][g];
    if (live) {
        if (!by_loop)
            for (int i = max(e_prev, 0); i < min(e, N); i++) par[i] = j;
    } else {
        par[j] = j; // (a row of zeros in sm_x)
    }
    if (by_loop && j == 0) { // rare: the group's first lane walks the loop / draws from cv::RNG
        if (degenerate) { // max weight 0 / NaN -> random indices from cv::RNG (src/pf2DRao.cpp:184-192)
            atomicOr(a.status + t, MKF_ST_POST_DEGENERATE);
            mkf_cvrng rng(s.seeds ? s.seeds[t * s.seed_stride + s.seed_off] : 1ull);
            (void)rng.uniform_int(0, N); // `int idx = rng.uniform(0, L);` drawn and discarded
            for (int i = 0; i < N; i++) par[i] = rng.uniform_int(0, N);
        } else {
            const double* wr = sm_w[wid][g];
            mkf_resample_sequential([&](int i) { return __ddiv_rn(wr[i], wsum); }, N, N, s.u_post[t],
                                    [&](int i, int idx) { par[i] = idx; });
        }
    }
    __syncwarp();
    if (tail && j == 0) {
        s.wsum[t] = wsum;
        if (s.unsorted_out) s.unsorted_out[t] = degenerate ? 1u : 0u;
    }
    if (tail && j < N) s.parent_out[s_] = par[j];

    // getEstimator: lane c sums element c of the children's means in the order of k_estimate_small's butterfly --
    // ((a0 + a8) + (a4 + a12)) + ((a2 + a10) + (a6 + a14)), the same for the odd children, then even + odd -- so the
    // estimate is bit-identical to that kernel's
    if (tail && j < D) {
        double a_[16];
        const int4* p4 = reinterpret_cast<const int4*>(par);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int4 pp = p4[q];
            a_[4 * q + 0] = 0.0 + sm_x[wid][g][pp.x][j];
            a_[4 * q + 1] = 0.0 + sm_x[wid][g][pp.y][j];
            a_[4 * q + 2] = 0.0 + sm_x[wid][g][pp.z][j];
            a_[4 * q + 3] = 0.0 + sm_x[wid][g][pp.w][j];
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1)
#pragma unroll
            for (int i = 0; i < o; i++) a_[i] = a_[i] + a_[i + o];
        sm_mean[wid][g][j] = a_[0] * step; // (step = fl(1 / N) = k_estimate_small's inv_n)
    }
    __syncwarp();
    if (tail) {
        double mean[D];
        const double2* m2 = reinterpret_cast<const double2*>(sm_mean[wid][g]);
#pragma unroll
        for (int p = 0; p < D / 2; p++) {
            const double2 q = m2[p];
            mean[2 * p] = q.x;
            mean[2 * p + 1] = q.y;
        }
        for (int r = j; r < R; r += 16) {
            double sacc = 0.0;
#pragma unroll
            for (int c = 0; c < D; c++) sacc = fma(coef[c * R + r], mean[c], sacc);
            if (r < s.Dpose) {
                const double pv = sacc + s.pmean[r];
                s.est_pose[t * s.Dpose + r] = pv;
                if (s.est_pose2) s.est_pose2[t * s.Dpose + r] = pv;
            } else {
                s.est_xbar[t * D + (r - s.Dpose)] = sacc;
            }
        }
    }
    if (a.ts && lane == 0) atomicMax(a.ts + 1, mkf_globaltimer());
}

#endif
