// mkf_runs.cuh -- the frame pipeline on RUN-LENGTH particle sets (included by mkf_api.cu).
//
// With one measurement column per track (MKF_MEAS_SHARED) and independent slots (MKF_ALIAS_INDEPENDENT), children that
// drew the same parent record and the same component are bit-identical Gaussians (mkf_kernels.cuh, record sharing), and
// because systematic resampling returns sorted parents they are consecutive slots.  The particle set of a track is
// therefore a short list of RUNS (record, multiplicity) -- ~50 for 500 slots in steady state -- and every step of
// ParticleFilter::update (src/pf2DRao.cpp:125-158) has a run-level form that never touches a per-slot array:
//
//   k_frame_heads     K -> N indicator draw (src/pf2DRao.cpp:128) in closed form -> K-1 cut positions; the frame's
//                     distinct Gaussians ("heads") are the runs cut at those positions.  One warp per track writes the
//                     head table {source record, component, multiplicity, first slot} and appends the heads to the
//                     batch-wide work list of k_slot_update_heads_direct (one atomicAdd per track).
//   k_slot_update_heads_direct   (mkf_kernels.cuh) predict + likelihood + update of every head, weight per head.
//   k_runs_repair     literal cv::Cholesky-failure semantics for flagged tracks (rare).
//   k_resample_runs   src/pf2DRao.cpp:139-156 on runs: wsum = sum m_h w_h, normalised weights, prefix sums at run ends in
//                     double-double, children per head from the closed-form count e(C) -> the next frame's run list.
//                     A run of m equal weights is one super-parent of weight m * w: which of its slots a threshold
//                     falls on does not change the child's Gaussian, so only the decisions AT run ends matter, and
//                     those are taken exactly as k_resample_block takes them (same ambiguity band, literal loop of
//                     src/pf2DRao.cpp:195-207 over the slots when undecidable, cv::RNG branch when max weight is 0).
//   k_estimate_runs   getEstimator (src/pf2DRao.cpp:23-31) = sum_runs multiplicity * x' / N, + PCA reconstruction.
//
// Per-slot views (parents, per-slot weights, states: mkf_batch_download; a following per-slot-measurement or
// association frame) are materialised on demand: k_expand_rep turns the head table into rep[] and the existing exact
// per-slot resampler k_resample_block replays the SAME resample (same head weights, the stored wsum, u and seed), so
// what it writes is what the run-level step decided.
#ifndef MKF_RUNS_CUH
#define MKF_RUNS_CUH

// one warp per track: the run list of a per-slot particle set (gi = record of every slot, sorted or not)
__global__ void __launch_bounds__(128) k_runs_from_slots(const int32_t* __restrict__ gi, long long T, int N,
                                                         int2* __restrict__ runs, int* __restrict__ nruns)
{
    mkf_pdl_launch_dependents();
    mkf_pdl_wait();
    const int lane = threadIdx.x & 31;
    const long long t = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (t >= T) return;
    const int32_t* __restrict__ g = gi + t * N;
    int2* __restrict__ rt = runs + t * N;
    int nr = 0, last = -2;
    for (int base = 0; base < N; base += 32) {
        const int j = base + lane;
        const int v = j < N ? g[j] : -1;
        int prev = __shfl_up_sync(0xffffffffu, v, 1);
        if (lane == 0) prev = last;
        const bool start = j < N && v != prev;
        const unsigned m = __ballot_sync(0xffffffffu, start);
        if (start) rt[nr + __popc(m & ((1u << lane) - 1u))] = make_int2(v, j); // .y = first slot for now
        nr += __popc(m);
        last = __shfl_sync(0xffffffffu, v, 31);
    }
    __syncwarp();
    for (int r0 = 0; r0 < nr; r0 += 32) { // first slots -> multiplicities
        const int r = r0 + lane;
        int a = 0, b = 0;
        if (r < nr) {
            a = rt[r].y;
            b = r + 1 < nr ? rt[r + 1].y : N;
        }
        __syncwarp();
        if (r < nr) rt[r].y = b - a;
        __syncwarp();
    }
    if (lane == 0) nruns[t] = nr;
}

struct FrameArgs {
    const double* __restrict__ u_ind;
    long long T;
    int N, K;
    const double* __restrict__ cw_hi;
    const double* __restrict__ cw_lo;
    const double* __restrict__ wprior;
    double wmax;
    int32_t* __restrict__ bounds;
    uint32_t* __restrict__ status;
    int clear_status;
    uint8_t* __restrict__ ind_tail;
    const int2* __restrict__ runs;
    const int* __restrict__ nruns;
    int4* __restrict__ hmeta; // T x N: head i of track t = {source record, component, multiplicity, first slot}
    int* __restrict__ nheads;
    int4* __restrict__ hd16;
    int* __restrict__ head_count;
    // MKF_MEAS_CAND (mkf_batch_associate with the update following): the slots of a track see the candidate their bin
    // selects, so the bin joins the key -- bins are sorted like components, given as run boundaries (k_resample_warp's
    // cut_out, 32 per (track, hand)); bins: the per-slot indices, read when the boundaries are flagged unusable
    const int32_t* __restrict__ bin_cuts; // null: one measurement per track
    const int32_t* __restrict__ bins;
    int cand_C, hand;
    // contiguous records in LIST order (mkf_heads_tma.cuh): the record of head i of track t is the one at the head's
    // position in the work list, lbase[t] + i.  lbase_prev: where the previous frame's heads (this frame's parents) lie,
    // lbase_cur (out): this frame's.  Null: records at t*N + i in the tile layout.
    const int* __restrict__ lbase_prev;
    int* __restrict__ lbase_cur;
    int dbg_frame;
    int pdl_late; // 1: the slot kernel's CTAs are released while the work list is written, not at this kernel's start
};

// component of slot j = number of cuts <= j (cuts sorted, nc <= 63): branch-free binary search
__device__ __forceinline__ int mkf_cuts_le(const int* cuts, int nc, int j)
{
    int lo = 0; // invariant: cuts[0..lo) <= j
#pragma unroll
    for (int s = 32; s > 0; s >>= 1)
        if (lo + s <= nc && cuts[lo + s - 1] <= j) lo += s;
    return lo;
}

constexpr int MKF_FH_WARPS = 8; // tracks per CTA of k_frame_heads

// the pieces of one run [a, b): cut at every component boundary and every candidate-bin boundary inside it.
// emit(key, length, first slot) per piece; returns the number of pieces.  key = component | bin << 8.
template <class F>
__device__ __forceinline__ int mkf_walk_pieces(const int* ccuts, int nc, const int* bcuts, int nb, int a, int b, int N,
                                               F emit)
{
    int kc = mkf_cuts_le(ccuts, nc, a), kb = nb > 0 ? mkf_cuts_le(bcuts, nb, a) : 0;
    int pos = a, np = 0;
    for (;;) {
        const int cc = kc < nc ? ccuts[kc] : N;
        const int cb = kb < nb ? bcuts[kb] : N;
        const int nx = cc < cb ? cc : cb;
        const int end = nx < b ? nx : b;
        emit(kc | (kb << 8), end - pos, pos, np);
        np++;
        if (nx >= b) break;
        pos = nx;
        while (kc < nc && ccuts[kc] <= pos) kc++;
        while (kb < nb && bcuts[kb] <= pos) kb++;
    }
    return np;
}

__global__ void __launch_bounds__(32 * MKF_FH_WARPS) k_frame_heads(const FrameArgs f)
{
    __shared__ int cuts_s[MKF_FH_WARPS][64];
    __shared__ int bcuts_s[MKF_FH_WARPS][32];
    __shared__ int nh_s[MKF_FH_WARPS];
    __shared__ int base_s;
    // A dependent released at this kernel's START sits on the SMs for its whole duration and then runs slower (the slot
    // kernel: 62 us instead of 54-55, profiles/r02_pdl_masks.txt); released while the work list is written it starts 1 us
    // after this kernel's end and runs at its normal speed.
    if (!f.pdl_late) mkf_pdl_launch_dependents();
    mkf_pdl_wait();
    MKF_TL_START(0, f.dbg_frame);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long t = (long long)blockIdx.x * MKF_FH_WARPS + wid;
    const bool live_t = t < f.T; // (a warp without a track still meets the CTA's barriers below)
    const int N = f.N, K = f.K;
    const int2* __restrict__ rt = f.runs + (live_t ? t : 0) * N;
    int4* __restrict__ hm = f.hmeta + (live_t ? t : 0) * N;
    const int nr = live_t ? f.nruns[t] : 0;
    int2 rn0 = make_int2(0, 0), rn1 = make_int2(0, 0); // (loads in flight during the indicator draw)
    if (lane < nr) rn0 = rt[lane];
    if (32 + lane < nr) rn1 = rt[32 + lane];
    int nh = 0;
    int mode = 0; // 1: fast path (pieces still to be written, in one walk with the work list), 2: head table written
    int a_[2] = {0, 0}, b_[2] = {0, 0}, h0_[2] = {0, 0};
    int* cuts = cuts_s[wid];
    int* bcuts = bcuts_s[wid];
    const int nc = K - 1; // cut q = first slot whose component exceeds q
    int nb = 0;           // likewise for the candidate bins
    const int tN = (int)((live_t ? t : 0) * N);
    if (live_t) {
        bool bins_per_slot = false;
        if (f.bin_cuts) {
            const int32_t* bc = f.bin_cuts + (t * 2 + f.hand) * 32;
            nb = f.cand_C - 1;
            const int v = lane < nb ? bc[lane] : 0;
            bins_per_slot = __shfl_sync(0xffffffffu, (lane == 0 ? bc[0] : 0), 0) < 0;
            if (lane < nb) bcuts[lane] = v;
        }
        // the K -> N indicator draw: the lanes of this warp are the group of k_indicator_bounds<32>
        int e_lo, e_hi;
        const bool closed = mkf_indicator_bounds_group<32>(t, lane, true, f.u_ind, N, K, f.cw_hi, f.cw_lo, f.wprior, f.wmax,
                                                           f.bounds, f.status, f.clear_status, f.ind_tail, e_lo, e_hi);
        const bool fast = __all_sync(0xffffffffu, closed);
        const int32_t* bt = f.bounds + t * (K + 2);
        int wrap_from = N;
        if (fast) { // the boundaries are still in registers
            if (lane < nc) cuts[lane] = e_lo;
            if (32 + lane < nc) cuts[32 + lane] = e_hi;
        } else {    // the literal loop wrote them (lane 0)
            __syncwarp();
            for (int q = lane; q < nc; q += 32) cuts[q] = __ldcg(bt + q);
            wrap_from = __ldcg(bt + K);
        }
        __syncwarp();
        const bool sorted_keys = wrap_from >= N && !bins_per_slot;
        if (sorted_keys && nr <= 64) {
            // the common case: the track's runs fit two per lane; count the pieces first
            mode = 1;
            int pos_carry = 0;
#pragma unroll
            for (int c = 0; c < 2; c++) {
                const int2 rn = c ? rn1 : rn0;
                const bool live = c * 32 + lane < nr;
                int inc = rn.y;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int n = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += n;
                }
                const int a = pos_carry + inc - rn.y, b = a + rn.y; // this run's slots [a, b)
                const int np = live ? mkf_walk_pieces(cuts, nc, bcuts, nb, a, b, N, [](int, int, int, int) {}) : 0;
                int pinc = np;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int n = __shfl_up_sync(0xffffffffu, pinc, o);
                    if (lane >= o) pinc += n;
                }
                a_[c] = a;
                b_[c] = b;
                h0_[c] = nh + pinc - np;
                nh += __shfl_sync(0xffffffffu, pinc, 31);
                pos_carry += __shfl_sync(0xffffffffu, inc, 31);
            }
        } else if (sorted_keys) {
            mode = 2;
            int pos_carry = 0;
            for (int r0 = 0; r0 < nr; r0 += 32) {
                const int r = r0 + lane;
                const int2 rn = r < nr ? rt[r] : make_int2(0, 0);
                int inc = rn.y;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int n = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += n;
                }
                const int a = pos_carry + inc - rn.y, b = a + rn.y; // this run's slots [a, b)
                const int np = r < nr ? mkf_walk_pieces(cuts, nc, bcuts, nb, a, b, N, [](int, int, int, int) {}) : 0;
                int pinc = np;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int n = __shfl_up_sync(0xffffffffu, pinc, o);
                    if (lane >= o) pinc += n;
                }
                if (r < nr) {
                    const int h0 = nh + pinc - np;
                    mkf_walk_pieces(cuts, nc, bcuts, nb, a, b, N, [&](int key, int len, int pos, int q) {
                        hm[h0 + q] = make_int4(rn.x, key, len, pos);
                    });
                }
                nh += __shfl_sync(0xffffffffu, pinc, 31);
                pos_carry += __shfl_sync(0xffffffffu, inc, 31);
            }
        } else {
            // keys that are not sorted in the slot index -- the indicator draw wrapped past the last component (prior
            // mass short of the thresholds), or the candidate bins came from the literal loop / cv::RNG: read per slot
            mode = 2;
            if (lane == 0) {
                const uint8_t* tail = f.ind_tail ? f.ind_tail + t * N : nullptr;
                const int32_t* bj = f.bins ? f.bins + (t * 2 + f.hand) * (long long)N : nullptr;
                auto key_of = [&](int j) { return mkf_component_of(bt, K, j, tail) | ((bj ? bj[j] : 0) << 8); };
                int pos = 0;
                for (int r = 0; r < nr; r++) {
                    const int2 rn = rt[r];
                    int j = pos;
                    const int b = pos + rn.y;
                    while (j < b) {
                        const int k = key_of(j);
                        int j2 = j + 1;
                        while (j2 < b && key_of(j2) == k) j2++;
                        hm[nh++] = make_int4(rn.x, k, j2 - j, j);
                        j = j2;
                    }
                    pos = b;
                }
            }
            nh = __shfl_sync(0xffffffffu, nh, 0);
        }
    }
    // one atomicAdd per CTA reserves its tracks' stretch of the batch-wide work list (order immaterial)
    if (lane == 0) {
        nh_s[wid] = nh;
        if (live_t) f.nheads[t] = nh;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
#pragma unroll
        for (int w = 0; w < MKF_FH_WARPS; w++) tot += nh_s[w];
        base_s = atomicAdd(f.head_count, tot);
    }
    __syncthreads();
    if (f.pdl_late) mkf_pdl_launch_dependents();
    int lb = base_s;
    for (int w = 0; w < wid; w++) lb += nh_s[w];
    // where the parents' records lie: the track's stretch of the previous list, or its own N slots
    const int src0 = f.lbase_prev ? (live_t ? f.lbase_prev[t] : 0) : tN;
    if (f.lbase_cur && live_t && lane == 0) f.lbase_cur[t] = lb;
    if (mode == 1) { // head table and work list in one walk
#pragma unroll
        for (int c = 0; c < 2; c++) {
            if (c * 32 + lane < nr) {
                const int rec = c ? rn1.x : rn0.x;
                const int h0 = h0_[c];
                mkf_walk_pieces(cuts, nc, bcuts, nb, a_[c], b_[c], N, [&](int key, int len, int pos, int q) {
                    hm[h0 + q] = make_int4(rec, key, len, pos);
                    f.hd16[lb + h0 + q] = make_int4(src0 + rec, tN + h0 + q, (int)t, key);
                });
            }
        }
    } else if (mode == 2) {
        __syncwarp();
        for (int i = lane; i < nh; i += 32) {
            const int4 m = __ldcg(hm + i);
            f.hd16[lb + i] = make_int4(src0 + m.x, tN + i, (int)t, m.y);
        }
    }
    MKF_TL_END(0, f.dbg_frame);
}

__device__ __forceinline__ dd dd_shfl_xor(dd v, int o)
{
    dd r;
    r.hi = __shfl_xor_sync(0xffffffffu, v.hi, o);
    r.lo = __shfl_xor_sync(0xffffffffu, v.lo, o);
    return r;
}
__device__ __forceinline__ dd dd_shfl_up(dd v, int o)
{
    dd r;
    r.hi = __shfl_up_sync(0xffffffffu, v.hi, o);
    r.lo = __shfl_up_sync(0xffffffffu, v.lo, o);
    return r;
}
// a + b for operands of the same sign (the weight sums): the error term of the low words is not re-split, error
// <= 2^-104 relative -- 11 operations instead of dd_add's 20
__device__ __forceinline__ dd dd_add_pos(dd a, dd b)
{
    dd s = dd_two_sum(a.hi, b.hi);
    s.lo = __dadd_rn(s.lo, __dadd_rn(a.lo, b.lo));
    return dd_fast_two_sum(s.hi, s.lo);
}
// m * w exactly (m an integer count) as a double-double
__device__ __forceinline__ dd dd_mul_exact(double m, double w)
{
    const double p = __dmul_rn(m, w);
    const double e = __fma_rn(m, w, -p);
    return dd_fast_two_sum(p, e);
}

struct ResampleRunsArgs {
    long long T;
    int N;
    const int4* __restrict__ hmeta;
    const int* __restrict__ nheads;
    const double* __restrict__ w_rec; // weight of head i of track t at t*N + i
    const double* __restrict__ u;
    const uint64_t* __restrict__ seeds;
    int seed_stride, seed_off;
    double* __restrict__ wsum_out;
    uint32_t* __restrict__ status;
    int2* __restrict__ runs; // out: the new particle set, {head index = record, children}
    int* __restrict__ nruns;
    double* __restrict__ u_keep;   // copies of the draw / seed of this resample, for the on-demand per-slot replay
    uint64_t* __restrict__ seed_keep;
    // getEstimator + reconstruction of the new set (src/pf2DRao.cpp:23-31, src/pfPose.cpp:347-348), left in the batch:
    // mkf_batch_estimate then only copies
    const double2* __restrict__ st_new; // the records the heads kernel just wrote
    int Dpose;
    const double* __restrict__ recon;
    const double* __restrict__ pmean;
    const double* __restrict__ tinv;
    double* __restrict__ est_xbar;  // T x d
    double* __restrict__ est_pose;  // T x Dpose
    double* __restrict__ est_pose2; // the association step's copy of the pose (or null)
    int aos;                        // st_new holds contiguous records (mkf_heads_tma.cuh) ...
    const int* __restrict__ lbase;  // ... in list order: head i of track t at lbase[t] + i
    // (aos) the heads' means x' once more, as [list position / 32][pair][list position % 32] double2: what the estimator
    // gathers, coalesced (the 96 bytes at the start of 720-byte records are a scattered read: 18 -> 29 us at 4096 x 500)
    const double2* __restrict__ xs;
    int dbg_frame;
};

// One track, by one warp (every lane enters).  coef: the reconstruction coefficients staged in shared memory as
// [c][r], r < Dpose + D (rows of recon, then rows of tinv).  The head table and the weights may have been written by
// this very warp a moment ago (k_frame_fused), so they are read past the L1 (ld.global.cg).
template <int D>
__device__ __forceinline__ void mkf_resample_runs_track(const ResampleRunsArgs& a, const long long t, const int lane,
                                                        const double* coef, const int nh)
{
    using L = SlotLay<D>;
    const int R = a.Dpose + D;
    const int N = a.N;
    const int4* hm = a.hmeta + t * N;
    const double* wr = a.w_rec + t * N;
    int2* rt = a.runs + t * N;
    const long long rbase = a.lbase ? (long long)a.lbase[t] : t * N; // the track's records in st_new
    // estimator of the new set (sum over its runs of children x mean): gathered as soon as a head's children are known
    double xs[D];
    bool xs_done = false;

    // pass 1: wsum = sum over slots (src/pf2DRao.cpp:139) = sum_h m_h w_h, accumulated in double-double and rounded once;
    // NaN-ignoring max (src/pf2DRao.cpp:161-172)
    dd acc = dd_make(0.0);
    double mx = 0.0, sq = 0.0;
    for (int i = lane; i < nh; i += 32) {
        const double w = __ldcg(wr + i), md = (double)__ldcg(&hm[i].z);
        acc = dd_add_pos(acc, dd_mul_exact(md, w));
        if (w > mx) mx = w;
        sq = fma(__dmul_rn(md, w), w, sq);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc = dd_add_pos(acc, dd_shfl_xor(acc, o));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
    }
    const double wsum = acc.hi;
    const double uu = a.u[t];
    const uint64_t seed = a.seeds ? a.seeds[t * a.seed_stride + a.seed_off] : 1ull;
    if (lane == 0) {
        a.wsum_out[t] = wsum;
        a.u_keep[t] = uu;
        a.seed_keep[t] = seed;
    }
    const double wmax_n = __ddiv_rn(mx, wsum);
    int nr = 0;
    if (!(wmax_n > 0.0)) { // max weight 0 / NaN -> N random parents from cv::RNG (src/pf2DRao.cpp:184-192)
        if (lane == 0) {
            atomicOr(a.status + t, MKF_ST_POST_DEGENERATE);
            mkf_cvrng rng(seed);
            (void)rng.uniform_int(0, N); // `int idx = rng.uniform(0, L);` drawn and discarded
            int cur = -1, cnt = 0;
            for (int i = 0; i < N; i++) {
                const int idx = rng.uniform_int(0, N); // a SLOT; its record is the head whose range holds it
                int lo = 0, hi = nh - 1;
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (__ldcg(&hm[mid].w) <= idx)
                        lo = mid;
                    else
                        hi = mid - 1;
                }
                if (lo == cur) {
                    cnt++;
                } else {
                    if (cnt) rt[nr++] = make_int2(cur, cnt);
                    cur = lo;
                    cnt = 1;
                }
            }
            if (cnt) rt[nr++] = make_int2(cur, cnt);
        }
        nr = __shfl_sync(0xffffffffu, nr, 0);
    } else {
        const double step = __ddiv_rn(1.0, (double)N);
        const double beta0 = __dmul_rn(uu, step);
        // the band in which the literal loop's accumulated rounding could change a decision (mkf_resample_tol*), plus
        // the rounding of mkf_count_le's remainder and of the double-double prefix sums (as k_resample_block's second
        // opinion)
        const double s2 = __ddiv_rn(sq, __dmul_rn(wsum, wsum)) * (1.0 + 1e-9);
        const double tol_loop = fmin(mkf_resample_tol(N, N, wmax_n, step), mkf_resample_tol_s2(N, N, s2, 1.0));
        const double tol2 = tol_loop + 8.0 * 1.1102230246251565e-16 * step + 8.0e-28 * 2.0;
        // pass 2: prefix sums at run ends -> children per head -> the new run list.  First with the prefix sums in plain
        // double (<= 8 roundings of partial sums <= 1 on the way to any run end: the same 16 x 2^-53 x (1 + mass) term
        // as k_resample_block adds to the band); a track with a threshold inside the band gets a second opinion with
        // the prefix sums in double-double (only the loop's own bound left), and then the literal loop.
        bool amb = false;
        for (int pass = 0; pass < 2; pass++) {
            const double tol = pass == 0 ? tol_loop + 16.0 * 1.1102230246251565e-16 * 2.0 : tol2;
            dd carry = dd_make(0.0);
            int e_carry = 0;
            amb = false;
            nr = 0;
#pragma unroll
            for (int e = 0; e < D; e++) xs[e] = 0.0;
            for (int i0 = 0; i0 < nh; i0 += 32) {
                const int i = i0 + lane;
                const bool valid = i < nh;
                const double wn = valid ? __ddiv_rn(__ldcg(wr + i), wsum) : 0.0; // normalised weight of each of the head's slots
                const double md = valid ? (double)__ldcg(&hm[i].z) : 0.0;
                dd C;
                if (pass == 0) {
                    double inc = __dmul_rn(md, wn);
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const double n = __shfl_up_sync(0xffffffffu, inc, o);
                        if (lane >= o) inc += n;
                    }
                    C = dd_make(carry.hi + inc);
                } else {
                    dd inc = dd_mul_exact(md, wn);
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const dd n = dd_shfl_up(inc, o);
                        if (lane >= o) inc = dd_add(n, inc);
                    }
                    C = dd_add(carry, inc);
                }
                int eh = N;
                if (valid) {
                    eh = mkf_count_le(C, beta0, step, N, tol, amb);
                    if (i == nh - 1 && eh < N) amb = true; // the literal loop would wrap past the last slot
                }
                int eprev = __shfl_up_sync(0xffffffffu, eh, 1);
                if (lane == 0) eprev = e_carry;
                const int c = valid ? eh - eprev : 0;
                const unsigned msk = __ballot_sync(0xffffffffu, c > 0);
                if (c > 0) {
                    rt[nr + __popc(msk & ((1u << lane) - 1u))] = make_int2(i, c);
                    const long long sp = rbase + i; // head i's record, just written by the slot kernel
                    const double2* __restrict__ src =
                        a.xs ? a.xs + (sp >> 5) * (D / 2 * 32) + (sp & 31) : a.st_new + mkf_rec_base<D>(sp, a.aos);
                    const double m = (double)c;
#pragma unroll
                    for (int p = 0; p < D / 2; p++) {
                        const double2 q = __ldg(src + (a.xs ? 32 * p : mkf_rec_off<D>(p, a.aos)));
                        xs[2 * p] = fma(m, q.x, xs[2 * p]);
                        xs[2 * p + 1] = fma(m, q.y, xs[2 * p + 1]);
                    }
                }
                nr += __popc(msk);
                const int lastl = (nh - i0 >= 32) ? 31 : (nh - i0 - 1);
                e_carry = __shfl_sync(0xffffffffu, eh, lastl);
                carry.hi = __shfl_sync(0xffffffffu, C.hi, 31); // lane 31's inclusive prefix (lanes beyond nh add zero)
                carry.lo = __shfl_sync(0xffffffffu, C.lo, 31);
            }
            amb = __any_sync(0xffffffffu, amb);
            if (!amb) break;
        }
        xs_done = !amb;
        if (amb) {
            // undecidable in closed form: the reference's loop itself (src/pf2DRao.cpp:195-207), slot by slot
            if (lane == 0) {
                atomicOr(a.status + t, MKF_ST_POST_FALLBACK);
                int h = 0, left = __ldcg(&hm[0].z);
                double wi = __ddiv_rn(__ldcg(wr), wsum);
                double beta = beta0;
                int cur = -1, cnt = 0;
                nr = 0;
                for (int i = 0; i < N; i++) {
                    while (beta > wi) {
                        beta = __dsub_rn(beta, wi);
                        if (--left == 0) { // idx = (idx + 1) % L moved on to the next head's first slot
                            h = (h + 1 == nh) ? 0 : h + 1;
                            left = __ldcg(&hm[h].z);
                            wi = __ddiv_rn(__ldcg(wr + h), wsum);
                        }
                    }
                    beta = __dadd_rn(beta, step);
                    if (h == cur) {
                        cnt++;
                    } else {
                        if (cnt) rt[nr++] = make_int2(cur, cnt);
                        cur = h;
                        cnt = 1;
                    }
                }
                if (cnt) rt[nr++] = make_int2(cur, cnt);
            }
            nr = __shfl_sync(0xffffffffu, nr, 0);
        }
    }
    if (lane == 0) a.nruns[t] = nr;
    __syncwarp();

    if (!xs_done) { // the rare branches (literal loop, cv::RNG indices) left only the run list: gather from it
#pragma unroll
        for (int e = 0; e < D; e++) xs[e] = 0.0;
        for (int r = lane; r < nr; r += 32) {
            const int2 rn = __ldcg(rt + r); // written by lane 0 just above
            const long long sp = rbase + rn.x;
            const double2* __restrict__ src =
                a.xs ? a.xs + (sp >> 5) * (D / 2 * 32) + (sp & 31) : a.st_new + mkf_rec_base<D>(sp, a.aos);
            const double m = (double)rn.y;
#pragma unroll
            for (int p = 0; p < D / 2; p++) {
                const double2 q = __ldg(src + (a.xs ? 32 * p : mkf_rec_off<D>(p, a.aos)));
                xs[2 * p] = fma(m, q.x, xs[2 * p]);
                xs[2 * p + 1] = fma(m, q.y, xs[2 * p + 1]);
            }
        }
    }
    const double inv_n = 1.0 / (double)N;
#pragma unroll
    for (int e = 0; e < D; e++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) xs[e] += __shfl_xor_sync(0xffffffffu, xs[e], o);
        xs[e] *= inv_n;
    }
    for (int r = lane; r < R; r += 32) {
        double sacc = 0.0;
#pragma unroll
        for (int c = 0; c < D; c++) sacc = fma(coef[c * R + r], xs[c], sacc);
        if (r < a.Dpose) {
            const double v = sacc + __ldg(a.pmean + r);
            a.est_pose[t * a.Dpose + r] = v;
            if (a.est_pose2) a.est_pose2[t * a.Dpose + r] = v;
        } else {
            a.est_xbar[t * D + (r - a.Dpose)] = sacc;
        }
    }
}

template <int D>
__global__ void __launch_bounds__(128, 7) k_resample_runs(const ResampleRunsArgs a)
{
    extern __shared__ double coef[]; // [c][r], r < Dpose + D: rows of recon (pose) then rows of tinv (xbar)
    mkf_pdl_launch_dependents();
    const int R = a.Dpose + D;
    for (int i = threadIdx.x; i < R * D; i += 128) { // model constants: safe before the dependency wait
        const int r = i / D, c = i - r * D;
        coef[c * R + r] = r < a.Dpose ? a.recon[r * D + c] : a.tinv[(r - a.Dpose) * D + c];
    }
    mkf_pdl_wait();
    MKF_TL_START(3, a.dbg_frame);
    __syncthreads();
    const long long t = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (t >= a.T) return;
    mkf_resample_runs_track<D>(a, t, threadIdx.x & 31, coef, a.nheads[t]);
    MKF_TL_END(3, a.dbg_frame);
}

// -----------------------------------------------------------------------------------------
// The whole frame in ONE launch.  k_frame_heads / k_slot_update_heads_direct / k_resample_runs are three grid-wide
// phases, but nothing in a frame couples two tracks: a warp can take a few tracks through all three phases on its own
// -- head table (K -> N draw, runs cut at the component boundaries), predict + likelihood + update of those heads (one
// lane per head, the heads of the warp's tracks packed into full steps of 32), weight sum / resample / estimate -- with
// no grid-wide dependency, no work list and no atomics.  The bookkeeping phases are latency chains of a few
// microseconds; here they run in the shadow of the other resident warps' slot arithmetic instead of as separate
// kernels in front of and behind it, and the warps of an SM drift out of phase, which the grid-wide version's
// load-all / compute-all / store-all rhythm never did.
//   Tracks are block-partitioned over the warps of a persistent grid (2 CTAs of 4 warps per SM); a warp works through
//   its block in groups of up to MKF_FUSE_G tracks.  The head table, weights and run lists live in global memory
//   exactly as in the three-kernel pipeline (they are L2-resident and the per-slot replay reads them later).
//   A cv::Cholesky failure (never seen on real data) only flags the track: k_runs_repair redoes it AND its resample.
// -----------------------------------------------------------------------------------------
constexpr int MKF_FUSE_G = 4;

// the head table of one track (same result as k_frame_heads, without the batch-wide work list); every lane enters
__device__ __forceinline__ int mkf_frame_heads_track(const FrameArgs& f, const long long t, const int lane, int* cuts,
                                                     int* bcuts)
{
    const int N = f.N, K = f.K;
    const int2* rt = f.runs + t * N;
    int4* hm = f.hmeta + t * N;
    const int nr = __ldcg(f.nruns + t);
    const int nc = K - 1;
    int nb = 0;
    bool bins_per_slot = false;
    if (f.bin_cuts) {
        const int32_t* bc = f.bin_cuts + (t * 2 + f.hand) * 32;
        nb = f.cand_C - 1;
        const int v = lane < nb ? bc[lane] : 0;
        bins_per_slot = __shfl_sync(0xffffffffu, (lane == 0 ? bc[0] : 0), 0) < 0;
        if (lane < nb) bcuts[lane] = v;
    }
    int e_lo, e_hi;
    const bool closed = mkf_indicator_bounds_group<32>(t, lane, true, f.u_ind, N, K, f.cw_hi, f.cw_lo, f.wprior, f.wmax,
                                                       f.bounds, f.status, f.clear_status, f.ind_tail, e_lo, e_hi);
    const bool fast = __all_sync(0xffffffffu, closed);
    const int32_t* bt = f.bounds + t * (K + 2);
    int wrap_from = N;
    if (fast) {
        if (lane < nc) cuts[lane] = e_lo;
        if (32 + lane < nc) cuts[32 + lane] = e_hi;
    } else {
        __syncwarp();
        for (int q = lane; q < nc; q += 32) cuts[q] = __ldcg(bt + q);
        wrap_from = __ldcg(bt + K);
    }
    __syncwarp();
    int nh = 0;
    if (wrap_from >= N && !bins_per_slot) {
        int pos_carry = 0;
        for (int r0 = 0; r0 < nr; r0 += 32) {
            const int r = r0 + lane;
            const int2 rn = r < nr ? __ldcg(rt + r) : make_int2(0, 0);
            int inc = rn.y;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int n = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += n;
            }
            const int a = pos_carry + inc - rn.y, b = a + rn.y;
            const int np = r < nr ? mkf_walk_pieces(cuts, nc, bcuts, nb, a, b, N, [](int, int, int, int) {}) : 0;
            int pinc = np;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int n = __shfl_up_sync(0xffffffffu, pinc, o);
                if (lane >= o) pinc += n;
            }
            if (r < nr) {
                const int h0 = nh + pinc - np;
                mkf_walk_pieces(cuts, nc, bcuts, nb, a, b, N,
                                [&](int key, int len, int pos, int q) { hm[h0 + q] = make_int4(rn.x, key, len, pos); });
            }
            nh += __shfl_sync(0xffffffffu, pinc, 31);
            pos_carry += __shfl_sync(0xffffffffu, inc, 31);
        }
    } else {
        if (lane == 0) {
            const uint8_t* tail = f.ind_tail ? f.ind_tail + t * N : nullptr;
            const int32_t* bj = f.bins ? f.bins + (t * 2 + f.hand) * (long long)N : nullptr;
            auto key_of = [&](int j) { return mkf_component_of(bt, K, j, tail) | ((bj ? bj[j] : 0) << 8); };
            int pos = 0;
            for (int r = 0; r < nr; r++) {
                const int2 rn = __ldcg(rt + r);
                int j = pos;
                const int b = pos + rn.y;
                while (j < b) {
                    const int k = key_of(j);
                    int j2 = j + 1;
                    while (j2 < b && key_of(j2) == k) j2++;
                    hm[nh++] = make_int4(rn.x, k, j2 - j, j);
                    j = j2;
                }
                pos = b;
            }
        }
        nh = __shfl_sync(0xffffffffu, nh, 0);
    }
    if (lane == 0) f.nheads[t] = nh;
    __syncwarp();
    return nh;
}

template <int D>
__global__ void __launch_bounds__(128, 2) k_frame_fused(const FrameArgs f, const SlotArgs a, const ResampleRunsArgs ra)
{
    using L = SlotLay<D>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* cst = reinterpret_cast<double*>(smem_raw);          // K x CS model constants (TMA)
    double* coef = cst + a.K * L::CS;                            // (Dpose + D) x D reconstruction coefficients
    __shared__ int cuts_s[4][64];
    __shared__ int bcuts_s[4][32];
    __shared__ __align__(8) uint64_t mbar;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t cbytes = (uint32_t)(a.K * L::CS * sizeof(double));
    if (tid == 0) mkf_mbar_init(&mbar, 1);
    __syncthreads();
    if (tid == 0) {
        mkf_mbar_expect_tx(&mbar, cbytes);
        mkf_tma_load_1d(cst, a.comp_const, cbytes, &mbar); // model constants: never written by the frame chain
    }
    const int R = ra.Dpose + D;
    for (int i = tid; i < R * D; i += 128) {
        const int r = i / D, c = i - r * D;
        coef[c * R + r] = r < ra.Dpose ? ra.recon[r * D + c] : ra.tinv[(r - ra.Dpose) * D + c];
    }
    mkf_pdl_launch_dependents();
    mkf_pdl_wait();
    __syncthreads();
    mkf_mbar_wait(&mbar, 0);

    // this warp's block of tracks
    const long long nwarps = (long long)gridDim.x * 4, w = (long long)blockIdx.x * 4 + wid;
    const long long t_begin = f.T * w / nwarps, t_end = f.T * (w + 1) / nwarps;
    const int N = f.N;
    for (long long t0 = t_begin; t0 < t_end; t0 += MKF_FUSE_G) {
        const int ng = (int)((t_end - t0) < MKF_FUSE_G ? (t_end - t0) : MKF_FUSE_G);
        // phase 1: head tables
        int nh[MKF_FUSE_G], hsum = 0;
#pragma unroll
        for (int g = 0; g < MKF_FUSE_G; g++) {
            nh[g] = 0;
            if (g < ng) nh[g] = mkf_frame_heads_track(f, t0 + g, lane, cuts_s[wid], bcuts_s[wid]);
            hsum += nh[g];
        }
        // phase 2: one lane per head, the group's heads packed into steps of 32
        for (int q0 = 0; q0 < hsum; q0 += 32) {
            int q = q0 + lane;
            if (q < hsum) {
                int g = 0;
#pragma unroll
                for (int gg = 0; gg < MKF_FUSE_G - 1; gg++)
                    if (g == gg && q >= nh[gg]) {
                        q -= nh[gg];
                        g = gg + 1;
                    }
                const long long t = t0 + g;
                const int4 m = __ldcg(f.hmeta + t * N + q);
                const long long sp = t * N + m.x, so_rec = t * N + q;
                const double2* __restrict__ src = a.st_in + (sp >> 5) * (long long)L::TILE2 + (sp & 31) * L::H;
                double v[L::NE];
#pragma unroll
                for (int p = 0; p < L::NP; p++) {
                    const double2 qq = __ldg(src + L::po(p));
                    v[2 * p] = qq.x;
                    if (2 * p + 1 < L::NE) v[2 * p + 1] = qq.y;
                }
                double zc[MKF_M];
                if (a.meas_layout == MKF_MEAS_CAND)
                    mkf_load_meas_cand(a, t, m.y >> 8, zc); // the column of the head's candidate bin
                else
                    mkf_load_meas(a, t, 0, zc); // shared layout: the track's column
                double wgt;
                const bool ok = slot_math<D, false>(v, cst + (m.y & 0xff) * L::CS, zc, a.r, a.chol_mode, a.stage, wgt);
                if (!ok) atomicOr(a.status + t, MKF_ST_CHOL_FAIL);
                double2* __restrict__ dst = a.st_out + (so_rec >> 5) * (long long)L::TILE2 + (so_rec & 31) * L::H;
#pragma unroll
                for (int p = 0; p < L::NP; p++) {
                    double2 qq;
                    qq.x = v[2 * p];
                    qq.y = (2 * p + 1 < L::NE) ? v[2 * p + 1] : 0.0;
                    __stcg(dst + L::po(p), qq); // (read back below for the estimate: keep it in L2)
                }
                __stcg(a.w_rec + so_rec, wgt);
            }
        }
        __syncwarp();
        // phase 3: weight sum, resample, estimate
#pragma unroll
        for (int g = 0; g < MKF_FUSE_G; g++) {
            if (g < ng) {
                const long long t = t0 + g;
                const uint32_t st = __ldcg(a.status + t); // (a failure flagged by one of this warp's lanes above)
                if (!(st & MKF_ST_CHOL_FAIL)) mkf_resample_runs_track<D>(ra, t, lane, coef, nh[g]);
            }
        }
    }
}

// literal cv::Cholesky-failure semantics for flagged tracks, head by head (see k_slot_update_repair)
// ra (optional): the fused frame kernel skipped the resample of a flagged track -- done here after its heads
template <int D>
__global__ void __launch_bounds__(128, 2) k_runs_repair(const SlotArgs a, const int4* __restrict__ hmeta,
                                                        const int* __restrict__ nheads, const ResampleRunsArgs ra,
                                                        int with_resample)
{
    using L = SlotLay<D>;
    extern __shared__ double coef_r[];
    __shared__ uint32_t flags[128];
    mkf_pdl_launch_dependents();
    mkf_pdl_wait();
    MKF_TL_START(2, a.dbg_frame);
    MKF_TL_END(2, a.dbg_frame);
    const long long T = a.total / a.N;
    const long long base = (long long)blockIdx.x * 128;
    {
        const long long t = base + threadIdx.x;
        uint32_t fl = 0;
        if (t < T) fl = (a.status[t] & MKF_ST_CHOL_FAIL) ? 1u : 0u;
        flags[threadIdx.x] = fl;
        if (!__syncthreads_or((int)fl)) return;
    }
    for (int q = 0; q < 128; q++) {
        if (!flags[q]) continue;
        const long long t = base + q;
        const int nh = nheads[t];
        for (int i = threadIdx.x; i < nh; i += 128) {
            const int4 m = hmeta[t * a.N + i];
            const long long so = t * a.N + i;
            const long long sp = a.lbase_prev ? (long long)a.lbase_prev[t] + m.x : t * a.N + m.x;
            const long long sd = a.lbase_cur ? (long long)a.lbase_cur[t] + i : so;
            const double2* src = a.st_in + mkf_rec_base<D>(sp, a.aos);
            double v[L::NE];
            for (int p = 0; p < L::NP; p++) {
                const double2 qq = src[mkf_rec_off<D>(p, a.aos)];
                v[2 * p] = qq.x;
                if (2 * p + 1 < L::NE) v[2 * p + 1] = qq.y;
            }
            double zc[MKF_M], w;
            if (a.meas_layout == MKF_MEAS_CAND)
                mkf_load_meas_cand(a, t, m.y >> 8, zc); // the column of the head's candidate bin
            else
                mkf_load_meas(a, t, 0, zc);
            slot_math<D, true>(v, a.comp_const + (long long)(m.y & 0xff) * L::CS, zc, a.r, a.chol_mode, a.stage, w);
            double2* dst = a.st_out + mkf_rec_base<D>(sd, a.aos);
            for (int p = 0; p < L::NP; p++) {
                double2 qq;
                qq.x = v[2 * p];
                qq.y = (2 * p + 1 < L::NE) ? v[2 * p + 1] : 0.0;
                dst[mkf_rec_off<D>(p, a.aos)] = qq;
                if (a.xs && p < D / 2) a.xs[(sd >> 5) * (D / 2 * 32) + 32 * p + (sd & 31)] = qq;
            }
            a.w_rec[so] = w;
        }
        if (with_resample) {
            __syncthreads();
            __threadfence_block();
            const int R = ra.Dpose + D;
            for (int i = threadIdx.x; i < R * D; i += 128) {
                const int r = i / D, c = i - r * D;
                coef_r[c * R + r] = r < ra.Dpose ? ra.recon[r * D + c] : ra.tinv[(r - ra.Dpose) * D + c];
            }
            __syncthreads();
            if (threadIdx.x < 32) mkf_resample_runs_track<D>(ra, t, threadIdx.x, coef_r, nh);
            __syncthreads();
        }
    }
}

// getEstimator + reconstruction from the run list: 4 tracks per CTA, one warp each
template <int D>
__global__ void __launch_bounds__(128) k_estimate_runs(const double2* __restrict__ st, const int2* __restrict__ runs,
                                                       const int* __restrict__ nruns, long long T, int N, int Dpose,
                                                       const double* __restrict__ recon, const double* __restrict__ pmean,
                                                       const double* __restrict__ tinv, double* __restrict__ xbar_out,
                                                       double* __restrict__ pose_out, double* __restrict__ pose_out2)
{
    using L = SlotLay<D>;
    extern __shared__ double coef[]; // [c][r], r < Dpose + D: rows of recon (pose) then rows of tinv (xbar)
    mkf_pdl_launch_dependents();
    const int R = Dpose + D;
    for (int i = threadIdx.x; i < R * D; i += 128) { // model constants: safe before the dependency wait
        const int r = i / D, c = i - r * D;
        coef[c * R + r] = r < Dpose ? recon[r * D + c] : tinv[(r - Dpose) * D + c];
    }
    mkf_pdl_wait();
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long t = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (t >= T) return;
    const int nr = nruns[t];
    const int2* __restrict__ rt = runs + t * N;
    double acc[D];
#pragma unroll
    for (int e = 0; e < D; e++) acc[e] = 0.0;
    for (int r = lane; r < nr; r += 32) {
        const int2 rn = rt[r];
        const long long sp = t * N + rn.x;
        const double2* __restrict__ src = st + (sp >> 5) * (long long)L::TILE2 + (sp & 31) * L::H;
        const double m = (double)rn.y;
#pragma unroll
        for (int p = 0; p < D / 2; p++) {
            const double2 q = __ldg(src + L::po(p));
            acc[2 * p] = fma(m, q.x, acc[2 * p]);
            acc[2 * p + 1] = fma(m, q.y, acc[2 * p + 1]);
        }
    }
    const double inv_n = 1.0 / (double)N;
#pragma unroll
    for (int e = 0; e < D; e++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], o);
        acc[e] *= inv_n;
    }
    for (int r = lane; r < R; r += 32) {
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < D; c++) s = fma(coef[c * R + r], acc[c], s);
        if (r < Dpose) {
            if (pose_out) pose_out[t * Dpose + r] = s + pmean[r];
            if (pose_out2) pose_out2[t * Dpose + r] = s + pmean[r];
        } else if (xbar_out) {
            xbar_out[t * D + (r - Dpose)] = s;
        }
    }
}

// head table -> rep[] (the record of st[cur] that holds every slot of the LAST frame), for the per-slot replay
__global__ void __launch_bounds__(128) k_expand_rep(const int4* __restrict__ hmeta, const int* __restrict__ nheads,
                                                    long long T, int N, int32_t* __restrict__ rep)
{
    const int lane = threadIdx.x & 31;
    const long long t = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (t >= T) return;
    const int nh = nheads[t];
    for (int i0 = 0; i0 < nh; i0 += 32) {
        // the lanes of a warp walk the slots of 32 consecutive heads together: a head of m slots costs ceil(m / 32) trips
        for (int q = 0; q < 32 && i0 + q < nh; q++) {
            const int4 m = hmeta[t * N + i0 + q];
            for (int j = lane; j < m.z; j += 32) rep[t * N + m.w + j] = i0 + q;
        }
    }
}

// run list -> src[] (the record of st[cur] that holds every slot of the CURRENT set) when no per-slot replay is
// possible (entry into run mode without an update: not needed, the per-slot arrays are still valid then)

#endif
