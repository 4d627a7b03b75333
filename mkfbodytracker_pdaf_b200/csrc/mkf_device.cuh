// mkf_device.cuh -- device-side building blocks: double-double arithmetic, the exact
// systematic-resampling count, TMA (1-D bulk copy) + mbarrier helpers, block scans.
#ifndef MKF_DEVICE_CUH
#define MKF_DEVICE_CUH

#include <cuda_runtime.h>
#include <stdint.h>

// ---------------------------------------------------------------------------------------------
// double-double (error-free transformations; no FMA contraction can alter pure add/sub chains)
// ---------------------------------------------------------------------------------------------
struct dd {
    double hi, lo;
};

__device__ __forceinline__ dd dd_make(double a)
{
    dd r;
    r.hi = a;
    r.lo = 0.0;
    return r;
}
__device__ __forceinline__ dd dd_two_sum(double a, double b)
{
    dd r;
    r.hi = __dadd_rn(a, b);
    double bb = __dsub_rn(r.hi, a);
    r.lo = __dadd_rn(__dsub_rn(a, __dsub_rn(r.hi, bb)), __dsub_rn(b, bb));
    return r;
}
__device__ __forceinline__ dd dd_fast_two_sum(double a, double b) // |a| >= |b|
{
    dd r;
    r.hi = __dadd_rn(a, b);
    r.lo = __dsub_rn(b, __dsub_rn(r.hi, a));
    return r;
}
__device__ __forceinline__ dd dd_add_d(dd a, double b)
{
    dd s = dd_two_sum(a.hi, b);
    s.lo = __dadd_rn(s.lo, a.lo);
    return dd_fast_two_sum(s.hi, s.lo);
}

__device__ __forceinline__ dd dd_add(dd a, dd b) // accurate variant: error <= 2 * 2^-106 relative
{
    dd s = dd_two_sum(a.hi, b.hi);
    const dd t = dd_two_sum(a.lo, b.lo);
    s.lo = __dadd_rn(s.lo, t.hi);
    s = dd_fast_two_sum(s.hi, s.lo);
    s.lo = __dadd_rn(s.lo, t.lo);
    return dd_fast_two_sum(s.hi, s.lo);
}

// ---------------------------------------------------------------------------------------------
// exact systematic resampling count (src/pf2DRao.cpp:195-207 in closed form)
//
// The reference walks thresholds T_i = beta0 + i*step (beta0 = fl(u*step), step = fl(1/N)) over the
// running sum of the weights.  In exact arithmetic, parent k receives the outputs i with
// C_{k-1} < T_i <= C_k, so with e_k = #{ i in [0,N) : T_i <= C_k } parent k owns [e_{k-1}, e_k).
// count_le returns e for a prefix sum C held in double-double and reports `amb` when a valid
// threshold lies within `tol` of C -- tol bounds the rounding error the reference's sequential
// `beta -= w; beta += step` loop can have accumulated, so outside that band the loop provably takes
// the same decision.  Ambiguous tracks are re-run by the exact sequential kernel.
// ---------------------------------------------------------------------------------------------
// df = C - beta0 (double-double)
__device__ __forceinline__ int mkf_count_le_df(dd df, double step, int N, double tol, bool& amb)
{
    // quotient estimate: step = fl(1/N), so df.hi * N is within one unit of df.hi / step; the remainder test
    // below corrects it (a double division here would cost more than the rest of the function)
    double qi = floor(__dmul_rn(df.hi, (double)N));
    double rem = __dadd_rn(__fma_rn(-qi, step, df.hi), df.lo); // C - T_qi
    if (rem < 0.0) {
        qi -= 1.0;
        rem = __dadd_rn(__fma_rn(-qi, step, df.hi), df.lo);
    } else if (rem >= step) {
        qi += 1.0;
        rem = __dadd_rn(__fma_rn(-qi, step, df.hi), df.lo);
    }
    // T_qi <= C < T_{qi+1}; distances rem and step - rem
    const double nm1 = (double)(N - 1);
    if (qi >= 0.0 && qi <= nm1 && rem <= tol) amb = true;
    if (qi + 1.0 >= 0.0 && qi + 1.0 <= nm1 && (step - rem) <= tol) amb = true;
    if (!(rem >= 0.0 && rem < step)) amb = true; // NaN / failed correction: let the exact loop decide
    double cnt = qi + 1.0;
    if (!(cnt > 0.0)) return 0;
    if (cnt >= (double)N) return N;
    return (int)cnt;
}
__device__ __forceinline__ int mkf_count_le(dd C, double beta0, double step, int N, double tol, bool& amb)
{
    return mkf_count_le_df(dd_add_d(C, -beta0), step, N, tol, amb);
}

// bound on the rounding error accumulated by the reference loop: every one of its <= N+L
// add/subtract results is <= wmax + step, each rounded with relative error 2^-53
__device__ __forceinline__ double mkf_resample_tol(int N, int L, double wmax, double step)
{
    return 1.0625 * 1.1102230246251565e-16 * (double)(N + L) * (wmax + step);
}

// A second bound on the same error, tighter when the weights are peaked.  beta never leaves [0, w_idx + step]: a
// subtraction happens only while beta > w_idx (result in (0, beta)), an addition only once beta <= w_idx.  So the
// result of the addition that emits output i is <= w_parent(i) + step and the result of the subtraction that leaves
// weight k is <= w_{k-1} + step; each is rounded with relative error 2^-53 and the errors add.  Parent k emits
// c_k <= N w_k + 2 outputs, hence
//     sum over additions    <= sum_k c_k w_k + N step <= N * S2 + 2 * mass + 1        (S2 = sum_k w_k^2)
//     sum over subtractions <= mass + L step           =  mass + L / N                 (no wrap; wraps are flagged)
// and the accumulated error is <= 2^-53 (N S2 + 3 mass + 1 + L/N).  For uniform weights this equals the bound above;
// at N = 65 536 with a few thousand effective parents it is ~50x smaller (the first bound charges w_max to every
// operation).  s2 must be an upper bound of S2 (callers inflate the computed sum).
__device__ __forceinline__ double mkf_resample_tol_s2(int N, int L, double s2, double mass)
{
    return 1.0625 * 1.1102230246251565e-16 * ((double)N * s2 + 4.0 * mass + 4.0 + (double)L / (double)N);
}

// cv::RNG (multiply-with-carry) as used by the degenerate fallback (src/pf2DRao.cpp:179-192)
struct mkf_cvrng {
    uint64_t state;
    __device__ __forceinline__ explicit mkf_cvrng(uint64_t s) : state(s ? s : 0xffffffffull) {}
    __device__ __forceinline__ unsigned next()
    {
        state = (uint64_t)(unsigned)state * 4164903690u + (unsigned)(state >> 32);
        return (unsigned)state;
    }
    __device__ __forceinline__ int uniform_int(int a, int b) { return a == b ? a : (int)(next() % (unsigned)(b - a) + a); }
    __device__ __forceinline__ double uniform_dbl()
    {
        unsigned t = next();
        return __dmul_rn((double)(((uint64_t)t << 32) | next()), 5.4210108624275221700372640043497e-20);
    }
};

// the literal sequential loop (bit-exact by construction); w(i) yields the i-th weight
// (out(i, idx) stores the i-th result)
template <class WF, class OF>
__device__ __forceinline__ void mkf_resample_sequential(WF w, int L, int N, double u, OF out)
{
    int idx = 0;
    const double step = __ddiv_rn(1.0, (double)N);
    double beta = __dmul_rn(u, step);
    double wi = w(0);
    for (int i = 0; i < N; i++) {
        while (beta > wi) {
            beta = __dsub_rn(beta, wi);
            idx = (idx + 1 == L) ? 0 : idx + 1; // (idx + 1) % L
            wi = w(idx);
        }
        beta = __dadd_rn(beta, step);
        out(i, idx);
    }
}

// The same loop run by a whole warp for one track: the lanes stage the (normalised) weights through shared memory
// CH at a time, lane 0 walks them.  The arithmetic and its order are those of mkf_resample_sequential; only the
// latency of the dependent weight loads changes (a flagged 65 536-slot track took 43 ms with one thread reading
// global memory).  `chunk` holds CH doubles private to the warp.
template <int CH, class WF, class OF>
__device__ __forceinline__ void mkf_resample_sequential_warp(WF w, int L, int N, double u, OF out, double* chunk)
{
    const int lane = threadIdx.x & 31;
    const double step = __ddiv_rn(1.0, (double)N);
    double beta = __dmul_rn(u, step);
    int idx = 0, i = 0, base = 0;
    for (;;) {
        for (int q = lane; q < CH; q += 32) chunk[q] = (base + q < L) ? w(base + q) : 0.0;
        __syncwarp();
        if (lane == 0) {
            double wi = chunk[idx - base];
            while (i < N) {
                if (beta > wi) {
                    beta = __dsub_rn(beta, wi);
                    idx = (idx + 1 == L) ? 0 : idx + 1;
                    if (idx < base || idx >= base + CH) break; // next weight lives in another chunk
                    wi = chunk[idx - base];
                } else {
                    beta = __dadd_rn(beta, step);
                    out(i++, idx);
                }
            }
        }
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= N) break;
        idx = __shfl_sync(0xffffffffu, idx, 0);
        base = idx / CH * CH;
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// TMA 1-D bulk copy global -> shared with mbarrier completion (sm_90+/sm_100a)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mkf_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mkf_mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mkf_smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mkf_mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mkf_smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mkf_tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     mkf_smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(mkf_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mkf_mbar_wait(uint64_t* bar, uint32_t phase)
{
    uint32_t ok = 0;
    const uint32_t a = mkf_smem_u32(bar);
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, "
                     "p;\n\t}"
                     : "=r"(ok)
                     : "r"(a), "r"(phase)
                     : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// programmatic dependent launch (PDL): every kernel of the per-frame chain lets its successor's CTAs be scheduled as
// soon as all of its own CTAs have started (launch_dependents, first thing) and blocks until the predecessor grid has
// completed and its writes are visible (wait) before it touches anything that grid produced -- or writes anything
// that grid may still read.  Without the launch attribute both are no-ops.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mkf_pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void mkf_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// warp / block collectives
// ---------------------------------------------------------------------------------------------
// block-wide exclusive scan of one double per thread (plain IEEE adds); also returns the block total.
// Rounding depth on the way to any prefix: 5 (warp scan) + BT/32 - 1 (<= 3 warp totals, BT <= 128) or + 5 (warp
// totals scanned by shuffles, BT > 128) + 1, i.e. <= 11 for every BT -- k_resample_block's tolerance relies on it.
template <int BT>
__device__ __forceinline__ double mkf_block_excl_scan_d(double v, double* scratch, double& total)
{
    constexpr int NW = BT / 32;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) scratch[wid] = inc;
    __syncthreads();
    double pre = 0.0, tot = 0.0;
    if constexpr (NW <= 4) {
#pragma unroll
        for (int w = 0; w < NW; w++) {
            const double s = scratch[w];
            if (w < wid) pre += s;
            tot += s;
        }
        __syncthreads();
    } else {
        if (wid == 0) {
            double s = lane < NW ? scratch[lane] : 0.0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double n = __shfl_up_sync(0xffffffffu, s, o);
                if (lane >= o) s += n;
            }
            if (lane < NW) scratch[lane] = s; // inclusive scan of the warp totals
        }
        __syncthreads();
        pre = wid > 0 ? scratch[wid - 1] : 0.0;
        tot = scratch[NW - 1];
        __syncthreads();
    }
    total = tot;
    double prev = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) prev = 0.0;
    return pre + prev;
}

// block-wide exclusive scan in double-double (the rare second-opinion pass of k_resample_block)
template <int BT>
__device__ __forceinline__ dd mkf_block_excl_scan_dd(dd v, double* scratch_hi, double* scratch_lo, dd& total)
{
    constexpr int NW = BT / 32;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    dd inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        dd n;
        n.hi = __shfl_up_sync(0xffffffffu, inc.hi, o);
        n.lo = __shfl_up_sync(0xffffffffu, inc.lo, o);
        if (lane >= o) inc = dd_add(n, inc);
    }
    if (lane == 31) {
        scratch_hi[wid] = inc.hi;
        scratch_lo[wid] = inc.lo;
    }
    __syncthreads();
    dd pre = dd_make(0.0), tot = dd_make(0.0);
    for (int w = 0; w < NW; w++) {
        dd sw;
        sw.hi = scratch_hi[w];
        sw.lo = scratch_lo[w];
        if (w < wid) pre = dd_add(pre, sw);
        tot = dd_add(tot, sw);
    }
    __syncthreads();
    total = tot;
    dd prev;
    prev.hi = __shfl_up_sync(0xffffffffu, inc.hi, 1);
    prev.lo = __shfl_up_sync(0xffffffffu, inc.lo, 1);
    if (lane == 0) prev = dd_make(0.0);
    return dd_add(pre, prev);
}

// block-wide exclusive max-scan of one int per thread (identity -1); returns block max in total
template <int BT>
__device__ __forceinline__ int mkf_block_excl_scan_max(int v, int* scratch, int& total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc = max(inc, n);
    }
    if (lane == 31) scratch[wid] = inc;
    __syncthreads();
    int pre = -1, tot = -1;
#pragma unroll
    for (int w = 0; w < BT / 32; w++) {
        int s = scratch[w];
        if (w < wid) pre = max(pre, s);
        tot = max(tot, s);
    }
    __syncthreads();
    total = tot;
    int prev = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) prev = -1;
    return max(pre, prev);
}

#endif
