// mkf_heads_tma.cuh -- the slot update of the run-length pipeline's heads with the records staged through shared
// memory by TMA bulk copies (KF_model::predict/update + mvnpdf of src/KF_model.cpp:11-25, src/pf2DRao.cpp:34-67,138:
// the same slot_math as every other slot kernel, so the results are bit-identical).
//
// Why: k_slot_update_heads_direct (one step of 32 heads per warp: entry -> 45 gathers -> ~1 300 dependent
// instructions -> 45 stores) keeps 8 warps per SM (255 registers) and each spends more than half its life waiting for
// its own gather (ncu: 6.7 of 12.1 cycles per issue on the long scoreboard, another 1.5 in the LSU queue) -- the
// kernel's rate is (warps per SM) / (memory wait + arithmetic), the same whether the records come from DRAM or from
// L2 (profiles/r02_l2_sweep.jsonl).  Here the memory wait leaves the warp's critical path:
//   * records live in a CONTIGUOUS layout (720 bytes each, record r at r * 720) in both ping-pong buffers while the
//     batch is in run-length mode (k_relayout converts when it enters / leaves), so the parents of a step -- sorted and
//     densely packed per track -- are a handful of contiguous stretches: one `cp.async.bulk` per stretch (SASS UBLKCP)
//     lands them in the warp's input stage, completion on an mbarrier; the step's measurement columns arrive next to
//     them by 8-byte cp.async (LDGSTS) counted on the same mbarrier;
//   * a persistent warp (4 per SM, one CTA per SM) reads its step from the stage into registers (LDS.128 at a stride of
//     45 x 16 bytes: conflict-free), immediately issues the fetch of its NEXT step into the same stage, computes, writes
//     the children to its output stage and hands that to one bulk store per stretch of consecutive destination records;
//     the list entries are read two steps ahead.
// A warp never waits for a global load it issued itself, issues no LDG / STG for the state at all, and the SM has
// 4 x 23 KB of reads in flight during the arithmetic.
#ifndef MKF_HEADS_TMA_CUH
#define MKF_HEADS_TMA_CUH

#include "mkf_kernels.cuh"

template <int D, int W, int O>
struct HeadsTmaLay {
    using L = SlotLay<D>;
    static constexpr int WARPS = W;                  // consumer warps per CTA (one CTA per SM)
    static constexpr int OUTS = O;                   // output stages, shared by the warps (held for ~10 % of a step)
    static constexpr int RB = L::NP * 16;            // bytes per record
    static constexpr int STAGE = 32 * RB;            // one step of records
    static constexpr int MEAS = 32 * 6 * 8;          // 6 doubles per lane
    static constexpr int WARP_BYTES = STAGE + MEAS + 128 + 16; // input stage | measurements | slots | full, empty barriers
    static constexpr int XS = (D / 2) * 32 * 16;     // the step's means once more, [pair][lane] (SlotArgs::xs)
    static constexpr int OUT_BYTES = STAGE + XS;     // output stage: the children's records | their means
    __host__ __device__ static constexpr size_t cst_bytes(int K) { return ((size_t)K * L::CS * 8 + 127) / 128 * 128; }
    __host__ __device__ static constexpr size_t smem_bytes(int K)
    {
        return cst_bytes(K) + (size_t)WARPS * WARP_BYTES + (size_t)OUTS * OUT_BYTES + 16 * OUTS;
    }
};

__device__ __forceinline__ void mkf_bulk_store(void* gmem_dst, const void* smem_src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(mkf_smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mkf_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void mkf_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void mkf_bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void mkf_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mkf_cp_async8(void* smem_dst, const void* gmem_src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(mkf_smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
// arrive on `bar` once every cp.async this thread issued so far has landed (the barrier's count includes the lane)
__device__ __forceinline__ void mkf_cp_async_arrive(uint64_t* bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(mkf_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ double2 mkf_lds128(const double2* p)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(mkf_smem_u32(p)));
    return v;
}

// MKF_TMA_PROF builds (csrc/Makefile `variant`) accumulate the clock cycles lane 0 of every warp spends per phase
#ifdef MKF_TMA_PROF
__device__ unsigned long long g_tma_prof[8];
#define MKF_TP_DECL long long tp_t = clock64(), tp_acc[7] = {0, 0, 0, 0, 0, 0, 0}
#define MKF_TP(i)                                                                                                     \
    do {                                                                                                              \
        const long long tp_n = clock64();                                                                             \
        tp_acc[i] += tp_n - tp_t;                                                                                     \
        tp_t = tp_n;                                                                                                  \
    } while (0)
#define MKF_TP_FLUSH(steps)                                                                                           \
    if (lane == 0) {                                                                                                  \
        for (int i_ = 0; i_ < 7; i_++) atomicAdd(&g_tma_prof[i_], (unsigned long long)tp_acc[i_]);                   \
        atomicAdd(&g_tma_prof[7], (unsigned long long)(steps));                                                      \
    }
#else
#define MKF_TP_DECL
#define MKF_TP(i)
#define MKF_TP_FLUSH(steps)
#endif

// Fetch of one step into the warp's input stage (every lane enters).  rec: the lane's list entry, valid: it has one
// (the valid lanes of a step are a prefix).  The parents of a step are sorted and nearly contiguous in the previous list
// (a parent without children leaves a gap, a run cut at a component boundary repeats its parent), and issuing a bulk
// copy costs a warp ~350 cycles whatever its size: parents at most MKF_TMA_GAP records apart are fetched by ONE copy,
// gaps included; when the spans would not fit the 32-record stage the step is packed exactly (one copy per stretch of
// consecutive parents).  Returns the lane's slot of the stage.
#ifndef MKF_TMA_GAP
#define MKF_TMA_GAP 6
#endif
template <int D>
__device__ __forceinline__ int mkf_heads_fetch(const SlotArgs& a, const double2* __restrict__ st_in, const int4 rec,
                                               const bool valid, double2* in_st, double* ms, uint64_t* fbar,
                                               const int lane, int* slots_out = nullptr)
{
    using L = SlotLay<D>;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int RB = L::NP * 16;
    const unsigned le = FULL >> (31 - lane); // lanes <= this one
    const int s = rec.x;
    const int sprev = __shfl_up_sync(FULL, s, 1);
    const int nvalid = __popc(__ballot_sync(FULL, valid));
    const int gap = s - sprev - 1; // records skipped since the previous lane's parent (-1: the same parent)
    int slot = 0, total = 0, x = 0, excl = 0;
    unsigned mseg = 0;
    for (int gmax = MKF_TMA_GAP;; gmax = 0) {
        const bool segst = valid && (lane == 0 || gap < -1 || gap > gmax);
        mseg = __ballot_sync(FULL, segst);
        const int l0 = 31 - __clz(mseg & le);                        // first lane of this lane's stretch
        const unsigned above = mseg & ~le;
        const int lend = above ? __ffs(above) - 2 : nvalid - 1;      // its last lane
        const int s0 = __shfl_sync(FULL, s, l0 < 0 ? 0 : l0);
        const int send = __shfl_sync(FULL, s, lend < 0 ? 0 : lend);
        x = segst ? send - s + 1 : 0;                                // records of the stretch, at its first lane
        int inc = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int nb = __shfl_up_sync(FULL, inc, o);
            if (lane >= o) inc += nb;
        }
        excl = inc - x;
        total = __shfl_sync(FULL, inc, 31);
        slot = __shfl_sync(FULL, excl, l0 < 0 ? 0 : l0) + (s - s0);
        if (total <= 32 || gmax == 0) break;
    }
    if (valid) { // the head's measurement column (raw: BH is subtracted after the wait)
        const long long t = rec.z;
        double* md = ms + lane * 6;
        if (a.meas_layout == MKF_MEAS_CAND) {
            const int bsel = rec.w >> 8;
            const double* __restrict__ px = a.cand + (t * 2 + a.hand) * 2 * (long long)a.cand_C;
#pragma unroll
            for (int r = 0; r < 4; r++) mkf_cp_async8(md + r, a.roi + t * 4 + r);
            mkf_cp_async8(md + 4, px + bsel);
            mkf_cp_async8(md + 5, px + a.cand_C + bsel);
        } else {
#pragma unroll
            for (int r = 0; r < MKF_M; r++) mkf_cp_async8(md + r, a.meas + t * MKF_M + r);
        }
    }
    if (slots_out) { // (a producer warp fetching for a consumer: the consumer reads its slot after the barrier's wait;
        slots_out[lane] = slot; // lane 0's arrive below releases these stores)
        __syncwarp();
    }
    mkf_cp_async_arrive(fbar);
    if (lane == 0) mkf_mbar_expect_tx(fbar, (uint32_t)(total * RB));
    while (mseg) {
        const int l0 = __ffs(mseg) - 1;
        mseg &= mseg - 1;
        const int s0 = __shfl_sync(FULL, s, l0);
        const int n0 = __shfl_sync(FULL, x, l0);
        const int b0 = __shfl_sync(FULL, excl, l0);
        if (lane == 0)
            mkf_tma_load_1d(in_st + b0 * L::NP, st_in + (long long)(unsigned)s0 * L::NP, (uint32_t)(n0 * RB), fbar);
    }
    return slot;
}

__device__ __forceinline__ void mkf_mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mkf_smem_u32(bar)) : "memory");
}

// W consumer warps, O output stages, P producer warps (P == 0: every consumer fetches its own next step).
// With producers, consumer c is served by producer c % P: the producer reads the list entries, plans and issues the
// fetch of the consumer's next step (mkf_heads_fetch: ~2 000 cycles of shuffles, cp.async and bulk-copy issue that a
// consumer would otherwise spend between two Gaussians) as soon as the consumer has signalled on its `empty` barrier
// that the stage has been read, and leaves every lane's slot of the stage next to it.
template <int D, int W, int O, int P>
__global__ void __launch_bounds__(32 * (W + P), 1) k_slot_update_heads_tma(const SlotArgs a, int* __restrict__ count_to_clear)
{
    using L = SlotLay<D>;
    using H = HeadsTmaLay<D, W, O>;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t cbar;
    double* cst = reinterpret_cast<double*>(smem_raw); // K x CS model constants (TMA)

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    unsigned char* const wbase = smem_raw + H::cst_bytes(a.K);
    auto in_of = [&](int c) { return reinterpret_cast<double2*>(wbase + (size_t)c * H::WARP_BYTES); };
    auto ms_of = [&](int c) { return reinterpret_cast<double*>(wbase + (size_t)c * H::WARP_BYTES + H::STAGE); };
    auto slots_of = [&](int c) { return reinterpret_cast<int*>(wbase + (size_t)c * H::WARP_BYTES + H::STAGE + H::MEAS); };
    auto fbar_of = [&](int c) {
        return reinterpret_cast<uint64_t*>(wbase + (size_t)c * H::WARP_BYTES + H::STAGE + H::MEAS + 128);
    };
    auto ebar_of = [&](int c) { return fbar_of(c) + 1; };
    unsigned char* ob = wbase + (size_t)W * H::WARP_BYTES;

    const uint32_t cbytes = (uint32_t)(a.K * L::CS * sizeof(double));
    if (tid == 0) mkf_mbar_init(&cbar, 1);
    if (wid < W && lane == 0) {
        mkf_mbar_init(fbar_of(wid), 33); // 32 lanes' cp.async arrivals + lane 0's expect_tx
        mkf_mbar_init(ebar_of(wid), 1);  // the consumer's "stage read"
        if (wid < O) *reinterpret_cast<int*>(ob + (size_t)O * H::OUT_BYTES + 16 * wid) = 0;
    }
    __syncthreads();
    if (tid == 0) {
        mkf_mbar_expect_tx(&cbar, cbytes);
        mkf_tma_load_1d(cst, a.comp_const, cbytes, &cbar); // model constants: never written by the frame chain
    }
    // Programmatic dependent launch: the dependents (k_runs_repair, and through it k_resample_runs) are released when a
    // warp of every CTA has finished its steps, not at the start.  Released at the start, their CTAs take their places
    // on the SMs during this kernel and it runs 63 us instead of 53 (device timeline of the pipelined loop, 4096 x 500:
    // tools/tma_timeline.py); released late, the hand-over costs ~2 us more and the frame is shorter.
    if (!a.pdl_late) mkf_pdl_launch_dependents();
    mkf_pdl_wait();

    const int n = *reinterpret_cast<const volatile int*>(a.head_count);
    if (blockIdx.x == 0 && tid == 0) *count_to_clear = 0; // the counter the next frame's k_frame_heads appends with
    if (a.ts && tid == 0) atomicMin(a.ts, mkf_globaltimer());
    MKF_TL_START(1, a.dbg_frame);
    const int S = (n + 31) >> 5;      // steps of 32 heads
    const int G = (int)gridDim.x * W; // consumer warps in the grid
    const int4 none = make_int4(-1, 0, 0, 0);
    auto entry = [&](int st) {
        int4 r = none;
        if (st < S && st * 32 + lane < n) r = __ldg(a.hd16 + st * 32 + lane);
        return r;
    };

    if (P > 0 && wid >= W) {
        // ---------------- producer: the fetches of consumers wid - W, wid - W + P, ... in step order ----------------
        const int p0 = wid - W;
        int c = p0, k = 0;                       // the next (consumer, round) to serve
        int step = (int)blockIdx.x * W + c;
        int4 rec = entry(step);
        while (step < S) {
            // the one after it (its entries are in flight while this one is planned)
            int c2 = c + P, k2 = k;
            if (c2 >= W) {
                c2 = p0;
                k2 = k + 1;
            }
            const int step2 = (int)blockIdx.x * W + c2 + k2 * G;
            const int4 rec2 = entry(step2);
            if (k > 0) mkf_mbar_wait(ebar_of(c), (uint32_t)((k - 1) & 1)); // consumer c has read step k - 1 out of its stage
            mkf_heads_fetch<D>(a, a.st_in, rec, rec.x >= 0, in_of(c), ms_of(c), fbar_of(c), lane, slots_of(c));
            // (consumers of this producer whose steps ran out are skipped: steps only grow with c and k)
            c = c2;
            k = k2;
            step = step2;
            rec = rec2;
            if (step >= S && c != p0) { // a later consumer of this round has no step left; neither has anyone after it
                break;
            }
        }
        if (a.pdl_late) mkf_pdl_launch_dependents();
        MKF_TL_END(1, a.dbg_frame);
        return;
    }

    // ---------------- consumer ----------------
    double2* in_st = in_of(wid);
    double* ms = ms_of(wid);
    uint64_t* fbar = fbar_of(wid);
    uint64_t* ebar = ebar_of(wid);
    // (== ebar: st_in_alias == st_in; the select below makes the arrive wait for the stage's last LDS)
    uint64_t* ebar_alias = ebar + (a.st_in_alias != a.st_in ? 1 : 0);
    double2* out_st = reinterpret_cast<double2*>(ob + (size_t)(wid % O) * H::OUT_BYTES); // shared by the warps wid % O
    double2* xs_st = out_st + 32 * L::NP;
    int* lock = reinterpret_cast<int*>(ob + (size_t)O * H::OUT_BYTES + 16 * (wid % O));
    int step = (int)blockIdx.x * W + wid; // this warp takes steps step, step + G, ...
    int4 rec = entry(step), rec_n = none;
    int slot = 0;
    if (P == 0 && step < S) slot = mkf_heads_fetch<D>(a, a.st_in, rec, rec.x >= 0, in_st, ms, fbar, lane);
    rec_n = entry(step + G);
    mkf_mbar_wait(&cbar, 0);

    uint32_t phase = 0;
    MKF_TP_DECL;
    int nsteps = 0;
    while (step < S) {
        const bool valid = rec.x >= 0;
        nsteps++;
        mkf_mbar_wait(fbar, phase);
        phase ^= 1;
        MKF_TP(0); // waiting for the input stage
        double v[L::NE];
        double zc[MKF_M];
        int dep = 0;
        if (P > 0) slot = *reinterpret_cast<volatile int*>(slots_of(wid) + lane);
        if (valid) {
            // (volatile LDS like the record's below, in front of them: the refill must not overtake these either)
            const double2* md2 = reinterpret_cast<const double2*>(ms + lane * 6);
            const double2 m01 = mkf_lds128(md2), m23 = mkf_lds128(md2 + 1), m45 = mkf_lds128(md2 + 2);
            if (a.meas_layout == MKF_MEAS_CAND) { // same operations as mkf_load_meas_cand
                const double rx = m01.x, ry = m01.y, rw = m23.x, rh = m23.y;
                const double cxv = __dadd_rn(rx, __ddiv_rn(rw, 2.0));
                zc[0] = cxv - a.bh[0];
                zc[1] = __dadd_rn(ry, __dmul_rn(0.5, rh)) - a.bh[1];
                zc[2] = m45.x - a.bh[2];
                zc[3] = m45.y - a.bh[3];
                zc[4] = cxv - a.bh[4];
                zc[5] = __dadd_rn(ry, __dmul_rn(a.neck, rh)) - a.bh[5];
            } else {
                zc[0] = m01.x - a.bh[0];
                zc[1] = m01.y - a.bh[1];
                zc[2] = m23.x - a.bh[2];
                zc[3] = m23.y - a.bh[3];
                zc[4] = m45.x - a.bh[4];
                zc[5] = m45.y - a.bh[5];
            }
            const double2* src = in_st + slot * L::NP;
#pragma unroll
            for (int p = 0; p < L::NP; p++) {
                const double2 q = mkf_lds128(src + p);
                v[2 * p] = q.x;
                if (2 * p + 1 < L::NE) v[2 * p + 1] = q.y;
            }
            // the stage may only be refilled once every LDS above has read it: the hand-over below goes through a
            // select on the last value loaded (both arms are the same address), so it cannot issue earlier
            dep = __double2hiint(v[L::NE - 1]) == 0x7ff7a5a5 ? 1 : 0;
        }
#ifdef MKF_TMA_PROF
        if (__shfl_sync(FULL, dep, 0) == 2) break; // (never: makes the stamp below wait for the LDS)
#endif
        MKF_TP(6); // stage -> registers
        // next step: fetch into the input stage (or tell the producer it may), entry of the step after it
        const int step_n = step + G;
        int slot_n = 0;
        if (P > 0) {
            if (lane == 0) mkf_mbar_arrive(dep ? ebar_alias : ebar);
        } else if (step_n < S) {
            const double2* base = __shfl_sync(FULL, dep, 0) ? a.st_in_alias : a.st_in;
            slot_n = mkf_heads_fetch<D>(a, base, rec_n, rec_n.x >= 0, in_st, ms, fbar, lane);
        }
        const int4 rec_nn = entry(step_n + G);
        MKF_TP(1); // next fetch issued

        double w = 0.0;
        if (valid) {
            const bool ok = slot_math<D, false>(v, cst + (rec.w & 0xff) * L::CS, zc, a.r, a.chol_mode, a.stage, w);
            if (!ok) atomicOr(a.status + rec.z, MKF_ST_CHOL_FAIL);
        }
        MKF_TP(2); // arithmetic
        // children -> an output stage -> one bulk store (the step's children are consecutive records of the new list)
        if (lane == 0) {
            if (O < W) {
                while (atomicCAS(lock, 0, 1) != 0) __nanosleep(32);
                __threadfence_block();
            } else
                mkf_bulk_wait_read(); // a private stage: the previous step's store has read it
        }
        __syncwarp();
        MKF_TP(3); // waiting for an output stage
        if (valid) {
            double2* dst = out_st + lane * L::NP;
#pragma unroll
            for (int p = 0; p < L::NP; p++) {
                double2 q;
                q.x = v[2 * p];
                q.y = (2 * p + 1 < L::NE) ? v[2 * p + 1] : 0.0;
                dst[p] = q;
            }
            if (a.xs) {
#pragma unroll
                for (int p = 0; p < D / 2; p++) xs_st[32 * p + lane] = make_double2(v[2 * p], v[2 * p + 1]);
            }
        }
        mkf_fence_async_smem();
        __syncwarp();
        {
            const int nvalid = __popc(__ballot_sync(FULL, valid));
            MKF_TP(4); // registers -> output stage
            if (lane == 0) {
                mkf_bulk_store(a.st_out + (long long)step * 32 * L::NP, out_st, (uint32_t)(nvalid * H::RB));
                // (the means as a second bulk store: six generic STG.128 per lane here cost the step ~1 300 cycles)
                if (a.xs) mkf_bulk_store(a.xs + (long long)step * (D / 2 * 32), xs_st, (uint32_t)H::XS);
                mkf_bulk_commit();
                if (O < W) { // hand the output stage back as soon as the store has read it
                    mkf_bulk_wait_read();
                    __threadfence_block();
                    atomicExch(lock, 0);
                }
            }
        }
        if (valid) { // (after the proxy fence, which would wait for these stores)
            a.w_rec[(unsigned)rec.y] = w;
        }
        MKF_TP(5); // (shared stage) waiting for the store to have read it
        rec = rec_n;
        rec_n = rec_nn;
        slot = slot_n;
        step = step_n;
    }
    MKF_TP_FLUSH(nsteps);
    if (a.pdl_late) mkf_pdl_launch_dependents();
    if (lane == 0) mkf_bulk_wait_all(); // shared memory must outlive the last store; the grid's end publishes it
    if (a.ts && lane == 0) atomicMax(a.ts + 1, mkf_globaltimer());
    MKF_TL_END(1, a.dbg_frame);
}

// Layout conversion of the live records of every track between the tile layout of the per-slot kernels
// (tile[pair][lane], record i of track t at t*N + i) and the contiguous records of k_slot_update_heads_tma (record i of
// track t at lbase[t] + i), out of place (into the idle ping-pong buffer).  to_aos: lbase[t] = t*N is written.
// count: records per track (null: all N).  One CTA per track.
template <int D>
__global__ void __launch_bounds__(128) k_relayout(const double2* __restrict__ in, double2* __restrict__ out,
                                                  const int* __restrict__ count, int N, int to_aos,
                                                  int* __restrict__ lbase)
{
    using L = SlotLay<D>;
    const long long t = blockIdx.x;
    const int cnt = count ? min(count[t], N) : N;
    const long long abase = to_aos ? t * N : (long long)lbase[t];
    if (to_aos && threadIdx.x == 0) lbase[t] = (int)(t * N);
    for (int idx = threadIdx.x; idx < cnt * L::NP; idx += blockDim.x) {
        const int i = idx / L::NP, p = idx - i * L::NP;
        const long long sp = t * N + i;
        const long long tile_off = (sp >> 5) * (long long)L::TILE2 + (sp & 31) * L::H + L::po(p);
        const long long aos_off = (abase + i) * L::NP + p;
        if (to_aos)
            out[aos_off] = in[tile_off];
        else
            out[tile_off] = in[aos_off];
    }
}

#endif
