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

template <int D>
struct HeadsTmaLay {
    using L = SlotLay<D>;
    static constexpr int WARPS = 4;
    static constexpr int RB = L::NP * 16;            // bytes per record
    static constexpr int STAGE = 32 * RB;            // one step of records
    static constexpr int MEAS = 32 * 8 * 8;          // 8 doubles per lane (6 used)
    static constexpr int WARP_BYTES = 2 * STAGE + MEAS + 16; // input stage | output stage | measurements | mbarrier
    __host__ __device__ static constexpr size_t cst_bytes(int K) { return ((size_t)K * L::CS * 8 + 127) / 128 * 128; }
    __host__ __device__ static constexpr size_t smem_bytes(int K) { return cst_bytes(K) + (size_t)WARPS * WARP_BYTES; }
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

// Fetch of one step into the warp's input stage (every lane enters).  rec: the lane's list entry, valid: it has one
// (the valid lanes of a step are a prefix).  Lanes that share their predecessor's parent record share its slot of the
// stage; a stretch of consecutive parent records is one bulk copy.  Returns the lane's slot.
template <int D>
__device__ __forceinline__ int mkf_heads_fetch(const SlotArgs& a, const double2* __restrict__ st_in, const int4 rec,
                                               const bool valid, double2* in_st, double* ms, uint64_t* fbar,
                                               const int lane)
{
    using L = SlotLay<D>;
    constexpr unsigned FULL = 0xffffffffu;
    const int s = rec.x;
    const int sprev = __shfl_up_sync(FULL, s, 1);
    const bool isnew = valid && (lane == 0 || s != sprev);
    const bool segst = isnew && (lane == 0 || s != sprev + 1);
    const unsigned mnew = __ballot_sync(FULL, isnew);
    unsigned mseg = __ballot_sync(FULL, segst);
    const int slot = __popc(mnew & (FULL >> (31 - lane))) - 1;
    const int total = __popc(mnew);
    if (valid) { // the head's measurement column (raw: BH is subtracted after the wait)
        const long long t = rec.z;
        double* md = ms + lane * 8;
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
    mkf_cp_async_arrive(fbar);
    if (lane == 0) mkf_mbar_expect_tx(fbar, (uint32_t)(total * HeadsTmaLay<D>::RB));
    while (mseg) {
        const int l0 = __ffs(mseg) - 1;
        mseg &= mseg - 1;
        const int l1 = mseg ? __ffs(mseg) - 1 : 0;
        const int s0 = __shfl_sync(FULL, s, l0);
        const int sl0 = __shfl_sync(FULL, slot, l0);
        const int sl1n = __shfl_sync(FULL, slot, l1);
        const int sl1 = mseg ? sl1n : total;
        if (lane == 0)
            mkf_tma_load_1d(in_st + sl0 * L::NP, st_in + (long long)(unsigned)s0 * L::NP,
                            (uint32_t)((sl1 - sl0) * HeadsTmaLay<D>::RB), fbar);
    }
    return slot;
}

template <int D>
__global__ void __launch_bounds__(128, 1) k_slot_update_heads_tma(const SlotArgs a, int* __restrict__ count_to_clear)
{
    using L = SlotLay<D>;
    using H = HeadsTmaLay<D>;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t cbar;
    double* cst = reinterpret_cast<double*>(smem_raw); // K x CS model constants (TMA)

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    unsigned char* wb = smem_raw + H::cst_bytes(a.K) + (size_t)wid * H::WARP_BYTES;
    double2* in_st = reinterpret_cast<double2*>(wb);
    double2* out_st = reinterpret_cast<double2*>(wb + H::STAGE);
    double* ms = reinterpret_cast<double*>(wb + 2 * H::STAGE);
    uint64_t* fbar = reinterpret_cast<uint64_t*>(wb + 2 * H::STAGE + H::MEAS);

    const uint32_t cbytes = (uint32_t)(a.K * L::CS * sizeof(double));
    if (tid == 0) mkf_mbar_init(&cbar, 1);
    if (lane == 0) mkf_mbar_init(fbar, 33); // 32 lanes' cp.async arrivals + lane 0's expect_tx
    __syncthreads();
    if (tid == 0) {
        mkf_mbar_expect_tx(&cbar, cbytes);
        mkf_tma_load_1d(cst, a.comp_const, cbytes, &cbar); // model constants: never written by the frame chain
    }
    mkf_pdl_launch_dependents();
    mkf_pdl_wait();

    const int n = *reinterpret_cast<const volatile int*>(a.head_count);
    if (blockIdx.x == 0 && tid == 0) *count_to_clear = 0; // the counter the next frame's k_frame_heads appends with
    if (a.ts && tid == 0) atomicMin(a.ts, mkf_globaltimer());
    const int S = (n + 31) >> 5;                 // steps of 32 heads
    const int G = (int)gridDim.x * H::WARPS;     // warps in the grid
    int step = (int)blockIdx.x * H::WARPS + wid; // this warp takes steps step, step + G, ...
    const int4 none = make_int4(-1, 0, 0, 0);
    int4 rec = none, rec_n = none;
    if (step < S && step * 32 + lane < n) rec = __ldg(a.hd16 + step * 32 + lane);
    int slot = 0;
    if (step < S) slot = mkf_heads_fetch<D>(a, a.st_in, rec, rec.x >= 0, in_st, ms, fbar, lane);
    if (step + G < S && (step + G) * 32 + lane < n) rec_n = __ldg(a.hd16 + (step + G) * 32 + lane);
    mkf_mbar_wait(&cbar, 0);

    uint32_t phase = 0;
    while (step < S) {
        const bool valid = rec.x >= 0;
        mkf_mbar_wait(fbar, phase);
        phase ^= 1;
        double v[L::NE];
        double zc[MKF_M];
        int dep = 0;
        if (valid) {
            // (volatile LDS like the record's below, in front of them: the refill must not overtake these either)
            const double2* md2 = reinterpret_cast<const double2*>(ms + lane * 8);
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
            // the stage may only be refilled once every LDS above has read it: the fetch below takes its base pointer
            // through a select on the last value loaded (both arms hold the same pointer), so it cannot issue earlier
            dep = __double2hiint(v[L::NE - 1]) == 0x7ff7a5a5 ? 1 : 0;
        }
        // next step: fetch into the input stage, entry of the step after it
        const int step_n = step + G;
        int slot_n = 0;
        if (step_n < S) {
            const double2* base = __shfl_sync(FULL, dep, 0) ? a.st_in_alias : a.st_in;
            slot_n = mkf_heads_fetch<D>(a, base, rec_n, rec_n.x >= 0, in_st, ms, fbar, lane);
        }
        int4 rec_nn = none;
        if (step_n + G < S && (step_n + G) * 32 + lane < n) rec_nn = __ldg(a.hd16 + (step_n + G) * 32 + lane);

        double w = 0.0;
        if (valid) {
            const bool ok = slot_math<D, false>(v, cst + (rec.w & 0xff) * L::CS, zc, a.r, a.chol_mode, a.stage, w);
            if (!ok) atomicOr(a.status + rec.z, MKF_ST_CHOL_FAIL);
        }
        // children -> output stage -> one bulk store per stretch of consecutive destination records
        if (lane == 0) mkf_bulk_wait_read(); // the previous step's stores have read the stage
        __syncwarp();
        if (valid) {
            double2* dst = out_st + lane * L::NP;
#pragma unroll
            for (int p = 0; p < L::NP; p++) {
                double2 q;
                q.x = v[2 * p];
                q.y = (2 * p + 1 < L::NE) ? v[2 * p + 1] : 0.0;
                dst[p] = q;
            }
            a.w_rec[(unsigned)rec.y] = w;
        }
        mkf_fence_async_smem();
        __syncwarp();
        {
            const int dq = rec.y;
            const int dprev = __shfl_up_sync(FULL, dq, 1);
            unsigned mst = __ballot_sync(FULL, valid && (lane == 0 || dq != dprev + 1));
            const int nvalid = __popc(__ballot_sync(FULL, valid));
            while (mst) {
                const int l0 = __ffs(mst) - 1;
                mst &= mst - 1;
                const int l1 = mst ? __ffs(mst) - 1 : nvalid;
                const int d0 = __shfl_sync(FULL, dq, l0);
                if (lane == 0)
                    mkf_bulk_store(a.st_out + (long long)(unsigned)d0 * L::NP, out_st + l0 * L::NP,
                                   (uint32_t)((l1 - l0) * H::RB));
            }
            if (lane == 0) mkf_bulk_commit();
        }
        rec = rec_n;
        rec_n = rec_nn;
        slot = slot_n;
        step = step_n;
    }
    if (lane == 0) mkf_bulk_wait_all(); // shared memory must outlive the last stores; the grid's end publishes them
    if (a.ts && lane == 0) atomicMax(a.ts + 1, mkf_globaltimer());
}

// Layout conversion of the live records of every track between the tile layout of the per-slot kernels
// (tile[pair][lane]) and the contiguous records of k_slot_update_heads_tma, out of place (into the idle ping-pong
// buffer).  count: records per track (null: all N).  One CTA per track.
template <int D>
__global__ void __launch_bounds__(128) k_relayout(const double2* __restrict__ in, double2* __restrict__ out,
                                                  const int* __restrict__ count, int N, int to_aos)
{
    using L = SlotLay<D>;
    const long long t = blockIdx.x;
    const int cnt = count ? min(count[t], N) : N;
    for (int idx = threadIdx.x; idx < cnt * L::NP; idx += blockDim.x) {
        const int i = idx / L::NP, p = idx - i * L::NP;
        const long long sp = t * N + i;
        const long long tile_off = (sp >> 5) * (long long)L::TILE2 + (sp & 31) * L::H + L::po(p);
        const long long aos_off = sp * L::NP + p;
        if (to_aos)
            out[aos_off] = in[tile_off];
        else
            out[tile_off] = in[aos_off];
    }
}

#endif
