// mkf_assoc.cuh -- the sample-based association ("PDAF") step of
// PFTracker::getMeasurementProposal (src/pfPose.cpp:238-326) for T persons, included by mkf_api.cu.
//
//   k_assoc_weights   one warp per (person, hand): gate (inside image, L != 0), proposal densities
//                     under both arms (getSampleProb / mvnpdf_multiple, src/pf2DRao.cpp:69-83,105-122),
//                     raw weight L / Z (src/pfPose.cpp:249-292)
//   run_resample      sum + normalise (src/pfPose.cpp:294-298) and C -> N systematic resample (:300-301)
//   k_assoc_meas      per-slot measurement columns (src/pfPose.cpp:303-323)
//   update_device     pf1->update / pf2->update (src/pfPose.cpp:325-326)
#ifndef MKF_ASSOC_CUH
#define MKF_ASSOC_CUH

struct AssocArgs {
    const double* __restrict__ cand_xy; // T x 2 x 2 x C
    const uint8_t* __restrict__ cand_L; // T x 2 x C
    const double* __restrict__ roi;     // T x 4
    const double* __restrict__ pose0;   // T x D0 (arm 0 reconstruction; [0],[1] = hand x,y)
    const double* __restrict__ pose1;   // T x D1
    double* __restrict__ w_raw;         // T x 2 x C
    uint8_t* __restrict__ gate;         // T x 2 x C
    long long T;
    int C, D0, D1, chol_mode, img_rows, img_cols;
    double pa, clutter, spread;
    // zeroed by k_assoc_weights (first kernel of the step; no memset nodes): per (person, hand) candidate-resample
    // flags, and the two arm batches' track status words
    uint32_t* __restrict__ as_status;
    uint32_t* __restrict__ status0;
    uint32_t* __restrict__ status1;
};

// 2-D isotropic proposal density exactly as mvnpdf_multiple evaluates it for cov = s2 * I:
// chol() of a diagonal matrix, 2x2 closed-form inverse, exp(q*-0.5 + (-lsd - log 2pi))
struct Iso2 {
    double ri00, ri11, shift;
};
__device__ __forceinline__ Iso2 mkf_iso2_setup(double s2, int chol_mode)
{
    // cv::Cholesky on diag(s2, s2): diagonal 1/sqrt(s2); chol(): R_ee = 1/elem
    const double inv = __ddiv_rn(1.0, sqrt(s2));
    double R;
    if (chol_mode == MKF_CHOL_CV3_LITERAL)
        R = inv; // OpenCV >= 3 leaves L_ee on the diagonal, so elem = L_ee and R_ee = 1/L_ee
    else if (chol_mode == MKF_CHOL_EXACT)
        R = __ddiv_rn(1.0, inv);
    else
        R = __ddiv_rn(1.0, inv);
    // cv::invert 2x2: d = 1/(R00*R11 - 0); inv00 = R11*d; inv11 = R00*d
    const double d = __ddiv_rn(1.0, __dmul_rn(R, R));
    Iso2 o;
    o.ri00 = __dmul_rn(R, d);
    o.ri11 = __dmul_rn(R, d);
    const double lsd = __dadd_rn(log(R), log(R));
    o.shift = __dsub_rn(-lsd, 1.8378770664093453); // 2*log(2*pi)/2
    return o;
}
__device__ __forceinline__ double mkf_iso2_pdf(const Iso2& g, double x, double y, double ux, double uy)
{
    const double v0 = __dmul_rn(__dsub_rn(x, ux), g.ri00);
    const double v1 = __dmul_rn(__dsub_rn(y, uy), g.ri11);
    const double q = __dadd_rn(__dmul_rn(v0, v0), __dmul_rn(v1, v1));
    return exp(__dadd_rn(__dmul_rn(q, -0.5), g.shift));
}

// one warp per (person, hand, stretch of MKF_ASSOC_SPAN candidates): with the reference's 5000 candidates per hand a
// (person, hand) is 40 warps, not one lane-strided loop of 157 trips (256 persons: 512 warps on 148 SMs)
constexpr int MKF_ASSOC_SPAN = 128;
__global__ void __launch_bounds__(128) k_assoc_weights(const AssocArgs a)
{
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int spans = (a.C + MKF_ASSOC_SPAN - 1) / MKF_ASSOC_SPAN;
    const long long wid = gw / spans;
    const int c_begin = (int)(gw - wid * spans) * MKF_ASSOC_SPAN;
    const int c_end = min(a.C, c_begin + MKF_ASSOC_SPAN);
    if (wid >= a.T * 2) return;
    const long long t = wid >> 1;
    const int h = (int)(wid & 1);
    if (c_begin == 0 && lane == 0) {
        a.as_status[wid] = 0u;
        (h ? a.status1 : a.status0)[t] = 0u;
    }
    const double scale = a.roi[t * 4 + 2]; // (double)msg->ROIs[0].width
    const Iso2 g = mkf_iso2_setup(__dmul_rn(__dmul_rn(a.spread, scale), 1.0), a.chol_mode);
    const double h0x = a.pose0[t * a.D0 + 0], h0y = a.pose0[t * a.D0 + 1];
    const double h1x = a.pose1[t * a.D1 + 0], h1y = a.pose1[t * a.D1 + 1];
    const double ownx = h ? h1x : h0x, owny = h ? h1y : h0y;
    const double othx = h ? h0x : h1x, othy = h ? h0y : h1y;
    const double zc = __dmul_rn(a.clutter, __dsub_rn(1.0, __dmul_rn(2.0, a.pa)));
    const double* __restrict__ px = a.cand_xy + (t * 2 + h) * 2 * (long long)a.C;
    const double* __restrict__ py = px + a.C;
    const uint8_t* __restrict__ pl = a.cand_L + (t * 2 + h) * (long long)a.C;
    double* __restrict__ wo = a.w_raw + (t * 2 + h) * (long long)a.C;
    uint8_t* __restrict__ go = a.gate + (t * 2 + h) * (long long)a.C;
    for (int c = c_begin + lane; c < c_end; c += 32) {
        const double x = px[c], y = py[c];
        double w = 0.0;
        uint8_t gt = 0;
        if ((y > 0) && (y < (double)a.img_rows) && (x > 0) && (x < (double)a.img_cols)) {
            const double Lk = __ddiv_rn((double)pl[c], 255.0);
            if (Lk != 0.0) {
                const double own = mkf_iso2_pdf(g, x, y, ownx, owny);
                const double oth = mkf_iso2_pdf(g, x, y, othx, othy);
                const double Z = __dadd_rn(__dadd_rn(__dmul_rn(own, a.pa), __dmul_rn(oth, a.pa)), zc);
                w = __ddiv_rn(Lk, Z);
                gt = 1;
            }
        }
        wo[c] = w;
        go[c] = gt;
    }
}

// measurement[:, i] = [roi.x + w/2, roi.y + 0.5 h, cand_x(bins[i]), cand_y(bins[i]), roi.x + w/2, roi.y + 1.65 h]
__global__ void k_assoc_meas(const double* __restrict__ cand_xy, const double* __restrict__ roi,
                             const int32_t* __restrict__ bins, const uint32_t* __restrict__ as_status, long long T,
                             int N, int C, double neck, double* __restrict__ meas0, double* __restrict__ meas1,
                             uint32_t* __restrict__ status0, uint32_t* __restrict__ status1)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T * 2 * N) return;
    const long long th = i / N;
    const int j = (int)(i - th * N);
    const long long t = th >> 1;
    const int h = (int)(th & 1);
    const double rx = roi[t * 4 + 0], ry = roi[t * 4 + 1], rw = roi[t * 4 + 2], rh = roi[t * 4 + 3];
    const int bsel = bins[th * N + j];
    const double* __restrict__ px = cand_xy + th * 2 * (long long)C;
    double* __restrict__ m = (h ? meas1 : meas0) + t * 6 * (long long)N;
    const double cxv = __dadd_rn(rx, __ddiv_rn(rw, 2.0));
    m[0 * N + j] = cxv;
    m[1 * N + j] = __dadd_rn(ry, __dmul_rn(0.5, rh));
    m[2 * N + j] = px[bsel];
    m[3 * N + j] = px[C + bsel];
    m[4 * N + j] = cxv;
    m[5 * N + j] = __dadd_rn(ry, __dmul_rn(neck, rh));
    if (j == 0) {
        const uint32_t st = as_status[th];
        if (st) atomicOr((h ? status1 : status0) + t, st);
    }
}

// the status part of k_assoc_meas alone: candidate-resample flags of (person, hand) -> the arm batch's track status
__global__ void k_assoc_status(const uint32_t* __restrict__ as_status, long long T, uint32_t* __restrict__ status0,
                               uint32_t* __restrict__ status1)
{
    const long long th = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (th >= T * 2) return;
    const uint32_t st = as_status[th];
    if (st) atomicOr(((th & 1) ? status1 : status0) + (th >> 1), st);
}

// the pose of every track of an arm batch (device, T x D): the copy the last estimate left, or a fresh one
static int posterior_pose_device(mkf_batch* b, const double** d_pose)
{
    int rc;
    if (!b->pose_cache_on) {
        if ((rc = b->pose_cache.ensure((size_t)b->T * b->m->D * sizeof(double)))) return rc;
        b->pose_cache_on = true;
        b->pose_valid = false;
    }
    static const bool reuse = [] { // MKF_POSE_CACHE=0: always recompute (A/B runs)
        const char* e = getenv("MKF_POSE_CACHE");
        return !(e && e[0] == '0');
    }();
    if ((!b->pose_valid || !reuse) && (rc = launch_estimate(b, nullptr, nullptr))) return rc; // writes the copy
    *d_pose = (const double*)b->pose_cache.p;
    return MKF_OK;
}

extern "C" int mkf_batch_associate(mkf_batch* a0, mkf_batch* a1, int C, const double* cand_xy, const uint8_t* cand_L,
                                   const double* roi, const double* u_cand, const double* u_ind, const double* u_post,
                                   const uint64_t* seeds, int do_update, int mem)
{
    if (!a0 || !a1 || !cand_xy || !cand_L || !roi || !u_cand || C <= 0) {
        mkf_set_error("mkf_batch_associate: null or invalid argument");
        return MKF_E_INVALID;
    }
    if (do_update && (!u_ind || !u_post)) {
        mkf_set_error("mkf_batch_associate: u_ind/u_post required when do_update != 0");
        return MKF_E_INVALID;
    }
    if (a0->T != a1->T || a0->N != a1->N || a0->device != a1->device || a0->stream != a1->stream) {
        mkf_set_error("mkf_batch_associate: the two arm batches must share T, N, device and stream");
        return MKF_E_INVALID;
    }
    CK(cudaSetDevice(a0->device));
    mkf_batch* b = a0;
    const long long T = b->T;
    const int N = b->N;
    const mkf_params& prm = b->m->prm;
    int rc;
    const double *d_cand, *d_roi, *d_uc, *d_ui = nullptr, *d_up = nullptr;
    const uint8_t* d_L;
    const uint64_t* d_seeds;
    if ((rc = in_ptr(b, cand_xy, (size_t)T * 4 * C, mem, b->as_cand, &d_cand))) return rc;
    if ((rc = in_ptr(b, cand_L, (size_t)T * 2 * C, mem, b->as_L, &d_L))) return rc;
    if ((rc = in_ptr(b, roi, (size_t)T * 4, mem, b->as_roi, &d_roi))) return rc;
    if ((rc = in_ptr(b, u_cand, (size_t)T * 2, mem, b->as_u, &d_uc))) return rc;
    if ((rc = in_ptr(b, seeds, (size_t)T * 6, mem, b->as_seed, &d_seeds))) return rc;
    if (do_update) {
        if ((rc = in_ptr(b, u_ind, (size_t)T * 2, mem, b->as_ui, &d_ui))) return rc;
        if ((rc = in_ptr(b, u_post, (size_t)T * 2, mem, b->as_up, &d_up))) return rc;
    }
    static const bool materialise = [] {
        const char* e = getenv("MKF_ASSOC_MATERIALISE");
        return e && e[0] == '1';
    }();
    const bool gather = do_update && !materialise; // see below, in front of the update
    // keep the device copy of the candidates alive for mkf_batch_assoc_results / the measurement gather
    if ((rc = b->as_w.ensure((size_t)T * 2 * C * sizeof(double))) || (rc = b->as_gate.ensure((size_t)T * 2 * C)) ||
        (rc = b->as_bins.ensure((size_t)T * 2 * N * sizeof(int32_t))) ||
        (rc = b->as_wsum.ensure((size_t)T * 2 * sizeof(double))) ||
        (rc = b->as_status.ensure((size_t)T * 2 * sizeof(uint32_t))) ||
        (rc = b->as_cuts.ensure((size_t)T * 2 * 32 * sizeof(int32_t))) ||
        (!gather && (rc = b->as_meas.ensure((size_t)T * 2 * 6 * N * sizeof(double)))))
        return rc;
    b->as_C = C;
    const double *d_pose0, *d_pose1;
    // posterior hand position of both arms: rows 0..1 of pca_proj^T xbar + pca_mean^T (src/pf2DRao.cpp:111-116)
    if ((rc = posterior_pose_device(a0, &d_pose0)) || (rc = posterior_pose_device(a1, &d_pose1))) return rc;
    AssocArgs aa;
    aa.cand_xy = d_cand;
    aa.cand_L = d_L;
    aa.roi = d_roi;
    aa.pose0 = d_pose0;
    aa.pose1 = d_pose1;
    aa.w_raw = (double*)b->as_w.p;
    aa.gate = (uint8_t*)b->as_gate.p;
    aa.T = T;
    aa.C = C;
    aa.D0 = a0->m->D;
    aa.D1 = a1->m->D;
    aa.chol_mode = prm.chol_mode;
    aa.img_rows = prm.img_rows;
    aa.img_cols = prm.img_cols;
    aa.pa = prm.assoc_pa;
    aa.clutter = prm.assoc_clutter;
    aa.spread = prm.proposal_spread;
    aa.as_status = (uint32_t*)b->as_status.p;
    aa.status0 = a0->status;
    aa.status1 = a1->status;
    k_assoc_weights<<<grid_for(T * 2 * 32 * ((C + MKF_ASSOC_SPAN - 1) / MKF_ASSOC_SPAN), 128), 128, 0, b->stream>>>(aa);
    MKF_LAUNCHED();
    CK(cudaGetLastError());
    bool have_cuts = false; // few candidates: the bins also come as run boundaries (the run-length frame pipeline's input)
    if ((rc = run_resample(b->stream, T * 2, (const double*)b->as_w.p, C, N, d_uc, 1, 1,
                           (double*)b->as_wsum.p, (int32_t*)b->as_bins.p, (uint32_t*)b->as_status.p, d_seeds, 3, 0,
                           MKF_ST_CAND_FALLBACK, MKF_ST_CAND_DEGENERATE, nullptr, nullptr, nullptr, nullptr, nullptr,
                           (int32_t*)b->as_cuts.p, &have_cuts)))
        return rc;
    // With the update following at once the per-slot columns are not materialised: the slot kernels assemble a
    // slot's column from the ROI and the candidate its bin selects (MKF_MEAS_CAND; 4 bytes per slot instead of 48
    // written and read back), and slots that drew the same parent record, component AND candidate share one child.
    // MKF_ASSOC_MATERIALISE=1 keeps the T x 6 x N columns (A/B runs).
    double* d_meas0 = (double*)b->as_meas.p;
    double* d_meas1 = d_meas0 + (size_t)T * 6 * N;
    if (gather) {
        k_assoc_status<<<grid_for(T * 2, 256), 256, 0, b->stream>>>((const uint32_t*)b->as_status.p, T, a0->status,
                                                                    a1->status);
    } else {
        k_assoc_meas<<<grid_for(T * 2 * N, 256), 256, 0, b->stream>>>(d_cand, d_roi, (const int32_t*)b->as_bins.p,
                                                                      (const uint32_t*)b->as_status.p, T, N, C,
                                                                      prm.neck_offset, d_meas0, d_meas1, a0->status,
                                                                      a1->status);
    }
    MKF_LAUNCHED();
    CK(cudaGetLastError());
    if (!do_update) return MKF_OK;
    // u_ind / u_post / seeds arrive as arm-major pairs per person: de-interleave with strided 2-D copies
    DevBuf& s0 = a0->in_u0;
    DevBuf& s1 = a0->in_u1;
    DevBuf& s2 = a1->in_u0;
    DevBuf& s3 = a1->in_u1;
    if ((rc = s0.ensure((size_t)T * 8)) || (rc = s1.ensure((size_t)T * 8)) || (rc = s2.ensure((size_t)T * 8)) ||
        (rc = s3.ensure((size_t)T * 8)))
        return rc;
    CK(cudaMemcpy2DAsync(s0.p, 8, d_ui, 16, 8, (size_t)T, cudaMemcpyDeviceToDevice, b->stream));
    CK(cudaMemcpy2DAsync(s2.p, 8, d_ui + 1, 16, 8, (size_t)T, cudaMemcpyDeviceToDevice, b->stream));
    CK(cudaMemcpy2DAsync(s1.p, 8, d_up, 16, 8, (size_t)T, cudaMemcpyDeviceToDevice, b->stream));
    CK(cudaMemcpy2DAsync(s3.p, 8, d_up + 1, 16, 8, (size_t)T, cudaMemcpyDeviceToDevice, b->stream));
    // seeds layout T x 2 x 3: per arm [candidate resample, indicator resample (unused), posterior resample]
    const int lay = gather ? MKF_MEAS_CAND : MKF_MEAS_PER_SLOT;
    mkf_batch* arms[2] = {a0, a1};
    const double* ui[2] = {(const double*)s0.p, (const double*)s2.p};
    const double* up[2] = {(const double*)s1.p, (const double*)s3.p};
    for (int h = 0; h < 2; h++) {
        mkf_batch* ab = arms[h];
        ab->cm_cand = d_cand;
        ab->cm_bins = (const int32_t*)b->as_bins.p;
        ab->cm_roi = d_roi;
        ab->cm_C = C;
        ab->cm_hand = h;
        ab->cm_cuts = have_cuts ? (const int32_t*)b->as_cuts.p : nullptr;
        rc = update_device(ab, gather ? nullptr : (h ? d_meas1 : d_meas0), lay, ui[h], up[h], 1, d_seeds, 6, h ? 5 : 2);
        ab->cm_cand = nullptr;
        ab->cm_cuts = nullptr;
        ab->cm_bins = nullptr;
        ab->cm_roi = nullptr;
        if (rc) return rc;
    }
    return MKF_OK;
}

extern "C" int mkf_batch_assoc_results(mkf_batch* b, uint8_t* gate, double* weights, int32_t* bins, int mem)
{
    if (!b || b->as_C <= 0) {
        mkf_set_error("mkf_batch_assoc_results: no association has run on this batch");
        return MKF_E_INVALID;
    }
    CK(cudaSetDevice(b->device));
    const size_t nc = (size_t)b->T * 2 * b->as_C;
    DevBuf tmp;
    int rc = MKF_OK;
    if (weights) {
        // normalised weights = raw / sum (src/pfPose.cpp:294-298); one "track" = one (person, hand)
        OutPtr<double> ow;
        if ((rc = ow.init(b, weights, nc, mem, tmp))) return rc;
        k_aux_outputs<<<grid_for((long long)nc, 256), 256, 0, b->stream>>>((const double*)b->as_w.p,
                                                                          (const double*)b->as_wsum.p, nullptr,
                                                                          (long long)nc, b->as_C, 0, ow.devp, nullptr, nullptr);
        MKF_LAUNCHED();
        if (cudaGetLastError() != cudaSuccess) {
            tmp.release();
            mkf_set_error("k_aux_outputs launch failed");
            return MKF_E_CUDA;
        }
        if ((rc = ow.finish(b))) {
            tmp.release();
            return rc;
        }
    }
    auto copy_out = [&](void* dst, const void* src, size_t bytes) -> int {
        if (!dst) return MKF_OK;
        cudaMemcpyKind kind = is_device_ptr(dst, mem) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
        if (cudaMemcpyAsync(dst, src, bytes, kind, b->stream) != cudaSuccess) {
            mkf_set_error("cudaMemcpyAsync failed in mkf_batch_assoc_results");
            return MKF_E_CUDA;
        }
        return MKF_OK;
    };
    if ((rc = copy_out(gate, b->as_gate.p, nc)) ||
        (rc = copy_out(bins, b->as_bins.p, (size_t)b->T * 2 * b->N * sizeof(int32_t)))) {
        tmp.release();
        return rc;
    }
    cudaError_t e = cudaStreamSynchronize(b->stream);
    tmp.release();
    if (e != cudaSuccess) {
        mkf_set_error("cudaStreamSynchronize failed: %s", cudaGetErrorString(e));
        return MKF_E_CUDA;
    }
    return MKF_OK;
}

#endif
