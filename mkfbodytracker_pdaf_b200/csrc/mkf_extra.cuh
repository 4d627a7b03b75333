// mkf_extra.cuh -- single-object entry points behind the reference-named host shims
// (include/mkf_shims.hpp): KF_model::predict / KF_model::update on explicit Gaussians and
// ParticleFilter::getSampleProb.  Included by mkf_api.cu.
#ifndef MKF_EXTRA_CUH
#define MKF_EXTRA_CUH

__global__ void k_set_bounds_from_comp(const int32_t* __restrict__ comp, long long n, int K, int32_t* __restrict__ bounds,
                                       int32_t* __restrict__ parent)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int k = comp[t];
    int32_t* bt = bounds + t * (K + 2);
    for (int q = 0; q < K; q++) bt[q] = (q >= k) ? 1 : 0; // N = 1: slot 0 has component k
    bt[K] = 1;
    bt[K + 1] = 0;
    parent[t] = 0;
}

// KF_model::predict (stage 1, src/KF_model.cpp:11-15) and/or the innovation likelihood + KF_model::update
// (stage 2, src/pf2DRao.cpp:138 + src/KF_model.cpp:17-25) for n independent Gaussians with explicit
// component indices.  x n x d, P n x d x d in/out; z n x 6; w_out n (likelihood, stage 2).  Host pointers.
extern "C" int mkf_kf_apply(const mkf_model* m, int n, const int32_t* comp, int stage, double* x, double* P,
                            const double* z, double* w_out, int device)
{
    if (!m || n <= 0 || !comp || !x || !P || stage < 1 || stage > 3 || ((stage & 2) && !z)) {
        mkf_set_error("mkf_kf_apply: invalid argument");
        return MKF_E_INVALID;
    }
    for (int i = 0; i < n; i++)
        if (comp[i] < 0 || comp[i] >= m->K) {
            mkf_set_error("mkf_kf_apply: component index %d out of range", comp[i]);
            return MKF_E_INVALID;
        }
    mkf_batch* b = nullptr;
    int rc = mkf_batch_create(&b, m, n, 1, device, nullptr);
    if (rc) return rc;
    auto done = [&](int code) {
        mkf_batch_destroy(b);
        return code;
    };
    if ((rc = mkf_batch_upload(b, x, P, MKF_MEM_HOST))) return done(rc);
    DevBuf dcomp, dz, du;
    auto done2 = [&](int code) {
        dcomp.release();
        dz.release();
        du.release();
        return done(code);
    };
    if ((rc = dcomp.ensure((size_t)n * 4)) || (rc = dz.ensure((size_t)n * 6 * 8)) || (rc = du.ensure((size_t)n * 8)))
        return done2(rc);
    cudaMemcpyAsync(dcomp.p, comp, (size_t)n * 4, cudaMemcpyHostToDevice, b->stream);
    if (z)
        cudaMemcpyAsync(dz.p, z, (size_t)n * 6 * 8, cudaMemcpyHostToDevice, b->stream);
    else
        cudaMemsetAsync(dz.p, 0, (size_t)n * 6 * 8, b->stream);
    k_set_bounds_from_comp<<<grid_for(n, 128), 128, 0, b->stream>>>((const int32_t*)dcomp.p, n, m->K, b->bounds,
                                                                     b->parent);
    MKF_LAUNCHED();
    // run the slot kernel alone (no indicator draw, no resampling)
    SlotArgs a;
    a.st_in = b->st[b->cur];
    a.st_out = b->st[b->cur ^ 1];
    a.parent = b->parent;
    a.bounds = b->bounds;
    a.meas = (const double*)dz.p;
    a.comp_const = b->d_comp;
    a.w_raw = b->w_raw;
    a.status = b->status;
    a.total = b->total;
    a.N = 1;
    a.K = m->K;
    a.meas_layout = MKF_MEAS_SHARED;
    a.chol_mode = m->prm.chol_mode;
    a.stage = stage;
    a.alias_chain = 0;
    a.unsorted = nullptr;
    for (int r = 0; r < MKF_M; r++) a.bh[r] = m->BH[r];
    a.r = m->prm.meas_noise_var;
    const size_t smem = (size_t)m->K * b->lay.cs * sizeof(double);
    if (m->d == 12)
        k_slot_update<12, false><<<grid_for(b->total, 128), 128, smem, b->stream>>>(a);
    else
        k_slot_update<10, false><<<grid_for(b->total, 128), 128, smem, b->stream>>>(a);
    MKF_LAUNCHED();
    if (m->d == 12)
        k_slot_update_repair<12><<<grid_for(b->T, 128), 128, 0, b->stream>>>(a, nullptr);
    else
        k_slot_update_repair<10><<<grid_for(b->T, 128), 128, 0, b->stream>>>(a, nullptr);
    MKF_LAUNCHED();
    if (cudaGetLastError() != cudaSuccess) {
        mkf_set_error("mkf_kf_apply: kernel launch failed");
        return done2(MKF_E_CUDA);
    }
    b->cur ^= 1;
    rc = mkf_batch_download(b, x, P, w_out, nullptr, nullptr, nullptr, nullptr, nullptr, MKF_MEM_HOST);
    return done2(rc);
}

__global__ void k_sample_prob(const double* __restrict__ pose, int D, long long track, const double* __restrict__ cand,
                              int C, double s2, int chol_mode, double* __restrict__ out)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const Iso2 g = mkf_iso2_setup(s2, chol_mode);
    out[c] = mkf_iso2_pdf(g, cand[c], cand[C + c], pose[track * D + 0], pose[track * D + 1]);
}

// ParticleFilter::getSampleProb (src/pf2DRao.cpp:105-122) for one track: density of C candidate
// positions (cand_xy 2 x C, row 0 = x) under N(hand estimate, 0.8*scale*I).  Host pointers.
extern "C" int mkf_batch_sample_prob(mkf_batch* b, int64_t track, const double* cand_xy, int C, double scale,
                                     double* out)
{
    if (!b || !cand_xy || !out || C <= 0 || track < 0 || track >= b->T) {
        mkf_set_error("mkf_batch_sample_prob: invalid argument");
        return MKF_E_INVALID;
    }
    CK(cudaSetDevice(b->device));
    int rc;
    if ((rc = b->as_hand.ensure((size_t)b->T * b->m->D * 8)) || (rc = b->as_cand.ensure((size_t)2 * C * 8)) ||
        (rc = b->as_w.ensure((size_t)C * 8)))
        return rc;
    if ((rc = estimate_pose_device(b, (double*)b->as_hand.p))) return rc;
    CK(cudaMemcpyAsync(b->as_cand.p, cand_xy, (size_t)2 * C * 8, cudaMemcpyHostToDevice, b->stream));
    const mkf_params& prm = b->m->prm;
    k_sample_prob<<<grid_for(C, 128), 128, 0, b->stream>>>((const double*)b->as_hand.p, b->m->D, track,
                                                            (const double*)b->as_cand.p, C,
                                                            prm.proposal_spread * scale * 1.0, prm.chol_mode,
                                                            (double*)b->as_w.p);
    MKF_LAUNCHED();
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, b->as_w.p, (size_t)C * 8, cudaMemcpyDeviceToHost, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    b->as_C = 0; // association scratch was reused
    return MKF_OK;
}

#endif
