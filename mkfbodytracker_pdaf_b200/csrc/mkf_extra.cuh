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
    cudaError_t ce = cudaMemcpyAsync(dcomp.p, comp, (size_t)n * 4, cudaMemcpyHostToDevice, b->stream);
    if (ce == cudaSuccess)
        ce = z ? cudaMemcpyAsync(dz.p, z, (size_t)n * 6 * 8, cudaMemcpyHostToDevice, b->stream)
               : cudaMemsetAsync(dz.p, 0, (size_t)n * 6 * 8, b->stream);
    if (ce != cudaSuccess) {
        mkf_set_error("mkf_kf_apply: %s", cudaGetErrorString(ce));
        return done2(MKF_E_CUDA);
    }
    k_set_bounds_from_comp<<<grid_for(n, 128), 128, 0, b->stream>>>((const int32_t*)dcomp.p, n, m->K, b->bounds,
                                                                     b->parent);
    MKF_LAUNCHED();
    // run the slot kernel alone (no indicator draw, no resampling)
    SlotArgs a{};
    a.st_in = b->st[b->cur];
    a.st_out = b->st[b->cur ^ 1];
    a.parent = b->parent;
    a.src = b->parent;
    a.rep = nullptr;
    a.dedup = 0;
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
        k_slot_update_repair<12, false><<<grid_for(b->T, 128), 128, 0, b->stream>>>(a, nullptr, SmallTailArgs{});
    else
        k_slot_update_repair<10, false><<<grid_for(b->T, 128), 128, 0, b->stream>>>(a, nullptr, SmallTailArgs{});
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
    if ((rc = b->as_cand.ensure((size_t)2 * C * 8)) || (rc = b->as_w.ensure((size_t)C * 8))) return rc;
    const double* d_pose; // the estimate the caller's loop has already computed, if the state has not changed since
    if ((rc = posterior_pose_device(b, &d_pose))) return rc;
    CK(cudaMemcpyAsync(b->as_cand.p, cand_xy, (size_t)2 * C * 8, cudaMemcpyHostToDevice, b->stream));
    const mkf_params& prm = b->m->prm;
    k_sample_prob<<<grid_for(C, 128), 128, 0, b->stream>>>(d_pose, b->m->D, track,
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

// ------------------------------------------------------------------------------------------------
// output back-end: PFTracker::rpy / get3Dpose / publishTFtree / publish2Dpos (src/pfPose.cpp:84-208)
// ------------------------------------------------------------------------------------------------
struct CamK {
    double k[9];
};

// get3Dpose (src/pfPose.cpp:93-127) of one estimate e (>= 21 entries): pos3D 3 x 5 (row-major)
__device__ __forceinline__ void mkf_get3dpose(const double* __restrict__ e, const CamK& cam, double* pos3D)
{
    const double roll = e[16], pitch = e[17], yaw = e[15]; // rpy(est(0,16), est(0,17), est(0,15))
    const double cr = cos(roll), sr = sin(roll), cp = cos(pitch), sp = sin(pitch), cy = cos(yaw), sy = sin(yaw);
    const double R1[9] = {1, 0, 0, 0, cr, -sr, 0, sr, cr};
    const double R2[9] = {cp, 0, sp, 0, 1, 0, -sp, 0, cp};
    const double R3[9] = {cy, -sy, 0, sy, cy, 0, 0, 0, 1};
    double R32[9], R[9], T[12], P[12];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
            double s = 0.0;
            for (int k = 0; k < 3; k++) s += R3[r * 3 + k] * R2[k * 3 + c];
            R32[r * 3 + c] = s;
        }
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
            double s = 0.0;
            for (int k = 0; k < 3; k++) s += R32[r * 3 + k] * R1[k * 3 + c];
            R[r * 3 + c] = s;
        }
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) T[r * 4 + c] = R[r * 3 + c];
        T[r * 4 + 3] = e[18 + r];
    }
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) {
            double s = 0.0;
            for (int k = 0; k < 3; k++) s += cam.k[r * 3 + k] * T[k * 4 + c];
            P[r * 4 + c] = s;
        }
#define SD(r, c) P[(r)*4 + (c)]
    double d = SD(0, 0) * (SD(1, 1) * SD(2, 2) - SD(1, 2) * SD(2, 1)) - SD(0, 1) * (SD(1, 0) * SD(2, 2) - SD(1, 2) * SD(2, 0)) +
               SD(0, 2) * (SD(1, 0) * SD(2, 1) - SD(1, 1) * SD(2, 0));
    double PI[9];
    if (d != 0.0) {
        d = 1.0 / d;
        PI[0] = (SD(1, 1) * SD(2, 2) - SD(1, 2) * SD(2, 1)) * d;
        PI[1] = (SD(0, 2) * SD(2, 1) - SD(0, 1) * SD(2, 2)) * d;
        PI[2] = (SD(0, 1) * SD(1, 2) - SD(0, 2) * SD(1, 1)) * d;
        PI[3] = (SD(1, 2) * SD(2, 0) - SD(1, 0) * SD(2, 2)) * d;
        PI[4] = (SD(0, 0) * SD(2, 2) - SD(0, 2) * SD(2, 0)) * d;
        PI[5] = (SD(0, 2) * SD(1, 0) - SD(0, 0) * SD(1, 2)) * d;
        PI[6] = (SD(1, 0) * SD(2, 1) - SD(1, 1) * SD(2, 0)) * d;
        PI[7] = (SD(0, 1) * SD(2, 0) - SD(0, 0) * SD(2, 1)) * d;
        PI[8] = (SD(0, 0) * SD(1, 1) - SD(0, 1) * SD(1, 0)) * d;
    } else {
        for (int i = 0; i < 9; i++) PI[i] = 0.0; // cv::invert: dst = 0 when singular
    }
#undef SD
    for (int k = 0; k < 5; k++) {
        const double z = e[3 * k + 2];
        const double im[3] = {e[3 * k] * z - P[3], e[3 * k + 1] * z - P[7], z - P[11]};
        for (int r = 0; r < 3; r++) {
            double s = 0.0;
            for (int c = 0; c < 3; c++) s += PI[r * 3 + c] * im[c];
            pos3D[r * 5 + k] = s;
        }
    }
}

__global__ void k_pose3d(const double* __restrict__ pose, int D, long long T, CamK cam, double* __restrict__ pos3d)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    double out[15];
    mkf_get3dpose(pose + t * D, cam, out);
    for (int i = 0; i < 15; i++) pos3d[t * 15 + i] = out[i];
}

__global__ void k_skeleton(const double* __restrict__ pose1, int D1, const double* __restrict__ pose2, int D2,
                           long long T, CamK cam, double* __restrict__ tf, double* __restrict__ joints2d)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const double* __restrict__ e1 = pose1 + t * D1;
    const double* __restrict__ e2 = pose2 + t * D2;
    double p1[15], p2[15];
    mkf_get3dpose(e1, cam, p1);
    mkf_get3dpose(e2, cam, p2);
#define P1(r, c) p1[(r)*5 + (c)]
#define P2(r, c) p2[(r)*5 + (c)]
    if (tf) {
        double* o = tf + t * 30;
        for (int k = 0; k < 2; k++) {
            *o++ = P1(0, k) - P1(0, k + 1);
            *o++ = P1(2, k) - P1(2, k + 1);
            *o++ = -P1(1, k) + P1(1, k + 1);
            *o++ = P2(0, k) - P2(0, k + 1);
            *o++ = P2(2, k) - P2(2, k + 1);
            *o++ = -P2(1, k) + P2(1, k + 1);
        }
        const double neck_x = (P1(0, 4) + P2(0, 4)) / 2.0, neck_y = (P1(1, 4) + P2(1, 4)) / 2.0,
                     neck_z = (P1(2, 4) + P2(2, 4)) / 2.0;
        const double head_x = (P1(0, 3) + P2(0, 3)) / 2.0, head_y = (P1(1, 3) + P2(1, 3)) / 2.0,
                     head_z = (P1(2, 3) + P2(2, 3)) / 2.0;
        *o++ = P1(0, 2) - neck_x;
        *o++ = P1(2, 2) - neck_z;
        *o++ = -P1(1, 2) + neck_y;
        *o++ = P2(0, 2) - neck_x;
        *o++ = P2(2, 2) - neck_z;
        *o++ = -P2(1, 2) + neck_y;
        *o++ = neck_x - head_x;
        *o++ = neck_z - head_z;
        *o++ = -neck_y + head_y;
        *o++ = head_x;
        *o++ = head_z;
        *o++ = -head_y;
        *o++ = -e1[18];
        *o++ = -e1[20];
        *o++ = e1[19];
        *o++ = -e1[16];
        *o++ = -e1[17];
        *o++ = -e1[15];
    }
#undef P1
#undef P2
    if (joints2d) {
        double* j = joints2d + t * 16;
        j[0] = e1[0];
        j[1] = e1[1];
        j[2] = e2[0];
        j[3] = e2[1];
        j[4] = 0.5 * (e1[9] + e2[9]);
        j[5] = 0.5 * (e1[10] + e2[10]);
        j[6] = 0.5 * (e1[12] + e2[12]);
        j[7] = 0.5 * (e1[13] + e2[13]);
        j[8] = e1[3];
        j[9] = e1[4];
        j[10] = e2[3];
        j[11] = e2[4];
        j[12] = e1[6];
        j[13] = e1[7];
        j[14] = e2[6];
        j[15] = e2[7];
    }
}

static CamK make_cam(const double* Kcam)
{
    CamK c;
    // Kinect literal of src/pfPose.cpp:101 (the webcam matrix of cal.yml is commented out at :98)
    const double kinect[9] = {525.0, 0.0, 319.5, 0.0, 525.0, 239.5, 0.0, 0.0, 1.0};
    for (int i = 0; i < 9; i++) c.k[i] = Kcam ? Kcam[i] : kinect[i];
    return c;
}

extern "C" int mkf_batch_pose3d(mkf_batch* b, const double* Kcam, double* pos3d, int mem)
{
    if (!b || !pos3d) {
        mkf_set_error("mkf_batch_pose3d: null argument");
        return MKF_E_INVALID;
    }
    if (b->m->D < 21) {
        mkf_set_error("mkf_batch_pose3d: pose vector too short (D=%d)", b->m->D);
        return MKF_E_INVALID;
    }
    CK(cudaSetDevice(b->device));
    int rc;
    const double* d_pose;
    if ((rc = posterior_pose_device(b, &d_pose))) return rc;
    OutPtr<double> o;
    if ((rc = o.init(b, pos3d, (size_t)b->T * 15, mem, b->out_a))) return rc;
    k_pose3d<<<grid_for(b->T, 128), 128, 0, b->stream>>>(d_pose, b->m->D, b->T, make_cam(Kcam),
                                                         o.devp);
    MKF_LAUNCHED();
    CK(cudaGetLastError());
    if ((rc = o.finish(b))) return rc;
    if (o.host) CK(cudaStreamSynchronize(b->stream));
    return MKF_OK;
}

extern "C" int mkf_batch_skeleton(mkf_batch* a0, mkf_batch* a1, const double* Kcam, double* tf, double* joints2d,
                                  int mem)
{
    if (!a0 || !a1 || (!tf && !joints2d)) {
        mkf_set_error("mkf_batch_skeleton: null argument");
        return MKF_E_INVALID;
    }
    if (a0->T != a1->T || a0->device != a1->device || a0->stream != a1->stream || a0->m->D < 21 || a1->m->D < 21) {
        mkf_set_error("mkf_batch_skeleton: the two arm batches must share T, device and stream (D >= 21)");
        return MKF_E_INVALID;
    }
    CK(cudaSetDevice(a0->device));
    int rc;
    const double *p1, *p2;
    if ((rc = posterior_pose_device(a0, &p1)) || (rc = posterior_pose_device(a1, &p2))) return rc;
    OutPtr<double> otf, oj;
    if ((rc = otf.init(a0, tf, (size_t)a0->T * 30, mem, a0->out_a))) return rc;
    if ((rc = oj.init(a0, joints2d, (size_t)a0->T * 16, mem, a0->out_b))) return rc;
    k_skeleton<<<grid_for(a0->T, 128), 128, 0, a0->stream>>>(p1, a0->m->D, p2, a1->m->D, a0->T, make_cam(Kcam), otf.devp,
                                                             oj.devp);
    MKF_LAUNCHED();
    CK(cudaGetLastError());
    if ((rc = otf.finish(a0)) || (rc = oj.finish(a0))) return rc;
    if (otf.host || oj.host) CK(cudaStreamSynchronize(a0->stream));
    return MKF_OK;
}

// ------------------------------------------------------------------------------------------------
// candidate generation front-end (the step before the path): getSamples / first-frame box / likelihood
// image lookup (src/pf2DRao.cpp:85-103, src/pfPose.cpp:216-236,254) on the device
// ------------------------------------------------------------------------------------------------
__global__ void k_propose(const double* __restrict__ pose0, int D0, const double* __restrict__ pose1, int D1,
                          const double* __restrict__ roi, const uint8_t* __restrict__ tracking,
                          const uint8_t* __restrict__ like, int n_img, int rows, int cols, double spread, uint64_t seed,
                          uint64_t frame, long long track0, long long T, int C, double* __restrict__ cand_xy,
                          uint8_t* __restrict__ cand_L)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T * 2 * C) return;
    const long long th = i / C;
    const int c = (int)(i - th * C);
    const long long t = th >> 1;
    const int h = (int)(th & 1);
    const double r4[4] = {roi[t * 4], roi[t * 4 + 1], roi[t * 4 + 2], roi[t * 4 + 3]};
    const double hx = h ? pose1[t * D1] : pose0[t * D0], hy = h ? pose1[t * D1 + 1] : pose0[t * D0 + 1];
    double x, y;
    mkf_synth_proposal(seed, (uint64_t)(track0 + t), frame, h, C, c, tracking ? tracking[t] : 1, hx, hy, r4, rows, cols,
                       spread, &x, &y);
    cand_xy[(th * 2 + 0) * C + c] = x;
    cand_xy[(th * 2 + 1) * C + c] = y;
    if (cand_L) {
        const uint8_t* img = like + (n_img == 1 ? 0 : t) * (long long)rows * cols;
        cand_L[th * C + c] = mkf_likelihood_lookup(img, rows, cols, x, y);
    }
}

extern "C" int mkf_batch_propose(mkf_batch* a0, mkf_batch* a1, int C, const double* roi, const uint8_t* tracking,
                                 const uint8_t* like, int n_img, uint64_t seed, uint64_t frame, int64_t track0,
                                 double* cand_xy, uint8_t* cand_L, int mem)
{
    if (!a0 || !a1 || !roi || !cand_xy || C <= 0 || (cand_L && (!like || (n_img != 1 && n_img != a0->T)))) {
        mkf_set_error("mkf_batch_propose: null or invalid argument");
        return MKF_E_INVALID;
    }
    if (a0->T != a1->T || a0->device != a1->device || a0->stream != a1->stream) {
        mkf_set_error("mkf_batch_propose: the two arm batches must share T, device and stream");
        return MKF_E_INVALID;
    }
    CK(cudaSetDevice(a0->device));
    mkf_batch* b = a0;
    const long long T = b->T;
    const mkf_params& prm = b->m->prm;
    int rc;
    const double* d_roi;
    const uint8_t *d_trk, *d_like;
    if ((rc = in_ptr(b, roi, (size_t)T * 4, mem, b->as_roi, &d_roi))) return rc;
    if ((rc = in_ptr(b, tracking, (size_t)T, mem, b->as_u, &d_trk))) return rc;
    if ((rc = in_ptr(b, like, cand_L ? (size_t)n_img * prm.img_rows * prm.img_cols : 0, mem, b->as_L, &d_like)))
        return rc;
    const double *p0, *p1;
    if ((rc = posterior_pose_device(a0, &p0)) || (rc = posterior_pose_device(a1, &p1))) return rc;
    OutPtr<double> oxy;
    OutPtr<uint8_t> oL;
    if ((rc = oxy.init(b, cand_xy, (size_t)T * 4 * C, mem, b->out_a))) return rc;
    if ((rc = oL.init(b, cand_L, (size_t)T * 2 * C, mem, b->out_b))) return rc;
    k_propose<<<grid_for(T * 2 * C, 256), 256, 0, b->stream>>>(p0, a0->m->D, p1, a1->m->D, d_roi, d_trk, d_like, n_img,
                                                               prm.img_rows, prm.img_cols, prm.proposal_spread, seed,
                                                               frame, track0, T, C, oxy.devp, oL.devp);
    MKF_LAUNCHED();
    CK(cudaGetLastError());
    if ((rc = oxy.finish(b)) || (rc = oL.finish(b))) return rc;
    if (oxy.host || oL.host) CK(cudaStreamSynchronize(b->stream));
    return MKF_OK;
}

#endif
