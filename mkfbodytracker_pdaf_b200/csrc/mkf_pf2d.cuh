// mkf_pf2d.cuh -- the legacy plain particle filter (src/pf2D.{h,cpp}; not compiled by the reference's
// CMakeLists.txt:29 but part of the path: "particle likelihoods/sec"), included by mkf_api.cu.
//
//   k_pf2d_weight           w_i = [sum_k w_k c_k expf(float(-1/2 d^T Sigma_k^-1 d))] * N2(p[6:8]-z0; 15 I) * N2(p[0:2]-z1; 15 I)
//                           (src/pf2D.cpp:157-172, :105-109, :124-128) -- one thread per particle
//   run_resample            normalise (:174-177) + systematic resample (:225-268)
//   k_pf2d_resample_predict particles.row(i) = old.row(parent_i) (+ N(0,5) per dimension, :90-102)
#ifndef MKF_PF2D_CUH
#define MKF_PF2D_CUH

struct PfRand {
    uint64_t seed = 0, epoch = 0;
    long long track0 = 0;
    const uint8_t* side = nullptr; // T flags on the device (nullptr: all 0)
    int im_w = 640, im_h = 480;
};

struct mkf_pf2d {
    long long T = 0;
    int N = 0, d = 0, K = 0, device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    double* part[2] = {nullptr, nullptr}; // T x N x d, ping-pong
    int cur = 0;
    double *w_raw = nullptr, *wsum = nullptr;
    int32_t* parent = nullptr;
    uint32_t* status = nullptr;
    double* gmm = nullptr; // K x (d + d*d + 2): mean, sigma_i, det_s, weight
    int gstride = 0;
    DevBuf in_meas, in_u, in_noise, in_part;
    // particle randomisation of the constructor and of resample()'s degenerate branch (src/pf2D.cpp:58-70,232-250),
    // drawn from the counter generator keyed (seed, track0 + t, epoch, particle, dim); epoch 0 = constructor,
    // epoch n = the n-th update
    PfRand rnd;
    uint8_t* d_side = nullptr;
    // optional per-kernel timing (mkf_pf2d_profile): start | weights | normalise + resample | gather + predict
    bool prof_on = false;
    std::vector<cudaEvent_t> prof_ev;
    int prof_n = 0, prof_cap = 0;
};

template <int D>
__global__ void __launch_bounds__(128) k_pf2d_weight(const double* __restrict__ part, const double* __restrict__ meas,
                                                      const double* __restrict__ gmm, int K, int gstride, long long T,
                                                      int N, double* __restrict__ w_raw)
{
    extern __shared__ double sg[];
    for (int i = threadIdx.x; i < K * gstride; i += blockDim.x) sg[i] = gmm[i];
    __syncthreads();
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= T * N) return;
    const long long t = s / N;
    double x[D];
    const double2* __restrict__ src = reinterpret_cast<const double2*>(part + s * D);
#pragma unroll
    for (int p = 0; p < D / 2; p++) {
        const double2 q = __ldg(src + p);
        x[2 * p] = q.x;
        x[2 * p + 1] = q.y;
    }
    double prior = 0.0;
    for (int k = 0; k < K; k++) {
        const double* __restrict__ mu = sg + k * gstride;
        const double* __restrict__ Si = mu + D;
        double xu[D];
#pragma unroll
        for (int c = 0; c < D; c++) xu[c] = __dsub_rn(x[c], mu[c]);
        // temp = -0.5*x_u*sigma_i (gemm, alpha = -0.5) then * x_u.t() (GEMM_2_T: 4 interleaved partial sums)
        double s4[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int c = 0; c < D; c++) {
            double acc = 0.0;
#pragma unroll
            for (int r = 0; r < D; r++) acc = __dadd_rn(acc, __dmul_rn(xu[r], Si[r * D + c]));
            const double tc = __dmul_rn(acc, -0.5);
            if (c < (D / 4) * 4)
                s4[c & 3] = __dadd_rn(s4[c & 3], __dmul_rn(tc, xu[c]));
            else
                s4[0] = __dadd_rn(s4[0], __dmul_rn(tc, xu[c]));
        }
        const double q = __dadd_rn(__dadd_rn(__dadd_rn(s4[0], s4[1]), s4[2]), s4[3]);
        // quirk B12: expf(float(q)) -- glibc's expf, restated operation by operation in include/mkf_expf.h (shared
        // with the oracle; 0 mismatches against the host libm over all 2^32 arguments)
        const double e = (double)mkf_expf((float)q);
        prior = __dadd_rn(prior, __dmul_rn(__dmul_rn(mu[D + D * D + 1], mu[D + D * D]), e));
    }
    // eyemvnpdf(x_u, 15): alpha = -0.5*1.0/15; 1/pow(2 pi 15, 1) * exp(alpha*(dx^2 + dy^2))
    const double alpha = __ddiv_rn(__dmul_rn(-0.5, 1.0), 15.0);
    const double nrm = __ddiv_rn(1.0, __dmul_rn(__dmul_rn(2.0, 3.14159265358979323846), 15.0)); // 1/pow(2 pi 15, 1)
    const double* __restrict__ mz = meas + t * 4;
    double lik = 1.0;
    {
        const double dx = __dsub_rn(x[6], mz[0]), dy = __dsub_rn(x[7], mz[1]);
        const double q = __dadd_rn(__dmul_rn(__dmul_rn(dx, alpha), dx), __dmul_rn(__dmul_rn(dy, alpha), dy));
        lik = __dmul_rn(nrm, exp(q));
    }
    {
        const double dx = __dsub_rn(x[0], mz[2]), dy = __dsub_rn(x[1], mz[3]);
        const double q = __dadd_rn(__dmul_rn(__dmul_rn(dx, alpha), dx), __dmul_rn(__dmul_rn(dy, alpha), dy));
        lik = __dmul_rn(lik, __dmul_rn(nrm, exp(q)));
    }
    w_raw[s] = __dmul_rn(prior, lik);
}

__global__ void k_pf2d_resample_predict(const double* __restrict__ old_p, double* __restrict__ new_p,
                                        int32_t* __restrict__ parent, const uint32_t* __restrict__ status,
                                        const double* __restrict__ noise, long long T, int N, int d, const PfRand rnd)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T * N * d) return;
    const long long s = i / d;
    const int c = (int)(i - s * d);
    const long long t = s / N;
    double v;
    if (status[t] & MKF_ST_POST_DEGENERATE) {
        // max weight 0: every particle is re-randomised across the image (src/pf2D.cpp:232-244)
        v = mkf_synth_pf2d_uniform(rnd.seed, (uint64_t)(rnd.track0 + t), rnd.epoch, (int)(s - t * N), c,
                                   rnd.side ? (int)rnd.side[t] : 0, rnd.im_w, rnd.im_h);
        if (c == d - 1) parent[s] = (int)(s - t * N); // no parent in this branch: reported as the particle itself
                                                      // (the resampler's cv::RNG draw belongs to pf2DRao, not to pf2D)
    } else {
        const long long sp = t * N + parent[s];
        v = __dadd_rn(0.0, old_p[sp * d + c]); // particles.row(i) = zeros + old_particles.row(idx)
    }
    if (noise && c < 8) v = __dadd_rn(v, __dmul_rn(noise[i], 5.0));
    new_p[i] = v;
}

// the same, one thread per particle for even compile-time D: D/2 16-byte gathers in flight per thread instead of one
// 8-byte load (the element-wise version reached 2.8-3.1 TB/s at 86 % occupancy, latency-bound: ncu long-scoreboard 17)
template <int D>
__global__ void __launch_bounds__(256) k_pf2d_resample_predict_v(const double* __restrict__ old_p,
                                                                  double* __restrict__ new_p,
                                                                  int32_t* __restrict__ parent,
                                                                  const uint32_t* __restrict__ status,
                                                                  const double* __restrict__ noise, long long T, int N,
                                                                  const PfRand rnd)
{
    static_assert(D % 2 == 0, "vector path needs an even dimension");
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= T * N) return;
    const long long t = s / N;
    double2 v[D / 2], nz[D / 2];
    if (__ldg(status + t) & MKF_ST_POST_DEGENERATE) {
        // max weight 0: every particle is re-randomised across the image (src/pf2D.cpp:232-244)
        const int side = rnd.side ? (int)rnd.side[t] : 0, j = (int)(s - t * N);
        parent[s] = j; // no parent in this branch: reported as the particle itself
#pragma unroll
        for (int p = 0; p < D / 2; p++) {
            // "- 0.0": the common tail below adds 0.0 to what it takes for a gathered row (x + 0.0 == x here)
            v[p].x = mkf_synth_pf2d_uniform(rnd.seed, (uint64_t)(rnd.track0 + t), rnd.epoch, j, 2 * p, side, rnd.im_w, rnd.im_h);
            v[p].y = mkf_synth_pf2d_uniform(rnd.seed, (uint64_t)(rnd.track0 + t), rnd.epoch, j, 2 * p + 1, side, rnd.im_w, rnd.im_h);
        }
    } else {
        const long long sp = t * N + __ldg(parent + s);
        const double2* __restrict__ src = reinterpret_cast<const double2*>(old_p + sp * D);
#pragma unroll
        for (int p = 0; p < D / 2; p++) v[p] = __ldg(src + p);
    }
    if (noise) {
        const double2* __restrict__ ns = reinterpret_cast<const double2*>(noise + s * D);
#pragma unroll
        for (int p = 0; p < D / 2; p++) nz[p] = __ldg(ns + p);
    }
    double2* __restrict__ dst = reinterpret_cast<double2*>(new_p + s * D);
#pragma unroll
    for (int p = 0; p < D / 2; p++) {
        double2 o;
        o.x = __dadd_rn(0.0, v[p].x);
        o.y = __dadd_rn(0.0, v[p].y);
        if (noise && 2 * p < 8) {
            o.x = __dadd_rn(o.x, __dmul_rn(nz[p].x, 5.0));
            o.y = __dadd_rn(o.y, __dmul_rn(nz[p].y, 5.0));
        }
        __stcs(dst + p, o);
    }
}

// the constructor's randomisation (src/pf2D.cpp:58-70): one thread per matrix element
__global__ void k_pf2d_randomise(double* __restrict__ part, long long T, int N, int d, const PfRand rnd)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T * N * d) return;
    const long long s = i / d;
    const int c = (int)(i - s * d);
    const long long t = s / N;
    part[i] = mkf_synth_pf2d_uniform(rnd.seed, (uint64_t)(rnd.track0 + t), rnd.epoch, (int)(s - t * N), c,
                                     rnd.side ? (int)rnd.side[t] : 0, rnd.im_w, rnd.im_h);
}

extern "C" void mkf_pf2d_destroy(mkf_pf2d* p)
{
    if (!p) return;
    cudaSetDevice(p->device);
    if (p->stream) cudaStreamSynchronize(p->stream);
    void* ptrs[] = {p->part[0], p->part[1], p->w_raw, p->wsum, p->parent, p->status, p->gmm, p->d_side};
    for (void* q : ptrs)
        if (q) cudaFree(q);
    p->in_meas.release();
    p->in_u.release();
    p->in_noise.release();
    p->in_part.release();
    for (cudaEvent_t e : p->prof_ev) cudaEventDestroy(e);
    if (p->own_stream && p->stream) cudaStreamDestroy(p->stream);
    delete p;
}

// host: sigma_i = inv(s) (DECOMP_CHOLESKY), det_s = 1/(pow(2 pi, d/2) sqrt(det s))  (src/pf2D.cpp:28-37)
static bool pf2d_gaussian(int d, const double* s, double* inv, double* det_s)
{
    std::vector<double> Lm(s, s + (size_t)d * d);
    for (int i = 0; i < d; i++) {
        for (int j = 0; j < i; j++) {
            double t = Lm[i * d + j];
            for (int k = 0; k < j; k++) t -= Lm[i * d + k] * Lm[j * d + k];
            Lm[i * d + j] = t * Lm[j * d + j];
        }
        double t = Lm[i * d + i];
        for (int k = 0; k < i; k++) t -= Lm[i * d + k] * Lm[i * d + k];
        if (!(t >= 2.220446049250313e-16)) return false;
        Lm[i * d + i] = 1.0 / std::sqrt(t);
    }
    double det = 1.0;
    for (int c = 0; c < d; c++) {
        std::vector<double> y(d);
        for (int i = 0; i < d; i++) {
            double t = (i == c) ? 1.0 : 0.0;
            for (int k = 0; k < i; k++) t -= Lm[i * d + k] * y[k];
            y[i] = t * Lm[i * d + i];
        }
        for (int i = d - 1; i >= 0; i--) {
            double t = y[i];
            for (int k = d - 1; k > i; k--) t -= Lm[k * d + i] * inv[k * d + c];
            inv[i * d + c] = t * Lm[i * d + i];
        }
    }
    {   // cv::determinant(s) is a partially pivoted LU on a copy (not the Cholesky factor above): the pivots' reciprocals
        // are multiplied up, then inverted once -- the last bits of det_s follow that order
        std::vector<double> U(s, s + (size_t)d * d);
        double acc = 1.0;
        for (int c = 0; c < d && acc != 0.0; c++) {
            int piv = c;
            for (int r = c + 1; r < d; r++)
                if (std::fabs(U[r * d + c]) > std::fabs(U[piv * d + c])) piv = r;
            if (std::fabs(U[piv * d + c]) < 2.220446049250313e-16) { acc = 0.0; break; }
            if (piv != c) {
                for (int q = c; q < d; q++) std::swap(U[c * d + q], U[piv * d + q]);
                acc = -acc;
            }
            const double neg_rcp = -1 / U[c * d + c];
            for (int r = c + 1; r < d; r++) {
                const double f = U[r * d + c] * neg_rcp;
                for (int q = c + 1; q < d; q++) U[r * d + q] += f * U[c * d + q];
            }
            U[c * d + c] = -neg_rcp;
        }
        if (acc != 0.0) {
            // sign first, then the reciprocal pivots in order
            double r = acc;
            for (int c = 0; c < d; c++) r *= U[c * d + c];
            det = 1. / r;
        } else
            det = 0.0;
    }
    *det_s = 1.0 / (std::pow(2.0 * 3.14159265358979323846, d / 2.0) * std::sqrt(det));
    return true;
}

extern "C" int mkf_pf2d_create(mkf_pf2d** out, int64_t T, int N, int d, int K, const double* means, const double* covs,
                               const double* weights, int device, void* stream)
{
    if (!out || !means || !covs || !weights || T <= 0 || N <= 0 || K <= 0) {
        mkf_set_error("mkf_pf2d_create: invalid argument");
        return MKF_E_INVALID;
    }
    *out = nullptr;
    if (!(d == 8 || d == 10 || d == 12)) {
        mkf_set_error("mkf_pf2d_create: d=%d not built (d in {8,10,12}; the reference needs d >= 8)", d);
        return MKF_E_UNSUPPORTED;
    }
    int ndev = mkf_device_count();
    if (ndev <= 0) {
        mkf_set_error("no CUDA device available: libmkf_b200 has no CPU fallback");
        return MKF_E_CUDA;
    }
    if (device < 0 || device >= ndev) {
        mkf_set_error("mkf_pf2d_create: bad device");
        return MKF_E_INVALID;
    }
    CK(cudaSetDevice(device));
    mkf_pf2d* p = new (std::nothrow) mkf_pf2d;
    if (!p) return MKF_E_NOMEM;
    p->T = T;
    p->N = N;
    p->d = d;
    p->K = K;
    p->device = device;
    p->gstride = d + d * d + 2;
    std::vector<double> g((size_t)K * p->gstride);
    for (int k = 0; k < K; k++) {
        double* gk = &g[(size_t)k * p->gstride];
        memcpy(gk, means + (size_t)k * d, sizeof(double) * d);
        if (!pf2d_gaussian(d, covs + (size_t)k * d * d, gk + d, gk + d + d * d)) {
            mkf_set_error("mkf_pf2d_create: covariance %d is not positive definite", k);
            delete p;
            return MKF_E_INVALID;
        }
        gk[d + d * d + 1] = weights[k];
    }
    if (stream) {
        p->stream = (cudaStream_t)stream;
    } else {
        if (cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking) != cudaSuccess) {
            mkf_set_error("cudaStreamCreate failed");
            delete p;
            return MKF_E_CUDA;
        }
        p->own_stream = true;
    }
    const size_t tot = (size_t)T * N;
    cudaError_t e;
    if ((e = cudaMalloc((void**)&p->part[0], tot * d * 8)) || (e = cudaMalloc((void**)&p->part[1], tot * d * 8)) ||
        (e = cudaMalloc((void**)&p->w_raw, tot * 8)) || (e = cudaMalloc((void**)&p->wsum, (size_t)T * 8)) ||
        (e = cudaMalloc((void**)&p->parent, tot * 4)) || (e = cudaMalloc((void**)&p->status, (size_t)T * 4)) ||
        (e = cudaMalloc((void**)&p->gmm, g.size() * 8))) {
        cudaGetLastError();
        mkf_set_error("mkf_pf2d_create: cudaMalloc failed (%s)", cudaGetErrorString(e));
        mkf_pf2d_destroy(p);
        return MKF_E_NOMEM;
    }
    cudaMemset(p->part[0], 0, tot * d * 8);
    cudaMemset(p->w_raw, 0, tot * 8);
    cudaMemset(p->wsum, 0, (size_t)T * 8);
    cudaMemset(p->parent, 0, tot * 4);
    cudaMemset(p->status, 0, (size_t)T * 4);
    cudaMemcpy(p->gmm, g.data(), g.size() * 8, cudaMemcpyHostToDevice);
    *out = p;
    return MKF_OK;
}

// ParticleFilter::getEstimator of the legacy filter (src/pf2D.cpp:79-88): sum_i weights[i] * particles.row(i), with
// the weights as update() left them -- normalised but NOT reset by resample() (src/pf2D.cpp:174-177,225-268), so they
// pair the pre-resample weights with the resampled, predicted particles exactly as the reference does; 1/N after the
// degenerate branch (:246-250).  One CTA per filter, tree sum (the reference sums in index order).
__global__ void __launch_bounds__(256) k_pf2d_estimate(const double* __restrict__ part, const double* __restrict__ w_raw,
                                                        const double* __restrict__ wsum,
                                                        const uint32_t* __restrict__ status, int N, int d,
                                                        double* __restrict__ est)
{
    __shared__ double red[8][12];
    const long long t = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const double ws = wsum[t];
    const bool flat = (status[t] & MKF_ST_POST_DEGENERATE) != 0 || !(ws > 0.0); // (no update yet: weights are 1/N)
    const double inv_n = __ddiv_rn(1.0, (double)N);
    double acc[12];
#pragma unroll
    for (int c = 0; c < 12; c++) acc[c] = 0.0;
    for (int i = tid; i < N; i += 256) {
        const double w = flat ? inv_n : __ddiv_rn(w_raw[t * N + i], ws);
        const double* __restrict__ x = part + (t * N + i) * d;
#pragma unroll
        for (int c = 0; c < 12; c++)
            if (c < d) acc[c] = fma(w, x[c], acc[c]);
    }
#pragma unroll
    for (int c = 0; c < 12; c++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
        if (lane == 0) red[wid][c] = acc[c];
    }
    __syncthreads();
    if (tid < d) {
        double s = 0.0;
        for (int q = 0; q < 8; q++) s += red[q][tid];
        est[t * d + tid] = s;
    }
}

template <class Tp>
static int pf_in_ptr(mkf_pf2d* p, const Tp* ptr, size_t count, int mem, DevBuf& stage, const Tp** out)
{
    if (!ptr) {
        *out = nullptr;
        return MKF_OK;
    }
    if (is_device_ptr(ptr, mem)) {
        *out = ptr;
        return MKF_OK;
    }
    int rc = stage.ensure(count * sizeof(Tp));
    if (rc) return rc;
    CK(cudaMemcpyAsync(stage.p, ptr, count * sizeof(Tp), cudaMemcpyHostToDevice, p->stream));
    *out = (const Tp*)stage.p;
    return MKF_OK;
}

extern "C" int mkf_pf2d_set_particles(mkf_pf2d* p, const double* particles, int mem)
{
    if (!p || !particles) {
        mkf_set_error("mkf_pf2d_set_particles: null argument");
        return MKF_E_INVALID;
    }
    CK(cudaSetDevice(p->device));
    const size_t bytes = (size_t)p->T * p->N * p->d * 8;
    CK(cudaMemcpyAsync(p->part[p->cur], particles, bytes,
                       is_device_ptr(particles, mem) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return MKF_OK;
}

extern "C" int mkf_pf2d_set_random(mkf_pf2d* p, uint64_t seed, int64_t track0, const uint8_t* side, int im_width,
                                   int im_height)
{
    if (!p || im_width < 2 || im_height < 2) {
        mkf_set_error("mkf_pf2d_set_random: invalid argument");
        return MKF_E_INVALID;
    }
    CK(cudaSetDevice(p->device));
    p->rnd.seed = seed;
    p->rnd.track0 = track0;
    p->rnd.im_w = im_width;
    p->rnd.im_h = im_height;
    p->rnd.side = nullptr;
    if (side) {
        if (!p->d_side) CK(cudaMalloc((void**)&p->d_side, (size_t)p->T));
        CK(cudaMemcpyAsync(p->d_side, side, (size_t)p->T, cudaMemcpyDefault, p->stream));
        CK(cudaStreamSynchronize(p->stream)); // `side` may be pageable host memory
        p->rnd.side = p->d_side;
    }
    return MKF_OK;
}

extern "C" int mkf_pf2d_randomise(mkf_pf2d* p)
{
    if (!p) {
        mkf_set_error("null pf2d");
        return MKF_E_INVALID;
    }
    CK(cudaSetDevice(p->device));
    p->rnd.epoch = 0;
    const long long n = p->T * p->N * p->d;
    k_pf2d_randomise<<<grid_for(n, 256), 256, 0, p->stream>>>(p->part[p->cur], p->T, p->N, p->d, p->rnd);
    MKF_LAUNCHED();
    CK(cudaGetLastError());
    // uniform weights (src/pf2D.cpp:53-56): wsum = 0 makes k_pf2d_estimate take 1/N
    CK(cudaMemsetAsync(p->wsum, 0, (size_t)p->T * 8, p->stream));
    CK(cudaMemsetAsync(p->status, 0, (size_t)p->T * 4, p->stream));
    return MKF_OK;
}

extern "C" int mkf_pf2d_update(mkf_pf2d* p, const double* meas, const double* u, const double* noise, int mem)
{
    if (!p || !meas || !u) {
        mkf_set_error("mkf_pf2d_update: null argument");
        return MKF_E_INVALID;
    }
    CK(cudaSetDevice(p->device));
    const double *d_meas, *d_u, *d_noise;
    int rc;
    const long long tot = p->T * p->N;
    if ((rc = pf_in_ptr(p, meas, (size_t)p->T * 4, mem, p->in_meas, &d_meas))) return rc;
    if ((rc = pf_in_ptr(p, u, (size_t)p->T, mem, p->in_u, &d_u))) return rc;
    if ((rc = pf_in_ptr(p, noise, (size_t)tot * p->d, mem, p->in_noise, &d_noise))) return rc;
    CK(cudaMemsetAsync(p->status, 0, (size_t)p->T * 4, p->stream));
    cudaEvent_t* pe = (p->prof_on && p->prof_n < p->prof_cap) ? &p->prof_ev[(size_t)p->prof_n * 4] : nullptr;
    if (pe) cudaEventRecord(pe[0], p->stream);
    const size_t smem = (size_t)p->K * p->gstride * sizeof(double);
    if (smem > 200 * 1024) {
        mkf_set_error("mkf_pf2d_update: GMM too large for shared memory");
        return MKF_E_UNSUPPORTED;
    }
#define LAUNCH_W(DD)                                                                                         \
    do {                                                                                                     \
        if (smem > 48 * 1024)                                                                                \
            CK(cudaFuncSetAttribute(k_pf2d_weight<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k_pf2d_weight<DD><<<grid_for(tot, 128), 128, smem, p->stream>>>(p->part[p->cur], d_meas, p->gmm, p->K,  \
                                                                       p->gstride, p->T, p->N, p->w_raw);    \
    } while (0)
    if (p->d == 8)
        LAUNCH_W(8);
    else if (p->d == 10)
        LAUNCH_W(10);
    else
        LAUNCH_W(12);
#undef LAUNCH_W
    MKF_LAUNCHED();
    CK(cudaGetLastError());
    if (pe) cudaEventRecord(pe[1], p->stream);
    if ((rc = run_resample(p->stream, p->T, p->w_raw, p->N, p->N, d_u, 1, 1, p->wsum, p->parent, p->status,
                           nullptr, 1, 0, MKF_ST_POST_FALLBACK, MKF_ST_POST_DEGENERATE)))
        return rc;
    if (pe) cudaEventRecord(pe[2], p->stream);
    p->rnd.epoch++; // epoch n = the n-th update (the degenerate branch's draws are keyed on it)
    if (p->d == 8)
        k_pf2d_resample_predict_v<8><<<grid_for(tot, 256), 256, 0, p->stream>>>(p->part[p->cur], p->part[p->cur ^ 1],
                                                                                p->parent, p->status, d_noise, p->T, p->N,
                                                                                p->rnd);
    else if (p->d == 12)
        k_pf2d_resample_predict_v<12><<<grid_for(tot, 256), 256, 0, p->stream>>>(p->part[p->cur], p->part[p->cur ^ 1],
                                                                                 p->parent, p->status, d_noise, p->T, p->N,
                                                                                 p->rnd);
    else
        k_pf2d_resample_predict<<<grid_for(tot * p->d, 256), 256, 0, p->stream>>>(p->part[p->cur], p->part[p->cur ^ 1],
                                                                                  p->parent, p->status, d_noise, p->T,
                                                                                  p->N, p->d, p->rnd);
    MKF_LAUNCHED();
    CK(cudaGetLastError());
    if (pe) {
        cudaEventRecord(pe[3], p->stream);
        p->prof_n++;
    }
    p->cur ^= 1;
    return MKF_OK;
}

// per-kernel device timing of mkf_pf2d_update (CUDA events on the filter's stream; bench.py)
extern "C" int mkf_pf2d_profile(mkf_pf2d* p, int max_updates)
{
    if (!p || max_updates < 0) {
        mkf_set_error("mkf_pf2d_profile: invalid argument");
        return MKF_E_INVALID;
    }
    CK(cudaSetDevice(p->device));
    CK(cudaStreamSynchronize(p->stream));
    for (cudaEvent_t e : p->prof_ev) cudaEventDestroy(e);
    p->prof_ev.clear();
    p->prof_n = 0;
    p->prof_cap = max_updates;
    p->prof_on = max_updates > 0;
    for (int i = 0; i < max_updates * 4; i++) {
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        p->prof_ev.push_back(e);
    }
    return MKF_OK;
}
extern "C" int mkf_pf2d_profile_read(mkf_pf2d* p, double* ms /* 3: weights, normalise+resample, gather+predict */,
                                     int* n_updates)
{
    if (!p || !ms) {
        mkf_set_error("mkf_pf2d_profile_read: null argument");
        return MKF_E_INVALID;
    }
    CK(cudaSetDevice(p->device));
    CK(cudaStreamSynchronize(p->stream));
    ms[0] = ms[1] = ms[2] = 0.0;
    for (int i = 0; i < p->prof_n; i++)
        for (int k = 0; k < 3; k++) {
            float t = 0;
            CK(cudaEventElapsedTime(&t, p->prof_ev[(size_t)i * 4 + k], p->prof_ev[(size_t)i * 4 + k + 1]));
            ms[k] += t;
        }
    if (n_updates) *n_updates = p->prof_n;
    p->prof_n = 0;
    return MKF_OK;
}

extern "C" int mkf_pf2d_get(mkf_pf2d* p, double* particles, double* w_norm, int32_t* parents, int mem)
{
    if (!p) {
        mkf_set_error("null pf2d");
        return MKF_E_INVALID;
    }
    CK(cudaSetDevice(p->device));
    const size_t tot = (size_t)p->T * p->N;
    auto kind = [&](void* dst) { return is_device_ptr(dst, mem) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost; };
    if (particles) CK(cudaMemcpyAsync(particles, p->part[p->cur], tot * p->d * 8, kind(particles), p->stream));
    if (parents) CK(cudaMemcpyAsync(parents, p->parent, tot * 4, kind(parents), p->stream));
    DevBuf tmp;
    if (w_norm) {
        double* dst = w_norm;
        const bool host = !is_device_ptr(w_norm, mem);
        if (host) {
            int rc = tmp.ensure(tot * 8);
            if (rc) return rc;
            dst = (double*)tmp.p;
        }
        k_aux_outputs<<<grid_for((long long)tot, 256), 256, 0, p->stream>>>(p->w_raw, p->wsum, nullptr, (long long)tot,
                                                                           p->N, 0, dst, nullptr, nullptr);
        MKF_LAUNCHED();
        if (host) cudaMemcpyAsync(w_norm, dst, tot * 8, cudaMemcpyDeviceToHost, p->stream);
    }
    cudaError_t e = cudaStreamSynchronize(p->stream);
    tmp.release();
    if (e != cudaSuccess) {
        mkf_set_error("mkf_pf2d_get: %s", cudaGetErrorString(e));
        return MKF_E_CUDA;
    }
    return MKF_OK;
}

extern "C" int mkf_pf2d_estimate(mkf_pf2d* p, double* est, int mem)
{
    if (!p || !est) {
        mkf_set_error("mkf_pf2d_estimate: null argument");
        return MKF_E_INVALID;
    }
    CK(cudaSetDevice(p->device));
    const bool host = !is_device_ptr(est, mem);
    const size_t bytes = (size_t)p->T * p->d * 8;
    DevBuf tmp;
    double* dst = est;
    if (host) {
        int rc = tmp.ensure(bytes);
        if (rc) return rc;
        dst = (double*)tmp.p;
    }
    k_pf2d_estimate<<<(unsigned)p->T, 256, 0, p->stream>>>(p->part[p->cur], p->w_raw, p->wsum, p->status, p->N, p->d, dst);
    MKF_LAUNCHED();
    CK(cudaGetLastError());
    if (host) {
        CK(cudaMemcpyAsync(est, dst, bytes, cudaMemcpyDeviceToHost, p->stream));
        cudaError_t e = cudaStreamSynchronize(p->stream);
        tmp.release();
        if (e != cudaSuccess) {
            mkf_set_error("mkf_pf2d_estimate: %s", cudaGetErrorString(e));
            return MKF_E_CUDA;
        }
    }
    return MKF_OK;
}

extern "C" int mkf_pf2d_sync(mkf_pf2d* p)
{
    if (!p) {
        mkf_set_error("null pf2d");
        return MKF_E_INVALID;
    }
    CK(cudaSetDevice(p->device));
    CK(cudaStreamSynchronize(p->stream));
    return MKF_OK;
}

#endif
