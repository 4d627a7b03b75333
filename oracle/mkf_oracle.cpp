// mkf_oracle.cpp -- CPU restatement of the per-frame filtering hot path of
// mgb45/mkfbodytracker_pdaf.  TEST INFRASTRUCTURE ONLY (see mkf_oracle.h): the product
// library never links or calls this file.
//
// PINNING.  The reference ships no golden vectors / tests, so the anchor is the reference
// ITSELF run here: oracle/_ref/libref.so is the reference's own src/{KF_model,my_gmm,pf2DRao,pfPose}.cpp compiled
// in place (oracle/Makefile) against oracle/cvshim, and tests/test_ref_sources.py + tests/test_ref_tracker.py
// require this file to agree with it BIT FOR BIT (filter states, weights, indices, poses) through whole frames of
// the node.  What remains unpinned is the third-party arithmetic under both: OpenCV `core` (unpinned, 2.4-era API,
// not installed here) is restated -- here and in cvshim -- from the published OpenCV 2.4.x algorithms
// (GEMMSingleMul operation order, LUImpl, CholImpl, cv::RNG), from memory of the OpenCV sources; that restatement
// only influences results at O(1e-16) relative and is cross-checked against OpenCV-python 4.13 where exported.
//
// Everything is IEEE double evaluated in the order the reference's cv::MatExpr tree evaluates
// it; build with -ffp-contract=off so that no multiply-add is fused.
//
// File:line citations refer to /root/reference/.

#include "mkf_oracle.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/mkf_synth.h"
#include "../include/mkf_expf.h"

namespace {

// ---------------------------------------------------------------------------------------------
// OpenCV core primitives, restated
// ---------------------------------------------------------------------------------------------

// cv::gemm for CV_64F, the GEMMSingleMul path (all matrices here are far below the blocking
// thresholds).  D (ar x bc) = alpha * A(ar x ac) * op(B) + beta * C.
// A*B: every output is the plain k-ordered sum; A*B^T (GEMM_2_T): four interleaved partial
// sums over k, combined as (s0+s1+s2+s3)*alpha.  D must not alias A or B (cv::gemm uses a
// temporary in that case; callers pass a scratch buffer).
void gemm_nn(const double* A, int ar, int ac, const double* B, int bc, double alpha, const double* C, double beta,
             double* D)
{
    for (int i = 0; i < ar; i++) {
        const double* a = A + (size_t)i * ac;
        for (int j = 0; j < bc; j++) {
            double s = 0;
            for (int k = 0; k < ac; k++) s += a[k] * B[(size_t)k * bc + j];
            s = s * alpha;
            D[(size_t)i * bc + j] = C ? s + C[(size_t)i * bc + j] * beta : s;
        }
    }
}

// D (ar x br) = alpha * A(ar x ac) * B(br x ac)^T + beta * C
void gemm_nt(const double* A, int ar, int ac, const double* B, int br, double alpha, const double* C, double beta,
             double* D)
{
    for (int i = 0; i < ar; i++) {
        const double* a = A + (size_t)i * ac;
        for (int j = 0; j < br; j++) {
            const double* b = B + (size_t)j * ac;
            double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
            int k = 0;
            for (; k <= ac - 4; k += 4) {
                s0 += a[k] * b[k];
                s1 += a[k + 1] * b[k + 1];
                s2 += a[k + 2] * b[k + 2];
                s3 += a[k + 3] * b[k + 3];
            }
            for (; k < ac; k++) s0 += a[k] * b[k];
            s0 = (s0 + s1 + s2 + s3) * alpha;
            D[(size_t)i * br + j] = C ? s0 + C[(size_t)i * br + j] * beta : s0;
        }
    }
}

// cv::LU (LUImpl<double>, OpenCV 2.4.x) applied as cv::invert(DECOMP_LU) does for n > 3:
// A is destroyed, b starts as the identity and ends as the inverse.  returns 0 when singular.
int lu_impl(double* A, int m, double* b, int n)
{
    int p = 1;
    for (int i = 0; i < m; i++) {
        int k = i;
        for (int j = i + 1; j < m; j++)
            if (std::abs(A[j * m + i]) > std::abs(A[k * m + i])) k = j;
        if (std::abs(A[k * m + i]) < std::numeric_limits<double>::epsilon()) return 0;
        if (k != i) {
            for (int j = i; j < m; j++) std::swap(A[i * m + j], A[k * m + j]);
            if (b)
                for (int j = 0; j < n; j++) std::swap(b[i * n + j], b[k * n + j]);
            p = -p;
        }
        double d = -1 / A[i * m + i];
        for (int j = i + 1; j < m; j++) {
            double alpha = A[j * m + i] * d;
            for (k = i + 1; k < m; k++) A[j * m + k] += alpha * A[i * m + k];
            if (b)
                for (k = 0; k < n; k++) b[j * n + k] += alpha * b[i * n + k];
        }
        A[i * m + i] = -d;
    }
    if (b) {
        for (int i = m - 1; i >= 0; i--)
            for (int j = 0; j < n; j++) {
                double s = b[i * n + j];
                for (int k = i + 1; k < m; k++) s -= A[i * m + k] * b[k * n + j];
                b[i * n + j] = s * A[i * m + i];
            }
    }
    return p;
}

// cv::invert(src, dst, DECOMP_LU) for a square CV_64F matrix: closed forms for n <= 3
// (only n == 2 occurs on this path), LU otherwise.  On failure dst is zeroed (cv::invert does
// `dst = Scalar(0)`).
int invert_lu(int n, const double* in, double* out)
{
    if (n == 2) {
        double d = in[0] * in[3] - in[1] * in[2];
        if (d != 0.) {
            d = 1. / d;
            double t0 = in[0] * d, t1 = in[3] * d;
            out[3] = t0;
            out[0] = t1;
            t0 = -in[1] * d;
            t1 = -in[2] * d;
            out[1] = t0;
            out[2] = t1;
            return 1;
        }
        std::fill(out, out + 4, 0.0);
        return 0;
    }
    if (n == 1) {
        if (in[0] != 0.) {
            out[0] = 1. / in[0];
            return 1;
        }
        out[0] = 0;
        return 0;
    }
    std::vector<double> a(in, in + (size_t)n * n);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) out[i * n + j] = (i == j) ? 1.0 : 0.0;
    int ok = lu_impl(a.data(), n, out, n) != 0;
    if (!ok) std::fill(out, out + (size_t)n * n, 0.0);
    return ok;
}

// cv::Cholesky(double* A, step, m, b = NULL, ...) (CholImpl<double>): in place on the lower
// triangle; the upper triangle is left untouched.  cv24 == true : diagonal left as 1/L_ii
// (OpenCV 2.4.x); cv24 == false: diagonal re-inverted to L_ii (OpenCV >= 3.0).
bool cv_cholesky(double* A, int m, bool cv24)
{
    for (int i = 0; i < m; i++) {
        int j;
        double s;
        for (j = 0; j < i; j++) {
            s = A[i * m + j];
            for (int k = 0; k < j; k++) s -= A[i * m + k] * A[j * m + k];
            A[i * m + j] = s * A[j * m + j];
        }
        s = A[i * m + i];
        for (int k = 0; k < j; k++) {
            double t = A[i * m + k];
            s -= t * t;
        }
        if (s < std::numeric_limits<double>::epsilon()) return false;
        A[i * m + i] = 1. / std::sqrt(s);
    }
    if (!cv24)
        for (int i = 0; i < m; i++) A[i * m + i] = 1 / A[i * m + i];
    return true;
}

// cv::RNG (multiply-with-carry), OpenCV core/operations.hpp
struct CvRng {
    uint64_t state;
    explicit CvRng(uint64_t s) : state(s ? s : 0xffffffffull) {}
    unsigned next()
    {
        state = (uint64_t)(unsigned)state * 4164903690U + (unsigned)(state >> 32);
        return (unsigned)state;
    }
    int uniform_int(int a, int b) { return a == b ? a : (int)(next() % (unsigned)(b - a) + a); }
    double uniform_dbl(double a, double b)
    {
        unsigned t = next();
        double v = (double)(((uint64_t)t << 32) | next()) * 5.4210108624275221700372640043497e-20;
        return v * (b - a) + a;
    }
};

// ---------------------------------------------------------------------------------------------
// ParticleFilter::chol  (src/pf2DRao.cpp:34-53)
// ---------------------------------------------------------------------------------------------
// sigma_i = in.clone(); if (Cholesky(...)) { for each e: row(e) *= diag(e); at(e,e) = 1/diag(e);
// if (e>0) zero the e-th sub-diagonal }.  `diagElem` is a *view*, so `elem` is read after the
// previous rows were modified -- only row e's own diagonal matters, which is untouched until
// step e.  On failure the (partially factored: cv::Cholesky works in place on the clone)
// matrix is returned as is.
bool chol_wrapper(int n, const double* in, double* out, int mode)
{
    std::memcpy(out, in, sizeof(double) * n * n);
    if (mode == ORC_CHOL_EXACT) {
        // the author's evident intent (MATLAB mvnpdf): R = upper Cholesky factor
        std::vector<double> a(in, in + (size_t)n * n);
        if (!cv_cholesky(a.data(), n, false)) return false;
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) out[i * n + j] = (j >= i) ? a[j * n + i] : 0.0;
        return true;
    }
    if (!cv_cholesky(out, n, mode == ORC_CHOL_CV24_LITERAL)) return false;
    for (int e = 0; e < n; e++) {
        double elem = out[e * n + e];
        for (int j = 0; j < n; j++) out[e * n + j] *= elem; // sigma_i.row(e) *= elem
        out[e * n + e] = 1.0 / elem;
        if (e > 0)
            for (int i = 0; i + e < n; i++) out[(i + e) * n + i] = 0.0; // zeros -> diag(-e)
    }
    return true;
}

// ---------------------------------------------------------------------------------------------
// ParticleFilter::mvnpdf  (src/pf2DRao.cpp:56-67)
// ---------------------------------------------------------------------------------------------
double mvnpdf(int n, const double* x, const double* u, const double* sigma, int mode, int* chol_ok)
{
    double R[36], Rinv[36], xu[6], v[6];
    bool ok = chol_wrapper(n, sigma, R, mode);
    if (chol_ok) *chol_ok = ok ? 1 : 0;
    invert_lu(n, R, Rinv);                       // R.inv()
    for (int i = 0; i < n; i++) xu[i] = x[i] - u[i]; // (x - u)
    gemm_nn(xu, 1, n, Rinv, n, 1.0, nullptr, 0.0, v); // (x-u).t() * R.inv()  (1 x n)
    double lsd = 0;                                   // cv::sum(log(R.diag(0)))
    for (int i = 0; i < n; i++) lsd += std::log(R[i * n + i]);
    // cv::pow(x_u,2) then reduce(.,1,CV_REDUCE_SUM): two interleaved accumulators (reduceC_)
    double q;
    if (n == 1) {
        q = v[0] * v[0];
    } else {
        double a0 = v[0] * v[0], a1 = v[1] * v[1];
        int i = 2;
        for (; i <= n - 2; i += 2) {
            a0 = a0 + v[i] * v[i];
            a1 = a1 + v[i + 1] * v[i + 1];
        }
        for (; i < n; i++) a0 = a0 + v[i] * v[i];
        q = a0 + a1;
    }
    return std::exp(-0.5 * q - lsd - n * std::log(2.0 * M_PI) / 2.0);
}

// ParticleFilter::mvnpdf_multiple (src/pf2DRao.cpp:69-83) for the 2-D proposal density:
// x is 2 x C row-major (row 0 = x coordinates), u 2, sigma 2 x 2; out C.
void mvnpdf_multiple2(int C, const double* x, const double* u, const double* sigma, int mode, double* out)
{
    double R[4], Rinv[4];
    chol_wrapper(2, sigma, R, mode);
    invert_lu(2, R, Rinv);
    double lsd = std::log(R[0]);
    lsd += std::log(R[3]);
    double shift = -lsd - 2 * std::log(2 * M_PI) / 2; // scalar part of the MatExpr
    for (int c = 0; c < C; c++) {
        double dx = x[c] - u[0], dy = x[C + c] - u[1];
        double v0 = dx * Rinv[0] + dy * Rinv[2];
        double v1 = dx * Rinv[1] + dy * Rinv[3];
        double q = v0 * v0 + v1 * v1;
        out[c] = std::exp(q * -0.5 + shift);
    }
}

// ---------------------------------------------------------------------------------------------
// ParticleFilter::maxWeight + resample  (src/pf2DRao.cpp:161-210)
// ---------------------------------------------------------------------------------------------
int resample(const double* w, int L, int N, double u, uint64_t seed, int32_t* out)
{
    CvRng rng(seed);                 // cv::RNG rng(cv::getTickCount());
    int idx = rng.uniform_int(0, L); // drawn, then unused
    (void)idx;
    double mw = 0;
    for (int i = 0; i < L; i++)
        if (w[i] > mw) mw = w[i];
    if (mw == 0) {
        for (int i = 0; i < N; i++) out[i] = rng.uniform_int(0, L);
        return 1;
    }
    idx = 0;
    double step = 1.0 / (double)N;
    double draw = (u >= 0.0) ? u : rng.uniform_dbl(0.0, 1.0);
    double beta = draw * step;
    for (int i = 0; i < N; i++) {
        while (beta > w[idx]) {
            beta -= w[idx];
            idx = (idx + 1) % L;
        }
        beta += step;
        out[i] = idx;
    }
    return 0;
}

} // namespace

// ---------------------------------------------------------------------------------------------
// my_gmm::loadGaussian  (src/my_gmm.cpp:45-75) -- the model
// ---------------------------------------------------------------------------------------------
struct orc_model {
    int K, d, D;
    std::vector<double> mean;   // K x d  (gmm.mean[k], 1 x d rows)
    std::vector<double> cov;    // K x d x d
    std::vector<double> weight; // K
    std::vector<double> Q, F;   // K x d x d
    std::vector<double> B;      // K x d
    std::vector<double> R;      // 6 x 6
    std::vector<double> H;      // 6 x d
    std::vector<double> BH;     // 6
    std::vector<double> proj;   // d x D (h_pca)
    std::vector<double> pmean;  // D     (m_pca)
};

extern "C" orc_model* orc_model_create(int K, int d, int D, const double* means, const double* covs,
                                       const double* weights, const double* gamma, const double* pca_proj,
                                       const double* pca_mean)
{
    if (K <= 0 || d <= 0 || d > 12 || D < 14 || D > 64) return nullptr;
    orc_model* m = new orc_model;
    m->K = K;
    m->d = d;
    m->D = D;
    m->mean.assign(means, means + (size_t)K * d);
    m->cov.assign(covs, covs + (size_t)K * d * d);
    m->weight.assign(weights, weights + K);
    m->proj.assign(pca_proj, pca_proj + (size_t)d * D);
    m->pmean.assign(pca_mean, pca_mean + D);
    m->Q.resize((size_t)K * d * d);
    m->F.assign((size_t)K * d * d, 0.0);
    m->B.resize((size_t)K * d);
    m->R.assign(36, 0.0);
    for (int i = 0; i < 6; i++) m->R[i * 6 + i] = 100 * 1.0; // 100*eye(6,6)
    // H1 = zeros(6, D) with ones at (0,9) (1,10) (2,0) (3,1) (4,12) (5,13)
    std::vector<double> H1((size_t)6 * D, 0.0);
    const int sel[6] = {9, 10, 0, 1, 12, 13};
    for (int r = 0; r < 6; r++) H1[(size_t)r * D + sel[r]] = 1;
    m->H.resize((size_t)6 * d);
    m->BH.resize(6);
    // tracker.H = H1*H.t()  (gemm, GEMM_2_T);  tracker.BH = H1*m.t()
    gemm_nt(H1.data(), 6, D, m->proj.data(), d, 1.0, nullptr, 0.0, m->H.data());
    gemm_nt(H1.data(), 6, D, m->pmean.data(), 1, 1.0, nullptr, 0.0, m->BH.data());
    for (int k = 0; k < K; k++) {
        double g = gamma[k];
        for (int e = 0; e < d * d; e++) m->Q[(size_t)k * d * d + e] = m->cov[(size_t)k * d * d + e] * (1 - g * g);
        for (int i = 0; i < d; i++) m->F[(size_t)k * d * d + i * d + i] = 1.0 * g;
        for (int i = 0; i < d; i++) m->B[(size_t)k * d + i] = m->mean[(size_t)k * d + i] * (1.0 - g);
    }
    return m;
}

extern "C" void orc_model_destroy(orc_model* m) { delete m; }

extern "C" void orc_model_get(const orc_model* m, double* H, double* BH, double* Q, double* B, double* R)
{
    if (H) std::copy(m->H.begin(), m->H.end(), H);
    if (BH) std::copy(m->BH.begin(), m->BH.end(), BH);
    if (Q) std::copy(m->Q.begin(), m->Q.end(), Q);
    if (B) std::copy(m->B.begin(), m->B.end(), B);
    if (R) std::copy(m->R.begin(), m->R.end(), R);
}

// ---------------------------------------------------------------------------------------------
// KF_model::predict / update  (src/KF_model.cpp:11-25)
// ---------------------------------------------------------------------------------------------
namespace {

void kf_predict(const orc_model* m, int k, double* x, double* P)
{
    const int d = m->d;
    const double* F = &m->F[(size_t)k * d * d];
    double t1[12], t2[144], t3[144];
    // state = F*state + B           -> gemm(F, state, 1, B, 1)
    gemm_nn(F, d, d, x, 1, 1.0, &m->B[(size_t)k * d], 1.0, t1);
    std::memcpy(x, t1, sizeof(double) * d);
    // cov = F*cov*F.t() + Q         -> tmp = F*cov; gemm(tmp, F, 1, Q, 1, GEMM_2_T)
    gemm_nn(F, d, d, P, d, 1.0, nullptr, 0.0, t2);
    gemm_nt(t2, d, d, F, d, 1.0, &m->Q[(size_t)k * d * d], 1.0, t3);
    std::memcpy(P, t3, sizeof(double) * d * d);
}

// zhat = H*state + BH, S = H*cov*H.t() + R  (src/pf2DRao.cpp:138 and src/KF_model.cpp:19-20)
void innovation_stats(const orc_model* m, const double* x, const double* P, double* zhat, double* S)
{
    const int d = m->d;
    double HP[72];
    gemm_nn(m->H.data(), 6, d, x, 1, 1.0, m->BH.data(), 1.0, zhat);
    gemm_nn(m->H.data(), 6, d, P, d, 1.0, nullptr, 0.0, HP);
    gemm_nt(HP, 6, d, m->H.data(), 6, 1.0, m->R.data(), 1.0, S);
}

void kf_update(const orc_model* m, const double* z, double* x, double* P)
{
    const int d = m->d;
    double zhat[6], S[36], Sinv[36], y[6], PHt[72], K[72], t1[12], KH[144], IKH[144], t3[144];
    innovation_stats(m, x, P, zhat, S);
    for (int i = 0; i < 6; i++) y[i] = z[i] - zhat[i]; // y = measurement - (H*state + BH)
    // K = cov*H.t()*S.inv()        -> tmp = gemm(cov, H, GEMM_2_T); inv = LU; gemm(tmp, inv)
    gemm_nt(P, d, d, m->H.data(), 6, 1.0, nullptr, 0.0, PHt);
    invert_lu(6, S, Sinv);
    gemm_nn(PHt, d, 6, Sinv, 6, 1.0, nullptr, 0.0, K);
    // state = state + K*y          -> gemm(K, y, 1, state, 1)
    gemm_nn(K, d, 6, y, 1, 1.0, x, 1.0, t1);
    std::memcpy(x, t1, sizeof(double) * d);
    // cov = (eye - K*H)*cov
    gemm_nn(K, d, 6, m->H.data(), d, 1.0, nullptr, 0.0, KH);
    for (int i = 0; i < d; i++)
        for (int j = 0; j < d; j++) IKH[i * d + j] = ((i == j) ? 1.0 : 0.0) - KH[i * d + j];
    gemm_nn(IKH, d, d, P, d, 1.0, nullptr, 0.0, t3);
    std::memcpy(P, t3, sizeof(double) * d * d);
}

} // namespace

extern "C" void orc_kf_predict(const orc_model* m, int k, double* x, double* P) { kf_predict(m, k, x, P); }
extern "C" void orc_kf_update(const orc_model* m, int k, const double* z, double* x, double* P)
{
    (void)k; // H, BH, R are the same for every component (src/my_gmm.cpp:54,61-72)
    kf_update(m, z, x, P);
}
extern "C" double orc_mvnpdf(int n, const double* x, const double* u, const double* sigma, int chol_mode, int* chol_ok)
{
    if (n < 1 || n > 6) return std::numeric_limits<double>::quiet_NaN();
    return mvnpdf(n, x, u, sigma, chol_mode, chol_ok);
}
extern "C" int orc_chol(int n, const double* in, double* out, int chol_mode)
{
    return chol_wrapper(n, in, out, chol_mode) ? 1 : 0;
}
extern "C" int orc_invert_lu(int n, const double* in, double* out) { return invert_lu(n, in, out); }
extern "C" int orc_resample(const double* w, int L, int N, double u, uint64_t seed, int32_t* out)
{
    return resample(w, L, N, u, seed, out);
}
extern "C" void orc_cvrng(uint64_t seed, int L, int n_int, int32_t* out_int, int n_dbl, double* out_dbl)
{
    CvRng rng(seed);
    for (int i = 0; i < n_int; i++) out_int[i] = rng.uniform_int(0, L);
    for (int i = 0; i < n_dbl; i++) out_dbl[i] = rng.uniform_dbl(0.0, 1.0);
}

// ---------------------------------------------------------------------------------------------
// ParticleFilter (pf2DRao)  (src/pf2DRao.cpp:13-31, 125-158; src/my_gmm.cpp:11-16, 30-42)
// ---------------------------------------------------------------------------------------------
struct orc_filter {
    const orc_model* m;
    int N, chol_mode, alias_mode;
    // buffer pool: pool_x[b], pool_P[b]; slot j uses buffer buf[j].  INDEPENDENT: buf[j]==j.
    std::vector<double> pool_x, pool_P, tmp_x, tmp_P;
    std::vector<int> buf;
    std::vector<double> w;
    std::vector<int32_t> ind;
};

extern "C" orc_filter* orc_filter_create(const orc_model* m, int N, int chol_mode, int alias_mode)
{
    if (!m || N <= 0) return nullptr;
    orc_filter* f = new orc_filter;
    f->m = m;
    f->N = N;
    f->chol_mode = chol_mode;
    f->alias_mode = alias_mode;
    const int d = m->d;
    f->pool_x.assign((size_t)N * d, 0.0);
    f->pool_P.assign((size_t)N * d * d, 0.0);
    f->tmp_x.assign((size_t)N * d, 0.0);
    f->tmp_P.assign((size_t)N * d * d, 0.0);
    f->buf.resize(N);
    for (int j = 0; j < N; j++) f->buf[j] = j;
    f->w.assign(N, 0.0);
    f->ind.assign(N, 0);
    return f;
}
extern "C" void orc_filter_destroy(orc_filter* f) { delete f; }

extern "C" int orc_filter_reset(orc_filter* f, double u, uint64_t seed)
{
    const orc_model* m = f->m;
    const int d = m->d, N = f->N;
    std::vector<int32_t> bins(N);
    int deg = resample(m->weight.data(), m->K, N, u, seed, bins.data());
    // resetTracker: tracks[i] = { mean[bins[i]].t(), cov[bins[i]], 1/N }, deep-copied by push_back
    for (int j = 0; j < N; j++) {
        std::memcpy(&f->pool_x[(size_t)j * d], &m->mean[(size_t)bins[j] * d], sizeof(double) * d);
        std::memcpy(&f->pool_P[(size_t)j * d * d], &m->cov[(size_t)bins[j] * d * d], sizeof(double) * d * d);
        f->buf[j] = j;
    }
    return deg;
}

static int filter_update_impl(orc_filter* f, const double* meas, int meas_stride_row, int meas_stride_col,
                              double u_ind, uint64_t seed_ind, double u_post, uint64_t seed_post, double* w_raw,
                              double* w_norm, int32_t* indicators, int32_t* parents, double* wsum_out)
{
    const orc_model* m = f->m;
    const int d = m->d, N = f->N;
    int status = 0;
    // std::vector<int> indicators = resample(gmm.weight, gmm.nParticles);
    if (resample(m->weight.data(), m->K, N, u_ind, seed_ind, f->ind.data())) status |= 1;
    if (indicators) std::copy(f->ind.begin(), f->ind.end(), indicators);
    double wsum = 0;
    for (int j = 0; j < N; j++) {
        const int i = f->ind[j];
        const int b = f->buf[j];
        double* x = &f->pool_x[(size_t)b * d];
        double* P = &f->pool_P[(size_t)b * d * d];
        double z[6], zhat[6], S[36];
        for (int r = 0; r < 6; r++) z[r] = meas[(size_t)r * meas_stride_row + (size_t)j * meas_stride_col];
        kf_predict(m, i, x, P); // gmm.KFtracker[i].predict(state, cov)
        innovation_stats(m, x, P, zhat, S);
        int ok = 1;
        f->w[j] = mvnpdf(6, z, zhat, S, f->chol_mode, &ok); // weights.push_back(mvnpdf(...))
        if (!ok) status |= 4;
        wsum = wsum + f->w[j];
        kf_update(m, z, x, P); // gmm.KFtracker[i].update(measurement.col(j), state, cov)
        // temp.push_back(gmm.tracks[j]) -- deep copy (state_params copy ctor, src/my_gmm.cpp:11-16)
        std::memcpy(&f->tmp_x[(size_t)j * d], x, sizeof(double) * d);
        std::memcpy(&f->tmp_P[(size_t)j * d * d], P, sizeof(double) * d * d);
    }
    if (w_raw) std::copy(f->w.begin(), f->w.end(), w_raw);
    if (wsum_out) *wsum_out = wsum;
    for (int i = 0; i < N; i++) f->w[i] = f->w[i] / wsum;
    if (w_norm) std::copy(f->w.begin(), f->w.end(), w_norm);
    // indicators = resample(weights, gmm.nParticles); tracks[j] = temp[indicators[j]]
    if (resample(f->w.data(), N, N, u_post, seed_post, f->ind.data())) status |= 2;
    if (parents) std::copy(f->ind.begin(), f->ind.end(), parents);
    if (f->alias_mode == ORC_ALIAS_CV_SHALLOW_LITERAL) {
        // implicit operator= is a shallow cv::Mat assignment: slot j now *shares* temp[parent]'s
        // buffers; duplicates of a parent are chained in place on the next frame (quirk B3).
        f->pool_x.swap(f->tmp_x);
        f->pool_P.swap(f->tmp_P);
        for (int j = 0; j < N; j++) f->buf[j] = f->ind[j];
    } else {
        for (int j = 0; j < N; j++) {
            std::memcpy(&f->pool_x[(size_t)j * d], &f->tmp_x[(size_t)f->ind[j] * d], sizeof(double) * d);
            std::memcpy(&f->pool_P[(size_t)j * d * d], &f->tmp_P[(size_t)f->ind[j] * d * d],
                        sizeof(double) * d * d);
            f->buf[j] = j;
        }
    }
    return status;
}

extern "C" int orc_filter_update(orc_filter* f, const double* meas, double u_ind, uint64_t seed_ind, double u_post,
                                 uint64_t seed_post, double* w_raw, double* w_norm, int32_t* indicators,
                                 int32_t* parents, double* wsum)
{
    return filter_update_impl(f, meas, f->N, 1, u_ind, seed_ind, u_post, seed_post, w_raw, w_norm, indicators,
                              parents, wsum);
}
extern "C" int orc_filter_update_shared(orc_filter* f, const double* z6, double u_ind, uint64_t seed_ind,
                                        double u_post, uint64_t seed_post, double* w_raw, double* w_norm,
                                        int32_t* indicators, int32_t* parents, double* wsum)
{
    return filter_update_impl(f, z6, 1, 0, u_ind, seed_ind, u_post, seed_post, w_raw, w_norm, indicators, parents,
                              wsum);
}

extern "C" void orc_filter_get_state(const orc_filter* f, double* x, double* P)
{
    const int d = f->m->d;
    for (int j = 0; j < f->N; j++) {
        if (x) std::memcpy(x + (size_t)j * d, &f->pool_x[(size_t)f->buf[j] * d], sizeof(double) * d);
        if (P) std::memcpy(P + (size_t)j * d * d, &f->pool_P[(size_t)f->buf[j] * d * d], sizeof(double) * d * d);
    }
}
extern "C" void orc_filter_set_state(orc_filter* f, const double* x, const double* P)
{
    const int d = f->m->d;
    for (int j = 0; j < f->N; j++) {
        f->buf[j] = j;
        std::memcpy(&f->pool_x[(size_t)j * d], x + (size_t)j * d, sizeof(double) * d);
        std::memcpy(&f->pool_P[(size_t)j * d * d], P + (size_t)j * d * d, sizeof(double) * d * d);
    }
}

// getEstimator: estimate = estimate + 1.0/N * state  (cv::scaleAdd(state, 1/N, estimate)),
// starting from an empty Mat (quirk B8: first term is just state*(1/N)).
static void estimator(const orc_filter* f, double* xbar)
{
    const int d = f->m->d;
    const double a = 1.0 / (double)f->N;
    for (int i = 0; i < d; i++) xbar[i] = 0.0;
    for (int j = 0; j < f->N; j++) {
        const double* x = &f->pool_x[(size_t)f->buf[j] * d];
        for (int i = 0; i < d; i++) xbar[i] = x[i] * a + xbar[i];
    }
}

// e = h_pca.t()*xbar + m_pca.t()  (src/pfPose.cpp:347-348): T-expr x Mat -> gemm(h, xbar, 1, m^T, 1, GEMM_1_T)
static void reconstruct(const orc_model* m, const double* xbar, double* pose)
{
    for (int c = 0; c < m->D; c++) {
        double s = 0;
        for (int k = 0; k < m->d; k++) s += m->proj[(size_t)k * m->D + c] * xbar[k];
        pose[c] = s * 1.0 + m->pmean[c] * 1.0;
    }
}

extern "C" void orc_filter_estimate(const orc_filter* f, double* xbar, double* pose)
{
    double xb[12];
    estimator(f, xb);
    if (xbar) std::memcpy(xbar, xb, sizeof(double) * f->m->d);
    if (pose) reconstruct(f->m, xb, pose);
}

// ---------------------------------------------------------------------------------------------
// output back-end: PFTracker::rpy / get3Dpose / publishTFtree / publish2Dpos  (src/pfPose.cpp:84-208)
// ---------------------------------------------------------------------------------------------
namespace {
// cv::invert(DECOMP_LU) 3x3 closed form (determinant + adjugate), OpenCV 2.4 modules/core/src/lapack.cpp
bool invert3(const double* S, double* Dst)
{
#define SD(r, c) S[(r)*3 + (c)]
    double d = SD(0, 0) * (SD(1, 1) * SD(2, 2) - SD(1, 2) * SD(2, 1)) - SD(0, 1) * (SD(1, 0) * SD(2, 2) - SD(1, 2) * SD(2, 0)) +
               SD(0, 2) * (SD(1, 0) * SD(2, 1) - SD(1, 1) * SD(2, 0));
    if (d == 0.) {
        for (int i = 0; i < 9; i++) Dst[i] = 0;
        return false;
    }
    d = 1. / d;
    double t[9];
    t[0] = (SD(1, 1) * SD(2, 2) - SD(1, 2) * SD(2, 1)) * d;
    t[1] = (SD(0, 2) * SD(2, 1) - SD(0, 1) * SD(2, 2)) * d;
    t[2] = (SD(0, 1) * SD(1, 2) - SD(0, 2) * SD(1, 1)) * d;
    t[3] = (SD(1, 2) * SD(2, 0) - SD(1, 0) * SD(2, 2)) * d;
    t[4] = (SD(0, 0) * SD(2, 2) - SD(0, 2) * SD(2, 0)) * d;
    t[5] = (SD(0, 2) * SD(1, 0) - SD(0, 0) * SD(1, 2)) * d;
    t[6] = (SD(1, 0) * SD(2, 1) - SD(1, 1) * SD(2, 0)) * d;
    t[7] = (SD(0, 1) * SD(2, 0) - SD(0, 0) * SD(2, 1)) * d;
    t[8] = (SD(0, 0) * SD(1, 1) - SD(0, 1) * SD(1, 0)) * d;
#undef SD
    for (int i = 0; i < 9; i++) Dst[i] = t[i];
    return true;
}

// rpy(roll, pitch, yaw) = R3*R2*R1  (src/pfPose.cpp:84-91)
void rpy(double roll, double pitch, double yaw, double* R)
{
    const double R1[9] = {1, 0, 0, 0, std::cos(roll), -std::sin(roll), 0, std::sin(roll), std::cos(roll)};
    const double R2[9] = {std::cos(pitch), 0, std::sin(pitch), 0, 1, 0, -std::sin(pitch), 0, std::cos(pitch)};
    const double R3[9] = {std::cos(yaw), -std::sin(yaw), 0, std::sin(yaw), std::cos(yaw), 0, 0, 0, 1};
    double R32[9];
    gemm_nn(R3, 3, 3, R2, 3, 1.0, nullptr, 0.0, R32); // (R3*R2) evaluated first, then *R1
    gemm_nn(R32, 3, 3, R1, 3, 1.0, nullptr, 0.0, R);
}

// get3Dpose (src/pfPose.cpp:93-127): estimate is the D-vector e; Kc the 3x3 camera matrix; out pos3D 3 x 5
void get3dpose(const double* estimate, const double* Kc, double* pos3D)
{
    double est[32];
    for (int i = 0; i < 22; i++) est[i] = estimate[i];
    double R[9], T[12], P[12];
    rpy(est[16], est[17], est[15], R);
    for (int r = 0; r < 3; r++) { // hconcat(R, t, T)
        for (int c = 0; c < 3; c++) T[r * 4 + c] = R[r * 3 + c];
        T[r * 4 + 3] = est[18 + r];
    }
    gemm_nn(Kc, 3, 3, T, 4, 1.0, nullptr, 0.0, P); // P = K*T
    double im[15];
    for (int k = 0; k < 5; k++) {
        est[3 * k] = est[3 * k] * est[3 * k + 2];
        est[3 * k + 1] = est[3 * k + 1] * est[3 * k + 2];
        for (int c = 0; c < 3; c++) im[k * 3 + c] = est[3 * k + c];
    }
    double P3[9], PI[9];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) P3[r * 3 + c] = P[r * 4 + c];
    invert3(P3, PI);
    // pos3D = P_I*(im_points - repeat(P.col(3).t(),5,1)).t()  -> AddEx evaluated, transposed, gemm(P_I, .., GEMM_2_T)
    double diff[15];
    for (int k = 0; k < 5; k++)
        for (int c = 0; c < 3; c++) diff[k * 3 + c] = im[k * 3 + c] - P[c * 4 + 3];
    gemm_nt(PI, 3, 3, diff, 5, 1.0, nullptr, 0.0, pos3D);
}
} // namespace

extern "C" void orc_get3dpose(const double* estimate, const double* Kc, double* pos3D) { get3dpose(estimate, Kc, pos3D); }

// publishTFtree (src/pfPose.cpp:129-171) translations, in broadcast order: for k = 0,1 {arm1 joint k+1 -> k,
// arm2 joint k+1 -> k}, Neck->arm1 shoulder, Neck->arm2 shoulder, Head->Neck, world->Head, world->cam; plus the
// camera Euler angles (setEuler(-e1[16], -e1[17], -e1[15])).  tf: 10 x 3 (9 translations + 1 euler triple last).
// publish2Dpos (src/pfPose.cpp:173-208): 8 joints x (x, y).
extern "C" void orc_skeleton(const double* e1, const double* e2, const double* Kc, double* tf, double* joints2d)
{
    double p1[15], p2[15];
    get3dpose(e1, Kc, p1);
    get3dpose(e2, Kc, p2);
#define P1(r, c) p1[(r)*5 + (c)]
#define P2(r, c) p2[(r)*5 + (c)]
    int o = 0;
    for (int k = 0; k < 2; k++) {
        tf[o++] = P1(0, k) - P1(0, k + 1);
        tf[o++] = P1(2, k) - P1(2, k + 1);
        tf[o++] = -P1(1, k) + P1(1, k + 1);
        tf[o++] = P2(0, k) - P2(0, k + 1);
        tf[o++] = P2(2, k) - P2(2, k + 1);
        tf[o++] = -P2(1, k) + P2(1, k + 1);
    }
    const double neck_x = (P1(0, 4) + P2(0, 4)) / 2.0, neck_y = (P1(1, 4) + P2(1, 4)) / 2.0, neck_z = (P1(2, 4) + P2(2, 4)) / 2.0;
    const double head_x = (P1(0, 3) + P2(0, 3)) / 2.0, head_y = (P1(1, 3) + P2(1, 3)) / 2.0, head_z = (P1(2, 3) + P2(2, 3)) / 2.0;
    tf[o++] = P1(0, 2) - neck_x;
    tf[o++] = P1(2, 2) - neck_z;
    tf[o++] = -P1(1, 2) + neck_y;
    tf[o++] = P2(0, 2) - neck_x;
    tf[o++] = P2(2, 2) - neck_z;
    tf[o++] = -P2(1, 2) + neck_y;
    tf[o++] = neck_x - head_x;
    tf[o++] = neck_z - head_z;
    tf[o++] = -neck_y + head_y;
    tf[o++] = head_x;
    tf[o++] = head_z;
    tf[o++] = -head_y;
    tf[o++] = -e1[18];
    tf[o++] = -e1[20];
    tf[o++] = e1[19];
    tf[o++] = -e1[16];
    tf[o++] = -e1[17];
    tf[o++] = -e1[15];
#undef P1
#undef P2
    if (joints2d) {
        const double j[16] = {e1[0], e1[1], e2[0], e2[1], 0.5 * (e1[9] + e2[9]), 0.5 * (e1[10] + e2[10]),
                              0.5 * (e1[12] + e2[12]), 0.5 * (e1[13] + e2[13]), e1[3], e1[4], e2[3], e2[4],
                              e1[6], e1[7], e2[6], e2[7]};
        for (int i = 0; i < 16; i++) joints2d[i] = j[i];
    }
}

// ---------------------------------------------------------------------------------------------
// association: PFTracker::getMeasurementProposal  (src/pfPose.cpp:238-323) with
// getSampleProb (src/pf2DRao.cpp:105-122)
// ---------------------------------------------------------------------------------------------
extern "C" int orc_associate(const orc_filter* armL, const orc_filter* armR, int C, const double* cand_xy,
                             const uint8_t* cand_L, const double* roi, int img_rows, int img_cols,
                             const double* u_cand, const uint64_t* seed_cand, uint8_t* gate, double* weights,
                             int32_t* bins, double* meas)
{
    const orc_filter* arm[2] = {armL, armR};
    const int N = armL->N;
    const double scale = roi[2]; // (double)msg->ROIs[0].width
    // getSampleProb: state = H*full_state + M with H = h_pca.t(), M = m_pca.t(); cov = 0.8*scale*eye(2,2)
    double hand[2][2];
    for (int a = 0; a < 2; a++) {
        double xb[12], pose[64];
        estimator(arm[a], xb);
        reconstruct(arm[a]->m, xb, pose);
        hand[a][0] = pose[0];
        hand[a][1] = pose[1];
    }
    double cov[4] = {0.8 * scale * 1.0, 0.8 * scale * 0.0, 0.8 * scale * 0.0, 0.8 * scale * 1.0};
    // p[a][h][c]: density of hand h's candidate c under arm a
    std::vector<double> p((size_t)4 * C);
    for (int a = 0; a < 2; a++)
        for (int h = 0; h < 2; h++)
            mvnpdf_multiple2(C, cand_xy + (size_t)h * 2 * C, hand[a], cov, armL->chol_mode,
                             &p[((size_t)a * 2 + h) * C]);
    const double Pa = 0.05;
    int status = 0;
    for (int h = 0; h < 2; h++) {
        const double* px = cand_xy + (size_t)h * 2 * C;
        const double* py = px + C;
        double* w = weights + (size_t)h * C;
        double sum = 0;
        for (int j = 0; j < C; j++) {
            uint8_t g = 0;
            if ((py[j] > 0) && (py[j] < img_rows) && (px[j] > 0) && (px[j] < img_cols)) {
                double L = (double)cand_L[(size_t)h * C + j] / 255.0;
                if (L == 0) {
                    w[j] = 0;
                } else {
                    // weights1: p1_x_1*Pa + p1_x_2*Pa + 1e-4*(1-2*Pa); weights2: p2_x_2*Pa + p2_x_1*Pa + ...
                    double own = p[((size_t)h * 2 + h) * C + j], other = p[((size_t)(1 - h) * 2 + h) * C + j];
                    double Z = own * Pa + other * Pa + 1e-4 * (1 - 2 * Pa);
                    w[j] = L / Z;
                    sum = sum + w[j];
                    g = 1;
                }
            } else {
                w[j] = 0;
            }
            if (gate) gate[(size_t)h * C + j] = g;
        }
        for (int j = 0; j < C; j++) w[j] = w[j] / sum;
        int32_t* b = bins + (size_t)h * N;
        if (resample(w, C, N, u_cand[h], seed_cand ? seed_cand[h] : 1, b)) status |= (1 << h);
        if (meas) {
            double* ms = meas + (size_t)h * 6 * N;
            for (int i = 0; i < N; i++) {
                ms[0 * N + i] = roi[0] + roi[2] / 2.0;
                ms[1 * N + i] = roi[1] + 0.5 * roi[3];
                ms[2 * N + i] = px[b[i]];
                ms[3 * N + i] = py[b[i]];
                ms[4 * N + i] = roi[0] + roi[2] / 2.0;
                ms[5 * N + i] = roi[1] + 1.65 * roi[3];
            }
        }
    }
    return status;
}

// ---------------------------------------------------------------------------------------------
// legacy plain particle filter (src/pf2D.cpp; not compiled by the reference's CMakeLists.txt:29)
// ---------------------------------------------------------------------------------------------
struct orc_pf2d {
    int N, d, K;
    std::vector<double> mean, sigma_i, det_s, weight; // gmm
    std::vector<double> particles, weights;           // N x d, N
    // randomisation of the constructor / degenerate branch (src/pf2D.cpp:58-70,232-250) from the counter generator
    uint64_t seed = 0, track = 0, epoch = 0;
    int side = 0, im_w = 640, im_h = 480;
    int noise_scaled = 0; // 1: `noise` holds what predict() adds (cv::randn(.., 0, 5) values), not standard normals
};
static void pf2d_randomise(orc_pf2d* p)
{
    for (int c = 0; c < p->d; c++)      // `for i < d: cv::randu(particles.col(i), lo, hi)`
        for (int i = 0; i < p->N; i++)
            p->particles[(size_t)i * p->d + c] =
                mkf_synth_pf2d_uniform(p->seed, p->track, p->epoch, i, c, p->side, p->im_w, p->im_h);
}
extern "C" void orc_pf2d_set_random(orc_pf2d* p, uint64_t seed, uint64_t track, int side, int im_w, int im_h)
{
    p->seed = seed;
    p->track = track;
    p->side = side;
    p->im_w = im_w;
    p->im_h = im_h;
}
// ParticleFilter(numParticles, numDims, side1) (src/pf2D.cpp:44-71): uniform weights, particles across the image
extern "C" void orc_pf2d_set_noise_scaled(orc_pf2d* p, int on) { p->noise_scaled = on; }
extern "C" void orc_pf2d_randomise(orc_pf2d* p)
{
    p->epoch = 0;
    p->weights.assign(p->N, 1.0 / (double)p->N);
    pf2d_randomise(p);
}

// cv::invert(DECOMP_CHOLESKY) = Cholesky solve against the identity; cv::determinant = pivoted LU (as OpenCV 2.4 does):
// sigma_i = inv(s), det_s = 1/(pow(2 pi, d/2) * sqrt(det(s)))   (src/pf2D.cpp:28-37)
extern "C" orc_pf2d* orc_pf2d_create(int N, int d, int K, const double* means, const double* covs,
                                     const double* weights)
{
    if (d < 8 || d > 12) return nullptr;
    orc_pf2d* p = new orc_pf2d;
    p->N = N;
    p->d = d;
    p->K = K;
    p->mean.assign(means, means + (size_t)K * d);
    p->weight.assign(weights, weights + K);
    p->sigma_i.resize((size_t)K * d * d);
    p->det_s.resize(K);
    for (int k = 0; k < K; k++) {
        const double* s = covs + (size_t)k * d * d;
        // Cholesky solve against the identity (cv::invert DECOMP_CHOLESKY = Cholesky(A, b=I))
        std::vector<double> L(s, s + (size_t)d * d), inv((size_t)d * d, 0.0);
        bool ok = cv_cholesky(L.data(), d, true); // diag holds 1/L_ii
        double det = 1;
        if (ok) {
            for (int c = 0; c < d; c++) {
                std::vector<double> y(d);
                for (int i = 0; i < d; i++) { // L y = e_c
                    double t = (i == c) ? 1.0 : 0.0;
                    for (int kk = 0; kk < i; kk++) t -= L[i * d + kk] * y[kk];
                    y[i] = t * L[i * d + i];
                }
                for (int i = d - 1; i >= 0; i--) { // L^T x = y
                    double t = y[i];
                    for (int kk = d - 1; kk > i; kk--) t -= L[kk * d + i] * inv[kk * d + c];
                    inv[i * d + c] = t * L[i * d + i];
                }
            }
        }
        {   // cv::determinant(s) for n > 3 (OpenCV 2.4 lapack.cpp): LU on a copy, sign / prod(stored reciprocal pivots)
            std::vector<double> a(s, s + (size_t)d * d);
            double r = lu_impl(a.data(), d, nullptr, 0);
            if (r) {
                for (int i = 0; i < d; i++) r *= a[i * d + i];
                r = 1. / r;
            }
            det = r;
        }
        std::copy(inv.begin(), inv.end(), p->sigma_i.begin() + (size_t)k * d * d);
        p->det_s[k] = 1.0 / (std::pow(2.0 * M_PI, d / 2.0) * std::sqrt(det));
    }
    p->particles.assign((size_t)N * d, 0.0);
    p->weights.assign(N, 1.0 / (double)N);
    return p;
}
extern "C" void orc_pf2d_destroy(orc_pf2d* p) { delete p; }
extern "C" void orc_pf2d_set_particles(orc_pf2d* p, const double* particles)
{
    p->particles.assign(particles, particles + (size_t)p->N * p->d);
}
extern "C" void orc_pf2d_get_particles(const orc_pf2d* p, double* particles, double* weights)
{
    if (particles) std::copy(p->particles.begin(), p->particles.end(), particles);
    if (weights) std::copy(p->weights.begin(), p->weights.end(), weights);
}
// getEstimator (src/pf2D.cpp:79-88): `estimate = estimate + weights[i]*particles.row(i)` from an empty Mat, taken as
// zeros(1, d); the weights are whatever update() / the constructor left (resample() does not reset them)
extern "C" void orc_pf2d_estimate(const orc_pf2d* p, double* est)
{
    for (int c = 0; c < p->d; c++) est[c] = 0.0;
    for (int i = 0; i < p->N; i++)
        for (int c = 0; c < p->d; c++) est[c] = est[c] + p->weights[i] * p->particles[(size_t)i * p->d + c];
}
extern "C" void orc_pf2d_get_gmm(const orc_pf2d* p, double* sigma_i, double* det_s)
{
    if (sigma_i) std::copy(p->sigma_i.begin(), p->sigma_i.end(), sigma_i);
    if (det_s) std::copy(p->det_s.begin(), p->det_s.end(), det_s);
}

// eyemvnpdf (src/pf2D.cpp:124-128): temp = -0.5*x_u*1.0/scale*eye(2,2)*x_u.t();
// 1.0/pow(2 pi scale, cols/2.0) * exp(temp)
static double eyemvnpdf(double dx, double dy, double scale)
{
    // MatExpr: ((x_u * -0.5) * (1.0/scale... evaluated as scaled gemm chain: alpha = -0.5*1.0/scale
    double alpha = -0.5 * 1.0 / scale;
    // (alpha * x_u) * eye  -> 1 x 2 ; then * x_u.t()
    double t0 = (dx * 1.0 + dy * 0.0) * alpha, t1 = (dx * 0.0 + dy * 1.0) * alpha;
    double q = t0 * dx + t1 * dy;
    return 1.0 / (std::pow(2.0 * M_PI * scale, 2 / 2.0)) * std::exp(q);
}

extern "C" int orc_pf2d_update(orc_pf2d* p, const double* meas, double u, const double* noise, double* w_norm,
                               int32_t* parents)
{
    const int N = p->N, d = p->d, K = p->K;
    double weightSum = 0;
    std::vector<double> xu(d), t(d);
    for (int i = 0; i < N; i++) {
        const double* x = &p->particles[(size_t)i * d];
        double prior = 0;
        for (int j = 0; j < K; j++) {
            // gmmmvnpdf: temp = -0.5*x_u*sigma_i*x_u.t(); return expf(float(temp))  (src/pf2D.cpp:105-109)
            const double* Si = &p->sigma_i[(size_t)j * d * d];
            for (int c = 0; c < d; c++) xu[c] = x[c] - p->mean[(size_t)j * d + c];
            for (int c = 0; c < d; c++) {
                double s = 0;
                for (int k = 0; k < d; k++) s += xu[k] * Si[k * d + c];
                t[c] = s * -0.5;
            }
            double q;
            gemm_nt(t.data(), 1, d, xu.data(), 1, 1.0, nullptr, 0.0, &q); // (..)*x_u.t()
            double e = (double)mkf_expf((float)q); // glibc's expf, restated in include/mkf_expf.h (shared with the device)
            prior = prior + p->weight[j] * p->det_s[j] * e;
        }
        // likelihood = eyemvnpdf(p[6:8] - meas.row(0), 15) * eyemvnpdf(p[0:2] - meas.row(1), 15)
        double lik = eyemvnpdf(x[6] - meas[0], x[7] - meas[1], 15) * eyemvnpdf(x[0] - meas[2], x[1] - meas[3], 15);
        p->weights[i] = prior * lik;
        weightSum += p->weights[i];
    }
    for (int i = 0; i < N; i++) p->weights[i] = p->weights[i] / weightSum;
    if (w_norm) std::copy(p->weights.begin(), p->weights.end(), w_norm);
    // resample() (src/pf2D.cpp:225-268)
    double mw = 0;
    for (int i = 0; i < N; i++)
        if (p->weights[i] > mw) mw = p->weights[i];
    int status = 0;
    std::vector<double> old(p->particles);
    p->epoch++; // epoch n = the n-th update
    if (mw == 0) {
        // src/pf2D.cpp:232-250: every particle re-randomised over the image (cv::randu there, the counter generator
        // here: cv::theRNG() cannot be seeded through the class), weights back to 1/N
        status = 1;
        pf2d_randomise(p);
        for (int i = 0; i < N; i++) {
            p->weights[i] = 1.0 / (double)N;
            if (parents) parents[i] = i;
        }
    } else {
        int idx = 0;
        double step = 1.0 / (double)N;
        double beta = u * step;
        for (int i = 0; i < N; i++) {
            while (beta > p->weights[idx]) {
                beta -= p->weights[idx];
                idx = (idx + 1) % N;
            }
            beta += step;
            if (parents) parents[i] = idx;
            // particles.row(i) = zeros + old_particles.row(idx)
            for (int c = 0; c < d; c++) p->particles[(size_t)i * d + c] = 0.0 + old[(size_t)idx * d + c];
        }
    }
    // predict(): particles.row(i) += randn(0,5) per dimension (src/pf2D.cpp:90-102); dims >= 8 of
    // the temp row are never written by randn in the reference (uninitialised); taken as 0 here.
    if (noise) {
        for (int i = 0; i < N; i++)
            for (int c = 0; c < d && c < 8; c++)
                p->particles[(size_t)i * d + c] =
                    p->particles[(size_t)i * d + c] +
                    (p->noise_scaled ? noise[(size_t)i * d + c] : noise[(size_t)i * d + c] * 5.0);
    }
    return status;
}

// ---------------------------------------------------------------------------------------------
// CPU baseline: one track per task, single-threaded within a track (as the reference is),
// OpenMP over tracks.
// ---------------------------------------------------------------------------------------------
extern "C" double orc_bench_tracks(const orc_model* m, int64_t T, int N, int frames, int per_slot, uint64_t seed,
                                   int jitter, int chol_mode, int alias_mode, int threads, double* pose_out,
                                   int* threads_used)
{
    int nthreads = 1;
#ifdef _OPENMP
    nthreads = threads > 0 ? threads : omp_get_max_threads();
#else
    (void)threads;
#endif
    if (threads_used) *threads_used = nthreads;
    std::vector<orc_filter*> filt((size_t)T);
    for (int64_t t = 0; t < T; t++) {
        filt[t] = orc_filter_create(m, N, chol_mode, alias_mode);
        orc_filter_reset(filt[t], mkf_synth_u(seed, (uint64_t)t, MKF_SYNTH_NO_FRAME, MKF_SYNTH_LANE_U_INIT), 1);
    }
    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (int64_t t = 0; t < T; t++) {
        std::vector<double> meas(per_slot ? (size_t)6 * N : 6);
        for (int fr = 0; fr < frames; fr++) {
            double ui = mkf_synth_u(seed, (uint64_t)t, (uint64_t)fr, MKF_SYNTH_LANE_U_IND);
            double up = mkf_synth_u(seed, (uint64_t)t, (uint64_t)fr, MKF_SYNTH_LANE_U_POST);
            if (per_slot) {
                for (int j = 0; j < N; j++) {
                    double z[6];
                    mkf_synth_meas(seed, (uint64_t)t, (uint64_t)fr, j, jitter, z);
                    for (int r = 0; r < 6; r++) meas[(size_t)r * N + j] = z[r];
                }
                orc_filter_update(filt[t], meas.data(), ui, 1, up, 1, nullptr, nullptr, nullptr, nullptr, nullptr);
            } else {
                mkf_synth_meas(seed, (uint64_t)t, (uint64_t)fr, -1, jitter, meas.data());
                orc_filter_update_shared(filt[t], meas.data(), ui, 1, up, 1, nullptr, nullptr, nullptr, nullptr,
                                         nullptr);
            }
        }
        if (pose_out) orc_filter_estimate(filt[t], nullptr, pose_out + (size_t)t * m->D);
    }
    auto t1 = std::chrono::steady_clock::now();
    for (int64_t t = 0; t < T; t++) orc_filter_destroy(filt[t]);
    return std::chrono::duration<double>(t1 - t0).count();
}

// candidate generation front-end for one person (src/pfPose.cpp:216-236 with getSamples, src/pf2DRao.cpp:85-103),
// driven by the shared counter generator.  like: rows x cols uint8 or NULL.
extern "C" void orc_propose(const orc_filter* armL, const orc_filter* armR, int C, const double* roi, int tracking,
                            const uint8_t* like, int rows, int cols, uint64_t seed, uint64_t track, uint64_t frame,
                            double* cand_xy, uint8_t* cand_L)
{
    const orc_filter* arm[2] = {armL, armR};
    for (int h = 0; h < 2; h++) {
        double xb[12], pose[64];
        estimator(arm[h], xb);       // getSamples: full_state = getEstimator(); state = H*full_state + M
        reconstruct(arm[h]->m, xb, pose);
        for (int c = 0; c < C; c++) {
            double x, y;
            mkf_synth_proposal(seed, track, frame, h, C, c, tracking, pose[0], pose[1], roi, rows, cols, 0.8, &x, &y);
            cand_xy[((size_t)h * 2 + 0) * C + c] = x;
            cand_xy[((size_t)h * 2 + 1) * C + c] = y;
            if (cand_L) cand_L[(size_t)h * C + c] = mkf_likelihood_lookup(like, rows, cols, x, y);
        }
    }
}

extern "C" float orc_expf(float x) { return mkf_expf(x); }
extern "C" float orc_libm_expf(float x) { return expf(x); }
extern "C" uint64_t orc_expf_compare(const uint32_t* bits, uint64_t n)
{
    uint64_t bad = 0;
#pragma omp parallel for reduction(+ : bad)
    for (uint64_t i = 0; i < n; i++) {
        float x, a, b;
        std::memcpy(&x, &bits[i], 4);
        a = expf(x);
        b = mkf_expf(x);
        uint32_t ua, ub;
        std::memcpy(&ua, &a, 4);
        std::memcpy(&ub, &b, 4);
        if (ua != ub && !(a != a && b != b)) bad++;
    }
    return bad;
}

// expose the shared synthetic generator so numpy-free tests can pin it
extern "C" void orc_synth_meas(uint64_t seed, uint64_t track, uint64_t frame, int64_t slot, int jitter, double* z6)
{
    mkf_synth_meas(seed, track, frame, slot, jitter, z6);
}
extern "C" double orc_synth_u(uint64_t seed, uint64_t track, uint64_t frame, uint32_t which)
{
    return mkf_synth_u(seed, track, frame, which);
}
extern "C" void orc_synth_candidate(uint64_t seed, uint64_t track, uint64_t frame, int hand, int C, int c, int jitter,
                                    double* cx, double* cy, uint8_t* L)
{
    mkf_synth_candidate(seed, track, frame, hand, C, c, jitter, cx, cy, L);
}
