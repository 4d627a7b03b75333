// oracle/cvshim/ref_glue.cpp -- TEST INFRASTRUCTURE: C entry points around the reference's own
// classes (compiled from /root/reference/src in place, see oracle/Makefile) for the Python tests and
// for bench.py's CPU baseline (kind "reference").
#define protected public
#define private public
#include "pf2DRao.h"
#undef protected
#undef private

#include <chrono>
#include <cstring>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../../include/mkf_synth.h"

static cv::Mat wrap(const double* p, int rows, int cols)
{
    cv::Mat m(rows, cols, CV_64F);
    for (int r = 0; r < rows; r++)
        for (int c = 0; c < cols; c++) m.at<double>(r, c) = p[(size_t)r * cols + c];
    return m;
}

extern "C" {

void* ref_pf_create(int nParticles) { return new ParticleFilter(nParticles); }
void ref_pf_destroy(void* pf) { delete (ParticleFilter*)pf; }

// the constructor loop of src/pfPose.cpp:61-65 for one arm
void ref_pf_load_model(void* pfv, int K, int d, int D, const double* means, const double* covs, const double* weights,
                       const double* gamma, const double* pca_proj, const double* pca_mean)
{
    ParticleFilter* pf = (ParticleFilter*)pfv;
    cv::Mat means1 = wrap(means, K, d), covs1 = wrap(covs, K * d, d), weights1 = wrap(weights, 1, K),
            g1 = wrap(gamma, K, 1);
    cv::Mat h1_pca = wrap(pca_proj, d, D), m1_pca = wrap(pca_mean, 1, D);
    for (int i = 0; i < means1.rows; i++)
        pf->gmm.loadGaussian(means1.row(i), covs1(cv::Range(covs1.cols * i, covs1.cols * (i + 1)), cv::Range(0, covs1.cols)),
                             h1_pca, m1_pca, weights1.at<double>(0, i), g1.at<double>(0, i));
}

// derived KF_model members of component k: Q d x d, B d, H 6 x d, BH 6, R 6 x 6, F d x d
void ref_pf_get_kf(void* pfv, int k, double* Q, double* B, double* H, double* BH, double* R, double* F)
{
    ParticleFilter* pf = (ParticleFilter*)pfv;
    KF_model& t = pf->gmm.KFtracker[k];
    auto out = [](const cv::Mat& m, double* p) {
        if (!p) return;
        for (int r = 0; r < m.rows; r++)
            for (int c = 0; c < m.cols; c++) p[(size_t)r * m.cols + c] = m.at<double>(r, c);
    };
    out(t.Q, Q);
    out(t.B, B);
    out(t.H, H);
    out(t.BH, BH);
    out(t.R, R);
    out(t.F, F);
}

// bins = pf->resample(gmm.weight, N); gmm.resetTracker(bins)  (src/pfPose.cpp:68-71); `tick` is what
// cv::getTickCount() returns inside resample
void ref_pf_reset(void* pfv, int64_t tick)
{
    ParticleFilter* pf = (ParticleFilter*)pfv;
    cv::cvshim_push_tick(tick);
    std::vector<int> bins = pf->resample(pf->gmm.weight, pf->gmm.nParticles);
    pf->gmm.resetTracker(bins);
}

// ParticleFilter::update with the two cv::getTickCount() values its resample() calls will see
void ref_pf_update(void* pfv, const double* meas /* 6 x N */, int64_t tick_ind, int64_t tick_post)
{
    ParticleFilter* pf = (ParticleFilter*)pfv;
    cv::cvshim_push_tick(tick_ind);
    cv::cvshim_push_tick(tick_post);
    pf->update(wrap(meas, 6, pf->gmm.nParticles));
}

void ref_pf_get_state(void* pfv, double* x, double* P)
{
    ParticleFilter* pf = (ParticleFilter*)pfv;
    const int N = pf->gmm.nParticles;
    for (int j = 0; j < N; j++) {
        const cv::Mat& s = pf->gmm.tracks[j].state;
        const cv::Mat& c = pf->gmm.tracks[j].cov;
        const int d = s.rows;
        if (x)
            for (int i = 0; i < d; i++) x[(size_t)j * d + i] = s.at<double>(i, 0);
        if (P)
            for (int r = 0; r < d; r++)
                for (int q = 0; q < d; q++) P[((size_t)j * d + r) * d + q] = c.at<double>(r, q);
    }
}

void ref_pf_estimate(void* pfv, double* xbar)
{
    ParticleFilter* pf = (ParticleFilter*)pfv;
    cv::Mat e = pf->getEstimator();
    for (int i = 0; i < e.rows; i++) xbar[i] = e.at<double>(i, 0);
}

int ref_pf_resample(void* pfv, const double* w, int L, int N, int64_t tick, int32_t* out)
{
    ParticleFilter* pf = (ParticleFilter*)pfv;
    cv::cvshim_push_tick(tick);
    std::vector<int> r = pf->resample(std::vector<double>(w, w + L), N);
    for (int i = 0; i < N; i++) out[i] = r[i];
    return 0;
}

void ref_pf_chol(void* pfv, int n, const double* in, double* out)
{
    ParticleFilter* pf = (ParticleFilter*)pfv;
    cv::Mat r = pf->chol(wrap(in, n, n));
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) out[i * n + j] = r.at<double>(i, j);
}

double ref_pf_mvnpdf(void* pfv, int n, const double* x, const double* u, const double* sigma)
{
    ParticleFilter* pf = (ParticleFilter*)pfv;
    return pf->mvnpdf(wrap(x, n, 1), wrap(u, n, 1), wrap(sigma, n, n));
}

void ref_kf_predict(void* pfv, int k, double* x, double* P)
{
    ParticleFilter* pf = (ParticleFilter*)pfv;
    const int d = pf->gmm.mean[0].cols;
    cv::Mat s = wrap(x, d, 1), c = wrap(P, d, d);
    pf->gmm.KFtracker[k].predict(s, c);
    for (int i = 0; i < d; i++) x[i] = s.at<double>(i, 0);
    for (int r = 0; r < d; r++)
        for (int q = 0; q < d; q++) P[r * d + q] = c.at<double>(r, q);
}

void ref_kf_update(void* pfv, int k, const double* z, double* x, double* P)
{
    ParticleFilter* pf = (ParticleFilter*)pfv;
    const int d = pf->gmm.mean[0].cols;
    cv::Mat s = wrap(x, d, 1), c = wrap(P, d, d);
    pf->gmm.KFtracker[k].update(wrap(z, 6, 1), s, c);
    for (int i = 0; i < d; i++) x[i] = s.at<double>(i, 0);
    for (int r = 0; r < d; r++)
        for (int q = 0; q < d; q++) P[r * d + q] = c.at<double>(r, q);
}

// getSampleProb (src/pf2DRao.cpp:105-122): H = h_pca.t() (D x d), M = m_pca.t() (D x 1)
void ref_pf_sample_prob(void* pfv, int d, int D, const double* pca_proj, const double* pca_mean, const double* in1, int C1,
                        const double* in2, int C2, double scale, double* w1, double* w2)
{
    ParticleFilter* pf = (ParticleFilter*)pfv;
    cv::Mat h = wrap(pca_proj, d, D), m = wrap(pca_mean, 1, D);
    std::vector<double> a, b;
    pf->getSampleProb(h.t(), m.t(), wrap(in1, 2, C1), wrap(in2, 2, C2), a, b, scale);
    std::memcpy(w1, a.data(), sizeof(double) * C1);
    std::memcpy(w2, b.data(), sizeof(double) * C2);
}

void ref_pf_get_samples_mean(void* pfv, int d, int D, const double* pca_proj, const double* pca_mean, int N, double scale,
                             double* mean_xy, double* sd_xy)
{
    ParticleFilter* pf = (ParticleFilter*)pfv;
    cv::Mat h = wrap(pca_proj, d, D), m = wrap(pca_mean, 1, D);
    cv::Mat s = pf->getSamples(h.t(), m.t(), N, scale);
    for (int r = 0; r < 2; r++) {
        double mu = 0, v = 0;
        for (int i = 0; i < N; i++) mu += s.at<double>(r, i) / N;
        for (int i = 0; i < N; i++) v += (s.at<double>(r, i) - mu) * (s.at<double>(r, i) - mu) / N;
        mean_xy[r] = mu;
        sd_xy[r] = std::sqrt(v);
    }
}

// CPU baseline on the reference's own code: T independent ParticleFilter objects, one track per
// OpenMP task, the synthetic workload of include/mkf_synth.h.  Returns wall seconds of the frame loop.
double ref_bench_tracks(int K, int d, int D, const double* means, const double* covs, const double* weights,
                        const double* gamma, const double* pca_proj, const double* pca_mean, int64_t T, int N, int frames,
                        int per_slot, uint64_t seed, int jitter, int threads, int* threads_used)
{
    int nthreads = 1;
#ifdef _OPENMP
    nthreads = threads > 0 ? threads : omp_get_max_threads();
#endif
    if (threads_used) *threads_used = nthreads;
    std::vector<ParticleFilter*> pfs((size_t)T);
    for (int64_t t = 0; t < T; t++) {
        pfs[t] = new ParticleFilter(N);
        ref_pf_load_model(pfs[t], K, d, D, means, covs, weights, gamma, pca_proj, pca_mean);
        ref_pf_reset(pfs[t], (int64_t)(mkf_hash4(seed, (uint64_t)t, MKF_SYNTH_NO_FRAME, MKF_SYNTH_LANE_U_INIT) | 1));
    }
    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (int64_t t = 0; t < T; t++) {
        std::vector<double> meas((size_t)6 * N);
        for (int fr = 0; fr < frames; fr++) {
            for (int j = 0; j < N; j++) {
                double z[6];
                mkf_synth_meas(seed, (uint64_t)t, (uint64_t)fr, per_slot ? j : -1, jitter, z);
                for (int r = 0; r < 6; r++) meas[(size_t)r * N + j] = z[r];
            }
            ref_pf_update(pfs[t], meas.data(), (int64_t)(mkf_hash4(seed, (uint64_t)t, (uint64_t)fr, MKF_SYNTH_LANE_U_IND) | 1),
                          (int64_t)(mkf_hash4(seed, (uint64_t)t, (uint64_t)fr, MKF_SYNTH_LANE_U_POST) | 1));
        }
    }
    auto t1 = std::chrono::steady_clock::now();
    for (int64_t t = 0; t < T; t++) delete pfs[t];
    return std::chrono::duration<double>(t1 - t0).count();
}

} // extern "C"
