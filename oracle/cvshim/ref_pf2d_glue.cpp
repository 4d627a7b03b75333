// ref_pf2d_glue.cpp -- C entry points around the reference's LEGACY plain particle filter, compiled from the
// reference's own src/pf2D.cpp (read in place, never copied) against the OpenCV-subset shim -> oracle/_ref/libref_pf2d.so.
// A separate library: src/pf2D.h re-defines my_gmm / ParticleFilter, which clash with src/my_gmm.h / src/pf2DRao.h
// (the reference does not compile this file either, CMakeLists.txt:29).  Test infrastructure only.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "pf2D.h"

namespace {
struct RefPf2d : public ParticleFilter { // the filter's state is protected: a subclass reads it
    RefPf2d(int n, int dims, bool side1) : ParticleFilter(n, dims, side1) {}
    int n() const { return N; }
    int dims() const { return d; }
    cv::Mat& parts() { return particles; }
    std::vector<double>& w() { return weights; }
};
} // namespace

extern "C" {
// ParticleFilter(numParticles, numDims, side1) (src/pf2D.cpp:44-71); rng_seed seeds the shim's global generator
// behind cv::randu / cv::randn
void* refpf_create(int N, int d, int side, unsigned long long rng_seed)
{
    cv::cvshim_seed_the_rng(rng_seed);
    cv::cvshim_random_log().clear();
    return new RefPf2d(N, d, side != 0);
}
void refpf_destroy(void* h) { delete (RefPf2d*)h; }
// my_gmm::loadGaussian(mean 1 x d, sigma d x d, weight) (src/pf2D.cpp:28-37)
void refpf_load_gaussian(void* h, const double* mean, const double* cov, double w)
{
    RefPf2d* p = (RefPf2d*)h;
    const int d = p->dims();
    cv::Mat u(1, d, CV_64F), s(d, d, CV_64F);
    for (int c = 0; c < d; c++) u.at<double>(0, c) = mean[c];
    for (int r = 0; r < d; r++)
        for (int c = 0; c < d; c++) s.at<double>(r, c) = cov[r * d + c];
    p->gmm.loadGaussian(u, s, w);
}
void refpf_get_gmm(void* h, double* sigma_i, double* det_s)
{
    RefPf2d* p = (RefPf2d*)h;
    const int d = p->dims();
    for (int k = 0; k < p->gmm.N; k++) {
        for (int r = 0; r < d; r++)
            for (int c = 0; c < d; c++) sigma_i[((size_t)k * d + r) * d + c] = p->gmm.sigma_i[k].at<double>(r, c);
        det_s[k] = p->gmm.det_s[k];
    }
}
void refpf_set_particles(void* h, const double* x)
{
    RefPf2d* p = (RefPf2d*)h;
    for (int i = 0; i < p->n(); i++)
        for (int c = 0; c < p->dims(); c++) p->parts().at<double>(i, c) = x[(size_t)i * p->dims() + c];
}
void refpf_get(void* h, double* x, double* w)
{
    RefPf2d* p = (RefPf2d*)h;
    if (x)
        for (int i = 0; i < p->n(); i++)
            for (int c = 0; c < p->dims(); c++) x[(size_t)i * p->dims() + c] = p->parts().at<double>(i, c);
    if (w)
        for (int i = 0; i < p->n(); i++) w[i] = p->w()[i];
}
// ParticleFilter::update(measurement 2 x 2) (src/pf2D.cpp:148-210) with the C library generator seeded so that the
// uniform of resample() (`rand() % N` drawn and unused, then `rand() / RAND_MAX`, :228,:255) is known: returned.
// noise_out (N x d, may be null): what predict() added to every particle (cv::randn draws, src/pf2D.cpp:90-102; zero
// for dimensions >= 8).  *degenerate = 1 when resample() took the `mw == 0` branch.
double refpf_update(void* h, const double* meas, unsigned srand_seed, double* noise_out, int* degenerate)
{
    RefPf2d* p = (RefPf2d*)h;
    const int N = p->n(), d = p->dims();
    srand(srand_seed);
    (void)rand();
    const double u = (double)rand() / RAND_MAX;
    srand(srand_seed);
    cv::Mat z(2, 2, CV_64F);
    for (int i = 0; i < 4; i++) z.at<double>(i / 2, i % 2) = meas[i];
    cv::cvshim_random_log().clear();
    p->update(z);
    const std::vector<std::vector<double> >& log = cv::cvshim_random_log();
    const size_t extra = log.size() - (size_t)3 * N; // d randu calls in front when the degenerate branch ran
    if (degenerate) *degenerate = extra > 0 ? 1 : 0;
    if (noise_out) {
        std::memset(noise_out, 0, sizeof(double) * (size_t)N * d);
        for (int i = 0; i < N; i++) {
            const std::vector<double>&a = log[extra + 3 * i], &b = log[extra + 3 * i + 1], &c = log[extra + 3 * i + 2];
            for (int q = 0; q < 4; q++) noise_out[(size_t)i * d + 2 + q] = a[q]; // temp.colRange(2, 6)
            for (int q = 0; q < 2; q++) noise_out[(size_t)i * d + 6 + q] = b[q]; // temp.colRange(6, 8)
            for (int q = 0; q < 2; q++) noise_out[(size_t)i * d + q] = c[q];     // temp.colRange(0, 2)
        }
    }
    return u;
}
// ParticleFilter::getEstimator (src/pf2D.cpp:79-88)
void refpf_estimate(void* h, double* est)
{
    RefPf2d* p = (RefPf2d*)h;
    cv::Mat e = p->getEstimator();
    for (int c = 0; c < p->dims(); c++) est[c] = e.at<double>(0, c);
}
}
