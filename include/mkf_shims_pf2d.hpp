// mkf_shims_pf2d.hpp -- the reference's LEGACY plain particle filter classes as drop-in host code over libmkf_b200
//
//   class my_gmm          src/pf2D.h:12-23   (loadGaussian; mean, sigma_i, det_s, weight, N)
//   class ParticleFilter  src/pf2D.h:25-51   (ParticleFilter(numParticles, numDims, side1), predict, update,
//                                             getEstimator, gmm)
//
// src/pf2D.{h,cpp} is not compiled by the reference (CMakeLists.txt:29 lists pf2DRao.cpp) and its class names clash
// with the ones of src/pf2DRao.h / src/my_gmm.h.  Here both sets can be used side by side: the legacy classes live in
// namespace mkf_legacy (write mkf_legacy::ParticleFilter, mkf_legacy::my_gmm).  The arithmetic (weights, normalise,
// systematic resample, random-walk predict, estimator) runs on the GPU through mkf_pf2d_*; there is no CPU fallback.
//
// Random draws.  resample() uses the C library generator exactly as the reference does (`rand() % N` drawn and
// discarded, then `rand()/RAND_MAX`, src/pf2D.cpp:228,255), so srand() reproduces the reference's uniform.  cv::randu
// (constructor) and cv::randn (predict) draw from OpenCV's global generator, which cannot be reproduced without
// OpenCV: the shim draws from its own counter generator (seedable: mkf_legacy::the_stream()), same distributions.
// The re-randomisation inside resample() when every weight is 0 (src/pf2D.cpp:232-250) happens on the device, from
// the counter generator of mkf_synth.h (mkf_pf2d_set_random).
#ifndef MKF_SHIMS_PF2D_HPP
#define MKF_SHIMS_PF2D_HPP

#include <cmath>
#include <cstdlib>

#include "mkf_shims.hpp" // cv::Mat (OpenCV or the built-in subset), mkf::check / flat / unflat

namespace mkf_legacy {

inline uint64_t& the_stream()
{
    static uint64_t s = 0x243F6A8885A308D3ull;
    return s;
}
inline uint64_t next_u64()
{
    uint64_t z = (the_stream() += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
inline double uniform01() { return (double)(next_u64() >> 11) * (1.0 / 9007199254740992.0); }
inline double normal01()
{
    const double u1 = 1.0 - uniform01(), u2 = uniform01(); // u1 in (0, 1]
    return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586476925286766559 * u2);
}

// GMM storage class (src/pf2D.h:12-23, src/pf2D.cpp:11-37)
class my_gmm {
  public:
    my_gmm() { N = 0; }
    ~my_gmm() {}
    // src/pf2D.cpp:28-37: sigma_i = invert(s, DECOMP_CHOLESKY), det_s = 1 / (pow(2 pi, cols/2) sqrt(det s))
    void loadGaussian(cv::Mat u, cv::Mat s, double w)
    {
        const int d = u.cols;
        if (u.rows != 1 || s.rows != d || s.cols != d)
            throw mkf::Error(MKF_E_INVALID, "my_gmm::loadGaussian: mean must be 1 x d and sigma d x d");
        N++;
        mean.push_back(u);
        // lower Cholesky factor, then the inverse column by column; determinant from the diagonal
        std::vector<double> L = mkf::flat(s), inv((size_t)d * d, 0.0);
        double det = 1.0;
        for (int i = 0; i < d; i++) {
            for (int j = 0; j <= i; j++) {
                double t = L[(size_t)i * d + j];
                for (int k = 0; k < j; k++) t -= L[(size_t)i * d + k] * L[(size_t)j * d + k];
                if (i == j) {
                    if (!(t > 0.0)) throw mkf::Error(MKF_E_INVALID, "my_gmm::loadGaussian: sigma is not positive definite");
                    L[(size_t)i * d + i] = std::sqrt(t);
                    det *= t;
                } else {
                    L[(size_t)i * d + j] = t / L[(size_t)j * d + j];
                }
            }
        }
        std::vector<double> y(d);
        for (int c = 0; c < d; c++) {
            for (int i = 0; i < d; i++) {
                double t = (i == c) ? 1.0 : 0.0;
                for (int k = 0; k < i; k++) t -= L[(size_t)i * d + k] * y[k];
                y[i] = t / L[(size_t)i * d + i];
            }
            for (int i = d - 1; i >= 0; i--) {
                double t = y[i];
                for (int k = d - 1; k > i; k--) t -= L[(size_t)k * d + i] * inv[(size_t)k * d + c];
                inv[(size_t)i * d + c] = t / L[(size_t)i * d + i];
            }
        }
        sigma_i.push_back(mkf::unflat(inv.data(), d, d));
        det_s.push_back(1.0 / (std::pow(2.0 * M_PI, d / 2.0) * std::sqrt(det)));
        weight.push_back(w);
        sigma_.push_back(mkf::flat(s)); // what mkf_pf2d_create takes (it derives sigma_i / det_s itself)
    }
    std::vector<cv::Mat> mean;
    std::vector<cv::Mat> sigma_i;
    std::vector<double> det_s;
    std::vector<double> weight;
    int N;

    // ---- additions ----
    const std::vector<std::vector<double>>& sigma_raw() const { return sigma_; }

  private:
    std::vector<std::vector<double>> sigma_;
};

class ParticleFilter {
  public:
    ParticleFilter() {}
    // src/pf2D.cpp:44-71: uniform weights, particles randomised across the 640 x 480 image (column 6 -- the head x
    // -- over the half of the image `side1` selects)
    ParticleFilter(int numParticles, int numDims, bool side1)
    {
        if (numParticles < 1 || numDims < 8 || numDims > 12 || (numDims & 1))
            throw mkf::Error(MKF_E_UNSUPPORTED, "legacy ParticleFilter: numDims must be 8, 10 or 12 and numParticles >= 1");
        N = numParticles;
        d = numDims;
        side = side1;
        im_height = 480;
        im_width = 640;
        host_.assign((size_t)N * d, 0.0);
        randomise();
        dirty_ = true;
    }
    ~ParticleFilter()
    {
        if (h_) mkf_pf2d_destroy(h_);
    }
    ParticleFilter(const ParticleFilter&) = delete;
    ParticleFilter& operator=(const ParticleFilter&) = delete;

    // Random walk motion model (src/pf2D.cpp:90-102): N(0, 5) on the first eight dimensions.  update() already
    // ends with it (on the device); a stand-alone call goes through the host copy of the particles.
    void predict()
    {
        pull();
        for (int i = 0; i < N; i++)
            for (int c = 0; c < 8; c++) host_[(size_t)i * d + c] += 5.0 * normal01();
        dirty_ = true;
    }
    // src/pf2D.cpp:148-210: weights (GMM prior x two isotropic likelihoods), normalise, resample, predict.
    // measurement is 2 x 2: row 0 against particle columns 6:8, row 1 against columns 0:2.
    void update(cv::Mat measurement)
    {
        if (measurement.rows != 2 || measurement.cols != 2)
            throw mkf::Error(MKF_E_INVALID, "legacy ParticleFilter::update: measurement must be 2 x 2");
        ensure();
        push();
        const std::vector<double> z = mkf::flat(measurement);
        (void)(rand() % N);                            // `int idx = rand() % N;` drawn and unused (src/pf2D.cpp:228)
        last_u = (double)rand() / RAND_MAX;            // src/pf2D.cpp:255
        last_noise.resize((size_t)N * d);
        for (size_t i = 0; i < last_noise.size(); i++) last_noise[i] = normal01();
        mkf::check(mkf_pf2d_update(h_, z.data(), &last_u, last_noise.data(), MKF_MEM_HOST));
    }
    // Weighted average pose estimate (src/pf2D.cpp:79-88), 1 x d
    cv::Mat getEstimator()
    {
        ensure();
        push();
        std::vector<double> e(d);
        mkf::check(mkf_pf2d_estimate(h_, e.data(), MKF_MEM_HOST));
        return mkf::unflat(e.data(), 1, d);
    }
    my_gmm gmm;

    // ---- additions (tests): the particle matrix, the draws of the last update ----
    cv::Mat getParticles()
    {
        pull();
        return mkf::unflat(host_.data(), N, d);
    }
    void setParticles(const cv::Mat& p)
    {
        if (p.rows != N || p.cols != d) throw mkf::Error(MKF_E_INVALID, "setParticles: N x d expected");
        host_ = mkf::flat(p);
        dirty_ = true;
    }
    double last_u = 0;
    std::vector<double> last_noise;
    uint64_t random_seed = 0; // seed of the degenerate branch's draws (mkf_pf2d_set_random)

  protected:
    void randomise()
    {
        for (int c = 0; c < d; c++) {
            double lo, hi; // cv::randu(particles.col(i), lo, hi): uniform on [lo, hi)
            if (c == 6) {
                lo = im_width / 2.0 * side + 1;
                hi = im_width / 2.0 + im_width / 2.0 * side;
            } else {
                lo = 1;
                hi = ((c % 2) == 0) * im_width + (((c + 1) % 2) == 0) * im_height;
            }
            for (int i = 0; i < N; i++) host_[(size_t)i * d + c] = lo + (hi - lo) * uniform01();
        }
    }
    void ensure() // the device filter is created at first use, once the GMM has been loaded (as PFTracker would)
    {
        if (h_) return;
        if (gmm.N < 1) throw mkf::Error(MKF_E_INVALID, "legacy ParticleFilter: gmm.loadGaussian has not been called");
        std::vector<double> means, covs;
        for (int k = 0; k < gmm.N; k++) {
            if (gmm.mean[k].cols != d) throw mkf::Error(MKF_E_INVALID, "legacy ParticleFilter: GMM dimension != numDims");
            const std::vector<double> m = mkf::flat(gmm.mean[k]);
            means.insert(means.end(), m.begin(), m.end());
            covs.insert(covs.end(), gmm.sigma_raw()[k].begin(), gmm.sigma_raw()[k].end());
        }
        mkf::check(mkf_pf2d_create(&h_, 1, N, d, gmm.N, means.data(), covs.data(), gmm.weight.data(), 0, nullptr));
        // arm resample()'s degenerate branch (max weight 0: particles re-drawn across the image, src/pf2D.cpp:232-250)
        // with this filter's side / image size and a seed taken from the shim's stream
        const uint8_t sd = side ? 1 : 0;
        random_seed = next_u64();
        mkf::check(mkf_pf2d_set_random(h_, random_seed, 0, &sd, im_width, im_height));
        dirty_ = true;
    }
    void push()
    {
        if (!dirty_) return;
        mkf::check(mkf_pf2d_set_particles(h_, host_.data(), MKF_MEM_HOST));
        dirty_ = false;
    }
    void pull()
    {
        if (!h_ || dirty_) return; // the host copy is the current one
        mkf::check(mkf_pf2d_get(h_, host_.data(), nullptr, nullptr, MKF_MEM_HOST));
    }
    int N = 0;
    int d = 0;
    bool side = false;
    int im_width = 640, im_height = 480;

  private:
    mkf_pf2d* h_ = nullptr;
    std::vector<double> host_;
    bool dirty_ = false;
};

} // namespace mkf_legacy

#endif
