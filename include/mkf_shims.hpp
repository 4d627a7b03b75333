// mkf_shims.hpp -- the reference's C++ class interfaces for the hot path, as drop-in host code
// over the C ABI of mkf_b200.h:
//
//   class KF_model      src/KF_model.h:8-16      (Q, R, F, B, H, BH; predict; update)
//   class state_params  src/my_gmm.h:9-18        (state, cov, weight; deep-copy constructor)
//   class my_gmm        src/my_gmm.h:20-33       (loadGaussian, resetTracker, mean, cov, weight, KFtracker, tracks, nParticles)
//   class ParticleFilter src/pf2DRao.h:13-31     (update, getEstimator, getSamples, getSampleProb, gmm, resample)
//
// Same names, members and argument meaning as the reference, so that pfPose.cpp-style driver code
// compiles against this header unchanged apart from the include.  Errors surface
// as mkf::Error (a std::runtime_error), the counterpart of the cv::Exception the reference throws.
//
// Matrix type: with -DMKF_HAVE_OPENCV the shims use the real cv::Mat (CV_64F); otherwise the
// ~100-line row-major f64 subset below (rows, cols, at<double>, ptr<double>, clone, t, row, col,
// zeros, eye, empty, shallow copies with shared buffers like cv::Mat).
//
// All arithmetic runs on the GPU through libmkf_b200.so; nothing here computes the filter on the host.
#ifndef MKF_SHIMS_HPP
#define MKF_SHIMS_HPP

#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "mkf_b200.h"

#ifdef MKF_HAVE_OPENCV
#include <opencv2/core/core.hpp>
#else
namespace cv {
class Mat {
  public:
    int rows = 0, cols = 0;
    Mat() {}
    Mat(int r, int c) : rows(r), cols(c), buf_(new double[(size_t)r * c], std::default_delete<double[]>())
    {
        std::memset(buf_.get(), 0, sizeof(double) * (size_t)r * c);
    }
    static Mat zeros(int r, int c) { return Mat(r, c); }
    static Mat eye(int r, int c)
    {
        Mat m(r, c);
        for (int i = 0; i < r && i < c; i++) m.at<double>(i, i) = 1.0;
        return m;
    }
    bool empty() const { return rows == 0 || cols == 0; }
    template <class T>
    T& at(int r, int c)
    {
        static_assert(sizeof(T) == sizeof(double), "f64 only");
        return buf_.get()[(size_t)r * cols + c];
    }
    template <class T>
    const T& at(int r, int c) const
    {
        return buf_.get()[(size_t)r * cols + c];
    }
    template <class T>
    T* ptr(int r = 0)
    {
        return buf_.get() + (size_t)r * cols;
    }
    template <class T>
    const T* ptr(int r = 0) const
    {
        return buf_.get() + (size_t)r * cols;
    }
    Mat clone() const
    {
        Mat m(rows, cols);
        if (!empty()) std::memcpy(m.buf_.get(), buf_.get(), sizeof(double) * (size_t)rows * cols);
        return m;
    }
    Mat t() const
    {
        Mat m(cols, rows);
        for (int r = 0; r < rows; r++)
            for (int c = 0; c < cols; c++) m.at<double>(c, r) = at<double>(r, c);
        return m;
    }
    Mat row(int r) const // copy (the subset has no strided views)
    {
        Mat m(1, cols);
        std::memcpy(m.buf_.get(), ptr<double>(r), sizeof(double) * cols);
        return m;
    }
    Mat col(int c) const
    {
        Mat m(rows, 1);
        for (int r = 0; r < rows; r++) m.at<double>(r, 0) = at<double>(r, c);
        return m;
    }
    Mat rowRange(int r0, int r1) const
    {
        Mat m(r1 - r0, cols);
        std::memcpy(m.buf_.get(), ptr<double>(r0), sizeof(double) * (size_t)(r1 - r0) * cols);
        return m;
    }

  private:
    std::shared_ptr<double> buf_; // shallow copies share the buffer, as cv::Mat headers do
};
} // namespace cv
#endif

namespace mkf {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};
inline void check(int rc)
{
    if (rc < 0) throw Error(rc, std::string("libmkf_b200: ") + mkf_last_error());
}
inline std::vector<double> flat(const cv::Mat& m)
{
    std::vector<double> v((size_t)m.rows * m.cols);
    for (int r = 0; r < m.rows; r++)
        for (int c = 0; c < m.cols; c++) v[(size_t)r * m.cols + c] = m.at<double>(r, c);
    return v;
}
inline cv::Mat unflat(const double* p, int rows, int cols)
{
    cv::Mat m = cv::Mat::zeros(rows, cols
#ifdef MKF_HAVE_OPENCV
                               ,
                               CV_64F
#endif
    );
    for (int r = 0; r < rows; r++)
        for (int c = 0; c < cols; c++) m.at<double>(r, c) = p[(size_t)r * cols + c];
    return m;
}
// the reference seeds cv::RNG(cv::getTickCount()) in every resample() call (src/pf2DRao.cpp:179)
inline uint64_t tick_seed()
{
#ifdef MKF_HAVE_OPENCV
    return (uint64_t)cv::getTickCount();
#else
    return (uint64_t)std::chrono::steady_clock::now().time_since_epoch().count();
#endif
}
// getSamples draws from a process-wide stream with a fixed start, as cv::randn does from cv::theRNG()
// (src/pf2DRao.cpp:91); assign to it to reseed.  It does not touch the tick clock.
inline uint64_t& the_sample_stream()
{
    static uint64_t s = 0x9E3779B97F4A7C15ull;
    return s;
}
// observer of every getSamples() result (2 x N), for tests that need to know the candidates a driver drew
inline std::function<void(const cv::Mat&)>& sample_hook()
{
    static std::function<void(const cv::Mat&)> h;
    return h;
}
// parameters given to every my_gmm constructed afterwards (drivers such as PFTracker construct their
// ParticleFilter objects themselves, src/pfPose.cpp:58-59).  The class shims are the DROP-IN tier, so their default
// alias mode is what the reference binary computes: MKF_ALIAS_CV_SHALLOW_LITERAL (quirk B3, src/pf2DRao.cpp:153-156:
// slots that drew the same parent share its cv::Mat and are filtered sequentially in place).  The batch C ABI keeps
// MKF_ALIAS_INDEPENDENT as its default (mkf_params_default); set default_params().alias_mode to it to opt in here.
inline mkf_params& default_params()
{
    static mkf_params p = [] {
        mkf_params q;
        mkf_params_default(&q);
        q.alias_mode = MKF_ALIAS_CV_SHALLOW_LITERAL;
        return q;
    }();
    return p;
}
// cv::RNG draw order of resample(): one discarded int, then uniform(0.0, 1.0)
inline double cvrng_uniform_after_int(uint64_t seed)
{
    uint64_t st = seed ? seed : 0xffffffffull;
    auto next = [&]() {
        st = (uint64_t)(unsigned)st * 4164903690u + (unsigned)(st >> 32);
        return (unsigned)st;
    };
    (void)next();
    unsigned t = next();
    return (double)(((uint64_t)t << 32) | next()) * 5.4210108624275221700372640043497e-20;
}
} // namespace mkf

// ------------------------------------------------------------------------------------------------
class my_gmm;

class KF_model {
  public:
    KF_model() {}
    ~KF_model() {}
    cv::Mat Q, R, F, B, H, BH; // filled by my_gmm::loadGaussian exactly as src/my_gmm.cpp:53-72
    // x <- F x + B, P <- F P F^T + Q   (src/KF_model.cpp:11-15), on the device
    void predict(cv::Mat& state, cv::Mat& cov) { apply(1, nullptr, state, cov); }
    // y = z - (H x + BH), S = H P H^T + R, K = P H^T S^-1, x <- x + K y, P <- (I - K H) P   (src/KF_model.cpp:17-25)
    void update(cv::Mat measurement, cv::Mat& state, cv::Mat& cov) { apply(2, &measurement, state, cov); }

    // binding to the owning model (set by my_gmm)
    my_gmm* owner = nullptr;
    int component = -1;

  private:
    inline void apply(int stage, const cv::Mat* z, cv::Mat& state, cv::Mat& cov);
};

class state_params {
  public:
    state_params() {}
    ~state_params() {}
    state_params(const state_params& other) : state(other.state.clone()), cov(other.cov.clone()), weight(other.weight) {}
    state_params& operator=(const state_params&) = default; // shallow, like the reference's implicit operator=
    cv::Mat state;
    cv::Mat cov;
    double weight = 0.0;
};

class my_gmm {
  public:
    my_gmm() {}
    ~my_gmm() { release(); }
    my_gmm(const my_gmm&) = delete;
    my_gmm& operator=(const my_gmm&) = delete;

    // my_gmm::loadGaussian (src/my_gmm.cpp:45-75): u 1 x d, s d x d, H = pca_proj (d x D), m = pca_mean (1 x D)
    void loadGaussian(cv::Mat u, cv::Mat s, cv::Mat& H, cv::Mat& m, double w, double g)
    {
        mean.push_back(u);
        cov.push_back(s);
        weight.push_back(w);
        gamma_.push_back(g);
        if (proj_.empty()) {
            proj_ = mkf::flat(H);
            pmean_ = mkf::flat(m);
            d_ = H.rows;
            D_ = H.cols;
        }
        release(); // the device model is rebuilt lazily with the new component
        KF_model tracker;
        tracker.owner = this;
        tracker.component = (int)mean.size() - 1;
        KFtracker.push_back(tracker);
        members_stale_ = true;
    }
    // my_gmm::resetTracker (src/my_gmm.cpp:30-42)
    void resetTracker(std::vector<int> bins)
    {
        tracks.clear();
        for (int i = 0; i < nParticles; i++) {
            state_params temp;
            temp.state = mean[bins[i]].t();
            temp.cov = cov[bins[i]];
            temp.weight = 1.0 / (double)nParticles;
            tracks.push_back(temp); // deep copy through the copy constructor, as in the reference
        }
        upload_tracks();
    }
    std::vector<cv::Mat> mean;
    std::vector<cv::Mat> cov;
    std::vector<double> weight;
    std::vector<KF_model> KFtracker;
    std::vector<state_params> tracks; // host mirror; refreshed by syncTracks()
    int nParticles = 0;

    // ---- additions (not in the reference) ----
    mkf_params params = mkf::default_params();
    int device = 0;
    // refresh the host mirror `tracks` from the device (the reference keeps the state on the host)
    void syncTracks()
    {
        ensure_batch();
        const int d = d_;
        std::vector<double> x((size_t)nParticles * d), P((size_t)nParticles * d * d);
        mkf::check(mkf_batch_download(batch_, x.data(), P.data(), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                      MKF_MEM_HOST));
        tracks.resize(nParticles);
        for (int j = 0; j < nParticles; j++) {
            tracks[j].state = mkf::unflat(&x[(size_t)j * d], d, 1);
            tracks[j].cov = mkf::unflat(&P[(size_t)j * d * d], d, d);
            tracks[j].weight = 1.0 / (double)nParticles;
        }
    }
    void upload_tracks()
    {
        ensure_batch();
        const int d = d_;
        std::vector<double> x((size_t)nParticles * d), P((size_t)nParticles * d * d);
        for (int j = 0; j < nParticles; j++) {
            for (int i = 0; i < d; i++) x[(size_t)j * d + i] = tracks[j].state.at<double>(i, 0);
            for (int r = 0; r < d; r++)
                for (int c = 0; c < d; c++) P[((size_t)j * d + r) * d + c] = tracks[j].cov.at<double>(r, c);
        }
        mkf::check(mkf_batch_upload(batch_, x.data(), P.data(), MKF_MEM_HOST));
    }
    mkf_model* model()
    {
        if (!model_) {
            const int K = (int)mean.size(), d = d_;
            if (K == 0) throw mkf::Error(MKF_E_INVALID, "my_gmm: no Gaussian loaded");
            std::vector<double> mu((size_t)K * d), cv_((size_t)K * d * d);
            for (int k = 0; k < K; k++) {
                for (int i = 0; i < d; i++) mu[(size_t)k * d + i] = mean[k].at<double>(0, i);
                for (int r = 0; r < d; r++)
                    for (int c = 0; c < d; c++) cv_[((size_t)k * d + r) * d + c] = cov[k].at<double>(r, c);
            }
            mkf::check(mkf_model_create(&model_, K, d, D_, mu.data(), cv_.data(), weight.data(), gamma_.data(),
                                        proj_.data(), pmean_.data(), &params));
        }
        if (members_stale_) fill_kf_members();
        return model_;
    }
    mkf_batch* batch()
    {
        ensure_batch();
        return batch_;
    }
    int d() const { return d_; }
    int D() const { return D_; }

  private:
    void ensure_batch()
    {
        if (!batch_) mkf::check(mkf_batch_create(&batch_, model(), 1, nParticles, device, nullptr));
    }
    void fill_kf_members()
    {
        members_stale_ = false;
        const int K = (int)mean.size(), d = d_;
        std::vector<double> Q((size_t)K * d * d), B((size_t)K * d), H(6 * (size_t)d), BH(6);
        mkf::check(mkf_model_get(model_, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, Q.data(), B.data(),
                                 H.data(), BH.data()));
        for (int k = 0; k < K; k++) {
            KF_model& t = KFtracker[k];
            t.owner = this;
            t.component = k;
            t.Q = mkf::unflat(&Q[(size_t)k * d * d], d, d);
            t.B = mkf::unflat(&B[(size_t)k * d], d, 1);
            t.H = mkf::unflat(H.data(), 6, d);
            t.BH = mkf::unflat(BH.data(), 6, 1);
            t.R = cv::Mat::eye(6, 6
#ifdef MKF_HAVE_OPENCV
                               ,
                               CV_64F
#endif
            );
            for (int i = 0; i < 6; i++) t.R.at<double>(i, i) = params.meas_noise_var;
            t.F = cv::Mat::eye(d, d
#ifdef MKF_HAVE_OPENCV
                               ,
                               CV_64F
#endif
            );
            for (int i = 0; i < d; i++) t.F.at<double>(i, i) = gamma_[k];
        }
    }
    void release()
    {
        if (batch_) mkf_batch_destroy(batch_);
        if (model_) mkf_model_destroy(model_);
        batch_ = nullptr;
        model_ = nullptr;
    }
    std::vector<double> gamma_, proj_, pmean_;
    int d_ = 0, D_ = 0;
    mkf_model* model_ = nullptr;
    mkf_batch* batch_ = nullptr;
    bool members_stale_ = false;
};

inline void KF_model::apply(int stage, const cv::Mat* z, cv::Mat& state, cv::Mat& cov)
{
    if (!owner) throw mkf::Error(MKF_E_INVALID, "KF_model is not bound to a my_gmm (construct it through loadGaussian)");
    const int d = owner->d();
    if (state.rows * state.cols != d || cov.rows != d || cov.cols != d || (z && z->rows * z->cols != 6))
        throw mkf::Error(MKF_E_INVALID, "KF_model: size mismatch"); // the reference: cv::Exception from gemm
    std::vector<double> x = mkf::flat(state), P = mkf::flat(cov), zz;
    if (z) zz = mkf::flat(*z);
    int32_t comp = component;
    mkf::check(mkf_kf_apply(owner->model(), 1, &comp, stage, x.data(), P.data(), z ? zz.data() : nullptr, nullptr,
                            owner->device));
    // write in place: cv::Mat assignment from a MatExpr reuses the destination buffer (quirk B3 relies on it)
    for (int i = 0; i < d; i++) state.at<double>(i, 0) = x[i];
    for (int r = 0; r < d; r++)
        for (int c = 0; c < d; c++) cov.at<double>(r, c) = P[(size_t)r * d + c];
}

class ParticleFilter {
  public:
    ParticleFilter(int nParticles) { gmm.nParticles = nParticles; } // src/pf2DRao.cpp:13-16
    ~ParticleFilter() {}

    // ParticleFilter::update (src/pf2DRao.cpp:125-158): measurement is 6 x N, one column per slot
    void update(cv::Mat measurement)
    {
        const int N = gmm.nParticles;
        if (measurement.rows != 6 || measurement.cols != N)
            throw mkf::Error(MKF_E_INVALID, "ParticleFilter::update: measurement must be 6 x nParticles");
        std::vector<double> z = mkf::flat(measurement);
        const uint64_t s_ind = next_seed(), s_post = next_seed();
        double u_ind = mkf::cvrng_uniform_after_int(s_ind), u_post = mkf::cvrng_uniform_after_int(s_post);
        uint64_t seeds[2] = {s_ind, s_post};
        mkf::check(mkf_batch_update(gmm.batch(), z.data(), MKF_MEAS_PER_SLOT, &u_ind, &u_post, seeds, MKF_MEM_HOST));
        last_u_ind = u_ind;
        last_u_post = u_post;
    }
    // ParticleFilter::getEstimator (src/pf2DRao.cpp:23-31): d x 1
    cv::Mat getEstimator()
    {
        std::vector<double> xb(gmm.d());
        mkf::check(mkf_batch_estimate(gmm.batch(), xb.data(), nullptr, MKF_MEM_HOST));
        return mkf::unflat(xb.data(), gmm.d(), 1);
    }
    // ParticleFilter::getSamples (src/pf2DRao.cpp:85-103): N proposals ~ N(hand estimate, (0.8 scale)^2) per axis
    // (cv::randn takes C = 0.8*scale*I as a standard-deviation matrix, quirk B10).  Drawn on the device by
    // mkf_batch_propose from the counter generator of mkf_synth.h (cv::randn's stream is not reproducible);
    // distribution-equivalent to second order, not bit-equivalent, to cv::randn.
    cv::Mat getSamples(cv::Mat H, cv::Mat M, int N, double scale)
    {
        (void)H;
        (void)M; // the model already holds pca_proj / pca_mean
        const double roi[4] = {0.0, 0.0, scale, scale};
        std::vector<double> xy((size_t)4 * N); // 2 hands x 2 rows x N; this filter plays both arms, hand 0 is used
        uint64_t& stream = mkf::the_sample_stream();
        stream = stream * 6364136223846793005ull + 1442695040888963407ull;
        mkf::check(mkf_batch_propose(gmm.batch(), gmm.batch(), N, roi, nullptr, nullptr, 1, stream, sample_calls_++, 0,
                                     xy.data(), nullptr, MKF_MEM_HOST));
        cv::Mat out = mkf::unflat(xy.data(), 2, N);
        if (mkf::sample_hook()) mkf::sample_hook()(out);
        return out;
    }
    // ParticleFilter::getSampleProb (src/pf2DRao.cpp:105-122)
    void getSampleProb(cv::Mat H, cv::Mat M, cv::Mat input1, cv::Mat input2, std::vector<double>& weight1,
                       std::vector<double>& weight2, double scale)
    {
        (void)H;
        (void)M;
        weight1.assign(input1.cols, 0);
        weight2.assign(input2.cols, 0);
        std::vector<double> c1 = mkf::flat(input1), c2 = mkf::flat(input2);
        mkf::check(mkf_batch_sample_prob(gmm.batch(), 0, c1.data(), input1.cols, scale, weight1.data()));
        mkf::check(mkf_batch_sample_prob(gmm.batch(), 0, c2.data(), input2.cols, scale, weight2.data()));
    }
    my_gmm gmm;
    // ParticleFilter::resample (src/pf2DRao.cpp:175-210)
    std::vector<int> resample(std::vector<double> weights, int N)
    {
        std::vector<int32_t> out(N);
        const uint64_t seed = next_seed();
        last_u = mkf::cvrng_uniform_after_int(seed);
        mkf::check(mkf_resample(weights.data(), (int)weights.size(), N, -1.0, seed, out.data(), gmm.device));
        return std::vector<int>(out.begin(), out.end());
    }

    // ---- additions: deterministic seeding for tests (the reference is clock-seeded) ----
    void setSeed(uint64_t s)
    {
        seeded_ = true;
        seed_ = s;
    }
    double last_u = 0, last_u_ind = 0, last_u_post = 0;

  protected:
    uint64_t next_seed()
    {
        if (!seeded_) return mkf::tick_seed();
        seed_ = seed_ * 6364136223846793005ull + 1442695040888963407ull;
        return seed_ | 1ull;
    }
    bool seeded_ = false;
    uint64_t seed_ = 0;
    uint64_t sample_calls_ = 0;
};

#endif // MKF_SHIMS_HPP
