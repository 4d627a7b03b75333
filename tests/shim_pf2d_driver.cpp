// Drives the legacy plain-particle-filter shims (include/mkf_shims_pf2d.hpp: my_gmm, ParticleFilter of src/pf2D.h)
// the way a node would drive the reference classes, and prints a trace tests/test_gpu_shims.py replays through the
// CPU oracle:  shim_pf2d_driver N d frames seed
#include <cstdio>
#include <cstdlib>

#include "../include/mkf_shims_pf2d.hpp"

using mkf_legacy::normal01;
using mkf_legacy::the_stream;
using mkf_legacy::uniform01;

static void dump(const char* tag, const cv::Mat& m)
{
    printf("%s %d %d", tag, m.rows, m.cols);
    for (int r = 0; r < m.rows; r++)
        for (int c = 0; c < m.cols; c++) printf(" %.17g", m.at<double>(r, c));
    printf("\n");
}

int main(int argc, char** argv)
{
    const int N = argc > 1 ? atoi(argv[1]) : 300, d = argc > 2 ? atoi(argv[2]) : 8, frames = argc > 3 ? atoi(argv[3]) : 3;
    const unsigned seed = argc > 4 ? (unsigned)atoi(argv[4]) : 7u;
    try {
        srand(seed);
        the_stream() = 0x1234ull + seed;
        mkf_legacy::ParticleFilter pf(N, d, true); // the reference's `ParticleFilter(numParticles, numDims, side1)`
        const int K = 5;
        for (int k = 0; k < K; k++) {
            cv::Mat u(1, d), s(d, d);
            cv::Mat a(d, d);
            for (int c = 0; c < d; c++) u.at<double>(0, c) = 100.0 + 300.0 * uniform01();
            for (int r = 0; r < d; r++)
                for (int c = 0; c < d; c++) a.at<double>(r, c) = normal01();
            for (int r = 0; r < d; r++)
                for (int c = 0; c < d; c++) {
                    double t = (r == c) ? (double)d : 0.0;
                    for (int q = 0; q < d; q++) t += a.at<double>(r, q) * a.at<double>(c, q);
                    s.at<double>(r, c) = 900.0 * t;
                }
            pf.gmm.loadGaussian(u, s, 1.0 / K);
            dump("MEAN", u);
            dump("SIGMA", s);
            dump("SIGMA_I", pf.gmm.sigma_i[k]);
            printf("DET_S %.17g\n", pf.gmm.det_s[k]);
        }
        // particles around the components rather than all over the image, so that weights do not underflow
        cv::Mat p0 = pf.getParticles();
        printf("RANGE %.17g %.17g %.17g %.17g\n", p0.at<double>(0, 6), p0.at<double>(1, 6), p0.at<double>(0, 0), p0.at<double>(0, 1));
        for (int i = 0; i < N; i++)
            for (int c = 0; c < d; c++) p0.at<double>(i, c) = pf.gmm.mean[i % K].at<double>(0, c) + 25.0 * normal01();
        pf.setParticles(p0);
        dump("EST0", pf.getEstimator());
        for (int f = 0; f < frames; f++) {
            cv::Mat cur = pf.getParticles();
            dump("PART", cur);
            cv::Mat z(2, 2);
            for (int i = 0; i < N; i++) {
                z.at<double>(0, 0) += cur.at<double>(i, 6) / N;
                z.at<double>(0, 1) += cur.at<double>(i, 7) / N;
                z.at<double>(1, 0) += cur.at<double>(i, 0) / N;
                z.at<double>(1, 1) += cur.at<double>(i, 1) / N;
            }
            dump("MEAS", z);
            pf.update(z);
            printf("U %.17g\n", pf.last_u);
            printf("NOISE %d %d", N, d);
            for (double v : pf.last_noise) printf(" %.17g", v);
            printf("\n");
            dump("AFTER", pf.getParticles());
            dump("EST", pf.getEstimator());
        }
        pf.predict(); // stand-alone call: host path
        dump("PRED", pf.getParticles());
    } catch (const std::exception& e) {
        fprintf(stderr, "shim_pf2d_driver: %s\n", e.what());
        return 1;
    }
    return 0;
}
