// shim_driver.cpp -- ROS-free re-enactment of PFTracker's use of the hot path
// (constructor src/pfPose.cpp:34-71, per-frame src/pfPose.cpp:303-326,347-348) written against
// include/mkf_shims.hpp exactly as pfPose.cpp is written against the reference's headers.
// Prints a line-oriented trace that tests/test_gpu_shims.py checks against the CPU oracle.
#include <cstdio>
#include <cstdlib>
#include <string>

#include "../include/mkf_shims.hpp"
#include "../include/mkf_synth.h"

static cv::Mat from(const std::vector<double>& v, int off, int rows, int cols)
{
    return mkf::unflat(v.data() + off, rows, cols);
}

int main(int argc, char** argv)
{
    if (argc < 5) {
        fprintf(stderr, "usage: shim_driver left.yml right.yml frames nParticles\n");
        return 2;
    }
    const std::string left = argv[1], right = argv[2];
    const int frames = atoi(argv[3]), numParticles = atoi(argv[4]);
    try {
        // --- cv::FileStorage part of the constructor: both arms, gamma from the right-arm file (quirk B4)
        mkf_model* loader = nullptr;
        mkf::check(mkf_model_load_yaml(&loader, left.c_str(), right.c_str(), nullptr));
        int K, d, D;
        mkf_model_dims(loader, &K, &d, &D);
        std::vector<double> means((size_t)K * d), covs((size_t)K * d * d), weights(K), g(K), h_pca_v((size_t)d * D),
            m_pca_v(D);
        mkf::check(mkf_model_get(loader, means.data(), covs.data(), weights.data(), g.data(), h_pca_v.data(),
                                 m_pca_v.data(), nullptr, nullptr, nullptr, nullptr));
        mkf_model_destroy(loader);
        cv::Mat h1_pca = from(h_pca_v, 0, d, D), m1_pca = from(m_pca_v, 0, 1, D);

        ParticleFilter* pf1 = new ParticleFilter(numParticles); // left arm pf
        pf1->setSeed(0x5EED0001);
        for (int i = 0; i < K; i++)
            pf1->gmm.loadGaussian(from(means, i * d, 1, d), from(covs, i * d * d, d, d), h1_pca, m1_pca, weights[i], g[i]);

        // Initialise particle filter
        std::vector<int> bins1 = pf1->resample(pf1->gmm.weight, numParticles);
        printf("INIT_U %.17g\n", pf1->last_u);
        pf1->gmm.resetTracker(bins1);

        // KF_model members as the reference exposes them
        printf("KF0_Q00 %.17g KF0_B0 %.17g H00 %.17g BH0 %.17g R00 %.17g F00 %.17g\n",
               pf1->gmm.KFtracker[0].Q.at<double>(0, 0), pf1->gmm.KFtracker[0].B.at<double>(0, 0),
               pf1->gmm.KFtracker[0].H.at<double>(0, 0), pf1->gmm.KFtracker[0].BH.at<double>(0, 0),
               pf1->gmm.KFtracker[0].R.at<double>(0, 0), pf1->gmm.KFtracker[0].F.at<double>(0, 0));

        // standalone KF_model::predict / update on component 3 starting from (mean_3, cov_3)
        {
            cv::Mat x = pf1->gmm.mean[3].t(), P = pf1->gmm.cov[3].clone();
            pf1->gmm.KFtracker[3].predict(x, P);
            printf("KFPRED");
            for (int i = 0; i < d; i++) printf(" %.17g", x.at<double>(i, 0));
            printf(" %.17g %.17g\n", P.at<double>(0, 0), P.at<double>(d - 1, 2));
            cv::Mat z(6, 1);
            double zz[6];
            mkf_synth_meas(0x5EED0001, 0, 0, -1, 0, zz);
            for (int r = 0; r < 6; r++) z.at<double>(r, 0) = zz[r];
            pf1->gmm.KFtracker[3].update(z, x, P);
            printf("KFUPD");
            for (int i = 0; i < d; i++) printf(" %.17g", x.at<double>(i, 0));
            printf(" %.17g %.17g\n", P.at<double>(0, 0), P.at<double>(d - 1, 2));
        }

        for (int fr = 0; fr < frames; fr++) {
            cv::Mat measurement1(6, numParticles);
            for (int i = 0; i < numParticles; i++) {
                double z[6];
                mkf_synth_meas(0x5EED0001, 0, (uint64_t)fr, i, 0, z);
                for (int r = 0; r < 6; r++) measurement1.at<double>(r, i) = z[r];
            }
            pf1->update(measurement1); // particle filter measurement left arm
            // cv::Mat e1 = h1_pca.t()*pf1->getEstimator() + m1_pca.t();
            cv::Mat xb = pf1->getEstimator();
            printf("FRAME %d U %.17g %.17g XBAR", fr, pf1->last_u_ind, pf1->last_u_post);
            for (int i = 0; i < d; i++) printf(" %.17g", xb.at<double>(i, 0));
            printf("\n");
        }
        // getSampleProb on a few candidate positions
        {
            cv::Mat in1(2, 3), in2(2, 2);
            double c1[6] = {380, 390, 100, 250, 260, 400}, c2[4] = {388, 10, 250, 20};
            for (int i = 0; i < 3; i++) {
                in1.at<double>(0, i) = c1[i];
                in1.at<double>(1, i) = c1[3 + i];
            }
            for (int i = 0; i < 2; i++) {
                in2.at<double>(0, i) = c2[i];
                in2.at<double>(1, i) = c2[2 + i];
            }
            std::vector<double> w1, w2;
            pf1->getSampleProb(h1_pca.t(), m1_pca.t(), in1, in2, w1, w2, 47.0);
            printf("PROB %.17g %.17g %.17g %.17g %.17g\n", w1[0], w1[1], w1[2], w2[0], w2[1]);
            cv::Mat smp = pf1->getSamples(h1_pca.t(), m1_pca.t(), 2000, 47.0);
            double mx = 0, my = 0;
            for (int i = 0; i < 2000; i++) {
                mx += smp.at<double>(0, i) / 2000;
                my += smp.at<double>(1, i) / 2000;
            }
            printf("SAMPLES_MEAN %.17g %.17g\n", mx, my);
        }
        pf1->gmm.syncTracks();
        printf("TRACK0");
        for (int i = 0; i < d; i++) printf(" %.17g", pf1->gmm.tracks[0].state.at<double>(i, 0));
        printf("\n");
        // state_params copy semantics (src/my_gmm.cpp:11-16): copy-construct = deep, assign = shallow
        state_params a = pf1->gmm.tracks[0];
        state_params b2;
        b2 = pf1->gmm.tracks[0];
        pf1->gmm.tracks[0].state.at<double>(0, 0) = 12345.0;
        printf("COPY deep %d shallow %d\n", a.state.at<double>(0, 0) != 12345.0, b2.state.at<double>(0, 0) == 12345.0);
        // error behaviour: a wrongly sized measurement raises (the reference: cv::Exception)
        try {
            pf1->update(cv::Mat(6, 3));
            printf("ERR none\n");
        } catch (const mkf::Error& e) {
            printf("ERR %d\n", e.code);
        }
        delete pf1;
    } catch (const std::exception& e) {
        printf("FATAL %s\n", e.what());
        return 1;
    }
    return 0;
}
