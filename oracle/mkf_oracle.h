/* mkf_oracle.h -- C interface of the CPU oracle (TEST INFRASTRUCTURE, not product code).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  The product (libmkf_b200.so) never links or calls it.
 *
 * PINNING: mgb45/mkfbodytracker_pdaf ships no tests, golden vectors or recorded outputs (SURVEY.md
 * section 4).  The restatement is pinned against the reference itself run here -- oracle/_ref: the
 * reference's own .cpp files compiled in place against oracle/cvshim (oracle/Makefile), bit-exact
 * agreement required by tests/test_ref_sources.py and tests/test_ref_tracker.py -- and cross-checked
 * against an independent numpy restatement, analytic known-answer tests and OpenCV-python 4.13
 * primitives (gemm / invert(DECOMP_LU) / FileStorage).  Unpinned: the OpenCV `core` arithmetic
 * underneath (version not pinned by the reference, not installed here), restated from the published
 * 2.4.x algorithms.
 */
#ifndef MKF_ORACLE_H
#define MKF_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* semantics switches, SURVEY.md section 8(c) */
enum { ORC_CHOL_CV24_LITERAL = 0, ORC_CHOL_CV3_LITERAL = 1, ORC_CHOL_EXACT = 2 };
enum { ORC_ALIAS_INDEPENDENT = 0, ORC_ALIAS_CV_SHALLOW_LITERAL = 1 };

typedef struct orc_model orc_model;
typedef struct orc_filter orc_filter;

/* my_gmm::loadGaussian for all K components (src/my_gmm.cpp:45-75, src/pfPose.cpp:61-65).
 * means K x d, covs K x d x d (stacked), weights K, gamma K, pca_proj d x D, pca_mean D. */
orc_model* orc_model_create(int K, int d, int D, const double* means, const double* covs, const double* weights,
                            const double* gamma, const double* pca_proj, const double* pca_mean);
void orc_model_destroy(orc_model* m);
/* copies out derived constants (any pointer may be NULL): H 6 x d, BH 6, Q K x d x d, B K x d, R 6 x 6 */
void orc_model_get(const orc_model* m, double* H, double* BH, double* Q, double* B, double* R);

/* ParticleFilter(int nParticles) (src/pf2DRao.cpp:13-16) bound to a model */
orc_filter* orc_filter_create(const orc_model* m, int N, int chol_mode, int alias_mode);
void orc_filter_destroy(orc_filter* f);

/* bins = resample(gmm.weight, N); gmm.resetTracker(bins) (src/pfPose.cpp:68-71, src/my_gmm.cpp:30-42).
 * u >= 0: injected uniform draw; u < 0: drawn from cv::RNG(seed) as the reference does. */
int orc_filter_reset(orc_filter* f, double u, uint64_t seed);

/* ParticleFilter::update (src/pf2DRao.cpp:125-158).  meas is the 6 x N row-major matrix the
 * reference passes (column j = slot j).  Optional outputs (NULL to skip): raw (un-normalised)
 * weights N, normalised weights N, component indicators N, resampled parents N, wsum.
 * returns status bits: 1 = indicator resample degenerate, 2 = posterior resample degenerate,
 * 4 = at least one chol() failure. */
int orc_filter_update(orc_filter* f, const double* meas, double u_ind, uint64_t seed_ind, double u_post,
                      uint64_t seed_post, double* w_raw, double* w_norm, int32_t* indicators, int32_t* parents,
                      double* wsum);
/* same with one shared measurement column z[6] replicated to all N slots */
int orc_filter_update_shared(orc_filter* f, const double* z6, double u_ind, uint64_t seed_ind, double u_post,
                             uint64_t seed_post, double* w_raw, double* w_norm, int32_t* indicators,
                             int32_t* parents, double* wsum);

/* state access: x N x d, P N x d x d (row-major full matrices) */
void orc_filter_get_state(const orc_filter* f, double* x, double* P);
void orc_filter_set_state(orc_filter* f, const double* x, const double* P);

/* getEstimator (src/pf2DRao.cpp:23-31) and e = pca_proj^T xbar + pca_mean^T (src/pfPose.cpp:347-348) */
void orc_filter_estimate(const orc_filter* f, double* xbar, double* pose);

/* ParticleFilter::resample (src/pf2DRao.cpp:175-210).  returns 1 if the degenerate (max weight 0/NaN)
 * fallback was taken, else 0. */
int orc_resample(const double* w, int L, int N, double u, uint64_t seed, int32_t* out);

/* single pieces, for known-answer tests */
void orc_kf_predict(const orc_model* m, int k, double* x, double* P);                 /* src/KF_model.cpp:11-15 */
void orc_kf_update(const orc_model* m, int k, const double* z, double* x, double* P); /* src/KF_model.cpp:17-25 */
/* mvnpdf (src/pf2DRao.cpp:56-67) on an n x n sigma; returns the weight; *chol_ok = cv::Cholesky result */
double orc_mvnpdf(int n, const double* x, const double* u, const double* sigma, int chol_mode, int* chol_ok);
/* chol() wrapper (src/pf2DRao.cpp:34-53): out n x n */
int orc_chol(int n, const double* in, double* out, int chol_mode);
/* cv::invert(DECOMP_LU) restatement; returns 0 if singular */
int orc_invert_lu(int n, const double* in, double* out);

/* cv::RNG restatement (OpenCV core/operations.hpp, from memory): fills out_int with
 * n_int draws of uniform(0,L) and then out_dbl with n_dbl draws of uniform(0.0,1.0) */
void orc_cvrng(uint64_t seed, int L, int n_int, int32_t* out_int, int n_dbl, double* out_dbl);

/* association ("PDAF") step of PFTracker::getMeasurementProposal (src/pfPose.cpp:238-323) for one
 * person: two arm filters, C candidates per hand.  cand_xy: 2 hands x 2 rows(x,y) x C (the
 * reference's 2 x N props matrices), cand_L: 2 x C uint8 likelihood-image samples, roi: x,y,w,h,
 * img_rows/img_cols for the gate.  Outputs: gate 2 x C (0/1), weights 2 x C (normalised),
 * bins 2 x N, meas 2 x 6 x N.  u_cand[2] / seed_cand[2] as for orc_resample.  xbar* override the
 * filters' estimators when non-NULL (used for teacher-forced tests). */
int orc_associate(const orc_filter* armL, const orc_filter* armR, int C, const double* cand_xy,
                  const uint8_t* cand_L, const double* roi, int img_rows, int img_cols, const double* u_cand,
                  const uint64_t* seed_cand, uint8_t* gate, double* weights, int32_t* bins, double* meas);

/* candidate generation front-end for one person (src/pfPose.cpp:216-236, src/pf2DRao.cpp:85-103) on the shared
 * counter generator: cand_xy 2 x 2 x C, cand_L 2 x C (NULL with like NULL) */
void orc_propose(const orc_filter* armL, const orc_filter* armR, int C, const double* roi, int tracking,
                 const uint8_t* like, int rows, int cols, uint64_t seed, uint64_t track, uint64_t frame, double* cand_xy,
                 uint8_t* cand_L);

/* output back-end (src/pfPose.cpp:84-208): get3Dpose of one D-vector estimate with camera matrix Kc (3 x 3) ->
 * pos3D 3 x 5; skeleton = publishTFtree translations (9 x 3, broadcast order) + camera Euler triple (tf 10 x 3)
 * and publish2Dpos joints (8 x 2) from the two arms' estimates */
void orc_get3dpose(const double* estimate, const double* Kc, double* pos3D);
void orc_skeleton(const double* e1, const double* e2, const double* Kc, double* tf, double* joints2d);

/* legacy plain particle filter (src/pf2D.cpp, uncompiled in the reference) */
typedef struct orc_pf2d orc_pf2d;
/* gmm: K components over d dims: means K x d, covs K x d x d, weights K (my_gmm::loadGaussian, src/pf2D.cpp:28-37) */
orc_pf2d* orc_pf2d_create(int N, int d, int K, const double* means, const double* covs, const double* weights);
void orc_pf2d_destroy(orc_pf2d* p);
void orc_pf2d_set_particles(orc_pf2d* p, const double* particles /* N x d */);
/* constructor / degenerate-branch randomisation (src/pf2D.cpp:44-71,232-250) from the counter generator of mkf_synth.h */
void orc_pf2d_set_random(orc_pf2d* p, uint64_t seed, uint64_t track, int side, int im_w, int im_h);
void orc_pf2d_randomise(orc_pf2d* p);
/* 1: the `noise` of orc_pf2d_update is what predict() adds (the reference's cv::randn(.., 0, 5) values), not N(0,1) draws */
void orc_pf2d_set_noise_scaled(orc_pf2d* p, int on);
/* include/mkf_expf.h (glibc's expf restated) and the host libm's expf, for tests/test_expf.py */
float orc_expf(float x);
float orc_libm_expf(float x);
/* n arguments given as float bit patterns; returns the number of bit-level mismatches between the two */
uint64_t orc_expf_compare(const uint32_t* bits, uint64_t n);
void orc_pf2d_get_particles(const orc_pf2d* p, double* particles, double* weights);
/* ParticleFilter::getEstimator (src/pf2D.cpp:79-88): est[d] = sum_i weights[i] * particles.row(i), index order */
void orc_pf2d_estimate(const orc_pf2d* p, double* est);
void orc_pf2d_get_gmm(const orc_pf2d* p, double* sigma_i /* K x d x d */, double* det_s /* K */);
/* update(measurement 2 x 2) = weight + normalise + resample + predict (src/pf2D.cpp:148-210).
 * u: injected uniform for resample (replaces rand()/RAND_MAX); noise N x d standard normals scaled by 5
 * inside (replaces cv::randn(.,0,5)); outputs optional: normalised weights N, parents N. */
int orc_pf2d_update(orc_pf2d* p, const double* meas, double u, const double* noise, double* w_norm, int32_t* parents);

/* batched, OpenMP-parallel-over-tracks CPU baseline on the synthetic workload of include/mkf_synth.h.
 * Runs T independent filters of N slots for `frames` frames (shared measurement column if per_slot==0)
 * and returns wall seconds of the frame loop; pose_out (T x D) may be NULL.  threads<=0: all cores. */
double orc_bench_tracks(const orc_model* m, int64_t T, int N, int frames, int per_slot, uint64_t seed, int jitter,
                        int chol_mode, int alias_mode, int threads, double* pose_out, int* threads_used);

#ifdef __cplusplus
}
#endif
#endif
