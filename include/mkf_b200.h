/* mkf_b200.h -- C ABI of libmkf_b200.so: the B200-native (sm_100a) implementation of the
 * per-frame filtering hot path of mgb45/mkfbodytracker_pdaf.
 *
 * The reference has no FFI/plugin boundary: the path is ordinary C++ classes linked into the
 * `poseTracker` node (CMakeLists.txt:29) and called from PFTracker (src/pfPose.cpp:58-71,
 * 222-223, 241-242, 300-301, 325-326, 347-348).  This header is the boundary a maintainer
 * would bind instead; every entry point cites the reference interface it replaces
 * (file:line relative to the reference repository).  include/mkf_shims.hpp layers the
 * reference's own class names (KF_model, my_gmm, state_params, ParticleFilter) on top.
 *
 * Conventions
 *  - plain C: pointers + sizes, no C++/torch types; 0 = success, negative MKF_E_* = failure
 *    (the reference signals errors with cv::Exception; no exception crosses this boundary);
 *    mkf_last_error() returns a thread-local message.
 *  - the caller owns every buffer it passes; the library owns device memory behind the
 *    opaque handles.  Every data pointer may be host or device memory (`mem` argument).
 *  - work is enqueued on the batch's CUDA stream; functions that fill HOST buffers
 *    synchronise that stream before returning, all others are asynchronous
 *    (mkf_batch_sync waits).  One handle = one host thread at a time.
 *  - "track" = one ParticleFilter (one arm filter), "slot" = one particle (one Gaussian
 *    (x, P) of dimension d).  All reals are IEEE double, as in the reference (CV_64F).
 *  - uniform draws are INPUTS (the reference seeds cv::RNG from the clock inside resample,
 *    src/pf2DRao.cpp:179, which is not reproducible); seeds only feed the degenerate
 *    random-index fallback (src/pf2DRao.cpp:184-192).
 *  - there is no CPU fallback: without a CUDA device every compute entry point fails with
 *    MKF_E_CUDA.
 */
#ifndef MKF_B200_H
#define MKF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MKF_ABI_VERSION 1

/* error codes */
#define MKF_OK 0
#define MKF_E_INVALID (-1)     /* bad argument (what cv::Exception size/type asserts would catch) */
#define MKF_E_CUDA (-2)        /* CUDA runtime failure / no device */
#define MKF_E_NOMEM (-3)       /* host or device allocation failed */
#define MKF_E_IO (-4)          /* model file unreadable */
#define MKF_E_PARSE (-5)       /* model file malformed */
#define MKF_E_UNSUPPORTED (-6) /* valid in the reference but not built here */

/* semantics switches; defaults in mkf_params_default */
#define MKF_CHOL_CV24_LITERAL 0 /* ParticleFilter::chol as it behaves on OpenCV 2.4 (src/pf2DRao.cpp:34-53) */
#define MKF_CHOL_CV3_LITERAL 1  /* ... on OpenCV >= 3.0 (diagonal convention of cv::Cholesky differs) */
#define MKF_CHOL_EXACT 2        /* true upper Cholesky factor (MATLAB mvnpdf, the evident intent) */
#define MKF_ALIAS_INDEPENDENT 0 /* each slot owns its Gaussian (north_star: one KF per particle) */
#define MKF_ALIAS_CV_SHALLOW_LITERAL 1 /* quirk B3 (src/pf2DRao.cpp:153-156): slots that drew the same parent
                                          share its buffer and are filtered sequentially in place, as the
                                          reference binary really does */

/* memory space of caller pointers */
#define MKF_MEM_AUTO 0 /* ask the driver (cudaPointerGetAttributes) */
#define MKF_MEM_HOST 1
#define MKF_MEM_DEVICE 2
#define MKF_MEM_HOST_ASYNC 3 /* PINNED host memory, no synchronisation; honoured by mkf_batch_update and
                                mkf_batch_estimate.  Inputs must hold their values when the call is made; the
                                library copies them on an internal copy stream (overlapping the previous frame's
                                kernels), orders its kernels after them, and copies results back on a second
                                internal stream.  The caller waits (mkf_batch_sync, which drains all three
                                streams) before reusing the input buffers or reading the outputs. */

/* measurement layouts accepted by mkf_batch_update */
#define MKF_MEAS_SHARED 0   /* T x 6       : one column per track, replicated to its N slots */
#define MKF_MEAS_PER_SLOT 1 /* T x 6 x N   : the reference's 6 x N cv::Mat per track (src/pf2DRao.cpp:125) */

/* per-track status bits (mkf_batch_download / mkf_batch_status) */
#define MKF_ST_IND_FALLBACK 0x1u    /* indicator resample resolved by the exact sequential loop */
#define MKF_ST_POST_FALLBACK 0x2u   /* posterior resample resolved by the exact sequential loop */
#define MKF_ST_POST_DEGENERATE 0x4u /* max weight 0/NaN: random-index fallback taken (src/pf2DRao.cpp:184-192) */
#define MKF_ST_CHOL_FAIL 0x8u       /* cv::Cholesky failed for at least one slot (src/pf2DRao.cpp:37) */
#define MKF_ST_CAND_FALLBACK 0x10u  /* candidate resample resolved by the exact sequential loop */
#define MKF_ST_CAND_DEGENERATE 0x20u /* no candidate passed the gate: random candidates (quirk B11) */
#define MKF_ST_IND_WRAP 0x40u       /* indicator resample wrapped past the last component */

typedef struct mkf_model mkf_model; /* one arm model: GMM prior + per-component KF constants */
typedef struct mkf_batch mkf_batch; /* T independent tracks x N slots on one GPU */

/* every hard-coded literal of the reference on this path, with the reference value as default */
typedef struct mkf_params {
    int chol_mode;         /* MKF_CHOL_*                                  default CV24_LITERAL */
    int alias_mode;        /* MKF_ALIAS_*                                 default INDEPENDENT  */
    double meas_noise_var; /* R = 100 * I6              src/my_gmm.cpp:54                      */
    double assoc_pa;       /* Pa = 0.05                 src/pfPose.cpp:247                     */
    double assoc_clutter;  /* 1e-4                      src/pfPose.cpp:261                     */
    double proposal_spread;/* 0.8 (x roi width)         src/pf2DRao.cpp:90,114                 */
    double neck_offset;    /* 1.65 (x roi height)       src/pfPose.cpp:313                     */
    int img_rows, img_cols;/* 480 x 640 gate            src/pfPose.cpp:251                     */
} mkf_params;
void mkf_params_default(mkf_params* p);

const char* mkf_last_error(void);
int mkf_abi_version(void);
/* number of CUDA devices visible (0 when there is none; never fails) */
int mkf_device_count(void);

/* ---- model: my_gmm::loadGaussian for all K components (src/my_gmm.cpp:45-75, src/pfPose.cpp:61-65) ----
 * means K x d, covs K x d x d (the reference's stacked (K*d) x d matrix), weights K, gamma K,
 * pca_proj d x D, pca_mean D (already widened to f64 as src/pfPose.cpp:44-51 does).  Host pointers.
 * Supported: d in {10, 12} with 6 measurement rows, K <= 64, 14 <= D <= 32. */
int mkf_model_create(mkf_model** out, int K, int d, int D, const double* means, const double* covs,
                     const double* weights, const double* gamma, const double* pca_proj, const double* pca_mean,
                     const mkf_params* params);
/* cv::FileStorage load of one arm model (src/pfPose.cpp:34-55): keys means, covs, weights,
 * pca_proj, pca_mean, gamma in OpenCV-YAML-1.0.  gamma_path: file to take `gamma` from, NULL =
 * same file.  (The reference reads BOTH arms' gamma from the right-arm file, src/pfPose.cpp:52-53,
 * quirk B4: pass the right-arm path here to reproduce it.) */
int mkf_model_load_yaml(mkf_model** out, const char* path, const char* gamma_path, const mkf_params* params);
/* the inverse: writes the model's GMM/PCA arrays in the same OpenCV-YAML-1.0 schema (the files the reference's
 * launch parameters left_arm_training / right_arm_training name, bodyTrackingBag.launch:2-6; produced upstream by
 * the gmm_training package, README.md:45-46).  f64 values round-trip bit-exactly; pca_proj / pca_mean keep `dt: f`
 * when they hold widened floats.  `gamma` is the one the model uses (see gamma_path above). */
int mkf_model_save_yaml(const mkf_model* m, const char* path);
void mkf_model_destroy(mkf_model* m);
int mkf_model_dims(const mkf_model* m, int* K, int* d, int* D);
/* copies of the model arrays (any pointer may be NULL): the loaded GMM and the derived
 * KF_model members Q (K x d x d), B (K x d), H (6 x d), BH (6) of src/my_gmm.cpp:53-72 */
int mkf_model_get(const mkf_model* m, double* means, double* covs, double* weights, double* gamma, double* pca_proj,
                  double* pca_mean, double* Q, double* B, double* H, double* BH);

/* ---- batch of tracks ---- */
/* replaces `new ParticleFilter(numParticles)` x T (src/pfPose.cpp:57-59).  stream: a cudaStream_t
 * to enqueue on (e.g. torch's current stream) or NULL for a private stream.  T * N <= 2^31 - 1 (32-bit slot
 * indices on the device).  The batch keeps a pointer to `m`: destroy the model after every batch made from it. */
int mkf_batch_create(mkf_batch** out, const mkf_model* m, int64_t T, int N, int device, void* stream);
void mkf_batch_destroy(mkf_batch* b);
int mkf_batch_sync(mkf_batch* b);
/* orders the batch's stream after every MKF_MEM_HOST_ASYNC copy issued so far (device-side waits, no host
 * synchronisation): work or events the caller enqueues on that stream afterwards see all results delivered */
int mkf_batch_join(mkf_batch* b);

/* bins = resample(gmm.weight, N); gmm.resetTracker(bins)  (src/pfPose.cpp:68-71, src/my_gmm.cpp:30-42).
 * u_init: T uniform draws in [0,1). */
int mkf_batch_reset(mkf_batch* b, const double* u_init, int mem);

/* ParticleFilter::update for every track (src/pf2DRao.cpp:125-158): indicator resample (K->N),
 * per-slot KF_model::predict (src/KF_model.cpp:11-15), innovation likelihood mvnpdf/chol
 * (src/pf2DRao.cpp:34-67), KF_model::update (src/KF_model.cpp:17-25), weight normalisation and
 * systematic resampling (src/pf2DRao.cpp:175-210).
 * meas: layout per meas_layout; u_ind, u_post: T draws each; seeds: T x 2 uint64 for the
 * degenerate fallback (NULL: seed 1 as cv::RNG(1)). */
int mkf_batch_update(mkf_batch* b, const double* meas, int meas_layout, const double* u_ind, const double* u_post,
                     const uint64_t* seeds, int mem);

/* ParticleFilter::getEstimator (src/pf2DRao.cpp:23-31) and e = pca_proj^T xbar + pca_mean^T
 * (src/pfPose.cpp:347-348).  xbar T x d, pose T x D; either may be NULL. */
int mkf_batch_estimate(mkf_batch* b, double* xbar, double* pose, int mem);

/* association ("PDAF") step of PFTracker::getMeasurementProposal for T persons
 * (src/pfPose.cpp:238-323; densities src/pf2DRao.cpp:69-83,105-122): gate, L/Z weights,
 * normalisation, C->N systematic resample and assembly of the per-slot measurement columns,
 * followed by ParticleFilter::update of both arms (src/pfPose.cpp:325-326) when do_update != 0.
 * arm[0]/arm[1]: the two arm batches (same T, N, device, stream).
 * cand_xy T x 2 hands x 2 (x row, y row) x C; cand_L T x 2 x C (likelihood-image samples,
 * uint8); roi T x 4 (x, y, w, h); u_cand T x 2; u_ind/u_post T x 2 (arm-major pairs);
 * seeds T x 2 x 3 uint64 or NULL: per arm [candidate, indicator (unused), posterior] fallback seeds. */
int mkf_batch_associate(mkf_batch* arm0, mkf_batch* arm1, int C, const double* cand_xy, const uint8_t* cand_L,
                        const double* roi, const double* u_cand, const double* u_ind, const double* u_post,
                        const uint64_t* seeds, int do_update, int mem);
/* results of the last mkf_batch_associate on arm0: gate T x 2 x C (0/1), weights T x 2 x C
 * (normalised), bins T x 2 x N; any may be NULL. */
int mkf_batch_assoc_results(mkf_batch* arm0, uint8_t* gate, double* weights, int32_t* bins, int mem);

/* state access in the reference's coordinates (parity tests, checkpoint/resume):
 * x T x N x d, P T x N x d x d (row-major full matrices, as gmm.tracks[j].state/.cov,
 * src/my_gmm.h:9-18), w_raw / w_norm T x N (weights of the last update before / after division
 * by wsum), indicators / parents T x N (component draw and resampled parent of the last
 * update), wsum T, status T.  Any pointer may be NULL. */
int mkf_batch_download(mkf_batch* b, double* x, double* P, double* w_raw, double* w_norm, int32_t* indicators,
                       int32_t* parents, double* wsum, uint32_t* status, int mem);
int mkf_batch_upload(mkf_batch* b, const double* x, const double* P, int mem);

/* KF_model::predict (stage 1, src/KF_model.cpp:11-15) and/or innovation likelihood + KF_model::update
 * (stage 2, src/pf2DRao.cpp:138 and src/KF_model.cpp:17-25; stage 3 = both) applied to n explicit
 * Gaussians with explicit component indices comp[n].  x n x d and P n x d x d in/out, z n x 6 (stage 2),
 * w_out n likelihoods or NULL.  Host pointers; synchronous.  Backs the KF_model shim. */
int mkf_kf_apply(const mkf_model* m, int n, const int32_t* comp, int stage, double* x, double* P, const double* z,
                 double* w_out, int device);

/* ParticleFilter::getSampleProb (src/pf2DRao.cpp:105-122) for one track of a batch: density of C
 * positions cand_xy (2 x C, row 0 = x) under N(posterior hand estimate, 0.8*scale*I).  Host pointers. */
int mkf_batch_sample_prob(mkf_batch* b, int64_t track, const double* cand_xy, int C, double scale, double* out);

/* ---- candidate generation front-end (the step before the path) ----
 * The proposal part of PFTracker::getMeasurementProposal (src/pfPose.cpp:216-236) with
 * ParticleFilter::getSamples (src/pf2DRao.cpp:85-103) and the likelihood-image lookup (:254) for T persons:
 * tracking[t] != 0 (NULL = all): C candidates per hand ~ N(posterior hand estimate, (0.8*roi_w)^2 I) (quirk B10);
 * tracking[t] == 0: uniform on the first-frame box around the face ROI.  Draws come from the counter generator
 * of mkf_synth.h keyed (seed, track0 + t, frame, hand, c) -- reproducible, unlike cv::randn/randu.
 * like: n_img (1 or T) likelihood images of img_rows x img_cols uint8 (already blurred by the caller,
 * src/pfPose.cpp:213), may be NULL with cand_L NULL.  Outputs in the layout mkf_batch_associate consumes:
 * cand_xy T x 2 x 2 x C, cand_L T x 2 x C. */
int mkf_batch_propose(mkf_batch* arm0, mkf_batch* arm1, int C, const double* roi, const uint8_t* tracking,
                      const uint8_t* like, int n_img, uint64_t seed, uint64_t frame, int64_t track0, double* cand_xy,
                      uint8_t* cand_L, int mem);

/* ---- output back-end (the step after the path) ----
 * PFTracker::get3Dpose (src/pfPose.cpp:93-127) of every track's current estimate: pos3d T x 3 x 5 (columns:
 * hand, elbow, shoulder, head, neck).  Kcam: 3 x 3 camera matrix (host, row-major) or NULL for the Kinect
 * literal the reference hard-codes (src/pfPose.cpp:101); mkf_load_camera_matrix reads a cal.yml. */
int mkf_batch_pose3d(mkf_batch* b, const double* Kcam, double* pos3d, int mem);
/* PFTracker::publishTFtree + publish2Dpos (src/pfPose.cpp:129-208) for T persons from both arms' estimates
 * (arm0 = e1, arm1 = e2): tf T x 10 x 3 = the nine broadcast translations in source order followed by the
 * camera Euler triple; joints2d T x 8 x 2.  Either output may be NULL. */
int mkf_batch_skeleton(mkf_batch* arm0, mkf_batch* arm1, const double* Kcam, double* tf, double* joints2d, int mem);
/* camera_matrix (3 x 3) of a ROS camera_calibration YAML such as the reference's cal.yml:4-7 */
int mkf_load_camera_matrix(const char* path, double* K9);

/* ParticleFilter::resample for one weight vector (src/pf2DRao.cpp:175-210) on the device:
 * w[L] (host), N outputs, u < 0 draws from cv::RNG(seed) as the reference does. */
int mkf_resample(const double* w, int L, int N, double u, uint64_t seed, int32_t* out, int device);

/* ---- legacy plain particle filter (src/pf2D.{h,cpp}, "particle likelihoods") ---- */
typedef struct mkf_pf2d mkf_pf2d;
/* T independent filters of N particles over d >= 8 dims with a K-component GMM prior
 * (my_gmm::loadGaussian, src/pf2D.cpp:28-37: means K x d, covs K x d x d, weights K). */
int mkf_pf2d_create(mkf_pf2d** out, int64_t T, int N, int d, int K, const double* means, const double* covs,
                    const double* weights, int device, void* stream);
void mkf_pf2d_destroy(mkf_pf2d* p);
int mkf_pf2d_set_particles(mkf_pf2d* p, const double* particles /* T x N x d */, int mem);
/* The particle randomisation of the constructor ParticleFilter(numParticles, numDims, side1) (src/pf2D.cpp:44-71) and of
 * resample()'s degenerate branch (max weight 0: every particle re-drawn across the image, weights back to 1/N,
 * src/pf2D.cpp:232-250).  The reference draws with cv::randu from the global cv::theRNG(), which the class interface
 * cannot seed; here the draws come from the counter generator of mkf_synth.h keyed (seed, track0 + t, epoch, particle,
 * dim), epoch 0 = constructor, epoch n = the n-th mkf_pf2d_update.  Column `dim` is uniform on [1, im_width) (even dim)
 * or [1, im_height) (odd dim); column 6 on [im_width/2*side + 1, im_width/2 + im_width/2*side).
 * mkf_pf2d_set_random fixes the parameters (side: T flags, host or device, NULL = all 0; defaults without the call:
 * seed 0, track0 0, side 0, 640 x 480); mkf_pf2d_randomise performs the constructor's draw (and the 1/N weights). */
int mkf_pf2d_set_random(mkf_pf2d* p, uint64_t seed, int64_t track0, const uint8_t* side, int im_width, int im_height);
int mkf_pf2d_randomise(mkf_pf2d* p);
int mkf_pf2d_get(mkf_pf2d* p, double* particles, double* w_norm, int32_t* parents, int mem);
/* ParticleFilter::update of src/pf2D.cpp:148-210: weights (GMM prior with float expf x two
 * isotropic 2-D likelihoods), normalise, systematic resample, random-walk predict.
 * meas T x 2 x 2, u T, noise T x N x d standard normals (NULL: no predict noise).  expf is glibc's algorithm, restated
 * in include/mkf_expf.h and shared with the oracle, so weights and resampled indices agree with the CPU path bit for bit.
 * A filter whose weights are all 0 takes the degenerate branch (status bit MKF_ST_POST_DEGENERATE; parents[i] = i). */
int mkf_pf2d_update(mkf_pf2d* p, const double* meas, const double* u, const double* noise, int mem);
/* ParticleFilter::getEstimator of src/pf2D.cpp:79-88: est (T x d) = sum_i weights[i] * particles.row(i), the weights
 * being the normalised ones the last update computed (resample() does not reset them, src/pf2D.cpp:225-268) and the
 * particles the resampled + predicted ones; before the first update, or after a degenerate one, weights are 1/N.
 * Device outputs are ordered on the filter's stream; host outputs are complete on return. */
int mkf_pf2d_estimate(mkf_pf2d* p, double* est, int mem);
int mkf_pf2d_sync(mkf_pf2d* p);
/* per-kernel device timing of mkf_pf2d_update (CUDA events on the filter's stream): arm for up to max_updates updates
 * (0 disables); read returns the summed milliseconds of ms[3] = {weights, normalise + resample, gather + predict} */
int mkf_pf2d_profile(mkf_pf2d* p, int max_updates);
int mkf_pf2d_profile_read(mkf_pf2d* p, double* ms, int* n_updates);

/* ---- multi-GPU: tracks shard over the GPUs of one box, one final gather of per-track summaries ----
 * The reference is one single-threaded process (src/pfPoseTracker.cpp:5-14) tracking one person; independent persons /
 * tracks have no coupling, so a batch per GPU needs no per-frame exchange.  The only collective is the
 * gather below: every rank contributes one row {pose[D], wsum, (double)status} per track and receives all ranks' rows
 * (ncclAllGather over NVLink / NVSwitch, enqueued on the batch's stream).  NCCL is bound at run time (dlopen of
 * libnccl.so.2 -- a copy already loaded into the process, e.g. torch's, is shared); without it these calls return
 * MKF_E_UNSUPPORTED and everything else works. */
typedef struct mkf_comm mkf_comm;
#define MKF_COMM_ID_BYTES 128
/* contiguous block partition of `total` tracks: rank r filters tracks [first, first + count) */
int mkf_shard_tracks(int64_t total, int world, int rank, int64_t* first, int64_t* count);
/* rank 0 makes an id (ncclGetUniqueId) and hands its 128 bytes to the other ranks by any host channel; every rank then
 * creates its communicator (ncclCommInitRank) for the device its batch lives on.  mkf_comm_wrap adopts an existing
 * ncclComm_t instead (not destroyed by mkf_comm_destroy). */
int mkf_comm_unique_id(void* id128);
int mkf_comm_create(mkf_comm** out, int nranks, int rank, const void* id128, int device);
int mkf_comm_wrap(mkf_comm** out, void* nccl_comm, int device);
void mkf_comm_destroy(mkf_comm* c);
int mkf_comm_info(const mkf_comm* c, int* nranks, int* rank, int* nccl_version);
/* this rank's summary rows: rows x (D + 2) doubles, rows >= T (rows beyond T are zero: padding for ragged shards).
 * Runs the estimator (src/pf2DRao.cpp:23-31, src/pfPose.cpp:347-348) for the pose columns. */
int mkf_batch_summaries(mkf_batch* b, int64_t rows, double* out, int mem);
/* the gather: out = nranks x rows_per_rank x (D + 2) doubles in rank order on every rank (rows_per_rank = the largest
 * shard; 0 = this batch's T when all shards are equal).  Device `out`: asynchronous on the batch's stream, packed and
 * gathered in place.  Host `out`: complete on return. */
int mkf_batch_gather_summaries(mkf_batch* b, mkf_comm* c, int64_t rows_per_rank, double* out, int mem);

/* ---- synthetic workload (include/mkf_synth.h), generated on the device ---- */
/* fills meas (device, layout per meas_layout) and u_ind/u_post (device, T each) for `frame` */
int mkf_synth_fill(mkf_batch* b, uint64_t seed, int64_t track0, uint64_t frame, int jitter, int meas_layout,
                   double* meas_dev, double* u_ind_dev, double* u_post_dev);

/* per-kernel device timing of mkf_batch_update, CUDA events on the batch's stream (bench.py roofline):
 * mkf_batch_profile arms recording for up to max_updates updates (0 disables);
 * mkf_batch_profile_read synchronises and returns the summed milliseconds of the three stages
 * (indicator bounds, fused slot update, normalise+resample) over n_updates updates, then rearms. */
int mkf_batch_profile(mkf_batch* b, int max_updates);
/* the same, sampling every `every`-th update (>= 1) for up to max_samples samples: the four event records of a
 * sampled update cost ~3 us each and suspend the kernels' programmatic overlap, so a timed region samples sparsely */
int mkf_batch_profile_every(mkf_batch* b, int max_samples, int every);
int mkf_batch_profile_read(mkf_batch* b, double* ms_bounds, double* ms_slot_update, double* ms_resample,
                           int* n_updates);
/* the same per kernel: ms[5] = indicator bounds, record-sharing keys (0 when the frame has none), the slot kernel
 * (k_slot_update / k_slot_update_heads_direct / k_slot_update_shared), the repair pass, normalise+resample */
int mkf_batch_profile_read_stages(mkf_batch* b, double* ms, int* n_updates);
/* the slot kernel's own span on the device over the same sampled updates: earliest CTA start to latest CTA end in
 * %globaltimer time, summed (ms) -- what the kernel takes without the launch gap an event-bracketed kernel pays when the
 * event records suspend the programmatic overlap.  Call before mkf_batch_profile_read_stages (which rearms sampling). */
int mkf_batch_profile_read_slot_span(mkf_batch* b, double* ms, int* n_updates);

/* Record sharing.  When the slots of a track see one measurement (MKF_MEAS_SHARED, MKF_ALIAS_INDEPENDENT), children
 * that drew the same parent and the same component are bit-identical Gaussians; the device computes and stores such a
 * group once and lets its slots refer to the one record.  The same holds inside mkf_batch_associate(do_update != 0)
 * for slots that also drew the same candidate (few candidates per hand, src/pfPose.cpp:300-323).  Nothing observable changes (per-slot weights,
 * parents, states and estimates are those of N independent slots); this call reports how many distinct records the
 * last update stored for how many slots. */
int mkf_batch_shared_records(mkf_batch* b, int64_t* records, int64_t* slots);

/* Name of the kernel that ran the distinct Gaussians ("heads") of the last run-length frame -- k_slot_update_heads_tma<d,
 * consumer warps, output stages, producer warps> (contiguous records staged through shared memory by TMA bulk copies;
 * the default) or k_slot_update_heads_direct<d> (MKF_HEADS_TMA=0, or a model whose constants leave no room for the
 * stages) -- empty before the first such frame.  For bench.py's roofline label. */
int mkf_batch_heads_kernel(mkf_batch* b, char* name, int len);

/* number of kernel launches issued through this library since load (bench.py "gpu_launches") */
uint64_t mkf_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* MKF_B200_H */
