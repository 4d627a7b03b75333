/* mkf_synth.h -- counter-based synthetic head/hands measurement generator.
 *
 * Shared verbatim by the CUDA kernels (nvcc, device side), the C-ABI host code and
 * the CPU oracle (gcc), so that "identical inputs and identical uniform draws"
 * (BASELINE.json north_star) holds by construction at any batch size without
 * storing the inputs.  Everything is integer hashing plus a handful of IEEE double
 * add/sub/mul that are never contracted into FMAs (explicit _rn intrinsics on the
 * device, -ffp-contract=off on the host) and no libm call, hence bit-identical on
 * CPU and GPU.
 *
 * The scenario follows SURVEY.md section 8(d): a fixed face ROI (300,51,47,47) as
 * facetracking would publish it (consumed by the reference at src/pfPose.cpp:308-313)
 * and a hand moving on a Lissajous-like curve around the PCA mean hand position
 * (data13D_PCA_100000_15_12.yml pca_mean = (388.0, 280.9)).
 */
#ifndef MKF_SYNTH_H
#define MKF_SYNTH_H

#include <stdint.h>

#if defined(__CUDACC__)
#define MKF_HD __host__ __device__ __forceinline__
#else
#define MKF_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define MKF_SMUL(a, b) __dmul_rn((a), (b))
#define MKF_SADD(a, b) __dadd_rn((a), (b))
#define MKF_SSUB(a, b) __dsub_rn((a), (b))
#define MKF_SFLOOR(a) floor(a)
#else
#include <math.h>
#define MKF_SMUL(a, b) ((a) * (b))
#define MKF_SADD(a, b) ((a) + (b))
#define MKF_SSUB(a, b) ((a) - (b))
#define MKF_SFLOOR(a) floor(a)
#endif

/* stream ids ("lane" of the counter) */
#define MKF_SYNTH_LANE_U_IND 0x1001u
#define MKF_SYNTH_LANE_U_POST 0x1002u
#define MKF_SYNTH_LANE_U_INIT 0x1003u
#define MKF_SYNTH_LANE_U_CAND 0x1004u /* +hand */
#define MKF_SYNTH_LANE_TRACK 0x2000u  /* +0..3 per-track jitter */
#define MKF_SYNTH_LANE_SHARED 0x3000u /* +0,1 shared hand noise */
#define MKF_SYNTH_LANE_SLOT 0x100000u /* +2*slot, +2*slot+1 per-slot hand noise */
#define MKF_SYNTH_LANE_CAND 0x40000000u /* + (hand*C + c)*4 + {0:x,1:y,2:L,3:kind} */
#define MKF_SYNTH_NO_FRAME 0xFFFFFFFFFFFFull

MKF_HD uint64_t mkf_mix64(uint64_t z)
{
    z ^= z >> 30;
    z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27;
    z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}

/* hash of the 4-word counter (seed, track, frame, lane) */
MKF_HD uint64_t mkf_hash4(uint64_t seed, uint64_t track, uint64_t frame, uint64_t lane)
{
    uint64_t h = mkf_mix64(seed + 0x9E3779B97F4A7C15ull);
    h = mkf_mix64(h ^ (track * 0xD6E8FEB86659FD93ull + 0x2545F4914F6CDD1Dull));
    h = mkf_mix64(h ^ (frame * 0xA0761D6478BD642Full + 0xE7037ED1A0B428DBull));
    h = mkf_mix64(h ^ (lane * 0x8EBC6AF09C88C6E3ull + 0x589965CC75374CC3ull));
    return h;
}

/* uniform in [0,1) with 53 random bits (exact) */
MKF_HD double mkf_u01(uint64_t h) { return (double)(h >> 11) * 1.1102230246251565404e-16; }

/* approximately N(0,1): Irwin-Hall sum of eight 16-bit uniforms, variance-normalised.
 * All intermediate values are small integers (exact); one final rounding. */
MKF_HD double mkf_gauss(uint64_t h)
{
    uint64_t h2 = mkf_mix64(h ^ 0xC2B2AE3D27D4EB4Full);
    uint32_t s = 0;
    s += (uint32_t)(h & 0xFFFF) + (uint32_t)((h >> 16) & 0xFFFF) + (uint32_t)((h >> 32) & 0xFFFF) +
         (uint32_t)((h >> 48) & 0xFFFF);
    s += (uint32_t)(h2 & 0xFFFF) + (uint32_t)((h2 >> 16) & 0xFFFF) + (uint32_t)((h2 >> 32) & 0xFFFF) +
         (uint32_t)((h2 >> 48) & 0xFFFF);
    /* mean of the sum is 8*(65535/2) = 262140; sd = 65536*sqrt(8/12) (to 1e-5) */
    double c = (double)((int32_t)s - 262140);
    return MKF_SMUL(c, 1.8688208922269842e-05); /* 1/(65536*sqrt(2/3)) */
}

/* sine-like wave of period 1 in p, piecewise parabolic, amplitude 1 */
MKF_HD double mkf_sinlike(double p)
{
    double f = MKF_SSUB(p, MKF_SFLOOR(p));
    if (f < 0.5) {
        return MKF_SMUL(MKF_SMUL(16.0, f), MKF_SSUB(0.5, f));
    }
    double g = MKF_SSUB(f, 0.5);
    return -MKF_SMUL(MKF_SMUL(16.0, g), MKF_SSUB(0.5, g));
}

/* true hand position of a track at a frame.  jitter==0 reproduces config 1 exactly
 * for every track; jitter!=0 draws a per-track phase/amplitude (configs 2..5). */
MKF_HD void mkf_synth_hand_truth(uint64_t seed, uint64_t track, uint64_t frame, int jitter, double* hx,
                                 double* hy)
{
    double phx = 0.0, phy = 0.16666666666666666, ax = 60.0, ay = 70.0;
    if (jitter) {
        phx = mkf_u01(mkf_hash4(seed, track, MKF_SYNTH_NO_FRAME, MKF_SYNTH_LANE_TRACK + 0));
        phy = mkf_u01(mkf_hash4(seed, track, MKF_SYNTH_NO_FRAME, MKF_SYNTH_LANE_TRACK + 1));
        ax = MKF_SADD(40.0, MKF_SMUL(40.0, mkf_u01(mkf_hash4(seed, track, MKF_SYNTH_NO_FRAME,
                                                              MKF_SYNTH_LANE_TRACK + 2))));
        ay = MKF_SADD(50.0, MKF_SMUL(40.0, mkf_u01(mkf_hash4(seed, track, MKF_SYNTH_NO_FRAME,
                                                              MKF_SYNTH_LANE_TRACK + 3))));
    }
    double t = (double)frame;
    double px = MKF_SADD(MKF_SMUL(t, 0.013333333333333334), phx); /* t/75 */
    double py = MKF_SADD(MKF_SMUL(t, 0.02), phy);                 /* t/50 */
    *hx = MKF_SADD(388.0, MKF_SMUL(ax, mkf_sinlike(px)));
    *hy = MKF_SADD(250.0, MKF_SMUL(ay, mkf_sinlike(py)));
}

/* measurement column in the reference's order [head_x, head_y, hand_x, hand_y, neck_x, neck_y]
 * (src/my_gmm.cpp:62-67, src/pfPose.cpp:308-313) for ROI (300,51,47,47).
 * slot < 0: one shared column per track-frame (config 2/5); slot >= 0: per-slot column (config 1/4). */
MKF_HD void mkf_synth_meas(uint64_t seed, uint64_t track, uint64_t frame, int64_t slot, int jitter,
                           double z[6])
{
    double hx, hy;
    mkf_synth_hand_truth(seed, track, frame, jitter, &hx, &hy);
    uint64_t l0 = slot < 0 ? (uint64_t)MKF_SYNTH_LANE_SHARED : (uint64_t)MKF_SYNTH_LANE_SLOT + 2ull * (uint64_t)slot;
    double nx = mkf_gauss(mkf_hash4(seed, track, frame, l0));
    double ny = mkf_gauss(mkf_hash4(seed, track, frame, l0 + 1));
    z[0] = 323.5;  /* roi.x + w/2   */
    z[1] = 74.5;   /* roi.y + 0.5 h */
    z[2] = MKF_SADD(hx, MKF_SMUL(3.0, nx));
    z[3] = MKF_SADD(hy, MKF_SMUL(3.0, ny));
    z[4] = 323.5;  /* roi.x + w/2   */
    z[5] = 128.55; /* roi.y + 1.65 h */
}

/* uniform draw for a resampling call (which = MKF_SYNTH_LANE_U_*) */
MKF_HD double mkf_synth_u(uint64_t seed, uint64_t track, uint64_t frame, uint32_t which)
{
    return mkf_u01(mkf_hash4(seed, track, frame, which));
}

/* association candidates (config 3): candidate 0 of each hand is the detection
 * (truth + N(0,3^2), L in {200..255}); the others are clutter, uniform on
 * [-32,672) x [-24,504) with L = 0 w.p. 1/2 else {1..128}. */
MKF_HD void mkf_synth_candidate(uint64_t seed, uint64_t track, uint64_t frame, int hand, int C, int c,
                                int jitter, double* cx, double* cy, uint8_t* L)
{
    uint64_t base = (uint64_t)MKF_SYNTH_LANE_CAND + 4ull * ((uint64_t)hand * (uint64_t)C + (uint64_t)c);
    uint64_t hxh = mkf_hash4(seed, track, frame, base + 0);
    uint64_t hyh = mkf_hash4(seed, track, frame, base + 1);
    uint64_t hl = mkf_hash4(seed, track, frame, base + 2);
    if (c == 0) {
        double hx, hy;
        /* the two hands of a person follow mirrored curves: hand 1 is offset by -140 px */
        mkf_synth_hand_truth(seed, track, frame, jitter, &hx, &hy);
        if (hand == 1) hx = MKF_SSUB(hx, 140.0);
        *cx = MKF_SADD(hx, MKF_SMUL(3.0, mkf_gauss(hxh)));
        *cy = MKF_SADD(hy, MKF_SMUL(3.0, mkf_gauss(hyh)));
        *L = (uint8_t)(200u + (uint32_t)(hl % 56u));
    } else {
        *cx = MKF_SADD(-32.0, MKF_SMUL(704.0, mkf_u01(hxh)));
        *cy = MKF_SADD(-24.0, MKF_SMUL(528.0, mkf_u01(hyh)));
        *L = (hl & 1u) ? (uint8_t)(1u + (uint32_t)((hl >> 8) % 128u)) : (uint8_t)0;
    }
}

/* Candidate proposal front-end (src/pfPose.cpp:216-236, src/pf2DRao.cpp:85-103) driven by the counter
 * generator instead of cv::randn / cv::randu (whose streams are not reproducible):
 *  tracking != 0: x = hx + (spread*scale) * n1, y = hy + (spread*scale) * n2 -- cv::randn is handed
 *                 C = 0.8*scale*I as a STANDARD-DEVIATION matrix (quirk B10), scale = roi width;
 *  tracking == 0: first frame after (re)acquiring the face: uniform on the box
 *                 [max(x-4w,0), min(x+5w,cols)) x [min(y+h,rows), min(y+7h,rows))  (integer-truncated). */
#define MKF_SYNTH_LANE_PROP 0x50000000u /* + (hand*C + c)*2 + {0:x,1:y} */
MKF_HD void mkf_synth_proposal(uint64_t seed, uint64_t track, uint64_t frame, int hand, int C, int c, int tracking,
                               double hx, double hy, const double roi[4], int rows, int cols, double spread,
                               double* px, double* py)
{
    const uint64_t lane = (uint64_t)MKF_SYNTH_LANE_PROP + 2ull * ((uint64_t)hand * (uint64_t)C + (uint64_t)c);
    const uint64_t h0 = mkf_hash4(seed, track, frame, lane), h1 = mkf_hash4(seed, track, frame, lane + 1);
    if (tracking) {
        const double sd = MKF_SMUL(MKF_SMUL(spread, roi[2]), 1.0);
        *px = MKF_SADD(hx, MKF_SMUL(sd, mkf_gauss(h0)));
        *py = MKF_SADD(hy, MKF_SMUL(sd, mkf_gauss(h1)));
    } else {
        int xmin = (int)roi[0] - (int)(4 * roi[2]);
        if (xmin < 0) xmin = 0;
        int xmax = (int)roi[0] + (int)(5 * roi[2]);
        if (xmax > cols) xmax = cols;
        int ymin = (int)roi[1] + (int)(roi[3]);
        if (ymin > rows) ymin = rows;
        int ymax = (int)roi[1] + (int)(7 * roi[3]);
        if (ymax > rows) ymax = rows;
        *px = MKF_SADD((double)xmin, MKF_SMUL(mkf_u01(h0), (double)(xmax - xmin)));
        *py = MKF_SADD((double)ymin, MKF_SMUL(mkf_u01(h1), (double)(ymax - ymin)));
    }
}

/* Legacy pf2D particle randomisation (constructor, src/pf2D.cpp:58-70, and the degenerate branch of resample(),
 * src/pf2D.cpp:232-250): column `dim` of the N x d particle matrix is cv::randu on [1, im_width) for even dim,
 * [1, im_height) for odd dim, and [im_width/2*side + 1, im_width/2 + im_width/2*side) for dim 6.  cv::randu draws
 * from the global cv::theRNG(), which is not reproducible through the class interface, so the draw comes from the
 * counter generator keyed (seed, track, epoch, particle, dim): epoch 0 = constructor, epoch n = the n-th update.
 * Value = lo + u * (hi - lo), the affine map cv::randu applies to its unit draw. */
#define MKF_SYNTH_LANE_PF2D 0x60000000u /* + particle*16 + dim */
MKF_HD double mkf_synth_pf2d_uniform(uint64_t seed, uint64_t track, uint64_t epoch, int particle, int dim, int side,
                                     int im_width, int im_height)
{
    double lo = 1.0, hi;
    const double half = MKF_SMUL((double)im_width, 0.5); /* im_width/2.0 */
    if (dim == 6) {
        lo = MKF_SADD(MKF_SMUL(half, (double)(side ? 1 : 0)), 1.0);
        hi = MKF_SADD(half, MKF_SMUL(half, (double)(side ? 1 : 0)));
    } else {
        hi = (dim % 2 == 0) ? (double)im_width : (double)im_height;
    }
    const double u = mkf_u01(mkf_hash4(seed, track, epoch,
                                       (uint64_t)MKF_SYNTH_LANE_PF2D + 16ull * (uint64_t)particle + (uint64_t)dim));
    return MKF_SADD(lo, MKF_SMUL(u, MKF_SSUB(hi, lo)));
}

/* likelihood.at<uchar>(y, x) with the implicit double -> int truncation of src/pfPose.cpp:254, guarded by the
 * inside-image test of :251 (outside candidates never reach the lookup; they get L = 0 here) */
MKF_HD uint8_t mkf_likelihood_lookup(const uint8_t* img, int rows, int cols, double x, double y)
{
    if (!((y > 0) && (y < (double)rows) && (x > 0) && (x < (double)cols))) return 0;
    return img[(long long)(int)y * cols + (int)x];
}

#endif /* MKF_SYNTH_H */
