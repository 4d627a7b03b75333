// mkf_internal.h -- host-side structures shared by the C-ABI translation units.
#ifndef MKF_INTERNAL_H
#define MKF_INTERNAL_H

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/mkf_b200.h"

#define MKF_M 6 /* measurement rows: head x,y, hand x,y, neck x,y (src/my_gmm.cpp:61-67) */

// Packed per-slot layout in the measurement-aligned basis x' = T x, T = [H; N]:
//   [ x'(d) | A = P'11 lower-packed (21) | B = P'21 row-major (d2 x 6) | C = P'22 lower-packed ]
struct mkf_layout {
    int d, d2, na, nb, nc, ne, np; // ne = doubles per slot, np = double2 pairs per slot
    int ox, oa, ob, oc;
    int cs; // doubles per component constant block: g, g^2, b'(d), Q' packed (ne - d), padded even
};
inline mkf_layout mkf_make_layout(int d)
{
    mkf_layout L;
    L.d = d;
    L.d2 = d - MKF_M;
    L.na = MKF_M * (MKF_M + 1) / 2;
    L.nb = L.d2 * MKF_M;
    L.nc = L.d2 * (L.d2 + 1) / 2;
    L.ne = d + L.na + L.nb + L.nc;
    L.np = (L.ne + 1) / 2;
    L.ox = 0;
    L.oa = d;
    L.ob = d + L.na;
    L.oc = d + L.na + L.nb;
    L.cs = 2 + L.ne;
    if (L.cs & 1) L.cs++;
    return L;
}

struct mkf_model {
    int K, d, D;
    mkf_params prm;
    mkf_layout lay;
    // as loaded (src/pfPose.cpp:34-55)
    std::vector<double> means, covs, weights, gamma, proj, pmean;
    // KF_model members derived exactly as src/my_gmm.cpp:53-72 (reference coordinates)
    std::vector<double> Q, B, H, BH;
    // measurement-aligned basis
    std::vector<double> Tm, Tinv; // d x d
    std::vector<double> comp_const; // K x cs : g, g^2, b', Q' packed
    std::vector<double> init_const; // K x ne : mu'_k, Sigma'_k packed (resetTracker)
    std::vector<double> cw_hi, cw_lo; // K : double-double inclusive prefix sums of the prior weights
    double prior_wmax;
    std::vector<double> recon; // D x d : pca_proj^T * Tinv
};

void mkf_set_error(const char* fmt, ...);
int mkf_model_finalize(mkf_model* m); // derive everything from the loaded arrays

// OpenCV-YAML-1.0 matrix reader (src/pfPose.cpp:34-55 uses cv::FileStorage)
struct mkf_yaml_mat {
    int rows = 0, cols = 0;
    char dt = 'd';
    std::vector<double> data; // f32 entries already widened as (double)(float)value
};
int mkf_yaml_read(const char* path, const char* key, mkf_yaml_mat* out);

#endif
