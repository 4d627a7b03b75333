// mkf_model.cpp -- host side of the arm model: loading (OpenCV-YAML-1.0), the KF_model
// constants of my_gmm::loadGaussian (src/my_gmm.cpp:45-75) and their image in the
// measurement-aligned basis used by the kernels ("basis").
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

#include "mkf_internal.h"

static thread_local char g_err[512] = "";

void mkf_set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* mkf_last_error(void) { return g_err; }
extern "C" int mkf_abi_version(void) { return MKF_ABI_VERSION; }

extern "C" void mkf_params_default(mkf_params* p)
{
    if (!p) return;
    p->chol_mode = MKF_CHOL_CV24_LITERAL;
    p->alias_mode = MKF_ALIAS_INDEPENDENT;
    p->meas_noise_var = 100.0;  // src/my_gmm.cpp:54
    p->assoc_pa = 0.05;         // src/pfPose.cpp:247
    p->assoc_clutter = 1e-4;    // src/pfPose.cpp:261
    p->proposal_spread = 0.8;   // src/pf2DRao.cpp:90,114
    p->neck_offset = 1.65;      // src/pfPose.cpp:313
    p->img_rows = 480;          // 640 x 480 likelihood image (cal.yml:1-2)
    p->img_cols = 640;
}

// ------------------------------------------------------------------------------------------
// OpenCV-YAML-1.0 "!!opencv-matrix" reader
// ------------------------------------------------------------------------------------------
int mkf_yaml_read(const char* path, const char* key, mkf_yaml_mat* out)
{
    std::ifstream f(path);
    if (!f) {
        mkf_set_error("cannot open model file '%s'", path);
        return MKF_E_IO;
    }
    std::stringstream ss;
    ss << f.rdbuf();
    const std::string txt = ss.str();
    // find "<key>:" at the beginning of a line
    const std::string pat = std::string(key) + ":";
    size_t pos = 0;
    for (;;) {
        pos = txt.find(pat, pos);
        if (pos == std::string::npos) {
            mkf_set_error("key '%s' not found in '%s'", key, path);
            return MKF_E_PARSE;
        }
        if (pos == 0 || txt[pos - 1] == '\n') break;
        pos += pat.size();
    }
    size_t end = txt.find(']', pos);
    if (end == std::string::npos) {
        mkf_set_error("key '%s' in '%s': missing ']'", key, path);
        return MKF_E_PARSE;
    }
    const std::string blk = txt.substr(pos, end - pos + 1);
    if (blk.find("!!opencv-matrix") == std::string::npos) {
        mkf_set_error("key '%s' in '%s' is not an !!opencv-matrix", key, path);
        return MKF_E_PARSE;
    }
    auto field = [&](const char* name, std::string* val) -> bool {
        size_t p = blk.find(name);
        if (p == std::string::npos) return false;
        p += strlen(name);
        size_t e = blk.find('\n', p);
        *val = blk.substr(p, e == std::string::npos ? std::string::npos : e - p);
        return true;
    };
    std::string v;
    if (!field("rows:", &v)) goto bad;
    out->rows = atoi(v.c_str());
    if (!field("cols:", &v)) goto bad;
    out->cols = atoi(v.c_str());
    if (!field("dt:", &v)) goto bad;
    {
        size_t q = v.find_first_not_of(" \t\"");
        if (q == std::string::npos) goto bad;
        out->dt = v[q];
    }
    if (out->rows <= 0 || out->cols <= 0 || (out->dt != 'd' && out->dt != 'f')) goto bad;
    {
        size_t p = blk.find("data:");
        if (p == std::string::npos) goto bad;
        p = blk.find('[', p);
        if (p == std::string::npos) goto bad;
        const char* c = blk.c_str() + p + 1;
        out->data.clear();
        out->data.reserve((size_t)out->rows * out->cols);
        for (;;) {
            while (*c == ' ' || *c == ',' || *c == '\n' || *c == '\r' || *c == '\t') c++;
            if (*c == ']' || *c == 0) break;
            char* e = nullptr;
            double val = strtod(c, &e);
            if (e == c) goto bad;
            // cv::FileStorage parses the text as double and stores dt 'f' as float;
            // src/pfPose.cpp:44-51 then widens with convertTo(CV_64F)
            if (out->dt == 'f') val = (double)(float)val;
            out->data.push_back(val);
            c = e;
        }
        if (out->data.size() != (size_t)out->rows * out->cols) {
            mkf_set_error("key '%s' in '%s': %zu values for a %d x %d matrix", key, path, out->data.size(),
                          out->rows, out->cols);
            return MKF_E_PARSE;
        }
    }
    return MKF_OK;
bad:
    mkf_set_error("key '%s' in '%s': malformed !!opencv-matrix block", key, path);
    return MKF_E_PARSE;
}

// ------------------------------------------------------------------------------------------
// derive constants
// ------------------------------------------------------------------------------------------
typedef long double ld;

static bool invert_ld(int n, const std::vector<ld>& A, std::vector<ld>& inv)
{
    std::vector<ld> a(A);
    inv.assign((size_t)n * n, 0.0L);
    for (int i = 0; i < n; i++) inv[(size_t)i * n + i] = 1.0L;
    for (int c = 0; c < n; c++) {
        int p = c;
        for (int r = c + 1; r < n; r++)
            if (fabsl(a[(size_t)r * n + c]) > fabsl(a[(size_t)p * n + c])) p = r;
        if (fabsl(a[(size_t)p * n + c]) < 1e-300L) return false;
        if (p != c)
            for (int j = 0; j < n; j++) {
                std::swap(a[(size_t)p * n + j], a[(size_t)c * n + j]);
                std::swap(inv[(size_t)p * n + j], inv[(size_t)c * n + j]);
            }
        ld piv = 1.0L / a[(size_t)c * n + c];
        for (int j = 0; j < n; j++) {
            a[(size_t)c * n + j] *= piv;
            inv[(size_t)c * n + j] *= piv;
        }
        for (int r = 0; r < n; r++) {
            if (r == c) continue;
            ld f = a[(size_t)r * n + c];
            if (f == 0.0L) continue;
            for (int j = 0; j < n; j++) {
                a[(size_t)r * n + j] -= f * a[(size_t)c * n + j];
                inv[(size_t)r * n + j] -= f * inv[(size_t)c * n + j];
            }
        }
    }
    return true;
}

// packed position of P'(r, c), r >= c, in the [A | B | C] part of the slot layout (offset from oa)
static int packed_index(const mkf_layout& L, int r, int c)
{
    if (r < c) std::swap(r, c);
    if (r < MKF_M) return r * (r + 1) / 2 + c;                       // A
    if (c < MKF_M) return L.na + (r - MKF_M) * MKF_M + c;            // B[r-6][c]
    int i = r - MKF_M, j = c - MKF_M;
    return L.na + L.nb + i * (i + 1) / 2 + j;                        // C
}

int mkf_model_finalize(mkf_model* m)
{
    const int K = m->K, d = m->d, D = m->D;
    if (!(d == 10 || d == 12)) {
        mkf_set_error("unsupported PCA dimension d=%d (built for d in {10,12})", d);
        return MKF_E_UNSUPPORTED;
    }
    if (K < 1 || K > 64 || D < 14 || D > 32) {
        mkf_set_error("unsupported model shape K=%d D=%d (need 1<=K<=64, 14<=D<=32)", K, D);
        return MKF_E_INVALID;
    }
    if (m->prm.alias_mode != MKF_ALIAS_INDEPENDENT && m->prm.alias_mode != MKF_ALIAS_CV_SHALLOW_LITERAL) {
        mkf_set_error("invalid alias_mode %d", m->prm.alias_mode);
        return MKF_E_INVALID;
    }
    if (m->prm.chol_mode < 0 || m->prm.chol_mode > 2 || !(m->prm.meas_noise_var > 0)) {
        mkf_set_error("invalid mkf_params");
        return MKF_E_INVALID;
    }
    m->lay = mkf_make_layout(d);
    const mkf_layout& L = m->lay;
    // --- reference-coordinate constants, evaluated as src/my_gmm.cpp:53-72 does ---
    m->Q.resize((size_t)K * d * d);
    m->B.resize((size_t)K * d);
    m->H.assign((size_t)MKF_M * d, 0.0);
    m->BH.assign(MKF_M, 0.0);
    const int sel[MKF_M] = {9, 10, 0, 1, 12, 13}; // H1 ones (src/my_gmm.cpp:62-67)
    for (int r = 0; r < MKF_M; r++) {
        for (int c = 0; c < d; c++) m->H[(size_t)r * d + c] = m->proj[(size_t)c * D + sel[r]];
        m->BH[r] = m->pmean[sel[r]];
    }
    double wmax = 0;
    for (int k = 0; k < K; k++) {
        const double g = m->gamma[k];
        for (int e = 0; e < d * d; e++) m->Q[(size_t)k * d * d + e] = m->covs[(size_t)k * d * d + e] * (1 - g * g);
        for (int i = 0; i < d; i++) m->B[(size_t)k * d + i] = m->means[(size_t)k * d + i] * (1.0 - g);
        if (m->weights[k] > wmax) wmax = m->weights[k];
    }
    if (!(wmax > 0)) {
        mkf_set_error("GMM prior weights have no positive entry");
        return MKF_E_INVALID;
    }
    m->prior_wmax = wmax;
    // --- measurement-aligned basis T = [H; N], N an orthonormal basis of null(H) ---
    std::vector<ld> Tl((size_t)d * d, 0.0L), basis; // basis: orthonormalised rows so far
    basis.reserve((size_t)d * d);
    auto residual = [&](std::vector<ld>& v) {
        const int nb = (int)(basis.size() / d);
        for (int pass = 0; pass < 2; pass++)
            for (int b = 0; b < nb; b++) {
                ld dot = 0;
                for (int c = 0; c < d; c++) dot += basis[(size_t)b * d + c] * v[c];
                for (int c = 0; c < d; c++) v[c] -= dot * basis[(size_t)b * d + c];
            }
        ld n2 = 0;
        for (int c = 0; c < d; c++) n2 += v[c] * v[c];
        return sqrtl(n2);
    };
    for (int r = 0; r < MKF_M; r++) {
        std::vector<ld> v(d);
        for (int c = 0; c < d; c++) {
            v[c] = m->H[(size_t)r * d + c];
            Tl[(size_t)r * d + c] = v[c];
        }
        ld n = residual(v);
        if (!(n > 1e-9L)) {
            mkf_set_error("measurement matrix H is rank deficient (row %d)", r);
            return MKF_E_INVALID;
        }
        for (int c = 0; c < d; c++) basis.push_back(v[c] / n);
    }
    for (int r = MKF_M; r < d; r++) {
        int best = -1;
        ld bestn = -1;
        std::vector<ld> bestv;
        for (int e = 0; e < d; e++) {
            std::vector<ld> v(d, 0.0L);
            v[e] = 1.0L;
            ld n = residual(v);
            if (n > bestn) {
                bestn = n;
                best = e;
                bestv = v;
            }
        }
        (void)best;
        for (int c = 0; c < d; c++) {
            bestv[c] /= bestn;
            basis.push_back(bestv[c]);
            Tl[(size_t)r * d + c] = bestv[c];
        }
    }
    std::vector<ld> Til;
    if (!invert_ld(d, Tl, Til)) {
        mkf_set_error("basis matrix is singular");
        return MKF_E_INVALID;
    }
    m->Tm.resize((size_t)d * d);
    m->Tinv.resize((size_t)d * d);
    for (int e = 0; e < d * d; e++) {
        m->Tm[e] = (double)Tl[e];
        m->Tinv[e] = (double)Til[e];
    }
    // --- per-component constants in the new basis ---
    auto xform_vec = [&](const double* v, double* out) {
        for (int r = 0; r < d; r++) {
            ld s = 0;
            for (int c = 0; c < d; c++) s += Tl[(size_t)r * d + c] * (ld)v[c];
            out[r] = (double)s;
        }
    };
    auto xform_sym_packed = [&](const double* P, double* out) { // out: ne - d packed values
        std::vector<ld> TP((size_t)d * d), TPT((size_t)d * d);
        for (int r = 0; r < d; r++)
            for (int c = 0; c < d; c++) {
                ld s = 0;
                for (int k = 0; k < d; k++) s += Tl[(size_t)r * d + k] * (ld)P[(size_t)k * d + c];
                TP[(size_t)r * d + c] = s;
            }
        for (int r = 0; r < d; r++)
            for (int c = 0; c < d; c++) {
                ld s = 0;
                for (int k = 0; k < d; k++) s += TP[(size_t)r * d + k] * Tl[(size_t)c * d + k];
                TPT[(size_t)r * d + c] = s;
            }
        for (int r = 0; r < d; r++)
            for (int c = 0; c <= r; c++)
                out[packed_index(L, r, c)] = (double)(0.5L * (TPT[(size_t)r * d + c] + TPT[(size_t)c * d + r]));
    };
    m->comp_const.assign((size_t)K * L.cs, 0.0);
    m->init_const.assign((size_t)K * L.ne, 0.0);
    for (int k = 0; k < K; k++) {
        double* cc = &m->comp_const[(size_t)k * L.cs];
        const double g = m->gamma[k];
        cc[0] = g;
        cc[1] = g * g;
        xform_vec(&m->B[(size_t)k * d], cc + 2);
        xform_sym_packed(&m->Q[(size_t)k * d * d], cc + 2 + d);
        double* ic = &m->init_const[(size_t)k * L.ne];
        xform_vec(&m->means[(size_t)k * d], ic);
        xform_sym_packed(&m->covs[(size_t)k * d * d], ic + d);
    }
    // --- prior weights: double-double inclusive prefix sums (exact to ~1e-32) ---
    m->cw_hi.resize(K);
    m->cw_lo.resize(K);
    {
        ld acc = 0; // x87 80-bit: 64-bit mantissa; K <= 64 terms of 53 bits -> use error-free sums instead
        double hi = 0, lo = 0;
        for (int k = 0; k < K; k++) {
            // two_sum(hi, w) then fold lo
            double w = m->weights[k];
            double s = hi + w;
            double bb = s - hi;
            double e = (hi - (s - bb)) + (w - bb);
            e += lo;
            double s2 = s + e;
            lo = e - (s2 - s);
            hi = s2;
            m->cw_hi[k] = hi;
            m->cw_lo[k] = lo;
            acc += w;
        }
        (void)acc;
    }
    // --- pose reconstruction in the new basis: e = pca_proj^T (Tinv x') + pca_mean ---
    m->recon.resize((size_t)D * d);
    for (int r = 0; r < D; r++)
        for (int c = 0; c < d; c++) {
            ld s = 0;
            for (int k = 0; k < d; k++) s += (ld)m->proj[(size_t)k * D + r] * Til[(size_t)k * d + c];
            m->recon[(size_t)r * d + c] = (double)s;
        }
    return MKF_OK;
}

// ------------------------------------------------------------------------------------------
// C ABI: model
// ------------------------------------------------------------------------------------------
extern "C" int mkf_model_create(mkf_model** out, int K, int d, int D, const double* means, const double* covs,
                                const double* weights, const double* gamma, const double* pca_proj,
                                const double* pca_mean, const mkf_params* params)
{
    if (!out || !means || !covs || !weights || !gamma || !pca_proj || !pca_mean || K <= 0 || d <= 0 || D <= 0) {
        mkf_set_error("mkf_model_create: null or non-positive argument");
        return MKF_E_INVALID;
    }
    *out = nullptr;
    mkf_model* m = new (std::nothrow) mkf_model;
    if (!m) return MKF_E_NOMEM;
    m->K = K;
    m->d = d;
    m->D = D;
    if (params)
        m->prm = *params;
    else
        mkf_params_default(&m->prm);
    m->means.assign(means, means + (size_t)K * d);
    m->covs.assign(covs, covs + (size_t)K * d * d);
    m->weights.assign(weights, weights + K);
    m->gamma.assign(gamma, gamma + K);
    m->proj.assign(pca_proj, pca_proj + (size_t)d * D);
    m->pmean.assign(pca_mean, pca_mean + D);
    int rc = mkf_model_finalize(m);
    if (rc != MKF_OK) {
        delete m;
        return rc;
    }
    *out = m;
    return MKF_OK;
}

extern "C" int mkf_model_load_yaml(mkf_model** out, const char* path, const char* gamma_path,
                                   const mkf_params* params)
{
    if (!out || !path) {
        mkf_set_error("mkf_model_load_yaml: null argument");
        return MKF_E_INVALID;
    }
    *out = nullptr;
    mkf_yaml_mat means, covs, weights, proj, pmean, gamma;
    int rc;
    if ((rc = mkf_yaml_read(path, "means", &means))) return rc;
    if ((rc = mkf_yaml_read(path, "covs", &covs))) return rc;
    if ((rc = mkf_yaml_read(path, "weights", &weights))) return rc;
    if ((rc = mkf_yaml_read(path, "pca_proj", &proj))) return rc;
    if ((rc = mkf_yaml_read(path, "pca_mean", &pmean))) return rc;
    if ((rc = mkf_yaml_read(gamma_path ? gamma_path : path, "gamma", &gamma))) return rc;
    const int K = means.rows, d = means.cols, D = proj.cols;
    // the same shape relations the reference relies on (src/pfPose.cpp:61-65)
    if (covs.rows != K * d || covs.cols != d || weights.rows * weights.cols != K || gamma.rows * gamma.cols < K ||
        proj.rows != d || pmean.rows * pmean.cols != D) {
        mkf_set_error("model '%s': inconsistent matrix shapes (means %dx%d covs %dx%d weights %dx%d pca_proj %dx%d "
                      "pca_mean %dx%d gamma %dx%d)",
                      path, means.rows, means.cols, covs.rows, covs.cols, weights.rows, weights.cols, proj.rows,
                      proj.cols, pmean.rows, pmean.cols, gamma.rows, gamma.cols);
        return MKF_E_PARSE;
    }
    return mkf_model_create(out, K, d, D, means.data.data(), covs.data.data(), weights.data.data(), gamma.data.data(),
                            proj.data.data(), pmean.data.data(), params);
}

// ------------------------------------------------------------------------------------------
// OpenCV-YAML-1.0 writer: the schema src/pfPose.cpp:34-55 reads (and the gmm_training package of README.md:45-46
// emits): means K x d, covs (K*d) x d, weights 1 x K, pca_proj d x D, pca_mean 1 x D, gamma K x 1.
// Doubles are written with 17 significant digits (round-trip exact); pca_proj / pca_mean keep the reference
// files' `dt: f` when every entry is a float widened to double, else `dt: d` (cv::FileStorage reads either and
// the reference converts to CV_64F afterwards, src/pfPose.cpp:44-51).
// ------------------------------------------------------------------------------------------
static void yaml_write_mat(FILE* f, const char* key, int rows, int cols, const std::vector<double>& v, bool allow_f32)
{
    bool f32 = allow_f32;
    for (size_t i = 0; f32 && i < v.size(); ++i) f32 = ((double)(float)v[i] == v[i]);
    fprintf(f, "%s: !!opencv-matrix\n   rows: %d\n   cols: %d\n   dt: %c\n   data: [ ", key, rows, cols, f32 ? 'f' : 'd');
    int col = 11;
    for (size_t i = 0; i < v.size(); ++i) {
        char buf[40];
        int n = snprintf(buf, sizeof buf, f32 ? "%.9g" : "%.17g", v[i]);
        // cv::FileStorage wants a real-number token: make sure integers carry a '.'
        if (!strpbrk(buf, ".eEn")) {
            buf[n++] = '.';
            buf[n] = 0;
        }
        if (col + n + 2 > 100) {
            fputs("\n       ", f);
            col = 7;
        }
        fputs(buf, f);
        col += n;
        if (i + 1 < v.size()) {
            fputs(", ", f);
            col += 2;
        }
    }
    fputs(" ]\n", f);
}

extern "C" int mkf_model_save_yaml(const mkf_model* m, const char* path)
{
    if (!m || !path) {
        mkf_set_error("mkf_model_save_yaml: null argument");
        return MKF_E_INVALID;
    }
    FILE* f = fopen(path, "w");
    if (!f) {
        mkf_set_error("cannot write model file '%s'", path);
        return MKF_E_IO;
    }
    fputs("%YAML:1.0\n", f);
    yaml_write_mat(f, "means", m->K, m->d, m->means, false);
    yaml_write_mat(f, "covs", m->K * m->d, m->d, m->covs, false);
    yaml_write_mat(f, "weights", 1, m->K, m->weights, false);
    yaml_write_mat(f, "pca_proj", m->d, m->D, m->proj, true);
    yaml_write_mat(f, "pca_mean", 1, m->D, m->pmean, true);
    yaml_write_mat(f, "gamma", m->K, 1, m->gamma, false);
    const bool bad = ferror(f) != 0;
    if (fclose(f) != 0 || bad) {
        mkf_set_error("write error on model file '%s'", path);
        return MKF_E_IO;
    }
    return MKF_OK;
}

extern "C" void mkf_model_destroy(mkf_model* m) { delete m; }

extern "C" int mkf_model_dims(const mkf_model* m, int* K, int* d, int* D)
{
    if (!m) {
        mkf_set_error("null model");
        return MKF_E_INVALID;
    }
    if (K) *K = m->K;
    if (d) *d = m->d;
    if (D) *D = m->D;
    return MKF_OK;
}

extern "C" int mkf_model_get(const mkf_model* m, double* means, double* covs, double* weights, double* gamma,
                             double* pca_proj, double* pca_mean, double* Q, double* B, double* H, double* BH)
{
    if (!m) {
        mkf_set_error("null model");
        return MKF_E_INVALID;
    }
    auto cp = [](const std::vector<double>& v, double* o) {
        if (o) memcpy(o, v.data(), v.size() * sizeof(double));
    };
    cp(m->means, means);
    cp(m->covs, covs);
    cp(m->weights, weights);
    cp(m->gamma, gamma);
    cp(m->proj, pca_proj);
    cp(m->pmean, pca_mean);
    cp(m->Q, Q);
    cp(m->B, B);
    cp(m->H, H);
    cp(m->BH, BH);
    return MKF_OK;
}

// camera_matrix of a ROS camera_calibration YAML (cal.yml:4-7); used only by the get3Dpose back-end
extern "C" int mkf_load_camera_matrix(const char* path, double* K9)
{
    if (!path || !K9) {
        mkf_set_error("mkf_load_camera_matrix: null argument");
        return MKF_E_INVALID;
    }
    std::ifstream f(path);
    if (!f) {
        mkf_set_error("cannot open '%s'", path);
        return MKF_E_IO;
    }
    std::stringstream ss;
    ss << f.rdbuf();
    const std::string txt = ss.str();
    size_t p = txt.find("camera_matrix:");
    if (p == std::string::npos || (p = txt.find("data:", p)) == std::string::npos ||
        (p = txt.find('[', p)) == std::string::npos) {
        mkf_set_error("'%s': no camera_matrix data", path);
        return MKF_E_PARSE;
    }
    const char* c = txt.c_str() + p + 1;
    for (int i = 0; i < 9; i++) {
        while (*c == ' ' || *c == ',' || *c == '\n' || *c == '\r' || *c == '\t') c++;
        char* e = nullptr;
        K9[i] = strtod(c, &e);
        if (e == c) {
            mkf_set_error("'%s': camera_matrix needs 9 numbers", path);
            return MKF_E_PARSE;
        }
        c = e;
    }
    return MKF_OK;
}
