// oracle/cvshim/cvshim.cpp -- TEST INFRASTRUCTURE: implementation of the OpenCV-2.4 subset declared
// in opencv2/core/core.hpp (see the header for scope and caveats).
#include <deque>
#include <fstream>
#include <sstream>

#include "opencv2/core/core.hpp"

namespace cv {

// ------------------------------------------------------------------------------------------------
// tick source controlled by the harness
// ------------------------------------------------------------------------------------------------
static thread_local std::deque<int64> g_ticks;
static thread_local int64 g_tick_counter = 1;
int64 getTickCount()
{
    if (!g_ticks.empty()) {
        int64 t = g_ticks.front();
        g_ticks.pop_front();
        return t;
    }
    return g_tick_counter++;
}
void cvshim_push_tick(int64 t) { g_ticks.push_back(t); }

// ------------------------------------------------------------------------------------------------
// gemm: GEMMSingleMul operation order (modules/core/src/matmul.cpp) for small CV_64F matrices
// ------------------------------------------------------------------------------------------------
void gemm(const Mat& matA, const Mat& matB, double alpha, const Mat& matC, double beta, Mat& matD, int flags)
{
    Mat A = matA, B = matB, C = (beta != 0) ? matC : Mat();
    const bool at = (flags & GEMM_1_T) != 0, bt = (flags & GEMM_2_T) != 0, ct = (flags & GEMM_3_T) != 0;
    const int a_rows = at ? A.cols : A.rows, a_cols = at ? A.rows : A.cols;
    const int b_rows = bt ? B.cols : B.rows, b_cols = bt ? B.rows : B.cols;
    cvshim_assert(a_cols == b_rows, "gemm inner dimensions");
    if (!C.empty()) cvshim_assert((ct ? C.cols : C.rows) == a_rows && (ct ? C.rows : C.cols) == b_cols, "gemm C size");
    matD.create(a_rows, b_cols, CV_64F);
    Mat D = matD;
    Mat tmp;
    if (D.data == A.data || D.data == B.data) { // cv::gemm computes into a temporary, then copies
        tmp.create(a_rows, b_cols, CV_64F);
        D = tmp;
    }
    const int n = a_cols;
    std::vector<double> abuf;
    for (int i = 0; i < a_rows; i++) {
        // row i of op(A) (A^T rows are gathered into a_buf)
        abuf.resize(n);
        for (int k = 0; k < n; k++) abuf[k] = at ? A.el(k, i) : A.el(i, k);
        const double* a = abuf.data();
        for (int j = 0; j < b_cols; j++) {
            double s0;
            if (bt) {
                const double* b = B.ptr<double>(j);
                double t0 = 0, t1 = 0, t2 = 0, t3 = 0;
                int k = 0;
                for (; k <= n - 4; k += 4) {
                    t0 += a[k] * b[k];
                    t1 += a[k + 1] * b[k + 1];
                    t2 += a[k + 2] * b[k + 2];
                    t3 += a[k + 3] * b[k + 3];
                }
                for (; k < n; k++) t0 += a[k] * b[k];
                s0 = (t0 + t1 + t2 + t3) * alpha;
            } else {
                double t0 = 0;
                for (int k = 0; k < n; k++) t0 += a[k] * B.el(k, j);
                s0 = t0 * alpha;
            }
            if (C.empty())
                D.el(i, j) = s0;
            else
                D.el(i, j) = s0 + (ct ? C.el(j, i) : C.el(i, j)) * beta;
        }
    }
    if (D.data != matD.data) D.copyTo(matD);
}

// CholImpl<double> of OpenCV 2.4.x: lower triangle in place, diagonal left as 1/L_ii when b == NULL
bool Cholesky(double* A, size_t astep, int m, double* b, size_t bstep, int n)
{
    double* L = A;
    int i, j, k;
    double s;
    astep /= sizeof(A[0]);
    bstep /= sizeof(double);
    for (i = 0; i < m; i++) {
        for (j = 0; j < i; j++) {
            s = A[i * astep + j];
            for (k = 0; k < j; k++) s -= L[i * astep + k] * L[j * astep + k];
            L[i * astep + j] = s * L[j * astep + j];
        }
        s = A[i * astep + i];
        for (k = 0; k < j; k++) {
            double t = L[i * astep + k];
            s -= t * t;
        }
        if (s < std::numeric_limits<double>::epsilon()) return false;
        L[i * astep + i] = 1. / std::sqrt(s);
    }
    if (!b) return true;
    // LLt x = b: forward then backward substitution
    for (i = 0; i < m; i++)
        for (j = 0; j < n; j++) {
            s = b[i * bstep + j];
            for (k = 0; k < i; k++) s -= L[i * astep + k] * b[k * bstep + j];
            b[i * bstep + j] = s * L[i * astep + i];
        }
    for (i = m - 1; i >= 0; i--)
        for (j = 0; j < n; j++) {
            s = b[i * bstep + j];
            for (k = m - 1; k > i; k--) s -= L[k * astep + i] * b[k * bstep + j];
            b[i * bstep + j] = s * L[i * astep + i];
        }
    return true;
}

// LUImpl<double> of OpenCV 2.4.x
int LU(double* A, size_t astep, int m, double* b, size_t bstep, int n)
{
    int i, j, k, p = 1;
    astep /= sizeof(A[0]);
    bstep /= sizeof(double);
    for (i = 0; i < m; i++) {
        k = i;
        for (j = i + 1; j < m; j++)
            if (std::abs(A[j * astep + i]) > std::abs(A[k * astep + i])) k = j;
        if (std::abs(A[k * astep + i]) < std::numeric_limits<double>::epsilon()) return 0;
        if (k != i) {
            for (j = i; j < m; j++) std::swap(A[i * astep + j], A[k * astep + j]);
            if (b)
                for (j = 0; j < n; j++) std::swap(b[i * bstep + j], b[k * bstep + j]);
            p = -p;
        }
        double d = -1 / A[i * astep + i];
        for (j = i + 1; j < m; j++) {
            double alpha = A[j * astep + i] * d;
            for (k = i + 1; k < m; k++) A[j * astep + k] += alpha * A[i * astep + k];
            if (b)
                for (k = 0; k < n; k++) b[j * bstep + k] += alpha * b[i * bstep + k];
        }
        A[i * astep + i] = -d;
    }
    if (b) {
        for (i = m - 1; i >= 0; i--)
            for (j = 0; j < n; j++) {
                double s = b[i * bstep + j];
                for (k = i + 1; k < m; k++) s -= A[i * astep + k] * b[k * bstep + j];
                b[i * bstep + j] = s * A[i * astep + i];
            }
    }
    return p;
}

// cv::determinant (core/src/lapack.cpp, OpenCV 2.4): closed forms for 1 x 1 .. 3 x 3, otherwise LU on a copy with
// det = sign / prod(stored diagonal) -- LU() leaves the RECIPROCALS of the pivots on the diagonal
double determinant(const Mat& m)
{
    cvshim_assert(m.rows == m.cols && m.rows > 0, "determinant needs a square matrix");
    const int n = m.rows;
#define Md(y, x) m.at<double>(y, x)
    if (n == 1) return Md(0, 0);
    if (n == 2) return Md(0, 0) * Md(1, 1) - Md(0, 1) * Md(1, 0);
    if (n == 3)
        return Md(0, 0) * (Md(1, 1) * Md(2, 2) - Md(1, 2) * Md(2, 1)) - Md(0, 1) * (Md(1, 0) * Md(2, 2) - Md(1, 2) * Md(2, 0)) +
               Md(0, 2) * (Md(1, 0) * Md(2, 1) - Md(1, 1) * Md(2, 0));
#undef Md
    Mat a = m.clone();
    double result = LU(a.ptr<double>(), a.step, n, 0, 0, 0);
    if (result) {
        for (int i = 0; i < n; i++) result *= a.at<double>(i, i);
        result = 1. / result;
    }
    return result;
}

double invert(const Mat& src, Mat& dst, int method)
{
    cvshim_assert(src.rows == src.cols, "invert needs a square matrix");
    cvshim_assert(method == DECOMP_LU || method == DECOMP_CHOLESKY, "invert method");
    const int n = src.rows;
    Mat s = src.clone();
    dst.create(n, n, CV_64F);
    bool result = false;
    if (n <= 3 && method == DECOMP_LU) {
        if (n == 2) {
            double d = s.el(0, 0) * s.el(1, 1) - s.el(0, 1) * s.el(1, 0);
            if (d != 0.) {
                result = true;
                d = 1. / d;
                double t0, t1;
                t0 = s.el(0, 0) * d;
                t1 = s.el(1, 1) * d;
                dst.el(1, 1) = t0;
                dst.el(0, 0) = t1;
                t0 = -s.el(0, 1) * d;
                t1 = -s.el(1, 0) * d;
                dst.el(0, 1) = t0;
                dst.el(1, 0) = t1;
            }
        } else if (n == 1) {
            double d = s.el(0, 0);
            if (d != 0.) {
                result = true;
                dst.el(0, 0) = 1. / d;
            }
        } else { // n == 3: determinant + adjugate
            double d = s.el(0, 0) * (s.el(1, 1) * s.el(2, 2) - s.el(1, 2) * s.el(2, 1)) -
                       s.el(0, 1) * (s.el(1, 0) * s.el(2, 2) - s.el(1, 2) * s.el(2, 0)) +
                       s.el(0, 2) * (s.el(1, 0) * s.el(2, 1) - s.el(1, 1) * s.el(2, 0));
            if (d != 0.) {
                result = true;
                d = 1. / d;
                double t[9];
                t[0] = (s.el(1, 1) * s.el(2, 2) - s.el(1, 2) * s.el(2, 1)) * d;
                t[1] = (s.el(0, 2) * s.el(2, 1) - s.el(0, 1) * s.el(2, 2)) * d;
                t[2] = (s.el(0, 1) * s.el(1, 2) - s.el(0, 2) * s.el(1, 1)) * d;
                t[3] = (s.el(1, 2) * s.el(2, 0) - s.el(1, 0) * s.el(2, 2)) * d;
                t[4] = (s.el(0, 0) * s.el(2, 2) - s.el(0, 2) * s.el(2, 0)) * d;
                t[5] = (s.el(0, 2) * s.el(1, 0) - s.el(0, 0) * s.el(1, 2)) * d;
                t[6] = (s.el(1, 0) * s.el(2, 1) - s.el(1, 1) * s.el(2, 0)) * d;
                t[7] = (s.el(0, 1) * s.el(2, 0) - s.el(0, 0) * s.el(2, 1)) * d;
                t[8] = (s.el(0, 0) * s.el(1, 1) - s.el(0, 1) * s.el(1, 0)) * d;
                for (int i = 0; i < 9; i++) dst.el(i / 3, i % 3) = t[i];
            }
        }
    } else {
        setIdentity(dst);
        if (method == DECOMP_LU)
            result = LU(s.ptr<double>(), s.step, n, dst.ptr<double>(), dst.step, n) != 0;
        else
            result = Cholesky(s.ptr<double>(), s.step, n, dst.ptr<double>(), dst.step, n);
    }
    if (!result)
        for (int r = 0; r < n; r++)
            for (int c = 0; c < n; c++) dst.el(r, c) = 0;
    return result;
}

void transpose(const Mat& src, Mat& dst)
{
    Mat s = (dst.data == src.data) ? src.clone() : src;
    dst.create(src.cols, src.rows, CV_64F);
    for (int r = 0; r < s.rows; r++)
        for (int c = 0; c < s.cols; c++) dst.el(c, r) = s.el(r, c);
}

void setIdentity(Mat& m, const Scalar& s)
{
    for (int r = 0; r < m.rows; r++)
        for (int c = 0; c < m.cols; c++) m.el(r, c) = (r == c) ? s[0] : 0.0;
}

static void binary(const Mat& a, const Mat& b, Mat& dst, int kind, double alpha, double beta, double gamma)
{
    cvshim_assert(a.rows == b.rows && a.cols == b.cols && a.channels() == b.channels(), "size mismatch in arithmetic op");
    Mat aa = a, bb = b;
    dst.create(a.rows, a.cols, a.type());
    for (int r = 0; r < a.rows; r++) {
        const double* pa = aa.ptr<double>(r);
        const double* pb = bb.ptr<double>(r);
        double* pd = dst.ptr<double>(r);
        for (int c = 0; c < a.cols * a.channels(); c++) {
            switch (kind) {
            case 0: pd[c] = pa[c] + pb[c]; break;
            case 1: pd[c] = pa[c] - pb[c]; break;
            case 2: pd[c] = pa[c] * alpha + pb[c]; break;                // scaleAdd
            default: pd[c] = pa[c] * alpha + pb[c] * beta + gamma; break; // addWeighted
            }
        }
    }
}
void add(const Mat& a, const Mat& b, Mat& dst) { binary(a, b, dst, 0, 0, 0, 0); }
void subtract(const Mat& a, const Mat& b, Mat& dst) { binary(a, b, dst, 1, 0, 0, 0); }
void scaleAdd(const Mat& a, double alpha, const Mat& b, Mat& dst) { binary(a, b, dst, 2, alpha, 0, 0); }
void addWeighted(const Mat& a, double alpha, const Mat& b, double beta, double gamma, Mat& dst)
{
    binary(a, b, dst, 3, alpha, beta, gamma);
}

// cv::log / cv::exp are table-driven in OpenCV (accurate to ~1e-16 relative); libm here
void log(const Mat& src, Mat& dst)
{
    Mat s = src;
    dst.create(src.rows, src.cols, CV_64F);
    for (int r = 0; r < s.rows; r++)
        for (int c = 0; c < s.cols; c++) dst.el(r, c) = std::log(s.el(r, c));
}
void exp(const Mat& src, Mat& dst)
{
    Mat s = src;
    dst.create(src.rows, src.cols, CV_64F);
    for (int r = 0; r < s.rows; r++)
        for (int c = 0; c < s.cols; c++) dst.el(r, c) = std::exp(s.el(r, c));
}
void pow(const Mat& src, double power, Mat& dst)
{
    Mat s = src;
    dst.create(src.rows, src.cols, CV_64F);
    for (int r = 0; r < s.rows; r++)
        for (int c = 0; c < s.cols; c++) {
            const double v = s.el(r, c);
            dst.el(r, c) = (power == 2) ? v * v : std::pow(v, power); // integer power 2: multiply(src, src)
        }
}
Scalar sum(const Mat& src)
{
    double s0 = 0;
    if (src.isContinuous()) { // sum_: unrolled by 4
        const int len = src.rows * src.cols;
        const double* p = src.ptr<double>();
        int i = 0;
        for (; i <= len - 4; i += 4) s0 += p[i] + p[i + 1] + p[i + 2] + p[i + 3];
        for (; i < len; i++) s0 += p[i];
    } else {
        for (int r = 0; r < src.rows; r++)
            for (int c = 0; c < src.cols; c++) s0 += src.el(r, c);
    }
    return Scalar(s0);
}
void reduce(const Mat& src, Mat& dst, int dim, int rtype, int)
{
    cvshim_assert(rtype == CV_REDUCE_SUM, "reduce: only CV_REDUCE_SUM");
    Mat s = (dst.data == src.data) ? src.clone() : src;
    if (dim == 1) { // to a single column: reduceC_, two interleaved accumulators
        dst.create(s.rows, 1, CV_64F);
        for (int r = 0; r < s.rows; r++) {
            const double* p = s.ptr<double>(r);
            const int w = s.cols;
            if (w == 1) {
                dst.el(r, 0) = p[0];
                continue;
            }
            double a0 = p[0], a1 = p[1];
            int i = 2;
            for (; i <= w - 2; i += 2) {
                a0 = a0 + p[i];
                a1 = a1 + p[i + 1];
            }
            for (; i < w; i++) a0 = a0 + p[i];
            dst.el(r, 0) = a0 + a1;
        }
    } else { // to a single row: column sums, row by row
        dst.create(1, s.cols, CV_64F);
        for (int c = 0; c < s.cols; c++) {
            double a = s.el(0, c);
            for (int r = 1; r < s.rows; r++) a += s.el(r, c);
            dst.el(0, c) = a;
        }
    }
}
Mat repeat(const Mat& src, int ny, int nx)
{
    Mat d(src.rows * ny, src.cols * nx, CV_64F);
    for (int r = 0; r < d.rows; r++)
        for (int c = 0; c < d.cols; c++) d.el(r, c) = src.el(r % src.rows, c % src.cols);
    return d;
}
void split(const Mat& src, std::vector<Mat>& mv)
{
    const int cn = src.channels();
    mv.clear();
    for (int k = 0; k < cn; k++) {
        Mat m(src.rows, src.cols, CV_64F);
        for (int r = 0; r < src.rows; r++)
            for (int c = 0; c < src.cols; c++) m.el(r, c) = src.ptr<double>(r)[c * cn + k];
        mv.push_back(m);
    }
}
void vconcat(const Mat& a, const Mat& b, Mat& dst)
{
    cvshim_assert(a.cols == b.cols, "vconcat width");
    Mat aa = a.clone(), bb = b.clone();
    dst.create(a.rows + b.rows, a.cols, CV_64F);
    for (int r = 0; r < aa.rows; r++)
        for (int c = 0; c < aa.cols; c++) dst.el(r, c) = aa.el(r, c);
    for (int r = 0; r < bb.rows; r++)
        for (int c = 0; c < bb.cols; c++) dst.el(aa.rows + r, c) = bb.el(r, c);
}
// cv::randn with a cn x cn "stddev" matrix: dst = mean + stddev * N(0, I) per element.  The Gaussian
// source (Box-Muller on a cv::RNG) is NOT OpenCV's Ziggurat: distribution-equivalent only.
static thread_local RNG g_the_rng(0x12345678ULL);
void cvshim_seed_the_rng(uint64 seed) { g_the_rng = RNG(seed); } // cv::theRNG().state = seed
void randn(Mat& dst, const Mat& mean, const Mat& stddev)
{
    const int cn = dst.channels();
    cvshim_assert(stddev.rows == cn && stddev.cols == cn && mean.rows * mean.cols == cn, "randn: matrix stddev form only");
    std::vector<double> z(cn);
    for (int r = 0; r < dst.rows; r++)
        for (int c = 0; c < dst.cols; c++) {
            for (int k = 0; k < cn; k++) {
                double u1 = g_the_rng.uniform(0.0, 1.0), u2 = g_the_rng.uniform(0.0, 1.0);
                z[k] = std::sqrt(-2.0 * std::log(1.0 - u1)) * std::cos(2 * M_PI * u2);
            }
            for (int k = 0; k < cn; k++) {
                double v = mean.at<double>(k);
                for (int q = 0; q < cn; q++) v += stddev.el(k, q) * z[q];
                dst.ptr<double>(r)[c * cn + k] = v;
            }
        }
    std::vector<double> rec;
    for (int r = 0; r < dst.rows; r++)
        for (int c = 0; c < dst.cols * cn; c++) rec.push_back(dst.ptr<double>(r)[c]);
    cvshim_random_log().push_back(rec);
}
// scalar form cv::randn(dst, 0, 5) on a view (src/pf2D.cpp:96-98): dst = mean + stddev * N(0, 1) per element, same
// Gaussian source as above
void randn(const Mat& dstc, double mean, double stddev)
{
    Mat dst = dstc;
    std::vector<double> rec;
    for (int r = 0; r < dst.rows; r++)
        for (int c = 0; c < dst.cols; c++) {
            const double u1 = g_the_rng.uniform(0.0, 1.0), u2 = g_the_rng.uniform(0.0, 1.0);
            const double z = std::sqrt(-2.0 * std::log(1.0 - u1)) * std::cos(2 * M_PI * u2);
            const double v = z * stddev + mean;
            dst.el(r, c) = v;
            rec.push_back(v);
        }
    cvshim_random_log().push_back(rec);
}
std::vector<std::vector<double> >& cvshim_random_log()
{
    static thread_local std::vector<std::vector<double> > log;
    return log;
}
// cv::randu(dst, low, high) for CV_64F: low + (high - low) * U[0,1) from the global RNG (not OpenCV's stream)
void randu(const Mat& dstc, double low, double high)
{
    Mat dst = dstc;
    std::vector<double> rec;
    for (int r = 0; r < dst.rows; r++)
        for (int c = 0; c < dst.cols; c++) {
            const double v = low + (high - low) * g_the_rng.uniform(0.0, 1.0);
            dst.el(r, c) = v;
            rec.push_back(v);
        }
    cvshim_random_log().push_back(rec);
}
void hconcat(const Mat& a, const Mat& b, Mat& dst)
{
    cvshim_assert(a.rows == b.rows, "hconcat height");
    Mat aa = a.clone(), bb = b.clone();
    dst.create(a.rows, a.cols + b.cols, CV_64F);
    for (int r = 0; r < aa.rows; r++) {
        for (int c = 0; c < aa.cols; c++) dst.el(r, c) = aa.el(r, c);
        for (int c = 0; c < bb.cols; c++) dst.el(r, aa.cols + c) = bb.el(r, c);
    }
}

// GaussianBlur for CV_8UC1 (the likelihood image, src/pfPose.cpp:213): separable, BORDER_REFLECT_101, the 8-bit
// fixed-point path of OpenCV 2.4's filter engine (kernel scaled by 2^8 per pass, rounding shift by 16 at the end).
// Restated from memory; the product does not blur (the caller hands over the blurred image).
void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sigmaX, double sigmaY, int)
{
    cvshim_assert(src.depth() == CV_8U && src.channels() == 1, "GaussianBlur: CV_8UC1 only");
    if (sigmaY <= 0) sigmaY = sigmaX;
    auto kernel = [](int n, double sigma) {
        std::vector<double> k(n);
        double sum = 0;
        for (int i = 0; i < n; i++) {
            const double x = i - (n - 1) * 0.5;
            k[i] = std::exp(-0.5 * x * x / (sigma * sigma));
            sum += k[i];
        }
        std::vector<int> ki(n);
        for (int i = 0; i < n; i++) ki[i] = (int)std::lrint((double)(float)(k[i] / sum) * 256.0);
        return ki;
    };
    const std::vector<int> kx = kernel(ksize.width, sigmaX), ky = kernel(ksize.height, sigmaY);
    auto refl = [](int p, int n) {
        if (n == 1) return 0;
        while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
        return p;
    };
    Mat s = src.clone();
    const int R = s.rows, C = s.cols, rx = ksize.width / 2, ry = ksize.height / 2;
    std::vector<int> tmp((size_t)R * C);
    for (int r = 0; r < R; r++)
        for (int c = 0; c < C; c++) {
            int acc = 0;
            for (int k = -rx; k <= rx; k++) acc += kx[k + rx] * (int)s.at<uchar>(r, refl(c + k, C));
            tmp[(size_t)r * C + c] = acc;
        }
    dst.create(R, C, CV_8UC1);
    for (int r = 0; r < R; r++)
        for (int c = 0; c < C; c++) {
            int acc = 0;
            for (int k = -ry; k <= ry; k++) acc += ky[k + ry] * tmp[(size_t)refl(r + k, R) * C + c];
            const int v = (acc + (1 << 15)) >> 16;
            dst.at<uchar>(r, c) = (uchar)(v < 0 ? 0 : (v > 255 ? 255 : v));
        }
}
void cvtColor(const Mat& src, Mat& dst, int code)
{
    cvshim_assert(code == CV_GRAY2RGB && src.depth() == CV_8U && src.channels() == 1, "cvtColor: GRAY2RGB only");
    Mat s = src;
    dst.create(s.rows, s.cols, CV_8UC3);
    for (int r = 0; r < s.rows; r++)
        for (int c = 0; c < s.cols; c++) {
            const uchar v = s.at<uchar>(r, c);
            uchar* d = dst.data + (size_t)r * dst.step + (size_t)c * 3;
            d[0] = d[1] = d[2] = v;
        }
}

// cv::FileStorage (OpenCV-YAML-1.0), just enough for `fs["key"] >> mat` on !!opencv-matrix nodes
FileStorage::FileStorage(const std::string& path, int)
{
    std::ifstream f(path.c_str());
    if (f) {
        std::stringstream ss;
        ss << f.rdbuf();
        txt_ = ss.str();
    }
}
FileNode FileStorage::operator[](const char* key) const
{
    FileNode n;
    const std::string pat = std::string(key) + ":";
    size_t pos = 0;
    for (;;) {
        pos = txt_.find(pat, pos);
        if (pos == std::string::npos) return n;
        if (pos == 0 || txt_[pos - 1] == '\n') break;
        pos += pat.size();
    }
    const size_t end = txt_.find(']', pos);
    if (end != std::string::npos) n.text = txt_.substr(pos, end - pos + 1);
    return n;
}
void operator>>(const FileNode& n, Mat& m)
{
    if (n.text.empty()) {
        m = Mat();
        return;
    }
    auto num = [&](const char* name) {
        const size_t p = n.text.find(name);
        cvshim_assert(p != std::string::npos, "FileNode: missing field");
        return std::atoi(n.text.c_str() + p + std::strlen(name));
    };
    const int rows = num("rows:"), cols = num("cols:");
    size_t p = n.text.find("dt:");
    cvshim_assert(p != std::string::npos, "FileNode: missing dt");
    p += 3;
    while (n.text[p] == ' ' || n.text[p] == '"') p++;
    const char dt = n.text[p];
    cvshim_assert(dt == 'd' || dt == 'f', "FileNode: dt must be d or f");
    m.create(rows, cols, dt == 'd' ? CV_64F : CV_32F);
    p = n.text.find('[', n.text.find("data:"));
    const char* c = n.text.c_str() + p + 1;
    for (int i = 0; i < rows * cols; i++) {
        while (*c == ' ' || *c == ',' || *c == '\n' || *c == '\r' || *c == '\t') c++;
        char* e = 0;
        const double v = std::strtod(c, &e);
        cvshim_assert(e != c, "FileNode: malformed number");
        if (dt == 'd')
            m.at<double>(i / cols, i % cols) = v;
        else
            m.at<float>(i / cols, i % cols) = (float)v;
        c = e;
    }
}
Mat& operator*=(Mat& a, double s) // a.convertTo(a, a.type(), s): in place, through views
{
    for (int r = 0; r < a.rows; r++) {
        double* p = a.ptr<double>(r);
        for (int c = 0; c < a.cols * a.channels(); c++) p[c] = p[c] * s + 0;
    }
    return a;
}

// ------------------------------------------------------------------------------------------------
// MatExpr (modules/core/src/matop.cpp)
// ------------------------------------------------------------------------------------------------
static inline bool isIdentity(const MatExpr& e) { return e.op == MatExpr::OP_IDENTITY; }
static inline bool isAddEx(const MatExpr& e) { return e.op == MatExpr::OP_ADDEX; }
static inline bool isScaled(const MatExpr& e) { return isAddEx(e) && (!e.b.data || e.beta == 0) && e.s == Scalar(); }
static inline bool isT(const MatExpr& e) { return e.op == MatExpr::OP_T; }
static inline bool isMatProd(const MatExpr& e) { return e.op == MatExpr::OP_GEMM && (!e.c.data || e.beta == 0); }

static MatExpr makeAddEx(const Mat& a, const Mat& b, double alpha, double beta, const Scalar& s = Scalar())
{
    MatExpr e;
    e.op = MatExpr::OP_ADDEX;
    e.a = a;
    e.b = b;
    e.alpha = alpha;
    e.beta = beta;
    e.s = s;
    return e;
}
static MatExpr makeGemm(int flags, const Mat& a, const Mat& b, double alpha = 1, const Mat& c = Mat(), double beta = 1)
{
    MatExpr e;
    e.op = MatExpr::OP_GEMM;
    e.flags = flags;
    e.a = a;
    e.b = b;
    e.c = c;
    e.alpha = alpha;
    e.beta = beta;
    return e;
}
static MatExpr makeT(const Mat& a, double alpha = 1)
{
    MatExpr e;
    e.op = MatExpr::OP_T;
    e.a = a;
    e.alpha = alpha;
    return e;
}

void MatExpr::assign(Mat& m) const
{
    switch (op) {
    case OP_IDENTITY: m = a; break; // shallow
    case OP_ADDEX:
        if (b.data) {
            if (s == Scalar() || !s.isReal()) {
                if (alpha == 1) {
                    if (beta == 1)
                        cv::add(a, b, m);
                    else if (beta == -1)
                        cv::subtract(a, b, m);
                    else
                        cv::scaleAdd(b, beta, a, m);
                } else if (beta == 1) {
                    if (alpha == -1)
                        cv::subtract(b, a, m);
                    else
                        cv::scaleAdd(a, alpha, b, m);
                } else
                    cv::addWeighted(a, alpha, b, beta, 0, m);
            } else
                cv::addWeighted(a, alpha, b, beta, s[0], m);
        } else {
            // b empty: convertTo(m, type, alpha, s[0]) -- also how `empty + alpha*X` ends up as alpha*X (quirk B8)
            a.convertTo(m, a.type(), alpha, s[0]);
        }
        break;
    case OP_GEMM: cv::gemm(a, b, alpha, c, beta, m, flags); break;
    case OP_T:
        cv::transpose(a, m);
        if (alpha != 1) m.convertTo(m, m.type(), alpha);
        break;
    case OP_INVERT: cv::invert(a, m, flags); break;
    case OP_INIT:
        m.create(irows, icols, itype);
        if (flags == 'I')
            setIdentity(m, Scalar(alpha));
        else
            for (int r = 0; r < m.rows; r++) std::memset(m.data + (size_t)r * m.step, 0, m.elemSize() * (size_t)m.cols);
        break;
    }
}

Mat& Mat::operator=(const MatExpr& e)
{
    e.assign(*this);
    return *this;
}
MatExpr Mat::zeros(int r, int c, int type)
{
    MatExpr e;
    e.op = MatExpr::OP_INIT;
    e.flags = '0';
    e.irows = r;
    e.icols = c;
    e.itype = type;
    return e;
}
MatExpr Mat::eye(int r, int c, int type)
{
    MatExpr e = zeros(r, c, type);
    e.flags = 'I';
    e.alpha = 1;
    return e;
}
MatExpr Mat::t() const { return makeT(*this, 1); }
MatExpr Mat::inv(int method) const
{
    MatExpr e;
    e.op = MatExpr::OP_INVERT;
    e.flags = method;
    e.a = *this;
    return e;
}
MatExpr MatExpr::t() const
{
    // MatOp_T::transpose / MatOp_AddEx::transpose / MatOp::transpose
    if (op == OP_T) {
        if (alpha == 1) return MatExpr(a);
        return makeAddEx(a, Mat(), alpha, 0);
    }
    if (isScaled(*this)) return makeT(a, alpha);
    if (op == OP_GEMM) { // MatOp_GEMM::transpose: (A*B)^T = B^T * A^T
        MatExpr e = *this;
        e.flags = (!(flags & CV_GEMM_A_T) ? CV_GEMM_B_T : 0) | (!(flags & CV_GEMM_B_T) ? CV_GEMM_A_T : 0) |
                  (!(flags & CV_GEMM_C_T) ? CV_GEMM_C_T : 0);
        std::swap(e.a, e.b);
        return e;
    }
    Mat m;
    assign(m);
    return makeT(m, 1);
}
MatExpr MatExpr::inv(int method) const
{
    Mat m;
    assign(m);
    return m.inv(method);
}

// MatOp::add / MatOp_GEMM::add
static MatExpr expr_add(const MatExpr& e1, const MatExpr& e2)
{
    if (e1.op == MatExpr::OP_GEMM || e2.op == MatExpr::OP_GEMM) {
        const bool i1 = isIdentity(e1), i2 = isIdentity(e2);
        const double alpha1 = i1 ? 1 : e1.alpha, alpha2 = i2 ? 1 : e2.alpha;
        if (isMatProd(e1) && (i2 || isScaled(e2) || isT(e2)))
            return makeGemm((e1.flags & ~CV_GEMM_C_T) | (isT(e2) ? CV_GEMM_C_T : 0), e1.a, e1.b, alpha1, e2.a, alpha2);
        if (isMatProd(e2) && (i1 || isScaled(e1) || isT(e1)))
            return makeGemm((e2.flags & ~CV_GEMM_C_T) | (isT(e1) ? CV_GEMM_C_T : 0), e2.a, e2.b, alpha2, e1.a, alpha1);
    }
    double alpha = 1, beta = 1;
    Scalar s;
    Mat m1, m2;
    if (isAddEx(e1) && (!e1.b.data || e1.beta == 0)) {
        m1 = e1.a;
        alpha = e1.alpha;
        s = e1.s;
    } else
        e1.assign(m1);
    if (isAddEx(e2) && (!e2.b.data || e2.beta == 0)) {
        m2 = e2.a;
        beta = e2.alpha;
        s[0] += e2.s[0];
    } else
        e2.assign(m2);
    return makeAddEx(m1, m2, alpha, beta, s);
}
// MatOp::subtract / MatOp_GEMM::subtract
static MatExpr expr_sub(const MatExpr& e1, const MatExpr& e2)
{
    if (e1.op == MatExpr::OP_GEMM || e2.op == MatExpr::OP_GEMM) {
        const bool i1 = isIdentity(e1), i2 = isIdentity(e2);
        const double alpha1 = i1 ? 1 : e1.alpha, alpha2 = i2 ? 1 : e2.alpha;
        if (isMatProd(e1) && (i2 || isScaled(e2) || isT(e2)))
            return makeGemm((e1.flags & ~CV_GEMM_C_T) | (isT(e2) ? CV_GEMM_C_T : 0), e1.a, e1.b, alpha1, e2.a, -alpha2);
        if (isMatProd(e2) && (i1 || isScaled(e1) || isT(e1)))
            return makeGemm((e2.flags & ~CV_GEMM_C_T) | (isT(e1) ? CV_GEMM_C_T : 0), e2.a, e2.b, -alpha2, e1.a, alpha1);
    }
    double alpha = 1, beta = -1;
    Scalar s;
    Mat m1, m2;
    if (isAddEx(e1) && (!e1.b.data || e1.beta == 0)) {
        m1 = e1.a;
        alpha = e1.alpha;
        s = e1.s;
    } else
        e1.assign(m1);
    if (isAddEx(e2) && (!e2.b.data || e2.beta == 0)) {
        m2 = e2.a;
        beta = -e2.alpha;
        s[0] -= e2.s[0];
    } else
        e2.assign(m2);
    return makeAddEx(m1, m2, alpha, beta, s);
}
// MatOp::matmul (MatOp_Invert::matmul reduces to it for the products that occur here)
static MatExpr expr_matmul(const MatExpr& e1, const MatExpr& e2)
{
    double scale = 1;
    int flags = 0;
    Mat m1, m2;
    if (isT(e1)) {
        flags = CV_GEMM_A_T;
        scale = e1.alpha;
        m1 = e1.a;
    } else if (isScaled(e1)) {
        scale = e1.alpha;
        m1 = e1.a;
    } else
        e1.assign(m1);
    if (isT(e2)) {
        flags |= CV_GEMM_B_T;
        scale *= e2.alpha;
        m2 = e2.a;
    } else if (isScaled(e2)) {
        scale *= e2.alpha;
        m2 = e2.a;
    } else
        e2.assign(m2);
    return makeGemm(flags, m1, m2, scale);
}
// MatOp_*::multiply(e, s)
static MatExpr expr_scale(const MatExpr& e, double s)
{
    MatExpr r = e;
    switch (e.op) {
    case MatExpr::OP_ADDEX:
        r.alpha *= s;
        r.beta *= s;
        r.s[0] *= s;
        return r;
    case MatExpr::OP_GEMM:
        r.alpha *= s;
        r.beta *= s;
        return r;
    case MatExpr::OP_T:
    case MatExpr::OP_INIT: r.alpha *= s; return r;
    case MatExpr::OP_IDENTITY: return makeAddEx(e.a, Mat(), s, 0);
    default: {
        Mat m;
        e.assign(m);
        return makeAddEx(m, Mat(), s, 0);
    }
    }
}

MatExpr operator+(const Mat& a, const Mat& b) { return makeAddEx(a, b, 1, 1); }
MatExpr operator+(const Mat& a, const MatExpr& e) { return expr_add(e, MatExpr(a)); } // e.op->add(e, MatExpr(a), en)
MatExpr operator+(const MatExpr& e, const Mat& b) { return expr_add(e, MatExpr(b)); }
MatExpr operator+(const MatExpr& e1, const MatExpr& e2) { return expr_add(e1, e2); }
MatExpr operator-(const Mat& a, const Mat& b) { return makeAddEx(a, b, 1, -1); }
MatExpr operator-(const Mat& a, const MatExpr& e) { return expr_sub(MatExpr(a), e); }
MatExpr operator-(const MatExpr& e, const Mat& b) { return expr_sub(e, MatExpr(b)); }
MatExpr operator-(const MatExpr& e1, const MatExpr& e2) { return expr_sub(e1, e2); }
MatExpr operator-(const MatExpr& e, double s)
{
    if (isAddEx(e)) { // MatOp_AddEx::add(e, -s)
        MatExpr r = e;
        r.s[0] += -s;
        return r;
    }
    Mat m;
    e.assign(m);
    return makeAddEx(m, Mat(), 1, 0, Scalar(-s));
}
MatExpr operator-(const Mat& a, double s) { return makeAddEx(a, Mat(), 1, 0, Scalar(-s)); }
MatExpr operator*(const Mat& a, const Mat& b) { return makeGemm(0, a, b); }
MatExpr operator*(const Mat& a, const MatExpr& e) { return expr_matmul(MatExpr(a), e); }
MatExpr operator*(const MatExpr& e, const Mat& b) { return expr_matmul(e, MatExpr(b)); }
MatExpr operator*(const MatExpr& e1, const MatExpr& e2) { return expr_matmul(e1, e2); }
MatExpr operator*(double s, const Mat& a) { return makeAddEx(a, Mat(), s, 0); }
MatExpr operator*(const Mat& a, double s) { return makeAddEx(a, Mat(), s, 0); }
MatExpr operator*(double s, const MatExpr& e) { return expr_scale(e, s); }
MatExpr operator*(const MatExpr& e, double s) { return expr_scale(e, s); }
MatExpr operator/(const MatExpr& e, double s) { return expr_scale(e, 1. / s); }
MatExpr operator/(const Mat& a, double s) { return makeAddEx(a, Mat(), 1. / s, 0); }

} // namespace cv
