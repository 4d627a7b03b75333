// oracle/cvshim/opencv2/core/core.hpp -- TEST INFRASTRUCTURE.
//
// A minimal restatement of the subset of OpenCV 2.4 `core` that the reference's hot-path sources
// (src/KF_model.cpp, src/my_gmm.cpp, src/pf2DRao.cpp) use, so that those files can be compiled IN
// PLACE, unmodified, into oracle/_ref/libref.so (see oracle/Makefile) without OpenCV or ROS.
// It is NOT OpenCV: it follows the published OpenCV 2.4.x algorithms and, for cv::MatExpr, the
// lazy-evaluation rules of modules/core/src/matop.cpp (which products are fused into one gemm,
// when a destination buffer is reused in place, ...) as far as those files exercise them.
// Written from memory of the OpenCV sources (OpenCV is not installed here); double precision only.
//
// What this buys: the reference's OWN control flow, operation sequence and cv::Mat aliasing
// (shallow operator=, in-place MatExpr assignment -> quirk B3, diag()/row() views in chol())
// drive the arithmetic, which pins the oracle restatement's reading of the reference.
#ifndef CVSHIM_CORE_HPP
#define CVSHIM_CORE_HPP

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define CV_8U 0
#define CV_32F 5
#define CV_64F 6
#define CV_8UC1 0
#define CV_8UC3 16
#define CV_32FC1 5
#define CV_64FC1 6
#define CV_64FC2 14
#define CV_MAT_DEPTH(t) ((t)&7)
#define CV_MAT_CN(t) ((((t) >> 3) & 63) + 1)
#define CV_MAKETYPE(depth, cn) (((depth)&7) + (((cn)-1) << 3))
#define CV_REDUCE_SUM 0
#define CV_GRAY2RGB 8
#define CV_GEMM_A_T 1
#define CV_GEMM_B_T 2
#define CV_GEMM_C_T 4

namespace cv {

typedef unsigned char uchar;
typedef int64_t int64;
typedef uint64_t uint64;

enum { GEMM_1_T = 1, GEMM_2_T = 2, GEMM_3_T = 4 };
enum { DECOMP_LU = 0, DECOMP_SVD = 1, DECOMP_EIG = 2, DECOMP_CHOLESKY = 3 };

struct Exception : std::runtime_error {
    explicit Exception(const std::string& s) : std::runtime_error(s) {}
};
inline void cvshim_assert(bool ok, const char* what)
{
    if (!ok) throw Exception(std::string("cvshim assertion failed: ") + what);
}

// the test harness controls what cv::getTickCount() returns (queue of values, then a counter)
int64 getTickCount();
void cvshim_push_tick(int64 t);
void cvshim_seed_the_rng(uint64_t seed); // state of the global generator behind randn / randu

struct Range {
    int start, end;
    Range() : start(0), end(0) {}
    Range(int s, int e) : start(s), end(e) {}
    static Range all() { return Range(std::numeric_limits<int>::min(), std::numeric_limits<int>::max()); }
};

struct Scalar {
    double val[4];
    Scalar() { val[0] = val[1] = val[2] = val[3] = 0; }
    Scalar(double v0, double v1 = 0, double v2 = 0, double v3 = 0)
    {
        val[0] = v0;
        val[1] = v1;
        val[2] = v2;
        val[3] = v3;
    }
    static Scalar all(double v) { return Scalar(v, v, v, v); }
    double& operator[](int i) { return val[i]; }
    const double& operator[](int i) const { return val[i]; }
    bool isReal() const { return val[1] == 0 && val[2] == 0 && val[3] == 0; }
    bool operator==(const Scalar& o) const
    {
        return val[0] == o.val[0] && val[1] == o.val[1] && val[2] == o.val[2] && val[3] == o.val[3];
    }
};

class MatExpr;

struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
};
struct Point {
    int x, y;
    Point() : x(0), y(0) {}
    Point(int x_, int y_) : x(x_), y(y_) {}
};
struct Vec3b {
    uchar val[3];
    uchar& operator[](int i) { return val[i]; }
    const uchar& operator[](int i) const { return val[i]; }
};

// matrix of depth CV_8U / CV_32F / CV_64F with 1..3 channels, row-major with a byte step; headers share the
// buffer.  The arithmetic (gemm, invert, MatExpr ...) is CV_64F only; the other depths exist for images and for
// the f32 model matrices, i.e. create / at / copy / convertTo / blur / cvtColor.
class Mat {
  public:
    int rows, cols;
    size_t step; // bytes between rows
    uchar* data;
    Mat() : rows(0), cols(0), step(0), data(0), cn_(1), depth_(CV_64F) {}
    Mat(int r, int c, int type) : rows(0), cols(0), step(0), data(0), cn_(1), depth_(CV_64F) { create(r, c, type); }
    Mat(Size sz, int type) : rows(0), cols(0), step(0), data(0), cn_(1), depth_(CV_64F) { create(sz.height, sz.width, type); }
    Mat(const Mat& m)
        : rows(m.rows), cols(m.cols), step(m.step), data(m.data), cn_(m.cn_), depth_(m.depth_), buf_(m.buf_)
    {
    }
    Mat& operator=(const Mat& m)
    {
        rows = m.rows;
        cols = m.cols;
        step = m.step;
        data = m.data;
        cn_ = m.cn_;
        depth_ = m.depth_;
        buf_ = m.buf_;
        return *this;
    }
    static size_t depthBytes(int depth) { return depth == CV_8U ? 1 : (depth == CV_32F ? 4 : 8); }
    size_t elemSize() const { return depthBytes(depth_) * cn_; }
    int depth() const { return depth_; }
    Mat& operator=(const MatExpr& e);

    // Mat::create: keeps the current buffer when shape and type already match (this is what makes
    // `state = F*state + B` write through to every header sharing the buffer)
    void create(int r, int c, int type)
    {
        const int cn = CV_MAT_CN(type), dp = CV_MAT_DEPTH(type);
        cvshim_assert((dp == CV_8U || dp == CV_32F || dp == CV_64F) && cn >= 1 && cn <= 3, "unsupported Mat type");
        if (data && rows == r && cols == c && cn_ == cn && depth_ == dp) return;
        rows = r;
        cols = c;
        cn_ = cn;
        depth_ = dp;
        step = elemSize() * (size_t)c;
        const size_t bytes = std::max<size_t>(step * (size_t)r, 8);
        buf_.reset(new double[(bytes + 7) / 8], std::default_delete<double[]>());
        data = (uchar*)buf_.get();
    }
    int type() const { return CV_MAKETYPE(depth_, cn_); }
    int channels() const { return cn_; }
    bool empty() const { return data == 0 || rows * cols == 0; }
    bool isContinuous() const { return step == elemSize() * (size_t)cols || rows <= 1; }
    Size size() const { return Size(cols, rows); }

    static MatExpr zeros(int r, int c, int type);
    static MatExpr eye(int r, int c, int type);

    template <class T>
    T& at(int r, int c)
    {
        return *(T*)(data + (size_t)r * step + (size_t)c * sizeof(T));
    }
    template <class T>
    const T& at(int r, int c) const
    {
        return *(const T*)(data + (size_t)r * step + (size_t)c * sizeof(T));
    }
    template <class T>
    T& at(int i) // single-index access of a row or column vector
    {
        return rows == 1 ? at<T>(0, i) : at<T>(i, 0);
    }
    template <class T>
    const T& at(int i) const
    {
        return rows == 1 ? at<T>(0, i) : at<T>(i, 0);
    }
    template <class T>
    T* ptr(int r = 0)
    {
        return (T*)(data + (size_t)r * step);
    }
    template <class T>
    const T* ptr(int r = 0) const
    {
        return (const T*)(data + (size_t)r * step);
    }
    double& el(int r, int c) { return at<double>(r, c); }
    const double& el(int r, int c) const { return at<double>(r, c); }

    Mat clone() const
    {
        Mat m;
        copyTo(m);
        return m;
    }
    // copyTo(OutputArray): a named Mat is (re)created; a temporary header (e.g. m.diag(-e)) is a
    // fixed-size destination that is written through
    void copyTo(Mat& dst) const
    {
        if (empty()) {
            dst = Mat();
            return;
        }
        dst.create(rows, cols, type());
        copy_elements(dst);
    }
    void copyTo(const Mat& dst) const
    {
        cvshim_assert(dst.rows == rows && dst.cols == cols && dst.cn_ == cn_ && dst.depth_ == depth_, "copyTo fixed-size destination");
        copy_elements(const_cast<Mat&>(dst));
    }
    void copyTo(std::vector<double>& v) const
    {
        v.resize((size_t)rows * cols);
        for (int r = 0; r < rows; r++)
            for (int c = 0; c < cols; c++) v[(size_t)r * cols + c] = el(r, c);
    }
    double getAs(int r, int i) const // i-th scalar of row r, any depth
    {
        const uchar* p = data + (size_t)r * step;
        return depth_ == CV_8U ? (double)p[i] : (depth_ == CV_32F ? (double)((const float*)p)[i] : ((const double*)p)[i]);
    }
    // convertTo(dst, rtype, alpha, beta): dst = saturate_cast<rtype>(src*alpha + beta); rtype < 0 keeps the depth
    void convertTo(Mat& dst, int rtype, double alpha = 1, double beta = 0) const
    {
        Mat src = *this; // keep the source header alive if dst aliases it
        const int dd = rtype < 0 ? depth_ : CV_MAT_DEPTH(rtype);
        dst.create(rows, cols, CV_MAKETYPE(dd, cn_));
        for (int r = 0; r < rows; r++) {
            uchar* d = dst.data + (size_t)r * dst.step;
            for (int c = 0; c < cols * cn_; c++) {
                const double v = src.getAs(r, c) * alpha + beta;
                if (dd == CV_64F)
                    ((double*)d)[c] = v;
                else if (dd == CV_32F)
                    ((float*)d)[c] = (float)v;
                else {
                    const long iv = std::lrint(v); // cvRound
                    d[c] = (uchar)(iv < 0 ? 0 : (iv > 255 ? 255 : iv));
                }
            }
        }
    }

    MatExpr t() const;
    MatExpr inv(int method = DECOMP_LU) const;

    Mat row(int r) const { return view(r, 0, 1, cols, step); }
    Mat col(int c) const { return view(0, c, rows, 1, step); }
    Mat rowRange(int a, int b) const { return view(a, 0, b - a, cols, step); }
    Mat rowRange(const Range& r) const { return rowRange(r.start, r.end); }
    Mat colRange(int a, int b) const { return view(0, a, rows, b - a, step); }
    Mat operator()(const Range& rr, const Range& cr) const
    {
        const int r0 = rr.start == std::numeric_limits<int>::min() ? 0 : rr.start;
        const int r1 = rr.end == std::numeric_limits<int>::max() ? rows : rr.end;
        const int c0 = cr.start == std::numeric_limits<int>::min() ? 0 : cr.start;
        const int c1 = cr.end == std::numeric_limits<int>::max() ? cols : cr.end;
        return view(r0, c0, r1 - r0, c1 - c0, step);
    }
    // Mat::diag(d): a column-vector VIEW of the d-th diagonal (d < 0: below the main diagonal)
    Mat diag(int d = 0) const
    {
        int len, r0, c0;
        if (d >= 0) {
            len = std::min(cols - d, rows);
            r0 = 0;
            c0 = d;
        } else {
            len = std::min(rows + d, cols);
            r0 = -d;
            c0 = 0;
        }
        cvshim_assert(len > 0, "diag out of range");
        return view(r0, c0, len, 1, step + elemSize());
    }

  private:
    Mat view(int r0, int c0, int nr, int nc, size_t st) const
    {
        Mat m;
        m.rows = nr;
        m.cols = nc;
        m.step = st;
        m.cn_ = cn_;
        m.depth_ = depth_;
        m.buf_ = buf_;
        m.data = data + (size_t)r0 * step + (size_t)c0 * elemSize();
        return m;
    }
    void copy_elements(Mat& dst) const
    {
        for (int r = 0; r < rows; r++) {
            const uchar* s = data + (size_t)r * step;
            uchar* d = dst.data + (size_t)r * dst.step;
            if (s != d) std::memmove(d, s, elemSize() * (size_t)cols);
        }
    }
    int cn_, depth_;
    std::shared_ptr<double> buf_;
};

// ------------------------------------------------------------------------------------------------
// array functions (CV_64F)
// ------------------------------------------------------------------------------------------------
void gemm(const Mat& A, const Mat& B, double alpha, const Mat& C, double beta, Mat& D, int flags = 0);
bool Cholesky(double* A, size_t astep, int m, double* b, size_t bstep, int n);
int LU(double* A, size_t astep, int m, double* b, size_t bstep, int n);
double invert(const Mat& src, Mat& dst, int method = DECOMP_LU);
void transpose(const Mat& src, Mat& dst);
void setIdentity(Mat& m, const Scalar& s = Scalar(1));
void add(const Mat& a, const Mat& b, Mat& dst);
void subtract(const Mat& a, const Mat& b, Mat& dst);
void scaleAdd(const Mat& a, double alpha, const Mat& b, Mat& dst); // dst = a*alpha + b
void addWeighted(const Mat& a, double alpha, const Mat& b, double beta, double gamma, Mat& dst);
void log(const Mat& src, Mat& dst);
void exp(const Mat& src, Mat& dst);
void pow(const Mat& src, double power, Mat& dst);
Scalar sum(const Mat& src);
void reduce(const Mat& src, Mat& dst, int dim, int rtype, int dtype = -1);
Mat repeat(const Mat& src, int ny, int nx);
void split(const Mat& src, std::vector<Mat>& mv);
void vconcat(const Mat& a, const Mat& b, Mat& dst);
void randn(Mat& dst, const Mat& mean, const Mat& stddev);
void randn(const Mat& dst, double mean, double stddev); // scalar form on a (view of a) CV_64F matrix (src/pf2D.cpp:96-98)
void randu(const Mat& dst, double low, double high); // fills a (view of a) CV_64F matrix
inline void randu(const Mat& dst, const Scalar& low, const Scalar& high) { randu(dst, low.val[0], high.val[0]); }
double determinant(const Mat& m); // core/src/lapack.cpp: closed forms up to 3 x 3, LU beyond
void hconcat(const Mat& a, const Mat& b, Mat& dst);
// every array produced by randn / randu, in call order (the harness replays the reference's internal draws)
std::vector<std::vector<double> >& cvshim_random_log();

// imgproc subset used by PFTracker::getMeasurementProposal / callback (src/pfPose.cpp:213-214, 253, 357-367)
enum { BORDER_DEFAULT = 4 };
void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_DEFAULT);
void cvtColor(const Mat& src, Mat& dst, int code);
inline void circle(Mat&, Point, int, const Scalar&, int = 1, int = 8, int = 0) {}
inline void line(Mat&, Point, Point, const Scalar&, int = 1, int = 8, int = 0) {}

// Mat_<T> with the comma initialiser of `(Mat_<double>(3,3) << a, b, ...)`
template <class T>
class Mat_ : public Mat {
  public:
    Mat_(int r, int c) : Mat(r, c, sizeof(T) == 8 ? CV_64F : (sizeof(T) == 4 ? CV_32F : CV_8U)) {}
};
template <class T>
class MatCommaInitializer_ {
  public:
    MatCommaInitializer_(const Mat_<T>& m) : m_(m), i_(0) {}
    template <class V>
    MatCommaInitializer_<T>& operator,(V v)
    {
        put((T)v);
        return *this;
    }
    void put(T v)
    {
        m_.template at<T>(i_ / m_.cols, i_ % m_.cols) = v;
        i_++;
    }
    operator Mat() const { return m_; }
    Mat_<T> m_;
    int i_;
};
template <class T, class V>
MatCommaInitializer_<T> operator<<(const Mat_<T>& m, V v)
{
    MatCommaInitializer_<T> ci(m);
    ci.put((T)v);
    return ci;
}

// cv::FileStorage reader for OpenCV-YAML-1.0 "!!opencv-matrix" nodes (src/pfPose.cpp:34-55)
class FileNode {
  public:
    std::string text; // the node's block
};
void operator>>(const FileNode& n, Mat& m);
class FileStorage {
  public:
    enum { READ = 0 };
    FileStorage(const std::string& path, int flags);
    FileNode operator[](const char* key) const;
    void release() {}
    bool isOpened() const { return !txt_.empty(); }

  private:
    std::string txt_;
};

Mat& operator*=(Mat& a, double s);
inline Mat& operator*=(Mat&& a, double s) { return operator*=(static_cast<Mat&>(a), s); }

class RNG {
  public:
    uint64 state;
    RNG() : state(0xffffffff) {}
    RNG(uint64 s) : state(s ? s : 0xffffffff) {}
    unsigned next()
    {
        state = (uint64)(unsigned)state * 4164903690U + (unsigned)(state >> 32);
        return (unsigned)state;
    }
    operator double()
    {
        unsigned t = next();
        return (((uint64)t << 32) | next()) * 5.4210108624275221700372640043497e-20;
    }
    int uniform(int a, int b) { return a == b ? a : (int)(next() % (b - a) + a); }
    double uniform(double a, double b) { return ((double)*this) * (b - a) + a; }
};

// ------------------------------------------------------------------------------------------------
// MatExpr: the lazy expression node of modules/core/src/matop.cpp, for the operators the
// reference uses.  `op` identifies the MatOp singleton.
// ------------------------------------------------------------------------------------------------
class MatExpr {
  public:
    enum Op { OP_IDENTITY, OP_ADDEX, OP_GEMM, OP_T, OP_INVERT, OP_INIT };
    Op op;
    int flags; // GEMM flags, invert method, or initializer kind ('0' zeros, 'I' eye)
    Mat a, b, c;
    double alpha, beta;
    Scalar s;
    int irows, icols, itype; // initializer shape

    MatExpr() : op(OP_IDENTITY), flags(0), alpha(1), beta(0), irows(0), icols(0), itype(CV_64F) {}
    MatExpr(const Mat& m) : op(OP_IDENTITY), flags(0), a(m), alpha(1), beta(0), irows(0), icols(0), itype(CV_64F) {}
    operator Mat() const
    {
        Mat m;
        assign(m);
        return m;
    }
    void assign(Mat& m) const; // op->assign(*this, m)
    MatExpr t() const;
    MatExpr inv(int method = DECOMP_LU) const;
    Mat diag(int d = 0) const { return ((Mat) * this).diag(d); }
    Mat row(int r) const { return ((Mat) * this).row(r); }
    Mat col(int c) const { return ((Mat) * this).col(c); }
};

MatExpr operator+(const Mat& a, const Mat& b);
MatExpr operator+(const Mat& a, const MatExpr& e);
MatExpr operator+(const MatExpr& e, const Mat& b);
MatExpr operator+(const MatExpr& e1, const MatExpr& e2);
MatExpr operator-(const Mat& a, const Mat& b);
MatExpr operator-(const Mat& a, const MatExpr& e);
MatExpr operator-(const MatExpr& e, const Mat& b);
MatExpr operator-(const MatExpr& e1, const MatExpr& e2);
MatExpr operator-(const MatExpr& e, double s);
MatExpr operator-(const Mat& a, double s);
MatExpr operator*(const Mat& a, const Mat& b);
MatExpr operator*(const Mat& a, const MatExpr& e);
MatExpr operator*(const MatExpr& e, const Mat& b);
MatExpr operator*(const MatExpr& e1, const MatExpr& e2);
MatExpr operator*(double s, const Mat& a);
MatExpr operator*(const Mat& a, double s);
MatExpr operator*(double s, const MatExpr& e);
MatExpr operator*(const MatExpr& e, double s);
MatExpr operator/(const MatExpr& e, double s); // MatOp::divide(e, s) = multiply(e, 1. / s)
MatExpr operator/(const Mat& a, double s);

} // namespace cv

#endif
