// mini_cv.h -- stand-in for the OpenCV 4.2 calls on the path of /root/reference/src/{correlation_flow,utils}.cc so that
// those files compile UNMODIFIED in an image without OpenCV's C++ library (oracle/_ref, see oracle/Makefile.ref).
// TEST INFRASTRUCTURE ONLY.
//   cv::warpPolar (correlation_flow.cc:234), cv::getRotationMatrix2D + cv::warpAffine (utils.cc:158-159, :168),
//   cv::cv2eigen / cv::eigen2cv (utils.cc:116-131), cv::Mat / Point2f / Size, imshow / waitKey (no-ops).
// The two warps run the fixed-point arithmetic of oracle/nislam_oracle.c (orc_warp_polar_rm / orc_warp_affine_inv_rm), which
// tests/test_oracle.py holds bit-exact (warpPolar) / within 1 ulp (warpAffine) against genuine cv2 4.13 vectors; the matrix set-up
// (getRotationMatrix2D, the inversion inside warpAffine) is written here from OpenCV's published formulas (imgwarp.cpp).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

extern "C" {
void orc_warp_polar_rm(const float* src, int H, int W, int D, int Cp, float cx, float cy, double maxRadius, float* dst);
void orc_warp_affine_inv_rm(const float* src, int H, int W, const double iM[6], float* dst);
}

#define CV_8U 0
#define CV_32F 5
#define CV_64F 6
#ifndef CV_PI
#define CV_PI 3.1415926535897932384626433832795
#endif

namespace cv {

enum { INTER_LINEAR = 1, WARP_FILL_OUTLIERS = 8, WARP_INVERSE_MAP = 16, WARP_POLAR_LOG = 256 };
enum { BORDER_CONSTANT = 0, BORDER_WRAP = 3 };

struct Size {
  int width, height;
  Size() : width(0), height(0) {}
  Size(int w, int h) : width(w), height(h) {}
};
struct Point2f {
  float x, y;
  Point2f() : x(0), y(0) {}
  Point2f(float x_, float y_) : x(x_), y(y_) {}
};

class Mat {                                  // single-channel, row-major, owning (shared) or wrapping user data
 public:
  int rows = 0, cols = 0;
  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(int r, int c, int type, void* user) : rows(r), cols(c), type_(type), ext_((unsigned char*)user) {}
  void create(int r, int c, int type) {
    rows = r; cols = c; type_ = type; ext_ = nullptr;
    buf_ = std::make_shared<std::vector<unsigned char>>((size_t)r * c * elem());
  }
  int type() const { return type_; }
  bool empty() const { return rows == 0 || cols == 0; }
  Size size() const { return Size(cols, rows); }
  size_t elem() const { return type_ == CV_8U ? 1 : (type_ == CV_32F ? 4 : 8); }
  unsigned char* data() const { return ext_ ? ext_ : (buf_ ? buf_->data() : nullptr); }
  template <class T> T& at(int r, int c) { return ((T*)data())[(size_t)r * cols + c]; }
  template <class T> const T& at(int r, int c) const { return ((const T*)data())[(size_t)r * cols + c]; }
  double get(int r, int c) const {
    if (type_ == CV_8U) return at<unsigned char>(r, c);
    if (type_ == CV_32F) return at<float>(r, c);
    return at<double>(r, c);
  }
 private:
  int type_ = CV_32F;
  std::shared_ptr<std::vector<unsigned char>> buf_;
  unsigned char* ext_ = nullptr;
};

// imgwarp.cpp cv::getRotationMatrix2D: angle in degrees, 2x3 CV_64F
inline Mat getRotationMatrix2D(Point2f center, double angle, double scale) {
  angle *= CV_PI / 180;
  const double alpha = std::cos(angle) * scale, beta = std::sin(angle) * scale;
  Mat M(2, 3, CV_64F);
  double* m = (double*)M.data();
  m[0] = alpha; m[1] = beta; m[2] = (1 - alpha) * center.x - beta * center.y;
  m[3] = -beta; m[4] = alpha; m[5] = beta * center.x + (1 - alpha) * center.y;
  return M;
}

// imgwarp.cpp cv::warpAffine: M -> double, inverted unless WARP_INVERSE_MAP, then the 10+5-bit fixed-point bilinear walk
inline void warpAffine(const Mat& src, Mat& dst, const Mat& M0, Size dsize, int flags = INTER_LINEAR, int borderMode = BORDER_CONSTANT) {
  if (src.type() != CV_32F || (flags & 7) != INTER_LINEAR || borderMode != BORDER_WRAP || dsize.width != src.cols || dsize.height != src.rows)
    throw std::runtime_error("mini_cv::warpAffine: only f32, INTER_LINEAR, BORDER_WRAP, same size (what the reference calls)");
  double M[6];
  for (int i = 0; i < 6; ++i) M[i] = M0.get(i / 3, i % 3);
  if (!(flags & WARP_INVERSE_MAP)) {
    double D = M[0] * M[4] - M[1] * M[3];
    D = D != 0 ? 1. / D : 0;
    const double A11 = M[4] * D, A22 = M[0] * D;
    M[0] = A11; M[1] *= -D; M[3] *= -D; M[4] = A22;
    const double b1 = -M[0] * M[2] - M[1] * M[5];
    const double b2 = -M[3] * M[2] - M[4] * M[5];
    M[2] = b1; M[5] = b2;
  }
  Mat out(src.rows, src.cols, CV_32F);
  orc_warp_affine_inv_rm((const float*)src.data(), src.rows, src.cols, M, (float*)out.data());
  dst = out;
}

// imgwarp.cpp cv::warpPolar (linear): remap with mapx = rho*Kmag*cos(phi*Kangle)+cx, mapy likewise, BORDER_CONSTANT(0)
inline void warpPolar(const Mat& src, Mat& dst, Size dsize, Point2f center, double maxRadius, int flags) {
  if (src.type() != CV_32F || (flags & 7) != INTER_LINEAR || (flags & WARP_POLAR_LOG) || (flags & WARP_INVERSE_MAP))
    throw std::runtime_error("mini_cv::warpPolar: only f32, INTER_LINEAR, linear, forward (what the reference calls)");
  Mat out(dsize.height, dsize.width, CV_32F);
  orc_warp_polar_rm((const float*)src.data(), src.rows, src.cols, dsize.height, dsize.width, center.x, center.y, maxRadius, (float*)out.data());
  dst = out;
}

inline void imshow(const std::string&, const Mat&) {}
inline int waitKey(int = 0) { return 0; }

// core/eigen.hpp: value-preserving transposing copies between row-major cv::Mat and column-major Eigen storage
template <class MatrixT> void cv2eigen(const Mat& src, MatrixT& dst) {
  dst.resize(src.rows, src.cols);
  for (int r = 0; r < src.rows; ++r)
    for (int c = 0; c < src.cols; ++c) dst(r, c) = (float)src.get(r, c);
}
template <class MatrixT> void eigen2cv(const MatrixT& src, Mat& dst) {
  dst.create((int)src.rows(), (int)src.cols(), CV_32F);
  float* p = (float*)dst.data();
  for (int r = 0; r < dst.rows; ++r)
    for (int c = 0; c < dst.cols; ++c) p[(size_t)r * dst.cols + c] = src(r, c);
}

}  // namespace cv
