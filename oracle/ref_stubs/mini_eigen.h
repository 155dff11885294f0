// mini_eigen.h -- the slice of Eigen 3.3 that /root/reference/src/{correlation_flow,loop_closure,utils,map,frame}.cc use,
// so that those files compile UNMODIFIED in an image without Eigen (oracle/_ref, see oracle/Makefile.ref).
//
// TEST INFRASTRUCTURE ONLY (same rule as nislam_oracle.c): nothing under ni_slam_b200/ may include or link this.
//
// Everything is evaluated eagerly (no expression templates).  Where Eigen's result depends on HOW it evaluates, this file
// follows Eigen 3.3.x built the way the reference builds it (-O3 -march=native on an AVX2 host, CMakeLists.txt:28-29):
//   * sum()/mean() of a real array: the linear vectorised reduction of Eigen/src/Core/Redux.h with 8-float packets, two
//     packet accumulators and the AVX/SSE3 horizontal add ((a0+a4)+(a1+a5))+((a2+a6)+(a3+a7));
//   * sum() of an expression that cannot be vectorised (|z^2| of a complex array): plain column-major running sum;
//   * complex * complex and complex / complex: the packet formulas of arch/AVX/Complex.h (pmul; pdiv = a*conj(b)/|b|^2),
//     each product and sum rounded separately (the recipe compiles with -ffp-contract=off);
//   * abs(complex) = std::abs = hypotf; pow(float array, int) = (float)std::pow(double, double) (scalar_pow_op is not
//     vectorised); exp() = std::exp(float) (Eigen's own pexp polynomial is NOT reproduced: the gaussian kernel is therefore
//     pinned in structure only);
//   * array / scalar is a true division by the scalar cast to the array's scalar type;
//   * maxCoeff(&row,&col): column-major visitor with strict '>' (first maximum wins).
#pragma once
#include <cmath>
#include <complex>
#include <cstddef>
#include <cstring>
#include <memory>
#include <ostream>
#include <type_traits>
#include <vector>

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW

namespace Eigen {

typedef std::ptrdiff_t Index;

template <class T> class aligned_allocator : public std::allocator<T> {
 public:
  template <class U> struct rebind { typedef aligned_allocator<U> other; };
  aligned_allocator() {}
  aligned_allocator(const aligned_allocator&) {}
  template <class U> aligned_allocator(const aligned_allocator<U>&) {}
};

namespace mini {
typedef std::complex<float> cf;
inline cf cmul(cf a, cf b) {          // arch/AVX/Complex.h pmul<Packet4cf>
  const float p0 = a.real() * b.real(), p1 = a.imag() * b.imag(), p2 = a.real() * b.imag(), p3 = a.imag() * b.real();
  return cf(p0 - p1, p2 + p3);
}
inline cf cdiv(cf a, cf b) {          // arch/AVX/Complex.h pdiv<Packet4cf>: pmul(a, pconj(b)) / (b.re^2 + b.im^2)
  const float p0 = a.real() * b.real(), p1 = a.imag() * b.imag(), p2 = a.imag() * b.real(), p3 = a.real() * b.imag();
  const float s0 = b.real() * b.real(), s1 = b.imag() * b.imag();
  const float den = s0 + s1;
  return cf((p0 + p1) / den, (p2 - p3) / den);
}
template <class T> struct is_cpx : std::false_type {};
template <class T> struct is_cpx<std::complex<T>> : std::true_type {};
template <class T> inline T mul(T a, T b) { return a * b; }
inline cf mul(cf a, cf b) { return cmul(a, b); }
template <class T> inline T div(T a, T b) { return a / b; }
inline cf div(cf a, cf b) { return cdiv(a, b); }
}  // namespace mini

template <class T> class Array2;

// a rectangular view used only as the two things the reference does with block(): read it in a sum, assign an array to it
template <class T> class Block2 {
 public:
  Block2(T* base, Index ld, Index r0, Index c0, Index nr, Index nc) : p_(base), ld_(ld), r0_(r0), c0_(c0), nr_(nr), nc_(nc) {}
  Index rows() const { return nr_; }
  Index cols() const { return nc_; }
  T get(Index r, Index c) const { return p_[(c0_ + c) * ld_ + r0_ + r]; }
  Block2& operator=(const Array2<typename std::remove_const<T>::type>& a) {
    for (Index c = 0; c < nc_; ++c)
      for (Index r = 0; r < nr_; ++r) p_[(c0_ + c) * ld_ + r0_ + r] = a(r, c);
    return *this;
  }
  Array2<typename std::remove_const<T>::type> operator+(const Block2& o) const {
    Array2<typename std::remove_const<T>::type> out(nr_, nc_);
    for (Index c = 0; c < nc_; ++c)
      for (Index r = 0; r < nr_; ++r) out(r, c) = get(r, c) + o.get(r, c);
    return out;
  }
 private:
  T* p_; Index ld_, r0_, c0_, nr_, nc_;
};

template <class T> class Array2 {
 public:
  typedef T Scalar;
  typedef Eigen::Index Index;
  typedef typename std::conditional<mini::is_cpx<T>::value, float, T>::type Real;

  Array2() : r_(0), c_(0) {}
  Array2(Index r, Index c) : r_(r), c_(c), d_((size_t)(r * c)) {}
  // real -> complex (Eigen 3.3 allows it: IFFT(fft_result.abs()) in correlation_flow.cc:92)
  template <class U, class = typename std::enable_if<!std::is_same<U, T>::value && mini::is_cpx<T>::value>::type>
  Array2(const Array2<U>& o) : r_(o.rows()), c_(o.cols()), d_((size_t)o.size()) {
    for (Index i = 0; i < size(); ++i) d_[(size_t)i] = T(o.data()[i]);
  }
  static Array2 Zero(Index r, Index c) { Array2 a(r, c); for (auto& v : a.d_) v = T(0); return a; }

  Index rows() const { return r_; }
  Index cols() const { return c_; }
  Index size() const { return r_ * c_; }
  T* data() { return d_.data(); }
  const T* data() const { return d_.data(); }
  void resize(Index r, Index c) { r_ = r; c_ = c; d_.resize((size_t)(r * c)); }
  T& operator()(Index r, Index c) { return d_[(size_t)(c * r_ + r)]; }
  const T& operator()(Index r, Index c) const { return d_[(size_t)(c * r_ + r)]; }

  Block2<T> block(Index r0, Index c0, Index nr, Index nc) { return Block2<T>(data(), r_, r0, c0, nr, nc); }
  Block2<const T> block(Index r0, Index c0, Index nr, Index nc) const { return Block2<const T>(data(), r_, r0, c0, nr, nc); }

  // ---- coefficient-wise unary
  template <class F> auto unary(F f) const -> Array2<decltype(f(T()))> {
    Array2<decltype(f(T()))> out(r_, c_);
    for (Index i = 0; i < size(); ++i) out.data()[i] = f(d_[(size_t)i]);
    return out;
  }
  Array2<Real> abs() const {
    Array2<Real> out = unary([](T v) -> Real { return std::abs(v); });
    out.set_novec(mini::is_cpx<T>::value);       // |complex| has no packet form: a reduction over it is not vectorised
    return out;
  }
  void set_novec(bool v) { novec_ = v; }
  Array2 square() const { return unary([](T v) -> T { return mini::mul(v, v); }); }
  Array2 conjugate() const { return unary([](T v) -> T { return conj_(v); }); }
  Array2 exp() const { return unary([](T v) -> T { return std::exp(v); }); }
  template <class E> Array2 pow(const E& e) const {
    return unary([e](T v) -> T { return (T)std::pow((double)v, (double)e); });
  }

  // ---- reductions
  T sum() const { return (mini::is_cpx<T>::value || novec_) ? seq_sum() : packet_sum(); }
  T seq_sum() const {                  // DefaultTraversal redux
    if (size() == 0) return T(0);
    T res = d_[0];
    for (Index i = 1; i < size(); ++i) res = res + d_[(size_t)i];
    return res;
  }
  T packet_sum() const {               // LinearVectorizedTraversal redux, PacketSize 8, no alignment peel (Eigen arrays are 32-byte aligned)
    const Index n = size(), P = 8;
    const Index a2 = (n / (2 * P)) * (2 * P), a1 = (n / P) * P;
    if (a1 == 0) return seq_sum();
    T p0[8], p1[8];
    for (Index k = 0; k < P; ++k) p0[k] = d_[(size_t)k];
    if (a1 > P) {
      for (Index k = 0; k < P; ++k) p1[k] = d_[(size_t)(P + k)];
      for (Index i = 2 * P; i < a2; i += 2 * P)
        for (Index k = 0; k < P; ++k) { p0[k] = p0[k] + d_[(size_t)(i + k)]; p1[k] = p1[k] + d_[(size_t)(i + P + k)]; }
      for (Index k = 0; k < P; ++k) p0[k] = p0[k] + p1[k];
      if (a1 > a2) for (Index k = 0; k < P; ++k) p0[k] = p0[k] + d_[(size_t)(a2 + k)];
    }
    T b[4];
    for (Index k = 0; k < 4; ++k) b[k] = p0[k] + p0[k + 4];
    T res = (b[0] + b[1]) + (b[2] + b[3]);
    for (Index i = a1; i < n; ++i) res = res + d_[(size_t)i];
    return res;
  }
  T mean() const { return sum() / T((Real)size()); }
  T maxCoeff() const {
    T m = d_[0];
    for (Index i = 1; i < size(); ++i) if (d_[(size_t)i] > m) m = d_[(size_t)i];
    return m;
  }
  template <class I> T maxCoeff(I* row, I* col) const {      // Visitor.h max_coeff_visitor: column-major, strict '>'
    Index best = 0;
    for (Index i = 1; i < size(); ++i) if (d_[(size_t)i] > d_[(size_t)best]) best = i;
    *row = (I)(best % r_); *col = (I)(best / r_);
    return d_[(size_t)best];
  }

  template <class F> Array2 zip(const Array2& o, F f) const {
    Array2 out(r_, c_);
    for (Index i = 0; i < size(); ++i) out.d_[(size_t)i] = f(d_[(size_t)i], o.d_[(size_t)i]);
    return out;
  }
  Array2 operator+(const Array2& o) const { return zip(o, [](T a, T b) { return a + b; }); }
  Array2 operator-(const Array2& o) const { return zip(o, [](T a, T b) { return a - b; }); }
  Array2 operator*(const Array2& o) const { return zip(o, [](T a, T b) { return mini::mul(a, b); }); }
  Array2 operator/(const Array2& o) const { return zip(o, [](T a, T b) { return mini::div(a, b); }); }
  // array (op) scalar: the scalar is cast to the array's (real) scalar type first, like Eigen's promote_scalar_arg
  template <class S, class = typename std::enable_if<std::is_arithmetic<S>::value>::type> Array2 operator+(S s) const {
    const Real v = (Real)s; return unary([v](T a) -> T { return a + v; });
  }
  template <class S, class = typename std::enable_if<std::is_arithmetic<S>::value>::type> Array2 operator-(S s) const {
    const Real v = (Real)s; return unary([v](T a) -> T { return a - v; });
  }
  template <class S, class = typename std::enable_if<std::is_arithmetic<S>::value>::type> Array2 operator*(S s) const {
    const Real v = (Real)s; return unary([v](T a) -> T { return a * v; });
  }
  template <class S, class = typename std::enable_if<std::is_arithmetic<S>::value>::type> Array2 operator/(S s) const {
    const Real v = (Real)s; return unary([v](T a) -> T { return a / v; });
  }
  Array2 operator-() const { return unary([](T a) -> T { return -a; }); }

 private:
  static float conj_(float v) { return v; }
  static double conj_(double v) { return v; }
  template <class U> static std::complex<U> conj_(std::complex<U> v) { return std::conj(v); }
  Index r_, c_;
  std::vector<T> d_;
  bool novec_ = false;
};

template <class S, class T, class = typename std::enable_if<std::is_arithmetic<S>::value>::type>
Array2<T> operator*(S s, const Array2<T>& a) { return a * s; }
template <class S, class T, class = typename std::enable_if<std::is_arithmetic<S>::value>::type>
Array2<T> operator+(S s, const Array2<T>& a) { return a + s; }
template <class S, class T, class = typename std::enable_if<std::is_arithmetic<S>::value>::type>
Array2<T> operator-(S s, const Array2<T>& a) {
  const typename Array2<T>::Real v = (typename Array2<T>::Real)s;
  return a.unary([v](T x) -> T { return v - x; });
}

typedef Array2<float> ArrayXXf;
typedef Array2<std::complex<float>> ArrayXXcf;

// MatrixXf: only the carrier between cv::cv2eigen / cv::eigen2cv and ArrayXXf (utils.cc:110-131)
class MatrixXf {
 public:
  MatrixXf() {}
  MatrixXf(const ArrayXXf& a) : a_(a) {}
  Index rows() const { return a_.rows(); }
  Index cols() const { return a_.cols(); }
  void resize(Index r, Index c) { a_.resize(r, c); }
  float& operator()(Index r, Index c) { return a_(r, c); }
  const float& operator()(Index r, Index c) const { return a_(r, c); }
  const ArrayXXf& array() const { return a_; }
 private:
  ArrayXXf a_;
};

// ---- small fixed-size matrices (Vector2d, Vector3d, Matrix2d, Matrix3d), column-major
template <class T, int R, int C> class Matrix;
template <class T, int R, int C> class TransposeView {
 public:
  explicit TransposeView(const Matrix<T, R, C>& m) : m_(m) {}
  const Matrix<T, R, C>& m_;
};
template <class T, int N> class HeadView {             // v.head(n) as an lvalue / rvalue of n leading coefficients
 public:
  HeadView(T* p, int n) : p_(p), n_(n) {}
  template <int M> HeadView& operator=(const Matrix<T, M, 1>& v) { for (int i = 0; i < n_; ++i) p_[i] = v[i]; return *this; }
  Matrix<T, 2, 1> eval2() const { return Matrix<T, 2, 1>(p_[0], p_[1]); }
  Matrix<T, 2, 1> operator-(const HeadView& o) const { return Matrix<T, 2, 1>(p_[0] - o.p_[0], p_[1] - o.p_[1]); }
  Matrix<T, 2, 1> operator+(const Matrix<T, 2, 1>& o) const { return Matrix<T, 2, 1>(p_[0] + o[0], p_[1] + o[1]); }
  T* p_; int n_;
};
template <class T, int R, int C> class Matrix {
 public:
  Matrix() { for (int i = 0; i < R * C; ++i) d_[i] = T(0); }
  Matrix(T a, T b) { static_assert(R * C == 2, "size"); d_[0] = a; d_[1] = b; }
  Matrix(T a, T b, T c) { static_assert(R * C == 3, "size"); d_[0] = a; d_[1] = b; d_[2] = c; }
  T& operator[](Index i) { return d_[i]; }
  const T& operator[](Index i) const { return d_[i]; }
  T& operator()(Index i) { return d_[i]; }
  const T& operator()(Index i) const { return d_[i]; }
  T& operator()(Index r, Index c) { return d_[c * R + r]; }
  const T& operator()(Index r, Index c) const { return d_[c * R + r]; }
  T sum() const { T s = d_[0]; for (int i = 1; i < R * C; ++i) s += d_[i]; return s; }
  TransposeView<T, R, C> transpose() const { return TransposeView<T, R, C>(*this); }
  HeadView<T, R> head(int n) { return HeadView<T, R>(d_, n); }
  Matrix operator-(const Matrix& o) const { Matrix m; for (int i = 0; i < R * C; ++i) m.d_[i] = d_[i] - o.d_[i]; return m; }
  Matrix operator+(const Matrix& o) const { Matrix m; for (int i = 0; i < R * C; ++i) m.d_[i] = d_[i] + o.d_[i]; return m; }
  T d_[R * C];
};
template <class T, int R, int C, int K> Matrix<T, R, K> operator*(const Matrix<T, R, C>& a, const Matrix<T, C, K>& b) {
  Matrix<T, R, K> m;
  for (int r = 0; r < R; ++r)
    for (int k = 0; k < K; ++k) { T s = T(0); for (int c = 0; c < C; ++c) s += a(r, c) * b(c, k); m(r, k) = s; }
  return m;
}
template <class T, int R, int C, int K> Matrix<T, C, K> operator*(const TransposeView<T, R, C>& a, const Matrix<T, R, K>& b) {
  Matrix<T, C, K> m;
  for (int c = 0; c < C; ++c)
    for (int k = 0; k < K; ++k) { T s = T(0); for (int r = 0; r < R; ++r) s += a.m_(r, c) * b(r, k); m(c, k) = s; }
  return m;
}
template <class T> Matrix<T, 2, 1> operator*(const Matrix<T, 2, 2>& a, const HeadView<T, 3>& h) { return a * h.eval2(); }
template <class T, int R, int C> std::ostream& operator<<(std::ostream& os, const TransposeView<T, R, C>& t) {
  for (int c = 0; c < C; ++c)
    for (int r = 0; r < R; ++r) os << (r || c ? " " : "") << t.m_(r, c);
  return os;
}
typedef Matrix<double, 2, 1> Vector2d;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 2, 2> Matrix2d;
typedef Matrix<double, 3, 3> Matrix3d;

}  // namespace Eigen
