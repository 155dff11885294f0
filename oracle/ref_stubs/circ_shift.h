// SHADOWS /root/reference/include/circ_shift.h (an Eigen expression-template view, 252 lines of Eigen internals that the
// mini Eigen cannot host).  Only fftshift is used by the compiled files (correlation_flow.cc:94); the view's coefficient
// rule (circ_shift.h:130-154 with the shifts of :238-244) is  out(r, c) = in((r - R/2) mod R, (c - C/2) mod C).
#pragma once
#include <Eigen/Core>
template <typename T> Eigen::Array2<T> fftshift(Eigen::Array2<T>& x) {
  const Eigen::Index R = x.rows(), C = x.cols(), rs = R / 2, cs = C / 2;
  Eigen::Array2<T> out(R, C);
  for (Eigen::Index c = 0; c < C; ++c)
    for (Eigen::Index r = 0; r < R; ++r) out(r, c) = x(((r - rs) % R + R) % R, ((c - cs) % C + C) % C);
  return out;
}
