// SHADOWS /root/reference/include/optimization_2d/pose_graph_2d_error_term.h (Ceres autodiff cost functors; Ceres is not in this
// image).  The compiled files use only RotationMatrix2D (:43-52 of the original) and NormalizeAngle (the reference's own
// normalize_angle.h, included unmodified below through the ceres/ceres.h stand-in).
#pragma once
#include <Eigen/Core>
#include "optimization_2d/normalize_angle.h"
namespace ceres {
namespace optimization_2d {
template <typename T> Eigen::Matrix<T, 2, 2> RotationMatrix2D(T yaw_radians) {
  const T cos_yaw = ceres::cos(yaw_radians);
  const T sin_yaw = ceres::sin(yaw_radians);
  Eigen::Matrix<T, 2, 2> rotation;
  rotation(0, 0) = cos_yaw; rotation(0, 1) = -sin_yaw; rotation(1, 0) = sin_yaw; rotation(1, 1) = cos_yaw;
  return rotation;
}
}  // namespace optimization_2d
}  // namespace ceres
