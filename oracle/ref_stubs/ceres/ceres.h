// stand-in for <ceres/ceres.h>: the compiled reference files only reach ceres::floor / cos / sin on doubles
// (include/optimization_2d/normalize_angle.h:44-48, used by ComputeRelativePose / ComputeAbsolutePose in src/utils.cc).
#pragma once
#include <cmath>
namespace ceres {
inline double floor(double x) { return std::floor(x); }
inline double cos(double x) { return std::cos(x); }
inline double sin(double x) { return std::sin(x); }
}  // namespace ceres
