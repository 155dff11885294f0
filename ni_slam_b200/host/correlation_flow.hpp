// correlation_flow.hpp -- header-only C++ shim that re-exposes the reference's host classes on top of the C ABI
// (include/nislam.h), so that src/map_builder.cc compiles unchanged against the B200 path.
//
//   reference                                              this shim
//   -----------------------------------------------------  ---------------------------------------------------------
//   class CorrelationFlow   include/correlation_flow.h:8    nislam::CorrelationFlowT<ArrayXXf, ArrayXXcf, Vector3d>
//   class LoopClosure       include/loop_closure.h:27       nislam::LoopClosureT<...>
//   struct CFConfig         include/read_configs.h:15       nislam::CFConfig (same fields, same order)
//   struct LoopClosureConfig include/read_configs.h:38      nislam::LoopClosureConfig
//   struct LoopClosureResult include/loop_closure.h:8       nislam::LoopClosureResultT<Vector3d>
//
// The array types are template parameters: with Eigen present (`__has_include(<Eigen/Core>)`) the aliases at the
// bottom instantiate them with Eigen::ArrayXXf / ArrayXXcf / Vector3d and the class names are the reference's.
// Any column-major container with rows(), cols(), data(), resize(r, c) works (tests/cpp/shim_test.cc uses a
// 30-line stand-in because Eigen is not installed in the build image).
//
// Error behaviour mirrors the reference: an invalid kernel id throws std::invalid_argument("Received invalid kernel
// type") from ComputePose (src/correlation_flow.cc:168); any other failure throws std::runtime_error (the reference
// has undefined behaviour there).  There is no CPU fallback: without a CUDA device the constructor throws.
#pragma once
#include <complex>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/nislam.h"

namespace nislam {

struct CFConfig {          // include/read_configs.h:15-25
  int width;
  int height;
  float lambda;
  int kernel;
  float sigma;
  float offset;
  int power;
  int rotation_divisor;
  int rotation_channel;
};

struct LoopClosureConfig { // include/read_configs.h:38-44
  bool to_find_loop;
  double position_response_thr;
  double angle_response_thr;
  int frame_gap_thr;
  double distance_thr;
};

inline void check(nis_ctx* ctx, int st) {
  if (st == NIS_OK) return;
  if (st == NIS_ERR_INVALID_KERNEL) throw std::invalid_argument("Received invalid kernel type");
  std::string msg = ctx ? nis_last_error(ctx) : "";
  if (msg.empty()) msg = nis_strerror(st);
  throw std::runtime_error("libnislam: " + msg);
}

template <class ArrayXXf, class ArrayXXcf, class Vector3d>
class CorrelationFlowT {
 public:
  // CorrelationFlow(CFConfig& cf_config, double& image_height, double& image_width)   include/correlation_flow.h:11
  CorrelationFlowT(CFConfig& cf_config, double& image_height, double& image_width, int device = 0) : cfg(cf_config) {
    cfg.height = int(image_height);      // src/correlation_flow.cc:40-41
    cfg.width = int(image_width);
    nis_cf_config c{cfg.lambda, cfg.kernel, cfg.sigma, cfg.offset, cfg.power, cfg.rotation_divisor, cfg.rotation_channel};
    check(nullptr, nis_create(&c, cfg.height, cfg.width, device, &ctx_));
  }
  ~CorrelationFlowT() { nis_destroy(ctx_); }
  CorrelationFlowT(const CorrelationFlowT&) = delete;
  CorrelationFlowT& operator=(const CorrelationFlowT&) = delete;

  // void ComputeIntermedium(const ArrayXXf& image, ArrayXXcf& fft_result, ArrayXXcf& fft_polar)   correlation_flow.h:12
  void ComputeIntermedium(const ArrayXXf& image, ArrayXXcf& fft_result, ArrayXXcf& fft_polar) {
    nis_frame* f = nullptr;
    check(ctx_, nis_features_f32(ctx_, image.data(), &f));
    fft_result.resize(cfg.height / 2 + 1, cfg.width);
    fft_polar.resize(cfg.rotation_divisor / 2 + 1, cfg.rotation_channel);
    int st = nis_frame_export(ctx_, f, reinterpret_cast<float*>(fft_result.data()), reinterpret_cast<float*>(fft_polar.data()));
    nis_frame_free(ctx_, f);
    check(ctx_, st);
  }

  // Optional GPU front end for MapBuilder::AddNewInput (src/map_builder.cc:31-33): hand over Camera's fixed-point maps once
  // (_map1.ptr<short>(), _map2.ptr<ushort>() of initUndistortRectifyMap(..., CV_16SC2), src/camera.cc:45-47) and feed the RAW
  // u8 cv::Mat; Camera::UndistortImage + ConvertMatToNormalizedArray + ComputeIntermedium then all run on the GPU.
  void SetUndistortMaps(const int16_t* map1_xy, const uint16_t* map2) { check(ctx_, nis_set_undistort_maps(ctx_, map1_xy, map2)); }
  void ComputeIntermediumU8(const uint8_t* image_rowmajor, ArrayXXcf& fft_result, ArrayXXcf& fft_polar) {
    nis_frame* f = nullptr;
    check(ctx_, nis_features_u8(ctx_, image_rowmajor, &f));
    fft_result.resize(cfg.height / 2 + 1, cfg.width);
    fft_polar.resize(cfg.rotation_divisor / 2 + 1, cfg.rotation_channel);
    int st = nis_frame_export(ctx_, f, reinterpret_cast<float*>(fft_result.data()), reinterpret_cast<float*>(fft_polar.data()));
    nis_frame_free(ctx_, f);
    check(ctx_, st);
  }

  // Vector3d ComputePose(last_fft_result, image, last_fft_polar, fft_polar, pose, not_large_rotation)   correlation_flow.h:13
  Vector3d ComputePose(const ArrayXXcf& last_fft_result, const ArrayXXf& image, const ArrayXXcf& last_fft_polar,
                       const ArrayXXcf& fft_polar, Vector3d& pose, bool not_large_rotation) {
    // the reference passes bare arrays; wrap them as device frames (the "last" frame needs no image, the current frame no fft_result)
    nis_frame *last = nullptr, *cur = nullptr;
    check(ctx_, nis_frame_import(ctx_, image.data(), reinterpret_cast<const float*>(last_fft_result.data()),
                                 reinterpret_cast<const float*>(last_fft_polar.data()), &last));
    int st = nis_frame_import(ctx_, image.data(), reinterpret_cast<const float*>(last_fft_result.data()),
                              reinterpret_cast<const float*>(fft_polar.data()), &cur);
    double p[3] = {0, 0, 0}, info[3] = {0, 0, 0};
    if (st == NIS_OK) st = nis_compute_pose(ctx_, last, cur, not_large_rotation ? 1 : 0, p, info, nullptr);
    nis_frame_free(ctx_, last);
    nis_frame_free(ctx_, cur);
    check(ctx_, st);
    Vector3d out;
    for (int i = 0; i < 3; ++i) { pose[i] = p[i]; out[i] = info[i]; }
    return out;
  }

  nis_ctx* handle() const { return ctx_; }
  const CFConfig& config() const { return cfg; }

 private:
  CFConfig cfg;
  nis_ctx* ctx_ = nullptr;
};

// LoopClosureResult (include/loop_closure.h:8-25); FramePtr -> slot / frame id of the GPU keyframe store
template <class Vector3d>
struct LoopClosureResultT {
  bool found = false;
  Vector3d response;
  int current_frame_id = -1;
  int loop_slot = -1;
  int loop_frame_id = -1;
  Vector3d relative_pose;
  LoopClosureResultT() { for (int i = 0; i < 3; ++i) { response[i] = -1.0; relative_pose[i] = 0.0; } }
};

// LoopClosure (include/loop_closure.h:27-38).  The reference reads the candidates' spectra out of Map/Frame; here the
// spectra live in the GPU keyframe store, filled by AddFrame (call it where MapBuilder calls _map->AddFrame,
// src/map_builder.cc:60) so a scan never moves 2.6 MB per candidate across PCIe.
template <class ArrayXXf, class ArrayXXcf, class Vector3d>
class LoopClosureT {
 public:
  typedef CorrelationFlowT<ArrayXXf, ArrayXXcf, Vector3d> CF;
  typedef LoopClosureResultT<Vector3d> Result;
  LoopClosureT(LoopClosureConfig& loop_closure_config, std::shared_ptr<CF> correlation_flow)
      : _loop_thr(loop_closure_config), _correlation_flow(correlation_flow) {}

  // Map::AddFrame + Map::SetFrameDistance for the arrays the scan reads
  int AddFrame(int frame_id, const ArrayXXf& image, const ArrayXXcf& fft_result, const ArrayXXcf& fft_polar, double acc_distance) {
    nis_ctx* ctx = _correlation_flow->handle();
    nis_frame* f = nullptr;
    check(ctx, nis_frame_import(ctx, image.data(), reinterpret_cast<const float*>(fft_result.data()),
                                reinterpret_cast<const float*>(fft_polar.data()), &f));
    int slot = -1;
    int st = nis_db_add(ctx, f, frame_id, acc_distance, &slot);
    nis_frame_free(ctx, f);
    check(ctx, st);
    return slot;
  }

  // FindLoopClosure(image, current_frame)  -- all frames in id order (src/loop_closure.cc:10-15)
  Result FindLoopClosure(const ArrayXXf& image, int current_frame_id, const ArrayXXcf& current_fft_result,
                         const ArrayXXcf& current_fft_polar, double current_distance) {
    return Scan(image, current_frame_id, current_fft_result, current_fft_polar, current_distance, nullptr, 0);
  }
  // Map::AddFrame's grid filing for slot (src/map.cc:27-30) and FindLoopClosure(image, current_frame, prior_pose) (:17-34)
  void SetPosition(int slot, const Vector3d& pose, double grid_scale) {
    nis_ctx* ctx = _correlation_flow->handle();
    check(ctx, nis_db_set_position(ctx, slot, pose[0], pose[1], grid_scale));
  }
  Result FindLoopClosure(const ArrayXXf& image, int current_frame_id, const ArrayXXcf& current_fft_result,
                         const ArrayXXcf& current_fft_polar, double current_distance, const Vector3d& prior_pose, double grid_scale) {
    nis_ctx* ctx = _correlation_flow->handle();
    nis_frame* q = nullptr;
    check(ctx, nis_frame_import(ctx, image.data(), reinterpret_cast<const float*>(current_fft_result.data()),
                                reinterpret_cast<const float*>(current_fft_polar.data()), &q));
    nis_loop_config c{_loop_thr.position_response_thr, _loop_thr.angle_response_thr, _loop_thr.frame_gap_thr, _loop_thr.distance_thr};
    nis_loop_result r;
    int st = nis_loop_scan_prior(ctx, q, current_frame_id, current_distance, &c, prior_pose[0], prior_pose[1], grid_scale, &r, nullptr, 0, nullptr);
    nis_frame_free(ctx, q);
    check(ctx, st);
    Result out;
    out.found = r.found != 0; out.current_frame_id = current_frame_id; out.loop_slot = r.slot; out.loop_frame_id = r.frame_id;
    for (int i = 0; i < 3; ++i) { out.response[i] = r.response[i]; out.relative_pose[i] = r.relative_pose[i]; }
    return out;
  }
  // FindLoopClosure(image, current_frame, frames) -- explicit candidate list, iteration order = list order (:36-73)
  Result FindLoopClosure(const ArrayXXf& image, int current_frame_id, const ArrayXXcf& current_fft_result,
                         const ArrayXXcf& current_fft_polar, double current_distance, const std::vector<int32_t>& candidate_slots) {
    return Scan(image, current_frame_id, current_fft_result, current_fft_polar, current_distance, candidate_slots.data(),
                (int)candidate_slots.size());
  }

 private:
  Result Scan(const ArrayXXf& image, int id, const ArrayXXcf& F, const ArrayXXcf& P, double dist, const int32_t* cand, int n) {
    nis_ctx* ctx = _correlation_flow->handle();
    nis_frame* q = nullptr;
    check(ctx, nis_frame_import(ctx, image.data(), reinterpret_cast<const float*>(F.data()), reinterpret_cast<const float*>(P.data()), &q));
    nis_loop_config c{_loop_thr.position_response_thr, _loop_thr.angle_response_thr, _loop_thr.frame_gap_thr, _loop_thr.distance_thr};
    nis_loop_result r;
    int st = nis_loop_scan(ctx, q, id, dist, &c, cand, n, &r, nullptr);
    nis_frame_free(ctx, q);
    check(ctx, st);
    Result out;
    out.found = r.found != 0;
    out.current_frame_id = id;
    out.loop_slot = r.slot;
    out.loop_frame_id = r.frame_id;
    for (int i = 0; i < 3; ++i) { out.response[i] = r.response[i]; out.relative_pose[i] = r.relative_pose[i]; }
    return out;
  }
  LoopClosureConfig _loop_thr;
  std::shared_ptr<CF> _correlation_flow;
};

}  // namespace nislam

#if defined(__has_include)
#if __has_include(<Eigen/Core>)
#include <Eigen/Core>
// the reference's names, for src/map_builder.cc
typedef nislam::CFConfig CFConfig;
typedef nislam::LoopClosureConfig LoopClosureConfig;
typedef nislam::CorrelationFlowT<Eigen::ArrayXXf, Eigen::ArrayXXcf, Eigen::Vector3d> CorrelationFlow;
typedef nislam::LoopClosureT<Eigen::ArrayXXf, Eigen::ArrayXXcf, Eigen::Vector3d> LoopClosure;
typedef nislam::LoopClosureResultT<Eigen::Vector3d> LoopClosureResult;
typedef std::shared_ptr<CorrelationFlow> CorrelationFlowPtr;
typedef std::shared_ptr<LoopClosure> LoopClosurePtr;
#endif
#endif
