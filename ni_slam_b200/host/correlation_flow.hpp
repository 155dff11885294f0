// correlation_flow.hpp -- header-only C++ shim that re-exposes the reference's host classes on top of the C ABI
// (include/nislam.h), so that src/map_builder.cc compiles unchanged against the B200 path.
//
//   reference                                              this shim
//   -----------------------------------------------------  ---------------------------------------------------------
//   class CorrelationFlow   include/correlation_flow.h:8    nislam::CorrelationFlowT<ArrayXXf, ArrayXXcf, Vector3d>
//   class LoopClosure       include/loop_closure.h:27       nislam::LoopClosureT<..., FramePtr, MapPtr, GridLocation> (the three reference signatures)
//   struct CFConfig         include/read_configs.h:15       nislam::CFConfig (same fields, same order)
//   struct LoopClosureConfig include/read_configs.h:38      nislam::LoopClosureConfig
//   struct LoopClosureResult include/loop_closure.h:8       nislam::LoopClosureResultT<Vector3d, FramePtr> (found, response, current_frame, loop_frame, relative_pose)
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
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdlib>
#include <map>
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
  ~CorrelationFlowT() { if (last_) nis_frame_free(ctx_, last_); nis_destroy(ctx_); }
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
  // The reference passes bare arrays.  Only what ComputePose reads is moved: (last_fft_result, last_fft_polar) of the last frame
  // -- kept on the device between calls, because MapBuilder::Tracking passes the SAME keyframe arrays for every frame until the next
  // keyframe (map_builder.cc:99-106, :129) -- and (image, fft_polar) of the current one; the layout conversion runs on the GPU.
  Vector3d ComputePose(const ArrayXXcf& last_fft_result, const ArrayXXf& image, const ArrayXXcf& last_fft_polar,
                       const ArrayXXcf& fft_polar, Vector3d& pose, bool not_large_rotation) {
    const uint64_t key = fingerprint(last_fft_result) * 1099511628211ull ^ fingerprint(last_fft_polar);
    if (!last_ || !cache_last_ || key != last_key_ || last_fft_result.data() != last_ptr_) {
      if (last_) { nis_frame_free(ctx_, last_); last_ = nullptr; }
      check(ctx_, nis_frame_import_ex(ctx_, nullptr, reinterpret_cast<const float*>(last_fft_result.data()),
                                      reinterpret_cast<const float*>(last_fft_polar.data()), 1, &last_));
      last_key_ = key; last_ptr_ = last_fft_result.data();
    }
    nis_frame* cur = nullptr;
    check(ctx_, nis_frame_import_ex(ctx_, image.data(), nullptr, reinterpret_cast<const float*>(fft_polar.data()), 0, &cur));
    double p[3] = {0, 0, 0}, info[3] = {0, 0, 0};
    int st = nis_compute_pose(ctx_, last_, cur, not_large_rotation ? 1 : 0, p, info, nullptr);
    nis_frame_free(ctx_, cur);
    check(ctx_, st);
    Vector3d out;
    for (int i = 0; i < 3; ++i) { pose[i] = p[i]; out[i] = info[i]; }
    return out;
  }
  // The device copy of the last keyframe's operands is reused when the caller passes the same buffer with the same content
  // (fingerprint = FNV-1a over 2048 samples spread over the array plus its size); SetOperandCache(false) re-uploads every call.
  void SetOperandCache(bool on) { cache_last_ = on; }

  nis_ctx* handle() const { return ctx_; }
  const CFConfig& config() const { return cfg; }

 private:
  static uint64_t fingerprint(const ArrayXXcf& a) {
    const size_t n = (size_t)a.rows() * (size_t)a.cols();
    const uint64_t* w = reinterpret_cast<const uint64_t*>(a.data());       // one complex<float> = 8 bytes
    uint64_t h = 1469598103934665603ull ^ n;
    const size_t step = n > 2048 ? n / 2048 : 1;
    for (size_t i = 0; i < n; i += step) { h ^= w[i]; h *= 1099511628211ull; }
    return h;
  }
  CFConfig cfg;
  nis_ctx* ctx_ = nullptr;
  nis_frame* last_ = nullptr;
  uint64_t last_key_ = 0;
  const void* last_ptr_ = nullptr;
  bool cache_last_ = true;
};

// LoopClosureResult (include/loop_closure.h:8-25), same fields, FramePtr a template parameter
template <class Vector3d, class FramePtr>
struct LoopClosureResultT {
  bool found;
  Vector3d response;
  FramePtr current_frame;
  FramePtr loop_frame;
  Vector3d relative_pose;
  LoopClosureResultT() : found(false), current_frame(), loop_frame() { for (int i = 0; i < 3; ++i) { response[i] = -1.0; relative_pose[i] = 0.0; } }
};

// LoopClosure (include/loop_closure.h:27-38) with the reference's three signatures.  Frame / Map are the CALLER'S classes (template
// parameters; the aliases at the bottom bind the reference's when its headers are present): candidates, frame ids and accumulated
// distances are read through their own getters exactly as src/loop_closure.cc does, so candidate order -- including the unordered_set
// order of GetFramesInGrids -- and the filters are the reference's.  A candidate's spectra are uploaded the first time the frame is
// seen (Frame::GetFFTResult, src/frame.cc:53-57) and then live in the GPU keyframe store; a scan moves the query image and its
// fft_polar only.  Frames are assumed immutable once constructed, as in the reference (MapBuilder never calls SetFFTResult).
template <class ArrayXXf, class ArrayXXcf, class Vector3d, class FramePtr, class MapPtr, class GridLocation>
class LoopClosureT {
 public:
  typedef CorrelationFlowT<ArrayXXf, ArrayXXcf, Vector3d> CF;
  typedef LoopClosureResultT<Vector3d, FramePtr> Result;
  // LoopClosure(LoopClosureConfig&, CorrelationFlowPtr, MapPtr)   include/loop_closure.h:29
  LoopClosureT(LoopClosureConfig& loop_closure_config, std::shared_ptr<CF> correlation_flow, MapPtr map)
      : _loop_thr(loop_closure_config), _correlation_flow(correlation_flow), _map(map) {}

  // FindLoopClosure(image, current_frame): all frames of the map in id order   src/loop_closure.cc:10-15
  Result FindLoopClosure(ArrayXXf& image, FramePtr& current_frame) {
    std::vector<FramePtr> frames;
    _map->GetAllFrames(frames);
    return FindLoopClosure(image, current_frame, frames);
  }
  // FindLoopClosure(image, current_frame, prior_pose): the frames filed in the 3 x 3 grid cells around the prior   :17-34
  Result FindLoopClosure(ArrayXXf& image, FramePtr& current_frame, Vector3d& prior_pose) {
    std::vector<GridLocation> grid_locations;
    GridLocation grid_location = _map->ComputeGridLocation(prior_pose);
    for (int i = -1; i <= 1; i++) {
      for (int j = -1; j <= 1; j++) {
        GridLocation gl = grid_location;
        gl.x += i;
        gl.y += j;
        grid_locations.emplace_back(gl);
      }
    }
    std::vector<FramePtr> frames;
    _map->GetFramesInGrids(frames, grid_locations);
    return FindLoopClosure(image, current_frame, frames);
  }
  // FindLoopClosure(image, current_frame, frames)   :36-73
  Result FindLoopClosure(ArrayXXf& image, FramePtr& current_frame, std::vector<FramePtr>& frames) {
    nis_ctx* ctx = _correlation_flow->handle();
    Result result;
    result.current_frame = current_frame;
    std::vector<int32_t> slots;
    std::vector<FramePtr> kept;
    for (FramePtr frame : frames) {
      if (_loop_thr.frame_gap_thr > 0 && std::abs((current_frame->GetFrameId() - frame->GetFrameId())) < _loop_thr.frame_gap_thr) continue;
      if (_loop_thr.distance_thr > 0) {
        double d1 = _map->GetFrameDistance(current_frame);
        double d2 = _map->GetFrameDistance(frame);
        if (std::abs((d1 - d2)) < _loop_thr.distance_thr) continue;
      }
      slots.push_back(SlotOf(frame));
      kept.push_back(frame);
    }
    if (!slots.empty()) {
      ArrayXXcf current_fft_result, current_fft_polar;
      current_frame->GetFFTResult(current_fft_result, current_fft_polar);
      nis_frame* q = nullptr;
      check(ctx, nis_frame_import_ex(ctx, image.data(), nullptr, reinterpret_cast<const float*>(current_fft_polar.data()), 0, &q));
      nis_loop_config c{_loop_thr.position_response_thr, _loop_thr.angle_response_thr, 0, 0.0};     // filters already applied above
      nis_loop_result r;
      int st = nis_loop_scan(ctx, q, current_frame->GetFrameId(), 0.0, &c, slots.data(), (int)slots.size(), &r, nullptr);
      nis_frame_free(ctx, q);
      check(ctx, st);
      if (r.slot >= 0) {
        for (size_t i = 0; i < slots.size(); ++i)
          if (slots[i] == r.slot) { result.loop_frame = kept[i]; break; }       // first in iteration order, like the strict '>' of :61
        for (int i = 0; i < 3; ++i) { result.response[i] = r.response[i]; result.relative_pose[i] = r.relative_pose[i]; }
      }
    }
    bool c1 = (result.response[0] > _loop_thr.position_response_thr);
    bool c2 = (result.response[2] > _loop_thr.angle_response_thr);
    result.found = (c1 && c2);
    return result;
  }
  int StoredFrames() const { return (int)_slot_of.size(); }

 private:
  int32_t SlotOf(FramePtr& frame) {
    auto it = _slot_of.find((const void*)&*frame);
    if (it != _slot_of.end()) return it->second;
    nis_ctx* ctx = _correlation_flow->handle();
    ArrayXXcf fft_result, fft_polar;
    frame->GetFFTResult(fft_result, fft_polar);
    int slot = -1;
    check(ctx, nis_db_add_spectra(ctx, reinterpret_cast<const float*>(fft_result.data()), reinterpret_cast<const float*>(fft_polar.data()),
                                  frame->GetFrameId(), 0.0, &slot));
    _slot_of[(const void*)&*frame] = slot;
    _held.push_back(frame);             // keeps the pointer identity valid for the lifetime of the store
    return slot;
  }
  LoopClosureConfig _loop_thr;
  std::shared_ptr<CF> _correlation_flow;
  MapPtr _map;
  std::map<const void*, int32_t> _slot_of;
  std::vector<FramePtr> _held;
};

}  // namespace nislam

#if defined(__has_include)
#if __has_include(<Eigen/Core>)
#include <Eigen/Core>
// the reference's names, for src/map_builder.cc
typedef nislam::CFConfig CFConfig;
typedef nislam::LoopClosureConfig LoopClosureConfig;
typedef nislam::CorrelationFlowT<Eigen::ArrayXXf, Eigen::ArrayXXcf, Eigen::Vector3d> CorrelationFlow;
// Frame / Map / GridLocation are the reference's own (include/frame.h, include/map.h): include those before this header
typedef nislam::LoopClosureT<Eigen::ArrayXXf, Eigen::ArrayXXcf, Eigen::Vector3d, FramePtr, MapPtr, GridLocation> LoopClosure;
typedef nislam::LoopClosureResultT<Eigen::Vector3d, FramePtr> LoopClosureResult;
typedef std::shared_ptr<CorrelationFlow> CorrelationFlowPtr;
typedef std::shared_ptr<LoopClosure> LoopClosurePtr;
#endif
#endif
