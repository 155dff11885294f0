// Drives the reference-facing C++ shim (ni_slam_b200/host/correlation_flow.hpp) the way MapBuilder drives the reference
// classes (src/map_builder.cc:72-75, :127-138, :172-182), with a stand-in for the Eigen array types (Eigen is not in the
// build image).  Known answers: SURVEY.md Appendix C.4/C.5/C.7 (circular rolls, identity, first-wins scan).
// Build: g++ -std=c++17 tests/cpp/shim_test.cc -L ni_slam_b200/lib -lnislam -Wl,-rpath,... ; needs a GPU to run.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <map>
#include <memory>
#include <set>
#include <vector>

#include "../../ni_slam_b200/host/correlation_flow.hpp"
#include "../../ni_slam_b200/host/map_stitcher.hpp"

template <class T> struct ColMajor {           // minimal Eigen::Array<T,Dynamic,Dynamic> stand-in (column-major)
  int r = 0, c = 0;
  std::vector<T> v;
  void resize(int rows, int cols) { r = rows; c = cols; v.assign((size_t)rows * cols, T()); }
  int rows() const { return r; }
  int cols() const { return c; }
  T* data() { return v.data(); }
  const T* data() const { return v.data(); }
  T& operator()(int i, int j) { return v[(size_t)j * r + i]; }
  const T& operator()(int i, int j) const { return v[(size_t)j * r + i]; }
};
struct Vec3 { double d[3]; double& operator[](int i) { return d[i]; } const double& operator[](int i) const { return d[i]; } };
typedef ColMajor<float> ArrayXXf;
typedef ColMajor<std::complex<float>> ArrayXXcf;

// ---- stand-ins for the reference's host data model (include/frame.h, include/map.h, include/camera.h): same member functions, same
// semantics, no Eigen / OpenCV.  The shim only ever touches them through these getters.
namespace Eigen { typedef ::ArrayXXf ArrayXXf; typedef ::ArrayXXcf ArrayXXcf; typedef ::Vec3 Vector3d; }
class Frame {                                                     // include/frame.h:10-42
 public:
  Frame(int frame_id, double timestamp, Eigen::ArrayXXf& frame, Eigen::ArrayXXcf& fft_result, Eigen::ArrayXXcf& fft_polar)
      : _frame_id(frame_id), _timestamp(timestamp), _frame(frame), _fft_result(fft_result), _fft_polar(fft_polar) {}
  void SetFrameId(int frame_id) { _frame_id = frame_id; }
  int GetFrameId() { return _frame_id; }
  void GetFFTResult(Eigen::ArrayXXcf& fft_result, Eigen::ArrayXXcf& fft_polar) { fft_result = _fft_result; fft_polar = _fft_polar; }
  void SetPose(Eigen::Vector3d& pose) { _pose = pose; }
  void GetPose(Eigen::Vector3d& pose) { pose = _pose; }
 private:
  int _frame_id; double _timestamp;
  Eigen::ArrayXXf _frame; Eigen::ArrayXXcf _fft_result, _fft_polar; Eigen::Vector3d _pose{};
};
typedef std::shared_ptr<Frame> FramePtr;
struct GridLocation { int x = 0, y = 0; };                        // include/map.h:15-32
class Map {                                                       // include/map.h:48-76, src/map.cc
 public:
  explicit Map(double grid_scale) : _grid_scale(grid_scale) {}
  void AddFrame(FramePtr& frame) {
    if (_frames.size() < 1) frame->SetFrameId(0);
    _frames[frame->GetFrameId()] = frame;
    Eigen::Vector3d pose; frame->GetPose(pose);
    GridLocation g = ComputeGridLocation(pose);
    _grid_map[{g.x, g.y}].insert(frame);
  }
  void SetFrameDistance(FramePtr& frame, double distance) { _frame_distanses[frame] = distance; }
  int GetAllFrames(std::vector<FramePtr>& frames) { for (auto kv : _frames) frames.emplace_back(kv.second); return (int)frames.size(); }
  double GetFrameDistance(FramePtr& frame) { return _frame_distanses.count(frame) > 0 ? _frame_distanses[frame] : -1; }
  GridLocation ComputeGridLocation(Eigen::Vector3d pose) {
    GridLocation g; g.x = static_cast<int>(pose[0] / _grid_scale); g.y = static_cast<int>(pose[1] / _grid_scale); return g;
  }
  int GetFramesInGrids(std::vector<FramePtr>& frames, std::vector<GridLocation>& grid_locations) {
    for (auto g : grid_locations) {
      auto it = _grid_map.find({g.x, g.y});
      if (it != _grid_map.end()) frames.insert(frames.end(), it->second.begin(), it->second.end());
    }
    return (int)frames.size();
  }
 private:
  std::map<int, FramePtr> _frames;
  std::map<FramePtr, double> _frame_distanses;
  double _grid_scale;
  std::map<std::pair<int, int>, std::set<FramePtr>> _grid_map;
};
typedef std::shared_ptr<Map> MapPtr;
class Camera {                                                    // the one conversion the two call sites use (src/camera.cc:148-158)
 public:
  Eigen::Vector3d ConvertCenterToPrincipal(const Eigen::Vector3d& image_center_pose) { return image_center_pose; }   // principal point at the centre
};
typedef std::shared_ptr<Camera> CameraPtr;
struct KeyframeSelectionConfig { double max_distance, max_angle, lower_response_thr, upper_response_thr; };
struct Vec3Call : Vec3 { double& operator()(int i) { return d[i]; } };

typedef nislam::CorrelationFlowT<ArrayXXf, ArrayXXcf, Vec3> CorrelationFlow;
typedef nislam::LoopClosureT<ArrayXXf, ArrayXXcf, Vec3, FramePtr, MapPtr, GridLocation> LoopClosure;
typedef nislam::LoopClosureResultT<Vec3, FramePtr> LoopClosureResult;
typedef std::shared_ptr<CorrelationFlow> CorrelationFlowPtr;
typedef std::shared_ptr<LoopClosure> LoopClosurePtr;

// ---- the reference's two call sites of the hot path.  The bodies of ComputeFFTResult's second line, Tracking's ComputePose call and
// FindLoopClosure are the text of /root/reference/src/map_builder.cc:74, :129-130 and :172-182, character for character: that this
// translation unit compiles and passes is the drop-in claim of INTEGRATION.md.
class MapBuilder {
 public:
  CameraPtr _camera;
  CorrelationFlowPtr _correlation_flow;
  MapPtr _map;
  LoopClosurePtr _loop_closure;
  Eigen::ArrayXXf _image_array;
  Eigen::ArrayXXcf _fft_result, _fft_polar, _last_fft_result, _last_fft_polar;
  FramePtr _current_frame;
  Eigen::Vector3d _current_pose{};
  std::vector<LoopClosureResult> _loop_matches;
  void ComputeFFTResult();
  Eigen::Vector3d TrackingCall(Eigen::Vector3d& relative_pose);
  bool FindLoopClosure();
};
void MapBuilder::ComputeFFTResult(){
  _correlation_flow->ComputeIntermedium(_image_array, _fft_result, _fft_polar); 
}
Eigen::Vector3d MapBuilder::TrackingCall(Eigen::Vector3d& relative_pose){
  Eigen::Vector3d response;
  response = _correlation_flow->ComputePose(
      _last_fft_result, _image_array, _last_fft_polar, _fft_polar, relative_pose, true);
  return response;
}
bool MapBuilder::FindLoopClosure(){
  LoopClosureResult loop_closure_result = _loop_closure->FindLoopClosure(_image_array, _current_frame, _current_pose); 
  if(loop_closure_result.found){
    loop_closure_result.relative_pose = _camera->ConvertCenterToPrincipal(loop_closure_result.relative_pose);
    _loop_matches.emplace_back(loop_closure_result);
    std::cout << "Find a loop edge, current frame = " << loop_closure_result.current_frame->GetFrameId()
              << ", loop frame = " << loop_closure_result.loop_frame->GetFrameId() << std::endl;
  }

  return loop_closure_result.found;
}

#define EXPECT(cond)                                                          \
  do {                                                                        \
    if (!(cond)) { printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); return 1; } \
  } while (0)

int main() {
  const int H = 480, W = 640;
  // smooth random texture quantised like a u8 image / 255
  ArrayXXf a; a.resize(H, W);
  std::vector<float> noise((size_t)H * W);
  srand(7);
  for (auto& x : noise) x = (float)rand() / RAND_MAX;
  for (int i = 0; i < H; ++i)
    for (int j = 0; j < W; ++j) {
      float s = 0;
      for (int di = -2; di <= 2; ++di)
        for (int dj = -2; dj <= 2; ++dj) s += noise[(size_t)((i + di + H) % H) * W + (j + dj + W) % W];
      a(i, j) = (float)((double)(float)(int)(s / 25.f * 255.f + 0.5f) / 255.0);
    }
  nislam::CFConfig cfg{0, 0, 0.1f, 0, 0.2f, 0.1f, 3, 720, 480};
  double h = H, w = W;
  auto cf = std::make_shared<CorrelationFlow>(cfg, h, w);
  EXPECT(cf->config().height == H && cf->config().width == W);           // ctor overrides (correlation_flow.cc:40-41)
  ArrayXXcf Fa, Pa;
  cf->ComputeIntermedium(a, Fa, Pa);
  EXPECT(Fa.rows() == H / 2 + 1 && Fa.cols() == W && Pa.rows() == 361 && Pa.cols() == 480);
  // DC bin = sum of the image
  double sum = 0;
  for (float x : a.v) sum += x;
  EXPECT(std::fabs(Fa(0, 0).real() - sum) < 1e-3 * sum && std::fabs(Fa(0, 0).imag()) < 1e-3);
  const int rolls[4][2] = {{0, 0}, {3, 0}, {0, -9}, {17, 25}};             // (sy, sx)
  Vec3 base{};
  for (int t = 0; t < 4; ++t) {
    const int sy = rolls[t][0], sx = rolls[t][1];
    ArrayXXf b; b.resize(H, W);
    for (int i = 0; i < H; ++i)
      for (int j = 0; j < W; ++j) b((i + sy + H) % H, (j + sx + W) % W) = a(i, j);
    ArrayXXcf Fb, Pb;
    cf->ComputeIntermedium(b, Fb, Pb);
    for (int mode = 0; mode < 2; ++mode) {
      Vec3 pose{};
      Vec3 info = cf->ComputePose(Fa, b, Pa, Pb, pose, mode == 1);
      EXPECT(pose[0] == -sx && pose[1] == -sy);
      double th = std::fmod(std::fabs(pose[2]), 2 * M_PI);
      EXPECT(th < 1e-6 || std::fabs(th - 2 * M_PI) < 1e-6);
      EXPECT(info[0] > 60 && info[2] > 60 && info[0] == info[1]);
      if (t == 0 && mode == 1) base = info;
      if (mode == 1) EXPECT(std::fabs(info[0] - base[0]) < 1e-3 * base[0]);
    }
  }
  // invalid kernel id -> std::invalid_argument from ComputePose, not from the ctor (correlation_flow.cc:157-169)
  {
    nislam::CFConfig bad = cfg; bad.kernel = 5;
    CorrelationFlow cfb(bad, h, w);
    Vec3 pose{};
    bool threw = false;
    try { cfb.ComputePose(Fa, a, Pa, Pa, pose, true); } catch (const std::invalid_argument& e) { threw = std::string(e.what()) == "Received invalid kernel type"; }
    EXPECT(threw);
  }
  // ---- LoopClosure with the reference's own signatures, driven through the verbatim MapBuilder call sites above
  {
    nislam::LoopClosureConfig lcfg{true, 60, 60, 0, 0};
    MapBuilder mb;
    mb._camera = std::make_shared<Camera>();
    mb._correlation_flow = cf;
    mb._map = std::make_shared<Map>(2.0);
    mb._loop_closure = std::shared_ptr<LoopClosure>(new LoopClosure(lcfg, mb._correlation_flow, mb._map));     // map_builder.cc:25-26
    ArrayXXf z; z.resize(H, W);
    for (int i = 0; i < H; ++i) for (int j = 0; j < W; ++j) z(i, j) = a((i * 7) % H, (j * 3) % W);   // unrelated texture
    ArrayXXcf Fz, Pz;
    cf->ComputeIntermedium(z, Fz, Pz);
    // keyframes: unrelated, a, a again (three map cells apart in x so that the prior-pose overload sees only some of them)
    struct K { ArrayXXf* img; ArrayXXcf* F; ArrayXXcf* P; double x; };
    K ks[3] = {{&z, &Fz, &Pz, 0.5}, {&a, &Fa, &Pa, 2.5}, {&a, &Fa, &Pa, 9.0}};
    std::vector<FramePtr> frames;
    for (int k = 0; k < 3; ++k) {
      FramePtr f = std::make_shared<Frame>(40 + k, 0.0, *ks[k].img, *ks[k].F, *ks[k].P);
      Vec3 pose{{ks[k].x, 0.5, 0.0}};
      f->SetPose(pose);
      mb._map->AddFrame(f);                       // the first frame is renamed to id 0 (map.cc:19-22)
      mb._map->SetFrameDistance(f, (double)k);
      frames.push_back(f);
    }
    mb._image_array.resize(H, W);
    for (int i = 0; i < H; ++i) for (int j = 0; j < W; ++j) mb._image_array((i + 5) % H, (j + 11) % W) = a(i, j);
    mb.ComputeFFTResult();                                                                          // map_builder.cc:72-75
    mb._current_frame = std::make_shared<Frame>(99, 1.0, mb._image_array, mb._fft_result, mb._fft_polar);
    mb._map->SetFrameDistance(mb._current_frame, 50.0);
    // Tracking's ComputePose call against keyframe a (map_builder.cc:129-130), twice: the second call reuses the device copy of `last`
    mb._last_fft_result = Fa; mb._last_fft_polar = Pa;
    Vec3 rel{}, rel2{};
    Vec3 resp = mb.TrackingCall(rel);
    Vec3 resp2 = mb.TrackingCall(rel2);
    EXPECT(rel[0] == -11 && rel[1] == -5 && resp[0] > 60 && resp[2] > 60);
    EXPECT(rel2[0] == rel[0] && rel2[1] == rel[1] && resp2[0] == resp[0] && resp2[2] == resp[2]);
    mb._last_fft_result = Fz; mb._last_fft_polar = Pz;                  // same buffers, new keyframe content -> re-uploaded
    Vec3 rel3{};
    Vec3 resp3 = mb.TrackingCall(rel3);
    EXPECT(resp3[0] < 30);
    // prior pose in cell (1, 0): neighbourhood cells x = 0..2 -> frames 0 (z) and 1 (a); frame 2 (x = 9 -> cell 4) is out of reach
    mb._current_pose = Vec3{{2.2, 0.4, 0.0}};
    EXPECT(mb.FindLoopClosure());                                                                   // map_builder.cc:172-182 verbatim
    EXPECT(mb._loop_matches.size() == 1 && mb._loop_matches[0].loop_frame == frames[1]);
    EXPECT(mb._loop_matches[0].loop_frame->GetFrameId() == 41 && mb._loop_matches[0].current_frame->GetFrameId() == 99);
    EXPECT(mb._loop_matches[0].relative_pose[0] == -11 && mb._loop_matches[0].relative_pose[1] == -5);
    EXPECT(mb._loop_closure->StoredFrames() == 2);                      // only the two candidates were uploaded
    // all-frames overload (loop_closure.cc:10-15): id order 0, 41, 42 -> the FIRST copy of a wins (strict '>', :61)
    LoopClosureResult r_all = mb._loop_closure->FindLoopClosure(mb._image_array, mb._current_frame);
    EXPECT(r_all.found && r_all.loop_frame == frames[1] && mb._loop_closure->StoredFrames() == 3);
    // explicit list in another order (:36-73): iteration order decides the tie
    std::vector<FramePtr> lst{frames[2], frames[1], frames[0]};
    LoopClosureResult r_lst = mb._loop_closure->FindLoopClosure(mb._image_array, mb._current_frame, lst);
    EXPECT(r_lst.found && r_lst.loop_frame == frames[2]);
    EXPECT(r_lst.response[0] == r_all.response[0] && r_lst.response[2] == r_all.response[2]);
    // filters (:43-53): a frame gap of 1000 removes every candidate -> initial best (-1,-1,-1), no loop frame
    nislam::LoopClosureConfig far{true, 60, 60, 1000, 0};
    LoopClosure lc2(far, cf, mb._map);
    LoopClosureResult r_far = lc2.FindLoopClosure(mb._image_array, mb._current_frame);
    EXPECT(!r_far.found && !r_far.loop_frame && r_far.response[0] == -1.0);
    // accumulated-distance filter: |50 - d| < 49.5 drops d = 1, 2 -> only frame 0 (unrelated) is evaluated
    nislam::LoopClosureConfig near{true, 60, 60, 0, 49.5};
    LoopClosure lc3(near, cf, mb._map);
    LoopClosureResult r_near = lc3.FindLoopClosure(mb._image_array, mb._current_frame);
    EXPECT(!r_near.found && r_near.loop_frame == frames[0]);
    // online latency through the reference's per-call surface (main.cpp:51-86 feeds one frame at a time): ComputeIntermedium +
    // ComputePose(tracking) per frame with host arrays in and out, the keyframe's operands cached on the device
    mb._last_fft_result = Fa; mb._last_fft_polar = Pa;
    std::vector<double> ms;
    for (int it = 0; it < 60; ++it) {
      auto t0 = std::chrono::steady_clock::now();
      mb.ComputeFFTResult();
      Vec3 r{};
      mb.TrackingCall(r);
      auto t1 = std::chrono::steady_clock::now();
      if (it >= 10) ms.push_back(std::chrono::duration<double, std::milli>(t1 - t0).count());
    }
    std::sort(ms.begin(), ms.end());
    printf("shim_online_ms p50 %.3f p99 %.3f\n", ms[ms.size() / 2], ms[(ms.size() * 99) / 100]);
  }
  // Multi-GPU scan entry points from a C++ host, on a one-rank communicator: NCCL id -> nis_comm_init -> ONE call per query
  // (image in; broadcast, features, local scan, all-gather and reduction inside the library) == nis_features_u8 + nis_loop_scan
  {
    nis_cf_config c{0.1f, 0, 0.2f, 0.1f, 3, 720, 480};
    nis_ctx* ctx = nullptr;
    EXPECT(nis_create(&c, H, W, 0, &ctx) == 0);
    std::vector<uint8_t> imgs((size_t)3 * H * W), query((size_t)H * W);
    // keyframes: the texture mirrored top-bottom, the texture itself, the texture mirrored left-right (mirror images do not correlate
    // under translation); the query is the middle one moved by (sy, sx) = (3, -9)
    for (int i = 0; i < H; ++i)
      for (int j = 0; j < W; ++j) {
        imgs[((size_t)0 * H + i) * W + j] = (uint8_t)std::lrint(a(H - 1 - i, j) * 255.f);
        imgs[((size_t)1 * H + i) * W + j] = (uint8_t)std::lrint(a(i, j) * 255.f);
        imgs[((size_t)2 * H + i) * W + j] = (uint8_t)std::lrint(a(i, W - 1 - j) * 255.f);
      }
    for (int i = 0; i < H; ++i)
      for (int j = 0; j < W; ++j) query[(size_t)i * W + j] = (uint8_t)std::lrint(a((i - 3 + H) % H, (j + 9 + W) % W) * 255.f);
    const int ids[3] = {10, 11, 12};
    EXPECT(nis_db_add_images(ctx, imgs.data(), 3, ids, nullptr) == 0 && nis_db_size(ctx) == 3);
    char id[128];
    EXPECT(nis_nccl_unique_id(id) == 0);
    EXPECT(nis_comm_init(ctx, id, 0, 1) == 0);
    nis_loop_config lcfg{60.0, 60.0, 0, 0.0};
    nis_loop_result sharded, local, direct;
    int winner_rank = -7;
    EXPECT(nis_loop_scan_sharded(ctx, query.data(), 0, 99, 0.0, &lcfg, 5000, &sharded, &winner_rank, &local) == 0);
    nis_frame* qf = nullptr;
    EXPECT(nis_features_u8(ctx, query.data(), &qf) == 0);
    EXPECT(nis_loop_scan(ctx, qf, 99, 0.0, &lcfg, nullptr, 0, &direct, nullptr) == 0);
    EXPECT(winner_rank == 0 && direct.found && direct.slot == 1 && direct.frame_id == 11 && direct.evaluated == 3);
    EXPECT(local.slot == direct.slot && sharded.slot == 5000 + direct.slot && sharded.frame_id == direct.frame_id && sharded.found == direct.found);
    EXPECT(direct.relative_pose[0] == 9.0 && direct.relative_pose[1] == -3.0);          // content rolled by (sy, sx) = (3, -9) -> pose (-sx, -sy, 0)
    for (int i = 0; i < 3; ++i) EXPECT(sharded.relative_pose[i] == direct.relative_pose[i] && sharded.response[i] == direct.response[i]);
    nis_frame_free(ctx, qf);
    EXPECT(nis_comm_destroy(ctx) == 0);
    nis_destroy(ctx);
  }
  // MapStitcher shim (host/map_stitcher.hpp) driven like MapBuilder drives the reference (map_builder.cc:37, :62, :113): identity pose,
  // principal point at the centre -> pixel (i, j) lands on ground (i - W/2, j - H/2); first insert stores the scaled pixel itself
  {
    struct FakeFrame { Vec3 p; void GetPose(Vec3& out) const { out = p; } };
    typedef nislam::MapStitcherT<Vec3, ColMajor<int>> MapStitcher;
    nis_camera_model cam{500.0, 500.0, W / 2.0, H / 2.0, 1.0, {1, 0, 0, 0, 1, 0, 0, 0, 1}};
    MapStitcher ms(1000, true, cam, H, W, -1, -1, 2, 2);
    std::vector<uint8_t> img((size_t)H * W);
    for (int j = 0; j < H; ++j) for (int i = 0; i < W; ++i) img[(size_t)j * W + i] = (uint8_t)((i + 2 * j) & 255);
    FakeFrame f0{{{0.0, 0.0, 0.0}}};
    const FakeFrame* fp = &f0;
    ms.InsertFrame(fp, img.data());
    ColMajor<int> d, wgt;
    EXPECT(ms.GetCell(0, 0, d, wgt) && ms.DroppedPixels() == 0);
    // ground (0, 0) <- pixel (i, j) = (W/2, H/2); scaled value = rint(v * 100/255)
    const int v = img[(size_t)(H / 2) * W + W / 2];
    EXPECT(wgt(0, 0) == 1 && d(0, 0) == (int)std::lrint((float)v * (float)(100.0 / 255.0)));
    EXPECT(ms.GetCell(-1, -1, d, wgt) && wgt(999, 999) == 1 && wgt(0, 0) == 0);      // ground (-1, -1) <- pixel (W/2 - 1, H/2 - 1)
    // second identical insert: existing cells average, (d*1 + d*1) / 2 = d, weight 2
    ms.InsertFrame(fp, img.data());
    EXPECT(ms.GetCell(0, 0, d, wgt) && wgt(0, 0) == 2 && d(0, 0) == (int)std::lrint((float)v * (float)(100.0 / 255.0)));
    std::vector<const FakeFrame*> frames{fp, fp};
    ms.RecomputeOccupancy(frames);
    EXPECT(ms.GetCell(0, 0, d, wgt) && wgt(0, 0) == 2);
    MapStitcher off(1000, false, cam, H, W, -1, -1, 2, 2);                           // stitch_map: false -> every call is a no-op (:15)
    off.InsertFrame(fp, img.data());
    EXPECT(!off.GetCell(0, 0, d, wgt));
  }
  printf("shim_test ok\n");
  return 0;
}
