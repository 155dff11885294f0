// Drives the reference-facing C++ shim (ni_slam_b200/host/correlation_flow.hpp) the way MapBuilder drives the reference
// classes (src/map_builder.cc:72-75, :127-138, :172-182), with a stand-in for the Eigen array types (Eigen is not in the
// build image).  Known answers: SURVEY.md Appendix C.4/C.5/C.7 (circular rolls, identity, first-wins scan).
// Build: g++ -std=c++17 tests/cpp/shim_test.cc -L ni_slam_b200/lib -lnislam -Wl,-rpath,... ; needs a GPU to run.
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
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
typedef nislam::CorrelationFlowT<ArrayXXf, ArrayXXcf, Vec3> CorrelationFlow;
typedef nislam::LoopClosureT<ArrayXXf, ArrayXXcf, Vec3> LoopClosure;

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
  // scan: three copies of keyframe a -> the first inserted wins (strict '>', loop_closure.cc:61)
  nislam::LoopClosureConfig lcfg{true, 60, 60, 0, 0};
  LoopClosure lc(lcfg, cf);
  ArrayXXf z; z.resize(H, W);
  for (int i = 0; i < H; ++i) for (int j = 0; j < W; ++j) z(i, j) = a((i * 7) % H, (j * 3) % W);   // unrelated texture
  ArrayXXcf Fz, Pz;
  cf->ComputeIntermedium(z, Fz, Pz);
  EXPECT(lc.AddFrame(40, z, Fz, Pz, 0.0) == 0);
  EXPECT(lc.AddFrame(41, a, Fa, Pa, 1.0) == 1);
  EXPECT(lc.AddFrame(42, a, Fa, Pa, 2.0) == 2);
  ArrayXXf b; b.resize(H, W);
  for (int i = 0; i < H; ++i) for (int j = 0; j < W; ++j) b((i + 5) % H, (j + 11) % W) = a(i, j);
  ArrayXXcf Fb, Pb;
  cf->ComputeIntermedium(b, Fb, Pb);
  auto res = lc.FindLoopClosure(b, 99, Fb, Pb, 50.0);
  EXPECT(res.found && res.loop_slot == 1 && res.loop_frame_id == 41);
  EXPECT(res.relative_pose[0] == -11 && res.relative_pose[1] == -5);
  auto res2 = lc.FindLoopClosure(b, 99, Fb, Pb, 50.0, std::vector<int32_t>{2, 1, 0});
  EXPECT(res2.loop_slot == 2);
  nislam::LoopClosureConfig far{true, 60, 60, 1000, 0};
  LoopClosure lc2(far, cf);
  auto res3 = lc2.FindLoopClosure(b, 99, Fb, Pb, 50.0);
  EXPECT(!res3.found && res3.loop_slot == -1 && res3.response[0] == -1.0);
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
