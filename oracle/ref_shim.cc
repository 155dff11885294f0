// ref_shim.cc -- C entry points over the reference's OWN CorrelationFlow / LoopClosure / Map / Frame classes, compiled from
// /root/reference/src/{correlation_flow,loop_closure,utils,map,frame}.cc UNMODIFIED against the stand-in headers of
// oracle/ref_stubs (Eigen, FFTW3, OpenCV, yaml-cpp, Ceres are not in this image).  Output: oracle/_ref/libnislam_ref.so.
// TEST INFRASTRUCTURE ONLY: used by tests/ (oracle pinning) and bench.py's CPU legs; never by ni_slam_b200/.
// Arrays cross this interface in the reference's layout (Eigen column-major, half spectra (R/2+1) x C complex).
#include <cstdio>
#include <cstring>
#include <iostream>
#include <stdexcept>

#include "fftw3.h"
#include "loop_closure.h"      // the reference's (pulls correlation_flow.h, map.h, frame.h, read_configs.h, utils.h)

extern "C" {
void orc_fft2(const float* x, int R, int C, float* xf);
void orc_ifft2_raw(const float* xf, int R, int C, float* x);

// ---- fftw3.h stand-in ------------------------------------------------------------------------------------------
fftwf_plan fftwf_plan_dft_r2c_2d(int n0, int n1, float* in, fftwf_complex* out, unsigned) {
  return new fftwf_plan_s{0, n0, n1, in, out};
}
fftwf_plan fftwf_plan_dft_c2r_2d(int n0, int n1, fftwf_complex* in, float* out, unsigned) {
  return new fftwf_plan_s{1, n0, n1, in, out};
}
void fftwf_execute(const fftwf_plan p) {
  // row-major n0 x n1 == column-major R = n1 rows x C = n0 columns ("reverse order for column major", correlation_flow.cc:57)
  if (p->kind == 0) orc_fft2((const float*)p->in, p->n1, p->n0, (float*)p->out);
  else orc_ifft2_raw((const float*)p->in, p->n1, p->n0, (float*)p->out);
}
void fftwf_destroy_plan(fftwf_plan p) { delete p; }
void fftw_cleanup(void) {}
}

namespace {
struct RefCtx {
  CFConfig cfg;
  int H, W;
  CorrelationFlowPtr cf;
};
// ComputePose prints two lines per call (correlation_flow.cc:139-140): std::cout of this shared object's process is silenced once
// (thread-safe: callers time the library from several threads); Python's own stdout is a different stream
struct Quiet {
  Quiet() { static const bool once = (std::cout.rdbuf(nullptr), true); (void)once; }
};
Eigen::ArrayXXf real_in(const float* p, int R, int C) {
  Eigen::ArrayXXf a(R, C);
  memcpy(a.data(), p, sizeof(float) * (size_t)R * C);
  return a;
}
Eigen::ArrayXXcf spec_in(const float* p, int R, int C) {
  Eigen::ArrayXXcf a(R / 2 + 1, C);
  memcpy((void*)a.data(), p, sizeof(float) * 2 * (size_t)(R / 2 + 1) * C);
  return a;
}
}  // namespace

extern "C" {

struct ref_cf_config { int height, width; float lambda; int kernel; float sigma, offset; int power, rotation_divisor, rotation_channel; };
struct ref_loop_config { double position_response_thr, angle_response_thr; int frame_gap_thr; double distance_thr; };
struct ref_loop_result { int found, index, frame_id; double relative_pose[3], response[3]; };

void* ref_create(const ref_cf_config* c, double image_height, double image_width) {
  RefCtx* r = new RefCtx();
  r->cfg.width = c->width; r->cfg.height = c->height; r->cfg.lambda = c->lambda; r->cfg.kernel = c->kernel; r->cfg.sigma = c->sigma;
  r->cfg.offset = c->offset; r->cfg.power = c->power; r->cfg.rotation_divisor = c->rotation_divisor; r->cfg.rotation_channel = c->rotation_channel;
  r->H = (int)image_height; r->W = (int)image_width;
  r->cf = std::make_shared<CorrelationFlow>(r->cfg, image_height, image_width);        // correlation_flow.cc:37-44
  return r;
}
void ref_destroy(void* h) { delete (RefCtx*)h; }

// utils.cc:110-118 through a u8 cv::Mat, like MapBuilder::ComputeFFTResult (map_builder.cc:72-75)
void ref_normalize_u8(const uint8_t* img_rowmajor, int H, int W, float* out_colmajor) {
  cv::Mat m(H, W, CV_8U, (void*)img_rowmajor);
  Eigen::ArrayXXf a;
  ConvertMatToNormalizedArray(m, a);
  memcpy(out_colmajor, a.data(), sizeof(float) * (size_t)H * W);
}
double ref_normalize_degree(double a) { return NormalizeDegree(a); }                   // utils.cc:173-175
void ref_rotate(const float* img, int H, int W, float degree, float* out) {             // utils.cc:154-161
  Eigen::ArrayXXf r = RotateArray(real_in(img, H, W), degree);
  memcpy(out, r.data(), sizeof(float) * (size_t)H * W);
}

void ref_compute_intermedium(void* h, const float* image, float* fft_result, float* fft_polar) {
  RefCtx* r = (RefCtx*)h;
  Eigen::ArrayXXcf F, P;
  r->cf->ComputeIntermedium(real_in(image, r->H, r->W), F, P);                          // correlation_flow.cc:89-95
  memcpy(fft_result, (const void*)F.data(), sizeof(float) * 2 * (size_t)F.size());
  memcpy(fft_polar, (const void*)P.data(), sizeof(float) * 2 * (size_t)P.size());
}

// returns 0, or -1 when the reference throws std::invalid_argument (bad kernel id, correlation_flow.cc:168)
int ref_compute_pose(void* h, const float* last_fft_result, const float* image, const float* last_fft_polar, const float* fft_polar,
                     int not_large_rotation, double pose[3], double info[3]) {
  RefCtx* r = (RefCtx*)h;
  const int D = r->cfg.rotation_divisor, Cp = r->cfg.rotation_channel;
  Quiet q;
  try {
    Eigen::Vector3d p;
    Eigen::Vector3d i = r->cf->ComputePose(spec_in(last_fft_result, r->H, r->W), real_in(image, r->H, r->W), spec_in(last_fft_polar, D, Cp),
                                           spec_in(fft_polar, D, Cp), p, not_large_rotation != 0);   // correlation_flow.cc:97-143
    for (int k = 0; k < 3; ++k) { pose[k] = p[k]; info[k] = i[k]; }
  } catch (const std::invalid_argument&) {
    return -1;
  }
  return 0;
}

// LoopClosure::FindLoopClosure through the reference's own Map / Frame objects (loop_closure.cc:10-73, map.cc:18-101).
// mode 0: explicit frame list in the given order (:36-73); 1: all frames of the map, id order (:10-15); 2: 3x3 grid cells around
// prior_pose (:17-34; unordered_set order).  poses: 3 doubles per keyframe (grid filing, Map::AddFrame); dists: accumulated
// distance per keyframe or NULL (Map::GetFrameDistance then returns -1).  result.index = position in the input arrays.
int ref_find_loop_closure(void* h, const ref_loop_config* thr, double grid_scale, const float* image, const float* cur_fft_result,
                          const float* cur_fft_polar, int cur_id, double cur_dist, int has_cur_dist, int n, const float* const* fft_results,
                          const float* const* fft_polars, const int* frame_ids, const double* dists, const double* poses, int mode,
                          const double* prior_pose, ref_loop_result* out) {
  RefCtx* r = (RefCtx*)h;
  const int D = r->cfg.rotation_divisor, Cp = r->cfg.rotation_channel;
  LoopClosureConfig lc;
  lc.to_find_loop = true; lc.position_response_thr = thr->position_response_thr; lc.angle_response_thr = thr->angle_response_thr;
  lc.frame_gap_thr = thr->frame_gap_thr; lc.distance_thr = thr->distance_thr;
  MapConfig mc; mc.grid_scale = grid_scale;
  MapPtr map = std::make_shared<Map>(mc);
  LoopClosure loop(lc, r->cf, map);
  Eigen::ArrayXXf img = real_in(image, r->H, r->W);
  std::vector<FramePtr> frames;
  for (int i = 0; i < n; ++i) {
    Eigen::ArrayXXf none;
    Eigen::ArrayXXcf F = spec_in(fft_results[i], r->H, r->W), P = spec_in(fft_polars[i], D, Cp);
    FramePtr f = std::make_shared<Frame>(frame_ids[i], 0.0, none, F, P);
    Eigen::Vector3d pose(poses ? poses[3 * i] : 0.0, poses ? poses[3 * i + 1] : 0.0, poses ? poses[3 * i + 2] : 0.0);
    f->SetPose(pose);
    if (mode != 0) {
      map->AddFrame(f);                       // note: the first frame added gets id 0 (map.cc:19-22), like in the reference
    }
    if (dists) map->SetFrameDistance(f, dists[i]);
    frames.push_back(f);
  }
  Eigen::ArrayXXf none;
  Eigen::ArrayXXcf cF = spec_in(cur_fft_result, r->H, r->W), cP = spec_in(cur_fft_polar, D, Cp);
  FramePtr cur = std::make_shared<Frame>(cur_id, 0.0, none, cF, cP);
  if (has_cur_dist) map->SetFrameDistance(cur, cur_dist);
  Quiet q;
  LoopClosureResult res;
  try {
    if (mode == 0) res = loop.FindLoopClosure(img, cur, frames);
    else if (mode == 1) res = loop.FindLoopClosure(img, cur);
    else {
      Eigen::Vector3d prior(prior_pose[0], prior_pose[1], prior_pose[2]);
      res = loop.FindLoopClosure(img, cur, prior);
    }
  } catch (const std::invalid_argument&) {
    return -1;
  }
  out->found = res.found ? 1 : 0;
  out->index = -1; out->frame_id = -1;
  if (res.loop_frame) {
    out->frame_id = res.loop_frame->GetFrameId();
    for (int i = 0; i < n; ++i) if (frames[i] == res.loop_frame) out->index = i;
  }
  for (int k = 0; k < 3; ++k) { out->response[k] = res.response[k]; out->relative_pose[k] = res.loop_frame ? res.relative_pose[k] : 0.0; }
  return 0;
}

}  // extern "C"
