// nis_stitch.cu -- MapStitcher on the GPU (src/map_stitcher.cc:14-145): the occupancy mosaic of keyframe images, sm_100a.
// Integer / byte work: per frame one u8 image in, per-frame (sum, count) pairs scattered with one 64-bit atomic per pixel into the
// frame's bounding box (which the merge kernel leaves zeroed again), merged into the dense cell window with the reference's integer rules.
//   stitch_normalize_kernel : InsertFrame's  image * (100.0 / 255.0)  as u8 (cv::Mat scaling: float multiply, round half to even)
//   stitch_scatter_kernel   : the pixel loop of AddImageToOccupancy (:95-111): x = (int)(Wx(i) + Hx(j)), y = (int)(Wy(i) + Hy(j))
//   stitch_merge_kernel     : the per-cell merge (:113-132), element-wise; untouched elements are provably unchanged by it
//   stitch_commit_kernel    : cells that received pixels now exist (the `_occupancy_data.count(loc)` test of the NEXT frame)
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <vector>

#include "../../include/nislam.h"
#include "../host/pose_math.hpp"

namespace {

struct PlaceArgs {          // AddImageToOccupancy's per-frame constants (:44-66)
  double r00, r01, r10, r11, X, Y, cx, cy;
  int min_x, min_y, bw, bh;
};

__global__ void stitch_normalize_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float scale = (float)(100.0 / 255.0);                         // convertTo's alpha is applied in float for 8u -> 8u
  const int v = __float2int_rn(__fmul_rn((float)in[i], scale));       // saturate_cast<uchar>(float) = cvRound
  out[i] = (uint8_t)min(255, max(0, v));
}

__device__ __forceinline__ void ground_xy(const PlaceArgs& a, int i, int j, int& x, int& y) {
  const double wi = (double)i - a.cx, hj = (double)j - a.cy;
  const double wx = __dadd_rn(__dmul_rn(a.r00, wi), a.X), wy = __dadd_rn(__dmul_rn(a.r10, wi), a.Y);     // Wx, Wy (:61-62)
  const double hx = __dmul_rn(a.r01, hj), hy = __dmul_rn(a.r11, hj);                                     // Hx, Hy (:63-64)
  x = __double2int_rz(__dadd_rn(wx, hx));                                                                // static_cast<int>: toward zero
  y = __double2int_rz(__dadd_rn(wy, hy));
}

// also commits the PREVIOUS frame's touched cells (this kernel does not read `present`; the merge kernel that does runs after it)
__global__ void stitch_scatter_kernel(const uint8_t* __restrict__ img, int H, int W, PlaceArgs a, unsigned long long* __restrict__ tbox,
                                      int* __restrict__ present, int* __restrict__ touched, int ncells) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (j == 0)
    for (int c = i; c < ncells; c += gridDim.x * blockDim.x)
      if (touched[c]) { present[c] = 1; touched[c] = 0; }
  if (i >= W) return;
  int x, y;
  ground_xy(a, i, j, x, y);
  const int b = (y - a.min_y) * a.bw + (x - a.min_x);
  atomicAdd(&tbox[b], ((unsigned long long)img[(size_t)j * W + i] << 32) | 1ull);      // per-frame (sum, count) in one 64-bit atomic
}

__device__ __forceinline__ int cell_of(int x, int cs, int& in) {      // ComputeCellPosition (:24-34): floor division
  const int c = x >= 0 ? x / cs : (x - cs + 1) / cs;
  in = x - c * cs;
  return c;
}

__global__ void stitch_merge_kernel(PlaceArgs a, unsigned long long* __restrict__ tbox, int cs, int cell_x0, int cell_y0,
                                    int cells_x, int cells_y, int* __restrict__ data, int* __restrict__ weight, const int* __restrict__ present,
                                    int* __restrict__ touched, unsigned long long* __restrict__ dropped) {
  const int bx = blockIdx.x * blockDim.x + threadIdx.x, by = blockIdx.y;
  if (bx >= a.bw) return;
  const unsigned long long t = tbox[by * a.bw + bx];
  const int cnt = (int)(unsigned)(t & 0xffffffffull), sum = (int)(unsigned)(t >> 32);
  if (cnt == 0) return;                           // untouched elements: (d*w + 0*0)/w = d for w >= 1, and d = 0 where w = 0
  tbox[by * a.bw + bx] = 0ull;                    // the box is all zero again for the next frame: no per-frame memset
  int inx, iny;
  const int ccx = cell_of(a.min_x + bx, cs, inx) - cell_x0, ccy = cell_of(a.min_y + by, cs, iny) - cell_y0;
  if (ccx < 0 || ccy < 0 || ccx >= cells_x || ccy >= cells_y) { atomicAdd(dropped, (unsigned long long)cnt); return; }
  const int cell = ccy * cells_x + ccx;
  const size_t e = (size_t)cell * cs * cs + (size_t)iny * cs + inx;
  if (present[cell]) {                            // :117-127
    const unsigned d = (unsigned)data[e] * (unsigned)weight[e] + (unsigned)sum * (unsigned)cnt;     // int arithmetic, wrapping like the reference's
    const int w = weight[e] + cnt;
    data[e] = w < 1 ? (int)d : (int)d / w;
    weight[e] = w;
  } else {                                        // :128-132: a new cell takes the raw per-frame sums
    data[e] = sum;
    weight[e] = cnt;
  }
  touched[cell] = 1;
}

__global__ void stitch_commit_kernel(int* __restrict__ present, int* __restrict__ touched, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (touched[i]) { present[i] = 1; touched[i] = 0; }
}

}  // namespace

struct nis_stitcher {
  int device = 0, H = 0, W = 0, cs = 0, cell_x0 = 0, cell_y0 = 0, cells_x = 0, cells_y = 0;
  int* data = nullptr; int* weight = nullptr; int* present = nullptr; int* touched = nullptr;
  unsigned long long* tbox = nullptr; size_t tcap = 0;   // per-frame (sum << 32 | count) over the frame's bounding box
  unsigned long long* dropped = nullptr;
  uint8_t* staging = nullptr;
  std::vector<uint8_t*> chunks;        // normalised images, kFramesPerChunk per allocation (_raw_images)
  int frames = 0;
  cudaStream_t stream = nullptr;
};
static const int kFramesPerChunk = 256;

#define SCU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return e_ == cudaErrorMemoryAllocation ? NIS_ERR_OUT_OF_MEMORY : NIS_ERR_CUDA; } while (0)

static const uint8_t* frame_ptr(const nis_stitcher* st, int slot) {
  return st->chunks[slot / kFramesPerChunk] + (size_t)(slot % kFramesPerChunk) * st->H * st->W;
}

static cudaError_t commit(nis_stitcher* st) {      // cells touched by the last frame exist from now on
  const int nc = st->cells_x * st->cells_y;
  stitch_commit_kernel<<<(nc + 255) / 256, 256, 0, st->stream>>>(st->present, st->touched, nc);
  return cudaGetLastError();
}

// AddImageToOccupancy(frame) for the stored image `slot` placed at `robot_pose`
static int add_image(nis_stitcher* st, int slot, const double robot_pose[3], const nis_camera_model* cam) {
  using namespace nis::pose;
  const int H = st->H, W = st->W;
  P3 ip = robot_to_image_plane(*cam, P3{robot_pose[0], robot_pose[1], robot_pose[2]});     // :39-40
  ip = principal_to_center(*cam, W, H, ip);                                                 // :41
  const double c = cos(ip.th), s = sin(ip.th);
  PlaceArgs a{c, -s, s, c, ip.x, ip.y, (double)W / 2, (double)H / 2, 0, 0, 0, 0};
  // corners (:67-79); the same rounding sequence as the device code (no FMA contraction on the host: plain x86-64 doubles)
  int xs[4], ys[4], k = 0;
  for (int i : {0, W - 1})
    for (int j : {0, H - 1}) {
      const double wi = (double)i - a.cx, hj = (double)j - a.cy;
      volatile double wx = a.r00 * wi; wx = wx + a.X;
      volatile double wy = a.r10 * wi; wy = wy + a.Y;
      volatile double hx = a.r01 * hj, hy = a.r11 * hj;
      xs[k] = (int)(wx + hx); ys[k] = (int)(wy + hy); ++k;
    }
  a.min_x = *std::min_element(xs, xs + 4); a.min_y = *std::min_element(ys, ys + 4);
  a.bw = *std::max_element(xs, xs + 4) - a.min_x + 1; a.bh = *std::max_element(ys, ys + 4) - a.min_y + 1;
  const size_t need = (size_t)a.bw * a.bh;
  if (need > st->tcap) {
    SCU(cudaStreamSynchronize(st->stream));
    if (st->tbox) { cudaFree(st->tbox); st->tbox = nullptr; st->tcap = 0; }
    SCU(cudaMalloc(&st->tbox, need * sizeof(unsigned long long)));
    SCU(cudaMemsetAsync(st->tbox, 0, need * sizeof(unsigned long long), st->stream));     // zeroed once; the merge kernel restores the zeros it consumed
    st->tcap = need;
  }
  const int nc = st->cells_x * st->cells_y;
  stitch_scatter_kernel<<<dim3((W + 255) / 256, H), 256, 0, st->stream>>>(frame_ptr(st, slot), H, W, a, st->tbox, st->present, st->touched, nc);
  stitch_merge_kernel<<<dim3((a.bw + 255) / 256, a.bh), 256, 0, st->stream>>>(a, st->tbox, st->cs, st->cell_x0, st->cell_y0, st->cells_x,
                                                                              st->cells_y, st->data, st->weight, st->present, st->touched, st->dropped);
  SCU(cudaGetLastError());
  return NIS_OK;
}

extern "C" {

int nis_stitcher_create(int device, int image_height, int image_width, int cell_size, int cell_x0, int cell_y0, int cells_x, int cells_y,
                        nis_stitcher** out) {
  if (!out || image_height < 1 || image_width < 1 || cell_size < 1 || cells_x < 1 || cells_y < 1) return NIS_ERR_INVALID_ARGUMENT;
  if (cudaSetDevice(device) != cudaSuccess) return NIS_ERR_CUDA;
  nis_stitcher* st = new nis_stitcher();
  st->device = device; st->H = image_height; st->W = image_width; st->cs = cell_size;
  st->cell_x0 = cell_x0; st->cell_y0 = cell_y0; st->cells_x = cells_x; st->cells_y = cells_y;
  const size_t n = (size_t)cells_x * cells_y * cell_size * cell_size, nc = (size_t)cells_x * cells_y;
  bool ok = cudaStreamCreateWithFlags(&st->stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaMalloc(&st->data, n * sizeof(int)) == cudaSuccess && cudaMalloc(&st->weight, n * sizeof(int)) == cudaSuccess &&
            cudaMalloc(&st->present, nc * sizeof(int)) == cudaSuccess && cudaMalloc(&st->touched, nc * sizeof(int)) == cudaSuccess &&
            cudaMalloc(&st->dropped, sizeof(unsigned long long)) == cudaSuccess &&
            cudaMalloc(&st->staging, (size_t)image_height * image_width) == cudaSuccess;
  ok = ok && cudaMemsetAsync(st->data, 0, n * sizeof(int), st->stream) == cudaSuccess &&
       cudaMemsetAsync(st->weight, 0, n * sizeof(int), st->stream) == cudaSuccess &&
       cudaMemsetAsync(st->present, 0, nc * sizeof(int), st->stream) == cudaSuccess &&
       cudaMemsetAsync(st->touched, 0, nc * sizeof(int), st->stream) == cudaSuccess &&
       cudaMemsetAsync(st->dropped, 0, sizeof(unsigned long long), st->stream) == cudaSuccess &&
       cudaStreamSynchronize(st->stream) == cudaSuccess;
  if (!ok) { nis_stitcher_destroy(st); return NIS_ERR_OUT_OF_MEMORY; }
  *out = st;
  return NIS_OK;
}

int nis_stitcher_destroy(nis_stitcher* st) {
  if (!st) return NIS_OK;
  cudaSetDevice(st->device);
  if (st->stream) cudaStreamSynchronize(st->stream);
  void* bufs[] = {st->data, st->weight, st->present, st->touched, st->tbox, st->dropped, st->staging};
  for (void* b : bufs) if (b) cudaFree(b);
  for (uint8_t* c : st->chunks) cudaFree(c);
  if (st->stream) cudaStreamDestroy(st->stream);
  delete st;
  return NIS_OK;
}

int nis_stitcher_insert(nis_stitcher* st, const uint8_t* image_u8, const double robot_pose[3], const nis_camera_model* cam, int* frame_slot) {
  if (!st || !image_u8 || !robot_pose || !cam || cam->height <= 0 || cam->fx == 0 || cam->fy == 0) return NIS_ERR_INVALID_ARGUMENT;
  // a NaN / infinite / absurdly distant pose would make the integer ground positions undefined behaviour and the scatter box enormous
  for (int i = 0; i < 3; ++i)
    if (!isfinite(robot_pose[i]) || fabs(robot_pose[i]) > 1e8) return NIS_ERR_INVALID_ARGUMENT;
  SCU(cudaSetDevice(st->device));
  const size_t npx = (size_t)st->H * st->W;
  const int slot = st->frames;
  if (slot / kFramesPerChunk >= (int)st->chunks.size()) {
    uint8_t* c = nullptr;
    SCU(cudaMalloc(&c, npx * kFramesPerChunk));
    st->chunks.push_back(c);
  }
  SCU(cudaMemcpyAsync(st->staging, image_u8, npx, cudaMemcpyHostToDevice, st->stream));
  stitch_normalize_kernel<<<(unsigned)((npx + 255) / 256), 256, 0, st->stream>>>(st->staging, const_cast<uint8_t*>(frame_ptr(st, slot)), (int)npx);
  // the frame only counts once it has been merged: a failure below (e.g. no memory for the scatter box) leaves the stitcher as it was
  const int rc = add_image(st, slot, robot_pose, cam);
  if (rc != NIS_OK) return rc;
  SCU(commit(st));
  SCU(cudaStreamSynchronize(st->stream));          // the caller's image buffer is free again, like the reference's synchronous call
  st->frames = slot + 1;
  if (frame_slot) *frame_slot = slot;
  return NIS_OK;
}

int nis_stitcher_recompute(nis_stitcher* st, const double* robot_poses, const nis_camera_model* cam) {
  if (!st || (!robot_poses && st->frames > 0) || !cam || cam->height <= 0 || cam->fx == 0 || cam->fy == 0) return NIS_ERR_INVALID_ARGUMENT;
  SCU(cudaSetDevice(st->device));
  const size_t n = (size_t)st->cells_x * st->cells_y * st->cs * st->cs, nc = (size_t)st->cells_x * st->cells_y;
  SCU(cudaMemsetAsync(st->data, 0, n * sizeof(int), st->stream));            // _occupancy_data.clear() (:137)
  SCU(cudaMemsetAsync(st->weight, 0, n * sizeof(int), st->stream));
  SCU(cudaMemsetAsync(st->present, 0, nc * sizeof(int), st->stream));
  SCU(cudaMemsetAsync(st->touched, 0, nc * sizeof(int), st->stream));
  SCU(cudaMemsetAsync(st->dropped, 0, sizeof(unsigned long long), st->stream));
  for (int f = 0; f < st->frames; ++f) {
    const int rc = add_image(st, f, robot_poses + 3 * f, cam);
    if (rc != NIS_OK) return rc;
  }
  SCU(commit(st));
  SCU(cudaStreamSynchronize(st->stream));
  return NIS_OK;
}

int nis_stitcher_frames(const nis_stitcher* st) { return st ? st->frames : 0; }

int nis_stitcher_cell(nis_stitcher* st, int cell_x, int cell_y, int32_t* data, int32_t* weight, int* present) {
  if (!st || !present) return NIS_ERR_INVALID_ARGUMENT;
  SCU(cudaSetDevice(st->device));
  const int cx = cell_x - st->cell_x0, cy = cell_y - st->cell_y0;
  *present = 0;
  if (cx < 0 || cy < 0 || cx >= st->cells_x || cy >= st->cells_y) return NIS_OK;
  const int cell = cy * st->cells_x + cx;
  SCU(cudaMemcpyAsync(present, st->present + cell, sizeof(int), cudaMemcpyDeviceToHost, st->stream));
  SCU(cudaStreamSynchronize(st->stream));
  if (!*present) return NIS_OK;
  const size_t e = (size_t)st->cs * st->cs;
  if (data) SCU(cudaMemcpyAsync(data, st->data + cell * e, e * sizeof(int), cudaMemcpyDeviceToHost, st->stream));
  if (weight) SCU(cudaMemcpyAsync(weight, st->weight + cell * e, e * sizeof(int), cudaMemcpyDeviceToHost, st->stream));
  SCU(cudaStreamSynchronize(st->stream));
  return NIS_OK;
}

int nis_stitcher_dropped(nis_stitcher* st, long long* pixels_outside_window) {
  if (!st || !pixels_outside_window) return NIS_ERR_INVALID_ARGUMENT;
  SCU(cudaSetDevice(st->device));
  unsigned long long v = 0;
  SCU(cudaMemcpyAsync(&v, st->dropped, sizeof v, cudaMemcpyDeviceToHost, st->stream));
  SCU(cudaStreamSynchronize(st->stream));
  *pixels_outside_window = (long long)v;
  return NIS_OK;
}

}  // extern "C"
