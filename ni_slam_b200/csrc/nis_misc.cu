// nis_misc.cu -- the two OpenCV-exact resampling kernels and the per-pair bookkeeping kernels (sm_100a).
//   polar_tma_kernel + rzc_fix_kernel : RemoveZeroComponent + fftshift + cv::warpPolar   (correlation_flow.cc:79-87, :94, :228-236)
//   rotate_kernel  : RotateArray = getRotationMatrix2D + warpAffine   (utils.cc:154-161)
//   polar_select / pose_finalize / scan_reduce : ComputePose control flow and the FindLoopClosure arg-max
//   (correlation_flow.cc:97-138, loop_closure.cc:61-71) kept on the device so a batch never syncs with the host.
// All interpolation arithmetic is written with explicit round-to-nearest intrinsics (no FMA contraction) so the
// fixed-point coordinates and the 4-tap sums equal OpenCV's scalar code bit for bit.
#include <cuda.h>
#include <limits.h>

#include "nis_device.cuh"
#include "nis_internal.h"
#include "nis_warp.cuh"

namespace nis {

// ---------------------------------------------------------------------------------------------------------
// tiled polar gather over a TMA-staged tile.
// A ray of the polar grid crosses image rows, so a direct gather touches ~8 different 128-byte lines per warp request and the L1
// tag stage is the limiter.  Here the inverse column pass before it (EpiStoreShift) writes power = IFFT(|F|) ALREADY fftshift-ed
// into `hp`, rzc_fix_kernel applies RemoveZeroComponent to it in place, and one CTA per polar cell of kPolarTA angles x kPolarTR
// radii loads the bounding box of its source footprint as ONE 2-D tile: cp.async.bulk.tensor (TMA) straight into shared memory,
// completion on an mbarrier, out-of-range taps zero-filled by the TMA unit -- which is exactly cv::warpPolar's
// WARP_FILL_OUTLIERS / BORDER_CONSTANT(0).  While the tile is in flight the threads fetch their table entries; the four taps then
// come from shared memory with the 1/32-px weights of the per-context table.  Same fixed-point arithmetic as polar_pixel, bit for bit.
// Per-context tables (built once): tiles[t] = {y0, x0, bh, bw} of the cell's box in shifted coordinates, and per output pixel
// (dy * pitch + dx) | fx << 16 | fy << 21 relative to that box (pitch = the common tile width).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void polar_fixed_point(int H, int W, double cp, double sp, float rf, int& ix, int& iy, int& fx, int& fy) {
  const float cx = (float)W / 2, cy = (float)H / 2;
  const float mx = (float)__dadd_rn(__dmul_rn((double)rf, cp), (double)cx);
  const float my = (float)__dadd_rn(__dmul_rn((double)rf, sp), (double)cy);
  const int sx = __float2int_rn(mx * 32.f), sy = __float2int_rn(my * 32.f);
  ix = sat_short(sx >> 5); iy = sat_short(sy >> 5); fx = sx & 31; fy = sy & 31;
}

__global__ void __launch_bounds__(256) polar_tile_bbox_kernel(int4* __restrict__ tiles, int H, int W, int D, int Cp,
                                                              const double* __restrict__ cs, const float* __restrict__ rho_tab) {
  __shared__ int red[4];
  if (threadIdx.x == 0) { red[0] = INT_MAX; red[1] = INT_MAX; red[2] = INT_MIN; red[3] = INT_MIN; }
  __syncthreads();
  int ymin = INT_MAX, xmin = INT_MAX, ymax = INT_MIN, xmax = INT_MIN;
  for (int i = threadIdx.x; i < kPolarTA * kPolarTR; i += blockDim.x) {
    const int phi = blockIdx.y * kPolarTA + i / kPolarTR, rho = blockIdx.x * kPolarTR + i % kPolarTR;
    if (phi >= D || rho >= Cp) continue;
    int ix, iy, fx, fy;
    polar_fixed_point(H, W, cs[2 * phi], cs[2 * phi + 1], rho_tab[rho], ix, iy, fx, fy);
    ymin = min(ymin, iy); xmin = min(xmin, ix); ymax = max(ymax, iy + 1); xmax = max(xmax, ix + 1);
  }
  atomicMin(&red[0], ymin); atomicMin(&red[1], xmin); atomicMax(&red[2], ymax); atomicMax(&red[3], xmax);
  __syncthreads();
  if (threadIdx.x == 0) {
    // TMA: the innermost start coordinate must be 16-byte aligned (4 floats); x & ~3 floors negative values too
    // (probed on the B200 with tools/ubench/tma_tile.cu: an unaligned start raises "illegal instruction", an aligned one -- negative or
    // partly outside the image -- zero-fills)
    const int x0 = red[1] & ~3;
    tiles[blockIdx.y * gridDim.x + blockIdx.x] = make_int4(red[0], x0, red[2] - red[0] + 1, red[3] - x0 + 1);
  }
}

__global__ void polar_tile_table_kernel(uint32_t* __restrict__ table, const int4* __restrict__ tiles, int tiles_r, int pitch, int H, int W,
                                        int D, int Cp, const double* __restrict__ cs, const float* __restrict__ rho_tab) {
  const int phi = blockIdx.y, rho = blockIdx.x * blockDim.x + threadIdx.x;
  if (rho >= Cp) return;
  int ix, iy, fx, fy;
  polar_fixed_point(H, W, cs[2 * phi], cs[2 * phi + 1], rho_tab[rho], ix, iy, fx, fy);
  const int4 t = tiles[(phi / kPolarTA) * tiles_r + rho / kPolarTR];
  table[(size_t)phi * Cp + rho] = (uint32_t)((iy - t.x) * pitch + (ix - t.y)) | ((uint32_t)fx << 16) | ((uint32_t)fy << 21);
}

// RemoveZeroComponent (correlation_flow.cc:79-87) on the fftshift-ed image, in place: unshifted row 0 / column 0 are shifted row H/2 /
// column W/2.  y.row(0) = (x.row(1) + x.row(R-1)) / 2, y.col(0) = (x.col(1) + x.col(C-1)) / 2, both from the ORIGINAL x, and y(0,0)
// comes from the column rule: all reads first, then the row, then the column.  One CTA per image.
__global__ void __launch_bounds__(1024) rzc_fix_kernel(Dst<float> hp, int H, int W) {
  float* p = hp.at(blockIdx.x);
  const int hs = H / 2, ws = W / 2;
  auto sh = [&](int r, int c) -> float* {                 // unshifted (r, c) -> its place in the shifted image
    const int y = r + hs - (r + hs >= H ? H : 0), x = c + ws - (c + ws >= W ? W : 0);
    return p + (size_t)y * W + x;
  };
  float rowv[2], colv[2];
  int nr = 0, nc = 0;
  for (int c = threadIdx.x; c < W && nr < 2; c += blockDim.x) rowv[nr++] = __fadd_rn(*sh(1, c), *sh(H - 1, c)) * 0.5f;
  for (int r = threadIdx.x; r < H && nc < 2; r += blockDim.x) colv[nc++] = __fadd_rn(*sh(r, 1), *sh(r, W - 1)) * 0.5f;
  __syncthreads();
  nr = 0;
  for (int c = threadIdx.x; c < W && nr < 2; c += blockDim.x) *sh(0, c) = rowv[nr++];
  __syncthreads();
  nc = 0;
  for (int r = threadIdx.x; r < H && nc < 2; r += blockDim.x) *sh(r, 0) = colv[nc++];
}
int launch_rzc_fix(Dst<float> hp, int H, int W, int B, cudaStream_t s) {
  if (B <= 0) return 0;
  if (W > 2048 || H > 2048) return -1;
  static int attr = set_carveout(rzc_fix_kernel);
  if (attr) return attr;
  rzc_fix_kernel<<<B, 1024, 0, s>>>(hp, H, W);
  return (int)cudaGetLastError();
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(256) polar_tma_kernel(const __grid_constant__ CUtensorMap hp_map, Dst<float> out, int D, int Cp,
                                                        const int4* __restrict__ tiles, const uint32_t* __restrict__ table, int pitch,
                                                        int tile_bytes) {
  extern __shared__ __align__(128) float box[];
  __shared__ __align__(8) unsigned long long mbar;
  const int b = blockIdx.z, tid = threadIdx.x;
  const uint32_t bar = smem_u32(&mbar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    const int4 t = __ldg(&tiles[blockIdx.y * gridDim.x + blockIdx.x]);        // y0, x0, bh, bw
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(tile_bytes) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(box)), "l"(&hp_map), "r"(t.y), "r"(t.x), "r"(b), "r"(bar) : "memory");
  }
  // table entries of this thread's four output pixels while the tile is in flight
  const int rho = blockIdx.x * kPolarTR + (tid % kPolarTR);
  uint32_t e[kPolarTA * kPolarTR / 256];
#pragma unroll
  for (int k = 0; k < kPolarTA * kPolarTR / 256; ++k) {
    const int phi = blockIdx.y * kPolarTA + tid / kPolarTR + k * (256 / kPolarTR);
    e[k] = (rho < Cp && phi < D) ? __ldg(table + (size_t)phi * Cp + rho) : 0u;
  }
  {
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar) : "memory");
  }
  if (rho >= Cp) return;
  float* o = out.at(b);
#pragma unroll
  for (int k = 0; k < kPolarTA * kPolarTR / 256; ++k) {
    const int phi = blockIdx.y * kPolarTA + tid / kPolarTR + k * (256 / kPolarTR);
    if (phi >= D) break;
    const float* q = box + (e[k] & 0xffffu);
    o[(size_t)phi * Cp + rho] = bilinear4(q[0], q[1], q[pitch], q[pitch + 1], (e[k] >> 16) & 31, (e[k] >> 21) & 31);
  }
}

int launch_polar_tile_bbox(int4* tiles, int H, int W, int D, int Cp, const double* cs_table, const float* rho_table, cudaStream_t s) {
  polar_tile_bbox_kernel<<<dim3((Cp + kPolarTR - 1) / kPolarTR, (D + kPolarTA - 1) / kPolarTA), 256, 0, s>>>(tiles, H, W, D, Cp, cs_table, rho_table);
  return (int)cudaGetLastError();
}
int launch_polar_tile_table(uint32_t* table, const int4* tiles, int pitch, int H, int W, int D, int Cp, const double* cs_table,
                            const float* rho_table, cudaStream_t s) {
  polar_tile_table_kernel<<<dim3((Cp + 127) / 128, D), 128, 0, s>>>(table, tiles, (Cp + kPolarTR - 1) / kPolarTR, pitch, H, W, D, Cp, cs_table, rho_table);
  return (int)cudaGetLastError();
}
int launch_polar_tma(const void* hp_map, Dst<float> out, int D, int Cp, const int4* tiles, const uint32_t* table, int pitch, int box_rows, int B,
                     cudaStream_t s) {
  if (B <= 0) return 0;
  const int tile_bytes = box_rows * pitch * (int)sizeof(float);
  static int attr = set_carveout(polar_tma_kernel);
  if (attr) return attr;
  static int attr_bytes = 0;
  if (tile_bytes > 48 * 1024 && tile_bytes > attr_bytes) {
    const cudaError_t e = cudaFuncSetAttribute(polar_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tile_bytes);
    if (e != cudaSuccess) return (int)e;
    attr_bytes = tile_bytes;
  }
  polar_tma_kernel<<<dim3((Cp + kPolarTR - 1) / kPolarTR, (D + kPolarTA - 1) / kPolarTA, B), 256, tile_bytes, s>>>(
      *reinterpret_cast<const CUtensorMap*>(hp_map), out, D, Cp, tiles, table, pitch, tile_bytes);
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// rotation (warpAffine, INTER_LINEAR, BORDER_WRAP, AB_BITS = 10, INTER_BITS = 5)
// ---------------------------------------------------------------------------------------------------------
template <bool U8>
__global__ void __launch_bounds__(256) rotate_kernel(Src<float> img_f32, Src<uint8_t> img_u8, const float* __restrict__ lut,
                                                     Dst<float> out, int H, int W, const double* __restrict__ mats,
                                                     const int* __restrict__ sel) {
  const int e = blockIdx.z;
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= W || y >= H) return;
  const double* M = mats + 6 * (size_t)sel[e];
  int X0, Y0;
  rotate_row_setup(M, y, X0, Y0);
  out.at(e)[(size_t)y * W + x] = rotate_pixel<U8>(U8 ? nullptr : img_f32.at(e), U8 ? img_u8.at(e) : nullptr, lut, H, W, M, X0, Y0, x);
}

int launch_rotate(Src<float> img_f32, Src<uint8_t> img_u8, const float* lut, Dst<float> out, int H, int W, const double* mats,
                  const int* sel, int E, cudaStream_t s) {
  if (E <= 0) return 0;
  const dim3 grid((W + 31) / 32, (H + 7) / 8, E);
  if (img_u8.base || img_u8.ptrs) rotate_kernel<true><<<grid, 256, 0, s>>>(img_f32, img_u8, lut, out, H, W, mats, sel);
  else rotate_kernel<false><<<grid, 256, 0, s>>>(img_f32, img_u8, lut, out, H, W, mats, sel);
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// Camera::UndistortImage (src/camera.cc:92-93): cv::remap(u8, map1 CV_16SC2, map2 CV_16UC1, INTER_LINEAR), BORDER_CONSTANT(0).
// OpenCV's u8 path is pure integer: weights (1-fy)(1-fx)... scaled to 2^15 (exact multiples of 32), rounded with
// (sum + 2^14) >> 15 -- reproduced bit for bit.  map1 = (x, y) integer source position, map2 = fy*32 + fx.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) undistort_kernel(Src<uint8_t> raw, Dst<uint8_t> out, int H, int W,
                                                        const short2* __restrict__ map1, const unsigned short* __restrict__ map2) {
  const int b = blockIdx.z;
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= W || y >= H) return;
  const short2 xy = __ldg(map1 + (size_t)y * W + x);
  const int a = __ldg(map2 + (size_t)y * W + x) & 1023;
  const int fx = a & 31, fy = a >> 5;
  const int sx = xy.x, sy = xy.y;
  const uint8_t* p = raw.at(b);
  auto tap = [&](int yy, int xx) -> int {
    return ((unsigned)xx < (unsigned)W && (unsigned)yy < (unsigned)H) ? (int)__ldg(p + yy * W + xx) : 0;
  };
  const int w0 = (32 - fy) * (32 - fx) * 32, w1 = (32 - fy) * fx * 32, w2 = fy * (32 - fx) * 32, w3 = fy * fx * 32;
  const int v = tap(sy, sx) * w0 + tap(sy, sx + 1) * w1 + tap(sy + 1, sx) * w2 + tap(sy + 1, sx + 1) * w3;
  out.at(b)[(size_t)y * W + x] = (uint8_t)min(255, max(0, (v + (1 << 14)) >> 15));
}
int launch_undistort(Src<uint8_t> raw, Dst<uint8_t> out, int H, int W, const void* map1, const void* map2, int B, cudaStream_t s) {
  if (B <= 0) return 0;
  undistort_kernel<<<dim3((W + 31) / 32, (H + 7) / 8, B), 256, 0, s>>>(raw, out, H, W, (const short2*)map1, (const unsigned short*)map2);
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// gaussian kernel helper: xf.square().abs().sum()/N over the stored half spectrum (correlation_flow.cc:184-185)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) spec_sqsum_kernel(Src<cpx> x, int count, double* out) {
  const int b = blockIdx.x;
  const cpx* p = x.at(b);
  double acc = 0.0;
  for (int i = threadIdx.x; i < count; i += blockDim.x) {
    const cpx v = __ldg(p + i);
    acc += (double)hypotf(v.x * v.x - v.y * v.y, 2.f * v.x * v.y);
  }
  __shared__ double red[8];
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int i = 0; i < 8; ++i) t += red[i];
    out[b] = t;          // raw sum; the consumer divides by n in f32 like the reference
  }
}
int launch_spec_sqsum(Src<cpx> x, int count, double* out, int B, cudaStream_t s) {
  if (B <= 0) return 0;
  static int attr = set_carveout(spec_sqsum_kernel);
  if (attr) return attr;
  spec_sqsum_kernel<<<B, 256, 0, s>>>(x, count, out);
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// ComputePose bookkeeping
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void key_to_rc(unsigned long long key, int R, int& row, int& col, float& val) {
  const uint32_t idx = 0xffffffffu - (uint32_t)(key & 0xffffffffull);
  col = (int)(idx / (uint32_t)R);
  row = (int)(idx % (uint32_t)R);
  val = ord2f((uint32_t)(key >> 32));
}

// GetInfo (correlation_flow.cc:238-243) from sum / sum of squares:  mean((g-m)^2) = (S2 - 2 m S1 + n m^2)/n
__device__ __forceinline__ float get_info(const PeakStats& st, float peak, double n) {
  const float m = ((float)st.sum - peak) / (float)(n - 1.0);
  const double md = (double)m;
  double var = (st.sumsq - 2.0 * md * st.sum + n * md * md) / n;
  var = var > 0.0 ? var : 0.0;
  const float sd = sqrtf((float)var);
  return (float)((double)(peak - m) / ((double)sd + 1e-7));
}

__global__ void polar_select_kernel(const PeakStats* __restrict__ polar, int D, int loop_mode, int* __restrict__ sel, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int row, col; float v;
  key_to_rc(polar[b].key, D, row, col, v);
  if (loop_mode) { sel[2 * b] = D + row; sel[2 * b + 1] = 2 * D + row; }
  else sel[b] = row;
}
int launch_polar_select(const PeakStats* polar, int D, int loop_mode, int* sel, int B, cudaStream_t s) {
  if (B <= 0) return 0;
  static int attr = set_carveout(polar_select_kernel);
  if (attr) return attr;
  polar_select_kernel<<<(B + 127) / 128, 128, 0, s>>>(polar, D, loop_mode, sel, B);
  return (int)cudaGetLastError();
}

__global__ void pose_finalize_kernel(const PeakStats* __restrict__ polar, const PeakStats* __restrict__ trans, AngleTables tabs,
                                     int H, int W, int D, int Cp, int loop_mode, int index0, PoseRecord* __restrict__ out, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int prow, pcol; float pval;
  key_to_rc(polar[b].key, D, prow, pcol, pval);
  const float info_rot = get_info(polar[b], pval, (double)D * Cp);
  int trow, tcol, hyp = 0, variant = 0; float tval, info_trans;
  if (!loop_mode) {
    key_to_rc(trans[b].key, H, trow, tcol, tval);
    info_trans = get_info(trans[b], tval, (double)H * W);
  } else {
    int r0, c0, r1, c1; float v0, v1;
    key_to_rc(trans[2 * b].key, H, r0, c0, v0);
    key_to_rc(trans[2 * b + 1].key, H, r1, c1, v1);
    const float i0 = get_info(trans[2 * b], v0, (double)H * W), i1 = get_info(trans[2 * b + 1], v1, (double)H * W);
    if (i0 > i1) { info_trans = i0; trow = r0; tcol = c0; hyp = 0; variant = 1; }     // correlation_flow.cc:121
    else { info_trans = i1; trow = r1; tcol = c1; hyp = 1; variant = 2; }
  }
  PoseRecord r;
  r.pose[0] = (double)(-(tcol - W / 2));      // pose[0] = trans[1] = -(col - width/2)
  r.pose[1] = (double)(-(trow - H / 2));      // pose[1] = trans[0] = -(row - height/2)
  r.pose[2] = tabs.theta[(size_t)variant * D + prow];
  r.info[0] = (double)info_trans; r.info[1] = (double)info_trans; r.info[2] = (double)info_rot;
  r.peak[0] = prow; r.peak[1] = pcol; r.peak[2] = trow; r.peak[3] = tcol;
  r.hyp = hyp; r.index = index0 + b;
  out[b] = r;
}
int launch_pose_finalize(const PeakStats* polar, const PeakStats* trans, AngleTables tabs, int H, int W, int D, int Cp,
                         int loop_mode, int index0, PoseRecord* out, int B, cudaStream_t s) {
  if (B <= 0) return 0;
  static int attr = set_carveout(pose_finalize_kernel);
  if (attr) return attr;
  pose_finalize_kernel<<<(B + 127) / 128, 128, 0, s>>>(polar, trans, tabs, H, W, D, Cp, loop_mode, index0, out, B);
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// boundary layout conversion on the device: reference column-major (C lines of R) <-> internal row-major [R][C].
// 32 x 32 tiles through shared memory; T = float (images) or cpx (half spectra).
// ---------------------------------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(256) transpose_kernel(const T* __restrict__ in, T* __restrict__ out, int rows_in, int cols_in) {
  __shared__ T tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8)
    if (r0 + i < rows_in && c0 + tx < cols_in) tile[i][tx] = in[(size_t)(r0 + i) * cols_in + c0 + tx];
  __syncthreads();
  for (int i = ty; i < 32; i += 8)
    if (c0 + i < cols_in && r0 + tx < rows_in) out[(size_t)(c0 + i) * rows_in + r0 + tx] = tile[tx][i];
}
int launch_transpose_f32(const float* in, float* out, int rows_in, int cols_in, cudaStream_t s) {
  transpose_kernel<float><<<dim3((cols_in + 31) / 32, (rows_in + 31) / 32), 256, 0, s>>>(in, out, rows_in, cols_in);
  return (int)cudaGetLastError();
}
int launch_transpose_cpx(const cpx* in, cpx* out, int rows_in, int cols_in, cudaStream_t s) {
  transpose_kernel<cpx><<<dim3((cols_in + 31) / 32, (rows_in + 31) / 32), 256, 0, s>>>(in, out, rows_in, cols_in);
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// candidate selection on the device (loop_closure.cc:43-53 filters; map.cc:81-101 + loop_closure.cc:17-34 grid neighbourhood).
// Input i in [0, n_in): slot = list ? list[i] : i.  A slot passes when the frame-gap and accumulated-distance filters keep it and,
// with a prior cell, when its cell lies in the 3 x 3 neighbourhood; its rank is the reference's iteration order over that
// neighbourhood (dx = -1..1 outer, dy = -1..1 inner), 0 without a prior.  Output: slots ordered by (rank, i), the input position
// of each, and the count -- three kernels (per-chunk counts, one-CTA exclusive scan, ordered write), no host loop over the DB.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int select_rank(const SelectArgs& a, int i) {
  const int slot = a.list ? a.list[i] : i;
  if (a.frame_gap_thr > 0 && abs(a.query_id - a.frame_id[slot]) < a.frame_gap_thr) return -1;
  if (a.distance_thr > 0 && fabs(a.query_dist - a.dist[slot]) < a.distance_thr) return -1;
  if (!a.use_prior) return 0;
  const int2 c = a.cell[slot];
  if (c.x == INT_MIN) return -1;                                  // never filed in the grid
  const int dx = c.x - a.cx, dy = c.y - a.cy;
  if (dx < -1 || dx > 1 || dy < -1 || dy > 1) return -1;
  return (dx + 1) * 3 + (dy + 1);
}
constexpr int kSelChunk = 2048, kSelThreads = 256, kSelPer = kSelChunk / kSelThreads;

__global__ void __launch_bounds__(kSelThreads) select_count_kernel(SelectArgs a, int* __restrict__ counts, int nchunks) {
  __shared__ int c[9];
  if (threadIdx.x < 9) c[threadIdx.x] = 0;
  __syncthreads();
  const int base = blockIdx.x * kSelChunk;
  for (int k = threadIdx.x; k < kSelChunk; k += kSelThreads) {
    const int i = base + k;
    if (i < a.n_in) { const int r = select_rank(a, i); if (r >= 0) atomicAdd(&c[r], 1); }
  }
  __syncthreads();
  if (threadIdx.x < 9) counts[threadIdx.x * nchunks + blockIdx.x] = c[threadIdx.x];
}
// exclusive scan of counts[9 * nchunks] in (rank, chunk) order; total -> *n_out
__global__ void __launch_bounds__(1024) select_scan_kernel(int* __restrict__ counts, int total_entries, int* __restrict__ n_out) {
  __shared__ int part[1024];
  const int per = (total_entries + 1023) / 1024;
  const int lo = min(threadIdx.x * per, total_entries), hi = min(lo + per, total_entries);
  int s = 0;
  for (int i = lo; i < hi; ++i) s += counts[i];
  part[threadIdx.x] = s;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const int v = threadIdx.x >= o ? part[threadIdx.x - o] : 0;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  int run = part[threadIdx.x] - s;
  for (int i = lo; i < hi; ++i) { const int v = counts[i]; counts[i] = run; run += v; }
  if (threadIdx.x == 1023) *n_out = part[1023];
}
__global__ void __launch_bounds__(kSelThreads) select_write_kernel(SelectArgs a, const int* __restrict__ offsets, int nchunks,
                                                                   int* __restrict__ cand, int* __restrict__ pos) {
  __shared__ int wsum[kSelThreads / 32];
  const int base = blockIdx.x * kSelChunk + threadIdx.x * kSelPer;       // each thread owns kSelPer consecutive inputs
  int rk[kSelPer];
#pragma unroll
  for (int k = 0; k < kSelPer; ++k) rk[k] = (base + k < a.n_in) ? select_rank(a, base + k) : -1;
  const int nranks = a.use_prior ? 9 : 1;
  for (int r = 0; r < nranks; ++r) {
    int mine = 0;
#pragma unroll
    for (int k = 0; k < kSelPer; ++k) mine += (rk[k] == r);
    int incl = mine;                                                      // inclusive scan over the CTA's threads
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if ((threadIdx.x & 31) >= o) incl += v; }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
    __syncthreads();
    int before = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) before += wsum[w];
    __syncthreads();
    int o = offsets[r * nchunks + blockIdx.x] + before + incl - mine;
#pragma unroll
    for (int k = 0; k < kSelPer; ++k)
      if (rk[k] == r) { const int i = base + k; cand[o] = a.list ? a.list[i] : i; pos[o] = i; ++o; }
  }
}
int select_scratch_ints(int n_in) { return 9 * ((n_in + kSelChunk - 1) / kSelChunk) + 1; }
int launch_select(SelectArgs a, int* scratch, int* cand, int* pos, int* n_out, cudaStream_t s) {
  const int nchunks = (a.n_in + kSelChunk - 1) / kSelChunk;
  if (nchunks <= 0) return (int)cudaMemsetAsync(n_out, 0, sizeof(int), s);
  select_count_kernel<<<nchunks, kSelThreads, 0, s>>>(a, scratch, nchunks);
  select_scan_kernel<<<1, 1024, 0, s>>>(scratch, 9 * nchunks, n_out);
  select_write_kernel<<<nchunks, kSelThreads, 0, s>>>(a, scratch, nchunks, cand, pos);
  return (int)cudaGetLastError();
}

// best = first record with the strictly largest response.sum(); initial best = (-1,-1,-1) (loop_closure.h:15)
__global__ void __launch_bounds__(256) scan_reduce_kernel(const PoseRecord* __restrict__ recs, int n, PoseRecord* __restrict__ best) {
  __shared__ double ssum[256];
  __shared__ int sidx[256];
  double bs = -3.0; int bi = -1;
  for (int i = threadIdx.x; i < n; i += 256) {
    const double s = recs[i].info[0] + recs[i].info[1] + recs[i].info[2];
    if (s > bs) { bs = s; bi = i; }          // per-thread indices ascend, so '>' keeps the first
  }
  ssum[threadIdx.x] = bs; sidx[threadIdx.x] = bi;
  __syncthreads();
  for (int o = 128; o; o >>= 1) {
    if (threadIdx.x < o) {
      const double s2 = ssum[threadIdx.x + o]; const int i2 = sidx[threadIdx.x + o];
      const double s1 = ssum[threadIdx.x]; const int i1 = sidx[threadIdx.x];
      const bool take = (i2 >= 0) && (i1 < 0 || s2 > s1 || (s2 == s1 && i2 < i1));
      if (take) { ssum[threadIdx.x] = s2; sidx[threadIdx.x] = i2; }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (sidx[0] >= 0) *best = recs[sidx[0]];
    else {
      PoseRecord r;
      r.pose[0] = r.pose[1] = r.pose[2] = 0.0;
      r.info[0] = r.info[1] = r.info[2] = -1.0;
      r.peak[0] = r.peak[1] = r.peak[2] = r.peak[3] = -1; r.hyp = 0; r.index = -1;
      *best = r;
    }
  }
}
int launch_scan_reduce(const PoseRecord* recs, int n, PoseRecord* best, cudaStream_t s) {
  scan_reduce_kernel<<<1, 256, 0, s>>>(recs, n, best);
  return (int)cudaGetLastError();
}

}  // namespace nis
