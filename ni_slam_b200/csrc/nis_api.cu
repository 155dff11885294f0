// nis_api.cu -- context, keyframe database and the C ABI (include/nislam.h) over the sm_100a kernels.
// Host-side mirror of CorrelationFlow (src/correlation_flow.cc) and LoopClosure (src/loop_closure.cc): the host
// only builds constant tables, sizes batches and enqueues kernels; all per-pixel arithmetic runs on the GPU and
// there is no CPU fallback.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <climits>
#include <map>
#include <string>
#include <vector>

#include "../../include/nislam.h"
#include "nis_internal.h"
#include "../host/pose_math.hpp"

using namespace nis;

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

// -------------------------------------------------------------------------------------------------------------
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  // Grows the buffer (contents are scratch and not preserved).  Transactional: the new block is allocated first and the old one
  // is kept if that fails, so a caller that retries with a smaller size after NIS_ERR_OUT_OF_MEMORY still owns a valid buffer of
  // `bytes` bytes.  Only when both blocks cannot coexist is the old one released before a second attempt; if that fails too the
  // buffer is left empty (p == nullptr, bytes == 0) and every capacity derived from `bytes` reads 0.
  int reserve(size_t n) {
    if (n <= bytes) return 0;
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, n);
    if (e != cudaSuccess && p) {
      (void)cudaGetLastError();
      cudaFree(p);
      p = nullptr; bytes = 0;
      e = cudaMalloc(&q, n);
    }
    if (e != cudaSuccess) return (int)e;
    if (p) cudaFree(p);
    p = q; bytes = n;
    return 0;
  }
  void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct nis_frame {
  uint8_t* img_u8 = nullptr;   // exactly one of img_u8 / img_f32 is set (row-major H x W)
  float* img_f32 = nullptr;
  cpx* F = nullptr;            // fft_result  [H/2+1][W]
  cpx* P = nullptr;            // fft_polar   [D/2+1][Cp]
  cpx* Ht = nullptr;           // T/(kernel(F)/max + lambda): the keyframe-only factor of EstimateTrans, translation stage
  cpx* Hp = nullptr;           // same for the polar stage
  void* block = nullptr;       // single allocation backing all of the above
  // what the buffers hold: frames made by nis_features_* carry everything, imported ones what the caller handed over
  bool has_image = true, has_spectra = true /* F */, has_polar = true /* P */, has_h = true /* Ht, Hp */;
};

struct SizeClass {             // one 2-D transform size: R rows (halved) x C cols
  int R, C;
  size_t spec, real;           // elements
  Twiddles col;
  RowTwiddles row;             // plan A and plan B tables (b = a when the length has no plan B)
};

// One lane = one CUDA stream with its own workspace.  Batches are dealt round-robin to the lanes so that several small
// (L2-resident) batches are in flight at once and the tail of one launch overlaps the head of another.
struct Lane {
  cudaStream_t stream = nullptr;
  cudaEvent_t ev = nullptr;
  int cap = 0;                 // workspace capacity in pairs
  DevBuf t1, real, pol, maxp, maxt, maxh, stats_p, stats_t, sel, xx, zz;
  CUtensorMap hp_map;          // TMA descriptor over `real` = the lane's fftshift-ed power images [cap][H][W], box = one polar cell's footprint
  DevBuf dbrec;                // scan scratch of the compact store modes: F, P, Ht, Hp of one batch of candidates
  int db_cap = 0;
};

struct nis_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;   // == lanes[0].stream: the stream callers time on; the other lanes fork from / join into it
  std::vector<Lane> lanes;
  int active_lanes = 1;            // lanes batches are dealt to (<= lanes.size())
  cudaStream_t prof_stream = nullptr;
  cudaEvent_t fork_ev = nullptr;
  std::vector<void*> frame_pool;               // freed nis_frame blocks, reused by the next frame_alloc
  std::vector<nis_frame*> live_frames;          // frames handed out and not yet freed: released by nis_destroy
  cudaStream_t copy_stream = nullptr;          // host->device uploads of a stream run back to back here, ahead of the compute lanes
  std::vector<cudaEvent_t> up_ev, feat_ev;     // per batch of a stream: upload done / features done
  nis_cf_config cfg{};
  int H = 0, W = 0, D = 0, Cp = 0;
  SizeClass sz[2];             // 0: H x W, 1: D x Cp
  size_t maxspec = 0, maxreal = 0;
  std::string err;
  long long launches = 0;
  int batch = 16, default_batch = 16;
  int ptile_pitch = 0, ptile_rows = 0;
  // constant tables
  DevBuf tw, lut, cs, rho, mats, theta, ptiles, ptab2, rowtab;
  DevBuf recs, best, cand, stage, qgather;
  // multi-GPU scan (nis_comm_init / nis_loop_scan_sharded): NCCL communicator of this context's rank
  void* nccl_comm = nullptr;
  int rank = 0, n_ranks = 1;
  int recs_cap = 0, cand_cap = 0;
  // stream slabs
  DevBuf sF, sP, sHt, sHp, sImg, sUnd, kfrec;
  // undistort front end (Camera::UndistortImage): fixed-point remap maps handed over by the caller
  DevBuf umap1, umap2;
  bool undistort = false;
  // scan: spectra of the query image rotated by every angle the polar stage can select (2 hypotheses x D rows), built once
  // per query when the candidate list is long; indexed by the same `sel` the rotation matrices use (minus D)
  DevBuf rotc, rotc_xx, rotc_sel;
  bool use_rot_cache = false;
  int rot_cache_min = 1024;
  // keyframe DB
  std::vector<void*> chunks;
  int chunk_slots = 64;
  std::vector<void*> slot_ptr;         // host copy of the device pointer table (record start per slot)
  int db_mode = 0;                     // NIS_DB_FULL / NIS_DB_SPECTRA / NIS_DB_IMAGE
  bool meta_dirty = true;              // frame ids / distances / cells changed since the last upload
  DevBuf d_fid, d_dist, d_cell, cand_in, cand_pos, sel_scratch;
  std::vector<int> slot_frame_id;
  std::vector<double> slot_dist;
  std::vector<std::pair<int, int>> slot_cell;    // grid cell at insertion time (Map::AddFrame), INT_MIN = not filed
  DevBuf d_slot_ptr;
  int d_slot_cap = 0;
  // pinned staging
  void* pin = nullptr; size_t pin_bytes = 0;
  // optional per-kernel-family event timing (nis_profile_begin / nis_profile_end)
  bool profiling = false;
  struct ProfEv { const char* name; cudaEvent_t a, b; };
  std::vector<ProfEv> prof;
};

static void prof_before(nis_ctx* ctx, const char* name) {
  nis_ctx::ProfEv ev{name, nullptr, nullptr};
  cudaEventCreate(&ev.a); cudaEventCreate(&ev.b);
  cudaEventRecord(ev.a, ctx->prof_stream ? ctx->prof_stream : ctx->stream);
  ctx->prof.push_back(ev);
}
static void prof_after(nis_ctx* ctx) { cudaEventRecord(ctx->prof.back().b, ctx->prof_stream ? ctx->prof_stream : ctx->stream); }

static int fail(nis_ctx* c, int status, const char* what, int cuda_err = 0) {
  if (c) {
    char buf[512];
    if (cuda_err > 0) {
      snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString((cudaError_t)cuda_err));
      (void)cudaGetLastError();          // a failed cudaMalloc leaves a non-sticky last error: clear it so the next launch check is not blamed for it
    }
    else snprintf(buf, sizeof buf, "%s", what);
    c->err = buf;
  }
  return status;
}

#define CU(call)                                                              \
  do {                                                                        \
    cudaError_t e_ = (call);                                                  \
    if (e_ != cudaSuccess) return fail(ctx, NIS_ERR_CUDA, #call, (int)e_);    \
  } while (0)
// kernel launch through a launcher returning cudaError_t-as-int (or -1 for an unsupported size)
#define LAUNCH(call)                                                                      \
  do {                                                                                    \
    if (ctx->profiling) prof_before(ctx, #call);                                          \
    int e_ = (call);                                                                      \
    if (ctx->profiling) prof_after(ctx);                                                  \
    if (e_ == -1) return fail(ctx, NIS_ERR_UNSUPPORTED_SIZE, "unsupported transform size: " #call); \
    if (e_ != 0) return fail(ctx, NIS_ERR_CUDA, #call, e_);                                \
    ctx->launches++;                                                                      \
  } while (0)
#define RESERVE(buf, n)                                                                   \
  do {                                                                                    \
    int e_ = (buf).reserve(n);                                                            \
    if (e_) return fail(ctx, NIS_ERR_OUT_OF_MEMORY, "cudaMalloc " #buf, e_);               \
  } while (0)
#define TRY(call)                \
  do {                           \
    int s_ = (call);             \
    if (s_ != NIS_OK) return s_; \
  } while (0)

// Host->device copies go on the context's (non-blocking) stream: a plain cudaMemcpy from pageable memory may return
// before the DMA lands and is not ordered against kernels on a non-blocking stream.
static cudaError_t h2d(nis_ctx* ctx, void* dst, const void* src, size_t bytes) {
  cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream);
  if (e != cudaSuccess) return e;
  return cudaStreamSynchronize(ctx->stream);
}

template <class T> static Src<T> src_slab(const T* base, long long stride, int shift = 0) {
  return Src<T>{base, stride, nullptr, 0, nullptr, shift};
}
template <class T> static Src<T> src_null() { return Src<T>{nullptr, 0, nullptr, 0, nullptr, 0}; }

// -------------------------------------------------------------------------------------------------------------
// constant tables
// -------------------------------------------------------------------------------------------------------------
static void host_twiddles(const int r[3], std::vector<cpx>& out, size_t& off1, size_t& off2) {     // row passes (Stockham)
  const int R0 = r[0], R1 = r[1], R2 = r[2], N = R0 * R1 * R2;
  off1 = out.size();
  out.resize(off1 + (size_t)std::max(R1 - 1, 1) * R0, make_float2(1.f, 0.f));
  for (int q = 1; q < R1; ++q)
    for (int k = 0; k < R0; ++k) {
      const double a = -2.0 * M_PI * (double)q * k / (double)(R0 * R1);
      out[off1 + (size_t)(q - 1) * R0 + k] = make_float2((float)cos(a), (float)sin(a));
    }
  off2 = out.size();
  out.resize(off2 + (size_t)std::max(R2 - 1, 1) * R0 * R1, make_float2(1.f, 0.f));
  for (int q = 1; q < R2; ++q)
    for (int k = 0; k < R0 * R1; ++k) {
      const double a = -2.0 * M_PI * (double)q * k / (double)N;
      out[off2 + (size_t)(q - 1) * R0 * R1 + k] = make_float2((float)cos(a), (float)sin(a));
    }
}
// column passes (in place, see nis_fft.cuh): tw1[(b-1)*C + d2] = exp(-2 pi i b d2 / (B C)), tw2[(a-1)*BC + j] = exp(-2 pi i a j / N)
static void host_twiddles_col(const int r[3], std::vector<cpx>& out, size_t& off1, size_t& off2) {
  const int A = r[0], B = r[1], C = r[2], N = A * B * C;
  off1 = out.size();
  out.resize(off1 + (size_t)std::max(B - 1, 1) * C, make_float2(1.f, 0.f));
  for (int q = 1; q < B; ++q)
    for (int k = 0; k < C; ++k) {
      const double a = -2.0 * M_PI * (double)q * k / (double)(B * C);
      out[off1 + (size_t)(q - 1) * C + k] = make_float2((float)cos(a), (float)sin(a));
    }
  off2 = out.size();
  out.resize(off2 + (size_t)std::max(A - 1, 1) * B * C, make_float2(1.f, 0.f));
  for (int q = 1; q < A; ++q)
    for (int k = 0; k < B * C; ++k) {
      const double a = -2.0 * M_PI * (double)q * k / (double)N;
      out[off2 + (size_t)(q - 1) * B * C + k] = make_float2((float)cos(a), (float)sin(a));
    }
}

// utils.cc:173-175
static double normalize_degree(double a) { return a - 360.0 * floor((a + 180.0) / 360.0); }

// cv::getRotationMatrix2D(Point2f(W/2., H/2.), degree, 1) followed by warpAffine's inversion (utils.cc:157-159)
static void rotation_inverse(int H, int W, double degree, double* M) {
  const float cxf = (float)(W / 2.), cyf = (float)(H / 2.);
  const double a = degree * (M_PI / 180.0);
  const double alpha = cos(a), beta = sin(a);
  M[0] = alpha; M[1] = beta; M[2] = (1 - alpha) * cxf - beta * cyf;
  M[3] = -beta; M[4] = alpha; M[5] = beta * cxf + (1 - alpha) * cyf;
  double Dt = M[0] * M[4] - M[1] * M[3];
  Dt = Dt != 0 ? 1. / Dt : 0;
  const double A11 = M[4] * Dt, A22 = M[0] * Dt;
  M[0] = A11; M[1] *= -Dt; M[3] *= -Dt; M[4] = A22;
  const double b1 = -M[0] * M[2] - M[1] * M[5];
  const double b2 = -M[3] * M[2] - M[4] * M[5];
  M[2] = b1; M[5] = b2;
}

static void rotate_row_table(const double* M, int H, int2* out) {
  for (int y = 0; y < H; ++y) {
    out[y].x = (int)lrint((M[1] * (double)y + M[2]) * 1024.0) + 16;
    out[y].y = (int)lrint((M[4] * (double)y + M[5]) * 1024.0) + 16;
  }
}

static int build_tables(nis_ctx* ctx) {
  const int H = ctx->H, W = ctx->W, D = ctx->D, Cp = ctx->Cp;
  // twiddles
  std::vector<cpx> tw;
  size_t o[8], ob[4];
  bool has_b[2];
  int r[3];
  for (int s = 0; s < 2; ++s) {
    SizeClass& z = ctx->sz[s];
    plan_radices_col(z.R, r); host_twiddles_col(r, tw, o[4 * s + 0], o[4 * s + 1]);
    plan_radices_row(z.C, r); host_twiddles(r, tw, o[4 * s + 2], o[4 * s + 3]);
    has_b[s] = plan_radices_row_b(z.C, r);
    if (has_b[s]) host_twiddles(r, tw, ob[2 * s], ob[2 * s + 1]);
  }
  RESERVE(ctx->tw, tw.size() * sizeof(cpx));
  CU(h2d(ctx, ctx->tw.p, tw.data(), tw.size() * sizeof(cpx)));
  const cpx* base = ctx->tw.as<cpx>();
  for (int s = 0; s < 2; ++s) {
    SizeClass& z = ctx->sz[s];
    z.col = Twiddles{base + o[4 * s + 0], base + o[4 * s + 1]};
    z.row.a = Twiddles{base + o[4 * s + 2], base + o[4 * s + 3]};
    z.row.b = has_b[s] ? Twiddles{base + ob[2 * s], base + ob[2 * s + 1]} : z.row.a;
  }
  // u8 -> f32/255 (utils.cc:117: matrix.array()/255.0)
  float lut[256];
  for (int u = 0; u < 256; ++u) lut[u] = (float)((double)(float)u / 255.0);
  RESERVE(ctx->lut, sizeof lut);
  CU(h2d(ctx, ctx->lut.p, lut, sizeof lut));
  // warpPolar angle / radius tables (correlation_flow.cc:228-236 -> cv::warpPolar)
  std::vector<double> cs(2 * (size_t)D);
  const double Kangle = 2.0 * M_PI / D;
  for (int phi = 0; phi < D; ++phi) { cs[2 * phi] = cos(Kangle * phi); cs[2 * phi + 1] = sin(Kangle * phi); }
  std::vector<float> rho(Cp);
  const double maxRadius = (double)std::min(H / 2, W / 2), Kmag = maxRadius / Cp;
  for (int q = 0; q < Cp; ++q) rho[q] = (float)(q * Kmag);
  RESERVE(ctx->cs, cs.size() * sizeof(double));
  RESERVE(ctx->rho, rho.size() * sizeof(float));
  CU(h2d(ctx, ctx->cs.p, cs.data(), cs.size() * sizeof(double)));
  CU(h2d(ctx, ctx->rho.p, rho.data(), rho.size() * sizeof(float)));
  {
    // polar cells (kPolarTA angles x kPolarTR radii): bounding box of every cell's source footprint; all cells load a tile of the common
    // size pitch x rows through TMA (out-of-range parts arrive as zeros), the per-pixel table holds offsets inside the cell's tile
    const int tr = (Cp + kPolarTR - 1) / kPolarTR, ta = (D + kPolarTA - 1) / kPolarTA;
    RESERVE(ctx->ptiles, (size_t)tr * ta * sizeof(int4));
    RESERVE(ctx->ptab2, (size_t)D * Cp * sizeof(uint32_t));
    int e_ = launch_polar_tile_bbox(ctx->ptiles.as<int4>(), H, W, D, Cp, ctx->cs.as<double>(), ctx->rho.as<float>(), ctx->stream);
    if (e_ != 0) return fail(ctx, NIS_ERR_CUDA, "launch_polar_tile_bbox", e_);
    std::vector<int4> tiles((size_t)tr * ta);
    CU(cudaMemcpyAsync(tiles.data(), ctx->ptiles.p, tiles.size() * sizeof(int4), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    int bh = 1, bw = 1;
    for (const int4& t : tiles) { bh = std::max(bh, t.z); bw = std::max(bw, t.w); }
    int pitch = (bw + 3) & ~3;                                       // TMA: the inner box extent must be a multiple of 16 bytes
    if (pitch % 8 == 0) pitch += 4;                                  // pitch = 4 (mod 8): rows of a tile repeat their banks only every 8th row
    ctx->ptile_pitch = pitch;
    ctx->ptile_rows = bh;
    if (pitch > 256 || bh > 256 || (size_t)bh * pitch > 65535 || (size_t)bh * pitch * sizeof(float) > 160 * 1024)
      return fail(ctx, NIS_ERR_UNSUPPORTED_SIZE, "polar cell footprint too large for a TMA tile");
    e_ = launch_polar_tile_table(ctx->ptab2.as<uint32_t>(), ctx->ptiles.as<int4>(), ctx->ptile_pitch, H, W, D, Cp, ctx->cs.as<double>(),
                                 ctx->rho.as<float>(), ctx->stream);
    if (e_ != 0) return fail(ctx, NIS_ERR_CUDA, "launch_polar_tile_table", e_);
    CU(cudaStreamSynchronize(ctx->stream));
  }
  // per polar-peak-row angle tables (correlation_flow.cc:105-136): every float/double step of the reference, once
  std::vector<double> mats(3 * (size_t)D * 6), theta(3 * (size_t)D);
  for (int row = 0; row < D; ++row) {
    const int rots0 = -(row - D / 2);                                   // :176
    float degree = (float)((double)rots0 * (2.0 / D) * 180);            // :105
    degree = (float)normalize_degree((double)degree);                   // :106
    // tracking (:108-109)
    float dt = fabsf(degree) > 90 ? degree - 180 : degree;
    rotation_inverse(H, W, (double)(-dt), &mats[(0 * (size_t)D + row) * 6]);
    float fin = dt > 180 ? dt - 360 : dt;                               // :134
    theta[0 * (size_t)D + row] = (double)(float)((double)(fin / 180) * M_PI);
    // loop, "-deg" (:116) and "-deg+180" (:117)
    rotation_inverse(H, W, (double)(-degree), &mats[(1 * (size_t)D + row) * 6]);
    rotation_inverse(H, W, (double)(-degree + 180), &mats[(2 * (size_t)D + row) * 6]);
    float d0 = degree; d0 = d0 > 180 ? d0 - 360 : d0;
    float d1 = degree + 180; d1 = d1 > 180 ? d1 - 360 : d1;             // :130, :134
    theta[1 * (size_t)D + row] = (double)(float)((double)(d0 / 180) * M_PI);
    theta[2 * (size_t)D + row] = (double)(float)((double)(d1 / 180) * M_PI);
  }
  // one extra matrix slot (index 3*D) for nis_debug_rotate
  mats.resize(mats.size() + 6, 0.0);
  // per (matrix slot, output row): X0, Y0 of warpAffine's fixed-point walk (AB_BITS = 10, round_delta = 16), the same two roundings
  // as OpenCV's per-row set-up; the fused rotation prologue reads them instead of redoing the double arithmetic per column pair
  std::vector<int2> rowtab((size_t)(3 * D + 1) * H);
  for (size_t slot = 0; slot < (size_t)3 * D; ++slot) rotate_row_table(&mats[slot * 6], H, &rowtab[slot * H]);
  RESERVE(ctx->rowtab, rowtab.size() * sizeof(int2));
  CU(h2d(ctx, ctx->rowtab.p, rowtab.data(), rowtab.size() * sizeof(int2)));
  RESERVE(ctx->mats, mats.size() * sizeof(double));
  RESERVE(ctx->theta, theta.size() * sizeof(double));
  CU(h2d(ctx, ctx->mats.p, mats.data(), mats.size() * sizeof(double)));
  CU(h2d(ctx, ctx->theta.p, theta.data(), theta.size() * sizeof(double)));
  return NIS_OK;
}

// TMA descriptor (cuTensorMapEncodeTiled, resolved through the runtime's driver entry point: no link against libcuda) over the lane's
// shifted power images: rank 3 {W, H, images} f32, box {pitch, rows, 1}, no swizzle, out-of-range elements read as zero
static int encode_hp_map(nis_ctx* ctx, Lane& L, int pairs) {
  static PFN_cuTensorMapEncodeTiled encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
      return fail(ctx, NIS_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    encode = (PFN_cuTensorMapEncodeTiled)fn;
  }
  const cuuint64_t dims[3] = {(cuuint64_t)ctx->W, (cuuint64_t)ctx->H, (cuuint64_t)pairs};
  const cuuint64_t strides[2] = {(cuuint64_t)ctx->W * sizeof(float), (cuuint64_t)ctx->W * ctx->H * sizeof(float)};
  const cuuint32_t box[3] = {(cuuint32_t)ctx->ptile_pitch, (cuuint32_t)ctx->ptile_rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = encode(&L.hp_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, L.real.p, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ctx, NIS_ERR_CUDA, "cuTensorMapEncodeTiled failed");
  return NIS_OK;
}

static int ensure_workspace(nis_ctx* ctx, Lane& L, int pairs) {
  if (pairs <= L.cap) return NIS_OK;
  CU(cudaStreamSynchronize(L.stream));
  L.cap = 0;                                   // stays 0 if any of the reservations below fails (no stale capacity after an OOM)
  const size_t E = 2 * (size_t)pairs;
  RESERVE(L.t1, E * ctx->maxspec * sizeof(cpx));          // the only full-size scratch: half-transformed spectra, in place
  RESERVE(L.real, (size_t)pairs * ctx->maxreal * sizeof(float));   // power = IFFT(|F|), stored fftshift-ed: what the polar gather's TMA tiles read
  RESERVE(L.pol, (size_t)pairs * ctx->sz[1].real * sizeof(float)); // polar image
  TRY(encode_hp_map(ctx, L, pairs));
  RESERVE(L.maxp, (size_t)pairs * sizeof(unsigned));
  RESERVE(L.maxt, E * sizeof(unsigned));
  RESERVE(L.maxh, (size_t)pairs * sizeof(unsigned));
  RESERVE(L.stats_p, (size_t)pairs * sizeof(PeakStats));
  RESERVE(L.stats_t, E * sizeof(PeakStats));
  RESERVE(L.sel, E * sizeof(int));
  RESERVE(L.xx, E * sizeof(double));
  RESERVE(L.zz, (size_t)pairs * sizeof(double));
  L.cap = pairs;
  return NIS_OK;
}

// lanes 1.. start after everything already queued on the main stream; the main stream resumes after all lanes
static int fork_lanes(nis_ctx* ctx) {
  if (ctx->active_lanes <= 1) return NIS_OK;
  CU(cudaEventRecord(ctx->fork_ev, ctx->stream));
  for (int i = 1; i < ctx->active_lanes; ++i) CU(cudaStreamWaitEvent(ctx->lanes[i].stream, ctx->fork_ev, 0));
  return NIS_OK;
}
static int join_lanes(nis_ctx* ctx) {
  for (int i = 1; i < ctx->active_lanes; ++i) {
    CU(cudaEventRecord(ctx->lanes[i].ev, ctx->lanes[i].stream));
    CU(cudaStreamWaitEvent(ctx->stream, ctx->lanes[i].ev, 0));
  }
  return NIS_OK;
}

static int ensure_recs(nis_ctx* ctx, int n) {
  if (n <= ctx->recs_cap) return NIS_OK;
  CU(cudaDeviceSynchronize());
  ctx->recs_cap = 0;
  RESERVE(ctx->recs, (size_t)n * sizeof(PoseRecord));
  RESERVE(ctx->best, sizeof(PoseRecord));
  ctx->recs_cap = n;
  return NIS_OK;
}

static int ensure_pinned(nis_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->pin_bytes) return NIS_OK;
  if (ctx->pin) cudaFreeHost(ctx->pin);
  ctx->pin = nullptr; ctx->pin_bytes = 0;
  CU(cudaMallocHost(&ctx->pin, bytes));
  ctx->pin_bytes = bytes;
  return NIS_OK;
}

// -------------------------------------------------------------------------------------------------------------
// batched stages
// -------------------------------------------------------------------------------------------------------------
// KernelFn for size class s (polynomial / gaussian, correlation_flow.cc:181-226)
static KernelFn kernel_fn(const nis_ctx* ctx, int s, const double* xx, const double* zz, int zz_shift, unsigned* maxbuf,
                          const int* xx_idx = nullptr) {
  const nis_cf_config& c = ctx->cfg;
  return KernelFn{(float)(unsigned)ctx->sz[s].real, c.kernel, c.offset, c.power, -1.f / (c.sigma * c.sigma), xx, zz, zz_shift, maxbuf, xx_idx};
}

// H = T / (kernel(Z)/max + lambda) for B keyframe spectra of size class s (the keyframe-only factor of EstimateTrans,
// correlation_flow.cc:157-171: Kzz = kernel(z); H = output_fft/(Kzz + lambda)).  3 kernels, nothing real-valued is stored.
// hzz_tail: the last two of them, for a caller whose own kernel has already left the row half of IFFT(Z conj Z) in the lane scratch.
static int hzz_tail(nis_ctx* ctx, Lane& L, int s, int B, Dst<cpx> Hout) {
  const SizeClass& z = ctx->sz[s];
  Dst<cpx> t1{L.t1.as<cpx>(), (long long)z.spec};
  Src<cpx> t1s = src_slab<cpx>(t1.base, t1.stride);
  LAUNCH(launch_colcol(z.R, z.col, t1s, t1, kernel_fn(ctx, s, L.zz.as<double>(), L.zz.as<double>(), 0, L.maxh.as<unsigned>()),
                       z.C, B, L.stream));
  LAUNCH(launch_row_fwd_h(z.C, z.row, ProSpec{t1s}, EpiHStore{Hout, L.maxh.as<unsigned>(), ctx->cfg.lambda}, z.R / 2 + 1, B, L.stream));
  return NIS_OK;
}
static int hzz_batch(nis_ctx* ctx, Lane& L, int s, Src<cpx> Z, int B, Dst<cpx> Hout) {
  ctx->prof_stream = L.stream;
  const SizeClass& z = ctx->sz[s];
  if (ctx->cfg.kernel != 0 && ctx->cfg.kernel != 1) return fail(ctx, NIS_ERR_INVALID_KERNEL, "Received invalid kernel type");
  Dst<cpx> t1{L.t1.as<cpx>(), (long long)z.spec};
  if (ctx->cfg.kernel == 1) LAUNCH(launch_spec_sqsum(Z, (int)z.spec, L.zz.as<double>(), B, L.stream));
  CU(cudaMemsetAsync(L.maxh.p, 0, sizeof(unsigned) * B, L.stream));
  // polar size: features_batch computes this factor through the fused store-and-square row kernel; stored spectra (compact store
  // modes, imported frames) take the same plan here so that both routes give the same bits
  LAUNCH(launch_row_inv_mulconj(z.C, z.row, ProMulConj{Z, Z}, EpiSpecStore{t1}, z.R / 2 + 1, B, L.stream, /*match_fused=*/s == 1));
  return hzz_tail(ctx, L, s, B, Hout);
}

#ifndef NIS_FUSE_PSQ
#define NIS_FUSE_PSQ 1      // A/B switch: 0 = fft_polar by a plain row pass, its keyframe factor by the three-kernel chain
#endif
// ComputeIntermedium (correlation_flow.cc:89-95) for B images, plus the cached H factors of both stages.
// with_h = false skips the H factors (frames that are only ever "current").
static int features_batch(nis_ctx* ctx, Lane& L, Src<float> f32, Src<uint8_t> u8, bool is_u8, int B, Dst<cpx> F, Dst<cpx> P,
                          Dst<cpx> Ht, Dst<cpx> Hp, bool with_h) {
  ctx->prof_stream = L.stream;
  TRY(ensure_workspace(ctx, L, B));
  const SizeClass& zt = ctx->sz[0];
  const SizeClass& zp = ctx->sz[1];
  Dst<cpx> t1{L.t1.as<cpx>(), (long long)zt.spec};
  Src<cpx> t1s = src_slab<cpx>(t1.base, t1.stride);
  if (is_u8) LAUNCH(launch_col_fwd_u8(zt.R, zt.col, ProRealU8{u8, zt.C, ctx->lut.as<float>()}, t1, zt.C, B, L.stream));
  else LAUNCH(launch_col_fwd_f32(zt.R, zt.col, ProRealF32{f32, zt.C}, t1, zt.C, B, L.stream));
  // fft_result = FFT(image) stored from registers; the same kernel continues with IFFT(|fft_result|) along the rows
  LAUNCH(launch_rowrow_storeabs(zt.C, zt.row, t1s, t1, MidStoreAbs{F}, zt.R / 2 + 1, B, L.stream));
  // power = IFFT(|fft_result|) lands fftshift-ed; RemoveZeroComponent in place; polar cells gather from TMA-staged tiles; the polar
  // image (it stays in L2) then takes a plain r2c column pass   (correlation_flow.cc:92-94)
  Dst<float> power{L.real.as<float>(), (long long)zt.real};
  LAUNCH(launch_col_inv_store_shift(zt.R, zt.col, t1s, EpiStoreShift{power, zt.R, zt.C, (float)zt.real}, zt.C, B, L.stream));
  LAUNCH(launch_rzc_fix(power, ctx->H, ctx->W, B, L.stream));
  Dst<cpx> t1p{L.t1.as<cpx>(), (long long)zp.spec};
  Dst<float> pol{L.pol.as<float>(), (long long)zp.real};
  LAUNCH(launch_polar_tma(&L.hp_map, pol, ctx->D, ctx->Cp, ctx->ptiles.as<int4>(), ctx->ptab2.as<uint32_t>(), ctx->ptile_pitch, ctx->ptile_rows, B,
                          L.stream));
  LAUNCH(launch_col_fwd_f32(zp.R, zp.col, ProRealF32{src_slab<float>(pol.base, pol.stride), zp.C}, t1p, zp.C, B, L.stream));
  if (NIS_FUSE_PSQ && with_h && (ctx->cfg.kernel == 0 || ctx->cfg.kernel == 1)) {   // an invalid kernel id only throws in ComputePose (:168)
    // fft_polar is stored from registers and the same kernel continues with the row half of IFFT(P conj P): the polar-stage keyframe
    // factor starts without re-reading P
    LAUNCH(launch_rowrow_storesq(zp.C, zp.row, src_slab<cpx>(t1p.base, t1p.stride), t1p, MidStoreSq{P}, zp.R / 2 + 1, B, L.stream));
    if (ctx->cfg.kernel == 1) LAUNCH(launch_spec_sqsum(src_slab<cpx>(P.base, P.stride), (int)zp.spec, L.zz.as<double>(), B, L.stream));
    CU(cudaMemsetAsync(L.maxh.p, 0, sizeof(unsigned) * B, L.stream));
    TRY(hzz_tail(ctx, L, 1, B, Hp));
    TRY(hzz_batch(ctx, L, 0, src_slab<cpx>(F.base, F.stride), B, Ht));
  } else {
    LAUNCH(launch_row_fwd(zp.C, zp.row, ProSpec{src_slab<cpx>(t1p.base, t1p.stride)}, EpiSpecStore{P}, zp.R / 2 + 1, B, L.stream));
    if (with_h && (ctx->cfg.kernel == 0 || ctx->cfg.kernel == 1)) {
      TRY(hzz_batch(ctx, L, 0, src_slab<cpx>(F.base, F.stride), B, Ht));
      TRY(hzz_batch(ctx, L, 1, src_slab<cpx>(P.base, P.stride), B, Hp));
    }
  }
  return NIS_OK;
}

// the current-frame half of EstimateTrans (correlation_flow.cc:160, :171-178) on half-transformed input already in t1:
//   [t1 holds row-inverse of X*conj(Z)] -> colcol (IFFT cols, kernel fn, FFT cols) -> rowrow (FFT rows, *H/max, IFFT rows)
//   -> col_inv_peak (IFFT cols, arg-max + sums)
static int correlate_tail(nis_ctx* ctx, Lane& L, int s, Src<cpx> Hz, int E, int zshift, unsigned* maxbuf, PeakStats* stats, float* g_debug,
                          const double* xx_override = nullptr, const int* xx_idx = nullptr) {
  const SizeClass& z = ctx->sz[s];
  Dst<cpx> t1{L.t1.as<cpx>(), (long long)z.spec};
  Src<cpx> t1s = src_slab<cpx>(t1.base, t1.stride);
  LAUNCH(launch_colcol(z.R, z.col, t1s, t1,
                       kernel_fn(ctx, s, xx_override ? xx_override : L.xx.as<double>(), L.zz.as<double>(), zshift, maxbuf, xx_idx), z.C, E,
                       L.stream));
  Src<cpx> Hze = Hz; Hze.shift = zshift;
  LAUNCH(launch_rowrow_filter(z.C, z.row, t1s, t1, MidFilterH{Hze, maxbuf}, z.R / 2 + 1, E, L.stream));
  EpiPeak ep{stats, z.R, (float)z.real, g_debug, (long long)z.real, z.C};
  LAUNCH(launch_col_inv_peak(z.R, z.col, t1s, ep, z.C, E, L.stream));
  return NIS_OK;
}

// EstimateTrans with both spectra stored (the polar stage): E == number of keyframes, no hypothesis sharing
static int estimate_trans_stored(nis_ctx* ctx, Lane& L, int s, Src<cpx> Z, Src<cpx> Hz, Src<cpx> X, int B, unsigned* maxbuf,
                                 PeakStats* stats, float* g_debug) {
  ctx->prof_stream = L.stream;
  const SizeClass& z = ctx->sz[s];
  if (ctx->cfg.kernel != 0 && ctx->cfg.kernel != 1) return fail(ctx, NIS_ERR_INVALID_KERNEL, "Received invalid kernel type");
  if (ctx->cfg.kernel == 1) {
    LAUNCH(launch_spec_sqsum(X, (int)z.spec, L.xx.as<double>(), B, L.stream));
    LAUNCH(launch_spec_sqsum(Z, (int)z.spec, L.zz.as<double>(), B, L.stream));
  }
  CU(cudaMemsetAsync(maxbuf, 0, sizeof(unsigned) * B, L.stream));
  CU(cudaMemsetAsync(stats, 0, sizeof(PeakStats) * B, L.stream));
  Dst<cpx> t1{L.t1.as<cpx>(), (long long)z.spec};
  LAUNCH(launch_row_inv_mulconj(z.C, z.row, ProMulConj{X, Z}, EpiSpecStore{t1}, z.R / 2 + 1, B, L.stream));
  return correlate_tail(ctx, L, s, Hz, B, 0, maxbuf, stats, g_debug);
}

// scan only: FFT(RotateArray(query, -deg)) and FFT(RotateArray(query, -deg+180)) for every polar row (2 D spectra), so that a
// candidate's two hypotheses cost one spectrum read each.  Same kernels, same arithmetic as the per-candidate path.
static int build_rot_cache(nis_ctx* ctx, Src<float> img_f32, Src<uint8_t> img_u8, bool is_u8) {
  const SizeClass& zt = ctx->sz[0];
  const int D = ctx->D, n = 2 * D;
  RESERVE(ctx->rotc, (size_t)n * zt.spec * sizeof(cpx));
  RESERVE(ctx->rotc_xx, (size_t)n * sizeof(double));
  if (ctx->rotc_sel.bytes < (size_t)n * sizeof(int)) {
    RESERVE(ctx->rotc_sel, (size_t)n * sizeof(int));
    std::vector<int> sel(n);
    for (int i = 0; i < n; ++i) sel[i] = D + i;          // matrix slots D..3D-1 = the two loop variants
    CU(h2d(ctx, ctx->rotc_sel.p, sel.data(), (size_t)n * sizeof(int)));
  }
  const int NL = ctx->active_lanes, B = std::max(1, ctx->batch) * 2;
  const bool gauss = ctx->cfg.kernel == 1;
  TRY(fork_lanes(ctx));
  for (int i0 = 0, k = 0; i0 < n; i0 += B, ++k) {
    Lane& L = ctx->lanes[k % NL];
    ctx->prof_stream = L.stream;
    const int nb = std::min(B, n - i0);
    TRY(ensure_workspace(ctx, L, (nb + 1) / 2));
    Dst<cpx> t1{L.t1.as<cpx>(), (long long)zt.spec};
    Src<float> i32 = img_f32; Src<uint8_t> i8 = img_u8;
    if (is_u8) i32 = src_null<float>(); else i8 = src_null<uint8_t>();
    RotateArgs ra{i32, i8, is_u8, ctx->lut.as<float>(), ctx->H, ctx->W, ctx->mats.as<double>(), ctx->rotc_sel.as<int>() + i0, ctx->rowtab.as<int2>()};
    LAUNCH(launch_col_fwd_rotate(zt.R, zt.col, ra, t1, zt.C, nb, L.stream));
    Dst<cpx> out{ctx->rotc.as<cpx>() + (size_t)i0 * zt.spec, (long long)zt.spec};
    LAUNCH(launch_row_fwd(zt.C, zt.row, ProSpec{src_slab<cpx>(t1.base, t1.stride)}, EpiSpecStore{out}, zt.R / 2 + 1, nb, L.stream, /*match_fused=*/true));
    if (gauss) LAUNCH(launch_spec_sqsum(src_slab<cpx>(out.base, out.stride), (int)zt.spec, ctx->rotc_xx.as<double>() + i0, nb, L.stream));
  }
  TRY(join_lanes(ctx));
  return NIS_OK;
}

// ComputePose (correlation_flow.cc:97-138) for B pairs -> device records
static int compute_pose_batch(nis_ctx* ctx, Lane& L, bool loop_mode, Src<cpx> Fz, Src<cpx> Pz, Src<cpx> Htz, Src<cpx> Hpz, Src<cpx> Px,
                              Src<float> img_f32, Src<uint8_t> img_u8, bool is_u8, int B, int index0, PoseRecord* recs) {
  ctx->prof_stream = L.stream;
  TRY(ensure_workspace(ctx, L, B));
  PeakStats* sp = L.stats_p.as<PeakStats>();
  PeakStats* st = L.stats_t.as<PeakStats>();
  // rotation: EstimateTrans(last_fft_polar, fft_polar, ...)   (:103)
  TRY(estimate_trans_stored(ctx, L, 1, Pz, Hpz, Px, B, L.maxp.as<unsigned>(), sp, nullptr));
  // translation: FFT(RotateArray(image, -deg [+180])) never leaves the chip; one or two hypotheses per pair   (:107-132)
  const SizeClass& zt = ctx->sz[0];
  const int shift = loop_mode ? 1 : 0, E = B << shift;
  Src<float> i32 = img_f32; i32.shift = shift;
  Src<uint8_t> i8 = img_u8; i8.shift = shift;
  if (is_u8) i32 = src_null<float>(); else i8 = src_null<uint8_t>();
  Dst<cpx> t1{L.t1.as<cpx>(), (long long)zt.spec};
  Src<cpx> t1s = src_slab<cpx>(t1.base, t1.stride);
  const bool gauss = ctx->cfg.kernel == 1;
  if (gauss) {
    CU(cudaMemsetAsync(L.xx.p, 0, sizeof(double) * E, L.stream));
    LAUNCH(launch_spec_sqsum(Fz, (int)zt.spec, L.zz.as<double>(), B, L.stream));
  }
  CU(cudaMemsetAsync(L.maxt.p, 0, sizeof(unsigned) * E, L.stream));
  CU(cudaMemsetAsync(st, 0, sizeof(PeakStats) * E, L.stream));
  Src<cpx> Fze = Fz; Fze.shift = shift;
  if (loop_mode && ctx->use_rot_cache) {
    // FFT(RotateArray(query, angle)) comes from the per-query cache: one inverse row pass instead of warp + 3 passes; the cache is
    // indexed through sel[] (two hypotheses per candidate), written by the select kernel
    LAUNCH(launch_polar_select(sp, ctx->D, 1, L.sel.as<int>(), B, L.stream));
    Src<cpx> Xc{ctx->rotc.as<cpx>() - (size_t)ctx->D * zt.spec, (long long)zt.spec, nullptr, 0, L.sel.as<int>(), 0};
    LAUNCH(launch_row_inv_mulconj(zt.C, zt.row, ProMulConj{Xc, Fze}, EpiSpecStore{t1}, zt.R / 2 + 1, E, L.stream, /*match_fused=*/true));
    TRY(correlate_tail(ctx, L, 0, Htz, E, shift, L.maxt.as<unsigned>(), st, nullptr, ctx->rotc_xx.as<double>() - ctx->D, L.sel.as<int>()));
  } else {
    // the rotation prologue reads the polar-stage peak itself: no select launch between the two stages
    RotateArgs ra{i32, i8, is_u8, ctx->lut.as<float>(), ctx->H, ctx->W, ctx->mats.as<double>(), nullptr, ctx->rowtab.as<int2>(), sp, ctx->D, shift};
    LAUNCH(launch_col_fwd_rotate(zt.R, zt.col, ra, t1, zt.C, E, L.stream));
    LAUNCH(launch_rowrow_mulconj(zt.C, zt.row, t1s, t1, MidMulConjZ{Fze, gauss ? L.xx.as<double>() : nullptr}, zt.R / 2 + 1, E, L.stream));
    TRY(correlate_tail(ctx, L, 0, Htz, E, shift, L.maxt.as<unsigned>(), st, nullptr));
  }
  AngleTables tabs{ctx->mats.as<double>(), ctx->theta.as<double>()};
  LAUNCH(launch_pose_finalize(sp, st, tabs, ctx->H, ctx->W, ctx->D, ctx->Cp, loop_mode ? 1 : 0, index0, recs, B, L.stream));
  return NIS_OK;
}

// -------------------------------------------------------------------------------------------------------------
// layout conversion at the boundary (reference column-major <-> internal row-major); boundary utilities only
// -------------------------------------------------------------------------------------------------------------
template <class T> static void transpose_to(const T* in, int rows_in, int cols_in, T* out) {   // in[rows_in][cols_in] -> out[cols_in][rows_in]
  for (int r = 0; r < rows_in; ++r)
    for (int c = 0; c < cols_in; ++c) out[(size_t)c * rows_in + r] = in[(size_t)r * cols_in + c];
}

// Map::ComputeGridLocation (src/map.cc:81-85): static_cast<int>(x / grid_scale), truncation toward zero
static std::pair<int, int> grid_cell(double x, double y, double scale) { return {(int)(x / scale), (int)(y / scale)}; }

// -------------------------------------------------------------------------------------------------------------
// C ABI
// -------------------------------------------------------------------------------------------------------------
extern "C" {

const char* nis_strerror(int status) {
  switch (status) {
    case NIS_OK: return "ok";
    case NIS_ERR_INVALID_ARGUMENT: return "invalid argument";
    case NIS_ERR_INVALID_KERNEL: return "Received invalid kernel type";
    case NIS_ERR_UNSUPPORTED_SIZE: return "unsupported transform size";
    case NIS_ERR_CUDA: return "CUDA error";
    case NIS_ERR_OUT_OF_MEMORY: return "out of device memory";
  }
  return "unknown status";
}

const char* nis_last_error(const nis_ctx* ctx) { return ctx ? ctx->err.c_str() : ""; }

int nis_create(const nis_cf_config* cfg, int image_height, int image_width, int device, nis_ctx** out) {
  if (!cfg || !out) return NIS_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  const int H = image_height, W = image_width, D = cfg->rotation_divisor, Cp = cfg->rotation_channel;
  if (H <= 0 || W <= 0 || D <= 0 || Cp <= 0 || (H & 1) || (D & 1) || (W % 32) || (Cp % 16)) return NIS_ERR_INVALID_ARGUMENT;
  if (!col_size_supported(H) || !col_size_supported(D) || !row_size_supported(W) || !row_size_supported(Cp))
    return NIS_ERR_UNSUPPORTED_SIZE;
  // the fused rotation wraps source coordinates with one conditional add/subtract (exact while the half diagonal stays
  // below 1.5 x the shorter side, i.e. aspect ratio <= 2.8)
  if ((double)std::max(H, W) > 2.8 * (double)std::min(H, W)) return NIS_ERR_UNSUPPORTED_SIZE;
  // the kernel id is only checked when EstimateTrans runs, like the reference (correlation_flow.cc:157-169)
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return NIS_ERR_CUDA;
  if (cudaSetDevice(device) != cudaSuccess) return NIS_ERR_CUDA;
  nis_ctx* ctx = new nis_ctx();
  ctx->device = device; ctx->cfg = *cfg;
  ctx->H = H; ctx->W = W; ctx->D = D; ctx->Cp = Cp;
  ctx->sz[0].R = H; ctx->sz[0].C = W; ctx->sz[1].R = D; ctx->sz[1].C = Cp;
  for (int s = 0; s < 2; ++s) {
    ctx->sz[s].spec = (size_t)(ctx->sz[s].R / 2 + 1) * ctx->sz[s].C;
    ctx->sz[s].real = (size_t)ctx->sz[s].R * ctx->sz[s].C;
  }
  ctx->maxspec = std::max(ctx->sz[0].spec, ctx->sz[1].spec);
  ctx->maxreal = std::max(ctx->sz[0].real, ctx->sz[1].real);
  const int nlanes = 8;
  ctx->active_lanes = 3;
  const char* el = getenv("NIS_LANES");
  if (el && atoi(el) > 0) ctx->active_lanes = std::min(atoi(el), nlanes);
  ctx->lanes.resize(nlanes);
  for (Lane& L : ctx->lanes)
    if (cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&L.ev, cudaEventDisableTiming) != cudaSuccess) { nis_destroy(ctx); return NIS_ERR_CUDA; }
  ctx->stream = ctx->lanes[0].stream;
  if (cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming) != cudaSuccess) { nis_destroy(ctx); return NIS_ERR_CUDA; }
  if (cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { nis_destroy(ctx); return NIS_ERR_CUDA; }
  {
    // default batch: one column-pass launch (W/32 CTAs per image, 2 resident CTAs per SM) should fill the GPU just once
    cudaDeviceProp prop;
    int sms = 148;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) sms = prop.multiProcessorCount;
    // four waves of column CTAs per launch on three lanes: measured best both HBM-resident and end to end (sweep in DESIGN.md)
    ctx->default_batch = std::max(4, std::min(64, 4 * ((2 * sms) / std::max(1, W / 32))));
    ctx->batch = ctx->default_batch;
  }
  const char* er = getenv("NIS_ROT_CACHE_MIN");      // candidates from which a scan builds the rotated-query cache (0 = never)
  if (er) ctx->rot_cache_min = atoi(er);
  const char* eb = getenv("NIS_BATCH");
  if (eb && atoi(eb) > 0) ctx->batch = atoi(eb);
  int st = build_tables(ctx);
  if (st != NIS_OK) { nis_destroy(ctx); return st; }
  *out = ctx;
  return NIS_OK;
}

int nis_destroy(nis_ctx* ctx) {
  if (!ctx) return NIS_OK;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  DevBuf* bufs[] = {&ctx->tw, &ctx->lut, &ctx->cs, &ctx->rho, &ctx->mats, &ctx->theta, &ctx->ptiles, &ctx->ptab2, &ctx->rowtab, &ctx->recs, &ctx->best, &ctx->cand, &ctx->sF,
                    &ctx->sP, &ctx->sHt, &ctx->sHp, &ctx->sImg, &ctx->sUnd, &ctx->umap1, &ctx->umap2, &ctx->d_slot_ptr, &ctx->rotc, &ctx->rotc_xx, &ctx->rotc_sel,
                    &ctx->d_fid, &ctx->d_dist, &ctx->d_cell, &ctx->cand_in, &ctx->cand_pos, &ctx->sel_scratch, &ctx->stage, &ctx->kfrec, &ctx->qgather};
  for (DevBuf* b : bufs) b->release();
  for (Lane& L : ctx->lanes) {
    DevBuf* lb[] = {&L.t1, &L.real, &L.pol, &L.maxp, &L.maxt, &L.maxh, &L.stats_p, &L.stats_t, &L.sel, &L.xx, &L.zz, &L.dbrec};
    for (DevBuf* b : lb) b->release();
    if (L.ev) cudaEventDestroy(L.ev);
    if (L.stream) cudaStreamDestroy(L.stream);
  }
  if (ctx->fork_ev) cudaEventDestroy(ctx->fork_ev);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  for (nis_frame* f : ctx->live_frames) { if (f->block) cudaFree(f->block); f->block = nullptr; }   // the handles themselves stay valid to free
  ctx->live_frames.clear();
  nis_comm_destroy(ctx);
  for (void* b : ctx->frame_pool) cudaFree(b);
  for (cudaEvent_t e : ctx->up_ev) cudaEventDestroy(e);
  for (cudaEvent_t e : ctx->feat_ev) cudaEventDestroy(e);
  for (void* c : ctx->chunks) cudaFree(c);
  if (ctx->pin) cudaFreeHost(ctx->pin);
  delete ctx;
  return NIS_OK;
}

void* nis_stream(nis_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int nis_synchronize(nis_ctx* ctx) {
  if (!ctx) return NIS_ERR_INVALID_ARGUMENT;
  CU(cudaStreamSynchronize(ctx->stream));
  return NIS_OK;
}
long long nis_kernel_launches(const nis_ctx* ctx) { return ctx ? ctx->launches : 0; }
int nis_set_lanes(nis_ctx* ctx, int lanes) {
  if (!ctx || lanes < 0) return NIS_ERR_INVALID_ARGUMENT;
  CU(cudaSetDevice(ctx->device));
  CU(cudaDeviceSynchronize());
  ctx->active_lanes = lanes > 0 ? std::min(lanes, (int)ctx->lanes.size()) : 3;
  return NIS_OK;
}
int nis_set_batch(nis_ctx* ctx, int batch) {
  if (!ctx || batch < 0) return NIS_ERR_INVALID_ARGUMENT;
  ctx->batch = batch > 0 ? batch : ctx->default_batch;
  return NIS_OK;
}

// ---- frames -------------------------------------------------------------------------------------------------
static int frame_alloc(nis_ctx* ctx, bool u8, nis_frame** out) {
  nis_frame* f = new nis_frame();
  const size_t bF = ctx->sz[0].spec * sizeof(cpx), bP = ctx->sz[1].spec * sizeof(cpx);
  const size_t bI = ctx->sz[0].real * (u8 ? 1 : sizeof(float));
  // every frame block has room for an f32 image, so freed blocks are interchangeable and come back from a small pool: cudaMalloc /
  // cudaFree per frame cost up to hundreds of ms next to a large keyframe store (measured), the per-frame MapBuilder loop hits both
  if (!ctx->frame_pool.empty()) {
    f->block = ctx->frame_pool.back();
    ctx->frame_pool.pop_back();
  } else {
    cudaError_t e = cudaMalloc(&f->block, 2 * (bF + bP) + ctx->sz[0].real * sizeof(float));
    if (e != cudaSuccess) { delete f; return fail(ctx, NIS_ERR_OUT_OF_MEMORY, "cudaMalloc frame", (int)e); }
  }
  char* p = (char*)f->block;
  f->F = (cpx*)p; f->P = (cpx*)(p + bF); f->Ht = (cpx*)(p + bF + bP); f->Hp = (cpx*)(p + 2 * bF + bP);
  if (u8) f->img_u8 = (uint8_t*)(p + 2 * (bF + bP)); else f->img_f32 = (float*)(p + 2 * (bF + bP));
  ctx->live_frames.push_back(f);
  *out = f;
  return NIS_OK;
}

int nis_frame_free(nis_ctx* ctx, nis_frame* f) {
  if (!f) return NIS_OK;
  if (ctx) {
    cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream);
    auto it = std::find(ctx->live_frames.begin(), ctx->live_frames.end(), f);
    if (it != ctx->live_frames.end()) { *it = ctx->live_frames.back(); ctx->live_frames.pop_back(); }
  }
  if (f->block) {
    if (ctx && ctx->frame_pool.size() < 16) ctx->frame_pool.push_back(f->block);
    else cudaFree(f->block);
  }
  delete f;
  return NIS_OK;
}

int nis_set_undistort_maps(nis_ctx* ctx, const int16_t* map1_xy, const uint16_t* map2) {
  if (!ctx || ((map1_xy == nullptr) != (map2 == nullptr))) return NIS_ERR_INVALID_ARGUMENT;
  CU(cudaSetDevice(ctx->device));
  CU(cudaDeviceSynchronize());
  if (!map1_xy) { ctx->undistort = false; return NIS_OK; }
  const size_t npx = ctx->sz[0].real;
  RESERVE(ctx->umap1, npx * 2 * sizeof(int16_t));
  RESERVE(ctx->umap2, npx * sizeof(uint16_t));
  CU(h2d(ctx, ctx->umap1.p, map1_xy, npx * 2 * sizeof(int16_t)));
  CU(h2d(ctx, ctx->umap2.p, map2, npx * sizeof(uint16_t)));
  ctx->undistort = true;
  return NIS_OK;
}

// undistort B raw u8 images on stream `st` (batched Camera::UndistortImage)
static int undistort_batch(nis_ctx* ctx, cudaStream_t st, const uint8_t* raw, uint8_t* out, int B) {
  ctx->prof_stream = st;
  const long long npx = (long long)ctx->sz[0].real;
  LAUNCH(launch_undistort(src_slab<uint8_t>(raw, npx), Dst<uint8_t>{out, npx}, ctx->H, ctx->W, ctx->umap1.p, ctx->umap2.p, B, st));
  return NIS_OK;
}

int nis_undistort_u8(nis_ctx* ctx, const uint8_t* raw, uint8_t* out) {
  if (!ctx || !raw || !out) return NIS_ERR_INVALID_ARGUMENT;
  if (!ctx->undistort) return fail(ctx, NIS_ERR_INVALID_ARGUMENT, "no undistort maps set");
  CU(cudaSetDevice(ctx->device));
  const size_t npx = ctx->sz[0].real;
  RESERVE(ctx->sImg, 2 * npx);
  uint8_t* d = ctx->sImg.as<uint8_t>();
  CU(h2d(ctx, d, raw, npx));
  TRY(undistort_batch(ctx, ctx->stream, d, d + npx, 1));
  CU(cudaStreamSynchronize(ctx->stream));
  CU(cudaMemcpy(out, d + npx, npx, cudaMemcpyDeviceToHost));
  return NIS_OK;
}

int nis_features_u8(nis_ctx* ctx, const uint8_t* image, nis_frame** out) {
  if (!ctx || !image || !out) return NIS_ERR_INVALID_ARGUMENT;
  CU(cudaSetDevice(ctx->device));
  nis_frame* f = nullptr;
  TRY(frame_alloc(ctx, true, &f));
  cudaError_t e;
  if (ctx->undistort) {                       // raw camera image: undistort into the frame's image buffer first
    const size_t npx = ctx->sz[0].real;
    if (ctx->sImg.reserve(npx)) { nis_frame_free(ctx, f); return fail(ctx, NIS_ERR_OUT_OF_MEMORY, "cudaMalloc raw staging"); }
    e = cudaMemcpyAsync(ctx->sImg.p, image, npx, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && undistort_batch(ctx, ctx->stream, ctx->sImg.as<uint8_t>(), f->img_u8, 1) != NIS_OK) e = cudaErrorUnknown;
  } else {
    e = cudaMemcpyAsync(f->img_u8, image, ctx->sz[0].real, cudaMemcpyHostToDevice, ctx->stream);
  }
  int st = e == cudaSuccess ? features_batch(ctx, ctx->lanes[0], src_null<float>(), src_slab<uint8_t>(f->img_u8, 0), true, 1,
                                              Dst<cpx>{f->F, 0}, Dst<cpx>{f->P, 0}, Dst<cpx>{f->Ht, 0}, Dst<cpx>{f->Hp, 0}, true)
                            : fail(ctx, NIS_ERR_CUDA, "cudaMemcpyAsync image", (int)e);
  if (st == NIS_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) st = fail(ctx, NIS_ERR_CUDA, "features sync", (int)cudaGetLastError());
  if (st != NIS_OK) { nis_frame_free(ctx, f); return st; }
  *out = f;
  return NIS_OK;
}

static int upload_colmajor_f32(nis_ctx* ctx, const float* host, int R, int C, float* dst);

int nis_features_f32(nis_ctx* ctx, const float* image_colmajor, nis_frame** out) {
  if (!ctx || !image_colmajor || !out) return NIS_ERR_INVALID_ARGUMENT;
  CU(cudaSetDevice(ctx->device));
  nis_frame* f = nullptr;
  TRY(frame_alloc(ctx, false, &f));
  int st = upload_colmajor_f32(ctx, image_colmajor, ctx->H, ctx->W, f->img_f32);          // [W][H] lines -> [H][W] on the device
  if (st == NIS_OK)
    st = features_batch(ctx, ctx->lanes[0], src_slab<float>(f->img_f32, 0), src_null<uint8_t>(), false, 1, Dst<cpx>{f->F, 0}, Dst<cpx>{f->P, 0},
                        Dst<cpx>{f->Ht, 0}, Dst<cpx>{f->Hp, 0}, true);
  if (st == NIS_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) st = fail(ctx, NIS_ERR_CUDA, "features sync", (int)cudaGetLastError());
  if (st != NIS_OK) { nis_frame_free(ctx, f); return st; }
  *out = f;
  return NIS_OK;
}

// Frame::GetFFTResult: row-major device spectra -> reference layout, transposed on the device into a staging buffer, one D2H each
int nis_frame_export(nis_ctx* ctx, const nis_frame* f, float* fft_result, float* fft_polar) {
  if (!ctx || !f) return NIS_ERR_INVALID_ARGUMENT;
  CU(cudaSetDevice(ctx->device));
  for (int s = 0; s < 2; ++s) {
    float* dst = s == 0 ? fft_result : fft_polar;
    if (!dst) continue;
    if (s == 0 ? !f->has_spectra : !f->has_polar) return fail(ctx, NIS_ERR_INVALID_ARGUMENT, "frame export: the frame does not hold that array");
    const SizeClass& z = ctx->sz[s];
    RESERVE(ctx->stage, z.spec * sizeof(cpx));
    ctx->prof_stream = ctx->stream;
    LAUNCH(launch_transpose_cpx(s == 0 ? f->F : f->P, ctx->stage.as<cpx>(), z.R / 2 + 1, z.C, ctx->stream));       // [half][C] -> C lines of half
    CU(cudaMemcpyAsync(dst, ctx->stage.p, z.spec * sizeof(cpx), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
  }
  return NIS_OK;
}

// reference-layout arrays -> device frame; the column-major -> row-major conversion runs on the GPU (32 x 32 shared-memory tiles) out of
// a staging copy, so an import costs the PCIe transfer plus a few microseconds
static int upload_colmajor_f32(nis_ctx* ctx, const float* host, int R, int C, float* dst) {
  RESERVE(ctx->stage, (size_t)R * C * sizeof(float));
  CU(cudaMemcpyAsync(ctx->stage.p, host, (size_t)R * C * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  ctx->prof_stream = ctx->stream;
  LAUNCH(launch_transpose_f32(ctx->stage.as<float>(), dst, C, R, ctx->stream));           // [C][R] lines -> [R][C]
  CU(cudaStreamSynchronize(ctx->stream));                                                   // the staging buffer is reused by the next upload
  return NIS_OK;
}
static int upload_colmajor_cpx(nis_ctx* ctx, const float* host, int half, int C, cpx* dst) {
  RESERVE(ctx->stage, (size_t)half * C * sizeof(cpx));
  CU(cudaMemcpyAsync(ctx->stage.p, host, (size_t)half * C * sizeof(cpx), cudaMemcpyHostToDevice, ctx->stream));
  ctx->prof_stream = ctx->stream;
  LAUNCH(launch_transpose_cpx(ctx->stage.as<cpx>(), dst, C, half, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return NIS_OK;
}
static int h_factors(nis_ctx* ctx, cpx* F, cpx* P, cpx* Ht, cpx* Hp) {
  // the keyframe-only factors H = T/(kernel(.)/max + lambda) of both stages (the kernel id is validated at ComputePose, like the reference)
  if (ctx->cfg.kernel != 0 && ctx->cfg.kernel != 1) return NIS_OK;
  Lane& L = ctx->lanes[0];
  TRY(ensure_workspace(ctx, L, 1));
  TRY(hzz_batch(ctx, L, 0, src_slab<cpx>(F, 0), 1, Dst<cpx>{Ht, 0}));
  TRY(hzz_batch(ctx, L, 1, src_slab<cpx>(P, 0), 1, Dst<cpx>{Hp, 0}));
  return NIS_OK;
}

int nis_frame_import_ex(nis_ctx* ctx, const float* image_colmajor, const float* fft_result, const float* fft_polar, int with_h, nis_frame** out) {
  if (!ctx || !out || (!image_colmajor && !fft_result && !fft_polar) || (with_h && (!fft_result || !fft_polar))) return NIS_ERR_INVALID_ARGUMENT;
  CU(cudaSetDevice(ctx->device));
  nis_frame* f = nullptr;
  TRY(frame_alloc(ctx, false, &f));
  f->has_image = image_colmajor != nullptr;
  f->has_spectra = fft_result != nullptr;
  f->has_polar = fft_polar != nullptr;
  f->has_h = false;
  int st = NIS_OK;
  if (image_colmajor) st = upload_colmajor_f32(ctx, image_colmajor, ctx->H, ctx->W, f->img_f32);
  if (st == NIS_OK && fft_result) st = upload_colmajor_cpx(ctx, fft_result, ctx->H / 2 + 1, ctx->W, f->F);
  if (st == NIS_OK && fft_polar) st = upload_colmajor_cpx(ctx, fft_polar, ctx->D / 2 + 1, ctx->Cp, f->P);
  if (st == NIS_OK && with_h) {
    st = h_factors(ctx, f->F, f->P, f->Ht, f->Hp);
    if (st == NIS_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) st = fail(ctx, NIS_ERR_CUDA, "import sync", (int)cudaGetLastError());
    f->has_h = st == NIS_OK;
  }
  if (st != NIS_OK) { nis_frame_free(ctx, f); return st; }
  *out = f;
  return NIS_OK;
}

int nis_frame_import(nis_ctx* ctx, const float* image_colmajor, const float* fft_result, const float* fft_polar, nis_frame** out) {
  if (!image_colmajor || !fft_result || !fft_polar) return NIS_ERR_INVALID_ARGUMENT;
  return nis_frame_import_ex(ctx, image_colmajor, fft_result, fft_polar, 1, out);
}

// ---- ComputePose ----------------------------------------------------------------------------------------------
static void record_out(const PoseRecord& r, double pose[3], double info[3], int32_t peak[4]) {
  for (int i = 0; i < 3; ++i) { if (pose) pose[i] = r.pose[i]; if (info) info[i] = r.info[i]; }
  if (peak) for (int i = 0; i < 4; ++i) peak[i] = r.peak[i];
}

int nis_compute_pose(nis_ctx* ctx, const nis_frame* last, const nis_frame* cur, int not_large_rotation, double pose[3],
                     double info[3], int32_t peak_rc[4]) {
  if (!ctx || !last || !cur || !pose || !info) return NIS_ERR_INVALID_ARGUMENT;
  if (!last->has_spectra || !last->has_polar || !last->has_h) return fail(ctx, NIS_ERR_INVALID_ARGUMENT, "ComputePose: the last frame needs fft_result, fft_polar and its H factors");
  if (!cur->has_image || !cur->has_polar) return fail(ctx, NIS_ERR_INVALID_ARGUMENT, "ComputePose: the current frame needs its image and fft_polar");
  CU(cudaSetDevice(ctx->device));
  ctx->use_rot_cache = false;
  TRY(ensure_recs(ctx, 1));
  const bool u8 = cur->img_u8 != nullptr;
  TRY(compute_pose_batch(ctx, ctx->lanes[0], !not_large_rotation, src_slab<cpx>(last->F, 0), src_slab<cpx>(last->P, 0),
                         src_slab<cpx>(last->Ht, 0), src_slab<cpx>(last->Hp, 0), src_slab<cpx>(cur->P, 0),
                         src_slab<float>(cur->img_f32, 0), src_slab<uint8_t>(cur->img_u8, 0), u8, 1, 0, ctx->recs.as<PoseRecord>()));
  PoseRecord r;
  CU(cudaMemcpyAsync(&r, ctx->recs.p, sizeof r, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  record_out(r, pose, info, peak_rc);
  return NIS_OK;
}

// ---- stream tracking --------------------------------------------------------------------------------------------
static int track_stream_window(nis_ctx* ctx, const uint8_t* frames, bool on_host, int n, double* poses, double* infos);
// Device memory is bounded: the stream is processed in windows of at most kStreamWindow pairs (5.5 MB of features per frame, so
// <= 6 GB of slabs however long the stream is); consecutive windows share one frame, whose features are recomputed.
static const int kStreamWindow = 1024;
static int track_stream_impl(nis_ctx* ctx, const uint8_t* frames, bool on_host, int n, double* poses, double* infos) {
  if (!ctx || !frames || n < 1 || (n > 1 && (!poses || !infos))) return NIS_ERR_INVALID_ARGUMENT;
  if (n <= kStreamWindow + 1) return track_stream_window(ctx, frames, on_host, n, poses, infos);
  const size_t npx = ctx->sz[0].real;
  for (int t0 = 0; t0 < n - 1; t0 += kStreamWindow) {
    const int m = std::min(kStreamWindow + 1, n - t0);
    TRY(track_stream_window(ctx, frames + (size_t)t0 * npx, on_host, m, poses + 3 * (size_t)t0, infos + 3 * (size_t)t0));
  }
  return NIS_OK;
}
#ifndef NIS_STREAM_RAMP
#define NIS_STREAM_RAMP 1
#endif
static int track_stream_window(nis_ctx* ctx, const uint8_t* frames, bool on_host, int n, double* poses, double* infos) {
  CU(cudaSetDevice(ctx->device));
  const size_t npx = ctx->sz[0].real, spt = ctx->sz[0].spec, spp = ctx->sz[1].spec;
  RESERVE(ctx->sF, (size_t)n * spt * sizeof(cpx));
  RESERVE(ctx->sP, (size_t)n * spp * sizeof(cpx));
  RESERVE(ctx->sHt, (size_t)n * spt * sizeof(cpx));
  RESERVE(ctx->sHp, (size_t)n * spp * sizeof(cpx));
  const uint8_t* d_frames = frames;
  if (on_host) {
    RESERVE(ctx->sImg, (size_t)n * npx);
    d_frames = ctx->sImg.as<uint8_t>();      // uploaded batch by batch below, each on its lane, so copies overlap compute
  }
  const uint8_t* d_raw = d_frames;
  if (ctx->undistort) {                        // raw camera frames: undistorted batch by batch into their own slab
    RESERVE(ctx->sUnd, (size_t)n * npx);
    d_frames = ctx->sUnd.as<uint8_t>();
  }
  TRY(ensure_recs(ctx, std::max(n - 1, 1)));
  const int B = ctx->batch;
  const int NL = ctx->active_lanes;
  cpx* F = ctx->sF.as<cpx>(); cpx* P = ctx->sP.as<cpx>();
  cpx* Ht = ctx->sHt.as<cpx>(); cpx* Hp = ctx->sHp.as<cpx>();
  // batch k covers frames [starts[k], starts[k+1]).  Frames coming from the host start with a quarter batch: nothing can be computed
  // before the first upload has landed, so the pipeline is primed with a small one (NIS_STREAM_RAMP=0 builds without)
  std::vector<int> starts{0};
  if (NIS_STREAM_RAMP && on_host && n > B && B >= 8) starts.push_back(B / 4);
  while (starts.back() < n) starts.push_back(std::min(n, starts.back() + B));
  const int nbatch = (int)starts.size() - 1;
  while ((int)ctx->up_ev.size() < nbatch) {
    cudaEvent_t a = nullptr, b = nullptr;
    CU(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
    ctx->up_ev.push_back(a);
    CU(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
    ctx->feat_ev.push_back(b);
  }
  // One software pipeline over the batches: uploads run back to back on the copy stream from the start; batch k's features go to
  // lane k % NL as soon as its upload has landed; the pose solves of batch k-1 (which need the first frame of batch k) follow on
  // the same lane, so copies overlap ALL of the compute and a batch's features are still in L2 when its solves read them.
  TRY(fork_lanes(ctx));
  if (on_host) {
    CU(cudaEventRecord(ctx->fork_ev, ctx->stream));               // the copy stream, too, starts after everything already queued
    CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->fork_ev, 0));
    for (int k = 0; k < nbatch; ++k) {
      const int t0 = starts[k], nb = starts[k + 1] - t0;
      CU(cudaMemcpyAsync(ctx->sImg.as<uint8_t>() + (size_t)t0 * npx, frames + (size_t)t0 * npx, (size_t)nb * npx, cudaMemcpyHostToDevice,
                         ctx->copy_stream));
      CU(cudaEventRecord(ctx->up_ev[k], ctx->copy_stream));
    }
  }
  auto pose_of_batch = [&](int j, Lane& L) -> int {                 // pairs (t, t+1) with t in batch j
    const int p0 = starts[j], nb = std::min(starts[j + 1] - p0, n - 1 - p0);
    if (nb <= 0) return NIS_OK;
    return compute_pose_batch(ctx, L, false, src_slab<cpx>(F + (size_t)p0 * spt, (long long)spt),
                              src_slab<cpx>(P + (size_t)p0 * spp, (long long)spp), src_slab<cpx>(Ht + (size_t)p0 * spt, (long long)spt),
                              src_slab<cpx>(Hp + (size_t)p0 * spp, (long long)spp), src_slab<cpx>(P + (size_t)(p0 + 1) * spp, (long long)spp),
                              src_null<float>(), src_slab<uint8_t>(d_frames + (size_t)(p0 + 1) * npx, (long long)npx), true, nb, p0,
                              ctx->recs.as<PoseRecord>() + p0);
  };
  for (int k = 0; k < nbatch; ++k) {
    const int t0 = starts[k], nb = starts[k + 1] - t0;
    Lane& L = ctx->lanes[k % NL];
    if (on_host) CU(cudaStreamWaitEvent(L.stream, ctx->up_ev[k], 0));
    if (ctx->undistort) TRY(undistort_batch(ctx, L.stream, d_raw + (size_t)t0 * npx, ctx->sUnd.as<uint8_t>() + (size_t)t0 * npx, nb));
    TRY(features_batch(ctx, L, src_null<float>(), src_slab<uint8_t>(d_frames + (size_t)t0 * npx, (long long)npx), true, nb,
                       Dst<cpx>{F + (size_t)t0 * spt, (long long)spt}, Dst<cpx>{P + (size_t)t0 * spp, (long long)spp},
                       Dst<cpx>{Ht + (size_t)t0 * spt, (long long)spt}, Dst<cpx>{Hp + (size_t)t0 * spp, (long long)spp}, true));
    CU(cudaEventRecord(ctx->feat_ev[k], L.stream));
    if (k >= 1) {                                                   // solves of batch k-1: features of batches k-1 (other lane) and k (this lane)
      if (NL > 1) CU(cudaStreamWaitEvent(L.stream, ctx->feat_ev[k - 1], 0));
      TRY(pose_of_batch(k - 1, L));
    }
  }
  TRY(pose_of_batch(nbatch - 1, ctx->lanes[(nbatch - 1) % NL]));    // the last batch's own pairs
  TRY(join_lanes(ctx));
  if (n > 1) {
    TRY(ensure_pinned(ctx, (size_t)(n - 1) * sizeof(PoseRecord)));
    CU(cudaMemcpyAsync(ctx->pin, ctx->recs.p, (size_t)(n - 1) * sizeof(PoseRecord), cudaMemcpyDeviceToHost, ctx->stream));
  }
  CU(cudaStreamSynchronize(ctx->stream));
  const PoseRecord* r = (const PoseRecord*)ctx->pin;
  for (int i = 0; i < n - 1; ++i) record_out(r[i], poses + 3 * i, infos + 3 * i, nullptr);
  return NIS_OK;
}
int nis_track_stream(nis_ctx* ctx, const uint8_t* frames_host, int n, double* poses, double* infos) {
  return track_stream_impl(ctx, frames_host, true, n, poses, infos);
}
int nis_track_stream_dev(nis_ctx* ctx, const uint8_t* frames_dev, int n, double* poses, double* infos) {
  return track_stream_impl(ctx, frames_dev, false, n, poses, infos);
}

// ---- stream tracking under the reference's keyframe policy (map_builder.cc:30-70 without loop closure / optimisation) ----
int nis_track_stream_keyframes(nis_ctx* ctx, const uint8_t* frames_host, int n, const nis_kfs_config* kfs, const nis_camera_model* cam,
                               nis_track_result* out) {
  if (!ctx || !frames_host || n < 1 || !kfs || !cam || !out || cam->height < 0 || cam->fx == 0 || cam->fy == 0) return NIS_ERR_INVALID_ARGUMENT;
  CU(cudaSetDevice(ctx->device));
  const size_t npx = ctx->sz[0].real, spt = ctx->sz[0].spec, spp = ctx->sz[1].spec;
  // Bounded device memory: frames are processed in windows of kStreamWindow; features AND keyframe factors H of a whole window are
  // computed in full batches (any frame may become a keyframe; one frame's H on demand inside the sequential loop would cost six
  // latency-bound launches per keyframe).  The current keyframe's record moves to a dedicated slot when its window is left.
  const int M = std::min(n, kStreamWindow);
  RESERVE(ctx->sF, (size_t)M * spt * sizeof(cpx));
  RESERVE(ctx->sP, (size_t)M * spp * sizeof(cpx));
  RESERVE(ctx->sHt, (size_t)M * spt * sizeof(cpx));
  RESERVE(ctx->sHp, (size_t)M * spp * sizeof(cpx));
  RESERVE(ctx->sImg, (size_t)M * npx);
  RESERVE(ctx->kfrec, 2 * (spt + spp) * sizeof(cpx));
  if (ctx->undistort) RESERVE(ctx->sUnd, (size_t)M * npx);
  const int B = ctx->batch, NL = ctx->active_lanes;
  TRY(ensure_recs(ctx, B));
  TRY(ensure_pinned(ctx, (size_t)B * sizeof(PoseRecord)));
  cpx* F = ctx->sF.as<cpx>(); cpx* P = ctx->sP.as<cpx>();
  cpx* Ht = ctx->sHt.as<cpx>(); cpx* Hp = ctx->sHp.as<cpx>();
  const int max_batches = (M + B - 1) / B;
  while ((int)ctx->up_ev.size() < max_batches) {
    cudaEvent_t a = nullptr, b = nullptr;
    CU(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
    ctx->up_ev.push_back(a);
    CU(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
    ctx->feat_ev.push_back(b);
  }
  Lane& L0 = ctx->lanes[0];
  pose::TrackerState st;
  pose::initialize(*cam, st, out[0]);
  pose::snapshot(st, out[0]);
  // keyframe operands: pointers into the window slabs, or into the dedicated slot once the window has moved on
  const cpx *kF = nullptr, *kP = nullptr, *kHt = nullptr, *kHp = nullptr;
  int K = 0, spec = std::min(B, 8), since_kf = 0;
  for (int w0 = 0; w0 < n; w0 += M) {
    const int m = std::min(M, n - w0);
    const uint8_t* d_raw = ctx->sImg.as<uint8_t>();
    const uint8_t* d_frames = ctx->undistort ? ctx->sUnd.as<uint8_t>() : d_raw;
    if (w0 > 0) {                                           // park the keyframe before its slab entry is overwritten
      cpx* kr = ctx->kfrec.as<cpx>();
      if (kF != kr) {
        CU(cudaMemcpyAsync(kr, kF, spt * sizeof(cpx), cudaMemcpyDeviceToDevice, ctx->stream));
        CU(cudaMemcpyAsync(kr + spt, kP, spp * sizeof(cpx), cudaMemcpyDeviceToDevice, ctx->stream));
        CU(cudaMemcpyAsync(kr + spt + spp, kHt, spt * sizeof(cpx), cudaMemcpyDeviceToDevice, ctx->stream));
        CU(cudaMemcpyAsync(kr + 2 * spt + spp, kHp, spp * sizeof(cpx), cudaMemcpyDeviceToDevice, ctx->stream));
        kF = kr; kP = kr + spt; kHt = kr + spt + spp; kHp = kr + 2 * spt + spp;
      }
    }
    const int nbatch = (m + B - 1) / B;
    TRY(fork_lanes(ctx));
    CU(cudaEventRecord(ctx->fork_ev, ctx->stream));
    CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->fork_ev, 0));
    for (int k = 0; k < nbatch; ++k) {
      const int t0 = k * B, nb = std::min(B, m - t0);
      CU(cudaMemcpyAsync(ctx->sImg.as<uint8_t>() + (size_t)t0 * npx, frames_host + (size_t)(w0 + t0) * npx, (size_t)nb * npx, cudaMemcpyHostToDevice,
                         ctx->copy_stream));
      CU(cudaEventRecord(ctx->up_ev[k], ctx->copy_stream));
    }
    for (int k = 0; k < nbatch; ++k) {
      const int t0 = k * B, nb = std::min(B, m - t0);
      Lane& L = ctx->lanes[k % NL];
      CU(cudaStreamWaitEvent(L.stream, ctx->up_ev[k], 0));
      if (ctx->undistort) TRY(undistort_batch(ctx, L.stream, d_raw + (size_t)t0 * npx, ctx->sUnd.as<uint8_t>() + (size_t)t0 * npx, nb));
      TRY(features_batch(ctx, L, src_null<float>(), src_slab<uint8_t>(d_frames + (size_t)t0 * npx, (long long)npx), true, nb,
                         Dst<cpx>{F + (size_t)t0 * spt, (long long)spt}, Dst<cpx>{P + (size_t)t0 * spp, (long long)spp},
                         Dst<cpx>{Ht + (size_t)t0 * spt, (long long)spt}, Dst<cpx>{Hp + (size_t)t0 * spp, (long long)spp}, true));
    }
    TRY(join_lanes(ctx));
    if (w0 == 0) { kF = F; kP = P; kHt = Ht; kHp = Hp; }     // frame 0 is the first keyframe (map_builder.cc:35-40)
    int t = w0 == 0 ? 1 : w0;
    while (t < w0 + m) {
      const int nb = std::min(spec, w0 + m - t), lt = t - w0;
      // frames t .. t+nb-1 against keyframe K: the keyframe operands are one slab with stride 0
      TRY(compute_pose_batch(ctx, L0, false, src_slab<cpx>(kF, 0), src_slab<cpx>(kP, 0), src_slab<cpx>(kHt, 0), src_slab<cpx>(kHp, 0),
                             src_slab<cpx>(P + (size_t)lt * spp, (long long)spp), src_null<float>(),
                             src_slab<uint8_t>(d_frames + (size_t)lt * npx, (long long)npx), true, nb, t, ctx->recs.as<PoseRecord>()));
      CU(cudaMemcpyAsync(ctx->pin, ctx->recs.p, (size_t)nb * sizeof(PoseRecord), cudaMemcpyDeviceToHost, ctx->stream));
      CU(cudaStreamSynchronize(ctx->stream));
      const PoseRecord* r = (const PoseRecord*)ctx->pin;
      int used = nb;
      for (int i = 0; i < nb; ++i) {
        nis_track_result& o = out[t + i];
        o.keyframe = K;
        const bool ins = pose::step(*cam, *kfs, ctx->W, ctx->H, r[i].pose, r[i].info, st, o);
        pose::snapshot(st, o);
        ++since_kf;
        if (ins) {                                   // frames behind it were solved against the wrong keyframe: redo them
          K = t + i;
          const size_t lk = (size_t)(K - w0);
          kF = F + lk * spt; kP = P + lk * spp; kHt = Ht + lk * spt; kHp = Hp + lk * spp;
          used = i + 1;
          spec = std::max(4, std::min(B, 2 * since_kf));
          since_kf = 0;
          break;
        }
      }
      if (used == nb && since_kf >= spec) spec = std::min(B, 2 * spec);
      t += used;
    }
  }
  return NIS_OK;
}

// ---- keyframe DB ------------------------------------------------------------------------------------------------
// Record per keyframe, by store mode (nis_db_set_mode):
//   NIS_DB_FULL    [F][P][Ht][Hp]  5.24 MB @640x480: a candidate costs only its own solve
//   NIS_DB_SPECTRA [F][P]          2.62 MB: the reference's own Frame payload (include/frame.h:35-36); Ht, Hp recomputed per batch
//   NIS_DB_IMAGE   u8 image        0.31 MB: everything recomputed per batch (100 k keyframes = 31 GB, fits one GPU)
// The recomputation runs the same kernels on the same data, so a scan returns the same bits in every mode.
static size_t db_record_bytes(const nis_ctx* ctx) {
  const size_t spec = (ctx->sz[0].spec + ctx->sz[1].spec) * sizeof(cpx);
  return ctx->db_mode == NIS_DB_FULL ? 2 * spec : (ctx->db_mode == NIS_DB_SPECTRA ? spec : ctx->sz[0].real);
}

static void db_unreserve(nis_ctx* ctx, int old) { ctx->slot_ptr.resize(old); }      // chunks stay allocated for the next insert

// Reserves n_new record slots.  Transactional: on any failure the slot table and the chunk list are exactly as before the call.
static int db_reserve_slots(nis_ctx* ctx, int n_new, int* first) {
  const size_t rec = db_record_bytes(ctx);
  const int old = (int)ctx->slot_ptr.size();
  const size_t old_chunks = ctx->chunks.size();
  *first = old;
  auto rollback = [&]() {
    ctx->slot_ptr.resize(old);
    while (ctx->chunks.size() > old_chunks) { cudaFree(ctx->chunks.back()); ctx->chunks.pop_back(); }
  };
  for (int i = 0; i < n_new; ++i) {
    const int slot = old + i;
    const int ch = slot / ctx->chunk_slots, within = slot % ctx->chunk_slots;
    if (ch >= (int)ctx->chunks.size()) {
      void* p = nullptr;
      cudaError_t e = cudaMalloc(&p, (size_t)ctx->chunk_slots * rec);
      if (e != cudaSuccess) { rollback(); return fail(ctx, NIS_ERR_OUT_OF_MEMORY, "cudaMalloc DB chunk", (int)e); }
      ctx->chunks.push_back(p);
    }
    ctx->slot_ptr.push_back((char*)ctx->chunks[ch] + (size_t)within * rec);
  }
  const int total = old + n_new;
  cudaError_t e = cudaSuccess;
  if (total > ctx->d_slot_cap) {
    int cap = std::max(4096, ctx->d_slot_cap);
    while (cap < total) cap *= 2;
    e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) {
      ctx->d_slot_cap = 0;
      if (ctx->d_slot_ptr.reserve((size_t)cap * sizeof(void*))) { rollback(); return fail(ctx, NIS_ERR_OUT_OF_MEMORY, "cudaMalloc DB slot table"); }
      ctx->d_slot_cap = cap;
      e = h2d(ctx, ctx->d_slot_ptr.p, ctx->slot_ptr.data(), (size_t)total * sizeof(void*));
    }
  } else {
    e = h2d(ctx, ctx->d_slot_ptr.as<void*>() + old, ctx->slot_ptr.data() + old, (size_t)n_new * sizeof(void*));
  }
  if (e != cudaSuccess) { rollback(); return fail(ctx, NIS_ERR_CUDA, "DB slot table upload", (int)e); }
  return NIS_OK;
}

static void db_push_meta(nis_ctx* ctx, int frame_id, double dist) {
  ctx->slot_frame_id.push_back(frame_id);
  ctx->slot_dist.push_back(dist);
  ctx->meta_dirty = true;
}

int nis_db_set_mode(nis_ctx* ctx, int mode) {
  if (!ctx || mode < NIS_DB_FULL || mode > NIS_DB_IMAGE) return NIS_ERR_INVALID_ARGUMENT;
  if (nis_db_size(ctx) != 0) return fail(ctx, NIS_ERR_INVALID_ARGUMENT, "nis_db_set_mode: the keyframe store must be empty");
  if (mode != ctx->db_mode) {                       // chunk size depends on the record size
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    for (void* c : ctx->chunks) cudaFree(c);
    ctx->chunks.clear();
  }
  ctx->db_mode = mode;
  return NIS_OK;
}
int nis_db_mode(const nis_ctx* ctx) { return ctx ? ctx->db_mode : -1; }

int nis_db_add(nis_ctx* ctx, const nis_frame* f, int frame_id, double acc_distance, int* slot) {
  if (!ctx || !f) return NIS_ERR_INVALID_ARGUMENT;
  if (ctx->db_mode == NIS_DB_IMAGE && !f->img_u8) return fail(ctx, NIS_ERR_INVALID_ARGUMENT, "NIS_DB_IMAGE stores u8 images: the frame has none");
  if (ctx->db_mode != NIS_DB_IMAGE && !(f->has_spectra && f->has_polar)) return fail(ctx, NIS_ERR_INVALID_ARGUMENT, "the frame carries no spectra");
  if (ctx->db_mode == NIS_DB_FULL && !f->has_h) return fail(ctx, NIS_ERR_INVALID_ARGUMENT, "the frame was imported without its H factors");
  CU(cudaSetDevice(ctx->device));
  int s0 = 0;
  TRY(db_reserve_slots(ctx, 1, &s0));
  // a frame's F, P, Ht, Hp are contiguous in the same order as a record
  const void* src = ctx->db_mode == NIS_DB_IMAGE ? (const void*)f->img_u8 : (const void*)f->F;
  cudaError_t e = cudaMemcpyAsync(ctx->slot_ptr[s0], src, db_record_bytes(ctx), cudaMemcpyDeviceToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) { db_unreserve(ctx, s0); return fail(ctx, NIS_ERR_CUDA, "DB record copy", (int)e); }
  db_push_meta(ctx, frame_id, acc_distance);
  if (slot) *slot = s0;
  return NIS_OK;
}

// Map::AddFrame for a keyframe whose spectra the caller holds as reference-layout arrays (Frame::GetFFTResult): straight into a
// record, H factors computed in place; no image needed (a keyframe's image is never read by the scan).  Not for NIS_DB_IMAGE.
int nis_db_add_spectra(nis_ctx* ctx, const float* fft_result, const float* fft_polar, int frame_id, double acc_distance, int* slot) {
  if (!ctx || !fft_result || !fft_polar) return NIS_ERR_INVALID_ARGUMENT;
  if (ctx->db_mode == NIS_DB_IMAGE) return fail(ctx, NIS_ERR_INVALID_ARGUMENT, "NIS_DB_IMAGE stores u8 images, not spectra");
  CU(cudaSetDevice(ctx->device));
  int s0 = 0;
  TRY(db_reserve_slots(ctx, 1, &s0));
  cpx* F = (cpx*)ctx->slot_ptr[s0];
  cpx* P = F + ctx->sz[0].spec;
  int st = upload_colmajor_cpx(ctx, fft_result, ctx->H / 2 + 1, ctx->W, F);
  if (st == NIS_OK) st = upload_colmajor_cpx(ctx, fft_polar, ctx->D / 2 + 1, ctx->Cp, P);
  if (st == NIS_OK && ctx->db_mode == NIS_DB_FULL) st = h_factors(ctx, F, P, P + ctx->sz[1].spec, P + ctx->sz[1].spec + ctx->sz[0].spec);
  if (st == NIS_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) st = fail(ctx, NIS_ERR_CUDA, "db_add_spectra sync", (int)cudaGetLastError());
  if (st != NIS_OK) { db_unreserve(ctx, s0); return st; }
  db_push_meta(ctx, frame_id, acc_distance);
  if (slot) *slot = s0;
  return NIS_OK;
}

static int db_add_images_impl(nis_ctx* ctx, const uint8_t* images, bool on_host, int n, const int* ids, const double* dists) {
  if (!ctx || (!images && n > 0) || n < 0) return NIS_ERR_INVALID_ARGUMENT;
  if (n == 0) return NIS_OK;
  CU(cudaSetDevice(ctx->device));
  const size_t npx = ctx->sz[0].real, spt = ctx->sz[0].spec, spp = ctx->sz[1].spec, rec = db_record_bytes(ctx) / sizeof(cpx);
  int s0 = 0;
  TRY(db_reserve_slots(ctx, n, &s0));
  auto body = [&]() -> int {
    const int B = ctx->batch;
    const int NL = ctx->active_lanes;
    const uint8_t* d_images = images;
    if (on_host) {                      // stage the whole upload once (u8 images are 1/8.5 of the features they turn into)
      RESERVE(ctx->sImg, (size_t)n * npx);
      CU(cudaMemcpyAsync(ctx->sImg.p, images, (size_t)n * npx, cudaMemcpyHostToDevice, ctx->stream));
      d_images = ctx->sImg.as<uint8_t>();
    }
    const uint8_t* d_raw = d_images;
    if (ctx->undistort) {
      RESERVE(ctx->sUnd, (size_t)n * npx);
      d_images = ctx->sUnd.as<uint8_t>();
    }
    TRY(fork_lanes(ctx));
    for (int i0 = 0, k = 0; i0 < n; ++k) {
      const int slot = s0 + i0;
      const int room = ctx->chunk_slots - slot % ctx->chunk_slots;       // stay inside one chunk (contiguous records)
      const int nb = std::min(std::min(B, n - i0), room);
      Lane& L = ctx->lanes[k % NL];
      if (ctx->undistort) TRY(undistort_batch(ctx, L.stream, d_raw + (size_t)i0 * npx, ctx->sUnd.as<uint8_t>() + (size_t)i0 * npx, nb));
      if (ctx->db_mode == NIS_DB_IMAGE) {
        CU(cudaMemcpyAsync(ctx->slot_ptr[slot], d_images + (size_t)i0 * npx, (size_t)nb * npx, cudaMemcpyDeviceToDevice, L.stream));
      } else {
        cpx* base = (cpx*)ctx->slot_ptr[slot];
        const bool full = ctx->db_mode == NIS_DB_FULL;
        TRY(features_batch(ctx, L, src_null<float>(), src_slab<uint8_t>(d_images + (size_t)i0 * npx, (long long)npx), true, nb,
                           Dst<cpx>{base, (long long)rec}, Dst<cpx>{base + spt, (long long)rec},
                           Dst<cpx>{full ? base + spt + spp : nullptr, (long long)rec}, Dst<cpx>{full ? base + 2 * spt + spp : nullptr, (long long)rec},
                           full));
      }
      i0 += nb;
    }
    TRY(join_lanes(ctx));
    CU(cudaStreamSynchronize(ctx->stream));
    return NIS_OK;
  };
  const int st = body();
  if (st != NIS_OK) { cudaDeviceSynchronize(); db_unreserve(ctx, s0); return st; }
  for (int i = 0; i < n; ++i) db_push_meta(ctx, ids ? ids[i] : s0 + i, dists ? dists[i] : 0.0);
  return NIS_OK;
}
int nis_db_add_images(nis_ctx* ctx, const uint8_t* images_host, int n, const int* frame_ids, const double* acc_distances) {
  return db_add_images_impl(ctx, images_host, true, n, frame_ids, acc_distances);
}
int nis_db_add_images_dev(nis_ctx* ctx, const uint8_t* images_dev, int n, const int* frame_ids, const double* acc_distances) {
  return db_add_images_impl(ctx, images_dev, false, n, frame_ids, acc_distances);
}
int nis_db_size(const nis_ctx* ctx) { return ctx ? (int)ctx->slot_frame_id.size() : 0; }
int nis_db_clear(nis_ctx* ctx) {
  if (!ctx) return NIS_ERR_INVALID_ARGUMENT;
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  for (void* c : ctx->chunks) cudaFree(c);
  ctx->chunks.clear(); ctx->slot_ptr.clear(); ctx->slot_frame_id.clear(); ctx->slot_dist.clear();
  ctx->slot_cell.clear(); ctx->meta_dirty = true;
  return NIS_OK;
}

// keyframe meta data the device-side candidate selection reads (frame id, accumulated distance, grid cell), re-uploaded when dirty
static int db_sync_meta(nis_ctx* ctx) {
  const size_t n = ctx->slot_frame_id.size();
  if (!ctx->meta_dirty || n == 0) return NIS_OK;
  ctx->slot_cell.resize(n, {INT_MIN, INT_MIN});
  CU(cudaStreamSynchronize(ctx->stream));
  size_t cap = std::max<size_t>(4096, ctx->d_fid.bytes / sizeof(int));
  while (cap < n) cap *= 2;
  RESERVE(ctx->d_fid, cap * sizeof(int));
  RESERVE(ctx->d_dist, cap * sizeof(double));
  RESERVE(ctx->d_cell, cap * sizeof(int2));
  std::vector<int2> cells(n);
  for (size_t i = 0; i < n; ++i) cells[i] = make_int2(ctx->slot_cell[i].first, ctx->slot_cell[i].second);
  CU(h2d(ctx, ctx->d_fid.p, ctx->slot_frame_id.data(), n * sizeof(int)));
  CU(h2d(ctx, ctx->d_dist.p, ctx->slot_dist.data(), n * sizeof(double)));
  CU(h2d(ctx, ctx->d_cell.p, cells.data(), n * sizeof(int2)));
  ctx->meta_dirty = false;
  return NIS_OK;
}

static void loop_result_init(nis_loop_result* out) {
  memset(out, 0, sizeof *out);
  out->slot = -1; out->frame_id = -1;
  out->response[0] = out->response[1] = out->response[2] = -1.0;                                                       // loop_closure.h:15
  for (int i = 0; i < 4; ++i) out->peak[i] = -1;
}

// materialise what the store mode does not keep for the nb candidates idx[0..nb) into the lane's record scratch
static int db_materialize(nis_ctx* ctx, Lane& L, const int* idx, int nb, Src<cpx>& Fz, Src<cpx>& Pz, Src<cpx>& Htz, Src<cpx>& Hpz) {
  const long long spt = (long long)ctx->sz[0].spec, spp = (long long)ctx->sz[1].spec;
  const void* const* ptrs = ctx->d_slot_ptr.as<const void*>();
  if (ctx->db_mode == NIS_DB_FULL) {
    const cpx* const* p = (const cpx* const*)ptrs;
    Fz = Src<cpx>{nullptr, 0, p, 0, idx, 0};
    Pz = Src<cpx>{nullptr, 0, p, spt, idx, 0};
    Htz = Src<cpx>{nullptr, 0, p, spt + spp, idx, 0};
    Hpz = Src<cpx>{nullptr, 0, p, 2 * spt + spp, idx, 0};
    return NIS_OK;
  }
  if (nb > L.db_cap) {
    CU(cudaStreamSynchronize(L.stream));
    L.db_cap = 0;
    const int cap = std::max(nb, ctx->batch);
    RESERVE(L.dbrec, (size_t)cap * 2 * (size_t)(spt + spp) * sizeof(cpx));
    L.db_cap = cap;
  }
  cpx* F = L.dbrec.as<cpx>();
  cpx* P = F + (size_t)L.db_cap * spt;
  cpx* Ht = P + (size_t)L.db_cap * spp;
  cpx* Hp = Ht + (size_t)L.db_cap * spt;
  TRY(ensure_workspace(ctx, L, nb));
  if (ctx->db_mode == NIS_DB_IMAGE) {
    Src<uint8_t> img{nullptr, 0, (const uint8_t* const*)ptrs, 0, idx, 0};
    TRY(features_batch(ctx, L, src_null<float>(), img, true, nb, Dst<cpx>{F, spt}, Dst<cpx>{P, spp}, Dst<cpx>{Ht, spt}, Dst<cpx>{Hp, spp}, true));
    Fz = src_slab<cpx>(F, spt); Pz = src_slab<cpx>(P, spp);
  } else {
    const cpx* const* p = (const cpx* const*)ptrs;
    Fz = Src<cpx>{nullptr, 0, p, 0, idx, 0};
    Pz = Src<cpx>{nullptr, 0, p, spt, idx, 0};
    TRY(hzz_batch(ctx, L, 0, Fz, nb, Dst<cpx>{Ht, spt}));
    TRY(hzz_batch(ctx, L, 1, Pz, nb, Dst<cpx>{Hp, spp}));
  }
  Htz = src_slab<cpx>(Ht, spt); Hpz = src_slab<cpx>(Hp, spp);
  return NIS_OK;
}

// ---- loop-closure scan (loop_closure.cc:36-73) --------------------------------------------------------------------
// prior != nullptr: candidates = the keyframes filed in the 3 x 3 grid cells around prior (x, y), loop_closure.cc:17-34
static int loop_scan_impl(nis_ctx* ctx, const nis_frame* query, int query_frame_id, double query_acc_distance, const nis_loop_config* cfg,
                          const int32_t* candidate_slots, int n_candidates, const double* prior, double grid_scale, nis_loop_result* out,
                          double* all_responses, nis_scan_record* records, int32_t* candidates_out, int max_candidates, int* n_candidates_out) {
  if (!ctx || !query || !cfg || !out || n_candidates < 0) return NIS_ERR_INVALID_ARGUMENT;
  if (!query->has_image || !query->has_polar) return fail(ctx, NIS_ERR_INVALID_ARGUMENT, "FindLoopClosure: the query needs its image and fft_polar");
  CU(cudaSetDevice(ctx->device));
  const int ndb = nis_db_size(ctx);
  const int n_in = candidate_slots ? n_candidates : ndb;
  loop_result_init(out);
  if (all_responses) for (int i = 0; i < 3 * n_in; ++i) all_responses[i] = -1.0;
  if (records) for (int i = 0; i < n_in; ++i) { memset(&records[i], 0, sizeof records[i]); for (int k = 0; k < 3; ++k) records[i].response[k] = -1.0; }
  if (n_candidates_out) *n_candidates_out = 0;
  if (candidate_slots)
    for (int i = 0; i < n_in; ++i)
      if (candidate_slots[i] < 0 || candidate_slots[i] >= ndb) return fail(ctx, NIS_ERR_INVALID_ARGUMENT, "candidate slot out of range");
  if (n_in == 0 || ndb == 0) return NIS_OK;
  TRY(db_sync_meta(ctx));
  // device-side candidate selection: filters (:43-53) and, with a prior pose, the grid neighbourhood (map.cc:81-101)
  if (n_in > ctx->cand_cap) {
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->cand_cap = 0;
    const int cap = std::max(n_in, 4096);
    RESERVE(ctx->cand, (size_t)cap * sizeof(int));
    RESERVE(ctx->cand_in, (size_t)cap * sizeof(int));
    RESERVE(ctx->cand_pos, (size_t)cap * sizeof(int));
    RESERVE(ctx->sel_scratch, (size_t)select_scratch_ints(cap) * sizeof(int));
    ctx->cand_cap = cap;
  }
  if (candidate_slots) CU(cudaMemcpyAsync(ctx->cand_in.p, candidate_slots, (size_t)n_in * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  SelectArgs sa{n_in, candidate_slots ? ctx->cand_in.as<int>() : nullptr, ctx->d_fid.as<int>(), ctx->d_dist.as<double>(), ctx->d_cell.as<int2>(),
                query_frame_id, query_acc_distance, cfg->frame_gap_thr, cfg->distance_thr, prior ? 1 : 0, 0, 0};
  if (prior) { const auto c0 = grid_cell(prior[0], prior[1], grid_scale); sa.cx = c0.first; sa.cy = c0.second; }
  int* d_n = ctx->sel_scratch.as<int>() + select_scratch_ints(ctx->cand_cap) - 1;
  ctx->prof_stream = ctx->stream;
  LAUNCH(launch_select(sa, ctx->sel_scratch.as<int>(), ctx->cand.as<int>(), ctx->cand_pos.as<int>(), d_n, ctx->stream));
  ctx->launches += 2;
  int n = 0;
  CU(cudaMemcpyAsync(&n, d_n, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  out->evaluated = n;
  if (n_candidates_out) *n_candidates_out = n;
  std::vector<int> pos;
  if (all_responses || records || candidates_out) {
    pos.resize(n);
    if (n) CU(cudaMemcpyAsync(pos.data(), ctx->cand_pos.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  }
  std::vector<int> cand_h;
  if (candidates_out && n) {
    cand_h.resize(n);
    CU(cudaMemcpyAsync(cand_h.data(), ctx->cand.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < n && i < max_candidates; ++i) candidates_out[i] = cand_h[i];
  }
  if (n == 0) return NIS_OK;
  TRY(ensure_recs(ctx, n));
  const bool u8 = query->img_u8 != nullptr;
  const int B = ctx->batch;
  const int NL = ctx->active_lanes;
  ctx->use_rot_cache = ctx->rot_cache_min > 0 && n >= ctx->rot_cache_min;
  if (ctx->use_rot_cache) TRY(build_rot_cache(ctx, src_slab<float>(query->img_f32, 0), src_slab<uint8_t>(query->img_u8, 0), u8));
  TRY(fork_lanes(ctx));
  for (int b0 = 0, k = 0; b0 < n; b0 += B, ++k) {
    const int nb = std::min(B, n - b0);
    const int* idx = ctx->cand.as<int>() + b0;
    Lane& L = ctx->lanes[k % NL];
    Src<cpx> Fz, Pz, Htz, Hpz;
    TRY(db_materialize(ctx, L, idx, nb, Fz, Pz, Htz, Hpz));
    TRY(compute_pose_batch(ctx, L, true, Fz, Pz, Htz, Hpz, src_slab<cpx>(query->P, 0), src_slab<float>(query->img_f32, 0),
                           src_slab<uint8_t>(query->img_u8, 0), u8, nb, b0, ctx->recs.as<PoseRecord>() + b0));
  }
  TRY(join_lanes(ctx));
  ctx->use_rot_cache = false;
  ctx->prof_stream = ctx->stream;
  LAUNCH(launch_scan_reduce(ctx->recs.as<PoseRecord>(), n, ctx->best.as<PoseRecord>(), ctx->stream));
  PoseRecord best;
  CU(cudaMemcpyAsync(&best, ctx->best.p, sizeof best, cudaMemcpyDeviceToHost, ctx->stream));
  std::vector<PoseRecord> all;
  if (all_responses || records) {
    all.resize(n);
    CU(cudaMemcpyAsync(all.data(), ctx->recs.p, (size_t)n * sizeof(PoseRecord), cudaMemcpyDeviceToHost, ctx->stream));
  }
  CU(cudaStreamSynchronize(ctx->stream));
  if (all_responses)
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < 3; ++k) all_responses[3 * pos[i] + k] = all[i].info[k];
  if (records)
    for (int i = 0; i < n; ++i) {
      nis_scan_record& r = records[pos[i]];
      r.evaluated = 1; r.hyp = all[i].hyp;
      for (int k = 0; k < 3; ++k) { r.relative_pose[k] = all[i].pose[k]; r.response[k] = all[i].info[k]; }
      for (int k = 0; k < 4; ++k) r.peak[k] = all[i].peak[k];
    }
  if (best.index >= 0) {
    int best_slot = -1;
    CU(cudaMemcpyAsync(&best_slot, ctx->cand.as<int>() + best.index, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    out->slot = best_slot;
    out->frame_id = ctx->slot_frame_id[out->slot];
    out->hyp = best.hyp;
    for (int k = 0; k < 3; ++k) { out->relative_pose[k] = best.pose[k]; out->response[k] = best.info[k]; }
    for (int k = 0; k < 4; ++k) out->peak[k] = best.peak[k];
  }
  out->found = (out->response[0] > cfg->position_response_thr) && (out->response[2] > cfg->angle_response_thr);        // :68-71
  return NIS_OK;
}

int nis_loop_scan(nis_ctx* ctx, const nis_frame* query, int query_frame_id, double query_acc_distance, const nis_loop_config* cfg,
                  const int32_t* candidate_slots, int n_candidates, nis_loop_result* out, double* all_responses) {
  return loop_scan_impl(ctx, query, query_frame_id, query_acc_distance, cfg, candidate_slots, n_candidates, nullptr, 0.0, out, all_responses,
                        nullptr, nullptr, 0, nullptr);
}
int nis_loop_scan_records(nis_ctx* ctx, const nis_frame* query, int query_frame_id, double query_acc_distance, const nis_loop_config* cfg,
                          const int32_t* candidate_slots, int n_candidates, nis_loop_result* out, nis_scan_record* records) {
  return loop_scan_impl(ctx, query, query_frame_id, query_acc_distance, cfg, candidate_slots, n_candidates, nullptr, 0.0, out, nullptr,
                        records, nullptr, 0, nullptr);
}

int nis_db_set_position(nis_ctx* ctx, int slot, double x, double y, double grid_scale) {
  if (!ctx || slot < 0 || slot >= nis_db_size(ctx) || !(grid_scale > 0)) return NIS_ERR_INVALID_ARGUMENT;
  ctx->slot_cell.resize(ctx->slot_frame_id.size(), {INT_MIN, INT_MIN});
  if (ctx->slot_cell[slot].first != INT_MIN) return fail(ctx, NIS_ERR_INVALID_ARGUMENT, "slot already filed in the grid (Map::AddFrame files a frame once)");
  ctx->slot_cell[slot] = grid_cell(x, y, grid_scale);
  ctx->meta_dirty = true;
  return NIS_OK;
}

int nis_loop_scan_prior(nis_ctx* ctx, const nis_frame* query, int query_frame_id, double query_acc_distance, const nis_loop_config* cfg,
                        double prior_x, double prior_y, double grid_scale, nis_loop_result* out, int32_t* candidates_out,
                        int max_candidates, int* n_candidates_out) {
  if (!ctx || !query || !cfg || !out || !(grid_scale > 0)) return NIS_ERR_INVALID_ARGUMENT;
  const double prior[2] = {prior_x, prior_y};
  return loop_scan_impl(ctx, query, query_frame_id, query_acc_distance, cfg, nullptr, 0, prior, grid_scale, out, nullptr, nullptr, candidates_out,
                        max_candidates, n_candidates_out);
}

int nis_loop_reduce(const nis_loop_result* per_rank, const int64_t* order, int n_ranks, const nis_loop_config* cfg,
                    nis_loop_result* out, int* winner_rank) {
  if (!per_rank || !cfg || !out || n_ranks <= 0) return NIS_ERR_INVALID_ARGUMENT;
  int best = -1;
  double bs = -3.0;
  int evaluated = 0;
  for (int r = 0; r < n_ranks; ++r) {
    evaluated += per_rank[r].evaluated;
    if (per_rank[r].slot < 0) continue;
    const double s = per_rank[r].response[0] + per_rank[r].response[1] + per_rank[r].response[2];
    const bool earlier = best >= 0 && s == bs && (order ? order[r] < order[best] : r < best);
    if (s > bs || earlier) { bs = s; best = r; }
  }
  if (best >= 0) *out = per_rank[best];
  else {
    memset(out, 0, sizeof *out);
    out->slot = -1; out->frame_id = -1;
    out->response[0] = out->response[1] = out->response[2] = -1.0;
    for (int i = 0; i < 4; ++i) out->peak[i] = -1;
  }
  out->evaluated = evaluated;
  out->found = (out->response[0] > cfg->position_response_thr) && (out->response[2] > cfg->angle_response_thr);
  if (winner_rank) *winner_rank = best;
  return NIS_OK;
}

// ---- multi-GPU scan in the library: the keyframe store is sharded by index over the ranks (one context per GPU, one
// process per GPU); a query is ONE call on every rank: ncclBroadcast of the 307 KB u8 query image from the root, local features +
// local scan, ONE ncclAllGather of the per-rank best records, the reference's strict-'>' / first-wins reduction on every rank.
// NCCL is resolved at run time (dlopen), so the library loads and runs single-GPU without it.
struct nis_nccl_id { char internal[128]; };
struct NcclApi {
  void* h = nullptr;
  int (*GetUniqueId)(nis_nccl_id*) = nullptr;
  int (*CommInitRank)(void**, int, nis_nccl_id, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi load_nccl() {
  NcclApi api;
  const char* names[] = {getenv("NIS_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    if (!nm || !*nm) continue;
    api.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (api.h) break;
  }
  if (!api.h) return api;
  api.GetUniqueId = (int (*)(nis_nccl_id*))dlsym(api.h, "ncclGetUniqueId");
  api.CommInitRank = (int (*)(void**, int, nis_nccl_id, int))dlsym(api.h, "ncclCommInitRank");
  api.CommDestroy = (int (*)(void*))dlsym(api.h, "ncclCommDestroy");
  api.Broadcast = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(api.h, "ncclBroadcast");
  api.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(api.h, "ncclAllGather");
  api.GetErrorString = (const char* (*)(int))dlsym(api.h, "ncclGetErrorString");
  if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.Broadcast || !api.AllGather) { dlclose(api.h); api.h = nullptr; }
  return api;
}
static NcclApi* nccl_api() {
  static NcclApi api = load_nccl();          // resolved once, thread-safe (C++11 static initialisation)
  return api.h ? &api : nullptr;
}
static int nccl_fail(nis_ctx* ctx, const char* what, int rc) {
  NcclApi* a = nccl_api();
  char buf[256];
  snprintf(buf, sizeof buf, "%s: %s", what, a && a->GetErrorString ? a->GetErrorString(rc) : "NCCL error");
  return fail(ctx, NIS_ERR_CUDA, buf);
}

int nis_nccl_unique_id(char id_out[128]) {
  NcclApi* a = nccl_api();
  if (!a || !id_out) return NIS_ERR_CUDA;
  nis_nccl_id id;
  if (a->GetUniqueId(&id) != 0) return NIS_ERR_CUDA;
  memcpy(id_out, id.internal, 128);
  return NIS_OK;
}

int nis_comm_init(nis_ctx* ctx, const char id[128], int rank, int n_ranks) {
  if (!ctx || !id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return NIS_ERR_INVALID_ARGUMENT;
  NcclApi* a = nccl_api();
  if (!a) return fail(ctx, NIS_ERR_CUDA, "libnccl.so.2 not found (set NIS_NCCL_LIB)");
  CU(cudaSetDevice(ctx->device));
  nis_comm_destroy(ctx);
  nis_nccl_id uid;
  memcpy(uid.internal, id, 128);
  const int rc = a->CommInitRank(&ctx->nccl_comm, n_ranks, uid, rank);
  if (rc != 0) { ctx->nccl_comm = nullptr; return nccl_fail(ctx, "ncclCommInitRank", rc); }
  ctx->rank = rank; ctx->n_ranks = n_ranks;
  return NIS_OK;
}

int nis_comm_destroy(nis_ctx* ctx) {
  if (!ctx) return NIS_ERR_INVALID_ARGUMENT;
  if (ctx->nccl_comm) { NcclApi* a = nccl_api(); if (a) a->CommDestroy(ctx->nccl_comm); ctx->nccl_comm = nullptr; }
  ctx->rank = 0; ctx->n_ranks = 1;
  return NIS_OK;
}

struct ShardRecord { nis_loop_result r; int64_t order; int64_t offset; };

int nis_loop_scan_sharded(nis_ctx* ctx, const uint8_t* query_image_rowmajor, int root, int query_frame_id, double query_acc_distance,
                          const nis_loop_config* cfg, long long global_slot_offset, nis_loop_result* out, int* winner_rank,
                          nis_loop_result* local_out) {
  if (!ctx || !cfg || !out || root < 0 || root >= ctx->n_ranks) return NIS_ERR_INVALID_ARGUMENT;
  if (ctx->n_ranks > 1 && !ctx->nccl_comm) return fail(ctx, NIS_ERR_INVALID_ARGUMENT, "nis_loop_scan_sharded: call nis_comm_init first");
  if (ctx->rank == root && !query_image_rowmajor) return NIS_ERR_INVALID_ARGUMENT;
  CU(cudaSetDevice(ctx->device));
  NcclApi* a = ctx->n_ranks > 1 ? nccl_api() : nullptr;
  const size_t npx = ctx->sz[0].real;
  nis_frame* q = nullptr;
  TRY(frame_alloc(ctx, true, &q));
  q->has_h = false;
  auto body = [&]() -> int {
    if (ctx->rank == root) CU(cudaMemcpyAsync(q->img_u8, query_image_rowmajor, npx, cudaMemcpyHostToDevice, ctx->stream));
    if (a) { const int rc = a->Broadcast(q->img_u8, q->img_u8, npx, 1 /* ncclUint8 */, root, ctx->nccl_comm, ctx->stream); if (rc) return nccl_fail(ctx, "ncclBroadcast", rc); }
    // every rank recomputes the query's features (cheaper than shipping 3.8 MB of spectra); a raw image is undistorted first when maps are set
    if (ctx->undistort) {
      RESERVE(ctx->sImg, npx);
      CU(cudaMemcpyAsync(ctx->sImg.p, q->img_u8, npx, cudaMemcpyDeviceToDevice, ctx->stream));
      TRY(undistort_batch(ctx, ctx->stream, ctx->sImg.as<uint8_t>(), q->img_u8, 1));
    }
    // From here on a rank that fails must still reach the all-gather, or the others wait in it forever: the local part runs in its own
    // scope, its status travels in the record (order = INT64_MIN), and every rank returns an error if any rank reported one.
    ShardRecord mine;
    memset(&mine, 0, sizeof mine);
    auto local_part = [&]() -> int {
      TRY(features_batch(ctx, ctx->lanes[0], src_null<float>(), src_slab<uint8_t>(q->img_u8, 0), true, 1, Dst<cpx>{q->F, 0}, Dst<cpx>{q->P, 0},
                         Dst<cpx>{q->Ht, 0}, Dst<cpx>{q->Hp, 0}, false));
      TRY(nis_loop_scan(ctx, q, query_frame_id, query_acc_distance, cfg, nullptr, 0, &mine.r, nullptr));
      return NIS_OK;
    };
    const int local_status = local_part();
    if (local_status == NIS_OK && local_out) *local_out = mine.r;
    mine.offset = global_slot_offset;
    mine.order = local_status != NIS_OK ? INT64_MIN : (mine.r.slot >= 0 ? global_slot_offset + mine.r.slot : INT64_MAX);
    const int G = ctx->n_ranks;
    std::vector<ShardRecord> all(G);
    if (a) {
      RESERVE(ctx->qgather, (size_t)(G + 1) * sizeof(ShardRecord));
      ShardRecord* d = ctx->qgather.as<ShardRecord>();
      CU(cudaMemcpyAsync(d + G, &mine, sizeof mine, cudaMemcpyHostToDevice, ctx->stream));
      const int rc = a->AllGather(d + G, d, sizeof(ShardRecord), 0 /* ncclChar */, ctx->nccl_comm, ctx->stream);
      if (rc) return nccl_fail(ctx, "ncclAllGather", rc);
      CU(cudaMemcpyAsync(all.data(), d, (size_t)G * sizeof(ShardRecord), cudaMemcpyDeviceToHost, ctx->stream));
      CU(cudaStreamSynchronize(ctx->stream));
    } else {
      all[0] = mine;
    }
    if (local_status != NIS_OK) return local_status;                 // this rank's own error (message already recorded)
    for (int i = 0; i < G; ++i)
      if (all[i].order == INT64_MIN) {
        char buf[96];
        snprintf(buf, sizeof buf, "nis_loop_scan_sharded: rank %d failed its local scan", i);
        return fail(ctx, NIS_ERR_CUDA, buf);
      }
    std::vector<nis_loop_result> rs(G);
    std::vector<int64_t> order(G);
    for (int i = 0; i < G; ++i) { rs[i] = all[i].r; order[i] = all[i].order; }
    int win = -1;
    TRY(nis_loop_reduce(rs.data(), order.data(), G, cfg, out, &win));
    if (win >= 0 && out->slot >= 0) out->slot = (int32_t)(all[win].offset + out->slot);       // global slot of the winner
    if (winner_rank) *winner_rank = win;
    return NIS_OK;
  };
  const int st = body();
  nis_frame_free(ctx, q);
  return st;
}

// ---- per-kernel-family timing (CUDA events around every launch; used by bench.py for the roofline line) ---------
int nis_profile_begin(nis_ctx* ctx) {
  if (!ctx) return NIS_ERR_INVALID_ARGUMENT;
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  for (auto& e : ctx->prof) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
  ctx->prof.clear();
  ctx->profiling = true;
  return NIS_OK;
}

int nis_profile_end(nis_ctx* ctx, char* json_out, int json_cap) {
  if (!ctx || !json_out || json_cap <= 0) return NIS_ERR_INVALID_ARGUMENT;
  CU(cudaSetDevice(ctx->device));
  ctx->profiling = false;
  CU(cudaStreamSynchronize(ctx->stream));
  std::map<std::string, std::pair<long long, double>> acc;      // family -> (launches, total ms)
  for (auto& e : ctx->prof) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e.a, e.b);
    std::string name(e.name);
    const size_t par = name.find('(');
    if (par != std::string::npos) name = name.substr(0, par);
    if (name.rfind("launch_", 0) == 0) name = name.substr(7);
    auto& a = acc[name];
    a.first++; a.second += ms;
    cudaEventDestroy(e.a); cudaEventDestroy(e.b);
  }
  ctx->prof.clear();
  std::string js = "{";
  bool first = true;
  for (auto& kv : acc) {
    char buf[256];
    snprintf(buf, sizeof buf, "%s\"%s\": {\"launches\": %lld, \"ms\": %.6f}", first ? "" : ", ", kv.first.c_str(), kv.second.first, kv.second.second);
    js += buf; first = false;
  }
  js += "}";
  if ((int)js.size() + 1 > json_cap) return fail(ctx, NIS_ERR_INVALID_ARGUMENT, "profile json buffer too small");
  memcpy(json_out, js.c_str(), js.size() + 1);
  return NIS_OK;
}

// ---- stage-level debug entry points (tests) ---------------------------------------------------------------------
int nis_debug_fft2(nis_ctx* ctx, int which, const float* real_in, float* spec_out) {
  if (!ctx || !real_in || !spec_out || which < 0 || which > 1) return NIS_ERR_INVALID_ARGUMENT;
  CU(cudaSetDevice(ctx->device));
  Lane& L = ctx->lanes[0];
  TRY(ensure_workspace(ctx, L, 1));
  ctx->prof_stream = L.stream;
  const SizeClass& z = ctx->sz[which];
  DevBuf out;
  RESERVE(out, z.spec * sizeof(cpx));
  CU(h2d(ctx, L.real.p, real_in, z.real * sizeof(float)));
  Dst<cpx> t1{L.t1.as<cpx>(), 0};
  LAUNCH(launch_col_fwd_f32(z.R, z.col, ProRealF32{src_slab<float>(L.real.as<float>(), 0), z.C}, t1, z.C, 1, L.stream));
  LAUNCH(launch_row_fwd(z.C, z.row, ProSpec{src_slab<cpx>(t1.base, 0)}, EpiSpecStore{Dst<cpx>{out.as<cpx>(), 0}}, z.R / 2 + 1, 1, L.stream));
  CU(cudaStreamSynchronize(ctx->stream));
  CU(cudaMemcpy(spec_out, out.p, z.spec * sizeof(cpx), cudaMemcpyDeviceToHost));
  out.release();
  return NIS_OK;
}

int nis_debug_ifft2(nis_ctx* ctx, int which, const float* spec_in, float* real_out) {
  if (!ctx || !spec_in || !real_out || which < 0 || which > 1) return NIS_ERR_INVALID_ARGUMENT;
  CU(cudaSetDevice(ctx->device));
  Lane& L = ctx->lanes[0];
  TRY(ensure_workspace(ctx, L, 1));
  ctx->prof_stream = L.stream;
  const SizeClass& z = ctx->sz[which];
  DevBuf in, one;
  RESERVE(in, z.spec * sizeof(cpx));
  RESERVE(one, z.spec * sizeof(cpx));
  CU(h2d(ctx, in.p, spec_in, z.spec * sizeof(cpx)));
  {
    // the inverse row pass has no plain-load instantiation: multiply by conj(1) through ProMulConj instead
    std::vector<cpx> ones(z.spec, make_float2(1.f, 0.f));
    CU(h2d(ctx, one.p, ones.data(), z.spec * sizeof(cpx)));
  }
  Dst<cpx> t1{L.t1.as<cpx>(), 0};
  LAUNCH(launch_row_inv_mulconj(z.C, z.row, ProMulConj{src_slab<cpx>(in.as<cpx>(), 0), src_slab<cpx>(one.as<cpx>(), 0)}, EpiSpecStore{t1},
                                z.R / 2 + 1, 1, L.stream));
  LAUNCH(launch_col_inv_store(z.R, z.col, src_slab<cpx>(t1.base, 0), EpiStore{Dst<float>{L.real.as<float>(), 0}, z.C, (float)z.real}, z.C, 1,
                              L.stream));
  CU(cudaStreamSynchronize(ctx->stream));
  CU(cudaMemcpy(real_out, L.real.p, z.real * sizeof(float), cudaMemcpyDeviceToHost));
  in.release(); one.release();
  return NIS_OK;
}

int nis_debug_polar(nis_ctx* ctx, const float* power_in, float* polar_out) {
  if (!ctx || !power_in || !polar_out) return NIS_ERR_INVALID_ARGUMENT;
  CU(cudaSetDevice(ctx->device));
  Lane& L = ctx->lanes[0];
  TRY(ensure_workspace(ctx, L, 1));
  ctx->prof_stream = L.stream;
  DevBuf out;
  RESERVE(out, ctx->sz[1].real * sizeof(float));
  {
    // the production path receives power already fftshift-ed from the inverse column pass (EpiStoreShift): shift here on the host
    const int H = ctx->H, W = ctx->W;
    std::vector<float> sh((size_t)H * W);
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) sh[(size_t)y * W + x] = power_in[(size_t)((y - H / 2 + H) % H) * W + (x - W / 2 + W) % W];
    CU(h2d(ctx, L.real.p, sh.data(), sh.size() * sizeof(float)));
  }
  LAUNCH(launch_rzc_fix(Dst<float>{L.real.as<float>(), 0}, ctx->H, ctx->W, 1, L.stream));
  LAUNCH(launch_polar_tma(&L.hp_map, Dst<float>{out.as<float>(), 0}, ctx->D, ctx->Cp, ctx->ptiles.as<int4>(), ctx->ptab2.as<uint32_t>(), ctx->ptile_pitch,
                          ctx->ptile_rows, 1, L.stream));
  CU(cudaStreamSynchronize(ctx->stream));
  CU(cudaMemcpy(polar_out, out.p, ctx->sz[1].real * sizeof(float), cudaMemcpyDeviceToHost));
  out.release();
  return NIS_OK;
}

int nis_debug_rotate(nis_ctx* ctx, const float* image_in, float degree, float* image_out) {
  if (!ctx || !image_in || !image_out) return NIS_ERR_INVALID_ARGUMENT;
  CU(cudaSetDevice(ctx->device));
  Lane& L = ctx->lanes[0];
  TRY(ensure_workspace(ctx, L, 1));
  ctx->prof_stream = L.stream;
  double M[6];
  rotation_inverse(ctx->H, ctx->W, (double)degree, M);
  const int slot = 3 * ctx->D;
  DevBuf out;
  RESERVE(out, ctx->sz[0].real * sizeof(float));
  CU(h2d(ctx, ctx->mats.as<double>() + 6 * (size_t)slot, M, sizeof M));
  CU(h2d(ctx, L.sel.p, &slot, sizeof(int)));
  CU(h2d(ctx, L.real.p, image_in, ctx->sz[0].real * sizeof(float)));
  LAUNCH(launch_rotate(src_slab<float>(L.real.as<float>(), 0), src_null<uint8_t>(), ctx->lut.as<float>(), Dst<float>{out.as<float>(), 0},
                       ctx->H, ctx->W, ctx->mats.as<double>(), L.sel.as<int>(), 1, L.stream));
  CU(cudaStreamSynchronize(ctx->stream));
  CU(cudaMemcpy(image_out, out.p, ctx->sz[0].real * sizeof(float), cudaMemcpyDeviceToHost));
  out.release();
  return NIS_OK;
}

int nis_debug_estimate_trans(nis_ctx* ctx, int which, const float* last_spec, const float* cur_spec, int32_t peak_rc[2], float* info,
                             float* g_out) {
  if (!ctx || !last_spec || !cur_spec || !peak_rc || !info || which < 0 || which > 1) return NIS_ERR_INVALID_ARGUMENT;
  CU(cudaSetDevice(ctx->device));
  Lane& L = ctx->lanes[0];
  TRY(ensure_workspace(ctx, L, 1));
  const SizeClass& z = ctx->sz[which];
  DevBuf dz, dx, dh, dg;
  RESERVE(dz, z.spec * sizeof(cpx)); RESERVE(dx, z.spec * sizeof(cpx)); RESERVE(dh, z.spec * sizeof(cpx));
  if (g_out) RESERVE(dg, z.real * sizeof(float));
  int st = NIS_OK;
  if (h2d(ctx, dz.p, last_spec, z.spec * sizeof(cpx)) != cudaSuccess || h2d(ctx, dx.p, cur_spec, z.spec * sizeof(cpx)) != cudaSuccess)
    st = fail(ctx, NIS_ERR_CUDA, "debug copy", (int)cudaGetLastError());
  PeakStats ps;
  if (st == NIS_OK) st = hzz_batch(ctx, L, which, src_slab<cpx>(dz.as<cpx>(), 0), 1, Dst<cpx>{dh.as<cpx>(), 0});
  if (st == NIS_OK) st = estimate_trans_stored(ctx, L, which, src_slab<cpx>(dz.as<cpx>(), 0), src_slab<cpx>(dh.as<cpx>(), 0),
                                               src_slab<cpx>(dx.as<cpx>(), 0), 1, L.maxt.as<unsigned>(), L.stats_t.as<PeakStats>(),
                                               g_out ? dg.as<float>() : nullptr);
  if (st == NIS_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) st = fail(ctx, NIS_ERR_CUDA, "debug sync", (int)cudaGetLastError());
  if (st == NIS_OK && cudaMemcpy(&ps, L.stats_t.p, sizeof ps, cudaMemcpyDeviceToHost) != cudaSuccess) st = fail(ctx, NIS_ERR_CUDA, "debug copy back");
  if (st == NIS_OK && g_out && cudaMemcpy(g_out, dg.p, z.real * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess) st = fail(ctx, NIS_ERR_CUDA, "debug g copy");
  dz.release(); dx.release(); dh.release(); dg.release();
  if (st != NIS_OK) return st;
  const uint32_t idx = 0xffffffffu - (uint32_t)(ps.key & 0xffffffffull);
  const int col = (int)(idx / (uint32_t)z.R), row = (int)(idx % (uint32_t)z.R);
  const float peak = ord2f((uint32_t)(ps.key >> 32));
  const double n = (double)z.real;
  const float m = ((float)ps.sum - peak) / (float)(n - 1.0);
  double var = (ps.sumsq - 2.0 * (double)m * ps.sum + n * (double)m * (double)m) / n;
  var = var > 0 ? var : 0;
  *info = (float)((double)(peak - m) / ((double)sqrtf((float)var) + 1e-7));
  peak_rc[0] = row; peak_rc[1] = col;
  return NIS_OK;
}

}  // extern "C"
