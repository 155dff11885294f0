// nis_internal.h -- launcher declarations shared by the .cu translation units (not part of the C-ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "nis_ops.cuh"

namespace nis {

// per-candidate / per-pair result written by pose_finalize (device and host POD)
struct PoseRecord {
  double pose[3];      // (dx_px, dy_px, theta_rad)          correlation_flow.cc:136-138
  double info[3];      // (info_trans, info_trans, info_rot)
  int32_t peak[4];     // polar row, polar col, trans row, trans col
  int32_t hyp;         // loop mode: 0 = "-deg" hypothesis kept, 1 = "-deg+180"
  int32_t index;       // candidate index in the scan order (or pair index)
};

// host-built per-polar-row tables (see build_tables in nis_api.cu); all indexed by the polar peak row in [0,D)
struct AngleTables {
  const double* rot_mats;   // [3][D][6]  inverse affine matrices: 0 tracking(-deg folded), 1 loop "-deg", 2 loop "-deg+180"
  const double* theta;      // [3][D]     final theta (float-rounded) per variant
};

bool col_size_supported(int N);
bool row_size_supported(int N);
void plan_radices_col(int N, int r[3]);
void plan_radices_row(int N, int r[3]);
bool plan_radices_row_b(int N, int r[3]);     // the size's two-stage plan B, if it has one (nis_sizes.h)
// twiddle tables of both row plans of one size; every row launcher picks the plan its kernel family measured faster with
struct RowTwiddles { Twiddles a, b; };

// ---- FFT passes (return cudaError_t as int; -1 = unsupported size) --------------------------------------
// rowtab[slot][y] = (X0, Y0) of cv::warpAffine's fixed-point walk for output row y under rotation-matrix slot `slot` (host-built beside mats)
// sel[e] = rotation-matrix slot of entry e; or, when `polar` is given, derived in the kernel from the polar-stage peak of pair e >> loop
// (slot = polar row for tracking, D + row / 2 D + row for the two loop hypotheses: correlation_flow.cc:105-121) -- no select launch
struct RotateArgs { Src<float> f32; Src<uint8_t> u8; bool is_u8; const float* lut; int H, W; const double* mats; const int* sel; const int2* rowtab;
                    const PeakStats* polar = nullptr; int D = 0; int loop = 0; };
int launch_col_fwd_f32(int N, Twiddles tw, ProRealF32 pro, Dst<cpx> out, int W, int B, cudaStream_t s);
int launch_col_fwd_u8(int N, Twiddles tw, ProRealU8 pro, Dst<cpx> out, int W, int B, cudaStream_t s);
int launch_col_fwd_rotate(int N, Twiddles tw, RotateArgs ra, Dst<cpx> out, int W, int E, cudaStream_t s);   // RotateArray fused into stage 0
int launch_col_inv_store(int N, Twiddles tw, Src<cpx> in, EpiStore epi, int W, int B, cudaStream_t s);
int launch_col_inv_peak(int N, Twiddles tw, Src<cpx> in, EpiPeak epi, int W, int B, cudaStream_t s);
// c2r -> kernel function -> r2c in one kernel (the real kernel image never leaves shared memory); in-place allowed
int launch_colcol(int N, Twiddles tw, Src<cpx> in, Dst<cpx> out, KernelFn fn, int W, int B, cudaStream_t s);
int launch_row_fwd(int N, RowTwiddles tw, ProSpec pro, EpiSpecStore epi, int nrows, int B, cudaStream_t s, bool match_fused = false);
int launch_row_fwd_h(int N, RowTwiddles tw, ProSpec pro, EpiHStore epi, int nrows, int B, cudaStream_t s);
int launch_row_inv_mulconj(int N, RowTwiddles tw, ProMulConj pro, EpiSpecStore epi, int nrows, int B, cudaStream_t s, bool match_fused = false);
// forward -> element-wise -> inverse in one kernel; in-place allowed (a CTA reads and writes only its own lines)
int launch_rowrow_mulconj(int N, RowTwiddles tw, Src<cpx> in, Dst<cpx> out, MidMulConjZ mid, int nrows, int B, cudaStream_t s);
int launch_rowrow_filter(int N, RowTwiddles tw, Src<cpx> in, Dst<cpx> out, MidFilterH mid, int nrows, int B, cudaStream_t s);
int launch_rowrow_storeabs(int N, RowTwiddles tw, Src<cpx> in, Dst<cpx> out, MidStoreAbs mid, int nrows, int B, cudaStream_t s);
int launch_rowrow_storesq(int N, RowTwiddles tw, Src<cpx> in, Dst<cpx> out, MidStoreSq mid, int nrows, int B, cudaStream_t s);

// ---- warps and bookkeeping kernels -----------------------------------------------------------------------
// tiled polar gather (nis_misc.cu): one CTA per cell of kPolarTA angles x kPolarTR radii, source box staged in shared memory by TMA
constexpr int kPolarTA = 16, kPolarTR = 64;
int launch_polar_tile_bbox(int4* tiles, int H, int W, int D, int Cp, const double* cs_table, const float* rho_table, cudaStream_t s);
int launch_polar_tile_table(uint32_t* table, const int4* tiles, int pitch, int H, int W, int D, int Cp, const double* cs_table,
                            const float* rho_table, cudaStream_t s);
// hp_map: CUtensorMap (cuTensorMapEncodeTiled) over the lane's shifted power images [B][H][W] f32 with box {pitch, box_rows, 1}
int launch_polar_tma(const void* hp_map, Dst<float> out, int D, int Cp, const int4* tiles, const uint32_t* table, int pitch, int box_rows, int B,
                     cudaStream_t s);
// RemoveZeroComponent on the fftshift-ed power image, in place
int launch_rzc_fix(Dst<float> hp, int H, int W, int B, cudaStream_t s);
int launch_col_inv_store_shift(int N, Twiddles tw, Src<cpx> in, EpiStoreShift epi, int W, int B, cudaStream_t s);
// rotate: out[e] = warpAffine(image[e], rot_mats[sel[e]]) with BORDER_WRAP.  Exactly one of img_f32 / img_u8 is used.
int launch_rotate(Src<float> img_f32, Src<uint8_t> img_u8, const float* lut, Dst<float> out, int H, int W, const double* mats,
                  const int* sel, int E, cudaStream_t s);
// Camera::UndistortImage: exact u8 cv::remap with the fixed-point maps the caller's Camera computed (CV_16SC2 + CV_16UC1)
int launch_undistort(Src<uint8_t> raw, Dst<uint8_t> out, int H, int W, const void* map1, const void* map2, int B, cudaStream_t s);
// gaussian kernel helper: out[b] = sum over the stored half spectrum of |x^2| (raw, double)
int launch_spec_sqsum(Src<cpx> x, int count, double* out, int B, cudaStream_t s);
// after the polar stage: per pair, pick the rotation-matrix slot(s) for the translation stage
int launch_polar_select(const PeakStats* polar, int D, int loop_mode, int* sel, int B, cudaStream_t s);
// per pair: info (GetInfo), hypothesis choice, pose
int launch_pose_finalize(const PeakStats* polar, const PeakStats* trans, AngleTables tabs, int H, int W, int D, int Cp,
                         int loop_mode, int index0, PoseRecord* out, int B, cudaStream_t s);
// boundary layout conversion (reference column-major <-> row-major): in[rows_in][cols_in] -> out[cols_in][rows_in]
int launch_transpose_f32(const float* in, float* out, int rows_in, int cols_in, cudaStream_t s);
int launch_transpose_cpx(const cpx* in, cpx* out, int rows_in, int cols_in, cudaStream_t s);
// device-side candidate selection (filters + optional 3x3 grid neighbourhood), see nis_misc.cu
struct SelectArgs {
  int n_in; const int* list;                       // list == nullptr: all slots in insertion order
  const int* frame_id; const double* dist; const int2* cell;
  int query_id; double query_dist; int frame_gap_thr; double distance_thr;
  int use_prior, cx, cy;
};
int select_scratch_ints(int n_in);
int launch_select(SelectArgs a, int* scratch, int* cand, int* pos, int* n_out, cudaStream_t s);
// scan: best record by response.sum(), strict '>', first in iteration order wins (loop_closure.cc:61)
int launch_scan_reduce(const PoseRecord* recs, int n, PoseRecord* best, cudaStream_t s);

}  // namespace nis
