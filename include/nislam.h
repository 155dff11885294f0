/*
 * nislam.h -- C ABI of the B200-native NI-SLAM tracking / loop-closure hot path (libnislam.so).
 *
 * The reference (sair-lab/ni-slam) has no FFI or plugin interface for this path: it sits behind two plain C++
 * classes.  Each entry point below names the reference interface it replaces (file:line under the reference
 * tree); ni_slam_b200/host/correlation_flow.hpp re-exposes the reference's class signatures on top of them and
 * INTEGRATION.md shows the binding a maintainer adds to MapBuilder.
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns an int status (NIS_OK = 0); no exceptions cross the ABI.
 *   - "reference layout" = Eigen column-major: a real R x C array is C lines of R floats; a half spectrum
 *     ArrayXXcf((R/2+1), C) is C lines of (R/2+1) interleaved (re,im) float pairs.
 *   - u8 images are cv::Mat-like row-major H x W.
 *   - one context per device; calls on one context are serialised on its CUDA stream; a context is not re-entrant
 *     (the reference's CorrelationFlow is not re-entrant either: FFTW planner calls, correlation_flow.cc:56-61).
 *   - there is no CPU fallback: every compute entry point fails with NIS_ERR_CUDA when no device is usable.
 */
#ifndef NISLAM_H_
#define NISLAM_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  NIS_OK = 0,
  NIS_ERR_INVALID_ARGUMENT = 1, /* NULL pointers, odd sizes ... */
  NIS_ERR_INVALID_KERNEL = 2,   /* cfg.kernel not in {0,1}: std::invalid_argument("Received invalid kernel type"), correlation_flow.cc:168 */
  NIS_ERR_UNSUPPORTED_SIZE = 3, /* transform length without an instantiated kernel (see csrc/nis_sizes.h) */
  NIS_ERR_CUDA = 4,             /* CUDA runtime error; text in nis_last_error() */
  NIS_ERR_OUT_OF_MEMORY = 5
};

typedef struct nis_ctx nis_ctx;     /* = CorrelationFlow + LoopClosure state on one GPU */
typedef struct nis_frame nis_frame; /* device-resident Frame payload: image + fft_result + fft_polar (include/frame.h:34-36) */

/* CFConfig (include/read_configs.h:15-25) minus width/height, which the ctor overwrites with the camera size
 * (correlation_flow.cc:40-41). */
typedef struct {
  float lambda;
  int kernel; /* 0 polynomial, 1 gaussian */
  float sigma;
  float offset;
  int power;
  int rotation_divisor;
  int rotation_channel;
} nis_cf_config;

/* LoopClosureConfig (include/read_configs.h:38-44) minus to_find_loop (a caller decision, map_builder.cc:63). */
typedef struct {
  double position_response_thr;
  double angle_response_thr;
  int frame_gap_thr;
  double distance_thr;
} nis_loop_config;

/* LoopClosureResult (include/loop_closure.h:8-25) with the FramePtr replaced by slot / frame id, plus the raw peaks. */
typedef struct {
  int32_t found;
  int32_t slot;       /* DB slot of the winner, -1 if no candidate was evaluated */
  int32_t frame_id;   /* its frame id */
  int32_t hyp;        /* 0: "-deg" hypothesis kept, 1: "-deg+180" (correlation_flow.cc:121-131) */
  double relative_pose[3];
  double response[3]; /* (-1,-1,-1) when nothing was evaluated (loop_closure.h:15) */
  int32_t peak[4];    /* polar row, polar col, translation row, translation col */
  int32_t evaluated;  /* candidates that passed the gap / distance filters */
} nis_loop_result;

/* ---- lifecycle: CorrelationFlow::CorrelationFlow(CFConfig&, double& h, double& w), include/correlation_flow.h:11 ---- */
int nis_create(const nis_cf_config* cfg, int image_height, int image_width, int device, nis_ctx** out);
int nis_destroy(nis_ctx* ctx);
const char* nis_last_error(const nis_ctx* ctx); /* never NULL */
const char* nis_strerror(int status);
/* the CUDA stream all of this context's kernels are launched on (cudaStream_t), for event timing by the caller */
void* nis_stream(nis_ctx* ctx);
int nis_synchronize(nis_ctx* ctx);
/* number of this library's kernels launched on the context so far */
long long nis_kernel_launches(const nis_ctx* ctx);
/* work-batch size (pairs / candidates in flight per kernel launch); 0 = library default */
int nis_set_batch(nis_ctx* ctx, int batch);
/* number of concurrent CUDA streams ("lanes") the batches are dealt to; all fork from / join into nis_stream(); 0 = default */
int nis_set_lanes(nis_ctx* ctx, int lanes);

/* ---- features: MapBuilder::ComputeFFTResult = ConvertMatToNormalizedArray + CorrelationFlow::ComputeIntermedium
 *      (src/map_builder.cc:72-75, src/utils.cc:110-118, include/correlation_flow.h:12) ---- */
int nis_features_u8(nis_ctx* ctx, const uint8_t* image_rowmajor, nis_frame** out);   /* cv::Mat u8 H x W */
int nis_features_f32(nis_ctx* ctx, const float* image_colmajor, nis_frame** out);    /* ArrayXXf H x W, reference layout */
/* Frame::GetFFTResult (src/frame.cc:53-57): copy both spectra out in reference layout; either pointer may be NULL */
int nis_frame_export(nis_ctx* ctx, const nis_frame* f, float* fft_result, float* fft_polar);
/* build a frame from reference-layout arrays the caller already holds (image, fft_result, fft_polar); also computes the frame's
 * keyframe factors H, so it can serve as `last` of nis_compute_pose or go into the store */
int nis_frame_import(nis_ctx* ctx, const float* image_colmajor, const float* fft_result, const float* fft_polar, nis_frame** out);
/* the same with any subset of the three arrays (NULL = absent) and the H factors optional: what CorrelationFlow::ComputePose actually
 * reads is (last_fft_result, last_fft_polar [+ H]) of the last frame and (image, fft_polar) of the current one (include/correlation_flow.h:13),
 * what FindLoopClosure reads of the query is (image, fft_polar) (src/loop_closure.cc:38-39, :58-59) */
int nis_frame_import_ex(nis_ctx* ctx, const float* image_colmajor, const float* fft_result, const float* fft_polar, int with_h, nis_frame** out);
int nis_frame_free(nis_ctx* ctx, nis_frame* f);

/* ---- undistort front end: Camera::UndistortImage (src/camera.cc:92-93) = cv::remap(u8, _map1, _map2, INTER_LINEAR).
 *      The caller's (unchanged) Camera computes the fixed-point maps once with initUndistortRectifyMap(..., CV_16SC2, ...)
 *      (src/camera.cc:45-47) and hands them over: map1_xy = H x W x 2 int16 (x, y), map2 = H x W uint16 (fy*32 + fx).
 *      While maps are set, every u8 image entering nis_features_u8 / nis_track_stream[_dev] / nis_db_add_images[_dev] is taken
 *      as the RAW camera image and undistorted on the GPU first, exactly like MapBuilder::AddNewInput (src/map_builder.cc:31-33).
 *      Pass NULL, NULL to switch the front end off.  nis_undistort_u8 runs the stage alone (host buffers). ---- */
int nis_set_undistort_maps(nis_ctx* ctx, const int16_t* map1_xy, const uint16_t* map2);
int nis_undistort_u8(nis_ctx* ctx, const uint8_t* raw_rowmajor, uint8_t* out_rowmajor);

/* ---- solve: CorrelationFlow::ComputePose(last_fft_result, image, last_fft_polar, fft_polar, pose, not_large_rotation)
 *      include/correlation_flow.h:13, src/correlation_flow.cc:97-143; returns info in info[3], pose = (dx_px, dy_px, theta_rad).
 *      peak_rc (may be NULL) = polar row, polar col, translation row, translation col of the integer arg-max. ---- */
int nis_compute_pose(nis_ctx* ctx, const nis_frame* last, const nis_frame* cur, int not_large_rotation, double pose[3],
                     double info[3], int32_t peak_rc[4]);

/* ---- batched tracking of a stream with "every frame is a keyframe" (SURVEY 8d): solve i pairs frame i (last) with
 *      frame i+1 (current), i.e. per frame ComputeIntermedium + ComputePose(..., true) exactly as
 *      MapBuilder::AddNewInput does (src/map_builder.cc:30-70).  frames: n u8 images, host (nis_track_stream) or
 *      device (nis_track_stream_dev) memory.  poses/infos: host, (n-1) x 3 doubles. ---- */
int nis_track_stream(nis_ctx* ctx, const uint8_t* frames_host, int n, double* poses, double* infos);
int nis_track_stream_dev(nis_ctx* ctx, const uint8_t* frames_dev, int n, double* poses, double* infos);

/* ---- tracking of a stream under the reference's keyframe policy: MapBuilder::AddNewInput without loop closure, optimisation
 *      and stitching (src/map_builder.cc:30-70): Tracking against the LAST KEYFRAME (:127-138), the confidence gate, pose
 *      composition (UpdateCurrentPose :118-125, ComputeAbsolutePose / ComputeRelativePose src/utils.cc:133-152,
 *      Camera::ConvertCenterToPrincipal / ConvertImagePlanePoseToCamera / ConvertCameraPoseToRobot src/camera.cc:148-218),
 *      ComputeRelativeDA (:157-166) and the keyframe test c1..c4 (:47-53).  Frames after a keyframe are solved against it
 *      speculatively in batches; the solves behind the first frame that becomes a keyframe are redone against the new one, so the
 *      results are exactly those of the frame-by-frame loop. ---- */
typedef struct {           /* KeyframeSelectionConfig, include/read_configs.h:27-32 */
  double max_distance;
  double max_angle;
  double lower_response_thr;
  double upper_response_thr;
} nis_kfs_config;
typedef struct {           /* the parts of Camera the pose conversions read (src/camera.cc:148-218) */
  double fx, fy, cx, cy;   /* _new_K(0,0), (1,1), (0,2), (1,2) */
  double height;           /* _height */
  double extrinsics[9];    /* _extrinsics, row-major 3 x 3 */
} nis_camera_model;
typedef struct {
  int32_t tracked;         /* good_tracking (map_builder.cc:132); 1 for the first frame */
  int32_t inserted;        /* return value of AddNewInput: the frame became a keyframe */
  int32_t keyframe;        /* index of the keyframe this frame was solved against (-1 for the first frame) */
  int32_t reserved;
  double response[3];      /* ComputePose return value */
  double relative_pose[3]; /* ComputePose pose after ConvertCenterToPrincipal (map_builder.cc:131) */
  double cf_pose[3];       /* _current_cf_pose after this frame */
  double pose[3];          /* _current_pose (robot frame) after this frame */
  double distance;         /* _distance after this frame */
} nis_track_result;
int nis_track_stream_keyframes(nis_ctx* ctx, const uint8_t* frames_host, int n, const nis_kfs_config* kfs,
                               const nis_camera_model* cam, nis_track_result* out /* n records */);

/* ---- MapStitcher (src/map_stitcher.cc, include/map_stitcher.h): the occupancy mosaic of keyframe images.  InsertFrame (:14-22:
 *      image * 100/255 as u8, kept for later) + AddImageToOccupancy (:36-133: every pixel lands on the integer ground position
 *      (int)(R (i - W/2, j - H/2) + t), truncated toward zero; per-frame sums and counts are merged into the cells with the
 *      reference's integer rules, including storing the raw per-frame SUM when a cell is first created) and RecomputeOccupancy
 *      (:135-145) after a pose-graph optimisation.  The reference keeps an unbounded hash of cell_size x cell_size cells; here the
 *      cells live in a dense window [cell_x0, cell_x0 + cells_x) x [cell_y0, cell_y0 + cells_y) in HBM, pixels outside it are counted
 *      (nis_stitcher_dropped) and otherwise ignored.  Frames are replayed in insertion order by nis_stitcher_recompute (the reference
 *      iterates an unordered_map keyed by pointer, i.e. in no defined order, although its merge rule is order dependent). ---- */
typedef struct nis_stitcher nis_stitcher;
int nis_stitcher_create(int device, int image_height, int image_width, int cell_size, int cell_x0, int cell_y0, int cells_x, int cells_y,
                        nis_stitcher** out);
int nis_stitcher_destroy(nis_stitcher* st);
/* InsertFrame(frame, image): image = the undistorted u8 frame (host, row-major H x W), robot_pose = Frame::GetPose */
int nis_stitcher_insert(nis_stitcher* st, const uint8_t* image_u8, const double robot_pose[3], const nis_camera_model* cam, int* frame_slot);
/* RecomputeOccupancy with the poses after optimisation: robot_poses = frames x 3 doubles in insertion order */
int nis_stitcher_recompute(nis_stitcher* st, const double* robot_poses, const nis_camera_model* cam);
int nis_stitcher_frames(const nis_stitcher* st);
/* one cell of GetOccupancyData(): data / weight = cell_size x cell_size int32, [in-cell y][in-cell x]; *present = 0 when the cell was
 * never created (buffers untouched) */
int nis_stitcher_cell(nis_stitcher* st, int cell_x, int cell_y, int32_t* data, int32_t* weight, int* present);
int nis_stitcher_dropped(nis_stitcher* st, long long* pixels_outside_window);

/* ---- keyframe database = Map::AddFrame for the arrays the scan reads (include/frame.h:35-36, src/map.cc) ----
 * Store mode (set while the store is empty): what a record keeps per keyframe.  A scan returns the same bits in every mode -- what is
 * not kept is recomputed per batch of candidates by the same kernels.
 *   NIS_DB_FULL    fft_result, fft_polar and the cached factors Ht, Hp     5.24 MB @640x480   (default; fastest scan)
 *   NIS_DB_SPECTRA fft_result, fft_polar = Frame::_fft_result/_fft_polar   2.62 MB            (the reference's own payload)
 *   NIS_DB_IMAGE   the u8 image                                            0.31 MB            (100 k keyframes on one GPU) */
enum { NIS_DB_FULL = 0, NIS_DB_SPECTRA = 1, NIS_DB_IMAGE = 2 };
int nis_db_set_mode(nis_ctx* ctx, int mode);
int nis_db_mode(const nis_ctx* ctx);
int nis_db_add(nis_ctx* ctx, const nis_frame* f, int frame_id, double acc_distance, int* slot);
/* a keyframe given as the reference-layout arrays Frame::GetFFTResult hands out (src/frame.cc:53-57); not for NIS_DB_IMAGE */
int nis_db_add_spectra(nis_ctx* ctx, const float* fft_result, const float* fft_polar, int frame_id, double acc_distance, int* slot);
/* bulk insert: n u8 images (host or device), features computed on the GPU straight into the DB */
int nis_db_add_images(nis_ctx* ctx, const uint8_t* images_host, int n, const int* frame_ids, const double* acc_distances);
int nis_db_add_images_dev(nis_ctx* ctx, const uint8_t* images_dev, int n, const int* frame_ids, const double* acc_distances);
int nis_db_size(const nis_ctx* ctx);
int nis_db_clear(nis_ctx* ctx);

/* ---- scan: LoopClosure::FindLoopClosure(image, current_frame, frames), src/loop_closure.cc:36-73.
 *      candidate_slots == NULL scans all slots in insertion order (the all-frames overload, loop_closure.cc:10-15);
 *      otherwise the given order is the iteration order (ties: first wins, strict '>').
 *      all_responses (may be NULL): n_candidates x 3 doubles, rows of skipped candidates are (-1,-1,-1). ---- */
int nis_loop_scan(nis_ctx* ctx, const nis_frame* query, int query_frame_id, double query_acc_distance,
                  const nis_loop_config* cfg, const int32_t* candidate_slots, int n_candidates, nis_loop_result* out,
                  double* all_responses);

/* the same scan with one record per entry of the candidate list (all slots when candidate_slots == NULL): records[i].evaluated = 0 for
 * candidates the gap / distance filters dropped (loop_closure.cc:43-53) */
typedef struct {
  int32_t evaluated;
  int32_t hyp;
  double relative_pose[3];
  double response[3];
  int32_t peak[4];
} nis_scan_record;
int nis_loop_scan_records(nis_ctx* ctx, const nis_frame* query, int query_frame_id, double query_acc_distance,
                          const nis_loop_config* cfg, const int32_t* candidate_slots, int n_candidates, nis_loop_result* out,
                          nis_scan_record* records);

/* ---- candidate selection by prior pose: LoopClosure::FindLoopClosure(image, current_frame, prior_pose), src/loop_closure.cc:17-34
 *      = Map::ComputeGridLocation + the 3x3 neighbourhood + Map::GetFramesInGrids (src/map.cc:81-101).
 *      nis_db_set_position files a keyframe under its grid cell ((int)(x/grid_scale), (int)(y/grid_scale)) -- like Map::AddFrame
 *      it is the pose at insertion time that counts.  nis_loop_scan_prior scans the keyframes of the 9 cells around the prior's cell,
 *      cells in the reference's order (dx = -1..1 outer, dy = -1..1 inner), slots ascending inside a cell (the reference iterates
 *      an unordered_set there, i.e. in no defined order).  candidates_out (may be NULL, capacity max_candidates) receives that list.
 *      Selection (grid cells and the gap / distance filters of nis_loop_scan alike) runs on the GPU over device-resident per-slot
 *      frame ids, distances and cells: counts per chunk, one exclusive scan, ordered compaction -- no host loop over the store. ---- */
int nis_db_set_position(nis_ctx* ctx, int slot, double x, double y, double grid_scale);
int nis_loop_scan_prior(nis_ctx* ctx, const nis_frame* query, int query_frame_id, double query_acc_distance,
                        const nis_loop_config* cfg, double prior_x, double prior_y, double grid_scale, nis_loop_result* out,
                        int32_t* candidates_out, int max_candidates, int* n_candidates_out);

/* ---- multi-GPU scan: the keyframe store is sharded by index, one context (= one process) per GPU.
 *      nis_nccl_unique_id on one rank -> ship the 128 bytes to the others by any means -> nis_comm_init on every rank; then one
 *      nis_loop_scan_sharded call per query on EVERY rank: the root's u8 query image (row-major H x W, raw when undistort maps are
 *      set) is broadcast (ncclBroadcast, 307 KB), every rank computes the query's features and scans its own shard, the per-rank
 *      best records are all-gathered (ONE ncclAllGather of 104 bytes per rank) and reduced with the reference's rule (strict '>',
 *      earliest global slot on ties, loop_closure.cc:61).  global_slot_offset = global index of this rank's slot 0; out->slot is the
 *      winner's GLOBAL slot, out->evaluated the total over all ranks; local_out (may be NULL) this rank's own best.
 *      NCCL is resolved with dlopen("libnccl.so.2") (override: NIS_NCCL_LIB) on the first call; single-GPU use never needs it. ---- */
int nis_nccl_unique_id(char id_out[128]);
int nis_comm_init(nis_ctx* ctx, const char id[128], int rank, int n_ranks);
int nis_comm_destroy(nis_ctx* ctx);
int nis_loop_scan_sharded(nis_ctx* ctx, const uint8_t* query_image_rowmajor, int root, int query_frame_id, double query_acc_distance,
                          const nis_loop_config* cfg, long long global_slot_offset, nis_loop_result* out, int* winner_rank,
                          nis_loop_result* local_out);
/* The reduction nis_loop_scan_sharded applies, exposed for hosts that move the records themselves: `order[i]` is the position of
 * rank i's winner in the global iteration order (ties: smallest wins); pass NULL to use the rank index. */
int nis_loop_reduce(const nis_loop_result* per_rank, const int64_t* order, int n_ranks, const nis_loop_config* cfg,
                    nis_loop_result* out, int* winner_rank);

/* ---- per-kernel-family timing: CUDA events around every launch between begin and end; `json_out` receives
 *      {"family": {"launches": n, "ms": total}, ...}.  Measurement aid for bench.py (not on the reference surface). ---- */
int nis_profile_begin(nis_ctx* ctx);
int nis_profile_end(nis_ctx* ctx, char* json_out, int json_cap);

/* ---- stage-level entry points used by the parity tests (host buffers, natural row-major numpy layout:
 *      real (R, C), spectrum (R/2+1, C) complex64).  which: 0 = image size H x W, 1 = polar size D x Cp. ---- */
int nis_debug_fft2(nis_ctx* ctx, int which, const float* real_in, float* spec_out);
int nis_debug_ifft2(nis_ctx* ctx, int which, const float* spec_in, float* real_out);
int nis_debug_polar(nis_ctx* ctx, const float* power_in, float* polar_out);         /* fftshift(RemoveZero(power)) -> warpPolar */
int nis_debug_rotate(nis_ctx* ctx, const float* image_in, float degree, float* image_out);
int nis_debug_estimate_trans(nis_ctx* ctx, int which, const float* last_spec, const float* cur_spec, int32_t peak_rc[2],
                             float* info, float* g_out /* may be NULL */);

#ifdef __cplusplus
}
#endif
#endif /* NISLAM_H_ */
