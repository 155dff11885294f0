"""Host-side mirror of the reference's CorrelationFlow / LoopClosure call surface over the C ABI (include/nislam.h).

The reference's host code is C++ (ni_slam_b200/host/correlation_flow.hpp is the drop-in shim for MapBuilder); this
module is the same surface for Python callers, tests and bench.py.  Everything computes on the GPU through
libnislam.so -- there is no CPU fallback: importing works anywhere, but creating a CorrelationFlow without the
built library or without a CUDA device raises.

Array conventions at this level are natural numpy: images (H, W); spectra (R/2+1, C) complex64 -- i.e. the same
(row, col) indexing as the reference's Eigen arrays.  The C ABI itself speaks the reference's column-major layout.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# NIS_LIB: alternative build of the same library (A/B measurements of compile-time variants)
LIB_PATH = os.environ.get("NIS_LIB") or os.path.join(_HERE, "lib", "libnislam.so")

NIS_OK, NIS_ERR_INVALID_ARGUMENT, NIS_ERR_INVALID_KERNEL, NIS_ERR_UNSUPPORTED_SIZE, NIS_ERR_CUDA, NIS_ERR_OOM = range(6)

SYMBOLS = [
    "nis_create", "nis_destroy", "nis_last_error", "nis_strerror", "nis_stream", "nis_synchronize", "nis_kernel_launches",
    "nis_set_batch", "nis_set_lanes", "nis_features_u8", "nis_features_f32", "nis_frame_export", "nis_frame_import", "nis_frame_free",
    "nis_set_undistort_maps", "nis_undistort_u8", "nis_compute_pose", "nis_track_stream", "nis_track_stream_dev", "nis_track_stream_keyframes", "nis_db_add", "nis_db_add_images",
    "nis_db_add_images_dev", "nis_db_size", "nis_db_clear", "nis_loop_scan", "nis_loop_reduce", "nis_db_set_position", "nis_loop_scan_prior", "nis_debug_fft2",
    "nis_debug_ifft2", "nis_debug_polar", "nis_debug_rotate", "nis_debug_estimate_trans", "nis_profile_begin",
    "nis_profile_end", "nis_stitcher_create", "nis_stitcher_destroy", "nis_stitcher_insert", "nis_stitcher_recompute",
    "nis_stitcher_frames", "nis_stitcher_cell", "nis_stitcher_dropped", "nis_frame_import_ex", "nis_db_set_mode", "nis_db_mode",
    "nis_db_add_spectra", "nis_loop_scan_records", "nis_nccl_unique_id", "nis_comm_init", "nis_comm_destroy", "nis_loop_scan_sharded",
]

DB_FULL, DB_SPECTRA, DB_IMAGE = 0, 1, 2      # nis_db_set_mode


class NisError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("libnislam status %d: %s" % (status, msg))
        self.status = status


class _CfConfigC(C.Structure):
    _fields_ = [("lam", C.c_float), ("kernel", C.c_int), ("sigma", C.c_float), ("offset", C.c_float), ("power", C.c_int),
                ("rotation_divisor", C.c_int), ("rotation_channel", C.c_int)]


class _LoopConfigC(C.Structure):
    _fields_ = [("position_response_thr", C.c_double), ("angle_response_thr", C.c_double), ("frame_gap_thr", C.c_int),
                ("distance_thr", C.c_double)]


class LoopResultC(C.Structure):
    _fields_ = [("found", C.c_int32), ("slot", C.c_int32), ("frame_id", C.c_int32), ("hyp", C.c_int32),
                ("relative_pose", C.c_double * 3), ("response", C.c_double * 3), ("peak", C.c_int32 * 4),
                ("evaluated", C.c_int32)]


SCAN_RECORD_DTYPE = np.dtype([("evaluated", np.int32), ("hyp", np.int32), ("relative_pose", np.float64, 3), ("response", np.float64, 3),
                              ("peak", np.int32, 4)])          # nis_scan_record


class _KfsConfigC(C.Structure):       # nis_kfs_config = KeyframeSelectionConfig (include/read_configs.h:27-32)
    _fields_ = [("max_distance", C.c_double), ("max_angle", C.c_double), ("lower_response_thr", C.c_double),
                ("upper_response_thr", C.c_double)]


class _CameraModelC(C.Structure):     # nis_camera_model: the parts of Camera the pose conversions read (src/camera.cc:148-218)
    _fields_ = [("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double), ("height", C.c_double),
                ("extrinsics", C.c_double * 9)]


# nis_track_result, one record per frame
TRACK_RESULT_DTYPE = np.dtype([("tracked", np.int32), ("inserted", np.int32), ("keyframe", np.int32), ("reserved", np.int32),
                               ("response", np.float64, 3), ("relative_pose", np.float64, 3), ("cf_pose", np.float64, 3),
                               ("pose", np.float64, 3), ("distance", np.float64)])


@dataclass
class KeyframeSelectionConfig:
    """include/read_configs.h:27-32; defaults = configs/config_ntu.yaml:19-23."""
    max_distance: float = 0.4
    max_angle: float = 0.052359877
    lower_response_thr: float = 30.0
    upper_response_thr: float = 90.0


@dataclass
class CameraModel:
    """What Camera's pose conversions read: new_K, height above ground, camera->robot extrinsics (src/camera.cc:148-218)."""
    fx: float = 1000.0
    fy: float = 1000.0
    cx: float = 320.0
    cy: float = 240.0
    height: float = 1.0
    extrinsics: tuple = (1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0)


@dataclass
class CFConfig:
    """include/read_configs.h:15-25 (width/height are overridden by the ctor, correlation_flow.cc:40-41)."""
    width: int = 640
    height: int = 480
    lam: float = 0.1
    kernel: int = 0
    sigma: float = 0.2
    offset: float = 0.1
    power: int = 3
    rotation_divisor: int = 720
    rotation_channel: int = 480


@dataclass
class LoopClosureConfig:
    """include/read_configs.h:38-44."""
    to_find_loop: bool = True
    position_response_thr: float = 60.0
    angle_response_thr: float = 60.0
    frame_gap_thr: int = 0
    distance_thr: float = 0.0


@dataclass
class LoopClosureResult:
    """include/loop_closure.h:8-25 (FramePtr -> slot / frame id)."""
    found: bool
    response: np.ndarray
    loop_slot: int
    loop_frame_id: int
    relative_pose: np.ndarray
    hyp: int
    peak: tuple
    evaluated: int
    raw: LoopResultC = None


_lib = None


def load_library():
    """Loads libnislam.so; raises if it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libnislam.so is not built (%s); run `python -m ni_slam_b200.build` -- this package has no CPU "
                           "fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i32, f64p = C.c_void_p, C.c_int, C.POINTER(C.c_double)
    lib.nis_create.argtypes = [C.POINTER(_CfConfigC), i32, i32, i32, C.POINTER(vp)]
    lib.nis_destroy.argtypes = [vp]
    lib.nis_last_error.argtypes = [vp]; lib.nis_last_error.restype = C.c_char_p
    lib.nis_strerror.argtypes = [i32]; lib.nis_strerror.restype = C.c_char_p
    lib.nis_stream.argtypes = [vp]; lib.nis_stream.restype = vp
    lib.nis_synchronize.argtypes = [vp]
    lib.nis_kernel_launches.argtypes = [vp]; lib.nis_kernel_launches.restype = C.c_longlong
    lib.nis_set_batch.argtypes = [vp, i32]
    lib.nis_set_lanes.argtypes = [vp, i32]
    lib.nis_track_stream_keyframes.argtypes = [vp, vp, i32, vp, vp, vp]
    lib.nis_stitcher_create.argtypes = [i32, i32, i32, i32, i32, i32, i32, i32, C.POINTER(vp)]
    lib.nis_stitcher_destroy.argtypes = [vp]
    lib.nis_stitcher_insert.argtypes = [vp, vp, vp, vp, vp]
    lib.nis_stitcher_recompute.argtypes = [vp, vp, vp]
    lib.nis_stitcher_frames.argtypes = [vp]
    lib.nis_stitcher_cell.argtypes = [vp, i32, i32, vp, vp, vp]
    lib.nis_stitcher_dropped.argtypes = [vp, vp]
    lib.nis_features_u8.argtypes = [vp, vp, C.POINTER(vp)]
    lib.nis_features_f32.argtypes = [vp, vp, C.POINTER(vp)]
    lib.nis_frame_export.argtypes = [vp, vp, vp, vp]
    lib.nis_frame_import.argtypes = [vp, vp, vp, vp, C.POINTER(vp)]
    lib.nis_frame_free.argtypes = [vp, vp]
    lib.nis_set_undistort_maps.argtypes = [vp, vp, vp]
    lib.nis_undistort_u8.argtypes = [vp, vp, vp]
    lib.nis_compute_pose.argtypes = [vp, vp, vp, i32, vp, vp, vp]
    lib.nis_track_stream.argtypes = [vp, vp, i32, vp, vp]
    lib.nis_track_stream_dev.argtypes = [vp, vp, i32, vp, vp]
    lib.nis_db_add.argtypes = [vp, vp, i32, C.c_double, C.POINTER(i32)]
    lib.nis_db_add_images.argtypes = [vp, vp, i32, vp, vp]
    lib.nis_db_add_images_dev.argtypes = [vp, vp, i32, vp, vp]
    lib.nis_db_size.argtypes = [vp]
    lib.nis_db_clear.argtypes = [vp]
    lib.nis_loop_scan.argtypes = [vp, vp, i32, C.c_double, C.POINTER(_LoopConfigC), vp, i32, C.POINTER(LoopResultC), vp]
    lib.nis_loop_reduce.argtypes = [vp, vp, i32, C.POINTER(_LoopConfigC), C.POINTER(LoopResultC), C.POINTER(i32)]
    lib.nis_profile_begin.argtypes = [vp]
    lib.nis_profile_end.argtypes = [vp, C.c_char_p, i32]
    lib.nis_db_set_position.argtypes = [vp, i32, C.c_double, C.c_double, C.c_double]
    lib.nis_loop_scan_prior.argtypes = [vp, vp, i32, C.c_double, C.POINTER(_LoopConfigC), C.c_double, C.c_double, C.c_double,
                                        C.POINTER(LoopResultC), vp, i32, C.POINTER(i32)]
    lib.nis_frame_import_ex.argtypes = [vp, vp, vp, vp, i32, C.POINTER(vp)]
    lib.nis_db_set_mode.argtypes = [vp, i32]
    lib.nis_db_mode.argtypes = [vp]
    lib.nis_db_add_spectra.argtypes = [vp, vp, vp, i32, C.c_double, C.POINTER(i32)]
    lib.nis_loop_scan_records.argtypes = [vp, vp, i32, C.c_double, C.POINTER(_LoopConfigC), vp, i32, C.POINTER(LoopResultC), vp]
    lib.nis_nccl_unique_id.argtypes = [C.c_char_p]
    lib.nis_comm_init.argtypes = [vp, C.c_char_p, i32, i32]
    lib.nis_comm_destroy.argtypes = [vp]
    lib.nis_loop_scan_sharded.argtypes = [vp, vp, i32, i32, C.c_double, C.POINTER(_LoopConfigC), C.c_longlong, C.POINTER(LoopResultC),
                                          C.POINTER(i32), C.POINTER(LoopResultC)]
    lib.nis_debug_fft2.argtypes = [vp, i32, vp, vp]
    lib.nis_debug_ifft2.argtypes = [vp, i32, vp, vp]
    lib.nis_debug_polar.argtypes = [vp, vp, vp]
    lib.nis_debug_rotate.argtypes = [vp, vp, C.c_float, vp]
    lib.nis_debug_estimate_trans.argtypes = [vp, i32, vp, vp, vp, C.POINTER(C.c_float), vp]
    _lib = lib
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Frame:
    """Device-resident Frame payload (include/frame.h:34-36): image + fft_result + fft_polar."""

    def __init__(self, cf, handle):
        self._cf, self._h = cf, handle

    def GetFFTResult(self):
        """Frame::GetFFTResult (src/frame.cc:53-57) -> (fft_result (H/2+1, W), fft_polar (D/2+1, Cp)) complex64."""
        cf = self._cf
        F = np.empty((cf.W, cf.H // 2 + 1), np.complex64)      # reference layout: W lines of H/2+1
        P = np.empty((cf.Cp, cf.D // 2 + 1), np.complex64)
        cf._check(cf._lib.nis_frame_export(cf._ctx, self._h, _p(F), _p(P)))
        return np.ascontiguousarray(F.T), np.ascontiguousarray(P.T)

    def free(self):
        # a context that is closed first releases the device blocks of its outstanding frames (nis_destroy); the handle is then
        # freed without a context
        if self._h is not None:
            self._cf._lib.nis_frame_free(self._cf._ctx, self._h)
        self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class CorrelationFlow:
    """include/correlation_flow.h:8-31.  ctor mirrors CorrelationFlow(CFConfig&, double& image_height, double& image_width)."""

    def __init__(self, cf_config: CFConfig, image_height, image_width, device: int = 0):
        self._lib = load_library()
        self._ctx = None
        self.cfg = cf_config
        self.H, self.W = int(image_height), int(image_width)        # correlation_flow.cc:40-41
        self.D, self.Cp = cf_config.rotation_divisor, cf_config.rotation_channel
        c = _CfConfigC(cf_config.lam, cf_config.kernel, cf_config.sigma, cf_config.offset, cf_config.power, self.D, self.Cp)
        ctx = C.c_void_p()
        st = self._lib.nis_create(C.byref(c), self.H, self.W, int(device), C.byref(ctx))
        if st != NIS_OK:
            raise NisError(st, self._lib.nis_strerror(st).decode())
        self._ctx = ctx

    # ---- plumbing
    def _check(self, st):
        if st == NIS_OK:
            return
        msg = self._lib.nis_last_error(self._ctx).decode() or self._lib.nis_strerror(st).decode()
        if st == NIS_ERR_INVALID_KERNEL:
            raise ValueError("Received invalid kernel type")        # std::invalid_argument, correlation_flow.cc:168
        raise NisError(st, msg)

    def close(self):
        if self._ctx is not None:
            self._lib.nis_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self) -> int:
        return int(self._lib.nis_stream(self._ctx) or 0)

    def synchronize(self):
        self._check(self._lib.nis_synchronize(self._ctx))

    def kernel_launches(self) -> int:
        return int(self._lib.nis_kernel_launches(self._ctx))

    def set_batch(self, b: int):
        self._check(self._lib.nis_set_batch(self._ctx, int(b)))

    def set_lanes(self, n: int):
        self._check(self._lib.nis_set_lanes(self._ctx, int(n)))

    def profile_begin(self):
        self._check(self._lib.nis_profile_begin(self._ctx))

    def profile_end(self) -> dict:
        """{kernel family: {"launches": n, "ms": total device time}} from CUDA events around every launch."""
        import json
        buf = C.create_string_buffer(1 << 16)
        self._check(self._lib.nis_profile_end(self._ctx, buf, len(buf)))
        return json.loads(buf.value.decode())

    # ---- reference surface
    def ComputeIntermedium(self, image) -> Frame:
        """MapBuilder::ComputeFFTResult (map_builder.cc:72-75): u8 (H, W) image (normalised on the GPU like
        ConvertMatToNormalizedArray) or f32 (H, W) array -> Frame holding fft_result and fft_polar."""
        image = np.asarray(image)
        if image.shape != (self.H, self.W):
            raise ValueError("image must be (%d, %d)" % (self.H, self.W))
        h = C.c_void_p()
        if image.dtype == np.uint8:
            img = np.ascontiguousarray(image)
            self._check(self._lib.nis_features_u8(self._ctx, _p(img), C.byref(h)))
        else:
            img = np.ascontiguousarray(image.astype(np.float32).T)          # reference layout (column-major)
            self._check(self._lib.nis_features_f32(self._ctx, _p(img), C.byref(h)))
        return Frame(self, h)

    def SetUndistortMaps(self, map1, map2):
        """Hand over the Camera's fixed-point remap maps (initUndistortRectifyMap(..., CV_16SC2), src/camera.cc:45-47): map1
        (H, W, 2) int16, map2 (H, W) uint16.  From then on every u8 image is treated as the RAW camera image and undistorted on
        the GPU first (Camera::UndistortImage, src/camera.cc:92-93).  Pass None, None to switch the front end off."""
        if map1 is None:
            self._check(self._lib.nis_set_undistort_maps(self._ctx, None, None))
            return
        m1 = np.ascontiguousarray(map1, np.int16)
        m2 = np.ascontiguousarray(map2, np.uint16)
        assert m1.shape == (self.H, self.W, 2) and m2.shape == (self.H, self.W)
        self._check(self._lib.nis_set_undistort_maps(self._ctx, _p(m1), _p(m2)))

    def UndistortImage(self, raw_u8):
        raw = np.ascontiguousarray(raw_u8, np.uint8)
        assert raw.shape == (self.H, self.W)
        out = np.empty_like(raw)
        self._check(self._lib.nis_undistort_u8(self._ctx, _p(raw), _p(out)))
        return out

    def ImportFrame(self, image_f32, fft_result, fft_polar) -> Frame:
        img = np.ascontiguousarray(np.asarray(image_f32, np.float32).T)
        F = np.ascontiguousarray(np.asarray(fft_result, np.complex64).T)
        P = np.ascontiguousarray(np.asarray(fft_polar, np.complex64).T)
        h = C.c_void_p()
        self._check(self._lib.nis_frame_import(self._ctx, _p(img), _p(F), _p(P), C.byref(h)))
        return Frame(self, h)

    def ImportFrameEx(self, image_f32=None, fft_result=None, fft_polar=None, with_h=False) -> Frame:
        """nis_frame_import_ex: any subset of (image, fft_result, fft_polar); what ComputePose reads is (fft_result, fft_polar, H) of the
        last frame and (image, fft_polar) of the current one."""
        img = np.ascontiguousarray(np.asarray(image_f32, np.float32).T) if image_f32 is not None else None
        F = np.ascontiguousarray(np.asarray(fft_result, np.complex64).T) if fft_result is not None else None
        P = np.ascontiguousarray(np.asarray(fft_polar, np.complex64).T) if fft_polar is not None else None
        h = C.c_void_p()
        self._check(self._lib.nis_frame_import_ex(self._ctx, _p(img), _p(F), _p(P), int(bool(with_h)), C.byref(h)))
        return Frame(self, h)

    # ---- multi-GPU scan plumbing (one context per rank)
    @staticmethod
    def NcclUniqueId() -> bytes:
        buf = C.create_string_buffer(128)
        st = load_library().nis_nccl_unique_id(buf)
        if st != NIS_OK:
            raise NisError(st, "ncclGetUniqueId failed (libnccl.so.2 not found? set NIS_NCCL_LIB)")
        return buf.raw

    def CommInit(self, unique_id: bytes, rank: int, n_ranks: int):
        self._check(self._lib.nis_comm_init(self._ctx, C.create_string_buffer(unique_id, 128), int(rank), int(n_ranks)))

    def ComputePose(self, last: Frame, cur: Frame, not_large_rotation: bool, return_peaks: bool = False):
        """CorrelationFlow::ComputePose (correlation_flow.cc:97-143): returns (info[3], pose[3]) like the reference
        returns `info` and fills `pose`; the current frame carries image and fft_polar."""
        pose = np.zeros(3, np.float64)
        info = np.zeros(3, np.float64)
        peak = np.zeros(4, np.int32)
        self._check(self._lib.nis_compute_pose(self._ctx, last._h, cur._h, int(bool(not_large_rotation)), _p(pose), _p(info), _p(peak)))
        if return_peaks:
            return info, pose, dict(polar=(int(peak[0]), int(peak[1])), trans=(int(peak[2]), int(peak[3])))
        return info, pose

    # ---- batched stream tracking (every frame a keyframe)
    def TrackStream(self, frames_u8):
        """frames (n, H, W) u8 in host memory -> (poses (n-1, 3), infos (n-1, 3))."""
        frames = np.ascontiguousarray(frames_u8, dtype=np.uint8)
        n = frames.shape[0]
        poses = np.zeros((max(n - 1, 0), 3), np.float64)
        infos = np.zeros((max(n - 1, 0), 3), np.float64)
        self._check(self._lib.nis_track_stream(self._ctx, _p(frames), n, _p(poses), _p(infos)))
        return poses, infos

    def TrackStreamKeyframes(self, frames_u8, kfs: "KeyframeSelectionConfig", cam: "CameraModel"):
        """MapBuilder::AddNewInput (tracking against the last keyframe, gate, pose composition, keyframe test; no loop closure)
        over a host stream -> structured array of n nis_track_result records (TRACK_RESULT_DTYPE)."""
        frames = np.ascontiguousarray(frames_u8, dtype=np.uint8)
        n = frames.shape[0]
        out = np.zeros(n, TRACK_RESULT_DTYPE)
        k = _KfsConfigC(kfs.max_distance, kfs.max_angle, kfs.lower_response_thr, kfs.upper_response_thr)
        c = _CameraModelC(cam.fx, cam.fy, cam.cx, cam.cy, cam.height, (C.c_double * 9)(*cam.extrinsics))
        self._check(self._lib.nis_track_stream_keyframes(self._ctx, _p(frames), n, C.byref(k), C.byref(c), _p(out)))
        return out

    def TrackStreamPtr(self, ptr: int, n: int, on_device: bool):
        """Same, frames given by raw pointer (pinned host or device memory, e.g. torch tensor .data_ptr())."""
        poses = np.zeros((max(n - 1, 0), 3), np.float64)
        infos = np.zeros((max(n - 1, 0), 3), np.float64)
        fn = self._lib.nis_track_stream_dev if on_device else self._lib.nis_track_stream
        self._check(fn(self._ctx, C.c_void_p(ptr), n, _p(poses), _p(infos)))
        return poses, infos

    # ---- stage-level entry points (parity tests)
    def debug_fft2(self, x, which=0):
        R, Cc = (self.H, self.W) if which == 0 else (self.D, self.Cp)
        x = np.ascontiguousarray(x, np.float32)
        assert x.shape == (R, Cc)
        out = np.empty((R // 2 + 1, Cc), np.complex64)
        self._check(self._lib.nis_debug_fft2(self._ctx, which, _p(x), _p(out)))
        return out

    def debug_ifft2(self, xf, which=0):
        R, Cc = (self.H, self.W) if which == 0 else (self.D, self.Cp)
        xf = np.ascontiguousarray(xf, np.complex64)
        assert xf.shape == (R // 2 + 1, Cc)
        out = np.empty((R, Cc), np.float32)
        self._check(self._lib.nis_debug_ifft2(self._ctx, which, _p(xf), _p(out)))
        return out

    def debug_polar(self, power):
        power = np.ascontiguousarray(power, np.float32)
        assert power.shape == (self.H, self.W)
        out = np.empty((self.D, self.Cp), np.float32)
        self._check(self._lib.nis_debug_polar(self._ctx, _p(power), _p(out)))
        return out

    def debug_rotate(self, img, degree):
        img = np.ascontiguousarray(img, np.float32)
        assert img.shape == (self.H, self.W)
        out = np.empty((self.H, self.W), np.float32)
        self._check(self._lib.nis_debug_rotate(self._ctx, _p(img), float(np.float32(degree)), _p(out)))
        return out

    def debug_estimate_trans(self, last_spec, cur_spec, which=0, want_g=False):
        R, Cc = (self.H, self.W) if which == 0 else (self.D, self.Cp)
        a = np.ascontiguousarray(last_spec, np.complex64)
        b = np.ascontiguousarray(cur_spec, np.complex64)
        assert a.shape == b.shape == (R // 2 + 1, Cc)
        peak = np.zeros(2, np.int32)
        info = C.c_float()
        g = np.empty((R, Cc), np.float32) if want_g else None
        self._check(self._lib.nis_debug_estimate_trans(self._ctx, which, _p(a), _p(b), _p(peak), C.byref(info), _p(g)))
        trans = (-(int(peak[0]) - R // 2), -(int(peak[1]) - Cc // 2))
        return float(info.value), trans, (int(peak[0]), int(peak[1])), g


class LoopClosure:
    """include/loop_closure.h:27-38 with the Map's frame store folded in (the keyframe DB lives on the GPU)."""

    def __init__(self, loop_closure_config: LoopClosureConfig, correlation_flow: CorrelationFlow):
        self._loop_thr = loop_closure_config
        self._cf = correlation_flow

    def _cfg_c(self):
        t = self._loop_thr
        return _LoopConfigC(t.position_response_thr, t.angle_response_thr, t.frame_gap_thr, t.distance_thr)

    def AddFrame(self, frame: Frame, frame_id: int, acc_distance: float = 0.0) -> int:
        """Map::AddFrame + Map::SetFrameDistance for the arrays the scan reads."""
        slot = C.c_int()
        self._cf._check(self._cf._lib.nis_db_add(self._cf._ctx, frame._h, int(frame_id), float(acc_distance), C.byref(slot)))
        return slot.value

    def AddImages(self, images_u8, frame_ids=None, acc_distances=None, ptr=None, n=None, on_device=False):
        """Bulk insert; features are computed on the GPU straight into the DB."""
        cf = self._cf
        if ptr is None:
            imgs = np.ascontiguousarray(images_u8, dtype=np.uint8)
            n, ptr = imgs.shape[0], imgs.ctypes.data
        ids = np.ascontiguousarray(frame_ids, np.int32) if frame_ids is not None else None
        ds = np.ascontiguousarray(acc_distances, np.float64) if acc_distances is not None else None
        fn = cf._lib.nis_db_add_images_dev if on_device else cf._lib.nis_db_add_images
        cf._check(fn(cf._ctx, C.c_void_p(ptr), int(n), _p(ids), _p(ds)))

    def AddSpectra(self, fft_result, fft_polar, frame_id: int, acc_distance: float = 0.0) -> int:
        """Map::AddFrame for a keyframe given as the arrays Frame::GetFFTResult hands out."""
        F = np.ascontiguousarray(np.asarray(fft_result, np.complex64).T)
        P = np.ascontiguousarray(np.asarray(fft_polar, np.complex64).T)
        slot = C.c_int()
        self._cf._check(self._cf._lib.nis_db_add_spectra(self._cf._ctx, _p(F), _p(P), int(frame_id), float(acc_distance), C.byref(slot)))
        return slot.value

    def SetMode(self, mode: int):
        """nis_db_set_mode: DB_FULL (F, P, Ht, Hp), DB_SPECTRA (F, P), DB_IMAGE (u8 image); only while the store is empty."""
        self._cf._check(self._cf._lib.nis_db_set_mode(self._cf._ctx, int(mode)))

    def size(self) -> int:
        return int(self._cf._lib.nis_db_size(self._cf._ctx))

    def FindLoopClosureRecords(self, current_frame: Frame, current_frame_id: int = 0, current_distance: float = 0.0, candidate_slots=None):
        """The scan with one nis_scan_record per entry of the candidate list (SCAN_RECORD_DTYPE): peaks, hypothesis, pose, response."""
        cf = self._cf
        cand = np.ascontiguousarray(candidate_slots, np.int32) if candidate_slots is not None else None
        n = int(cand.shape[0]) if cand is not None else 0
        recs = np.zeros(n if cand is not None else self.size(), SCAN_RECORD_DTYPE)
        out = LoopResultC()
        cfg = self._cfg_c()
        cf._check(cf._lib.nis_loop_scan_records(cf._ctx, current_frame._h, int(current_frame_id), float(current_distance), C.byref(cfg),
                                                _p(cand), n, C.byref(out), _p(recs)))
        return _result_from_c(out), recs

    def FindLoopClosureSharded(self, query_u8, root: int, global_slot_offset: int, current_frame_id: int = 0, current_distance: float = 0.0):
        """nis_loop_scan_sharded: one call per query on every rank -> (global result, winner rank, this rank's own best)."""
        cf = self._cf
        q = np.ascontiguousarray(query_u8, np.uint8) if query_u8 is not None else None
        out, loc, win = LoopResultC(), LoopResultC(), C.c_int(-1)
        cfg = self._cfg_c()
        cf._check(cf._lib.nis_loop_scan_sharded(cf._ctx, _p(q), int(root), int(current_frame_id), float(current_distance), C.byref(cfg),
                                                int(global_slot_offset), C.byref(out), C.byref(win), C.byref(loc)))
        return _result_from_c(out), win.value, _result_from_c(loc)

    def clear(self):
        self._cf._check(self._cf._lib.nis_db_clear(self._cf._ctx))

    def FindLoopClosure(self, current_frame: Frame, current_frame_id: int = 0, current_distance: float = 0.0,
                        candidate_slots=None, return_all: bool = False):
        """LoopClosure::FindLoopClosure(image, current_frame[, frames]) (loop_closure.cc:10-15, :36-73)."""
        cf = self._cf
        cand = np.ascontiguousarray(candidate_slots, np.int32) if candidate_slots is not None else None
        n = int(cand.shape[0]) if cand is not None else 0
        n_in = n if cand is not None else self.size()
        allr = np.zeros((n_in, 3), np.float64) if return_all else None
        out = LoopResultC()
        cfg = self._cfg_c()
        cf._check(cf._lib.nis_loop_scan(cf._ctx, current_frame._h, int(current_frame_id), float(current_distance), C.byref(cfg),
                                        _p(cand), n, C.byref(out), _p(allr)))
        res = _result_from_c(out)
        return (res, allr) if return_all else res

    def SetPosition(self, slot: int, x: float, y: float, grid_scale: float):
        """Map::AddFrame's grid filing (src/map.cc:27-30): the keyframe's pose at insertion time picks its cell."""
        self._cf._check(self._cf._lib.nis_db_set_position(self._cf._ctx, int(slot), float(x), float(y), float(grid_scale)))

    def FindLoopClosurePrior(self, current_frame: Frame, prior_pose, grid_scale: float, current_frame_id: int = 0,
                             current_distance: float = 0.0):
        """LoopClosure::FindLoopClosure(image, current_frame, prior_pose) (loop_closure.cc:17-34): only the keyframes filed in
        the 3x3 grid cells around the prior pose are scanned.  Returns (result, candidate slots in iteration order)."""
        cf = self._cf
        cap = max(self.size(), 1)
        cand = np.zeros(cap, np.int32)
        n = C.c_int()
        out = LoopResultC()
        cfg = self._cfg_c()
        cf._check(cf._lib.nis_loop_scan_prior(cf._ctx, current_frame._h, int(current_frame_id), float(current_distance), C.byref(cfg),
                                              float(prior_pose[0]), float(prior_pose[1]), float(grid_scale), C.byref(out), _p(cand),
                                              cap, C.byref(n)))
        return _result_from_c(out), cand[:n.value].copy()

    def Reduce(self, per_rank, order=None):
        """Multi-GPU: pick the winner among per-rank results with the reference's rule (strict '>', first wins)."""
        return loop_reduce(per_rank, order, self._loop_thr)


class MapStitcher:
    """MapStitcher (include/map_stitcher.h:24-41) on the GPU: InsertFrame / RecomputeOccupancy / GetOccupancyData over a dense window
    of cells [cell_x0, cell_x0 + cells_x) x [cell_y0, cell_y0 + cells_y)."""

    def __init__(self, cell_size: int, cam: CameraModel, image_height: int, image_width: int, cell_x0: int = -2, cell_y0: int = -2,
                 cells_x: int = 4, cells_y: int = 4, device: int = 0):
        self._lib = load_library()
        self._st = C.c_void_p()
        self.cs, self.H, self.W = int(cell_size), int(image_height), int(image_width)
        self._cam = _CameraModelC(cam.fx, cam.fy, cam.cx, cam.cy, cam.height, (C.c_double * 9)(*cam.extrinsics))
        self._check(self._lib.nis_stitcher_create(device, self.H, self.W, self.cs, cell_x0, cell_y0, cells_x, cells_y, C.byref(self._st)))

    def _check(self, status):
        if status != 0:
            raise NisError(status, self._lib.nis_strerror(status).decode())

    def InsertFrame(self, image_u8, robot_pose) -> int:
        img = np.ascontiguousarray(image_u8, np.uint8)
        assert img.shape == (self.H, self.W)
        pose = (C.c_double * 3)(*[float(v) for v in robot_pose])
        slot = C.c_int(-1)
        self._check(self._lib.nis_stitcher_insert(self._st, _p(img), pose, C.byref(self._cam), C.byref(slot)))
        return slot.value

    def RecomputeOccupancy(self, robot_poses):
        poses = np.ascontiguousarray(robot_poses, np.float64).reshape(-1, 3)
        assert poses.shape[0] == self.frames()
        self._check(self._lib.nis_stitcher_recompute(self._st, _p(poses), C.byref(self._cam)))

    def frames(self) -> int:
        return int(self._lib.nis_stitcher_frames(self._st))

    def cell(self, cell_x: int, cell_y: int):
        """-> (data, weight) int32 [cs, cs] indexed [in-cell y, in-cell x], or None when the cell does not exist."""
        data = np.zeros((self.cs, self.cs), np.int32)
        weight = np.zeros((self.cs, self.cs), np.int32)
        present = C.c_int(0)
        self._check(self._lib.nis_stitcher_cell(self._st, cell_x, cell_y, _p(data), _p(weight), C.byref(present)))
        return (data, weight) if present.value else None

    def dropped(self) -> int:
        v = C.c_longlong(0)
        self._check(self._lib.nis_stitcher_dropped(self._st, C.byref(v)))
        return int(v.value)

    def close(self):
        if self._st:
            self._lib.nis_stitcher_destroy(self._st)
            self._st = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def loop_reduce(per_rank, order, loop_cfg: LoopClosureConfig):
    """nis_loop_reduce: merge the all-gathered per-rank best records (host-only, needs no GPU).  `order[i]` = position of
    rank i's winner in the global iteration order (ties: smallest wins)."""
    lib = load_library()
    arr = (LoopResultC * len(per_rank))(*[r.raw if isinstance(r, LoopClosureResult) else r for r in per_rank])
    o = np.ascontiguousarray(order, np.int64) if order is not None else None
    out = LoopResultC()
    win = C.c_int()
    cfg = _LoopConfigC(loop_cfg.position_response_thr, loop_cfg.angle_response_thr, loop_cfg.frame_gap_thr, loop_cfg.distance_thr)
    st = lib.nis_loop_reduce(arr, _p(o), len(per_rank), C.byref(cfg), C.byref(out), C.byref(win))
    if st != NIS_OK:
        raise NisError(st, lib.nis_strerror(st).decode())
    return _result_from_c(out), win.value


def _result_from_c(out: LoopResultC) -> LoopClosureResult:
    return LoopClosureResult(found=bool(out.found), response=np.array(out.response[:]), loop_slot=out.slot,
                             loop_frame_id=out.frame_id, relative_pose=np.array(out.relative_pose[:]), hyp=out.hyp,
                             peak=tuple(out.peak[:]), evaluated=out.evaluated, raw=out)
