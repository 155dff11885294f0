"""ctypes loader for oracle/_ref/libnislam_ref.so: the reference's OWN src/{correlation_flow,loop_closure,utils,map,frame}.cc
compiled unmodified against the stand-in headers of oracle/ref_stubs (recipe: oracle/Makefile.ref, entry points: oracle/ref_shim.cc).

TEST INFRASTRUCTURE ONLY.  It pins the C restatement (oracle/nislam_oracle.c) and the Python one (oracle/nislam_ref.py) to the
reference's source text: control flow, quirks and operation order are the reference's; FFTW / OpenCV / Eigen arithmetic is
what the stand-ins document.  /root/reference does not exist on the GPU box: there only the prebuilt .so is used.

Same calling conventions as oracle_c (natural (row, col) numpy arrays in, layout conversion inside).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from oracle_c import CFConfigC, LoopConfigC, LoopResultC, _p, from_colmajor, make_cfg, to_colmajor  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libnislam_ref.so")
REFERENCE = os.environ.get("NIS_REFERENCE_DIR", "/root/reference")


def available() -> bool:
    """True when the .so exists or can be built here (the reference sources are present)."""
    return os.path.exists(_SO) or os.path.isdir(os.path.join(REFERENCE, "src"))


def build(force: bool = False) -> str:
    """make -f Makefile.ref when /root/reference is present; otherwise the prebuilt .so (it travels to the GPU box) is used as is."""
    if os.path.isdir(os.path.join(REFERENCE, "src")):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-f", "Makefile.ref", "REF=" + REFERENCE] + (["-B"] if force else []))
    if not os.path.exists(_SO):
        raise FileNotFoundError("oracle/_ref/libnislam_ref.so is not built and %s is absent" % REFERENCE)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.ref_create.restype = C.c_void_p
        _lib.ref_create.argtypes = [C.c_void_p, C.c_double, C.c_double]
        _lib.ref_destroy.argtypes = [C.c_void_p]
        _lib.ref_normalize_degree.restype = C.c_double
        _lib.ref_normalize_degree.argtypes = [C.c_double]
        _lib.ref_rotate.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p]
        _lib.ref_normalize_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        _lib.ref_compute_intermedium.argtypes = [C.c_void_p] * 4
        _lib.ref_compute_pose.argtypes = [C.c_void_p] * 5 + [C.c_int, C.c_void_p, C.c_void_p]
        _lib.ref_find_loop_closure.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                               C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    return _lib


class CorrelationFlow:
    """The reference's CorrelationFlow object (correlation_flow.cc:37-44 ctor: cfg.height / width := the camera's image size)."""

    def __init__(self, cfg: CFConfigC):
        self.cfg = cfg
        self.h = lib().ref_create(C.addressof(cfg), float(cfg.height), float(cfg.width))

    def close(self):
        if self.h:
            lib().ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def compute_intermedium(self, image):
        cfg = self.cfg
        F = np.zeros((cfg.width, cfg.height // 2 + 1), np.complex64)
        P = np.zeros((cfg.rotation_channel, cfg.rotation_divisor // 2 + 1), np.complex64)
        lib().ref_compute_intermedium(self.h, _p(to_colmajor(image.astype(np.float32))), _p(F), _p(P))
        return from_colmajor(F), from_colmajor(P)

    def compute_pose(self, last_fft_result, image, last_fft_polar, fft_polar, not_large_rotation):
        pose = (C.c_double * 3)()
        info = (C.c_double * 3)()
        rc = lib().ref_compute_pose(self.h, _p(to_colmajor(last_fft_result.astype(np.complex64))),
                                    _p(to_colmajor(image.astype(np.float32))), _p(to_colmajor(last_fft_polar.astype(np.complex64))),
                                    _p(to_colmajor(fft_polar.astype(np.complex64))), int(bool(not_large_rotation)), pose, info)
        if rc:
            raise ValueError("Received invalid kernel type")
        return np.array(info[:]), np.array(pose[:])

    def find_loop_closure(self, thr: LoopConfigC, image, cur_fft_result, cur_fft_polar, cur_id, cur_dist, keyframes, mode=0,
                          poses=None, prior_pose=None, grid_scale=1.0):
        """keyframes: list of (frame_id, fft_result, fft_polar, distance-or-None).  mode 0 = the given list in order
        (loop_closure.cc:36-73), 1 = all frames of the map in id order (:10-15), 2 = 3x3 grid cells around prior_pose (:17-34)."""
        n = len(keyframes)
        Fs = [to_colmajor(k[1].astype(np.complex64)) for k in keyframes]
        Ps = [to_colmajor(k[2].astype(np.complex64)) for k in keyframes]
        fp = (C.c_void_p * max(n, 1))(*[a.ctypes.data for a in Fs])
        pp = (C.c_void_p * max(n, 1))(*[a.ctypes.data for a in Ps])
        ids = np.array([k[0] for k in keyframes] or [0], np.int32)
        has_d = n > 0 and keyframes[0][3] is not None
        ds = np.array([k[3] for k in keyframes], np.float64) if has_d else None
        ps = np.ascontiguousarray(poses, np.float64) if poses is not None else None
        pr = np.ascontiguousarray(prior_pose, np.float64) if prior_pose is not None else None
        out = LoopResultC()
        rc = lib().ref_find_loop_closure(self.h, C.addressof(thr), float(grid_scale), _p(to_colmajor(image.astype(np.float32))),
                                         _p(to_colmajor(cur_fft_result.astype(np.complex64))),
                                         _p(to_colmajor(cur_fft_polar.astype(np.complex64))), int(cur_id),
                                         float(cur_dist if cur_dist is not None else 0.0), int(cur_dist is not None), n, fp, pp,
                                         _p(ids), _p(ds) if ds is not None else None, _p(ps) if ps is not None else None, int(mode),
                                         _p(pr) if pr is not None else None, C.addressof(out))
        if rc:
            raise ValueError("Received invalid kernel type")
        return dict(found=bool(out.found), index=out.index, frame_id=out.frame_id, relative_pose=np.array(out.relative_pose[:]),
                    response=np.array(out.response[:]))


def normalize_u8(img):
    H, W = img.shape
    out = np.zeros((W, H), np.float32)
    lib().ref_normalize_u8(_p(np.ascontiguousarray(img, dtype=np.uint8)), H, W, _p(out))
    return from_colmajor(out)


def normalize_degree(a):
    return float(lib().ref_normalize_degree(float(a)))


def rotate(x, degree):
    H, W = x.shape
    out = np.zeros((W, H), np.float32)
    lib().ref_rotate(_p(to_colmajor(x.astype(np.float32))), H, W, float(np.float32(degree)), _p(out))
    return from_colmajor(out)
