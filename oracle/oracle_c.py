"""ctypes loader for the C oracle (oracle/nislam_oracle.c).  TEST INFRASTRUCTURE ONLY (see that file's header).

Arrays cross this interface in the reference's layout: real R x C arrays column-major (numpy: shape (C, R)
C-contiguous, or equivalently an (R, C) Fortran array), half spectra as (C, R/2+1) complex64.
The helpers below take/return natural numpy (row, col)-indexed arrays and do the layout conversion.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libnislam_oracle.so")


class CFConfigC(C.Structure):
    _fields_ = [("height", C.c_int), ("width", C.c_int), ("lam", C.c_float), ("kernel", C.c_int),
                ("sigma", C.c_float), ("offset", C.c_float), ("power", C.c_int),
                ("rotation_divisor", C.c_int), ("rotation_channel", C.c_int)]


class LoopConfigC(C.Structure):
    _fields_ = [("position_response_thr", C.c_double), ("angle_response_thr", C.c_double),
                ("frame_gap_thr", C.c_int), ("distance_thr", C.c_double)]


class LoopResultC(C.Structure):
    _fields_ = [("found", C.c_int), ("index", C.c_int), ("frame_id", C.c_int),
                ("relative_pose", C.c_double * 3), ("response", C.c_double * 3)]


class PeaksC(C.Structure):
    _fields_ = [("polar_row", C.c_int), ("polar_col", C.c_int), ("trans_row", C.c_int), ("trans_col", C.c_int),
                ("hyp", C.c_int), ("degree", C.c_float)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "nislam_oracle.c")
    if force or not os.path.exists(_SO) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_SO)):
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_rotate.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p]
        _lib.orc_normalize_degree.restype = C.c_double
        _lib.orc_normalize_degree.argtypes = [C.c_double]
        _lib.orc_rotation_inverse.argtypes = [C.c_int, C.c_int, C.c_double, C.c_void_p]
        _lib.orc_find_loop_closure.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double,
                                               C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                               C.c_void_p]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def to_colmajor(x):
    """(R, C) natural array -> buffer in the reference layout (C lines of R)."""
    return np.ascontiguousarray(np.asarray(x).T)


def from_colmajor(buf):
    return np.ascontiguousarray(buf.T)


def make_cfg(height=480, width=640, lam=0.1, kernel=0, sigma=0.2, offset=0.1, power=3, rotation_divisor=720,
             rotation_channel=480):
    return CFConfigC(height, width, lam, kernel, sigma, offset, power, rotation_divisor, rotation_channel)


def fft2(x):
    R, Cc = x.shape
    xc = to_colmajor(x.astype(np.float32))
    out = np.zeros((Cc, R // 2 + 1), np.complex64)
    lib().orc_fft2(_p(xc), R, Cc, _p(out))
    return from_colmajor(out)


def ifft2(xf):
    half, Cc = xf.shape
    R = (half - 1) * 2
    xc = to_colmajor(xf.astype(np.complex64))
    out = np.zeros((Cc, R), np.float32)
    lib().orc_ifft2(_p(xc), R, Cc, _p(out))
    return from_colmajor(out)


def normalize_u8(img):
    H, W = img.shape
    out = np.zeros((W, H), np.float32)
    lib().orc_normalize_u8(_p(np.ascontiguousarray(img, dtype=np.uint8)), H, W, _p(out))
    return from_colmajor(out)


def remove_zero_component(x):
    R, Cc = x.shape
    out = np.zeros((Cc, R), np.float32)
    lib().orc_remove_zero_component(_p(to_colmajor(x.astype(np.float32))), R, Cc, _p(out))
    return from_colmajor(out)


def fftshift(x):
    R, Cc = x.shape
    out = np.zeros((Cc, R), np.float32)
    lib().orc_fftshift(_p(to_colmajor(x.astype(np.float32))), R, Cc, _p(out))
    return from_colmajor(out)


def polar(x, D=720, Cp=480):
    H, W = x.shape
    out = np.zeros((Cp, D), np.float32)
    lib().orc_polar(_p(to_colmajor(x.astype(np.float32))), H, W, D, Cp, _p(out))
    return from_colmajor(out)


def rotate(x, degree):
    H, W = x.shape
    out = np.zeros((W, H), np.float32)
    lib().orc_rotate(_p(to_colmajor(x.astype(np.float32))), H, W, float(np.float32(degree)), _p(out))
    return from_colmajor(out)


def undistort_u8(raw, map1, map2):
    H, W = raw.shape
    out = np.zeros((H, W), np.uint8)
    lib().orc_undistort_u8(_p(np.ascontiguousarray(raw, np.uint8)), H, W, _p(np.ascontiguousarray(map1, np.int16)),
                           _p(np.ascontiguousarray(map2, np.uint16)), _p(out))
    return out


def rotation_inverse(H, W, degree):
    m = np.zeros(6, np.float64)
    lib().orc_rotation_inverse(H, W, float(degree), _p(m))
    return m


def compute_intermedium(cfg, image):
    """image (H, W) f32 -> (fft_result (H/2+1, W), fft_polar (D/2+1, Cp)) complex64."""
    H, W, D, Cp = cfg.height, cfg.width, cfg.rotation_divisor, cfg.rotation_channel
    F = np.zeros((W, H // 2 + 1), np.complex64)
    P = np.zeros((Cp, D // 2 + 1), np.complex64)
    lib().orc_compute_intermedium(C.byref(cfg), _p(to_colmajor(image.astype(np.float32))), _p(F), _p(P))
    return from_colmajor(F), from_colmajor(P)


def estimate_trans(cfg, last_fft, cur_fft, R, Cc, want_g=False):
    trans = (C.c_int * 2)()
    peak = (C.c_int * 2)()
    info = C.c_float()
    g = np.zeros((Cc, R), np.float32) if want_g else None
    rc = lib().orc_estimate_trans(C.byref(cfg), _p(to_colmajor(last_fft.astype(np.complex64))),
                                  _p(to_colmajor(cur_fft.astype(np.complex64))), R, Cc, trans, peak, C.byref(info),
                                  _p(g) if want_g else None)
    if rc:
        raise ValueError("Received invalid kernel type")
    return float(info.value), (trans[0], trans[1]), (peak[0], peak[1]), (from_colmajor(g) if want_g else None)


def compute_pose(cfg, last_fft_result, image, last_fft_polar, fft_polar, not_large_rotation):
    pose = (C.c_double * 3)()
    info = (C.c_double * 3)()
    pk = PeaksC()
    rc = lib().orc_compute_pose(C.byref(cfg), _p(to_colmajor(last_fft_result.astype(np.complex64))),
                                _p(to_colmajor(image.astype(np.float32))),
                                _p(to_colmajor(last_fft_polar.astype(np.complex64))),
                                _p(to_colmajor(fft_polar.astype(np.complex64))), int(bool(not_large_rotation)), pose,
                                info, C.byref(pk))
    if rc:
        raise ValueError("Received invalid kernel type")
    peaks = dict(polar=(pk.polar_row, pk.polar_col), trans=(pk.trans_row, pk.trans_col), hyp=pk.hyp,
                 degree=float(pk.degree))
    return np.array(info[:]), np.array(pose[:]), peaks


def find_loop_closure(cfg, thr, image, cur_fft_polar, cur_id, cur_dist, keyframes, threads=1):
    """keyframes: list of (frame_id, fft_result, fft_polar, distance)."""
    n = len(keyframes)
    Fs = [to_colmajor(k[1].astype(np.complex64)) for k in keyframes]
    Ps = [to_colmajor(k[2].astype(np.complex64)) for k in keyframes]
    fp = (C.c_void_p * max(n, 1))(*[a.ctypes.data for a in Fs])
    pp = (C.c_void_p * max(n, 1))(*[a.ctypes.data for a in Ps])
    ids = np.array([k[0] for k in keyframes] or [0], np.int32)
    ds = np.array([k[3] for k in keyframes] or [0.0], np.float64)
    out = LoopResultC()
    rc = lib().orc_find_loop_closure(C.addressof(cfg), C.addressof(thr), _p(to_colmajor(image.astype(np.float32))),
                                     _p(to_colmajor(cur_fft_polar.astype(np.complex64))), int(cur_id), float(cur_dist),
                                     n, fp, pp, _p(ids), _p(ds), int(threads), C.addressof(out))
    if rc:
        raise ValueError("Received invalid kernel type")
    return dict(found=bool(out.found), index=out.index, frame_id=out.frame_id,
                relative_pose=np.array(out.relative_pose[:]), response=np.array(out.response[:]))


def track_stream(cfg, frames_u8, threads=1):
    n = frames_u8.shape[0]
    poses = np.zeros((max(n - 1, 0), 3), np.float64)
    infos = np.zeros((max(n - 1, 0), 3), np.float64)
    rc = lib().orc_track_stream(C.byref(cfg), _p(np.ascontiguousarray(frames_u8, dtype=np.uint8)), n, int(threads),
                                _p(poses), _p(infos))
    if rc:
        raise ValueError("Received invalid kernel type")
    return poses, infos


def num_threads():
    return int(lib().orc_num_threads())
