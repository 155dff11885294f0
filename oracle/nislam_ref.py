"""CPU restatement of NI-SLAM's tracking / loop-closure hot path (numpy + scipy.fft(f32) + genuine cv2).

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (ni_slam_b200/) may import this file; only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use oracle/.

PARITY UNPINNED: the reference (sair-lab/ni-slam @ 819f252) ships no tests, golden vectors or
fixtures, and cannot be compiled here (Eigen, FFTW3, OpenCV-C++ absent).  This file
restates the reference's arithmetic line by line; the third-party arithmetic it stands on is
  * FFTW3f r2c/c2r 2-D (version unpinned, CMakeLists.txt:21)  -> scipy.fft (pocketfft) on float32
  * OpenCV 4.2 warpPolar / getRotationMatrix2D / warpAffine    -> python cv2 4.13 (the genuine library)
  * Eigen 3 maxCoeff / array expressions                        -> numpy, column-major first maximum
It is pinned only by the analytic known-answer tests of SURVEY.md Appendix C (tests/test_oracle_*.py).

All arrays are numpy 2-D, indexed (row, col) like the Eigen arrays of the reference; "column-major"
only matters for argmax tie-breaks and for the byte layout at the C-ABI boundary.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import scipy.fft as sfft

try:  # cv2 is present in this image; the C oracle re-implements the warps without it
    import cv2
except Exception:  # pragma: no cover
    cv2 = None

f32 = np.float32


@dataclass
class CFConfig:
    """include/read_configs.h:15-25 (width/height are overridden by the ctor, correlation_flow.cc:40-41)."""
    height: int = 480
    width: int = 640
    lam: float = 0.1          # 'lambda'
    kernel: int = 0           # 0 polynomial, 1 gaussian
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


# ----------------------------------------------------------------------------------------------
# L1 helpers
# ----------------------------------------------------------------------------------------------
def convert_mat_to_normalized_array(u8: np.ndarray) -> np.ndarray:
    """src/utils.cc:110-118: cv2eigen (u8 -> float) then array/255.0 (double divide, stored as float)."""
    return (u8.astype(f32).astype(np.float64) / 255.0).astype(f32)


def normalize_degree(angle_degree: float) -> float:
    """src/utils.cc:173-175 (double)."""
    return angle_degree - 360.0 * math.floor((angle_degree + 180.0) / 360.0)


def rotate_array(array: np.ndarray, degree) -> np.ndarray:
    """src/utils.cc:154-161: getRotationMatrix2D((W/2., H/2.), degree, 1) + warpAffine(LINEAR, BORDER_WRAP)."""
    h, w = array.shape
    m = cv2.getRotationMatrix2D((float(f32(w / 2.0)), float(f32(h / 2.0))), float(degree), 1.0)
    return cv2.warpAffine(np.ascontiguousarray(array, dtype=f32), m, (w, h), flags=cv2.INTER_LINEAR,
                          borderMode=cv2.BORDER_WRAP)


def fftshift(x: np.ndarray) -> np.ndarray:
    """include/circ_shift.h:238-244: out(r,c) = in((r - R/2) mod R, (c - C/2) mod C)."""
    return np.roll(x, (x.shape[0] // 2, x.shape[1] // 2), axis=(0, 1))


# ----------------------------------------------------------------------------------------------
# L2 CorrelationFlow
# ----------------------------------------------------------------------------------------------
class CorrelationFlow:
    """src/correlation_flow.cc:37-243."""

    def __init__(self, cfg: CFConfig, image_height: float, image_width: float, verbose: bool = False):
        cfg = CFConfig(**vars(cfg))
        cfg.height = int(image_height)           # :40-41
        cfg.width = int(image_width)
        self.cfg = cfg
        self.verbose = verbose
        self.target_fft = self.get_target_fft(cfg.height, cfg.width)
        self.target_rotation_fft = self.get_target_fft(cfg.rotation_divisor, cfg.rotation_channel)
        self.last_stages = {}

    # :46-51
    def get_target_fft(self, rows, cols):
        t = np.zeros((rows, cols), f32)
        t[rows // 2, cols // 2] = 1
        return self.fft(t)

    # :53-63  r2c, halved along rows (FFTW called with n0=cols, n1=rows on column-major data)
    @staticmethod
    def fft(x):
        return sfft.rfft2(np.asarray(x, f32), axes=(1, 0)).astype(np.complex64, copy=False)

    # :65-77  c2r then / size
    @staticmethod
    def ifft(xf):
        rows = (xf.shape[0] - 1) * 2
        cols = xf.shape[1]
        return sfft.irfft2(np.asarray(xf, np.complex64), s=(cols, rows), axes=(1, 0)).astype(f32, copy=False)

    # :79-87
    @staticmethod
    def remove_zero_component(x):
        y = x.copy()
        rows, cols = x.shape
        y[0, :] = ((x[1, :] + x[rows - 1, :]).astype(np.float64) / 2.0).astype(f32)
        y[:, 0] = ((x[:, 1] + x[:, cols - 1]).astype(np.float64) / 2.0).astype(f32)
        return y

    # :228-236
    def polar(self, array):
        h, w = array.shape
        center = (float(f32(w) / f32(2)), float(f32(h) / f32(2)))
        radius = float(min(h // 2, w // 2))
        dsize = (self.cfg.rotation_channel, self.cfg.rotation_divisor)
        return cv2.warpPolar(np.ascontiguousarray(array, dtype=f32), dsize, center, radius,
                             cv2.INTER_LINEAR + cv2.WARP_FILL_OUTLIERS)

    # :89-95
    def compute_intermedium(self, image):
        fft_result = self.fft(image)
        power = self.ifft(np.abs(fft_result))
        high_power = self.remove_zero_component(power)
        polar_img = self.polar(fftshift(high_power))
        fft_polar = self.fft(polar_img)
        self.last_stages = dict(power=power, high_power=high_power, polar=polar_img)
        return fft_result, fft_polar

    # :208-226
    def polynomial_kernel(self, xf, zf=None):
        if zf is None:
            zf = xf
        xz = self.ifft(xf * np.conj(zf))
        base = xz + f32(self.cfg.offset)
        # Eigen 3.3 ArrayBase::pow(int) -> std::pow(float,int) -> double pow, rounded back to float
        kernel = np.power(base.astype(np.float64), int(self.cfg.power)).astype(f32)
        kernel = kernel / np.abs(kernel).max()
        return self.fft(kernel), kernel

    # :181-206
    def gaussian_kernel(self, xf, zf=None):
        n = self.cfg_n
        xx = f32(np.abs(np.square(xf)).astype(f32).sum(dtype=f32)) / f32(n)
        if zf is None:
            zz = xx
            zf = xf
        else:
            zz = f32(np.abs(np.square(zf)).astype(f32).sum(dtype=f32)) / f32(n)
        xz = self.ifft(xf * np.conj(zf))
        xxzz = (xx + zz - f32(2) * xz) / f32(n)
        kernel = np.exp(f32(-1.0 / (f32(self.cfg.sigma) * f32(self.cfg.sigma))) * xxzz).astype(f32)
        kernel = kernel / np.abs(kernel).max()
        return self.fft(kernel), kernel

    # :238-243
    @staticmethod
    def get_info(output, response):
        side_lobe_mean = f32((f32(output.sum(dtype=np.float64)) - response) / f32(output.size - 1))
        std = f32(math.sqrt(float(np.square((output - side_lobe_mean).astype(np.float64)).mean())))
        return f32((float(response) - float(side_lobe_mean)) / (float(std) + 1e-7))

    # :145-179
    def estimate_trans(self, last_fft_result, fft_result, output_fft, height, width):
        self.cfg_n = height * width
        if self.cfg.kernel == 0:
            kzz, _ = self.polynomial_kernel(last_fft_result)
            kxz, _ = self.polynomial_kernel(fft_result, last_fft_result)
        elif self.cfg.kernel == 1:
            kzz, _ = self.gaussian_kernel(last_fft_result)
            kxz, _ = self.gaussian_kernel(fft_result, last_fft_result)
        else:
            raise ValueError("Received invalid kernel type")     # std::invalid_argument :168
        h = output_fft / (kzz + f32(self.cfg.lam))
        g_hat = (h * kxz).astype(np.complex64)
        g = self.ifft(g_hat)
        # Eigen maxCoeff on a column-major array: first maximum in column-major order
        flat = np.argmax(g.T.reshape(-1))
        col, row = divmod(int(flat), g.shape[0])
        response = g[row, col]
        trans = (-(row - height // 2), -(col - width // 2))
        info = self.get_info(g, response)
        return info, trans, (row, col), g

    # :97-143
    def compute_pose(self, last_fft_result, image, last_fft_polar, fft_polar, not_large_rotation: bool):
        cfg = self.cfg
        info_rots, rots, peak_rot, _ = self.estimate_trans(
            last_fft_polar, fft_polar, self.target_rotation_fft, cfg.rotation_divisor, cfg.rotation_channel)
        degree = f32(rots[0] * (2.0 / cfg.rotation_divisor) * 180)
        degree = f32(normalize_degree(float(degree)))
        if not_large_rotation:
            degree = f32(degree - f32(180)) if abs(degree) > 90 else degree
            fft_rot_orig = self.fft(rotate_array(image, -degree))
            info_trans, trans, peak_t, _ = self.estimate_trans(
                last_fft_result, fft_rot_orig, self.target_fft, cfg.height, cfg.width)
            hyp = 0
        else:
            fft_rot_orig = self.fft(rotate_array(image, -degree))
            fft_rot_veri = self.fft(rotate_array(image, f32(-degree + f32(180))))
            io, to, po, _ = self.estimate_trans(last_fft_result, fft_rot_orig, self.target_fft, cfg.height, cfg.width)
            iv, tv, pv, _ = self.estimate_trans(last_fft_result, fft_rot_veri, self.target_fft, cfg.height, cfg.width)
            if io > iv:
                info_trans, trans, peak_t, hyp = io, to, po, 0
            else:
                info_trans, trans, peak_t, hyp = iv, tv, pv, 1
                degree = f32(degree + f32(180))
        if degree > 180:
            degree = f32(degree - f32(360))
        theta = f32(float(f32(degree / f32(180))) * math.pi)
        pose = np.array([trans[1], trans[0], float(theta)], np.float64)
        info = np.array([float(info_trans), float(info_trans), float(info_rots)], np.float64)
        self.last_peaks = dict(polar=peak_rot, trans=peak_t, degree=float(degree), hyp=hyp)
        return info, pose


# ----------------------------------------------------------------------------------------------
# L4 LoopClosure scan  (src/loop_closure.cc:36-73, include/loop_closure.h:15)
# ----------------------------------------------------------------------------------------------
@dataclass
class Keyframe:
    frame_id: int
    fft_result: np.ndarray
    fft_polar: np.ndarray
    distance: float = 0.0


@dataclass
class LoopClosureResult:
    found: bool = False
    response: np.ndarray = field(default_factory=lambda: np.array([-1.0, -1.0, -1.0]))
    loop_index: int = -1          # index into the candidate list (FramePtr in the reference)
    loop_frame_id: int = -1
    relative_pose: np.ndarray = field(default_factory=lambda: np.zeros(3))


def find_loop_closure(cf: CorrelationFlow, thr: LoopClosureConfig, image, cur: Keyframe, frames):
    result = LoopClosureResult()
    for idx, fr in enumerate(frames):
        if thr.frame_gap_thr > 0 and abs(cur.frame_id - fr.frame_id) < thr.frame_gap_thr:
            continue
        if thr.distance_thr > 0 and abs(cur.distance - fr.distance) < thr.distance_thr:
            continue
        response, rel = cf.compute_pose(fr.fft_result, image, fr.fft_polar, cur.fft_polar, False)
        if response.sum() > result.response.sum():
            result.response = response
            result.loop_index = idx
            result.loop_frame_id = fr.frame_id
            result.relative_pose = rel
    result.found = bool(result.response[0] > thr.position_response_thr and
                        result.response[2] > thr.angle_response_thr)
    return result


# ----------------------------------------------------------------------------------------------
# synthetic inputs (SURVEY.md section 8d / Appendix C fixture)
# ----------------------------------------------------------------------------------------------
def make_canvas(seed=0, shape=(960, 1280), sigma=2.0):
    c = np.random.default_rng(seed).random(shape).astype(f32)
    c = cv2.GaussianBlur(c, (0, 0), sigma)
    c = (c - c.min()) / (c.max() - c.min())
    return c.astype(f32)


def crop(canvas, cx, cy, ang, h=480, w=640):
    m = cv2.getRotationMatrix2D((float(cx), float(cy)), float(ang), 1.0)
    m[0, 2] += w / 2 - cx
    m[1, 2] += h / 2 - cy
    img = cv2.warpAffine(canvas, m, (w, h), flags=cv2.INTER_LINEAR)
    return np.clip(np.rint(img * 255.0), 0, 255).astype(np.uint8)
