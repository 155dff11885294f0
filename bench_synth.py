"""Synthetic inputs for bench.py / smoke: a blurred random ground-texture canvas and u8 camera frames
cropped from it at known poses.  Generated with torch on the GPU (or CPU) -- no datasets, no network.

crop convention = cv2.getRotationMatrix2D((cx, cy), ang, 1) shifted so the centre lands on (W/2, H/2), i.e.
    src_x = cx + cos(a) (x - W/2) - sin(a) (y - H/2),   src_y = cy + sin(a) (x - W/2) + cos(a) (y - H/2)
bilinear sampling, then rint(255 v) as u8 (the reference's input is a grayscale cv::Mat, src/dataset.cc:38-46).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


def make_canvas(size: int = 4096, seed: int = 0, sigma: float = 2.0, device="cpu") -> torch.Tensor:
    g = torch.Generator(device="cpu").manual_seed(seed)
    c = torch.rand((size, size), generator=g, dtype=torch.float32).to(device)
    r = int(math.ceil(3 * sigma))
    x = torch.arange(-r, r + 1, dtype=torch.float32, device=device)
    k = torch.exp(-0.5 * (x / sigma) ** 2)
    k = (k / k.sum()).view(1, 1, 1, -1)
    c = c.view(1, 1, size, size)
    c = F.conv2d(F.pad(c, (r, r, 0, 0), mode="circular"), k)
    c = F.conv2d(F.pad(c, (0, 0, r, r), mode="circular"), k.transpose(2, 3))
    c = (c - c.min()) / (c.max() - c.min())
    return c.view(size, size).contiguous()


def crops(canvas: torch.Tensor, cx, cy, ang_deg, H: int = 480, W: int = 640, chunk: int = 64) -> torch.Tensor:
    """canvas (S, S) f32; cx, cy, ang_deg: 1-D arrays of n poses -> (n, H, W) uint8 on canvas.device."""
    dev = canvas.device
    S = canvas.shape[0]
    cx = torch.as_tensor(np.asarray(cx), dtype=torch.float32, device=dev)
    cy = torch.as_tensor(np.asarray(cy), dtype=torch.float32, device=dev)
    a = torch.deg2rad(torch.as_tensor(np.asarray(ang_deg), dtype=torch.float32, device=dev))
    n = cx.shape[0]
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32, device=dev) - H / 2,
                            torch.arange(W, dtype=torch.float32, device=dev) - W / 2, indexing="ij")
    out = torch.empty((n, H, W), dtype=torch.uint8, device=dev)
    src = canvas.view(1, 1, S, S)
    for i0 in range(0, n, chunk):
        sl = slice(i0, min(n, i0 + chunk))
        ca, sa = torch.cos(a[sl]).view(-1, 1, 1), torch.sin(a[sl]).view(-1, 1, 1)
        sx = cx[sl].view(-1, 1, 1) + ca * xs - sa * ys
        sy = cy[sl].view(-1, 1, 1) + sa * xs + ca * ys
        grid = torch.stack((2 * sx / (S - 1) - 1, 2 * sy / (S - 1) - 1), dim=-1)
        v = F.grid_sample(src.expand(grid.shape[0], 1, S, S), grid, mode="bilinear", padding_mode="zeros", align_corners=True)
        out[sl] = torch.clamp(torch.round(v[:, 0] * 255.0), 0, 255).to(torch.uint8)
    return out


def stream_poses(n: int, seed: int = 0, size: int = 4096, H: int = 480, W: int = 640):
    """Cumulative camera poses of the tracking stream: per-frame dx, dy ~ U{-20..20} px, dang ~ U(-5, 5) deg in 0.5 deg
    steps; the walk is reflected off a margin so every crop stays inside the canvas."""
    rng = np.random.default_rng(seed)
    margin = int(math.hypot(H, W) / 2) + 8
    cx, cy, ang = np.empty(n), np.empty(n), np.empty(n)
    x, y, a = size / 2.0, size / 2.0, 0.0
    for t in range(n):
        if t:
            dx, dy = rng.integers(-20, 21, 2)
            da = rng.integers(-10, 11) * 0.5
            if not (margin <= x + dx <= size - margin):
                dx = -dx
            if not (margin <= y + dy <= size - margin):
                dy = -dy
            x, y, a = x + dx, y + dy, a + da
        cx[t], cy[t], ang[t] = x, y, a
    return cx, cy, ang


def db_poses(n: int, seed: int = 1, size: int = 4096, H: int = 480, W: int = 640):
    rng = np.random.default_rng(seed)
    margin = int(math.hypot(H, W) / 2) + 8
    cx = rng.uniform(margin, size - margin, n)
    cy = rng.uniform(margin, size - margin, n)
    ang = rng.integers(-360, 360, n) * 0.5
    return cx, cy, ang
