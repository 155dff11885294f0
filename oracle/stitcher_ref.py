"""CPU restatement of MapStitcher (src/map_stitcher.cc:14-145) -- TEST INFRASTRUCTURE ONLY.  Parity unpinned by the reference (no
tests); pinned by the known answers in tests/test_oracle.py and, for the u8 scaling, by genuine OpenCV where cv2 is importable.

  InsertFrame          :14-22   norm = image * (100.0/255.0) as a u8 cv::Mat (float multiply, cvRound), kept in _raw_images
  ComputeCellPosition  :24-34   floor division into cells of cell_size
  AddImageToOccupancy  :36-133  pose -> image plane -> centre; x = (int)(Wx(i) + Hx(j)), y = (int)(Wy(i) + Hy(j)) (truncation toward
                                zero); per-frame sums / counts per cell; merge: an existing cell gets (data*weight + sum*count) / (weight
                                + count) in int arithmetic, a new cell takes the raw sums (not divided) -- both kept as they are
  RecomputeOccupancy   :135-145 clear, replay every stored frame (here: in insertion order; the reference's order is that of an
                                unordered_map keyed by pointer)
Cells are dicts {(cell_x, cell_y): (data[cs, cs] int32, weight[cs, cs] int32)} indexed [in-cell y, in-cell x].
"""
import math

import numpy as np


def normalize_image(image_u8):
    scale = np.float32(100.0 / 255.0)
    return np.clip(np.rint(image_u8.astype(np.float32) * scale), 0, 255).astype(np.uint8)      # np.rint = round half to even = cvRound


def robot_to_image_plane(cam, robot_pose):          # camera.cc:211-222, :177-194, :233-241
    c = np.linalg.inv(cam.E) @ np.asarray(robot_pose, np.float64)
    c[0] /= cam.height
    c[1] /= cam.height
    return np.array([cam.fx * c[0], cam.fy * c[1], c[2]])


def principal_to_center(cam, p):                    # camera.cc:136-146
    c, s = math.cos(p[2]), math.sin(p[2])
    R = np.array([[c, -s], [s, c]])
    o_bias = np.array([cam.W * 0.5 - cam.cx, cam.H * 0.5 - cam.cy])
    out = np.array(p, np.float64)
    out[:2] = p[:2] - (np.eye(2) - R) @ o_bias
    return out


class MapStitcher:
    def __init__(self, cell_size, camera):
        self.cs, self.cam = int(cell_size), camera
        self.raw = []                                # normalised images in insertion order
        self.cells = {}

    def insert_frame(self, image_u8, robot_pose):
        self.raw.append(normalize_image(image_u8))
        self.add_image_to_occupancy(len(self.raw) - 1, robot_pose)

    def add_image_to_occupancy(self, slot, robot_pose):
        data = self.raw[slot].astype(np.int64)
        H, W = data.shape
        ip = principal_to_center(self.cam, robot_to_image_plane(self.cam, robot_pose))
        c, s = math.cos(ip[2]), math.sin(ip[2])
        r00, r01, r10, r11 = c, -s, s, c
        w_idx = np.arange(W, dtype=np.float64) - W / 2.0
        h_idx = np.arange(H, dtype=np.float64) - H / 2.0
        wx, wy = r00 * w_idx + ip[0], r10 * w_idx + ip[1]
        hx, hy = r01 * h_idx, r11 * h_idx
        x = np.trunc(wx[None, :] + hx[:, None]).astype(np.int64)          # [j, i]
        y = np.trunc(wy[None, :] + hy[:, None]).astype(np.int64)
        cs = self.cs
        cxs, cys = np.floor_divide(x, cs), np.floor_divide(y, cs)
        inx, iny = x - cxs * cs, y - cys * cs
        tmp = {}
        for key in set(zip(cxs.ravel().tolist(), cys.ravel().tolist())):
            m = (cxs == key[0]) & (cys == key[1])
            d = np.zeros((cs, cs), np.int64)
            w = np.zeros((cs, cs), np.int64)
            np.add.at(d, (iny[m], inx[m]), data[m])
            np.add.at(w, (iny[m], inx[m]), 1)
            tmp[key] = (d, w)
        for key, (d, w) in tmp.items():
            if key in self.cells:
                od, ow = self.cells[key]
                nd = _wrap32(od.astype(np.int64) * ow + d * w)
                nw = ow.astype(np.int64) + w
                q = np.where(nw >= 1, _trunc_div(nd, np.maximum(nw, 1)), nd)
                self.cells[key] = (q.astype(np.int32), nw.astype(np.int32))
            else:
                self.cells[key] = (d.astype(np.int32), w.astype(np.int32))

    def recompute_occupancy(self, robot_poses):
        self.cells = {}
        for slot, p in enumerate(robot_poses):
            self.add_image_to_occupancy(slot, p)


def _wrap32(v):
    return ((v + 2 ** 31) % 2 ** 32) - 2 ** 31


def _trunc_div(a, b):                                # C++ int division: toward zero
    return (np.abs(a) // b) * np.sign(a)
