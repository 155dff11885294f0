"""Generates the committed golden fixtures from oracle/nislam_ref.py (scipy pocketfft f32 + the genuine cv2 4.13).

Run here (CPU container):  python tests/golden/make_golden.py
The reference (sair-lab/ni-slam) has no tests or vectors and cannot be built in this image, so these
vectors pin the *restatement* with genuine OpenCV arithmetic for the two warps; they are not outputs of the
reference binary ("parity unpinned").
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import cv2  # noqa: E402
import nislam_ref as ref  # noqa: E402

H, W = 480, 640


def main():
    canvas = ref.make_canvas(0)
    cases = [(7, 0, 0), (11, -6, -4.5), (20, 13, 10), (-30, 22, 7.5), (5, 5, -33), (-41, -17, 152.0)]
    imgs = [ref.crop(canvas, 640, 480, 0)]
    for dx, dy, ang in cases:
        imgs.append(ref.crop(canvas, 640 + dx, 480 + dy, ang))
    imgs.append(ref.crop(canvas, 1140, 830, 0))        # no overlap with keyframe a, partly off-canvas
    imgs = np.stack(imgs)
    assert hashlib.sha1(imgs[0].tobytes()).hexdigest().startswith("a9e89e5362ad6c3f")

    cf = ref.CorrelationFlow(ref.CFConfig(), H, W)
    a = ref.convert_mat_to_normalized_array(imgs[0])
    Fa, Pa = cf.compute_intermedium(a)
    out = dict(images=imgs, cases=np.array(cases, np.float64))
    # checksums of the keyframe's features (layout independent: plain sums)
    out["a_fft_result_abs_sum"] = np.float64(np.abs(Fa).astype(np.float64).sum())
    out["a_fft_polar_abs_sum"] = np.float64(np.abs(Pa).astype(np.float64).sum())
    out["a_polar_image"] = cf.last_stages["polar"][::16, ::16].copy()       # subsampled 45x30 probe of the polar image
    rows = []
    for i in range(1, imgs.shape[0]):
        b = ref.convert_mat_to_normalized_array(imgs[i])
        Fb, Pb = cf.compute_intermedium(b)
        for mode in (1, 0):
            info, pose = cf.compute_pose(Fa, b, Pa, Pb, bool(mode))
            pk = cf.last_peaks
            rows.append([i, mode, pose[0], pose[1], pose[2], info[0], info[1], info[2], pk["polar"][0], pk["polar"][1],
                         pk["trans"][0], pk["trans"][1], pk["hyp"], pk["degree"]])
    out["pose_rows"] = np.array(rows, np.float64)
    out["pose_cols"] = np.array(["img", "not_large_rotation", "x", "y", "theta", "info0", "info1", "info2", "polar_row",
                                 "polar_col", "trans_row", "trans_col", "hyp", "degree"])
    # gaussian kernel variant on one pair
    cfg_g = ref.CFConfig(kernel=1)
    cfg = ref.CorrelationFlow(cfg_g, H, W)
    b = ref.convert_mat_to_normalized_array(imgs[2])
    Fb, Pb = cfg.compute_intermedium(b)
    info, pose = cfg.compute_pose(Fa, b, Pa, Pb, True)
    out["gauss_row"] = np.array([2, 1, *pose, *info, *cfg.last_peaks["polar"], *cfg.last_peaks["trans"]], np.float64)
    np.savez_compressed(os.path.join(HERE, "golden_pairs.npz"), **out)

    # --- stage-level vectors from genuine cv2 on small seeded inputs -------------------------------------------
    st = {}
    rng = np.random.default_rng(1234)
    src = rng.random((48, 64)).astype(np.float32)
    st["src"] = src
    cfs = ref.CorrelationFlow(ref.CFConfig(rotation_divisor=72, rotation_channel=40), 48, 64)
    st["polar_72x40"] = cfs.polar(src)
    degs = np.array([0.5, -3.0, 10.0, -45.5, 90.0, 180.0, 179.5, -352.5, 33.0, 123.5], np.float32)
    st["rot_degrees"] = degs
    st["rot_out"] = np.stack([ref.rotate_array(src, d) for d in degs])
    src2 = rng.random((30, 22)).astype(np.float32)      # odd-ish non-square case, H < W swapped
    st["src2"] = src2
    cfs2 = ref.CorrelationFlow(ref.CFConfig(rotation_divisor=36, rotation_channel=16), 30, 22)
    st["polar2_36x16"] = cfs2.polar(src2)
    st["rot2_out"] = np.stack([ref.rotate_array(src2, d) for d in degs])
    # full-size inverse rotation matrices (doubles) as cv2 computes them: M = getRotationMatrix2D, then invertAffine
    mats = []
    for d in degs:
        m = cv2.getRotationMatrix2D((W / 2.0, H / 2.0), float(d), 1.0)
        mats.append(cv2.invertAffineTransform(m).reshape(-1))
    st["inv_mats_640x480"] = np.stack(mats)
    np.savez_compressed(os.path.join(HERE, "golden_stages.npz"), **st)
    # --- undistort front end: Camera::UndistortImage = remap with the CV_16SC2 maps of initUndistortRectifyMap (camera.cc:45-47,92-93)
    und = {}
    hs, ws = 96, 128
    Ks = np.array([[104.0, 0, 63.2], [0, 103.0, 48.7], [0, 0, 1]])
    Ds = np.array([-0.28, 0.09, 0.0007, -0.0004, -0.012])
    newK, _ = cv2.getOptimalNewCameraMatrix(Ks, Ds, (ws, hs), 0, (ws, hs))
    m1, m2 = cv2.initUndistortRectifyMap(Ks, Ds, None, newK, (ws, hs), cv2.CV_16SC2)
    raw = np.random.default_rng(77).integers(0, 256, (hs, ws)).astype(np.uint8)
    und.update(K=Ks, D=Ds, map1=m1, map2=m2, raw=raw, out=cv2.remap(raw, m1, m2, cv2.INTER_LINEAR))
    np.savez_compressed(os.path.join(HERE, "golden_undistort.npz"), **und)
    print("wrote golden_pairs.npz, golden_stages.npz, golden_undistort.npz")
    print(out["pose_rows"])


if __name__ == "__main__":
    main()
