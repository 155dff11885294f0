"""Generates tests/golden/golden_ref.npz from oracle/_ref: the reference's OWN sources
(/root/reference/src/{correlation_flow,loop_closure,utils,map,frame}.cc, compiled unmodified by oracle/Makefile.ref against the
stand-in headers of oracle/ref_stubs).  These vectors are outputs of the reference's code run in this container; they pin the C
oracle (tests/test_oracle_ref.py, CPU) and the CUDA path (tests/test_gpu_parity.py, GPU box, where /root/reference is absent).

Run here (CPU container, /root/reference present):  python tests/golden/make_golden_ref.py
Inputs: the 8 seeded 640x480 u8 frames of golden_pairs.npz (tests/golden/make_golden.py; SURVEY.md Appendix C fixture).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import oracle_c as oc  # noqa: E402  (config structs / layout helpers only)
import oracle_ref as orf  # noqa: E402

COLS = ["kernel", "last", "cur", "not_large_rotation", "x", "y", "theta", "info0", "info1", "info2"]


def main():
    g = np.load(os.path.join(HERE, "golden_pairs.npz"))
    imgs_u8 = g["images"]
    out = {}
    imgs = [orf.normalize_u8(u) for u in imgs_u8]
    out["normalized_probe"] = np.stack([im[::40, ::40] for im in imgs])
    feats = {}
    rows = []
    for kernel in (0, 1):
        cfg = oc.make_cfg(kernel=kernel)
        cf = orf.CorrelationFlow(cfg)
        if kernel == 0:
            for i, im in enumerate(imgs):
                feats[i] = cf.compute_intermedium(im)
            out["fft_result_probe"] = np.stack([feats[i][0][::16, ::16] for i in range(len(imgs))])     # 16 x 40 complex probes
            out["fft_polar_probe"] = np.stack([feats[i][1][::19, ::16] for i in range(len(imgs))])
            out["fft_result_abs_sum"] = np.array([np.abs(feats[i][0]).astype(np.float64).sum() for i in range(len(imgs))])
            out["fft_polar_abs_sum"] = np.array([np.abs(feats[i][1]).astype(np.float64).sum() for i in range(len(imgs))])
        pairs = [(0, j) for j in range(1, len(imgs))] + [(2, 1), (3, 5)]
        for (a, b) in pairs:
            for mode in (1, 0):
                info, pose = cf.compute_pose(feats[a][0], imgs[b], feats[a][1], feats[b][1], bool(mode))
                rows.append([kernel, a, b, mode, *pose, *info])
        if kernel == 0:
            # LoopClosure::FindLoopClosure, the three overloads, through the reference's own Map / Frame
            thr = oc.LoopConfigC(30.0, 60.0, 0, 0.0)
            order = [3, 0, 5, 1, 4, 6]
            kfs = [(10 + k, feats[i][0], feats[i][1], float(k)) for k, i in enumerate(order)]
            q = 2
            r0 = cf.find_loop_closure(thr, imgs[q], feats[q][0], feats[q][1], 100, 50.0, kfs, mode=0)
            # mode 1 (all frames, id order): Map::AddFrame renames the first frame added to id 0 (map.cc:19-22)
            poses = np.array([[0.5 + 1.2 * k, 0.5 - 0.7 * k, 0.0] for k in range(len(order))])
            r1 = cf.find_loop_closure(thr, imgs[q], feats[q][0], feats[q][1], 100, 50.0, kfs, mode=1, poses=poses, grid_scale=2.0)
            r2 = cf.find_loop_closure(thr, imgs[q], feats[q][0], feats[q][1], 100, 50.0, kfs, mode=2, poses=poses, grid_scale=2.0,
                                      prior_pose=np.array([1.0, 0.2, 0.0]))
            # frame-gap and distance filters (loop_closure.cc:43-53)
            thr_f = oc.LoopConfigC(30.0, 60.0, 89, 47.5)
            r3 = cf.find_loop_closure(thr_f, imgs[q], feats[q][0], feats[q][1], 100, 50.0, kfs, mode=0)
            # duplicated keyframe: strict '>' keeps the first (loop_closure.cc:61)
            kfs_d = [(7, feats[1][0], feats[1][1], 0.0), (8, feats[1][0], feats[1][1], 1.0), (9, feats[5][0], feats[5][1], 2.0)]
            r4 = cf.find_loop_closure(thr, imgs[q], feats[q][0], feats[q][1], 100, 50.0, kfs_d, mode=0)
            out["scan_order"] = np.array(order)
            out["scan_query"] = np.array(q)
            out["scan_poses"] = poses
            for name, r in (("list", r0), ("all", r1), ("prior", r2), ("filtered", r3), ("dup", r4)):
                out["scan_" + name] = np.array([r["found"], r["index"], r["frame_id"], *r["relative_pose"], *r["response"]], np.float64)
        cf.close()
    out["pose_rows"] = np.array(rows, np.float64)
    out["pose_cols"] = np.array(COLS)
    # RotateArray / NormalizeDegree / ConvertMatToNormalizedArray through the reference's utils.cc
    degs = np.array([0.5, -3.0, 10.0, -45.5, 90.0, 180.0, 179.5, -352.5, 33.0, 123.5], np.float32)
    out["rot_degrees"] = degs
    out["rot_probe"] = np.stack([orf.rotate(imgs[1], d)[::24, ::32] for d in degs])
    angs = np.array([180.0, 190.0, -180.0, 359.5, -540.25, 0.0, 720.0, -179.99])
    out["nd_in"] = angs
    out["nd_out"] = np.array([orf.normalize_degree(a) for a in angs])
    np.savez_compressed(os.path.join(HERE, "golden_ref.npz"), **out)
    print("wrote golden_ref.npz")
    np.set_printoptions(linewidth=200, suppress=True)
    print(out["pose_rows"])
    for k in out:
        if k.startswith("scan_"):
            print(k, out[k])


if __name__ == "__main__":
    main()
