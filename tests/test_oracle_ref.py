"""CPU tests (-m "not gpu") that pin the C oracle to the REFERENCE'S OWN SOURCE TEXT.

oracle/_ref = /root/reference/src/{correlation_flow,loop_closure,utils,map,frame}.cc compiled unmodified against the stand-in headers
of oracle/ref_stubs (oracle/Makefile.ref).  tests/golden/golden_ref.npz holds its outputs on the seeded fixture frames
(tests/golden/make_golden_ref.py).  Bar: integer results exact, theta exact, info within 1e-6 relative (observed: bit-identical).
  * golden tests run everywhere (the GPU box has no /root/reference);
  * the live _ref-vs-C comparisons run wherever oracle/_ref/libnislam_ref.so exists or can be built.
"""
import os

import numpy as np
import pytest

import oracle_c as oc
import oracle_ref as orf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
H, W, D, CP = 480, 640, 720, 480
needs_ref = pytest.mark.skipif(not orf.available(), reason="oracle/_ref not built and /root/reference absent")


@pytest.fixture(scope="module")
def gref():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_ref.npz"))


@pytest.fixture(scope="module")
def frames(golden_pairs):
    return [oc.normalize_u8(u) for u in golden_pairs["images"]]


@pytest.fixture(scope="module")
def feats(frames):
    cfg = oc.make_cfg()
    return [oc.compute_intermedium(cfg, im) for im in frames]


def test_normalized_frames_and_features_match_reference(gref, frames, feats):
    for i, im in enumerate(frames):
        assert np.array_equal(im[::40, ::40], gref["normalized_probe"][i])            # utils.cc:110-118
        assert np.array_equal(feats[i][0][::16, ::16], gref["fft_result_probe"][i])     # correlation_flow.cc:91
        Pp = gref["fft_polar_probe"][i]
        assert np.abs(feats[i][1][::19, ::16] - Pp).max() <= 1e-6 * np.abs(Pp).max()    # :92-94
        assert abs(np.abs(feats[i][0]).astype(np.float64).sum() / gref["fft_result_abs_sum"][i] - 1) < 1e-7
        assert abs(np.abs(feats[i][1]).astype(np.float64).sum() / gref["fft_polar_abs_sum"][i] - 1) < 1e-7


def test_compute_pose_matches_reference_goldens(gref, frames, feats):
    """ComputePose, both modes, polynomial AND gaussian kernel, 9 pairs (correlation_flow.cc:97-243)."""
    rows = gref["pose_rows"]
    assert len(rows) == 36
    for r in rows:
        kernel, a, b, mode = (int(v) for v in r[:4])
        cfg = oc.make_cfg(kernel=kernel)
        info, pose, _ = oc.compute_pose(cfg, feats[a][0], frames[b], feats[a][1], feats[b][1], mode)
        assert np.array_equal(pose, r[4:7]), (r[:4], pose, r[4:7])                       # dx, dy exact; theta exact (no mod 2 pi needed)
        assert np.allclose(info, r[7:10], rtol=1e-6, atol=0), (r[:4], info, r[7:10])


def _c_scan(thr, frames, feats, q, kfs, order=None):
    cfg = oc.make_cfg()
    ks = kfs if order is None else [kfs[i] for i in order]
    r = oc.find_loop_closure(cfg, thr, frames[q], feats[q][1], 100, 50.0, ks)
    if order is not None and r["index"] >= 0:
        r["index"] = order[r["index"]]
    return r


def _same_scan(r, row):
    assert (int(r["found"]), r["index"], r["frame_id"]) == tuple(int(v) for v in row[:3]), (r, row)
    if r["index"] >= 0:
        assert np.array_equal(r["relative_pose"], row[3:6])
    assert np.allclose(r["response"], row[6:9], rtol=1e-6, atol=0)


def test_find_loop_closure_matches_reference_goldens(gref, frames, feats):
    """LoopClosure::FindLoopClosure through the reference's own Map / Frame (loop_closure.cc:10-73): explicit list, filters,
    first-wins on a duplicated keyframe, all frames in id order, 3x3 grid cells around a prior pose."""
    order = [int(i) for i in gref["scan_order"]]
    q = int(gref["scan_query"])
    kfs = [(10 + k, feats[i][0], feats[i][1], float(k)) for k, i in enumerate(order)]
    thr = oc.LoopConfigC(30.0, 60.0, 0, 0.0)
    _same_scan(_c_scan(thr, frames, feats, q, kfs), gref["scan_list"])
    _same_scan(_c_scan(oc.LoopConfigC(30.0, 60.0, 89, 47.5), frames, feats, q, kfs), gref["scan_filtered"])
    kfs_d = [(7, feats[1][0], feats[1][1], 0.0), (8, feats[1][0], feats[1][1], 1.0), (9, feats[5][0], feats[5][1], 2.0)]
    _same_scan(_c_scan(thr, frames, feats, q, kfs_d), gref["scan_dup"])
    # all frames: Map::AddFrame renames the first frame to id 0 (map.cc:19-22) and GetAllFrames iterates in id order (map.cc:51-56)
    kfs_all = [(0 if k == 0 else kf[0], kf[1], kf[2], kf[3]) for k, kf in enumerate(kfs)]
    _same_scan(_c_scan(thr, frames, feats, q, kfs_all, order=sorted(range(len(kfs_all)), key=lambda i: kfs_all[i][0])), gref["scan_all"])
    # prior pose (1.0, 0.2) at grid_scale 2 -> cell (0, 0); 3x3 neighbourhood = cells -1..1 (loop_closure.cc:19-28, map.cc:81-85)
    poses = gref["scan_poses"]
    cells = [(int(p[0] / 2.0), int(p[1] / 2.0)) for p in poses]
    near = [i for i, c in enumerate(cells) if abs(c[0]) <= 1 and abs(c[1]) <= 1]
    _same_scan(_c_scan(thr, frames, feats, q, kfs_all, order=near), gref["scan_prior"])


def test_utils_match_reference_goldens(gref, frames):
    for a, want in zip(gref["nd_in"], gref["nd_out"]):
        assert oc.lib().orc_normalize_degree(float(a)) == want                           # utils.cc:173-175
    for d, want in zip(gref["rot_degrees"], gref["rot_probe"]):
        assert np.array_equal(oc.rotate(frames[1], d)[::24, ::32], want), d              # utils.cc:154-161


# ---------------------------------------------------------------- live: the compiled reference beside the C restatement
@needs_ref
def test_ref_library_exports():
    l = orf.lib()
    for s in ("ref_create", "ref_destroy", "ref_compute_intermedium", "ref_compute_pose", "ref_find_loop_closure", "ref_rotate",
              "ref_normalize_u8", "ref_normalize_degree"):
        assert hasattr(l, s)


@needs_ref
def test_ref_equals_c_oracle_on_random_pairs(golden_pairs):
    """Seeded crops of a fresh canvas (not the golden frames): features, ComputePose in both modes, both kernels."""
    rng = np.random.default_rng(20261017)
    canvas = rng.random((700, 900)).astype(np.float32)
    k = np.array([1, 4, 6, 4, 1], np.float32) / 16
    for ax in (0, 1):
        canvas = sum(np.roll(canvas, s - 2, axis=ax) * k[s] for s in range(5))
    canvas = (canvas - canvas.min()) / (canvas.max() - canvas.min())
    def crop(y, x):
        return np.ascontiguousarray(np.rint(canvas[y:y + H, x:x + W] * 255).astype(np.uint8))
    a_u8, b_u8 = crop(100, 120), crop(91, 137)
    assert np.array_equal(oc.normalize_u8(a_u8), orf.normalize_u8(a_u8))
    a, b = oc.normalize_u8(a_u8), oc.normalize_u8(np.ascontiguousarray(np.rot90(b_u8, 2)))
    for kernel in (0, 1):
        cfg = oc.make_cfg(kernel=kernel)
        cf = orf.CorrelationFlow(cfg)
        Fa, Pa = oc.compute_intermedium(cfg, a)
        Fb, Pb = oc.compute_intermedium(cfg, b)
        Fr, Pr = cf.compute_intermedium(a)
        assert np.array_equal(Fa, Fr) and np.abs(Pa - Pr).max() <= 1e-6 * np.abs(Pa).max()
        for mode in (True, False):
            io, po, _ = oc.compute_pose(cfg, Fa, b, Pa, Pb, mode)
            ir, pr = cf.compute_pose(Fa, b, Pa, Pb, mode)
            assert np.array_equal(po, pr), (kernel, mode, po, pr)
            assert np.allclose(io, ir, rtol=1e-6, atol=0), (kernel, mode, io, ir)
        cf.close()


@needs_ref
def test_ref_invalid_kernel_throws_like_reference():
    cfg = oc.make_cfg(kernel=7)
    cf = orf.CorrelationFlow(cfg)
    z = np.zeros((H // 2 + 1, W), np.complex64)
    zp = np.zeros((D // 2 + 1, CP), np.complex64)
    with pytest.raises(ValueError):
        cf.compute_pose(z, np.zeros((H, W), np.float32), zp, zp, True)                  # correlation_flow.cc:168
    cf.close()
