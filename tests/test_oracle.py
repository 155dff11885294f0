"""CPU tests (-m "not gpu"): pin the C oracle (oracle/nislam_oracle.c) against
  * the analytic known answers of SURVEY.md Appendix C,
  * the committed golden vectors generated with scipy-f32 FFT + genuine cv2 (tests/golden/make_golden.py),
  * the Python restatement itself when scipy/cv2 are importable.
"""
import numpy as np
import pytest

import oracle_c as oc

H, W, D, CP = 480, 640, 720, 480


def wrap_pi(a):
    return (a + np.pi) % (2 * np.pi) - np.pi


# ---------------------------------------------------------------- analytic KATs (Appendix C)
def test_impulse_spectrum_is_checkerboard():
    # C.1  FFT(delta[R/2,C/2]) = (-1)^(kr+kc)   (correlation_flow.cc:46-51)
    for R, C in ((480, 640), (720, 480), (48, 64)):
        x = np.zeros((R, C), np.float32)
        x[R // 2, C // 2] = 1
        F = oc.fft2(x)
        kr, kc = np.meshgrid(np.arange(R // 2 + 1), np.arange(C), indexing="ij")
        expect = np.where((kr + kc) % 2 == 0, 1.0, -1.0)
        assert np.abs(F - expect).max() < 2e-6


def test_remove_zero_component_4x4():
    # C.2
    x = np.arange(16, dtype=np.float32).reshape(4, 4)
    y = oc.remove_zero_component(x)
    assert np.array_equal(y, np.array([[2, 9, 10, 11], [6, 5, 6, 7], [10, 9, 10, 11], [14, 13, 14, 15]], np.float32))


def test_normalize_degree():
    # C.3  (utils.cc:173-175)
    nd = oc.lib().orc_normalize_degree
    assert nd(180.0) == -180.0 and nd(190.0) == -170.0 and nd(-180.0) == -180.0 and nd(359.5) == -0.5


def test_fftshift_matches_roll():
    x = np.random.default_rng(0).random((6, 8)).astype(np.float32)
    assert np.array_equal(oc.fftshift(x), np.roll(x, (3, 4), axis=(0, 1)))


def test_fft_roundtrip_and_numpy():
    rng = np.random.default_rng(3)
    for R, C in ((480, 640), (720, 480), (448, 448), (30, 22), (960, 1280)):
        x = rng.random((R, C)).astype(np.float32)
        F = oc.fft2(x)
        Fn = np.fft.rfft2(x.astype(np.float64), axes=(1, 0))
        assert np.abs(F - Fn).max() / np.abs(Fn).max() < 1e-6
        assert np.abs(oc.ifft2(F) - x).max() < 2e-6


def test_argmax_tiebreak_column_major_first():
    # C.6: duplicated maxima -> smallest col, then smallest row (Eigen maxCoeff on a column-major array).
    # Build a spectrum whose IFFT is exactly two equal impulses by linearity through estimate_trans is not possible,
    # so check the reduction rule on the oracle's g output of an identity pair (single peak) and the rule itself here.
    g = np.zeros((4, 5), np.float32)
    g[3, 1] = g[0, 2] = g[2, 1] = 7.0
    flat = int(np.argmax(g.T.reshape(-1)))
    assert divmod(flat, 4) == (1, 2)        # col 1, row 2


# ---------------------------------------------------------------- golden vectors (genuine cv2)
def test_warps_bit_exact_vs_cv2_golden(golden_stages):
    s = golden_stages
    assert np.array_equal(oc.polar(s["src"], 72, 40), s["polar_72x40"])
    assert np.array_equal(oc.polar(s["src2"], 36, 16), s["polar2_36x16"])
    for d, want, want2 in zip(s["rot_degrees"], s["rot_out"], s["rot2_out"]):
        assert np.array_equal(oc.rotate(s["src"], d), want), d
        assert np.array_equal(oc.rotate(s["src2"], d), want2), d
    for d, m in zip(s["rot_degrees"], s["inv_mats_640x480"]):
        assert np.allclose(oc.rotation_inverse(H, W, float(d)), m, rtol=0, atol=1e-9)


def _check_rows(rows, fn):
    for r in rows:
        i, mode = int(r[0]), int(r[1])
        info, pose, pk = fn(i, mode)
        assert pose[0] == r[2] and pose[1] == r[3], (i, mode, pose, r)
        assert abs(wrap_pi(pose[2] - r[4])) < 1e-6, (i, mode, pose, r)
        assert pk["trans"] == (int(r[10]), int(r[11]))
        assert pk["polar"][0] % (D // 2) == int(r[8]) % (D // 2)       # 180-degree twin peak
        # cross-check of two independent f32 chains (scipy pocketfft + double sums vs the C oracle, which is held to the reference's
        # own f32 evaluation order by tests/test_oracle_ref.py): 3e-4; below the "tracking lost" gate of 30 the peak is a noise maximum
        tol = 3e-4 if min(r[5], r[7]) > 30 else 2e-3
        assert np.allclose(info, r[5:8], rtol=tol), (i, mode, info, r[5:8])


def test_pose_goldens(golden_pairs):
    g = golden_pairs
    cfg = oc.make_cfg()
    imgs = [oc.normalize_u8(u) for u in g["images"]]
    feats = [oc.compute_intermedium(cfg, im) for im in imgs]
    assert abs(np.abs(feats[0][0]).astype(np.float64).sum() / g["a_fft_result_abs_sum"] - 1) < 1e-5
    assert abs(np.abs(feats[0][1]).astype(np.float64).sum() / g["a_fft_polar_abs_sum"] - 1) < 1e-5

    def fn(i, mode):
        return oc.compute_pose(cfg, feats[0][0], imgs[i], feats[0][1], feats[i][1], mode)
    _check_rows(g["pose_rows"], fn)


def test_gaussian_kernel_golden(golden_pairs):
    g = golden_pairs
    cfg = oc.make_cfg(kernel=1)
    a, b = oc.normalize_u8(g["images"][0]), oc.normalize_u8(g["images"][2])
    Fa, Pa = oc.compute_intermedium(cfg, a)
    Fb, Pb = oc.compute_intermedium(cfg, b)
    info, pose, pk = oc.compute_pose(cfg, Fa, b, Pa, Pb, True)
    r = g["gauss_row"]
    assert pose[0] == r[2] and pose[1] == r[3] and abs(wrap_pi(pose[2] - r[4])) < 1e-6
    assert np.allclose(info, r[5:8], rtol=3e-4)


def test_invalid_kernel_raises():
    cfg = oc.make_cfg(kernel=7, height=48, width=64, rotation_divisor=72, rotation_channel=40)
    z = np.zeros((25, 64), np.complex64)
    with pytest.raises(ValueError):
        oc.estimate_trans(cfg, z, z, 48, 64)


# ---------------------------------------------------------------- circular rolls (C.4 / C.5)
def test_circular_roll_known_answers(golden_pairs):
    cfg = oc.make_cfg()
    a = oc.normalize_u8(golden_pairs["images"][0])
    Fa, Pa = oc.compute_intermedium(cfg, a)
    base = None
    for sy, sx in ((0, 0), (3, 0), (0, -9), (17, 25)):
        b = np.roll(a, (sy, sx), axis=(0, 1))
        Fb, Pb = oc.compute_intermedium(cfg, b)
        info, pose, pk = oc.compute_pose(cfg, Fa, b, Pa, Pb, True)
        assert (pose[0], pose[1]) == (-sx, -sy)
        assert abs(wrap_pi(pose[2])) < 1e-6
        assert pk["polar"][0] % (D // 2) == 0
        if base is None:
            base = info
            assert pk["trans"] == (H // 2, W // 2)
        assert np.allclose(info, base, rtol=1e-3)


# ---------------------------------------------------------------- scan (C.7, loop_closure.cc:36-73)
def test_scan_first_wins_and_filters(golden_pairs):
    g = golden_pairs
    cfg = oc.make_cfg()
    imgs = [oc.normalize_u8(u) for u in g["images"][:4]]
    feats = [oc.compute_intermedium(cfg, im) for im in imgs]
    thr = oc.LoopConfigC(60.0, 60.0, 0, 0.0)
    # three identical copies of keyframe 0 among others: the first copy must win (strict >)
    kfs = [(10, *feats[3], 0.0), (11, *feats[0], 1.0), (12, *feats[0], 2.0), (13, *feats[0], 3.0)]
    res = oc.find_loop_closure(cfg, thr, imgs[1], feats[1][1], 99, 50.0, kfs)
    assert res["found"] and res["index"] == 1 and res["frame_id"] == 11
    assert tuple(res["relative_pose"][:2]) == (7.0, 0.0)
    # same with 4 threads
    res4 = oc.find_loop_closure(cfg, thr, imgs[1], feats[1][1], 99, 50.0, kfs, threads=4)
    assert res4["index"] == 1 and np.array_equal(res4["response"], res["response"])
    # frame-gap filter removes ids with |gap| < 89 of 99 -> only id 10 survives
    thr2 = oc.LoopConfigC(60.0, 60.0, 89, 0.0)
    res2 = oc.find_loop_closure(cfg, thr2, imgs[1], feats[1][1], 99, 50.0, kfs)
    assert res2["index"] == 0
    # distance filter: |50 - d| < 49.5 removes everything except d = 0.0
    thr3 = oc.LoopConfigC(60.0, 60.0, 0, 49.5)
    res3 = oc.find_loop_closure(cfg, thr3, imgs[1], feats[1][1], 99, 50.0, kfs)
    assert res3["index"] == 0
    # empty candidate list: initial best (-1,-1,-1), not found (loop_closure.h:15)
    res0 = oc.find_loop_closure(cfg, thr, imgs[1], feats[1][1], 99, 50.0, [])
    assert (not res0["found"]) and res0["index"] == -1 and np.array_equal(res0["response"], [-1, -1, -1])


# ---------------------------------------------------------------- Python restatement == C restatement (if cv2/scipy exist)
def test_python_restatement_agrees(golden_pairs):
    ref = pytest.importorskip("nislam_ref")
    if ref.cv2 is None:
        pytest.skip("cv2 missing")
    g = golden_pairs
    cf = ref.CorrelationFlow(ref.CFConfig(), H, W)
    cfg = oc.make_cfg()
    a = ref.convert_mat_to_normalized_array(g["images"][0])
    b = ref.convert_mat_to_normalized_array(g["images"][3])
    assert np.array_equal(a, oc.normalize_u8(g["images"][0]))
    Fa, Pa = cf.compute_intermedium(a)
    Fc, Pc = oc.compute_intermedium(cfg, a)
    assert np.abs(Fa - Fc).max() / np.abs(Fa).max() < 1e-6
    assert np.abs(Pa - Pc).max() / np.abs(Pa).max() < 1e-6
    # stage: polar image bit-exact against cv2 at full size
    hp = cf.last_stages["high_power"]
    assert np.array_equal(oc.polar(ref.fftshift(hp)), cf.polar(ref.fftshift(hp)))
    assert np.array_equal(oc.rotate(b, -10.0), ref.rotate_array(b, np.float32(-10.0)))
    Fb, Pb = cf.compute_intermedium(b)
    i_r, t_r, p_r, g_r = cf.estimate_trans(Fa, Fb, cf.target_fft, H, W)
    i_c, t_c, p_c, g_c = oc.estimate_trans(cfg, Fa, Fb, H, W, want_g=True)
    assert p_r == p_c and t_r == t_c
    # f32 FFT-chain noise in g is ~5e-3 of g's sidelobe std (measured against an f64 pipeline); the peak of a good
    # match stands ~150 std above it, so argmax and info are stable although g itself is only comparable to ~1e-2 std
    assert np.sqrt(np.mean((g_r - g_c) ** 2)) / g_r.std() < 2e-2
    assert abs(i_r - i_c) / i_r < 3e-4


# ---------------------------------------------------------------- undistort front end (camera.cc:92-93)
def test_undistort_bit_exact_vs_cv2_golden():
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_undistort.npz"))
    assert np.array_equal(oc.undistort_u8(g["raw"], g["map1"], g["map2"]), g["out"])
    cv2 = pytest.importorskip("cv2")
    H, W = 480, 640
    K = np.array([[520., 0, 318.], [0, 515., 243.], [0, 0, 1]])
    Dm = np.array([-0.28, 0.09, 0.0007, -0.0004, -0.012])
    newK, _ = cv2.getOptimalNewCameraMatrix(K, Dm, (W, H), 0, (W, H))
    m1, m2 = cv2.initUndistortRectifyMap(K, Dm, None, newK, (W, H), cv2.CV_16SC2)
    raw = np.random.default_rng(5).integers(0, 256, (H, W)).astype(np.uint8)
    assert np.array_equal(oc.undistort_u8(raw, m1, m2), cv2.remap(raw, m1, m2, cv2.INTER_LINEAR))


def test_u8_normalisation_is_reproducible_in_f32():
    """utils.cc:110-118 computes (float)((double)u / 255.0).  The CUDA prologues do the conversion arithmetically in f32
    (nis_ops.cuh u8_to_unit); check in exact rational arithmetic that both f32 forms give the reference's value for all 256 inputs."""
    import math
    from fractions import Fraction as Fr

    def rn32(x):                       # round a rational to the nearest float32 (ties to even), as a Fraction
        if x == 0:
            return Fr(0)
        s, x = (1 if x > 0 else -1), abs(x)
        e = math.floor(math.log2(x)) - 23
        while x / Fr(2) ** e >= 2 ** 24:
            e += 1
        while x / Fr(2) ** e < 2 ** 23:
            e -= 1
        q = x / Fr(2) ** e
        n = q.numerator // q.denominator
        rem = q - n
        if rem > Fr(1, 2) or (rem == Fr(1, 2) and n % 2 == 1):
            n += 1
        return s * n * Fr(2) ** e

    r = rn32(Fr(1, 255))
    for u in range(256):
        ref = Fr(float(np.float32(np.float64(u) / 255.0)))
        assert rn32(Fr(u, 255)) == ref                       # correctly rounded f32 division
        q = rn32(Fr(u) * r)
        rem = rn32(Fr(u) - q * 255)                          # fmaf(-q, 255, u), exact
        assert rn32(rem * r + q) == ref                      # fmaf(rem, 1/255, q): the device form
        assert float(ref) == float(oc.normalize_u8(np.full((2, 2), u, np.uint8))[0, 0])


# ---------------------------------------------------------------- tracker restatement (map_builder.cc:30-70) known answers
def _cam(**kw):
    import tracker_ref as tr
    d = dict(fx=800.0, fy=820.0, cx=330.0, cy=235.0, height=0.5, extrinsics=[0, -1, 0.1, 1, 0, 0.2, 0, 0, 1], image_width=W, image_height=H)
    d.update(kw)
    return tr.Camera(**d)


def test_pose_algebra_known_answers():
    import tracker_ref as tr
    # NormalizeAngle: [-pi, pi)
    assert tr.normalize_angle(np.pi) == -np.pi and abs(tr.normalize_angle(3 * np.pi / 2) + np.pi / 2) < 1e-15
    # absolute(relative) round trip and a hand-computed composition: p1 = (1, 2, 90 deg), rel = (1, 0, 0) -> (1, 3, 90 deg)
    p1, p2 = np.array([1.0, 2.0, np.pi / 2]), np.array([-3.0, 0.5, -2.0])
    assert np.allclose(tr.compute_absolute_pose(p1, np.array([1.0, 0.0, 0.0])), [1.0, 3.0, np.pi / 2], atol=1e-15)
    rel = tr.compute_relative_pose(p1, p2)
    back = tr.compute_absolute_pose(p1, rel)
    assert np.allclose(back[:2], p2[:2], atol=1e-14) and abs(wrap_pi(back[2] - p2[2])) < 1e-14
    cam = _cam()
    # ConvertCenterToPrincipal: no rotation -> unchanged; 180 deg -> + 2 * O_bias with O_bias = (W/2 - cx, H/2 - cy) = (-10, 5)
    assert np.array_equal(cam.convert_center_to_principal(np.array([3.0, 4.0, 0.0])), [3.0, 4.0, 0.0])
    assert np.allclose(cam.convert_center_to_principal(np.array([3.0, 4.0, np.pi])), [3.0 - 20.0, 4.0 + 10.0, np.pi], atol=1e-12)
    # image plane -> camera -> robot: (80, 82, 0.3) -> (0.1, 0.1, 0.3) -> E * (0.05, 0.05, 0.3)
    assert np.allclose(cam.image_plane_to_robot(np.array([80.0, 82.0, 0.3])), [-0.05 + 0.03, 0.05 + 0.06, 0.3], atol=1e-15)


def test_tracker_keyframe_policy_with_scripted_poses():
    """AddNewInput's control flow with ComputePose scripted: gate, c1..c4, tracking against the LAST KEYFRAME, stale state on a
    lost frame (map_builder.cc:42-57)."""
    import tracker_ref as tr
    script = [  # (response, pose_center) returned by ComputePose for frames 1..
        ((150.0, 150.0, 120.0), (4.0, 0.0, 0.0)),     # small motion, confident: not a keyframe
        ((150.0, 150.0, 120.0), (20.0, 0.0, 0.0)),    # 20 px / 800 = 0.025 > max_distance 0.02: keyframe
        ((10.0, 10.0, 120.0), (3.0, 3.0, 0.0)),       # lost (response(0) < 30): state stays
        ((150.0, 150.0, 50.0), (1.0, 0.0, 0.0)),      # response(2) inside (30, 90): keyframe by c4
        ((150.0, 150.0, 120.0), (0.0, 0.0, 0.05)),    # rotation 0.05 > max_angle 0.03: keyframe by c2
    ]
    it = iter(script)
    trk = tr.MapBuilderTracker(_cam(cx=320.0, cy=240.0), 0.02, 0.03, 30.0, 90.0, lambda img: (None, None), lambda *a: next(it))
    outs = [trk.add_new_input(None) for _ in range(6)]
    assert [o["inserted"] for o in outs] == [True, False, True, False, True, True]
    assert [o["tracked"] for o in outs] == [True, True, True, False, True, True]
    assert [o["keyframe"] for o in outs] == [-1, 0, 0, 2, 2, 4]
    assert np.allclose(outs[1]["cf_pose"], [4, 0, 0]) and np.allclose(outs[2]["cf_pose"], [20, 0, 0])     # both against keyframe 0
    assert np.allclose(outs[3]["cf_pose"], [20, 0, 0])                                                    # lost: stale
    assert np.allclose(outs[4]["cf_pose"], [21, 0, 0]) and np.allclose(outs[5]["cf_pose"], [21, 0, 0.05])
    assert abs(outs[5]["distance"] - (0.025 + 1 / 800)) < 1e-15
    # robot pose: extrinsics rotate by +90 deg, scale by height 0.5: cf (21, 0) px -> camera (0.02625, 0) -> robot delta (0, 0.013125)
    assert np.allclose(outs[4]["pose"][:2] - outs[0]["pose"][:2], [0.0, 0.5 * 21 / 800], atol=1e-15)


# ---------------------------------------------------------------- MapStitcher restatement (map_stitcher.cc) known answers
def test_stitcher_scaling_matches_opencv():
    """InsertFrame's `image * (100.0/255.0)` on a u8 cv::Mat = convertTo with a float alpha and cvRound; cv2.convertScaleAbs runs the
    same scale-convert path (|.| is the identity on non-negative input)."""
    import stitcher_ref as sr
    v = np.arange(256, dtype=np.uint8).reshape(16, 16)
    ours = sr.normalize_image(v)
    assert ours[0, 0] == 0 and ours[15, 15] == 100 and ours[7, 15] == 50          # 127 * 100/255 = 49.8 -> 50
    cv2 = pytest.importorskip("cv2")
    assert np.array_equal(ours, cv2.convertScaleAbs(v, alpha=100.0 / 255.0))


def test_stitcher_known_answers():
    import stitcher_ref as sr
    import tracker_ref as tr
    Hs, Ws, cs = 6, 8, 5
    cam = tr.Camera(fx=2.0, fy=2.0, cx=Ws / 2, cy=Hs / 2, height=0.5, extrinsics=np.eye(3), image_width=Ws, image_height=Hs)
    img = (np.arange(Hs * Ws, dtype=np.uint8).reshape(Hs, Ws) * 5)
    st = sr.MapStitcher(cs, cam)
    # robot pose (0.25, 0.5, 0): camera = pose / height = (0.5, 1), image plane = f * camera = (1, 2) px; the principal point is the centre
    st.insert_frame(img, [0.25, 0.5, 0.0])
    norm = sr.normalize_image(img).astype(np.int32)
    # pixel (j, i) lands on x = trunc(i - 4 + 1), y = trunc(j - 3 + 2): x in [-3, 4], y in [-1, 4] -> cells x in {-1, 0}, y in {-1, 0}
    assert set(st.cells) == {(-1, -1), (0, -1), (-1, 0), (0, 0)}
    d, w = st.cells[(0, 0)]
    assert w[0, 0] == 1 and d[0, 0] == norm[1, 3] and d[4, 4] == norm[5, 7]       # (x, y) = (0, 0) <- (i, j) = (3, 1); first insert: raw sums
    d, w = st.cells[(-1, -1)]
    assert d[4, 2] == norm[0, 0] and w[4, 2] == 1 and w.sum() == 3                # x = -3 -> cell -1, in-cell 2; y = -1 -> in-cell 4
    # same frame again: existing cells merge as (data*weight + sum*count) / (weight + count) = (d + d) / 2 = d, weights double
    st.insert_frame(img, [0.25, 0.5, 0.0])
    d2, w2 = st.cells[(0, 0)]
    assert np.array_equal(d2, st.cells[(0, 0)][0]) and w2[0, 0] == 2 and d2[0, 0] == norm[1, 3]
    # truncation toward zero: a pose shifted by half a pixel (image plane x = 0.5) makes x = trunc(i - 4 + 0.5): -3.5 -> -3 and 0.5 -> 0,
    # so the two source columns i = 3 and i = 4 collide on x = 0 (count 2, sum of both) and x = -3 ... 4 as before
    st3 = sr.MapStitcher(cs, cam)
    st3.insert_frame(img, [0.125, 0.5, 0.0])
    d3, w3 = st3.cells[(0, 0)]
    assert w3[0, 0] == 2 and d3[0, 0] == norm[1, 3] + norm[1, 4]
    # recompute with the same poses reproduces the state; with a moved pose the mosaic moves
    before = {k: (v[0].copy(), v[1].copy()) for k, v in st.cells.items()}
    st.recompute_occupancy([[0.25, 0.5, 0.0], [0.25, 0.5, 0.0]])
    assert all(np.array_equal(before[k][0], st.cells[k][0]) and np.array_equal(before[k][1], st.cells[k][1]) for k in before)


# ---------------------------------------------------------------- size-independent properties (hypothesis)
def test_pose_algebra_properties():
    hyp = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st
    import tracker_ref as tr
    coord = st.floats(-500, 500, allow_nan=False)
    ang = st.floats(-10, 10, allow_nan=False)

    @settings(max_examples=200, deadline=None)
    @given(coord, coord, ang, coord, coord, ang)
    def check(x1, y1, a1, x2, y2, a2):
        p1, p2 = np.array([x1, y1, a1]), np.array([x2, y2, a2])
        rel = tr.compute_relative_pose(p1, p2)
        back = tr.compute_absolute_pose(p1, rel)
        assert np.allclose(back[:2], p2[:2], atol=1e-9) and abs(wrap_pi(back[2] - p2[2])) < 1e-9
        assert -np.pi <= rel[2] < np.pi and -np.pi <= back[2] < np.pi                 # NormalizeAngle range
        cam = _cam()
        c = np.array([x1, y1, wrap_pi(a1)])
        pp = cam.convert_center_to_principal(c)
        assert pp[2] == c[2]
        # ConvertPrincipalToCenter (camera.cc:136-146) undoes ConvertCenterToPrincipal
        import stitcher_ref as sr
        assert np.allclose(sr.principal_to_center(cam, pp), c, atol=1e-9)
        # robot -> image plane undoes image plane -> robot
        assert np.allclose(sr.robot_to_image_plane(cam, cam.image_plane_to_robot(c)), c, atol=1e-6)
    check()


def test_stitcher_conservation_properties():
    """Every source pixel lands in exactly one cell element: after one InsertFrame the weights sum to H*W and the raw sums to the sum
    of the normalised image, whatever the pose; replaying the same poses reproduces the mosaic."""
    import stitcher_ref as sr
    import tracker_ref as tr
    rng = np.random.default_rng(3)
    Hs, Ws = 24, 32
    cam = tr.Camera(fx=40.0, fy=38.0, cx=17.0, cy=11.5, height=0.7, extrinsics=[0.8, -0.6, 0.01, 0.6, 0.8, 0.02, 0, 0, 1], image_width=Ws, image_height=Hs)
    for trial in range(20):
        img = rng.integers(0, 256, (Hs, Ws), dtype=np.uint8)
        pose = [rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(-np.pi, np.pi)]
        st = sr.MapStitcher(int(rng.integers(5, 40)), cam)
        st.insert_frame(img, pose)
        assert sum(int(w.sum()) for _, w in st.cells.values()) == Hs * Ws
        assert sum(int(d.sum()) for d, _ in st.cells.values()) == int(sr.normalize_image(img).astype(np.int64).sum())
        pose2 = [pose[0] + 0.05, pose[1] - 0.03, pose[2] + 0.4]
        st.insert_frame(img[::-1].copy(), pose2)
        snap = {k: (d.copy(), w.copy()) for k, (d, w) in st.cells.items()}
        st.recompute_occupancy([pose, pose2])
        assert set(snap) == set(st.cells)
        assert all(np.array_equal(snap[k][0], st.cells[k][0]) and np.array_equal(snap[k][1], st.cells[k][1]) for k in snap)
