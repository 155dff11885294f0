"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI (ni_slam_b200.api -> libnislam.so), against
the CPU oracle (oracle/nislam_oracle.c) on the same seeded inputs and against the committed golden vectors.

Parity bar: translation-stage (row, col) bit-exact; polar row equal mod D/2 (the polar peak has an
exact 180-degree twin); dx, dy exact; theta equal mod 2 pi to 1e-6; info within 3e-4 relative (two f32
FFT chains agree to ~6e-5 on info; measured between the scipy and the C restatement); scan winner identical.
"""
import numpy as np
import pytest

import oracle_c as oc

pytestmark = pytest.mark.gpu

H, W, D, CP = 480, 640, 720, 480
INFO_RTOL = 3e-4
# The polar-stage confidence is the noisiest output of the path: T/(Kzz + lambda) amplifies f32 rounding where Kzz ~ -lambda and the
# polar peak has an exact twin D/2 rows away, so even the two CPU restatements (C port vs scipy-f32) differ by up to 1.8e-4 on it over the
# seeded random pairs below (measured; info_trans agrees to 3e-5).  It only gates thresholds of 30/60 on values of ~100.
INFO_ROT_RTOL = 1e-3


def wrap_pi(a):
    return (a + np.pi) % (2 * np.pi) - np.pi


def info_close(got, want):
    """(info_trans, info_trans, info_rot): translation-stage confidence within INFO_RTOL, polar-stage within INFO_ROT_RTOL.  The oracle
    sums GetInfo in f32 in Eigen's packet order like the compiled reference (oracle/_ref); that order alone moves the polar confidence
    by up to ~3e-4 against an exact sum, which is what the GPU's double accumulators are close to."""
    got, want = np.asarray(got), np.asarray(want)
    return bool(np.allclose(got[:2], want[:2], rtol=INFO_RTOL) and np.allclose(got[2], want[2], rtol=INFO_ROT_RTOL))


@pytest.fixture(scope="module")
def cf():
    import ni_slam_b200 as nis
    c = nis.CorrelationFlow(nis.CFConfig(), H, W)
    yield c
    c.close()


@pytest.fixture(scope="module")
def cfg():
    return oc.make_cfg()


@pytest.fixture(scope="module")
def imgs(golden_pairs):
    return golden_pairs["images"]


# ------------------------------------------------------------------ stages
def test_fft2_and_ifft2_match_oracle(cf):
    rng = np.random.default_rng(0)
    for which, (R, C) in enumerate(((H, W), (D, CP))):
        x = rng.random((R, C)).astype(np.float32)
        F = cf.debug_fft2(x, which)
        Fo = oc.fft2(x)
        assert np.abs(F - Fo).max() / np.abs(Fo).max() < 1e-6
        xb = cf.debug_ifft2(F, which)
        assert np.abs(xb - x).max() < 2e-6
        assert np.abs(xb - oc.ifft2(Fo)).max() < 2e-6


def test_impulse_spectrum_known_answer(cf):
    x = np.zeros((H, W), np.float32)
    x[H // 2, W // 2] = 1
    F = cf.debug_fft2(x, 0)
    kr, kc = np.meshgrid(np.arange(H // 2 + 1), np.arange(W), indexing="ij")
    assert np.abs(F - np.where((kr + kc) % 2 == 0, 1.0, -1.0)).max() < 2e-6


def test_polar_warp_bit_exact(cf):
    rng = np.random.default_rng(1)
    power = rng.random((H, W)).astype(np.float32)
    want = oc.polar(oc.fftshift(oc.remove_zero_component(power)), D, CP)
    got = cf.debug_polar(power)
    assert np.array_equal(got, want)


def test_rotate_bit_exact(cf, golden_stages):
    rng = np.random.default_rng(2)
    img = rng.random((H, W)).astype(np.float32)
    for d in golden_stages["rot_degrees"]:
        assert np.array_equal(cf.debug_rotate(img, d), oc.rotate(img, d)), d


def test_estimate_trans_stage(cf, cfg, imgs):
    a, b = oc.normalize_u8(imgs[0]), oc.normalize_u8(imgs[1])
    Fa, Pa = oc.compute_intermedium(cfg, a)
    Fb, Pb = oc.compute_intermedium(cfg, b)
    for which, (za, zb, R, C) in enumerate(((Fa, Fb, H, W), (Pa, Pb, D, CP))):
        io, to, po, go = oc.estimate_trans(cfg, za, zb, R, C, want_g=True)
        ig, tg, pg, gg = cf.debug_estimate_trans(za, zb, which, want_g=True)
        if which == 0:
            assert pg == po and tg == to
        else:
            assert pg[0] % (D // 2) == po[0] % (D // 2)
        assert abs(ig - io) / io < INFO_RTOL
        assert np.sqrt(np.mean((gg - go) ** 2)) / go.std() < 2e-2


# ------------------------------------------------------------------ features
def test_features_match_oracle(cf, cfg, imgs):
    for i in (0, 3):
        fr = cf.ComputeIntermedium(imgs[i])                          # u8 path (ConvertMatToNormalizedArray on the GPU)
        F, P = fr.GetFFTResult()
        Fo, Po = oc.compute_intermedium(cfg, oc.normalize_u8(imgs[i]))
        assert np.abs(F - Fo).max() / np.abs(Fo).max() < 1e-6
        assert np.abs(P - Po).max() / np.abs(Po).max() < 2e-6
        fr2 = cf.ComputeIntermedium(oc.normalize_u8(imgs[i]))        # f32 path
        F2, P2 = fr2.GetFFTResult()
        assert np.array_equal(F, F2) and np.array_equal(P, P2)


def test_feature_checksums_golden(cf, golden_pairs, imgs):
    F, P = cf.ComputeIntermedium(imgs[0]).GetFFTResult()
    assert abs(np.abs(F).astype(np.float64).sum() / golden_pairs["a_fft_result_abs_sum"] - 1) < 1e-5
    assert abs(np.abs(P).astype(np.float64).sum() / golden_pairs["a_fft_polar_abs_sum"] - 1) < 1e-5


# ------------------------------------------------------------------ ComputePose
def _check(info, pose, pk, r):
    assert pose[0] == r[2] and pose[1] == r[3], (pose, r)
    assert abs(wrap_pi(pose[2] - r[4])) < 1e-6, (pose, r)
    assert pk["trans"] == (int(r[10]), int(r[11]))
    assert pk["polar"][0] % (D // 2) == int(r[8]) % (D // 2)
    # below the reference's own "tracking lost" gate (lower_response_thr = 30, map_builder.cc:132) the peak is a noise
    # maximum a few sigma high and info is correspondingly more sensitive to f32 rounding
    # (these rows come from the scipy restatement, a cross-check; the rows pinned to the compiled reference are in golden_ref.npz)
    rtol = INFO_RTOL if min(r[5], r[7]) > 30 else 3e-3
    assert np.allclose(info[:2], r[5:7], rtol=rtol), (info, r[5:8])
    assert np.allclose(info[2], r[7], rtol=max(rtol, INFO_ROT_RTOL)), (info, r[5:8])      # polar-stage confidence: see INFO_ROT_RTOL


def test_compute_pose_golden(cf, golden_pairs, imgs):
    frames = [cf.ComputeIntermedium(u) for u in imgs]
    for r in golden_pairs["pose_rows"]:
        i, mode = int(r[0]), int(r[1])
        info, pose, pk = cf.ComputePose(frames[0], frames[i], bool(mode), return_peaks=True)
        _check(info, pose, pk, r)


def test_compute_pose_matches_oracle_on_random_pairs(cf, cfg, golden_pairs, imgs):
    # random circular rolls + the golden crops as "last" frames: exercises other keyframes than image 0
    rng = np.random.default_rng(5)
    for trial in range(4):
        ia, ib = rng.choice(len(imgs) - 1, 2, replace=False)
        a, b = oc.normalize_u8(imgs[ia]), oc.normalize_u8(imgs[ib])
        Fa, Pa = oc.compute_intermedium(cfg, a)
        Fb, Pb = oc.compute_intermedium(cfg, b)
        fa, fb = cf.ComputeIntermedium(imgs[ia]), cf.ComputeIntermedium(imgs[ib])
        for mode in (True, False):
            io, po, pko = oc.compute_pose(cfg, Fa, b, Pa, Pb, mode)
            ig, pg, pkg = cf.ComputePose(fa, fb, mode, return_peaks=True)
            if pkg["polar"][0] % (D // 2) != pko["polar"][0] % (D // 2):
                continue          # uncorrelated pair: polar peak is noise, nothing to compare
            same_twin = pkg["polar"][0] == pko["polar"][0]
            if same_twin or not mode:
                assert pkg["trans"] == pko["trans"], (ia, ib, mode)
                assert (pg[0], pg[1]) == (po[0], po[1])
                assert info_close(ig, io), (ig, io)


def test_imported_reference_layout_frames(cf, cfg, imgs):
    # frames built from reference-layout arrays (as MapBuilder holds them) give the same answer
    a, b = oc.normalize_u8(imgs[0]), oc.normalize_u8(imgs[2])
    Fa, Pa = oc.compute_intermedium(cfg, a)
    Fb, Pb = oc.compute_intermedium(cfg, b)
    fa, fb = cf.ImportFrame(a, Fa, Pa), cf.ImportFrame(b, Fb, Pb)
    F2, P2 = fa.GetFFTResult()
    assert np.array_equal(F2, Fa) and np.array_equal(P2, Pa)
    io, po, pko = oc.compute_pose(cfg, Fa, b, Pa, Pb, True)
    ig, pg, pkg = cf.ComputePose(fa, fb, True, return_peaks=True)
    assert (pg[0], pg[1]) == (po[0], po[1]) == (11.0, -6.0)
    assert info_close(ig, io), (ig, io)


def test_circular_roll_known_answers(cf, imgs):
    a = imgs[0]
    fa = cf.ComputeIntermedium(a)
    base = None
    for sy, sx in ((0, 0), (3, 0), (0, -9), (17, 25)):
        fb = cf.ComputeIntermedium(np.roll(a, (sy, sx), axis=(0, 1)))
        info, pose, pk = cf.ComputePose(fa, fb, True, return_peaks=True)
        assert (pose[0], pose[1]) == (-sx, -sy)
        assert abs(wrap_pi(pose[2])) < 1e-6
        assert pk["polar"][0] % (D // 2) == 0
        if base is None:
            base = info
            assert pk["trans"] == (H // 2, W // 2)
        assert np.allclose(info, base, rtol=1e-3)


def test_gaussian_kernel_and_invalid_kernel(golden_pairs, imgs):
    import ni_slam_b200 as nis
    c = nis.CorrelationFlow(nis.CFConfig(kernel=1), H, W)
    fa, fb = c.ComputeIntermedium(imgs[0]), c.ComputeIntermedium(imgs[2])
    info, pose = c.ComputePose(fa, fb, True)
    r = golden_pairs["gauss_row"]
    assert pose[0] == r[2] and pose[1] == r[3] and abs(wrap_pi(pose[2] - r[4])) < 1e-6
    # the gaussian kernel exponentiates (xx + zz - 2 xz)/n where xx, zz are f32 sums of ~1e5 magnitude: its response is
    # inherently sensitive to f32 summation order (the reference's Eigen order is yet another one), hence 2e-3
    assert np.allclose(info, r[5:8], rtol=2e-3)
    c.close()
    bad = nis.CorrelationFlow(nis.CFConfig(kernel=7), H, W)          # ctor succeeds, EstimateTrans throws (:168)
    fa, fb = bad.ComputeIntermedium(imgs[0]), bad.ComputeIntermedium(imgs[2])
    with pytest.raises(ValueError, match="Received invalid kernel type"):
        bad.ComputePose(fa, fb, True)
    bad.close()


def test_unsupported_sizes_are_rejected():
    import ni_slam_b200 as nis
    with pytest.raises(nis.NisError):
        nis.CorrelationFlow(nis.CFConfig(), 481, 640)      # odd height
    with pytest.raises(nis.NisError):
        nis.CorrelationFlow(nis.CFConfig(), 500, 640)      # no instantiated column plan


# ------------------------------------------------------------------ stream
def test_track_stream_matches_oracle(cf, cfg, imgs):
    frames = np.stack([imgs[0], imgs[1], imgs[2], imgs[3], imgs[0], imgs[5]])
    poses, infos = cf.TrackStream(frames)
    po, io = oc.track_stream(cfg, frames, threads=4)
    assert poses.shape == (5, 3)
    for t in range(5):
        if io[t, 0] < 30 or io[t, 2] < 30:        # tracking lost in the oracle: peaks are noise (map_builder.cc:132)
            continue
        assert (poses[t, 0], poses[t, 1]) == (po[t, 0], po[t, 1]), t
        assert abs(wrap_pi(poses[t, 2] - po[t, 2])) < 1e-6
        assert info_close(infos[t], io[t]), (t, infos[t], io[t])
    # batch-size independence: same stream, batch of 2 (ragged last batch)
    cf.set_batch(2)
    poses2, infos2 = cf.TrackStream(frames)
    cf.set_batch(0)
    assert np.array_equal(poses, poses2) and np.allclose(infos, infos2, rtol=1e-6)
    # single frame: nothing to solve
    p1, i1 = cf.TrackStream(frames[:1])
    assert p1.shape == (0, 3)


# ------------------------------------------------------------------ loop-closure scan
def test_loop_scan_matches_oracle(cf, cfg, imgs):
    import ni_slam_b200 as nis
    lc = nis.LoopClosure(nis.LoopClosureConfig(position_response_thr=60, angle_response_thr=60), cf)
    lc.clear()
    order = [3, 0, 0, 5, 0, 6, 2]                       # three identical copies of keyframe 0: the FIRST must win
    ids = [10 + k for k in range(len(order))]
    dists = [float(k) for k in range(len(order))]
    lc.AddImages(imgs[order], ids, dists)
    assert lc.size() == len(order)
    q = cf.ComputeIntermedium(imgs[1])
    res, allr = lc.FindLoopClosure(q, 99, 50.0, return_all=True)
    feats = [oc.compute_intermedium(cfg, oc.normalize_u8(imgs[i])) for i in order]
    qi = oc.normalize_u8(imgs[1])
    _, qP = oc.compute_intermedium(cfg, qi)
    kfs = [(ids[k], feats[k][0], feats[k][1], dists[k]) for k in range(len(order))]
    thr = oc.LoopConfigC(60.0, 60.0, 0, 0.0)
    ro = oc.find_loop_closure(cfg, thr, qi, qP, 99, 50.0, kfs)
    assert res.found and ro["found"]
    assert res.loop_slot == ro["index"] == 1 and res.loop_frame_id == ro["frame_id"] == 11
    assert tuple(res.relative_pose[:2]) == tuple(ro["relative_pose"][:2]) == (7.0, 0.0)
    assert info_close(res.response, ro["response"]), (res.response, ro["response"])
    assert res.evaluated == len(order)
    assert np.array_equal(allr[1], allr[2]) and np.array_equal(allr[1], allr[4])     # identical keyframes, identical bits
    # explicit candidate list in a different order: iteration order decides ties
    res2 = lc.FindLoopClosure(q, 99, 50.0, candidate_slots=[6, 4, 2, 1])
    assert res2.loop_slot == 4
    # filters (loop_closure.cc:43-53)
    lc2 = nis.LoopClosure(nis.LoopClosureConfig(60, 60, frame_gap_thr=89, distance_thr=0.0), cf)
    r3 = lc2.FindLoopClosure(q, 99, 50.0)
    assert r3.loop_slot == 0 and r3.evaluated == 1
    lc3 = nis.LoopClosure(nis.LoopClosureConfig(60, 60, frame_gap_thr=0, distance_thr=49.5), cf)
    r4 = lc3.FindLoopClosure(q, 99, 50.0)
    assert r4.loop_slot == 0 and r4.evaluated == 1
    # everything filtered / empty list: initial best (-1,-1,-1), not found
    lc4 = nis.LoopClosure(nis.LoopClosureConfig(60, 60, frame_gap_thr=1000), cf)
    r5 = lc4.FindLoopClosure(q, 99, 50.0)
    assert (not r5.found) and r5.loop_slot == -1 and np.array_equal(r5.response, [-1, -1, -1]) and r5.evaluated == 0
    # batch-size independence
    cf.set_batch(3)
    res6 = lc.FindLoopClosure(q, 99, 50.0)
    cf.set_batch(0)
    assert res6.loop_slot == res.loop_slot and np.array_equal(res6.response, res.response)
    # rank reduction helper: shard [0..3] and [4..6], global order = slot
    ra = lc.FindLoopClosure(q, 99, 50.0, candidate_slots=[0, 1, 2, 3])
    rb = lc.FindLoopClosure(q, 99, 50.0, candidate_slots=[4, 5, 6])
    red, win = lc.Reduce([rb, ra], order=[rb.loop_slot, ra.loop_slot])
    assert win == 1 and red.loop_slot == 1 and red.found and red.evaluated == 7
    lc.clear()


def test_scan_rotated_query_cache_is_result_preserving(cf, imgs, monkeypatch):
    """Long scans take FFT(RotateArray(query, angle)) from a per-query cache of all 2 D angles instead of re-rotating per
    candidate; the responses must not change."""
    import ni_slam_b200 as nis
    order = [3, 0, 5, 6, 2, 4, 0, 7]
    lc = nis.LoopClosure(nis.LoopClosureConfig(60, 60), cf)
    lc.clear()
    lc.AddImages(imgs[order])
    q = cf.ComputeIntermedium(imgs[1])
    res, allr = lc.FindLoopClosure(q, 99, 50.0, return_all=True)
    lc.clear()
    monkeypatch.setenv("NIS_ROT_CACHE_MIN", "1")
    cf2 = nis.CorrelationFlow(nis.CFConfig(), H, W)
    lc2 = nis.LoopClosure(nis.LoopClosureConfig(60, 60), cf2)
    lc2.AddImages(imgs[order])
    q2 = cf2.ComputeIntermedium(imgs[1])
    res2, allr2 = lc2.FindLoopClosure(q2, 99, 50.0, return_all=True)
    assert res2.loop_slot == res.loop_slot == 1 and res2.peak == res.peak and res2.hyp == res.hyp
    assert np.array_equal(res2.relative_pose, res.relative_pose)
    assert np.allclose(allr2, allr, rtol=2e-6)
    cf2.close()


def test_sharded_entry_point_single_rank(cf, imgs):
    """nis_loop_scan_sharded on a one-rank communicator (no NCCL call): query image in, features on the device, local scan, the
    reduction -- same winner, pose and response bits as ComputeIntermedium + FindLoopClosure; the slot comes back as a GLOBAL slot."""
    import ni_slam_b200 as nis
    order = [3, 0, 5, 6, 2, 4, 7]
    lc = nis.LoopClosure(nis.LoopClosureConfig(60, 60), cf)
    lc.clear()
    lc.AddImages(imgs[order])
    ref = lc.FindLoopClosure(cf.ComputeIntermedium(imgs[1]), 99, 50.0)
    res, win, mine = lc.FindLoopClosureSharded(imgs[1], 0, 1000, current_frame_id=99, current_distance=50.0)
    assert win == 0 and res.found == ref.found and res.loop_slot == 1000 + ref.loop_slot and mine.loop_slot == ref.loop_slot
    assert res.evaluated == ref.evaluated == len(order)
    assert np.array_equal(res.relative_pose, ref.relative_pose) and np.array_equal(res.response, ref.response)
    lc.clear()
    empty, win0, _ = lc.FindLoopClosureSharded(imgs[1], 0, 0)            # empty shard: no winner, nothing found
    assert win0 == -1 and not empty.found and empty.loop_slot == -1 and empty.evaluated == 0


def test_prior_pose_candidate_selection(cf, imgs):
    """FindLoopClosure(image, frame, prior_pose): Map::ComputeGridLocation + 3x3 cells + GetFramesInGrids (loop_closure.cc:17-34)."""
    import ni_slam_b200 as nis
    lc = nis.LoopClosure(nis.LoopClosureConfig(60, 60), cf)
    lc.clear()
    order = [3, 0, 5, 0, 2, 6]
    lc.AddImages(imgs[order])
    scale = 0.1
    positions = [(0.05, 0.05), (0.12, 0.18), (0.35, 0.05), (0.26, 0.11), (-0.04, -0.08), (0.19, 0.29)]
    for slot, (x, y) in enumerate(positions):
        lc.SetPosition(slot, x, y, scale)
    q = cf.ComputeIntermedium(imgs[1])
    prior = (0.15, 0.15, 0.0)                                   # cell (1, 1): neighbourhood = cells 0..2 x 0..2

    def cell(x, y):
        return (int(x / scale), int(y / scale))                 # truncation toward zero like static_cast<int>
    want = []
    for i in (-1, 0, 1):
        for j in (-1, 0, 1):
            want += [s for s, (x, y) in enumerate(positions) if cell(x, y) == (1 + i, 1 + j)]
    res, cand = lc.FindLoopClosurePrior(q, prior, scale, 99, 50.0)
    assert list(cand) == want == [0, 4, 1, 5, 3]                # slot 2 is in cell (3, 0): outside; (-0.04,-0.08) truncates to (0, 0)
    ref = lc.FindLoopClosure(q, 99, 50.0, candidate_slots=want)
    assert res.loop_slot == ref.loop_slot == 1 and np.array_equal(res.response, ref.response)
    # a prior far away: no candidates, initial result
    res2, cand2 = lc.FindLoopClosurePrior(q, (5.0, 5.0, 0.0), scale, 99, 50.0)
    assert len(cand2) == 0 and not res2.found and res2.loop_slot == -1
    lc.clear()


def test_large_rotation_loop_mode(cf, golden_pairs, imgs):
    # image 6 is rotated by 152 degrees: only loop mode (two hypotheses) recovers it
    fa, fb = cf.ComputeIntermedium(imgs[0]), cf.ComputeIntermedium(imgs[6])
    info, pose, pk = cf.ComputePose(fa, fb, False, return_peaks=True)
    assert (pose[0], pose[1]) == (-41.0, -17.0)
    assert abs(wrap_pi(pose[2] - np.deg2rad(152.0))) < 1e-6


# ------------------------------------------------------------------ reference-facing C++ shim
def test_cpp_shim_known_answers():
    """ni_slam_b200/host/correlation_flow.hpp driven like MapBuilder drives the reference classes (tests/cpp/shim_test.cc)."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "cpp", "_build", "shim_test")
    src = os.path.join(root, "tests", "cpp", "shim_test.cc")
    if not os.path.exists(exe) or os.path.getmtime(src) > os.path.getmtime(exe):
        os.makedirs(os.path.dirname(exe), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++17", "-O1", src, "-o", exe, "-L", os.path.join(root, "ni_slam_b200", "lib"), "-lnislam",
                               "-Wl,-rpath," + os.path.join(root, "ni_slam_b200", "lib")])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "shim_test ok" in out.stdout, out.stdout + out.stderr


# ------------------------------------------------------------------ other image sizes (BASELINE configs[4] = 1280x960; a tiny config)
@pytest.mark.parametrize("h,w,d,cp", [(96, 128, 80, 64), (960, 1280, 720, 480)])
def test_other_sizes_match_oracle(h, w, d, cp):
    import ni_slam_b200 as nis
    rng = np.random.default_rng(h)
    base = rng.random((h + 64, w + 64)).astype(np.float32)
    k = np.ones(5, np.float32) / 5                       # cheap separable blur so that shifts correlate
    for ax in (0, 1):
        base = np.apply_along_axis(lambda v: np.convolve(v, k, mode="same"), ax, base)
    base = (base - base.min()) / (base.max() - base.min())
    a_u8 = np.clip(np.rint(base[32:32 + h, 32:32 + w] * 255), 0, 255).astype(np.uint8)
    b_u8 = np.clip(np.rint(base[32 + 3:32 + 3 + h, 32 - 5:32 - 5 + w] * 255), 0, 255).astype(np.uint8)      # window moved by (-5, +3)
    cfg = oc.make_cfg(height=h, width=w, rotation_divisor=d, rotation_channel=cp)
    c = nis.CorrelationFlow(nis.CFConfig(rotation_divisor=d, rotation_channel=cp), h, w)
    a, b = oc.normalize_u8(a_u8), oc.normalize_u8(b_u8)
    Fa, Pa = oc.compute_intermedium(cfg, a)
    Fb, Pb = oc.compute_intermedium(cfg, b)
    fa, fb = c.ComputeIntermedium(a_u8), c.ComputeIntermedium(b_u8)
    F, P = fa.GetFFTResult()
    assert np.abs(F - Fa).max() / np.abs(Fa).max() < 2e-6 and np.abs(P - Pa).max() / np.abs(Pa).max() < 3e-6
    for mode in (True, False):
        io, po, pko = oc.compute_pose(cfg, Fa, b, Pa, Pb, mode)
        ig, pg, pkg = c.ComputePose(fa, fb, mode, return_peaks=True)
        assert pkg["polar"][0] % (d // 2) == pko["polar"][0] % (d // 2)
        assert pkg["trans"] == pko["trans"] and (pg[0], pg[1]) == (po[0], po[1]) == (-5.0, 3.0)
        assert abs(wrap_pi(pg[2] - po[2])) < 1e-6 or not mode
        assert np.allclose(ig, io, rtol=1e-3)
    c.close()


# ------------------------------------------------------------------ undistort front end (Camera::UndistortImage)
def _synthetic_maps(h, w):
    """fixed-point remap maps of a mild radial distortion, built without cv2: (x, y) int16 + (fy*32 + fx) uint16"""
    ys, xs = np.meshgrid(np.arange(h, dtype=np.float64), np.arange(w, dtype=np.float64), indexing="ij")
    xn, yn = (xs - w / 2) / (0.8 * w), (ys - h / 2) / (0.8 * w)
    r2 = xn * xn + yn * yn
    f = 1 - 0.25 * r2 + 0.08 * r2 * r2
    sx = np.rint(((xn * f) * 0.8 * w + w / 2) * 32).astype(np.int64)
    sy = np.rint(((yn * f) * 0.8 * w + h / 2) * 32).astype(np.int64)
    m1 = np.stack([sx >> 5, sy >> 5], axis=-1).astype(np.int16)
    m2 = ((sy & 31) * 32 + (sx & 31)).astype(np.uint16)
    return m1, m2


def test_undistort_front_end(imgs):
    import os
    import ni_slam_b200 as nis
    # stage: bit-exact against the cv2-generated golden (small camera) and the C oracle (full size)
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_undistort.npz"))
    small = nis.CorrelationFlow(nis.CFConfig(rotation_divisor=80, rotation_channel=64), 96, 128)
    small.SetUndistortMaps(g["map1"], g["map2"])
    assert np.array_equal(small.UndistortImage(g["raw"]), g["out"])
    small.close()
    c = nis.CorrelationFlow(nis.CFConfig(), H, W)
    m1, m2 = _synthetic_maps(H, W)
    c.SetUndistortMaps(m1, m2)
    raw = imgs[:4]
    und = np.stack([oc.undistort_u8(r, m1, m2) for r in raw])
    for r, u in zip(raw, und):
        assert np.array_equal(c.UndistortImage(r), u)
    # end to end: raw frames through the front end == undistorted frames without it (features, stream poses, scan)
    F_raw, P_raw = c.ComputeIntermedium(raw[0]).GetFFTResult()
    poses_raw, infos_raw = c.TrackStream(raw)
    lc = nis.LoopClosure(nis.LoopClosureConfig(60, 60), c)
    lc.clear()
    lc.AddImages(raw[[2, 0, 3]])
    res_raw = lc.FindLoopClosure(c.ComputeIntermedium(raw[1]), 99, 50.0)
    lc.clear()
    c.SetUndistortMaps(None, None)
    F_u, P_u = c.ComputeIntermedium(und[0]).GetFFTResult()
    poses_u, infos_u = c.TrackStream(und)
    lc.AddImages(und[[2, 0, 3]])
    res_u = lc.FindLoopClosure(c.ComputeIntermedium(und[1]), 99, 50.0)
    assert np.array_equal(F_raw, F_u) and np.array_equal(P_raw, P_u)
    assert np.array_equal(poses_raw, poses_u) and np.allclose(infos_raw, infos_u, rtol=1e-6)
    assert res_raw.loop_slot == res_u.loop_slot and np.allclose(res_raw.response, res_u.response, rtol=1e-6)
    lc.clear()
    c.close()


# ------------------------------------------------------------------ ABI edge cases and error behaviour
def test_abi_edge_cases(cf, imgs):
    import ctypes as C
    import ni_slam_b200 as nis
    from ni_slam_b200 import api
    lib = nis.load_library()
    # NULL arguments -> NIS_ERR_INVALID_ARGUMENT, never a crash
    assert lib.nis_features_u8(cf._ctx, None, None) == api.NIS_ERR_INVALID_ARGUMENT
    assert lib.nis_compute_pose(cf._ctx, None, None, 1, None, None, None) == api.NIS_ERR_INVALID_ARGUMENT
    assert lib.nis_track_stream(cf._ctx, None, 3, None, None) == api.NIS_ERR_INVALID_ARGUMENT
    assert lib.nis_destroy(None) == api.NIS_OK and lib.nis_frame_free(cf._ctx, None) == api.NIS_OK
    assert lib.nis_strerror(api.NIS_ERR_INVALID_KERNEL) == b"Received invalid kernel type"
    # wrong image shape is rejected on the Python side before reaching the ABI
    with pytest.raises(ValueError):
        cf.ComputeIntermedium(np.zeros((H, W + 1), np.uint8))
    # empty bulk insert, scan of an empty store, candidate slot out of range
    lc = nis.LoopClosure(nis.LoopClosureConfig(60, 60), cf)
    lc.clear()
    lc.AddImages(np.zeros((0, H, W), np.uint8))
    assert lc.size() == 0
    q = cf.ComputeIntermedium(imgs[1])
    r = lc.FindLoopClosure(q, 1, 0.0)
    assert not r.found and r.loop_slot == -1 and r.evaluated == 0
    with pytest.raises(nis.NisError):
        lc.FindLoopClosure(q, 1, 0.0, candidate_slots=[0])
    # two-frame stream (one solve), and a constant image: finite outputs, no crash (info is 0/eps by construction)
    poses, infos = cf.TrackStream(np.stack([imgs[0], imgs[1]]))
    assert poses.shape == (1, 3) and (poses[0, 0], poses[0, 1]) == (7.0, 0.0)
    flat = np.full((2, H, W), 128, np.uint8)
    p2, i2 = cf.TrackStream(flat)
    assert np.all(np.isfinite(p2)) and np.all(np.isfinite(i2))
    # launches are counted (the driver reads this as evidence that the CUDA path ran)
    assert cf.kernel_launches() > 0


# ------------------------------------------------------------------ seeded random pairs (SURVEY 8d synthetic inputs)
def test_random_pairs_match_oracle(cf, cfg):
    """12 seeded pairs (|d| <= 60 px, |ang| <= 15 deg) cropped from the SURVEY App. C canvas: translation peak bit-exact, polar
    row equal mod D/2, dx/dy exact, info within tolerance, in both modes."""
    ref = pytest.importorskip("nislam_ref")
    if ref.cv2 is None:
        pytest.skip("cv2 missing")
    canvas = ref.make_canvas(0)
    rng = np.random.default_rng(2024)
    a_u8 = ref.crop(canvas, 640, 480, 0)
    a = oc.normalize_u8(a_u8)
    Fa, Pa = oc.compute_intermedium(cfg, a)
    fa = cf.ComputeIntermedium(a_u8)
    n_checked = 0
    for _ in range(12):
        dx, dy = rng.integers(-60, 61, 2)
        ang = rng.integers(-30, 31) * 0.5
        b_u8 = ref.crop(canvas, 640 + dx, 480 + dy, ang)
        b = oc.normalize_u8(b_u8)
        Fb, Pb = oc.compute_intermedium(cfg, b)
        fb = cf.ComputeIntermedium(b_u8)
        for mode in (True, False):
            io, po, pko = oc.compute_pose(cfg, Fa, b, Pa, Pb, mode)
            ig, pg, pkg = cf.ComputePose(fa, fb, mode, return_peaks=True)
            assert pkg["polar"][0] % (D // 2) == pko["polar"][0] % (D // 2), (dx, dy, ang, mode)
            assert pkg["trans"] == pko["trans"], (dx, dy, ang, mode)
            assert (pg[0], pg[1]) == (po[0], po[1]) == (float(dx), float(dy)), (dx, dy, ang, mode, pg, po)
            assert abs(wrap_pi(pg[2] - po[2])) < 1e-6, (dx, dy, ang, mode, pg, po)
            assert abs(wrap_pi(pg[2] - np.deg2rad(ang))) < np.deg2rad(0.5) + 1e-6      # ground truth to one polar bin (360/D)
            assert np.allclose(ig[:2], io[:2], rtol=INFO_RTOL), (ig, io)
            assert np.allclose(ig[2], io[2], rtol=INFO_ROT_RTOL), (ig, io)
            n_checked += 1
    assert n_checked == 24


# ------------------------------------------------------------------ keyframe-policy tracking (map_builder.cc:30-70)
def test_track_stream_keyframes_matches_map_builder_restatement(cf, cfg):
    """nis_track_stream_keyframes (speculative batches against the last keyframe) against the frame-by-frame restatement of
    MapBuilder::AddNewInput driven by the C oracle: same tracked / inserted / keyframe sequence, integer pixel shifts exact, angles
    equal mod 2 pi, composed poses to 1e-9, confidences within the info tolerances."""
    ref = pytest.importorskip("nislam_ref")
    if ref.cv2 is None:
        pytest.skip("cv2 missing")
    import ni_slam_b200 as nis
    import tracker_ref as tr
    canvas = ref.make_canvas(0)
    rng = np.random.default_rng(11)
    n = 36
    steps = np.stack([rng.integers(-7, 8, n), rng.integers(-7, 8, n)], 1)
    steps[0] = 0
    xy = np.cumsum(steps, 0)
    ang = np.cumsum(np.r_[0, rng.integers(-2, 3, n - 1) * 0.5])
    frames = np.stack([ref.crop(canvas, 640 + int(xy[t, 0]), 480 + int(xy[t, 1]), float(ang[t])) for t in range(n)])
    E = [0.0, -1.0, 0.1, 1.0, 0.0, 0.2, 0.0, 0.0, 1.0]
    cam = nis.CameraModel(fx=800.0, fy=820.0, cx=330.0, cy=235.0, height=0.5, extrinsics=tuple(E))
    kfs = nis.KeyframeSelectionConfig(max_distance=0.02, max_angle=0.03, lower_response_thr=30.0, upper_response_thr=90.0)
    got = cf.TrackStreamKeyframes(frames, kfs, cam)
    assert got.shape == (n,)

    def cpose(lF, img, lP, P):
        info, pose, _ = oc.compute_pose(cfg, lF, img, lP, P, True)
        return info, pose
    trk = tr.MapBuilderTracker(tr.Camera(cam.fx, cam.fy, cam.cx, cam.cy, cam.height, E, W, H), kfs.max_distance, kfs.max_angle,
                               kfs.lower_response_thr, kfs.upper_response_thr, lambda img: oc.compute_intermedium(cfg, img), cpose)
    n_kf = 0
    for t in range(n):
        o = trk.add_new_input(oc.normalize_u8(frames[t]))
        g = got[t]
        assert (bool(g["tracked"]), bool(g["inserted"]), int(g["keyframe"])) == (o["tracked"], o["inserted"], o["keyframe"]), t
        if t:
            assert np.allclose(g["response"][:2], o["response"][:2], rtol=INFO_RTOL) and np.allclose(g["response"][2], o["response"][2], rtol=INFO_ROT_RTOL), t
            assert abs(wrap_pi(g["relative_pose"][2] - o["relative_pose"][2])) < 1e-6, t
            assert np.allclose(g["relative_pose"][:2], o["relative_pose"][:2], rtol=0, atol=1e-9), t
        assert np.allclose(g["cf_pose"][:2], o["cf_pose"][:2], rtol=0, atol=1e-9) and abs(wrap_pi(g["cf_pose"][2] - o["cf_pose"][2])) < 1e-9, t
        assert np.allclose(g["pose"][:2], o["pose"][:2], rtol=0, atol=1e-9) and abs(wrap_pi(g["pose"][2] - o["pose"][2])) < 1e-9, t
        assert abs(g["distance"] - o["distance"]) < 1e-12
        n_kf += int(o["inserted"])
    assert 3 <= n_kf < n            # the policy really alternates between keyframes and tracked-only frames
    # a one-frame stream is just Initialize
    one = cf.TrackStreamKeyframes(frames[:1], kfs, cam)
    assert one.shape == (1,) and one[0]["inserted"] == 1 and one[0]["keyframe"] == -1


# ------------------------------------------------------------------ MapStitcher (map_stitcher.cc): integer work, bit-exact
def _stitch_compare(st_gpu, st_ref, x0, y0, nx, ny):
    seen = 0
    for cy in range(y0, y0 + ny):
        for cx in range(x0, x0 + nx):
            got = st_gpu.cell(cx, cy)
            exp = st_ref.cells.get((cx, cy))
            assert (got is None) == (exp is None), (cx, cy)
            if got is not None:
                assert np.array_equal(got[1], exp[1]), ("weight", cx, cy)
                assert np.array_equal(got[0], exp[0]), ("data", cx, cy)
                seen += 1
    assert seen == len(st_ref.cells)


def test_map_stitcher_bit_exact():
    import ni_slam_b200 as nis
    import stitcher_ref as sr
    import tracker_ref as tr
    rng = np.random.default_rng(5)
    E = [0.6, -0.8, 0.05, 0.8, 0.6, -0.02, 0.0, 0.0, 1.0]
    for (Hs, Ws, cs, nfr, spread) in ((96, 128, 50, 14, 0.12), (480, 640, 1000, 4, 0.9)):
        camd = dict(fx=410.0, fy=395.0, cx=Ws / 2 - 3.5, cy=Hs / 2 + 2.25, height=0.8)
        cam = nis.CameraModel(extrinsics=tuple(E), **camd)
        ref = sr.MapStitcher(cs, tr.Camera(extrinsics=E, image_width=Ws, image_height=Hs, **camd))
        gpu = nis.MapStitcher(cs, cam, Hs, Ws, cell_x0=-4, cell_y0=-4, cells_x=8, cells_y=8)
        poses = []
        for f in range(nfr):
            img = rng.integers(0, 256, (Hs, Ws), dtype=np.uint8)
            pose = [rng.uniform(-spread, spread), rng.uniform(-spread, spread), rng.uniform(-np.pi, np.pi)]
            if f == 0:
                pose = [0.0, 0.0, 0.0]                  # Initialize: exact integer ground positions
            if f == 3:
                pose = list(poses[2])                   # a revisit: every element of the footprint merges
            poses.append(pose)
            assert gpu.InsertFrame(img, pose) == f
            ref.insert_frame(img, pose)
        assert gpu.frames() == nfr and gpu.dropped() == 0
        _stitch_compare(gpu, ref, -4, -4, 8, 8)
        # RecomputeOccupancy after "optimisation": perturbed poses, frames replayed in insertion order
        poses2 = [[p[0] + rng.normal(0, 0.01), p[1] + rng.normal(0, 0.01), p[2] + rng.normal(0, 0.02)] for p in poses]
        gpu.RecomputeOccupancy(poses2)
        ref.recompute_occupancy(poses2)
        assert gpu.dropped() == 0
        _stitch_compare(gpu, ref, -4, -4, 8, 8)
        gpu.close()
    # a window that is too small: pixels outside are counted, cells outside do not exist
    cam = nis.CameraModel(fx=400.0, fy=400.0, cx=64.0, cy=48.0, height=1.0)
    small = nis.MapStitcher(50, cam, 96, 128, cell_x0=0, cell_y0=0, cells_x=1, cells_y=1)
    small.InsertFrame(np.full((96, 128), 255, np.uint8), [0.0, 0.0, 0.0])
    d, w = small.cell(0, 0)
    assert w.sum() == 50 * 48 and small.dropped() == 96 * 128 - 50 * 48 and small.cell(-1, 0) is None and d.max() == 100
    small.close()


def test_map_stitcher_conservation_full_size():
    """Size-independent property at the full 640x480 / 1000-cell geometry: every source pixel lands in exactly one cell element, so
    after one InsertFrame the weights sum to H*W (none dropped) and the data to the sum of the scaled image, for any pose."""
    import ni_slam_b200 as nis
    import stitcher_ref as sr
    rng = np.random.default_rng(17)
    cam = nis.CameraModel(fx=900.0, fy=880.0, cx=W / 2 + 6.0, cy=H / 2 - 3.0, height=1.3, extrinsics=(0.0, -1.0, 0.3, 1.0, 0.0, -0.1, 0.0, 0.0, 1.0))
    for trial in range(3):
        st = nis.MapStitcher(1000, cam, H, W, cell_x0=-2, cell_y0=-2, cells_x=4, cells_y=4)
        img = rng.integers(0, 256, (H, W), dtype=np.uint8)
        st.InsertFrame(img, [rng.uniform(-0.8, 0.8), rng.uniform(-0.8, 0.8), rng.uniform(-np.pi, np.pi)])
        assert st.dropped() == 0
        wsum = dsum = 0
        for cy in range(-2, 2):
            for cx in range(-2, 2):
                c = st.cell(cx, cy)
                if c is not None:
                    dsum += int(c[0].astype(np.int64).sum())
                    wsum += int(c[1].astype(np.int64).sum())
        assert wsum == H * W and dsum == int(sr.normalize_image(img).astype(np.int64).sum())
        st.close()


# ------------------------------------------------------------------ pinned to the reference's own source (oracle/_ref goldens)
@pytest.fixture(scope="module")
def gref():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_ref.npz"))


def test_compute_pose_matches_compiled_reference_goldens(gref, imgs):
    """tests/golden/golden_ref.npz = outputs of /root/reference/src/correlation_flow.cc compiled unmodified (oracle/_ref): 9 pairs x
    both modes x both kernels.  dx, dy exact; theta mod 2 pi (the polar twin); info within INFO_RTOL (2e-3 below the
    tracking-lost gate and for the gaussian kernel, whose response hangs on f32 sums of ~1e5 magnitude)."""
    import ni_slam_b200 as nis
    for kernel in (0, 1):
        c = nis.CorrelationFlow(nis.CFConfig(kernel=kernel), H, W)
        frames = [c.ComputeIntermedium(u) for u in imgs]
        rows = [r for r in gref["pose_rows"] if int(r[0]) == kernel]
        assert len(rows) == 18
        worst = 0.0
        for r in rows:
            a, b, mode = int(r[1]), int(r[2]), int(r[3])
            info, pose = c.ComputePose(frames[a], frames[b], bool(mode))
            confident = min(r[7], r[9]) > 30
            if not confident and (pose[0], pose[1]) != (r[4], r[5]):
                # below the reference's own "tracking lost" gate (30, map_builder.cc:132) the arg-max sits on a noise maximum a few sigma
                # high; two of them a rounding error apart may swap (seen: gaussian kernel on the pair without overlap)
                assert np.allclose(info[:2], r[7:9], rtol=2e-2), (r[:4], info, r[7:10])
                continue
            assert pose[0] == r[4] and pose[1] == r[5], (r[:4], pose, r[4:7])
            assert abs(wrap_pi(pose[2] - r[6])) < 1e-6, (r[:4], pose, r[4:7])
            # gaussian kernel on a pair without overlap ("tracking lost"): exp() of f32 sums of ~1e5 magnitude on a response that is noise
            rtol = (INFO_RTOL if kernel == 0 else 2e-3) if confident else (5e-3 if kernel == 0 else 1e-2)
            assert np.allclose(info[:2], r[7:9], rtol=rtol), (r[:4], info, r[7:10])
            assert np.allclose(info[2], r[9], rtol=max(rtol, INFO_ROT_RTOL)), (r[:4], info, r[7:10])
            if confident:
                worst = max(worst, float(np.max(np.abs(info - r[7:10]) / r[7:10])))
        print("kernel %d: worst info deviation from the compiled reference on confident pairs: %.2e" % (kernel, worst))
        c.close()


def test_scan_matches_compiled_reference_goldens(cf, gref, imgs):
    """LoopClosure::FindLoopClosure through the reference's own Map / Frame (golden_ref.npz): list, filters, duplicate, all, prior."""
    import ni_slam_b200 as nis
    order = [int(i) for i in gref["scan_order"]]
    q = cf.ComputeIntermedium(imgs[int(gref["scan_query"])])

    def same(res, row, slot_to_index=lambda s: s):
        assert (int(res.found), slot_to_index(res.loop_slot), res.loop_frame_id) == tuple(int(v) for v in row[:3]), (res, row)
        assert tuple(res.relative_pose[:2]) == tuple(row[3:5]) and abs(wrap_pi(res.relative_pose[2] - row[5])) < 1e-6
        assert np.allclose(res.response, row[6:9], rtol=INFO_ROT_RTOL)
    lc = nis.LoopClosure(nis.LoopClosureConfig(position_response_thr=30, angle_response_thr=60), cf)
    lc.clear()
    ids = [10 + k for k in range(len(order))]
    lc.AddImages(imgs[order], ids, [float(k) for k in range(len(order))])
    same(lc.FindLoopClosure(q, 100, 50.0), gref["scan_list"])
    lcf = nis.LoopClosure(nis.LoopClosureConfig(30, 60, frame_gap_thr=89, distance_thr=47.5), cf)
    same(lcf.FindLoopClosure(q, 100, 50.0), gref["scan_filtered"])
    # prior pose (1.0, 0.2), grid_scale 2: the reference's Map files frames by pose at insertion (map.cc:27-30)
    for slot, p in enumerate(gref["scan_poses"]):
        lc.SetPosition(slot, p[0], p[1], 2.0)
    res, cand = lc.FindLoopClosurePrior(q, (1.0, 0.2, 0.0), 2.0, 100, 50.0)
    row = gref["scan_prior"].copy()
    if int(row[1]) == 0:
        row[2] = 10          # Map::AddFrame renames the first frame to id 0 (map.cc:19-22); the store keeps the caller's ids
    same(res, row)
    lc.clear()
    lc.AddImages(imgs[[1, 1, 5]], [7, 8, 9], [0.0, 1.0, 2.0])
    same(lc.FindLoopClosure(q, 100, 50.0), gref["scan_dup"])                    # strict '>': the first of two identical keyframes
    lc.clear()


# ------------------------------------------------------------------ the scan at DB scale: every candidate against the oracle
def _db_fixture(n, h=H, w=W, size=2048):
    import torch
    import bench_synth as bs
    canvas = bs.make_canvas(size, seed=11, device="cuda")
    cx, cy, ang = bs.db_poses(n, seed=12, size=size, H=h, W=w)
    db = bs.crops(canvas, cx, cy, ang, h, w).cpu().numpy()
    j = n // 3
    q = bs.crops(canvas, [cx[j] + 13], [cy[j] - 7], [ang[j] + 4.5], h, w).cpu().numpy()[0]
    torch.cuda.synchronize()
    return db, q, j


def _oracle_records(cfg, db, q_u8, workers=16):
    from concurrent.futures import ThreadPoolExecutor
    qi = oc.normalize_u8(q_u8)
    _, qP = oc.compute_intermedium(cfg, qi)

    def one(k):
        Fk, Pk = oc.compute_intermedium(cfg, oc.normalize_u8(db[k]))
        info, pose, pk = oc.compute_pose(cfg, Fk, qi, Pk, qP, False)
        return info, pose, pk
    with ThreadPoolExecutor(workers) as ex:
        return list(ex.map(one, range(len(db))))


def _compare_records(recs, orc, kernel=0):
    """Every candidate.  Correlated candidates (both confidences above the reference's own gate of 30, map_builder.cc:132): translation
    peak bit-exact, polar row mod D/2, hypothesis, pose, response.  Uncorrelated ones: the arg-max sits on a noise maximum a few sigma
    high and two of them a rounding error apart may swap, so positions may differ; the responses still have to agree loosely.
    Returns (confident candidates compared exactly, uncorrelated candidates whose noise peak moved)."""
    exact = moved = 0
    for k, (info, pose, pk) in enumerate(orc):
        r = recs[k]
        assert r["evaluated"] == 1
        confident = min(info[0], info[2]) > 30
        same_polar = int(r["peak"][0]) % (D // 2) == pk["polar"][0] % (D // 2)
        same_trans = tuple(int(v) for v in r["peak"][2:]) == pk["trans"]
        if not confident and not (same_polar and same_trans):
            assert r["response"][0] < 30, (k, r, info)                 # the GPU agrees that this candidate is not a match
            assert np.allclose(r["response"][2], info[2], rtol=5e-2), (k, r, info)
            moved += 1
            continue
        assert same_polar and same_trans, (k, r, info, pk)
        # the polar twin (row +- D/2) names the same two rotations the other way round: "-deg" of one is "-deg+180" of the other
        twin = int(r["peak"][0]) != pk["polar"][0]
        assert int(r["hyp"]) == (1 - pk["hyp"] if twin else pk["hyp"]), (k, r, pk)
        assert (r["relative_pose"][0], r["relative_pose"][1]) == (pose[0], pose[1])
        assert abs(wrap_pi(r["relative_pose"][2] - pose[2])) < 1e-6
        rtol = (INFO_RTOL if kernel == 0 else 2e-3) if confident else (5e-3 if kernel == 0 else 1e-2)
        assert np.allclose(r["response"][:2], info[:2], rtol=rtol), (k, r["response"], info)
        assert np.allclose(r["response"][2], info[2], rtol=max(rtol, INFO_ROT_RTOL)), (k, r["response"], info)
        exact += confident
    return exact, moved


def test_scan_256_keyframes_every_candidate_vs_oracle(cfg, monkeypatch):
    """256-keyframe store (SURVEY 8d subsample): multi-batch, multi-lane scan with per-candidate records against the C oracle's
    ComputePose(..., false) for EVERY candidate -- once with the per-candidate rotation, once with the rotated-query cache (the
    configuration the bench uses), and in all three store modes (same bits)."""
    import ni_slam_b200 as nis
    db, q_u8, j = _db_fixture(256)
    orc = _oracle_records(cfg, db, q_u8)
    best = int(np.argmax([i.sum() for i, _, _ in orc]))            # first maximum = strict '>' in iteration order
    assert best == j
    base = None
    for rot_min, mode in (("0", nis.DB_FULL), ("1", nis.DB_FULL), ("1", nis.DB_SPECTRA), ("1", nis.DB_IMAGE)):
        monkeypatch.setenv("NIS_ROT_CACHE_MIN", rot_min)
        c = nis.CorrelationFlow(nis.CFConfig(), H, W)
        c.set_batch(24)                                              # 256 / 24: eleven batches over three lanes, ragged tail
        lc = nis.LoopClosure(nis.LoopClosureConfig(30, 60), c)
        lc.SetMode(mode)
        lc.AddImages(db, list(range(1000, 1256)), [float(k) for k in range(256)])
        q = c.ComputeIntermedium(q_u8)
        res, recs = lc.FindLoopClosureRecords(q, 5000, 1e6)
        assert res.evaluated == 256 and res.loop_slot == j and res.loop_frame_id == 1000 + j and res.found
        assert tuple(res.relative_pose[:2]) == tuple(orc[j][1][:2])      # (13, -7) px of canvas motion seen in the keyframe's rotated frame
        exact, moved = _compare_records(recs, orc)
        n_conf = sum(1 for i, _, _ in orc if min(i[0], i[2]) > 30)
        assert exact == n_conf and n_conf >= 3 and moved <= 256 // 4, (exact, n_conf, moved)
        print("store mode %d, rot cache min %s: %d correlated candidates exact, %d of %d uncorrelated noise peaks moved" % (mode, rot_min, exact, moved, 256 - n_conf))
        if base is None:
            base = recs.copy()
        else:       # cached vs per-candidate rotation, and the compact store modes: same peaks and poses, responses to f32 round-off
            assert np.array_equal(recs["peak"], base["peak"]) and np.array_equal(recs["hyp"], base["hyp"])
            assert np.array_equal(recs["relative_pose"], base["relative_pose"])
            assert np.allclose(recs["response"], base["response"], rtol=2e-6)
        if mode != nis.DB_FULL:
            full = prev_full
            assert np.array_equal(recs["response"], full["response"])        # compact modes recompute with the same kernels: same bits
        if mode == nis.DB_FULL and rot_min == "1":
            prev_full = recs.copy()
        # device-side candidate selection: explicit list order and filters (loop_closure.cc:43-53)
        if mode == nis.DB_FULL and rot_min == "0":
            lst = [j + 5, j, 3, j - 7, 200]
            r2, rec2 = lc.FindLoopClosureRecords(q, 5000, 1e6, candidate_slots=lst)
            assert r2.loop_slot == j and np.array_equal(rec2["response"], recs["response"][lst])
            lcf = nis.LoopClosure(nis.LoopClosureConfig(30, 60, frame_gap_thr=0, distance_thr=10.5), c)
            r3, rec3 = lcf.FindLoopClosureRecords(q, 5000, float(j))             # |d - j| < 10.5 dropped: 21 keyframes around j
            assert r3.evaluated == 256 - 21 and rec3["evaluated"].sum() == 256 - 21 and rec3["evaluated"][j] == 0
            assert r3.loop_slot != j
            lcg = nis.LoopClosure(nis.LoopClosureConfig(30, 60, frame_gap_thr=250, distance_thr=0.0), c)
            r4 = lcg.FindLoopClosure(q, 1000, 1e6)                               # |1000 - id| < 250 dropped: ids 1250..1255 remain
            assert r4.evaluated == 6 and 250 <= r4.loop_slot <= 255
        c.close()


def test_scan_1280x960_vs_oracle():
    """BASELINE configs[4] image size through the loop-mode scan (960-/1200-point column plans, 1280-point row plan): 8 keyframes x 2
    hypotheses, every candidate against the C oracle, full and image-only store modes (same bits)."""
    import ni_slam_b200 as nis
    h, w = 960, 1280
    db, q_u8, j = _db_fixture(8, h, w, size=4096)
    cfg4 = oc.make_cfg(height=h, width=w)
    orc = _oracle_records(cfg4, db, q_u8, workers=8)
    assert int(np.argmax([i.sum() for i, _, _ in orc])) == j
    prev = None
    for mode in (nis.DB_FULL, nis.DB_IMAGE):
        c = nis.CorrelationFlow(nis.CFConfig(), h, w)
        c.set_batch(3)                                               # three batches, ragged tail
        lc = nis.LoopClosure(nis.LoopClosureConfig(30, 60), c)
        lc.SetMode(mode)
        lc.AddImages(db, list(range(8)))
        res, recs = lc.FindLoopClosureRecords(c.ComputeIntermedium(q_u8), 99, 1e6)
        assert res.evaluated == 8 and res.loop_slot == j and res.found
        _compare_records(recs, orc)
        if prev is not None:
            assert np.array_equal(recs["response"], prev["response"]) and np.array_equal(recs["relative_pose"], prev["relative_pose"])
        prev = recs.copy()
        c.close()


def test_gaussian_kernel_scan_vs_oracle(imgs):
    """gaussian kernel through the loop-mode scan against the C oracle (which tests/test_oracle_ref.py holds equal to the compiled
    reference): 7 keyframes x 2 hypotheses."""
    import ni_slam_b200 as nis
    cfg_g = oc.make_cfg(kernel=1)
    c = nis.CorrelationFlow(nis.CFConfig(kernel=1), H, W)
    lc = nis.LoopClosure(nis.LoopClosureConfig(30, 60), c)
    order = [3, 0, 5, 1, 4, 6, 7]
    lc.AddImages(imgs[order])
    q = c.ComputeIntermedium(imgs[2])
    res, recs = lc.FindLoopClosureRecords(q, 99, 50.0)
    orc = _oracle_records(cfg_g, imgs[order], imgs[2], workers=7)
    _compare_records(recs, orc, kernel=1)
    best = int(np.argmax([i.sum() for i, _, _ in orc]))
    assert res.loop_slot == best
    c.close()
