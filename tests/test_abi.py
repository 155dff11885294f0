"""CPU tests: the C-ABI library builds, loads and exports every symbol include/nislam.h declares; the emulated CUDA
FFT passes (tests/cpp/emu_fft.cc replays each CTA thread by thread on the host) agree with numpy."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from ni_slam_b200 import build, api
    lib_path = build.build()
    lib = C.CDLL(lib_path)
    hdr = open(os.path.join(ROOT, "include", "nislam.h")).read()
    declared = sorted(set(re.findall(r"\b(nis_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(api.SYMBOLS) == declared


def test_no_gpu_fails_loudly_not_silently():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import ni_slam_b200 as nis
    with pytest.raises(nis.NisError):
        nis.CorrelationFlow(nis.CFConfig(), 480, 640)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "ni_slam_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cc")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.lower(), os.path.join(dp, f)


@pytest.fixture(scope="module")
def emu():
    so = os.path.join(ROOT, "tests", "cpp", "_build", "libemu.so")
    src = os.path.join(ROOT, "tests", "cpp", "emu_fft.cc")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    deps = [src] + [os.path.join(ROOT, "ni_slam_b200", "csrc", f) for f in ("nis_fft.cuh", "nis_ops.cuh", "nis_sizes.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-std=c++17", "-I", os.path.join(ROOT, "ni_slam_b200", "csrc"),
                               src, "-o", so])
    return C.CDLL(so)


def P(a):
    return a.ctypes.data_as(C.c_void_p)


def test_emulated_register_dfts(emu):
    rng = np.random.default_rng(0)
    for R in (1, 2, 3, 4, 5, 6, 8, 9, 10, 12, 15, 16, 20, 32):
        for inv in (0, 1):
            v = (rng.standard_normal(R) + 1j * rng.standard_normal(R)).astype(np.complex64)
            w = v.copy()
            assert emu.emu_dft(P(w), R, inv) == 0
            ref = np.fft.ifft(v) * R if inv else np.fft.fft(v)
            assert np.abs(w - ref).max() / np.abs(ref).max() < 1e-6, (R, inv)


@pytest.mark.parametrize("N,plan", [(640, "a"), (480, "a"), (1280, "a"), (128, "a"), (64, "a"), (640, "b"), (480, "b")])
def test_emulated_row_pass(emu, N, plan):
    """plan a = three stages (radix 16 first; 128 / 64 are two-stage with R2 = 1), plan b = the two-stage 32 x 20 / 32 x 15 plans"""
    rng = np.random.default_rng(N)
    nl = 11          # not a multiple of the CTA's line count: exercises the ragged last CTA
    x = (rng.standard_normal((nl, N)) + 1j * rng.standard_normal((nl, N))).astype(np.complex64)
    fn = emu.emu_row if plan == "a" else emu.emu_row_b
    for inv in (0, 1):
        out = np.zeros_like(x)
        assert fn(P(x), nl, N, inv, P(out)) == 0
        ref = np.fft.ifft(x.astype(np.complex128), axis=1) * N if inv else np.fft.fft(x.astype(np.complex128), axis=1)
        assert np.abs(out - ref).max() / np.abs(ref).max() < 2e-6


@pytest.mark.parametrize("N,W", [(480, 64), (720, 32), (960, 32), (1200, 16), (96, 64), (80, 32), (64, 16)])
def test_emulated_column_pass(emu, N, W):
    rng = np.random.default_rng(N)
    B = 2
    x = rng.standard_normal((B, N, W)).astype(np.float32)
    out = np.zeros((B, N // 2 + 1, W), np.complex64)
    assert emu.emu_col_fwd_f32(P(x), B, N, W, P(out)) == 0
    ref = np.fft.rfft(x.astype(np.float64), axis=1)
    assert np.abs(out - ref).max() / np.abs(ref).max() < 2e-6
    if W % 32 == 0:          # the 16-lane geometry of the rotation pass: same arithmetic per column, so the same bits
        out16 = np.zeros_like(out)
        assert emu.emu_col_fwd_f32_l16(P(x), B, N, W, P(out16)) == 0
        assert np.array_equal(out16, out)
    back = np.zeros((B, N, W), np.float32)
    assert emu.emu_col_inv_store(P(out), B, N, W, P(back)) == 0
    assert np.abs(back * W - x).max() < 1e-5
    # peak epilogue: arg-max (column-major first), sum, sum of squares
    stats = np.zeros((B, 3), np.uint64)
    g = np.zeros((B, N, W), np.float32)
    assert emu.emu_col_inv_peak(P(out), B, N, W, P(stats), P(g)) == 0
    for b in range(B):
        key = int(stats[b, 0])
        col, row = divmod(0xFFFFFFFF - (key & 0xFFFFFFFF), N)
        assert (col, row) == divmod(int(np.argmax(g[b].T.reshape(-1))), N)
        s, q = stats[b, 1:].view(np.float64)
        assert abs(q / np.square(g[b].astype(np.float64)).sum() - 1) < 1e-5


def test_argmax_tiebreak_rule_in_kernel_key(emu):
    # duplicated maxima -> smallest col, then smallest row.  Feed a spectrum whose inverse is exactly
    # constant (only the DC bin set) so every element ties: the winner must be (row 0, col 0).
    N, W = 96, 64
    spec = np.zeros((1, N // 2 + 1, W), np.complex64)
    spec[0, 0, 0] = N          # after the (omitted) row pass this is "DC per column": every column gets the same constant
    spec[0, 0, :] = N
    stats = np.zeros((1, 3), np.uint64)
    g = np.zeros((1, N, W), np.float32)
    assert emu.emu_col_inv_peak(P(spec), 1, N, W, P(stats), P(g)) == 0
    assert np.all(g == g[0, 0, 0])
    key = int(stats[0, 0])
    assert divmod(0xFFFFFFFF - (key & 0xFFFFFFFF), N) == (0, 0)


def test_cpp_shim_compiles_against_the_abi():
    """The reference-facing C++ shim (host/correlation_flow.hpp) and its driver compile and link against libnislam.so."""
    from ni_slam_b200 import build
    build.build()
    exe = os.path.join(ROOT, "tests", "cpp", "_build", "shim_test")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O1", os.path.join(ROOT, "tests", "cpp", "shim_test.cc"), "-o", exe, "-L",
                           os.path.join(ROOT, "ni_slam_b200", "lib"), "-lnislam",
                           "-Wl,-rpath," + os.path.join(ROOT, "ni_slam_b200", "lib")])
    assert os.path.exists(exe)


@pytest.mark.parametrize("N,W", [(480, 32), (720, 32), (960, 16), (96, 64), (80, 16), (64, 16)])
def test_emulated_fused_colcol(emu, N, W):
    """inverse column pass -> (x/n + offset)^3 -> forward column pass, the real kernel image only in shared memory"""
    rng = np.random.default_rng(N + 1)
    B = 2
    real = rng.standard_normal((B, N, W)).astype(np.float32) * 50
    spec = np.fft.rfft(real.astype(np.float64), axis=1).astype(np.complex64)
    out = np.zeros_like(spec)
    mx = np.zeros(B, np.uint32)
    assert emu.emu_colcol_poly(P(spec), B, N, W, P(out), P(mx), C.c_float(0.1), 3) == 0
    n = float(N * W)
    k = ((real.astype(np.float64) * N / n).astype(np.float32) + np.float32(0.1)).astype(np.float64) ** 3
    want = np.fft.rfft(k, axis=1)
    assert np.abs(out - want).max() / np.abs(want).max() < 3e-6
    assert np.allclose(mx.view(np.float32), np.abs(k).reshape(B, -1).max(axis=1), rtol=1e-5)


@pytest.mark.parametrize("N,plan", [(640, "a"), (480, "a"), (128, "a"), (640, "b"), (480, "b")])
def test_emulated_fused_rowrow(emu, N, plan):
    """forward row pass -> element-wise (x conj z | H x / max) -> inverse row pass"""
    mulconj = emu.emu_rowrow_mulconj if plan == "a" else emu.emu_rowrow_mulconj_b
    filt = emu.emu_rowrow_filter if plan == "a" else emu.emu_rowrow_filter_b
    rng = np.random.default_rng(N + 2)
    B, nrows = 2, 7
    x = (rng.standard_normal((B, nrows, N)) + 1j * rng.standard_normal((B, nrows, N))).astype(np.complex64)
    z = (rng.standard_normal((B, nrows, N)) + 1j * rng.standard_normal((B, nrows, N))).astype(np.complex64)
    out = np.zeros_like(x)
    xx = np.zeros(B, np.float64)
    assert mulconj(P(x), P(z), B, nrows, N, P(out), P(xx)) == 0
    X = np.fft.fft(x.astype(np.complex128), axis=2)
    want = np.fft.ifft(X * np.conj(z), axis=2) * N
    assert np.abs(out - want).max() / np.abs(want).max() < 3e-6
    assert np.allclose(xx, (np.abs(X) ** 2).sum(axis=(1, 2)), rtol=1e-5)
    mx = np.array([2.0, 0.5], np.float32).view(np.uint32)
    out2 = np.zeros_like(x)
    assert filt(P(x), P(z), P(mx), B, nrows, N, P(out2)) == 0
    want2 = np.fft.ifft(X * z / np.array([2.0, 0.5]).reshape(B, 1, 1), axis=2) * N
    assert np.abs(out2 - want2).max() / np.abs(want2).max() < 3e-6


@pytest.mark.parametrize("N,plan_b", [(480, 0), (480, 1), (640, 1)])
def test_emulated_fused_store_and_square(emu, N, plan_b):
    """forward row pass -> store the spectrum, continue with |.|^2 -> inverse row pass (fft_polar and the first step of its Kzz)"""
    rng = np.random.default_rng(N + 3)
    B, nrows = 2, 5
    x = (rng.standard_normal((B, nrows, N)) + 1j * rng.standard_normal((B, nrows, N))).astype(np.complex64)
    f = np.zeros_like(x)
    out = np.zeros_like(x)
    assert emu.emu_rowrow_storesq(P(x), B, nrows, N, P(f), P(out), plan_b) == 0
    X = np.fft.fft(x.astype(np.complex128), axis=2)
    assert np.abs(f - X).max() / np.abs(X).max() < 2e-6
    want = np.fft.ifft(np.abs(X) ** 2, axis=2) * N
    assert np.abs(out - want).max() / np.abs(want).max() < 3e-6


def test_host_pose_math_matches_tracker_restatement():
    """ni_slam_b200/host/pose_math.hpp (gate, pose composition, keyframe test of nis_track_stream_keyframes) against the Python
    restatement of map_builder.cc:30-70 on 300 scripted ComputePose outputs: decisions identical, poses to 1e-12."""
    import numpy as np
    import tracker_ref as tr
    exe = os.path.join(ROOT, "tests", "cpp", "_build", "pose_math_test")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O1", os.path.join(ROOT, "tests", "cpp", "pose_math_test.cc"), "-o", exe])
    rng = np.random.default_rng(7)
    n = 300
    E = [0.6, -0.8, 0.1, 0.8, 0.6, -0.2, 0.0, 0.0, 1.0]
    cam = dict(fx=812.5, fy=790.25, cx=331.5, cy=236.25, height=0.37)
    kfs = (0.02, 0.03, 30.0, 90.0)
    resp = np.stack([rng.uniform(5, 200, n), np.zeros(n), rng.uniform(5, 200, n)], 1)
    resp[:, 1] = resp[:, 0]
    pose = np.stack([rng.integers(-25, 26, n), rng.integers(-25, 26, n), np.deg2rad(rng.integers(-720, 721, n) * 0.5)], 1).astype(np.float64)
    text = "%r %r %r %r %r " % (cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["height"]) + " ".join(repr(e) for e in E)
    text += " %r %r %r %r 640 480 %d\n" % (kfs + (n,))
    text += "\n".join(" ".join(repr(float(v)) for v in list(resp[i]) + list(pose[i])) for i in range(n)) + "\n"
    out = subprocess.run([exe], input=text, capture_output=True, text=True, check=True).stdout.split("\n")
    it = iter(range(n))
    trk = tr.MapBuilderTracker(tr.Camera(extrinsics=E, image_width=640, image_height=480, **cam), *kfs, lambda img: (None, None),
                               lambda *a: (lambda i: (resp[i], pose[i]))(next(it)))
    n_ins = 0
    for i in range(n + 1):
        o = trk.add_new_input(None)
        v = [float(x) for x in out[i].split()]
        assert (int(v[0]), int(v[1]), int(v[2])) == (int(o["tracked"]), int(o["inserted"]), o["keyframe"]), i
        assert np.allclose(v[3:6], o["cf_pose"], rtol=0, atol=1e-9) and np.allclose(v[6:9], o["pose"], rtol=0, atol=1e-9), i
        assert abs(v[9] - o["distance"]) < 1e-12
        if i:
            assert np.allclose(v[10:13], o["relative_pose"], rtol=0, atol=1e-9)
        n_ins += int(o["inserted"])
    assert 10 < n_ins < n
