"""Small end-to-end exercise of every kernel family for compute-sanitizer (GPU box):
   compute-sanitizer --tool memcheck python tools/sanitize_probe.py        (also --tool racecheck / initcheck)
Tiny geometry (96x128 images, 80x64 polar grid) so the instrumented run stays short; plus one 640x480 pair."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ni_slam_b200 as nis

rng = np.random.default_rng(0)


def texture(h, w, n, step=3):
    big = rng.random((h + 64, w + 64)).astype(np.float32)
    k = np.ones((5, 5), np.float32) / 25
    from numpy.lib.stride_tricks import sliding_window_view
    sm = (sliding_window_view(np.pad(big, 2, mode="wrap"), (5, 5)) * k).sum((-1, -2))
    sm = (sm - sm.min()) / (sm.max() - sm.min())
    return np.stack([np.rint(sm[8 + step * t:8 + step * t + h, 8 + 2 * t:8 + 2 * t + w] * 255).astype(np.uint8) for t in range(n)])


def run(h, w, d, cp, nframes):
    cf = nis.CorrelationFlow(nis.CFConfig(rotation_divisor=d, rotation_channel=cp), h, w, device=0)
    fr = texture(h, w, nframes)
    poses, infos = cf.TrackStream(fr)
    fa, fb = cf.ComputeIntermedium(fr[0]), cf.ComputeIntermedium(fr[1])
    for mode in (True, False):
        cf.ComputePose(fa, fb, mode)
    lc = nis.LoopClosure(nis.LoopClosureConfig(), cf)
    lc.AddImages(fr[: max(3, nframes // 2)])
    lc.FindLoopClosure(fb, 100, 10.0)
    kf = cf.TrackStreamKeyframes(fr, nis.KeyframeSelectionConfig(max_distance=0.004, max_angle=0.03), nis.CameraModel(cx=w / 2 - 1, cy=h / 2 + 1))
    st = nis.MapStitcher(50, nis.CameraModel(fx=300.0, fy=300.0, cx=w / 2, cy=h / 2), h, w, cell_x0=-6, cell_y0=-6, cells_x=12, cells_y=12)
    for t in range(3):
        st.InsertFrame(fr[t], [0.05 * t, -0.02 * t, 0.3 * t])
    st.RecomputeOccupancy([[0.0, 0.0, 0.1 * t] for t in range(3)])
    print("%dx%d: poses %s keyframes %d stitched frames %d dropped %d" % (w, h, poses[0], int(kf["inserted"].sum()), st.frames(), st.dropped()), flush=True)
    st.close(); cf.close()


run(96, 128, 80, 64, 9)
if "--full" in sys.argv:
    run(480, 640, 720, 480, 3)
print("probe done")
