"""Times the phases of a loop-closure query (features of the query, scan) over a synthetic keyframe store (GPU box)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ni_slam_b200 as nis, bench_synth as bs
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 12
dev = torch.device("cuda:0")
cf = nis.CorrelationFlow(nis.CFConfig(), 480, 640, device=0)
lc = nis.LoopClosure(nis.LoopClosureConfig(position_response_thr=60, angle_response_thr=60), cf)
canvas = bs.make_canvas(4096, seed=0, device=dev)
gcx, gcy, gang = bs.db_poses(n, seed=1)
for c0 in range(0, n, 2048):
    c1 = min(n, c0 + 2048)
    imgs = bs.crops(canvas, gcx[c0:c1], gcy[c0:c1], gang[c0:c1])
    lc.AddImages(None, np.arange(c0, c1, dtype=np.int32), None, ptr=imgs.data_ptr(), n=c1 - c0, on_device=True)
    del imgs
j = n // 2 + 3
q = bs.crops(canvas, [gcx[j] + 13], [gcy[j] - 7], [gang[j] + 4.5])[0].cpu().numpy()
for i in range(nq):
    t0 = time.perf_counter()
    qf = cf.ComputeIntermedium(q)
    t1 = time.perf_counter()
    res = lc.FindLoopClosure(qf, current_frame_id=10 ** 9)
    t2 = time.perf_counter()
    del qf
    t3 = time.perf_counter()
    print("query %2d features %.2f ms scan %.2f ms free %.2f ms winner %d" % (i, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, res.loop_frame_id), flush=True)
