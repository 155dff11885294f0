#!/usr/bin/env python
"""bench.py -- headline benchmark of the NI-SLAM tracking / loop-closure hot path on B200.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference ...        # the reference's CPU path (oracle port) on the host cores

Workload (BASELINE.json configs[1]): a synthetic 640x480 u8 stream, every frame a keyframe; one STEP =
one pass of the hot path over the whole stream: per frame ComputeIntermedium + ComputePose(tracking) = one 3-DoF pose
solve.  `value` = pose solves/s with the frames resident in HBM; `e2e` = the same through the public call with the
frames in pinned HOST memory (H2D of the frames and D2H of the poses inside the timed region).
Per-frame tracking does not shard: N > 1 runs N independent replicas (weak scaling, no collective).
The loop-closure scan (configs[2..3]) is reported beside it in `loop_closure`: a keyframe DB sharded by index over the
ranks, one broadcast query, one NCCL all-gather of the per-rank best records.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, D, CP = 480, 640, 720, 480
BYTES_PER_SOLVE = H * W + 2 * ((H // 2 + 1) * W * 8 + (D // 2 + 1) * CP * 8)        # 5 547 520
BYTES_PER_CANDIDATE = (H // 2 + 1) * W * 8 + (D // 2 + 1) * CP * 8                  # 2 620 160
METRIC = "pose_solves_per_sec_640x480"


NCU_FAMILY = {"colcol": ("colcol_kernel",), "rowrow_filter": ("rowrow_kernel", "MidFilterH"), "rowrow_storeabs": ("rowrow_kernel", "MidStoreAbs"),
              "rowrow_mulconj": ("rowrow_kernel", "MidMulConjZ"), "rowrow_storesq": ("rowrow_kernel", "MidStoreSq"), "row_inv_mulconj": ("row_kernel", "ProMulConj"), "row_fwd_h": ("row_kernel", "EpiHStore"),
              "row_fwd": ("row_kernel", "ProSpec, EpiSpecStore"), "col_fwd_rotate": ("col_fwd_kernel", "ProRotate"), "col_fwd_u8": ("col_fwd_kernel", "ProRealU8"),
              "col_fwd_f32": ("col_fwd_kernel", "ProRealF32"), "col_inv_peak": ("col_inv_kernel", "EpiPeak"),
              "col_inv_store_shift": ("col_inv_kernel", "EpiStoreShift"), "polar_tma": ("polar_tma_kernel",), "rzc_fix": ("rzc_fix_kernel",)}
NCU_SUMMARY = "ncu_r02_summary.json"


def ncu_traffic(kernel_family):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel family, from the committed ncu --set full
    capture of the same command (profiles/ncu_r02_summary.json: cold-cache replay, 56 images per launch = the default batch; mean over
    the family's instantiations, e.g. the 480- and the 720-point colcol).  None if no capture is committed."""
    try:
        with open(os.path.join(ROOT, "profiles", NCU_SUMMARY)) as f:
            cap = json.load(f)["tracking"]["full_capture"]
        pat = NCU_FAMILY[kernel_family]
        vals = [(v["dram_read_MB"] + v["dram_write_MB"]) * 1e6 for k, v in cap.items() if all(p in k for p in pat)]
        return float(np.mean(vals)) if vals else None
    except Exception:
        return None


def bind_to_gpu_cpus(torch, local_rank):
    """Multi-rank runs: pin this rank to the CPUs NVML reports as local to its GPU, so that the pinned host buffers of the e2e leg
    are first-touched on that GPU's NUMA node (eight ranks otherwise stream 2.5 GB of frames per step through one socket's memory
    controllers).  Returns the number of CPUs bound, or None when the topology is not available.  Not applied at N = 1, where the
    CPU baseline of the same invocation must see every core."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(local_rank)
        bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (w >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=self.tmp, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, mx, reasons, pw = [], [], set(), []
        for line in self.tmp.read().splitlines():
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9 or p[0] != str(self.index):
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2])); pw.append(float(p[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v == "Active":
                    reasons.add(name)
        try:
            os.unlink(self.tmp.name)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU implementation of the path (oracle port; the reference cannot be compiled in
# this image -- no Eigen / FFTW / OpenCV C++ -- see DESIGN.md) on all host cores
# ------------------------------------------------------------------------------------------------------------------
def cpu_frames(n, seed=0):
    import torch
    import bench_synth as bs
    canvas = bs.make_canvas(2048, seed=seed, device="cpu")
    cx, cy, ang = bs.stream_poses(n, seed=seed, size=2048)
    return bs.crops(canvas, cx, cy, ang, H, W).numpy()


def time_oracle_stream(frames, threads):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_c as oc
    cfg = oc.make_cfg(height=H, width=W)
    t0 = time.perf_counter()
    oc.track_stream(cfg, frames, threads=threads)
    dt = time.perf_counter() - t0
    return (frames.shape[0] - 1) / dt, dt


def time_python_stream(frames, threads):
    """The same stream through oracle/nislam_ref.py (scipy pocketfft f32 + genuine cv2), pairs spread over a thread pool
    (scipy.fft, cv2 and the big numpy ops release the GIL).  Returns (solves/s, seconds) or None if scipy/cv2 are missing."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        import nislam_ref as ref
        if ref.cv2 is None:
            return None
        ref.cv2.setNumThreads(1)
    except Exception:
        return None
    from concurrent.futures import ThreadPoolExecutor
    import threading
    n = frames.shape[0]
    local = threading.local()

    def cf():
        if not hasattr(local, "cf"):
            local.cf = ref.CorrelationFlow(ref.CFConfig(), H, W)
        return local.cf

    def feat(t):
        img = ref.convert_mat_to_normalized_array(frames[t])
        return (img,) + tuple(cf().compute_intermedium(img))

    def pose(t):
        return cf().compute_pose(feats[t - 1][1], feats[t][0], feats[t - 1][2], feats[t][2], True)

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=max(1, threads)) as ex:
        feats = list(ex.map(feat, range(n)))
        list(ex.map(pose, range(1, n)))
    dt = time.perf_counter() - t0
    return (n - 1) / dt, dt


_MP = {}


def _mp_init():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import nislam_ref as ref
    ref.cv2.setNumThreads(1)
    _MP["ref"] = ref
    _MP["cf"] = ref.CorrelationFlow(ref.CFConfig(), H, W)


def _mp_chunk(frames):
    """one worker: features of its frames, then the pairs inside its chunk (chunks overlap by one frame)"""
    ref, cf = _MP["ref"], _MP["cf"]
    feats = []
    for f in frames:
        img = ref.convert_mat_to_normalized_array(f)
        feats.append((img,) + tuple(cf.compute_intermedium(img)))
    out = []
    for t in range(1, len(feats)):
        out.append(cf.compute_pose(feats[t - 1][1], feats[t][0], feats[t - 1][2], feats[t][2], True))
    return len(out)


def time_python_stream_mp(frames, procs, pool=None):
    """The scipy+cv2 restatement on `procs` worker PROCESSES (no GIL): the stream is cut into contiguous chunks that overlap by
    one frame.  Pool start-up and imports are outside the timed region."""
    import multiprocessing as mp
    try:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import nislam_ref as ref
        if ref.cv2 is None:
            return None
    except Exception:
        return None
    n = frames.shape[0]
    procs = max(1, min(procs, n - 1))
    own = pool is None
    if own:
        pool = mp.get_context("spawn").Pool(procs, initializer=_mp_init)
        pool.map(_mp_chunk, [frames[:2]] * procs)                 # warm-up: imports, FFT plans, cv2 init
    bounds = np.linspace(0, n - 1, procs + 1).astype(int)
    chunks = [frames[bounds[i]:bounds[i + 1] + 1] for i in range(procs) if bounds[i + 1] > bounds[i]]
    t0 = time.perf_counter()
    done = sum(pool.map(_mp_chunk, chunks))
    dt = time.perf_counter() - t0
    if own:
        pool.close()
        pool.join()
    assert done == n - 1
    return (n - 1) / dt, dt


def ref_available():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        import oracle_ref as orf
        if not orf.available():
            return False
        orf.lib()
        return True
    except Exception:
        return False


def time_ref_stream(frames, threads):
    """The stream through oracle/_ref = the reference's OWN correlation_flow.cc / utils.cc compiled unmodified (oracle/Makefile.ref):
    per frame ConvertMatToNormalizedArray + ComputeIntermedium, per pair ComputePose(tracking), exactly MapBuilder's call sequence
    (map_builder.cc:72-75, :129); one CorrelationFlow object per thread (the reference's is not re-entrant), frames / pairs spread
    over a thread pool (the library calls release the GIL).  Returns (solves/s, seconds)."""
    import threading
    from concurrent.futures import ThreadPoolExecutor
    import oracle_c as oc
    import oracle_ref as orf
    cfg = oc.make_cfg(height=H, width=W)
    local = threading.local()

    def cf():
        if not hasattr(local, "cf"):
            local.cf = orf.CorrelationFlow(cfg)
        return local.cf

    def feat(t):
        img = orf.normalize_u8(frames[t])
        return (img,) + tuple(cf().compute_intermedium(img))

    def pose(t):
        return cf().compute_pose(feats[t - 1][1], feats[t][0], feats[t - 1][2], feats[t][2], True)
    n = frames.shape[0]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=max(1, threads)) as ex:
        feats = list(ex.map(feat, range(n)))
        list(ex.map(pose, range(1, n)))
    dt = time.perf_counter() - t0
    return (n - 1) / dt, dt


def time_cpu_scan(db_u8, q_u8, threads, use_ref):
    """LoopClosure::FindLoopClosure over len(db_u8) candidates on the CPU (loop_closure.cc:36-73: per candidate ComputePose(..., false)):
    oracle/_ref when built, else the C port; candidates spread over `threads` threads.  Keyframe features are computed outside the timed
    region (the reference reads them from its Frames).  Returns (candidates/s, seconds, winner index)."""
    from concurrent.futures import ThreadPoolExecutor
    import threading
    import oracle_c as oc
    cfg = oc.make_cfg(height=H, width=W)
    local = threading.local()
    if use_ref:
        import oracle_ref as orf

        def cf():
            if not hasattr(local, "cf"):
                local.cf = orf.CorrelationFlow(cfg)
            return local.cf
        feat = lambda u: cf().compute_intermedium(oc.normalize_u8(u))
        solve = lambda k: cf().compute_pose(feats[k][0], qi, feats[k][1], qP, False)[0]
    else:
        feat = lambda u: oc.compute_intermedium(cfg, oc.normalize_u8(u))
        solve = lambda k: oc.compute_pose(cfg, feats[k][0], qi, feats[k][1], qP, False)[0]
    with ThreadPoolExecutor(max_workers=max(1, threads)) as ex:
        feats = list(ex.map(feat, db_u8))
        qi = oc.normalize_u8(q_u8)
        qP = feat(q_u8)[1]
        t0 = time.perf_counter()
        infos = list(ex.map(solve, range(len(db_u8))))
        dt = time.perf_counter() - t0
    return len(db_u8) / dt, dt, int(np.argmax([i.sum() for i in infos]))


def best_cpu_stream(frames, threads):
    """Times both CPU restatements on the same frames and returns the faster one: (solves/s, seconds, label)."""
    v_c, dt_c = time_oracle_stream(frames, threads)
    best = (v_c, dt_c, "oracle/nislam_oracle.c (dependency-free C, OpenMP)")
    py = time_python_stream(frames, threads)
    if py is not None and py[0] > best[0]:
        best = (py[0], py[1], "oracle/nislam_ref.py (scipy pocketfft f32 + cv2, thread pool)")
    mp_ = time_python_stream_mp(frames, threads) if threads > 1 else None
    if mp_ is not None and mp_[0] > best[0]:
        best = (mp_[0], mp_[1], "oracle/nislam_ref.py (scipy pocketfft f32 + cv2, %d worker processes)" % threads)
    return best + ({"c_port": v_c, "python_scipy_cv2_threads": None if py is None else py[0],
                    "python_scipy_cv2_processes": None if mp_ is None else mp_[0]},)


def workload_config(n_frames, world):
    """The `config` object both arms print (same workload, same keys): BASELINE.json configs[1]."""
    return {"workload": "tracking stream %dx%d u8, every frame a keyframe, ComputeIntermedium + ComputePose(tracking) per frame " % (W, H) +
                        ("(BASELINE.json configs[1])" if (W, H) == (640, 480) else "(BASELINE.json configs[4] image size)"), "frames_per_step_per_gpu": n_frames, "solves_per_step": (n_frames - 1) * world,
            "rotation_divisor": D, "rotation_channel": CP, "kernel": "polynomial",
            "parallelism": "replicas only (tracking does not shard)" if world > 1 else "1 GPU",
            "l2": "inputs larger than L2: %.0f MB of frames + %.0f MB of features per step vs 126 MB L2" %
                  (n_frames * H * W / 1e6, n_frames * BYTES_PER_CANDIDATE / 1e6)}


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path on the host cores.  oracle/_ref (the reference's
    correlation_flow.cc / utils.cc compiled unmodified, kind "reference") when it is built, else the C port (kind "port").  The
    workload and `config` are the GPU arm's; each step times a bounded sample of that stream (args.ref_frames frames), the value is
    per-solve throughput, so the sample length does not enter it."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    threads = cores
    n = args.ref_frames
    frames = cpu_frames(n)
    use_ref = ref_available()
    warm = frames[:min(n, 2 * threads + 1)]
    others = {}
    if use_ref:
        time_ref_stream(warm, threads)
        timer, kind, label = time_ref_stream, "reference", "oracle/_ref = /root/reference/src/{correlation_flow,utils}.cc compiled unmodified (oracle/Makefile.ref)"
        try:
            others["c_port_solves_per_sec"] = time_oracle_stream(warm, threads)[0]
        except Exception:
            pass
    else:
        _, _, label, others = best_cpu_stream(warm, threads)
        kind = "port"
        timer = time_python_stream if label.startswith("oracle/nislam_ref.py (scipy pocketfft f32 + cv2, thread") else time_oracle_stream
    times = []
    for _ in range(args.steps):
        _, dt = timer(frames, threads)
        times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = (n - 1) / (ms / 1e3)
    sample = "%d-frame sample of the stream (%d solves) per step, %s, %d threads; other CPU restatements on the warm-up frames: %s" % (
        n, n - 1, label, threads, json.dumps(others))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args.frames, max(1, args.gpus)),
            "cpu_baseline": {"value": value, "unit": "solves/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def protect_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner / NCCL_DEBUG output to
    stdout), so file descriptor 1 is pointed at stderr for the whole run and the JSON line goes to a private duplicate of the
    original stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def set_size(size):
    """Image geometry of the run (module globals: every helper reads them at call time)."""
    global H, W, BYTES_PER_SOLVE, BYTES_PER_CANDIDATE, METRIC
    W, H = (int(v) for v in size.lower().split("x"))
    BYTES_PER_CANDIDATE = (H // 2 + 1) * W * 8 + (D // 2 + 1) * CP * 8
    BYTES_PER_SOLVE = H * W + 2 * BYTES_PER_CANDIDATE
    METRIC = "pose_solves_per_sec_%dx%d" % (W, H)


def main():
    protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", default="640x480", help="image size WxH: 640x480 (BASELINE configs[1..3], the headline) or 1280x960 (configs[4])")
    ap.add_argument("--frames", type=int, default=1000, help="frames per step (stream length)")
    ap.add_argument("--batch", type=int, default=0, help="pairs in flight per kernel launch (0 = library default)")
    ap.add_argument("--lanes", type=int, default=0, help="concurrent CUDA streams the batches are dealt to (0 = library default)")
    ap.add_argument("--db", type=int, default=-1, help="loop-closure keyframes PER GPU (0 = skip the scan section; -1 = BASELINE configs: 10k on one GPU, "
                                                     "100k sharded over N GPUs, capped by free HBM)")
    ap.add_argument("--queries", type=int, default=5)
    ap.add_argument("--ref-frames", type=int, default=129, help="frames per step of the CPU reference arm")
    ap.add_argument("--cpu-frames", type=int, default=129, help="frames of the cpu_baseline sample (0 = skip)")
    ap.add_argument("--no-cfg4", dest="cfg4", action="store_false", help="skip the BASELINE configs[4] block (1280x960 stream + 50k-keyframe store)")
    ap.add_argument("--no-extras", dest="extras", action="store_false", help="skip the next-row extras (online latency, undistort, keyframe policy, stitcher)")
    args = ap.parse_args()
    set_size(args.size)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return 0

    import torch
    import torch.distributed as dist
    import bench_synth as bs
    import ni_slam_b200 as nis
    from ni_slam_b200 import build as nis_build

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    nis_build.build()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_cpus(torch, local_rank) if world > 1 else None      # before any pinned allocation (first touch)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_size(args):
        """One full measurement at the current image size (module globals H, W ...); returns the JSON line (rank 0) or None."""
        cf = nis.CorrelationFlow(nis.CFConfig(), H, W, device=local_rank)
        if args.batch:
            cf.set_batch(args.batch)
        if args.lanes:
            cf.set_lanes(args.lanes)
        ext = torch.cuda.ExternalStream(cf.stream, device=dev)

        # ---- synthetic stream (different walk per rank)
        n = args.frames
        canvas = bs.make_canvas(4096, seed=0, device=dev)
        cx, cy, ang = bs.stream_poses(n, seed=100 + rank)
        frames = bs.crops(canvas, cx, cy, ang, H, W)                        # (n, H, W) u8 in HBM
        frames_host = torch.empty((n, H, W), dtype=torch.uint8, pin_memory=True)
        frames_host.copy_(frames)
        torch.cuda.synchronize()

        def step_dev():
            return cf.TrackStreamPtr(frames.data_ptr(), n, on_device=True)

        def step_host():
            return cf.TrackStreamPtr(frames_host.data_ptr(), n, on_device=False)

        for _ in range(args.warmup):
            poses, infos = step_dev()

        # sanity of the answers against the synthetic motion (not timed)
        dth = np.deg2rad(np.diff(ang))
        dmag = np.hypot(np.diff(cx), np.diff(cy))
        ok = (np.abs((poses[:, 2] - dth + np.pi) % (2 * np.pi) - np.pi) < np.deg2rad(0.75)) & \
             (np.abs(np.hypot(poses[:, 0], poses[:, 1]) - dmag) < 2.0)
        pose_ok_frac = float(ok.mean())

        # ---- timed region: HBM-resident inputs
        clocks = ClockSampler(local_rank)
        clocks.start()
        for _ in range(2):                 # keep the GPU under the same load while nvidia-smi spins up (samples every 100 ms)
            step_dev()
        l0 = cf.kernel_launches()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        for _ in range(args.steps):
            step_dev()
        e1.record(ext)
        barrier()
        clk = clocks.stop()
        ms_total = e0.elapsed_time(e1)
        launches = cf.kernel_launches() - l0
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step = float(t.item()) / args.steps
        solves_per_step = (n - 1) * world
        value = solves_per_step / (ms_step / 1e3)

        # ---- e2e: same call with HOST buffers (pinned), H2D + D2H inside
        for _ in range(1):
            step_host()
        barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record(ext)
        t_host0 = time.perf_counter()
        for _ in range(args.steps):
            step_host()
        e3.record(ext)
        barrier()
        wall_e2e = (time.perf_counter() - t_host0) / args.steps
        t2 = torch.tensor([max(e2.elapsed_time(e3) / args.steps, wall_e2e * 1e3)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        e2e_value = solves_per_step / (float(t2.item()) / 1e3)

        # ---- per-kernel-family device time (one extra, untimed-for-value step with events around every launch)
        cf.set_lanes(1)                 # one lane so that the event pairs of different kernels do not overlap
        cf.profile_begin()
        step_dev()
        prof = cf.profile_end()
        cf.set_lanes(args.lanes)
        tot = sum(v["ms"] for v in prof.values()) or 1.0
        kernels = {k: {"launches": v["launches"], "ms": round(v["ms"], 4), "share": round(v["ms"] / tot, 4)} for k, v in
                   sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
        dom = next(iter(kernels))
        peak, peak_src = measured_peak()
        achieved = (value / world) * BYTES_PER_SOLVE / 1e9                  # per GPU
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(dom),
                    "traffic_note": "DRAM bytes per launch of the dominant kernel family (ncu --set full of this command, cold cache, 56 images per "
                                    "launch = the default batch, profiles/%s); the same capture sums to 38.6 MB of DRAM traffic and 8.6 M warp "
                                    "instructions per solve over the whole multi-kernel step" % NCU_SUMMARY,
                    "definition": "per-GPU solves/s x %d algorithmic B/solve (SURVEY 8d) over the whole multi-kernel step" % BYTES_PER_SOLVE,
                    "peak_source": peak_src, "dominant_kernel": dom, "dominant_kernel_share": kernels[dom]["share"],
                    "dominant_kernel_avg_launch_ms": kernels[dom]["ms"] / max(kernels[dom]["launches"], 1), "kernels": kernels}
        if dom == "colcol":     # the fused column kernel reads and writes one half spectrum per image: its own compulsory bytes per launch
            c_t, c_p = (H // 2 + 1) * W * 8, (D // 2 + 1) * CP * 8
            per_step = n * 4 * (c_t + c_p)                                   # per frame 2 launches per size (Kzz and Kxz), each in + out
            gbs = per_step / (kernels[dom]["ms"] * 1e-3) / 1e9
            roofline["dominant_kernel_hbm"] = {"what": "colcol: one half spectrum in + out per image and launch (%d B translation size, %d B polar size), "
                                                       "all launches of one step over their summed "
                                                       "live device time" % (2 * c_t, 2 * c_p), "achieved": gbs, "unit": "GB/s", "frac": gbs / peak}

        # ---- next-row measurement: the same stream entering as RAW camera frames through the undistort front end
        front = None
        try:
            if not args.extras:
                raise StopIteration
            hh, ww = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
            xn, yn = (ww - W / 2) / (0.8 * W), (hh - H / 2) / (0.8 * W)
            r2 = xn * xn + yn * yn
            fdist = 1 - 0.25 * r2 + 0.08 * r2 * r2
            sx = np.rint(((xn * fdist) * 0.8 * W + W / 2) * 32).astype(np.int64)
            sy = np.rint(((yn * fdist) * 0.8 * W + H / 2) * 32).astype(np.int64)
            cf.SetUndistortMaps(np.stack([sx >> 5, sy >> 5], axis=-1).astype(np.int16), ((sy & 31) * 32 + (sx & 31)).astype(np.uint16))
            step_dev()
            barrier()
            eu0, eu1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            eu0.record(ext)
            for _ in range(3):
                step_dev()
            eu1.record(ext)
            barrier()
            front = {"what": "Camera::UndistortImage (exact u8 remap) + tracking, frames resident in HBM", "value": (n - 1) / (eu0.elapsed_time(eu1) / 3e3),
                     "unit": "solves/s per GPU"}
        except StopIteration:
            front = None
        finally:
            cf.SetUndistortMaps(None, None)

        # ---- next-row measurement: the same stream under the reference's keyframe policy (MapBuilder::AddNewInput:
        # tracking against the LAST KEYFRAME, gate, pose composition, keyframe test), from pinned host memory, wall clock around the call
        policy = None
        try:
            if not args.extras:
                raise StopIteration
            kfs = nis.KeyframeSelectionConfig(max_distance=0.06, max_angle=0.0873, lower_response_thr=30.0, upper_response_thr=90.0)
            cam = nis.CameraModel(fx=1000.0, fy=1000.0, cx=W / 2 - 7.0, cy=H / 2 + 4.0, height=1.0)
            cf.TrackStreamKeyframes(frames_host.numpy(), kfs, cam)
            t0 = time.perf_counter()
            res = cf.TrackStreamKeyframes(frames_host.numpy(), kfs, cam)
            dt = time.perf_counter() - t0
            policy = {"what": "nis_track_stream_keyframes: AddNewInput (no loop closure) over the same stream, keyframe when > 60 px or > 5 deg "
                              "from the last keyframe or a confidence inside (30, 90); host frames in, per-frame records out",
                      "value": n / dt, "unit": "frames/s per GPU", "keyframes": int(res["inserted"].sum()), "tracked": int(res["tracked"].sum()),
                      "frames": int(n)}
        except StopIteration:
            policy = None
        except Exception as e:                                   # a next-row extra must never take the headline down
            policy = {"error": str(e)[:200]}

        # ---- next-row measurement: MapStitcher -- InsertFrame of keyframes into the occupancy mosaic, then one
        # RecomputeOccupancy (what follows every pose-graph optimisation); integer work, HBM bound
        stitch = None
        try:
            if not args.extras:
                raise StopIteration
            nst = 64
            cam_s = nis.CameraModel(fx=1000.0, fy=1000.0, cx=W / 2, cy=H / 2, height=1.0)
            ms_ = nis.MapStitcher(1000, cam_s, H, W, cell_x0=-3, cell_y0=-3, cells_x=6, cells_y=6, device=local_rank)
            imgs_s = frames_host.numpy()[:nst]
            rs = np.random.default_rng(3)
            poses_s = np.stack([rs.uniform(-1.5, 1.5, nst), rs.uniform(-1.5, 1.5, nst), rs.uniform(-np.pi, np.pi, nst)], 1)
            ms_.InsertFrame(imgs_s[0], poses_s[0])                # first insert allocates the image chunk and the scatter box: not timed
            t0 = time.perf_counter()
            for f in range(1, nst):
                ms_.InsertFrame(imgs_s[f], poses_s[f])
            t_ins = (time.perf_counter() - t0) * nst / (nst - 1)
            ms_.RecomputeOccupancy(poses_s)
            t0 = time.perf_counter()
            ms_.RecomputeOccupancy(poses_s)
            t_rec = time.perf_counter() - t0
            bytes_frame = H * W * (1 + 16)                       # u8 image in + read-modify-write of data and weight (2 x int32 x 2) per pixel
            stitch = {"what": "MapStitcher: %d keyframes 640x480 into 1000x1000-cell mosaic; InsertFrame from host memory (synchronous per "
                              "frame, like the reference) and RecomputeOccupancy from the stored images" % nst,
                      "insert_frames_per_sec": nst / t_ins, "recompute_frames_per_sec": nst / t_rec,
                      "recompute_algorithmic_GBps": nst / t_rec * bytes_frame / 1e9, "dropped_pixels": ms_.dropped()}
            ms_.close()
        except StopIteration:
            stitch = None
        except Exception as e:
            stitch = {"error": str(e)[:200]}

        # ---- online latency (the reference's per-call surface, main.cpp:51-86 feeds one frame at a time): one frame in, one pose out
        online = None
        try:
            if not args.extras:
                raise StopIteration
            fa = cf.ComputeIntermedium(frames_host[0].numpy())
            lat = []
            for t in range(1, 61):
                t0 = time.perf_counter()
                fb = cf.ComputeIntermedium(frames_host[t].numpy())             # nis_features_u8: H2D of the frame + 12 kernels + sync
                cf.ComputePose(fa, fb, True)                                   # nis_compute_pose: 11 kernels + D2H of the record
                lat.append((time.perf_counter() - t0) * 1e3)
                fa.free()
                fa = fb
            lat = np.sort(lat[10:])
            online = {"what": "nis_features_u8 + nis_compute_pose per frame at batch 1, pageable host frame in, pose out, wall clock",
                      "p50_ms": float(lat[len(lat) // 2]), "p99_ms": float(lat[min(len(lat) - 1, int(0.99 * len(lat)))]),
                      "frames_per_sec": float(1e3 / np.mean(lat))}
            exe = os.path.join(ROOT, "tests", "cpp", "_build", "shim_test")
            if rank == 0 and os.path.exists(exe):                              # the same through the C++ shim's reference signatures
                import re
                out = subprocess.run([exe], capture_output=True, text=True, timeout=120).stdout
                m = re.search(r"shim_online_ms p50 ([0-9.]+) p99 ([0-9.]+)", out)
                if m:
                    online["shim"] = {"what": "CorrelationFlow::ComputeIntermedium + ComputePose through host/correlation_flow.hpp (Eigen-layout host "
                                              "arrays in and out, keyframe operands cached on the device), tests/cpp/shim_test.cc",
                                      "p50_ms": float(m.group(1)), "p99_ms": float(m.group(2))}
        except StopIteration:
            online = None
        except Exception as e:
            online = {"error": str(e)[:200]}

        # ---- loop-closure scan: keyframe store sharded by index over the ranks; ONE library call per query on every rank
        # (nis_loop_scan_sharded: ncclBroadcast of the query image, local scan, one ncclAllGather of the best records, reduction)
        loop = None
        if args.db != 0:
            lc = nis.LoopClosure(nis.LoopClosureConfig(position_response_thr=60, angle_response_thr=60), cf)
            if world > 1:
                ids = [nis.CorrelationFlow.NcclUniqueId() if rank == 0 else None]
                dist.broadcast_object_list(ids, src=0)
                cf.CommInit(ids[0], rank, world)
            free_b = torch.cuda.mem_get_info(dev)[0]
            rec_bytes = {nis.DB_FULL: 2 * BYTES_PER_CANDIDATE, nis.DB_SPECTRA: BYTES_PER_CANDIDATE, nis.DB_IMAGE: H * W}
            mode_name = {nis.DB_FULL: "full (F, P, Ht, Hp: %.2f MB)" % (2 * BYTES_PER_CANDIDATE / 1e6),
                         nis.DB_SPECTRA: "spectra (F, P = the reference's Frame payload: %.2f MB; H recomputed per batch)" % (BYTES_PER_CANDIDATE / 1e6),
                         nis.DB_IMAGE: "image (u8: %.2f MB; features recomputed per batch)" % (H * W / 1e6)}

            def pick_mode(per_gpu):
                for m in (nis.DB_FULL, nis.DB_SPECTRA, nis.DB_IMAGE):
                    if per_gpu * rec_bytes[m] <= 0.80 * free_b:
                        return m
                return nis.DB_IMAGE

            def run_series(total, label):
                """Builds a `total`-keyframe store over the ranks in the richest mode that fits, times args.queries sharded queries."""
                per_gpu = -(-total // world)
                mode = pick_mode(per_gpu)
                if world > 1:
                    tm = torch.tensor([mode], dtype=torch.int64, device=dev)
                    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
                    mode = int(tm.item())
                lc.clear()
                lc.SetMode(mode)
                gcx, gcy, gang = bs.db_poses(per_gpu * world, seed=1)
                g0 = rank * per_gpu
                for c0 in range(0, per_gpu, 2048):                              # keyframes are generated and added in chunks (bounded temporaries)
                    c1 = min(per_gpu, c0 + 2048)
                    db_imgs = bs.crops(canvas, gcx[g0 + c0:g0 + c1], gcy[g0 + c0:g0 + c1], gang[g0 + c0:g0 + c1], H, W)
                    torch.cuda.synchronize()                                     # torch's stream and the library's (non-blocking) streams are not ordered
                    lc.AddImages(None, np.arange(g0 + c0, g0 + c1, dtype=np.int32), None, ptr=db_imgs.data_ptr(), n=c1 - c0, on_device=True)
                    del db_imgs
                jstar = (per_gpu * world) // 2 + 3                                  # planted keyframe (global id)
                q_dev = bs.crops(canvas, [gcx[jstar] + 13], [gcy[jstar] - 7], [gang[jstar] + 4.5], H, W)
                q_host = torch.empty((H, W), dtype=torch.uint8, pin_memory=True)
                q_host.copy_(q_dev[0])
                torch.cuda.synchronize()

                def query():
                    return lc.FindLoopClosureSharded(q_host.numpy() if rank == 0 else None, 0, g0, current_frame_id=10 ** 9)
                res, win, mine = query()
                barrier()
                l1 = cf.kernel_launches()
                per_query = []
                for _ in range(args.queries):                     # each query timed on its own; the median guards against host hiccups
                    barrier()
                    t0 = time.perf_counter()
                    e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e4.record(ext)
                    res, win, mine = query()
                    e5.record(ext)
                    torch.cuda.synchronize()
                    per_query.append(max(e4.elapsed_time(e5), (time.perf_counter() - t0) * 1e3))
                barrier()
                q_ms = float(np.median(per_query))
                tq = torch.tensor([q_ms], dtype=torch.float64, device=dev)
                if world > 1:
                    dist.all_reduce(tq, op=dist.ReduceOp.MAX)
                q_ms = float(tq.item())
                cand_per_s = per_gpu * world / (q_ms / 1e3)
                # the store is dense (many keyframes overlap the query), so the winner need not be the planted one: check that the
                # returned relative pose agrees with the winner's true pose (rotation to 0.75 deg, translation length to 2 px)
                winner_ok = None
                if res.loop_frame_id >= 0:
                    w = int(res.loop_frame_id)
                    qx, qy, qa = gcx[jstar] + 13, gcy[jstar] - 7, gang[jstar] + 4.5
                    dth = np.deg2rad(qa - gang[w])
                    winner_ok = bool(abs((res.relative_pose[2] - dth + np.pi) % (2 * np.pi) - np.pi) < np.deg2rad(0.75) and
                                     abs(np.hypot(res.relative_pose[0], res.relative_pose[1]) - np.hypot(qx - gcx[w], qy - gcy[w])) < 2.0)
                # where one query's device time goes (one extra query with events around every launch, one lane): the query's own
                # features, the rotated-query cache (col_fwd_rotate + row_fwd over all 2 D angles, a fixed cost per query) and the scan
                fam = None
                try:
                    cf.set_lanes(1)
                    cf.profile_begin()
                    query()
                    prof = cf.profile_end()
                    cf.set_lanes(args.lanes)
                    fam = {k: round(v["ms"], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
                    fam["rotated_query_cache_ms"] = round(prof.get("col_fwd_rotate", {"ms": 0.0})["ms"] + prof.get("row_fwd", {"ms": 0.0})["ms"], 3)
                except Exception as e:
                    fam = {"error": str(e)[:200]}
                out = {"workload": label, "db_keyframes": per_gpu * world, "keyframes_per_gpu": per_gpu, "store_mode": mode_name[mode],
                       "kernel_ms_per_query_one_lane": fam,
                       "value": 1e3 / q_ms, "unit": "queries/s", "ms_per_query": q_ms, "ms_per_query_all": [round(x, 3) for x in per_query],
                       "candidates_per_sec": cand_per_s, "winner_frame_id": int(res.loop_frame_id), "winner_global_slot": int(res.loop_slot),
                       "winner_rank": int(win), "planted_frame_id": int(jstar), "found": bool(res.found),
                       "relative_pose": [float(x) for x in res.relative_pose], "winner_consistent_with_ground_truth": winner_ok,
                       "gpu_launches_per_query": (cf.kernel_launches() - l1) // max(args.queries, 1),
                       "roofline": {"bound": "hbm", "achieved": cand_per_s / world * BYTES_PER_CANDIDATE / 1e9, "peak": peak, "unit": "GB/s",
                                    "frac": cand_per_s / world * BYTES_PER_CANDIDATE / 1e9 / peak,
                                    "definition": "per-GPU candidates/s x %d algorithmic B/candidate (SURVEY 8d)" % BYTES_PER_CANDIDATE}}
                return out, (res, mine, gcx, gcy, gang, q_host, per_gpu, mode)

            nfix = args.db if args.db > 0 else (10000 if (W, H) == (640, 480) else 4000)
            total_true = 100000 if (W, H) == (640, 480) else 50000
            series = {}
            st_fixed = None
            try:
                if not args.extras:
                    raise StopIteration
                series["fixed_per_gpu"], st_fixed = run_series(nfix * world, "%d keyframes PER GPU (the same problem per GPU at every N; N = 1 is BASELINE configs[2])" % nfix)
            except StopIteration:
                pass
            except Exception as e:
                series["fixed_per_gpu"] = {"error": str(e)[:300]}
                st_fixed = None
            # ---- N-rank answer == single-rank answer (outside every timed region): rank 0 re-scans, alone and in the full store mode, a
            # 256-keyframe subsample of the GLOBAL store that contains every rank's local winner, and compares records
            check = None
            if st_fixed is not None:
                try:
                    res, mine, gcx, gcy, gang, q_host, per_gpu, mode = st_fixed
                    locals_ = [int(mine.loop_slot + rank * per_gpu) if mine.loop_slot >= 0 else -1]
                    resp_ = [[float(x) for x in mine.response]]
                    if world > 1:
                        gl = [None] * world
                        dist.all_gather_object(gl, (locals_[0], resp_[0]))
                        locals_, resp_ = [g[0] for g in gl], [g[1] for g in gl]
                    if rank == 0:
                        sub = sorted(set([g for g in locals_ if g >= 0]) | set(np.linspace(0, per_gpu * world - 1, 256 - world).astype(int).tolist()))
                        lc.clear()
                        lc.SetMode(nis.DB_FULL)
                        imgs_sub = bs.crops(canvas, gcx[sub], gcy[sub], gang[sub], H, W)
                        torch.cuda.synchronize()
                        lc.AddImages(None, np.asarray(sub, np.int32), None, ptr=imgs_sub.data_ptr(), n=len(sub), on_device=True)
                        qf = cf.ComputeIntermedium(q_host.numpy())
                        r1, recs = lc.FindLoopClosureRecords(qf, 10 ** 9, 0.0)
                        same_winner = int(r1.loop_frame_id) == int(res.loop_frame_id)
                        same_pose = bool(np.array_equal(r1.relative_pose, res.relative_pose))
                        bits = bool(np.array_equal(r1.response, res.response))
                        # the sharded scan of >= 1024 candidates takes FFT(RotateArray(query)) from the per-query cache, the 256-keyframe re-scan
                        # rotates per candidate: same peaks and poses, responses equal to f32 round-off (tests/test_gpu_parity.py, 2e-6)
                        close = bool(np.allclose(r1.response, res.response, rtol=5e-6))
                        locals_ok = all(np.allclose(recs["response"][sub.index(g)], rp, rtol=5e-6) for g, rp in zip(locals_, resp_) if g >= 0)
                        check = {"scan_equals_single_rank": bool(same_winner and same_pose and close and locals_ok), "identical_response_bits": bits,
                                 "same_winner": same_winner, "same_pose": same_pose, "responses_equal_to_5e-6": close, "local_bests_reproduced": locals_ok,
                                 "response_sharded": [float(x) for x in res.response], "response_single": [float(x) for x in r1.response],

                                 "what": "rank 0 alone re-scanned %d keyframes of the global store (every rank's local winner + an even subsample, full store "
                                         "mode) with per-candidate records: same winner, same pose, same response as the %d-rank sharded call, and every rank's "
                                         "local best record reproduced" % (len(sub), world)}
                        lc.clear()
                except Exception as e:
                    check = {"error": str(e)[:300]}
            if world > 1:
                barrier()
            if args.db < 0:
                try:
                    series["true_total"], _ = run_series(total_true, "%dk keyframes in total over %d GPU(s) (BASELINE configs[%d]), MEASURED" % (
                        total_true // 1000, world, 3 if total_true == 100000 else 4))
                except Exception as e:
                    series["true_total"] = {"error": str(e)[:300]}
            lc.clear()
            lc.SetMode(nis.DB_FULL)
            head = series.get("true_total") if "value" in (series.get("true_total") or {}) else series.get("fixed_per_gpu")
            loop = {"metric": "loop_closure_queries_per_sec", "series": series, "n_rank_check": check,
                    "collective": ("one ncclBroadcast (%d B u8 query image) + one ncclAllGather (104 B per rank) inside nis_loop_scan_sharded" % (H * W))
                    if world > 1 else "none (1 rank)", "rotated_query_cache": "on (>= 1024 candidates)"}
            if head and "value" in head:
                loop.update({k: head[k] for k in ("value", "unit", "db_keyframes", "keyframes_per_gpu", "ms_per_query", "candidates_per_sec", "store_mode", "roofline")})

        # ---- CPU baselines (rank 0, N = 1 only), bounded samples of the same workloads on the host cores: oracle/_ref (the reference's own
        # sources, kind "reference") when built, else the C port
        cpu = None
        if rank == 0 and world == 1 and args.cpu_frames > 1:
            cores = os.cpu_count() or 1
            use_ref = ref_available()
            try:
                sample = frames_host[:args.cpu_frames].numpy()
                if use_ref:
                    time_ref_stream(sample[:min(args.cpu_frames, cores + 1)], cores)                 # warm-up: per-thread contexts, FFT plans
                    v_all, dt_all = time_ref_stream(sample, cores)
                    v_1, dt_1 = time_ref_stream(sample[:7], 1)
                    v_c, _ = time_oracle_stream(sample, cores)
                    cpu = {"value": v_all, "unit": "solves/s", "cores": cores, "kind": "reference",
                           "sample": "first %d frames of the same stream (%d solves, %.1f s) through oracle/_ref = the reference's correlation_flow.cc / "
                                     "utils.cc compiled unmodified against stand-in headers (FFT = the C oracle's, warps = cv2-verified fixed point), "
                                     "%d threads; 1 thread on 7 frames: %.2f solves/s (how the reference itself runs); the dependency-free C port on "
                                     "the same frames, %d threads: %.1f solves/s" % (args.cpu_frames, args.cpu_frames - 1, dt_all, cores, v_1, cores, v_c),
                           "value_1_thread": v_1, "c_port_all_threads": v_c}
                else:
                    v_all, dt_all, label, both = best_cpu_stream(sample, cores)
                    v_1, dt_1, label1, both1 = best_cpu_stream(sample[:min(9, args.cpu_frames)], 1)
                    cpu = {"value": v_all, "unit": "solves/s", "cores": cores, "kind": "port",
                           "sample": "first %d frames of the same stream (%d solves, %.1f s), %s, %d threads; 1 thread: %.2f solves/s" % (
                               args.cpu_frames, args.cpu_frames - 1, dt_all, label, cores, v_1), "value_1_thread": v_1, "all_threads_both": both}
            except Exception as e:          # the checker must not take the bench down
                cpu = {"value": None, "unit": "solves/s", "cores": 0, "kind": "port", "sample": "failed: %r" % (e,)}
            if loop is not None:
                try:
                    ncand = 256
                    gcx, gcy, gang = bs.db_poses(ncand, seed=1)
                    db_s = bs.crops(canvas, gcx, gcy, gang, H, W).cpu().numpy()
                    q_s = bs.crops(canvas, [gcx[ncand // 2] + 13], [gcy[ncand // 2] - 7], [gang[ncand // 2] + 4.5], H, W).cpu().numpy()[0]
                    c_all, dt_s, win_s = time_cpu_scan(db_s, q_s, cores, use_ref)
                    c_1, dt_s1, _ = time_cpu_scan(db_s[:8], q_s, 1, use_ref)
                    loop["cpu_baseline"] = {"value": c_all / 1e5, "unit": "queries/s over 100k keyframes (extrapolated linearly from the sample: constant cost per candidate)",
                                            "candidates_per_sec": c_all, "candidates_per_sec_1_thread": c_1, "cores": cores,
                                            "kind": "reference" if use_ref else "port",
                                            "sample": "LoopClosure::FindLoopClosure over %d seeded keyframes (%.1f s on %d threads; 8 candidates on 1 thread: %.1f s), "
                                                      "per candidate ComputePose(..., false) = loop_closure.cc:58-59; winner = planted keyframe: %s" % (
                                                          ncand, dt_s, cores, dt_s1, win_s == ncand // 2)}
                except Exception as e:
                    loop["cpu_baseline"] = {"error": str(e)[:300]}

        if rank == 0:
            line = {"metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                    "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                    "data": "synthetic",
                    "config": workload_config(n, world), "tuning": {"batch": args.batch or "default", "lanes": args.lanes or "default", "cpus_bound_per_rank": numa},
                    "clocks": clk,
                    "e2e": {"value": e2e_value, "unit": "solves/s", "h2d_bytes_per_step": n * H * W,
                            "d2h_bytes_per_step": (n - 1) * 72},
                    "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "loop_closure": loop, "online": online, "undistort_front_end": front, "keyframe_policy": policy, "map_stitcher": stitch,
                    "pose_ok_frac": pose_ok_frac}
        cf.close()
        return line if rank == 0 else None

    line = run_size(args)
    # ---- BASELINE configs[4] beside the headline, in the same invocation: a 1280x960 stream and the 50k-keyframe store at this N
    if args.cfg4 and (W, H) == (640, 480):
        try:
            sub = argparse.Namespace(**vars(args))
            sub.frames, sub.steps, sub.warmup, sub.queries, sub.cpu_frames, sub.extras, sub.db = 250, 3, 3, 2, 0, False, (0 if args.db == 0 else -1)
            set_size("1280x960")
            l4 = run_size(sub)
            if line is not None and l4 is not None:
                keep = ("metric", "value", "unit", "ms_per_step", "config", "e2e", "gpu_launches", "pose_ok_frac")
                c4 = {k: l4[k] for k in keep}
                c4["roofline"] = {k: l4["roofline"][k] for k in ("bound", "achieved", "peak", "unit", "frac", "definition")}
                c4["loop_closure"] = l4["loop_closure"]
                line["configs4_1280x960"] = c4
        except Exception as e:
            if line is not None:
                line["configs4_1280x960"] = {"error": str(e)[:300]}
        finally:
            set_size(args.size)
    if line is not None:
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
