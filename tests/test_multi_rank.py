"""CPU test of the N > 1 scan path (gloo, world_size 2): every rank holds the best record of its keyframe shard, the
ranks all-gather the 96-byte records and merge them with nis_loop_reduce -- the same plumbing bench.py runs over NCCL.
The merge must reproduce the reference's rule over the GLOBAL iteration order: strictly greater response.sum() wins,
first in order wins ties (src/loop_closure.cc:61), thresholds applied to the winner (:68-71)."""
import ctypes as C
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import ni_slam_b200 as nis


def make_record(slot, frame_id, response, pose=(1.0, 2.0, 0.5), evaluated=4):
    r = nis.LoopResultC()
    r.found = 0
    r.slot, r.frame_id, r.hyp = slot, frame_id, 0
    for i in range(3):
        r.response[i] = response[i]
        r.relative_pose[i] = pose[i]
    r.evaluated = evaluated
    return r


CASES = [
    # (rank0 record, rank1 record, expected winner rank, expected found)
    (dict(slot=3, frame_id=3, response=(150.0, 150.0, 110.0)), dict(slot=1, frame_id=1001, response=(160.0, 160.0, 100.0)), 1, True),
    # exact tie on response.sum(): the record earlier in the global order (rank 0's frame 3) wins
    (dict(slot=3, frame_id=3, response=(150.0, 150.0, 110.0)), dict(slot=1, frame_id=1001, response=(150.0, 150.0, 110.0)), 0, True),
    # rank 1 evaluated nothing (all candidates filtered): its (-1,-1,-1) record never wins
    (dict(slot=2, frame_id=2, response=(70.0, 70.0, 65.0)), dict(slot=-1, frame_id=-1, response=(-1.0, -1.0, -1.0), evaluated=0), 0, True),
    # winner below the thresholds -> not found
    (dict(slot=2, frame_id=2, response=(20.0, 20.0, 65.0)), dict(slot=5, frame_id=1005, response=(10.0, 10.0, 30.0)), 0, False),
    # nobody evaluated anything
    (dict(slot=-1, frame_id=-1, response=(-1.0, -1.0, -1.0), evaluated=0), dict(slot=-1, frame_id=-1, response=(-1.0, -1.0, -1.0), evaluated=0), -1, False),
]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = nis.LoopClosureConfig(position_response_thr=60, angle_response_thr=60)
    nbytes = C.sizeof(nis.LoopResultC)
    out = []
    for case in CASES:
        mine = make_record(**case[rank])
        t = torch.frombuffer(bytearray(bytes(mine)), dtype=torch.uint8)
        gathered = [torch.empty(nbytes, dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(gathered, t)
        recs = [nis.LoopResultC.from_buffer_copy(g.numpy().tobytes()) for g in gathered]
        order = [r.frame_id if r.slot >= 0 else 2 ** 62 for r in recs]
        res, win = nis.loop_reduce(recs, order, cfg)
        out.append((win, res.found, res.loop_frame_id, res.evaluated, tuple(res.response)))
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_scan_merge_gloo():
    nis.load_library()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results[0] == results[1]                    # every rank reaches the same decision
    for (c0, c1, want_rank, want_found), (win, found, fid, evaluated, resp) in zip(CASES, results[0]):
        assert win == want_rank and found == want_found
        if want_rank >= 0:
            assert fid == (c0, c1)[want_rank]["frame_id"]
        else:
            assert resp == (-1.0, -1.0, -1.0)
        assert evaluated == c0.get("evaluated", 4) + c1.get("evaluated", 4)


def test_reduce_without_order_uses_rank_index():
    cfg = nis.LoopClosureConfig(60, 60)
    a = make_record(1, 11, (100.0, 100.0, 100.0))
    b = make_record(2, 12, (100.0, 100.0, 100.0))
    res, win = nis.loop_reduce([a, b], None, cfg)
    assert win == 0 and res.loop_frame_id == 11 and res.found
