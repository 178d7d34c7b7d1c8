"""world_size-2 CPU test (gloo) of the N>1 path: tile dealing + the end-of-frame gather.  The ranks
render their tiles with the CPU oracle (no GPU here); rank 0 must end up with exactly the single-process
image."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path):
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_lib as O
    from psychopath_b200 import api, sharding
    from scenes import motion_instances_text
    scene = api.Scene.from_psy_string(motion_instances_text())
    rx, ry = scene.resolution
    spp, seed = 2, 7
    bw = sharding.bucket_dim(256, spp)
    buckets = sharding.bucket_list(rx, ry, bw)
    mine = sharding.buckets_for_rank(buckets, rank, world)
    img = np.zeros((ry, rx, 3), dtype=np.float32)
    for (x, y, w, h) in mine:
        xyz, _, _ = O.render_bucket(scene.desc, x, y, w, h, spp, seed)
        img[y:y + h, x:x + w] = xyz
    t = torch.from_numpy(img)
    sharding.gather_image(t, dst=0)
    counts = torch.tensor([len(mine)], dtype=torch.int64)
    dist.all_reduce(counts)
    if rank == 0:
        np.save(out_path, np.concatenate([t.numpy().reshape(-1), [float(counts.item()), float(len(buckets))]]))
    dist.barrier()
    dist.destroy_process_group()


def test_bucket_assignment_is_a_balanced_partition():
    """Every bucket goes to exactly one rank, each rank keeps Hilbert order, and no rank carries more than one
    bucket's worth of pixels above the mean (the longest-processing-time-first bound)."""
    sys.path.insert(0, REPO)
    from psychopath_b200 import sharding
    for (rx, ry, bw) in ((3840, 2160, 412), (1920, 1080, 300), (96, 64, 11)):
        buckets = sharding.bucket_list(rx, ry, bw)
        assert sum(w * h for (_, _, w, h) in buckets) == rx * ry
        for world in (1, 2, 3, 8):
            shares = [sharding.buckets_for_rank(buckets, r, world) for r in range(world)]
            assert sorted(b for s in shares for b in s) == sorted(buckets)
            for s in shares:
                idx = [buckets.index(b) for b in s]
                assert idx == sorted(idx)
            loads = [sum(w * h for (_, _, w, h) in s) for s in shares]
            biggest = max(w * h for (_, _, w, h) in buckets)
            assert max(loads) <= rx * ry / world + biggest


def test_two_rank_tile_gather(tmp_path):
    out = str(tmp_path / "img.npy")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    data = np.load(out)
    img, n_done, n_total = data[:-2], data[-2], data[-1]
    assert n_done == n_total  # every bucket rendered exactly once across the ranks
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import oracle_lib as O
    from psychopath_b200 import api
    from scenes import motion_instances_text
    scene = api.Scene.from_psy_string(motion_instances_text())
    ref, _, _ = O.render_image(scene.desc, 2, 7, spb=256, threads=1)
    assert np.array_equal(img.reshape(ref.shape).astype(np.float32), ref)


def _puller(key, frames, n_jobs, q, barrier):
    sys.path.insert(0, REPO)
    import ctypes as C
    from psychopath_b200 import abi
    lib = abi.load_cuda()
    h = C.c_void_p()
    assert lib.psy_work_queue_open(key, C.byref(h)) == 0
    got = {}
    for f in frames:
        barrier.wait()  # the per-frame collective of the real thing (the tile gather) keeps the ranks in step
        mine = []
        while True:
            idx = C.c_uint64()
            assert lib.psy_work_queue_pull(h, f, C.byref(idx)) == 0
            if idx.value >= n_jobs:
                break
            mine.append(idx.value)
        got[f] = mine
    q.put(got)
    barrier.wait()
    lib.psy_work_queue_close(h)


def test_shared_work_counter_hands_every_bucket_to_exactly_one_process():
    """The pull scheduler of the N > 1 path (psy_render_image_shard with work_queue_key): processes that open the same
    key draw 0, 1, 2, ... from one shared-memory counter per frame — no bucket twice, none lost, over 70 frames
    (more than the 64 counter slots, so slot reuse is covered).  The library is loaded without a GPU."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    key = (os.getpid() * 2654435761) & 0x7FFFFFFF or 1
    n_jobs, frames = 37, list(range(70))
    q = ctx.Queue()
    barrier = ctx.Barrier(3)
    procs = [ctx.Process(target=_puller, args=(key, frames, n_jobs, q, barrier)) for _ in range(3)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for f in frames:
        drawn = sorted(i for r in results for i in r[f])
        assert drawn == list(range(n_jobs)), f
