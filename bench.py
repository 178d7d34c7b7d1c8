#!/usr/bin/env python3
"""bench.py — Mrays/s (extend + shadow) of the render core on BASELINE.json's headline configuration.

Workload (config C5): the synthetic hierarchical-instancing scene (100 cluster sub-assemblies x
(10 instances of 4 unique 500k-triangle assets + 10 area lights) ~ 5e8 effective triangles, 1000 lights
through nested light trees, moving cluster instances with 2 transform keys) at 3840x2160, 256 spp.

A step = one full render of that image.  With N > 1 ranks (torchrun, one process per GPU) the
Hilbert-ordered tiles are dealt to the ranks by sample count (scene replicated per GPU, no exchange while
rendering) and the disjoint device tiles are summed onto rank 0 with one NCCL reduce per step;
total work is fixed, so scaling is "strong".

  value     whole-job Mrays/s, scene resident in HBM, image left on the device
  e2e       the same metric through the C ABI with host buffers: every step uploads the scene
            description (H2D) and fetches the finished image (D2H)
  roofline  k_trace_pool (the dominant kernel): algorithmic bytes of SURVEY.md 8(d)'s formula, with visit
            counts measured by the counting trace kernel, over the CUDA-event time of the trace launches;
            traffic = the kernel's measured DRAM bytes per ray (profiles/r2_trace_dram.json, keyed to the kernel
            sources' hash) x rays per launch
  cpu_baseline / --impl reference: the CPU oracle (restated reference path; the Rust reference cannot be
            built here) on the host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

WORKLOADS = {
    # name: (scene kwargs, resolution, spp, seed)
    # motion_pct=1250: the cluster instances move far enough over the shutter (drift up to 5 units, rotation up to 0.6
    # rad) that the root assembly's BVH keeps 2-key bounds on its lower nodes; round 1 ran the
    # same scene with motion_pct=100, which the builder's 1.4x union rule collapses to single-key nodes ("C5r1")
    "C5": dict(asset_quads=500, n_assets=4, n_clusters=100, inst_per_cluster=10, lights_per_cluster=10,
               cluster_xform_keys=2, res=(3840, 2160), spp=256, seed=5, motion_pct=1250),
    "C5r1": dict(asset_quads=500, n_assets=4, n_clusters=100, inst_per_cluster=10, lights_per_cluster=10,
                 cluster_xform_keys=2, res=(3840, 2160), spp=256, seed=5, motion_pct=100),
    "C4": dict(asset_quads=500, n_assets=4, n_clusters=100, inst_per_cluster=10, lights_per_cluster=10,
               cluster_xform_keys=1, res=(1920, 1080), spp=64, seed=4),
    "tiny": dict(asset_quads=40, n_assets=4, n_clusters=16, inst_per_cluster=6, lights_per_cluster=4,
                 cluster_xform_keys=2, res=(320, 180), spp=8, seed=5),
}
WORKLOAD_DESC = {
    "C5": "C5: synthetic hierarchical-instancing scene (100 clusters x (10 instances of 4 x 500k-tri assets + 10 area "
          "lights), 2 transform keys per cluster instance, motion-blurred cluster nodes), 3840x2160, 256 spp",
    "C5r1": "C5 as round 1 ran it (cluster motion small enough to collapse to single-key BVH nodes), 3840x2160, 256 spp",
    "C4": "C4: synthetic hierarchical-instancing scene, 1920x1080, 64 spp",
    "tiny": "tiny instanced scene (harness self-test), 320x180, 8 spp",
}


def peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def build_scene(name):
    from psychopath_b200 import api
    kw = dict(WORKLOADS[name])
    t0 = time.time()
    scene = api.Scene.synth_instanced(**kw)
    return scene, time.time() - t0


def algorithmic_bytes_per_ray(st, keys_node=1.0, keys_vert=1, keys_xform=1):
    """B_ray = N_node*(16+96*k_n) + N_tri*(16+48*k_v) + N_inst*(32+48*k_x) + 48 + 32."""
    rays = max(int(st["ray_count"]), 1)
    n_node = st["accel_node_visits"] / rays
    n_motion = st["motion_node_visits"] / rays
    n_tri = st["tri_tests"] / rays
    n_inst = st["instance_visits"] / rays
    b = (n_node - n_motion) * (16 + 96) + n_motion * (16 + 96 * 2) + n_tri * (16 + 48 * keys_vert) + \
        n_inst * (32 + 48 * keys_xform) + 48 + 32
    return b, {"nodes_per_ray": n_node, "motion_nodes_per_ray": n_motion, "tris_per_ray": n_tri,
               "instances_per_ray": n_inst}


def native_oracle():
    """the CPU baseline is timed with `-march=native`.  The oracle library that travels with the repo is
    built for x86-64-v3 (it must run on whatever box the tests land on); for timing, it is rebuilt here, on the box
    that runs the bench, for that box's own CPU (same sources, same -ffp-contract=off; ~6 s)."""
    if os.environ.get("PSY_ORACLE_LIB"):
        return os.environ["PSY_ORACLE_LIB"]
    odir = os.path.join(REPO, "oracle")
    out = os.path.join(odir, "_native", "liboracle.so")
    try:
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++17", "-O3", "-march=native", "-ffp-contract=off", "-fno-fast-math", "-fPIC",
                               "-pthread", "-shared", "-o", out, os.path.join(odir, "oracle.cpp")],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=300)
        os.environ["PSY_ORACLE_LIB"] = out
        return out
    except Exception:
        return None


def cpu_sample(scene, spp, seed, target_s=15.0, threads=0):
    """Times the oracle on a bounded number of Hilbert-ordered buckets of the same workload."""
    native_oracle()
    import oracle_lib as O
    # Buckets are drawn evenly from the whole Hilbert list (oracle_render_image), so the rate of a small
    # probe extrapolates; grow the sample until it takes about target_s (at most 3 rounds, bounded).
    ncores = os.cpu_count() or 1
    nb = 8 * ncores
    for _ in range(3):
        _, st, done = O.render_image(scene.desc, spp, seed, spb=4096, threads=threads, max_buckets=nb)
        if st.total_time >= 0.5 * target_s:
            break
        grow = min(max(0.9 * target_s / max(st.total_time, 1e-3), 1.5), 16.0)
        nb = int(nb * grow)
    rate = st.ray_count / max(st.total_time, 1e-9)
    return {"mrays_s": rate / 1e6, "msamples_s": st.samples / max(st.total_time, 1e-9) / 1e6, "seconds": st.total_time,
            "buckets": done, "rays": int(st.ray_count), "samples": int(st.samples), "threads": int(st.threads)}


def c3_leg(api, device):
    """BASELINE config 2 on one GPU: 10.0 M-triangle displaced mesh, 4 deformation + 4 transform keys, 1920x1080,
    128 spp.  Two timed frames after one warm-up, scene resident; visit counts from the counters build at 8 spp."""
    t0 = time.time()
    s = api.Scene.synth_displaced_grid(2240, n_vert_keys=4, n_xform_keys=4, res=(1920, 1080), spp=128, seed=3)
    build_s = time.time() - t0
    t0 = time.time()
    r = api.Renderer(s, device=device)
    upload_s = time.time() - t0
    r.render(device_resident_only=True)
    t0 = time.perf_counter()
    rays = 0
    trace_s = 0.0
    for _ in range(2):
        st = r.render(device_resident_only=True)[1]
        rays += int(st["ray_count"])
        trace_s += st["trace_time"]
    dt = time.perf_counter() - t0
    _, cst = r.render(spp=8, counters=True, device_resident_only=True)
    b_ray, per_ray = algorithmic_bytes_per_ray(cst, keys_vert=2, keys_xform=2)
    pk, _ = peaks()
    out = {"workload": "C3: 10.0 M-triangle displaced mesh, 4 deformation + 4 transform keys, 1920x1080, 128 spp",
           "value": rays / dt / 1e6, "unit": "Mrays/s", "ms_per_frame": dt / 2 * 1e3, "trace_ms_per_frame": trace_s / 2 * 1e3,
           "msamples_s": 2 * 1920 * 1080 * 128 / dt / 1e6, "per_ray": per_ray, "bytes_per_ray": b_ray,
           "roofline_frac_trace": (rays * b_ray) / max(trace_s, 1e-12) / 1e9 / pk["hbm_gbs"],
           "host_build_s": build_s, "upload_s": upload_s, "triangles": int(s.desc.contents.n_tris)}
    r.close()
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    name = args.workload
    kw = WORKLOADS[name]
    scene, build_s = build_scene(name)
    vals, last = [], None
    per_step = max(3.0, min(20.0, 150.0 / max(args.steps + args.warmup, 1)))
    for i in range(args.warmup + args.steps):
        last = cpu_sample(scene, kw["spp"], kw["seed"], target_s=per_step)
        if i >= args.warmup:
            vals.append(last)
    v = sum(x["mrays_s"] for x in vals) / len(vals)
    ms = sum(x["seconds"] for x in vals) / len(vals) * 1e3
    sample = (f"{last['buckets']} Hilbert-ordered 4x4-pixel buckets x {kw['spp']} spp of the same scene "
              f"({last['samples']} samples, {last['rays']} rays per step)")
    line = {"impl": "reference", "metric": "Mrays/s (primary+shadow)", "value": v, "unit": "Mrays/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_DESC[name], "timing": "bounded CPU sample per step"},
            "msamples_s": sum(x["msamples_s"] for x in vals) / len(vals),
            "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": last["threads"], "kind": "port", "sample": sample,
                             "build": "-O3 -march=native" if os.environ.get("PSY_ORACLE_LIB") else "-O3 -march=x86-64-v3",
                             "note": "restated reference CPU path (oracle/); the Rust reference cannot be built here"},
            "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "host_build_s": build_s}
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("PSY_BENCH_WORKLOAD", "C5"), choices=sorted(WORKLOADS))
    ap.add_argument("--pool", type=int, default=0,
                    help="path-pool size (paths in flight per GPU); 0 = the library's default: the whole image up to "
                         "64 Mi paths, at least ~6 buckets per GPU when tiled across GPUs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--static-deal", action="store_true",
                    help="N > 1: deal the buckets to the ranks by sample count instead of letting them pull (A/B)")
    ap.add_argument("--buckets-per-gpu", type=int, default=0, help="N > 1: buckets per GPU (0 = the library's 12)")
    ap.add_argument("--no-c3", action="store_true",
                    help="skip the extra leg on BASELINE config 2 (10 M-triangle mesh, 4+4 motion keys: every BVH node a "
                         "motion node)")
    ap.add_argument("--counter-spp", type=int, default=8)
    ap.add_argument("--spp", type=int, default=0, help="override spp (profiling runs only; not the headline config)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        print(f"note: warmup {args.warmup} < 3 breaks the timing rules; use only for smoke runs", file=sys.stderr)
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    g.build()
    from psychopath_b200 import api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the render core has no CPU fallback")
    torch.cuda.set_device(local_rank)
    wq_key = 0
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        # the ranks pull buckets from a shared counter; its name is a fresh random key rank 0 hands out
        k = torch.randint(1, 2 ** 31 - 1, (1,), dtype=torch.int64, device="cuda")
        dist.broadcast(k, src=0)
        wq_key = 0 if args.static_deal else int(k.item())

    name = args.workload
    kw = WORKLOADS[name]
    if args.spp:
        kw["spp"] = args.spp
        WORKLOAD_DESC[name] += f" [spp overridden to {args.spp}]"
    spp, seed = kw["spp"], kw["seed"]
    # what psy_render_image / _shard choose when max_samples_per_bucket is 0 (default_pool in psy_cuda.cu)
    total_samples = kw["res"][0] * kw["res"][1] * spp
    pool_paths = args.pool or min(64 << 20, max(2 << 20, total_samples // ((args.buckets_per_gpu or 12) * world)
                                                if world > 1 else total_samples))
    scene, build_s = build_scene(name)
    t0 = time.time()
    r = api.Renderer(scene, device=local_rank)
    upload_s = time.time() - t0
    rx, ry = scene.resolution

    dev_ptr, n_floats = r.device_image()

    class _Wrap:
        __cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (dev_ptr, False), "version": 2}

    image_t = torch.as_tensor(_Wrap(), device=torch.device("cuda", local_rank))
    lat = torch.zeros(1, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        st = r.render_shard(rank, world, max_samples_per_bucket=args.pool, work_queue_key=wq_key,
                            buckets_per_gpu=args.buckets_per_gpu) if world > 1 else \
            r.render(max_samples_per_bucket=args.pool, device_resident_only=True)[1]
        if world > 1:
            dist.reduce(image_t, dst=0, op=dist.ReduceOp.SUM)  # tiles are disjoint, the rest is zero
        return st

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rays = 0
    launches = 0
    trace_ms = 0.0
    trace_launches = 0
    stage = {"generate_ms": 0.0, "trace_ms": 0.0, "shade_ms": 0.0, "film_ms": 0.0}
    barrier()
    e0.record()
    tw0 = time.perf_counter()
    for _ in range(args.steps):
        st = step()
        rays += int(st["ray_count"])
        launches += int(st["kernel_launches"])
        ms, nl = r.last_trace_timing()
        trace_ms += ms
        trace_launches += nl
        stage["generate_ms"] += st["initial_ray_generation_time"] * 1e3 / args.steps
        stage["trace_ms"] += st["trace_time"] * 1e3 / args.steps
        stage["shade_ms"] += st["ray_generation_time"] * 1e3 / args.steps
        stage["film_ms"] += st["sample_writing_time"] * 1e3 / args.steps
    e1.record()
    barrier()
    wall_s = time.perf_counter() - tw0
    # The library renders on its own stream and synchronises it before returning, so the torch events on
    # the current stream bracket the same region as the wall clock; use the larger of the two.
    dev_s = max(e0.elapsed_time(e1) * 1e-3, wall_s)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([dev_s, float(rays), float(launches), trace_ms, float(trace_launches)], dtype=torch.float64,
                     device="cuda")
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dev_s = float(tmax[0])
        rays = int(tsum[1])
        launches = int(tsum[2])
    samples_per_step = rx * ry * spp
    value = rays / dev_s / 1e6
    msamples = samples_per_step * args.steps / dev_s / 1e6

    # ---- end-to-end through the C ABI with host buffers: every step each rank uploads the scene description
    # (H2D), renders its tiles, the tiles are gathered, and rank 0 reads the finished image back (D2H) ----
    # host buffers of the end-to-end arm are page-locked, as the timing contract asks: the scene arrays (registered once)
    # and the image the result is read into (reused from step to step)
    scene.pin()
    e2e_img = torch.empty((ry, rx, 3), dtype=torch.float32, pin_memory=True).numpy() if rank == 0 else None

    e2e_phases = []

    def e2e_step():
        tp = [time.perf_counter()]
        rr = api.Renderer(scene, device=local_rank)          # H2D: the whole scene description
        tp.append(time.perf_counter())
        if world == 1:
            img_, st_ = rr.render(max_samples_per_bucket=args.pool, out=e2e_img)   # D2H: the finished image
            tp.append(time.perf_counter())
        else:
            st_ = rr.render_shard(rank, world, max_samples_per_bucket=args.pool, work_queue_key=wq_key,
                                  buckets_per_gpu=args.buckets_per_gpu)
            tp.append(time.perf_counter())
            ptr_, n_ = rr.device_image()

            class _W:
                __cuda_array_interface__ = {"shape": (n_,), "typestr": "<f4", "data": (ptr_, False), "version": 2}
            t_ = torch.as_tensor(_W(), device=torch.device("cuda", local_rank))
            dist.reduce(t_, dst=0, op=dist.ReduceOp.SUM)
            torch.cuda.synchronize()
            tp.append(time.perf_counter())
            img_ = rr.fetch_image(out=e2e_img) if rank == 0 else None          # D2H on rank 0
            tp.append(time.perf_counter())
        rr.close()
        tp.append(time.perf_counter())
        e2e_phases.append([round((b - a) * 1e3, 2) for a, b in zip(tp, tp[1:])])
        return img_, st_
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    e_rays = 0
    n_e2e = max(1, min(args.steps, 2))
    for _ in range(n_e2e):
        img, st2 = e2e_step()
        e_rays += int(st2["ray_count"])
    barrier()
    e_s = time.perf_counter() - t0
    if os.environ.get("PSY_E2E_PHASES"):  # upload | render | (reduce | fetch |) close, ms per step, every rank
        print(f"[e2e phases] rank {rank}: {e2e_phases}", file=sys.stderr, flush=True)
    et = torch.tensor([e_s, float(e_rays)], dtype=torch.float64, device="cuda")
    if world > 1:
        emax = et.clone()
        dist.all_reduce(emax, op=dist.ReduceOp.MAX)
        esum = et.clone()
        dist.all_reduce(esum, op=dist.ReduceOp.SUM)
        e_s, e_rays = float(emax[0]), int(esum[1])
    e2e = None
    if rank == 0:
        e2e = {"value": e_rays / e_s / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": int(scene.host_bytes()) * world,
               "d2h_bytes_per_step": int(rx * ry * 12), "steps": n_e2e, "ms_per_step": e_s / n_e2e * 1e3,
               "image_mean_xyz": [float(x) for x in np.nanmean(np.where(np.isfinite(img), img, np.nan).reshape(-1, 3), axis=0)],
               "non_finite_pixels": int((~np.isfinite(img).all(axis=-1)).sum())}

    # ---- roofline of k_trace: visit counts from the counters build at reduced spp (per-ray averages) ----
    roofline = None
    cpu = None
    if rank == 0:
        pk, pk_kind = peaks()
        _, cst = r.render(max_samples_per_bucket=args.pool, spp=min(args.counter_spp, spp), counters=True,
                          device_resident_only=True)
        # every instance path of these scenes carries the cluster's transform keys (32 + 48*k_x per entry)
        b_ray, per_ray = algorithmic_bytes_per_ray(cst, keys_xform=int(kw.get("cluster_xform_keys", 1)))
        my_rays = int(t[1])
        achieved = (my_rays * b_ray) / max(float(t[3]) * 1e-3, 1e-12) / 1e9
        # DRAM bytes of the trace kernel per launch: bytes per ray from the committed ncu capture of the same
        # scene (profiles/r1_trace_dram.json) x the rays one launch of this run traces on average
        # (ncu cannot run inside this process: the capture is keyed to a hash of the kernel's sources and is reported
        # only while it still describes the kernel that ran; tools/trace_traffic.py refreshes it)
        traffic, traffic_note = None, None
        tp = os.path.join(REPO, "profiles", "r2_trace_dram.json")
        if os.path.exists(tp) and name == "C5":
            sys.path.insert(0, os.path.join(REPO, "tools"))
            from trace_traffic import kernel_source_hash
            td = json.load(open(tp))
            if td.get("kernel_source_sha256") == kernel_source_hash():
                traffic = td["dram_bytes_per_ray"] * my_rays / max(float(t[4]), 1.0)
                traffic_note = "ncu dram__bytes_read+write per ray (%.0f B, profiles/r2_trace_dram.json, same kernel sources) x " \
                               "rays per launch; algorithmic bytes per launch: %.3g" % (
                                   td["dram_bytes_per_ray"], b_ray * my_rays / max(float(t[4]), 1.0))
            else:
                traffic_note = "profiles/r2_trace_dram.json was captured for other kernel sources: not reported" 
        roofline = {"bound": "hbm", "kernel": "k_trace_pool", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "frac": achieved / pk["hbm_gbs"], "traffic": traffic, "traffic_note": traffic_note,
                    "peak_kind": pk_kind,
                    "bytes_per_ray": b_ray, "per_ray": per_ray,
                    "trace_ms_per_step": float(t[3]) / args.steps, "trace_launches_per_step": float(t[4]) / args.steps,
                    "trace_share_of_step": float(t[3]) * 1e-3 / dev_s,
                    "roofline_mrays_s": pk["hbm_gbs"] * 1e9 / b_ray / 1e6,
                    "note": "visit counts measured by the counters build at %d spp" % min(args.counter_spp, spp)}
        if not args.no_cpu_baseline:
            c = cpu_sample(scene, spp, seed, target_s=15.0)
            cpu = {"value": c["mrays_s"], "unit": "Mrays/s", "cores": c["threads"], "kind": "port",
                   "build": "-O3 -march=native" if os.environ.get("PSY_ORACLE_LIB") else "-O3 -march=x86-64-v3",
                   "msamples_s": c["msamples_s"], "seconds": c["seconds"],
                   "sample": f"{c['buckets']} Hilbert-ordered 4x4-pixel buckets x {spp} spp of the same scene "
                             f"({c['samples']} samples, {c['rays']} rays)"}

    # ---- extra leg (rank 0, one GPU): BASELINE config 2, where EVERY node visit lerps two time keys ----
    c3 = None
    if rank == 0 and world == 1 and name == "C5" and not args.no_c3 and not args.spp:
        c3 = c3_leg(api, local_rank)

    if rank == 0:
        line = {"metric": "Mrays/s (primary+shadow)", "value": value, "unit": "Mrays/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_s / args.steps * 1e3,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD_DESC[name], "resolution": [rx, ry], "spp": spp, "seed": seed,
                           "paths_in_flight": pool_paths, "parallelism": f"tiles x{world}",
                           "tile_assignment": ("n/a" if world == 1 else "static deal by sample count" if not wq_key else
                                               "ranks pull buckets from a shared-memory counter, largest first"),
                           "l2": "working set (path pool %d MB + scene) exceeds the 126 MB L2" % (pool_paths * 390 >> 20)},
                "msamples_s": msamples, "rays_per_step": rays / args.steps, "samples_per_step": samples_per_step,
                "gpu_launches": launches, "clocks": clocks, "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu,
                "stage_ms_per_step_rank0": stage, "host_build_s": build_s, "upload_s": upload_s,
                "motion_nodes_per_ray": roofline["per_ray"]["motion_nodes_per_ray"] if roofline else None,
                "c3": c3}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        if rank == 0 and wq_key:
            # remove the shared-memory work counter's name (the library unlinks it when a handle is closed)
            import ctypes
            lib = api.cuda_lib()
            h = ctypes.c_void_p()
            if lib.psy_work_queue_open(wq_key, ctypes.byref(h)) == 0:
                lib.psy_work_queue_close(h)
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
