"""host SAH build vs GPU LBVH topology (host bounds + collapse) vs the whole BVH4 on the device, for
BASELINE config 2 (10 M triangles, 4 + 4 keys): build seconds, then render speed on each tree (the LBVH tree is of
lower quality: more node visits per ray; the two LBVH modes give the same tree)."""
import os, sys, time, json
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from psychopath_b200 import api
out = {}
for mode in ("sah", "lbvh", "lbvh_bvh4"):
    if mode != "sah":
        api.enable_gpu_build(min_tris=1 << 16, mode="bvh4" if mode == "lbvh_bvh4" else "topology")
        api.gpu_build_seconds()
    t0 = time.time()
    s = api.Scene.synth_displaced_grid(2240, n_vert_keys=4, n_xform_keys=4, res=(1920, 1080), spp=32, seed=3)
    build_s = time.time() - t0
    gsec, _ = api.gpu_build_seconds() if mode != "sah" else (0.0, 0)
    api.disable_gpu_build()
    r = api.Renderer(s)
    r.render(device_resident_only=True)
    t0 = time.perf_counter()
    _, st = r.render(device_resident_only=True)
    dt = time.perf_counter() - t0
    _, cst = r.render(spp=4, counters=True, device_resident_only=True)
    out[mode] = {"scene_build_s": build_s, "of_which_gpu_builder_s": gsec, "library_build_seconds": s.build_seconds,
                 "nodes": int(s.desc.contents.n_nodes), "mrays_s": st["ray_count"] / dt / 1e6, "ms_per_frame_32spp": dt * 1e3,
                 "nodes_per_ray": cst["accel_node_visits"] / cst["ray_count"], "tris_per_ray": cst["tri_tests"] / cst["ray_count"]}
    print(mode, json.dumps(out[mode]), flush=True)
    r.close()
json.dump(out, open(os.path.join(REPO, "gpurun_out", "r2_gpu_build.json"), "w"), indent=1)
