import os, sys, time, json
sys.path.insert(0, '.')
from psychopath_b200 import api
api.enable_gpu_build(min_tris=1 << 16, mode="bvh4")
t0 = time.time()
s = api.Scene.synth_displaced_grid(2240, n_vert_keys=4, n_xform_keys=4, res=(1920, 1080), spp=32, seed=3)
gsec, _ = api.gpu_build_seconds()
api.disable_gpu_build()
r = api.Renderer(s)
r.render(device_resident_only=True)
t0 = time.perf_counter()
_, st = r.render(device_resident_only=True)
dt = time.perf_counter() - t0
_, cst = r.render(spp=4, counters=True, device_resident_only=True)
print(os.environ.get("PSY_CUDA_LIB", "default")[-14:], "gpu builder s %.3f" % gsec, "nodes", s.desc.contents.n_nodes, "Mrays/s %.1f" % (st["ray_count"] / dt / 1e6),
      "nodes/ray %.2f" % (cst["accel_node_visits"] / cst["ray_count"]), "tris/ray %.2f" % (cst["tri_tests"] / cst["ray_count"]), "depth", s.desc.contents.meshes[0].bvh_depth)
