"""Times BASELINE.json's other configurations once each on one GPU (not the bench line; for BASELINE.md section 6).

    python tools/run_configs.py            # all
    python tools/run_configs.py 2 3        # only configs 2 and 3
"""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from psychopath_b200 import api
from scenes import cornell_text, cube_text

only = set(sys.argv[1:])


def want(i):
    return not only or str(i) in only


def run(name, scene):
    r = api.Renderer(scene)
    r.render(device_resident_only=True)  # warm-up
    t0 = time.perf_counter()
    _, st = r.render(device_resident_only=True)
    dt = time.perf_counter() - t0
    print(f"{name}: {st['ray_count'] / dt / 1e6:.1f} Mrays/s, {st['samples'] / dt / 1e6:.1f} Msamples/s, {dt * 1e3:.1f} ms/frame, "
          f"trace {st['trace_time'] * 1e3:.1f} ms, rays {st['ray_count']}", flush=True)


if want(0):
    s = api.Scene.from_psy_string(cornell_text())
    s.set_render(res=(512, 512), spp=16, seed=1)
    run("config 0: cornell_box.psy 512x512 16 spp", s)
if want(1):
    s = api.Scene.from_psy_string(cube_text(res=(1920, 1080), spp=64, camera_motion=True))
    run("config 1: cube.psy 1080p 64 spp, DoF + camera motion", s)
if want(2):
    t0 = time.time()
    s = api.Scene.synth_displaced_grid(2240, n_vert_keys=4, n_xform_keys=4, res=(1920, 1080), spp=128, seed=3)
    print(f"  (10.0 M triangles, host BVH build {time.time() - t0:.1f} s)")
    run("config 2: 10M-triangle displaced mesh 1080p 128 spp, 4+4 motion keys", s)
if want(3):
    s = api.Scene.synth_instanced(asset_quads=500, n_assets=4, n_clusters=100, inst_per_cluster=10, lights_per_cluster=10,
                                  cluster_xform_keys=1, res=(1920, 1080), spp=64, seed=4)
    run("config 3: instanced scene (1000 instances of 500k-tri assets, 1000 lights) 1080p 64 spp", s)
