"""Where the end-to-end overhead per frame goes: scene create (validate + H2D + device-side gathers), render, image
fetch (D2H), destroy.  C5 scene at a small spp so that the fixed costs stand out."""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench
from psychopath_b200 import api
kw = dict(bench.WORKLOADS["C5"]); kw["spp"] = 1
scene = api.Scene.synth_instanced(**kw)
r0 = api.Renderer(scene); r0.render(); r0.close()
for i in range(4):
    t0 = time.perf_counter(); r = api.Renderer(scene); t1 = time.perf_counter()
    img, st = r.render(); t2 = time.perf_counter()
    r.close(); t3 = time.perf_counter()
    print(f"create {1e3*(t1-t0):.1f} ms  render+fetch {1e3*(t2-t1):.1f} ms (device {1e3*(st['trace_time']+st['ray_generation_time']+st['initial_ray_generation_time']+st['sample_writing_time']):.1f})  destroy {1e3*(t3-t2):.1f} ms")
print("--- pinned scene arrays and pinned result image")
scene.pin()
r = api.Renderer(scene); out = r.pinned_image(); r.close()
import numpy as np
for i in range(4):
    t0 = time.perf_counter(); r = api.Renderer(scene); t1 = time.perf_counter()
    img, st = r.render(out=out); t2 = time.perf_counter()
    r.close(); t3 = time.perf_counter()
    print(f"create {1e3*(t1-t0):.1f} ms  render+fetch {1e3*(t2-t1):.1f} ms (device {1e3*(st['trace_time']+st['ray_generation_time']+st['initial_ray_generation_time']+st['sample_writing_time']):.1f})  destroy {1e3*(t3-t2):.1f} ms")
