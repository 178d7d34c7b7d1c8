import os, sys, time
sys.path.insert(0, '/root/repo')
import bench, numpy as np
from psychopath_b200 import api
kw = dict(bench.WORKLOADS["C5"]); kw["spp"] = int(sys.argv[1]) if len(sys.argv) > 1 else 64
scene = api.Scene.synth_instanced(**kw)
r = api.Renderer(scene)
for i in range(2):
    t0=time.perf_counter(); st = r.render(device_resident_only=True)[1]; t1=time.perf_counter()
    print(f"resident render {1e3*(t1-t0):.1f} ms  device stages {1e3*(st['trace_time']+st['ray_generation_time']+st['initial_ray_generation_time']+st['sample_writing_time']):.1f}")
def loop(tag, out=None):
    for i in range(3):
        t0 = time.perf_counter(); rr = api.Renderer(scene); t1 = time.perf_counter()
        img, st = rr.render(out=out); t2 = time.perf_counter()
        rr.close(); t3 = time.perf_counter()
        print(f"{tag}: create {1e3*(t1-t0):.1f}  render+fetch {1e3*(t2-t1):.1f} (device stages {1e3*(st['trace_time']+st['ray_generation_time']+st['initial_ray_generation_time']+st['sample_writing_time']):.1f}, total_time {1e3*st['total_time']:.1f})  destroy {1e3*(t3-t2):.1f}")
loop("pageable")
scene.pin()
import torch
out = torch.empty((2160, 3840, 3), dtype=torch.float32, pin_memory=True).numpy()
loop("pinned", out)
