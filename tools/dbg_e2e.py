"""Debug helper: where the end-to-end overhead of one frame goes (scene upload, first render, image download, teardown)."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from psychopath_b200 import api
scene, _ = bench.build_scene("C5")
scene.set_render(res=(3840, 2160), spp=int(sys.argv[1]) if len(sys.argv) > 1 else 4, seed=5)
for it in range(3):
    t0 = time.perf_counter()
    r = api.Renderer(scene, device=0)
    t1 = time.perf_counter()
    img, st = r.render(max_samples_per_bucket=4 << 20)
    t2 = time.perf_counter()
    _, st2 = r.render(max_samples_per_bucket=4 << 20, device_resident_only=True)
    t3 = time.perf_counter()
    r.close()
    t4 = time.perf_counter()
    print(f"create {t1 - t0:.3f}s  first render+download {t2 - t1:.3f}s  resident render {t3 - t2:.3f}s  close {t4 - t3:.3f}s")
