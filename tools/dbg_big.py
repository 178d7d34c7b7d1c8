"""Sanity run of BASELINE config 2 at full size: a ~10M-triangle displaced mesh with 4 deformation and 4 transform
keys.  Checks that the scene uploads, renders, is deterministic, and that probe rays through the production kernel
agree with the one-ray-per-lane kernel is left to the tests; prints build / upload / render times."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from psychopath_b200 import api
q = int(sys.argv[1]) if len(sys.argv) > 1 else 2240
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 8
t0 = time.time()
s = api.Scene.synth_displaced_grid(q, n_vert_keys=4, n_xform_keys=4, res=(1920, 1080), spp=spp, seed=3)
d = s.desc.contents
t1 = time.time()
r = api.Renderer(s)
t2 = time.time()
print(f"triangles {2 * q * q / 1e6:.1f} M, nodes {d.n_nodes}, host build {t1 - t0:.1f}s, upload {t2 - t1:.1f}s", flush=True)
img, st = r.render()
t3 = time.time()
img2, st2 = r.render(max_samples_per_bucket=1 << 20)
print(f"render {t3 - t2:.2f}s  rays {st['ray_count']}  Mrays/s {st['ray_count'] / (t3 - t2) / 1e6:.1f}  trace_time {st['trace_time']:.2f}s")
print("deterministic across pool sizes:", np.array_equal(img, img2, equal_nan=True), " mean", float(np.nanmean(img)))
