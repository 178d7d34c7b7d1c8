"""Debug helper: renders a bench workload on the GPU, locates non-finite pixels and compares those
pixels' per-sample radiance between the CUDA path and the CPU oracle."""
import sys, os
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import numpy as np
import bench
from psychopath_b200 import api
import oracle_lib as O

name = sys.argv[1] if len(sys.argv) > 1 else "C4"
kw = dict(bench.WORKLOADS[name])
if len(sys.argv) > 2:
    kw["spp"] = int(sys.argv[2])
scene = api.Scene.synth_instanced(**kw)
r = api.Renderer(scene)
img, st = r.render()
bad = ~np.isfinite(img).all(axis=-1)
ys, xs = np.where(bad)
print("non-finite pixels:", len(xs), "of", bad.size, "rays", st["ray_count"])
for x, y in list(zip(xs, ys))[:6]:
    gs, _ = r.render_bucket_samples(int(x), int(y), 1, 1, kw["spp"], kw["seed"])
    _, osamp, _ = O.render_bucket(scene.desc, int(x), int(y), 1, 1, kw["spp"], kw["seed"], want_samples=True)
    gi = np.where(~np.isfinite(gs.reshape(-1, 4)).all(axis=-1))[0]
    oi = np.where(~np.isfinite(osamp.reshape(-1, 4)).all(axis=-1))[0]
    print("pixel", x, y, "gpu bad samples", gi.tolist(), "oracle bad samples", oi.tolist())
    for i in gi[:2]:
        print("   gpu", gs.reshape(-1, 4)[i], "oracle", osamp.reshape(-1, 4)[i])
