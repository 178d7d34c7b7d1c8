"""Locates the non-finite pixels of a bench workload's frame (strict build, so that the oracle's strict mode can
be asked about the same samples) and compares those pixels' per-sample radiance with the CPU oracle.

    python tools/find_nan.py C5r1 256
"""
import os
import sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
import numpy as np
import bench
from psychopath_b200 import api
import oracle_lib as O

name = sys.argv[1] if len(sys.argv) > 1 else "C5"
kw = dict(bench.WORKLOADS[name])
if len(sys.argv) > 2:
    kw["spp"] = int(sys.argv[2])
scene = api.Scene.synth_instanced(**kw)
for strict in (True, False):
    r = api.Renderer(scene, strict=strict)
    img, st = r.render()
    bad = ~np.isfinite(img).all(axis=-1)
    ys, xs = np.where(bad)
    print("strict" if strict else "product", "build: non-finite pixels:", len(xs), "of", bad.size, "rays", st["ray_count"], flush=True)
    for x, y in list(zip(xs, ys))[:8]:
        gs, _ = r.render_bucket_samples(int(x), int(y), 1, 1, kw["spp"], kw["seed"])
        if strict:
            with O.strict_math():
                _, osamp, _ = O.render_bucket(scene.desc, int(x), int(y), 1, 1, kw["spp"], kw["seed"], want_samples=True)
        else:
            _, osamp, _ = O.render_bucket(scene.desc, int(x), int(y), 1, 1, kw["spp"], kw["seed"], want_samples=True)
        gi = np.where(~np.isfinite(gs.reshape(-1, 4)).all(axis=-1))[0]
        oi = np.where(~np.isfinite(osamp.reshape(-1, 4)).all(axis=-1))[0]
        print("  pixel", x, y, "gpu bad samples", gi.tolist(), "oracle bad samples", oi.tolist(),
              "all samples bit-equal" if np.array_equal(gs.view(np.uint32), osamp.view(np.uint32)) or
              ((gs.view(np.uint32) == osamp.view(np.uint32)) | (np.isnan(gs) & np.isnan(osamp))).all() else "SAMPLES DIFFER", flush=True)
    r.close()
