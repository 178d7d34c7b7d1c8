"""A wider sweep of the fuzz parity check than the test suite runs: seeds A..B of tests/scenes.py::random_scene_text
(and its many_keys variant), strict build against the oracle, every sample; prints the seeds that differ."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
import numpy as np
import oracle_lib as O
from psychopath_b200 import api
from scenes import random_scene_text

a, b = int(sys.argv[1]), int(sys.argv[2])
bad = []
rays = 0
for seed in range(a, b):
    for many in (False, True):
        scene = api.Scene.from_psy_string(random_scene_text(seed, many_keys=many))
        d = scene.desc.contents
        r = api.Renderer(scene, strict=True)
        gs, st = r.render_bucket_samples(0, 0, d.res_x, d.res_y, d.spp, d.seed)
        r.close()
        with O.strict_math():
            _, osamp, ost = O.render_bucket(scene.desc, 0, 0, d.res_x, d.res_y, d.spp, d.seed, want_samples=True)
        au, bu = gs.view(np.uint32), osamp.view(np.uint32)
        ok = ((au == bu) | (np.isnan(gs) & np.isnan(osamp))).all() and st["ray_count"] == ost.ray_count
        rays += int(ost.ray_count)
        if not ok:
            bad.append((seed, many))
            print("DIFFERS", seed, many, flush=True)
print(f"seeds {a}..{b - 1} x 2 variants: {len(bad)} differ; {rays} rays")
