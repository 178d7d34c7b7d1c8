"""Measured parity numbers of the PRODUCT build (libdevice libm) against the oracle (glibc libm), per scene:
image relative-L2 and RMSE in linear XYZ, share of samples beyond 1e-3, ray-count difference, non-finite pixels.
Needs a GPU.  Writes one JSON document (default gpurun_out/r2_parity.json); the copy under profiles/ is what the
gates in tests/test_gpu_parity.py are derived from (gate = 2 x measured, rounded up)."""
import json
import os
import sys
import time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
import numpy as np
import oracle_lib as O
from psychopath_b200 import api
from scenes import cornell_text, cube_text
from test_gpu_parity import SCENES, make_scene

out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, "gpurun_out", "r2_parity.json")
rows = []


def image_row(name, scene, spp, seed, crop=None, spb=4096):
    r = api.Renderer(scene)
    t0 = time.time()
    g, gst = r.render(spp=spp, seed=seed, crop=crop)
    tg = time.time() - t0
    t0 = time.time()
    o, ost, _ = O.render_image(scene.desc, spp, seed, spb=spb, crop=crop)
    to = time.time() - t0
    fin = np.isfinite(g).all(axis=-1) & np.isfinite(o).all(axis=-1)
    diff = (g - o)[fin]
    row = {"scene": name, "resolution": list(scene.resolution), "spp": spp, "crop": crop,
           "rel_l2": float(np.linalg.norm(diff) / max(np.linalg.norm(o[fin]), 1e-30)),
           "rmse": float(np.sqrt((diff.astype(np.float64) ** 2).mean())),
           "max_abs": float(np.abs(diff).max()), "mean_xyz": [float(v) for v in o[fin].mean(axis=0)],
           "non_finite_pixels_gpu": int((~np.isfinite(g).all(axis=-1)).sum()),
           "non_finite_pixels_oracle": int((~np.isfinite(o).all(axis=-1)).sum()),
           "rays_gpu": int(gst["ray_count"]), "rays_oracle": int(ost.ray_count), "gpu_s": tg, "oracle_s": to}
    rows.append(row)
    print(json.dumps(row), flush=True)


def sample_row(name, scene):
    d = scene.desc.contents
    w, h = min(d.res_x, 40), min(d.res_y, 24)
    x0, y0 = (d.res_x - w) // 2, (d.res_y - h) // 2
    r = api.Renderer(scene)
    gs, st = r.render_bucket_samples(x0, y0, w, h, 4, d.seed)
    _, osamp, ost = O.render_bucket(scene.desc, x0, y0, w, h, 4, d.seed, want_samples=True)
    err = np.abs(gs - osamp).max(axis=-1) / (np.abs(osamp).max(axis=-1) + 1e-3)
    row = {"scene": name, "kind": "per-sample radiance, 4 spp", "samples": int(err.size),
           "bit_identical_share": float((gs.view(np.uint32) == osamp.view(np.uint32)).all(axis=-1).mean()),
           "share_beyond_1e-3": float((err > 1e-3).mean()), "share_beyond_1e-5": float((err > 1e-5).mean()),
           "rays_gpu": int(st["ray_count"]), "rays_oracle": int(ost.ray_count)}
    rows.append(row)
    print(json.dumps(row), flush=True)


for name in SCENES:
    s = make_scene(name)
    sample_row(name, s)
    image_row(name, s, s.spp, s.seed)
    image_row(name + " @64spp", s, 64, s.seed)

s = api.Scene.from_psy_string(cornell_text())
s.set_render(res=(512, 512), spp=16, seed=1)
image_row("C1 cornell_box.psy 512x512 16spp (BASELINE config 0, full size)", s, 16, 1)
s = api.Scene.from_psy_string(cube_text(res=(1920, 1080), spp=64, camera_motion=True))
image_row("C2 cube.psy 1080p 64spp crop (BASELINE config 1)", s, 64, s.seed, crop=(704, 348, 704 + 511, 348 + 383))
import bench
kw = dict(bench.WORKLOADS["C5"])
s = api.Scene.synth_instanced(**kw)
image_row("C5 scene, 256x144 crop of the 4K frame at 64 spp", s, 64, kw["seed"], crop=(1792, 1008, 1792 + 255, 1008 + 143))
kw = dict(bench.WORKLOADS["C5r1"])
s = api.Scene.synth_instanced(**kw)
image_row("C5 scene of round 1 (single-key nodes), same crop at 64 spp", s, 64, kw["seed"], crop=(1792, 1008, 1792 + 255, 1008 + 143))
with open(out_path, "w") as f:
    json.dump({"what": "product build (libdevice) vs oracle (glibc), same seed / spp / sample sequences", "rows": rows}, f, indent=1)
