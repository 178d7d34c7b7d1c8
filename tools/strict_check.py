"""Strict-mode parity report (needs a GPU): per scene, per-sample radiance of libpsy_cuda_strict.so against the
oracle in strict-math mode (expected: bit-identical), and the strict-film image against the oracle's image."""
import os
import sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
import numpy as np
import oracle_lib as O
from psychopath_b200 import api
from test_gpu_parity import SCENES, make_scene

names = sys.argv[1:] or SCENES
for name in names:
    scene = make_scene(name)
    d = scene.desc.contents
    w, h = min(d.res_x, 40), min(d.res_y, 24)
    x0, y0 = (d.res_x - w) // 2, (d.res_y - h) // 2
    r = api.Renderer(scene, strict=True)
    gs, st = r.render_bucket_samples(x0, y0, w, h, 4, d.seed)
    with O.strict_math():
        oxyz, osamp, ost = O.render_bucket(scene.desc, x0, y0, w, h, 4, d.seed, want_samples=True)
        oimg, oist, _ = O.render_image(scene.desc, d.spp, d.seed, spb=4096)
    same = (gs.view(np.uint32) == osamp.view(np.uint32)).all(axis=-1) | (np.isnan(gs).all(axis=-1) & np.isnan(osamp).all(axis=-1))
    gimg, gst = r.render(strict_film=True, max_samples_per_bucket=4096)
    eq = np.array_equal(gimg.view(np.uint32), oimg.view(np.uint32)) or np.array_equal(gimg, oimg, equal_nan=True)
    nd = int((gimg.view(np.uint32) != oimg.view(np.uint32)).any(axis=-1).sum())
    fin = np.isfinite(gimg).all(axis=-1) & np.isfinite(oimg).all(axis=-1)
    rel = float(np.linalg.norm((gimg - oimg)[fin]) / max(np.linalg.norm(oimg[fin]), 1e-30))
    print(f"{name}: samples {same.size} differing {int((~same).sum())}  rays gpu {st['ray_count']} oracle {ost.ray_count} | "
          f"image bit-equal {eq} differing pixels {nd} rel-L2 {rel:.3e} rays gpu {gst['ray_count']} oracle {oist.ray_count}", flush=True)
    bad = np.argwhere(~same)
    for b in bad[:3]:
        print("   sample", tuple(b), "gpu", gs[tuple(b)], "oracle", osamp[tuple(b)])
