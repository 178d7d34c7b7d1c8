"""Fuzz parity: seeded random scenes in the reference's text format (tests/scenes.py::random_scene_text), loaded by
the same host code on both sides, rendered by the CUDA path and by the oracle.

The hand-written fixtures pin the feature combinations somebody thought of; these pin the ones nobody did: time keys
on every quantity that can have them, all colour kinds on all shader and light kinds, polygons that the loader fans,
assemblies three deep instanced with 0-3 transform keys.  Bars: the strict build has the oracle's bits for every
sample and (film in the reference's order) every pixel; the product build's closest hits are the oracle's."""
import numpy as np
import pytest

import oracle_lib as O
from psychopath_b200 import abi, api
from scenes import random_scene_text
from test_gpu_parity import probe_rays
from test_gpu_strict import same_or_both_nan

pytestmark = pytest.mark.gpu

SEEDS = list(range(48))


@pytest.mark.parametrize("seed", SEEDS)
def test_random_scene_strict_build_has_the_oracles_bits(seed):
    scene = api.Scene.from_psy_string(random_scene_text(seed))
    d = scene.desc.contents
    r = api.Renderer(scene, strict=True)
    gs, st = r.render_bucket_samples(0, 0, d.res_x, d.res_y, d.spp, d.seed)
    gimg, gst = r.render(strict_film=True, max_samples_per_bucket=4096)
    with O.strict_math():
        _, osamp, ost = O.render_bucket(scene.desc, 0, 0, d.res_x, d.res_y, d.spp, d.seed, want_samples=True)
        oimg, oist, _ = O.render_image(scene.desc, d.spp, d.seed, spb=4096)
    assert st["ray_count"] == ost.ray_count and gst["ray_count"] == oist.ray_count
    assert ost.ray_count > d.res_x * d.res_y * d.spp  # the paths do go on after the camera ray
    ok = same_or_both_nan(gs, osamp)
    assert ok.all(), f"{(~ok).any(axis=-1).sum()} of {ok.shape[0]} samples differ"
    ok = same_or_both_nan(gimg, oimg)
    assert ok.all(), f"{(~ok).any(axis=-1).sum()} pixels differ"


@pytest.mark.parametrize("seed", SEEDS[::3])
def test_random_scene_probe_rays(seed):
    scene = api.Scene.from_psy_string(random_scene_text(seed))
    rays = probe_rays(scene, n_random=8000, seed=1000 + seed)
    gh = api.Tracer(scene).trace(rays)
    oh = O.trace_rays(scene.desc, rays)[0]
    same = (gh["kind"] == oh["kind"]) & (gh["prim_id"] == oh["prim_id"]) & (gh["inst_id"] == oh["inst_id"]) & \
           (gh["path_id"] == oh["path_id"])
    assert (~same).mean() <= 1e-4, f"{(~same).sum()} of {len(rays)} rays disagree on ids"
    tri = same & (oh["kind"] == abi.HIT_TRIANGLE)
    assert tri.sum() > 100
    for k in ("t", "b0", "b1", "b2"):
        assert np.array_equal(gh[k][tri], oh[k][tri]), k


@pytest.mark.parametrize("seed", SEEDS[1::6])
def test_random_scene_product_build_within_the_gate(seed):
    scene = api.Scene.from_psy_string(random_scene_text(seed))
    scene.set_render(spp=16)
    d = scene.desc.contents
    gimg, st = api.Renderer(scene).render()
    oimg, ost, _ = O.render_image(scene.desc, d.spp, d.seed, spb=4096)
    assert abs(int(st["ray_count"]) - int(ost.ray_count)) <= 0.001 * ost.ray_count + 4
    rel_l2 = np.linalg.norm(gimg - oimg) / max(np.linalg.norm(oimg), 1e-20)
    print(f"[fuzz parity] seed {seed}: rel-L2 {rel_l2:.3e}")
    assert rel_l2 <= 2e-3


@pytest.mark.parametrize("seed", range(100, 108))
def test_random_scene_with_nine_and_twelve_time_keys(seed):
    """Meshes and instance transforms with more than eight time keys: the trace kernel's node references hold the key
    count only up to seven; beyond that it comes from the node header (dpool.cuh, kc == 7)."""
    scene = api.Scene.from_psy_string(random_scene_text(seed, many_keys=True))
    d = scene.desc.contents
    assert max(d.nodes[i].n_keys for i in range(d.n_nodes)) >= 8
    r = api.Renderer(scene, strict=True)
    gs, st = r.render_bucket_samples(0, 0, d.res_x, d.res_y, d.spp, d.seed)
    with O.strict_math():
        _, osamp, ost = O.render_bucket(scene.desc, 0, 0, d.res_x, d.res_y, d.spp, d.seed, want_samples=True)
    assert st["ray_count"] == ost.ray_count
    ok = same_or_both_nan(gs, osamp)
    assert ok.all(), f"{(~ok).any(axis=-1).sum()} of {ok.shape[0]} samples differ"
    rays = probe_rays(scene, n_random=6000, seed=2000 + seed)
    gh = api.Tracer(scene).trace(rays)
    oh = O.trace_rays(scene.desc, rays)[0]
    for k in ("kind", "prim_id", "inst_id", "path_id", "t", "b0", "b1", "b2"):
        assert np.array_equal(gh[k], oh[k]), k
    # the device BVH build refuses meshes of more than eight keys; the scene must still build (topology path) and hit the same
    api.enable_gpu_build(min_tris=4, mode="bvh4")
    try:
        gpu = api.Scene.from_psy_string(random_scene_text(seed, many_keys=True))
    finally:
        api.disable_gpu_build()
    closest = (rays["flags"] & 1) == 0
    bh = api.Tracer(gpu).trace(rays)
    assert np.array_equal(bh["kind"][closest], oh["kind"][closest])
    hit = closest & (oh["kind"] == abi.HIT_TRIANGLE)
    assert np.array_equal(bh["t"][hit], oh["t"][hit])
