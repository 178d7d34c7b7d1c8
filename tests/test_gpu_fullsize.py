"""Probe-ray parity through the BASELINE configurations at their FULL scene sizes (the oracle traverses the same
10 M-triangle / 500 M-effective-triangle scenes; a probe set costs it milliseconds once the scene is built).

Bar: ids, t and barycentrics bit-identical (the tie exemption of the north star is not needed)."""
import numpy as np
import pytest

import oracle_lib as O
from psychopath_b200 import abi, api
from test_gpu_parity import probe_rays

pytestmark = pytest.mark.gpu


def check(scene, n_random):
    tr = api.Tracer(scene)
    rays = probe_rays(scene, n_random=n_random)
    gh, gg = tr.trace(rays, want_geo=True)
    oh, og, _ = O.trace_rays(scene.desc, rays, want_geo=True)
    for k in ("kind", "prim_id", "inst_id", "path_id"):
        assert np.array_equal(gh[k], oh[k]), (k, int((gh[k] != oh[k]).sum()))
    hit = (oh["kind"] != abi.HIT_MISS) & (oh["kind"] != abi.HIT_OCCLUDED)
    assert hit.sum() > 1000
    tri = hit & (oh["kind"] == abi.HIT_TRIANGLE)
    for k in ("t", "b0", "b1", "b2"):
        assert np.array_equal(gh[k][tri], oh[k][tri]), k
    rel_t = np.abs(gh["t"][hit] - oh["t"][hit]) / np.abs(oh["t"][hit])
    assert rel_t.max() <= 1e-5
    for k in ("pos", "nor", "nor_g"):
        a, b = gg[k][hit], og[k][hit]
        scale = np.maximum(np.abs(b).max(axis=1, keepdims=True), 1e-20)
        assert (np.abs(a - b) / scale).max() <= 1e-5, k
    return int(hit.sum())


def test_config_c3_ten_million_triangles_four_plus_four_keys():
    """BASELINE config 2: 10.0 M-triangle displaced mesh, 4 deformation keys + 4 transform keys, 1080p."""
    scene = api.Scene.synth_displaced_grid(2240, n_vert_keys=4, n_xform_keys=4, res=(1920, 1080), spp=128, seed=3)
    assert scene.desc.contents.n_tris >= 9_999_000
    check(scene, 20000)


def test_config_c4_c5_full_instanced_scene():
    """BASELINE configs 3 / 4: 100 clusters x (10 instances of 4 x 500k-triangle assets + 10 lights); C5 adds two
    transform keys per cluster instance and the 4K resolution."""
    import bench
    for name in ("C4", "C5"):
        scene = api.Scene.synth_instanced(**bench.WORKLOADS[name])
        assert scene.desc.contents.n_tris >= 4 * 500_000  # + the ground plane
        check(scene, 20000)
