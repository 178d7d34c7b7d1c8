"""mesh BVHs whose triangle order and topology come from the GPU LBVH build (psy_build_mesh_topology,
dbuild.cuh).  The tree differs from the host's SAH tree, the hits must not: the CUDA path and the oracle traverse the
SAME GPU-built tree bit for bit, and both find the closest hits of the SAH-built scene."""
import numpy as np
import pytest

import oracle_lib as O
from psychopath_b200 import abi, api
from test_gpu_parity import probe_rays
from test_host_and_abi import check_bvh

pytestmark = pytest.mark.gpu


def build_pair(make):
    sah = make()
    api.enable_gpu_build(min_tris=64)
    try:
        gpu = make()
    finally:
        api.disable_gpu_build()
    secs, n = api.gpu_build_seconds()
    assert n >= 1 and secs > 0
    return sah, gpu


@pytest.mark.parametrize("make", [
    lambda: api.Scene.synth_displaced_grid(160, n_vert_keys=4, n_xform_keys=4, res=(192, 108), spp=4, seed=3),
    lambda: api.Scene.synth_instanced(asset_quads=80, n_assets=3, n_clusters=9, inst_per_cluster=5, lights_per_cluster=4,
                                      cluster_xform_keys=2, res=(96, 54), spp=4, seed=4, motion_pct=1250),
])
def test_gpu_built_tree_is_valid_and_gives_the_same_hits(make):
    sah, gpu = build_pair(make)
    d = gpu.desc.contents
    for m in range(d.n_meshes):
        mesh = d.meshes[m]
        check_bvh(d, mesh.root, True, mesh.tri_base, mesh.tri_count)
    rays = probe_rays(gpu, n_random=12000)
    gh = api.Tracer(gpu).trace(rays)
    oh, _, cnt_gpu_tree = O.trace_rays(gpu.desc, rays)
    for k in ("kind", "prim_id", "inst_id", "path_id", "t", "b0", "b1", "b2"):
        assert np.array_equal(gh[k], oh[k]), k  # CUDA path == oracle on the same GPU-built tree
    sh, _, cnt_sah_tree = O.trace_rays(sah.desc, rays)
    closest = (rays["flags"] & 1) == 0  # occlusion rays stop at the first hit found, which depends on the tree
    assert np.array_equal(oh["kind"][closest], sh["kind"][closest])
    hit = closest & (oh["kind"] == abi.HIT_TRIANGLE)
    assert hit.sum() > 500
    assert np.array_equal(oh["t"][hit], sh["t"][hit])
    assert (oh["prim_id"][hit] != sh["prim_id"][hit]).mean() < 0.01  # exact ties only
    # the LBVH tree is allowed to cost more visits than the SAH tree, within reason
    print(f"[gpu build] node visits per ray: LBVH {cnt_gpu_tree[0] / len(rays):.1f}  SAH {cnt_sah_tree[0] / len(rays):.1f}")
    assert cnt_gpu_tree[0] < 3.0 * cnt_sah_tree[0]
    # and the rendered images agree to within sampling-independent float noise (same samples, same hits)
    a, _ = api.Renderer(sah).render()
    b, _ = api.Renderer(gpu).render()
    rel = np.linalg.norm(a - b) / np.linalg.norm(a)
    assert rel < 2e-3, rel
