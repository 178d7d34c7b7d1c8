"""mesh BVHs built on the GPU (dbuild.cuh) — the whole BVH4 (psy_mesh_bvh_build: LBVH topology, bounds,
union rule, collapse) or only triangle order and topology (psy_build_mesh_topology, the host doing the rest).  The tree
differs from the host's SAH tree, the hits must not: the CUDA path and the oracle traverse the SAME GPU-built tree bit
for bit, and both find the closest hits of the SAH-built scene.  And the two GPU modes must agree with each other:
same topology, same arithmetic, hence the same nodes bit for bit, in another order."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
from psychopath_b200 import abi, api
from test_gpu_parity import probe_rays
from test_host_and_abi import check_bvh

pytestmark = pytest.mark.gpu


def build_pair(make, mode="bvh4"):
    sah = make()
    api.enable_gpu_build(min_tris=64, mode=mode)
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
@pytest.mark.parametrize("mode", ["bvh4", "topology"])
def test_gpu_built_tree_is_valid_and_gives_the_same_hits(make, mode):
    sah, gpu = build_pair(make, mode)
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


def node_bits(d, ref):
    """Everything of a BVH4 node but its child references, as bytes: code, counts, bounds and motion keys."""
    n = d.nodes[ref]
    out = [bytes(C.string_at(C.addressof(n.bounds), C.sizeof(n.bounds))),
           (n.traversal_code, n.n_keys, n.n_children).__repr__().encode()]
    if n.n_keys > 1:
        for k in range(n.n_keys):
            kb = d.motion_bounds[n.key_base + k]
            out.append(bytes(C.string_at(C.addressof(kb), C.sizeof(kb))))
    return b"".join(out)


def same_tree(da, root_a, db, root_b):
    """Walks two BVH4s in step: equal node contents, equal leaves, whatever the node numbering."""
    stack = [(root_a, root_b)]
    nodes = 0
    while stack:
        a, b = stack.pop()
        leaf_a, leaf_b = bool(a & 0x80000000), bool(b & 0x80000000)
        assert leaf_a == leaf_b
        if leaf_a:
            assert a == b
            continue
        assert node_bits(da, a) == node_bits(db, b), f"node {a} / {b}"
        nodes += 1
        na, nb = da.nodes[a], db.nodes[b]
        for c in range(4):
            if c < na.n_children:
                stack.append((na.child[c], nb.child[c]))
            else:
                assert na.child[c] == nb.child[c] == abi.PSY_NONE
    return nodes


@pytest.mark.parametrize("make", [
    # 4 vertex keys: nodes with one key and with four, the union rule deciding at every level
    lambda: api.Scene.synth_displaced_grid(120, n_vert_keys=4, n_xform_keys=2, res=(64, 36), spp=1, seed=3),
    lambda: api.Scene.synth_displaced_grid(90, n_vert_keys=1, n_xform_keys=1, res=(64, 36), spp=1, seed=9),
    lambda: api.Scene.synth_displaced_grid(70, n_vert_keys=3, n_xform_keys=1, res=(64, 36), spp=1, seed=5),
    lambda: api.Scene.synth_instanced(asset_quads=60, n_assets=4, n_clusters=4, inst_per_cluster=3, lights_per_cluster=2,
                                      cluster_xform_keys=2, res=(64, 36), spp=1, seed=4),
])
def test_device_built_bvh4_is_the_hosts_bvh4_of_the_same_topology(make):
    scenes = {}
    for mode in ("bvh4", "topology"):
        api.enable_gpu_build(min_tris=64, mode=mode)
        try:
            scenes[mode] = make()
        finally:
            api.disable_gpu_build()
        assert api.gpu_build_seconds()[1] >= 1
    da, db = scenes["bvh4"].desc.contents, scenes["topology"].desc.contents
    assert da.n_meshes == db.n_meshes and da.n_nodes == db.n_nodes and da.n_motion_bounds == db.n_motion_bounds
    assert da.n_tris == db.n_tris
    ta = np.ctypeslib.as_array(da.tris, shape=(da.n_tris * 4,))
    tb = np.ctypeslib.as_array(db.tris, shape=(db.n_tris * 4,))
    assert np.array_equal(ta, tb)  # the same triangle order
    multi_key = walked = 0
    for m in range(da.n_meshes):
        ma, mb = da.meshes[m], db.meshes[m]
        assert (ma.tri_base, ma.tri_count, ma.bvh_depth, ma.bounds_count) == (mb.tri_base, mb.tri_count, mb.bvh_depth,
                                                                             mb.bounds_count)
        for k in range(ma.bounds_count):
            ba, bb = da.obj_bounds[ma.bounds_base + k], db.obj_bounds[mb.bounds_base + k]
            assert bytes(C.string_at(C.addressof(ba), C.sizeof(ba))) == bytes(C.string_at(C.addressof(bb), C.sizeof(bb)))
        walked += same_tree(da, ma.root, db, mb.root)  # (0 for a mesh that is one leaf, like the ground plane)
        multi_key += sum(1 for j in range(da.n_nodes) if da.nodes[j].n_keys > 1) if m == 0 else 0
    assert walked > 100
    if da.meshes[0].n_keys > 1:
        assert multi_key > 0  # the motion-key path of the device build is exercised


@pytest.mark.parametrize("seed", range(16))
def test_random_scenes_build_the_same_on_device_and_host_and_hit_the_same(seed):
    """The fuzz scenes' small meshes (boxes, fans, ragged patches of quads and triangles, 1-3 time keys): Morton ties,
    leaves of one, two and three triangles at every depth, subtrees that end at odd and even depth."""
    from scenes import random_scene_text
    text = random_scene_text(seed)
    scenes = {"sah": api.Scene.from_psy_string(text)}
    for mode in ("bvh4", "topology"):
        api.enable_gpu_build(min_tris=4, mode=mode)
        try:
            scenes[mode] = api.Scene.from_psy_string(text)
        finally:
            api.disable_gpu_build()
        assert api.gpu_build_seconds()[1] >= 1
    da, db = scenes["bvh4"].desc.contents, scenes["topology"].desc.contents
    assert (da.n_meshes, da.n_nodes, da.n_motion_bounds, da.n_tris) == (db.n_meshes, db.n_nodes, db.n_motion_bounds, db.n_tris)
    assert np.array_equal(np.ctypeslib.as_array(da.tris, shape=(da.n_tris * 4,)),
                          np.ctypeslib.as_array(db.tris, shape=(db.n_tris * 4,)))
    for m in range(da.n_meshes):
        ma, mb = da.meshes[m], db.meshes[m]
        assert (ma.bvh_depth, ma.bounds_count) == (mb.bvh_depth, mb.bounds_count)
        same_tree(da, ma.root, db, mb.root)
        check_bvh(da, ma.root, True, ma.tri_base, ma.tri_count)
    rays = probe_rays(scenes["bvh4"], n_random=6000, seed=500 + seed)
    gh = api.Tracer(scenes["bvh4"]).trace(rays)
    oh = O.trace_rays(da, rays)[0]
    for k in ("kind", "prim_id", "inst_id", "path_id", "t", "b0", "b1", "b2"):
        assert np.array_equal(gh[k], oh[k]), k
    sh = O.trace_rays(scenes["sah"].desc, rays)[0]
    closest = (rays["flags"] & 1) == 0
    assert np.array_equal(oh["kind"][closest], sh["kind"][closest])
    hit = closest & (oh["kind"] == abi.HIT_TRIANGLE)
    assert np.array_equal(oh["t"][hit], sh["t"][hit])


def test_cli_gpu_build_flag(tmp_path):
    """`psychopath --synth C3 --gpu-build`: BASELINE config 2's 10 M triangles with both mesh BVH4s from the device, and
    the picture of the SAH-built scene (same samples, same closest hits; only exact ties may pick another triangle)."""
    import os
    import subprocess
    import cv2
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(repo, "bin", "psychopath")
    imgs = {}
    for name, flags in (("sah", []), ("gpu", ["--gpu-build"])):
        d = tmp_path / name
        d.mkdir()
        out = subprocess.run([exe, "--synth", "C3", "--spp", "2"] + flags, cwd=d, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stderr
        if flags:
            assert "2 mesh BVHs built on the GPU" in out.stdout, out.stdout
        pngs = [f for f in os.listdir(d) if f.endswith(".png")]
        assert len(pngs) == 1
        imgs[name] = cv2.imread(str(d / pngs[0])).astype(np.int32)
    assert imgs["sah"].shape == imgs["gpu"].shape == (1080, 1920, 3) and imgs["sah"].mean() > 2
    assert (np.abs(imgs["sah"] - imgs["gpu"]).max(axis=-1) > 2).mean() < 1e-3
