"""SURVEY.md N3, host half: a mesh BVH whose triangle order and binary topology come from an external builder
(on a GPU box: psy_build_mesh_topology, the LBVH build; here a ctypes stand-in that sorts by centroid and builds a
balanced tree).  The host derives bounds / the BVH4 collapse from it; the resulting scene must be a valid BVH4 and give
the same closest hits as the SAH-built scene."""
import ctypes as C

import numpy as np

import oracle_lib as O
from psychopath_b200 import abi, api
from test_host_and_abi import check_bvh

TOPO_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.POINTER(C.c_float), C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.c_uint32,
                      C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_uint8),
                      C.POINTER(C.c_uint32), C.POINTER(C.c_int32), C.POINTER(C.c_uint32))
calls = []


def stand_in(ctx, verts, n_verts, n_keys, tri_idx, n_tris, per_leaf, order, left, right, axis, n_internal, root, depth):
    v = np.ctypeslib.as_array(verts, shape=(n_verts, n_keys, 3))
    t = np.ctypeslib.as_array(tri_idx, shape=(n_tris, 3))
    cent = v[t][:, :, 0, :].mean(axis=1)  # key 0 centroids
    perm = np.lexsort((cent[:, 2], cent[:, 1], cent[:, 0])).astype(np.uint32)
    np.ctypeslib.as_array(order, shape=(n_tris,))[:] = perm
    nxt = [0]
    dmax = [0]

    def build(lo, hi, d):  # sorted positions [lo, hi)
        if hi - lo <= per_leaf:
            dmax[0] = max(dmax[0], d)
            return ~((lo << 2) | (hi - lo))  # leaf: (first position << 2) | count
        me = nxt[0]
        nxt[0] += 1
        mid = (lo + hi) // 2
        axis[me] = d % 3
        left[me] = build(lo, mid, d + 1)
        right[me] = build(mid, hi, d + 1)
        return me
    root[0] = build(0, n_tris, 0)
    n_internal[0] = nxt[0]
    depth[0] = dmax[0]
    calls.append(int(n_tris))
    return 0


def closest_hits(scene, rays):
    h, _, _ = O.trace_rays(scene.desc, rays)
    return h


def test_mesh_bvh_from_an_external_topology_is_valid_and_finds_the_same_hits():
    host = abi.load_host()
    sah = api.Scene.synth_displaced_grid(24, n_vert_keys=4, n_xform_keys=2, res=(64, 36), spp=1, seed=3)
    cb = TOPO_FN(stand_in)
    try:
        host.psyh_set_mesh_topology_builder(C.cast(cb, C.c_void_p), None, 16)
        ext = api.Scene.synth_displaced_grid(24, n_vert_keys=4, n_xform_keys=2, res=(64, 36), spp=1, seed=3)
        inst = api.Scene.synth_instanced(asset_quads=12, n_assets=2, n_clusters=4, inst_per_cluster=3, lights_per_cluster=2,
                                         res=(48, 27), spp=1)
    finally:
        host.psyh_set_mesh_topology_builder(None, None, 0)
    assert len(calls) >= 3 and min(calls) >= 16
    n = C.c_uint32()
    assert host.psyh_take_mesh_topology_seconds(C.byref(n)) > 0 and n.value == len(calls)
    for s in (ext, inst):
        d = s.desc.contents
        for m in range(d.n_meshes):
            mesh = d.meshes[m]
            depth = check_bvh(d, mesh.root, True, mesh.tri_base, mesh.tri_count)
            assert depth <= mesh.bvh_depth + 1
            orig = sorted(d.tris[4 * (mesh.tri_base + t) + 3] for t in range(mesh.tri_count))
            assert orig == list(range(mesh.tri_count))
    da, db = sah.desc.contents, ext.desc.contents
    assert da.n_tris == db.n_tris and da.n_nodes > 0 and db.n_nodes > 0
    rays = O.camera_rays(sah.desc, 0, 0, 64, 36, 2, 3)
    rng = np.random.default_rng(5)
    rays["time"] = rng.random(len(rays)).astype(np.float32)
    ha, hb = closest_hits(sah, rays), closest_hits(ext, rays)
    assert (ha["kind"] == abi.HIT_TRIANGLE).sum() > 500
    assert np.array_equal(ha["kind"], hb["kind"])
    hit = ha["kind"] == abi.HIT_TRIANGLE
    # a different tree changes visit order, so exact ties may resolve to another triangle; t agrees regardless
    assert np.array_equal(ha["t"][hit], hb["t"][hit])
    assert (ha["prim_id"][hit] != hb["prim_id"][hit]).mean() < 0.01
