"""Independent check of what the oracle and the CUDA path SHARE: the host-built scene (BVH4s, leaf order, ids,
instance tables, composed transforms) and the conventions behind it.

A second, deliberately different ray caster is written here in numpy float64: no BVH, no code from oracle/ or
csrc/host, a different triangle test (Moeller-Trumbore instead of Woop), rays taken INTO object space instead
of triangles out to world space (SURVEY.md F6 — the ray parameter t is invariant under the affine map, so
both formulations must agree), instance hierarchies walked recursively from the assembly / instance / object
tables.  It sees every triangle of every instance, so a triangle the builder lost, a wrong id, a transposed
matrix, a world->local / local->world mix-up or a wrong time-key interpolation in the shared build shows up as a
disagreement with the oracle's BVH traversal.  (tests/test_gpu_parity.py then ties the CUDA path to the oracle bit
for bit on the same scenes.)

Bars: same hit / miss, same (instance id chain's leaf id, original triangle index) unless the two nearest
candidates are within 1e-6 relative (tie exemption of the north star), |t - t_oracle| <= 1e-5 t.
"""
import numpy as np
import pytest

import oracle_lib as O
from psychopath_b200 import abi, api
from scenes import motion_instances_text, nested_assemblies_text


def arr(ptr, n, dtype, cols=None):
    if n == 0:
        return np.zeros((0,) if cols is None else (0, cols), dtype=dtype)
    a = np.ctypeslib.as_array(ptr, shape=(n,))
    a = np.frombuffer(a, dtype=dtype).copy() if a.dtype != dtype else a.copy()
    return a if cols is None else a.reshape(-1, cols)


class FlatScene:
    def __init__(self, scene):
        d = scene.desc.contents
        self.d = d
        self.tris = np.array([d.tris[i] for i in range(4 * d.n_tris)], dtype=np.int64).reshape(-1, 4) if d.n_tris < 5000 \
            else np.frombuffer((abi.u32 * (4 * d.n_tris)).from_address(abi.C.addressof(d.tris.contents)), dtype=np.uint32).astype(np.int64).reshape(-1, 4)
        nv = 3 * d.n_vert_keys
        self.verts = np.frombuffer((abi.f32 * nv).from_address(abi.C.addressof(d.verts.contents)), dtype=np.float32).astype(np.float64).reshape(-1, 3)
        self.xforms = np.zeros((d.n_xforms, 4, 4))
        for i in range(d.n_xforms):
            m = np.array([[d.xforms[i].r[r][c] for c in range(4)] for r in range(3)], dtype=np.float64)
            self.xforms[i, :3, :] = m
            self.xforms[i, 3, 3] = 1.0


def lerp_keys(keys, t):
    """src/lerp.rs:20-35 semantics in float64: piecewise linear over n equally spaced keys."""
    n = len(keys)
    if n == 1 or t >= 1.0:
        return keys[-1]
    x = t * (n - 1)
    i = int(np.floor(x))
    a = x - i
    return keys[i] * (1.0 - a) + keys[i + 1] * a


def compose(parent, child):
    """TransformStack::push (transform_stack.rs:33-62): key-wise product, the shorter series resampled."""
    if parent is None:
        return child
    if child is None:
        return parent
    n = max(len(parent), len(child))

    def at(keys, i):
        return keys[i] if len(keys) == n else lerp_keys(keys, i / (n - 1))
    # world->local of the nested instance = (parent-local -> child-local) after (world -> parent-local)
    return [at(child, i) @ at(parent, i) for i in range(n)]


def moller_trumbore(o, d, p0, p1, p2, max_t):
    """All triangles at once; returns (t, index) of the nearest hit with 0 < t <= max_t, or (inf, -1)."""
    e1, e2 = p1 - p0, p2 - p0
    pv = np.cross(d, e2)
    det = np.einsum("ij,ij->i", e1, pv)
    ok = np.abs(det) > 0.0
    inv = np.where(ok, 1.0 / np.where(ok, det, 1.0), 0.0)
    tv = o - p0
    u = np.einsum("ij,ij->i", tv, pv) * inv
    qv = np.cross(tv, e1)
    v = np.einsum("ij,ij->i", qv, np.broadcast_to(d, qv.shape)) * inv
    t = np.einsum("ij,ij->i", e2, qv) * inv
    hit = ok & (u >= 0) & (v >= 0) & (u + v <= 1) & (t > 0) & (t <= max_t)
    t = np.where(hit, t, np.inf)
    return t


def brute_force(fs, ray):
    """Nearest and second nearest mesh hit of one ray: [(t, inst_id, prim_id)], by walking every instance."""
    d = fs.d
    o, dr, time, max_t = ray["orig"].astype(np.float64), ray["dir"].astype(np.float64), float(ray["time"]), float(ray["max_t"])
    cands = []

    def walk(asm_index, keys):
        a = d.assemblies[asm_index]
        for i in range(a.instance_count):
            inst = d.instances[a.instance_base + i]
            own = [fs.xforms[inst.xform_base + k] for k in range(inst.xform_count)] if inst.xform_count else None
            k2 = compose(keys, own)
            if inst.kind == 1:
                walk(inst.data_index, k2)
                continue
            obj = d.objects[inst.data_index]
            if obj.kind != 0:
                continue
            m = d.meshes[obj.index]
            ol, dl = o, dr
            if k2 is not None:
                M = lerp_keys(np.array(k2), time)
                ol = M[:3, :3] @ o + M[:3, 3]
                dl = M[:3, :3] @ dr
            tri = fs.tris[m.tri_base:m.tri_base + m.tri_count]
            nk = m.n_keys

            def vert(col):
                idx = m.vert_base + tri[:, col] * nk
                if nk == 1:
                    return fs.verts[idx]
                x = min(time, 1.0) * (nk - 1)
                i0 = min(int(np.floor(x)), nk - 2) if time < 1.0 else nk - 2
                al = x - i0
                return fs.verts[idx + i0] * (1.0 - al) + fs.verts[idx + i0 + 1] * al
            t = moller_trumbore(ol, dl, vert(0), vert(1), vert(2), max_t)
            order = np.argsort(t)[:2]
            for j in order:
                if np.isfinite(t[j]):
                    cands.append((float(t[j]), int(inst.id), int(tri[j, 3]), int(obj.index)))
    walk(d.root_assembly, None)
    cands.sort()
    return cands[:2]


def make(name):
    if name == "fixture":
        return api.Scene.from_psy_string(motion_instances_text())
    if name == "nested":
        return api.Scene.from_psy_string(nested_assemblies_text())
    if name == "instanced_motion":
        return api.Scene.synth_instanced(asset_quads=12, n_assets=3, n_clusters=6, inst_per_cluster=4,
                                         lights_per_cluster=2, cluster_xform_keys=2, res=(48, 27), spp=2, seed=4)
    if name == "deforming_grid":
        return api.Scene.synth_displaced_grid(20, n_vert_keys=4, n_xform_keys=4, res=(48, 27), spp=2, seed=3)
    raise KeyError(name)


@pytest.mark.parametrize("name", ["fixture", "nested", "instanced_motion", "deforming_grid"])
def test_bvh_traversal_agrees_with_independent_brute_force(name):
    check_against_brute_force(make(name), 80)


@pytest.mark.parametrize("seed,many_keys", [(0, False), (3, False), (7, False), (11, False), (101, True), (104, True)])
def test_random_scenes_agree_with_independent_brute_force(seed, many_keys):
    """The fuzz scenes of tests/test_gpu_fuzz.py (assemblies three deep, 0-3 or up to 12 time keys on vertices and
    transforms, fanned polygons): what the host built from them, seen from outside."""
    from scenes import random_scene_text
    check_against_brute_force(api.Scene.from_psy_string(random_scene_text(seed, many_keys=many_keys)), 60)


def check_against_brute_force(scene, min_hits):
    d = scene.desc.contents
    fs = FlatScene(scene)
    cam = O.camera_rays(scene.desc, 0, 0, d.res_x, d.res_y, 1, d.seed)
    rng = np.random.default_rng(11)
    rays = cam[rng.choice(len(cam), size=min(len(cam), 260), replace=False)].copy()
    rays["time"][:20] = 1.0
    rays["time"][20:40] = 0.0
    # a few rays from inside the scene in random directions
    extra = rays[:60].copy()
    dirs = rng.normal(0, 1, (60, 3))
    extra["dir"] = (dirs / np.linalg.norm(dirs, axis=1, keepdims=True)).astype(np.float32)
    extra["orig"] = (extra["orig"] * 0.5).astype(np.float32)
    extra["time"] = rng.random(60).astype(np.float32)
    rays = np.concatenate([rays, extra])
    hits, _, _ = O.trace_rays(scene.desc, rays)
    n_hit = n_tie = 0
    for r, h in zip(rays, hits):
        c = brute_force(fs, r)
        if h["kind"] in (abi.HIT_SPHERE_LIGHT, abi.HIT_RECT_LIGHT):
            assert not c or c[0][0] >= h["t"] * (1 - 1e-5)  # the light is in front of every mesh
            continue
        if not c:
            assert h["kind"] == abi.HIT_MISS, (r, h)
            continue
        assert h["kind"] == abi.HIT_TRIANGLE, (r, h, c)
        n_hit += 1
        assert abs(h["t"] - c[0][0]) <= 1e-5 * c[0][0] + 1e-7, (h["t"], c)
        tie = len(c) > 1 and abs(c[1][0] - c[0][0]) <= 1e-6 * c[0][0]
        if tie:
            n_tie += 1
            assert (int(h["inst_id"]), int(h["prim_id"])) in {(x[1], x[2]) for x in c}
        else:
            assert (int(h["inst_id"]), int(h["prim_id"])) == (c[0][1], c[0][2]), (h, c)
    assert n_hit >= min_hits, n_hit
