"""CPU-only tests: the C ABI library loads and exports what include/psy_cuda.h declares, the ctypes
mirror matches the C structs, the host scene builders produce valid acceleration structures, and
the BVH traversal of the oracle agrees with brute force."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest

import oracle_lib as O
from psychopath_b200 import abi, api
from scenes import cornell_text, cube_text, motion_instances_text

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cuda_library_exports_every_declared_symbol():
    header = open(os.path.join(REPO, "include", "psy_cuda.h")).read()
    declared = set(re.findall(r"\b(psy_[a-z_]+)\s*\(", header)) - {"psy_bucket_cb"}
    assert declared == set(abi.CUDA_SYMBOLS), declared ^ set(abi.CUDA_SYMBOLS)
    lib = C.CDLL(abi.CUDA_LIB)
    for name in declared:
        assert hasattr(lib, name), name
    lib.psy_abi_version.restype = C.c_int32
    assert lib.psy_abi_version() == 2


def test_integration_doc_binds_every_declared_symbol():
    """INTEGRATION.md's reference-side binding (the Rust `extern "C"` block a maintainer would add) names exactly the
    functions include/psy_cuda.h declares."""
    header = open(os.path.join(REPO, "include", "psy_cuda.h")).read()
    declared = set(re.findall(r"\b(psy_[a-z_]+)\s*\(", header)) - {"psy_bucket_cb"}
    doc = open(os.path.join(REPO, "INTEGRATION.md")).read()
    bound = set(re.findall(r"pub fn (psy_[a-z_]+)\(", doc))
    assert bound == declared, bound ^ declared


def test_struct_layout_matches_c():
    """Compiles a tiny C program against the header and compares sizeof() with the ctypes mirror."""
    names = ["PsyXform", "PsyBBox", "PsyBBox4", "PsyBvh4Node", "PsyColor", "PsyShader", "PsyMesh", "PsyLight",
             "PsyDistantLight", "PsyObject", "PsyInstance", "PsyLightNode", "PsyAssembly", "PsyCamera", "PsySceneDesc",
             "PsyRay", "PsyHit", "PsyHitGeo", "PsyStats", "PsyRenderParams"]
    src = '#include <stdio.h>\n#include "psy_cuda.h"\nint main(void){' + "".join(
        f'printf("{n} %zu\\n", sizeof({n}));' for n in names) + "return 0;}"
    with tempfile.TemporaryDirectory() as td:
        cfile = os.path.join(td, "s.c")
        open(cfile, "w").write(src)
        exe = os.path.join(td, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(REPO, "include"), cfile, "-o", exe])
        out = subprocess.check_output([exe], text=True)
    sizes = dict(line.split() for line in out.strip().splitlines())
    for n in names:
        assert int(sizes[n]) == C.sizeof(getattr(abi, n)), n
    assert C.sizeof(abi.PsyBvh4Node) == 128
    assert abi.RAY_DTYPE.itemsize == C.sizeof(abi.PsyRay)
    assert abi.HIT_DTYPE.itemsize == C.sizeof(abi.PsyHit)
    assert abi.GEO_DTYPE.itemsize == C.sizeof(abi.PsyHitGeo)


def test_no_cpu_fallback():
    """Without a GPU the product path must fail loudly (PSY_ERR_NO_DEVICE), never fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    scene = api.Scene.from_psy_string(cornell_text())
    with pytest.raises(api.PsyError, match="no CUDA device|no CPU fallback"):
        api.Renderer(scene)


def test_cli_without_a_gpu_fails_loudly():
    """bin/psychopath parses its arguments like src/main.rs and, with no CUDA device, stops with the library's
    error instead of rendering on the CPU."""
    import subprocess
    import torch
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(repo, "bin", "psychopath")
    out = subprocess.run([exe, "--dev"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and "PsyBvh4Node size: 128 bytes" in out.stdout
    out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert out.returncode != 0 and "--input" in out.stderr
    out = subprocess.run([exe, "-i", "x.psy", "--crop", "5", "1", "2", "3"], capture_output=True, text=True, timeout=60)
    assert out.returncode != 0 and "X1 must be less than or equal to X2" in out.stderr
    if torch.cuda.is_available():
        return
    scene = os.path.join(repo, "tests", "golden", "cube.psy")
    out = subprocess.run([exe, "-i", scene], capture_output=True, text=True, timeout=120)
    assert out.returncode != 0 and "no CPU fallback" in out.stderr


def test_product_package_never_touches_the_oracle():
    pkg = os.path.join(REPO, "psychopath_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(root, f), errors="ignore").read()
                assert "liboracle" not in text and "oracle_lib" not in text and "oracle/" not in text.replace(
                    "oracle/osample.h for the", ""), f


@pytest.mark.parametrize("text,tris,lights,instances", [(cornell_text(), 36, 1, 8), (cube_text(), 14, 1, 3)])
def test_example_scenes_load(text, tris, lights, instances):
    """New-layout example scenes: counts of SURVEY.md 10."""
    s = api.Scene.from_psy_string(text)
    d = s.desc.contents
    assert (d.n_tris, d.n_lights, d.n_instances) == (tris, lights, instances)
    assert d.n_light_instances == 1 and d.assemblies[d.root_assembly].light_energy > 0
    # object-level SurfaceShaderBind is carried to the instance
    for i in range(d.n_instances):
        inst = d.instances[i]
        if d.objects[inst.data_index].kind == 0:
            assert inst.shader_index != abi.PSY_NONE


def test_loader_errors():
    with pytest.raises(api.PsyError, match="Output"):
        api.Scene.from_psy_string("Scene $s { RenderSettings { Resolution [4 4] SamplesPerPixel [1] } }")
    bad = motion_instances_text().replace("Data [$floor]", "Data [$nothing]")
    with pytest.raises(api.PsyError, match="doesn't exist"):
        api.Scene.from_psy_string(bad)
    with pytest.raises(api.PsyError, match="parse error"):
        api.Scene.from_psy_string("Scene $s { Output { Path [\"a.png\"] ")


def check_bvh(d, root, is_mesh, first, count):
    """Every object of [first, first+count) appears in exactly one leaf; child counts and codes are sane."""
    seen = np.zeros(count, dtype=np.int32)
    depth = 0
    stack = [(root, 1)]
    while stack:
        ref, dep = stack.pop()
        depth = max(depth, dep)
        if ref == abi.PSY_NONE:
            continue
        if ref & 0x80000000:
            start, n = ref & 0x0FFFFFFF, (ref >> 28) & 7
            assert 1 <= n <= (3 if is_mesh else 1)
            seen[start - first:start - first + n] += 1
            continue
        node = d.nodes[ref]
        assert 2 <= node.n_children <= 4 and node.traversal_code < 48 and node.n_keys >= 1
        for c in range(4):
            if c < node.n_children:
                stack.append((node.child[c], dep + 1))
            else:
                assert node.child[c] == abi.PSY_NONE
                assert node.bounds.v[0][c] == float("inf") and node.bounds.v[1][c] == float("-inf")
    assert (seen == 1).all()
    return depth


@pytest.mark.parametrize("make", [
    lambda: api.Scene.from_psy_string(motion_instances_text()),
    lambda: api.Scene.synth_instanced(asset_quads=20, n_assets=3, n_clusters=5, inst_per_cluster=4, lights_per_cluster=3,
                                      res=(32, 32), spp=1),
    lambda: api.Scene.synth_displaced_grid(40, 4, 4, res=(32, 32), spp=1),
])
def test_acceleration_structures_are_valid(make):
    s = make()
    d = s.desc.contents
    for m in range(d.n_meshes):
        mesh = d.meshes[m]
        check_bvh(d, mesh.root, True, mesh.tri_base, mesh.tri_count)
        # original indices are a permutation
        orig = sorted(d.tris[4 * (mesh.tri_base + t) + 3] for t in range(mesh.tri_count))
        assert orig == list(range(mesh.tri_count))
    for a in range(d.n_assemblies):
        asm = d.assemblies[a]
        check_bvh(d, asm.root, False, asm.instance_base, asm.instance_count)
        ids = sorted(d.instances[asm.instance_base + i].id for i in range(asm.instance_count))
        assert ids == list(range(asm.instance_count))
        # light tree: every light instance is reachable exactly once, energies add up
        if asm.light_root != abi.PSY_NONE:
            leaves, stack = [], [asm.light_root]
            while stack:
                n = d.light_nodes[stack.pop()]
                assert n.child_count <= 8
                if n.child_count == 0:
                    leaves.append(n.light_index)
                else:
                    kids = [d.light_nodes[n.child_base + i] for i in range(n.child_count)]
                    assert abs(sum(k.energy for k in kids) - n.energy) <= 1e-3 * max(n.energy, 1e-9)
                    stack.extend(n.child_base + i for i in range(n.child_count))
            assert sorted(leaves) == list(range(asm.light_instance_count))


def test_oracle_bvh_matches_brute_force():
    """Closest hit through the BVH == closest hit over all triangles (untransformed mesh, so the
    triangles are already in world space)."""
    text = """
Scene $s { Output { Path ["x.png"] } RenderSettings { Resolution [16 16] SamplesPerPixel [1] }
 Camera { Fov [60] Transform [1 0 0 0 0 1 0 0 0 0 1 0 0 0 0 1] }
 World { BackgroundShader { Type [Color] Color [rec709, 0 0 0] } }
 Assembly {
  SurfaceShader $m { Type [Lambert] Color [rec709, 0.5 0.5 0.5] }
  MeshSurface $g { Vertices [%s] FaceVertCounts [%s] FaceVertIndices [%s] }
  Instance { Data [$g] SurfaceShaderBind [$m] }
 } }"""
    rng = np.random.default_rng(5)
    n = 12
    zs = rng.random((n + 1, n + 1)) * 0.8
    verts = " ".join(f"{x / n * 4 - 2:.5f} {y / n * 4 - 2:.5f} {zs[y, x]:.5f}" for y in range(n + 1) for x in range(n + 1))
    counts = " ".join(["4"] * (n * n))
    idx = " ".join(f"{y * (n + 1) + x} {y * (n + 1) + x + 1} {(y + 1) * (n + 1) + x + 1} {(y + 1) * (n + 1) + x}"
                   for y in range(n) for x in range(n))
    s = api.Scene.from_psy_string(text % (verts, counts, idx))
    d = s.desc.contents
    assert d.n_tris == 2 * n * n
    rays = np.zeros(300, dtype=abi.RAY_DTYPE)
    rays["orig"] = np.c_[rng.uniform(-2.5, 2.5, 300), rng.uniform(-2.5, 2.5, 300), rng.uniform(1.5, 4, 300)]
    dirs = np.c_[rng.normal(0, 0.5, 300), rng.normal(0, 0.5, 300), -np.ones(300)]
    rays["dir"] = dirs / np.linalg.norm(dirs, axis=1, keepdims=True)
    rays["max_t"] = np.inf
    hits, _, _ = O.trace_rays(s.desc, rays)
    lib = O.load()
    out = (C.c_float * 4)()
    tri = np.array([d.tris[i] for i in range(4 * d.n_tris)], dtype=np.uint32).reshape(-1, 4)
    vs = np.array([d.verts[i] for i in range(3 * d.n_vert_keys)], dtype=np.float32).reshape(-1, 3)
    f3 = lambda a: (C.c_float * len(a))(*[float(x) for x in a])
    for r, h in zip(rays, hits):
        best_t, best = np.inf, None
        for k in range(len(tri)):
            t9 = np.concatenate([vs[tri[k, 0]], vs[tri[k, 1]], vs[tri[k, 2]]])
            if lib.oracle_tri_intersect(f3(r["orig"]), f3(r["dir"]), float(best_t), f3(t9), out):
                best_t, best = out[0], int(tri[k, 3])
        if best is None:
            assert h["kind"] == abi.HIT_MISS
        else:
            assert h["kind"] == abi.HIT_TRIANGLE and h["t"] == np.float32(best_t)


def test_bucket_list_matches_reference_scheduler():
    """Renderer::render's Hilbert bucket list (renderer.rs:138-180): covers the crop exactly once."""
    h = abi.load_host()
    cap = 4096
    buf = (C.c_uint32 * (4 * cap))()
    for (w, hh, sx, sy, bw) in [(512, 512, 0, 0, 16), (100, 37, 5, 9, 8), (7, 3, 0, 0, 16), (1920, 1080, 0, 0, 64)]:
        n = h.psyh_bucket_list(w, hh, sx, sy, bw, bw, buf, cap)
        assert 0 < n <= cap
        cover = np.zeros((hh, w), dtype=np.int32)
        for i in range(n):
            x, y, bw_, bh_ = buf[4 * i:4 * i + 4]
            cover[y - sy:y - sy + bh_, x - sx:x - sx + bw_] += 1
        assert (cover == 1).all()
    # bucket_dim = floor(sqrt(spb / spp)) and the Hilbert order starts at the origin
    n = h.psyh_bucket_list(512, 512, 0, 0, 16, 16, buf, cap)
    assert n == 32 * 32 and list(buf[0:4]) == [0, 0, 16, 16] and list(buf[4:8]) == [16, 0, 16, 16]


def test_oracle_edge_cases():
    s = api.Scene.from_psy_string(cornell_text())
    hits, _, _ = O.trace_rays(s.desc, np.zeros(0, dtype=abi.RAY_DTYPE))
    assert len(hits) == 0
    rays = np.zeros(3, dtype=abi.RAY_DTYPE)
    rays["orig"] = (-2.78, -8.0, 2.73)
    rays["dir"] = (0, 1, 0)
    rays["max_t"] = (np.inf, 0.0, np.inf)
    rays["flags"] = (0, 0, 1)
    hits, geo, cnt = O.trace_rays(s.desc, rays, want_geo=True)
    assert hits["kind"][0] == abi.HIT_TRIANGLE and hits["t"][0] > 8.0
    assert hits["kind"][1] == abi.HIT_MISS
    assert hits["kind"][2] == abi.HIT_OCCLUDED
    assert cnt[0] > 0 and cnt[1] > 0
