"""Pins on what the `.psy` loader MEANS, with expectations computed by hand from the
reference's parse code — the loader is shared by the oracle and the CUDA path, so a wrong convention here would
cancel out in every GPU-vs-oracle comparison.

  parse_matrix      src/parse/psy.rs:556-572     16 numbers, column-major, 4th row dropped
  parse_color       src/parse/psy.rs:582-615     rec709 -> rec709_e_to_xyz (sub_crates/color/build.rs:130-161)
  parse_camera      src/parse/psy.rs:353-453, src/camera.rs:22-72   degrees -> radians (f32), tfov, DoF defaults
  mesh faces        src/parse/psy_mesh_surface.rs:96-115 (fan), src/surface/triangle_mesh.rs:80 ((v0, v2, v1) winding)
  instance xforms   psychoblend/assembly.py:317  the file holds world -> local
"""
import numpy as np

import oracle_lib as O
from psychopath_b200 import abi, api

F = np.float32

HEAD = """Scene $s { Output { Path ["out/img.exr"] } RenderSettings { Resolution [40 30] SamplesPerPixel [3] Seed [9] DicingRate [2.5] }
 Camera { %s }
 World { BackgroundShader { Type [Color] Color [%s] } }
 Assembly {
  SurfaceShader $red { Type [Lambert] Color [rec709, 1 0 0] }
  SurfaceShader $mix { Type [GGX] Color [rec709, 0.25 0.5 0.75] Roughness [0.3] Fresnel [0.6] }
  SurfaceShader $hot { Type [Emit] Color [blackbody, 3200 0.5] }
  SurfaceShader $warm { Type [Emit] Color [color_temperature, 4500 2.0] }
  %s
 } }"""
CAM = "Fov [90] Transform [1 2 3 0  4 5 6 0  7 8 9 0  10 11 12 1]"
QUAD = """MeshSurface $q { Vertices [-1 -1 0  1 -1 0  1 1 0  -1 1 0  0 2 0] FaceVertCounts [4 3 5] FaceVertIndices [0 1 2 3  0 1 4  0 1 2 3 4]
             Normals [0 0 2  0 3 0  4 0 0  1 1 0  0 0 1] }
  Instance { Data [$q] SurfaceShaderBind [$red] Transform [1 0 0 0  0 1 0 0  0 0 1 0  -5 0 0 1] }"""


def load(cam=CAM, bg="rec709, 0 0 0", body=QUAD):
    return api.Scene.from_psy_string(HEAD % (cam, bg, body))


def xf(x):
    return np.array([[x.r[r][c] for c in range(4)] for r in range(3)], dtype=np.float32)


def test_render_settings_and_output():
    s = load()
    d = s.desc.contents
    assert (d.res_x, d.res_y, d.spp, d.seed) == (40, 30, 3, 9)
    assert s.output_file == "out/img.exr"


def test_matrix_is_column_major_with_the_fourth_row_dropped():
    s = load()
    d = s.desc.contents
    # file order: columns (1 2 3 0) (4 5 6 0) (7 8 9 0) (10 11 12 1)  ->  rows (1 4 7 | 10) (2 5 8 | 11) (3 6 9 | 12)
    want = np.array([[1, 4, 7, 10], [2, 5, 8, 11], [3, 6, 9, 12]], dtype=np.float32)
    assert np.array_equal(xf(d.camera.xforms[0]), want)
    # a point through it: M p + T (math3d/src/point.rs:132-141)
    out = (abi.f32 * 3)()
    O.load().oracle_xform_point(d.camera.xforms[0], (abi.f32 * 3)(1, 1, 1), out)
    assert list(out) == [22.0, 26.0, 30.0]


def test_camera_fov_and_depth_of_field_defaults():
    s = load()
    d = s.desc.contents
    half = F(F(90) * F(np.pi / 180.0)) / F(2)
    assert d.camera.n_tfovs == 1 and abs(d.camera.tfovs[0] - F(np.sin(half) / np.cos(half))) <= 2e-7
    # no aperture / focus given -> (0, 1)  (camera.rs:33-57)
    assert (d.camera.n_aperture_radii, d.camera.aperture_radii[0]) == (1, 0.0)
    assert (d.camera.n_focus_distances, d.camera.focus_distances[0]) == (1, 1.0)
    # aperture without focus distance, and a zero focus distance, both disable focal blur
    for cam in (CAM + " ApertureRadius [0.3]", CAM + " ApertureRadius [0.3] FocalDistance [0]"):
        sc = load(cam=cam)
        c = sc.desc.contents.camera
        assert (c.aperture_radii[0], c.focus_distances[0]) == (0.0, 1.0)
    sc = load(cam=CAM + " ApertureRadius [0.3] FocalDistance [7.5] Fov [60]")
    c = sc.desc.contents.camera
    assert (c.aperture_radii[0], c.focus_distances[0], c.n_tfovs) == (F(0.3), 7.5, 2)


def test_colours():
    s = load(bg="rec709, 0 1 0")
    d = s.desc.contents
    m = np.array([[0.4969210174, 0.3390896921, 0.1639892905], [0.2562248996, 0.6781793842, 0.0655957162],
                  [0.0232931727, 0.1130298974, 0.8636769299]], dtype=np.float32)  # rec709_e_to_xyz, SURVEY.md 10
    sh = {i: d.shaders[i] for i in range(d.n_shaders)}
    red = sh[0].color
    assert (red.kind, red.a, red.b, red.c) == (0, m[0, 0], m[1, 0], m[2, 0])
    mix = sh[1]
    c = np.array([0.25, 0.5, 0.75], dtype=np.float32)
    want = [F(F(F(c[0] * m[r, 0]) + F(c[1] * m[r, 1])) + F(c[2] * m[r, 2])) for r in range(3)]  # build.rs:78-85
    assert [mix.color.a, mix.color.b, mix.color.c] == want
    assert (mix.kind, mix.roughness, mix.fresnel) == (2, F(0.3), F(0.6))  # PSY_SHADER_GGX
    assert (sh[2].color.kind, sh[2].color.a, sh[2].color.b) == (1, 3200.0, 0.5)  # blackbody: temperature, factor
    assert (sh[3].color.kind, sh[3].color.a, sh[3].color.b) == (2, 4500.0, 2.0)  # color_temperature
    assert (d.background.a, d.background.b, d.background.c) == (m[0, 1], m[1, 1], m[2, 1])


def test_faces_are_fan_triangulated_with_swapped_winding():
    s = load()
    d = s.desc.contents
    assert d.n_tris == 2 + 1 + 3
    tris = sorted(tuple(d.tris[4 * i + k] for k in (3, 0, 1, 2)) for i in range(d.n_tris))  # by original index
    # fan: quad (0 1 2)(0 2 3), triangle (0 1 4), pentagon (0 1 2)(0 2 3)(0 3 4); stored as (v0, v2, v1)
    assert [t[1:] for t in tris] == [(0, 2, 1), (0, 3, 2), (0, 4, 1), (0, 2, 1), (0, 3, 2), (0, 4, 3)]
    assert [t[0] for t in tris] == list(range(6))
    # normals are normalised at parse (psy_mesh_surface.rs:64)
    n = np.array([d.normals[i] for i in range(15)], dtype=np.float32).reshape(5, 3)
    assert np.allclose(n[:3], np.eye(3)[[2, 1, 0]], atol=1e-7) and np.allclose(n[3], [2 ** -0.5, 2 ** -0.5, 0], atol=1e-7)


def test_instance_transform_is_world_to_local():
    """`Transform [.. -5 0 0 1]` is world -> local: the object sits at x = +5 in the world."""
    s = load()
    d = s.desc.contents
    assert np.array_equal(xf(d.xforms[d.instances[0].xform_base])[:, 3], [-5, 0, 0])
    rays = np.zeros(2, dtype=abi.RAY_DTYPE)
    rays["orig"] = [(5.2, 0.3, 4.0), (-4.8, 0.3, 4.0)]
    rays["dir"] = (0, 0, -1)
    rays["max_t"] = np.inf
    hits, geo, _ = O.trace_rays(s.desc, rays, want_geo=True)
    assert hits["kind"][0] == abi.HIT_TRIANGLE and hits["t"][0] == 4.0 and hits["kind"][1] == abi.HIT_MISS
    assert np.allclose(geo["pos"][0], (5.2, 0.3, 0.0), atol=1e-6)
    # (v0, v2, v1) winding: nor_g = cross(p0 - p1, p0 - p2) of the STORED order points to -z for a
    # counter-clockwise face seen from +z (triangle_mesh.rs:271)
    assert geo["nor_g"][0][2] < 0 and abs(geo["nor_g"][0][0]) < 1e-6


def test_old_and_new_layout_give_the_same_scene():
    """the example scenes' layout (Shaders / Objects blocks, binding on the object) and the
    PsychoBlend layout (everything inside Assembly, binding on the instance) describe the same scene."""
    new = """Scene $s { Output { Path ["a.png"] } RenderSettings { Resolution [8 8] SamplesPerPixel [1] Seed [1] }
 Camera { Fov [40] Transform [1 0 0 0 0 1 0 0 0 0 1 0 0 0 0 1] }
 World { BackgroundShader { Type [Color] Color [rec709, 0 0 0] } }
 Shaders { SurfaceShader $red { Type [Lambert] Color [rec709, 1 0 0] } }
 Objects { MeshSurface $q { SurfaceShaderBind [$red] Vertices [-1 -1 0  1 -1 0  1 1 0  -1 1 0] FaceVertCounts [4] FaceVertIndices [0 1 2 3] } }
 Assembly { Instance { Data [$q] Transform [1 0 0 0  0 1 0 0  0 0 1 0  0 0 3 1] } } }"""
    old = """Scene $s { Output { Path ["a.png"] } RenderSettings { Resolution [8 8] SamplesPerPixel [1] Seed [1] }
 Camera { Fov [40] Transform [1 0 0 0 0 1 0 0 0 0 1 0 0 0 0 1] }
 World { BackgroundShader { Type [Color] Color [rec709, 0 0 0] } }
 Assembly { SurfaceShader $red { Type [Lambert] Color [rec709, 1 0 0] }
  MeshSurface $q { Vertices [-1 -1 0  1 -1 0  1 1 0  -1 1 0] FaceVertCounts [4] FaceVertIndices [0 1 2 3] }
  Instance { Data [$q] SurfaceShaderBind [$red] Transform [1 0 0 0  0 1 0 0  0 0 1 0  0 0 3 1] } } }"""
    sa, sb = api.Scene.from_psy_string(new), api.Scene.from_psy_string(old)
    a, b = sa.desc.contents, sb.desc.contents
    assert (a.n_tris, a.n_instances, a.n_shaders) == (b.n_tris, b.n_instances, b.n_shaders) == (2, 1, 1)
    assert a.instances[0].shader_index == b.instances[0].shader_index == 0
    assert [a.tris[i] for i in range(8)] == [b.tris[i] for i in range(8)]
    assert np.array_equal(xf(a.xforms[0]), xf(b.xforms[0]))
