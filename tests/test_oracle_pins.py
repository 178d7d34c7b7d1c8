"""Pins the CPU oracle against every known-answer test the reference's own unit tests hold for the hot
path and against the self-derived vectors of SURVEY.md 10.  CPU only.

Reference tests restated here (file:line in /root/reference):
  sub_crates/math3d/src/transform.rs:153-177   multiply_test, inverse_test
  sub_crates/math3d/src/point.rs:166-201       mul_matrix_1, mul_matrix_3
  sub_crates/math3d/src/vector.rs:190-290      add, sub, mul_scalar, mul_matrix_1, div, length, length2, normalized, dot, cross
  sub_crates/math3d/src/normal.rs:209-220      mul_matrix_1 (<= 4 ulp)
  src/lerp.rs:140-322                          lerp1-3, lerp_slice1-6, lerp_matrix, lerp_point
  src/fp_utils.rs:84-131                       inc/dec ulp tests
  src/hilbert.rs:65-77                         reversible
"""
import ctypes as C
import hashlib

import numpy as np

import oracle_lib as O
from psychopath_b200 import abi


def xf(vals):
    """Transform::new_from_values(a..l): rows (a b c d / e f g h / i j k l)."""
    m = abi.PsyXform()
    for i in range(3):
        for j in range(4):
            m.r[i][j] = vals[4 * i + j]
    return m


def xf_vals(m):
    return [m.r[i][j] for i in range(3) for j in range(4)]


A = [1.0, 2.0, 2.0, 1.5, 3.0, 6.0, 7.0, 8.0, 9.0, 2.0, 11.0, 12.0]


def test_transform_multiply_exact():
    lib = O.load()
    b = [1.0, 5.0, 9.0, 13.0, 2.0, 6.0, 10.0, 14.0, 3.0, 7.0, 11.0, 15.0]
    c = [97.0, 50.0, 136.0, 162.5, 110.0, 60.0, 156.0, 185.0, 123.0, 70.0, 176.0, 207.5]
    out = abi.PsyXform()
    lib.oracle_xform_mul(C.byref(xf(A)), C.byref(xf(b)), C.byref(out))
    assert xf_vals(out) == c


def test_transform_inverse():
    lib = O.load()
    a = xf([1.0, 0.33, 0.0, -2.0, 0.0, 1.0, 0.0, 0.0, 2.1, 0.7, 1.3, 0.0])
    inv = abi.PsyXform()
    lib.oracle_xform_inverse(C.byref(a), C.byref(inv))
    prod = abi.PsyXform()
    lib.oracle_xform_mul(C.byref(a), C.byref(inv), C.byref(prod))
    ident = [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0]
    assert np.allclose(xf_vals(prod), ident, atol=1e-7)


def _apply(fn, m, v):
    out = (C.c_float * 3)()
    fn(C.byref(m), (C.c_float * 3)(*v), out)
    return list(out)


def test_point_vector_normal_times_matrix():
    lib = O.load()
    assert _apply(lib.oracle_xform_point, xf(A), (1.0, 2.5, 4.0)) == [15.5, 54.0, 70.0]
    assert _apply(lib.oracle_xform_vector, xf(A), (1.0, 2.5, 4.0)) == [14.0, 46.0, 58.0]
    n = np.array(_apply(lib.oracle_xform_normal, xf(A), (1.0, 2.5, 4.0)), dtype=np.float32)
    want = np.array([-4.0625, 1.78125, -0.03125], dtype=np.float32)
    ulps = np.abs(n.view(np.int32) - want.view(np.int32))
    assert ulps.max() <= 4


def test_vector_arithmetic_known_answers():
    """sub_crates/math3d/src/vector.rs:190-290 (add, sub, mul_scalar, div, length, length2, normalized, dot_test,
    cross_test); normal.rs and point.rs hold the same cases for their types."""
    lib = O.load()
    lib.oracle_vec3_scalar.restype = C.c_float

    def op(k, a, b=None):
        out = (C.c_float * 3)()
        lib.oracle_vec3_op(k, (C.c_float * 3)(*a), (C.c_float * 3)(*b) if b else None, out)
        return list(out)

    def scalar(k, a, b=None):
        return lib.oracle_vec3_scalar(k, (C.c_float * 3)(*a), (C.c_float * 3)(*b) if b else None)

    v1, v2 = (1.0, 2.0, 3.0), (1.5, 4.5, 2.5)
    assert op(0, v1, v2) == [2.5, 6.5, 5.5]
    assert op(1, v1, v2) == [-0.5, -2.5, 0.5]
    assert op(2, v1, (2.0, 0, 0)) == [2.0, 4.0, 6.0]
    assert op(3, v1, (2.0, 0, 0)) == [0.5, 1.0, 1.5]
    assert abs(scalar(1, v1) - 3.7416573867739413) < 0.000001
    assert scalar(2, v1) == 14.0
    n = op(5, v1)
    for got, want in zip(n, (0.2672612419124244, 0.5345224838248488, 0.8017837257372732)):
        assert abs(got - want) < 0.000001
    assert scalar(0, v1, v2) == 18.0
    assert op(4, (1.0, 0.0, 0.0), (0.0, 1.0, 0.0)) == [0.0, 0.0, 1.0]


def test_point_matrix_composition():
    lib = O.load()
    m1, m2 = xf(A), xf([4.0, 1.0, 2.0, 3.5, 3.0, 6.0, 5.0, 2.0, 2.0, 2.0, 4.0, 12.0])
    m12 = abi.PsyXform()
    lib.oracle_xform_mul(C.byref(m1), C.byref(m2), C.byref(m12))
    p = (1.0, 2.5, 4.0)
    a = np.array(_apply(lib.oracle_xform_point, m12, p))
    b = np.array(_apply(lib.oracle_xform_point, m2, _apply(lib.oracle_xform_point, m1, p)))
    assert ((a - b) ** 2).sum() <= 0.00001


def test_lerp_matrix_exact_quarters():
    lib = O.load()
    a = [0.0, 2.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0, 8.0, 9.0, 10.0, 11.0]
    b = [-1.0, 1.0, 3.0, 4.0, 5.0, 6.0, 7.0, 8.0, 9.0, 10.0, 11.0, 12.0]
    want = {0.0: a, 1.0: b,
            0.25: [-0.25, 1.75, 2.25, 3.25, 4.25, 5.25, 6.25, 7.25, 8.25, 9.25, 10.25, 11.25],
            0.5: [-0.5, 1.5, 2.5, 3.5, 4.5, 5.5, 6.5, 7.5, 8.5, 9.5, 10.5, 11.5],
            0.75: [-0.75, 1.25, 2.75, 3.75, 4.75, 5.75, 6.75, 7.75, 8.75, 9.75, 10.75, 11.75]}
    for alpha, w in want.items():
        out = abi.PsyXform()
        lib.oracle_xform_lerp(C.byref(xf(a)), C.byref(xf(b)), alpha, C.byref(out))
        assert xf_vals(out) == w


def test_lerp_slice_known_answers():
    lib = O.load()
    s = (C.c_float * 5)(0.0, 1.0, 2.0, 3.0, 4.0)
    for alpha, want in ((0.0, 0.0), (1.0, 4.0), (0.5, 2.0), (0.25, 1.0), (0.75, 3.0), (0.625, 2.5)):
        assert lib.oracle_lerp_slice_f32(s, 5, alpha) == want
    two = (C.c_float * 2)(1.0, 2.0)
    assert lib.oracle_lerp_slice_f32(two, 2, 0.0) == 1.0
    assert lib.oracle_lerp_slice_f32(two, 2, 1.0) == 2.0
    assert lib.oracle_lerp_slice_f32(two, 2, 0.5) == 1.5
    # lerp_point_1..3, lerp_normal_1..3, lerp_vector_1..3 (lerp.rs:243-320): (1, 2, 1) -> (-2, 1, -1), component-wise
    for a, b, mid in ((1.0, -2.0, -0.5), (2.0, 1.0, 1.5), (1.0, -1.0, 0.0)):
        pair = (C.c_float * 2)(a, b)
        assert lib.oracle_lerp_slice_f32(pair, 2, 0.0) == a
        assert lib.oracle_lerp_slice_f32(pair, 2, 1.0) == b
        assert lib.oracle_lerp_slice_f32(pair, 2, 0.5) == mid


def test_ulp_helpers():
    lib = O.load()
    inc, dec = lib.oracle_increment_ulp, lib.oracle_decrement_ulp
    assert inc(1.0) > 1.0 and inc(-1.0) > -1.0
    assert dec(1.0) < 1.0 and dec(-1.0) < -1.0
    assert inc(0.0) > 0.0 and inc(-0.0) > 0.0
    assert dec(0.0) < 0.0 and dec(-0.0) < 0.0
    for v in (1.0, -1.0, 1.2, -1.2):
        v32 = float(np.float32(v))
        assert dec(inc(v32)) == v32
        assert inc(dec(v32)) == v32


def test_hilbert_known_and_reversible():
    lib = O.load()
    want = [(0, 0), (1, 0), (1, 1), (0, 1), (0, 2), (0, 3), (1, 3), (1, 2), (2, 2), (2, 3), (3, 3), (3, 2), (3, 1),
            (2, 1), (2, 0), (3, 0)]
    x, y = C.c_uint32(), C.c_uint32()
    got = []
    for d in range(16):
        lib.oracle_hilbert_d2xy(d, C.byref(x), C.byref(y))
        got.append((x.value, y.value))
    assert got == want
    # reversible (hilbert.rs:65-77): d2xy is a bijection on a 64x64 block
    seen = set()
    for d in range(64 * 64):
        lib.oracle_hilbert_d2xy(d, C.byref(x), C.byref(y))
        assert x.value < 64 and y.value < 64
        seen.add((x.value, y.value))
    assert len(seen) == 64 * 64


def test_hash_and_golden_ratio_vectors():
    lib = O.load()
    assert lib.oracle_hash_u32(0, 0) == 0x00000000
    assert lib.oracle_hash_u32(1, 0) == 0xE7DE5BAE
    assert lib.oracle_hash_u32(0x12345678, 1) == 0xD624D67C
    h = lib.oracle_hash_u32((5 << 16) ^ 7, 1)
    assert h == 0xC5383481
    assert np.float32(lib.oracle_golden_ratio_sample(0, h)) == np.float32(0.07516894)
    assert np.float32(lib.oracle_golden_ratio_sample(3, h)) == np.float32(0.9292709)


def test_traversal_table_sha256():
    lib = O.load()
    buf = (C.c_uint8 * 384)()
    lib.oracle_traversal_table(buf)
    t = bytes(buf)
    assert hashlib.sha256(t).hexdigest() == "cdb2d257f878635ce8d1ddbbd46bb2f4594db6459d2857a7d30bbe692fe04a5d"
    assert list(t[0:27]) == [228] * 27 and list(t[27:45]) == [36] * 18 and list(t[45:48]) == [4] * 3
    assert list(t[48:52]) == [27, 177, 177, 75]
    assert list(t[7 * 48:7 * 48 + 27]) == [27] * 27


def test_fp_gamma_bits():
    lib = O.load()
    for n, bits in ((2, 0x34000001), (3, 0x34400002), (5, 0x34A00003), (7, 0x34E00006)):
        assert np.float32(lib.oracle_fp_gamma(n)).view(np.uint32) == bits


def test_probit_vectors():
    lib = O.load()
    w = float(np.float32(1.0) / np.float32(3.0))
    assert lib.oracle_probit(0.5, w) == 0.0
    for n, want in ((0.25, -0.22482991), (0.75, 0.22482991), (0.01, -0.77544928), (0.99, 0.77544945)):
        got = np.float32(lib.oracle_probit(n, w))
        assert abs(int(got.view(np.int32)) - int(np.float32(want).view(np.int32))) <= 1, (n, got)


def test_sampler_is_a_stratified_sequence():
    """sobol_burley is not vendored (parity with upstream unpinned): check the structural properties the
    restatement must have — values in [0,1), determinism, and (0,2)-net stratification of each Owen-
    scrambled Sobol dimension over the first 2^k samples."""
    lib = O.load()
    out = (C.c_float * 4)()
    pts = np.zeros((256, 8), dtype=np.float32)
    for i in range(256):
        for ds in range(2):
            lib.oracle_get_sample_4d(i, ds, 17, 33, 12345, out)
            pts[i, 4 * ds:4 * ds + 4] = list(out)
    assert (pts >= 0).all() and (pts < 1).all()
    for k in (4, 6, 8):
        n = 1 << k
        for d in range(8):
            cells = np.floor(pts[:n, d] * n).astype(int)
            assert len(set(cells.tolist())) == n, (k, d)
    lib.oracle_get_sample_4d(5, 0, 17, 33, 12345, out)
    assert list(out) == [float(v) for v in pts[5, :4]]
    lib.oracle_get_sample_4d(5, 0, 18, 33, 12345, out)
    assert list(out) != [float(v) for v in pts[5, :4]]
    # dimension sets >= 2 fall back to the hash (renderer.rs:714-723)
    lib.oracle_get_sample_4d(5, 3, 17, 33, 12345, out)
    assert all(0.0 <= v <= 1.0 for v in out)


def test_triangle_known_answers():
    """Woop test on hand-checkable cases (the reference has no triangle tests: analytic pins)."""
    lib = O.load()
    f3 = lambda *v: (C.c_float * len(v))(*v)
    out = (C.c_float * 4)()
    tri = f3(0, 0, 5, 1, 0, 5, 0, 1, 5)
    assert lib.oracle_tri_intersect(f3(0.25, 0.25, 0), f3(0, 0, 1), float("inf"), tri, out) == 1
    assert out[0] == 5.0 and abs(out[1] - 0.5) < 1e-7 and abs(out[2] - 0.25) < 1e-7 and abs(out[3] - 0.25) < 1e-7
    # beyond max_t, behind the origin, outside the triangle, parallel
    assert lib.oracle_tri_intersect(f3(0.25, 0.25, 0), f3(0, 0, 1), 4.0, tri, out) == 0
    assert lib.oracle_tri_intersect(f3(0.25, 0.25, 6), f3(0, 0, 1), float("inf"), tri, out) == 0
    assert lib.oracle_tri_intersect(f3(0.9, 0.9, 0), f3(0, 0, 1), float("inf"), tri, out) == 0
    assert lib.oracle_tri_intersect(f3(0.25, 0.25, 0), f3(1, 0, 0), float("inf"), tri, out) == 0
    # watertight: a ray through the shared edge of two triangles hits at least one of them
    t1 = f3(0, 0, 5, 1, 0, 5, 0, 1, 5)
    t2 = f3(1, 0, 5, 1, 1, 5, 0, 1, 5)
    hits = 0
    for s in np.linspace(0.0, 1.0, 257):
        o = f3(float(1 - s), float(s), 0)
        hits += (lib.oracle_tri_intersect(o, f3(0, 0, 1), float("inf"), t1, out) |
                 lib.oracle_tri_intersect(o, f3(0, 0, 1), float("inf"), t2, out))
    assert hits == 257


def test_spectral_roundtrip_and_white_furnace():
    """Meng upsampling: XYZ -> 4 hero wavelengths -> XYZ integrates back to roughly the input colour, and an
    equal-energy white of luminance 1 upsamples to a flat spectrum of 1 (EQUAL_ENERGY_REFLECTANCE)."""
    lib = O.load()
    col = abi.PsyColor(0, 1.0, 1.0, 1.0)
    e = (C.c_float * 4)()
    for hero in (380.0, 455.5, 612.25, 700.0):
        lib.oracle_color_to_spectral(C.byref(col), hero, e)
        assert np.allclose(list(e), 1.0, atol=2e-3), list(e)
    # Monte-Carlo round trip of a saturated-ish colour
    col = abi.PsyColor(0, 0.35, 0.45, 0.2)
    acc = np.zeros(3)
    n = 4096
    xyz = (C.c_float * 3)()
    for i in range(n):
        hero = 380.0 + 320.0 * ((i + 0.5) / n)
        lib.oracle_color_to_spectral(C.byref(col), hero, e)
        lib.oracle_xyz_from_spectral(e, hero, xyz)
        acc += np.array(list(xyz))
    acc /= n
    assert np.allclose(acc, [0.35, 0.45, 0.2], rtol=0.08, atol=0.01), acc


def test_sphere_lambert_boundary_is_where_the_references_nan_comes_from():
    """The two non-finite pixels of round 1's 4K frame (BENCH_r01 `e2e.non_finite_pixels: 2`), root-caused.

    `sphere_lambert` (surface_closure.rs:351-377) evaluates `ysin = rcos / nlsin; ycos = (1 - ysin^2).sqrt()` and the
    g / h terms BEFORE it picks a case from `nl = acos(nlcos)`, `r = acos(rcos)`.  On the boundary between the fully
    visible and the partially occluded case (nl = pi/2 - r) rounding makes ysin > 1, so g and h are NaN; whether the
    NaN is returned hangs on the last bit of the two acos results.  LightTree::select (light_tree.rs:146-170) adds
    the children's estimates without a guard, and neither the integrator nor the film (renderer.rs:299-310) has
    one, so the NaN reaches the pixel.  Inputs below: pixel (1087, 1757), sample 161 of that frame (profiles/
    r2_find_nan.txt): glibc keeps the child on the finite side by one ulp, CUDA's acosf does not — the product build
    shows a NaN there, the oracle does not, and with the shared deterministic libm (strict mode) both agree."""
    lib = O.load()
    out = (C.c_float * 4)()
    v = lib.oracle_sphere_lambert(0.563733935, 0.825956464, out)
    g, h, nl, r = list(out)
    assert np.isnan(g) and np.isnan(h)                      # the terms of the partially occluded case are NaN here ...
    margin = np.float32(nl) - (np.float32(np.pi / 2) - np.float32(r))
    assert abs(margin) <= 1.2e-7                            # ... one ulp from the case boundary
    assert np.isfinite(v) == bool(margin < 0)               # and the result is finite only on the near side of it
    # inputs on the same boundary where glibc's acos lands on the far side: the reference's formula returns NaN
    # (this is what pixel (1566, 495) sample 210 of that frame hit on the CPU and the GPU alike)
    for nlcos, rcos in [(0.42701658606529236, 0.9042438268661499), (0.46583878993988037, 0.8848696351051331),
                        (0.8665500283241272, 0.4990902245044708)]:
        assert np.isnan(lib.oracle_sphere_lambert(nlcos, rcos, out))
        assert abs(np.float32(out[2]) - (np.float32(np.pi / 2) - np.float32(out[3]))) <= 2.4e-7


def test_white_furnace_through_the_integrator():
    """An analytic pin of the path the reference has no test for: a convex white Lambert box in a uniform background,
    no lights.  A path that hits the box escapes with its first bounce, and with cosine-weighted sampling the bounce
    weight is exactly the albedo, so EVERY sample must carry the background's radiance — whatever the box does (motion
    blur, depth of field).  This ties together camera rays, traversal, robust ray origins (a self-hit would go dark),
    Lambert sample / pdf (surface_closure.rs:284-332), the miss shader and the Meng upsampling of white.  With albedo
    0.5 the box's samples are exactly half; GGX creates no energy."""
    from psychopath_b200 import api
    from scenes import furnace_text

    def samples(shader, **kw):
        s = api.Scene.from_psy_string(furnace_text(shader, **kw))
        d = s.desc.contents
        _, samp, st = O.render_bucket(s.desc, 0, 0, d.res_x, d.res_y, d.spp, d.seed, want_samples=True)
        hit_share = st.ray_count / (d.res_x * d.res_y * d.spp) - 1.0
        return samp.reshape(-1, 4), hit_share

    bg = 0.5
    for kw in ({}, {"motion": True, "aperture": 0.2}):
        samp, hit_share = samples("Type [Lambert] Color [rec709, 1 1 1]", **kw)
        assert 0.15 < hit_share < 0.3                       # a fifth of the paths do go through the box
        assert np.abs(samp - bg).max() < 3e-4, (samp.min(), samp.max())
    samp, hit_share = samples("Type [Lambert] Color [rec709, 0.5 0.5 0.5]")
    half = np.abs(samp - 0.5 * bg).max(axis=1) < 3e-4
    full = np.abs(samp - bg).max(axis=1) < 3e-4
    assert (half | full).all() and abs(half.mean() - hit_share) < 1e-6
    samp, _ = samples("Type [GGX] Color [rec709, 1 1 1] Roughness [0.4] Fresnel [1.0]")
    assert samp.mean() < bg and samp.max() < bg * 1.5


def _lit_floor(**kw):
    from psychopath_b200 import api
    from scenes import lit_floor_text
    s = api.Scene.from_psy_string(lit_floor_text(**kw))
    d = s.desc.contents
    rays = O.camera_rays(s.desc, 0, 0, d.res_x, d.res_y, d.spp, d.seed)
    hits, geo, _ = O.trace_rays(s.desc, rays, want_geo=True)
    _, samp, _ = O.render_bucket(s.desc, 0, 0, d.res_x, d.res_y, d.spp, d.seed, want_samples=True)
    on = hits["kind"] == abi.HIT_TRIANGLE
    assert on.all()  # the floor fills the view
    return geo["pos"].astype(np.float64), samp.reshape(-1, 4), d.spp


LIGHT_AT = np.array([0.5, 1.0, 3.0])
LIGHT_INSTANCE = "Instance {{ Data [${}] Transform [1 0 0 0  0 1 0 0  0 0 1 0  -0.5 -1.0 -3.0 1] }}"  # world -> local


def test_direct_light_from_a_distant_disk_is_the_analytic_value():
    """Direct lighting has a closed form on a Lambert floor under a black sky, and nothing in the reference tests it:
    DistantDiskLight (distant_disk_light.rs:59-85) delivers irradiance `colour` at normal incidence whatever its radius
    (radiance colour / solid angle, pdf 1 / solid angle), so every sample is albedo / pi * colour * cos(theta).  Pins the
    world-light selection, the cone sampling and its pdf, Lambert evaluation and the MIS weight's limit."""
    _, samp, _ = _lit_floor(world_light="DistantDiskLight $sun { Direction [-0.3 0.4 -0.8] Color [rec709, 2 2 2] Radius [0.001] }")
    want = 2.0 * 0.6 / np.pi * 0.8 / np.sqrt(0.09 + 0.16 + 0.64)
    assert np.abs(samp / want - 1.0).max() < 2e-3, (samp.min(), samp.max(), want)


def test_direct_light_from_a_rectangle_is_the_analytic_value():
    """RectangleLight (rectangle_light.rs:121-243): radiance colour / area * 0.5 (the power is shared by the two
    sides, :191), sampled uniformly over the solid angle of its two spherical triangles.  The pixel means must
    integrate to albedo / pi * E with E the irradiance of that rectangle, here by quadrature over its area."""
    for w, h in ((0.4, 0.3), (2.0, 1.0)):
        k = 30.0
        pos, samp, spp = _lit_floor(objects=f"RectangleLight $panel {{ Color [rec709, {k} {k} {k}] Dimensions [{w} {h}] }}",
                                    instances=LIGHT_INSTANCE.format("panel"))
        g = (np.arange(24) + 0.5) / 24 - 0.5
        qx, qy = np.meshgrid(g * w, g * h)
        q = np.stack([LIGHT_AT[0] + qx.ravel(), LIGHT_AT[1] + qy.ravel(), np.full(qx.size, LIGHT_AT[2])], 1)
        radiance = k / (w * h) * 0.5
        e = np.zeros(len(pos))
        for a in range(0, len(pos), 2048):
            v = q[None, :, :] - pos[a:a + 2048, None, :]
            d2 = (v ** 2).sum(2)
            cos = v[:, :, 2] / np.sqrt(d2)     # the same at the floor and at the light: both are horizontal
            e[a:a + 2048] = (radiance * cos * cos / d2).mean(1) * (w * h)
        want = 0.6 / np.pi * e
        got = samp[:, 0]
        assert (got > 0).all()
        assert abs(got.mean() / want.mean() - 1.0) < 3e-3
        per_pixel = got.reshape(-1, spp).mean(1) / want.reshape(-1, spp).mean(1)
        assert np.abs(per_pixel - 1.0).max() < 0.03


def test_direct_light_from_a_sphere_and_the_references_skewed_basis():
    """SphereLight (sphere_light.rs:85-192): radiance colour / (4 pi r^2), cone sampling, so a small sphere at distance d
    gives E = colour * cos / (4 d^2) and the samples that reach it are exact.  But a seventh of them do not: the
    reference hands the UN-normalised vector to the light's centre to coordinate_system_from_vector (:106-111), whose
    construction (math.rs:26-34, Duff et al.) assumes unit length; the tangent axes come out skewed, grazing cone
    samples land on the far side of the sphere, and their shadow rays are stopped by the light itself.  The oracle and
    the CUDA path reproduce this: it is the reference's arithmetic, and parity is judged against it."""
    k, r = 40.0, 0.05
    pos, samp, _ = _lit_floor(objects=f"SphereLight $bulb {{ Color [rec709, {k} {k} {k}] Radius [{r}] }}",
                              instances=LIGHT_INSTANCE.format("bulb"))
    v = LIGHT_AT - pos
    d2 = (v ** 2).sum(1)
    want = 0.6 / np.pi * (k / (4 * np.pi * r * r)) * np.pi * r * r / d2 * (v[:, 2] / np.sqrt(d2))
    ratio = samp[:, 0] / want
    lost = ratio == 0.0
    assert 0.08 < lost.mean() < 0.2, lost.mean()
    assert abs(np.median(ratio[~lost]) - 1.0) < 2e-3
    lo, hi = np.percentile(ratio[~lost], [1, 99])
    assert 0.97 < lo and hi < 1.03, (lo, hi)


def _rect_irradiance(pos, centre, ex, ey, w, h, radiance, n=24):
    """Irradiance on the floor (normal +z) at `pos` from a two-sided rectangle, by midpoint quadrature over its area."""
    g = (np.arange(n) + 0.5) / n - 0.5
    qx, qy = np.meshgrid(g * w, g * h)
    q = centre[None, :] + qx.ravel()[:, None] * ex[None, :] + qy.ravel()[:, None] * ey[None, :]
    nl = np.cross(ex, ey)
    e = np.zeros(len(pos))
    for a in range(0, len(pos), 2048):
        v = q[None, :, :] - pos[a:a + 2048, None, :]
        d2 = (v ** 2).sum(2)
        dist = np.sqrt(d2)
        e[a:a + 2048] = (radiance * np.clip(v[:, :, 2] / dist, 0, None) * np.abs((v * nl).sum(2)) / dist / d2).mean(1) * (w * h)
    return e


def test_light_tree_selection_is_divided_out():
    """Two rectangle lights of very different power and distance: the floor must receive the SUM of their analytic
    contributions — LightTree::select's probabilities (light_tree.rs:135-201) are importance estimates and must cancel."""
    lights = [dict(c=np.array([0.5, 1.0, 3.0]), w=0.4, h=0.3, k=30.0), dict(c=np.array([-2.0, 4.0, 2.0]), w=1.0, h=0.5, k=4.0)]
    objs = "".join(f"RectangleLight $p{i} {{ Color [rec709, {l['k']} {l['k']} {l['k']}] Dimensions [{l['w']} {l['h']}] }}\n"
                   for i, l in enumerate(lights))
    insts = "".join(f"Instance {{ Data [$p{i}] Transform [1 0 0 0  0 1 0 0  0 0 1 0  {-l['c'][0]} {-l['c'][1]} {-l['c'][2]} 1] }}\n"
                    for i, l in enumerate(lights))
    pos, samp, spp = _lit_floor(objects=objs, instances=insts, spp=64)
    ex, ey = np.array([1.0, 0, 0]), np.array([0, 1.0, 0])
    e = sum(_rect_irradiance(pos, l["c"], ex, ey, l["w"], l["h"], l["k"] / (l["w"] * l["h"]) * 0.5) for l in lights)
    want, got = 0.6 / np.pi * e, samp[:, 0]
    assert abs(got.mean() / want.mean() - 1.0) < 4e-3
    per_pixel = got.reshape(-1, spp).mean(1) / want.reshape(-1, spp).mean(1)
    assert np.abs(per_pixel - 1.0).max() < 0.05


def test_light_inside_a_rotated_sub_assembly():
    """Assembly::sample_lights (assembly.rs:46-140) takes the shading point into the instance's space and the light
    sample back out: a rectangle light inside a sub-assembly that is instanced with a rotation and a translation must
    light the floor like the same rectangle placed there directly."""
    a = np.radians(40.0)
    rot = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
    shift, local = np.array([1.0, 2.0, 0.5]), np.array([0.3, -0.2, 2.5])
    cols = [rot.T[:, 0], rot.T[:, 1], rot.T[:, 2], -rot.T @ shift]  # world -> local, column-major, fourth row dropped
    xf = " ".join(f"{v:.9g}" for j, col in enumerate(cols) for v in list(col) + [1.0 if j == 3 else 0.0])
    w, h, k = 0.6, 0.4, 25.0
    objs = f"""Assembly $sub {{
            RectangleLight $panel {{ Color [rec709, {k} {k} {k}] Dimensions [{w} {h}] }}
            Instance {{ Data [$panel] Transform [1 0 0 0  0 1 0 0  0 0 1 0  {-local[0]} {-local[1]} {-local[2]} 1] }}
        }}"""
    pos, samp, spp = _lit_floor(objects=objs, instances=f"Instance {{ Data [$sub] Transform [{xf}] }}", spp=32)
    e = _rect_irradiance(pos, rot @ local + shift, rot @ np.array([1.0, 0, 0]), rot @ np.array([0, 1.0, 0]), w, h,
                         k / (w * h) * 0.5)
    want, got = 0.6 / np.pi * e, samp[:, 0]
    assert abs(got.mean() / want.mean() - 1.0) < 4e-3
    per_pixel = got.reshape(-1, spp).mean(1) / want.reshape(-1, spp).mean(1)
    assert np.abs(per_pixel - 1.0).max() < 0.04


def test_lights_seen_directly_and_the_references_factor_of_two():
    """A camera ray that hits a light picks up its Emit closure: colour / (4 pi r^2) for a sphere
    (sphere_light.rs:318-323) — the radiance its sampling uses — and colour / area for a rectangle
    (rectangle_light.rs:316-320), which is TWICE the colour / area * 0.5 its sampling uses (:191).  The reference is
    inconsistent there (a floor close under a large rectangle light comes out brighter than either convention alone);
    both sides of the parity reproduce it."""
    from psychopath_b200 import api
    from scenes import lit_floor_text
    for obj, name, want, kind in (("RectangleLight $l { Color [rec709, 10 10 10] Dimensions [3.0 2.0] }", "l", 10.0 / 6.0, abi.HIT_RECT_LIGHT),
                                  ("SphereLight $l { Color [rec709, 10 10 10] Radius [0.8] }", "l", 10.0 / (4 * np.pi * 0.64), abi.HIT_SPHERE_LIGHT)):
        text = lit_floor_text(objects=obj, instances="Instance { Data [$l] Transform [1 0 0 0  0 1 0 0  0 0 1 0  -0.5 -1.0 -0.9 1] }", spp=4)
        s = api.Scene.from_psy_string(text)
        d = s.desc.contents
        rays = O.camera_rays(s.desc, 0, 0, d.res_x, d.res_y, d.spp, d.seed)
        hits, _, _ = O.trace_rays(s.desc, rays)
        _, samp, _ = O.render_bucket(s.desc, 0, 0, d.res_x, d.res_y, d.spp, d.seed, want_samples=True)
        direct = hits["kind"] == kind
        assert direct.mean() > 0.05
        assert np.abs(samp.reshape(-1, 4)[direct] / want - 1.0).max() < 1e-3


def test_camera_rays_are_the_pinhole_and_thin_lens_projection():
    """Camera::generate_ray and the image-plane mapping (camera.rs:74-102, renderer.rs:214-255) in closed form: with a
    closed aperture every ray leaves the camera position and, in camera space, dir.xy / dir.z of a pixel's samples
    averages to the pixel centre on the plane z = 1 (x in [-tan(fov/2), tan(fov/2)], y scaled by H / W, row 0 at the
    top); with an open aperture all rays of a pixel sample still cross at the pixel's point on the focus plane."""
    from psychopath_b200 import api
    from scenes import lit_floor_text
    w, h, spp = 32, 24, 64
    for aperture in (0.0, 0.3):
        text = lit_floor_text(res=(w, h), spp=spp).replace("ApertureRadius [0.0]", f"ApertureRadius [{aperture}]")
        s = api.Scene.from_psy_string(text)
        rays = O.camera_rays(s.desc, 0, 0, w, h, spp, 5)
        # the fixture's camera: world <- camera columns X = (1,0,0), Y = (0,0.6,0.8), Z = (0,0.8,-0.6), at (0,-6,5)
        m = np.array([[1, 0, 0], [0, 0.6, 0.8], [0, 0.8, -0.6]], dtype=np.float64).T   # columns X Y Z
        o = (rays["orig"].astype(np.float64) - np.array([0, -6.0, 5.0])) @ m           # camera-space origin
        d = rays["dir"].astype(np.float64) @ m
        assert (d[:, 2] > 0).all()
        tfov = np.tan(np.radians(35.0) / 2)
        focus = 9.0
        # where each ray crosses the focus plane z = focus (z = 1 for the pinhole), as image-plane coordinates
        t = ((focus if aperture else 1.0) - o[:, 2]) / d[:, 2]
        p = (o + d * t[:, None])[:, :2] / ((focus if aperture else 1.0) * tfov)
        p = p.reshape(h, w, spp, 2)
        px = ((np.arange(w) + 0.5) / w - 0.5) * 2.0
        py = (0.5 - (np.arange(h) + 0.5) / h) * 2.0 * (h / w)
        assert np.abs(p[..., 0].mean(2) - px[None, :]).max() < 0.25 * 2.0 / w      # pixel centre, within a quarter pixel
        assert np.abs(p[..., 1].mean(2) - py[:, None]).max() < 0.25 * 2.0 / w
        assert np.abs(p[..., 0] - px[None, :, None]).max() < 2.5 * 2.0 / w          # the filter's footprint: a few pixels
        if aperture == 0.0:
            assert np.abs(o).max() < 1e-5
        else:
            r = np.hypot(o[:, 0], o[:, 1])
            assert np.abs(o[:, 2]).max() < 1e-5 and r.max() <= aperture * 1.0001 and r.max() > 0.9 * aperture


def test_sphere_lambert_against_quadrature():
    """The light tree's importance estimate for Lambert surfaces (surface_closure.rs:351-377, Snyder's clamped-cosine
    integral over a spherical cap) against a direct quadrature of max(cos, 0) / pi over the cap: exact where the cap is
    wholly above the horizon and where its centre is below it.  In between the reference writes
    `nlcos * rsin2 + g - h` where Snyder's formula has `(g - h) / pi`; the oracle (and the CUDA path) keep the
    reference's expression — it only steers which light a path picks, the pick's probability is divided out again
    (test_light_tree_selection_is_divided_out)."""
    lib = O.load()
    lib.oracle_sphere_lambert.restype = C.c_float
    lib.oracle_sphere_lambert.argtypes = [C.c_float, C.c_float, C.POINTER(C.c_float)]
    out = (C.c_float * 4)()

    def quadrature(nl, r, n=1200):
        th = (np.arange(n) + 0.5) / n * r
        ph = (np.arange(n) + 0.5) / n * 2 * np.pi
        t, p = np.meshgrid(th, ph, indexing="ij")
        z = -np.sin(t) * np.cos(p) * np.sin(nl) + np.cos(t) * np.cos(nl)   # cap around an axis at angle nl from the normal
        return (np.clip(z, 0, None) * np.sin(t)).sum() * (r / n) * (2 * np.pi / n) / np.pi

    def f(nl, r):
        return lib.oracle_sphere_lambert(np.float32(np.cos(nl)), np.float32(np.cos(r)), out)

    for nl, r in ((0.3, 0.2), (1.0, 0.4), (0.0, 1.2), (0.9, 0.6)):       # wholly above the horizon
        assert abs(f(nl, r) / quadrature(nl, r) - 1.0) < 1e-5
        assert abs(f(nl, r) / (np.cos(nl) * np.sin(r) ** 2) - 1.0) < 1e-5
    for nl, r in ((1.7, 0.4), (2.0, 0.6), (1.6, 1.0)):                   # centre below the horizon, cap still reaching above
        assert abs(f(nl, r) / quadrature(nl, r) - 1.0) < 1e-4
    assert f(2.2, 0.5) == 0.0                                               # wholly below
    for nl, r in ((1.4, 0.3), (1.5, 0.5), (1.2, 0.7)):                   # straddling, centre above: the reference's form
        got, exact = f(nl, r), quadrature(nl, r)
        flat = np.cos(nl) * np.sin(r) ** 2
        assert abs((flat + (got - flat) / np.pi) / exact - 1.0) < 1e-4      # Snyder's value once (g - h) is divided by pi
        assert got > exact                                                  # as written, an over-estimate


def test_fast_exp_and_plancks_law_in_closed_form():
    """`fastapprox::fast::exp` is an un-vendored dependency restated from its published algorithm: its
    error against exp() must be the library's (a few 1e-5 relative; a wrong constant shows up as 1e-2 or worse).
    Blackbody colours (color.rs:336-383) follow Planck's law in W sr^-1 m^-2 um^-1... times 1e-6, at the hero wavelength
    and three more a quarter of the range apart, zero below 200 K; Temperature colours are that times 4e7 / T^4."""
    lib = O.load()
    lib.oracle_fast_exp.restype = C.c_float
    lib.oracle_fast_exp.argtypes = [C.c_float]
    xs = np.linspace(-80, 80, 4001).astype(np.float32)
    rel = np.array([lib.oracle_fast_exp(x) / np.exp(np.float64(x)) - 1.0 for x in xs])
    assert np.abs(rel).max() < 1e-4, np.abs(rel).max()

    def planck(t, wl_nm):
        c, h, kb = 299792458.0, 6.62607015e-34, 1.38064852e-23
        wl = wl_nm * 1e-9
        return 2 * h * c * c / wl ** 5 / (np.exp(h * c / (wl * kb * t)) - 1.0) * 1e-6

    e = (C.c_float * 4)()
    for t in (1000.0, 2500.0, 3200.0, 5600.0, 6500.0, 9000.0, 15000.0):
        for hero in (380.0, 455.0, 530.0, 612.0, 699.0):
            wls = [380.0 + (hero - 380.0 + 80.0 * i) % 320.0 for i in range(4)]
            lib.oracle_color_to_spectral(C.byref(abi.PsyColor(1, t, 2.5, 0.0)), C.c_float(hero), e)
            for i in range(4):
                assert abs(e[i] / (2.5 * planck(t, wls[i])) - 1.0) < 1.5e-4, (t, hero, i)
            lib.oracle_color_to_spectral(C.byref(abi.PsyColor(2, t, 2.5, 0.0)), C.c_float(hero), e)
            for i in range(4):
                assert abs(e[i] / (2.5 * planck(t, wls[i]) * 4.0e7 / t ** 4) - 1.0) < 1.5e-4
    lib.oracle_color_to_spectral(C.byref(abi.PsyColor(1, 150.0, 1.0, 0.0)), C.c_float(500.0), e)
    assert list(e) == [0.0, 0.0, 0.0, 0.0]


def test_power_heuristic_is_the_pdf_over_the_mis_weight():
    """mis.rs:12-21 returns a / (a^2 / (a^2 + b^2)), the pdf divided by the power-heuristic weight (so dividing a
    contribution by it applies the weight and the pdf at once), and passes an infinite a through (delta lights)."""
    lib = O.load()
    f = lib.oracle_power_heuristic
    for a, b in ((1.0, 1.0), (3.0, 0.5), (0.25, 8.0), (1146.0, 0.3), (2.0, 0.0)):
        w = a * a / (a * a + b * b)
        assert abs(f(a, b) / (a / w) - 1.0) < 1e-6
    assert f(float("inf"), 3.0) == float("inf")
    # the two weights of a pair of techniques add up to one
    for a, b in ((1.0, 1.0), (3.0, 0.5), (0.25, 8.0)):
        assert abs(a / f(a, b) + b / f(b, a) - 1.0) < 1e-6
