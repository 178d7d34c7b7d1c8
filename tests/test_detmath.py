"""include/psy_detmath.h — the deterministic sin / cos / acos / asin / log both strict builds share.

Accuracy against numpy (glibc-class libm); the GPU side of "same bits on CPU and GPU" is what
tests/test_gpu_strict.py proves by comparing whole renders bit for bit."""
import numpy as np

import oracle_lib as O

F64 = {"sin": (0, np.sin, (-13.0, 13.0)), "cos": (1, np.cos, (-13.0, 13.0)), "acos": (2, np.arccos, (-1.0, 1.0)),
       "asin": (3, np.arcsin, (-1.0, 1.0)), "log": (4, np.log, (1e-300, 1e3))}


def _f64(fn, xs):
    lib = O.load()
    return np.array([lib.oracle_det_f64(fn, float(v)) for v in xs])


def _f32(fn, xs):
    lib = O.load()
    return np.array([lib.oracle_det_f32(fn, float(v)) for v in xs], dtype=np.float32)


def test_f64_functions_within_two_ulp():
    rng = np.random.default_rng(7)
    n = 20000
    for name, (fn, ref, dom) in F64.items():
        x = rng.uniform(dom[0], dom[1], n)
        if name == "log":
            x = np.concatenate([x, 10.0 ** rng.uniform(-300, 300, n), rng.uniform(0.5, 2.0, n)])
        if name in ("acos", "asin"):
            edge = 1.0 - 10.0 ** rng.uniform(-16, 0, n)
            x = np.concatenate([x, edge, -edge])
        got, want = _f64(fn, x), ref(x)
        ulp = np.abs(got - want) / np.spacing(np.abs(want))
        assert ulp.max() <= (1.0 if name != "asin" else 2.0), (name, float(ulp.max()), float(x[ulp.argmax()]))


def test_f32_functions_correctly_rounded():
    rng = np.random.default_rng(8)
    n = 20000
    for fn, ref, dom in [(0, np.sin, (-13, 13)), (1, np.cos, (-13, 13)), (2, np.arccos, (-1, 1)), (3, np.arcsin, (-1, 1))]:
        x = rng.uniform(dom[0], dom[1], n).astype(np.float32)
        want = ref(x.astype(np.float64)).astype(np.float32)
        assert np.array_equal(_f32(fn, x), want)


def test_special_values():
    lib = O.load()
    assert lib.oracle_det_f64(2, 1.0) == 0.0 and abs(lib.oracle_det_f64(2, -1.0) - np.pi) < 1e-15
    assert lib.oracle_det_f64(4, 1.0) == 0.0 and lib.oracle_det_f64(4, 0.0) == -np.inf
    assert np.isnan(lib.oracle_det_f64(4, -1.0)) and np.isnan(lib.oracle_det_f64(2, 1.5)) and np.isnan(lib.oracle_det_f64(0, np.inf))
    assert lib.oracle_det_f64(0, 0.0) == 0.0 and lib.oracle_det_f64(1, 0.0) == 1.0
    assert lib.oracle_det_f32(3, 1.0) == np.float32(np.pi / 2)


def test_strict_and_libm_oracle_images_differ_only_by_libm_noise():
    """What the default-mode image tolerance measures: the oracle against itself with the other libm."""
    from psychopath_b200 import api
    from scenes import cornell_text
    s = api.Scene.from_psy_string(cornell_text())
    s.set_render(res=(64, 64), spp=16, seed=1)
    a, sa, _ = O.render_image(s.desc, 16, 1)
    with O.strict_math():
        b, sb, _ = O.render_image(s.desc, 16, 1)
    rel = np.linalg.norm(a - b) / np.linalg.norm(a)
    assert rel < 2e-3, rel
    assert abs(int(sa.ray_count) - int(sb.ray_count)) <= 0.002 * sa.ray_count
