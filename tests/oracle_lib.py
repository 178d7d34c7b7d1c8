"""ctypes bindings of the CPU oracle (oracle/liboracle.so) — test infrastructure only.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module; nothing under psychopath_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from psychopath_b200 import abi

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(REPO, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "liboracle.so")


class OracleStats(C.Structure):
    _fields_ = [("trace_time", C.c_double), ("initial_ray_generation_time", C.c_double),
                ("ray_generation_time", C.c_double), ("sample_writing_time", C.c_double),
                ("total_time", C.c_double), ("ray_count", C.c_uint64), ("accel_node_visits", C.c_uint64),
                ("tri_tests", C.c_uint64), ("instance_visits", C.c_uint64), ("motion_node_visits", C.c_uint64),
                ("samples", C.c_uint64), ("threads", C.c_uint32)]


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("PSY_ORACLE_LIB", ORACLE_LIB)  # bench.py: the -march=native build made on the box
    if not os.path.exists(path):
        build_oracle()
        path = ORACLE_LIB
    lib = C.CDLL(path)
    u32, f32 = C.c_uint32, C.c_float
    D = C.POINTER(abi.PsySceneDesc)
    lib.oracle_trace_rays.argtypes = [D, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.oracle_render_bucket.argtypes = [D, u32, u32, u32, u32, u32, u32, C.c_void_p, C.c_void_p, C.POINTER(OracleStats)]
    lib.oracle_render_image.argtypes = [D, C.c_void_p, u32, u32, u32, u32, u32, C.c_void_p, C.POINTER(OracleStats),
                                        C.POINTER(u32)]
    lib.oracle_camera_rays.argtypes = [D, u32, u32, u32, u32, u32, u32, C.c_void_p]
    lib.oracle_hash_u32.argtypes = [u32, u32]
    lib.oracle_hash_u32.restype = u32
    lib.oracle_golden_ratio_sample.argtypes = [u32, u32]
    lib.oracle_golden_ratio_sample.restype = f32
    lib.oracle_get_sample_4d.argtypes = [u32, u32, u32, u32, u32, C.POINTER(f32)]
    lib.oracle_get_sample_4d.restype = None
    lib.oracle_probit.argtypes = [f32, f32]
    lib.oracle_probit.restype = f32
    lib.oracle_fp_gamma.argtypes = [u32]
    lib.oracle_fp_gamma.restype = f32
    for n in ("oracle_increment_ulp", "oracle_decrement_ulp", "oracle_fast_exp"):
        getattr(lib, n).argtypes = [f32]
        getattr(lib, n).restype = f32
    lib.oracle_power_heuristic.argtypes = [f32, f32]
    lib.oracle_power_heuristic.restype = f32
    lib.oracle_hilbert_d2xy.argtypes = [u32, C.POINTER(u32), C.POINTER(u32)]
    lib.oracle_hilbert_d2xy.restype = None
    lib.oracle_traversal_table.argtypes = [C.c_void_p]
    lib.oracle_traversal_table.restype = None
    X = C.POINTER(abi.PsyXform)
    lib.oracle_xform_mul.argtypes = [X, X, X]
    lib.oracle_xform_inverse.argtypes = [X, X]
    lib.oracle_xform_lerp.argtypes = [X, X, f32, X]
    F = C.POINTER(f32)
    for n in ("oracle_xform_point", "oracle_xform_vector", "oracle_xform_normal"):
        getattr(lib, n).argtypes = [X, F, F]
        getattr(lib, n).restype = None
    lib.oracle_lerp_slice_f32.argtypes = [F, u32, f32]
    lib.oracle_lerp_slice_f32.restype = f32
    lib.oracle_color_to_spectral.argtypes = [C.POINTER(abi.PsyColor), f32, F]
    lib.oracle_color_to_spectral.restype = None
    lib.oracle_xyz_from_spectral.argtypes = [F, f32, F]
    lib.oracle_xyz_from_spectral.restype = None
    lib.oracle_tri_intersect.argtypes = [F, F, f32, F, F]
    lib.oracle_sphere_lambert.argtypes = [f32, f32, F]
    lib.oracle_sphere_lambert.restype = f32
    lib.oracle_set_strict_math.argtypes = [C.c_int]
    lib.oracle_set_strict_math.restype = None
    lib.oracle_set_debug.argtypes = [u32, u32, u32]
    lib.oracle_set_debug.restype = None
    lib.oracle_det_f64.argtypes = [C.c_int, C.c_double]
    lib.oracle_det_f64.restype = C.c_double
    lib.oracle_det_f32.argtypes = [C.c_int, f32]
    lib.oracle_det_f32.restype = f32
    _lib = lib
    return lib


class strict_math:
    """Context manager: the oracle calls include/psy_detmath.h instead of glibc for sin/cos/acos/asin/log, like
    the strict build of the CUDA library (libpsy_cuda_strict.so)."""

    def __enter__(self):
        load().oracle_set_strict_math(1)
        return self

    def __exit__(self, *exc):
        load().oracle_set_strict_math(0)
        return False


def trace_rays(desc, rays, want_geo=False):
    """rays: numpy array of abi.RAY_DTYPE.  Returns (hits, geo|None, counters[4])."""
    lib = load()
    rays = np.ascontiguousarray(rays)
    hits = np.zeros(len(rays), dtype=abi.HIT_DTYPE)
    geo = np.zeros(len(rays), dtype=abi.GEO_DTYPE) if want_geo else None
    counters = np.zeros(4, dtype=np.uint64)
    rc = lib.oracle_trace_rays(desc, rays.ctypes.data, len(rays), hits.ctypes.data,
                               geo.ctypes.data if want_geo else None, counters.ctypes.data)
    assert rc == 0
    return hits, geo, counters


def camera_rays(desc, x, y, w, h, spp, seed):
    lib = load()
    rays = np.zeros(w * h * spp, dtype=abi.RAY_DTYPE)
    lib.oracle_camera_rays(desc, x, y, w, h, spp, seed, rays.ctypes.data)
    return rays


def render_bucket(desc, x, y, w, h, spp, seed, want_samples=False):
    lib = load()
    xyz = np.zeros((h, w, 3), dtype=np.float32)
    samples = np.zeros((h, w, spp, 4), dtype=np.float32) if want_samples else None
    st = OracleStats()
    rc = lib.oracle_render_bucket(desc, x, y, w, h, spp, seed, xyz.ctypes.data,
                                  samples.ctypes.data if want_samples else None, C.byref(st))
    assert rc == 0
    return xyz, samples, st


def render_image(desc, spp, seed, spb=4096, threads=0, crop=None, max_buckets=0):
    lib = load()
    d = desc.contents
    xyz = np.zeros((d.res_y, d.res_x, 3), dtype=np.float32)
    st = OracleStats()
    done = C.c_uint32(0)
    cr = None
    if crop is not None:
        cr = np.asarray(crop, dtype=np.uint32)
    rc = lib.oracle_render_image(desc, cr.ctypes.data if cr is not None else None, spp, seed, spb, threads,
                                 max_buckets, xyz.ctypes.data, C.byref(st), C.byref(done))
    assert rc == 0
    return xyz, st, done.value
