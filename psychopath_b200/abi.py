"""ctypes mirror of include/psy_cuda.h (the C ABI) and of libpsy_host.so's entry points.

Keep in sync with the header: tests/test_abi.py checks struct sizes against values the C side
reports and that every declared symbol is exported.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)

PSY_NONE = 0xFFFFFFFF
PSY_RAY_OCCLUSION = 1
HIT_MISS, HIT_TRIANGLE, HIT_SPHERE_LIGHT, HIT_RECT_LIGHT, HIT_OCCLUDED = range(5)

u32 = C.c_uint32
f32 = C.c_float


class PsyXform(C.Structure):
    _fields_ = [("r", (f32 * 4) * 3)]


class PsyBBox(C.Structure):
    _fields_ = [("min", f32 * 3), ("max", f32 * 3)]


class PsyBBox4(C.Structure):
    _fields_ = [("v", (f32 * 4) * 6)]


class PsyBvh4Node(C.Structure):
    _fields_ = [("child", u32 * 4), ("traversal_code", u32), ("n_keys", u32), ("key_base", u32),
                ("n_children", u32), ("bounds", PsyBBox4)]


class PsyColor(C.Structure):
    _fields_ = [("kind", u32), ("a", f32), ("b", f32), ("c", f32)]


class PsyShader(C.Structure):
    _fields_ = [("kind", u32), ("color", PsyColor), ("roughness", f32), ("fresnel", f32)]


class PsyMesh(C.Structure):
    _fields_ = [("root", u32), ("tri_base", u32), ("tri_count", u32), ("n_keys", u32), ("vert_base", u32),
                ("vert_count", u32), ("has_normals", u32), ("bounds_base", u32), ("bounds_count", u32),
                ("bvh_depth", u32)]


class PsyLight(C.Structure):
    _fields_ = [("kind", u32), ("n_shape", u32), ("shape_base", u32), ("n_colors", u32), ("color_base", u32),
                ("bounds_base", u32), ("bounds_count", u32)]


class PsyDistantLight(C.Structure):
    _fields_ = [("n_radii", u32), ("radii_base", u32), ("n_dirs", u32), ("dir_base", u32), ("n_colors", u32),
                ("color_base", u32)]


class PsyObject(C.Structure):
    _fields_ = [("kind", u32), ("index", u32)]


class PsyInstance(C.Structure):
    _fields_ = [("kind", u32), ("data_index", u32), ("shader_index", u32), ("id", u32), ("xform_base", u32),
                ("xform_count", u32)]


class PsyLightNode(C.Structure):
    _fields_ = [("child_base", u32), ("child_count", u32), ("light_index", u32), ("bounds_base", u32),
                ("bounds_count", u32), ("energy", f32)]


class PsyAssembly(C.Structure):
    _fields_ = [("root", u32), ("instance_base", u32), ("instance_count", u32), ("light_instance_base", u32),
                ("light_instance_count", u32), ("light_root", u32), ("light_energy", f32), ("bounds_base", u32),
                ("bounds_count", u32), ("bvh_depth", u32)]


class PsyCamera(C.Structure):
    _fields_ = [("xforms", C.POINTER(PsyXform)), ("n_xforms", u32), ("tfovs", C.POINTER(f32)), ("n_tfovs", u32),
                ("aperture_radii", C.POINTER(f32)), ("n_aperture_radii", u32),
                ("focus_distances", C.POINTER(f32)), ("n_focus_distances", u32)]


class PsySceneDesc(C.Structure):
    _fields_ = [
        ("abi_version", u32), ("res_x", u32), ("res_y", u32), ("spp", u32), ("seed", u32),
        ("camera", PsyCamera),
        ("background", PsyColor),
        ("distant_lights", C.POINTER(PsyDistantLight)), ("n_distant_lights", u32),
        ("world_floats", C.POINTER(f32)), ("n_world_floats", u32),
        ("nodes", C.POINTER(PsyBvh4Node)), ("n_nodes", u32),
        ("motion_bounds", C.POINTER(PsyBBox4)), ("n_motion_bounds", u32),
        ("meshes", C.POINTER(PsyMesh)), ("n_meshes", u32),
        ("tris", C.POINTER(u32)), ("n_tris", u32),
        ("verts", C.POINTER(f32)), ("normals", C.POINTER(f32)), ("n_vert_keys", u32),
        ("lights", C.POINTER(PsyLight)), ("n_lights", u32),
        ("light_shape", C.POINTER(f32)), ("n_light_shape", u32),
        ("light_colors", C.POINTER(PsyColor)), ("n_light_colors", u32),
        ("objects", C.POINTER(PsyObject)), ("n_objects", u32),
        ("shaders", C.POINTER(PsyShader)), ("n_shaders", u32),
        ("assemblies", C.POINTER(PsyAssembly)), ("n_assemblies", u32), ("root_assembly", u32),
        ("instances", C.POINTER(PsyInstance)), ("n_instances", u32),
        ("light_instances", C.POINTER(PsyInstance)), ("n_light_instances", u32),
        ("xforms", C.POINTER(PsyXform)), ("n_xforms", u32),
        ("light_nodes", C.POINTER(PsyLightNode)), ("n_light_nodes", u32),
        ("lt_bounds", C.POINTER(PsyBBox)), ("n_lt_bounds", u32),
        ("obj_bounds", C.POINTER(PsyBBox)), ("n_obj_bounds", u32),
    ]


class PsyRay(C.Structure):
    _fields_ = [("orig", f32 * 3), ("dir", f32 * 3), ("time", f32), ("max_t", f32), ("flags", u32)]


class PsyHit(C.Structure):
    _fields_ = [("kind", u32), ("t", f32), ("prim_id", u32), ("inst_id", u32), ("path_id", u32), ("b0", f32),
                ("b1", f32), ("b2", f32)]


class PsyHitGeo(C.Structure):
    _fields_ = [("pos", f32 * 3), ("pos_err", f32), ("nor", f32 * 3), ("sample_pdf", f32), ("nor_g", f32 * 3),
                ("closure_kind", u32)]


class PsyStats(C.Structure):
    _fields_ = [("trace_time", C.c_double), ("initial_ray_generation_time", C.c_double),
                ("ray_generation_time", C.c_double), ("sample_writing_time", C.c_double),
                ("total_time", C.c_double), ("ray_count", C.c_uint64), ("accel_node_visits", C.c_uint64),
                ("tri_tests", C.c_uint64), ("instance_visits", C.c_uint64), ("motion_node_visits", C.c_uint64),
                ("samples", C.c_uint64), ("kernel_launches", C.c_uint64), ("buckets", C.c_uint64)]


class PsyRenderParams(C.Structure):
    _fields_ = [("crop_enabled", u32), ("crop_x1", u32), ("crop_y1", u32), ("crop_x2", u32), ("crop_y2", u32),
                ("spp", u32), ("seed_override_enabled", u32), ("seed", u32), ("max_samples_per_bucket", u32),
                ("collect_counters", u32), ("device_resident_only", u32), ("strict_film_order", u32),
                ("buckets_per_gpu", u32), ("work_queue_key", u32)]


BUCKET_CB = C.CFUNCTYPE(None, C.c_void_p, u32, u32, u32, u32, C.POINTER(f32))

# numpy dtypes matching PsyRay / PsyHit / PsyHitGeo (all fields 4 bytes, no padding)
import numpy as _np  # noqa: E402

RAY_DTYPE = _np.dtype([("orig", "<f4", 3), ("dir", "<f4", 3), ("time", "<f4"), ("max_t", "<f4"), ("flags", "<u4")])
HIT_DTYPE = _np.dtype([("kind", "<u4"), ("t", "<f4"), ("prim_id", "<u4"), ("inst_id", "<u4"), ("path_id", "<u4"),
                       ("b0", "<f4"), ("b1", "<f4"), ("b2", "<f4")])
GEO_DTYPE = _np.dtype([("pos", "<f4", 3), ("pos_err", "<f4"), ("nor", "<f4", 3), ("sample_pdf", "<f4"),
                       ("nor_g", "<f4", 3), ("closure_kind", "<u4")])

HOST_LIB = os.path.join(HERE, "libpsy_host.so")
CUDA_LIB = os.environ.get("PSY_CUDA_LIB", os.path.join(HERE, "libpsy_cuda.so"))  # override: kernel experiments
CUDA_LIB_STRICT = os.path.join(HERE, "libpsy_cuda_strict.so")  # the same sources built with -DPSY_STRICT=1
PSY_BUILD_STRICT_MATH = 1

# Every symbol include/psy_cuda.h declares.
CUDA_SYMBOLS = [
    "psy_last_error", "psy_abi_version", "psy_build_flags", "psy_init", "psy_shutdown", "psy_scene_create", "psy_scene_destroy",
    "psy_trace_rays", "psy_render_bucket", "psy_render_bucket_samples", "psy_render_image",
    "psy_render_image_shard", "psy_device_image", "psy_scene_stream", "psy_fetch_image", "psy_last_trace_timing",
    "psy_work_queue_open", "psy_work_queue_pull", "psy_work_queue_close", "psy_build_mesh_topology",
    "psy_host_register", "psy_host_unregister", "psy_mesh_bvh_build", "psy_mesh_bvh_take",
]


def load_host():
    lib = C.CDLL(HOST_LIB)
    lib.psyh_last_error.restype = C.c_char_p
    lib.psyh_load_psy_string.argtypes = [C.c_char_p, u32, C.POINTER(C.c_void_p)]
    lib.psyh_load_psy_file.argtypes = [C.c_char_p, u32, C.POINTER(C.c_void_p)]
    lib.psyh_count_scenes.argtypes = [C.c_char_p, C.POINTER(u32)]
    lib.psyh_synth_displaced_grid.argtypes = [u32] * 7 + [C.POINTER(C.c_void_p)]
    lib.psyh_synth_instanced.argtypes = [u32] * 11 + [C.POINTER(C.c_void_p)]
    lib.psyh_scene_desc.argtypes = [C.c_void_p]
    lib.psyh_scene_desc.restype = C.POINTER(PsySceneDesc)
    lib.psyh_build_seconds.argtypes = [C.c_void_p]
    lib.psyh_build_seconds.restype = C.c_double
    lib.psyh_output_path.argtypes = [C.c_void_p]
    lib.psyh_output_path.restype = C.c_char_p
    lib.psyh_scene_set_render.argtypes = [C.c_void_p, u32, u32, u32, u32]
    lib.psyh_scene_set_render.restype = None
    lib.psyh_scene_free.argtypes = [C.c_void_p]
    lib.psyh_scene_free.restype = None
    lib.psyh_hilbert_d2xy.argtypes = [u32, C.POINTER(u32), C.POINTER(u32)]
    lib.psyh_hilbert_d2xy.restype = None
    lib.psyh_bucket_list.argtypes = [u32] * 6 + [C.POINTER(u32), u32]
    lib.psyh_bucket_list.restype = u32
    lib.psyh_set_mesh_topology_builder.argtypes = [C.c_void_p, C.c_void_p, u32]
    lib.psyh_set_mesh_topology_builder.restype = None
    lib.psyh_test_datatree_dump.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t]
    lib.psyh_test_datatree_dump.restype = C.c_long
    lib.psyh_test_parse_floats.argtypes = [C.c_char_p, C.POINTER(f32), C.c_size_t, C.POINTER(C.c_size_t)]
    lib.psyh_test_parse_floats.restype = C.c_size_t
    lib.psyh_test_parse_uints.argtypes = [C.c_char_p, C.POINTER(u32), C.c_size_t, C.POINTER(C.c_size_t)]
    lib.psyh_test_parse_uints.restype = C.c_size_t
    lib.psyh_test_quick_select.argtypes = [C.POINTER(C.c_int32), C.c_size_t, C.c_size_t]
    lib.psyh_test_quick_select.restype = None
    lib.psyh_test_log2_64.argtypes = [C.c_uint64]
    lib.psyh_test_log2_64.restype = C.c_uint64
    lib.psyh_set_mesh_bvh_builder.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, u32]
    lib.psyh_set_mesh_bvh_builder.restype = None
    lib.psyh_take_mesh_topology_seconds.argtypes = [C.POINTER(u32)]
    lib.psyh_take_mesh_topology_seconds.restype = C.c_double
    return lib


def load_cuda(path=None):
    """Loads the CUDA library.  There is no fallback: a missing library is an error."""
    path = path or CUDA_LIB
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(path)
    lib.psy_last_error.restype = C.c_char_p
    lib.psy_abi_version.restype = C.c_int32
    lib.psy_build_flags.restype = C.c_uint32
    lib.psy_init.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]
    lib.psy_shutdown.argtypes = [C.c_void_p]
    lib.psy_scene_create.argtypes = [C.c_void_p, C.POINTER(PsySceneDesc), C.POINTER(C.c_void_p)]
    lib.psy_scene_destroy.argtypes = [C.c_void_p]
    lib.psy_trace_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    lib.psy_render_bucket.argtypes = [C.c_void_p, u32, u32, u32, u32, u32, u32, C.c_void_p, C.POINTER(PsyStats)]
    lib.psy_render_bucket_samples.argtypes = lib.psy_render_bucket.argtypes
    lib.psy_render_image.argtypes = [C.c_void_p, C.POINTER(PsyRenderParams), C.c_void_p, C.POINTER(PsyStats),
                                     C.c_void_p, C.c_void_p]
    lib.psy_render_image_shard.argtypes = [C.c_void_p, C.POINTER(PsyRenderParams), u32, u32, C.POINTER(PsyStats)]
    lib.psy_device_image.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    lib.psy_scene_stream.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    lib.psy_fetch_image.argtypes = [C.c_void_p, C.c_void_p]
    lib.psy_last_trace_timing.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    lib.psy_work_queue_open.argtypes = [u32, C.POINTER(C.c_void_p)]
    lib.psy_work_queue_pull.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
    lib.psy_work_queue_close.argtypes = [C.c_void_p]
    lib.psy_host_register.argtypes = [C.c_void_p, C.c_size_t]
    lib.psy_host_unregister.argtypes = [C.c_void_p]
    for name in CUDA_SYMBOLS:
        if name not in ("psy_last_error", "psy_build_flags"):
            getattr(lib, name).restype = C.c_int32
    return lib
