"""Host-side mirror of the reference's interface to the render core, bound to the CUDA library
through the C ABI (include/psy_cuda.h).

  Scene      <- what `parse_scene` returns inside `Renderer` (src/parse/psy.rs:95-200): the built scene
  Renderer   <- `Renderer { output_file, resolution, spp, seed, scene }` and `Renderer::render(
                max_samples_per_bucket, crop, thread_count, do_blender_output) -> (Image, RenderStats)`
                (src/renderer.rs:32-38,76-191)
  Tracer     <- `Tracer::from_assembly` / `Tracer::trace(&mut RayBatch) -> &[SurfaceIntersection]`
                (src/tracer.rs:22-37)

There is no CPU path here: if libpsy_cuda.so is missing or no GPU is visible the constructors raise.
"""
import ctypes as C

import numpy as np

from . import abi

_host = None
_cuda = {}


def host_lib():
    global _host
    if _host is None:
        _host = abi.load_host()
    return _host


def cuda_lib(strict=False):
    """The CUDA library; strict=True loads libpsy_cuda_strict.so (the same kernels built with -DPSY_STRICT=1: the
    deterministic libm replacements of include/psy_detmath.h, for bit-level parity runs against the oracle)."""
    if strict not in _cuda:
        lib = abi.load_cuda(abi.CUDA_LIB_STRICT if strict else abi.CUDA_LIB)
        flags = lib.psy_build_flags()
        if bool(flags & abi.PSY_BUILD_STRICT_MATH) != bool(strict) and "PSY_CUDA_LIB" not in __import__("os").environ:
            raise PsyError(f"library build flags {flags:#x} do not match strict={strict}")
        _cuda[strict] = lib
    return _cuda[strict]


class PsyError(RuntimeError):
    pass


def enable_gpu_build(min_tris=1 << 20, device=0, mode="bvh4"):
    """meshes of at least `min_tris` triangles are built on the GPU instead of by the host's SAH builder.
    mode "bvh4": the whole BVH4 of the mesh (LBVH topology, time-sampled bounds, union rule, BVH4 collapse:
    psy_mesh_bvh_build); the host only reorders the triangle list and appends the nodes.  mode "topology": triangle
    order and binary topology only (psy_build_mesh_topology), bounds and collapse on the host — the same nodes, bit
    for bit.  Scenes built afterwards (Scene.from_* / synth_*) use it; disable_gpu_build() restores the SAH builder."""
    lib = cuda_lib()
    ctx = _context(device)
    if mode not in ("bvh4", "topology"):
        raise ValueError(mode)
    host_lib().psyh_set_mesh_topology_builder(C.cast(lib.psy_build_mesh_topology, C.c_void_p), ctx, int(min_tris))
    if mode == "bvh4":
        host_lib().psyh_set_mesh_bvh_builder(C.cast(lib.psy_mesh_bvh_build, C.c_void_p),
                                             C.cast(lib.psy_mesh_bvh_take, C.c_void_p), ctx, int(min_tris))
    else:
        host_lib().psyh_set_mesh_bvh_builder(None, None, None, 0)


def disable_gpu_build():
    host_lib().psyh_set_mesh_topology_builder(None, None, 0)
    host_lib().psyh_set_mesh_bvh_builder(None, None, None, 0)


def gpu_build_seconds():
    """(seconds inside the GPU builder, meshes it built) since the last call."""
    n = C.c_uint32()
    s = host_lib().psyh_take_mesh_topology_seconds(C.byref(n))
    return s, n.value


class Scene:
    """A parsed and built scene living in host memory (flat PsySceneDesc)."""

    def __init__(self, handle):
        self._h = handle
        self.desc = host_lib().psyh_scene_desc(handle)

    def __del__(self):
        try:
            self.unpin()
            if self._h:
                host_lib().psyh_scene_free(self._h)
                self._h = None
        except Exception:
            pass

    @staticmethod
    def _check(rc):
        if rc != 0:
            raise PsyError(host_lib().psyh_last_error().decode())

    @classmethod
    def from_psy_file(cls, path, index=0):
        h = C.c_void_p()
        cls._check(host_lib().psyh_load_psy_file(str(path).encode(), index, C.byref(h)))
        return cls(h)

    @classmethod
    def from_psy_string(cls, text, index=0):
        h = C.c_void_p()
        cls._check(host_lib().psyh_load_psy_string(text.encode(), index, C.byref(h)))
        return cls(h)

    @classmethod
    def synth_displaced_grid(cls, quads_per_side, n_vert_keys=4, n_xform_keys=4, res=(1920, 1080), spp=128, seed=3):
        h = C.c_void_p()
        cls._check(host_lib().psyh_synth_displaced_grid(quads_per_side, n_vert_keys, n_xform_keys, res[0], res[1], spp,
                                                        seed, C.byref(h)))
        return cls(h)

    @classmethod
    def synth_instanced(cls, asset_quads=500, n_assets=4, n_clusters=100, inst_per_cluster=10, lights_per_cluster=10,
                        cluster_xform_keys=2, res=(1920, 1080), spp=64, seed=4, motion_pct=100):
        """motion_pct scales the cluster instances' motion over the shutter (100: drift <= 0.4 units, which the
        builder's 1.4x union rule merges into single-key bounds; 1250: BASELINE config 4's motion-blurred scene)."""
        h = C.c_void_p()
        cls._check(host_lib().psyh_synth_instanced(asset_quads, n_assets, n_clusters, inst_per_cluster,
                                                   lights_per_cluster, cluster_xform_keys, res[0], res[1], spp, seed,
                                                   motion_pct, C.byref(h)))
        return cls(h)

    def set_render(self, res=None, spp=0, seed=None):
        d = self.desc.contents
        host_lib().psyh_scene_set_render(self._h, res[0] if res else 0, res[1] if res else 0, spp,
                                         d.seed if seed is None else seed)

    @property
    def resolution(self):
        d = self.desc.contents
        return (d.res_x, d.res_y)

    @property
    def spp(self):
        return self.desc.contents.spp

    @property
    def seed(self):
        return self.desc.contents.seed

    @property
    def output_file(self):
        return host_lib().psyh_output_path(self._h).decode()

    @property
    def build_seconds(self):
        return host_lib().psyh_build_seconds(self._h)

    def pin(self):
        """Page-locks the large scene arrays (nodes, time-key bounds, triangles, vertices) so that every later upload
        of this scene (Renderer(scene)) copies at the link's rate.  One-time cost; undone when the scene is freed."""
        if getattr(self, "_pinned", None):
            return
        lib = cuda_lib()
        d = self.desc.contents
        regs = []
        for ptr, nbytes in ((d.nodes, d.n_nodes * 128), (d.motion_bounds, d.n_motion_bounds * 96), (d.tris, d.n_tris * 16),
                            (d.verts, d.n_vert_keys * 12), (d.normals, d.n_vert_keys * 12)):
            addr = C.cast(ptr, C.c_void_p).value
            if addr and nbytes >= (1 << 20):
                if lib.psy_host_register(addr, nbytes) == 0:
                    regs.append(addr)
        self._pinned = regs

    def unpin(self):
        for addr in getattr(self, "_pinned", None) or []:
            cuda_lib().psy_host_unregister(addr)
        self._pinned = None

    def host_bytes(self):
        """Bytes of scene data the upload copies host->device."""
        d = self.desc.contents
        return (d.n_nodes * 128 + d.n_motion_bounds * 96 + d.n_tris * 16 + d.n_vert_keys * 24 + d.n_instances * 24 +
                d.n_xforms * 48 + d.n_light_nodes * 24 + d.n_lt_bounds * 24)


_ctx = {}


def _context(device, strict=False):
    """device: a CUDA device index, or a tuple of them (one context driving several GPUs from this process)."""
    lib = cuda_lib(strict)
    if (device, strict) not in _ctx:
        ctx = C.c_void_p()
        devs = tuple(device) if isinstance(device, (tuple, list)) else (device,)
        ids = (C.c_int * len(devs))(*devs)
        rc = lib.psy_init(ids, len(devs), C.byref(ctx))
        if rc != 0:
            raise PsyError(f"psy_init failed ({rc}): {lib.psy_last_error().decode()}")
        _ctx[(device, strict)] = ctx
    return _ctx[(device, strict)]


def stats_dict(st):
    return {k: getattr(st, k) for k, _ in abi.PsyStats._fields_}


class DeviceScene:
    """The scene uploaded to one GPU (PsyScene*)."""

    def __init__(self, scene, device=0, strict=False):
        """device: one CUDA device index, or a tuple of indices: the scene is then replicated on all of them and
        render() shares the image's buckets between them (gathered onto the first with NCCL)."""
        if isinstance(device, list):
            device = tuple(device)
        self.scene = scene
        self.strict = bool(strict)
        self.lib = cuda_lib(self.strict)
        self.ctx = _context(device, self.strict)
        self.device = device
        self._s = C.c_void_p()
        rc = self.lib.psy_scene_create(self.ctx, scene.desc, C.byref(self._s))
        if rc != 0:
            raise PsyError(f"psy_scene_create failed ({rc}): {self.lib.psy_last_error().decode()}")

    def close(self):
        if self._s:
            self.lib.psy_scene_destroy(self._s)
            self._s = C.c_void_p()
        for img in getattr(self, "_pinned_images", []):
            self.lib.psy_host_unregister(img.ctypes.data)
        self._pinned_images = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise PsyError(f"{what} failed ({rc}): {self.lib.psy_last_error().decode()}")


class Tracer(DeviceScene):
    """Tracer::trace over a batch of probe rays."""

    def trace(self, rays, want_geo=False):
        rays = np.ascontiguousarray(rays, dtype=abi.RAY_DTYPE)
        hits = np.zeros(len(rays), dtype=abi.HIT_DTYPE)
        geo = np.zeros(len(rays), dtype=abi.GEO_DTYPE) if want_geo else None
        self._check(self.lib.psy_trace_rays(self._s, rays.ctypes.data, len(rays), hits.ctypes.data,
                                            geo.ctypes.data if want_geo else None), "psy_trace_rays")
        return (hits, geo) if want_geo else hits


class Renderer(Tracer):
    """Renderer::render and the per-bucket body of render_job."""

    @property
    def resolution(self):
        return self.scene.resolution

    @property
    def spp(self):
        return self.scene.spp

    @property
    def seed(self):
        return self.scene.seed

    @property
    def output_file(self):
        return self.scene.output_file

    def _params(self, max_samples_per_bucket, crop, spp, seed, counters, resident, strict_film=False,
                buckets_per_gpu=0, work_queue_key=0):
        p = abi.PsyRenderParams()
        p.strict_film_order = 1 if strict_film else 0
        p.buckets_per_gpu = int(buckets_per_gpu or 0)
        p.work_queue_key = int(work_queue_key or 0) & 0xFFFFFFFF
        if crop is not None:
            p.crop_enabled = 1
            p.crop_x1, p.crop_y1, p.crop_x2, p.crop_y2 = [int(c) for c in crop]
        p.spp = int(spp or 0)
        if seed is not None:
            p.seed_override_enabled = 1
            p.seed = int(seed)
        p.max_samples_per_bucket = int(max_samples_per_bucket or 0)
        p.collect_counters = 1 if counters else 0
        p.device_resident_only = 1 if resident else 0
        return p

    def render(self, max_samples_per_bucket=0, crop=None, thread_count=None, do_blender_output=False, spp=None,
               seed=None, counters=False, on_bucket=None, device_resident_only=False, strict_film=False,
               buckets_per_gpu=0, out=None):
        """Returns (image, stats): image is float32 (res_y, res_x, 3) XYZ, row 0 = top (None when
        device_resident_only).  thread_count is accepted for signature parity and ignored (the GPU is the
        worker pool).  do_blender_output=True streams every finished bucket to stdout in the reference's
        protocol (renderer.rs:312-346) unless the caller installs its own on_bucket.
        strict_film=True adds the samples to the film in the reference's path order (PsyRenderParams)."""
        p = self._params(max_samples_per_bucket, crop, spp, seed, counters, device_resident_only, strict_film,
                         buckets_per_gpu)
        rx, ry = self.resolution
        if device_resident_only:
            img = None
        elif out is not None:
            # a caller-owned image, e.g. page-locked memory reused from frame to frame (pinned_image())
            assert out.dtype == np.float32 and out.shape == (ry, rx, 3) and out.flags["C_CONTIGUOUS"]
            img = out
        else:
            img = np.zeros((ry, rx, 3), dtype=np.float32)
        st = abi.PsyStats()
        cb = None
        if do_blender_output and on_bucket is None:
            on_bucket = self._blender_stream(crop)
        if on_bucket is not None:
            def _cb(user, x0, y0, x1, y1, ptr):
                n = (x1 - x0) * (y1 - y0) * 3
                on_bucket(x0, y0, x1, y1, np.ctypeslib.as_array(ptr, shape=(n,)).reshape(y1 - y0, x1 - x0, 3).copy())
            cb = abi.BUCKET_CB(_cb)
        self._check(self.lib.psy_render_image(self._s, C.byref(p), img.ctypes.data if img is not None else None,
                                              C.byref(st), C.cast(cb, C.c_void_p) if cb else None, None),
                    "psy_render_image")
        return img, stats_dict(st)

    def _blender_stream(self, crop):
        """Default bucket sink for do_blender_output: `DIV / pct% / x0 y0 x1 y1 / base64 RGBA / BUCKET_END / DIV`."""
        import sys
        rx, ry = self.resolution
        if crop is not None:
            x1, y1, x2, y2 = [int(c) for c in crop]
            total = (min(x2, rx - 1) - min(x1, rx - 1) + 1) * (min(y2, ry - 1) - min(y1, ry - 1) + 1)
        else:
            total = rx * ry
        state = {"done": 0}
        hl = host_lib()
        hl.psyh_bucket_base64.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_char_p, C.c_size_t]
        hl.psyh_bucket_base64.restype = C.c_size_t

        def sink(x0, y0, x1, y1, xyz):
            w, h = x1 - x0, y1 - y0
            state["done"] += w * h
            cap = ((w * h * 16 + 2) // 3) * 4 + 8
            buf = C.create_string_buffer(cap)
            xyz = np.ascontiguousarray(xyz, dtype=np.float32)
            n = hl.psyh_bucket_base64(xyz.ctypes.data, w, h, buf, cap)
            pct = state["done"] * 100.0 / max(total, 1)
            sys.stdout.write("DIV\n%.2f%%\n%d %d %d %d\n%s\nBUCKET_END\nDIV\n" % (pct, x0, y0, x1, y1, buf.raw[:n].decode()))
            sys.stdout.flush()
        return sink

    def render_shard(self, rank, world, max_samples_per_bucket=0, spp=None, seed=None, work_queue_key=0,
                     buckets_per_gpu=0):
        """One process per GPU: renders this rank's share of the buckets into the device image.  work_queue_key
        (the same non-zero value on every rank of the node, fresh per job) makes the ranks pull buckets from a shared
        counter; 0 deals them by sample count."""
        p = self._params(max_samples_per_bucket, None, spp, seed, False, True, False, buckets_per_gpu, work_queue_key)
        st = abi.PsyStats()
        self._check(self.lib.psy_render_image_shard(self._s, C.byref(p), rank, world, C.byref(st)),
                    "psy_render_image_shard")
        return stats_dict(st)

    def pinned_image(self):
        """A page-locked (res_y, res_x, 3) float32 image to pass as render(out=...) / fetch_image(out=...)."""
        rx, ry = self.resolution
        img = np.zeros((ry, rx, 3), dtype=np.float32)
        self._check(self.lib.psy_host_register(img.ctypes.data, img.nbytes), "psy_host_register")
        self._pinned_images = getattr(self, "_pinned_images", []) + [img]
        return img

    def device_image(self):
        ptr = C.c_void_p()
        n = C.c_size_t()
        self._check(self.lib.psy_device_image(self._s, C.byref(ptr), C.byref(n)), "psy_device_image")
        return ptr.value, n.value

    def fetch_image(self, out=None):
        rx, ry = self.resolution
        img = out if out is not None else np.zeros((ry, rx, 3), dtype=np.float32)
        self._check(self.lib.psy_fetch_image(self._s, img.ctypes.data), "psy_fetch_image")
        return img

    def render_bucket(self, x, y, w, h, spp=None, seed=None):
        xyz = np.zeros((h, w, 3), dtype=np.float32)
        st = abi.PsyStats()
        self._check(self.lib.psy_render_bucket(self._s, x, y, w, h, spp or self.spp, self.seed if seed is None else seed,
                                               xyz.ctypes.data, C.byref(st)), "psy_render_bucket")
        return xyz, stats_dict(st)

    def render_bucket_samples(self, x, y, w, h, spp=None, seed=None):
        spp = spp or self.spp
        out = np.zeros((h, w, spp, 4), dtype=np.float32)
        st = abi.PsyStats()
        self._check(self.lib.psy_render_bucket_samples(self._s, x, y, w, h, spp, self.seed if seed is None else seed,
                                                       out.ctypes.data, C.byref(st)), "psy_render_bucket_samples")
        return out, stats_dict(st)

    def last_trace_timing(self):
        ms = C.c_double()
        n = C.c_uint64()
        self._check(self.lib.psy_last_trace_timing(self._s, C.byref(ms), C.byref(n)), "psy_last_trace_timing")
        return ms.value, n.value
