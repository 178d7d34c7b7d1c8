/* psy_cuda.h — C ABI of the B200 render core for Psychopath.
 *
 * This is the drop-in boundary.  The reference has no FFI seam; the render core
 * sits behind three Rust-internal call sites, and each entry point below names the one it replaces:
 *
 *   psy_scene_create   <- the immutable scene graph handed to `Renderer` (src/renderer.rs:32-38),
 *                         i.e. what `parse_scene` (src/parse/psy.rs:95-200) builds: Camera, World,
 *                         Assembly{instances, xforms, shaders, objects, BVH4, LightTree}
 *                         (src/scene/assembly.rs:20-40).  Passed as flat SoA arrays (PsySceneDesc).
 *   psy_trace_rays     <- `Tracer::trace(&mut RayBatch) -> &[SurfaceIntersection]` (src/tracer.rs:34-37)
 *   psy_render_bucket  <- body of `Renderer::render_job` for one BucketJob (src/renderer.rs:239-310)
 *   psy_render_image   <- `Renderer::render(spb, crop, threads, blender) -> (Image, RenderStats)`
 *                         (src/renderer.rs:76-191)
 *   PsyStats           <- `RenderStats` (src/renderer.rs:40-49)
 *
 * Conventions: every function returns int32_t status (PSY_OK == 0), never throws or aborts across
 * the boundary; `psy_last_error()` returns a thread-local message.  The caller owns all host memory;
 * `psy_scene_create` copies what it needs to the device.  Handles are opaque.  No torch types.
 * There is no CPU fallback: without a CUDA device `psy_init` fails with PSY_ERR_NO_DEVICE.
 */
#ifndef PSY_CUDA_H
#define PSY_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSY_ABI_VERSION 2

/* status codes */
#define PSY_OK 0
#define PSY_ERR_INVALID 1      /* bad argument / malformed scene description */
#define PSY_ERR_NO_DEVICE 2    /* no usable CUDA device (there is no CPU fallback) */
#define PSY_ERR_CUDA 3         /* a CUDA runtime call failed; see psy_last_error() */
#define PSY_ERR_LIMIT 4        /* scene exceeds a compiled-in limit (stack depth, path count) */
#define PSY_ERR_NCCL 5

#define PSY_NONE 0xFFFFFFFFu

/* ---- BVH4 child reference (32 bit) ------------------------------------------------------------
 * Mirrors `BVH4Node::{Internal,Leaf}` (src/accel/bvh4.rs:45-55), with the leaf folded into the
 * parent's child slot so that a leaf costs no extra fetch:
 *   PSY_NONE                      empty slot / empty tree
 *   bit31 = 0                     index of an internal node in PsySceneDesc.nodes
 *   bit31 = 1                     leaf: bits 28..30 = object count (1..7), bits 0..27 = first object
 *                                 (triangle index for a mesh BVH, instance index for an assembly BVH)
 */
#define PSY_CHILD_LEAF_BIT 0x80000000u
#define PSY_CHILD_LEAF(start, count) (PSY_CHILD_LEAF_BIT | ((uint32_t)(count) << 28) | (uint32_t)(start))
#define PSY_CHILD_IS_LEAF(c) (((c) & PSY_CHILD_LEAF_BIT) != 0u && (c) != PSY_NONE)
#define PSY_CHILD_LEAF_START(c) ((c) & 0x0FFFFFFFu)
#define PSY_CHILD_LEAF_COUNT(c) (((c) >> 28) & 7u)
#define PSY_MAX_LEAF_START 0x0FFFFFFEu

/* 3x4 affine transform, row-major: row i = (X_i, Y_i, Z_i, T_i) where X,Y,Z are glam's column
 * axes and T the translation (sub_crates/math3d/src/transform.rs:12).  A point transforms as
 * out_i = ((X_i*p.x + Y_i*p.y) + Z_i*p.z) + T_i  — the op order of glam's Affine3A. */
typedef struct PsyXform {
    float r[3][4];
} PsyXform;

/* Axis-aligned box (src/bbox.rs:16-20) */
typedef struct PsyBBox {
    float min[3];
    float max[3];
} PsyBBox;

/* One time key of 4-wide bounds (src/bbox4.rs:18-22): xmin[4],xmax[4],ymin[4],ymax[4],zmin[4],zmax[4].
 * Empty child slots hold +inf/-inf (src/accel/bvh4.rs:290-299). */
typedef struct PsyBBox4 {
    float v[6][4];
} PsyBBox4;

/* BVH4 internal node, 128 bytes, 16-byte aligned. */
typedef struct PsyBvh4Node {
    uint32_t child[4];       /* child references, see above */
    uint32_t traversal_code; /* 0..47, index into TRAVERSAL_TABLE[ray_code] (sub_crates/bvh_order) */
    uint32_t n_keys;         /* number of time keys of the bounds (>=1) */
    uint32_t key_base;       /* if n_keys > 1: first key in PsySceneDesc.motion_bounds */
    uint32_t n_children;     /* 2..4 */
    PsyBBox4 bounds;         /* the bounds when n_keys == 1; key 0 otherwise */
} PsyBvh4Node;

/* Colour (src/color.rs:47-60) */
#define PSY_COLOR_XYZ 0
#define PSY_COLOR_BLACKBODY 1
#define PSY_COLOR_TEMPERATURE 2
typedef struct PsyColor {
    uint32_t kind;
    float a, b, c; /* XYZ: x,y,z.  Blackbody/Temperature: a = kelvin, b = factor */
} PsyColor;

/* Surface shader (src/shading/mod.rs:28-40) */
#define PSY_SHADER_EMIT 0
#define PSY_SHADER_LAMBERT 1
#define PSY_SHADER_GGX 2
typedef struct PsyShader {
    uint32_t kind;
    PsyColor color;
    float roughness;
    float fresnel;
} PsyShader;

/* Triangle mesh (src/surface/triangle_mesh.rs:19-26).  Vertices are stored with the time keys of
 * one vertex contiguous; `tris` entries are (v0, v2, v1, original_index) in BVH leaf order. */
typedef struct PsyMesh {
    uint32_t root;        /* child reference of the BVH4 root (may itself be a leaf) */
    uint32_t tri_base;    /* first entry in PsySceneDesc.tris */
    uint32_t tri_count;
    uint32_t n_keys;      /* time samples per vertex */
    uint32_t vert_base;   /* first vertex-key in PsySceneDesc.verts (units of 3 floats) */
    uint32_t vert_count;  /* vertices (not vertex-keys) */
    uint32_t has_normals; /* normals stored like verts, in PsySceneDesc.normals at vert_base */
    uint32_t bounds_base; /* object bounds keys in PsySceneDesc.obj_bounds */
    uint32_t bounds_count;
    uint32_t bvh_depth;   /* BVH4 depth: (binary depth / 2) + 1 (src/accel/bvh4.rs:87) */
} PsyMesh;

/* Surface lights (src/light/sphere_light.rs:27-31, rectangle_light.rs:24-28) */
#define PSY_LIGHT_SPHERE 0
#define PSY_LIGHT_RECT 1
typedef struct PsyLight {
    uint32_t kind;
    uint32_t n_shape;    /* time keys of radius (sphere) or dimensions (rect) */
    uint32_t shape_base; /* into PsySceneDesc.light_shape: 1 float/key (sphere), 2 floats/key (rect) */
    uint32_t n_colors;
    uint32_t color_base; /* into PsySceneDesc.light_colors */
    uint32_t bounds_base;
    uint32_t bounds_count;
} PsyLight;

/* World distant-disk light (src/light/distant_disk_light.rs:16-20) */
typedef struct PsyDistantLight {
    uint32_t n_radii, radii_base;  /* into world_floats */
    uint32_t n_dirs, dir_base;     /* into world_floats, 3 floats per key */
    uint32_t n_colors, color_base; /* into light_colors */
} PsyDistantLight;

/* Object table entry (src/scene/assembly.rs:397-401) */
#define PSY_OBJECT_MESH 0
#define PSY_OBJECT_LIGHT 1
typedef struct PsyObject {
    uint32_t kind;
    uint32_t index; /* into meshes or lights */
} PsyObject;

/* Instance (src/scene/assembly.rs:403-417) */
#define PSY_INSTANCE_OBJECT 0
#define PSY_INSTANCE_ASSEMBLY 1
typedef struct PsyInstance {
    uint32_t kind;
    uint32_t data_index;   /* global index into objects or assemblies */
    uint32_t shader_index; /* global index into shaders, or PSY_NONE */
    uint32_t id;           /* `Instance.id`: position before the BVH build reordered the list */
    uint32_t xform_base;   /* into PsySceneDesc.xforms */
    uint32_t xform_count;  /* 0 = untransformed */
} PsyInstance;

/* Light tree node (src/accel/light_tree.rs:27-39), arity <= 8 */
typedef struct PsyLightNode {
    uint32_t child_base;  /* into light_nodes; children are contiguous */
    uint32_t child_count; /* 0 for a leaf */
    uint32_t light_index; /* leaf: index into the assembly's light_instances (relative) */
    uint32_t bounds_base; /* into lt_bounds */
    uint32_t bounds_count;
    float energy;
} PsyLightNode;

/* Assembly (src/scene/assembly.rs:20-40) */
typedef struct PsyAssembly {
    uint32_t root; /* child reference of the object BVH4 root, PSY_NONE when empty */
    uint32_t instance_base, instance_count;
    uint32_t light_instance_base, light_instance_count; /* into light_instances */
    uint32_t light_root;                                /* into light_nodes, PSY_NONE when no lights */
    float light_energy;                                 /* LightTree::approximate_energy */
    uint32_t bounds_base, bounds_count;                 /* into obj_bounds */
    uint32_t bvh_depth;
} PsyAssembly;

/* Camera (src/camera.rs:12-19); tfov = sin(fov/2)/cos(fov/2) precomputed as in Camera::new */
typedef struct PsyCamera {
    const PsyXform* xforms;
    uint32_t n_xforms;
    const float* tfovs;
    uint32_t n_tfovs;
    const float* aperture_radii;
    uint32_t n_aperture_radii;
    const float* focus_distances;
    uint32_t n_focus_distances;
} PsyCamera;

typedef struct PsySceneDesc {
    uint32_t abi_version;
    /* render settings (src/renderer.rs:32-38) */
    uint32_t res_x, res_y, spp, seed;

    PsyCamera camera;

    /* world (src/scene/world.rs) */
    PsyColor background;
    const PsyDistantLight* distant_lights;
    uint32_t n_distant_lights;
    const float* world_floats;
    uint32_t n_world_floats;

    /* acceleration structures */
    const PsyBvh4Node* nodes;
    uint32_t n_nodes;
    const PsyBBox4* motion_bounds;
    uint32_t n_motion_bounds;

    /* geometry */
    const PsyMesh* meshes;
    uint32_t n_meshes;
    const uint32_t* tris; /* 4 per triangle */
    uint32_t n_tris;
    const float* verts; /* 3 per vertex-key */
    const float* normals; /* same indexing as verts; only valid where has_normals */
    uint32_t n_vert_keys;

    const PsyLight* lights;
    uint32_t n_lights;
    const float* light_shape;
    uint32_t n_light_shape;
    const PsyColor* light_colors;
    uint32_t n_light_colors;

    const PsyObject* objects;
    uint32_t n_objects;
    const PsyShader* shaders;
    uint32_t n_shaders;

    /* assemblies */
    const PsyAssembly* assemblies;
    uint32_t n_assemblies;
    uint32_t root_assembly;
    const PsyInstance* instances;
    uint32_t n_instances;
    const PsyInstance* light_instances;
    uint32_t n_light_instances;
    const PsyXform* xforms;
    uint32_t n_xforms;
    const PsyLightNode* light_nodes;
    uint32_t n_light_nodes;
    const PsyBBox* lt_bounds;
    uint32_t n_lt_bounds;
    const PsyBBox* obj_bounds;
    uint32_t n_obj_bounds;
} PsySceneDesc;

/* Probe ray (src/ray.rs:15-21).  flags bit0 = occlusion ray (src/ray.rs:9). */
#define PSY_RAY_OCCLUSION 1u
typedef struct PsyRay {
    float orig[3];
    float dir[3];
    float time;
    float max_t;
    uint32_t flags;
} PsyRay;

/* Result of a probe ray.  The reference's SurfaceIntersection carries no ids (src/surface/mod.rs:70-92);
 * they are added here as SURVEY.md F5 describes. */
#define PSY_HIT_MISS 0
#define PSY_HIT_TRIANGLE 1
#define PSY_HIT_SPHERE_LIGHT 2
#define PSY_HIT_RECT_LIGHT 3
#define PSY_HIT_OCCLUDED 4
typedef struct PsyHit {
    uint32_t kind;
    float t;
    uint32_t prim_id; /* triangle: original triangle index; rect light: 0/1; sphere: 0 */
    uint32_t inst_id; /* Instance.id of the leaf instance inside its assembly */
    uint32_t path_id; /* index of the expanded instance path (unique per nested instance chain) */
    float b0, b1, b2; /* barycentrics (triangles / rect light) */
} PsyHit;

/* Optional geometric detail of a closest hit: SurfaceIntersectionData (src/surface/mod.rs:81-92) */
typedef struct PsyHitGeo {
    float pos[3];
    float pos_err;
    float nor[3];
    float sample_pdf;
    float nor_g[3];
    uint32_t closure_kind; /* PSY_SHADER_* of the closure at the hit */
} PsyHitGeo;

/* RenderStats (src/renderer.rs:40-49) + device counters */
typedef struct PsyStats {
    double trace_time;                  /* seconds in extend/shadow kernels */
    double initial_ray_generation_time; /* seconds in the generate kernel */
    double ray_generation_time;         /* seconds in shade kernels */
    double sample_writing_time;         /* seconds in the film kernel */
    double total_time;
    uint64_t ray_count;         /* rays traced (extend + shadow), like Tracer::rays_traced */
    uint64_t accel_node_visits; /* internal node visits (counters build only, else 0) */
    uint64_t tri_tests;         /* triangle tests (counters build only, else 0) */
    uint64_t instance_visits;   /* instance entries (counters build only, else 0) */
    uint64_t motion_node_visits; /* node visits that lerped two keys (counters) */
    uint64_t samples;           /* w*h*spp */
    uint64_t kernel_launches;   /* number of kernels this call launched */
    uint64_t buckets;           /* buckets this call rendered (with several GPUs: how the work was shared) */
} PsyStats;

typedef struct PsyRenderParams {
    uint32_t crop_enabled;
    uint32_t crop_x1, crop_y1, crop_x2, crop_y2; /* inclusive, as `--crop` (src/main.rs:104-119) */
    uint32_t spp;                 /* 0 = scene's */
    uint32_t seed_override_enabled;
    uint32_t seed;
    uint32_t max_samples_per_bucket; /* `--spb`; on the GPU this is the path-pool size (paths in flight, ~390 B each);
                                        0 = default: the whole image up to 64 Mi paths, ~6 buckets per GPU when the
                                        image is tiled across GPUs, 4 Mi when buckets are streamed to a callback */
    uint32_t collect_counters;    /* 1 = run the counters build of the trace kernel */
    uint32_t device_resident_only; /* 1 = leave the image on the device, skip the D2H copy (bench `value`) */
    uint32_t strict_film_order;   /* 1 = add every bucket's samples to the film in the reference's path order after its
                                     swap compaction (src/renderer.rs:286-310): the image then depends on the bucket
                                     list exactly as the reference's does, so max_samples_per_bucket is the reference's
                                     `--spb` (0 = its default 4096) and a bucket holds at most 65536 paths (src/ray.rs:7) */
    uint32_t buckets_per_gpu;     /* image shared between GPUs, max_samples_per_bucket == 0: buckets per GPU (0 = 12) */
    uint32_t work_queue_key;      /* psy_render_image_shard: non-zero = the ranks of this node pull their buckets from
                                     a counter in POSIX shared memory named by this key (any value all ranks agree on,
                                     fresh per job) instead of being dealt a fixed share */
} PsyRenderParams;

typedef void (*psy_bucket_cb)(void* user, uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1,
                              const float* xyz /* (x1-x0)*(y1-y0)*3, row-major */);

typedef struct PsyContext PsyContext;
typedef struct PsyScene PsyScene;

const char* psy_last_error(void);
int32_t psy_abi_version(void);
/* bit 0: the library was built with -DPSY_STRICT=1 (libpsy_cuda_strict.so): sin/cos/acos/asin/log come from
 * include/psy_detmath.h instead of libdevice, like the oracle's strict mode (parity tests only; ~2x slower shading) */
#define PSY_BUILD_STRICT_MATH 1u
uint32_t psy_build_flags(void);

/* device_ids == NULL -> device 0.  One context per process per GPU set.  With n_devices > 1 the context drives all of
 * them from this process: scenes are replicated, psy_render_image shares the buckets between the GPUs (one host thread
 * each, pulling from a common counter) and gathers the tiles onto the first device with NCCL (ncclCommInitAll here,
 * one grouped ncclReduce per frame); psy_trace_rays / psy_render_bucket use the first device. */
int32_t psy_init(const int* device_ids, int n_devices, PsyContext** out);
int32_t psy_shutdown(PsyContext* ctx);

int32_t psy_scene_create(PsyContext* ctx, const PsySceneDesc* desc, PsyScene** out);
int32_t psy_scene_destroy(PsyScene* scene);

/* Host-buffer probe trace: copies rays H2D, traces, copies hits D2H.  `geo` may be NULL. */
int32_t psy_trace_rays(PsyScene* scene, const PsyRay* rays, size_t n, PsyHit* hits, PsyHitGeo* geo);

/* One bucket (the body of render_job).  xyz_out: w*h*3 floats, row-major, the final per-pixel mean. */
int32_t psy_render_bucket(PsyScene* scene, uint32_t x, uint32_t y, uint32_t w, uint32_t h, uint32_t spp,
                          uint32_t seed, float* xyz_out, PsyStats* stats);

/* Per-sample radiance for parity tests: out[(py*w+px)*spp+si][4] = LightPath.color after the path
 * terminates (src/renderer.rs:387), before the film transform. */
int32_t psy_render_bucket_samples(PsyScene* scene, uint32_t x, uint32_t y, uint32_t w, uint32_t h,
                                  uint32_t spp, uint32_t seed, float* samples_out, PsyStats* stats);

/* Whole image.  xyz_out: res_x*res_y*3 floats (row 0 = top, like Image in src/image.rs); may be NULL
 * when params->device_resident_only.  on_bucket may be NULL. */
int32_t psy_render_image(PsyScene* scene, const PsyRenderParams* params, float* xyz_out, PsyStats* stats,
                         psy_bucket_cb on_bucket, void* user);

/* One process per GPU (torchrun): renders this rank's share of the bucket list into a zero-initialised device image
 * that the caller gathers with a collective (psy_device_image gives the pointer).  With params->work_queue_key the ranks
 * pull buckets from a shared counter (largest buckets first) until the list is empty; without it the buckets are dealt
 * by sample count, largest first, each to the rank with the least so far, the same on every rank.  Every rank must
 * call it once per frame with the same arguments. */
int32_t psy_render_image_shard(PsyScene* scene, const PsyRenderParams* params, uint32_t rank, uint32_t world,
                               PsyStats* stats);
int32_t psy_device_image(PsyScene* scene, void** dev_ptr, size_t* n_floats);

/* SURVEY.md N3 — the GPU build of a mesh BVH, behind a flag (`psychopath --gpu-build`, api.enable_gpu_build()).
 * Replaces the ORDERING and TOPOLOGY half of BVHBase::from_objects (src/accel/bvh_base.rs:57-215): triangles sorted along
 * a 63-bit Morton curve of their centroids at shutter time 0.5, binary tree over the sorted triangles by parallel
 * locally-ordered clustering (PLOC; the environment variable PSY_GPU_BUILDER=lbvh selects Karras' radix tree instead),
 * subtrees of at most `tris_per_leaf` (<= 3) triangles as leaves, triangles in the leaf order of the finished tree.
 * Host arrays in and out:
 *   verts   n_verts * n_keys * 3 floats, the time keys of a vertex contiguous (PsySceneDesc.verts layout)
 *   tri_idx 3 vertex indices per triangle
 *   order   n_tris: the triangle that ends up at each position of the leaf-ordered list
 *   left / right / axis   room for n_tris - 1 entries; *n_internal are written, one per internal node.  A child >= 0
 *           is an internal node; a child < 0 is a leaf: code = ~child, positions [code >> 2, (code >> 2) + (code & 3))
 *           of `order`
 *   root    reference of the root in the same encoding (0, or a leaf when the mesh has <= tris_per_leaf triangles);
 *   depth   depth of the deepest leaf (root = 0)
 * The host derives time-sampled bounds, the 1.4x union rule and the BVH4 collapse from it exactly as for its SAH
 * trees (psyh_set_mesh_topology_builder in libpsy_host.so takes this function's address). */
int32_t psy_build_mesh_topology(PsyContext* ctx, const float* verts, uint32_t n_verts, uint32_t n_keys, const uint32_t* tri_idx,
                                uint32_t n_tris, uint32_t tris_per_leaf, uint32_t* order, int32_t* left, int32_t* right,
                                uint8_t* axis, uint32_t* n_internal, int32_t* root, uint32_t* depth);

/* SURVEY.md N3, the whole of it — a mesh's BVH4 built on the device, replacing BVHBase::from_objects
 * (src/accel/bvh_base.rs:57-215) AND BVH4::construct_from_base (src/accel/bvh4.rs:186-358) for that mesh: the
 * topology above, then per-node time-sampled bounds (bottom-up, bvh_base.rs:119-132's 1.4x union rule at every node),
 * the BVH2 -> BVH4 collapse with its traversal codes, and the time-sampled child bounds of every BVH4 node.  Same
 * arithmetic as the host builder: given the same topology the nodes are the host's bit for bit, in another order
 * (tests/test_gpu_build.py).
 *   psy_mesh_bvh_build  inputs as psy_build_mesh_topology; `obj_base` is added to the leaves' triangle positions
 *                       (PsyMesh.tri_base), `node_base` to the references of internal children and `key_base` to the
 *                       key_base of nodes with time keys: where this mesh's nodes / keys will start in
 *                       PsySceneDesc.nodes / .motion_bounds, so that psy_mesh_bvh_take can write them in place.
 *                       The result stays on the device behind *out; *info says how much room the
 *                       caller's arrays need.  Meshes of <= tris_per_leaf triangles or more than 8 time keys are refused
 *                       (PSY_ERR_INVALID): the host builds those.
 *   psy_mesh_bvh_take   copies the result out and frees the handle (also on error; all-NULL arrays just free it):
 *                       order[n_tris]; nodes[n_nodes] (the first is the root); motion[n_motion];
 *                       root_bounds[n_root_bounds]. */
typedef struct PsyMeshBvh PsyMeshBvh;
typedef struct PsyMeshBvhInfo {
    uint32_t n_nodes;       /* BVH4 nodes */
    uint32_t n_motion;      /* PsyBBox4 keys of the nodes with n_keys > 1 */
    uint32_t n_root_bounds; /* keys of the whole mesh's bounds (1 or n_keys) */
    uint32_t root;          /* child reference of the root (= node_base) */
    uint32_t depth;         /* BVH4 depth as construct_from_base reports it: binary depth / 2 + 1 */
    uint32_t binary_depth;  /* depth of the deepest leaf of the binary tree (root = 0) */
} PsyMeshBvhInfo;
int32_t psy_mesh_bvh_build(PsyContext* ctx, const float* verts, uint32_t n_verts, uint32_t n_keys, const uint32_t* tri_idx,
                           uint32_t n_tris, uint32_t tris_per_leaf, uint32_t obj_base, uint32_t node_base, uint32_t key_base,
                           PsyMeshBvh** out, PsyMeshBvhInfo* info);
int32_t psy_mesh_bvh_take(PsyMeshBvh* bvh, uint32_t* order, PsyBvh4Node* nodes, PsyBBox4* motion, PsyBBox* root_bounds);

/* Page-locks / releases caller memory (cudaHostRegister): scene arrays that are uploaded every frame, result images
 * that are fetched every frame.  Optional — every entry point also takes pageable memory, at a lower copy rate. */
int32_t psy_host_register(void* ptr, size_t bytes);
int32_t psy_host_unregister(void* ptr);

/* The work counter behind work_queue_key, usable on its own by a host that schedules buckets itself (the reference's
 * Renderer::render, src/renderer.rs:138-180, with psy_render_bucket as the body of render_job): every process of the
 * node that opens the same key shares 64 counters, one per frame in flight; psy_work_queue_pull returns 0, 1, 2, ...
 * across all of them.  Frames must be used in increasing order and at most 32 apart between processes. */
typedef struct PsyWorkQueue PsyWorkQueue;
int32_t psy_work_queue_open(uint32_t key, PsyWorkQueue** out);
int32_t psy_work_queue_pull(PsyWorkQueue* q, uint64_t frame, uint64_t* index);
int32_t psy_work_queue_close(PsyWorkQueue* q);
int32_t psy_scene_stream(PsyScene* scene, void** cuda_stream);
int32_t psy_fetch_image(PsyScene* scene, float* xyz_out);

/* Average duration (ms) of the trace kernel launches of the last render call, measured with CUDA
 * events on the library's stream, and the number of trace launches. */
int32_t psy_last_trace_timing(PsyScene* scene, double* total_ms, uint64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* PSY_CUDA_H */
