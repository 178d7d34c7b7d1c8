// dscene.cuh — the scene as it lives in HBM (DESIGN.md "Data layout").
#pragma once
#include "../../../include/psy_cuda.h"
#include "dmath.cuh"

namespace psyd {

// One expanded instance path (a chain of instances from the root assembly to this instance).
// Paths are numbered breadth-first: children of assembly-instance path p are child_base + i.
struct DPath {
    uint32_t inst;        // global instance index
    uint32_t child_base;  // first child path if the instance is an assembly, else PSY_NONE
    uint32_t xf_base;     // composed world->local keys in path_xforms (units of transforms)
    uint32_t xf_count;    // 0 = no transform anywhere on the chain (space.is_empty())
};

// Flattened per-path record read on instance entry by the trace kernel (2 x uint4):
//   r0 = { kind | own_xform << 3 | xf_count << 8, xf_base, child_base, instance id }
//   r1 = mesh:     { BVH root ref, tri_base, tv_base, n_keys }
//        assembly: { BVH root ref, instance_base, assembly index, 0 }
//        light:    { light index, 0, 0, 0 }
// plus, at path_static[6*p .. 6*p+6], the single composed world->local key (3 rows) and its inverse
// (3 rows) for static paths, so that all loads of an entry have addresses known from `p` alone.
constexpr uint32_t PATH_MESH = 0, PATH_SPHERE_LIGHT = 1, PATH_RECT_LIGHT = 2, PATH_ASSEMBLY = 3, PATH_EMPTY = 4;

struct DMesh {
    uint32_t root, tri_base, tri_count, n_keys;
    uint32_t tv_base;    // first float4 of this mesh's gathered triangle vertices
    uint32_t vert_base;  // vertex-key base for the normals
    uint32_t has_normals, pad;
};

// k_trace_pool fetches BVH4 nodes through the texture pipe (tex1Dfetch on a linear uint4 texture over `nodes`):
// its LSU pipe is ~70 % busy with the scattered state / stack traffic, the TEX pipe is otherwise idle (+4 % on C5).
#ifndef PSY_POOL_NODE_TEX
#define PSY_POOL_NODE_TEX 1
#endif

// k_trace_pool's node store ("blob"): variable-size records addressed in 16-byte units, the time keys of a node
// directly behind its header, so that a motion node costs ONE memory round trip instead of header -> keys:
//   unit 0      child references (inner children re-encoded, see below; leaves as in the ABI)
//   unit 1      traversal_code, n_keys, 0, n_children
//   unit 2+6k.. bounds of time key k: xmin[4] xmax[4] ymin[4] ymax[4] zmin[4] zmax[4]
// An inner child reference carries what the fetch needs before the node is seen:
//   bit 31 = 0 | bits 28-30: min(n_keys, 8) - 1 | bits 0-27: unit offset of the node
// (key count code 7 = "8 or more keys": the count is then read from the header, a dependent fetch).
// A static node is 8 units = 128 bytes, exactly the ABI's PsyBvh4Node.
// Device-only colour kind (product build): an XYZ colour that does not vary over the shutter, with its Meng spectrum
// interpolated over the chromaticity grid ONCE at upload (the lookup is linear in the cell's spectra,
// meng.rs:254-272,319-320) — a = index of its 95-sample spectrum in DScene::meng_baked (as bits), b = x + y + z times
// whatever scale has been applied since.  Shading then costs one lerp per wavelength instead of the cell search and
// up to six spectra per wavelength.  Float-reassociated against the reference's order (~1e-7 relative), so the strict
// build does not bake.
constexpr uint32_t PSY_COLOR_BAKED = 3u;
constexpr uint32_t NODE2_OFF_MASK = 0x0FFFFFFFu;
constexpr uint32_t NODE2_KC_SHIFT = 28;

struct DScene {
    // acceleration structures
    const uint4* nodes;          // PsyBvh4Node as 8 x 16 B (k_trace, the one-ray-per-lane kernel)
    const uint4* nodes2;         // the blob described above (k_trace_pool)
    uint32_t root_bvh2;          // root assembly's BVH root as a blob reference
    unsigned long long nodes_tex; // the blob as a linear uint4 texture object (PSY_POOL_NODE_TEX)
    const float4* motion_bounds; // PsyBBox4 as 6 x float4 per key
    // geometry: per triangle (leaf order) 3*k float4: [v0 keys][v1 keys][v2 keys]; v0-key0 .w = original index
    const float4* tri_verts;
    const uint4* tris;     // (v0, v2, v1, original index)
    const float4* normals; // per vertex-key
    const DMesh* meshes;
    const PsyLight* lights;
    const float* light_shape;
    const PsyColor* light_colors;
    const PsyObject* objects;
    const PsyShader* shaders;
    const PsyAssembly* assemblies;
    const PsyInstance* instances;
    const PsyInstance* light_instances;
    const uint32_t* light_inst_local;  // light instance -> index inside its assembly's instance list
    const float4* xforms;       // instance transforms, 3 rows each
    const PsyLightNode* light_nodes;
    const PsyBBox* lt_bounds;
    // instance paths
    const DPath* paths;
    const float4* path_xforms;  // composed keys, 3 rows each
    const float4* path_inv;     // per path: inverse of the single composed key (valid when xf_count == 1)
    const uint4* path_recs;     // 2 per path, see PATH_* above
    const float4* path_static;  // 6 per path
    uint32_t root_assembly;
    uint32_t root_bvh;            // child reference of the root assembly's BVH4
    uint32_t root_instance_base;  // first instance of the root assembly
    uint32_t n_paths;
    // camera
    const float4* cam_xforms;
    const float* cam_tfovs;
    const float* cam_aperture;
    const float* cam_focus;
    uint32_t n_cam_xforms, n_cam_tfovs, n_cam_aperture, n_cam_focus;
    // world
    PsyColor background;
    const PsyDistantLight* distant_lights;
    const float* world_floats;
    uint32_t n_distant_lights;
    float world_light_energy_sum;  // fold of approximate_energy over world lights
    // spectral tables
    const float* meng_spectra;  // [186][95]
    const float* meng_baked;    // [n baked colours][95]: see PSY_COLOR_BAKED
    uint32_t res_x, res_y;
};

// Hit record written by the trace kernel (32 B)
struct HitRec {
    uint32_t kind;  // PSY_HIT_*
    float t;
    uint32_t prim;  // triangle: global triangle index; rect light: 0/1
    uint32_t path;
    float b0, b1, b2;
    uint32_t pad;
};

struct TraceCounters {
    unsigned long long node_visits, tri_tests, instance_visits, motion_node_visits;
};

}  // namespace psyd
