// host_api.cpp — C entry points of libpsy_host.so: the host side above the CUDA C ABI.
// (The reference's host is Rust; no Rust toolchain exists in this image, so the host that parses,
// builds acceleration structures and schedules buckets is restated in C++ — SURVEY.md F1.)
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>

#include "bvh_build.h"
#include "scene_ir.h"

using namespace psyh;

namespace {
thread_local std::string g_err;
struct HostScene {
    std::unique_ptr<FlatScene> flat;
};
template <typename F>
int guarded(F f) {
    try {
        f();
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    } catch (...) {
        g_err = "unknown error";
        return 1;
    }
}
}  // namespace

extern "C" {

const char* psyh_last_error(void) { return g_err.c_str(); }

// Parses .psy text and builds scene number `scene_index` (src/main.rs:241-253 loops over every
// top-level Scene node; the CLI calls this once per index).
int psyh_load_psy_string(const char* text, uint32_t scene_index, void** out) {
    return guarded([&] {
        std::vector<SceneIR> scenes = load_psy_text(text);
        if (scene_index >= scenes.size()) throw std::runtime_error("scene index out of range");
        auto* hs = new HostScene();
        hs->flat = flatten_scene(scenes[scene_index]);
        *out = hs;
    });
}

int psyh_count_scenes(const char* text, uint32_t* count) {
    return guarded([&] { *count = (uint32_t)load_psy_text(text).size(); });
}

int psyh_load_psy_file(const char* path, uint32_t scene_index, void** out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) {
        g_err = std::string("cannot open ") + path;
        return 1;
    }
    std::stringstream ss;
    ss << f.rdbuf();
    std::string text = ss.str();
    return psyh_load_psy_string(text.c_str(), scene_index, out);
}

int psyh_synth_displaced_grid(uint32_t quads_per_side, uint32_t n_vert_keys, uint32_t n_xform_keys, uint32_t res_x,
                              uint32_t res_y, uint32_t spp, uint32_t seed, void** out) {
    return guarded([&] {
        SceneIR s = synth_displaced_grid(quads_per_side, n_vert_keys, n_xform_keys, res_x, res_y, spp, seed);
        auto* hs = new HostScene();
        hs->flat = flatten_scene(s);
        *out = hs;
    });
}

int psyh_synth_instanced(uint32_t asset_quads, uint32_t n_assets, uint32_t n_clusters, uint32_t inst_per_cluster,
                         uint32_t lights_per_cluster, uint32_t cluster_xform_keys, uint32_t res_x, uint32_t res_y,
                         uint32_t spp, uint32_t seed, uint32_t motion_pct, void** out) {
    return guarded([&] {
        SceneIR s = synth_instanced(asset_quads, n_assets, n_clusters, inst_per_cluster, lights_per_cluster,
                                    cluster_xform_keys, res_x, res_y, spp, seed, motion_pct ? motion_pct : 100);
        auto* hs = new HostScene();
        hs->flat = flatten_scene(s);
        *out = hs;
    });
}

// registers the builder of mesh BVH topologies (psy_build_mesh_topology of the CUDA library) for meshes
// of at least `min_tris` triangles; fn == NULL restores the host SAH builder for everything.
void psyh_set_mesh_topology_builder(void* fn, void* ctx, uint32_t min_tris) {
    set_mesh_topology_builder(reinterpret_cast<MeshTopologyFn>(fn), ctx, min_tris);
}
double psyh_take_mesh_topology_seconds(uint32_t* meshes_built) { return take_mesh_topology_seconds(meshes_built); }
// ... and of whole mesh BVH4s (psy_mesh_bvh_build / psy_mesh_bvh_take); NULLs unregister.  Tried before the topology
// builder; its time and mesh count are reported by psyh_take_mesh_topology_seconds as well.
void psyh_set_mesh_bvh_builder(void* build, void* take, void* ctx, uint32_t min_tris) {
    set_mesh_bvh_builder(reinterpret_cast<MeshBvhBuildFn>(build), reinterpret_cast<MeshBvhTakeFn>(take), ctx, min_tris);
}

// ---- test hooks: the reference's own unit tests for the parser and the build helpers are restated against these
// (tests/test_parse_semantics.py, tests/test_host_and_abi.py)
long psyh_test_datatree_dump(const char* text, char* out, size_t cap) {
    std::string r;
    if (guarded([&]() { r = datatree_dump(text); }) != 0) return -1;
    if (out && cap) {
        const size_t n = std::min(cap - 1, r.size());
        std::memcpy(out, r.data(), n);
        out[n] = 0;
    }
    return (long)r.size();
}
size_t psyh_test_parse_floats(const char* text, float* out, size_t cap, size_t* rest) {
    return test_parse_floats(text, out, cap, rest);
}
size_t psyh_test_parse_uints(const char* text, uint32_t* out, size_t cap, size_t* rest) {
    return test_parse_uints(text, out, cap, rest);
}
void psyh_test_quick_select(int32_t* list, size_t len, size_t n) {
    quick_select_ref(list, len, n, [](int32_t a, int32_t b) { return a < b; });
}
uint64_t psyh_test_log2_64(uint64_t n) { return log2_64(n); }

const PsySceneDesc* psyh_scene_desc(void* scene) { return &static_cast<HostScene*>(scene)->flat->desc; }
double psyh_build_seconds(void* scene) { return static_cast<HostScene*>(scene)->flat->build_seconds; }
const char* psyh_output_path(void* scene) { return static_cast<HostScene*>(scene)->flat->output_path.c_str(); }
void psyh_scene_set_render(void* scene, uint32_t res_x, uint32_t res_y, uint32_t spp, uint32_t seed) {
    PsySceneDesc& d = static_cast<HostScene*>(scene)->flat->desc;
    if (res_x) d.res_x = res_x;
    if (res_y) d.res_y = res_y;
    if (spp) d.spp = spp;
    d.seed = seed;
}
void psyh_scene_free(void* scene) { delete static_cast<HostScene*>(scene); }

// hilbert::d2xy (src/hilbert.rs:47-63)
void psyh_hilbert_d2xy(uint32_t d, uint32_t* xo, uint32_t* yo) {
    uint32_t x = 0, y = 0, t = d;
    for (uint32_t s = 1; s < (1u << 16); s <<= 1) {
        uint32_t rx = 1 & (t >> 1);
        uint32_t ry = 1 & (t ^ rx);
        if (ry == 0) {
            if (rx == 1) {
                x = (s - 1) - x;
                y = (s - 1) - y;
            }
            uint32_t tmp = x;
            x = y;
            y = tmp;
        }
        x += s * rx;
        y += s * ry;
        t >>= 2;
    }
    *xo = x;
    *yo = y;
}

// Bucket list of Renderer::render (src/renderer.rs:138-180): Hilbert-ordered (x,y,w,h) jobs.
// Returns the number of jobs; writes at most `cap` of them (4 uint32 each).
uint32_t psyh_bucket_list(uint32_t width, uint32_t height, uint32_t start_x, uint32_t start_y, uint32_t bucket_w,
                          uint32_t bucket_h, uint32_t* out, uint32_t cap) {
    uint32_t cx = width / bucket_w + 1, cy = height / bucket_h + 1;
    uint32_t larger = cx > cy ? cx : cy;
    uint32_t pow2 = 1;
    while (pow2 < larger) pow2 <<= 1;
    uint64_t bucket_n = (uint64_t)pow2 * pow2;
    uint32_t n = 0;
    for (uint64_t d = 0; d < bucket_n; ++d) {
        uint32_t bx, by;
        psyh_hilbert_d2xy((uint32_t)d, &bx, &by);
        uint64_t x = (uint64_t)bx * bucket_w, y = (uint64_t)by * bucket_h;
        if (x >= width || y >= height) continue;
        uint32_t w = (uint32_t)std::min<uint64_t>(bucket_w, width - x);
        uint32_t h = (uint32_t)std::min<uint64_t>(bucket_h, height - y);
        if (w == 0 || h == 0) continue;
        if (n < cap) {
            out[4 * n + 0] = start_x + (uint32_t)x;
            out[4 * n + 1] = start_y + (uint32_t)y;
            out[4 * n + 2] = w;
            out[4 * n + 3] = h;
        }
        ++n;
    }
    return n;
}

}  // extern "C"
