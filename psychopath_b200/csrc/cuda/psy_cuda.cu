// psy_cuda.cu — implementation of the C ABI in include/psy_cuda.h: scene upload, the wavefront
// driver and the probe-ray entry.  Everything that computes runs in the CUDA kernels of
// kernels.cuh; there is no CPU fallback (psy_init fails without a device).
#include <dlfcn.h>
#include <fcntl.h>
#include <nccl.h>  // types and enums only: the library is bound at run time (NcclApi below)
#include <sys/mman.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../host/hmath.h"
#include "kernels.cuh"
#include "dbuild.cuh"

#define MENG_TABLE_QUAL static const
#include "../../../tables/meng_tables.inc"
#define SOBOL_TABLE_QUAL static const
#include "../../../tables/sobol_tables.inc"

using namespace psyd;

namespace {

thread_local std::string g_err;
int32_t fail(int32_t code, const std::string& msg) {
    g_err = msg;
    return code;
}
#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(PSY_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));               \
    } while (0)

// Freed device buffers are kept (per device, up to 8 GB) and handed out again: a renderer that uploads a scene per
// frame (Blender, bench.py's end-to-end arm) otherwise pays for ~35 cudaMalloc per create and, now and then, for a
// cudaFree that takes hundreds of milliseconds when the driver returns a large block to the system.
struct BufCache {
    std::mutex mu;
    std::map<int, std::multimap<size_t, void*>> free_;
    std::map<int, size_t> held;
    static constexpr size_t LIMIT = 8ull << 30, BIGGEST = 2ull << 30;
    void* take(int dev, size_t n, size_t& cap) {
        std::lock_guard<std::mutex> g(mu);
        auto& m = free_[dev];
        auto it = m.lower_bound(n);
        if (it == m.end() || it->first > 2 * n + (1u << 20)) return nullptr;
        void* p = it->second;
        cap = it->first;
        held[dev] -= cap;
        m.erase(it);
        return p;
    }
    bool put(int dev, void* p, size_t cap) {
        std::lock_guard<std::mutex> g(mu);
        if (cap > BIGGEST || held[dev] + cap > LIMIT) return false;
        free_[dev].emplace(cap, p);
        held[dev] += cap;
        return true;
    }
    void drop(int dev) {  // psy_shutdown
        std::lock_guard<std::mutex> g(mu);
        for (auto& e : free_[dev]) cudaFree(e.second);
        free_[dev].clear();
        held[dev] = 0;
    }
};
BufCache& buf_cache() {
    static BufCache c;
    return c;
}

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;  // what the owner asked for
    size_t cap = 0;    // what the allocation holds
    int dev = -1;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), bytes(o.bytes), cap(o.cap), dev(o.dev) { o.p = nullptr, o.bytes = 0, o.cap = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) {
            release();
            p = o.p, bytes = o.bytes, cap = o.cap, dev = o.dev;
            o.p = nullptr, o.bytes = 0, o.cap = 0;
        }
        return *this;
    }
    ~DevBuf() { release(); }
    void release() {
        if (p && !buf_cache().put(dev, p, cap)) cudaFree(p);
        p = nullptr, bytes = 0, cap = 0;
    }
    cudaError_t alloc(size_t n) {
        const size_t want = n ? n : 16;
        if (p && cap >= want && cap <= 2 * want + (1u << 20)) {
            bytes = n;
            return cudaSuccess;
        }
        release();
        cudaGetDevice(&dev);
        p = buf_cache().take(dev, want, cap);
        if (!p) {
            cap = (want + 65535) & ~(size_t)65535;
            cudaError_t e = cudaMalloc(&p, cap);
            if (e != cudaSuccess) {
                // make room: give the cached buffers back to the driver and try once more
                cudaGetLastError();
                buf_cache().drop(dev);
                e = cudaMalloc(&p, cap);
            }
            if (e != cudaSuccess) {
                p = nullptr, cap = 0;
                return e;
            }
        }
        bytes = n;
        return cudaSuccess;
    }
    template <typename T>
    cudaError_t upload(const T* src, size_t count, cudaStream_t s) {
        cudaError_t e = alloc(count * sizeof(T));
        if (e != cudaSuccess) return e;
        if (count) return cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyHostToDevice, s);
        return cudaSuccess;
    }
    template <typename T>
    T* as() const {
        return reinterpret_cast<T*>(p);
    }
};

}  // namespace

// The path pool and the trace scratch area do not depend on the scene (25 GB at the default size): a destroyed
// scene leaves them with its context, and the next scene that needs a pool no larger takes them over instead of
// allocating — rendering frame after frame (Blender, bench.py's end-to-end arm) pays for cudaMalloc / cudaFree once.
struct PoolCache {
    std::vector<DevBuf> bufs;
    size_t cap = 0;
    uint32_t trace_blocks = 0;
};

// NCCL, bound with dlopen so that the library loads (and one-GPU rendering works) on a machine without it.
// Used by contexts that drive several GPUs: the end-of-frame gather of the per-GPU tiles onto the first device.
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load(std::string& err) {
        if (handle) return true;
        handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!handle) handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!handle) {
            err = std::string("cannot load NCCL: ") + dlerror();
            return false;
        }
        auto sym = [&](const char* n) { return dlsym(handle, n); };
        CommInitAll = reinterpret_cast<decltype(CommInitAll)>(sym("ncclCommInitAll"));
        CommDestroy = reinterpret_cast<decltype(CommDestroy)>(sym("ncclCommDestroy"));
        Reduce = reinterpret_cast<decltype(Reduce)>(sym("ncclReduce"));
        GroupStart = reinterpret_cast<decltype(GroupStart)>(sym("ncclGroupStart"));
        GroupEnd = reinterpret_cast<decltype(GroupEnd)>(sym("ncclGroupEnd"));
        GetErrorString = reinterpret_cast<decltype(GetErrorString)>(sym("ncclGetErrorString"));
        if (!CommInitAll || !CommDestroy || !Reduce || !GroupStart || !GroupEnd || !GetErrorString) {
            err = "libnccl.so.2 lacks a required symbol";
            return false;
        }
        return true;
    }
};

// Work counter shared by the processes of one node (one process per GPU): the ranks PULL buckets from it instead of
// being dealt a fixed share, so a rank whose tiles are cheap (sky) takes more of them.  A POSIX shared-memory
// object of 64 counters, one per frame in flight (frame f uses slot f % 64 and clears slot (f + 32) % 64: the
// per-frame collective that gathers the tiles keeps the ranks less than 32 frames apart).
struct ShmQueue {
    struct Layout {
        std::atomic<uint64_t> next[64];
    };
    Layout* mem = nullptr;
    int fd = -1;
    std::string name;
    bool open(uint32_t key, std::string& err) {
        char buf[64];
        std::snprintf(buf, sizeof(buf), "/psy_wq_%08x", key);
        name = buf;
        fd = shm_open(name.c_str(), O_CREAT | O_RDWR, 0600);
        if (fd < 0 || ftruncate(fd, sizeof(Layout)) != 0) {
            err = "shm_open / ftruncate failed for " + name;
            return false;
        }
        void* m = mmap(nullptr, sizeof(Layout), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        if (m == MAP_FAILED) {
            err = "mmap failed for " + name;
            return false;
        }
        mem = static_cast<Layout*>(m);  // (a new object is zero-filled)
        return true;
    }
    void begin_frame(uint64_t frame) { mem->next[(frame + 32) % 64].store(0, std::memory_order_relaxed); }
    uint64_t pull(uint64_t frame) { return mem->next[frame % 64].fetch_add(1, std::memory_order_relaxed); }
    ~ShmQueue() {
        if (mem) munmap(mem, sizeof(Layout));
        if (fd >= 0) close(fd);
        if (!name.empty()) shm_unlink(name.c_str());  // the mapping of the other ranks stays valid
    }
};

struct PsyContext {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    std::mutex cache_mu;
    PoolCache cache;
    // a context over several GPUs: this is the first device's; the others follow (psy_init with n_devices > 1)
    std::vector<PsyContext*> peers;
    std::vector<ncclComm_t> comms;  // comms[i]: device i of {this, peers...}
    NcclApi nccl;
    // process-per-GPU sharding (psy_render_image_shard): frames rendered so far, the shared work counter
    uint64_t shard_frame = 0;
    uint32_t wq_key = 0;
    std::unique_ptr<ShmQueue> wq;
    ~PsyContext() {
        for (ncclComm_t c : comms)
            if (c) nccl.CommDestroy(c);
        for (PsyContext* p : peers) {
            cudaSetDevice(p->device);
            if (p->stream) cudaStreamDestroy(p->stream);
            delete p;
        }
    }
};

struct PsyScene {
    PsyContext* ctx = nullptr;
    DScene d{};
    uint32_t res_x = 0, res_y = 0, spp = 0, seed = 0;
    uint64_t upload_bytes = 0;
    DevBuf meng_baked;
    DevBuf nodes, node_blob, motion, tri_verts, tris, normals, meshes, lights, light_shape, light_colors, objects, shaders,
        assemblies, instances, light_instances, light_inst_local, xforms, light_nodes, lt_bounds, paths, path_xforms,
        path_inv, path_recs, path_static, cam_xforms, cam_tfovs, cam_aperture, cam_focus, distant, world_floats, meng;
    // path pool
    size_t pool_cap = 0;
    DevBuf pool_ray_o, pool_ray_d, pool_flags, pool_sseed, pool_wl, pool_cpdf, pool_nb_o, pool_nb_d, pool_next_atten,
        pool_light_atten, pool_pending, pool_color, pool_hits, pool_light_pick, pool_light_uv, pool_sh_o, pool_sh_d, pool_ic,
        active_a, active_b;
    DevBuf counts, fetch, qcounts, queues, counters, image, bucket_out, samples_out;
    DevBuf pool_cls, pool_alive, part_tmp;  // per-slot byte columns and the block counts of the ordered partition
    DevBuf pool_whist;  // strict film order: the waves each path survived (allocated on first use)
    DevBuf trace_scratch;  // k_trace_pool: child-reference stacks + saved parent spaces, [warp][level][ray]
    uint32_t pool_trace_blocks = 0;
    size_t counts_cap = 0;
    std::vector<cudaEvent_t> events;
    size_t events_used = 0;
    double last_trace_ms = 0.0;
    uint64_t last_trace_launches = 0;
    std::mutex mu;
    // the copies of the scene on the other GPUs of a multi-GPU context (empty otherwise)
    std::vector<PsyScene*> replicas;
    ~PsyScene() {
        if (d.nodes_tex) cudaDestroyTextureObject((cudaTextureObject_t)d.nodes_tex);
        for (auto e : events) cudaEventDestroy(e);
    }
};

namespace {

using HXf = psyh::Xf;

HXf xf_from_abi(const PsyXform& p) {
    HXf m;
    m.X = psyh::v3(p.r[0][0], p.r[1][0], p.r[2][0]);
    m.Y = psyh::v3(p.r[0][1], p.r[1][1], p.r[2][1]);
    m.Z = psyh::v3(p.r[0][2], p.r[1][2], p.r[2][2]);
    m.T = psyh::v3(p.r[0][3], p.r[1][3], p.r[2][3]);
    return m;
}

// spectrum_xyz_to_p_4 (meng.rs:177-328) evaluated AT the 95 table wavelengths for one XYZ colour: what the product
// build interpolates once per colour at upload instead of once per shading call (PSY_COLOR_BAKED, dscene.cuh).
// Returns false when the colour has no spectrum to bake (x + y + z = 0 or not finite: the exact path returns 0).
bool meng_bake(float x, float y, float z, float* out95, float& inv_norm_out) {
    const float inv_norm = x + y + z;
    const float norm = 1.0f / inv_norm;
    if (!(norm < 3.402823466e+38f)) return false;
    inv_norm_out = inv_norm;
    for (int j = 0; j < MENG_NUM_SAMPLES; ++j) out95[j] = 0.0f;
    const float xyy0 = x * norm, xyy1 = y * norm;
    const float uv0 = MENG_MAT_XY_TO_UV[0] * xyy0 + MENG_MAT_XY_TO_UV[1] * xyy1 + MENG_MAT_XY_TO_UV[2];
    const float uv1 = MENG_MAT_XY_TO_UV[3] * xyy0 + MENG_MAT_XY_TO_UV[4] * xyy1 + MENG_MAT_XY_TO_UV[5];
    if (uv0 < 0.0f || uv0 >= (float)MENG_GRID_WIDTH || uv1 < 0.0f || uv1 >= (float)MENG_GRID_HEIGHT) return true;
    const int uvi0 = (int)uv0, uvi1 = (int)uv1;
    const short* cell = MENG_GRID[uvi0 + MENG_GRID_WIDTH * uvi1];
    const bool inside = cell[0] != 0;
    const int num = cell[1];
    if (num == 0) return true;
    const short* idx = cell + 2;
    if (inside) {
        const float u2 = uv0 - (float)uvi0, v2 = uv1 - (float)uvi1;
        for (int j = 0; j < MENG_NUM_SAMPLES; ++j)
            out95[j] = MENG_SPECTRA[idx[0]][j] * ((1.0f - u2) * (1.0f - v2)) + MENG_SPECTRA[idx[2]][j] * ((1.0f - u2) * v2) +
                       MENG_SPECTRA[idx[3]][j] * (u2 * v2) + MENG_SPECTRA[idx[1]][j] * (u2 * (1.0f - v2));
        return true;
    }
    const float ex = uv0 - MENG_UV[idx[0]][0], ey = uv1 - MENG_UV[idx[0]][1];
    float e0x = MENG_UV[idx[1]][0] - MENG_UV[idx[0]][0], e0y = MENG_UV[idx[1]][1] - MENG_UV[idx[0]][1];
    float uu = e0x * ey - ex * e0y;
    for (int i = 0; i < num - 1; ++i) {
        const int last = (i == num - 2) ? 1 : i + 2;
        const float e1x = MENG_UV[idx[last]][0] - MENG_UV[idx[0]][0], e1y = MENG_UV[idx[last]][1] - MENG_UV[idx[0]][1];
        const float vv = ex * e1y - e1x * ey;
        const float area = e0x * e1y - e1x * e0y;
        const float u = uu / area, v = vv / area;
        const float w = 1.0f - u - v;
        if (u < 0.0f || v < 0.0f || w < 0.0f) {
            uu = -vv;
            e0x = e1x;
            e0y = e1y;
            continue;
        }
        for (int j = 0; j < MENG_NUM_SAMPLES; ++j)
            out95[j] = MENG_SPECTRA[idx[0]][j] * w + MENG_SPECTRA[idx[i + 1]][j] * v + MENG_SPECTRA[idx[last]][j] * u;
        break;
    }
    return true;
}

// TransformStack::push (transform_stack.rs:33-62): compose parent keys with child keys
std::vector<HXf> compose(const std::vector<HXf>& parent, const std::vector<HXf>& child) {
    if (parent.empty()) return child;
    size_t n1 = parent.size(), n = child.size();
    std::vector<HXf> out(std::max(n1, n));
    if (n1 == n) {
        for (size_t i = 0; i < n; ++i) out[i] = psyh::xf_mul(parent[i], child[i]);
    } else if (n1 > n) {
        float s = (float)(n1 - 1);
        for (size_t i = 0; i < n1; ++i) out[i] = psyh::xf_mul(parent[i], psyh::lerp_slice(child.data(), n, (float)i / s));
    } else {
        float s = (float)(n - 1);
        for (size_t i = 0; i < n; ++i) out[i] = psyh::xf_mul(psyh::lerp_slice(parent.data(), n1, (float)i / s), child[i]);
    }
    return out;
}

struct HostPath {
    DPath d;
    std::vector<HXf> keys;  // composed keys (empty = none)
};

// [0, n) in contiguous pieces on the host's cores (pieces of at least `grain`); f(begin, end)
template <typename F>
void host_parallel_for(size_t n, size_t grain, F f) {
    const unsigned hw = std::thread::hardware_concurrency();
    const size_t nt = std::min<size_t>(std::min<unsigned>(hw ? hw : 1u, 16u), (n + grain - 1) / std::max<size_t>(grain, 1));
    if (nt <= 1) {
        if (n) f((size_t)0, n);
        return;
    }
    std::vector<std::thread> th;
    const size_t per = (n + nt - 1) / nt;
    for (size_t t = 0; t < nt; ++t) {
        const size_t b = t * per, e = std::min(n, b + per);
        if (b >= e) break;
        th.emplace_back([=]() { f(b, e); });
    }
    for (auto& x : th) x.join();
}

int32_t validate(const PsySceneDesc* s) {
    if (!s) return fail(PSY_ERR_INVALID, "null scene description");
    if (s->abi_version != PSY_ABI_VERSION) return fail(PSY_ERR_INVALID, "PsySceneDesc.abi_version mismatch");
    if (s->res_x == 0 || s->res_y == 0 || s->res_x > 65535 || s->res_y > 65535)
        return fail(PSY_ERR_INVALID, "resolution must be in 1..65535");
    if (s->n_assemblies == 0 || s->root_assembly >= s->n_assemblies) return fail(PSY_ERR_INVALID, "bad root assembly");
    if (s->camera.n_xforms == 0 || s->camera.n_tfovs == 0 || s->camera.n_aperture_radii == 0 ||
        s->camera.n_focus_distances == 0)
        return fail(PSY_ERR_INVALID, "camera needs transforms, fovs, aperture and focus keys");
    for (uint32_t i = 0; i < s->n_instances; ++i) {
        const PsyInstance& in = s->instances[i];
        if (in.kind == PSY_INSTANCE_OBJECT ? in.data_index >= s->n_objects : in.data_index >= s->n_assemblies)
            return fail(PSY_ERR_INVALID, "instance data_index out of range");
        if (in.shader_index != PSY_NONE && in.shader_index >= s->n_shaders)
            return fail(PSY_ERR_INVALID, "instance shader_index out of range");
        if ((uint64_t)in.xform_base + in.xform_count > s->n_xforms) return fail(PSY_ERR_INVALID, "instance xform range");
    }
    for (uint32_t i = 0; i < s->n_objects; ++i) {
        const PsyObject& o = s->objects[i];
        if (o.kind == PSY_OBJECT_MESH ? o.index >= s->n_meshes : o.index >= s->n_lights)
            return fail(PSY_ERR_INVALID, "object index out of range");
    }
    {
        // every node once, on all host cores (the scene is validated at every upload: 57 MB of nodes in the headline scene)
        std::atomic<int> bad{0};
        host_parallel_for(s->n_nodes, 32768, [&](size_t b0, size_t e0) {
            int mine = 0;
            for (size_t i = b0; i < e0 && !mine; ++i) {
                const PsyBvh4Node& n = s->nodes[i];
                if (n.n_children < 2 || n.n_children > 4 || n.traversal_code >= 48 || n.n_keys == 0) {
                    mine = 1;
                    break;
                }
                if (n.n_keys > 1 && (uint64_t)n.key_base + n.n_keys > s->n_motion_bounds) mine = 2;
                for (uint32_t c = 0; c < n.n_children; ++c) {
                    uint32_t r = n.child[c];
                    if (r == PSY_NONE)
                        mine = 3;
                    else if (!PSY_CHILD_IS_LEAF(r) && r >= s->n_nodes)
                        mine = 4;
                }
            }
            if (mine) {
                int expect = 0;
                bad.compare_exchange_strong(expect, mine);
            }
        });
        switch (bad.load()) {
            case 1: return fail(PSY_ERR_INVALID, "malformed BVH4 node");
            case 2: return fail(PSY_ERR_INVALID, "BVH4 node key range");
            case 3: return fail(PSY_ERR_INVALID, "BVH4 node with missing child");
            case 4: return fail(PSY_ERR_INVALID, "BVH4 child out of range");
            default: break;
        }
    }
    for (uint32_t i = 0; i < s->n_assemblies; ++i) {
        const PsyAssembly& a = s->assemblies[i];
        if ((uint64_t)a.instance_base + a.instance_count > s->n_instances ||
            (uint64_t)a.light_instance_base + a.light_instance_count > s->n_light_instances)
            return fail(PSY_ERR_INVALID, "assembly instance range");
    }
    return PSY_OK;
}

// Worst-case traversal stack depth (see dtrace.cuh): 3*depth+2 entries per BVH4 plus one marker.
uint32_t stack_need(const PsySceneDesc* s, uint32_t asm_index, std::vector<int64_t>& memo, int nest, int& max_nest) {
    if (memo[asm_index] >= 0) return (uint32_t)memo[asm_index];
    max_nest = std::max(max_nest, nest);
    const PsyAssembly& a = s->assemblies[asm_index];
    uint32_t own = 3 * a.bvh_depth + 2 + 1;
    uint32_t worst = 0;
    if (nest < 64) {
        for (uint32_t i = 0; i < a.instance_count; ++i) {
            const PsyInstance& in = s->instances[a.instance_base + i];
            uint32_t c = 0;
            if (in.kind == PSY_INSTANCE_ASSEMBLY) {
                c = stack_need(s, in.data_index, memo, nest + 1, max_nest);
            } else if (s->objects[in.data_index].kind == PSY_OBJECT_MESH) {
                c = 3 * s->meshes[s->objects[in.data_index].index].bvh_depth + 2 + 1;
            }
            worst = std::max(worst, c);
        }
    }
    memo[asm_index] = own + worst;
    return own + worst;
}

PathPool make_pool(PsyScene* sc) {
    PathPool P;
    P.ray_o = sc->pool_ray_o.as<float4>();
    P.ray_d = sc->pool_ray_d.as<float4>();
    P.flags = sc->pool_flags.as<uint32_t>();
    P.sseed = sc->pool_sseed.as<uint32_t>();
    P.wavelength = sc->pool_wl.as<float>();
    P.closure_pdf = sc->pool_cpdf.as<float>();
    P.nb_o = sc->pool_nb_o.as<float4>();
    P.nb_d = sc->pool_nb_d.as<float4>();
    P.next_atten = sc->pool_next_atten.as<float4>();
    P.light_atten = sc->pool_light_atten.as<float4>();
    P.pending = sc->pool_pending.as<float4>();
    P.color = sc->pool_color.as<float4>();
    P.hits = sc->pool_hits.as<HitRec>();
    P.light_pick = sc->pool_light_pick.as<uint4>();
    P.light_uv = sc->pool_light_uv.as<float2>();
    P.sh_o = sc->pool_sh_o.as<float4>();
    P.sh_d = sc->pool_sh_d.as<float4>();
    for (int k = 0; k < 6; ++k) P.ic[k] = sc->pool_ic.as<float4>() + (size_t)k * sc->pool_cap;
    P.cls = sc->pool_cls.as<uint8_t>();
    P.alive = sc->pool_alive.as<uint8_t>();
    P.whist = sc->pool_whist.as<uint8_t>();
    return P;
}

// the scene-independent buffers, in a fixed order (PoolCache::bufs uses the same order)
std::vector<DevBuf*> pool_buffers(PsyScene* sc) {
    return {&sc->pool_ray_o, &sc->pool_ray_d, &sc->pool_flags, &sc->pool_sseed, &sc->pool_wl, &sc->pool_cpdf,
            &sc->pool_nb_o, &sc->pool_nb_d, &sc->pool_next_atten, &sc->pool_light_atten, &sc->pool_pending,
            &sc->pool_color, &sc->pool_hits, &sc->pool_light_pick, &sc->pool_light_uv, &sc->pool_sh_o, &sc->pool_sh_d,
            &sc->pool_ic, &sc->pool_cls, &sc->pool_alive, &sc->part_tmp, &sc->active_a, &sc->active_b, &sc->queues,
            &sc->trace_scratch};
}

// psy_scene_destroy: leave the pool with the context if it is the larger one
void donate_pool(PsyScene* sc) {
    if (sc->pool_cap == 0) return;
    PsyContext* ctx = sc->ctx;
    std::lock_guard<std::mutex> g(ctx->cache_mu);
    if (sc->pool_cap <= ctx->cache.cap) return;
    std::vector<DevBuf*> mine = pool_buffers(sc);
    ctx->cache.bufs.clear();  // frees a smaller cached pool
    ctx->cache.bufs.resize(mine.size());
    for (size_t i = 0; i < mine.size(); ++i) ctx->cache.bufs[i] = std::move(*mine[i]);
    ctx->cache.cap = sc->pool_cap;
    ctx->cache.trace_blocks = sc->pool_trace_blocks;
    sc->pool_cap = 0;
}

int32_t ensure_pool(PsyScene* sc, size_t n) {
    if (n <= sc->pool_cap) return PSY_OK;
    if (sc->pool_cap == 0) {
        // a pool left behind by an earlier scene of this context?
        PsyContext* ctx = sc->ctx;
        std::lock_guard<std::mutex> g(ctx->cache_mu);
        if (ctx->cache.cap >= n) {
            std::vector<DevBuf*> mine = pool_buffers(sc);
            for (size_t i = 0; i < mine.size(); ++i) *mine[i] = std::move(ctx->cache.bufs[i]);
            sc->pool_cap = ctx->cache.cap;
            sc->pool_trace_blocks = ctx->cache.trace_blocks;
            ctx->cache.bufs.clear();
            ctx->cache.cap = 0;
            // the byte columns are clean after a completed render; make sure after an aborted one
            CK(cudaMemsetAsync(sc->pool_cls.p, 0xFF, sc->pool_cap + 16, ctx->stream));
            CK(cudaMemsetAsync(sc->pool_alive.p, 0xFF, sc->pool_cap + 16, ctx->stream));
            return PSY_OK;
        }
    }
    size_t cap = n;
    CK(sc->pool_ray_o.alloc(cap * 16));
    CK(sc->pool_ray_d.alloc(cap * 16));
    CK(sc->pool_flags.alloc(cap * 4));
    CK(sc->pool_sseed.alloc(cap * 4));
    CK(sc->pool_wl.alloc(cap * 4));
    CK(sc->pool_cpdf.alloc(cap * 4));
    CK(sc->pool_nb_o.alloc(cap * 16));
    CK(sc->pool_nb_d.alloc(cap * 16));
    CK(sc->pool_next_atten.alloc(cap * 16));
    CK(sc->pool_light_atten.alloc(cap * 16));
    CK(sc->pool_pending.alloc(cap * 16));
    CK(sc->pool_color.alloc(cap * 16));
    CK(sc->pool_hits.alloc(cap * sizeof(HitRec)));
    CK(sc->pool_light_pick.alloc(cap * 16));
    CK(sc->pool_light_uv.alloc(cap * 8));
    CK(sc->pool_sh_o.alloc(cap * 16));
    CK(sc->pool_sh_d.alloc(cap * 16));
    CK(sc->pool_ic.alloc(cap * 16 * 6));
    CK(sc->pool_cls.alloc(cap + 16));
    CK(sc->pool_alive.alloc(cap + 16));
    CK(cudaMemsetAsync(sc->pool_cls.p, 0xFF, cap + 16, sc->ctx->stream));
    CK(cudaMemsetAsync(sc->pool_alive.p, 0xFF, cap + 16, sc->ctx->stream));
    CK(sc->part_tmp.alloc(((cap + PART_PER_BLOCK - 1) / PART_PER_BLOCK) * SHADE_CLASSES * 4));
    CK(sc->active_a.alloc(cap * 4));
    CK(sc->active_b.alloc(cap * 4));
    CK(sc->queues.alloc(cap * 4 * SHADE_CLASSES));
    if (!sc->trace_scratch.p) {
        sc->pool_trace_blocks = (uint32_t)sc->ctx->sm_count * PSY_POOL_MIN_BLOCKS;
        CK(sc->trace_scratch.alloc((size_t)sc->pool_trace_blocks * POOL_WARPS * POOL_SCRATCH_WORDS_PER_WARP * 4));
        CK(cudaFuncSetAttribute(k_trace_pool<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POOL_SMEM_BYTES));
        CK(cudaFuncSetAttribute(k_trace_pool<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POOL_SMEM_BYTES));
    }
    sc->pool_cap = cap;
    return PSY_OK;
}

constexpr int WAVES = 6;           // camera, shadow, bounce1, shadow, bounce2, shadow (renderer.rs:599)
constexpr int COUNT_SLOTS = 8;     // per bucket: counts[0..6]

int32_t ensure_counts(PsyScene* sc, size_t n_buckets) {
    if (n_buckets <= sc->counts_cap) return PSY_OK;
    CK(sc->counts.alloc(n_buckets * COUNT_SLOTS * 4));
    CK(sc->fetch.alloc(n_buckets * COUNT_SLOTS * 4));
    CK(sc->qcounts.alloc(n_buckets * COUNT_SLOTS * SHADE_CLASSES * 4));
    sc->counts_cap = n_buckets;
    return PSY_OK;
}

// (nullptr when the driver refuses another event: the callers then skip the time stamp, cudaEventRecord(nullptr)
// fails without side effects and sum_ms ignores pairs it cannot measure)
cudaEvent_t next_event(PsyScene* sc) {
    if (sc->events_used == sc->events.size()) {
        cudaEvent_t e = nullptr;
        if (cudaEventCreate(&e) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        sc->events.push_back(e);
    }
    return sc->events[sc->events_used++];
}

struct StageEvents {
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> gen, trace, shade, film;
};

// Enqueues the whole wavefront for one bucket on the scene's stream (no host synchronisation).
int32_t enqueue_bucket(PsyScene* sc, const Bucket& b, size_t bucket_index, float* out, uint32_t stride, uint32_t ox,
                       uint32_t oy, float* samples_out, bool counters, StageEvents* ev, uint64_t* launches,
                       bool strict_film = false) {
    cudaStream_t st = sc->ctx->stream;
    uint32_t n = b.w * b.h * b.spp;
    if (strict_film) {
        // the reference's own limit: RayBatch indices are u16 (ray.rs:7)
        if (n > STRICT_MAX_PATHS)
            return fail(PSY_ERR_LIMIT, "strict_film_order needs buckets of at most 65536 paths (pass the reference's --spb)");
        if (sc->pool_whist.bytes < sc->pool_cap) {
            CK(sc->pool_whist.alloc(sc->pool_cap));
            CK(cudaFuncSetAttribute(k_film_strict, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(STRICT_MAX_PATHS * 3)));
        }
        CK(cudaMemsetAsync(sc->pool_whist.p, 0, n, st));
    }
    PathPool P = make_pool(sc);
    uint32_t* counts = sc->counts.as<uint32_t>() + bucket_index * COUNT_SLOTS;
    uint32_t* fetch = sc->fetch.as<uint32_t>() + bucket_index * COUNT_SLOTS;
    uint32_t* act[2] = {sc->active_a.as<uint32_t>(), sc->active_b.as<uint32_t>()};
    auto stamp = [&](std::vector<std::pair<cudaEvent_t, cudaEvent_t>>* v, bool begin) {
        if (!ev) return;
        cudaEvent_t e = next_event(sc);
        if (e) cudaEventRecord(e, st);
        if (begin)
            v->push_back({e, nullptr});
        else
            v->back().second = e;
    };
    int sms = sc->ctx->sm_count;
    stamp(ev ? &ev->gen : nullptr, true);
    k_generate<<<(n + 255) / 256, 256, 0, st>>>(sc->d, P, b, n, act[0], counts);
    stamp(ev ? &ev->gen : nullptr, false);
    *launches += 1;
    uint32_t trace_blocks = (uint32_t)std::min<uint64_t>((n + 127) / 128, (uint64_t)sms * 12);
    uint32_t shade_blocks = (uint32_t)std::min<uint64_t>((n + 127) / 128, (uint64_t)sms * 16);
    // k_shade_bounce is small since it reads the hit / closure k_shade_light prepared: 3 x 256 threads per SM
    uint32_t bounce_blocks = (uint32_t)std::min<uint64_t>((n + 255) / 256, (uint64_t)sms * 3);
    // k_shade_light: one register-limited block per SM (see kernels.cuh)
    uint32_t light_blocks = (uint32_t)std::min<uint64_t>((n + PSY_LIGHT_THREADS - 1) / PSY_LIGHT_THREADS,
                                                         (uint64_t)sms * std::max(1, 512 / PSY_LIGHT_THREADS));
    ShadeQueues Q;
    static const TraceTuning tune = [] {
        TraceTuning t{64, 12, 8, 8};  // measured best on the C4 scene (profiles/README.md)
        if (const char* e = getenv("PSY_TRACE_LEAFMIN")) t.leaf_min = atoi(e);
        if (const char* e = getenv("PSY_TRACE_INSTMIN")) t.inst_min = atoi(e);
        if (const char* e = getenv("PSY_TRACE_BUDGET")) t.step_budget = atoi(e);
        if (const char* e = getenv("PSY_TRACE_REFILL")) t.refill_min = atoi(e);
        return t;
    }();
    // PSY_TRACE_KERNEL=warp selects the one-ray-per-lane kernel (k_trace) for A/B measurements
    static const bool pool_kernel = [] {
        const char* e = getenv("PSY_TRACE_KERNEL");
        return !(e && std::string(e) == "warp");
    }();
    // The visit counters come from the counting instantiation of the production kernel (k_trace_pool<true>: same
    // traversal, four more per-lane adds); PSY_POOL_COUNT=0 takes them from k_trace<true> instead (A/B: both kernels
    // visit exactly the same nodes / primitives per ray, tests/test_gpu_parity.py compares them).
    static const bool pool_counting = [] {
        const char* e = getenv("PSY_POOL_COUNT");
        return pool_kernel && !(e && e[0] == '0');
    }();
    for (int c = 0; c < SHADE_CLASSES; ++c) Q.q[c] = sc->queues.as<uint32_t>() + (size_t)c * sc->pool_cap;
    uint32_t* qcounts = sc->qcounts.as<uint32_t>() + bucket_index * COUNT_SLOTS * SHADE_CLASSES;
    // slot-ordered lists from the per-slot byte columns (dpartition.cuh)
    const uint32_t part_nb = (n + PART_PER_BLOCK - 1) / PART_PER_BLOCK;
    uint32_t* part_tmp = sc->part_tmp.as<uint32_t>();
    for (int w = 0; w < WAVES; ++w) {
        Q.count = qcounts + w * SHADE_CLASSES;
        stamp(ev ? &ev->trace : nullptr, true);
        if (pool_kernel && !counters)
            k_trace_pool<false><<<sc->pool_trace_blocks, 32 * POOL_WARPS, POOL_SMEM_BYTES, st>>>(
                sc->d, P, act[w & 1], counts + w, fetch + w, nullptr, Q, sc->trace_scratch.as<uint32_t>());
        else if (counters && pool_counting)
            k_trace_pool<true><<<sc->pool_trace_blocks, 32 * POOL_WARPS, POOL_SMEM_BYTES, st>>>(
                sc->d, P, act[w & 1], counts + w, fetch + w, sc->counters.as<TraceCounters>(), Q,
                sc->trace_scratch.as<uint32_t>());
        else if (counters)
            k_trace<true><<<trace_blocks, 128, 0, st>>>(sc->d, P, act[w & 1], counts + w, fetch + w,
                                                        sc->counters.as<TraceCounters>(), Q, tune);
        else
            k_trace<false><<<trace_blocks, 128, 0, st>>>(sc->d, P, act[w & 1], counts + w, fetch + w, nullptr, Q, tune);
        stamp(ev ? &ev->trace : nullptr, false);
        stamp(ev ? &ev->shade : nullptr, true);
        uint32_t* next_active = act[(w + 1) & 1];
        uint32_t* next_count = counts + w + 1;
        {
            // class queues of this wave, in slot order
            PartLists L{};
            for (int c = 0; c < SHADE_CLASSES; ++c) L.list[c] = Q.q[c], L.total[c] = Q.count + c;
            k_part_count<SHADE_CLASSES><<<part_nb, PART_THREADS, 0, st>>>(P.cls, n, part_tmp, part_nb);
            k_part_scan<SHADE_CLASSES><<<1, 1024, 0, st>>>(part_tmp, part_nb, L);
            k_part_scatter<SHADE_CLASSES><<<part_nb, PART_THREADS, 0, st>>>(P.cls, n, part_tmp, part_nb, L);
        }
        k_shade<SHADE_SIMPLE><<<shade_blocks, 128, 0, st>>>(sc->d, P, b, Q, next_active, next_count);
        k_shade<SHADE_EMIT><<<shade_blocks, 128, 0, st>>>(sc->d, P, b, Q, next_active, next_count);
        // surface hits: pick a light (tree walk), sample it + evaluate the closure, then bounce + path state
        k_shade_select<SHADE_LAMBERT><<<shade_blocks, 128, 0, st>>>(sc->d, P, b, Q);
        k_shade_light<SHADE_LAMBERT><<<light_blocks, PSY_LIGHT_THREADS, 0, st>>>(sc->d, P, b, Q);
        k_shade_bounce<SHADE_LAMBERT><<<bounce_blocks, 256, 0, st>>>(sc->d, P, b, Q, next_active, next_count);
        k_shade_select<SHADE_GGX><<<shade_blocks, 128, 0, st>>>(sc->d, P, b, Q);
        k_shade_light<SHADE_GGX><<<light_blocks, PSY_LIGHT_THREADS, 0, st>>>(sc->d, P, b, Q);
        k_shade_bounce<SHADE_GGX><<<bounce_blocks, 256, 0, st>>>(sc->d, P, b, Q, next_active, next_count);
        if (strict_film) k_record_alive<<<(n + 255) / 256, 256, 0, st>>>(P.alive, n, (uint32_t)w, P.whist);
        {
            // the next wave's live list, in slot order
            PartLists L{};
            L.list[0] = next_active, L.total[0] = next_count;
            k_part_count<1><<<part_nb, PART_THREADS, 0, st>>>(P.alive, n, part_tmp, part_nb);
            k_part_scan<1><<<1, 1024, 0, st>>>(part_tmp, part_nb, L);
            k_part_scatter<1><<<part_nb, PART_THREADS, 0, st>>>(P.alive, n, part_tmp, part_nb, L);
        }
        stamp(ev ? &ev->shade : nullptr, false);
        *launches += 15;
    }
    stamp(ev ? &ev->film : nullptr, true);
    if (strict_film) {
        // (act[0] is free by now: the rank of every path in the reference's final path order)
        k_film_strict<<<1, STRICT_THREADS, (size_t)n * 3, st>>>(P, b, out, stride, ox, oy, samples_out, act[0], (uint32_t)WAVES);
        *launches += WAVES;
    } else {
        // pixels per warp: as many as still give every SM a few warps
        const uint32_t n_pix = b.w * b.h;
        uint32_t ppw = 32;
        while (b.spp >= 16u && ppw > 2u && n_pix / ppw < (uint32_t)sms * 16u) ppw >>= 1;
        const uint32_t per_block = ppw * FILM_WARPS;
        k_film<<<(n_pix + per_block - 1) / per_block, 32 * FILM_WARPS, 0, st>>>(P, b, out, stride, ox, oy, samples_out, ppw);
    }
    stamp(ev ? &ev->film : nullptr, false);
    *launches += 1;
    CK(cudaGetLastError());
    return PSY_OK;
}

double sum_ms(const std::vector<std::pair<cudaEvent_t, cudaEvent_t>>& v) {
    double t = 0.0;
    for (auto& p : v) {
        float ms = 0.0f;
        if (p.first && p.second && cudaEventElapsedTime(&ms, p.first, p.second) == cudaSuccess) t += ms;
    }
    return t;
}

// Renderer::render's bucket list (renderer.rs:138-180) for a pool of `spb` paths
std::vector<Bucket> make_buckets(uint32_t width, uint32_t height, uint32_t start_x, uint32_t start_y, uint32_t spp,
                                 uint32_t seed, size_t spb) {
    double target = (double)spb / (double)spp;
    uint32_t bdim = std::sqrt(target) < 1.0 ? 1u : (uint32_t)std::sqrt(target);
    std::vector<Bucket> jobs;
    uint32_t cx = width / bdim + 1, cy = height / bdim + 1;
    uint32_t larger = std::max(cx, cy), pow2 = 1;
    while (pow2 < larger) pow2 <<= 1;
    uint64_t bucket_n = (uint64_t)pow2 * pow2;
    for (uint64_t hd = 0; hd < bucket_n; ++hd) {
        // hilbert::d2xy (hilbert.rs:47-63)
        uint32_t x = 0, y = 0, t = (uint32_t)hd;
        for (uint32_t s = 1; s < (1u << 16); s <<= 1) {
            uint32_t rx = 1 & (t >> 1);
            uint32_t ry = 1 & (t ^ rx);
            if (ry == 0) {
                if (rx == 1) {
                    x = (s - 1) - x;
                    y = (s - 1) - y;
                }
                std::swap(x, y);
            }
            x += s * rx;
            y += s * ry;
            t >>= 2;
        }
        uint64_t bx = (uint64_t)x * bdim, by = (uint64_t)y * bdim;
        if (bx >= width || by >= height) continue;
        uint32_t w = (uint32_t)std::min<uint64_t>(bdim, width - bx), h = (uint32_t)std::min<uint64_t>(bdim, height - by);
        if (w > 0 && h > 0) jobs.push_back(Bucket{start_x + (uint32_t)bx, start_y + (uint32_t)by, w, h, spp, seed});
    }
    return jobs;
}

// Where a render takes its buckets from: the whole list in order (one GPU), or a counter shared with the other GPUs
// of the job — threads of this process (std::atomic) or processes of this node (ShmQueue) — from which every GPU
// pulls the next bucket when it is about to run dry.
struct JobFeed {
    std::function<int64_t()> next;  // index into the job list, or -1 when the list is exhausted
    int max_in_flight = 0;          // buckets enqueued ahead of the GPU; 0 = no limit (static list)
};

JobFeed feed_all(size_t n_jobs, std::shared_ptr<size_t> pos = std::make_shared<size_t>(0)) {
    JobFeed f;
    f.next = [n_jobs, pos]() -> int64_t { return *pos < n_jobs ? (int64_t)(*pos)++ : -1; };
    return f;
}

// after an aborted render the byte columns may hold stale classes / survivor marks: clear them before the next one
void reset_after_error(PsyScene* sc) {
    cudaStreamSynchronize(sc->ctx->stream);
    cudaGetLastError();
    if (sc->pool_cap) {
        cudaMemsetAsync(sc->pool_cls.p, 0xFF, sc->pool_cap + 16, sc->ctx->stream);
        cudaMemsetAsync(sc->pool_alive.p, 0xFF, sc->pool_cap + 16, sc->ctx->stream);
        cudaStreamSynchronize(sc->ctx->stream);
    }
}

int32_t render_buckets_impl(PsyScene* sc, const std::vector<Bucket>& jobs, JobFeed& feed, bool counters, PsyStats* stats,
                            float* host_out, psy_bucket_cb cb, void* user, bool strict_film, std::mutex* cb_mu) {
    cudaStream_t st = sc->ctx->stream;
    auto t0 = std::chrono::steady_clock::now();
    size_t max_n = 0;
    for (auto& b : jobs) max_n = std::max(max_n, (size_t)b.w * b.h * b.spp);
    int32_t rc = ensure_pool(sc, max_n);
    if (rc) return rc;
    rc = ensure_counts(sc, jobs.size());
    if (rc) return rc;
    if (!sc->image.p) {
        CK(sc->image.alloc((size_t)sc->res_x * sc->res_y * 3 * sizeof(float)));
        CK(cudaMemsetAsync(sc->image.p, 0, sc->image.bytes, st));
    }
    if (!sc->counters.p) CK(sc->counters.alloc(sizeof(TraceCounters)));
    CK(cudaMemsetAsync(sc->counters.p, 0, sizeof(TraceCounters), st));
    if (!jobs.empty()) {
        CK(cudaMemsetAsync(sc->counts.p, 0, jobs.size() * COUNT_SLOTS * 4, st));
        CK(cudaMemsetAsync(sc->fetch.p, 0, jobs.size() * COUNT_SLOTS * 4, st));
        CK(cudaMemsetAsync(sc->qcounts.p, 0, jobs.size() * COUNT_SLOTS * SHADE_CLASSES * 4, st));
    }
    sc->events_used = 0;
    StageEvents ev;
    uint64_t launches = 0;
    std::vector<float> cb_buf;
    std::vector<size_t> mine;          // the buckets this GPU rendered, in the order it took them
    std::vector<cudaEvent_t> done_ev;  // (pulling feeds only) bucket k has left the GPU
    for (;;) {
        if (feed.max_in_flight > 0 && (int)mine.size() >= feed.max_in_flight) {
            // do not hoard: the next bucket is taken only when the GPU has at most max_in_flight - 1 left to do
            cudaEvent_t e = done_ev[mine.size() - (size_t)feed.max_in_flight];
            if (e) CK(cudaEventSynchronize(e));
        }
        const int64_t j = feed.next();
        if (j < 0) break;
        const Bucket& b = jobs[(size_t)j];
        rc = enqueue_bucket(sc, b, mine.size(), sc->image.as<float>(), sc->res_x, 0, 0, nullptr, counters, &ev, &launches,
                            strict_film);
        if (rc) return rc;
        mine.push_back((size_t)j);
        if (feed.max_in_flight > 0) {
            cudaEvent_t e = next_event(sc);
            if (e) CK(cudaEventRecord(e, st));
            done_ev.push_back(e);
        }
        if (cb) {
            // Blender-style streaming: hand the finished bucket to the caller (renderer.rs:312-346)
            cb_buf.resize((size_t)b.w * b.h * 3);
            CK(cudaMemcpy2DAsync(cb_buf.data(), (size_t)b.w * 12, sc->image.as<float>() + ((size_t)b.y * sc->res_x + b.x) * 3,
                                 (size_t)sc->res_x * 12, (size_t)b.w * 12, b.h, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            if (cb_mu) {
                std::lock_guard<std::mutex> g(*cb_mu);
                cb(user, b.x, b.y, b.x + b.w, b.y + b.h, cb_buf.data());
            } else {
                cb(user, b.x, b.y, b.x + b.w, b.y + b.h, cb_buf.data());
            }
        }
    }
    if (host_out)
        CK(cudaMemcpyAsync(host_out, sc->image.p, sc->image.bytes, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());

    std::vector<uint32_t> counts(mine.size() * COUNT_SLOTS);
    if (!mine.empty()) CK(cudaMemcpy(counts.data(), sc->counts.p, counts.size() * 4, cudaMemcpyDeviceToHost));
    uint64_t rays = 0, samples = 0;
    for (size_t k = 0; k < mine.size(); ++k) {
        for (int w = 0; w < WAVES; ++w) rays += counts[k * COUNT_SLOTS + w];
        samples += (uint64_t)jobs[mine[k]].w * jobs[mine[k]].h * jobs[mine[k]].spp;
    }
    TraceCounters tc{};
    if (counters) CK(cudaMemcpy(&tc, sc->counters.p, sizeof(tc), cudaMemcpyDeviceToHost));
    sc->last_trace_ms = sum_ms(ev.trace);
    sc->last_trace_launches = ev.trace.size();
    if (getenv("PSY_WAVE_STATS") && ev.trace.size() == mine.size() * WAVES && ev.shade.size() == ev.trace.size()) {
        // per wave over all buckets: rays, trace / shade milliseconds, rays per microsecond of the trace kernel —
        // the evidence for or against refilling dead path slots (later waves are thinner: are they slower per ray?)
        for (int w = 0; w < WAVES; ++w) {
            uint64_t r = 0;
            double tms = 0.0, sms = 0.0;
            for (size_t k = 0; k < mine.size(); ++k) {
                r += counts[k * COUNT_SLOTS + w];
                float ms = 0.0f;
                auto& pt = ev.trace[k * WAVES + w];
                if (pt.first && pt.second && cudaEventElapsedTime(&ms, pt.first, pt.second) == cudaSuccess) tms += ms;
                auto& ps = ev.shade[k * WAVES + w];
                if (ps.first && ps.second && cudaEventElapsedTime(&ms, ps.first, ps.second) == cudaSuccess) sms += ms;
            }
            std::fprintf(stderr, "[waves] wave %d: %12llu rays  trace %9.2f ms (%7.1f Mrays/s)  shade %9.2f ms\n", w + 1,
                         (unsigned long long)r, tms, tms > 0 ? (double)r / tms * 1e-3 : 0.0, sms);
        }
    }
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        stats->trace_time = sc->last_trace_ms * 1e-3;
        stats->initial_ray_generation_time = sum_ms(ev.gen) * 1e-3;
        stats->ray_generation_time = sum_ms(ev.shade) * 1e-3;
        stats->sample_writing_time = sum_ms(ev.film) * 1e-3;
        stats->total_time = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        stats->ray_count = rays;
        stats->accel_node_visits = tc.node_visits;
        stats->tri_tests = tc.tri_tests;
        stats->instance_visits = tc.instance_visits;
        stats->motion_node_visits = tc.motion_node_visits;
        stats->samples = samples;
        stats->kernel_launches = launches;
        stats->buckets = mine.size();
    }
    return PSY_OK;
}

int32_t render_buckets(PsyScene* sc, const std::vector<Bucket>& jobs, JobFeed& feed, bool counters, PsyStats* stats,
                       float* host_out, psy_bucket_cb cb, void* user, bool strict_film = false, std::mutex* cb_mu = nullptr) {
    int32_t rc = render_buckets_impl(sc, jobs, feed, counters, stats, host_out, cb, user, strict_film, cb_mu);
    if (rc) {
        const std::string msg = g_err;
        reset_after_error(sc);
        g_err = msg;
    }
    return rc;
}

}  // namespace

extern "C" {

const char* psy_last_error(void) { return g_err.c_str(); }
int32_t psy_abi_version(void) { return PSY_ABI_VERSION; }
uint32_t psy_build_flags(void) { return (PSY_STRICT ? PSY_BUILD_STRICT_MATH : 0u); }

// One device of a context: stream + the constant-memory tables (per device)
static int32_t init_device(int dev, std::unique_ptr<PsyContext>& out) {
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    std::unique_ptr<PsyContext> c(new PsyContext());
    c->device = dev;
    c->sm_count = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    // static tables
    uint8_t table[8][48];
    for (int raydir = 0; raydir < 8; ++raydir) {
        // TRAVERSAL_TABLE (sub_crates/bvh_order/build.rs:17-75)
        int ray[3] = {raydir & 1, (raydir >> 1) & 1, (raydir >> 2) & 1};
        int n = 0;
        for (int s2 = 0; s2 < 3; ++s2)
            for (int s1 = 0; s1 < 3; ++s1)
                for (int s0 = 0; s0 < 3; ++s0) {
                    int perm[4] = {0, 1, 2, 3};
                    if (ray[s1]) std::swap(perm[0], perm[1]);
                    if (ray[s2]) std::swap(perm[2], perm[3]);
                    if (ray[s0]) {
                        std::swap(perm[0], perm[2]);
                        std::swap(perm[1], perm[3]);
                    }
                    table[raydir][n++] = (uint8_t)(perm[0] + (perm[1] << 2) + (perm[2] << 4) + (perm[3] << 6));
                }
        for (int s1 = 0; s1 < 3; ++s1)
            for (int s0 = 0; s0 < 3; ++s0) {
                int perm[3] = {0, 1, 2};
                if (ray[s1]) std::swap(perm[0], perm[1]);
                if (ray[s0]) {
                    std::swap(perm[0], perm[1]);
                    std::swap(perm[0], perm[2]);
                }
                table[raydir][n++] = (uint8_t)(perm[0] + (perm[1] << 2) + (perm[2] << 4));
            }
        for (int s1 = 0; s1 < 3; ++s1)
            for (int s0 = 0; s0 < 3; ++s0) {
                int perm[3] = {0, 1, 2};
                if (ray[s1]) std::swap(perm[1], perm[2]);
                if (ray[s0]) {
                    std::swap(perm[0], perm[2]);
                    std::swap(perm[0], perm[1]);
                }
                table[raydir][n++] = (uint8_t)(perm[0] + (perm[1] << 2) + (perm[2] << 4));
            }
        for (int s0 = 0; s0 < 3; ++s0) {
            int perm[2] = {0, 1};
            if (ray[s0]) std::swap(perm[0], perm[1]);
            table[raydir][n++] = (uint8_t)(perm[0] + (perm[1] << 2));
        }
    }
    CK(cudaMemcpyToSymbol(c_traversal_table, table, sizeof(table)));
    CK(cudaMemcpyToSymbol(c_sobol_rev, SOBOL_REV_VECTORS, sizeof(SOBOL_REV_VECTORS)));
    CK(cudaMemcpyToSymbol(c_meng_grid, MENG_GRID, sizeof(MENG_GRID)));
    CK(cudaMemcpyToSymbol(c_meng_uv, MENG_UV, sizeof(MENG_UV)));
    CK(cudaMemcpyToSymbol(c_meng_mat, MENG_MAT_XY_TO_UV, sizeof(MENG_MAT_XY_TO_UV)));
    out = std::move(c);
    return PSY_OK;
}

int32_t psy_init(const int* device_ids, int n_devices, PsyContext** out) {
    if (!out) return fail(PSY_ERR_INVALID, "psy_init: out is null");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(PSY_ERR_NO_DEVICE, std::string("no CUDA device: ") + cudaGetErrorString(e) +
                                           " (this library has no CPU fallback)");
    std::vector<int> ids;
    if (device_ids && n_devices >= 1)
        ids.assign(device_ids, device_ids + n_devices);
    else
        ids.push_back(0);
    for (size_t i = 0; i < ids.size(); ++i) {
        if (ids[i] < 0 || ids[i] >= count) return fail(PSY_ERR_INVALID, "device id out of range");
        for (size_t j = 0; j < i; ++j)
            if (ids[j] == ids[i]) return fail(PSY_ERR_INVALID, "device listed twice");
    }
    std::unique_ptr<PsyContext> c;
    int32_t rc = init_device(ids[0], c);
    if (rc) return rc;
    for (size_t i = 1; i < ids.size(); ++i) {
        std::unique_ptr<PsyContext> p;
        rc = init_device(ids[i], p);
        if (rc) return rc;  // (c's destructor releases the peers made so far)
        c->peers.push_back(p.release());
    }
    if (ids.size() > 1) {
        // several GPUs in one process: their tiles are gathered with NCCL at the end of a frame (renderer.rs:76-191 is
        // the seam: one worker per GPU instead of one per CPU thread)
        std::string err;
        if (!c->nccl.load(err)) return fail(PSY_ERR_CUDA, err);
        c->comms.assign(ids.size(), nullptr);
        ncclResult_t nr = c->nccl.CommInitAll(c->comms.data(), (int)ids.size(), ids.data());
        if (nr != ncclSuccess) {
            c->comms.clear();
            return fail(PSY_ERR_CUDA, std::string("ncclCommInitAll: ") + c->nccl.GetErrorString(nr));
        }
    }
    CK(cudaSetDevice(ids[0]));
    *out = c.release();
    return PSY_OK;
}

int32_t psy_shutdown(PsyContext* ctx) {
    if (!ctx) return PSY_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    buf_cache().drop(ctx->device);
    for (PsyContext* p : ctx->peers) {
        cudaSetDevice(p->device);
        buf_cache().drop(p->device);
    }
    delete ctx;  // peers, communicators, work queue
    return PSY_OK;
}

static int32_t scene_create_one(PsyContext* ctx, const PsySceneDesc* s, PsyScene** out) {
    int32_t rc = PSY_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    // PSY_UPLOAD_TIMING=1: where the time of an upload goes (stderr)
    static const bool timing = getenv("PSY_UPLOAD_TIMING") != nullptr;
    auto tp0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[upload] %-28s %.2f ms\n", what, std::chrono::duration<double>(now - tp0).count() * 1e3);
        tp0 = now;
    };

    // stack requirement
    {
        std::vector<int64_t> memo(s->n_assemblies, -1);
        int max_nest = 0;
        uint32_t need = stack_need(s, s->root_assembly, memo, 0, max_nest);
        // (+1: k_trace_pool writes the nearest hit child to the stack before taking it back as the new top)
        if (need + 1 > (uint32_t)TRACE_STACK_SIZE)
            return fail(PSY_ERR_LIMIT, "scene needs a traversal stack of " + std::to_string(need + 1) +
                                           " entries, limit is " + std::to_string(TRACE_STACK_SIZE));
        if (max_nest >= TRACE_MAX_NEST) return fail(PSY_ERR_LIMIT, "assembly nesting deeper than TRACE_MAX_NEST");
    }

    // expand instance paths breadth-first and compose their transforms
    std::vector<HostPath> paths;
    {
        struct Item {
            uint32_t path, assembly;
        };
        std::vector<Item> queue;
        auto add_children = [&](uint32_t asm_index, const std::vector<HXf>& parent_keys) -> int32_t {
            const PsyAssembly& a = s->assemblies[asm_index];
            if (paths.size() + a.instance_count > (1u << 26)) return fail(PSY_ERR_LIMIT, "more than 2^26 instance paths");
            uint32_t base = (uint32_t)paths.size();
            for (uint32_t i = 0; i < a.instance_count; ++i) {
                const PsyInstance& in = s->instances[a.instance_base + i];
                HostPath hp;
                hp.d.inst = a.instance_base + i;
                hp.d.child_base = PSY_NONE;
                if (in.xform_count > 0) {
                    std::vector<HXf> own(in.xform_count);
                    for (uint32_t k = 0; k < in.xform_count; ++k) own[k] = xf_from_abi(s->xforms[in.xform_base + k]);
                    hp.keys = compose(parent_keys, own);
                } else {
                    hp.keys = parent_keys;
                }
                paths.push_back(std::move(hp));
                if (in.kind == PSY_INSTANCE_ASSEMBLY) queue.push_back({base + i, in.data_index});
            }
            return PSY_OK;
        };
        rc = add_children(s->root_assembly, {});
        if (rc) return rc;
        for (size_t q = 0; q < queue.size(); ++q) {
            Item it = queue[q];
            paths[it.path].d.child_base = (uint32_t)paths.size();
            std::vector<HXf> keys = paths[it.path].keys;
            rc = add_children(it.assembly, keys);
            if (rc) return rc;
        }
    }
    std::vector<DPath> dpaths(paths.size());
    std::vector<PsyXform> path_xf, path_inv(paths.size());
    for (size_t i = 0; i < paths.size(); ++i) {
        paths[i].d.xf_base = (uint32_t)path_xf.size();
        paths[i].d.xf_count = (uint32_t)paths[i].keys.size();
        for (auto& k : paths[i].keys) path_xf.push_back(psyh::to_abi(k));
        path_inv[i] = psyh::to_abi(paths[i].keys.size() == 1 ? psyh::xf_inverse(paths[i].keys[0]) : HXf());
        dpaths[i] = paths[i].d;
    }

    lap("instance paths");
    // gather triangle vertices per leaf-ordered triangle: [v0 keys][v1 keys][v2 keys]
    std::vector<DMesh> dmeshes(s->n_meshes);
    size_t tv_total = 0;
    for (uint32_t m = 0; m < s->n_meshes; ++m) tv_total += (size_t)s->meshes[m].tri_count * 3 * s->meshes[m].n_keys;
    if (tv_total > 0xFFFFFFFFull) return fail(PSY_ERR_LIMIT, "gathered triangle vertex array exceeds 2^32 entries");
    // (the gather itself runs on the GPU after the upload, k_gather_tri_verts; the host only checks the indices)
    {
        size_t o = 0;
        for (uint32_t m = 0; m < s->n_meshes; ++m) {
            const PsyMesh& pm = s->meshes[m];
            DMesh& dm = dmeshes[m];
            dm.root = pm.root, dm.tri_base = pm.tri_base, dm.tri_count = pm.tri_count, dm.n_keys = pm.n_keys;
            dm.tv_base = (uint32_t)o;
            dm.vert_base = pm.vert_base;
            dm.has_normals = pm.has_normals;
            dm.pad = 0;
            if ((uint64_t)pm.tri_base + pm.tri_count > s->n_tris) return fail(PSY_ERR_INVALID, "mesh triangle range");
            if (((uint64_t)pm.vert_base + (uint64_t)pm.vert_count * pm.n_keys) > s->n_vert_keys)
                return fail(PSY_ERR_INVALID, "mesh vertex range");
            if (pm.tri_count && pm.vert_count == 0) return fail(PSY_ERR_INVALID, "mesh with triangles but no vertices");
            // (the triangles' vertex indices are checked by the gather kernel on the device)
            o += (size_t)pm.tri_count * 3 * pm.n_keys;
        }
    }
    // k_trace_pool's node blob (dscene.cuh): where every node starts, in 16-byte units, and the re-encoded references
    std::vector<uint32_t> unit_off(s->n_nodes);
    uint64_t blob_units = 0;
    {
        // prefix sum of the node sizes in chunks: sizes per chunk, chunk starts, then the offsets (parallel passes)
        const size_t chunk = 32768, n_chunks = ((size_t)s->n_nodes + chunk - 1) / chunk;
        std::vector<uint64_t> chunk_units(n_chunks + 1, 0);
        host_parallel_for(n_chunks, 1, [&](size_t c0, size_t c1) {
            for (size_t c = c0; c < c1; ++c) {
                uint64_t u = 0;
                const size_t e = std::min<size_t>(s->n_nodes, (c + 1) * chunk);
                for (size_t i = c * chunk; i < e; ++i) u += 2u + 6u * (uint64_t)s->nodes[i].n_keys;
                chunk_units[c + 1] = u;
            }
        });
        for (size_t c = 0; c < n_chunks; ++c) chunk_units[c + 1] += chunk_units[c];
        blob_units = chunk_units[n_chunks];
        if (blob_units > (uint64_t)NODE2_OFF_MASK) return fail(PSY_ERR_LIMIT, "BVH4 node store exceeds 4 GB");
        host_parallel_for(n_chunks, 1, [&](size_t c0, size_t c1) {
            for (size_t c = c0; c < c1; ++c) {
                uint64_t u = chunk_units[c];
                const size_t e = std::min<size_t>(s->n_nodes, (c + 1) * chunk);
                for (size_t i = c * chunk; i < e; ++i) {
                    unit_off[i] = (uint32_t)u;
                    u += 2u + 6u * (uint64_t)s->nodes[i].n_keys;
                }
            }
        });
    }
    auto blob_ref = [&](uint32_t c) -> uint32_t {
        if (c & PSY_CHILD_LEAF_BIT) return c;  // leaves and PSY_NONE
        return ((std::min(s->nodes[c].n_keys, 8u) - 1u) << NODE2_KC_SHIFT) | unit_off[c];
    };
    // flattened path records for the trace kernel's instance entry (dscene.cuh PATH_*)
    std::vector<uint4> path_recs(paths.size() * 2);
    std::vector<PsyXform> path_static(paths.size() * 2);
    for (size_t i = 0; i < paths.size(); ++i) {
        const PsyInstance& in = s->instances[dpaths[i].inst];
        uint32_t kind = PATH_EMPTY;
        uint4 r1 = make_uint4(0, 0, 0, 0);
        if (in.kind == PSY_INSTANCE_ASSEMBLY) {
            const PsyAssembly& a = s->assemblies[in.data_index];
            if (a.root != PSY_NONE) {
                kind = PATH_ASSEMBLY;
                r1 = make_uint4(a.root, a.instance_base, in.data_index, 0);
            }
        } else {
            const PsyObject& o = s->objects[in.data_index];
            if (o.kind == PSY_OBJECT_MESH) {
                const DMesh& m = dmeshes[o.index];
                if (m.root != PSY_NONE) {
                    kind = PATH_MESH;
                    r1 = make_uint4(m.root, m.tri_base, m.tv_base, m.n_keys);
                }
            } else {
                kind = s->lights[o.index].kind == PSY_LIGHT_SPHERE ? PATH_SPHERE_LIGHT : PATH_RECT_LIGHT;
                r1 = make_uint4(o.index, 0, 0, 0);
            }
        }
        uint32_t own = in.xform_count > 0 ? 1u : 0u;
        // shading class of a triangle hit on this path (k_trace_pool files finished rays by it)
        uint32_t sh_kind = in.shader_index == PSY_NONE ? (uint32_t)PSY_SHADER_EMIT : s->shaders[in.shader_index].kind;
        uint32_t cls = sh_kind == PSY_SHADER_LAMBERT ? SHADE_LAMBERT : (sh_kind == PSY_SHADER_GGX ? SHADE_GGX : SHADE_EMIT);
        // r0.w: the BVH root of a mesh / assembly path as a blob reference (k_trace_pool); r1.x stays the ABI's (k_trace)
        const uint32_t root2 = (kind == PATH_MESH || kind == PATH_ASSEMBLY) ? blob_ref(r1.x) : 0u;
        path_recs[2 * i] = make_uint4(kind | (own << 3) | (cls << PREC_CLS_SHIFT) | (dpaths[i].xf_count << 8), dpaths[i].xf_base,
                                      dpaths[i].child_base, root2);
        path_recs[2 * i + 1] = r1;
        path_static[2 * i] = paths[i].keys.size() == 1 ? psyh::to_abi(paths[i].keys[0]) : psyh::to_abi(HXf());
        path_static[2 * i + 1] = path_inv[i];
    }

    std::vector<float4> normals4;
    bool any_normals = false;
    for (uint32_t m = 0; m < s->n_meshes; ++m) any_normals |= s->meshes[m].has_normals != 0;
    if (any_normals) {
        normals4.resize(s->n_vert_keys);
        host_parallel_for(s->n_vert_keys, 65536, [&](size_t b0, size_t e0) {
            for (size_t i = b0; i < e0; ++i)
                normals4[i] = make_float4(s->normals[3 * i], s->normals[3 * i + 1], s->normals[3 * i + 2], 0.0f);
        });
    }

    // light instance -> index of the same instance inside its assembly's instance list
    std::vector<uint32_t> light_local(s->n_light_instances, PSY_NONE);
    for (uint32_t a = 0; a < s->n_assemblies; ++a) {
        const PsyAssembly& pa = s->assemblies[a];
        for (uint32_t li = 0; li < pa.light_instance_count; ++li)
            for (uint32_t i = 0; i < pa.instance_count; ++i)
                if (s->instances[pa.instance_base + i].id == s->light_instances[pa.light_instance_base + li].id) {
                    light_local[pa.light_instance_base + li] = i;
                    break;
                }
    }

    lap("index checks, path records");
    PsyScene* sc = new PsyScene();
    sc->ctx = ctx;
    sc->res_x = s->res_x, sc->res_y = s->res_y, sc->spp = s->spp, sc->seed = s->seed;
#define UP(buf, ptr, n)                                                       \
    do {                                                                      \
        cudaError_t e_ = sc->buf.upload(ptr, n, st);                          \
        if (e_ != cudaSuccess) {                                              \
            delete sc;                                                        \
            return fail(PSY_ERR_CUDA, std::string("upload " #buf ": ") + cudaGetErrorString(e_)); \
        }                                                                     \
        sc->upload_bytes += sc->buf.bytes;                                    \
    } while (0)
    UP(nodes, s->nodes, (size_t)s->n_nodes);
    UP(motion, s->motion_bounds, (size_t)s->n_motion_bounds);
    UP(tris, s->tris, (size_t)s->n_tris * 4);
    UP(normals, normals4.data(), normals4.size());
    UP(meshes, dmeshes.data(), dmeshes.size());
    // Device copies of the colours; the product build bakes every XYZ colour that is constant over the shutter
    // (shaders, single-key surface lights, the background) into its own 95-sample spectrum (PSY_COLOR_BAKED).
    std::vector<PsyColor> dev_colors(s->light_colors, s->light_colors + s->n_light_colors);
    std::vector<PsyShader> dev_shaders(s->shaders, s->shaders + s->n_shaders);
    PsyColor dev_background = s->background;
    std::vector<float> baked;
#if !PSY_STRICT
    {
        auto bake = [&](PsyColor& c) {
            if (c.kind != PSY_COLOR_XYZ) return;
            float sp[MENG_NUM_SAMPLES], inv_norm = 0.0f;
            if (!meng_bake(c.a, c.b, c.c, sp, inv_norm)) return;
            const uint32_t index = (uint32_t)(baked.size() / MENG_NUM_SAMPLES);
            baked.insert(baked.end(), sp, sp + MENG_NUM_SAMPLES);
            c.kind = PSY_COLOR_BAKED;
            std::memcpy(&c.a, &index, 4);
            c.b = inv_norm;
            c.c = 0.0f;
        };
        for (auto& sh : dev_shaders) bake(sh.color);
        for (uint32_t i = 0; i < s->n_lights; ++i)
            if (s->lights[i].n_colors == 1 && s->lights[i].color_base < dev_colors.size()) bake(dev_colors[s->lights[i].color_base]);
        bake(dev_background);
    }
#endif
    UP(lights, s->lights, (size_t)s->n_lights);
    UP(light_shape, s->light_shape, (size_t)s->n_light_shape);
    UP(light_colors, dev_colors.data(), dev_colors.size());
    UP(objects, s->objects, (size_t)s->n_objects);
    UP(shaders, dev_shaders.data(), dev_shaders.size());
    UP(meng_baked, baked.data(), baked.size());
    UP(assemblies, s->assemblies, (size_t)s->n_assemblies);
    UP(instances, s->instances, (size_t)s->n_instances);
    UP(light_instances, s->light_instances, (size_t)s->n_light_instances);
    UP(light_inst_local, light_local.data(), light_local.size());
    UP(xforms, s->xforms, (size_t)s->n_xforms);
    UP(light_nodes, s->light_nodes, (size_t)s->n_light_nodes);
    UP(lt_bounds, s->lt_bounds, (size_t)s->n_lt_bounds);
    UP(paths, dpaths.data(), dpaths.size());
    UP(path_xforms, path_xf.data(), path_xf.size());
    UP(path_inv, path_inv.data(), path_inv.size());
    UP(path_recs, path_recs.data(), path_recs.size());
    UP(path_static, path_static.data(), path_static.size());
    UP(cam_xforms, s->camera.xforms, (size_t)s->camera.n_xforms);
    UP(cam_tfovs, s->camera.tfovs, (size_t)s->camera.n_tfovs);
    UP(cam_aperture, s->camera.aperture_radii, (size_t)s->camera.n_aperture_radii);
    UP(cam_focus, s->camera.focus_distances, (size_t)s->camera.n_focus_distances);
    UP(distant, s->distant_lights, (size_t)s->n_distant_lights);
    UP(world_floats, s->world_floats, (size_t)s->n_world_floats);
    UP(meng, &MENG_SPECTRA[0][0], (size_t)186 * 95);
#undef UP
    lap("cudaMalloc + H2D enqueue");
    // Triangle vertices in leaf order, [v0 keys][v1 keys][v2 keys] per triangle, gathered on the device from the
    // vertex keys and the index tuples: the upload carries each vertex once instead of once per use.
    DevBuf dverts, dbad;
    {
        const uint32_t zero = 0;
        cudaError_t e0_ = dbad.upload(&zero, 1, st);
        if (e0_ != cudaSuccess) {
            delete sc;
            return fail(PSY_ERR_CUDA, std::string("upload: ") + cudaGetErrorString(e0_));
        }
        cudaError_t e_ = sc->tri_verts.alloc(tv_total * sizeof(float4));
        if (e_ == cudaSuccess) e_ = dverts.upload(s->verts, (size_t)s->n_vert_keys * 3, st);
        if (e_ != cudaSuccess) {
            delete sc;
            return fail(PSY_ERR_CUDA, std::string("upload vertices: ") + cudaGetErrorString(e_));
        }
        sc->upload_bytes += dverts.bytes;
        for (uint32_t m = 0; m < s->n_meshes; ++m) {
            const DMesh& dm = dmeshes[m];
            const uint64_t items = (uint64_t)dm.tri_count * 3 * dm.n_keys;
            if (items == 0) continue;
            k_gather_tri_verts<<<(unsigned)((items + 255) / 256), 256, 0, st>>>(
                sc->tris.as<uint4>(), dverts.as<float>(), sc->tri_verts.as<float4>(), dm.tri_base, dm.tri_count, dm.n_keys,
                dm.vert_base, s->meshes[m].vert_count, dm.tv_base, dbad.as<uint32_t>());
        }
    }
    // the node blob, built on the device from the ABI arrays just uploaded
    {
        DevBuf doff;
        cudaError_t e_ = sc->node_blob.alloc((size_t)blob_units * 16);
        if (e_ == cudaSuccess) e_ = doff.upload(unit_off.data(), unit_off.size(), st);
        if (e_ != cudaSuccess) {
            delete sc;
            return fail(PSY_ERR_CUDA, std::string("node blob: ") + cudaGetErrorString(e_));
        }
        if (s->n_nodes)
            k_build_node_blob<<<(s->n_nodes + 255) / 256, 256, 0, st>>>(sc->nodes.as<uint4>(), sc->motion.as<float4>(),
                                                                      doff.as<uint32_t>(), s->n_nodes,
                                                                      sc->node_blob.as<uint4>());
        e_ = cudaStreamSynchronize(st);  // doff dies here
        if (e_ != cudaSuccess) {
            delete sc;
            return fail(PSY_ERR_CUDA, std::string("node blob build: ") + cudaGetErrorString(e_));
        }
    }
    uint32_t bad_index = 0;
    cudaError_t e = cudaMemcpyAsync(&bad_index, dbad.p, 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // the staging vectors above die at return
    if (e != cudaSuccess) {
        delete sc;
        return fail(PSY_ERR_CUDA, std::string("scene upload: ") + cudaGetErrorString(e));
    }
    if (bad_index) {
        delete sc;
        return fail(PSY_ERR_INVALID, "triangle vertex index out of range");
    }

    lap("vertex gather, node blob, sync");
    DScene& d = sc->d;
    d.nodes = sc->nodes.as<uint4>();
    d.nodes_tex = 0;
    {
        // A linear texture object over the node blob, when it fits the device's 1D linear texture limit (16-byte
        // texels); a larger scene leaves it 0 and k_trace_pool loads through the LSU pipe instead.
        // PSY_NODE_TEX=0 switches it off for A/B measurements.
        int max_texels = 0;
        CK(cudaDeviceGetAttribute(&max_texels, cudaDevAttrMaxTexture1DLinearWidth, ctx->device));
        auto env_on = [](const char* name) {
            const char* e = getenv(name);
            return !(e && e[0] == '0');
        };
        cudaTextureDesc td{};
        td.readMode = cudaReadModeElementType;
        if (PSY_POOL_NODE_TEX && env_on("PSY_NODE_TEX") && s->n_nodes && blob_units <= (uint64_t)max_texels) {
            cudaResourceDesc rd{};
            rd.resType = cudaResourceTypeLinear;
            rd.res.linear.devPtr = sc->node_blob.p;
            rd.res.linear.desc = cudaCreateChannelDesc<uint4>();
            rd.res.linear.sizeInBytes = (size_t)blob_units * 16;
            cudaTextureObject_t tex = 0;
            CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
            d.nodes_tex = tex;
        }
    }
    d.motion_bounds = sc->motion.as<float4>();
    d.tri_verts = sc->tri_verts.as<float4>();
    d.tris = sc->tris.as<uint4>();
    d.normals = sc->normals.as<float4>();
    d.meshes = sc->meshes.as<DMesh>();
    d.lights = sc->lights.as<PsyLight>();
    d.light_shape = sc->light_shape.as<float>();
    d.light_colors = sc->light_colors.as<PsyColor>();
    d.objects = sc->objects.as<PsyObject>();
    d.shaders = sc->shaders.as<PsyShader>();
    d.assemblies = sc->assemblies.as<PsyAssembly>();
    d.instances = sc->instances.as<PsyInstance>();
    d.light_instances = sc->light_instances.as<PsyInstance>();
    d.light_inst_local = sc->light_inst_local.as<uint32_t>();
    d.xforms = sc->xforms.as<float4>();
    d.light_nodes = sc->light_nodes.as<PsyLightNode>();
    d.lt_bounds = sc->lt_bounds.as<PsyBBox>();
    d.paths = sc->paths.as<DPath>();
    d.path_xforms = sc->path_xforms.as<float4>();
    d.path_inv = sc->path_inv.as<float4>();
    d.path_recs = sc->path_recs.as<uint4>();
    d.path_static = sc->path_static.as<float4>();
    d.nodes2 = sc->node_blob.as<uint4>();
    d.root_bvh = s->assemblies[s->root_assembly].root;
    d.root_bvh2 = blob_ref(d.root_bvh);
    d.root_instance_base = s->assemblies[s->root_assembly].instance_base;
    d.root_assembly = s->root_assembly;
    d.n_paths = (uint32_t)dpaths.size();
    d.cam_xforms = sc->cam_xforms.as<float4>();
    d.cam_tfovs = sc->cam_tfovs.as<float>();
    d.cam_aperture = sc->cam_aperture.as<float>();
    d.cam_focus = sc->cam_focus.as<float>();
    d.n_cam_xforms = s->camera.n_xforms;
    d.n_cam_tfovs = s->camera.n_tfovs;
    d.n_cam_aperture = s->camera.n_aperture_radii;
    d.n_cam_focus = s->camera.n_focus_distances;
    d.background = dev_background;
    d.distant_lights = sc->distant.as<PsyDistantLight>();
    d.world_floats = sc->world_floats.as<float>();
    d.n_distant_lights = s->n_distant_lights;
    {
        // fold of approximate_energy over the world lights (scene/mod.rs:44-53)
        float sum = 0.0f;
        for (uint32_t i = 0; i < s->n_distant_lights; ++i) {
            const PsyDistantLight& l = s->distant_lights[i];
            float a = 0.0f;
            for (uint32_t k = 0; k < l.n_colors; ++k) {
                const PsyColor& c = s->light_colors[l.color_base + k];
                float en = c.kind == PSY_COLOR_BLACKBODY ? (c.a * c.a) * (c.a * c.a) * c.b : c.b;
                a = a + en;
            }
            sum = sum + a / (float)l.n_colors;
        }
        d.world_light_energy_sum = sum;
    }
    d.meng_spectra = sc->meng.as<float>();
    d.meng_baked = sc->meng_baked.as<float>();
    d.res_x = s->res_x;
    d.res_y = s->res_y;
    lap("texture object, DScene");
    *out = sc;
    return PSY_OK;
}

static void scene_destroy_one(PsyScene* scene) {
    cudaSetDevice(scene->ctx->device);
    cudaStreamSynchronize(scene->ctx->stream);
    donate_pool(scene);
    delete scene;
}

int32_t psy_scene_create(PsyContext* ctx, const PsySceneDesc* s, PsyScene** out) {
    if (!ctx || !out) return fail(PSY_ERR_INVALID, "psy_scene_create: null argument");
    int32_t rc = validate(s);
    if (rc) return rc;
    PsyScene* first = nullptr;
    rc = scene_create_one(ctx, s, &first);
    if (rc) return rc;
    // a context over several GPUs replicates the scene (read-only) on each of them
    for (PsyContext* peer : ctx->peers) {
        PsyScene* r = nullptr;
        rc = scene_create_one(peer, s, &r);
        if (rc) {
            for (PsyScene* x : first->replicas) scene_destroy_one(x);
            scene_destroy_one(first);
            return rc;
        }
        first->replicas.push_back(r);
    }
    cudaSetDevice(ctx->device);
    *out = first;
    return PSY_OK;
}

int32_t psy_scene_destroy(PsyScene* scene) {
    if (!scene) return PSY_OK;
    for (PsyScene* r : scene->replicas) scene_destroy_one(r);
    scene_destroy_one(scene);
    return PSY_OK;
}

int32_t psy_trace_rays(PsyScene* sc, const PsyRay* rays, size_t n, PsyHit* hits, PsyHitGeo* geo) {
    if (!sc || (!rays && n) || (!hits && n)) return fail(PSY_ERR_INVALID, "psy_trace_rays: null argument");
    if (n == 0) return PSY_OK;
    if (n > 0x7FFFFFFFull) return fail(PSY_ERR_INVALID, "too many rays in one call");
    std::lock_guard<std::mutex> g(sc->mu);
    CK(cudaSetDevice(sc->ctx->device));
    cudaStream_t st = sc->ctx->stream;
    DevBuf drays, dhits, dgeo;
    CK(drays.upload(rays, n, st));
    CK(dhits.alloc(n * sizeof(PsyHit)));
    if (geo) CK(dgeo.alloc(n * sizeof(PsyHitGeo)));
    const char* kernel_env = getenv("PSY_TRACE_KERNEL");
    if (kernel_env && std::string(kernel_env) == "warp") {
        k_trace_probe<false><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(
            sc->d, drays.as<PsyRay>(), (uint32_t)n, dhits.as<PsyHit>(), geo ? dgeo.as<PsyHitGeo>() : nullptr, nullptr);
    } else {
        // the production path: the rays go through the path pool and k_trace_pool like a wave of a render
        int32_t rc = ensure_pool(sc, n);
        if (rc) return rc;
        rc = ensure_counts(sc, 1);
        if (rc) return rc;
        CK(cudaMemsetAsync(sc->fetch.p, 0, COUNT_SLOTS * 4, st));
        CK(cudaMemsetAsync(sc->qcounts.p, 0, COUNT_SLOTS * SHADE_CLASSES * 4, st));
        PathPool P = make_pool(sc);
        ShadeQueues Q;
        for (int c = 0; c < SHADE_CLASSES; ++c) Q.q[c] = sc->queues.as<uint32_t>() + (size_t)c * sc->pool_cap;
        Q.count = sc->qcounts.as<uint32_t>();
        k_probe_load<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(drays.as<PsyRay>(), (uint32_t)n, P,
                                                                 sc->active_a.as<uint32_t>(), sc->counts.as<uint32_t>());
        k_trace_pool<false><<<sc->pool_trace_blocks, 32 * POOL_WARPS, POOL_SMEM_BYTES, st>>>(
            sc->d, P, sc->active_a.as<uint32_t>(), sc->counts.as<uint32_t>(), sc->fetch.as<uint32_t>(), nullptr, Q,
            sc->trace_scratch.as<uint32_t>());
        k_probe_store<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(sc->d, drays.as<PsyRay>(), (uint32_t)n, P,
                                                                  dhits.as<PsyHit>(), geo ? dgeo.as<PsyHitGeo>() : nullptr);
        // the trace kernel left a shading class per slot for a shade stage that does not follow here
        CK(cudaMemsetAsync(sc->pool_cls.p, 0xFF, n, st));
    }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(hits, dhits.p, n * sizeof(PsyHit), cudaMemcpyDeviceToHost, st));
    if (geo) CK(cudaMemcpyAsync(geo, dgeo.p, n * sizeof(PsyHitGeo), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return PSY_OK;
}

static int32_t render_bucket_impl(PsyScene* sc, uint32_t x, uint32_t y, uint32_t w, uint32_t h, uint32_t spp, uint32_t seed,
                                  float* xyz_out, float* samples_out, PsyStats* stats) {
    if (!sc) return fail(PSY_ERR_INVALID, "null scene");
    if (w == 0 || h == 0 || spp == 0) return fail(PSY_ERR_INVALID, "empty bucket");
    if ((uint64_t)x + w > sc->res_x || (uint64_t)y + h > sc->res_y) return fail(PSY_ERR_INVALID, "bucket outside the image");
    uint64_t n = (uint64_t)w * h * spp;
    if (n > (1ull << 28)) return fail(PSY_ERR_LIMIT, "bucket has more than 2^28 paths");
    std::lock_guard<std::mutex> g(sc->mu);
    CK(cudaSetDevice(sc->ctx->device));
    cudaStream_t st = sc->ctx->stream;
    auto t0 = std::chrono::steady_clock::now();
    int32_t rc = ensure_pool(sc, n);
    if (rc) return rc;
    rc = ensure_counts(sc, 1);
    if (rc) return rc;
    CK(cudaMemsetAsync(sc->counts.p, 0, COUNT_SLOTS * 4, st));
    CK(cudaMemsetAsync(sc->fetch.p, 0, COUNT_SLOTS * 4, st));
    CK(cudaMemsetAsync(sc->qcounts.p, 0, COUNT_SLOTS * SHADE_CLASSES * 4, st));
    CK(sc->bucket_out.alloc((size_t)w * h * 3 * 4));
    if (samples_out) CK(sc->samples_out.alloc(n * 16));
    Bucket b{x, y, w, h, spp, seed};
    sc->events_used = 0;
    StageEvents ev;
    uint64_t launches = 0;
    rc = enqueue_bucket(sc, b, 0, sc->bucket_out.as<float>(), w, x, y, samples_out ? sc->samples_out.as<float>() : nullptr,
                        false, &ev, &launches);
    if (rc) return rc;
    if (xyz_out) CK(cudaMemcpyAsync(xyz_out, sc->bucket_out.p, (size_t)w * h * 12, cudaMemcpyDeviceToHost, st));
    if (samples_out) CK(cudaMemcpyAsync(samples_out, sc->samples_out.p, n * 16, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    uint32_t counts[COUNT_SLOTS];
    CK(cudaMemcpy(counts, sc->counts.p, sizeof(counts), cudaMemcpyDeviceToHost));
    sc->last_trace_ms = sum_ms(ev.trace);
    sc->last_trace_launches = ev.trace.size();
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        for (int wv = 0; wv < WAVES; ++wv) stats->ray_count += counts[wv];
        stats->trace_time = sc->last_trace_ms * 1e-3;
        stats->initial_ray_generation_time = sum_ms(ev.gen) * 1e-3;
        stats->ray_generation_time = sum_ms(ev.shade) * 1e-3;
        stats->sample_writing_time = sum_ms(ev.film) * 1e-3;
        stats->total_time = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        stats->samples = n;
        stats->kernel_launches = launches;
    }
    return PSY_OK;
}

int32_t psy_render_bucket(PsyScene* sc, uint32_t x, uint32_t y, uint32_t w, uint32_t h, uint32_t spp, uint32_t seed,
                          float* xyz_out, PsyStats* stats) {
    return render_bucket_impl(sc, x, y, w, h, spp, seed, xyz_out, nullptr, stats);
}

int32_t psy_render_bucket_samples(PsyScene* sc, uint32_t x, uint32_t y, uint32_t w, uint32_t h, uint32_t spp,
                                  uint32_t seed, float* samples_out, PsyStats* stats) {
    if (!samples_out) return fail(PSY_ERR_INVALID, "samples_out is null");
    return render_bucket_impl(sc, x, y, w, h, spp, seed, nullptr, samples_out, stats);
}

// Paths in flight when the caller does not say (max_samples_per_bucket = 0).  Every wave of a bucket is one launch
// of each kernel, and the later waves of a small bucket cannot fill 148 SMs x 20 warps x 80 rays: measured on the
// C5 scene, 4 Mi paths in flight give 562 Mrays/s, 16 Mi 650, 64 Mi 679.  So: as many as the image has, up to
// 64 Mi (25 GB of the 180 GB).  When the image is shared between GPUs the buckets are the unit of load balancing:
// `buckets_per_gpu` of them per GPU (default 12: the GPUs pull buckets as they finish, so the last one to finish is
// at most one bucket late, ~4 % on average; smaller buckets would cost more than that in thin later waves).
static size_t default_pool(uint64_t total_samples, uint32_t world, uint32_t buckets_per_gpu) {
    const uint64_t lo = 2ull << 20, hi = 64ull << 20;
    if (buckets_per_gpu == 0) buckets_per_gpu = 12;
    uint64_t want = world > 1 ? total_samples / ((uint64_t)buckets_per_gpu * world) : total_samples;
    return (size_t)std::min(hi, std::max(lo, want));
}

static int32_t image_jobs(PsyScene* sc, const PsyRenderParams* p, std::vector<Bucket>& jobs, uint32_t world = 1,
                          bool streaming = false) {
    uint32_t spp = (p && p->spp) ? p->spp : sc->spp;
    uint32_t seed = (p && p->seed_override_enabled) ? p->seed : sc->seed;
    if (spp == 0) return fail(PSY_ERR_INVALID, "spp is zero");
    uint32_t width = sc->res_x, height = sc->res_y, sx = 0, sy = 0;
    if (p && p->crop_enabled) {
        // renderer.rs:100-108
        uint32_t x1 = std::min(p->crop_x1, sc->res_x - 1), y1 = std::min(p->crop_y1, sc->res_y - 1);
        uint32_t x2 = std::min(p->crop_x2, sc->res_x - 1), y2 = std::min(p->crop_y2, sc->res_y - 1);
        if (x2 < x1 || y2 < y1) return fail(PSY_ERR_INVALID, "crop rectangle is inverted");
        width = x2 - x1 + 1, height = y2 - y1 + 1, sx = x1, sy = y1;
    }
    size_t spb = (p && p->max_samples_per_bucket)
                     ? p->max_samples_per_bucket
                     : default_pool((uint64_t)width * height * spp, world, p ? p->buckets_per_gpu : 0);
    if (!(p && p->max_samples_per_bucket)) {
        // ... and no more than the device has room for (~400 bytes per path incl. queues; a pool already allocated counts)
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
            const size_t fit = std::max(sc->pool_cap, (size_t)((double)free_b * 0.8 / 400.0));
            spb = std::min(spb, std::max(fit, (size_t)1 << 16));
        } else {
            cudaGetLastError();
        }
    }
    // the film order of the reference depends on its bucket list: strict renders use the reference's default --spb
    if (p && p->strict_film_order && !p->max_samples_per_bucket) spb = 4096;
    // a caller that wants every finished bucket (Blender's progressive display) gets buckets of a viewable size
    if (streaming && !(p && p->max_samples_per_bucket)) spb = std::min<size_t>(spb, 4u << 20);
    jobs = make_buckets(width, height, sx, sy, spp, seed, spb);
    return PSY_OK;
}

// Shared bucket lists are handed out largest first (stable: Hilbert order among equals), so that the partial
// buckets of the image border — the small change — are what the GPUs even out with at the end of the frame.
static std::vector<size_t> largest_first(const std::vector<Bucket>& all) {
    std::vector<size_t> order(all.size());
    for (size_t j = 0; j < all.size(); ++j) order[j] = j;
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) {
        return (uint64_t)all[a].w * all[a].h * all[a].spp > (uint64_t)all[b].w * all[b].h * all[b].spp;
    });
    return order;
}

static void add_stats(PsyStats* total, const PsyStats& s) {
    total->trace_time = std::max(total->trace_time, s.trace_time);
    total->initial_ray_generation_time = std::max(total->initial_ray_generation_time, s.initial_ray_generation_time);
    total->ray_generation_time = std::max(total->ray_generation_time, s.ray_generation_time);
    total->sample_writing_time = std::max(total->sample_writing_time, s.sample_writing_time);
    total->ray_count += s.ray_count;
    total->accel_node_visits += s.accel_node_visits;
    total->tri_tests += s.tri_tests;
    total->instance_visits += s.instance_visits;
    total->motion_node_visits += s.motion_node_visits;
    total->samples += s.samples;
    total->kernel_launches += s.kernel_launches;
    total->buckets += s.buckets;
}

// Renderer::render over the GPUs of one context (renderer.rs:76-191 with one worker per GPU): a host thread per GPU
// pulls buckets from a shared counter and renders them into its GPU's zero-initialised image; the disjoint tiles are
// then summed onto the first GPU with one grouped ncclReduce over NVLink, and fetched from there.
static int32_t render_image_multi(PsyScene* sc, const PsyRenderParams* p, float* xyz_out, PsyStats* stats, psy_bucket_cb cb,
                                  void* user) {
    PsyContext* ctx = sc->ctx;
    std::vector<PsyScene*> devs{sc};
    devs.insert(devs.end(), sc->replicas.begin(), sc->replicas.end());
    const uint32_t world = (uint32_t)devs.size();
    auto t0 = std::chrono::steady_clock::now();
    std::vector<Bucket> jobs;
    int32_t rc = image_jobs(sc, p, jobs, world, cb != nullptr);
    if (rc) return rc;
    const std::vector<size_t> order = largest_first(jobs);
    std::atomic<size_t> next{0};
    std::vector<int32_t> rcs(world, PSY_OK);
    std::vector<std::string> errs(world);
    std::vector<PsyStats> sts(world);
    std::mutex cb_mu;
    const bool counters = p && p->collect_counters, strict = p && p->strict_film_order;
    std::vector<std::thread> workers;
    for (uint32_t i = 0; i < world; ++i) {
        workers.emplace_back([&, i]() {
            PsyScene* d = devs[i];
            std::lock_guard<std::mutex> g(d->mu);
            auto run = [&]() -> int32_t {
                CK(cudaSetDevice(d->ctx->device));
                if (!d->image.p) CK(d->image.alloc((size_t)d->res_x * d->res_y * 3 * sizeof(float)));
                CK(cudaMemsetAsync(d->image.p, 0, d->image.bytes, d->ctx->stream));
                JobFeed feed;
                feed.max_in_flight = 2;
                feed.next = [&]() -> int64_t {
                    const size_t k = next.fetch_add(1);
                    return k < order.size() ? (int64_t)order[k] : -1;
                };
                return render_buckets(d, jobs, feed, counters, &sts[i], nullptr, cb, user, strict, &cb_mu);
            };
            rcs[i] = run();
            if (rcs[i]) errs[i] = g_err;
        });
    }
    for (auto& w : workers) w.join();
    for (uint32_t i = 0; i < world; ++i)
        if (rcs[i]) return fail(rcs[i], "GPU " + std::to_string(devs[i]->ctx->device) + ": " + errs[i]);
    // gather: the images are disjoint and zero elsewhere, so the sum onto device 0 is the image
    const size_t n_floats = (size_t)sc->res_x * sc->res_y * 3;
    ncclResult_t nr = ctx->nccl.GroupStart();
    for (uint32_t i = 0; i < world && nr == ncclSuccess; ++i)
        nr = ctx->nccl.Reduce(devs[i]->image.p, devs[i]->image.p, n_floats, ncclFloat, ncclSum, 0, ctx->comms[i],
                              devs[i]->ctx->stream);
    ncclResult_t ne = ctx->nccl.GroupEnd();
    if (nr == ncclSuccess) nr = ne;
    if (nr != ncclSuccess) return fail(PSY_ERR_CUDA, std::string("ncclReduce: ") + ctx->nccl.GetErrorString(nr));
    for (uint32_t i = world; i-- > 0;) {
        CK(cudaSetDevice(devs[i]->ctx->device));
        if (i == 0 && xyz_out && !(p && p->device_resident_only))
            CK(cudaMemcpyAsync(xyz_out, sc->image.p, sc->image.bytes, cudaMemcpyDeviceToHost, sc->ctx->stream));
        CK(cudaStreamSynchronize(devs[i]->ctx->stream));
    }
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        for (uint32_t i = 0; i < world; ++i) add_stats(stats, sts[i]);
        stats->total_time = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    sc->last_trace_ms = 0.0, sc->last_trace_launches = 0;
    for (uint32_t i = 0; i < world; ++i) {
        sc->last_trace_ms = std::max(sc->last_trace_ms, devs[i]->last_trace_ms);
        if (i) sc->last_trace_launches += devs[i]->last_trace_launches;
    }
    return PSY_OK;
}

int32_t psy_render_image(PsyScene* sc, const PsyRenderParams* p, float* xyz_out, PsyStats* stats, psy_bucket_cb cb,
                         void* user) {
    if (!sc) return fail(PSY_ERR_INVALID, "null scene");
    bool resident = p && p->device_resident_only;
    if (!xyz_out && !resident) return fail(PSY_ERR_INVALID, "xyz_out is null and device_resident_only is not set");
    if (!sc->replicas.empty()) return render_image_multi(sc, p, xyz_out, stats, cb, user);
    std::lock_guard<std::mutex> g(sc->mu);
    CK(cudaSetDevice(sc->ctx->device));
    std::vector<Bucket> jobs;
    int32_t rc = image_jobs(sc, p, jobs, 1, cb != nullptr);
    if (rc) return rc;
    if (sc->image.p) CK(cudaMemsetAsync(sc->image.p, 0, sc->image.bytes, sc->ctx->stream));
    JobFeed feed = feed_all(jobs.size());
    return render_buckets(sc, jobs, feed, p && p->collect_counters, stats, resident ? nullptr : xyz_out, cb, user,
                          p && p->strict_film_order);
}

int32_t psy_render_image_shard(PsyScene* sc, const PsyRenderParams* p, uint32_t rank, uint32_t world, PsyStats* stats) {
    if (!sc || world == 0 || rank >= world) return fail(PSY_ERR_INVALID, "bad shard arguments");
    if (!sc->replicas.empty()) return fail(PSY_ERR_INVALID, "psy_render_image_shard is the one-process-per-GPU entry");
    std::lock_guard<std::mutex> g(sc->mu);
    CK(cudaSetDevice(sc->ctx->device));
    PsyContext* ctx = sc->ctx;
    std::vector<Bucket> all;
    int32_t rc = image_jobs(sc, p, all, world);
    if (rc) return rc;
    if (sc->image.p) CK(cudaMemsetAsync(sc->image.p, 0, sc->image.bytes, ctx->stream));
    const std::vector<size_t> order = largest_first(all);
    const uint64_t frame = ctx->shard_frame++;
    JobFeed feed;
    if (p && p->work_queue_key && world > 1) {
        // The ranks of this node pull buckets from a counter in shared memory (every rank calls this function once per
        // frame, with the same arguments, and gathers the tiles with a collective before the next frame).
        if (!ctx->wq || ctx->wq_key != p->work_queue_key) {
            std::string err;
            std::unique_ptr<ShmQueue> q(new ShmQueue());
            if (!q->open(p->work_queue_key, err)) return fail(PSY_ERR_INVALID, err);
            ctx->wq = std::move(q);
            ctx->wq_key = p->work_queue_key;
        }
        ShmQueue* q = ctx->wq.get();
        q->begin_frame(frame);
        feed.max_in_flight = 2;
        feed.next = [q, frame, &order]() -> int64_t {
            const uint64_t k = q->pull(frame);
            return k < order.size() ? (int64_t)order[k] : -1;
        };
    } else {
        // No shared counter: a fixed deal by sample count, largest first, each bucket to the rank with the least so far
        // (every rank computes the same assignment); a rank renders its buckets in Hilbert order.
        std::vector<uint64_t> load(world, 0);
        auto mine = std::make_shared<std::vector<size_t>>();
        for (size_t j : order) {
            uint32_t r = 0;
            for (uint32_t k = 1; k < world; ++k)
                if (load[k] < load[r]) r = k;
            load[r] += (uint64_t)all[j].w * all[j].h * all[j].spp;
            if (r == rank) mine->push_back(j);
        }
        std::sort(mine->begin(), mine->end());
        auto pos = std::make_shared<size_t>(0);
        feed.next = [mine, pos]() -> int64_t { return *pos < mine->size() ? (int64_t)(*mine)[(*pos)++] : -1; };
    }
    return render_buckets(sc, all, feed, p && p->collect_counters, stats, nullptr, nullptr, nullptr,
                          p && p->strict_film_order);
}

// SURVEY.md N3 — LBVH topology of one mesh on the GPU (dbuild.cuh).  Host arrays in, host arrays out.
namespace {
// The LBVH topology of one mesh, left on the device: sorted triangle order, and (when the mesh is more than one leaf)
// the kept internal nodes' children / split axes, root = node 0.
struct LbvhDevice {
    DevBuf verts, idx, order, left, right, axis;
    uint32_t n_internal = 0, depth = 0;
    int32_t root = 0;
};
int32_t lbvh_device_topology(PsyContext* ctx, const float* verts, uint32_t n_verts, uint32_t n_keys, const uint32_t* tri_idx,
                             uint32_t n_tris, uint32_t tris_per_leaf, LbvhDevice& L) {
    cudaStream_t st = ctx->stream;
    DevBuf dcent, dbox, dkeys, dkeys2, dvals, dtmp, dleft, dright, daxis, drng, dnpar, dlpar, dkeep, dnew, ddepth;
    CK(L.verts.upload(verts, (size_t)n_verts * n_keys * 3, st));
    CK(L.idx.upload(tri_idx, (size_t)n_tris * 3, st));
    CK(dcent.alloc((size_t)n_tris * sizeof(float3)));
    const uint32_t box_init[6] = {~0u, ~0u, ~0u, 0u, 0u, 0u};
    CK(dbox.upload(box_init, 6, st));
    CK(dkeys.alloc((size_t)n_tris * 8));
    CK(dkeys2.alloc((size_t)n_tris * 8));
    CK(dvals.alloc((size_t)n_tris * 4));
    CK(L.order.alloc((size_t)n_tris * 4));
    const unsigned tb = (n_tris + 255) / 256;
    k_lbvh_centroids<<<tb, 256, 0, st>>>(L.verts.as<float>(), n_keys, L.idx.as<uint32_t>(), n_tris, dcent.as<float3>(),
                                         dbox.as<uint32_t>());
    k_lbvh_morton<<<tb, 256, 0, st>>>(dcent.as<float3>(), n_tris, dbox.as<uint32_t>(), dkeys.as<uint64_t>(), dvals.as<uint32_t>());
    size_t tmp_bytes = 0, scan_bytes = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dkeys.as<uint64_t>(), dkeys2.as<uint64_t>(), dvals.as<uint32_t>(),
                                       L.order.as<uint32_t>(), (int)n_tris, 0, 63, st));
    CK(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)n_tris, st));
    CK(dtmp.alloc(std::max(tmp_bytes, scan_bytes)));
    CK(cub::DeviceRadixSort::SortPairs(dtmp.p, tmp_bytes, dkeys.as<uint64_t>(), dkeys2.as<uint64_t>(), dvals.as<uint32_t>(),
                                       L.order.as<uint32_t>(), (int)n_tris, 0, 63, st));
    L.depth = 0;
    L.n_internal = 0;
    // PSY_GPU_BUILDER=lbvh keeps the radix tree; the default is PLOC over the same sorted triangles (dbuild.cuh)
    const char* builder_env = getenv("PSY_GPU_BUILDER");
    const bool use_ploc = !(builder_env && std::string(builder_env) == "lbvh");
    if (n_tris > tris_per_leaf && use_ploc) {
        const uint32_t n = n_tris, n_nodes = n_tris - 1;
        DevBuf box_a, box_b, cid_a, cid_b, nn, lead, lead_at, stay, stay_at, left, right, axis, cnt, parent, keep, keep_at, start,
            order2, ddepth;
        CK(box_a.alloc((size_t)n * sizeof(PlocBox)));
        CK(box_b.alloc((size_t)n * sizeof(PlocBox)));
        CK(cid_a.alloc((size_t)n * 4));
        CK(cid_b.alloc((size_t)n * 4));
        CK(nn.alloc((size_t)n * 4));
        CK(lead.alloc((size_t)n * 4));
        CK(lead_at.alloc((size_t)n * 4));
        CK(stay.alloc((size_t)n * 4));
        CK(stay_at.alloc((size_t)n * 4));
        CK(left.alloc((size_t)n_nodes * 4));
        CK(right.alloc((size_t)n_nodes * 4));
        CK(axis.alloc((size_t)n_nodes));
        CK(cnt.alloc((size_t)n_nodes * 4));
        CK(parent.alloc(((size_t)n + n_nodes) * 4));
        CK(keep.alloc((size_t)n_nodes * 4));
        CK(keep_at.alloc((size_t)n_nodes * 4));
        CK(start.alloc((size_t)n_nodes * 4));
        CK(order2.alloc((size_t)n * 4));
        CK(ddepth.alloc(4));
        CK(cudaMemsetAsync(ddepth.p, 0, 4, st));
        k_ploc_init<<<tb, 256, 0, st>>>(L.verts.as<float>(), n_keys, L.idx.as<uint32_t>(), L.order.as<uint32_t>(), n,
                                       box_a.as<PlocBox>(), cid_a.as<uint32_t>());
        DevBuf *bi = &box_a, *bo = &box_b, *ci = &cid_a, *co = &cid_b;
        uint32_t m = n, node_base = 0;
        int rounds = 0;
        while (m > 1) {
            const unsigned mb = (m + 255) / 256;
            k_ploc_nearest<<<mb, 256, 0, st>>>(bi->as<PlocBox>(), m, nn.as<uint32_t>());
            k_ploc_flags<<<mb, 256, 0, st>>>(nn.as<uint32_t>(), m, lead.as<uint32_t>(), stay.as<uint32_t>());
            CK(cub::DeviceScan::ExclusiveSum(dtmp.p, scan_bytes, lead.as<uint32_t>(), lead_at.as<uint32_t>(), (int)m, st));
            CK(cub::DeviceScan::ExclusiveSum(dtmp.p, scan_bytes, stay.as<uint32_t>(), stay_at.as<uint32_t>(), (int)m, st));
            k_ploc_merge<<<mb, 256, 0, st>>>(bi->as<PlocBox>(), ci->as<uint32_t>(), nn.as<uint32_t>(), lead.as<uint32_t>(),
                                            lead_at.as<uint32_t>(), stay.as<uint32_t>(), stay_at.as<uint32_t>(), m, n, node_base,
                                            bo->as<PlocBox>(), co->as<uint32_t>(), left.as<uint32_t>(), right.as<uint32_t>(),
                                            axis.as<uint8_t>(), cnt.as<uint32_t>(), parent.as<uint32_t>());
            uint32_t last[2] = {0, 0};
            CK(cudaMemcpyAsync(&last[0], lead.as<uint32_t>() + (m - 1), 4, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(&last[1], lead_at.as<uint32_t>() + (m - 1), 4, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            const uint32_t merges = last[0] + last[1];
            if (merges == 0 || ++rounds > 4096) return fail(PSY_ERR_CUDA, "PLOC made no progress");
            m -= merges;
            node_base += merges;
            std::swap(bi, bo);
            std::swap(ci, co);
        }
        if (node_base != n_nodes) return fail(PSY_ERR_CUDA, "PLOC node count");
        const uint32_t root_node = n_nodes - 1;  // created last
        const unsigned nb = (n_nodes + 255) / 256;
        k_ploc_order<<<tb, 256, 0, st>>>(L.order.as<uint32_t>(), n, root_node, left.as<uint32_t>(), cnt.as<uint32_t>(),
                                        parent.as<uint32_t>(), tris_per_leaf, order2.as<uint32_t>(), ddepth.as<uint32_t>());
        k_ploc_keep<<<nb, 256, 0, st>>>(n, n_nodes, root_node, left.as<uint32_t>(), cnt.as<uint32_t>(), parent.as<uint32_t>(),
                                       tris_per_leaf, keep.as<uint32_t>(), start.as<uint32_t>());
        CK(cub::DeviceScan::ExclusiveSum(dtmp.p, scan_bytes, keep.as<uint32_t>(), keep_at.as<uint32_t>(), (int)n_nodes, st));
        uint32_t last_keep = 0, last_at = 0;
        CK(cudaMemcpyAsync(&last_keep, keep.as<uint32_t>() + (n_nodes - 1), 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&last_at, keep_at.as<uint32_t>() + (n_nodes - 1), 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&L.depth, ddepth.p, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        const uint32_t n_keep = last_keep + last_at;
        CK(L.left.alloc((size_t)n_keep * 4));
        CK(L.right.alloc((size_t)n_keep * 4));
        CK(L.axis.alloc((size_t)n_keep));
        k_ploc_emit<<<nb, 256, 0, st>>>(n, n_nodes, n_keep, root_node, left.as<uint32_t>(), right.as<uint32_t>(),
                                       axis.as<uint8_t>(), cnt.as<uint32_t>(), parent.as<uint32_t>(), keep.as<uint32_t>(),
                                       keep_at.as<uint32_t>(), start.as<uint32_t>(), tris_per_leaf, L.left.as<int32_t>(),
                                       L.right.as<int32_t>(), L.axis.as<uint8_t>());
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(st));
        L.order = std::move(order2);
        L.n_internal = n_keep;
        L.root = 0;
    } else if (n_tris > tris_per_leaf) {
        const int n_nodes = (int)n_tris - 1;  // of the radix tree over single triangles
        CK(dleft.alloc((size_t)n_nodes * 4));
        CK(dright.alloc((size_t)n_nodes * 4));
        CK(daxis.alloc((size_t)n_nodes));
        CK(drng.alloc((size_t)n_nodes * sizeof(int2)));
        CK(dnpar.alloc((size_t)n_nodes * 4));
        CK(dlpar.alloc((size_t)n_tris * 4));
        CK(dkeep.alloc((size_t)n_nodes * 4));
        CK(dnew.alloc((size_t)n_nodes * 4));
        CK(L.left.alloc((size_t)n_nodes * 4));
        CK(L.right.alloc((size_t)n_nodes * 4));
        CK(L.axis.alloc((size_t)n_nodes));
        CK(ddepth.alloc(4));
        CK(cudaMemsetAsync(ddepth.p, 0, 4, st));
        const unsigned nb = ((unsigned)n_nodes + 255) / 256;
        k_lbvh_tree<<<nb, 256, 0, st>>>(dkeys2.as<uint64_t>(), (int)n_tris, dleft.as<int32_t>(), dright.as<int32_t>(),
                                        daxis.as<uint8_t>(), drng.as<int2>(), dnpar.as<int32_t>(), dlpar.as<int32_t>());
        k_lbvh_flag<<<nb, 256, 0, st>>>(drng.as<int2>(), n_nodes, (int)tris_per_leaf, dkeep.as<uint32_t>());
        CK(cub::DeviceScan::ExclusiveSum(dtmp.p, scan_bytes, dkeep.as<uint32_t>(), dnew.as<uint32_t>(), n_nodes, st));
        k_lbvh_emit<<<nb, 256, 0, st>>>(dleft.as<int32_t>(), dright.as<int32_t>(), daxis.as<uint8_t>(), drng.as<int2>(),
                                        dkeep.as<uint32_t>(), dnew.as<uint32_t>(), n_nodes, L.left.as<int32_t>(),
                                        L.right.as<int32_t>(), L.axis.as<uint8_t>());
        k_lbvh_depth<<<tb, 256, 0, st>>>(dnpar.as<int32_t>(), dlpar.as<int32_t>(), dkeep.as<uint32_t>(), (int)n_tris,
                                         ddepth.as<uint32_t>());
        CK(cudaGetLastError());
        uint32_t last_keep = 0, last_new = 0;
        CK(cudaMemcpyAsync(&last_keep, dkeep.as<uint32_t>() + (n_nodes - 1), 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&last_new, dnew.as<uint32_t>() + (n_nodes - 1), 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&L.depth, ddepth.p, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        L.n_internal = last_keep + last_new;  // (the root spans everything: it is kept, and is new node 0)
        L.root = 0;
    } else {
        L.root = ~(int32_t)((0u << 2) | n_tris);  // the whole mesh is one leaf
        CK(cudaStreamSynchronize(st));
    }
    CK(cudaGetLastError());
    return PSY_OK;
}
}  // namespace

int32_t psy_build_mesh_topology(PsyContext* ctx, const float* verts, uint32_t n_verts, uint32_t n_keys, const uint32_t* tri_idx,
                                uint32_t n_tris, uint32_t tris_per_leaf, uint32_t* order, int32_t* left, int32_t* right,
                                uint8_t* axis, uint32_t* n_internal, int32_t* root, uint32_t* depth) {
    if (!ctx || !verts || !tri_idx || !order || !root || !depth || !n_internal || n_keys == 0 || tris_per_leaf == 0 ||
        tris_per_leaf > 3)
        return fail(PSY_ERR_INVALID, "psy_build_mesh_topology: bad argument");
    if (n_tris == 0 || n_tris > (PSY_MAX_LEAF_START >> 2)) return fail(PSY_ERR_INVALID, "psy_build_mesh_topology: triangle count");
    if (n_tris > tris_per_leaf && (!left || !right || !axis)) return fail(PSY_ERR_INVALID, "psy_build_mesh_topology: null output");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    LbvhDevice L;
    const int32_t rc = lbvh_device_topology(ctx, verts, n_verts, n_keys, tri_idx, n_tris, tris_per_leaf, L);
    if (rc != PSY_OK) return rc;
    *depth = L.depth, *n_internal = L.n_internal, *root = L.root;
    if (L.n_internal) {
        CK(cudaMemcpyAsync(left, L.left.p, (size_t)L.n_internal * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(right, L.right.p, (size_t)L.n_internal * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(axis, L.axis.p, (size_t)L.n_internal, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaMemcpyAsync(order, L.order.p, (size_t)n_tris * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    return PSY_OK;
}

// The whole mesh BVH on the device: LBVH topology, then bounds, union rule and BVH4 collapse (dbuild.cuh), left there
// until psy_mesh_bvh_take copies it into the caller's arrays.
struct PsyMeshBvh {
    int device = 0;
    cudaStream_t stream = nullptr;
    uint32_t n_tris = 0;
    DevBuf order, nodes, motion;
    PsyMeshBvhInfo info{};
    float root_bounds[BUILD_MAX_KEYS * 6];
};

int32_t psy_mesh_bvh_build(PsyContext* ctx, const float* verts, uint32_t n_verts, uint32_t n_keys, const uint32_t* tri_idx,
                           uint32_t n_tris, uint32_t tris_per_leaf, uint32_t obj_base, uint32_t node_base, uint32_t key_base,
                           PsyMeshBvh** out, PsyMeshBvhInfo* info) {
    if (!ctx || !verts || !tri_idx || !out || !info || n_keys == 0 || tris_per_leaf == 0 || tris_per_leaf > 3)
        return fail(PSY_ERR_INVALID, "psy_mesh_bvh_build: bad argument");
    if (n_keys > (uint32_t)BUILD_MAX_KEYS) return fail(PSY_ERR_INVALID, "psy_mesh_bvh_build: more time samples than the device build handles");
    if (n_tris <= tris_per_leaf) return fail(PSY_ERR_INVALID, "psy_mesh_bvh_build: a mesh of one leaf has no nodes to build");
    if ((uint64_t)obj_base + n_tris > PSY_MAX_LEAF_START || n_tris > (PSY_MAX_LEAF_START >> 2) ||
        (uint64_t)node_base + n_tris >= 0x7FFFFFFFull)
        return fail(PSY_ERR_INVALID, "psy_mesh_bvh_build: triangle count");
    *out = nullptr;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    LbvhDevice L;
    int32_t rc = lbvh_device_topology(ctx, verts, n_verts, n_keys, tri_idx, n_tris, tris_per_leaf, L);
    if (rc != PSY_OK) return rc;
    const int n_int = (int)L.n_internal;
    if (n_int <= 0) return fail(PSY_ERR_INVALID, "psy_mesh_bvh_build: empty topology");
    std::unique_ptr<PsyMeshBvh> h(new PsyMeshBvh());
    h->device = ctx->device, h->stream = st, h->n_tris = n_tris;
    DevBuf parent, cb, cnk, ready, rootb, rootnk, is4, idx4, mcount, kbase, tmp;
    CK(parent.alloc((size_t)n_int * 4));
    CK(cb.alloc((size_t)n_int * 2 * n_keys * 6 * sizeof(float)));
    CK(cnk.alloc((size_t)n_int * 2));
    CK(ready.alloc((size_t)n_int * 4));
    CK(rootb.alloc(sizeof(float) * BUILD_MAX_KEYS * 6));
    CK(rootnk.alloc(4));
    CK(is4.alloc((size_t)n_int * 4));
    CK(idx4.alloc((size_t)n_int * 4));
    CK(mcount.alloc((size_t)n_int * 4));
    CK(kbase.alloc((size_t)n_int * 4));
    CK(cudaMemsetAsync(ready.p, 0, (size_t)n_int * 4, st));
    CK(cudaMemsetAsync(rootnk.p, 0, 4, st));
    const unsigned nb = ((unsigned)n_int + 255) / 256;
    k_b2_parents<<<nb, 256, 0, st>>>(L.left.as<int32_t>(), L.right.as<int32_t>(), n_int, parent.as<int32_t>());
    k_b2_refit<<<((unsigned)(2 * n_int) + 127) / 128, 128, 0, st>>>(
        L.verts.as<float>(), n_keys, L.idx.as<uint32_t>(), L.order.as<uint32_t>(), L.left.as<int32_t>(), L.right.as<int32_t>(),
        parent.as<int32_t>(), n_int, cb.as<float>(), cnk.as<uint8_t>(), ready.as<uint32_t>(), rootb.as<float>(),
        rootnk.as<uint32_t>());
    k_b4_mark<<<nb, 256, 0, st>>>(L.left.as<int32_t>(), L.right.as<int32_t>(), parent.as<int32_t>(), cnk.as<uint8_t>(), n_int,
                                  is4.as<uint32_t>(), mcount.as<uint32_t>());
    size_t scan_bytes = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (uint32_t*)nullptr, (uint32_t*)nullptr, n_int, st));
    CK(tmp.alloc(scan_bytes));
    CK(cub::DeviceScan::ExclusiveSum(tmp.p, scan_bytes, is4.as<uint32_t>(), idx4.as<uint32_t>(), n_int, st));
    CK(cub::DeviceScan::ExclusiveSum(tmp.p, scan_bytes, mcount.as<uint32_t>(), kbase.as<uint32_t>(), n_int, st));
    uint32_t last[4] = {0, 0, 0, 0}, root_nk = 0;
    CK(cudaMemcpyAsync(&last[0], is4.as<uint32_t>() + (n_int - 1), 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&last[1], idx4.as<uint32_t>() + (n_int - 1), 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&last[2], mcount.as<uint32_t>() + (n_int - 1), 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&last[3], kbase.as<uint32_t>() + (n_int - 1), 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&root_nk, rootnk.p, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h->root_bounds, rootb.p, sizeof(h->root_bounds), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    const uint32_t n4 = last[0] + last[1], nm = last[2] + last[3];
    if (root_nk == 0 || n4 == 0) return fail(PSY_ERR_CUDA, "psy_mesh_bvh_build: the refit did not reach the root");
    CK(h->nodes.alloc((size_t)n4 * sizeof(PsyBvh4Node)));
    CK(h->motion.alloc(std::max<size_t>((size_t)nm, 1) * sizeof(PsyBBox4)));
    k_b4_emit<<<((unsigned)n_int + 127) / 128, 128, 0, st>>>(
        L.left.as<int32_t>(), L.right.as<int32_t>(), L.axis.as<uint8_t>(), cb.as<float>(), cnk.as<uint8_t>(), is4.as<uint32_t>(),
        idx4.as<uint32_t>(), kbase.as<uint32_t>(), n_int, n_keys, obj_base, node_base, key_base, h->nodes.as<PsyBvh4Node>(),
        h->motion.as<PsyBBox4>());
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    h->order = std::move(L.order);
    h->info.n_nodes = n4;
    h->info.n_motion = nm;
    h->info.n_root_bounds = root_nk;
    h->info.root = node_base;              // the binary root is at depth 0: this mesh's first BVH4 node
    h->info.depth = L.depth / 2 + 1;       // like BVH4::construct_from_base (bvh4.rs:186-200)
    h->info.binary_depth = L.depth;
    *info = h->info;
    *out = h.release();
    return PSY_OK;
}

int32_t psy_mesh_bvh_take(PsyMeshBvh* h, uint32_t* order, PsyBvh4Node* nodes, PsyBBox4* motion, PsyBBox* root_bounds) {
    if (!h) return PSY_OK;
    std::unique_ptr<PsyMeshBvh> own(h);
    if (!order && !nodes && !motion && !root_bounds) return PSY_OK;  // just dropped
    if (!order || !nodes || !root_bounds || (h->info.n_motion && !motion)) return fail(PSY_ERR_INVALID, "psy_mesh_bvh_take: null output");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    CK(cudaMemcpyAsync(order, h->order.p, (size_t)h->n_tris * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(nodes, h->nodes.p, (size_t)h->info.n_nodes * sizeof(PsyBvh4Node), cudaMemcpyDeviceToHost, st));
    if (h->info.n_motion)
        CK(cudaMemcpyAsync(motion, h->motion.p, (size_t)h->info.n_motion * sizeof(PsyBBox4), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (uint32_t k = 0; k < h->info.n_root_bounds; ++k)
        for (int d = 0; d < 3; ++d) root_bounds[k].min[d] = h->root_bounds[6 * k + d], root_bounds[k].max[d] = h->root_bounds[6 * k + 3 + d];
    return PSY_OK;
}

// Page-locks caller memory (scene arrays, result images) so that the copies to and from the device run at the
// link's rate instead of through the driver's staging buffers; a convenience over cudaHostRegister for hosts that
// do not link the CUDA runtime themselves.
int32_t psy_host_register(void* ptr, size_t bytes) {
    if (!ptr || bytes == 0) return PSY_OK;
    cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
    if (e == cudaErrorHostMemoryAlreadyRegistered) {
        cudaGetLastError();
        return PSY_OK;
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(PSY_ERR_CUDA, std::string("cudaHostRegister: ") + cudaGetErrorString(e));
    }
    return PSY_OK;
}
int32_t psy_host_unregister(void* ptr) {
    if (!ptr) return PSY_OK;
    cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) cudaGetLastError();
    return PSY_OK;
}

// The shared work counter on its own (no GPU involved): for a host that schedules buckets itself, like the reference's
// Renderer::render with psy_render_bucket as its render_job body, and for the CPU tests.
int32_t psy_work_queue_open(uint32_t key, PsyWorkQueue** out) {
    if (!out || key == 0) return fail(PSY_ERR_INVALID, "psy_work_queue_open: null argument or zero key");
    std::unique_ptr<ShmQueue> q(new ShmQueue());
    std::string err;
    if (!q->open(key, err)) return fail(PSY_ERR_INVALID, err);
    *out = reinterpret_cast<PsyWorkQueue*>(q.release());
    return PSY_OK;
}
int32_t psy_work_queue_pull(PsyWorkQueue* q, uint64_t frame, uint64_t* index) {
    if (!q || !index) return fail(PSY_ERR_INVALID, "psy_work_queue_pull: null argument");
    ShmQueue* s = reinterpret_cast<ShmQueue*>(q);
    s->begin_frame(frame);
    *index = s->pull(frame);
    return PSY_OK;
}
int32_t psy_work_queue_close(PsyWorkQueue* q) {
    delete reinterpret_cast<ShmQueue*>(q);
    return PSY_OK;
}

int32_t psy_device_image(PsyScene* sc, void** dev_ptr, size_t* n_floats) {
    if (!sc || !dev_ptr || !n_floats) return fail(PSY_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> g(sc->mu);
    CK(cudaSetDevice(sc->ctx->device));
    if (!sc->image.p) {
        CK(sc->image.alloc((size_t)sc->res_x * sc->res_y * 3 * sizeof(float)));
        CK(cudaMemsetAsync(sc->image.p, 0, sc->image.bytes, sc->ctx->stream));
        CK(cudaStreamSynchronize(sc->ctx->stream));
    }
    *dev_ptr = sc->image.p;
    *n_floats = (size_t)sc->res_x * sc->res_y * 3;
    return PSY_OK;
}

int32_t psy_scene_stream(PsyScene* sc, void** cuda_stream) {
    if (!sc || !cuda_stream) return fail(PSY_ERR_INVALID, "null argument");
    *cuda_stream = (void*)sc->ctx->stream;
    return PSY_OK;
}

int32_t psy_fetch_image(PsyScene* sc, float* xyz_out) {
    if (!sc || !xyz_out) return fail(PSY_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> g(sc->mu);
    CK(cudaSetDevice(sc->ctx->device));
    if (!sc->image.p) return fail(PSY_ERR_INVALID, "no image has been rendered yet");
    CK(cudaMemcpyAsync(xyz_out, sc->image.p, sc->image.bytes, cudaMemcpyDeviceToHost, sc->ctx->stream));
    CK(cudaStreamSynchronize(sc->ctx->stream));
    return PSY_OK;
}

int32_t psy_last_trace_timing(PsyScene* sc, double* total_ms, uint64_t* launches) {
    if (!sc) return fail(PSY_ERR_INVALID, "null scene");
    if (total_ms) *total_ms = sc->last_trace_ms;
    if (launches) *launches = sc->last_trace_launches;
    return PSY_OK;
}

}  // extern "C"
