// kernels.cuh — the wavefront stages.  One path owns one slot of the SoA path pool; every wave
// traces exactly one ray per live path (camera/bounce ray or shadow ray), like the reference's
// wave loop (src/renderer.rs:279-297).  The kernels mark slots with one byte; the class queues and
// the next wave's live list are built from those bytes in slot order (dpartition.cuh).
//
//   k_generate      <- renderer.rs:241-275  (sampler, pixel filter, LightPath::new, Camera::generate_ray)
//   k_trace_pool    <- Tracer::trace        (src/tracer.rs:34-37; dpool.cuh — the production trace kernel)
//   k_trace         <- the same traversal with one ray per lane (dtraverse.cuh): counting build, A/B partner
//   k_shade<SIMPLE/EMIT>, k_shade_select / _light / _bounce<LAMBERT/GGX>
//                   <- LightPath::next      (renderer.rs:449-691)
//   k_film          <- renderer.rs:299-310  (XYZ::from_spectral_sample / spp, summed per pixel)
#pragma once
#include <math_constants.h>

#include "dshade.cuh"
#include "dtraverse.cuh"

namespace psyd {

// flags word of a path
constexpr uint32_t F_OCCLUSION = 1u;       // the ray now in the slot is a shadow ray
constexpr uint32_t F_EVENT_SHIFT = 1;      // 0 camera, 1 bounce, 2 shadow
constexpr uint32_t F_EVENT_MASK = 3u << 1;
constexpr uint32_t F_BOUNCE_SHIFT = 3;
constexpr uint32_t F_BOUNCE_MASK = 3u << 3;
constexpr uint32_t F_HAS_NEXT = 1u << 5;
constexpr uint32_t EV_CAMERA = 0, EV_BOUNCE = 1, EV_SHADOW = 2;

struct PathPool {
    float4* ray_o;  // orig.xyz, time
    float4* ray_d;  // dir.xyz, max_t
    uint32_t* flags;
    uint32_t* sseed;
    float* wavelength;
    float* closure_pdf;
    float4* nb_o;  // next bounce ray origin
    float4* nb_d;  // next bounce ray direction
    float4* next_atten;
    float4* light_atten;
    float4* pending;
    float4* color;
    HitRec* hits;
    uint4* light_pick;  // k_shade_select -> k_shade: LightPick (kind, index, parent path, selection pdf)
    float2* light_uv;   // the two light-sampling random numbers of the same Sobol vector
    float4* sh_o;       // k_shade_light -> k_shade: shadow ray origin, w = 1 when a light sample was found
    float4* sh_d;       // shadow ray direction, w = max_t
    // k_shade_light -> k_shade_bounce: what the bounce needs of the finalised hit and the prepared closure, so that
    // neither is computed twice: (pos, pos_err) (nor_g, roughness) (incoming, fresnel) (spectral colour)
    // (flipped shading normal) (flipped geometric normal)
    float4* ic[6];
    // one byte per slot, 0xFF when idle: the shading class of the ray traced in this wave (k_trace* -> class queues),
    // and 0 for a path that continues (shade kernels -> next wave's live list); see dpartition.cuh
    uint8_t* cls;
    uint8_t* alive;
    // strict film order only: bit w is set when the path continued after wave w (k_record_alive)
    uint8_t* whist;
};

struct Bucket {
    uint32_t x, y, w, h, spp, seed;
};

// Shading classes: k_trace sorts the paths of a wave into one queue per class so that the warps
// of the shade kernels run one code path (Laine et al.'s wavefront formulation, here on the
// reference's closure / event kinds).
constexpr int SHADE_SIMPLE = 0;   // shadow-ray result, or a camera/bounce ray that missed
constexpr int SHADE_EMIT = 1;     // hit an emitter (Emit closure) or light geometry
constexpr int SHADE_LAMBERT = 2;  // surface hit, Lambert closure
constexpr int SHADE_GGX = 3;      // surface hit, GGX closure
constexpr int SHADE_CLASSES = 4;
// Persistent trace kernel tuning: stack pops between two refill checks, idle lanes that trigger a refill
struct TraceTuning {
    int step_budget;  // default 24
    int refill_min;   // idle lanes that trigger a refill; 32 = classic "fetch a full warp of rays when the warp is empty"
    int leaf_min;     // parked triangle leaves that trigger their joint processing
    int inst_min;     // parked instance entries that trigger their joint processing
};

struct ShadeQueues {
    uint32_t* q[SHADE_CLASSES];  // path slots per class
    uint32_t* count;             // SHADE_CLASSES counters of the current wave
};

// Camera::generate_ray (camera.rs:74-102)
PSY_DEV void camera_ray(const DScene& S, float x, float y, float time, float u, float v, V3& orig, V3& dir) {
    Xf transform = lerp_slice_xf(S.cam_xforms, S.n_cam_xforms, time);
    float tfov = lerp_slice_f(S.cam_tfovs, S.n_cam_tfovs, time);
    float aperture_radius = lerp_slice_f(S.cam_aperture, S.n_cam_aperture, time);
    float focus_distance = lerp_slice_f(S.cam_focus, S.n_cam_focus, time);
    float cu, cv;
    square_to_circle((u * 2.0f) - 1.0f, (v * 2.0f) - 1.0f, cu, cv);
    V3 o = mk3(aperture_radius * cu, aperture_radius * cv, 0.0f);
    V3 d = normalized(mk3((x * tfov) - (o.x / focus_distance), (y * tfov) - (o.y / focus_distance), 1.0f));
    orig = xf_point(transform, o);
    dir = xf_vector(transform, d);
}

__global__ void __launch_bounds__(256) k_generate(DScene S, PathPool P, Bucket b, uint32_t n, uint32_t* active,
                                                  uint32_t* count0) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *count0 = n;
    if (i >= n) return;
    uint32_t q = i / b.spp, si = i - q * b.spp;
    uint32_t px = b.x + (q % b.w), py = b.y + (q / b.w);
    float s0[4], s1[4];
    get_sample_4d(si, 0, px, py, b.seed, s0);
    get_sample_4d(si, 1, px, py, b.seed, s1);
    float cmpx = 1.0f / (float)S.res_x, cmpy = 1.0f / (float)S.res_y;
    float min_y = -((float)S.res_y / (float)S.res_x), max_y = (float)S.res_y / (float)S.res_x;
    float x_extent = 1.0f - -1.0f, y_extent = max_y - min_y;
    float filter_x = probit(s0[3], 2.0f / 6.0f) + 0.5f;
    float filter_y = probit(s1[0], 2.0f / 6.0f) + 0.5f;
    float samp_x = (filter_x + (float)px) * cmpx;
    float samp_y = (filter_y + (float)py) * cmpy;
    float img_x = (samp_x - 0.5f) * x_extent;
    float img_y = (0.5f - samp_y) * y_extent;
    float wl = map_0_1_to_wavelength(golden_ratio_sample(si, hash_u32((px << 16) ^ py, b.seed)));
    V3 o, d;
    camera_ray(S, img_x, img_y, s0[0], s0[1], s0[2], o, d);
    P.ray_o[i] = make_float4(o.x, o.y, o.z, s0[0]);
    P.ray_d[i] = make_float4(d.x, d.y, d.z, CUDART_INF_F);
    P.flags[i] = (EV_CAMERA << F_EVENT_SHIFT);
    P.sseed[i] = b.seed ^ 0x40d4682bu;
    P.wavelength[i] = wl;
    P.closure_pdf[i] = 1.0f;
    P.next_atten[i] = make_float4(1, 1, 1, 1);
    P.light_atten[i] = make_float4(1, 1, 1, 1);
    P.pending[i] = make_float4(0, 0, 0, 0);
    P.color[i] = make_float4(0, 0, 0, 0);
    active[i] = i;
}

// Persistent trace kernel: warps fetch 32 rays at a time from a global counter.
template <bool COUNT>
#ifndef PSY_TRACE_MIN_BLOCKS
#define PSY_TRACE_MIN_BLOCKS 4
#endif
__global__ void __launch_bounds__(128, PSY_TRACE_MIN_BLOCKS) k_trace(DScene S, PathPool P, const uint32_t* __restrict__ active,
                                               const uint32_t* __restrict__ count, uint32_t* fetch,
                                               TraceCounters* counters, ShadeQueues Q, TraceTuning tune) {
    const uint32_t n = *count;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t FULL = 0xffffffffu;
    TraceCounters cnt;
    cnt.node_visits = cnt.tri_tests = cnt.instance_visits = cnt.motion_node_visits = 0;
    TraversalMem mem;
    Traversal<COUNT> tr;
    tr.attach(mem);
    bool busy = false;       // this lane holds a ray
    bool exhausted = false;  // warp-uniform: the global ray queue is empty
    uint32_t slot = 0;
    for (;;) {
        // Lane refill: finished lanes take new rays as soon as tune.refill_min of them are idle, so a
        // long ray does not hold 31 empty lanes hostage.
        const uint32_t idle = __ballot_sync(FULL, !busy);
        if (!exhausted && (idle == FULL || __popc(idle) >= tune.refill_min)) {
            const uint32_t want = __popc(idle);
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(fetch, want);
            base = __shfl_sync(FULL, base, 0);
            if (base + want >= n) exhausted = true;
            if (!busy) {
                uint32_t i = base + __popc(idle & ((1u << lane) - 1u));
                if (i < n) {
                    slot = active[i];
                    float4 ro = P.ray_o[slot], rd = P.ray_d[slot];
                    RayIn ray;
                    ray.orig = mk3(ro.x, ro.y, ro.z);
                    ray.dir = mk3(rd.x, rd.y, rd.z);
                    ray.time = ro.w;
                    ray.max_t = rd.w;
                    ray.occlusion = (P.flags[slot] & F_OCCLUSION) != 0u;
                    tr.begin(S, ray);
                    busy = true;
                }
            }
        }
        if (__ballot_sync(FULL, busy) == 0u) break;
        tr.run_warp(S, tune.step_budget, exhausted ? 33 : tune.refill_min, tune.leaf_min, tune.inst_min, cnt);
        int cls = -1;
        if (busy && tr.done()) {
            busy = false;
            const HitRec& hit = tr.hit;
            uint4* hp = reinterpret_cast<uint4*>(P.hits + slot);
            hp[0] = make_uint4(hit.kind, __float_as_uint(hit.t), hit.prim, hit.path);
            hp[1] = make_uint4(__float_as_uint(hit.b0), __float_as_uint(hit.b1), __float_as_uint(hit.b2), 0u);
            // shading class of this path for the next stage (see k_shade)
            if (tr.occlusion || hit.kind == PSY_HIT_MISS) {
                cls = SHADE_SIMPLE;
            } else if (hit.kind != PSY_HIT_TRIANGLE) {
                cls = SHADE_EMIT;
            } else {
                uint32_t sh = S.instances[S.paths[hit.path].inst].shader_index;
                uint32_t kind = sh == PSY_NONE ? (uint32_t)PSY_SHADER_EMIT : S.shaders[sh].kind;
                cls = kind == PSY_SHADER_LAMBERT ? SHADE_LAMBERT : (kind == PSY_SHADER_GGX ? SHADE_GGX : SHADE_EMIT);
            }
        }
        // the path's shading class; the class queues are built from these bytes in slot order (dpartition.cuh)
        if (cls >= 0) P.cls[slot] = (uint8_t)cls;
    }
    if (COUNT) {
        atomicAdd(&counters->node_visits, cnt.node_visits);
        atomicAdd(&counters->tri_tests, cnt.tri_tests);
        atomicAdd(&counters->instance_visits, cnt.instance_visits);
        atomicAdd(&counters->motion_node_visits, cnt.motion_node_visits);
    }
}

// HitRec -> the ABI's PsyHit / PsyHitGeo of probe ray i (psy_trace_rays)
PSY_DEV void probe_store(const DScene& S, const HitRec& h, const RayIn& ray, uint32_t i, PsyHit* hits, PsyHitGeo* geo) {
    PsyHit o;
    o.kind = h.kind;
    o.t = h.t;
    o.prim_id = 0;
    o.inst_id = 0;
    o.path_id = 0;
    o.b0 = h.b0, o.b1 = h.b1, o.b2 = h.b2;
    bool is_hit = h.kind == PSY_HIT_TRIANGLE || h.kind == PSY_HIT_SPHERE_LIGHT || h.kind == PSY_HIT_RECT_LIGHT;
    if (is_hit) {
        o.path_id = h.path;
        o.inst_id = S.instances[S.paths[h.path].inst].id;
        o.prim_id = h.kind == PSY_HIT_TRIANGLE ? S.tris[h.prim].w : h.prim;
    }
    hits[i] = o;
    if (geo) {
        PsyHitGeo g;
        g.pos[0] = g.pos[1] = g.pos[2] = 0.0f;
        g.pos_err = 0.0f;
        g.nor[0] = g.nor[1] = g.nor[2] = 0.0f;
        g.sample_pdf = 0.0f;
        g.nor_g[0] = g.nor_g[1] = g.nor_g[2] = 0.0f;
        g.closure_kind = 0;
        if (is_hit) {
            Isect is;
            finalize_hit(S, h, ray.orig, ray.dir, ray.time, is);
            g.pos[0] = is.pos.x, g.pos[1] = is.pos.y, g.pos[2] = is.pos.z;
            g.pos_err = is.pos_err;
            g.nor[0] = is.nor.x, g.nor[1] = is.nor.y, g.nor[2] = is.nor.z;
            g.sample_pdf = is.sample_pdf;
            g.nor_g[0] = is.nor_g.x, g.nor_g[1] = is.nor_g.y, g.nor_g[2] = is.nor_g.z;
            g.closure_kind = is.closure.kind;
        }
        geo[i] = g;
    }
}

// Scene upload: the leaf-ordered triangle vertex array of one mesh, [v0 keys][v1 keys][v2 keys] per triangle
// (the original triangle index rides in v0-key0 .w), from the mesh's vertex keys and (v0, v2, v1, orig) tuples.
// The vertex indices are validated here as well (a bad one raises *bad_index and reads vertex 0): the host does not
// walk the triangle list at upload.
__global__ void __launch_bounds__(256) k_gather_tri_verts(const uint4* __restrict__ tris, const float* __restrict__ verts,
                                                          float4* __restrict__ out, uint32_t tri_base, uint32_t tri_count,
                                                          uint32_t k, uint32_t vert_base, uint32_t vert_count,
                                                          uint32_t tv_base, uint32_t* __restrict__ bad_index) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (uint64_t)tri_count * 3u * k) return;
    const uint32_t key = (uint32_t)(i % k), j = (uint32_t)((i / k) % 3u), t = (uint32_t)(i / (3u * (uint64_t)k));
    const uint4 idx = tris[tri_base + t];
    uint32_t vi = j == 0u ? idx.x : (j == 1u ? idx.y : idx.z);
    if (vi >= vert_count) {
        *bad_index = 1u;
        vi = 0u;
    }
    const float* v = verts + ((size_t)vert_base + (size_t)vi * k + key) * 3u;
    const float w = (j == 0u && key == 0u) ? __uint_as_float(idx.w) : 0.0f;
    out[(size_t)tv_base + i] = make_float4(v[0], v[1], v[2], w);
}

// Scene upload: the node blob of k_trace_pool (dscene.cuh) from the ABI's node and motion-key arrays.
// unit_off[i] = first 16-byte unit of node i.  One thread per node.
PSY_DEV uint32_t node2_ref(uint32_t c, const uint4* __restrict__ nodes, const uint32_t* __restrict__ unit_off) {
    if (c & PSY_CHILD_LEAF_BIT) return c;  // leaves and PSY_NONE
    const uint32_t nk = nodes[(size_t)c * 8u + 1u].y;
    return ((min(nk, 8u) - 1u) << NODE2_KC_SHIFT) | unit_off[c];
}
__global__ void __launch_bounds__(256) k_build_node_blob(const uint4* __restrict__ nodes, const float4* __restrict__ motion,
                                                         const uint32_t* __restrict__ unit_off, uint32_t n_nodes,
                                                         uint4* __restrict__ blob) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    const uint4* src = nodes + (size_t)i * 8u;
    const uint4 child = src[0], meta = src[1];
    uint4* dst = blob + unit_off[i];
    dst[0] = make_uint4(node2_ref(child.x, nodes, unit_off), node2_ref(child.y, nodes, unit_off),
                        node2_ref(child.z, nodes, unit_off), node2_ref(child.w, nodes, unit_off));
    dst[1] = make_uint4(meta.x, meta.y, 0u, meta.w);
    if (meta.y == 1u) {
        for (int k = 0; k < 6; ++k) dst[2 + k] = src[2 + k];
    } else {
        const uint4* keys = reinterpret_cast<const uint4*>(motion) + (size_t)meta.z * 6u;
        for (uint32_t k = 0; k < meta.y * 6u; ++k) dst[2 + k] = keys[k];
    }
}

// Probe rays through the production trace kernel: load them into the path pool ...
__global__ void __launch_bounds__(256) k_probe_load(const PsyRay* __restrict__ rays, uint32_t n, PathPool P, uint32_t* active,
                                                    uint32_t* count0) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *count0 = n;
    if (i >= n) return;
    const PsyRay r = rays[i];
    P.ray_o[i] = make_float4(r.orig[0], r.orig[1], r.orig[2], r.time);
    P.ray_d[i] = make_float4(r.dir[0], r.dir[1], r.dir[2], r.max_t);
    P.flags[i] = (r.flags & PSY_RAY_OCCLUSION) ? F_OCCLUSION : 0u;
    active[i] = i;
}
// ... and turn the hit records k_trace_pool left in the pool into the ABI's structs
__global__ void __launch_bounds__(128) k_probe_store(DScene S, const PsyRay* __restrict__ rays, uint32_t n, PathPool P,
                                                     PsyHit* hits, PsyHitGeo* geo) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const PsyRay r = rays[i];
    RayIn ray;
    ray.orig = mk3(r.orig[0], r.orig[1], r.orig[2]);
    ray.dir = mk3(r.dir[0], r.dir[1], r.dir[2]);
    ray.time = r.time;
    ray.max_t = r.max_t;
    ray.occlusion = (r.flags & PSY_RAY_OCCLUSION) != 0u;
    probe_store(S, P.hits[i], ray, i, hits, geo);
}

// Probe rays: trace + finalise into the ABI's PsyHit / PsyHitGeo (psy_trace_rays)
template <bool COUNT>
__global__ void __launch_bounds__(128) k_trace_probe(DScene S, const PsyRay* __restrict__ rays, uint32_t n, PsyHit* hits,
                                                     PsyHitGeo* geo, TraceCounters* counters) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    TraceCounters cnt;
    cnt.node_visits = cnt.tri_tests = cnt.instance_visits = cnt.motion_node_visits = 0;
    TraversalMem mem;
    Traversal<COUNT> tr;
    tr.attach(mem);
    RayIn ray;
    ray.orig = ray.dir = mk3(0, 0, 0);
    ray.time = 0.0f, ray.max_t = 0.0f, ray.occlusion = false;
    if (i < n) {
        PsyRay r = rays[i];
        ray.orig = mk3(r.orig[0], r.orig[1], r.orig[2]);
        ray.dir = mk3(r.dir[0], r.dir[1], r.dir[2]);
        ray.time = r.time;
        ray.max_t = r.max_t;
        ray.occlusion = (r.flags & PSY_RAY_OCCLUSION) != 0u;
        tr.begin(S, ray);
    }
    tr.run_warp(S, 1 << 30, 33, 12, 12, cnt);  // every lane of the warp takes part (lanes past n hold no ray)
    if (i < n) probe_store(S, tr.hit, ray, i, hits, geo);
    if (COUNT) {
        atomicAdd(&counters->node_visits, cnt.node_visits);
        atomicAdd(&counters->tri_tests, cnt.tri_tests);
        atomicAdd(&counters->instance_visits, cnt.instance_visits);
        atomicAdd(&counters->motion_node_visits, cnt.motion_node_visits);
    }
}

// LightPath::next for one path of shading class CLS; returns true when the path stays alive (a new
// ray was written).  The class is decided by k_trace, so each instantiation only carries its own code:
//   SHADE_SIMPLE   shadow-ray results and camera/bounce rays that missed (renderer.rs:659-689)
//   SHADE_EMIT     hits on emitters and light geometry                    (renderer.rs:472-483)
//   SHADE_LAMBERT / SHADE_GGX   surface hits: NEE + bounce                (renderer.rs:485-657)
template <int CLS>
PSY_DEV bool path_next(const DScene& S, const PathPool& P, uint32_t slot, const Bucket& b) {
    uint32_t flags = P.flags[slot];
    const uint32_t event = (flags & F_EVENT_MASK) >> F_EVENT_SHIFT;
    float4 ro = P.ray_o[slot];
    const float time = ro.w;
    const float wl = P.wavelength[slot];

    if (CLS == SHADE_SIMPLE) {
        if (event == EV_SHADOW) {
            const uint32_t hit_kind = P.hits[slot].kind;
            if (hit_kind == PSY_HIT_MISS) P.color[slot] = to_f4(from_f4(P.color[slot]) + from_f4(P.pending[slot]));
            if (flags & F_HAS_NEXT) {
                float4 o = P.nb_o[slot], d = P.nb_d[slot];
                P.ray_o[slot] = make_float4(o.x, o.y, o.z, time);
                P.ray_d[slot] = make_float4(d.x, d.y, d.z, CUDART_INF_F);
                P.light_atten[slot] = to_f4(from_f4(P.light_atten[slot]) * from_f4(P.next_atten[slot]));
                flags = (flags & ~(F_EVENT_MASK | F_OCCLUSION)) | (EV_BOUNCE << F_EVENT_SHIFT);
                P.flags[slot] = flags;
                return true;
            }
            return false;
        }
        // background
        V4 bg = color_to_spectral(S, S.background, wl);
        P.color[slot] = to_f4(from_f4(P.color[slot]) + ((bg * from_f4(P.light_atten[slot])) / P.closure_pdf[slot]));
        return false;
    }

    const HitRec h = P.hits[slot];
    float4 rd = P.ray_d[slot];
    V3 orig = mk3(ro.x, ro.y, ro.z), dir = mk3(rd.x, rd.y, rd.z);
    V4 light_attenuation = from_f4(P.light_atten[slot]);
    float closure_sample_pdf = P.closure_pdf[slot];
    Isect is;
    finalize_hit(S, h, orig, dir, time, is);

    if (CLS == SHADE_EMIT) {
        V4 c = color_to_spectral(S, is.closure.color, wl);
        V4 col = from_f4(P.color[slot]);
        if (event == EV_CAMERA) {
            col = col + c;
        } else {
            float mis_pdf = power_heuristic(closure_sample_pdf, is.sample_pdf);
            col = col + ((c * light_attenuation) / mis_pdf);
        }
        P.color[slot] = to_f4(col);
        return false;
    }
    static_assert(CLS == SHADE_SIMPLE || CLS == SHADE_EMIT, "surface hits are shaded by k_shade_select/_light/_bounce");
    return false;
}

#ifndef PSY_LIGHT_THREADS
#define PSY_LIGHT_THREADS 512
#endif

// Last third of a surface hit (classes LAMBERT / GGX): the bounce sample and the path's new state
// (renderer.rs:598-657), after k_shade_select and k_shade_light.  Survivors are appended to the next wave's
// live list.  Like k_shade_light the block fills an SM and its warps are kept in step by barriers between the
// big out-of-line pieces, so that the SM's warps share the instruction cache lines they run.
template <int CLS>
__global__ void __launch_bounds__(PSY_LIGHT_THREADS) k_shade_bounce(DScene S, PathPool P, Bucket b, ShadeQueues Q,
                                                                    uint32_t* active_out, uint32_t* count_out) {
    constexpr uint32_t KIND = CLS == SHADE_LAMBERT ? (uint32_t)PSY_SHADER_LAMBERT : (uint32_t)PSY_SHADER_GGX;
    const uint32_t n = Q.count[CLS];
    const uint32_t* __restrict__ queue = Q.q[CLS];
    const uint32_t lane = threadIdx.x & 31u;
    for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        const bool on = i < n;
        const uint32_t slot = on ? queue[i] : 0u;
        float time = 0.0f;
        Isect is;
        ShadeCtx shade_ctx;
        if (on) {
            // the finalised hit and the prepared closure, as k_shade_light left them
            const float4 c0 = P.ic[0][slot], c1 = P.ic[1][slot], c2 = P.ic[2][slot], c3 = P.ic[3][slot],
                         c4 = P.ic[4][slot], c5 = P.ic[5][slot];
            time = P.ray_o[slot].w;
            is.pos = mk3(c0.x, c0.y, c0.z), is.pos_err = c0.w;
            is.nor_g = mk3(c1.x, c1.y, c1.z), is.closure.roughness = c1.w;
            is.incoming = mk3(c2.x, c2.y, c2.z), is.closure.fresnel = c2.w;
            shade_ctx.spec = from_f4(c3);
            shade_ctx.nn = mk3(c4.x, c4.y, c4.z);
            shade_ctx.fng = mk3(c5.x, c5.y, c5.z);
            shade_ctx.aa = mk3(0, 0, 0);
            if (KIND == (uint32_t)PSY_SHADER_GGX) shade_ctx.aa = -normalized(is.incoming);
        }
        uint32_t flags = 0, sseed = 0, bounce_count = 0;
        float b4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        bool sample_bounce = false;
        if (on) {
            flags = P.flags[slot];
            sseed = P.sseed[slot] + 1u;  // the light sample's sequence (k_shade_select)
            bounce_count = (flags & F_BOUNCE_MASK) >> F_BOUNCE_SHIFT;
            sample_bounce = bounce_count < 2u;
            if (sample_bounce) {
                const uint32_t q = slot / b.spp, si = slot - q * b.spp;
                const uint32_t px = b.x + (q % b.w), py = b.y + (q / b.w);
                bounce_count += 1u;
                sseed += 1u;
                get_sample_4d(si, 0, px, py, sseed, b4);
            }
        }
        __syncthreads();
        bool alive = false;
        if (on) {
            float closure_sample_pdf = P.closure_pdf[slot];
            const V4 light_attenuation = from_f4(P.light_atten[slot]) / closure_sample_pdf;
            // the shadow ray k_shade_light prepared (the pending contribution is already in P.pending)
            const float4 sh_o = P.sh_o[slot];
            const bool found_light = sh_o.w != 0.0f;
            // bounce (renderer.rs:598-644)
            bool do_bounce = false;
            V3 nb_o = mk3(0, 0, 0), nb_d = mk3(0, 0, 0);
            V4 next_atten = from_f4(P.next_atten[slot]);
            if (sample_bounce) {
                V3 bdir;
                V4 filter;
                float pdf;
                closure_sample<KIND>(is.closure, shade_ctx, is.incoming, b4[0], b4[1], bdir, filter, pdf);
                if ((pdf > 0.0f) && (max_element4(filter) > 0.0f)) {
                    next_atten = filter;
                    closure_sample_pdf = pdf;
                    nb_o = robust_ray_origin(is.pos, is.pos_err, normalized(is.nor_g), bdir);
                    nb_d = bdir;
                    flags |= F_HAS_NEXT;
                    do_bounce = true;
                    P.nb_o[slot] = make_float4(nb_o.x, nb_o.y, nb_o.z, 0.0f);
                    P.nb_d[slot] = make_float4(nb_d.x, nb_d.y, nb_d.z, 0.0f);
                    P.next_atten[slot] = to_f4(next_atten);
                }
            } else {
                flags &= ~F_HAS_NEXT;
            }
            flags = (flags & ~F_BOUNCE_MASK) | (bounce_count << F_BOUNCE_SHIFT);
            P.sseed[slot] = sseed;
            P.closure_pdf[slot] = closure_sample_pdf;
            if (found_light) {
                P.ray_o[slot] = make_float4(sh_o.x, sh_o.y, sh_o.z, time);
                P.ray_d[slot] = P.sh_d[slot];
                flags = (flags & ~F_EVENT_MASK) | (EV_SHADOW << F_EVENT_SHIFT) | F_OCCLUSION;
                P.light_atten[slot] = to_f4(light_attenuation);
                alive = true;
            } else if (do_bounce) {
                P.ray_o[slot] = make_float4(nb_o.x, nb_o.y, nb_o.z, time);
                P.ray_d[slot] = make_float4(nb_d.x, nb_d.y, nb_d.z, CUDART_INF_F);
                flags = (flags & ~(F_EVENT_MASK | F_OCCLUSION)) | (EV_BOUNCE << F_EVENT_SHIFT);
                P.light_atten[slot] = to_f4(light_attenuation * next_atten);
                alive = true;
            } else {
                P.light_atten[slot] = to_f4(light_attenuation);
            }
            P.flags[slot] = flags;
        }
        // survivors: one byte per slot; the next wave's live list is built from them in slot order
        // (dpartition.cuh; replaces the swap compaction of renderer.rs:286-295)
        if (alive) P.alive[slot] = 0;
    }
}

// Second third of a surface hit: sample the light k_shade_select picked and evaluate the closure towards it
// (renderer.rs:488-596).  Its own kernel for the same reason as k_shade_select: with the light sampling inline,
// k_shade<LAMBERT> was 195 KB of straight-line code, each warp streaming through it once, and spent 11 of 17
// cycles per issue waiting for instructions (profiles/r1n_launches.txt); the halves fit the instruction cache.
#ifndef PSY_LIGHT_THREADS
#define PSY_LIGHT_THREADS 512
#endif
// The block is as large as the registers allow (one block per SM) and its warps are kept in step by barriers
// between the big out-of-line pieces, so that at any time all warps of an SM execute the same few KB of code.
template <int CLS>
__global__ void __launch_bounds__(PSY_LIGHT_THREADS) k_shade_light(DScene S, PathPool P, Bucket b, ShadeQueues Q) {
    constexpr uint32_t KIND = CLS == SHADE_LAMBERT ? (uint32_t)PSY_SHADER_LAMBERT : (uint32_t)PSY_SHADER_GGX;
    const uint32_t n = Q.count[CLS];
    const uint32_t* __restrict__ queue = Q.q[CLS];
    for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        const bool on = i < n;
        const uint32_t slot = on ? queue[i] : 0u;
        float time = 0.0f, wl = 0.0f;
        V3 dir = mk3(0, 0, 0);
        Isect is;
        if (on) {
            // the finalised hit, as k_shade_select left it (finalize_hit runs once per surface hit, there)
            const float4 c0 = P.ic[0][slot], c1 = P.ic[1][slot], c2 = P.ic[2][slot], c4 = P.ic[4][slot], c5 = P.ic[5][slot];
            time = P.ray_o[slot].w;
            wl = P.wavelength[slot];
            is.kind = PSY_HIT_TRIANGLE;
            is.t = 0.0f, is.sample_pdf = 0.0f;
            is.pos = mk3(c0.x, c0.y, c0.z), is.pos_err = c0.w;
            is.nor_g = mk3(c1.x, c1.y, c1.z), is.closure.roughness = c1.w;
            is.incoming = mk3(c2.x, c2.y, c2.z), is.closure.fresnel = c2.w;
            is.nor = mk3(c4.x, c4.y, c4.z);
            is.closure.kind = KIND;
            is.closure.color.kind = __float_as_uint(c5.x), is.closure.color.a = c5.y, is.closure.color.b = c5.z;
            is.closure.color.c = c5.w;
            dir = is.incoming;
        }
        ShadeCtx shade_ctx;
        if (on) {
            shade_ctx = closure_prepare<KIND>(S, is.closure, dir, is.nor, is.nor_g, wl);
            P.ic[3][slot] = to_f4(shade_ctx.spec);
            P.ic[4][slot] = make_float4(shade_ctx.nn.x, shade_ctx.nn.y, shade_ctx.nn.z, 0.0f);
            P.ic[5][slot] = make_float4(shade_ctx.fng.x, shade_ctx.fng.y, shade_ctx.fng.z, 0.0f);
        }
        __syncthreads();
        LightSample li;
        li.kind = 0, li.pdf = 0.0f, li.selection_pdf = 0.0f;
        if (on) {
            LightPick pick;
            const uint4 lp = P.light_pick[slot];
            pick.kind = lp.x, pick.index = lp.y, pick.parent_path = lp.z, pick.selection_pdf = __uint_as_float(lp.w);
            const float2 luv = P.light_uv[slot];
            sample_picked_light(S, pick, luv.x, luv.y, wl, time, is.pos, li);
        }
        __syncthreads();
        if (!on) continue;
        const V4 light_attenuation = from_f4(P.light_atten[slot]) / P.closure_pdf[slot];
        float4 sh_o = make_float4(0, 0, 0, 0), sh_d = make_float4(0, 0, 0, 0);
        if (!(li.kind == 0 || li.pdf <= 0.0f || li.selection_pdf <= 0.0f)) {
            V4 attenuation;
            float closure_pdf;
            V3 so, sd;
            float smax;
            V3 ng = normalized(is.nor_g);
            V3 ldir = li.kind == 1 ? li.direction : li.pos - is.pos;
            closure_evaluate<KIND>(is.closure, shade_ctx, ldir, attenuation, closure_pdf);
            so = robust_ray_origin(is.pos, is.pos_err, ng, ldir);
            if (li.kind == 1) {
                sd = li.direction;
                smax = CUDART_INF_F;
            } else {
                V3 offset_end = robust_ray_origin(li.pos, li.err, normalized(li.nor), -ldir);
                sd = offset_end - so;
                smax = 1.0f;
            }
            if (!(max_element4(attenuation) <= 0.0f)) {
                float light_mis_pdf = power_heuristic(li.pdf, closure_pdf);
                V4 pend = ((li.color * attenuation) * light_attenuation) / (light_mis_pdf * li.selection_pdf);
                P.pending[slot] = to_f4(pend);
                sh_o = make_float4(so.x, so.y, so.z, 1.0f);
                sh_d = make_float4(sd.x, sd.y, sd.z, smax);
            }
        }
        P.sh_o[slot] = sh_o;
        P.sh_d[slot] = sh_d;
    }
}

// First half of a surface hit (classes LAMBERT / GGX): the light-tree walk of Scene::sample_lights.  It is
// its own kernel so that this loop-heavy, small piece of code is not evicted from the instruction cache by
// the f64 light-sampling and spectral code of k_shade (k_shade<LAMBERT> was stalled on instruction fetch
// for 12 of 17 cycles per issue when it contained both, profiles/r1_shade_lambert_r1g.txt).
// 8 blocks per SM (64 registers): the walk is a chain of dependent arithmetic per child (normalize, acos / asin,
// divisions) and the kernel ran at 40 % of the issue slots with 14 warps per SM; 32 warps hide it better than the
// few spilled registers cost (shade stage 253 -> 229 ms per C5 frame at 16 spp; 6 / 10 / 12 blocks: 237 / 229 / 233).
template <int CLS>
__global__ void __launch_bounds__(128, 8) k_shade_select(DScene S, PathPool P, Bucket b, ShadeQueues Q) {
    constexpr uint32_t KIND = CLS == SHADE_LAMBERT ? (uint32_t)PSY_SHADER_LAMBERT : (uint32_t)PSY_SHADER_GGX;
    const uint32_t n = Q.count[CLS];
    const uint32_t* __restrict__ queue = Q.q[CLS];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t slot = queue[i];
        const HitRec h = P.hits[slot];
        const float4 ro = P.ray_o[slot], rd = P.ray_d[slot];
        Isect is;
        finalize_hit(S, h, mk3(ro.x, ro.y, ro.z), mk3(rd.x, rd.y, rd.z), ro.w, is);
        const uint32_t q = slot / b.spp, si = slot - q * b.spp;
        const uint32_t px = b.x + (q % b.w), py = b.y + (q / b.w);
        float s4[4];
        get_sample_4d(si, 0, px, py, P.sseed[slot] + 1u, s4);  // next_lds_sequence(); next_lds_samp()
        LightPick pick;
        select_light<KIND>(S, s4[0], ro.w, is, pick);
        P.light_pick[slot] = make_uint4(pick.kind, pick.index, pick.parent_path, __float_as_uint(pick.selection_pdf));
        P.light_uv[slot] = make_float2(s4[1], s4[2]);
        // the finalised hit goes on to k_shade_light (and from there to k_shade_bounce): computed once per surface hit
        P.ic[0][slot] = make_float4(is.pos.x, is.pos.y, is.pos.z, is.pos_err);
        P.ic[1][slot] = make_float4(is.nor_g.x, is.nor_g.y, is.nor_g.z, is.closure.roughness);
        P.ic[2][slot] = make_float4(is.incoming.x, is.incoming.y, is.incoming.z, is.closure.fresnel);
        P.ic[4][slot] = make_float4(is.nor.x, is.nor.y, is.nor.z, 0.0f);  // (unflipped; k_shade_light replaces it)
        P.ic[5][slot] = make_float4(__uint_as_float(is.closure.color.kind), is.closure.color.a, is.closure.color.b,
                                    is.closure.color.c);
    }
}

// One launch per shading class; survivors of all classes are appended to the next wave's live list.
template <int CLS>
__global__ void __launch_bounds__(128) k_shade(DScene S, PathPool P, Bucket b, ShadeQueues Q, uint32_t* active_out,
                                               uint32_t* count_out) {
    const uint32_t n = Q.count[CLS];
    const uint32_t* __restrict__ queue = Q.q[CLS];
    const uint32_t lane = threadIdx.x & 31u;
    for (uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; base < n; base += gridDim.x * blockDim.x) {
        uint32_t i = base + lane;
        bool alive = false;
        uint32_t slot = 0;
        if (i < n) {
            slot = queue[i];
            alive = path_next<CLS>(S, P, slot, b);
        }
        if (alive) P.alive[slot] = 0;  // see k_shade_bounce
    }
}

// Film: sum XYZ::from_spectral_sample(path)/spp over the pixel's samples, in sample order like the reference
// (renderer.rs:299-310), so the sum has the same bits whatever the launch shape.
// A warp owns 32 pixels.  The samples of a pixel are consecutive pool slots, so one thread per pixel would read
// 32 different 16 B pieces per instruction (that version took 338 ms per 4K frame at 256 spp for 42 GB of reads).
// Instead the warp reads 32 consecutive samples of ONE pixel per instruction (512 B, coalesced), each lane converts
// its sample to XYZ/spp and parks it in a shared-memory tile; after 32 pixels the tile is full and lane p adds up
// pixel p's row in order.
// out: image with row stride `stride` pixels whose origin is pixel (ox, oy).
constexpr int FILM_WARPS = 2;
// `ppw` = pixels per warp (<= 32): buckets of many samples per pixel have few pixels, and 32 pixels per warp would
// leave most of the GPU idle (128 x 128 pixels at 256 spp = 512 warps); the host picks it so that the grid fills the SMs.
__global__ void __launch_bounds__(32 * FILM_WARPS) k_film(PathPool P, Bucket b, float* out, uint32_t stride, uint32_t ox,
                                                         uint32_t oy, float* samples_out, uint32_t ppw) {
    __shared__ float tile[FILM_WARPS][3][32][33];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t n_pix = b.w * b.h;
    const uint32_t q0 = (blockIdx.x * FILM_WARPS + warp) * ppw;  // first pixel of this warp's group
    if (q0 >= n_pix) return;
    const float fspp = (float)b.spp;
    V3 acc = mk3(0, 0, 0);
    if (b.spp < 16u) {  // (ppw == 32 here)
        // few samples per pixel: a pixel's samples share one or two sectors anyway, one thread per pixel is fine
        const uint32_t q = q0 + lane;
        if (q < n_pix) {
            for (uint32_t si = 0; si < b.spp; ++si) {
                const uint32_t slot = q * b.spp + si;
                const float4 c = P.color[slot];
                if (samples_out) reinterpret_cast<float4*>(samples_out)[slot] = c;
                const V3 xyz = xyz_from_spectral(from_f4(c), P.wavelength[slot]);
                acc = acc + mk3(xyz.x / fspp, xyz.y / fspp, xyz.z / fspp);
            }
        }
    } else
    for (uint32_t s0 = 0; s0 < b.spp; s0 += 32u) {
        const uint32_t si = s0 + lane;
        for (uint32_t j = 0; j < ppw; ++j) {
            const uint32_t q = q0 + j;
            V3 v = mk3(0, 0, 0);
            if (q < n_pix && si < b.spp) {
                const uint32_t slot = q * b.spp + si;
                const float4 c = P.color[slot];
                if (samples_out) reinterpret_cast<float4*>(samples_out)[slot] = c;
                const V3 xyz = xyz_from_spectral(from_f4(c), P.wavelength[slot]);
                v = mk3(xyz.x / fspp, xyz.y / fspp, xyz.z / fspp);
            }
            tile[warp][0][j][lane] = v.x, tile[warp][1][j][lane] = v.y, tile[warp][2][j][lane] = v.z;
        }
        __syncwarp();
        const uint32_t n_s = min(32u, b.spp - s0);
        if (lane < ppw)
            for (uint32_t s = 0; s < n_s; ++s)
                acc = acc + mk3(tile[warp][0][lane][s], tile[warp][1][lane][s], tile[warp][2][lane][s]);
        __syncwarp();
    }
    const uint32_t q = q0 + lane;
    if (out && q < n_pix && lane < ppw) {
        const uint32_t px = b.x + (q % b.w), py = b.y + (q / b.w);
        float* o = out + ((size_t)(py - oy) * stride + (px - ox)) * 3;
        o[0] = acc.x, o[1] = acc.y, o[2] = acc.z;
    }
}

// ---- strict film order (PsyRenderParams.strict_film_order) ---------------------------------------------------
// The reference adds a bucket's samples to the film in the order its path vector has after the wave loop's swap
// compaction (renderer.rs:286-310): `if paths[i].next(..) { paths.swap(new_end, i); new_end += 1 }` for every wave,
// then `for path in paths { pixel += xyz / spp }`.  Float addition does not commute with that permutation, so an image
// that is to equal the reference's to the last bit has to replay it.  k_record_alive keeps, per slot, which waves the
// path survived; k_film_strict replays the swaps for one bucket (<= 65536 paths, the reference's own limit: its ray
// indices are u16, ray.rs:7) and sums every pixel's samples in the resulting order.
__global__ void __launch_bounds__(256) k_record_alive(const uint8_t* __restrict__ alive, uint32_t n, uint32_t wave,
                                                      uint8_t* __restrict__ whist) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && alive[i] == 0) whist[i] |= (uint8_t)(1u << wave);
}

constexpr uint32_t STRICT_MAX_PATHS = 65536;
constexpr int STRICT_THREADS = 1024;
// order: u16[n] in shared memory (the path now at position i), hist: u8[n]; rank: u32[n] in global memory
__global__ void __launch_bounds__(STRICT_THREADS) k_film_strict(PathPool P, Bucket b, float* out, uint32_t stride,
                                                                uint32_t ox, uint32_t oy, float* samples_out,
                                                                uint32_t* rank, uint32_t n_waves) {
    extern __shared__ unsigned char strict_smem[];
    const uint32_t n = b.w * b.h * b.spp;
    uint16_t* order = reinterpret_cast<uint16_t*>(strict_smem);
    uint8_t* hist = strict_smem + (size_t)n * 2;
    const float fspp = (float)b.spp;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        order[i] = (uint16_t)i;
        hist[i] = P.whist[i];
        const float4 c = P.color[i];
        if (samples_out) reinterpret_cast<float4*>(samples_out)[i] = c;
        const V3 xyz = xyz_from_spectral(from_f4(c), P.wavelength[i]);
        P.pending[i] = make_float4(xyz.x / fspp, xyz.y / fspp, xyz.z / fspp, 0.0f);  // (the pool column is free by now)
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t pi = n;
        for (uint32_t w = 0; w < n_waves && pi > 0; ++w) {
            uint32_t new_end = 0;
            for (uint32_t i = 0; i < pi; ++i) {
                const uint16_t s = order[i];
                if ((hist[s] >> w) & 1u) {
                    order[i] = order[new_end];
                    order[new_end] = s;
                    ++new_end;
                }
            }
            pi = new_end;
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) rank[order[i]] = i;
    __syncthreads();
    // one thread per pixel: its samples in increasing rank
    const uint32_t n_pix = b.w * b.h;
    for (uint32_t q = threadIdx.x; q < n_pix; q += blockDim.x) {
        V3 acc = mk3(0, 0, 0);
        int64_t last = -1;
        for (uint32_t k = 0; k < b.spp; ++k) {
            uint32_t best = 0xFFFFFFFFu, best_slot = 0;
            for (uint32_t si = 0; si < b.spp; ++si) {
                const uint32_t r = rank[q * b.spp + si];
                if ((int64_t)r > last && r < best) best = r, best_slot = q * b.spp + si;
            }
            last = (int64_t)best;
            const float4 v = P.pending[best_slot];
            acc = acc + mk3(v.x, v.y, v.z);
        }
        if (out) {
            const uint32_t px = b.x + (q % b.w), py = b.y + (q / b.w);
            float* o = out + ((size_t)(py - oy) * stride + (px - ox)) * 3;
            o[0] = acc.x, o[1] = acc.y, o[2] = acc.z;
        }
    }
}

}  // namespace psyd

#include "dpartition.cuh"
#include "dpool.cuh"
