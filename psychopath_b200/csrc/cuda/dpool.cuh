// dpool.cuh — k_trace_pool: the trace stage as a warp-private pool of rays in shared memory.
//
// The traversal of one ray is a sequence of steps of three very different kinds: an internal BVH4
// node (bvh4.rs:118-166), one triangle of a leaf (triangle_mesh.rs:174-262), an instance entry
// (tracer.rs:94-149).  With one ray per lane (dtraverse.cuh, k_trace) the lanes of a warp want
// different kinds at the same time, and ncu shows the instance and triangle code running with 4-5
// of 32 lanes (profiles/r1_trace_r1i_phases.txt).  Here a warp owns POOL_RAYS rays whose state
// lives in shared memory (SoA, one column per ray), plus one small queue of ray numbers per step
// kind.  Every iteration the warp picks ONE queue, takes up to 32 rays from it, runs that step for
// all of them together, and files each ray into the queue of its next step.  Lanes are not tied to
// rays, so every step kind runs with (nearly) full warps whatever the mix is.  Everything is
// warp-synchronous: no block barrier, no atomics except the global ray fetch.
//
// Per ray the order of node visits, primitive tests and instance entries is exactly that of the
// reference's stack loop, so hits are bit-identical to k_trace and to the oracle.
//
// A step is a chain of dependent memory accesses, and with 20 warps per SM the kernel lives or dies
// by how many L2 round trips one iteration serialises (first version: 3-4 trips, 30 % issue
// utilisation).  So the state is laid out to make the common steps ONE trip:
//   * the top of the child-reference stack and a 4-entry ring window below it are in shared memory;
//     pushes also write through to the global scratch area, which only deeper pops read back;
//   * the ray is kept in the current assembly's and the current mesh's space at once: leaving a mesh
//     instance is detected by stack height and restores nothing from memory; leaving a first-level
//     assembly is detected the same way and reloads the root-space ray, written once per ray to the
//     scratch area (it used to sit in shared memory: 6 of 33 words, and with them the SM's L1 —
//     trace -5 %); deeper nesting saves / reloads through the scratch area;
//   * the triangle-vertex base of the current mesh is in the state, so a triangle step loads its
//     vertices (and the mesh's local->world matrix from scratch) straight away;
//   * node bounds are fetched together with the child words, before the motion-key test, and through
//     the texture pipe (the LSU pipe carries everything else).
#pragma once
// (included at the end of kernels.cuh: uses PathPool, ShadeQueues, the shading classes and dtrace.cuh)

namespace psyd {

// Rays per warp.  The state is 27 words + 6 queue bytes per ray in shared memory; 84 rays x 4 warps x 5 blocks is the
// most that stays under the 196 KB shared-memory carve-out, which leaves the SM 60 KB of L1 instead of 28 KB (C5 at 16
// spp, trace ms per frame: 64 rays 384, 72 359, 80 346, 84 343, 92 382 — the step up is the carve-out — 96 385).
#ifndef PSY_POOL_RAYS
#define PSY_POOL_RAYS 84
#endif
constexpr int POOL_RAYS = PSY_POOL_RAYS;  // rays per warp (multiple of 4, < 256)
constexpr int POOL_WARPS = 4;             // warps per block
#ifndef PSY_POOL_MIN_BLOCKS
#define PSY_POOL_MIN_BLOCKS 5
#endif
constexpr uint32_t POOL_WIN = 4;  // stack entries below the top that stay in shared memory
#ifndef PSY_POOL_NODE_REPS
#define PSY_POOL_NODE_REPS 1  // internal nodes a ray may take in one step (see the node step)
#endif
#ifndef PSY_POOL_RECYCLE
#define PSY_POOL_RECYCLE 24  // finished rays / free slots that trigger their recycling (trace ms per C5 frame at
                             // 16 spp: 8 -> 377, 16 -> 362, 24 -> 354, 32 -> 378)
#endif

enum PoolQueue { Q_NODE = 0, Q_TRI = 1, Q_INST = 2, Q_DONE = 3, Q_FREE = 4, Q_LIGHT = 5, Q_COUNT = 6 };

// columns of the per-ray state
enum PoolField {
    PF_SLOT = 0,
    PF_MAXT,
    PF_TIME,
    PF_OL,  // 3+3 words: ray origin and 1/direction in the current assembly's space
    PF_DI = PF_OL + 3,
    PF_MOL = PF_DI + 3,  // 3+3 words: the same in the current mesh instance's space
    PF_MDI = PF_MOL + 3,
    PF_FLAGS = PF_MDI + 3,
    PF_TOP,  // next child reference to process (top of the stack)
    PF_SP,   // entries below the top: count | (first one still in the window << 8) | (count at mesh entry << 16)
    PF_PATH_BASE,
    PF_INST_BASE,
    PF_MESH_PATH,
    PF_TVBASE,  // tri_verts index of triangle 0 of the current mesh, minus tri_base * 3 * keys
    PF_NK,      // vertex keys of the current mesh
    PF_STK,  // POOL_WIN words: ring window of the stack
    PF_WORDS = PF_STK + POOL_WIN
};

// bits of PF_FLAGS
constexpr uint32_t RF_OCCL = 1u << 0;
constexpr uint32_t RF_FINITE = 1u << 1;  // with the 3 octant bits above it: "space bits" of the assembly-level ray
constexpr uint32_t RF_SPACE_MASK = 15u << 1;
constexpr uint32_t RF_IN_MESH = 1u << 5;
constexpr uint32_t RF_SPACE_EMPTY = 1u << 6;
constexpr uint32_t RF_CTXN_SHIFT = 7;  // 4 bits
constexpr uint32_t RF_CTXN_MASK = 15u << RF_CTXN_SHIFT;
constexpr uint32_t RF_HIT_SHIFT = 11;  // 3 bits: PSY_HIT_*
constexpr uint32_t RF_HIT_MASK = 7u << RF_HIT_SHIFT;
constexpr uint32_t RF_PERM_SHIFT = 14;    // xi, yi, zi: 2 bits each
constexpr uint32_t RF_MSPACE_SHIFT = 20;  // space bits of the mesh-level ray (finite, octant), 4 bits
constexpr uint32_t RF_MSPACE_MASK = 15u << RF_MSPACE_SHIFT;
constexpr uint32_t RF_MCLS_SHIFT = 24;  // 2 bits: shading class of the current mesh instance
constexpr uint32_t RF_HCLS_SHIFT = 26;  // 2 bits: shading class of the closest triangle hit so far
constexpr uint32_t RF_WSPACE_SHIFT = 28;  // space bits of the root-level ray, 4 bits
// path record word r0.x (dscene.cuh): kind 0-2, own 3, shading class 4-5, transform keys 8+
constexpr uint32_t PREC_CLS_SHIFT = 4;

constexpr int POOL_CTX_WORDS = 9;
constexpr int POOL_MS_WORDS = 12;
constexpr int POOL_RC_WORDS = 6;  // per ray: origin in permuted axis order + shear constants of the watertight test
constexpr int POOL_WOL_WORDS = 6;  // per ray: origin and 1/direction in the root assembly's space (written once per ray;
                                   // what leaving a first-level assembly goes back to)
constexpr size_t POOL_SCRATCH_WORDS_PER_WARP =
    (size_t)(TRACE_STACK_SIZE + POOL_CTX_WORDS * TRACE_MAX_NEST + POOL_MS_WORDS + POOL_RC_WORDS + POOL_WOL_WORDS) * POOL_RAYS;
constexpr size_t POOL_SMEM_BYTES_PER_WARP = (size_t)PF_WORDS * POOL_RAYS * 4 + (size_t)Q_COUNT * POOL_RAYS;
constexpr size_t POOL_SMEM_BYTES = POOL_SMEM_BYTES_PER_WARP * POOL_WARPS;

struct LocalRay {
    V3 o, di;
    uint32_t bits;  // finite | octant << 1
};
// tracer.rs:66-69 / 118-121: the ray in the space of `x` (update_local)
PSY_DEV LocalRay make_local_ray(const Xf& x, V3 orig, V3 dir) {
    LocalRay l;
    l.o = xf_point(x, orig);
    l.di = recip3(xf_vector(x, dir));
    const float big = 3.402823466e+38f;
    const bool finite = fabsf(l.di.x) <= big && fabsf(l.di.y) <= big && fabsf(l.di.z) <= big && fabsf(l.o.x) <= big &&
                        fabsf(l.o.y) <= big && fabsf(l.o.z) <= big;
    l.bits = (finite ? 1u : 0u) | ((uint32_t)ray_code(l.di) << 1);
    return l;
}

// Rare code is kept OUT OF LINE: the kernel is ~48 KB of SASS and the warps of a scheduler are in different phases of
// it, so every cold block inside a hot one costs instruction-cache misses (stall_no_instruction 1.2 cycles per issue;
// moving the two blocks below out: trace -4 %).
// The slab test of one node for rays with an infinite 1/direction component (bbox4.rs:63-97 with SSE's min / max).
PSY_NOINLINE uint32_t slab4_sse(float4 bx0, float4 bx1, float4 by0, float4 by1, float4 bz0, float4 bz1, V3 ol, V3 di, float max_t,
                                uint32_t n_children) {
    const float ADJ = 1.00000024f;
    uint32_t hits = 0;
#define PSY_SLAB_S(LANE, BIT)                                                            \
    {                                                                                    \
        float t1x = (bx0.LANE - ol.x) * di.x;                                            \
        float t1y = (by0.LANE - ol.y) * di.y;                                            \
        float t1z = (bz0.LANE - ol.z) * di.z;                                            \
        float t2x = (bx1.LANE - ol.x) * di.x;                                            \
        float t2y = (by1.LANE - ol.y) * di.y;                                            \
        float t2z = (bz1.LANE - ol.z) * di.z;                                            \
        float fx = sse_max(t1x, t2x), fy = sse_max(t1y, t2y), fz = sse_max(t1z, t2z);    \
        float nx = sse_min(t1x, t2x), ny = sse_min(t1y, t2y), nz = sse_min(t1z, t2z);    \
        float far_t = sse_min(sse_min(fx, sse_min(fy, fz)) * ADJ, max_t);                \
        float near_t = sse_max(sse_max(nx, ny), sse_max(nz, 0.0f));                      \
        if (near_t < far_t) hits |= (1u << BIT);                                         \
    }
    PSY_SLAB_S(x, 0)
    PSY_SLAB_S(y, 1)
    if (n_children > 2u) PSY_SLAB_S(z, 2)
    if (n_children > 3u) PSY_SLAB_S(w, 3)
#undef PSY_SLAB_S
    return hits;
}

template <bool COUNT>
__global__ void __launch_bounds__(32 * POOL_WARPS, PSY_POOL_MIN_BLOCKS)
    k_trace_pool(DScene S, PathPool P, const uint32_t* __restrict__ active, const uint32_t* __restrict__ count,
                 uint32_t* fetch, TraceCounters* counters, ShadeQueues Q, uint32_t* scratch) {
    extern __shared__ uint32_t pool_smem[];
    constexpr int R = POOL_RAYS;
    const uint32_t FULL = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t* const st = pool_smem + warp * (POOL_SMEM_BYTES_PER_WARP / 4);
    uint8_t* const qs = reinterpret_cast<uint8_t*>(st + PF_WORDS * R);
    uint32_t* const stk = scratch + (size_t)(blockIdx.x * POOL_WARPS + warp) * POOL_SCRATCH_WORDS_PER_WARP;
    uint32_t* const ctxb = stk + TRACE_STACK_SIZE * R;            // saved parent spaces of nested assemblies
    uint32_t* const msb = ctxb + POOL_CTX_WORDS * TRACE_MAX_NEST * R;  // local->world of the current mesh
    uint32_t* const rcb = msb + POOL_MS_WORDS * R;                     // per-ray constants of the triangle test
    uint32_t* const wolb = rcb + POOL_RC_WORDS * R;                    // root-space copy of the ray
    const uint32_t n_rays = *count;

    struct {
        uint32_t node_visits, tri_tests, instance_visits, motion_node_visits;  // per lane and launch: far below 2^32
    } cnt;
    cnt.node_visits = cnt.tri_tests = cnt.instance_visits = cnt.motion_node_visits = 0;

    for (uint32_t i = lane; i < (uint32_t)R; i += 32) qs[Q_FREE * R + i] = (uint8_t)i;
    __syncwarp();
    // The queues are LIFO arrays of R bytes; their lengths are packed 8 bits per queue into one warp-uniform
    // 64-bit register, so "the length of queue k" is a shift, not a select chain.
    uint64_t qn = (uint64_t)R << (8 * Q_FREE);
    bool exhausted = n_rays == 0u;
#define QLEN(k) ((uint32_t)(qn >> (8 * (k))) & 255u)

#define ST(F, r) st[(F) * R + (r)]
#define STF(F, r) __uint_as_float(st[(F) * R + (r)])
// stack below the top: `sp` entries, of which [lo, sp) are in the shared-memory ring window
// (a push writes through to the scratch stack: branch-free, and the store is fire-and-forget)
#define PSY_PUSH(v)                                \
    {                                              \
        stk[sp * R + r] = (v);                     \
        ST(PF_STK + (sp % POOL_WIN), r) = (v);     \
        ++sp;                                      \
        lo = sp > lo + POOL_WIN ? sp - POOL_WIN : lo; \
    }
// Leaving the first-level assembly (tracer.rs:131-148): back to the root assembly's ray, which never left
// shared memory.
#define PSY_LEAVE_LEVEL1()                                                                          \
    {                                                                                               \
        flags = (flags & ~(RF_CTXN_MASK | RF_SPACE_MASK)) | (((flags >> RF_WSPACE_SHIFT) & 15u) << 1); \
        ST(PF_PATH_BASE, r) = 0u;                                                                   \
        ST(PF_INST_BASE, r) = S.root_instance_base;                                                 \
        _Pragma("unroll") for (int w_ = 0; w_ < 6; ++w_) ST(PF_OL + w_, r) = wolb[w_ * R + r];      \
    }
// (popping an entry from below the mesh entry's / first-level assembly's stack height leaves it: no marker needed)
#define PSY_POP(out)                                \
    {                                               \
        --sp;                                       \
        if (sp < msp) flags &= ~RF_IN_MESH;         \
        if ((flags & (RF_CTXN_MASK | RF_IN_MESH)) == (1u << RF_CTXN_SHIFT) && sp < asp) PSY_LEAVE_LEVEL1() \
        if (sp < lo) {                              \
            lo = sp;                                \
            out = stk[sp * R + r];                  \
        } else {                                    \
            out = ST(PF_STK + (sp % POOL_WIN), r);  \
        }                                           \
    }
#define PSY_LOAD_SP()                                          \
    {                                                          \
        const uint32_t spw = ST(PF_SP, r);                     \
        sp = spw & 255u, lo = (spw >> 8) & 255u, msp = (spw >> 16) & 255u, asp = spw >> 24; \
    }

    for (;;) {
        // ---- pick the step kind of this iteration (warp-uniform)
        // Finished rays and free slots are recycled as soon as PSY_POOL_RECYCLE of them have gathered; otherwise the
        // fullest of the four work queues runs (ties: triangle > instance > light > node).
        uint32_t k;
        {
            const uint32_t n_done = QLEN(Q_DONE), n_free = QLEN(Q_FREE);
#ifndef PSY_POOL_NODE_BIAS
#define PSY_POOL_NODE_BIAS 0
#endif
            const uint32_t key_node = QLEN(Q_NODE) ? (uint32_t)max((int)min(QLEN(Q_NODE), 32u) + PSY_POOL_NODE_BIAS, 1) * 8u : 0u,
                           key_light = min(QLEN(Q_LIGHT), 32u) * 8u + 1u,
                           key_inst = min(QLEN(Q_INST), 32u) * 8u + 2u, key_tri = min(QLEN(Q_TRI), 32u) * 8u + 3u;
            const uint32_t best = max(max(key_node, key_light), max(key_inst, key_tri));
            if (n_done >= (uint32_t)PSY_POOL_RECYCLE) {
                k = Q_DONE;
            } else if (!exhausted && n_free >= (uint32_t)PSY_POOL_RECYCLE) {
                k = Q_FREE;
            } else if (best >= 8u) {
                k = (0x1250u >> ((best & 7u) * 4u)) & 15u;  // tie-break rank -> queue: 0 node, 1 light, 2 inst, 3 tri
            } else if (n_done > 0u) {
                k = Q_DONE;
            } else if (!exhausted && n_free > 0u) {
                k = Q_FREE;
            } else {
                break;
            }
        }
        const uint32_t n = min(QLEN(k), 32u);
        bool act = lane < n;
        qn -= (uint64_t)n << (8u * k);
        const uint32_t r = act ? (uint32_t)qs[k * R + QLEN(k) + lane] : 0u;
        __syncwarp();  // the queue entries just read may be overwritten by this iteration's pushes

        // ---- finished rays: hit record, shading class, slot back to the free list
        if (k == Q_DONE) {
            int cls = -1;
            uint32_t slot = 0;
            if (act) {
                slot = ST(PF_SLOT, r);
                const uint32_t flags = ST(PF_FLAGS, r);
                const uint32_t hk = (flags & RF_HIT_MASK) >> RF_HIT_SHIFT;
                if (hk == PSY_HIT_MISS || hk == PSY_HIT_OCCLUDED) {
                    uint4* hp = reinterpret_cast<uint4*>(P.hits + slot);
                    hp[0] = make_uint4(hk, 0u, 0u, 0u);
                    hp[1] = make_uint4(0u, 0u, 0u, 0u);
                }
                if ((flags & RF_OCCL) || hk == PSY_HIT_MISS) {
                    cls = SHADE_SIMPLE;
                } else if (hk != PSY_HIT_TRIANGLE) {
                    cls = SHADE_EMIT;
                } else {
                    cls = (int)((flags >> RF_HCLS_SHIFT) & 3u);
                }
                qs[Q_FREE * R + QLEN(Q_FREE) + lane] = (uint8_t)r;
            }
            qn += (uint64_t)n << (8 * Q_FREE);
            // the path's shading class; the class queues are built from these bytes in slot order (dpartition.cuh)
            if (cls >= 0) P.cls[slot] = (uint8_t)cls;
            __syncwarp();
            continue;
        }

        // state every remaining step kind works on
        uint32_t flags = 0, top = PSY_NONE, sp = 0, lo = 0, msp = 0, asp = 0;
        bool to_light = false;  // an instance entry that turned out to be a light: handled by the light step

        if (k == Q_FREE) {
            // ---- new rays: the per-ray prologue of TracerInner::trace (tracer.rs:59-77)
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(fetch, (uint32_t)n);
            base = __shfl_sync(FULL, base, 0);
            if (base + (uint32_t)n >= n_rays) exhausted = true;
            const bool got = act && base + lane < n_rays;
            const uint32_t back = __ballot_sync(FULL, act && !got);
            if (act && !got) qs[Q_FREE * R + QLEN(Q_FREE) + __popc(back & lt)] = (uint8_t)r;
            qn += (uint64_t)__popc(back) << (8 * Q_FREE);
            act = got;
            if (act) {
                const uint32_t slot = active[base + lane];
                const float4 ro = P.ray_o[slot], rd = P.ray_d[slot];
                const V3 orig = mk3(ro.x, ro.y, ro.z), dir = mk3(rd.x, rd.y, rd.z);
                const RayTriPre pre = tri_precompute(dir);
                const V3 op = permute3(orig, pre);
                const LocalRay lr = make_local_ray(xf_identity(), orig, dir);  // evaluated literally, like the reference
                flags = ((P.flags[slot] & F_OCCLUSION) ? RF_OCCL : 0u) | (lr.bits << 1) | (lr.bits << RF_WSPACE_SHIFT) |
                        RF_SPACE_EMPTY | ((uint32_t)PSY_HIT_MISS << RF_HIT_SHIFT) |
                        (((uint32_t)pre.xi | ((uint32_t)pre.yi << 2) | ((uint32_t)pre.zi << 4)) << RF_PERM_SHIFT);
                ST(PF_SLOT, r) = slot;
                ST(PF_MAXT, r) = __float_as_uint(rd.w);
                ST(PF_TIME, r) = __float_as_uint(ro.w);
                ST(PF_OL + 0, r) = __float_as_uint(lr.o.x), ST(PF_OL + 1, r) = __float_as_uint(lr.o.y);
                ST(PF_OL + 2, r) = __float_as_uint(lr.o.z);
                ST(PF_DI + 0, r) = __float_as_uint(lr.di.x), ST(PF_DI + 1, r) = __float_as_uint(lr.di.y);
                ST(PF_DI + 2, r) = __float_as_uint(lr.di.z);
                ST(PF_PATH_BASE, r) = 0u;
                ST(PF_INST_BASE, r) = S.root_instance_base;
#pragma unroll
                for (int w = 0; w < 6; ++w) wolb[w * R + r] = ST(PF_OL + w, r);
                rcb[0 * R + r] = __float_as_uint(op.x), rcb[1 * R + r] = __float_as_uint(op.y);
                rcb[2 * R + r] = __float_as_uint(op.z);
                rcb[3 * R + r] = __float_as_uint(pre.sx), rcb[4 * R + r] = __float_as_uint(pre.sy);
                rcb[5 * R + r] = __float_as_uint(pre.sz);
                top = S.root_bvh2;
            }
        } else if (k == Q_NODE) {
            // ---- one internal node (bvh4.rs:118-166, bbox4.rs:63-97)
            if (act) {
                uint32_t ref = ST(PF_TOP, r);
                flags = ST(PF_FLAGS, r);
                PSY_LOAD_SP()
                const bool in_mesh = (flags & RF_IN_MESH) != 0u;
                const uint32_t fo = in_mesh ? (uint32_t)PF_MOL : (uint32_t)PF_OL;
                const V3 ol = mk3(STF(fo + 0, r), STF(fo + 1, r), STF(fo + 2, r));
                const V3 di = mk3(STF(fo + 3, r), STF(fo + 4, r), STF(fo + 5, r));
                const uint32_t space = in_mesh ? (flags >> RF_MSPACE_SHIFT) & 15u : (flags >> 1) & 15u;  // finite | octant << 1
                const float max_t = STF(PF_MAXT, r);
                // A ray whose nearest hit child is again an internal node takes that node in the same step (up to
                // PSY_POOL_NODE_REPS in a row): its next fetch is issued as soon as the slab test is done, instead of
                // after a trip through the queues.  Lanes whose next entry is a leaf (or nothing) sit the repeats out.
#pragma unroll 1
                for (int rep_i = 0;; ++rep_i) {
                // blob reference (dscene.cuh): unit offset of the node, and how many time keys follow its header
                const uint32_t unit = ref & NODE2_OFF_MASK, kc = ref >> NODE2_KC_SHIFT;
                const uint4* np = S.nodes2 + unit;
                // header and first key at once (for motion nodes the first key is replaced below; its memory is valid)
                uint4 child, meta;  // meta: traversal_code, n_keys, -, n_children
                float4 bx0, bx1, by0, by1, bz0, bz1;
                // motion nodes: which keys the ray's time falls between is known from the reference alone, so the keys
                // are requested together with the header — one round trip
                uint32_t mk_i1 = 0;
                float mk_a2 = 0.0f;
                bool mk_blend = false;
                if (kc != 0u && kc != 7u) mk_blend = lerp_slice_keys(kc + 1u, STF(PF_TIME, r), mk_i1, mk_a2);
                if (S.nodes_tex && kc == 0u) {
                    // node fetches go through the texture pipe, next to the LSU pipe that carries everything else
                    // (a scene too large for a linear texture has no texture object and takes the other branch)
                    const int tn = (int)unit;
                    child = tex1Dfetch<uint4>(S.nodes_tex, tn);
                    meta = tex1Dfetch<uint4>(S.nodes_tex, tn + 1);
#define PSY_TEXF4(i)                                                                                          \
    ([&] {                                                                                                    \
        const uint4 t = tex1Dfetch<uint4>(S.nodes_tex, tn + (i));                                             \
        return make_float4(__uint_as_float(t.x), __uint_as_float(t.y), __uint_as_float(t.z), __uint_as_float(t.w)); \
    }())
                    bx0 = PSY_TEXF4(2), bx1 = PSY_TEXF4(3), by0 = PSY_TEXF4(4), by1 = PSY_TEXF4(5), bz0 = PSY_TEXF4(6),
                    bz1 = PSY_TEXF4(7);
#undef PSY_TEXF4
                } else {
                    child = __ldg(np);
                    meta = __ldg(np + 1);
                    const float4* bp = reinterpret_cast<const float4*>(np + 2) + 6u * mk_i1;
                    bx0 = __ldg(bp), bx1 = __ldg(bp + 1), by0 = __ldg(bp + 2), by1 = __ldg(bp + 3), bz0 = __ldg(bp + 4),
                    bz1 = __ldg(bp + 5);
                    if (mk_blend) {
                        const float ia = 1.0f - mk_a2, a2 = mk_a2;
#define PSY_LERP4(A, B) \
    A = make_float4((A.x * ia) + (B.x * a2), (A.y * ia) + (B.y * a2), (A.z * ia) + (B.z * a2), (A.w * ia) + (B.w * a2))
                        const float4 c0 = __ldg(bp + 6), c1 = __ldg(bp + 7), c2 = __ldg(bp + 8), c3 = __ldg(bp + 9),
                                     c4 = __ldg(bp + 10), c5 = __ldg(bp + 11);
                        PSY_LERP4(bx0, c0);
                        PSY_LERP4(bx1, c1);
                        PSY_LERP4(by0, c2);
                        PSY_LERP4(by1, c3);
                        PSY_LERP4(bz0, c4);
                        PSY_LERP4(bz1, c5);
#undef PSY_LERP4
                    }
                }
                if (kc == 7u) {
                    // 8 or more time keys: the count comes from the header (a dependent fetch, like the ABI layout)
                    uint32_t i1;
                    float a2;
                    const bool blend = lerp_slice_keys(meta.y, STF(PF_TIME, r), i1, a2);
                    const float4* k0 = reinterpret_cast<const float4*>(np + 2) + 6u * i1;
                    bx0 = __ldg(k0), bx1 = __ldg(k0 + 1), by0 = __ldg(k0 + 2), by1 = __ldg(k0 + 3), bz0 = __ldg(k0 + 4),
                    bz1 = __ldg(k0 + 5);
                    if (blend) {
                        const float ia = 1.0f - a2;
#define PSY_LERP4(A, B) \
    A = make_float4((A.x * ia) + (B.x * a2), (A.y * ia) + (B.y * a2), (A.z * ia) + (B.z * a2), (A.w * ia) + (B.w * a2))
                        const float4 c0 = __ldg(k0 + 6), c1 = __ldg(k0 + 7), c2 = __ldg(k0 + 8), c3 = __ldg(k0 + 9),
                                     c4 = __ldg(k0 + 10), c5 = __ldg(k0 + 11);
                        PSY_LERP4(bx0, c0);
                        PSY_LERP4(bx1, c1);
                        PSY_LERP4(by0, c2);
                        PSY_LERP4(by1, c3);
                        PSY_LERP4(bz0, c4);
                        PSY_LERP4(bz1, c5);
#undef PSY_LERP4
                    }
                }
                if (COUNT && kc != 0u) cnt.motion_node_visits += 1;
                if (COUNT) cnt.node_visits += 1;
                const float ADJ = 1.00000024f;
                uint32_t hits = 0;
#define PSY_SLAB(LANE, BIT)                                                  \
    {                                                                        \
        float t1x = (bx0.LANE - ol.x) * di.x;                                \
        float t1y = (by0.LANE - ol.y) * di.y;                                \
        float t1z = (bz0.LANE - ol.z) * di.z;                                \
        float t2x = (bx1.LANE - ol.x) * di.x;                                \
        float t2y = (by1.LANE - ol.y) * di.y;                                \
        float t2z = (bz1.LANE - ol.z) * di.z;                                \
        float fx = SMAX(t1x, t2x), fy = SMAX(t1y, t2y), fz = SMAX(t1z, t2z); \
        float nx = SMIN(t1x, t2x), ny = SMIN(t1y, t2y), nz = SMIN(t1z, t2z); \
        float far_t = SMIN(SMIN(fx, SMIN(fy, fz)) * ADJ, max_t);             \
        float near_t = SMAX(SMAX(nx, ny), SMAX(nz, 0.0f));                   \
        if (near_t < far_t) hits |= (1u << BIT);                             \
    }
                if (space & 1u) {
                    // no slab parameter can be NaN: IEEE min/max gives the bits of SSE's minps/maxps
#define SMIN fminf
#define SMAX fmaxf
                    PSY_SLAB(x, 0)
                    PSY_SLAB(y, 1)
                    if (meta.w > 2u) PSY_SLAB(z, 2)
                    if (meta.w > 3u) PSY_SLAB(w, 3)
#undef SMIN
#undef SMAX
                } else {
                    // axis-parallel rays can produce 0 * inf = NaN: keep the operand order of minps/maxps (rare: out of line)
                    hits = slab4_sse(bx0, bx1, by0, by1, bz0, bz1, ol, di, max_t, meta.w);
                }
#undef PSY_SLAB
                top = PSY_NONE;
                {
                    // far -> near (bvh4.rs:152-165).  Every hit child is written to the stack, predicated and
                    // without nested branches; the nearest one is then taken back as the new top.
                    const uint32_t order_code = c_traversal_table[space >> 1][meta.x];
#pragma unroll
                    for (int i = 3; i >= 0; --i) {
                        const uint32_t child_i = (order_code >> (i * 2)) & 3u;
                        const uint32_t c_lo = (child_i & 1u) ? child.y : child.x;
                        const uint32_t c_hi = (child_i & 1u) ? child.w : child.z;
                        const uint32_t c = (child_i & 2u) ? c_hi : c_lo;
                        if ((uint32_t)i < meta.w && ((hits >> child_i) & 1u)) {
                            stk[sp * R + r] = c;
                            ST(PF_STK + (sp % POOL_WIN), r) = c;
                            ++sp;
                            top = c;
                        }
                    }
                    lo = sp > lo + POOL_WIN ? sp - POOL_WIN : lo;  // (the taken-back entry used a ring slot too)
                    if (hits) --sp;
                }
                if (rep_i + 1 >= PSY_POOL_NODE_REPS || (top & PSY_CHILD_LEAF_BIT)) break;  // (PSY_NONE has the leaf bit)
                ref = top;
                }
            }
        } else if (k == Q_TRI) {
            // ---- the triangles of the leaf on top (triangle_mesh.rs:174-262): all of them (<= 3) in this step, in
            // leaf order and against the running max_t like the reference's inner loop; the ray's constants and the
            // mesh's local->world matrix are fetched once per leaf
            if (act) {
                const uint32_t leaf = ST(PF_TOP, r);
                const uint32_t ti0 = PSY_CHILD_LEAF_START(leaf), count = PSY_CHILD_LEAF_COUNT(leaf);
                const uint32_t nk = ST(PF_NK, r);
                const float4* tv = S.tri_verts + (size_t)(ST(PF_TVBASE, r) + ti0 * 3u * nk);
                flags = ST(PF_FLAGS, r);
                PSY_LOAD_SP()
                top = PSY_NONE;
                RayTriPre pre;
                pre.xi = (int)((flags >> RF_PERM_SHIFT) & 3u);
                pre.yi = (int)((flags >> (RF_PERM_SHIFT + 2)) & 3u);
                pre.zi = (int)((flags >> (RF_PERM_SHIFT + 4)) & 3u);
                const V3 op = mk3(__uint_as_float(rcb[0 * R + r]), __uint_as_float(rcb[1 * R + r]),
                                  __uint_as_float(rcb[2 * R + r]));
                pre.sx = __uint_as_float(rcb[3 * R + r]), pre.sy = __uint_as_float(rcb[4 * R + r]);
                pre.sz = __uint_as_float(rcb[5 * R + r]);
                const bool has_space = !(flags & RF_SPACE_EMPTY);
                Xf ms = xf_identity();
                if (has_space) {
                    const uint32_t* m = msb + r;
                    ms.X = mk3(__uint_as_float(m[0 * R]), __uint_as_float(m[1 * R]), __uint_as_float(m[2 * R]));
                    ms.Y = mk3(__uint_as_float(m[3 * R]), __uint_as_float(m[4 * R]), __uint_as_float(m[5 * R]));
                    ms.Z = mk3(__uint_as_float(m[6 * R]), __uint_as_float(m[7 * R]), __uint_as_float(m[8 * R]));
                    ms.T = mk3(__uint_as_float(m[9 * R]), __uint_as_float(m[10 * R]), __uint_as_float(m[11 * R]));
                }
                uint32_t i1 = 0;
                float a2 = 0.0f;
                const bool blend = nk != 1u && lerp_slice_keys(nk, STF(PF_TIME, r), i1, a2);
                float max_t = STF(PF_MAXT, r);
                bool hit_any = false;
#pragma unroll 1
                for (uint32_t c = 0; c < count; ++c) {
                    const float4* v = tv + (size_t)c * 3u * nk + i1;
                    V3 p0 = f4_xyz(__ldg(v)), p1 = f4_xyz(__ldg(v + nk)), p2 = f4_xyz(__ldg(v + 2u * nk));
                    if (blend) {
                        p0 = lerp3(p0, f4_xyz(__ldg(v + 1)), a2);
                        p1 = lerp3(p1, f4_xyz(__ldg(v + nk + 1)), a2);
                        p2 = lerp3(p2, f4_xyz(__ldg(v + 2u * nk + 1)), a2);
                    }
                    if (has_space) {
                        p0 = xf_point(ms, p0), p1 = xf_point(ms, p1), p2 = xf_point(ms, p2);
                    } else {
                        p0 = permute3(p0, pre), p1 = permute3(p1, pre), p2 = permute3(p2, pre);
                    }
                    if (COUNT) cnt.tri_tests += 1;
                    float t, b0, b1, b2;
                    if (tri_intersect_perm(op, pre, max_t, p0, p1, p2, t, b0, b1, b2)) {
                        if (flags & RF_OCCL) {
                            flags = (flags & ~RF_HIT_MASK) | ((uint32_t)PSY_HIT_OCCLUDED << RF_HIT_SHIFT);
                            sp = 0, lo = 0;
                            break;  // triangle_mesh.rs:247-250: the ray is done
                        }
                        max_t = t;
                        hit_any = true;
                        uint4* hp = reinterpret_cast<uint4*>(P.hits + ST(PF_SLOT, r));
                        hp[0] = make_uint4(PSY_HIT_TRIANGLE, __float_as_uint(t), ti0 + c, ST(PF_MESH_PATH, r));
                        hp[1] = make_uint4(__float_as_uint(b0), __float_as_uint(b1), __float_as_uint(b2), 0u);
                    }
                }
                if (hit_any) {
                    ST(PF_MAXT, r) = __float_as_uint(max_t);
                    flags = (flags & ~(RF_HIT_MASK | (3u << RF_HCLS_SHIFT))) | ((uint32_t)PSY_HIT_TRIANGLE << RF_HIT_SHIFT) |
                            (((flags >> RF_MCLS_SHIFT) & 3u) << RF_HCLS_SHIFT);
                }
            }
        } else {
            // ---- instance entry (tracer.rs:94-149); assembly leaves hold exactly one instance.  Light
            // instances found by the Q_INST step are passed on to the Q_LIGHT step, so that their long,
            // rarely needed hit test runs for many rays at once instead of one or two.
            if (act) {
                const uint32_t start = PSY_CHILD_LEAF_START(ST(PF_TOP, r));
                const uint32_t p = ST(PF_PATH_BASE, r) + (start - ST(PF_INST_BASE, r));
                const uint4 r0 = __ldg(S.path_recs + 2 * (size_t)p);
                const uint4 r1 = __ldg(S.path_recs + 2 * (size_t)p + 1);
                const uint32_t slot = ST(PF_SLOT, r);
                const float4 ro = P.ray_o[slot], rd = P.ray_d[slot];
                flags = ST(PF_FLAGS, r);
                PSY_LOAD_SP()
                top = PSY_NONE;
                const uint32_t kind = r0.x & 7u;
                const uint32_t own = (r0.x >> 3) & 1u;
                const uint32_t xf_count = r0.x >> 8;
                const float time = STF(PF_TIME, r);
                const bool light = kind == PATH_SPHERE_LIGHT || kind == PATH_RECT_LIGHT;
                to_light = k == Q_INST && light;
                if (COUNT && !to_light) cnt.instance_visits += 1;
                // world->local transform of the instance at the ray's time
                Xf xform = xf_identity();
                if (xf_count == 1u) {
                    xform = xf_load(S.path_static + 6 * (size_t)p);
                } else if (xf_count > 1u) {
                    xform = lerp_slice_xf(S.path_xforms + 3 * (size_t)r0.y, xf_count, time);
                }
                const V3 orig = mk3(ro.x, ro.y, ro.z), dir = mk3(rd.x, rd.y, rd.z);
                const uint32_t ctx_n = (flags & RF_CTXN_MASK) >> RF_CTXN_SHIFT;
                const bool mesh = kind == PATH_MESH;
                if (to_light) {
                    top = ST(PF_TOP, r);
                } else if (mesh || (kind == PATH_ASSEMBLY && ctx_n < (uint32_t)TRACE_MAX_NEST)) {
                    // Mesh and assembly entries share everything expensive (the ray in the instance's space),
                    // so the lanes of both kinds run it together.
                    if (!mesh && ctx_n != 0u) {
                        // nested deeper than one level: save the parent space (leaving the assembly is a reload: same
                        // inputs, same bits as the reference's second update_local, tracer.rs:131-148).  The root
                        // level needs no saving: its ray stays in PF_WOL.
                        uint32_t* c = ctxb + (size_t)(ctx_n * POOL_CTX_WORDS) * R + r;
                        c[0 * R] = ST(PF_PATH_BASE, r);
                        c[1 * R] = ST(PF_INST_BASE, r);
                        c[2 * R] = own | (flags & RF_SPACE_MASK);
                        if (own) {
                            c[3 * R] = ST(PF_OL + 0, r), c[4 * R] = ST(PF_OL + 1, r), c[5 * R] = ST(PF_OL + 2, r);
                            c[6 * R] = ST(PF_DI + 0, r), c[7 * R] = ST(PF_DI + 1, r), c[8 * R] = ST(PF_DI + 2, r);
                        }
                    }
                    // a mesh keeps its own copy of the ray (leaving the mesh instance just drops it)
                    const uint32_t fo = mesh ? (uint32_t)PF_MOL : (uint32_t)PF_OL;
                    uint32_t space = (flags >> 1) & 15u;
                    if (own) {
                        const LocalRay lr = make_local_ray(xform, orig, dir);
                        ST(fo + 0, r) = __float_as_uint(lr.o.x), ST(fo + 1, r) = __float_as_uint(lr.o.y);
                        ST(fo + 2, r) = __float_as_uint(lr.o.z);
                        ST(fo + 3, r) = __float_as_uint(lr.di.x), ST(fo + 4, r) = __float_as_uint(lr.di.y);
                        ST(fo + 5, r) = __float_as_uint(lr.di.z);
                        space = lr.bits;
                    } else if (mesh) {
#pragma unroll
                        for (int w = 0; w < 6; ++w) ST(PF_MOL + w, r) = ST(PF_OL + w, r);
                    }
                    if (mesh) {
                        flags = (flags & ~(RF_MSPACE_MASK | RF_SPACE_EMPTY | (3u << RF_MCLS_SHIFT))) | (space << RF_MSPACE_SHIFT) |
                                (xf_count == 0u ? RF_SPACE_EMPTY : 0u) | RF_IN_MESH |
                                (((r0.x >> PREC_CLS_SHIFT) & 3u) << RF_MCLS_SHIFT);
                        ST(PF_MESH_PATH, r) = p;
                        ST(PF_TVBASE, r) = r1.z - r1.y * 3u * r1.w;
                        ST(PF_NK, r) = r1.w;
                        if (xf_count != 0u) {
                            // local->world for the vertices: precomputed for static paths, else the inverse of the
                            // lerped keys, per ray (triangle_mesh.rs:183-189).  Stored with its output components
                            // in the ray's permuted axis order (a rotation: xi = zi+1, yi = zi+2 mod 3), which is
                            // only a choice of store address.
                            const Xf inv = xf_count == 1u ? xf_load(S.path_static + 6 * (size_t)p + 3) : xf_inverse(xform);
                            const uint32_t zi = (flags >> (RF_PERM_SHIFT + 4)) & 3u;
                            const uint32_t o0 = ((5u - zi) % 3u) * R, o1 = ((6u - zi) % 3u) * R, o2 = ((7u - zi) % 3u) * R;
                            uint32_t* m = msb + r;
                            m[o0] = __float_as_uint(inv.X.x), m[o1] = __float_as_uint(inv.X.y), m[o2] = __float_as_uint(inv.X.z);
                            m += 3 * R;
                            m[o0] = __float_as_uint(inv.Y.x), m[o1] = __float_as_uint(inv.Y.y), m[o2] = __float_as_uint(inv.Y.z);
                            m += 3 * R;
                            m[o0] = __float_as_uint(inv.Z.x), m[o1] = __float_as_uint(inv.Z.y), m[o2] = __float_as_uint(inv.Z.z);
                            m += 3 * R;
                            m[o0] = __float_as_uint(inv.T.x), m[o1] = __float_as_uint(inv.T.y), m[o2] = __float_as_uint(inv.T.z);
                        }
                    } else {
                        flags = (flags & ~(RF_SPACE_MASK | RF_CTXN_MASK)) | (space << 1) | ((ctx_n + 1u) << RF_CTXN_SHIFT);
                        ST(PF_PATH_BASE, r) = r0.z;
                        ST(PF_INST_BASE, r) = r1.y;
                    }
                    if (mesh) {
                        msp = sp;  // whatever is popped from below this height is outside the mesh
                    } else if (ctx_n == 0u) {
                        asp = sp;  // ... and outside the first-level assembly
                    } else {
                        PSY_PUSH(MARK_POP_ASM)
                    }
                    top = r0.w;  // BVH root as a blob reference (r1.x is the ABI reference k_trace uses)
                } else if (light) {
                    const V3 ol = own ? xf_point(xform, orig) : mk3(STF(PF_OL + 0, r), STF(PF_OL + 1, r), STF(PF_OL + 2, r));
                    const float max_t = STF(PF_MAXT, r);
                    RayTriPre pre;
                    pre.xi = (int)((flags >> RF_PERM_SHIFT) & 3u);
                    pre.yi = (int)((flags >> (RF_PERM_SHIFT + 2)) & 3u);
                    pre.zi = (int)((flags >> (RF_PERM_SHIFT + 4)) & 3u);
                    pre.sx = __uint_as_float(rcb[3 * R + r]), pre.sy = __uint_as_float(rcb[4 * R + r]);
                    pre.sz = __uint_as_float(rcb[5 * R + r]);
                    const LightHit lh =
                        light_instance_hit(S.lights, S.light_shape, r1.x, kind, xform, orig, dir, ol, time, max_t, pre);
                    if (COUNT) cnt.tri_tests += lh.tri_tests;
                    if (lh.kind != PSY_HIT_MISS) {
                        if (flags & RF_OCCL) {
                            flags = (flags & ~RF_HIT_MASK) | ((uint32_t)PSY_HIT_OCCLUDED << RF_HIT_SHIFT);
                            sp = 0, lo = 0;
                        } else {
                            ST(PF_MAXT, r) = __float_as_uint(lh.t);
                            flags = (flags & ~RF_HIT_MASK) | (lh.kind << RF_HIT_SHIFT);
                            uint4* hp = reinterpret_cast<uint4*>(P.hits + slot);
                            hp[0] = make_uint4(lh.kind, __float_as_uint(lh.t), lh.prim, p);
                            hp[1] = make_uint4(__float_as_uint(lh.b0), __float_as_uint(lh.b1), __float_as_uint(lh.b2), 0u);
                        }
                    }
                }
            }
        }

        // ---- common tail: find the ray's next entry, leave instances on the way (tracer.rs:131-148),
        // and file the ray into the queue of its next step
        int dest = to_light ? (int)Q_LIGHT : -1;
        if (act && !to_light) {
            if (top == PSY_NONE && sp > 0u) PSY_POP(top)
            while (top == MARK_POP_ASM) {
                {
                    const uint32_t level = ((flags & RF_CTXN_MASK) >> RF_CTXN_SHIFT) - 1u;
                    flags = (flags & ~RF_CTXN_MASK) | (level << RF_CTXN_SHIFT);
                    const uint32_t* c = ctxb + (size_t)(level * POOL_CTX_WORDS) * R + r;
                    ST(PF_PATH_BASE, r) = c[0 * R];
                    ST(PF_INST_BASE, r) = c[1 * R];
                    const uint32_t c2 = c[2 * R];
                    if (c2 & 1u) {
                        flags = (flags & ~RF_SPACE_MASK) | (c2 & RF_SPACE_MASK);
                        ST(PF_OL + 0, r) = c[3 * R], ST(PF_OL + 1, r) = c[4 * R], ST(PF_OL + 2, r) = c[5 * R];
                        ST(PF_DI + 0, r) = c[6 * R], ST(PF_DI + 1, r) = c[7 * R], ST(PF_DI + 2, r) = c[8 * R];
                    }
                }
                top = PSY_NONE;
                if (sp > 0u) PSY_POP(top)
            }
            ST(PF_TOP, r) = top;
            ST(PF_SP, r) = sp | (lo << 8) | (msp << 16) | (asp << 24);
            ST(PF_FLAGS, r) = flags;
            dest = top == PSY_NONE ? Q_DONE : (!(top & PSY_CHILD_LEAF_BIT) ? Q_NODE : ((flags & RF_IN_MESH) ? Q_TRI : Q_INST));
        }
        {
            // lanes filing into the same queue find each other with one match; the new lengths are two warp sums
            const uint32_t m_mine = __match_any_sync(FULL, dest);
            // (FIFO queues + prefetch.global.L1/.L2 of the filed ray's next node were tried: no gain, r1 log in
            // profiles/README.md; the queues stay LIFO, which keeps a ray's nodes warm in L1.)
            if (dest >= 0) qs[(uint32_t)dest * R + QLEN((uint32_t)dest) + __popc(m_mine & lt)] = (uint8_t)r;
            const uint32_t add_lo = __reduce_add_sync(FULL, (dest >= 0 && dest < 4) ? 1u << (8 * dest) : 0u);
            const uint32_t add_hi = __reduce_add_sync(FULL, dest == Q_LIGHT ? 1u << (8 * (Q_LIGHT - 4)) : 0u);
            qn += (uint64_t)add_lo | ((uint64_t)add_hi << 32);
        }
        __syncwarp();
    }
#undef QLEN
#undef ST
#undef STF
#undef PSY_PUSH
#undef PSY_POP
#undef PSY_LEAVE_LEVEL1
#undef PSY_LOAD_SP
    if (COUNT) {
        atomicAdd(&counters->node_visits, (unsigned long long)cnt.node_visits);
        atomicAdd(&counters->tri_tests, (unsigned long long)cnt.tri_tests);
        atomicAdd(&counters->instance_visits, (unsigned long long)cnt.instance_visits);
        atomicAdd(&counters->motion_node_visits, (unsigned long long)cnt.motion_node_visits);
    }
}

}  // namespace psyd
