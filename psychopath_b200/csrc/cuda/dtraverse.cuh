// dtraverse.cuh — the resumable per-ray traversal state machine.
//
// `Traversal` holds everything one ray needs between steps (local-space ray, instance context,
// motion-lerped local->world matrix, the stack of child references), so a warp can stop after a
// bounded number of steps, retire finished lanes, hand them new rays and continue ("persistent
// threads with lane refill").  begin() = the per-ray prologue of TracerInner::trace
// (src/tracer.rs:59-77), run_warp() = BVH4::traverse (src/accel/bvh4.rs:101-184) with the instance
// push/pop of trace_assembly (src/tracer.rs:87-150) and the leaf tests of
// TriangleMesh::intersect_rays (src/surface/triangle_mesh.rs:174-262) folded into one loop.
#pragma once
#include "dtrace.cuh"

namespace psyd {

// Words saved per nesting level when the ray enters an instance: the parent's local-space ray, so
// that leaving the instance is a reload instead of the reference's recomputation
// (tracer.rs:131-148 calls update_local again; same inputs, same bits).
constexpr int CTX_WORDS = 10;

struct TraversalMem {
    uint32_t stack[TRACE_STACK_SIZE];
    uint32_t ctx[CTX_WORDS * (TRACE_MAX_NEST + 1)];  // the last slot belongs to the mesh level
};

template <bool COUNT>
struct Traversal {
    // the ray (world space)
    V3 orig, dir;
    float time;
    float max_t;
    bool occlusion;
    RayTriPre pre;
    V3 orig_p;  // origin in the ray's permuted axis order (Woop test)
    // current space
    V3 orig_local, dir_inv_local;
    bool finite_dir_inv;  // no inf/NaN in the local ray: the slab test cannot produce NaN
    uint32_t octant;      // ray_code(dir_inv_local): row of TRAVERSAL_TABLE (bvh4.rs:29-34,111-112)
    // assembly-level context
    uint32_t path_base;      // first path of the current assembly's instances
    uint32_t asm_inst_base;  // first instance of the current assembly
    int ctx_n;
    // mesh-level context
    bool in_mesh, space_empty;
    uint32_t mesh_path;
    uint32_t mesh_tri_base, mesh_tv_base, mesh_n_keys;
    Xf mat_space;  // local->world of the current mesh instance, output components already in the ray's
                   // permuted axis order so the transformed vertices feed the Woop test without selects
    // The dynamically indexed arrays live outside the struct (TraversalMem) so that every scalar member
    // can be promoted to a register.
    uint32_t* stack;
    uint32_t* ctx;
    int sp;
    uint32_t pend;  // parked leaf / instance reference, PSY_NONE when none
    HitRec hit;

    PSY_DEV void attach(TraversalMem& m) {
        stack = m.stack;
        ctx = m.ctx;
        sp = 0;
        pend = PSY_NONE;
    }

    PSY_DEV Xf xf_permuted(const Xf& m) const {
        Xf r;
        r.X = permute3(m.X, pre), r.Y = permute3(m.Y, pre), r.Z = permute3(m.Z, pre), r.T = permute3(m.T, pre);
        return r;
    }
    PSY_DEV void set_local(const Xf& x) {
        orig_local = xf_point(x, orig);
        dir_inv_local = recip3(xf_vector(x, dir));
        const float big = 3.402823466e+38f;
        finite_dir_inv = fabsf(dir_inv_local.x) <= big && fabsf(dir_inv_local.y) <= big && fabsf(dir_inv_local.z) <= big &&
                         fabsf(orig_local.x) <= big && fabsf(orig_local.y) <= big && fabsf(orig_local.z) <= big;
        octant = (uint32_t)ray_code(dir_inv_local);
    }
    PSY_DEV void save_ctx(int level, uint32_t own) {
        uint32_t* c = ctx + level * CTX_WORDS;
        c[0] = path_base;
        c[1] = asm_inst_base;
        c[2] = own | (finite_dir_inv ? 2u : 0u) | (octant << 2);
        c[3] = __float_as_uint(orig_local.x), c[4] = __float_as_uint(orig_local.y), c[5] = __float_as_uint(orig_local.z);
        c[6] = __float_as_uint(dir_inv_local.x), c[7] = __float_as_uint(dir_inv_local.y);
        c[8] = __float_as_uint(dir_inv_local.z);
    }
    PSY_DEV void load_ctx(int level) {
        const uint32_t* c = ctx + level * CTX_WORDS;
        path_base = c[0];
        asm_inst_base = c[1];
        if (c[2] & 1u) {  // the instance we are leaving had its own transform: restore the parent's local ray
            finite_dir_inv = (c[2] & 2u) != 0u;
            orig_local = mk3(__uint_as_float(c[3]), __uint_as_float(c[4]), __uint_as_float(c[5]));
            dir_inv_local = mk3(__uint_as_float(c[6]), __uint_as_float(c[7]), __uint_as_float(c[8]));
            octant = c[2] >> 2;
        }
    }

    PSY_DEV void begin(const DScene& S, const RayIn& ray) {
        orig = ray.orig, dir = ray.dir, time = ray.time, max_t = ray.max_t, occlusion = ray.occlusion;
        hit.kind = PSY_HIT_MISS;
        hit.t = 0.0f;
        hit.prim = 0;
        hit.path = 0;
        hit.b0 = hit.b1 = hit.b2 = 0.0f;
        hit.pad = 0;
        pre = tri_precompute(dir);
        orig_p = permute3(orig, pre);
        set_local(xf_identity());  // tracer.rs:66-69: update_local(i, identity), evaluated literally
        path_base = 0;
        asm_inst_base = S.root_instance_base;
        ctx_n = 0;
        in_mesh = false, space_empty = true;
        mesh_path = 0;
        mesh_tri_base = 0, mesh_tv_base = 0, mesh_n_keys = 1;
        mat_space = xf_identity();
        sp = 0;
        pend = PSY_NONE;
        if (S.root_bvh != PSY_NONE) stack[sp++] = S.root_bvh;
    }

    PSY_DEV bool done() const { return sp == 0 && pend == PSY_NONE; }

    PSY_DEV void load_tri_verts(const DScene& S, uint32_t ti, V3& p0, V3& p1, V3& p2) const {
        uint32_t k = mesh_n_keys;
        const float4* tv = S.tri_verts + mesh_tv_base + (size_t)(ti - mesh_tri_base) * 3u * k;
        if (k == 1) {
            p0 = f4_xyz(__ldg(tv));
            p1 = f4_xyz(__ldg(tv + 1));
            p2 = f4_xyz(__ldg(tv + 2));
        } else {
            uint32_t i1;
            float a2;
            if (!lerp_slice_keys(k, time, i1, a2)) {
                p0 = f4_xyz(__ldg(tv + i1));
                p1 = f4_xyz(__ldg(tv + k + i1));
                p2 = f4_xyz(__ldg(tv + 2 * k + i1));
            } else {
                p0 = lerp3(f4_xyz(__ldg(tv + i1)), f4_xyz(__ldg(tv + i1 + 1)), a2);
                p1 = lerp3(f4_xyz(__ldg(tv + k + i1)), f4_xyz(__ldg(tv + k + i1 + 1)), a2);
                p2 = lerp3(f4_xyz(__ldg(tv + 2 * k + i1)), f4_xyz(__ldg(tv + 2 * k + i1 + 1)), a2);
            }
        }
    }

    // ---- internal node (bvh4.rs:118-166, bbox4.rs:63-97)
    PSY_DEV void visit_node(const DScene& S, uint32_t ref, TraceCounters& cnt) {
        const uint4* np = S.nodes + (size_t)ref * 8u;
        const uint4 child = __ldg(np);
        const uint4 meta = __ldg(np + 1);  // traversal_code, n_keys, key_base, n_children
        float4 bx0, bx1, by0, by1, bz0, bz1;
        if (meta.y == 1u) {
            const float4* bp = reinterpret_cast<const float4*>(np + 2);
            bx0 = __ldg(bp), bx1 = __ldg(bp + 1), by0 = __ldg(bp + 2), by1 = __ldg(bp + 3), bz0 = __ldg(bp + 4),
            bz1 = __ldg(bp + 5);
        } else {
            if (COUNT) cnt.motion_node_visits += 1;
            uint32_t i1;
            float a2;
            bool blend = lerp_slice_keys(meta.y, time, i1, a2);
            const float4* k0 = S.motion_bounds + (size_t)(meta.z + i1) * 6u;
            bx0 = __ldg(k0), bx1 = __ldg(k0 + 1), by0 = __ldg(k0 + 2), by1 = __ldg(k0 + 3), bz0 = __ldg(k0 + 4),
            bz1 = __ldg(k0 + 5);
            if (blend) {
                const float4* k1 = k0 + 6;
                float ia = 1.0f - a2;
#define PSY_LERP4(A, B)                                                                                    \
    A = make_float4((A.x * ia) + (B.x * a2), (A.y * ia) + (B.y * a2), (A.z * ia) + (B.z * a2), (A.w * ia) + (B.w * a2))
                float4 c0 = __ldg(k1), c1 = __ldg(k1 + 1), c2 = __ldg(k1 + 2), c3 = __ldg(k1 + 3), c4 = __ldg(k1 + 4),
                       c5 = __ldg(k1 + 5);
                PSY_LERP4(bx0, c0);
                PSY_LERP4(bx1, c1);
                PSY_LERP4(by0, c2);
                PSY_LERP4(by1, c3);
                PSY_LERP4(bz0, c4);
                PSY_LERP4(bz1, c5);
#undef PSY_LERP4
            }
        }
        if (COUNT) cnt.node_visits += 1;
        const float ADJ = 1.00000024f;
        uint32_t hits = 0;
#define PSY_SLAB(LANE, BIT)                                                                 \
    {                                                                                       \
        float t1x = (bx0.LANE - orig_local.x) * dir_inv_local.x;                            \
        float t1y = (by0.LANE - orig_local.y) * dir_inv_local.y;                            \
        float t1z = (bz0.LANE - orig_local.z) * dir_inv_local.z;                            \
        float t2x = (bx1.LANE - orig_local.x) * dir_inv_local.x;                            \
        float t2y = (by1.LANE - orig_local.y) * dir_inv_local.y;                            \
        float t2z = (bz1.LANE - orig_local.z) * dir_inv_local.z;                            \
        float fx = SMAX(t1x, t2x), fy = SMAX(t1y, t2y), fz = SMAX(t1z, t2z);                \
        float nx = SMIN(t1x, t2x), ny = SMIN(t1y, t2y), nz = SMIN(t1z, t2z);                \
        float far_t = SMIN(SMIN(fx, SMIN(fy, fz)) * ADJ, max_t);                            \
        float near_t = SMAX(SMAX(nx, ny), SMAX(nz, 0.0f));                                  \
        if (near_t < far_t) hits |= (1u << BIT);                                            \
    }
        if (finite_dir_inv) {
            // No slab parameter can be NaN here ((finite - finite) * finite), so the single-instruction
            // IEEE min/max gives the same bits as SSE's minps/maxps.
#define SMIN fminf
#define SMAX fmaxf
            PSY_SLAB(x, 0)
            PSY_SLAB(y, 1)
            if (meta.w > 2u) PSY_SLAB(z, 2)
            if (meta.w > 3u) PSY_SLAB(w, 3)
#undef SMIN
#undef SMAX
        } else {
            // Axis-parallel rays (an infinite 1/dir component) can produce 0 * inf = NaN: keep the
            // operand-order semantics of minps/maxps that the reference's BBox4 test has.
#define SMIN sse_min
#define SMAX sse_max
            PSY_SLAB(x, 0)
            PSY_SLAB(y, 1)
            if (meta.w > 2u) PSY_SLAB(z, 2)
            if (meta.w > 3u) PSY_SLAB(w, 3)
#undef SMIN
#undef SMAX
        }
#undef PSY_SLAB
        if (hits) {
            // push far -> near: order position i = n-1 .. 0, so position 0 is popped first (bvh4.rs:152-165).
            // Fully unrolled and predicated: every lane runs the same four slots.
            const uint32_t order_code = c_traversal_table[octant][meta.x];
#ifdef PSY_PREFETCH_TOP
            uint32_t next_ref = PSY_NONE;
#endif
#pragma unroll
            for (int i = 3; i >= 0; --i) {
                const uint32_t child_i = (order_code >> (i * 2)) & 3u;
                const uint32_t lo = (child_i & 1u) ? child.y : child.x;
                const uint32_t hi = (child_i & 1u) ? child.w : child.z;
                const uint32_t c = (child_i & 2u) ? hi : lo;
                // (prefetch.global.L1 of all pushed children was tried here: 8 % slower on B200, see
                // profiles/README.md.)
                if ((uint32_t)i < meta.w && ((hits >> child_i) & 1u)) {
                    stack[sp++] = c;
#ifdef PSY_PREFETCH_TOP
                    next_ref = c;
#endif
                }
            }
#ifdef PSY_PREFETCH_TOP
            // the entry that will be popped next: start its cache line(s) on the way while the warp runs
            // the leaf phases and the loop bookkeeping
            if (!(next_ref & PSY_CHILD_LEAF_BIT)) {
                const char* a = reinterpret_cast<const char*>(S.nodes + (size_t)next_ref * 8u);
#pragma unroll
                for (int k = 0; k < PSY_PREFETCH_TOP; ++k) asm volatile("prefetch.global.L1 [%0];" ::"l"(a + 128 / PSY_PREFETCH_TOP * k));
            }
#endif
        }
    }

    // ---- one triangle of a leaf (triangle_mesh.rs:174-262); returns true when an occlusion ray is done.
    // A leaf with several triangles stays parked and is consumed one triangle per flush, in order, so
    // that every lane of a flush runs the same straight-line code.
    PSY_DEV bool visit_triangle(const DScene& S, uint32_t ti, TraceCounters& cnt) {
        V3 p0, p1, p2;
        load_tri_verts(S, ti, p0, p1, p2);
        if (!space_empty) {
            p0 = xf_point(mat_space, p0);
            p1 = xf_point(mat_space, p1);
            p2 = xf_point(mat_space, p2);
        } else {
            p0 = permute3(p0, pre), p1 = permute3(p1, pre), p2 = permute3(p2, pre);
        }
        if (COUNT) cnt.tri_tests += 1;
        float t, b0, b1, b2;
        if (tri_intersect_perm(orig_p, pre, max_t, p0, p1, p2, t, b0, b1, b2)) {
            if (occlusion) {
                hit.kind = PSY_HIT_OCCLUDED;
                return true;
            }
            max_t = t;
            hit.kind = PSY_HIT_TRIANGLE;
            hit.t = t;
            hit.prim = ti;
            hit.path = mesh_path;
            hit.b0 = b0, hit.b1 = b1, hit.b2 = b2;
        }
        return false;
    }

    // ---- instance (tracer.rs:94-149); assembly leaves hold exactly one instance.  Everything the entry
    // needs comes from one 32-byte path record (+ the static transform pair at an address that depends
    // only on the path number), i.e. one level of dependent loads.
    PSY_DEV bool visit_instance(const DScene& S, uint32_t start, TraceCounters& cnt) {
        const uint32_t p = path_base + (start - asm_inst_base);
        const uint4 r0 = __ldg(S.path_recs + 2 * (size_t)p);
        const uint4 r1 = __ldg(S.path_recs + 2 * (size_t)p + 1);
        const uint32_t kind = r0.x & 7u;
        const uint32_t own = (r0.x >> 3) & 1u;
        const uint32_t xf_count = r0.x >> 8;
        if (COUNT) cnt.instance_visits += 1;
        // world->local transform of the instance at the ray's time: shared by every kind, so the lanes of a
        // flush run it together whatever they are entering
        Xf xform = xf_identity();
        if (xf_count == 1u) {
            xform = xf_load(S.path_static + 6 * (size_t)p);
        } else if (xf_count > 1u) {
            xform = lerp_slice_xf(S.path_xforms + 3 * (size_t)r0.y, xf_count, time);
        }
        if (kind == PATH_MESH || kind == PATH_ASSEMBLY) {
            const bool mesh = kind == PATH_MESH;
            if (!mesh && ctx_n >= TRACE_MAX_NEST) return false;
            save_ctx(mesh ? TRACE_MAX_NEST : ctx_n, own);
            if (own) set_local(xform);
            if (mesh) {
                mesh_tri_base = r1.y, mesh_tv_base = r1.z, mesh_n_keys = r1.w;
                mesh_path = p;
                space_empty = xf_count == 0u;
                // local->world for the vertices: precomputed for static paths, else the inverse of the
                // lerped keys, per ray in registers (triangle_mesh.rs:183-189)
                if (xf_count == 1u) {
                    mat_space = xf_permuted(xf_load(S.path_static + 6 * (size_t)p + 3));
                } else if (xf_count > 1u) {
                    mat_space = xf_permuted(xf_inverse(xform));
                }
                in_mesh = true;
            } else {
                ++ctx_n;
                path_base = r0.z;
                asm_inst_base = r1.y;
            }
            stack[sp++] = mesh ? MARK_POP_MESH : MARK_POP_ASM;
            stack[sp++] = r1.x;
        } else if (kind == PATH_SPHERE_LIGHT || kind == PATH_RECT_LIGHT) {
            const V3 ol = own ? xf_point(xform, orig) : orig_local;
            const LightHit lh =
                light_instance_hit(S.lights, S.light_shape, r1.x, kind, xform, orig, dir, ol, time, max_t, pre);
            if (COUNT) cnt.tri_tests += lh.tri_tests;
            if (lh.kind != PSY_HIT_MISS) {
                if (occlusion) {
                    hit.kind = PSY_HIT_OCCLUDED;
                    return true;
                }
                max_t = lh.t;
                hit.kind = lh.kind;
                hit.t = lh.t;
                hit.prim = lh.prim;
                hit.path = p;
                hit.b0 = lh.b0, hit.b1 = lh.b1, hit.b2 = lh.b2;
            }
        }
        return false;
    }

    // Marker handling (leaving an object instance / a sub-assembly): tracer.rs:131-148
    PSY_DEV void pop_marker(uint32_t ref) {
        if (ref == MARK_POP_MESH) {
            in_mesh = false;
            load_ctx(TRACE_MAX_NEST);
        } else {
            --ctx_n;
            load_ctx(ctx_n);
        }
    }

    // Warp-synchronous stepping with postponed leaves.  EVERY lane of the warp must call this (idle
    // lanes with no work).  Per iteration every lane pops at most one entry:
    //   * internal nodes and pop-markers are handled at once (the cheap, common case);
    //   * a triangle leaf or an instance entry is parked in `pend` and the lane waits until
    //     `leaf_min` lanes hold a parked entry of that kind, or until no lane has node work left —
    //     then they run the expensive leaf / instance code together.
    // The per-ray order of node visits and primitive tests is exactly that of the plain stack loop, so
    // results are bit-identical; only the interleaving between lanes changes.  (Measured on B200, C4
    // scene: per-lane `while(sp>0)` loops 1.9 active threads/instruction, lock-step steps without
    // parking 6.5-9.2; see profiles/.)
    // Ends when no lane has work, after `budget` iterations, or when `refill_min` lanes are idle.
    PSY_DEV void run_warp(const DScene& S, int budget, int refill_min, int leaf_min, int inst_min, TraceCounters& cnt) {
        const uint32_t FULL = 0xffffffffu;
        for (int it = 0; it < budget; ++it) {
            const bool has_work = sp > 0 || pend != PSY_NONE;
            const uint32_t m = __ballot_sync(FULL, has_work);
            if (m == 0u) break;
            if (32 - __popc(m) >= refill_min) break;
            bool did_node = false;
            if (pend == PSY_NONE && sp > 0) {
                uint32_t ref = stack[--sp];
                if (!(ref & PSY_CHILD_LEAF_BIT)) {
                    visit_node(S, ref, cnt);
                    did_node = true;
                } else if (ref >= MARK_POP_ASM) {
                    pop_marker(ref);
                    did_node = true;
                } else {
                    pend = ref;
                }
            }
            const bool pend_tri = pend != PSY_NONE && in_mesh;
            const bool pend_inst = pend != PSY_NONE && !in_mesh;
            const uint32_t nm = __ballot_sync(FULL, did_node);
            const uint32_t tm = __ballot_sync(FULL, pend_tri);
            const uint32_t im = __ballot_sync(FULL, pend_inst);
            const bool flush_all = nm == 0u;
            if (tm && (flush_all || __popc(tm) >= leaf_min)) {
                if (pend_tri) {
                    const uint32_t start = PSY_CHILD_LEAF_START(pend), count = PSY_CHILD_LEAF_COUNT(pend);
                    const bool finished = count != 0u && visit_triangle(S, start, cnt);
                    pend = (count > 1u && !finished) ? PSY_CHILD_LEAF(start + 1u, count - 1u) : PSY_NONE;
                    if (finished) sp = 0;
                }
            }
            if (im && (flush_all || __popc(im) >= inst_min)) {
                if (pend_inst) {
                    bool finished = visit_instance(S, PSY_CHILD_LEAF_START(pend), cnt);
                    pend = PSY_NONE;
                    if (finished) sp = 0;
                }
            }
        }
    }
};

}  // namespace psyd
