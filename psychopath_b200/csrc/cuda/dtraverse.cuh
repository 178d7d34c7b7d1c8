// dtraverse.cuh — the resumable per-ray traversal state machine.
//
// `Traversal` holds everything one ray needs between steps (local-space ray, instance context,
// motion-lerped local->world matrix, the stack of child references), so a warp can stop after a
// bounded number of steps, retire finished lanes, hand them new rays and continue ("persistent
// threads with lane refill").  begin() = the per-ray prologue of TracerInner::trace
// (src/tracer.rs:59-77), run() = BVH4::traverse (src/accel/bvh4.rs:101-184) with the instance
// push/pop of trace_assembly (src/tracer.rs:87-150) and the leaf tests of
// TriangleMesh::intersect_rays (src/surface/triangle_mesh.rs:174-262) folded into one loop.
#pragma once
#include "dtrace.cuh"

namespace psyd {

struct TraversalMem {
    uint32_t stack[TRACE_STACK_SIZE];
    uint32_t ctx[3 * TRACE_MAX_NEST];
};

template <bool COUNT>
struct Traversal {
    PSY_DEV void attach(TraversalMem& m) {
        stack = m.stack;
        ctx_asm = m.ctx;
        ctx_base = m.ctx + TRACE_MAX_NEST;
        ctx_active = m.ctx + 2 * TRACE_MAX_NEST;
        sp = 0;
        pend = PSY_NONE;
    }
    // the ray (world space)
    V3 orig, dir;
    float time;
    float max_t;
    bool occlusion;
    RayTriPre pre;
    // current space
    V3 orig_local, dir_inv_local;
    bool finite_dir_inv;  // no inf/NaN in the local ray: the slab test cannot produce NaN
    // assembly-level context
    uint32_t cur_asm, path_base, active_path;
    // The dynamically indexed arrays live outside the struct (TraversalMem) so that every scalar member
    // can be promoted to a register.
    uint32_t *ctx_asm, *ctx_base, *ctx_active;
    int ctx_n;
    // mesh-level context
    bool in_mesh, mesh_own_xform, space_empty;
    uint32_t mesh_parent_active, mesh_path;
    uint32_t mesh_tri_base, mesh_tv_base, mesh_n_keys;
    Xf mat_space;
    // stack of child references / markers
    uint32_t* stack;
    int sp;
    uint32_t pend;  // parked leaf / instance reference, PSY_NONE when none
    HitRec hit;

    PSY_DEV void set_local(const Xf& x) {
        orig_local = xf_point(x, orig);
        dir_inv_local = recip3(xf_vector(x, dir));
        const float big = 3.402823466e+38f;
        finite_dir_inv = fabsf(dir_inv_local.x) <= big && fabsf(dir_inv_local.y) <= big && fabsf(dir_inv_local.z) <= big &&
                         fabsf(orig_local.x) <= big && fabsf(orig_local.y) <= big && fabsf(orig_local.z) <= big;
    }
    PSY_DEV void restore_local(const DScene& S) {
        Xf x = active_path == PSY_NONE ? xf_identity() : path_xform(S, S.paths[active_path], time);
        set_local(x);
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
        set_local(xf_identity());  // tracer.rs:66-69: update_local(i, identity), evaluated literally
        cur_asm = S.root_assembly;
        path_base = 0;
        active_path = PSY_NONE;
        ctx_n = 0;
        in_mesh = false, mesh_own_xform = false, space_empty = true;
        mesh_parent_active = PSY_NONE, mesh_path = 0;
        mesh_tri_base = 0, mesh_tv_base = 0, mesh_n_keys = 1;
        mat_space = xf_identity();
        sp = 0;
        pend = PSY_NONE;
        uint32_t root = S.assemblies[cur_asm].root;
        if (root != PSY_NONE) stack[sp++] = root;
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
            uint32_t order_code = c_traversal_table[ray_code(dir_inv_local)][meta.x];
            for (int i = (int)meta.w; i > 0;) {
                --i;
                uint32_t child_i = (order_code >> (i * 2)) & 3u;
                if (hits & (1u << child_i)) {
                    uint32_t c = child_i == 0 ? child.x : (child_i == 1 ? child.y : (child_i == 2 ? child.z : child.w));
                    stack[sp++] = c;
                }
            }
        }
    }

    // ---- triangle leaf (triangle_mesh.rs:174-262); returns true when an occlusion ray is done
    PSY_DEV bool visit_triangles(const DScene& S, uint32_t start, uint32_t count, TraceCounters& cnt) {
        for (uint32_t ti = start; ti < start + count; ++ti) {
            V3 p0, p1, p2;
            load_tri_verts(S, ti, p0, p1, p2);
            if (!space_empty) {
                p0 = xf_point(mat_space, p0);
                p1 = xf_point(mat_space, p1);
                p2 = xf_point(mat_space, p2);
            }
            if (COUNT) cnt.tri_tests += 1;
            float t, b0, b1, b2;
            if (tri_intersect(orig, pre, max_t, p0, p1, p2, t, b0, b1, b2)) {
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
        }
        return false;
    }

    // ---- instance (tracer.rs:94-149); assembly leaves hold exactly one instance
    PSY_DEV bool visit_instance(const DScene& S, uint32_t start, TraceCounters& cnt) {
        const PsyInstance inst = S.instances[start];
        uint32_t p = path_base + (start - S.assemblies[cur_asm].instance_base);
        const DPath P = S.paths[p];
        bool own = inst.xform_count > 0;
        if (COUNT) cnt.instance_visits += 1;
        if (inst.kind == PSY_INSTANCE_OBJECT) {
            const PsyObject obj = S.objects[inst.data_index];
            if (obj.kind == PSY_OBJECT_MESH) {
                const DMesh mesh = S.meshes[obj.index];
                if (mesh.root == PSY_NONE) return false;
                mesh_tri_base = mesh.tri_base, mesh_tv_base = mesh.tv_base, mesh_n_keys = mesh.n_keys;
                mesh_parent_active = active_path;
                mesh_own_xform = own;
                mesh_path = p;
                active_path = p;
                space_empty = P.xf_count == 0;
                if (own) set_local(path_xform(S, P, time));
                if (!space_empty) mat_space = path_mat_space(S, P, p, time);
                in_mesh = true;
                stack[sp++] = MARK_POP_MESH;
                stack[sp++] = mesh.root;
            } else {
                const PsyLight l = S.lights[obj.index];
                Xf xform = path_xform(S, P, time);
                if (l.kind == PSY_LIGHT_SPHERE) {
                    V3 ol = own ? xf_point(xform, orig) : orig_local;
                    float t;
                    if (sphere_light_hit(S, l, xform, ol, dir, time, max_t, t)) {
                        if (occlusion) {
                            hit.kind = PSY_HIT_OCCLUDED;
                            return true;
                        }
                        max_t = t;
                        hit.kind = PSY_HIT_SPHERE_LIGHT;
                        hit.t = t;
                        hit.prim = 0;
                        hit.path = p;
                        hit.b0 = hit.b1 = hit.b2 = 0.0f;
                    }
                } else {
                    float w, h;
                    rect_dims(S, l, time, w, h);
                    Xf space_inv = xf_inverse(xform);
                    V3 p1, p2, p3, p4;
                    rect_corners(space_inv, w, h, p1, p2, p3, p4);
#pragma unroll 1
                    for (int k = 0; k < 2; ++k) {
                        V3 a = k == 0 ? p1 : p3, b = k == 0 ? p2 : p4, c = k == 0 ? p3 : p1;
                        float t, b0, b1, b2;
                        if (COUNT) cnt.tri_tests += 1;
                        if (tri_intersect(orig, pre, max_t, a, b, c, t, b0, b1, b2)) {
                            if (t < max_t) {
                                if (occlusion) {
                                    hit.kind = PSY_HIT_OCCLUDED;
                                    return true;
                                }
                                max_t = t;
                                hit.kind = PSY_HIT_RECT_LIGHT;
                                hit.t = t;
                                hit.prim = (uint32_t)k;
                                hit.path = p;
                                hit.b0 = b0, hit.b1 = b1, hit.b2 = b2;
                                break;
                            }
                        }
                    }
                }
            }
        } else {
            // sub-assembly (tracer.rs:125-127)
            const PsyAssembly sub = S.assemblies[inst.data_index];
            if (sub.root == PSY_NONE || ctx_n >= TRACE_MAX_NEST) return false;
            ctx_asm[ctx_n] = cur_asm;
            ctx_base[ctx_n] = path_base;
            ctx_active[ctx_n] = active_path;
            ++ctx_n;
            cur_asm = inst.data_index;
            path_base = P.child_base;
            active_path = p;
            if (own) set_local(path_xform(S, P, time));
            stack[sp++] = MARK_POP_ASM;
            stack[sp++] = sub.root;
        }
        return false;
    }

    // Marker handling (leaving an object instance / a sub-assembly): tracer.rs:131-148
    PSY_DEV void pop_marker(const DScene& S, uint32_t ref) {
        if (ref == MARK_POP_MESH) {
            in_mesh = false;
            active_path = mesh_parent_active;
            if (mesh_own_xform) restore_local(S);
        } else {
            --ctx_n;
            bool own = S.instances[S.paths[active_path].inst].xform_count > 0;
            cur_asm = ctx_asm[ctx_n];
            path_base = ctx_base[ctx_n];
            active_path = ctx_active[ctx_n];
            if (own) restore_local(S);
        }
    }

    // Warp-synchronous stepping with postponed leaves.  EVERY lane of the warp must call this (idle
    // lanes with no work).  Per iteration every lane pops at most one entry:
    //   * internal nodes and pop-markers are handled at once (the cheap, common case);
    //   * a triangle leaf or an instance entry is parked in `pend` and the lane waits until
    //     `leaf_min` lanes hold a parked entry of that kind, or until no lane has node work left —
    //     then they run the expensive leaf / instance code together.
    // The per-ray order of node visits and primitive tests is exactly that of the plain stack loop, so
    // results are bit-identical; only the interleaving between lanes changes.  (Three measurements on
    // B200, C4 scene: per-lane `while(sp>0)` loops 1.9 active threads/instruction, lock-step steps
    // without parking 6.5-9.2, with parking see profiles/.)
    // Ends when no lane has work, after `budget` iterations, or when `refill_min` lanes are idle.
    PSY_DEV void run_warp(const DScene& S, int budget, int refill_min, int leaf_min, TraceCounters& cnt) {
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
                    pop_marker(S, ref);
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
                    bool finished = visit_triangles(S, PSY_CHILD_LEAF_START(pend), PSY_CHILD_LEAF_COUNT(pend), cnt);
                    pend = PSY_NONE;
                    if (finished) sp = 0;
                }
            }
            if (im && (flush_all || __popc(im) >= leaf_min)) {
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
