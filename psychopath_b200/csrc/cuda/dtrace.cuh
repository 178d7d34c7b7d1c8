// dtrace.cuh — closest-hit / occlusion traversal of one ray through the hierarchical-instance
// scene: assembly BVH4 -> instance (push composed transform, re-localise the ray) -> mesh BVH4 or
// sub-assembly -> pop.  One thread owns one ray; recursion of the reference becomes one explicit
// stack of 32-bit child references plus a small context stack for nested assemblies.
//
// Replaces (file:line in the reference):
//   src/tracer.rs:51-190            TracerInner::{trace, trace_assembly, trace_object}
//   src/accel/bvh4.rs:101-184       BVH4::traverse (ordered traversal via TRAVERSAL_TABLE)
//   src/bbox4.rs:63-97,119-136      BBox4::intersect_ray, time lerp of the bounds
//   src/ray.rs:122-127              RayBatch::update_local
//   src/surface/triangle_mesh.rs:124-262, src/surface/triangle.rs:15-149 (Woop watertight test)
//   src/light/sphere_light.rs:207-280, src/light/rectangle_light.rs:258-300 (hit tests)
#pragma once
#include "dscene.cuh"

namespace psyd {

constexpr int TRACE_STACK_SIZE = 96;
constexpr int TRACE_MAX_NEST = 8;
constexpr uint32_t MARK_POP_MESH = 0xFFFFFFFEu;
constexpr uint32_t MARK_POP_ASM = 0xFFFFFFFDu;

__constant__ uint8_t c_traversal_table[8][48];

struct RayIn {
    V3 orig, dir;
    float time, max_t;
    bool occlusion;
};

struct RayTriPre {
    int xi, yi, zi;
    float sx, sy, sz;
};
// RayTriPrecompute::new (triangle.rs:15-44)
PSY_DEV RayTriPre tri_precompute(V3 d) {
    RayTriPre p;
    float xabs = fabsf(d.x), yabs = fabsf(d.y), zabs = fabsf(d.z);
    if (xabs > yabs && xabs > zabs) {
        p.xi = 1, p.yi = 2, p.zi = 0;
    } else if (yabs > zabs) {
        p.xi = 2, p.yi = 0, p.zi = 1;
    } else {
        p.xi = 0, p.yi = 1, p.zi = 2;
    }
    float dx = comp(d, p.xi), dy = comp(d, p.yi), dz = comp(d, p.zi);
    p.sx = dx / dz;
    p.sy = dy / dz;
    p.sz = 1.0f / dz;
    return p;
}
PSY_DEV float max_abs_3(float a, float b, float c) {
    a = fabsf(a), b = fabsf(b), c = fabsf(c);
    if (a > b && a > c) return a;
    if (b > c) return b;
    return c;
}
// Components of v in the ray's permuted axis order (xi, yi, zi)
PSY_DEV V3 permute3(V3 v, const RayTriPre& pre) { return mk3(comp(v, pre.xi), comp(v, pre.yi), comp(v, pre.zi)); }

// triangle::intersect_ray (triangle.rs:58-149) on inputs that are already in the ray's permuted axis
// order: q_i = permute(tri_i), orig_p = permute(orig).  Per component the arithmetic is the reference's.
PSY_DEV bool tri_intersect_perm(V3 orig_p, const RayTriPre& pre, float max_t, V3 q0, V3 q1, V3 q2, float& t, float& b0,
                                float& b1, float& b2) {
    V3 p0 = q0 - orig_p, p1 = q1 - orig_p, p2 = q2 - orig_p;
    float p0k = p0.z, p1k = p1.z, p2k = p2.z;
    float p0x = p0.x - (pre.sx * p0k);
    float p0y = p0.y - (pre.sy * p0k);
    float p1x = p1.x - (pre.sx * p1k);
    float p1y = p1.y - (pre.sy * p1k);
    float p2x = p2.x - (pre.sx * p2k);
    float p2y = p2.y - (pre.sy * p2k);
    float e0 = (p1x * p2y) - (p1y * p2x);
    float e1 = (p2x * p0y) - (p2y * p0x);
    float e2 = (p0x * p1y) - (p0y * p1x);
    if (e0 == 0.0f || e1 == 0.0f || e2 == 0.0f) {
        e0 = (float)(((double)p1x * (double)p2y) - ((double)p1y * (double)p2x));
        e1 = (float)(((double)p2x * (double)p0y) - ((double)p2y * (double)p0x));
        e2 = (float)(((double)p0x * (double)p1y) - ((double)p0y * (double)p1x));
    }
    if ((e0 < 0.0f || e1 < 0.0f || e2 < 0.0f) && (e0 > 0.0f || e1 > 0.0f || e2 > 0.0f)) return false;
    float det = e0 + e1 + e2;
    if (det == 0.0f) return false;
    float p0z = pre.sz * p0k;
    float p1z = pre.sz * p1k;
    float p2z = pre.sz * p2k;
    float t_scaled = (e0 * p0z) + (e1 * p1z) + (e2 * p2z);
    if ((det > 0.0f && (t_scaled <= 0.0f || t_scaled > (max_t * det))) ||
        (det < 0.0f && (t_scaled >= 0.0f || t_scaled < (max_t * det))))
        return false;
    float inv_det = 1.0f / det;
    b0 = e0 * inv_det;
    b1 = e1 * inv_det;
    b2 = e2 * inv_det;
    t = t_scaled * inv_det;
    {
        float max_zt = max_abs_3(p0z, p1z, p2z);
        float dz = fp_gamma(3) * max_zt;
        float max_xt = max_abs_3(p0x, p1x, p2x);
        float max_yt = max_abs_3(p0y, p1y, p2y);
        float dx = fp_gamma(5) * (max_xt + max_zt);
        float dy = fp_gamma(5) * (max_yt + max_zt);
        float de = 2.0f * ((fp_gamma(2) * max_xt * max_yt) + (dy * max_xt + dx * max_yt));
        float max_e = max_abs_3(e0, e1, e2);
        float dt = 3.0f * ((fp_gamma(3) * max_e * max_zt) + (de * max_zt + dz * max_e)) * fabsf(inv_det);
        if (t <= dt) return false;
    }
    return true;
}
PSY_DEV bool tri_intersect(V3 orig, const RayTriPre& pre, float max_t, V3 t0, V3 t1, V3 t2, float& t, float& b0,
                           float& b1, float& b2) {
    return tri_intersect_perm(permute3(orig, pre), pre, max_t, permute3(t0, pre), permute3(t1, pre), permute3(t2, pre), t,
                              b0, b1, b2);
}

PSY_DEV V3 f4_xyz(float4 v) { return mk3(v.x, v.y, v.z); }

// The three vertices of gathered triangle `ti` of mesh `m` at `time`, in the mesh's local space.
PSY_DEV void load_tri(const DScene& S, const DMesh& m, uint32_t ti, float time, V3& p0, V3& p1, V3& p2) {
    uint32_t k = m.n_keys;
    const float4* tv = S.tri_verts + m.tv_base + (size_t)(ti - m.tri_base) * 3u * k;
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

// world->local transform of path p at `time` (lerp_slice of the composed keys); identity if none
PSY_DEV Xf path_xform(const DScene& S, const DPath& P, float time) {
    if (P.xf_count == 0) return xf_identity();
    return lerp_slice_xf(S.path_xforms + 3 * (size_t)P.xf_base, P.xf_count, time);
}
// local->world ("mat_space") of path p: precomputed inverse for static paths, per-ray inverse of the
// lerped keys under transform motion blur (triangle_mesh.rs:133-138,183-189)
PSY_DEV Xf path_mat_space(const DScene& S, const DPath& P, uint32_t p, float time) {
    if (P.xf_count == 1) return xf_load(S.path_inv + 3 * (size_t)p);
    if (P.xf_count == 0) return xf_identity();
    return xf_inverse(lerp_slice_xf(S.path_xforms + 3 * (size_t)P.xf_base, P.xf_count, time));
}

PSY_DEV int ray_code(V3 d) { return (d.x < 0.0f ? 1 : 0) + (d.y < 0.0f ? 2 : 0) + (d.z < 0.0f ? 4 : 0); }

// Rectangle-light dimensions at `time` (lerp of (w,h) tuples)
PSY_DEV void rect_dims(const float* __restrict__ light_shape, const PsyLight& l, float time, float& w, float& h) {
    const float* s = light_shape + l.shape_base;
    uint32_t i1;
    float a2;
    if (!lerp_slice_keys(l.n_shape, time, i1, a2)) {
        w = __ldg(s + 2 * i1);
        h = __ldg(s + 2 * i1 + 1);
    } else {
        w = lerpf(__ldg(s + 2 * i1), __ldg(s + 2 * i1 + 2), a2);
        h = lerpf(__ldg(s + 2 * i1 + 1), __ldg(s + 2 * i1 + 3), a2);
    }
}

// SphereLight hit test (sphere_light.rs:214-280); returns true and t on a hit
PSY_DEV bool sphere_light_hit(const float* __restrict__ light_shape, const PsyLight& l, const Xf& xform, V3 orig_local,
                              V3 world_dir, float time, float max_t, float& t_out) {
    float radius = lerp_slice_f(light_shape + l.shape_base, l.n_shape, time);
    V3 dir = xf_vector(xform, world_dir);
    float a = length2(dir);
    float b = 2.0f * dot(dir, orig_local);
    float c = length2(orig_local) - (radius * radius);
    float discriminant = (b * b) - (4.0f * a * c);
    if (discriminant < 0.0f) return false;
    discriminant = sqrtf(discriminant);
    float q = b < 0.0f ? -0.5f * (b - discriminant) : -0.5f * (b + discriminant);
    float t0 = q / a;
    float t1 = q != 0.0f ? c / q : max_t;
    if (t0 > t1) {
        float tmp = t0;
        t0 = t1;
        t1 = tmp;
    }
    if (t0 > max_t || t1 <= 0.0f) return false;
    if (t0 > 0.0f) {
        t_out = t0;
    } else if (t1 <= max_t) {
        t_out = t1;
    } else {
        return false;
    }
    return true;
}

// Corners of a rectangle light in world space (rectangle_light.rs:276-285)
PSY_DEV void rect_corners(const Xf& space_inv, float w, float h, V3& p1, V3& p2, V3& p3, V3& p4) {
    p1 = xf_point(space_inv, mk3(w * 0.5f, h * 0.5f, 0.0f));
    p2 = xf_point(space_inv, mk3(w * -0.5f, h * 0.5f, 0.0f));
    p3 = xf_point(space_inv, mk3(w * -0.5f, h * -0.5f, 0.0f));
    p4 = xf_point(space_inv, mk3(w * 0.5f, h * -0.5f, 0.0f));
}

// Hit test of a ray against one light instance (sphere_light.rs:214-280, rectangle_light.rs:287-331),
// out of line: light entries are a small share of the instance visits, and keeping their code (an
// inverse, two watertight triangle tests with f64 fallback, a sqrt) away from the mesh / assembly
// entry keeps the hot loop of k_trace compact in the instruction cache.
struct LightHit {
    uint32_t kind;  // PSY_HIT_MISS when nothing closer than max_t was hit
    uint32_t prim;
    float t, b0, b1, b2;
    uint32_t tri_tests;
};
PSY_DEV LightHit light_instance_hit_inline(const PsyLight* __restrict__ lights, const float* __restrict__ light_shape,
                                           uint32_t light_index, uint32_t path_kind, Xf xform, V3 orig, V3 dir,
                                           V3 orig_in_space, float time, float max_t, RayTriPre pre) {
    LightHit out;
    out.kind = PSY_HIT_MISS, out.prim = 0, out.t = 0.0f, out.b0 = out.b1 = out.b2 = 0.0f, out.tri_tests = 0;
    const PsyLight l = lights[light_index];
    if (path_kind == PATH_SPHERE_LIGHT) {
        float t;
        if (sphere_light_hit(light_shape, l, xform, orig_in_space, dir, time, max_t, t)) {
            out.kind = PSY_HIT_SPHERE_LIGHT;
            out.t = t;
        }
    } else {
        float w, h;
        rect_dims(light_shape, l, time, w, h);
        Xf space_inv = xf_inverse(xform);
        V3 p1, p2, p3, p4;
        rect_corners(space_inv, w, h, p1, p2, p3, p4);
#pragma unroll 1
        for (int k = 0; k < 2; ++k) {
            V3 a = k == 0 ? p1 : p3, b = k == 0 ? p2 : p4, c = k == 0 ? p3 : p1;
            float t, b0, b1, b2;
            out.tri_tests += 1;
            if (tri_intersect(orig, pre, max_t, a, b, c, t, b0, b1, b2)) {
                if (t < max_t) {
                    out.kind = PSY_HIT_RECT_LIGHT;
                    out.t = t;
                    out.prim = (uint32_t)k;
                    out.b0 = b0, out.b1 = b1, out.b2 = b2;
                    break;
                }
            }
        }
    }
    return out;
}
// out-of-line copy for k_trace (one ray per lane), where light entries run with one or two lanes
PSY_NOINLINE LightHit light_instance_hit(const PsyLight* __restrict__ lights, const float* __restrict__ light_shape,
                                         uint32_t light_index, uint32_t path_kind, Xf xform, V3 orig, V3 dir,
                                         V3 orig_in_space, float time, float max_t, RayTriPre pre) {
    return light_instance_hit_inline(lights, light_shape, light_index, path_kind, xform, orig, dir, orig_in_space, time,
                                     max_t, pre);
}

}  // namespace psyd
