// oracle/oracle.cpp — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement ("oracle") of Psychopath's render core operating on the same flat scene
// description (PsySceneDesc) the CUDA library takes.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load this library; the product path never
// does (psychopath_b200/ has no reference to oracle/).
//
// PARITY PINNING: the Rust reference cannot be built in this image (no rustc/cargo, 17 un-vendored
// crates), so the oracle is pinned against (a) every known-answer test the reference's own unit
// tests hold for this path (math3d transform/point/vector/normal tests, lerp tests, fp_utils ulp
// tests, hilbert round trip — tests/test_oracle_pins.py) and (b) the self-derived vectors of
// SURVEY.md §10.  Triangle/slab/BVH/light/closure/spectral code has NO reference test or golden
// image: for those parts parity is UNPINNED beyond the restatement itself.
// sobol_burley and fastapprox are third-party crates absent from /root/reference: unpinned.
//
// Follows (file:line in /root/reference):
//   src/tracer.rs:51-190           trace / trace_assembly / trace_object
//   src/transform_stack.rs:33-82   TransformStack
//   src/accel/bvh4.rs:101-184      BVH4::traverse      src/bbox4.rs:63-97  BBox4::intersect_ray
//   src/surface/triangle_mesh.rs:124-322, src/surface/triangle.rs:15-169
//   src/light/sphere_light.rs:207-338, src/light/rectangle_light.rs:258-337
//   src/scene/mod.rs:28-99, src/scene/assembly.rs:46-140, src/accel/light_tree.rs:135-201
//   src/renderer.rs:194-359,370-737, src/camera.rs:74-102
//
// Deviation (observable only on exact ties): the reference walks a *stream* of rays and orders a
// node's children by the first ray of the stream task (bvh4.rs:111-116); this restatement walks
// one ray at a time and orders children by that ray's own octant.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <thread>
#include <vector>

#include "oshade.h"

namespace orc {

// TRAVERSAL_TABLE (sub_crates/bvh_order/build.rs:17-75)
struct TraversalTable {
    uint8_t t[8][48];
    TraversalTable() {
        for (int raydir = 0; raydir < 8; ++raydir) {
            int ray[3] = {raydir & 1, (raydir >> 1) & 1, (raydir >> 2) & 1};
            int n = 0;
            for (int s2 = 0; s2 < 3; ++s2)
                for (int s1 = 0; s1 < 3; ++s1)
                    for (int s0 = 0; s0 < 3; ++s0) {
                        int perm[4] = {0, 1, 2, 3};
                        if (ray[s1] == 1) std::swap(perm[0], perm[1]);
                        if (ray[s2] == 1) std::swap(perm[2], perm[3]);
                        if (ray[s0] == 1) {
                            std::swap(perm[0], perm[2]);
                            std::swap(perm[1], perm[3]);
                        }
                        t[raydir][n++] = (uint8_t)(perm[0] + (perm[1] << 2) + (perm[2] << 4) + (perm[3] << 6));
                    }
            for (int s1 = 0; s1 < 3; ++s1)
                for (int s0 = 0; s0 < 3; ++s0) {
                    int perm[3] = {0, 1, 2};
                    if (ray[s1] == 1) std::swap(perm[0], perm[1]);
                    if (ray[s0] == 1) {
                        std::swap(perm[0], perm[1]);
                        std::swap(perm[0], perm[2]);
                    }
                    t[raydir][n++] = (uint8_t)(perm[0] + (perm[1] << 2) + (perm[2] << 4));
                }
            for (int s1 = 0; s1 < 3; ++s1)
                for (int s0 = 0; s0 < 3; ++s0) {
                    int perm[3] = {0, 1, 2};
                    if (ray[s1] == 1) std::swap(perm[1], perm[2]);
                    if (ray[s0] == 1) {
                        std::swap(perm[0], perm[2]);
                        std::swap(perm[0], perm[1]);
                    }
                    t[raydir][n++] = (uint8_t)(perm[0] + (perm[1] << 2) + (perm[2] << 4));
                }
            for (int s0 = 0; s0 < 3; ++s0) {
                int perm[2] = {0, 1};
                if (ray[s0] == 1) std::swap(perm[0], perm[1]);
                t[raydir][n++] = (uint8_t)(perm[0] + (perm[1] << 2));
            }
        }
    }
};
static const TraversalTable TRAVERSAL;

struct Counters {
    uint64_t node_visits = 0, tri_tests = 0, instance_visits = 0, motion_node_visits = 0;
};

// Expanded instance paths, numbered breadth-first: the children of path p (an assembly instance)
// are child_base[p] + (instance index within that assembly); the root's children start at 0.
struct Paths {
    std::vector<uint32_t> child_base;  // per path; PSY_NONE for object instances
    uint32_t count = 0;
    bool overflow = false;
};
static Paths expand_paths(const PsySceneDesc& d) {
    Paths p;
    struct Item {
        uint32_t path, assembly;
    };
    std::vector<Item> queue;
    const PsyAssembly& root = d.assemblies[d.root_assembly];
    p.child_base.assign(root.instance_count, PSY_NONE);
    for (uint32_t i = 0; i < root.instance_count; ++i) {
        const PsyInstance& in = d.instances[root.instance_base + i];
        if (in.kind == PSY_INSTANCE_ASSEMBLY) queue.push_back({i, in.data_index});
    }
    for (size_t q = 0; q < queue.size(); ++q) {
        Item it = queue[q];
        const PsyAssembly& a = d.assemblies[it.assembly];
        uint32_t base = (uint32_t)p.child_base.size();
        p.child_base[it.path] = base;
        if (p.child_base.size() + a.instance_count > (1u << 26)) {
            p.overflow = true;
            break;
        }
        p.child_base.resize(base + a.instance_count, PSY_NONE);
        for (uint32_t i = 0; i < a.instance_count; ++i) {
            const PsyInstance& in = d.instances[a.instance_base + i];
            if (in.kind == PSY_INSTANCE_ASSEMBLY) queue.push_back({base + i, in.data_index});
        }
    }
    p.count = (uint32_t)p.child_base.size();
    return p;
}

// TransformStack (transform_stack.rs)
struct TransformStack {
    std::vector<Xf> stack;
    std::vector<size_t> idx{0, 0};
    void clear() {
        stack.clear();
        idx.assign(2, 0);
    }
    void push(const PsyXform* xf, size_t n) {
        if (stack.empty()) {
            for (size_t i = 0; i < n; ++i) stack.push_back(from_abi(xf[i]));
        } else {
            size_t i1 = idx[idx.size() - 2], i2 = idx[idx.size() - 1];
            size_t n1 = i2 - i1;
            size_t maxlen = std::max(n, n1);
            std::vector<Xf> out(maxlen);
            // merge_slices_to(parent, child, |p, c| p * c)   (algorithm.rs:266-305)
            if (n1 == n) {
                for (size_t i = 0; i < n; ++i) out[i] = xf_mul(stack[i1 + i], from_abi(xf[i]));
            } else if (n1 > n) {
                float s = (float)(n1 - 1);
                for (size_t i = 0; i < n1; ++i) {
                    Xf c = lerp_slice_fn<Xf>(n, (float)i / s, [&](size_t k) { return from_abi(xf[k]); });
                    out[i] = xf_mul(stack[i1 + i], c);
                }
            } else {
                float s = (float)(n - 1);
                for (size_t i = 0; i < n; ++i) {
                    Xf p = lerp_slice(stack.data() + i1, n1, (float)i / s);
                    out[i] = xf_mul(p, from_abi(xf[i]));
                }
            }
            stack.insert(stack.end(), out.begin(), out.end());
        }
        idx.push_back(stack.size());
    }
    void pop() {
        size_t i1 = idx[idx.size() - 2], i2 = idx[idx.size() - 1];
        stack.resize(stack.size() - (i2 - i1));
        idx.pop_back();
    }
    const Xf* top() const { return stack.data() + idx[idx.size() - 2]; }
    size_t top_len() const { return idx[idx.size() - 1] - idx[idx.size() - 2]; }
};

struct Ray {
    V3 orig, dir;
    float time, max_t;
    bool occlusion, done;
    V3 orig_local, dir_inv_local;
    void update_local(const Xf& x) {
        orig_local = xf_point(x, orig);
        dir_inv_local = recip(xf_vector(x, dir));
    }
};

enum IsectKind { ISECT_MISS = 0, ISECT_OCCLUDE = 1, ISECT_HIT = 2 };
struct Isect {
    int kind = ISECT_MISS;
    V3 incoming{}, pos{}, nor{}, nor_g{};
    float pos_err = 0, t = 0, sample_pdf = 0;
    Closure closure{};
    // ids
    uint32_t hit_kind = PSY_HIT_MISS, prim_id = 0, inst_id = 0, path_id = 0;
    float b0 = 0, b1 = 0, b2 = 0;
};

inline int ray_code(V3 d) { return (d.x < 0.0f ? 1 : 0) + (d.y < 0.0f ? 2 : 0) + (d.z < 0.0f ? 4 : 0); }

// RayTriPrecompute + intersect_ray (triangle.rs:15-149)
struct RayTriPre {
    int xi, yi, zi;
    float sx, sy, sz;
};
inline RayTriPre tri_precompute(V3 d) {
    RayTriPre p;
    float xabs = std::fabs(d.x), yabs = std::fabs(d.y), zabs = std::fabs(d.z);
    if (xabs > yabs && xabs > zabs) {
        p.xi = 1, p.yi = 2, p.zi = 0;
    } else if (yabs > zabs) {
        p.xi = 2, p.yi = 0, p.zi = 1;
    } else {
        p.xi = 0, p.yi = 1, p.zi = 2;
    }
    float dx = d.get(p.xi), dy = d.get(p.yi), dz = d.get(p.zi);
    p.sx = dx / dz;
    p.sy = dy / dz;
    p.sz = 1.0f / dz;
    return p;
}
inline float max_abs_3(float a, float b, float c) {
    a = std::fabs(a), b = std::fabs(b), c = std::fabs(c);
    if (a > b && a > c) return a;
    if (b > c) return b;
    return c;
}
inline bool tri_intersect(V3 orig, const RayTriPre& pre, float max_t, V3 t0, V3 t1, V3 t2, float& t, float& b0, float& b1,
                          float& b2) {
    V3 p0 = t0 - orig, p1 = t1 - orig, p2 = t2 - orig;
    float p0x = p0.get(pre.xi) - (pre.sx * p0.get(pre.zi));
    float p0y = p0.get(pre.yi) - (pre.sy * p0.get(pre.zi));
    float p1x = p1.get(pre.xi) - (pre.sx * p1.get(pre.zi));
    float p1y = p1.get(pre.yi) - (pre.sy * p1.get(pre.zi));
    float p2x = p2.get(pre.xi) - (pre.sx * p2.get(pre.zi));
    float p2y = p2.get(pre.yi) - (pre.sy * p2.get(pre.zi));
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
    float p0z = pre.sz * p0.get(pre.zi);
    float p1z = pre.sz * p1.get(pre.zi);
    float p2z = pre.sz * p2.get(pre.zi);
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
        float dt = 3.0f * ((fp_gamma(3) * max_e * max_zt) + (de * max_zt + dz * max_e)) * std::fabs(inv_det);
        if (t <= dt) return false;
    }
    return true;
}
// surface_point (triangle.rs:155-169)
inline void surface_point(V3 t0, V3 t1, V3 t2, float b0, float b1, float b2, V3& pos, float& pos_err) {
    pos = ((t0 * b0) + (t1 * b1)) + (t2 * b2);
    V3 e = (((vabs(t0) * b0) + (vabs(t1) * b1)) + (vabs(t2) * b2)) * fp_gamma(7);
    pos_err = max_element(e);
}

struct Tracer {
    const PsySceneDesc& d;
    const Paths& paths;
    TransformStack xstack;
    Counters cnt;

    Tracer(const PsySceneDesc& desc, const Paths& p) : d(desc), paths(p) {}

    // lerp_slice(space, time); the reference indexes out of bounds (panics) for a light instance
    // with no transform anywhere above it — identity is used here instead.
    Xf top_lerp(float time) const {
        return xstack.top_len() > 0 ? lerp_slice(xstack.top(), xstack.top_len(), time) : Xf();
    }

    V3 vert(uint32_t key_index) const {
        const float* v = d.verts + (size_t)key_index * 3;
        return mk(v[0], v[1], v[2]);
    }
    V3 normal_at(uint32_t key_index) const {
        const float* v = d.normals + (size_t)key_index * 3;
        return mk(v[0], v[1], v[2]);
    }

    // BBox4::intersect_ray for the first n lanes (bbox4.rs:63-97); returns a lane mask
    static int bbox4_hits(const PsyBBox4& b, const Ray& r, int n) {
        const float ADJ = 1.00000024f;
        int mask = 0;
        for (int i = 0; i < n; ++i) {
            float t1x = (b.v[0][i] - r.orig_local.x) * r.dir_inv_local.x;
            float t1y = (b.v[2][i] - r.orig_local.y) * r.dir_inv_local.y;
            float t1z = (b.v[4][i] - r.orig_local.z) * r.dir_inv_local.z;
            float t2x = (b.v[1][i] - r.orig_local.x) * r.dir_inv_local.x;
            float t2y = (b.v[3][i] - r.orig_local.y) * r.dir_inv_local.y;
            float t2z = (b.v[5][i] - r.orig_local.z) * r.dir_inv_local.z;
            float fx = sse_max(t1x, t2x), fy = sse_max(t1y, t2y), fz = sse_max(t1z, t2z);
            float nx = sse_min(t1x, t2x), ny = sse_min(t1y, t2y), nz = sse_min(t1z, t2z);
            float far_t = sse_min(sse_min(fx, sse_min(fy, fz)) * ADJ, r.max_t);
            float near_t = sse_max(sse_max(nx, ny), sse_max(nz, 0.0f));
            if (near_t < far_t) mask |= 1 << i;
        }
        return mask;
    }
    static PsyBBox4 lerp_bbox4(const PsyBBox4& a, const PsyBBox4& b, float alpha) {
        PsyBBox4 o;
        for (int r = 0; r < 6; ++r)
            for (int i = 0; i < 4; ++i) o.v[r][i] = (a.v[r][i] * (1.0f - alpha)) + (b.v[r][i] * alpha);
        return o;
    }

    // BVH4::traverse for one ray (bvh4.rs:101-184)
    template <typename LeafFn>
    void traverse(uint32_t root, Ray& ray, LeafFn leaf_fn) {
        if (root == PSY_NONE) return;
        const uint8_t* table = TRAVERSAL.t[ray_code(ray.dir_inv_local)];
        uint32_t stack[(42 * 3) + 2];
        int sp = 0;
        stack[sp++] = root;
        while (sp > 0) {
            uint32_t ref = stack[--sp];
            if (PSY_CHILD_IS_LEAF(ref)) {
                leaf_fn(PSY_CHILD_LEAF_START(ref), PSY_CHILD_LEAF_COUNT(ref));
                continue;
            }
            const PsyBvh4Node& node = d.nodes[ref];
            cnt.node_visits += 1;
            int hits = 0;
            if (!ray.done) {
                if (node.n_keys == 1) {
                    hits = bbox4_hits(node.bounds, ray, (int)node.n_children);
                } else {
                    cnt.motion_node_visits += 1;
                    const PsyBBox4* keys = d.motion_bounds + node.key_base;
                    PsyBBox4 b;
                    size_t n = node.n_keys;
                    if (ray.time == 1.0f) {
                        b = keys[n - 1];
                    } else {
                        float tmp = ray.time * (float)(n - 1);
                        size_t i1 = (size_t)tmp;
                        float alpha2 = tmp - (float)i1;
                        b = lerp_bbox4(keys[i1], keys[i1 + 1], alpha2);
                    }
                    hits = bbox4_hits(b, ray, (int)node.n_children);
                }
            }
            if (hits) {
                uint8_t order_code = table[node.traversal_code];
                for (int i = (int)node.n_children; i > 0;) {
                    --i;
                    int child_i = (order_code >> (i * 2)) & 3;
                    if (hits & (1 << child_i)) stack[sp++] = node.child[child_i];
                }
            }
        }
    }

    Isect trace(Ray& ray) {
        Isect isect;
        xstack.clear();
        ray.done = false;
        ray.update_local(Xf());
        trace_assembly(d.root_assembly, 0, ray, isect);
        return isect;
    }

    void trace_assembly(uint32_t asm_index, uint32_t path_base, Ray& ray, Isect& isect) {
        const PsyAssembly& a = d.assemblies[asm_index];
        traverse(a.root, ray, [&](uint32_t start, uint32_t) {
            const PsyInstance& inst = d.instances[start];
            uint32_t path = path_base + (start - a.instance_base);
            cnt.instance_visits += 1;
            bool has_x = inst.xform_count > 0;
            if (has_x) {
                xstack.push(d.xforms + inst.xform_base, inst.xform_count);
                ray.update_local(lerp_slice(xstack.top(), xstack.top_len(), ray.time));
            }
            if (inst.kind == PSY_INSTANCE_OBJECT) {
                const PsyObject& o = d.objects[inst.data_index];
                if (o.kind == PSY_OBJECT_MESH)
                    intersect_mesh(d.meshes[o.index], inst, path, ray, isect);
                else if (d.lights[o.index].kind == PSY_LIGHT_SPHERE)
                    intersect_sphere_light(d.lights[o.index], inst, path, ray, isect);
                else
                    intersect_rect_light(d.lights[o.index], inst, path, ray, isect);
            } else {
                trace_assembly(inst.data_index, paths.child_base[path], ray, isect);
            }
            if (has_x) {
                xstack.pop();
                if (xstack.top_len() > 0)
                    ray.update_local(lerp_slice(xstack.top(), xstack.top_len(), ray.time));
                else
                    ray.update_local(Xf());
            }
        });
    }

    Closure shader_closure(uint32_t shader_index) const {
        Closure c{};
        if (shader_index == PSY_NONE) {
            // unassigned_shader: Emit rec709_to_xyz((1,0,1)) (tracer.rs:161-164; D65 matrix)
            c.kind = PSY_SHADER_EMIT;
            c.color.kind = PSY_COLOR_XYZ;
            c.color.a = (1.0f * 0.4123907993f) + (0.0f * 0.3575843394f) + (1.0f * 0.1804807884f);
            c.color.b = (1.0f * 0.2126390059f) + (0.0f * 0.7151686788f) + (1.0f * 0.0721923154f);
            c.color.c = (1.0f * 0.0193308187f) + (0.0f * 0.1191947798f) + (1.0f * 0.9505321522f);
            return c;
        }
        const PsyShader& s = d.shaders[shader_index];
        c.kind = s.kind;
        c.color = s.color;
        c.roughness = s.roughness;
        c.fresnel = s.fresnel;
        return c;
    }

    // TriangleMesh::intersect_rays for one ray (triangle_mesh.rs:124-322)
    void intersect_mesh(const PsyMesh& m, const PsyInstance& inst, uint32_t path, Ray& ray, Isect& isect) {
        const Xf* space = xstack.top();
        size_t space_len = xstack.top_len();
        Xf static_mat_space = space_len == 1 ? xf_inverse(lerp_slice(space, space_len, 0.0f)) : Xf();
        traverse(m.root, ray, [&](uint32_t start, uint32_t count) {
            if (ray.done) return;
            Xf mat_space = space_len > 1 ? xf_inverse(lerp_slice(space, space_len, ray.time)) : static_mat_space;
            bool non_shadow_hit = false;
            V3 h0{}, h1{}, h2{};
            const uint32_t* hit_idx = nullptr;
            float ht = 0, hb0 = 0, hb1 = 0, hb2 = 0;
            RayTriPre pre = tri_precompute(ray.dir);
            for (uint32_t ti = start; ti < start + count; ++ti) {
                const uint32_t* idx = d.tris + (size_t)ti * 4;
                V3 p0, p1, p2;
                if (m.n_keys == 1) {
                    p0 = vert(m.vert_base + idx[0]);
                    p1 = vert(m.vert_base + idx[1]);
                    p2 = vert(m.vert_base + idx[2]);
                } else {
                    uint32_t k = m.n_keys;
                    p0 = lerp_slice_fn<V3>(k, ray.time, [&](size_t j) { return vert(m.vert_base + idx[0] * k + (uint32_t)j); });
                    p1 = lerp_slice_fn<V3>(k, ray.time, [&](size_t j) { return vert(m.vert_base + idx[1] * k + (uint32_t)j); });
                    p2 = lerp_slice_fn<V3>(k, ray.time, [&](size_t j) { return vert(m.vert_base + idx[2] * k + (uint32_t)j); });
                }
                if (space_len > 0) {
                    p0 = xf_point(mat_space, p0);
                    p1 = xf_point(mat_space, p1);
                    p2 = xf_point(mat_space, p2);
                }
                cnt.tri_tests += 1;
                float t, b0, b1, b2;
                if (tri_intersect(ray.orig, pre, ray.max_t, p0, p1, p2, t, b0, b1, b2)) {
                    if (ray.occlusion) {
                        isect = Isect();
                        isect.kind = ISECT_OCCLUDE;
                        isect.hit_kind = PSY_HIT_OCCLUDED;
                        ray.done = true;
                        break;
                    } else {
                        non_shadow_hit = true;
                        ray.max_t = t;
                        h0 = p0, h1 = p1, h2 = p2;
                        hit_idx = idx;
                        ht = t, hb0 = b0, hb1 = b1, hb2 = b2;
                    }
                }
            }
            if (non_shadow_hit) {
                V3 pos;
                float pos_err;
                surface_point(h0, h1, h2, hb0, hb1, hb2, pos, pos_err);
                V3 geo_normal = cross(h0 - h1, h0 - h2);
                V3 shading_normal = geo_normal;
                if (m.has_normals) {
                    uint32_t k = m.n_keys;
                    V3 n0 = normalized(lerp_slice_fn<V3>(k, ray.time, [&](size_t j) { return normal_at(m.vert_base + hit_idx[0] * k + (uint32_t)j); }));
                    V3 n1 = normalized(lerp_slice_fn<V3>(k, ray.time, [&](size_t j) { return normal_at(m.vert_base + hit_idx[1] * k + (uint32_t)j); }));
                    V3 n2 = normalized(lerp_slice_fn<V3>(k, ray.time, [&](size_t j) { return normal_at(m.vert_base + hit_idx[2] * k + (uint32_t)j); }));
                    V3 s_nor = xf_normal(mat_space, ((n0 * hb0) + (n1 * hb1)) + (n2 * hb2));
                    shading_normal = dot(s_nor, geo_normal) >= 0.0f ? s_nor : -s_nor;
                }
                isect.kind = ISECT_HIT;
                isect.incoming = ray.dir;
                isect.t = ht;
                isect.pos = pos;
                isect.pos_err = pos_err;
                isect.nor = shading_normal;
                isect.nor_g = geo_normal;
                isect.sample_pdf = 0.0f;
                isect.closure = shader_closure(inst.shader_index);
                isect.hit_kind = PSY_HIT_TRIANGLE;
                isect.prim_id = hit_idx[3];
                isect.inst_id = inst.id;
                isect.path_id = path;
                isect.b0 = hb0, isect.b1 = hb1, isect.b2 = hb2;
            }
        });
    }

    PsyColor light_color(const PsyLight& l, float time) const {
        return lerp_slice(d.light_colors + l.color_base, l.n_colors, time);
    }

    // SphereLight::sample_pdf (sphere_light.rs:51-81)
    float sphere_sample_pdf(const PsyLight& l, const Xf& space, V3 arr, float time) const {
        V3 a = xf_point(space, arr);
        double radius = (double)lerp_slice(d.light_shape + l.shape_base, l.n_shape, time);
        double d2 = (double)length2(mk(0, 0, 0) - a);
        double dd = std::sqrt(d2);
        if (dd > radius) {
            double sin_theta_max2 = rmin((radius * radius) / d2, 1.0);
            double cos_theta_max = std::sqrt(1.0 - sin_theta_max2);
            return (float)uniform_sample_cone_pdf(cos_theta_max);
        }
        return (float)(1.0 / (4.0 * PI_64));
    }

    // SphereLight::intersect_rays (sphere_light.rs:207-338)
    void intersect_sphere_light(const PsyLight& l, const PsyInstance& inst, uint32_t path, Ray& ray, Isect& isect) {
        float time = ray.time;
        Xf xform = top_lerp(time);
        float radius = lerp_slice(d.light_shape + l.shape_base, l.n_shape, time);
        V3 orig = ray.orig_local;
        V3 dir = xf_vector(xform, ray.dir);
        float a = length2(dir);
        float b = 2.0f * dot(dir, orig);
        float c = length2(orig) - (radius * radius);
        float discriminant = (b * b) - (4.0f * a * c);
        if (discriminant < 0.0f) return;
        discriminant = std::sqrt(discriminant);
        float q = b < 0.0f ? -0.5f * (b - discriminant) : -0.5f * (b + discriminant);
        float t0 = q / a;
        float t1 = q != 0.0f ? c / q : ray.max_t;
        if (t0 > t1) std::swap(t0, t1);
        if (t0 > ray.max_t || t1 <= 0.0f) return;
        float t;
        if (t0 > 0.0f)
            t = t0;
        else if (t1 <= ray.max_t)
            t = t1;
        else
            return;
        if (ray.occlusion) {
            isect = Isect();
            isect.kind = ISECT_OCCLUDE;
            isect.hit_kind = PSY_HIT_OCCLUDED;
            ray.done = true;
        } else {
            Xf inv_xform = xf_inverse(xform);
            V3 t_pos = orig + (dir * t);
            V3 unit_pos = normalized(t_pos);
            V3 pos = xf_vector(inv_xform, unit_pos * radius);  // Vector * Transform, then into_point
            V3 normal = xf_normal(inv_xform, unit_pos);
            isect.kind = ISECT_HIT;
            isect.incoming = ray.dir;
            isect.t = t;
            isect.pos = pos;
            isect.pos_err = 0.001f;
            isect.nor = normal;
            isect.nor_g = normal;
            isect.sample_pdf = sphere_sample_pdf(l, xform, ray.orig, time);
            float inv_surface_area = (float)(1.0 / (4.0 * PI_64 * (double)radius * (double)radius));
            isect.closure = Closure{PSY_SHADER_EMIT, color_scale(light_color(l, time), inv_surface_area), 0, 0};
            isect.hit_kind = PSY_HIT_SPHERE_LIGHT;
            isect.prim_id = 0;
            isect.inst_id = inst.id;
            isect.path_id = path;
            isect.b0 = isect.b1 = isect.b2 = 0.0f;
            ray.max_t = t;
        }
    }

    void rect_dims(const PsyLight& l, float time, float& w, float& h) const {
        const float* s = d.light_shape + l.shape_base;
        size_t n = l.n_shape;
        if (n == 1 || time == 1.0f) {
            w = s[2 * (n - 1)];
            h = s[2 * (n - 1) + 1];
            return;
        }
        float tmp = time * (float)(n - 1);
        size_t i1 = (size_t)tmp;
        float alpha2 = tmp - (float)i1;
        w = lerp(s[2 * i1], s[2 * i1 + 2], alpha2);
        h = lerp(s[2 * i1 + 1], s[2 * i1 + 3], alpha2);
    }

    // RectangleLight::sample_pdf (rectangle_light.rs:52-96)
    float rect_sample_pdf(const PsyLight& l, const Xf& space, V3 arr, V3 sample_dir, V3 hit_point, float time) const {
        float w, h;
        rect_dims(l, time, w, h);
        Xf space_inv = xf_inverse(space);
        V3 p1 = xf_point(space_inv, mk(w * 0.5f, h * 0.5f, 0.0f));
        V3 p2 = xf_point(space_inv, mk(w * -0.5f, h * 0.5f, 0.0f));
        V3 p3 = xf_point(space_inv, mk(w * -0.5f, h * -0.5f, 0.0f));
        V3 p4 = xf_point(space_inv, mk(w * 0.5f, h * -0.5f, 0.0f));
        V3 sp1 = normalized(p1 - arr), sp2 = normalized(p2 - arr), sp3 = normalized(p3 - arr), sp4 = normalized(p4 - arr);
        float area_1 = spherical_triangle_solid_angle(sp2, sp1, sp3);
        float area_2 = spherical_triangle_solid_angle(sp4, sp1, sp3);
        V3 normal = xf_normal(space_inv, mk(0, 0, 1));
        if ((area_1 + area_2) < 0.01f) {
            float area = triangle_surface_area(p2, p1, p3) + triangle_surface_area(p4, p1, p3);
            return length2(hit_point - arr) / std::fabs(dot(normalized(sample_dir), normalized(normal))) / area;
        }
        return 1.0f / (area_1 + area_2);
    }

    // RectangleLight::intersect_rays (rectangle_light.rs:258-337)
    void intersect_rect_light(const PsyLight& l, const PsyInstance& inst, uint32_t path, Ray& ray, Isect& isect) {
        float time = ray.time;
        V3 orig = ray.orig, dir = ray.dir;
        float max_t = ray.max_t;
        float w, h;
        rect_dims(l, time, w, h);
        Xf xform = top_lerp(time);
        Xf space_inv = xf_inverse(xform);
        V3 p1 = xf_point(space_inv, mk(w * 0.5f, h * 0.5f, 0.0f));
        V3 p2 = xf_point(space_inv, mk(w * -0.5f, h * 0.5f, 0.0f));
        V3 p3 = xf_point(space_inv, mk(w * -0.5f, h * -0.5f, 0.0f));
        V3 p4 = xf_point(space_inv, mk(w * 0.5f, h * -0.5f, 0.0f));
        RayTriPre pre = tri_precompute(dir);
        V3 tris[2][3] = {{p1, p2, p3}, {p3, p4, p1}};
        for (int k = 0; k < 2; ++k) {
            float t, b0, b1, b2;
            cnt.tri_tests += 1;
            if (tri_intersect(orig, pre, max_t, tris[k][0], tris[k][1], tris[k][2], t, b0, b1, b2)) {
                if (t < max_t) {
                    if (ray.occlusion) {
                        isect = Isect();
                        isect.kind = ISECT_OCCLUDE;
                        isect.hit_kind = PSY_HIT_OCCLUDED;
                        ray.done = true;
                    } else {
                        V3 pos;
                        float pos_err;
                        surface_point(tris[k][0], tris[k][1], tris[k][2], b0, b1, b2, pos, pos_err);
                        V3 normal = cross(tris[k][0] - tris[k][1], tris[k][0] - tris[k][2]);
                        isect.kind = ISECT_HIT;
                        isect.incoming = dir;
                        isect.t = t;
                        isect.pos = pos;
                        isect.pos_err = pos_err;
                        isect.nor = normal;
                        isect.nor_g = normal;
                        isect.sample_pdf = rect_sample_pdf(l, xform, orig, dir, pos, time);
                        float inv_surface_area = (float)(1.0 / ((double)w * (double)h));
                        isect.closure = Closure{PSY_SHADER_EMIT, color_scale(light_color(l, time), inv_surface_area), 0, 0};
                        isect.hit_kind = PSY_HIT_RECT_LIGHT;
                        isect.prim_id = (uint32_t)k;
                        isect.inst_id = inst.id;
                        isect.path_id = path;
                        isect.b0 = b0, isect.b1 = b1, isect.b2 = b2;
                        ray.max_t = t;
                    }
                    break;
                }
            }
        }
    }
};

//------------------------------------------------------------------------------------------------
// Light sampling

static bool g_dbg_on = false;
struct LightSample {
    int kind = 0;  // 0 none, 1 distant, 2 surface
    V4 color{};
    V3 direction{};          // distant
    V3 pos{}, nor{};         // surface: sample_geo
    float err = 0;
    float pdf = 0, selection_pdf = 0;
};

struct LightSampler {
    const PsySceneDesc& d;
    TransformStack xstack;
    explicit LightSampler(const PsySceneDesc& desc) : d(desc) {}

    static V3 b3(const float* p) { return mk(p[0], p[1], p[2]); }

    // LightTree::select (light_tree.rs:135-201)
    bool select(const PsyAssembly& a, V3 inc, V3 pos, V3 nor, V3 nor_g, const Closure& sc, float time, float n,
                uint32_t& light_i, float& sel_pdf, float& whittled_n) const {
        if (a.light_root == PSY_NONE) return false;
        auto node_prob = [&](const PsyLightNode& nd) {
            V3 bmin, bmax;
            const PsyBBox* bb = d.lt_bounds + nd.bounds_base;
            size_t nb = nd.bounds_count;
            if (nb == 1 || time == 1.0f) {
                bmin = b3(bb[nb - 1].min);
                bmax = b3(bb[nb - 1].max);
            } else {
                float tmp = time * (float)(nb - 1);
                size_t i1 = (size_t)tmp;
                float alpha2 = tmp - (float)i1;
                bmin = lerp(b3(bb[i1].min), b3(bb[i1 + 1].min), alpha2);
                bmax = lerp(b3(bb[i1].max), b3(bb[i1 + 1].max), alpha2);
            }
            V3 center = lerp(bmin, bmax, 0.5f);
            V3 dd = center - pos;
            float r2 = length2(bmax - bmin) * 0.25f;
            float inv_surface_area = 1.0f / r2;
            float approx_contrib = closure_estimate_sphere(sc, inc, dd, r2, nor, nor_g);
            if (ORC_UNLIKELY(g_dbg_on))
                std::fprintf(stderr, "[oracle dbg]     child energy %.9g r2 %.9g dd (%.9g %.9g %.9g) contrib %.9g\n", (double)nd.energy,
                             (double)r2, (double)dd.x, (double)dd.y, (double)dd.z, (double)approx_contrib);
            return nd.energy * inv_surface_area * approx_contrib;
        };
        const PsyLightNode* node = d.light_nodes + a.light_root;
        float tot_prob = 1.0f;
        while (node->child_count > 0) {
            float ps[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            float total = 0.0f;
            uint32_t nc = node->child_count;
            for (uint32_t i = 0; i < nc; ++i) {
                float p = node_prob(d.light_nodes[node->child_base + i]);
                ps[i] = p;
                total += p;
            }
            if (total <= 0.0f) {
                float p = 1.0f / (float)nc;
                for (int i = 0; i < 8; ++i) ps[i] = p;
            } else {
                for (int i = 0; i < 8; ++i) ps[i] /= total;
            }
            float base = 0.0f;
            for (uint32_t i = 0; i < 8; ++i) {
                float p = ps[i];
                if ((n <= base + p) || (i == nc - 1)) {
                    tot_prob *= p;
                    node = d.light_nodes + node->child_base + i;
                    n = (n - base) / p;
                    break;
                } else {
                    base += p;
                }
            }
        }
        light_i = node->light_index;
        sel_pdf = tot_prob;
        whittled_n = n;
        return true;
    }

    // SphereLight::sample_from_point (sphere_light.rs:85-192)
    void sample_sphere(const PsyLight& l, const Xf& space, V3 arr_w, float u, float v, float wl, float time, V4& color,
                       V3& spos, V3& snor, float& serr, float& pdf) const {
        V3 arr = xf_point(space, arr_w);
        V3 pos = mk(0, 0, 0);
        Xf inv_space = xf_inverse(space);
        double radius = (double)lerp_slice(d.light_shape + l.shape_base, l.n_shape, time);
        PsyColor col = lerp_slice(d.light_colors + l.color_base, l.n_colors, time);
        double surface_area_inv = 1.0 / (4.0 * PI_64 * radius * radius);
        V3 z = pos - arr;
        double d2 = (double)length2(z);
        double dd = std::sqrt(d2);
        V3 x, y;
        coordinate_system_from_vector(z, x, y);
        x = normalized(x), y = normalized(y), z = normalized(z);
        {
            V3 vr = mk((float)radius, (float)radius, (float)radius);
            V3 v2 = xf_vector(inv_space, vr);
            serr = length(v2) * 0.001f;
        }
        if (dd > radius) {
            double sin_theta_max2 = rmin((radius * radius) / d2, 1.0);
            double cos_theta_max2 = 1.0 - sin_theta_max2;
            double sin_theta_max = std::sqrt(sin_theta_max2);
            double cos_theta_max = std::sqrt(cos_theta_max2);
            V3 sample = normalized(uniform_sample_cone(u, v, cos_theta_max));
            double cos_theta = (double)sample.z;
            double cos_theta2 = cos_theta * cos_theta;
            double sin_theta2 = rmax(1.0 - cos_theta2, 0.0);
            double sin_theta = std::sqrt(sin_theta2);
            double ddd = 1.0 - (d2 * sin_theta * sin_theta / (radius * radius));
            double cos_a = ddd <= 0.0 ? sin_theta_max : ((dd / radius) * sin_theta2) + (cos_theta * std::sqrt(ddd));
            double sin_a = std::sqrt(rmax(1.0 - (cos_a * cos_a), 0.0));
            double phi = (double)v * 2.0 * PI_64;
            V3 s2 = mk((float)(m_cos(phi) * sin_a * radius), (float)(m_sin(phi) * sin_a * radius),
                       (float)(dd - (cos_a * radius)));
            V3 sample_vec = ((x * s2.x) + (y * s2.y)) + (z * s2.z);
            V3 normal = normalized(arr + sample_vec);
            V3 point = normal * (float)radius;
            spos = xf_point(inv_space, point);
            snor = xf_normal(inv_space, normal);
            pdf = (float)uniform_sample_cone_pdf(cos_theta_max);
            color = color_to_spectral(col, wl) * (float)surface_area_inv;
        } else {
            V3 sample_vec = uniform_sample_sphere(u, v);
            V3 normal = normalized(arr + sample_vec);
            V3 point = normal * (float)radius;
            spos = xf_point(inv_space, point);
            snor = xf_normal(inv_space, normal);
            pdf = (float)(1.0 / (4.0 * PI_64));
            color = color_to_spectral(col, wl) * (float)surface_area_inv;
        }
    }

    // RectangleLight::sample_from_point (rectangle_light.rs:121-243)
    void sample_rect(const PsyLight& l, const Xf& space, V3 arr, float u, float v, float wl, float time, V4& color,
                     V3& spos, V3& snor, float& serr, float& pdf) const {
        float w, h;
        {
            const float* s = d.light_shape + l.shape_base;
            size_t n = l.n_shape;
            if (n == 1 || time == 1.0f) {
                w = s[2 * (n - 1)];
                h = s[2 * (n - 1) + 1];
            } else {
                float tmp = time * (float)(n - 1);
                size_t i1 = (size_t)tmp;
                float alpha2 = tmp - (float)i1;
                w = lerp(s[2 * i1], s[2 * i1 + 2], alpha2);
                h = lerp(s[2 * i1 + 1], s[2 * i1 + 3], alpha2);
            }
        }
        PsyColor col = lerp_slice(d.light_colors + l.color_base, l.n_colors, time);
        double surface_area = (double)w * (double)h;
        double surface_area_inv = 1.0 / surface_area;
        Xf space_inv = xf_inverse(space);
        V3 p1 = xf_point(space_inv, mk(w * 0.5f, h * 0.5f, 0.0f));
        V3 p2 = xf_point(space_inv, mk(w * -0.5f, h * 0.5f, 0.0f));
        V3 p3 = xf_point(space_inv, mk(w * -0.5f, h * -0.5f, 0.0f));
        V3 p4 = xf_point(space_inv, mk(w * 0.5f, h * -0.5f, 0.0f));
        V3 lp1 = p1 - arr, lp2 = p2 - arr, lp3 = p3 - arr, lp4 = p4 - arr;
        V3 sp1 = normalized(lp1), sp2 = normalized(lp2), sp3 = normalized(lp3), sp4 = normalized(lp4);
        float area_1 = spherical_triangle_solid_angle(sp2, sp1, sp3);
        float area_2 = spherical_triangle_solid_angle(sp4, sp1, sp3);
        V3 normal = xf_normal(space_inv, mk(0, 0, 1));
        snor = normal;
        serr = 0.0001f;
        if ((area_1 + area_2) < 0.01f) {
            float sa1 = triangle_surface_area(p2, p1, p3);
            float sa2 = triangle_surface_area(p4, p1, p3);
            float threshhold = sa1 / (sa1 + sa2);
            V3 sample_point;
            if (u < threshhold)
                sample_point = uniform_sample_triangle(p2, p1, p3, v, u / threshhold);
            else
                sample_point = uniform_sample_triangle(p4, p1, p3, v, (u - threshhold) / (1.0f - threshhold));
            V3 shadow_vec = sample_point - arr;
            color = color_to_spectral(col, wl) * (float)surface_area_inv * 0.5f;
            pdf = length2(sample_point - arr) / std::fabs(dot(normalized(shadow_vec), normalized(normal))) / (sa1 + sa2);
            spos = sample_point;
        } else {
            float prob_1 = std::isinf(area_1) ? 1.0f : (std::isinf(area_2) ? 0.0f : area_1 / (area_1 + area_2));
            float prob_2 = 1.0f - prob_1;
            V3 shadow_vec = u < prob_1 ? uniform_sample_spherical_triangle(sp2, sp1, sp3, v, u / prob_1)
                                       : uniform_sample_spherical_triangle(sp4, sp1, sp3, v, 1.0f - ((u - prob_1) / prob_2));
            V3 arr_local = xf_point(space, arr);
            V3 svl = xf_vector(space, shadow_vec);
            svl = svl * (-arr_local.z / svl.z);
            V3 spl = arr_local + svl;
            spl.x = rmin(rmax(spl.x, w * -0.5f), w * 0.5f);
            spl.y = rmin(rmax(spl.y, h * -0.5f), h * 0.5f);
            spl.z = 0.0f;
            spos = xf_point(space_inv, spl);
            pdf = 1.0f / (area_1 + area_2);
            color = color_to_spectral(col, wl) * (float)surface_area_inv * 0.5f;
        }
    }

    // Assembly::sample_lights (assembly.rs:46-140)
    bool sample_assembly(uint32_t asm_index, float n, float u, float v, float wl, float time, const Isect& intr,
                         LightSample& out) {
        const PsyAssembly& a = d.assemblies[asm_index];
        if (intr.kind != ISECT_HIT) return false;
        Xf sel_xform = xstack.top_len() > 0 ? lerp_slice(xstack.top(), xstack.top_len(), time) : Xf();
        uint32_t light_i;
        float sel_pdf, whittled_n;
        if (!select(a, xf_vector(sel_xform, intr.incoming), xf_point(sel_xform, intr.pos), xf_normal(sel_xform, intr.nor),
                    xf_normal(sel_xform, intr.nor_g), intr.closure, time, n, light_i, sel_pdf, whittled_n))
            return false;
        const PsyInstance& inst = d.light_instances[a.light_instance_base + light_i];
        if (inst.kind == PSY_INSTANCE_OBJECT) {
            const PsyObject& o = d.objects[inst.data_index];
            const PsyLight& l = d.lights[o.index];
            Xf xform;
            if (inst.xform_count > 0) {
                Xf own = lerp_slice_fn<Xf>(inst.xform_count, time, [&](size_t k) { return from_abi(d.xforms[inst.xform_base + k]); });
                xform = xstack.top_len() > 0 ? xf_mul(lerp_slice(xstack.top(), xstack.top_len(), time), own) : own;
            } else {
                xform = xstack.top_len() > 0 ? lerp_slice(xstack.top(), xstack.top_len(), time) : Xf();
            }
            out.kind = 2;
            if (l.kind == PSY_LIGHT_SPHERE)
                sample_sphere(l, xform, intr.pos, u, v, wl, time, out.color, out.pos, out.nor, out.err, out.pdf);
            else
                sample_rect(l, xform, intr.pos, u, v, wl, time, out.color, out.pos, out.nor, out.err, out.pdf);
            out.selection_pdf = sel_pdf;
            return true;
        }
        if (inst.xform_count > 0) xstack.push(d.xforms + inst.xform_base, inst.xform_count);
        bool ok = sample_assembly(inst.data_index, whittled_n, u, v, wl, time, intr, out);
        if (inst.xform_count > 0) xstack.pop();
        if (ok) out.selection_pdf = out.selection_pdf * sel_pdf;
        return ok;
    }

    float distant_energy(const PsyDistantLight& l) const {
        float a = 0.0f;
        for (uint32_t i = 0; i < l.n_colors; ++i) {
            const PsyColor& c = d.light_colors[l.color_base + i];
            float e = c.kind == PSY_COLOR_XYZ ? c.b : (c.kind == PSY_COLOR_BLACKBODY ? (c.a * c.a) * (c.a * c.a) * c.b : c.b);
            a = a + e;
        }
        return a / (float)l.n_colors;
    }

    // Scene::sample_lights (scene/mod.rs:28-99)
    LightSample sample_scene(float n, float u, float v, float w, float wl, float time, const Isect& intr) {
        (void)w;
        LightSample out;
        float wsum = 0.0f;
        for (uint32_t i = 0; i < d.n_distant_lights; ++i) wsum = wsum + distant_energy(d.distant_lights[i]);
        float wl_energy = wsum <= 0.0f ? 0.0f : 1.0f;
        float ll_energy = d.assemblies[d.root_assembly].light_energy <= 0.0f ? 0.0f : 1.0f;
        float tot_energy = wl_energy + ll_energy;
        if (tot_energy <= 0.0f) return out;
        float wl_prob = wl_energy / tot_energy;
        if (n < wl_prob) {
            float nn = n / wl_prob;
            // weighted_choice (algorithm.rs:17-36)
            float total_weight = 0.0f;
            for (uint32_t i = 0; i < d.n_distant_lights; ++i) total_weight = total_weight + distant_energy(d.distant_lights[i]);
            float target = nn * total_weight;
            float x = 0.0f, p = 0.0f;
            uint32_t pick = 0;
            for (uint32_t i = 0; i < d.n_distant_lights; ++i) {
                float wgt = distant_energy(d.distant_lights[i]);
                x += wgt;
                if (x > target || i == d.n_distant_lights) {
                    pick = i;
                    p = wgt / total_weight;
                    break;
                }
                if (i + 1 == d.n_distant_lights) return out;  // unreachable!() in the reference
            }
            // DistantDiskLight::sample_from_point (distant_disk_light.rs:59-85)
            const PsyDistantLight& l = d.distant_lights[pick];
            double radius = (double)lerp_slice(d.world_floats + l.radii_base, l.n_radii, time);
            V3 direction = lerp_slice_fn<V3>(l.n_dirs, time, [&](size_t k) { return b3(d.world_floats + l.dir_base + 3 * k); });
            PsyColor col = lerp_slice(d.light_colors + l.color_base, l.n_colors, time);
            double solid_angle_inv = 1.0 / (2.0 * PI_64 * (1.0 - m_cos(radius)));
            V3 z = -normalized(direction), x2, y2;
            coordinate_system_from_vector(z, x2, y2);
            double cos_theta_max = m_cos(radius);
            V3 sample = normalized(uniform_sample_cone(u, v, cos_theta_max));
            out.kind = 1;
            out.color = color_to_spectral(col, wl) * (float)solid_angle_inv;
            out.direction = ((x2 * sample.x) + (y2 * sample.y)) + (z * sample.z);
            out.pdf = (float)uniform_sample_cone_pdf(cos_theta_max);
            out.selection_pdf = p * wl_prob;
            return out;
        }
        float nn = (n - wl_prob) / (1.0f - wl_prob);
        xstack.clear();
        if (sample_assembly(d.root_assembly, nn, u, v, wl, time, intr, out)) {
            out.selection_pdf = out.selection_pdf * (1.0f - wl_prob);
            return out;
        }
        out.kind = 0;
        return out;
    }
};

//------------------------------------------------------------------------------------------------
// Camera + integrator

// Camera::generate_ray (camera.rs:74-102)
static Ray camera_ray(const PsySceneDesc& d, float x, float y, float time, float u, float v) {
    const PsyCamera& c = d.camera;
    Xf transform = lerp_slice_fn<Xf>(c.n_xforms, time, [&](size_t k) { return from_abi(c.xforms[k]); });
    float tfov = lerp_slice(c.tfovs, c.n_tfovs, time);
    float aperture_radius = lerp_slice(c.aperture_radii, c.n_aperture_radii, time);
    float focus_distance = lerp_slice(c.focus_distances, c.n_focus_distances, time);
    float cu, cv;
    square_to_circle((u * 2.0f) - 1.0f, (v * 2.0f) - 1.0f, cu, cv);
    V3 orig = mk(aperture_radius * cu, aperture_radius * cv, 0.0f);
    V3 dir = normalized(mk((x * tfov) - (orig.x / focus_distance), (y * tfov) - (orig.y / focus_distance), 1.0f));
    Ray r{};
    r.orig = xf_point(transform, orig);
    r.dir = xf_vector(transform, dir);
    r.time = time;
    r.max_t = std::numeric_limits<float>::infinity();
    r.occlusion = false;
    r.done = false;
    return r;
}

// Debug trace of one sample's path events (oracle_set_debug): prints every quantity LightPath::next derives.
static uint32_t g_dbg_px = 0xFFFFFFFFu, g_dbg_py = 0, g_dbg_si = 0;
static bool g_wave_stats = false;
#define ORC_DBG(...)                                                               \
    do {                                                                           \
        if (ORC_UNLIKELY(g_debug_trace) && px == g_dbg_px && py == g_dbg_py && sample_number == g_dbg_si) { \
            std::fprintf(stderr, "[oracle dbg] " __VA_ARGS__);                     \
            std::fprintf(stderr, "\n");                                            \
        }                                                                          \
    } while (0)
#define ORC_V4(q_) (double)(q_).v[0], (double)(q_).v[1], (double)(q_).v[2], (double)(q_).v[3]
#define ORC_V3(q_) (double)(q_).x, (double)(q_).y, (double)(q_).z

enum Event { EV_CAMERA, EV_BOUNCE, EV_SHADOW };
struct LightPath {
    Event event = EV_CAMERA;
    uint32_t bounce_count = 0;
    uint32_t sampling_seed = 0;
    uint32_t px = 0, py = 0, sample_number = 0, dim_offset = 0;
    float time = 0, wavelength = 0;
    bool has_next_bounce = false;
    Ray next_bounce_ray{};
    V4 next_attenuation_fac = splat4(1.0f);
    float closure_sample_pdf = 1.0f;
    V4 light_attenuation = splat4(1.0f);
    V4 pending_color_addition = splat4(0.0f);
    V4 color = splat4(0.0f);

    void next_lds_sequence() {
        dim_offset = 0;
        sampling_seed += 1;
    }
    void next_lds_samp(float out[4]) {
        uint32_t dimension = dim_offset;
        dim_offset += 1;
        get_sample_4d(sample_number, dimension, px, py, sampling_seed, out);
    }

    // LightPath::next (renderer.rs:449-691)
    bool next(LightSampler& ls, const PsySceneDesc& d, const Isect& isect, Ray& ray) {
        ORC_DBG("event %d bounce %u isect.kind %d hit_kind %u t %.9g prim %u path %u ray o (%.9g %.9g %.9g) d (%.9g %.9g %.9g) time %.9g wl %.9g",
                (int)event, bounce_count, (int)isect.kind, isect.hit_kind, (double)isect.t, isect.prim_id, isect.path_id, ORC_V3(ray.orig),
                ORC_V3(ray.dir), (double)time, (double)wavelength);
        if (event == EV_CAMERA || event == EV_BOUNCE) {
            if (isect.kind == ISECT_HIT) {
                const Closure& closure = isect.closure;
                if (closure.kind == PSY_SHADER_EMIT) {
                    V4 c = color_to_spectral(closure.color, wavelength);
                    if (event == EV_CAMERA) {
                        color = color + c;
                    } else {
                        float mis_pdf = power_heuristic(closure_sample_pdf, isect.sample_pdf);
                        color = color + ((c * light_attenuation) / mis_pdf);
                    }
                    return false;
                }
                ORC_DBG("  hit pos (%.9g %.9g %.9g) err %.9g nor (%.9g %.9g %.9g) nor_g (%.9g %.9g %.9g) closure %u atten (%.9g %.9g %.9g %.9g) cpdf %.9g",
                        ORC_V3(isect.pos), (double)isect.pos_err, ORC_V3(isect.nor), ORC_V3(isect.nor_g), closure.kind,
                        ORC_V4(light_attenuation), (double)closure_sample_pdf);
                light_attenuation = light_attenuation / closure_sample_pdf;

                next_lds_sequence();
                float s4[4];
                next_lds_samp(s4);
                if (ORC_UNLIKELY(g_debug_trace)) {
                    g_dbg_on = (px == g_dbg_px && py == g_dbg_py && sample_number == g_dbg_si);
                    sphere_lambert_debug() = g_dbg_on;
                }
                LightSample li = ls.sample_scene(s4[0], s4[1], s4[2], s4[3], wavelength, time, isect);
                if (ORC_UNLIKELY(g_debug_trace)) {
                    g_dbg_on = false;
                    sphere_lambert_debug() = false;
                }
                bool found_light = false;
                ORC_DBG("  light kind %d pdf %.9g sel %.9g color (%.9g %.9g %.9g %.9g) pos (%.9g %.9g %.9g) nor (%.9g %.9g %.9g) dir (%.9g %.9g %.9g) s4 (%.9g %.9g %.9g)",
                        li.kind, (double)li.pdf, (double)li.selection_pdf, ORC_V4(li.color), ORC_V3(li.pos), ORC_V3(li.nor),
                        ORC_V3(li.direction), (double)s4[0], (double)s4[1], (double)s4[2]);
                if (!(li.kind == 0 || li.pdf <= 0.0f || li.selection_pdf <= 0.0f)) {
                    V4 attenuation;
                    float closure_pdf;
                    Ray shadow{};
                    V3 world_dir = ray.dir;  // rays.dir(ray_idx): the ray that produced this hit
                    if (li.kind == 1) {
                        closure_evaluate(closure, world_dir, li.direction, isect.nor, isect.nor_g, wavelength, attenuation, closure_pdf);
                        V3 offset_pos = robust_ray_origin(isect.pos, isect.pos_err, normalized(isect.nor_g), li.direction);
                        shadow.orig = offset_pos;
                        shadow.dir = li.direction;
                        shadow.max_t = std::numeric_limits<float>::infinity();
                    } else {
                        V3 dir = li.pos - isect.pos;
                        closure_evaluate(closure, world_dir, dir, isect.nor, isect.nor_g, wavelength, attenuation, closure_pdf);
                        V3 offset_pos = robust_ray_origin(isect.pos, isect.pos_err, normalized(isect.nor_g), dir);
                        V3 offset_end = robust_ray_origin(li.pos, li.err, normalized(li.nor), -dir);
                        shadow.orig = offset_pos;
                        shadow.dir = offset_end - offset_pos;
                        shadow.max_t = 1.0f;
                    }
                    shadow.time = time;
                    shadow.occlusion = true;
                    ORC_DBG("  eval atten (%.9g %.9g %.9g %.9g) closure_pdf %.9g", ORC_V4(attenuation), (double)closure_pdf);
                    if (!(max_element(attenuation) <= 0.0f)) {
                        float light_mis_pdf = power_heuristic(li.pdf, closure_pdf);
                        pending_color_addition = ((li.color * attenuation) * light_attenuation) / (light_mis_pdf * li.selection_pdf);
                        ORC_DBG("  pending (%.9g %.9g %.9g %.9g) mis %.9g", ORC_V4(pending_color_addition), (double)light_mis_pdf);
                        ray = shadow;
                        found_light = true;
                    }
                }

                bool do_bounce = false;
                if (bounce_count < 2) {
                    bounce_count += 1;
                    next_lds_sequence();
                    float b4[4];
                    next_lds_samp(b4);
                    V3 dir;
                    V4 filter;
                    float pdf;
                    closure_sample(closure, isect.incoming, isect.nor, isect.nor_g, b4[0], b4[1], wavelength, dir, filter, pdf);
                    ORC_DBG("  bounce dir (%.9g %.9g %.9g) filter (%.9g %.9g %.9g %.9g) pdf %.9g", ORC_V3(dir), ORC_V4(filter), (double)pdf);
                    if ((pdf > 0.0f) && (max_element(filter) > 0.0f)) {
                        next_attenuation_fac = filter;
                        closure_sample_pdf = pdf;
                        V3 offset_pos = robust_ray_origin(isect.pos, isect.pos_err, normalized(isect.nor_g), dir);
                        next_bounce_ray = Ray{};
                        next_bounce_ray.orig = offset_pos;
                        next_bounce_ray.dir = dir;
                        next_bounce_ray.time = time;
                        next_bounce_ray.max_t = std::numeric_limits<float>::infinity();
                        next_bounce_ray.occlusion = false;
                        has_next_bounce = true;
                        do_bounce = true;
                    }
                } else {
                    has_next_bounce = false;
                }

                if (found_light) {
                    event = EV_SHADOW;
                    return true;
                } else if (do_bounce) {
                    ray = next_bounce_ray;
                    event = EV_BOUNCE;
                    light_attenuation = light_attenuation * next_attenuation_fac;
                    return true;
                }
                return false;
            } else {
                V4 bg = color_to_spectral(d.background, wavelength);
                color = color + ((bg * light_attenuation) / closure_sample_pdf);
                return false;
            }
        }
        // EV_SHADOW
        if (isect.kind == ISECT_MISS) color = color + pending_color_addition;
        if (has_next_bounce) {
            ray = next_bounce_ray;
            light_attenuation = light_attenuation * next_attenuation_fac;
            event = EV_BOUNCE;
            return true;
        }
        return false;
    }
};

struct BucketResult {
    uint64_t rays = 0;
    Counters cnt;
    double gen_s = 0, trace_s = 0, shade_s = 0, film_s = 0;
};

// render_job body for one bucket (renderer.rs:239-310).  xyz: accumulated (+=) into an image of
// stride `img_w` pixels whose origin is (img_x0, img_y0).  samples_out (optional): per-sample
// LightPath.color, indexed [((y-by)*w + (x-bx))*spp + si][4].
static void render_bucket(const PsySceneDesc& d, const Paths& paths, uint32_t bx, uint32_t by, uint32_t bw, uint32_t bh,
                          uint32_t spp, uint32_t seed, float* xyz, uint32_t img_w, uint32_t img_x0, uint32_t img_y0,
                          float* samples_out, BucketResult& res) {
    using clk = std::chrono::steady_clock;
    auto t0 = clk::now();
    std::vector<LightPath> lpaths;
    std::vector<Ray> rays;
    std::vector<uint32_t> order;  // original index of the path now at slot i
    lpaths.reserve((size_t)bw * bh * spp);
    rays.reserve((size_t)bw * bh * spp);
    float cmpx = 1.0f / (float)d.res_x, cmpy = 1.0f / (float)d.res_y;
    float min_x = -1.0f, max_x = 1.0f;
    float min_y = -((float)d.res_y / (float)d.res_x), max_y = (float)d.res_y / (float)d.res_x;
    float x_extent = max_x - min_x, y_extent = max_y - min_y;
    for (uint32_t y = by; y < by + bh; ++y)
        for (uint32_t x = bx; x < bx + bw; ++x)
            for (uint32_t si = 0; si < spp; ++si) {
                float s0[4], s1[4];
                get_sample_4d(si, 0, x, y, seed, s0);
                get_sample_4d(si, 1, x, y, seed, s1);
                float filter_x = probit(s0[3], 2.0f / 6.0f) + 0.5f;
                float filter_y = probit(s1[0], 2.0f / 6.0f) + 0.5f;
                float samp_x = (filter_x + (float)x) * cmpx;
                float samp_y = (filter_y + (float)y) * cmpy;
                float img_x = (samp_x - 0.5f) * x_extent;
                float img_y = (0.5f - samp_y) * y_extent;
                LightPath p;
                p.sampling_seed = seed ^ 0x40d4682bu;
                p.px = x, p.py = y;
                p.sample_number = si;
                p.time = s0[0];
                p.wavelength = map_0_1_to_wavelength(golden_ratio_sample(si, hash_u32((x << 16) ^ y, seed)));
                lpaths.push_back(p);
                rays.push_back(camera_ray(d, img_x, img_y, s0[0], s0[1], s0[2]));
                order.push_back((uint32_t)order.size());
            }
    auto t1 = clk::now();
    res.gen_s += std::chrono::duration<double>(t1 - t0).count();

    Tracer tracer(d, paths);
    LightSampler sampler(d);
    std::vector<Isect> isects;
    size_t pi = lpaths.size();
    while (pi > 0) {
        auto ta = clk::now();
        isects.resize(pi);
        const Counters before = tracer.cnt;
        for (size_t i = 0; i < pi; ++i) isects[i] = tracer.trace(rays[i]);
        res.rays += pi;
        if (ORC_UNLIKELY(g_wave_stats))  // oracle_set_wave_stats: per-wave visit counts (what kind of ray costs what)
            std::fprintf(stderr, "[oracle waves] %zu rays: %.2f node visits (%.2f two-key), %.2f triangle tests, %.2f instance entries per ray\n",
                         pi, (double)(tracer.cnt.node_visits - before.node_visits) / (double)pi,
                         (double)(tracer.cnt.motion_node_visits - before.motion_node_visits) / (double)pi,
                         (double)(tracer.cnt.tri_tests - before.tri_tests) / (double)pi,
                         (double)(tracer.cnt.instance_visits - before.instance_visits) / (double)pi);
        auto tb = clk::now();
        res.trace_s += std::chrono::duration<double>(tb - ta).count();
        size_t new_end = 0;
        for (size_t i = 0; i < pi; ++i) {
            if (lpaths[i].next(sampler, d, isects[i], rays[i])) {
                std::swap(lpaths[new_end], lpaths[i]);
                std::swap(rays[new_end], rays[i]);
                std::swap(order[new_end], order[i]);
                new_end += 1;
            }
        }
        pi = new_end;
        res.shade_s += std::chrono::duration<double>(clk::now() - tb).count();
    }
    auto tf = clk::now();
    for (size_t i = 0; i < lpaths.size(); ++i) {
        const LightPath& p = lpaths[i];
        float col[3];
        xyz_from_spectral(p.color, p.wavelength, col);
        if (xyz) {
            float* px = xyz + ((size_t)(p.py - img_y0) * img_w + (p.px - img_x0)) * 3;
            px[0] += col[0] / (float)spp;
            px[1] += col[1] / (float)spp;
            px[2] += col[2] / (float)spp;
        }
        if (samples_out) {
            float* so = samples_out + (size_t)order[i] * 4;
            for (int k = 0; k < 4; ++k) so[k] = p.color.v[k];
        }
    }
    res.film_s += std::chrono::duration<double>(clk::now() - tf).count();
    res.cnt.node_visits += tracer.cnt.node_visits;
    res.cnt.tri_tests += tracer.cnt.tri_tests;
    res.cnt.instance_visits += tracer.cnt.instance_visits;
    res.cnt.motion_node_visits += tracer.cnt.motion_node_visits;
}

static void hilbert_d2xy(uint32_t dd, uint32_t& x, uint32_t& y) {
    x = 0, y = 0;
    uint32_t t = dd;
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
}

}  // namespace orc

using namespace orc;

extern "C" {

struct OracleStats {
    double trace_time, initial_ray_generation_time, ray_generation_time, sample_writing_time, total_time;
    uint64_t ray_count, accel_node_visits, tri_tests, instance_visits, motion_node_visits, samples;
    uint32_t threads;
};

int oracle_trace_rays(const PsySceneDesc* d, const PsyRay* rays, size_t n, PsyHit* hits, PsyHitGeo* geo, uint64_t* counters) {
    Paths paths = expand_paths(*d);
    if (paths.overflow) return 4;
    Tracer tracer(*d, paths);
    for (size_t i = 0; i < n; ++i) {
        Ray r{};
        r.orig = mk(rays[i].orig[0], rays[i].orig[1], rays[i].orig[2]);
        r.dir = mk(rays[i].dir[0], rays[i].dir[1], rays[i].dir[2]);
        r.time = rays[i].time;
        r.max_t = rays[i].max_t;
        r.occlusion = (rays[i].flags & PSY_RAY_OCCLUSION) != 0;
        Isect is = tracer.trace(r);
        PsyHit& h = hits[i];
        h.kind = is.hit_kind;
        h.t = is.kind == ISECT_HIT ? is.t : 0.0f;
        h.prim_id = is.prim_id;
        h.inst_id = is.inst_id;
        h.path_id = is.path_id;
        h.b0 = is.b0, h.b1 = is.b1, h.b2 = is.b2;
        if (geo) {
            PsyHitGeo& g = geo[i];
            std::memset(&g, 0, sizeof(g));
            if (is.kind == ISECT_HIT) {
                g.pos[0] = is.pos.x, g.pos[1] = is.pos.y, g.pos[2] = is.pos.z;
                g.pos_err = is.pos_err;
                g.nor[0] = is.nor.x, g.nor[1] = is.nor.y, g.nor[2] = is.nor.z;
                g.sample_pdf = is.sample_pdf;
                g.nor_g[0] = is.nor_g.x, g.nor_g[1] = is.nor_g.y, g.nor_g[2] = is.nor_g.z;
                g.closure_kind = is.closure.kind;
            }
        }
    }
    if (counters) {
        counters[0] = tracer.cnt.node_visits;
        counters[1] = tracer.cnt.tri_tests;
        counters[2] = tracer.cnt.instance_visits;
        counters[3] = tracer.cnt.motion_node_visits;
    }
    return 0;
}

// One bucket; xyz_out (w*h*3) receives the per-pixel mean, samples_out (w*h*spp*4) may be NULL.
int oracle_render_bucket(const PsySceneDesc* d, uint32_t x, uint32_t y, uint32_t w, uint32_t h, uint32_t spp, uint32_t seed,
                         float* xyz_out, float* samples_out, OracleStats* st) {
    Paths paths = expand_paths(*d);
    if (paths.overflow) return 4;
    if (xyz_out) std::fill(xyz_out, xyz_out + (size_t)w * h * 3, 0.0f);
    BucketResult r;
    auto t0 = std::chrono::steady_clock::now();
    render_bucket(*d, paths, x, y, w, h, spp, seed, xyz_out, w, x, y, samples_out, r);
    if (st) {
        std::memset(st, 0, sizeof(*st));
        st->trace_time = r.trace_s;
        st->initial_ray_generation_time = r.gen_s;
        st->ray_generation_time = r.shade_s;
        st->sample_writing_time = r.film_s;
        st->total_time = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        st->ray_count = r.rays;
        st->accel_node_visits = r.cnt.node_visits;
        st->tri_tests = r.cnt.tri_tests;
        st->instance_visits = r.cnt.instance_visits;
        st->motion_node_visits = r.cnt.motion_node_visits;
        st->samples = (uint64_t)w * h * spp;
        st->threads = 1;
    }
    return 0;
}

// Renderer::render (renderer.rs:76-191): Hilbert-ordered buckets over `threads` workers.
// crop = (x1,y1,x2,y2) inclusive or NULL.  xyz_out: res_x*res_y*3 (row 0 = top).
// max_buckets > 0 renders only that many buckets, spread evenly over the Hilbert-ordered list (bounded
// CPU-baseline sample); buckets_done reports how many were rendered.
int oracle_render_image(const PsySceneDesc* d, const uint32_t* crop, uint32_t spp, uint32_t seed, uint32_t spb,
                        uint32_t threads, uint32_t max_buckets, float* xyz_out, OracleStats* st, uint32_t* buckets_done) {
    Paths paths = expand_paths(*d);
    if (paths.overflow) return 4;
    uint32_t img_w = d->res_x, img_h = d->res_y;
    if (xyz_out) std::fill(xyz_out, xyz_out + (size_t)img_w * img_h * 3, 0.0f);
    uint32_t width = img_w, height = img_h, start_x = 0, start_y = 0;
    if (crop) {
        uint32_t x1 = std::min(crop[0], img_w - 1), y1 = std::min(crop[1], img_h - 1);
        uint32_t x2 = std::min(crop[2], img_w - 1), y2 = std::min(crop[3], img_h - 1);
        width = x2 - x1 + 1, height = y2 - y1 + 1, start_x = x1, start_y = y1;
    }
    double target = (double)spb / (double)spp;
    uint32_t bdim = std::sqrt(target) < 1.0 ? 1u : (uint32_t)std::sqrt(target);
    struct Job {
        uint32_t x, y, w, h;
    };
    std::vector<Job> jobs;
    {
        uint32_t cx = width / bdim + 1, cy = height / bdim + 1;
        uint32_t larger = std::max(cx, cy), pow2 = 1;
        while (pow2 < larger) pow2 <<= 1;
        uint64_t bucket_n = (uint64_t)pow2 * pow2;
        for (uint64_t hd = 0; hd < bucket_n; ++hd) {
            uint32_t bx, by;
            hilbert_d2xy((uint32_t)hd, bx, by);
            uint64_t x = (uint64_t)bx * bdim, y = (uint64_t)by * bdim;
            if (x >= width || y >= height) continue;
            uint32_t w = (uint32_t)std::min<uint64_t>(bdim, width - x), h = (uint32_t)std::min<uint64_t>(bdim, height - y);
            if (w > 0 && h > 0) jobs.push_back({start_x + (uint32_t)x, start_y + (uint32_t)y, w, h});
        }
    }
    if (max_buckets > 0 && jobs.size() > max_buckets) {
        // bounded sample: every k-th bucket of the Hilbert list, so the sample covers the whole image
        // (the first N buckets alone would be one corner of it)
        std::vector<Job> picked(max_buckets);
        for (uint32_t i = 0; i < max_buckets; ++i) picked[i] = jobs[(size_t)i * jobs.size() / max_buckets];
        jobs.swap(picked);
    }
    if (threads == 0) threads = std::max(1u, std::thread::hardware_concurrency());
    std::atomic<size_t> next{0};
    std::vector<BucketResult> results(threads);
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (uint32_t t = 0; t < threads; ++t) {
        pool.emplace_back([&, t]() {
            for (;;) {
                size_t j = next.fetch_add(1);
                if (j >= jobs.size()) break;
                const Job& b = jobs[j];
                render_bucket(*d, paths, b.x, b.y, b.w, b.h, spp, seed, xyz_out, img_w, 0, 0, nullptr, results[t]);
            }
        });
    }
    for (auto& th : pool) th.join();
    if (buckets_done) *buckets_done = (uint32_t)jobs.size();
    if (st) {
        std::memset(st, 0, sizeof(*st));
        for (auto& r : results) {
            st->trace_time += r.trace_s;
            st->initial_ray_generation_time += r.gen_s;
            st->ray_generation_time += r.shade_s;
            st->sample_writing_time += r.film_s;
            st->ray_count += r.rays;
            st->accel_node_visits += r.cnt.node_visits;
            st->tri_tests += r.cnt.tri_tests;
            st->instance_visits += r.cnt.instance_visits;
            st->motion_node_visits += r.cnt.motion_node_visits;
        }
        st->total_time = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        uint64_t px = 0;
        for (auto& j : jobs) px += (uint64_t)j.w * j.h;
        st->samples = px * spp;
        st->threads = threads;
    }
    return 0;
}

// Camera rays of a bucket in generation order (for probe-ray parity tests).
int oracle_camera_rays(const PsySceneDesc* d, uint32_t x0, uint32_t y0, uint32_t w, uint32_t h, uint32_t spp, uint32_t seed,
                       PsyRay* out) {
    float cmpx = 1.0f / (float)d->res_x, cmpy = 1.0f / (float)d->res_y;
    float y_extent = ((float)d->res_y / (float)d->res_x) - -((float)d->res_y / (float)d->res_x);
    size_t k = 0;
    for (uint32_t y = y0; y < y0 + h; ++y)
        for (uint32_t x = x0; x < x0 + w; ++x)
            for (uint32_t si = 0; si < spp; ++si) {
                float s0[4], s1[4];
                get_sample_4d(si, 0, x, y, seed, s0);
                get_sample_4d(si, 1, x, y, seed, s1);
                float fx = probit(s0[3], 2.0f / 6.0f) + 0.5f, fy = probit(s1[0], 2.0f / 6.0f) + 0.5f;
                float img_x = (((fx + (float)x) * cmpx) - 0.5f) * 2.0f;
                float img_y = (0.5f - ((fy + (float)y) * cmpy)) * y_extent;
                Ray r = camera_ray(*d, img_x, img_y, s0[0], s0[1], s0[2]);
                PsyRay& o = out[k++];
                o.orig[0] = r.orig.x, o.orig[1] = r.orig.y, o.orig[2] = r.orig.z;
                o.dir[0] = r.dir.x, o.dir[1] = r.dir.y, o.dir[2] = r.dir.z;
                o.time = r.time;
                o.max_t = r.max_t;
                o.flags = 0;
            }
    return 0;
}

//---- unit-level entry points used to pin the restatement against the reference's own tests -------
void oracle_set_strict_math(int on) { strict_math() = on != 0; }
void oracle_set_wave_stats(int on) { g_wave_stats = on != 0; }
// sphere_lambert (surface_closure.rs:351-377) and the g / h terms it evaluates before branching
float oracle_sphere_lambert(float nlcos, float rcos, float* g_h_nl_r) {
    if (g_h_nl_r) {
        float nlsin = std::sqrt(1.0f - (nlcos * nlcos));
        float rsin2 = 1.0f - (rcos * rcos);
        float rsin = std::sqrt(rsin2);
        float ysin = rcos / nlsin;
        float ycos2 = 1.0f - (ysin * ysin);
        float ycos = std::sqrt(ycos2);
        g_h_nl_r[0] = (-2.0f * nlsin * rcos * ycos) + H_PI - m_asinf(ysin) + (ysin * ycos);
        g_h_nl_r[1] = nlcos * ((ycos * std::sqrt(rsin2 - ycos2)) + (rsin2 * m_asinf(ycos / rsin)));
        g_h_nl_r[2] = m_acosf(nlcos);
        g_h_nl_r[3] = m_acosf(rcos);
    }
    return sphere_lambert(nlcos, rcos);
}
double oracle_det_f64(int fn, double x) {
    return fn == 0 ? psy_det_sin(x) : fn == 1 ? psy_det_cos(x) : fn == 2 ? psy_det_acos(x) : fn == 3 ? psy_det_asin(x) : psy_det_log(x);
}
float oracle_det_f32(int fn, float x) {
    return fn == 0 ? psy_det_sinf(x) : fn == 1 ? psy_det_cosf(x) : fn == 2 ? psy_det_acosf(x) : psy_det_asinf(x);
}
void oracle_set_debug(uint32_t px, uint32_t py, uint32_t si) {
    g_dbg_px = px, g_dbg_py = py, g_dbg_si = si;
    g_debug_trace = px != 0xFFFFFFFFu;
}
uint32_t oracle_hash_u32(uint32_t n, uint32_t seed) { return hash_u32(n, seed); }
float oracle_golden_ratio_sample(uint32_t i, uint32_t scramble) { return golden_ratio_sample(i, scramble); }
void oracle_get_sample_4d(uint32_t i, uint32_t ds, uint32_t px, uint32_t py, uint32_t seed, float* out) {
    get_sample_4d(i, ds, px, py, seed, out);
}
float oracle_probit(float n, float width) { return probit(n, width); }
float oracle_fp_gamma(uint32_t n) { return fp_gamma(n); }
float oracle_increment_ulp(float v) { return increment_ulp(v); }
float oracle_decrement_ulp(float v) { return decrement_ulp(v); }
void oracle_hilbert_d2xy(uint32_t dd, uint32_t* x, uint32_t* y) { hilbert_d2xy(dd, *x, *y); }
void oracle_traversal_table(uint8_t* out) { std::memcpy(out, TRAVERSAL.t, 384); }
float oracle_fast_exp(float p) { return fast_exp(p); }
float oracle_power_heuristic(float a, float b) { return power_heuristic(a, b); }
void oracle_xform_mul(const PsyXform* a, const PsyXform* b, PsyXform* out) {
    Xf r = xf_mul(from_abi(*a), from_abi(*b));
    const float* c[4] = {&r.X.x, &r.Y.x, &r.Z.x, &r.T.x};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 4; ++j) out->r[i][j] = c[j][i];
}
void oracle_xform_inverse(const PsyXform* a, PsyXform* out) {
    Xf r = xf_inverse(from_abi(*a));
    const float* c[4] = {&r.X.x, &r.Y.x, &r.Z.x, &r.T.x};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 4; ++j) out->r[i][j] = c[j][i];
}
void oracle_xform_lerp(const PsyXform* a, const PsyXform* b, float alpha, PsyXform* out) {
    Xf r = lerp(from_abi(*a), from_abi(*b), alpha);
    const float* c[4] = {&r.X.x, &r.Y.x, &r.Z.x, &r.T.x};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 4; ++j) out->r[i][j] = c[j][i];
}
void oracle_xform_point(const PsyXform* a, const float* p, float* out) {
    V3 r = xf_point(from_abi(*a), mk(p[0], p[1], p[2]));
    out[0] = r.x, out[1] = r.y, out[2] = r.z;
}
void oracle_xform_vector(const PsyXform* a, const float* p, float* out) {
    V3 r = xf_vector(from_abi(*a), mk(p[0], p[1], p[2]));
    out[0] = r.x, out[1] = r.y, out[2] = r.z;
}
void oracle_xform_normal(const PsyXform* a, const float* p, float* out) {
    V3 r = xf_normal(from_abi(*a), mk(p[0], p[1], p[2]));
    out[0] = r.x, out[1] = r.y, out[2] = r.z;
}
float oracle_lerp_slice_f32(const float* s, uint32_t n, float alpha) { return lerp_slice(s, n, alpha); }
void oracle_color_to_spectral(const PsyColor* c, float hero, float* out4) {
    V4 r = color_to_spectral(*c, hero);
    for (int i = 0; i < 4; ++i) out4[i] = r.v[i];
}
void oracle_xyz_from_spectral(const float* e4, float hero, float* out3) {
    V4 e{{e4[0], e4[1], e4[2], e4[3]}};
    xyz_from_spectral(e, hero, out3);
}
int oracle_tri_intersect(const float* orig, const float* dir, float max_t, const float* tri9, float* out4) {
    RayTriPre pre = tri_precompute(mk(dir[0], dir[1], dir[2]));
    float t, b0, b1, b2;
    bool hit = tri_intersect(mk(orig[0], orig[1], orig[2]), pre, max_t, mk(tri9[0], tri9[1], tri9[2]),
                             mk(tri9[3], tri9[4], tri9[5]), mk(tri9[6], tri9[7], tri9[8]), t, b0, b1, b2);
    if (hit) out4[0] = t, out4[1] = b0, out4[2] = b1, out4[3] = b2;
    return hit ? 1 : 0;
}

// math3d's vector / normal / point unit tests (sub_crates/math3d/src/{vector,normal,point}.rs): op = 0 add, 1 sub,
// 2 mul scalar (b[0]), 3 div scalar (b[0]), 4 cross, 5 normalized; scalar results: 0 dot, 1 length, 2 length2
void oracle_vec3_op(int op, const float* a, const float* b, float* out) {
    const V3 x = mk(a[0], a[1], a[2]), y = b ? mk(b[0], b[1], b[2]) : mk(0, 0, 0);
    V3 r = x;
    switch (op) {
        case 0: r = x + y; break;
        case 1: r = x - y; break;
        case 2: r = x * b[0]; break;
        case 3: r = x / b[0]; break;
        case 4: r = cross(x, y); break;
        default: r = normalized(x); break;
    }
    out[0] = r.x, out[1] = r.y, out[2] = r.z;
}
float oracle_vec3_scalar(int op, const float* a, const float* b) {
    const V3 x = mk(a[0], a[1], a[2]), y = b ? mk(b[0], b[1], b[2]) : mk(0, 0, 0);
    return op == 0 ? dot(x, y) : (op == 1 ? length(x) : length2(x));
}

}  // extern "C"
