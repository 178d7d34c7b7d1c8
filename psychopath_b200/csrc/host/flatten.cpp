// flatten.cpp — SceneIR -> PsySceneDesc: builds every mesh BVH4, every assembly's instance BVH4
// and light tree the way the reference's AssemblyBuilder does, and lays the result out as the flat
// SoA arrays the C ABI takes.
//
// Reference: src/surface/triangle_mesh.rs:29-115 (from_verts_and_indices),
//            src/scene/assembly.rs:288-395 (build, instance_bounds),
//            src/accel/light_tree.rs:62-131,212-372 (LightTree::from_objects, LightTreeBuilder),
//            src/camera.rs:22-72 (Camera::new), src/light/*.rs (bounds, approximate_energy).
#include <chrono>
#include <cstring>
#include <map>
#include <numeric>
#include <stdexcept>

#include "bvh_build.h"
#include "scene_ir.h"

namespace psyh {

namespace {
std::mutex g_topo_mu;
MeshTopologyFn g_topo_fn = nullptr;
void* g_topo_ctx = nullptr;
uint32_t g_topo_min_tris = 0;
double g_topo_seconds = 0.0;
uint32_t g_topo_meshes = 0;
MeshBvhBuildFn g_bvh_build = nullptr;
MeshBvhTakeFn g_bvh_take = nullptr;
void* g_bvh_ctx = nullptr;
uint32_t g_bvh_min_tris = 0;
}  // namespace

void set_mesh_bvh_builder(MeshBvhBuildFn build, MeshBvhTakeFn take, void* ctx, uint32_t min_tris) {
    std::lock_guard<std::mutex> g(g_topo_mu);
    g_bvh_build = build, g_bvh_take = take, g_bvh_ctx = ctx, g_bvh_min_tris = min_tris;
}

void set_mesh_topology_builder(MeshTopologyFn fn, void* ctx, uint32_t min_tris) {
    std::lock_guard<std::mutex> g(g_topo_mu);
    g_topo_fn = fn, g_topo_ctx = ctx, g_topo_min_tris = min_tris;
}
double take_mesh_topology_seconds(uint32_t* meshes_built) {
    std::lock_guard<std::mutex> g(g_topo_mu);
    double s = g_topo_seconds;
    if (meshes_built) *meshes_built = g_topo_meshes;
    g_topo_seconds = 0.0, g_topo_meshes = 0;
    return s;
}

PsyColor color_from_rec709(float r, float g, float b) {
    PsyColor c;
    c.kind = PSY_COLOR_XYZ;
    c.a = (r * 0.4969210174f) + (g * 0.3390896921f) + (b * 0.1639892905f);
    c.b = (r * 0.2562248996f) + (g * 0.6781793842f) + (b * 0.0655957162f);
    c.c = (r * 0.0232931727f) + (g * 0.1130298974f) + (b * 0.8636769299f);
    return c;
}

// Color::approximate_energy (src/color.rs:127-142)
static float color_energy(const PsyColor& c) {
    switch (c.kind) {
        case PSY_COLOR_XYZ: return c.b;
        case PSY_COLOR_BLACKBODY: {
            float t2 = c.a * c.a;
            return t2 * t2 * c.b;
        }
        default: return c.b;
    }
}

static float colors_energy(const std::vector<PsyColor>& cols) {
    float a = 0.0f;
    for (auto& c : cols) a = a + color_energy(c);
    return a / (float)cols.size();
}

namespace {

struct LtbNode {
    bool is_leaf;
    uint32_t b0, b1;
    float energy;
    uint32_t child_index;
};

struct LightTreeBuilder {
    std::vector<LtbNode> nodes;
    std::vector<BBox> bounds;
    const std::vector<BBox>* tb;  // per light id, bounds at 0.5
    ObjBoundsView ob;
    const std::vector<float>* energy;

    // recursive_build (light_tree.rs:298-372)
    std::pair<uint32_t, std::pair<uint32_t, uint32_t>> build(uint32_t offset, uint32_t* objs, size_t n) {
        uint32_t me = (uint32_t)nodes.size();
        if (n == 0) return {0, {0, 0}};
        if (n == 1) {
            uint32_t bi = (uint32_t)bounds.size();
            const BBox* b = ob.ptr(objs[0]);
            bounds.insert(bounds.end(), b, b + ob.count(objs[0]));
            nodes.push_back(LtbNode{true, bi, (uint32_t)bounds.size(), (*energy)[objs[0]], offset});
            return {me, {bi, (uint32_t)bounds.size()}};
        }
        nodes.push_back(LtbNode{false, 0, 0, 0.0f, 0});
        auto sp = sah_split_ids(objs, n, tb->data());
        size_t split = sp.first;
        auto c1 = build(offset, objs, split);
        auto c2 = build(offset + (uint32_t)split, objs + split, n - split);
        uint32_t bi = (uint32_t)bounds.size();
        std::vector<BBox> merged;
        merge_slices_append(bounds.data() + c1.second.first, c1.second.second - c1.second.first,
                            bounds.data() + c2.second.first, c2.second.second - c2.second.first, merged,
                            [](const BBox& a, const BBox& b) { return bb_union(a, b); });
        bounds.insert(bounds.end(), merged.begin(), merged.end());
        float e = nodes[me + 1].energy + nodes[c2.first].energy;
        nodes[me] = LtbNode{false, bi, (uint32_t)bounds.size(), e, c2.first};
        return {me, {bi, (uint32_t)bounds.size()}};
    }

    size_t child_count(size_t level, uint32_t idx) const {
        if (level > 0) {
            if (nodes[idx].is_leaf) return 1;
            return child_count(level - 1, idx + 1) + child_count(level - 1, nodes[idx].child_index);
        }
        return 1;
    }
    std::pair<uint32_t, size_t> nth_child(size_t level, uint32_t idx, size_t n) const {
        if (level > 0 && !nodes[idx].is_leaf) {
            auto r = nth_child(level - 1, idx + 1, n);
            if (r.second == 0) return {r.first, 0};
            return nth_child(level - 1, nodes[idx].child_index, r.second - 1);
        }
        return {idx, n};
    }
};

struct Flattener {
    FlatScene& fs;
    std::map<const MeshIR*, uint32_t> mesh_cache;
    std::map<const LightIR*, uint32_t> light_cache;
    std::vector<float> light_energy;  // per global light

    explicit Flattener(FlatScene& f) : fs(f) {}

    void append_bvh(const Bvh4Flat& bvh, uint32_t& root_out) {
        uint32_t node_base = (uint32_t)fs.nodes.size();
        uint32_t key_base = (uint32_t)fs.motion_bounds.size();
        fs.nodes.resize((size_t)node_base + bvh.nodes.size());
        parallel_for(bvh.nodes.size(), 65536, [&](size_t b0, size_t e0) {
            for (size_t j = b0; j < e0; ++j) {
                PsyBvh4Node nd = bvh.nodes[j];
                for (int i = 0; i < 4; ++i) {
                    uint32_t c = nd.child[i];
                    if (c != PSY_NONE && !PSY_CHILD_IS_LEAF(c)) nd.child[i] = c + node_base;
                }
                if (nd.n_keys > 1) nd.key_base += key_base;
                fs.nodes[(size_t)node_base + j] = nd;
            }
        });
        fs.motion_bounds.insert(fs.motion_bounds.end(), bvh.motion.begin(), bvh.motion.end());
        uint32_t r = bvh.root;
        if (r != PSY_NONE && !PSY_CHILD_IS_LEAF(r)) r += node_base;
        root_out = r;
        if (fs.nodes.size() >= 0x7FFFFFFFu) throw std::runtime_error("too many BVH nodes");
    }

    uint32_t add_mesh(const MeshIR& m) {
        auto it = mesh_cache.find(&m);
        if (it != mesh_cache.end()) return it->second;
        if (m.verts.empty() || m.verts[0].empty()) throw std::runtime_error("mesh without vertices");
        size_t k = m.verts.size(), vc = m.verts[0].size(), nt = m.tris.size();
        for (auto& v : m.verts)
            if (v.size() != vc) throw std::runtime_error("mesh time samples differ in vertex count");
        if (!m.normals.empty()) {
            if (m.normals.size() != k) throw std::runtime_error("normal time samples differ from vertex time samples");
            for (auto& v : m.normals)
                if (v.size() != vc) throw std::runtime_error("normal count differs from vertex count");
        }
        PsyMesh pm{};
        pm.n_keys = (uint32_t)k;
        pm.vert_count = (uint32_t)vc;
        pm.vert_base = (uint32_t)(fs.verts.size() / 3);
        pm.has_normals = m.normals.empty() ? 0u : 1u;
        {
            // the time keys of a vertex contiguous (PsySceneDesc.verts)
            const size_t o0 = (size_t)pm.vert_base * 3;
            fs.verts.resize(o0 + vc * k * 3);
            fs.normals.resize(fs.verts.size(), 0.0f);
            parallel_for(vc, 65536, [&](size_t b0, size_t e0) {
                for (size_t vi = b0; vi < e0; ++vi)
                    for (size_t ti = 0; ti < k; ++ti) {
                        float* o = fs.verts.data() + o0 + (vi * k + ti) * 3;
                        const V3& p = m.verts[ti][vi];
                        o[0] = p.x, o[1] = p.y, o[2] = p.z;
                        if (pm.has_normals) {
                            float* q = fs.normals.data() + o0 + (vi * k + ti) * 3;
                            const V3& nn = m.normals[ti][vi];
                            q[0] = nn.x, q[1] = nn.y, q[2] = nn.z;
                        }
                    }
            });
        }
        const bool timing = getenv("PSY_BUILD_TIMING") != nullptr && nt > 100000;
        auto tm0 = std::chrono::steady_clock::now();
        for (size_t i = 0; i < nt; ++i) {
            auto& t = m.tris[i];
            if (t[0] >= vc || t[1] >= vc || t[2] >= vc) throw std::runtime_error("triangle index out of range");
        }
        std::vector<uint32_t> ids(nt);
        std::iota(ids.begin(), ids.end(), 0u);
        pm.tri_base = (uint32_t)(fs.tris.size() / 4);
        pm.tri_count = (uint32_t)nt;
        if ((size_t)pm.tri_base + nt > PSY_MAX_LEAF_START) throw std::runtime_error("too many triangles");
        Bvh4Flat bvh;
        bool built = false, in_place = false;
        {
            // the whole BVH4 of the mesh from the registered external (GPU) builder, written straight
            // into the scene's node / key arrays
            MeshBvhBuildFn fb;
            MeshBvhTakeFn ft;
            void* fctx;
            uint32_t fmin;
            {
                std::lock_guard<std::mutex> g(g_topo_mu);
                fb = g_bvh_build, ft = g_bvh_take, fctx = g_bvh_ctx, fmin = g_bvh_min_tris;
            }
            if (fb && ft && nt >= fmin && nt > 3) {
                auto t0 = std::chrono::steady_clock::now();
                PsyMeshBvh* h = nullptr;
                PsyMeshBvhInfo info{};
                const size_t node_base = fs.nodes.size(), key_base = fs.motion_bounds.size();
                int32_t rc = node_base + nt < 0x7FFFFFFFu
                                 ? fb(fctx, fs.verts.data() + (size_t)pm.vert_base * 3, (uint32_t)vc, (uint32_t)k, &m.tris[0][0],
                                      (uint32_t)nt, 3, pm.tri_base, (uint32_t)node_base, (uint32_t)key_base, &h, &info)
                                 : -1;
                if (rc == 0 && info.binary_depth <= BVH_MAX_DEPTH) {
                    fs.nodes.resize(node_base + info.n_nodes);
                    fs.motion_bounds.resize(key_base + info.n_motion);
                    std::vector<PsyBBox> rb(info.n_root_bounds);
                    rc = ft(h, ids.data(), fs.nodes.data() + node_base, fs.motion_bounds.data() + key_base, rb.data());
                    if (rc != 0) {
                        fs.nodes.resize(node_base);
                        fs.motion_bounds.resize(key_base);
                    } else {
                        in_place = true;
                        bvh.root = info.root;
                        bvh.depth = info.depth;
                        for (auto& b : rb) {
                            BBox x;
                            x.min = v3(b.min[0], b.min[1], b.min[2]);
                            x.max = v3(b.max[0], b.max[1], b.max[2]);
                            bvh.bounds.push_back(x);
                        }
                        built = true;
                        std::lock_guard<std::mutex> g(g_topo_mu);
                        g_topo_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                        g_topo_meshes += 1;
                    }
                } else if (rc == 0) {
                    ft(h, nullptr, nullptr, nullptr, nullptr);  // deeper than the reference allows: dropped
                }
                if (!built) {
                    bvh = Bvh4Flat();
                    std::iota(ids.begin(), ids.end(), 0u);
                }
                if (timing && built)
                    std::fprintf(stderr, "[build] mesh %zu triangles x %zu keys: BVH4 from the device builder %.3fs\n", nt, k,
                                 std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
            }
        }
        // per-triangle bounds for every time sample (triangle_mesh.rs:86-99)
        std::vector<BBox> bounds;
        std::vector<uint32_t> begin;
        if (!built) {
            tm0 = std::chrono::steady_clock::now();
            bounds.resize(nt * k);
            begin.resize(nt + 1);
            parallel_for(nt, 65536, [&](size_t b0, size_t e0) {
                for (size_t i = b0; i < e0; ++i) {
                    begin[i] = (uint32_t)(i * k);
                    auto& t = m.tris[i];
                    for (size_t ti = 0; ti < k; ++ti) {
                        V3 p0 = m.verts[ti][t[0]], p1 = m.verts[ti][t[1]], p2 = m.verts[ti][t[2]];
                        BBox b;
                        b.min = vmin(p0, vmin(p1, p2));
                        b.max = vmax(p0, vmax(p1, p2));
                        bounds[i * k + ti] = b;
                    }
                }
            });
            begin[nt] = (uint32_t)(nt * k);
            if (timing)
                std::fprintf(stderr, "[build] mesh %zu triangles x %zu keys: triangle bounds %.3fs\n", nt, k,
                             std::chrono::duration<double>(std::chrono::steady_clock::now() - tm0).count());
        }
        if (!built) {
            // triangle order + binary topology from the registered external (GPU LBVH) builder
            MeshTopologyFn fn;
            void* fctx;
            uint32_t fmin;
            {
                std::lock_guard<std::mutex> g(g_topo_mu);
                fn = g_topo_fn, fctx = g_topo_ctx, fmin = g_topo_min_tris;
            }
            if (fn && nt >= fmin && nt > 3) {
                std::vector<int32_t> left(nt - 1), right(nt - 1);
                std::vector<uint8_t> axis(nt - 1);
                Bvh2Topology topo;
                auto t0 = std::chrono::steady_clock::now();
                int32_t rc = fn(fctx, fs.verts.data() + (size_t)pm.vert_base * 3, (uint32_t)vc, (uint32_t)k, &m.tris[0][0],
                                (uint32_t)nt, 3, ids.data(), left.data(), right.data(), axis.data(), &topo.n_internal, &topo.root,
                                &topo.depth);
                double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                if (rc == 0 && topo.depth <= BVH_MAX_DEPTH) {
                    topo.left = left.data(), topo.right = right.data(), topo.axis = axis.data();
                    bvh = build_bvh4_from_topology(ids.data(), nt, ObjBoundsView{bounds.data(), begin.data()}, 3, pm.tri_base,
                                                   topo);
                    built = true;
                    std::lock_guard<std::mutex> g(g_topo_mu);
                    g_topo_seconds += dt, g_topo_meshes += 1;
                } else {
                    std::iota(ids.begin(), ids.end(), 0u);  // fall back to the SAH builder on the original order
                }
            }
        }
        if (!built) bvh = build_bvh4(ids.data(), nt, ObjBoundsView{bounds.data(), begin.data()}, 3, pm.tri_base);
        auto tm1 = std::chrono::steady_clock::now();
        const size_t tri_off = fs.tris.size();
        fs.tris.resize(tri_off + nt * 4);
        parallel_for(nt, 65536, [&](size_t b0, size_t e0) {
            for (size_t i = b0; i < e0; ++i) {
                auto& t = m.tris[ids[i]];
                // (v0, v2, v1, original index) — triangle_mesh.rs:80
                uint32_t* o = fs.tris.data() + tri_off + i * 4;
                o[0] = t[0], o[1] = t[2], o[2] = t[1], o[3] = ids[i];
            }
        });
        if (in_place)
            pm.root = bvh.root;
        else
            append_bvh(bvh, pm.root);
        pm.bvh_depth = bvh.depth;
        pm.bounds_base = (uint32_t)fs.obj_bounds.size();
        pm.bounds_count = (uint32_t)bvh.bounds.size();
        for (auto& b : bvh.bounds) fs.obj_bounds.push_back(to_abi(b));
        if (timing)
            std::fprintf(stderr, "[build] triangle list + node append %.3fs\n",
                         std::chrono::duration<double>(std::chrono::steady_clock::now() - tm1).count());
        uint32_t idx = (uint32_t)fs.meshes.size();
        fs.meshes.push_back(pm);
        mesh_cache[&m] = idx;
        return idx;
    }

    uint32_t add_light(const LightIR& l) {
        auto it = light_cache.find(&l);
        if (it != light_cache.end()) return it->second;
        PsyLight pl{};
        pl.kind = l.kind;
        size_t per = l.kind == PSY_LIGHT_SPHERE ? 1 : 2;
        if (l.shape.empty() || l.shape.size() % per != 0 || l.colors.empty())
            throw std::runtime_error("light needs at least one shape key and one colour");
        pl.n_shape = (uint32_t)(l.shape.size() / per);
        pl.shape_base = (uint32_t)fs.light_shape.size();
        fs.light_shape.insert(fs.light_shape.end(), l.shape.begin(), l.shape.end());
        pl.n_colors = (uint32_t)l.colors.size();
        pl.color_base = (uint32_t)fs.light_colors.size();
        fs.light_colors.insert(fs.light_colors.end(), l.colors.begin(), l.colors.end());
        pl.bounds_base = (uint32_t)fs.obj_bounds.size();
        pl.bounds_count = pl.n_shape;
        for (uint32_t i = 0; i < pl.n_shape; ++i) {
            BBox b;
            if (l.kind == PSY_LIGHT_SPHERE) {
                float r = l.shape[i];
                b.min = v3(-r, -r, -r);
                b.max = v3(r, r, r);
            } else {
                float w = l.shape[2 * i], h = l.shape[2 * i + 1];
                b.min = v3(w * -0.5f, h * -0.5f, 0.0f);
                b.max = v3(w * 0.5f, h * 0.5f, 0.0f);
            }
            fs.obj_bounds.push_back(to_abi(b));
        }
        uint32_t idx = (uint32_t)fs.lights.size();
        fs.lights.push_back(pl);
        light_energy.push_back(colors_energy(l.colors));
        light_cache[&l] = idx;
        return idx;
    }

    std::vector<BBox> bounds_of(uint32_t base, uint32_t count) const {
        std::vector<BBox> v(count);
        for (uint32_t i = 0; i < count; ++i) {
            const PsyBBox& b = fs.obj_bounds[base + i];
            v[i].min = v3(b.min[0], b.min[1], b.min[2]);
            v[i].max = v3(b.max[0], b.max[1], b.max[2]);
        }
        return v;
    }

    uint32_t add_assembly(const AssemblyIR& a) {
        uint32_t shader_base = (uint32_t)fs.shaders.size();
        fs.shaders.insert(fs.shaders.end(), a.shaders.begin(), a.shaders.end());

        std::vector<uint32_t> obj_global(a.objects.size());
        for (size_t i = 0; i < a.objects.size(); ++i) {
            const ObjectIR& o = a.objects[i];
            PsyObject po{};
            po.kind = o.kind;
            po.index = o.kind == PSY_OBJECT_MESH ? add_mesh(*o.mesh) : add_light(*o.light);
            obj_global[i] = (uint32_t)fs.objects.size();
            fs.objects.push_back(po);
        }
        std::vector<uint32_t> asm_global(a.assemblies.size());
        for (size_t i = 0; i < a.assemblies.size(); ++i) asm_global[i] = add_assembly(*a.assemblies[i]);

        // instances + their transforms (add_instance, assembly.rs:218-281)
        size_t ni = a.instances.size();
        std::vector<PsyInstance> insts(ni);
        std::vector<Xf> local_xf;  // this assembly's transforms, in instance order
        uint32_t xform_base = (uint32_t)fs.xforms.size();
        for (size_t i = 0; i < ni; ++i) {
            const InstanceIR& in = a.instances[i];
            PsyInstance pi{};
            pi.kind = in.is_assembly ? PSY_INSTANCE_ASSEMBLY : PSY_INSTANCE_OBJECT;
            if (in.is_assembly) {
                if (in.data_index >= asm_global.size()) throw std::runtime_error("instance of unknown assembly");
                pi.data_index = asm_global[in.data_index];
            } else {
                if (in.data_index >= obj_global.size()) throw std::runtime_error("instance of unknown object");
                pi.data_index = obj_global[in.data_index];
            }
            int sh = in.shader_index;
            if (sh < 0 && !in.is_assembly) sh = a.objects[in.data_index].default_shader;
            if (sh >= (int)a.shaders.size()) throw std::runtime_error("unknown surface shader");
            pi.shader_index = sh < 0 ? PSY_NONE : shader_base + (uint32_t)sh;
            pi.id = (uint32_t)i;
            pi.xform_base = xform_base + (uint32_t)local_xf.size();
            pi.xform_count = (uint32_t)in.xforms.size();
            local_xf.insert(local_xf.end(), in.xforms.begin(), in.xforms.end());
            insts[i] = pi;
        }
        for (auto& x : local_xf) fs.xforms.push_back(to_abi(x));

        // instance_bounds (assembly.rs:354-395)
        std::vector<uint32_t> bis(ni + 1, 0);
        std::vector<BBox> bbs;
        for (size_t i = 0; i < ni; ++i) {
            const PsyInstance& pi = insts[i];
            std::vector<BBox> in_b, out_b;
            if (pi.kind == PSY_INSTANCE_OBJECT) {
                const PsyObject& po = fs.objects[pi.data_index];
                if (po.kind == PSY_OBJECT_MESH)
                    in_b = bounds_of(fs.meshes[po.index].bounds_base, fs.meshes[po.index].bounds_count);
                else
                    in_b = bounds_of(fs.lights[po.index].bounds_base, fs.lights[po.index].bounds_count);
            } else {
                const PsyAssembly& pa = fs.assemblies[pi.data_index];
                in_b = bounds_of(pa.bounds_base, pa.bounds_count);
            }
            if (pi.xform_count > 0) {
                const Xf* xf = local_xf.data() + (pi.xform_base - xform_base);
                transform_bbox_slice_from(in_b, xf, pi.xform_count, out_b);
            } else {
                out_b = in_b;
            }
            if (out_b.empty()) throw std::runtime_error("instance of an empty object/assembly is not supported");
            bbs.insert(bbs.end(), out_b.begin(), out_b.end());
            bis[i + 1] = (uint32_t)bbs.size();
        }

        PsyAssembly pa{};
        pa.instance_base = (uint32_t)fs.instances.size();
        pa.instance_count = (uint32_t)ni;
        if ((size_t)pa.instance_base + ni > PSY_MAX_LEAF_START) throw std::runtime_error("too many instances");
        std::vector<uint32_t> ids(ni);
        std::iota(ids.begin(), ids.end(), 0u);
        Bvh4Flat bvh = build_bvh4(ids.data(), ni, ObjBoundsView{bbs.data(), bis.data()}, 1, pa.instance_base);
        std::vector<PsyInstance> ordered(ni);
        for (size_t i = 0; i < ni; ++i) ordered[i] = insts[ids[i]];
        append_bvh(bvh, pa.root);
        pa.bvh_depth = bvh.depth;
        pa.bounds_base = (uint32_t)fs.obj_bounds.size();
        pa.bounds_count = (uint32_t)bvh.bounds.size();
        for (auto& b : bvh.bounds) fs.obj_bounds.push_back(to_abi(b));

        // light instances (assembly.rs:297-318)
        std::vector<PsyInstance> lights;
        std::vector<float> energies;
        for (auto& pi : ordered) {
            if (pi.kind == PSY_INSTANCE_OBJECT) {
                const PsyObject& po = fs.objects[pi.data_index];
                if (po.kind == PSY_OBJECT_LIGHT) {
                    lights.push_back(pi);
                    energies.push_back(light_energy[po.index]);
                }
            } else if (fs.assemblies[pi.data_index].light_energy > 0.0f) {
                lights.push_back(pi);
                energies.push_back(fs.assemblies[pi.data_index].light_energy);
            }
        }
        pa.light_instance_base = (uint32_t)fs.light_instances.size();
        pa.light_instance_count = (uint32_t)lights.size();
        pa.light_root = PSY_NONE;
        pa.light_energy = 0.0f;
        if (!lights.empty()) {
            // light tree over the light instances (light_tree.rs:62-131)
            size_t nl = lights.size();
            std::vector<uint32_t> lids(nl);
            std::iota(lids.begin(), lids.end(), 0u);
            std::vector<BBox> lb;
            std::vector<uint32_t> lbegin(nl + 1, 0);
            std::vector<BBox> tb(nl);
            for (size_t i = 0; i < nl; ++i) {
                uint32_t id = lights[i].id;
                lb.insert(lb.end(), bbs.begin() + bis[id], bbs.begin() + bis[id + 1]);
                lbegin[i + 1] = (uint32_t)lb.size();
            }
            for (size_t i = 0; i < nl; ++i) tb[i] = lerp_slice(lb.data() + lbegin[i], lbegin[i + 1] - lbegin[i], 0.5f);
            LightTreeBuilder ltb;
            ltb.tb = &tb;
            ltb.ob = ObjBoundsView{lb.data(), lbegin.data()};
            ltb.energy = &energies;
            ltb.build(0, lids.data(), nl);
            for (size_t i = 0; i < nl; ++i) fs.light_instances.push_back(lights[lids[i]]);

            // construct_from_builder with 3-level collapse (light_tree.rs:89-131)
            pa.light_root = (uint32_t)fs.light_nodes.size();
            fs.light_nodes.emplace_back();
            emit_light_node(ltb, 0, pa.light_root);
            pa.light_energy = ltb.nodes[0].energy;
        }
        uint32_t idx = (uint32_t)fs.assemblies.size();
        fs.instances.insert(fs.instances.end(), ordered.begin(), ordered.end());
        fs.assemblies.push_back(pa);
        return idx;
    }

    void emit_light_node(const LightTreeBuilder& b, uint32_t bidx, uint32_t slot) {
        const LtbNode& n = b.nodes[bidx];
        PsyLightNode ln{};
        ln.bounds_base = (uint32_t)fs.lt_bounds.size();
        ln.bounds_count = n.b1 - n.b0;
        for (uint32_t i = n.b0; i < n.b1; ++i) fs.lt_bounds.push_back(to_abi(b.bounds[i]));
        ln.energy = n.energy;
        if (n.is_leaf) {
            ln.child_base = 0;
            ln.child_count = 0;
            ln.light_index = n.child_index;
            fs.light_nodes[slot] = ln;
            return;
        }
        size_t cc = b.child_count(3, bidx);
        ln.child_base = (uint32_t)fs.light_nodes.size();
        ln.child_count = (uint32_t)cc;
        ln.light_index = PSY_NONE;
        fs.light_nodes.resize(fs.light_nodes.size() + cc);
        fs.light_nodes[slot] = ln;
        for (size_t i = 0; i < cc; ++i) emit_light_node(b, b.nth_child(3, bidx, i).first, ln.child_base + (uint32_t)i);
    }
};

}  // namespace

void FlatScene::finalize() {
    desc.abi_version = PSY_ABI_VERSION;
    desc.camera.xforms = cam_xforms.data();
    desc.camera.n_xforms = (uint32_t)cam_xforms.size();
    desc.camera.tfovs = cam_tfovs.data();
    desc.camera.n_tfovs = (uint32_t)cam_tfovs.size();
    desc.camera.aperture_radii = cam_aperture.data();
    desc.camera.n_aperture_radii = (uint32_t)cam_aperture.size();
    desc.camera.focus_distances = cam_focus.data();
    desc.camera.n_focus_distances = (uint32_t)cam_focus.size();
    desc.distant_lights = distant_lights.data();
    desc.n_distant_lights = (uint32_t)distant_lights.size();
    desc.world_floats = world_floats.data();
    desc.n_world_floats = (uint32_t)world_floats.size();
    desc.nodes = nodes.data();
    desc.n_nodes = (uint32_t)nodes.size();
    desc.motion_bounds = motion_bounds.data();
    desc.n_motion_bounds = (uint32_t)motion_bounds.size();
    desc.meshes = meshes.data();
    desc.n_meshes = (uint32_t)meshes.size();
    desc.tris = tris.data();
    desc.n_tris = (uint32_t)(tris.size() / 4);
    desc.verts = verts.data();
    desc.normals = normals.data();
    desc.n_vert_keys = (uint32_t)(verts.size() / 3);
    desc.lights = lights.data();
    desc.n_lights = (uint32_t)lights.size();
    desc.light_shape = light_shape.data();
    desc.n_light_shape = (uint32_t)light_shape.size();
    desc.light_colors = light_colors.data();
    desc.n_light_colors = (uint32_t)light_colors.size();
    desc.objects = objects.data();
    desc.n_objects = (uint32_t)objects.size();
    desc.shaders = shaders.data();
    desc.n_shaders = (uint32_t)shaders.size();
    desc.assemblies = assemblies.data();
    desc.n_assemblies = (uint32_t)assemblies.size();
    desc.instances = instances.data();
    desc.n_instances = (uint32_t)instances.size();
    desc.light_instances = light_instances.data();
    desc.n_light_instances = (uint32_t)light_instances.size();
    desc.xforms = xforms.data();
    desc.n_xforms = (uint32_t)xforms.size();
    desc.light_nodes = light_nodes.data();
    desc.n_light_nodes = (uint32_t)light_nodes.size();
    desc.lt_bounds = lt_bounds.data();
    desc.n_lt_bounds = (uint32_t)lt_bounds.size();
    desc.obj_bounds = obj_bounds.data();
    desc.n_obj_bounds = (uint32_t)obj_bounds.size();
}

std::unique_ptr<FlatScene> flatten_scene(const SceneIR& s) {
    auto t0 = std::chrono::steady_clock::now();
    std::unique_ptr<FlatScene> fs(new FlatScene());
    fs->name = s.name;
    fs->output_path = s.output_path;
    fs->desc.res_x = s.res_x;
    fs->desc.res_y = s.res_y;
    fs->desc.spp = s.spp;
    fs->desc.seed = s.seed;

    // Camera::new (camera.rs:22-72)
    if (s.cam_xforms.empty()) throw std::runtime_error("Camera has no transform(s)!");
    if (s.cam_fovs.empty()) throw std::runtime_error("Camera has no fov(s)!");
    std::vector<float> ap = s.cam_aperture, fo = s.cam_focus;
    if (ap.empty() || fo.empty()) {
        ap = {0.0f};
        fo = {1.0f};
    }
    bool zero_focus = false;
    for (float d : fo) zero_focus |= (d == 0.0f);
    if (zero_focus) {
        ap = {0.0f};
        fo = {1.0f};
    }
    for (auto& x : s.cam_xforms) fs->cam_xforms.push_back(to_abi(x));
    for (float n : s.cam_fovs) fs->cam_tfovs.push_back(std::sin(n / 2.0f) / std::cos(n / 2.0f));
    fs->cam_aperture = ap;
    fs->cam_focus = fo;

    fs->desc.background = s.background;
    for (auto& dl : s.distant_lights) {
        if (dl.radii.empty() || dl.dirs.empty() || dl.colors.empty())
            throw std::runtime_error("DistantDiskLight needs Radius, Direction and Color");
        PsyDistantLight p{};
        p.n_radii = (uint32_t)dl.radii.size();
        p.radii_base = (uint32_t)fs->world_floats.size();
        fs->world_floats.insert(fs->world_floats.end(), dl.radii.begin(), dl.radii.end());
        p.n_dirs = (uint32_t)dl.dirs.size();
        p.dir_base = (uint32_t)fs->world_floats.size();
        for (auto& d : dl.dirs) {
            fs->world_floats.push_back(d.x);
            fs->world_floats.push_back(d.y);
            fs->world_floats.push_back(d.z);
        }
        p.n_colors = (uint32_t)dl.colors.size();
        p.color_base = (uint32_t)fs->light_colors.size();
        fs->light_colors.insert(fs->light_colors.end(), dl.colors.begin(), dl.colors.end());
        fs->distant_lights.push_back(p);
    }

    if (!s.root) throw std::runtime_error("scene has no root assembly");
    Flattener fl(*fs);
    fs->desc.root_assembly = fl.add_assembly(*s.root);
    fs->finalize();
    fs->build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return fs;
}

}  // namespace psyh
