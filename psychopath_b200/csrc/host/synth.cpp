// synth.cpp — seeded synthetic scenes of the shapes BASELINE.json names:
//   C3  displaced grid, deformation + transform motion blur
//   C4/C5  two-level instancing: clusters (sub-assemblies) of instanced assets + area lights,
//          nested BVH4s and nested light trees, optionally moving cluster instances.
// Geometry goes straight into the SceneIR (no .psy text for 10 M triangles).
#include <cmath>

#include "scene_ir.h"

namespace psyh {
namespace {

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 0x1234567ull) {}
    uint64_t next() {  // splitmix64
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    float uni() { return (float)(next() >> 40) * (1.0f / 16777216.0f); }
    float range(float a, float b) { return a + (b - a) * uni(); }
};

inline uint32_t hash2(int x, int y, uint32_t seed) {
    uint32_t h = (uint32_t)x * 0x8da6b343u ^ (uint32_t)y * 0xd8163841u ^ seed * 0xcb1ab31fu;
    h ^= h >> 15;
    h *= 0x2c1b3c6du;
    h ^= h >> 12;
    h *= 0x297a2d39u;
    h ^= h >> 15;
    return h;
}
inline float lattice(int x, int y, uint32_t seed) { return (float)(hash2(x, y, seed) >> 8) * (1.0f / 16777216.0f); }
float value_noise(float x, float y, uint32_t seed) {
    float fx = std::floor(x), fy = std::floor(y);
    int ix = (int)fx, iy = (int)fy;
    float tx = x - fx, ty = y - fy;
    tx = tx * tx * (3.0f - 2.0f * tx);
    ty = ty * ty * (3.0f - 2.0f * ty);
    float a = lattice(ix, iy, seed), b = lattice(ix + 1, iy, seed);
    float c = lattice(ix, iy + 1, seed), d = lattice(ix + 1, iy + 1, seed);
    return (a * (1 - tx) + b * tx) * (1 - ty) + (c * (1 - tx) + d * tx) * ty;
}
float fbm(float x, float y, uint32_t seed, int octaves) {
    float amp = 0.5f, sum = 0.0f, f = 1.0f;
    for (int o = 0; o < octaves; ++o) {
        sum += amp * value_noise(x * f, y * f, seed + (uint32_t)o * 101u);
        amp *= 0.5f;
        f *= 2.0f;
    }
    return sum;
}

Xf rot_z(float a) {
    Xf m;
    float c = std::cos(a), s = std::sin(a);
    m.X = v3(c, s, 0);
    m.Y = v3(-s, c, 0);
    return m;
}
Xf rot_x(float a) {
    Xf m;
    float c = std::cos(a), s = std::sin(a);
    m.Y = v3(0, c, s);
    m.Z = v3(0, -s, c);
    return m;
}
Xf translate(V3 t) {
    Xf m;
    m.T = t;
    return m;
}
Xf scale(float s) {
    Xf m;
    m.X = v3(s, 0, 0);
    m.Y = v3(0, s, 0);
    m.Z = v3(0, 0, s);
    return m;
}

// camera->world: +x right, +y up, +z forward (src/camera.rs:88-102)
Xf look_at(V3 eye, V3 target, V3 up) {
    V3 f = normalized(target - eye);
    V3 r = normalized(cross(f, up));
    V3 u = cross(r, f);
    Xf m;
    m.X = r;
    m.Y = u;
    m.Z = f;
    m.T = eye;
    return m;
}

PsyShader lambert(float r, float g, float b) {
    PsyShader s{};
    s.kind = PSY_SHADER_LAMBERT;
    s.color = color_from_rec709(r, g, b);
    return s;
}
PsyShader ggx(float r, float g, float b, float rough, float fresnel) {
    PsyShader s{};
    s.kind = PSY_SHADER_GGX;
    s.color = color_from_rec709(r, g, b);
    s.roughness = rough;
    s.fresnel = fresnel;
    return s;
}

void add_grid_tris(MeshIR& m, uint32_t q) {
    uint32_t stride = q + 1;
    m.tris.reserve((size_t)q * q * 2);
    for (uint32_t j = 0; j < q; ++j)
        for (uint32_t i = 0; i < q; ++i) {
            uint32_t a = j * stride + i, b = a + 1, c = a + stride, d = c + 1;
            m.tris.push_back({a, b, d});
            m.tris.push_back({a, d, c});
        }
}

// Asset: displaced sphere (kind 0) or torus (kind 1), q x q quads.
std::shared_ptr<MeshIR> make_asset(uint32_t q, int kind, uint32_t seed) {
    auto m = std::make_shared<MeshIR>();
    m->verts.resize(1);
    auto& v = m->verts[0];
    v.resize((size_t)(q + 1) * (q + 1));
    const float PI = 3.14159265358979f;
    for (uint32_t j = 0; j <= q; ++j)
        for (uint32_t i = 0; i <= q; ++i) {
            float u = (float)i / (float)q, w = (float)j / (float)q;
            float n = fbm(u * 12.0f, w * 12.0f, seed, 4);
            V3 p;
            if ((kind & 1) == 0) {
                float th = w * PI, ph = u * 2.0f * PI;
                float r = 1.0f + 0.15f * (n - 0.5f);
                p = v3(r * std::sin(th) * std::cos(ph), r * std::sin(th) * std::sin(ph), r * std::cos(th));
            } else {
                float a = u * 2.0f * PI, b = w * 2.0f * PI;
                float r = 0.35f + 0.08f * (n - 0.5f);
                p = v3((1.0f + r * std::cos(b)) * std::cos(a), (1.0f + r * std::cos(b)) * std::sin(a), r * std::sin(b));
            }
            v[(size_t)j * (q + 1) + i] = p;
        }
    add_grid_tris(*m, q);
    return m;
}

}  // namespace

SceneIR synth_displaced_grid(uint32_t q, uint32_t n_vert_keys, uint32_t n_xform_keys, uint32_t res_x,
                             uint32_t res_y, uint32_t spp, uint32_t seed) {
    SceneIR sc;
    sc.name = "$synth_displaced_grid";
    sc.output_path = "synth_displaced_grid.png";
    sc.res_x = res_x;
    sc.res_y = res_y;
    sc.spp = spp;
    sc.seed = seed;
    if (n_vert_keys < 1) n_vert_keys = 1;
    if (n_xform_keys < 1) n_xform_keys = 1;

    const float extent = 20.0f;
    uint32_t stride = q + 1;
    // base heights + deformation direction
    std::vector<V3> base((size_t)stride * stride), delta((size_t)stride * stride);
    for (uint32_t j = 0; j <= q; ++j)
        for (uint32_t i = 0; i <= q; ++i) {
            float x = ((float)i / (float)q - 0.5f) * extent, y = ((float)j / (float)q - 0.5f) * extent;
            float h = 3.0f * fbm(x * 0.35f + 17.0f, y * 0.35f + 5.0f, 0xC3u, 5);
            base[(size_t)j * stride + i] = v3(x, y, h);
            float w = 0.02f * extent;
            delta[(size_t)j * stride + i] =
                v3(w * 0.3f * std::sin(y * 0.7f), w * 0.3f * std::cos(x * 0.9f), w * std::sin(x * 0.5f + y * 0.3f));
        }
    // two meshes: checker of 16x16-quad blocks, Lambert vs GGX
    std::shared_ptr<MeshIR> mesh[2] = {std::make_shared<MeshIR>(), std::make_shared<MeshIR>()};
    for (int mi = 0; mi < 2; ++mi) {
        mesh[mi]->verts.resize(n_vert_keys);
        for (uint32_t k = 0; k < n_vert_keys; ++k) {
            float f = n_vert_keys > 1 ? (float)k / (float)(n_vert_keys - 1) : 0.0f;
            auto& v = mesh[mi]->verts[k];
            v.resize(base.size());
            for (size_t i = 0; i < base.size(); ++i) v[i] = base[i] + delta[i] * f;
        }
    }
    for (uint32_t j = 0; j < q; ++j)
        for (uint32_t i = 0; i < q; ++i) {
            uint32_t a = j * stride + i, b = a + 1, c = a + stride, d = c + 1;
            int mi = ((i / 16) + (j / 16)) & 1;
            mesh[mi]->tris.push_back({a, b, d});
            mesh[mi]->tris.push_back({a, d, c});
        }

    auto root = std::make_shared<AssemblyIR>();
    root->shaders.push_back(lambert(0.7f, 0.65f, 0.55f));
    root->shaders.push_back(ggx(0.9f, 0.85f, 0.8f, 0.3f, 0.5f));
    for (int mi = 0; mi < 2; ++mi) {
        if (mesh[mi]->tris.empty()) continue;
        ObjectIR o;
        o.kind = PSY_OBJECT_MESH;
        o.mesh = mesh[mi];
        root->objects.push_back(o);
        InstanceIR in;
        in.data_index = (uint32_t)root->objects.size() - 1;
        in.shader_index = mi;
        for (uint32_t k = 0; k < n_xform_keys; ++k) {
            float f = n_xform_keys > 1 ? (float)k / (float)(n_xform_keys - 1) : 0.0f;
            // local->world: rotate about z by 0..6 degrees and translate 0..0.5; stored world->local
            Xf l2w = affine_mul(translate(v3(0.5f * f, 0.0f, 0.0f)), rot_z(f * 6.0f * 3.14159265f / 180.0f));
            in.xforms.push_back(xf_inverse(l2w));
        }
        root->instances.push_back(in);
    }
    // sphere light
    {
        auto l = std::make_shared<LightIR>();
        l->kind = PSY_LIGHT_SPHERE;
        l->shape = {1.5f};
        l->colors = {color_from_rec709(400.0f, 380.0f, 340.0f)};
        ObjectIR o;
        o.kind = PSY_OBJECT_LIGHT;
        o.light = l;
        root->objects.push_back(o);
        InstanceIR in;
        in.data_index = (uint32_t)root->objects.size() - 1;
        in.xforms.push_back(xf_inverse(translate(v3(-4.0f, 3.0f, 9.0f))));
        root->instances.push_back(in);
    }
    DistantLightIR dl;
    dl.radii = {0.05f};
    dl.dirs = {v3(-0.4f, 0.3f, -0.85f)};
    dl.colors = {color_from_rec709(2.0f, 1.9f, 1.7f)};
    sc.distant_lights.push_back(dl);
    sc.background = color_from_rec709(0.08f, 0.1f, 0.15f);
    sc.root = root;

    // camera 35 degrees above the horizon
    float el = 35.0f * 3.14159265f / 180.0f, dist = 24.0f;
    V3 eye = v3(0.0f, -dist * std::cos(el), dist * std::sin(el) + 1.5f);
    sc.cam_xforms.push_back(look_at(eye, v3(0, 0, 1.5f), v3(0, 0, 1)));
    sc.cam_fovs = {45.0f * 3.14159265f / 180.0f};
    return sc;
}

SceneIR synth_instanced(uint32_t asset_quads, uint32_t n_assets, uint32_t n_clusters, uint32_t inst_per_cluster,
                        uint32_t lights_per_cluster, uint32_t cluster_xform_keys, uint32_t res_x, uint32_t res_y,
                        uint32_t spp, uint32_t seed, uint32_t motion_pct) {
    SceneIR sc;
    sc.name = "$synth_instanced";
    sc.output_path = "synth_instanced.png";
    sc.res_x = res_x;
    sc.res_y = res_y;
    sc.spp = spp;
    sc.seed = seed;
    if (n_assets < 1) n_assets = 1;
    if (cluster_xform_keys < 1) cluster_xform_keys = 1;
    Rng rng(0xC4C4ull + seed * 0);  // geometry does not depend on the render seed

    std::vector<std::shared_ptr<MeshIR>> assets;
    for (uint32_t a = 0; a < n_assets; ++a) assets.push_back(make_asset(asset_quads, (int)a, 0xA55E7u + a));

    auto root = std::make_shared<AssemblyIR>();
    // ground plane
    {
        auto g = std::make_shared<MeshIR>();
        g->verts.resize(1);
        float e = 90.0f;
        g->verts[0] = {v3(-e, -e, 0), v3(e, -e, 0), v3(-e, e, 0), v3(e, e, 0)};
        g->tris = {{0, 1, 3}, {0, 3, 2}};
        root->shaders.push_back(lambert(0.5f, 0.5f, 0.5f));
        ObjectIR o;
        o.kind = PSY_OBJECT_MESH;
        o.mesh = g;
        root->objects.push_back(o);
        InstanceIR in;
        in.data_index = 0;
        in.shader_index = 0;
        root->instances.push_back(in);
    }

    uint32_t side = (uint32_t)std::ceil(std::sqrt((double)n_clusters));
    const float spacing = 12.0f;
    for (uint32_t c = 0; c < n_clusters; ++c) {
        auto cl = std::make_shared<AssemblyIR>();
        cl->shaders.push_back(lambert(rng.range(0.2f, 0.9f), rng.range(0.2f, 0.9f), rng.range(0.2f, 0.9f)));
        cl->shaders.push_back(ggx(rng.range(0.5f, 0.95f), rng.range(0.5f, 0.95f), rng.range(0.5f, 0.95f),
                                  rng.range(0.1f, 0.6f), rng.range(0.2f, 0.9f)));
        for (uint32_t a = 0; a < n_assets; ++a) {
            ObjectIR o;
            o.kind = PSY_OBJECT_MESH;
            o.mesh = assets[a];
            cl->objects.push_back(o);
        }
        for (uint32_t i = 0; i < inst_per_cluster; ++i) {
            InstanceIR in;
            in.data_index = (uint32_t)(rng.next() % n_assets);
            in.shader_index = (int)(rng.next() & 1);
            float s = rng.range(0.6f, 1.4f);
            V3 pos = v3(rng.range(-4.0f, 4.0f), rng.range(-4.0f, 4.0f), s * 1.1f + rng.range(0.0f, 2.0f));
            Xf l2w = affine_mul(translate(pos),
                                affine_mul(rot_z(rng.range(0.0f, 6.28f)), affine_mul(rot_x(rng.range(0.0f, 6.28f)), scale(s))));
            in.xforms.push_back(xf_inverse(l2w));
            cl->instances.push_back(in);
        }
        for (uint32_t i = 0; i < lights_per_cluster; ++i) {
            auto l = std::make_shared<LightIR>();
            float energy = std::exp(rng.range(0.0f, 4.60517f));  // log-uniform 1..100
            float tint = rng.range(0.7f, 1.0f);
            if (i & 1) {
                l->kind = PSY_LIGHT_SPHERE;
                l->shape = {rng.range(0.08f, 0.3f)};
            } else {
                l->kind = PSY_LIGHT_RECT;
                l->shape = {rng.range(0.3f, 1.0f), rng.range(0.3f, 1.0f)};
            }
            l->colors = {color_from_rec709(energy, energy * tint, energy * tint * tint)};
            ObjectIR o;
            o.kind = PSY_OBJECT_LIGHT;
            o.light = l;
            cl->objects.push_back(o);
            InstanceIR in;
            in.data_index = (uint32_t)cl->objects.size() - 1;
            V3 pos = v3(rng.range(-4.5f, 4.5f), rng.range(-4.5f, 4.5f), rng.range(3.0f, 6.0f));
            // rectangle lights face down (their normal is +z in local space; flip with a rotation about x)
            Xf l2w = affine_mul(translate(pos), rot_x(3.14159265f + rng.range(-0.4f, 0.4f)));
            in.xforms.push_back(xf_inverse(l2w));
            cl->instances.push_back(in);
        }
        root->assemblies.push_back(cl);
        InstanceIR in;
        in.is_assembly = true;
        in.data_index = c;
        float cx = ((float)(c % side) - 0.5f * (float)(side - 1)) * spacing + rng.range(-2.0f, 2.0f);
        float cy = ((float)(c / side) - 0.5f * (float)(side - 1)) * spacing + rng.range(-2.0f, 2.0f);
        // motion over the shutter: drift up to 0.4 units and up to 0.15 rad of rotation at motion_pct = 100.  That
        // much collapses under the builder's 1.4x union rule (bvh_base.rs:119-132): every node of the root BVH gets
        // one merged box.  BASELINE config 4 ("motion-blurred instanced scene": 2-key bounds on the root assembly's
        // nodes) uses motion_pct = 1250: drift up to 5 units, rotation up to 0.6 rad.
        const float mscale = (float)motion_pct / 100.0f;
        float ang0 = rng.range(0.0f, 6.28f), dang = rng.range(-0.15f, 0.15f) * std::min(mscale, 4.0f);
        V3 drift = v3(rng.range(-0.4f, 0.4f), rng.range(-0.4f, 0.4f), 0.0f) * mscale;
        for (uint32_t k = 0; k < cluster_xform_keys; ++k) {
            float f = cluster_xform_keys > 1 ? (float)k / (float)(cluster_xform_keys - 1) : 0.0f;
            Xf l2w = affine_mul(translate(v3(cx, cy, 0.0f) + drift * f), rot_z(ang0 + dang * f));
            in.xforms.push_back(xf_inverse(l2w));
        }
        root->instances.push_back(in);
    }
    sc.root = root;
    sc.background = color_from_rec709(0.02f, 0.025f, 0.04f);
    float field = 0.5f * (float)side * spacing;
    V3 eye = v3(-0.2f * field, -1.55f * field - 8.0f, 0.55f * field + 6.0f);
    sc.cam_xforms.push_back(look_at(eye, v3(0, -0.1f * field, 1.0f), v3(0, 0, 1)));
    sc.cam_fovs = {50.0f * 3.14159265f / 180.0f};
    return sc;
}

}  // namespace psyh
