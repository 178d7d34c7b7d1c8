// psy_loader.cpp — `.psy` text -> SceneIR.
//
// Token/tree level follows src/parse/data_tree.rs:296-566 (`Type $ident { ... }` / `Type [leaf]`,
// `#` comments, `\` escapes, `$` kept as part of the identifier).  Scene level follows
// src/parse/psy.rs:95-615, psy_assembly.rs:17-188, psy_mesh_surface.rs:27-127, psy_light.rs,
// psy_surface_shader.rs:24-139, and additionally accepts the newer layout the shipped
// example_scenes use (`Shaders{}` / `Objects{}` at Scene level, `SurfaceShaderBind` on the object;
// SURVEY.md F2 / §11).
#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>

#include "scene_ir.h"

namespace psyh {
namespace {

struct Node {
    std::string type;
    std::string ident;  // includes the leading '$' when present
    bool has_ident = false;
    bool is_leaf = false;
    std::string contents;  // leaf text
    std::vector<Node> children;
    size_t offset = 0;
};

struct Lexer {
    const std::string& s;
    size_t p = 0;
    explicit Lexer(const std::string& t) : s(t) {}
    static bool is_ws(char c) { return c == '\n' || c == '\r' || c == '\t' || c == ' '; }
    static bool is_reserved(char c) {
        return c == '{' || c == '}' || c == '[' || c == ']' || c == '$' || c == '#' || c == '\\';
    }
    static bool is_ident_char(char c) { return !is_ws(c) && !is_reserved(c); }
    void skip_ws_comments() {
        for (;;) {
            size_t q = p;
            while (p < s.size() && is_ws(s[p])) ++p;
            if (p < s.size() && s[p] == '#')
                while (p < s.size() && s[p] != '\n' && s[p] != '\r') ++p;
            if (p == q) break;
        }
    }
    [[noreturn]] void fail(const char* what, size_t at) const {
        throw std::runtime_error(std::string("psy parse error: ") + what + " at byte " + std::to_string(at));
    }

    // parse_node (data_tree.rs:296-378); returns false when no node starts here.
    bool parse_node(Node& out) {
        skip_ws_comments();
        if (p >= s.size() || !is_ident_char(s[p])) return false;
        size_t start = p;
        while (p < s.size() && is_ident_char(s[p])) ++p;
        out.type = s.substr(start, p - start);
        out.offset = p;
        skip_ws_comments();
        if (p < s.size() && s[p] == '$') {
            size_t is = p++;
            bool esc = false;
            while (p < s.size()) {
                char c = s[p];
                if (esc)
                    esc = false;
                else if (c == '\\')
                    esc = true;
                else if (!is_ident_char(c))
                    break;
                ++p;
            }
            out.ident = s.substr(is, p - is);
            out.has_ident = true;
            skip_ws_comments();
            if (p >= s.size() || s[p] != '{') fail("expected '{' after identifier", p);
        }
        if (p < s.size() && s[p] == '{') {
            ++p;
            out.is_leaf = false;
            for (;;) {
                Node child;
                if (!parse_node(child)) break;
                out.children.push_back(std::move(child));
            }
            skip_ws_comments();
            if (p >= s.size() || s[p] != '}') fail("missing '}'", p);
            ++p;
            return true;
        }
        if (p < s.size() && s[p] == '[') {
            ++p;
            size_t cs = p;
            bool esc = false;
            while (p < s.size()) {
                char c = s[p];
                if (esc)
                    esc = false;
                else if (c == '\\')
                    esc = true;
                else if (c == ']')
                    break;
                ++p;
            }
            if (p >= s.size()) fail("missing ']'", cs);
            out.is_leaf = true;
            out.contents = s.substr(cs, p - cs);
            ++p;
            return true;
        }
        fail("expected '{' or '[' after type name", p);
    }
};

std::string trim(const std::string& s) {
    size_t a = 0, b = s.size();
    while (a < b && isspace((unsigned char)s[a])) ++a;
    while (b > a && isspace((unsigned char)s[b - 1])) --b;
    return s.substr(a, b - a);
}

// Sequence of whitespace-separated floats (basics.rs:16-18 ws_f32, repeated); stops at the first non-number.
// `rest`: where it stopped (trailing whitespace consumed, like ws_f32's multispace0).
std::vector<float> parse_floats(const std::string& text, size_t* rest = nullptr) {
    std::vector<float> out;
    const char* c = text.c_str();
    char* end = nullptr;
    for (;;) {
        while (*c && isspace((unsigned char)*c)) ++c;
        if (!*c) break;
        // nom 5's recognize_float: [sign] (digits ['.' [digits]] | '.' digits) [('e'|'E') [sign] digits] — no hex, no
        // "inf" / "nan"; the text it recognises goes to the correctly rounded conversion (Rust's f32::from_str / strtof)
        const char* q = (*c == '-' || *c == '+') ? c + 1 : c;
        const char* ds = q;
        while (isdigit((unsigned char)*q)) ++q;
        if (q > ds) {
            if (*q == '.') {
                ++q;
                while (isdigit((unsigned char)*q)) ++q;
            }
        } else if (*q == '.' && isdigit((unsigned char)q[1])) {
            ++q;
            while (isdigit((unsigned char)*q)) ++q;
        } else {
            break;
        }
        if (*q == 'e' || *q == 'E') {
            const char* e = q + 1;
            if (*e == '+' || *e == '-') ++e;
            if (isdigit((unsigned char)*e)) {
                while (isdigit((unsigned char)*e)) ++e;
                q = e;
            }
        }
        const std::string tok(c, q);
        out.push_back(strtof(tok.c_str(), &end));
        c = q;
    }
    if (rest) *rest = (size_t)(c - text.c_str());
    return out;
}

// Fields the reference reads with all_consuming(...) (psy.rs:277-407, psy_light.rs, psy_surface_shader.rs:80-96):
// exactly `count` numbers and nothing else in the leaf.
std::vector<float> parse_floats_exact(const Node& n, size_t count, const char* what) {
    size_t rest = 0;
    std::vector<float> v = parse_floats(n.contents, &rest);
    if (v.size() != count || rest != n.contents.size())
        throw std::runtime_error(std::string("psy parse error: ") + what + " at byte " + std::to_string(n.offset));
    return v;
}

// ws_u32 / ws_usize (basics.rs:20-26), repeated: digits only
std::vector<uint32_t> parse_uints(const std::string& text, size_t* rest = nullptr) {
    std::vector<uint32_t> out;
    const char* c = text.c_str();
    for (;;) {
        while (*c && isspace((unsigned char)*c)) ++c;
        if (!*c || !isdigit((unsigned char)*c)) break;
        char* end = nullptr;
        unsigned long v = strtoul(c, &end, 10);
        out.push_back((uint32_t)v);
        c = end;
    }
    if (rest) *rest = (size_t)(c - text.c_str());
    return out;
}
std::vector<uint32_t> parse_uints_exact(const std::string& text, size_t count, const char* what) {
    size_t rest = 0;
    std::vector<uint32_t> v = parse_uints(text, &rest);
    if (v.size() != count || rest != text.size()) throw std::runtime_error(what);
    return v;
}

// parse_matrix (psy.rs:556-572): 16 numbers, column-major, 4th row dropped.
Xf parse_matrix(const Node& n) {
    std::vector<float> v = parse_floats_exact(n, 16, "Transform should be sixteen numbers");
    Xf m;
    m.X = v3(v[0], v[1], v[2]);
    m.Y = v3(v[4], v[5], v[6]);
    m.Z = v3(v[8], v[9], v[10]);
    m.T = v3(v[12], v[13], v[14]);
    return m;
}

// parse_color (psy.rs:582-615)
PsyColor parse_color(const Node& n) {
    size_t comma = n.contents.find(',');
    if (comma == std::string::npos || n.contents.find(',', comma + 1) != std::string::npos)
        throw std::runtime_error("psy parse error: bad Color at byte " + std::to_string(n.offset));
    std::string kind = trim(n.contents.substr(0, comma));
    std::vector<float> v = parse_floats(n.contents.substr(comma + 1));
    if (kind == "rec709" && v.size() >= 3) return color_from_rec709(v[0], v[1], v[2]);
    if (kind == "blackbody" && v.size() >= 2) return PsyColor{PSY_COLOR_BLACKBODY, v[0], v[1], 0.0f};
    if (kind == "color_temperature" && v.size() >= 2) return PsyColor{PSY_COLOR_TEMPERATURE, v[0], v[1], 0.0f};
    throw std::runtime_error("psy parse error: bad Color at byte " + std::to_string(n.offset));
}

const Node* first_child(const Node& n, const char* type) {
    for (auto& c : n.children)
        if (c.type == type) return &c;
    return nullptr;
}
size_t count_children(const Node& n, const char* type) {
    size_t k = 0;
    for (auto& c : n.children) k += (c.type == type);
    return k;
}
const Node* first_leaf(const Node& n, const char* type) {
    for (auto& c : n.children)
        if (c.is_leaf && c.type == type) return &c;
    return nullptr;
}

PsyShader parse_surface_shader(const Node& n) {
    const Node* t = first_leaf(n, "Type");
    if (!t) throw std::runtime_error("Expected a Type field in SurfaceShader.");
    std::string type = trim(t->contents);
    const Node* col = first_leaf(n, "Color");
    if (!col) throw std::runtime_error("Expected a Color field in SurfaceShader.");
    PsyShader s{};
    s.color = parse_color(*col);
    if (type == "Lambert") {
        s.kind = PSY_SHADER_LAMBERT;
    } else if (type == "Emit") {
        s.kind = PSY_SHADER_EMIT;
    } else if (type == "GGX") {
        s.kind = PSY_SHADER_GGX;
        const Node* r = first_leaf(n, "Roughness");
        const Node* f = first_leaf(n, "Fresnel");
        if (!r) throw std::runtime_error("Expected a Roughness field in GTR SurfaceShader.");
        if (!f) throw std::runtime_error("Expected a Fresnel field in GTR SurfaceShader.");
        s.roughness = parse_floats_exact(*r, 1, "bad Roughness")[0];
        s.fresnel = parse_floats_exact(*f, 1, "bad Fresnel")[0];
    } else {
        throw std::runtime_error("unknown SurfaceShader Type '" + type + "'");
    }
    return s;
}

std::shared_ptr<MeshIR> parse_mesh(const Node& n) {
    auto m = std::make_shared<MeshIR>();
    for (auto& c : n.children) {
        if (!c.is_leaf) continue;
        if (c.type == "Vertices") {
            std::vector<float> v = parse_floats(c.contents);
            std::vector<V3> pts(v.size() / 3);
            for (size_t i = 0; i < pts.size(); ++i) pts[i] = v3(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
            m->verts.push_back(std::move(pts));
        } else if (c.type == "Normals") {
            std::vector<float> v = parse_floats(c.contents);
            std::vector<V3> ns(v.size() / 3);
            for (size_t i = 0; i < ns.size(); ++i) ns[i] = normalized(v3(v[3 * i], v[3 * i + 1], v[3 * i + 2]));
            m->normals.push_back(std::move(ns));
        }
    }
    if (m->verts.empty()) throw std::runtime_error("MeshSurface without Vertices");
    std::vector<uint32_t> counts, indices;
    if (const Node* c = first_leaf(n, "FaceVertCounts")) counts = parse_uints(c->contents);
    if (const Node* c = first_leaf(n, "FaceVertIndices")) indices = parse_uints(c->contents);
    size_t ii = 0;
    for (uint32_t fvc : counts) {
        if (fvc < 3) throw std::runtime_error("Cannot handle polygons with less than three vertices.");
        if (ii + fvc > indices.size()) throw std::runtime_error("FaceVertIndices too short");
        for (uint32_t vi = 0; vi + 2 < fvc; ++vi)
            m->tris.push_back({indices[ii], indices[ii + vi + 1], indices[ii + vi + 2]});
        ii += fvc;
    }
    return m;
}

std::shared_ptr<LightIR> parse_surface_light(const Node& n, uint32_t kind) {
    auto l = std::make_shared<LightIR>();
    l->kind = kind;
    for (auto& c : n.children) {
        if (!c.is_leaf) continue;
        if (kind == PSY_LIGHT_SPHERE && c.type == "Radius") {
            l->shape.push_back(parse_floats_exact(c, 1, "bad Radius")[0]);
        } else if (kind == PSY_LIGHT_RECT && c.type == "Dimensions") {
            auto v = parse_floats_exact(c, 2, "bad Dimensions");
            l->shape.push_back(v[0]);
            l->shape.push_back(v[1]);
        } else if (c.type == "Color") {
            l->colors.push_back(parse_color(c));
        }
    }
    return l;
}

struct NameTables {
    std::map<std::string, int> shaders;
    std::map<std::string, uint32_t> objects;
    std::map<std::string, uint32_t> assemblies;
};

void add_object_node(const Node& child, AssemblyIR& a, NameTables& names) {
    if (!child.has_ident) throw std::runtime_error(child.type + " needs a name");
    if (names.objects.count(child.ident) || names.assemblies.count(child.ident))
        throw std::runtime_error("Attempted to add object with a name that already exists: " + child.ident);
    ObjectIR o;
    if (child.type == "MeshSurface") {
        o.kind = PSY_OBJECT_MESH;
        o.mesh = parse_mesh(child);
        if (const Node* b = first_leaf(child, "SurfaceShaderBind")) {
            auto it = names.shaders.find(trim(b->contents));
            if (it == names.shaders.end())
                throw std::runtime_error("Unknown surface shader '" + trim(b->contents) + "'.");
            o.default_shader = it->second;
        }
    } else {
        o.kind = PSY_OBJECT_LIGHT;
        o.light = parse_surface_light(child, child.type == "SphereLight" ? PSY_LIGHT_SPHERE : PSY_LIGHT_RECT);
    }
    names.objects[child.ident] = (uint32_t)a.objects.size();
    a.objects.push_back(std::move(o));
}

void add_shader_node(const Node& child, AssemblyIR& a, NameTables& names) {
    if (!child.has_ident) throw std::runtime_error("SurfaceShader needs a name");
    if (names.shaders.count(child.ident))
        throw std::runtime_error("Attempted to add surface shader with a name that already exists.");
    names.shaders[child.ident] = (int)a.shaders.size();
    a.shaders.push_back(parse_surface_shader(child));
}

// parse_assembly (psy_assembly.rs:17-188).  `pre` holds scene-level Shaders/Objects children that
// the new layout declares outside the assembly; they are treated as if declared at its top.
std::shared_ptr<AssemblyIR> parse_assembly(const Node& tree, const std::vector<const Node*>& pre) {
    if (tree.is_leaf) throw std::runtime_error("Assembly must be an internal node");
    auto a = std::make_shared<AssemblyIR>();
    NameTables names;
    for (const Node* p : pre) {
        if (p->type == "SurfaceShader")
            add_shader_node(*p, *a, names);
        else if (p->type == "MeshSurface" || p->type == "SphereLight" || p->type == "RectangleLight")
            add_object_node(*p, *a, names);
    }
    for (auto& child : tree.children) {
        if (child.type == "Assembly") {
            if (!child.has_ident || child.is_leaf) throw std::runtime_error("sub-Assembly needs a name");
            if (names.objects.count(child.ident) || names.assemblies.count(child.ident))
                throw std::runtime_error("Attempted to add assembly with a name that already exists.");
            names.assemblies[child.ident] = (uint32_t)a->assemblies.size();
            a->assemblies.push_back(parse_assembly(child, {}));
        } else if (child.type == "Instance") {
            if (child.is_leaf) throw std::runtime_error("Instance must be an internal node");
            size_t nd = 0;
            const Node* data = nullptr;
            for (auto& c : child.children)
                if (c.is_leaf && c.type == "Data") {
                    ++nd;
                    if (!data) data = &c;
                }
            if (nd != 1) throw std::runtime_error("Instance needs exactly one Data");
            const std::string& name = data->contents;  // compared verbatim (psy_assembly.rs:46-51,77)
            InstanceIR in;
            if (const Node* b = first_leaf(child, "SurfaceShaderBind")) {
                auto it = names.shaders.find(b->contents);
                if (it == names.shaders.end()) it = names.shaders.find(trim(b->contents));
                if (it == names.shaders.end())
                    throw std::runtime_error("Unknown surface shader '" + b->contents + "'.");
                in.shader_index = it->second;
            }
            for (auto& c : child.children)
                if (c.is_leaf && c.type == "Transform") in.xforms.push_back(parse_matrix(c));
            auto io = names.objects.find(name);
            auto ia = names.assemblies.find(name);
            if (io == names.objects.end() && ia == names.assemblies.end()) {
                io = names.objects.find(trim(name));
                ia = names.assemblies.find(trim(name));
            }
            if (io != names.objects.end()) {
                in.is_assembly = false;
                in.data_index = io->second;
            } else if (ia != names.assemblies.end()) {
                in.is_assembly = true;
                in.data_index = ia->second;
            } else {
                throw std::runtime_error("Attempted to add instance for data with a name that doesn't exist: " +
                                         name + " at byte " + std::to_string(data->offset));
            }
            a->instances.push_back(std::move(in));
        } else if (child.type == "SurfaceShader") {
            add_shader_node(child, *a, names);
        } else if (child.type == "MeshSurface" || child.type == "SphereLight" || child.type == "RectangleLight") {
            add_object_node(child, *a, names);
        }
    }
    return a;
}

SceneIR parse_scene(const Node& tree) {
    auto need_one = [&](const char* t) {
        size_t c = count_children(tree, t);
        if (c != 1)
            throw std::runtime_error(std::string("Scene should have precisely one ") + t + " section (found " +
                                     std::to_string(c) + ").");
        return first_child(tree, t);
    };
    SceneIR sc;
    if (tree.has_ident) sc.name = tree.ident;
    const Node* output = need_one("Output");
    const Node* rs = need_one("RenderSettings");
    const Node* cam = need_one("Camera");
    const Node* world = need_one("World");
    const Node* asmb = need_one("Assembly");

    // parse_output_info (psy.rs:202-258)
    {
        bool found = false;
        for (auto& c : output->children)
            if (c.is_leaf && c.type == "Path") {
                std::string tc = trim(c.contents);
                if (tc.size() < 2 || tc.front() != '"' || tc.back() != '"')
                    throw std::runtime_error("File paths must be surrounded by quotes.");
                sc.output_path = tc.substr(1, tc.size() - 2);
                found = true;
            }
        if (!found) throw std::runtime_error("Output section must contain a Path.");
    }
    // parse_render_settings (psy.rs:260-351)
    {
        bool fr = false, fs = false;
        for (auto& c : rs->children) {
            if (!c.is_leaf) continue;
            if (c.type == "Resolution") {
                auto v = parse_uints_exact(c.contents, 2, "Resolution should be '[width height]'.");
                sc.res_x = v[0];
                sc.res_y = v[1];
                fr = true;
            } else if (c.type == "SamplesPerPixel") {
                auto v = parse_uints_exact(c.contents, 1, "SamplesPerPixel should be '[samples]'.");
                sc.spp = v[0];
                fs = true;
            } else if (c.type == "Seed") {
                auto v = parse_uints_exact(c.contents, 1, "Seed should be '[seed]'.");
                sc.seed = v[0];
            }
        }
        if (!fr || !fs) throw std::runtime_error("RenderSettings must have both Resolution and SamplesPerPixel.");
    }
    // parse_camera (psy.rs:353-453)
    for (auto& c : cam->children) {
        if (!c.is_leaf) continue;
        if (c.type == "Fov")
            sc.cam_fovs.push_back(parse_floats_exact(c, 1, "bad Fov")[0] * (3.14159265358979323846f / 180.0f));
        else if (c.type == "FocalDistance")
            sc.cam_focus.push_back(parse_floats_exact(c, 1, "bad FocalDistance")[0]);
        else if (c.type == "ApertureRadius")
            sc.cam_aperture.push_back(parse_floats_exact(c, 1, "bad ApertureRadius")[0]);
        else if (c.type == "Transform")
            sc.cam_xforms.push_back(parse_matrix(c));
    }
    // parse_world (psy.rs:455-554)
    {
        if (count_children(*world, "BackgroundShader") != 1)
            throw std::runtime_error("World should have precisely one BackgroundShader section.");
        const Node* bgs = first_child(*world, "BackgroundShader");
        const Node* t = first_leaf(*bgs, "Type");
        if (!t || count_children(*bgs, "Type") != 1)
            throw std::runtime_error("BackgroundShader should have precisely one Type specified.");
        if (trim(t->contents) != "Color")
            throw std::runtime_error("The specified BackgroundShader Type isn't a recognized type.");
        const Node* col = first_leaf(*bgs, "Color");
        if (!col) throw std::runtime_error("BackgroundShader's Type is Color, but no Color is specified.");
        sc.background = parse_color(*col);
        for (auto& c : world->children) {
            if (c.is_leaf || c.type != "DistantDiskLight") continue;
            DistantLightIR dl;
            for (auto& l : c.children) {
                if (!l.is_leaf) continue;
                if (l.type == "Radius")
                    dl.radii.push_back(parse_floats_exact(l, 1, "bad Radius")[0]);
                else if (l.type == "Direction") {
                    auto v = parse_floats_exact(l, 3, "bad Direction");
                    dl.dirs.push_back(v3(v[0], v[1], v[2]));
                } else if (l.type == "Color")
                    dl.colors.push_back(parse_color(l));
            }
            sc.distant_lights.push_back(std::move(dl));
        }
    }
    // New layout: scene-level Shaders{} / Objects{}
    std::vector<const Node*> pre;
    for (auto& c : tree.children)
        if (!c.is_leaf && c.type == "Shaders")
            for (auto& s : c.children) pre.push_back(&s);
    for (auto& c : tree.children)
        if (!c.is_leaf && c.type == "Objects")
            for (auto& s : c.children) pre.push_back(&s);
    sc.root = parse_assembly(*asmb, pre);
    return sc;
}

}  // namespace

std::vector<SceneIR> load_psy_text(const std::string& text) {
    Lexer lx(text);
    std::vector<SceneIR> scenes;
    for (;;) {
        Node n;
        if (!lx.parse_node(n)) break;
        if (n.type == "Scene" && !n.is_leaf) scenes.push_back(parse_scene(n));
    }
    lx.skip_ws_comments();
    if (lx.p < text.size()) lx.fail("unexpected text", lx.p);
    return scenes;
}


// ---- test hooks (tests/test_parse_semantics.py mirrors the reference's own parser tests with them)
// Canonical text of the parsed tree: `Type{...}` / `Type $ident{...}` / `Type[contents]`, siblings separated by ';'.
static void dump_node(const Node& n, std::string& out) {
    out += n.type;
    if (n.has_ident) out += " " + n.ident;
    if (n.is_leaf) {
        out += "[" + n.contents + "]";
    } else {
        out += "{";
        for (size_t i = 0; i < n.children.size(); ++i) {
            if (i) out += ";";
            dump_node(n.children[i], out);
        }
        out += "}";
    }
}
std::string datatree_dump(const std::string& text) {
    Lexer lx(text);
    std::string out;
    bool first = true;
    for (;;) {
        Node n;
        if (!lx.parse_node(n)) break;
        if (!first) out += ";";
        first = false;
        dump_node(n, out);
    }
    lx.skip_ws_comments();
    if (lx.p != text.size()) lx.fail("unexpected text", lx.p);
    return out;
}
size_t test_parse_floats(const std::string& text, float* out, size_t cap, size_t* rest) {
    std::vector<float> v = parse_floats(text, rest);
    for (size_t i = 0; i < v.size() && i < cap; ++i) out[i] = v[i];
    return v.size();
}
size_t test_parse_uints(const std::string& text, uint32_t* out, size_t cap, size_t* rest) {
    std::vector<uint32_t> v = parse_uints(text, rest);
    for (size_t i = 0; i < v.size() && i < cap; ++i) out[i] = v[i];
    return v.size();
}

}  // namespace psyh
