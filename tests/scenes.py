"""Scene fixtures shared by tests, smoke() and bench.py (inputs only; no reference code)."""
import math
import os
import re

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def cornell_text():
    """example_scenes/cornell_box.psy verbatim (BASELINE config C1)."""
    return open(os.path.join(GOLDEN, "cornell_box.psy")).read()


def cube_text(res=None, spp=None, camera_motion=False):
    """example_scenes/cube.psy; with camera_motion=True a second camera Transform key is added
    (first key composed with translate(0.3,0,0) and a 2 degree roll) as BASELINE config C2 asks."""
    text = open(os.path.join(GOLDEN, "cube.psy")).read()
    if res is not None:
        text = re.sub(r"Resolution \[[^\]]*\]", f"Resolution [{res[0]} {res[1]}]", text)
    if spp is not None:
        text = re.sub(r"SamplesPerPixel \[[^\]]*\]", f"SamplesPerPixel [{spp}]", text)
    if camera_motion:
        cam = re.search(r"Camera \{.*?\n    \}", text, re.S).group(0)
        m = re.search(r"Transform \[([^\]]*)\]", cam)
        v = [float(x) for x in m.group(1).split()]
        cols = [v[0:4], v[4:8], v[8:12], v[12:16]]  # column-major
        a = math.radians(2.0)
        c, s = math.cos(a), math.sin(a)
        # second key = first key * (translate(0.3,0,0) * rotZ(2deg)) in camera space
        local = [[c, s, 0, 0], [-s, c, 0, 0], [0, 0, 1, 0], [0.3, 0, 0, 1]]  # columns
        out = []
        for j in range(4):
            col = [sum(cols[k][i] * local[j][k] for k in range(4)) for i in range(4)]
            out.extend(col)
        key2 = "Transform [" + " ".join(f"{x:.6f}" for x in out) + "]"
        cam2 = cam.replace(m.group(0), m.group(0) + "\n        " + key2)
        text = text.replace(cam, cam2)
    return text


def motion_instances_text():
    """A small old-layout (PsychoBlend style) scene: nested assemblies, transform and deformation
    motion blur, vertex normals, GGX/Lambert/Emit shaders, sphere + rectangle + distant lights."""
    return """
Scene $fixture {
    Output { Path ["fixture.png"] }
    RenderSettings { Resolution [96 64] SamplesPerPixel [8] Seed [7] }
    Camera {
        Fov [50.0]
        Fov [52.0]
        FocalDistance [9.0]
        ApertureRadius [0.05]
        Transform [1 0 0 0  0 0 1 0  0 1 0 0  0 -9 2 1]
        Transform [1 0 0 0  0 0 1 0  0 1 0 0  0.2 -9 2.1 1]
    }
    World {
        BackgroundShader { Type [Color] Color [rec709, 0.05 0.06 0.09] }
        DistantDiskLight $sun {
            Direction [-0.3 0.4 -0.8]
            Color [rec709, 1.5 1.4 1.2]
            Radius [0.05]
        }
    }
    Assembly {
        SurfaceShader $grey { Type [Lambert] Color [rec709, 0.6 0.6 0.6] }
        SurfaceShader $metal { Type [GGX] Color [rec709, 0.9 0.7 0.4] Roughness [0.25] Fresnel [0.6] }
        SurfaceShader $glow { Type [Emit] Color [blackbody, 3200 0.00000002] }
        MeshSurface $floor {
            Vertices [-8 -8 0  8 -8 0  8 8 0  -8 8 0]
            FaceVertCounts [4]
            FaceVertIndices [0 1 2 3]
        }
        MeshSurface $tent {
            Vertices [-1 -1 0  1 -1 0  1 1 0  -1 1 0  0 0 1.5]
            Vertices [-1 -1 0  1 -1 0  1 1 0  -1 1 0  0.4 0.2 1.9]
            Normals [-1 -1 0.2  1 -1 0.2  1 1 0.2  -1 1 0.2  0 0 1]
            Normals [-1 -1 0.2  1 -1 0.2  1 1 0.2  -1 1 0.2  0.2 0.1 1]
            FaceVertCounts [3 3 3 3]
            FaceVertIndices [0 1 4  1 2 4  2 3 4  3 0 4]
        }
        SphereLight $bulb { Color [rec709, 30 28 25] Radius [0.3] }
        RectangleLight $panel { Color [rec709, 20 22 26] Dimensions [1.5 1.0] Dimensions [1.8 1.0] }
        Assembly $group {
            MeshSurface $box {
                Vertices [-0.5 -0.5 0  0.5 -0.5 0  0.5 0.5 0  -0.5 0.5 0  -0.5 -0.5 1  0.5 -0.5 1  0.5 0.5 1  -0.5 0.5 1]
                FaceVertCounts [4 4 4 4 4 4]
                FaceVertIndices [0 3 2 1  4 5 6 7  0 1 5 4  1 2 6 5  2 3 7 6  3 0 4 7]
            }
            SphereLight $lamp { Color [color_temperature, 4500 8.0] Radius [0.15] }
            SurfaceShader $red { Type [Lambert] Color [rec709, 0.7 0.1 0.1] }
            Instance { Data [$box] SurfaceShaderBind [$red] Transform [1 0 0 0  0 1 0 0  0 0 1 0  -1.2 0 0 1] }
            Instance { Data [$box] SurfaceShaderBind [$red]
                Transform [0.8 0.6 0 0  -0.6 0.8 0 0  0 0 1 0  1.2 0.3 0 1]
                Transform [0.6 0.8 0 0  -0.8 0.6 0 0  0 0 1 0  1.2 0.3 -0.3 1] }
            Instance { Data [$lamp] Transform [1 0 0 0  0 1 0 0  0 0 1 0  0 0 -2.2 1] }
        }
        Instance { Data [$floor] SurfaceShaderBind [$grey] }
        Instance { Data [$tent] SurfaceShaderBind [$metal] Transform [1 0 0 0  0 1 0 0  0 0 1 0  2.5 -1 0 1] }
        Instance { Data [$tent] SurfaceShaderBind [$glow]
            Transform [0.5 0 0 0  0 0.5 0 0  0 0 0.5 0  -1.5 -1.5 0 1]
            Transform [0.5 0 0 0  0 0.5 0 0  0 0 0.5 0  -1.7 -1.5 -0.2 1]
            Transform [0.5 0 0 0  0 0.5 0 0  0 0 0.5 0  -1.9 -1.4 -0.2 1] }
        Instance { Data [$bulb] Transform [1 0 0 0  0 1 0 0  0 0 1 0  3 2 -3.5 1] }
        Instance { Data [$panel] Transform [1 0 0 0  0 -1 0 0  0 0 -1 0  -2 1 4 1] }
        Instance { Data [$group] Transform [1 0 0 0  0 1 0 0  0 0 1 0  0 -2 0 1] Transform [1 0 0 0  0 1 0 0  0 0 1 0  -0.4 -2 0 1] }
        Instance { Data [$group] Transform [0 1 0 0  -1 0 0 0  0 0 1 0  3 3 0 1] }
    }
}
"""


def nested_assemblies_text():
    """Assemblies nested three deep (root -> $town -> $block -> $house -> meshes / lights), with and without
    transforms of their own and with transform motion at two levels: exercises every way of entering and leaving
    an instance (first level: kept in shared memory by k_trace_pool; deeper: saved to the scratch area)."""
    return """
Scene $nested {
    Output { Path ["nested.png"] }
    RenderSettings { Resolution [96 64] SamplesPerPixel [8] Seed [11] }
    Camera {
        Fov [55.0]
        FocalDistance [12.0]
        ApertureRadius [0.02]
        Transform [1 0 0 0  0 0 1 0  0 1 0 0  0 -12 3 1]
    }
    World {
        BackgroundShader { Type [Color] Color [rec709, 0.08 0.09 0.12] }
        DistantDiskLight $sun { Direction [-0.2 0.5 -0.8] Color [rec709, 1.2 1.1 1.0] Radius [0.04] }
    }
    Assembly {
        SurfaceShader $grey { Type [Lambert] Color [rec709, 0.6 0.6 0.6] }
        MeshSurface $ground {
            Vertices [-12 -12 0  12 -12 0  12 12 0  -12 12 0]
            FaceVertCounts [4]
            FaceVertIndices [0 1 2 3]
        }
        Assembly $town {
            SurfaceShader $brick { Type [Lambert] Color [rec709, 0.7 0.3 0.2] }
            Assembly $block {
                SurfaceShader $steel { Type [GGX] Color [rec709, 0.8 0.8 0.9] Roughness [0.3] Fresnel [0.5] }
                Assembly $house {
                    SurfaceShader $brick { Type [Lambert] Color [rec709, 0.7 0.3 0.2] }
                    SurfaceShader $steel { Type [GGX] Color [rec709, 0.8 0.8 0.9] Roughness [0.3] Fresnel [0.5] }
                    MeshSurface $walls {
                        Vertices [-0.5 -0.5 0  0.5 -0.5 0  0.5 0.5 0  -0.5 0.5 0  -0.5 -0.5 1  0.5 -0.5 1  0.5 0.5 1  -0.5 0.5 1]
                        FaceVertCounts [4 4 4 4 4 4]
                        FaceVertIndices [0 3 2 1  4 5 6 7  0 1 5 4  1 2 6 5  2 3 7 6  3 0 4 7]
                    }
                    MeshSurface $roof {
                        Vertices [-0.6 -0.6 1  0.6 -0.6 1  0.6 0.6 1  -0.6 0.6 1  0 0 1.7]
                        Vertices [-0.6 -0.6 1  0.6 -0.6 1  0.6 0.6 1  -0.6 0.6 1  0.1 0 1.8]
                        FaceVertCounts [3 3 3 3]
                        FaceVertIndices [0 1 4  1 2 4  2 3 4  3 0 4]
                    }
                    SphereLight $porch { Color [rec709, 6 5 4] Radius [0.1] }
                    Instance { Data [$walls] SurfaceShaderBind [$brick] }
                    Instance { Data [$roof] SurfaceShaderBind [$steel] Transform [1 0 0 0  0 1 0 0  0 0 1 0  0 0 -0.02 1] }
                    Instance { Data [$porch] Transform [1 0 0 0  0 1 0 0  0 0 1 0  0 0.8 -0.6 1] }
                }
                RectangleLight $lamp { Color [rec709, 9 9 10] Dimensions [0.6 0.4] }
                Instance { Data [$house] }
                Instance { Data [$house] Transform [0.8 0.6 0 0  -0.6 0.8 0 0  0 0 1 0  -1.6 0.2 0 1] }
                Instance { Data [$house]
                    Transform [0.7 0 0 0  0 0.7 0 0  0 0 0.7 0  1.2 0.4 0 1]
                    Transform [0.7 0 0 0  0 0.7 0 0  0 0 0.7 0  1.3 0.4 -0.1 1] }
                Instance { Data [$lamp] Transform [1 0 0 0  0 -1 0 0  0 0 -1 0  0 0 2.5 1] }
            }
            Instance { Data [$block] Transform [1 0 0 0  0 1 0 0  0 0 1 0  -2.5 0 0 1] }
            Instance { Data [$block] Transform [0 1 0 0  -1 0 0 0  0 0 1 0  2.5 1 0 1] Transform [0 1 0 0  -1 0 0 0  0 0 1 0  2.7 1 0 1] }
            Instance { Data [$block] }
        }
        Instance { Data [$ground] SurfaceShaderBind [$grey] }
        Instance { Data [$town] Transform [1 0 0 0  0 1 0 0  0 0 1 0  0 2 0 1] }
        Instance { Data [$town] Transform [0.9 0 0 0  0 0.9 0 0  0 0 0.9 0  6 -3 0 1] Transform [0.9 0 0 0  0 0.9 0 0  0 0 0.9 0  6 -3.3 0 1] }
        Instance { Data [$town] }
    }
}
"""


def random_scene_text(seed, many_keys=False):
    """A seeded random scene in the reference's text format, for the fuzz parity tests: every feature the path reads
    (time keys on cameras, vertices, normals, transforms, light sizes and colours; Lambert / GGX / Emit shaders;
    rec709 / blackbody / color_temperature colours; sphere, rectangle and distant lights; assemblies nested up to
    three deep, instanced with and without transforms) in combinations nobody wrote by hand.  many_keys: meshes and
    instances with 9 and 12 time keys as well (BVH nodes of eight or more keys take their own fetch path in the trace
    kernel, and the device BVH build hands meshes of more than eight keys back to the host)."""
    import random
    rnd = random.Random(seed)

    def f(x):
        return f"{x:.6g}"

    def xform(scale=None, spread=4.0, z=None):
        a = rnd.uniform(0, 2 * math.pi)
        s = scale if scale is not None else rnd.uniform(0.5, 1.3)
        c, sn = math.cos(a) * s, math.sin(a) * s
        tx, ty = rnd.uniform(-spread, spread), rnd.uniform(-spread, spread)
        tz = rnd.uniform(-0.2, 0.6) if z is None else z
        return [c, sn, 0, 0, -sn, c, 0, 0, 0, 0, s, 0, tx, ty, tz, 1]

    def xform_keys(n, **kw):
        base = xform(**kw)
        keys = [base]
        for _ in range(n - 1):
            k = list(keys[-1])
            k[12] += rnd.uniform(-0.4, 0.4)
            k[13] += rnd.uniform(-0.4, 0.4)
            k[14] += rnd.uniform(-0.2, 0.2)
            if rnd.random() < 0.5:  # a little extra rotation between keys
                a = rnd.uniform(-0.15, 0.15)
                c, s = math.cos(a), math.sin(a)
                k[0], k[1], k[4], k[5] = k[0] * c - k[1] * s, k[0] * s + k[1] * c, k[4] * c - k[5] * s, k[4] * s + k[5] * c
            keys.append(k)
        return "".join(" Transform [" + " ".join(f(v) for v in k) + "]" for k in keys)

    def colour(lo, hi):
        kind = rnd.choice(["rec709", "rec709", "rec709", "blackbody", "color_temperature"])
        if kind == "rec709":
            return "[rec709, " + " ".join(f(rnd.uniform(lo, hi)) for _ in range(3)) + "]"
        if kind == "blackbody":  # un-normalised Planck radiance: tiny factors
            return f"[blackbody, {f(rnd.uniform(2500, 7000))} {f(rnd.uniform(lo, hi) * 1e-9)}]"
        return f"[color_temperature, {f(rnd.uniform(2500, 9000))} {f(rnd.uniform(lo, hi))}]"

    def shader(name):
        kind = rnd.choice(["Lambert", "Lambert", "GGX", "GGX", "Emit"])
        if kind == "Lambert":
            return f"SurfaceShader ${name} {{ Type [Lambert] Color {colour(0.1, 0.9)} }}"
        if kind == "GGX":
            return (f"SurfaceShader ${name} {{ Type [GGX] Color {colour(0.2, 0.95)} Roughness [{f(rnd.uniform(0.02, 0.9))}] "
                    f"Fresnel [{f(rnd.uniform(0.0, 1.0))}] }}")
        return f"SurfaceShader ${name} {{ Type [Emit] Color {colour(0.2, 2.0)} }}"

    def mesh(name):
        n_keys = rnd.choice([1, 2, 9, 12] if many_keys else [1, 1, 2, 3])
        with_normals = rnd.random() < 0.4
        kind = rnd.choice(["patch", "patch", "box", "fan"])
        if kind == "patch":
            n = rnd.randint(2, 7)
            size = rnd.uniform(0.8, 2.5)
            base = [[(i / n - 0.5) * size, (j / n - 0.5) * size, rnd.uniform(0, 0.5)] for j in range(n + 1)
                    for i in range(n + 1)]
            counts, idx = [], []
            for j in range(n):
                for i in range(n):
                    a, b, c, d = j * (n + 1) + i, j * (n + 1) + i + 1, (j + 1) * (n + 1) + i + 1, (j + 1) * (n + 1) + i
                    if rnd.random() < 0.3:  # two triangles instead of a quad
                        counts += [3, 3]
                        idx += [a, b, c, a, c, d]
                    else:
                        counts.append(4)
                        idx += [a, b, c, d]
        elif kind == "box":
            sx, sy, sz = (rnd.uniform(0.3, 1.0) for _ in range(3))
            base = [[x * sx, y * sy, z * sz] for z in (0, 1) for (x, y) in ((-1, -1), (1, -1), (1, 1), (-1, 1))]
            counts = [4] * 6
            idx = [0, 3, 2, 1, 4, 5, 6, 7, 0, 1, 5, 4, 1, 2, 6, 5, 2, 3, 7, 6, 3, 0, 4, 7]
        else:  # an n-gon fan around an apex (polygons of more than four vertices are fanned by the loader)
            n = rnd.randint(5, 9)
            r = rnd.uniform(0.5, 1.5)
            base = [[r * math.cos(2 * math.pi * i / n), r * math.sin(2 * math.pi * i / n), 0.0] for i in range(n)]
            base.append([0.0, 0.0, rnd.uniform(0.5, 1.8)])
            counts = [n] + [3] * n
            idx = list(range(n))
            for i in range(n):
                idx += [i, (i + 1) % n, n]
        out = [f"MeshSurface ${name} {{"]
        verts = base
        for k in range(n_keys):
            if k:
                verts = [[v[0] + rnd.uniform(-0.08, 0.08), v[1] + rnd.uniform(-0.08, 0.08), v[2] + rnd.uniform(-0.08, 0.08)]
                         for v in verts]
            out.append("    Vertices [" + "  ".join(" ".join(f(c) for c in v) for v in verts) + "]")
        if with_normals:
            for k in range(n_keys):
                out.append("    Normals [" + "  ".join(
                    " ".join(f(c) for c in (rnd.uniform(-0.3, 0.3), rnd.uniform(-0.3, 0.3), 1.0)) for _ in base) + "]")
        out.append("    FaceVertCounts [" + " ".join(map(str, counts)) + "]")
        out.append("    FaceVertIndices [" + " ".join(map(str, idx)) + "]")
        out.append("}")
        return "\n".join(out)

    def light(name):
        n_keys = rnd.choice([1, 1, 2])
        cols = " ".join(f"Color {colour(4, 40)}" for _ in range(rnd.choice([1, 1, 2])))
        if rnd.random() < 0.5:
            return f"SphereLight ${name} {{ {cols} " + " ".join(f"Radius [{f(rnd.uniform(0.05, 0.5))}]" for _ in range(n_keys)) + " }"
        return f"RectangleLight ${name} {{ {cols} " + " ".join(
            f"Dimensions [{f(rnd.uniform(0.3, 2.0))} {f(rnd.uniform(0.3, 2.0))}]" for _ in range(n_keys)) + " }"

    counter = [0]

    def assembly(depth, indent):
        """The body of an assembly: shaders, meshes, lights, sub-assemblies, then the instances of all of them."""
        counter[0] += 1
        tag = counter[0]
        lines, things, shaders = [], [], []
        for i in range(rnd.randint(1, 3)):
            shaders.append(f"sh{tag}_{i}")
            lines.append(shader(shaders[-1]))
        for i in range(rnd.randint(1, 3)):
            things.append((f"m{tag}_{i}", "mesh"))
            lines.append(mesh(things[-1][0]))
        for i in range(rnd.randint(0 if depth else 1, 2)):
            things.append((f"l{tag}_{i}", "light"))
            lines.append(light(things[-1][0]))
        if depth < 2:
            for i in range(rnd.randint(0, 2) if depth else rnd.randint(1, 2)):
                name = f"a{tag}_{i}"
                things.append((name, "assembly"))
                lines.append(f"Assembly ${name} {{\n" + assembly(depth + 1, indent + "    ") + "\n}")
        spread = 4.0 if depth == 0 else 1.5
        for name, kind in things:
            for _ in range(rnd.randint(1, 3 if kind != "light" else 2)):
                n_keys = rnd.choice([0, 1, 2, 9] if many_keys else [0, 1, 1, 2, 3])
                bind = f" SurfaceShaderBind [${rnd.choice(shaders)}]" if kind == "mesh" else ""
                z = -rnd.uniform(1.5, 4.0) if kind == "light" else None  # (the fixtures' "up" is -z in instance space)
                xf = xform_keys(n_keys, spread=spread, z=z, scale=1.0 if kind == "light" else None) if n_keys else ""
                lines.append(f"Instance {{ Data [${name}]{bind}{xf} }}")
        return "\n".join(indent + ln for block in lines for ln in block.split("\n"))

    cam_keys = rnd.choice([1, 2, 3])
    cam = []
    for k in range(cam_keys):
        cam.append(f"        Transform [1 0 0 0  0 0 1 0  0 1 0 0  {f(0.15 * k)} {f(-11 + 0.1 * k)} {f(2.5 + 0.05 * k)} 1]")
    fovs = "\n".join(f"        Fov [{f(rnd.uniform(40, 60))}]" for _ in range(rnd.choice([1, 2])))
    world = [f"        BackgroundShader {{ Type [Color] Color {colour(0.02, 0.3)} }}"]
    for i in range(rnd.randint(0, 2)):
        n = rnd.choice([1, 2])
        world.append(f"        DistantDiskLight $sun{i} {{ "
                     + " ".join(f"Direction [{f(rnd.uniform(-0.5, 0.5))} {f(rnd.uniform(-0.5, 0.5))} -0.8]" for _ in range(n))
                     + " " + " ".join(f"Color {colour(0.5, 2.0)}" for _ in range(rnd.choice([1, 2])))
                     + " " + " ".join(f"Radius [{f(rnd.uniform(0.01, 0.2))}]" for _ in range(n)) + " }")
    floor = ("        SurfaceShader $floor_sh { Type [Lambert] Color [rec709, 0.5 0.5 0.5] }\n"
             "        MeshSurface $floor {\n            Vertices [-9 -9 0  9 -9 0  9 9 0  -9 9 0]\n"
             "            FaceVertCounts [4]\n            FaceVertIndices [0 1 2 3]\n        }\n"
             "        Instance { Data [$floor] SurfaceShaderBind [$floor_sh] }\n")
    return (f"Scene $fuzz{seed} {{\n    Output {{ Path [\"fuzz{seed}.png\"] }}\n"
            f"    RenderSettings {{ Resolution [64 48] SamplesPerPixel [4] Seed [{seed + 1}] }}\n"
            f"    Camera {{\n{fovs}\n        FocalDistance [{f(rnd.uniform(8, 13))}]\n"
            f"        ApertureRadius [{f(rnd.choice([0.0, 0.03, 0.1]))}]\n" + "\n".join(cam) + "\n    }\n"
            "    World {\n" + "\n".join(world) + "\n    }\n"
            "    Assembly {\n" + floor + assembly(0, "        ") + "\n    }\n}\n")


def edge_case_texts():
    """Small scenes at the edges of the path: nothing to hit, nothing to light with, a mesh that is a single leaf,
    only a distant light, an emitter as the only light source; odd resolutions and sample counts."""
    HEAD = """Scene $edge {
        Output { Path ["edge.png"] }
        RenderSettings { Resolution [%d %d] SamplesPerPixel [%d] Seed [3] }
        Camera { Fov [50.0] FocalDistance [9.0] ApertureRadius [0.0] Transform [1 0 0 0  0 0 1 0  0 1 0 0  0 -9 2 1] }
        World { BackgroundShader { Type [Color] Color [rec709, 0.3 0.4 0.5] } %s }
        Assembly { %s }
    }
    """
    TRI = """SurfaceShader $s { Type [Lambert] Color [rec709, 0.7 0.6 0.5] }
            MeshSurface $t { Vertices [-2 0 -1  2 0 -1  0 0 3] FaceVertCounts [3] FaceVertIndices [0 1 2] }
            Instance { Data [$t] SurfaceShaderBind [$s] }"""
    EMIT = """SurfaceShader $e { Type [Emit] Color [rec709, 2 1 0.5] }
            SurfaceShader $s { Type [GGX] Color [rec709, 0.7 0.6 0.5] Roughness [0.3] Fresnel [0.4] }
            MeshSurface $t { Vertices [-2 0 -1  2 0 -1  0 0 3] FaceVertCounts [3] FaceVertIndices [0 1 2] }
            MeshSurface $f { Vertices [-6 -6 -1  6 -6 -1  6 6 -1  -6 6 -1] FaceVertCounts [4] FaceVertIndices [0 1 2 3] }
            Instance { Data [$t] SurfaceShaderBind [$e] }
            Instance { Data [$f] SurfaceShaderBind [$s] }"""
    SUN = "DistantDiskLight $sun { Direction [-0.3 0.4 -0.8] Color [rec709, 1.5 1.4 1.2] Radius [0.05] }"
    cases = {
        "background_only": HEAD % (37, 23, 3, "", ""),
        "one_triangle_no_lights": HEAD % (37, 23, 3, "", TRI),
        "one_triangle_sun_only": HEAD % (40, 24, 1, SUN, TRI),
        "emitter_lights_a_ggx_floor": HEAD % (33, 21, 5, "", EMIT),
        "emitter_and_sun": HEAD % (32, 20, 4, SUN, EMIT),
    }
    cases["white_furnace_motion_dof"] = furnace_text("Type [Lambert] Color [rec709, 1 1 1]", motion=True, aperture=0.2)
    cases["ggx_furnace"] = furnace_text("Type [GGX] Color [rec709, 1 1 1] Roughness [0.4] Fresnel [1.0]")
    # the closed-form direct-lighting scenes of tests/test_oracle_pins.py
    at = "Transform [1 0 0 0  0 1 0 0  0 0 1 0  -0.5 -1.0 -3.0 1]"
    cases["floor_under_a_distant_disk"] = lit_floor_text(
        world_light="DistantDiskLight $sun { Direction [-0.3 0.4 -0.8] Color [rec709, 2 2 2] Radius [0.001] }", spp=4)
    cases["floor_under_a_sphere_light"] = lit_floor_text(
        objects="SphereLight $bulb { Color [rec709, 40 40 40] Radius [0.05] }", instances=f"Instance {{ Data [$bulb] {at} }}", spp=4)
    cases["floor_under_a_large_close_rectangle"] = lit_floor_text(
        objects="RectangleLight $l { Color [rec709, 10 10 10] Dimensions [3.0 2.0] }",
        instances="Instance { Data [$l] Transform [1 0 0 0  0 1 0 0  0 0 1 0  -0.5 -1.0 -0.9 1] }", spp=4)
    return cases


def furnace_text(shader, motion=False, aperture=0.0):
    """A convex box in a uniform background and nothing else: a path that hits the box leaves it with its first bounce,
    so with a white Lambert surface every sample must carry exactly the background's radiance (the white furnace)."""
    xf = "Transform [0.8 0.6 0 0  -0.6 0.8 0 0  0 0 1 0  0.3 0 0.5 1]"
    if motion:
        xf += " Transform [0.6 0.8 0 0  -0.8 0.6 0 0  0 0 1 0  0.5 0.2 0.3 1]"
    return f"""
Scene $furnace {{
    Output {{ Path ["furnace.png"] }}
    RenderSettings {{ Resolution [48 32] SamplesPerPixel [8] Seed [9] }}
    Camera {{ Fov [40.0] FocalDistance [9.0] ApertureRadius [{aperture}] Transform [1 0 0 0  0 0 1 0  0 1 0 0  0 -9 1 1] }}
    World {{ BackgroundShader {{ Type [Color] Color [rec709, 0.5 0.5 0.5] }} }}
    Assembly {{
        SurfaceShader $s {{ {shader} }}
        MeshSurface $box {{
            Vertices [-1 -1 0  1 -1 0  1 1 0  -1 1 0  -1 -1 2  1 -1 2  1 1 2  -1 1 2]
            FaceVertCounts [4 4 4 4 4 4]
            FaceVertIndices [0 3 2 1  4 5 6 7  0 1 5 4  1 2 6 5  2 3 7 6  3 0 4 7]
        }}
        Instance {{ Data [$box] SurfaceShaderBind [$s] {xf} }}
    }}
}}
"""


def lit_floor_text(world_light="", objects="", instances="", res=(32, 24), spp=16):
    """A large Lambert floor (albedo 0.6) under a black sky, seen from above, lit by whatever the caller adds: the
    analytic direct-lighting pins of tests/test_oracle_pins.py."""
    return f"""
Scene $nee {{
    Output {{ Path ["nee.png"] }}
    RenderSettings {{ Resolution [{res[0]} {res[1]}] SamplesPerPixel [{spp}] Seed [5] }}
    Camera {{ Fov [35.0] FocalDistance [9.0] ApertureRadius [0.0] Transform [1 0 0 0  0 0.6 0.8 0  0 0.8 -0.6 0  0 -6 5 1] }}
    World {{ BackgroundShader {{ Type [Color] Color [rec709, 0 0 0] }} {world_light} }}
    Assembly {{
        SurfaceShader $grey {{ Type [Lambert] Color [rec709, 0.6 0.6 0.6] }}
        MeshSurface $floor {{ Vertices [-40 -40 0  40 -40 0  40 40 0  -40 40 0] FaceVertCounts [4] FaceVertIndices [0 1 2 3] }}
        {objects}
        Instance {{ Data [$floor] SurfaceShaderBind [$grey] }}
        {instances}
    }}
}}
"""


def deeply_nested_text(levels):
    """`levels` assemblies inside one another below the root, each instanced twice with a small transform (one of them
    moving), a box and a lamp at the bottom: the nesting limit of the trace kernels' saved-space stack."""
    inner = """SurfaceShader $s { Type [Lambert] Color [rec709, 0.7 0.5 0.3] }
            MeshSurface $box {
                Vertices [-0.5 -0.5 0  0.5 -0.5 0  0.5 0.5 0  -0.5 0.5 0  -0.5 -0.5 1  0.5 -0.5 1  0.5 0.5 1  -0.5 0.5 1]
                FaceVertCounts [4 4 4 4 4 4]
                FaceVertIndices [0 3 2 1  4 5 6 7  0 1 5 4  1 2 6 5  2 3 7 6  3 0 4 7]
            }
            SphereLight $lamp { Color [rec709, 8 7 6] Radius [0.1] }
            Instance { Data [$box] SurfaceShaderBind [$s] }
            Instance { Data [$lamp] Transform [1 0 0 0  0 1 0 0  0 0 1 0  0.2 0.9 -1.6 1] }"""
    body = inner
    for k in range(levels, 0, -1):
        body = f"""Assembly $a{k} {{
            {body}
        }}
        Instance {{ Data [$a{k}] Transform [0.98 0.17 0 0  -0.17 0.98 0 0  0 0 1 0  0.35 0.1 0 1] }}
        Instance {{ Data [$a{k}] Transform [1 0 0 0  0 1 0 0  0 0 1 0  -0.9 -0.3 0 1] Transform [1 0 0 0  0 1 0 0  0 0 1 0  -1.0 -0.3 -0.05 1] }}"""
    return f"""
Scene $deep {{
    Output {{ Path ["deep.png"] }}
    RenderSettings {{ Resolution [48 32] SamplesPerPixel [4] Seed [2] }}
    Camera {{ Fov [55.0] FocalDistance [9.0] ApertureRadius [0.0] Transform [1 0 0 0  0 0 1 0  0 1 0 0  0 -14 1 1] }}
    World {{ BackgroundShader {{ Type [Color] Color [rec709, 0.1 0.1 0.12] }} }}
    Assembly {{
        SurfaceShader $g {{ Type [Lambert] Color [rec709, 0.5 0.5 0.5] }}
        MeshSurface $floor {{ Vertices [-20 -20 0  20 -20 0  20 20 0  -20 20 0] FaceVertCounts [4] FaceVertIndices [0 1 2 3] }}
        Instance {{ Data [$floor] SurfaceShaderBind [$g] }}
        {body}
    }}
}}
"""
