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
