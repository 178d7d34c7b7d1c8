"""Debug helper: renders a small image of a C5-like scene with and without counters (run under compute-sanitizer)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from psychopath_b200 import api
q = int(sys.argv[1])
s = api.Scene.synth_instanced(asset_quads=q, n_assets=4, n_clusters=100, inst_per_cluster=10, lights_per_cluster=10,
                              cluster_xform_keys=2, res=(int(sys.argv[2]), int(sys.argv[3])), spp=int(sys.argv[4]), seed=5)
r = api.Renderer(s)
mode = sys.argv[5] if len(sys.argv) > 5 else "both"
if mode != "nocount":
    img, st = r.render(counters=True)
    print("ok", st["ray_count"], st["accel_node_visits"], st["tri_tests"], st["instance_visits"])
if mode != "count":
    img2, st2 = r.render(counters=False)
    print("plain", st2["ray_count"])
if mode == "both":
    print("same image:", np.array_equal(img, img2, equal_nan=True))
