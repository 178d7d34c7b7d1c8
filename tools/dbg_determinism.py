"""Debug helper: renders the same image of a C5-like scene several times and compares the results bit for bit."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from psychopath_b200 import api
q = int(sys.argv[1])
s = api.Scene.synth_instanced(asset_quads=q, n_assets=4, n_clusters=100, inst_per_cluster=10, lights_per_cluster=10,
                              cluster_xform_keys=2, res=(int(sys.argv[2]), int(sys.argv[3])), spp=int(sys.argv[4]), seed=5)
r = api.Renderer(s)
ref, st = r.render(counters=False)
for i in range(3):
    img, st2 = r.render(counters=(i == 2))
    d = ~np.isclose(img, ref, rtol=0, atol=0, equal_nan=True)
    print(i, "rays", st2["ray_count"], st["ray_count"], "differing pixels:", int(d.any(axis=-1).sum()))
