"""Visit counters of the two counting kernels on the same scene: k_trace<true> (one ray per lane) and the counting
instantiation of k_trace_pool (PSY_POOL_COUNT=1).  Run each in its own process:
    python tools/dbg_counters.py            # k_trace<true>
    PSY_POOL_COUNT=1 python tools/dbg_counters.py"""
import os
import sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
from psychopath_b200 import api
s = api.Scene.synth_instanced(asset_quads=200, n_assets=4, n_clusters=40, inst_per_cluster=8, lights_per_cluster=6,
                              cluster_xform_keys=2, res=(480, 270), spp=8, seed=11, motion_pct=1250)
r = api.Renderer(s)
a, sa = r.render()
c, sc = r.render(counters=True)
print("image equal:", np.array_equal(a, c, equal_nan=True), "rays", sa["ray_count"], sc["ray_count"], "nodes", sc["accel_node_visits"],
      "motion", sc["motion_node_visits"], "tris", sc["tri_tests"], "inst", sc["instance_visits"])
