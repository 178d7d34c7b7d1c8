"""Per-wave rays and kernel times of one C5 frame (PSY_WAVE_STATS=1): is a thin later wave slower per ray than wave 1?
    PSY_WAVE_STATS=1 python tools/wave_stats.py [spp] [pool]"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench
from psychopath_b200 import api
kw = dict(bench.WORKLOADS["C5"])
kw["spp"] = int(sys.argv[1]) if len(sys.argv) > 1 else 256
pool = int(sys.argv[2]) if len(sys.argv) > 2 else 0
r = api.Renderer(api.Scene.synth_instanced(**kw))
os.environ.pop("PSY_WAVE_STATS", None)
r.render(device_resident_only=True, max_samples_per_bucket=pool)
os.environ["PSY_WAVE_STATS"] = "1"
print(f"--- C5 at {kw['spp']} spp, pool {pool or 'default'}", file=sys.stderr)
r.render(device_resident_only=True, max_samples_per_bucket=pool)
