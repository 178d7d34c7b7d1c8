"""DRAM / L2 traffic of k_trace_pool per ray, for bench.py's `roofline.traffic`.

  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -k regex:k_trace_pool \
      --csv --log-file gpurun_out/trace_dram_r2.csv python tools/trace_traffic.py run      # on the GPU box
  python tools/trace_traffic.py summarize gpurun_out/trace_dram_r2.csv                      # -> profiles/r2_trace_dram.json

`run` renders the C5 frame twice at 4 spp and records the rays traced (gpurun_out/trace_rays.json).  The summary is
keyed to a hash of the kernel's sources: bench.py reports `traffic` only while the hash still matches."""
import csv
import hashlib
import json
import os
import sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
KERNEL_SOURCES = ["dpool.cuh", "dtrace.cuh", "dmath.cuh", "dscene.cuh"]


def kernel_source_hash():
    h = hashlib.sha256()
    for f in KERNEL_SOURCES:
        h.update(open(os.path.join(REPO, "psychopath_b200", "csrc", "cuda", f), "rb").read())
    return h.hexdigest()


def run():
    import bench
    from psychopath_b200 import api
    kw = dict(bench.WORKLOADS["C5"])
    kw["spp"] = 4
    r = api.Renderer(api.Scene.synth_instanced(**kw))
    rays = 0
    for _ in range(2):
        rays += int(r.render(device_resident_only=True)[1]["ray_count"])
    json.dump({"rays": rays, "frames": 2, "spp": 4}, open(os.path.join(REPO, "gpurun_out", "trace_rays.json"), "w"))
    print("rays", rays)


def summarize(path):
    rays = json.load(open(os.path.join(REPO, "gpurun_out", "trace_rays.json")))
    tot = {"dram__bytes_read.sum": 0.0, "dram__bytes_write.sum": 0.0, "lts__t_bytes.sum": 0.0}
    launches = set()
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    for row in csv.DictReader(l for l in open(path) if not l.startswith("==")):
        m = row.get("Metric Name")
        if m in tot and "k_trace_pool" in row["Kernel Name"]:
            tot[m] += float(row["Metric Value"].replace(",", "")) * unit.get(row["Metric Unit"], 1.0)
            launches.add(row["ID"])
    dram = tot["dram__bytes_read.sum"] + tot["dram__bytes_write.sum"]
    out = {"kernel": "k_trace_pool", "dram_bytes_per_ray": dram / rays["rays"], "l2_bytes_per_ray": tot["lts__t_bytes.sum"] / rays["rays"],
           "kernel_source_sha256": kernel_source_hash(), "kernel_sources": KERNEL_SOURCES,
           "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -k "
                     "regex:k_trace_pool over tools/trace_traffic.py run (C5 scene, %d frames at %d spp): %d launches, %.1f GB "
                     "DRAM and %.1f GB L2 traffic for %.1f M rays" % (rays["frames"], rays["spp"], len(launches), dram / 1e9,
                                                                    tot["lts__t_bytes.sum"] / 1e9, rays["rays"] / 1e6)}
    json.dump(out, open(os.path.join(REPO, "profiles", "r2_trace_dram.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "run":
        run()
    else:
        summarize(sys.argv[2])
