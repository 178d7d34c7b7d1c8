"""The part of bench.py's contract that can be checked without a GPU: the reference arm (`--impl reference`, the
CPU restatement timed on host cores) prints one JSON line with the keys the driver reads, rank > 0 under torchrun
stays silent, and the product arm refuses to run without a CUDA device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(args, env_extra=None, timeout=600):
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(REPO, "bench.py")] + args, cwd=REPO, env=env, capture_output=True,
                          text=True, timeout=timeout)


def test_reference_arm_line():
    p = run_bench(["--impl", "reference", "--workload", "tiny", "--steps", "1", "--warmup", "0"])
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mrays/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_is_rank0_only():
    p = run_bench(["--impl", "reference", "--workload", "tiny", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                  {"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, timeout=120)
    assert p.returncode == 0 and not [l for l in p.stdout.splitlines() if l.startswith("{")]


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        return  # covered by the gpu suite and the bench run itself
    p = run_bench(["--workload", "tiny", "--steps", "1", "--warmup", "0"], timeout=300)
    assert p.returncode != 0
    assert "no CPU fallback" in (p.stderr + p.stdout)
