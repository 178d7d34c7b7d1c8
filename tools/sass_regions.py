#!/usr/bin/env python3
"""Attributes the executed instructions of one kernel in an ncu capture to the traversal phases.

ncu's source page attributes inlined helpers (dmath.cuh, dtrace.cuh) to their own lines, which hides
WHERE in the traversal they run.  `nvdisasm -gi` keeps the whole inline chain of every SASS
instruction, so each instruction is assigned to the outermost function of dtraverse.cuh / dshade.cuh
that is not the stepping loop itself; the ncu per-instruction counters (same library build!) are then
summed per phase.

  python tools/sass_regions.py gpurun_out/prof_trace_r1i.ncu-rep k_traceILb0 [psychopath_b200/libpsy_cuda.so]
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_DIR = os.path.join(REPO, "psychopath_b200", "csrc", "cuda")
LOOP_FUNCS = {"run_warp"}


def function_ranges(path):
    """[(first_line, last_line, name)] of the PSY_DEV / __global__ functions of one header."""
    starts = []
    if os.path.basename(path) == "dpool.cuh":
        # one kernel: its phases are delimited by "// ---- <name>" comments
        for i, l in enumerate(open(path), 1):
            m = re.match(r"\s*// ---- ([\w /-]+)", l)
            if m:
                starts.append((i, "pool: " + m.group(1).strip()[:24]))
            elif "__global__" in l:
                starts.append((i, "pool: setup"))
        n = sum(1 for _ in open(path))
        return [(s, (starts[k + 1][0] - 1 if k + 1 < len(starts) else n), name) for k, (s, name) in enumerate(starts)]
    for i, l in enumerate(open(path), 1):
        m = re.match(r"\s*(?:template\s*<[^>]*>\s*)?(?:PSY_DEV|PSY_NOINLINE|__global__|__device__)[^;(]*?(\w+)\s*\(", l)
        if m and not l.strip().startswith("//"):
            starts.append((i, "k_body" if "__global__" in l else m.group(1)))
    n = sum(1 for _ in open(path))
    return [(s, (starts[k + 1][0] - 1 if k + 1 < len(starts) else n), name) for k, (s, name) in enumerate(starts)]


def load_ranges():
    out = {}
    for f in os.listdir(CUDA_DIR):
        if f.endswith((".cuh", ".cu")):
            out[f] = function_ranges(os.path.join(CUDA_DIR, f))
    return out


def func_of(ranges, fname, line):
    for a, b, name in ranges.get(fname, []):
        if a <= line <= b:
            return name
    return None


def disasm_regions(lib, kernel_substr):
    ranges = load_ranges()
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=td, check=True, capture_output=True)
        cubin = [os.path.join(td, f) for f in os.listdir(td) if f.endswith(".cubin")][0]
        text = subprocess.run(["nvdisasm", "-gi", cubin], capture_output=True, text=True, check=True).stdout
    regions, chain, pending, inside = {}, [], [], False
    loc = re.compile(r'"([^"]+)", line (\d+)')
    for l in text.split("\n"):
        if l.startswith("//---") and ".text." in l:
            inside = kernel_substr in l
            chain, pending = [], []
            continue
        if not inside:
            continue
        s = l.strip()
        if s.startswith("//## File"):
            pending.append([(os.path.basename(a), int(b)) for a, b in loc.findall(s)])
            continue
        m = re.match(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", s)
        if not m:
            continue
        if pending:
            # the first annotation holds the innermost frame (+ its call site); the following ones the outer frames
            chain = [fl for ann in pending for fl in ann]
            pending = []
        funcs = [func_of(ranges, f, ln) for f, ln in chain if f in ("dtraverse.cuh", "dshade.cuh", "kernels.cuh", "dpool.cuh")]
        funcs = [f for f in funcs if f]
        outer = [f for f in funcs if f not in LOOP_FUNCS and not f.startswith("k_")]
        if outer:
            region = outer[-1]
        elif funcs:
            region = funcs[-1]
        else:
            region = "?"
        regions[int(m.group(1), 16)] = (region, m.group(2).split()[0] if not m.group(2).startswith("@") else m.group(2).split()[1])
    return regions


def main():
    rep, ksub = sys.argv[1], sys.argv[2]
    lib = sys.argv[3] if len(sys.argv) > 3 else os.path.join(REPO, "psychopath_b200", "libpsy_cuda.so")
    regions = disasm_regions(lib, ksub)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    ia, ii, it, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    base = None
    agg = collections.defaultdict(lambda: [0, 0, 0, 0])
    ops = collections.defaultdict(collections.Counter)
    missing = 0
    for r in rows[2:]:
        if len(r) <= it or not r[ia].startswith("0x"):
            continue
        addr = int(r[ia], 16)
        base = addr if base is None else base
        reg, op = regions.get(addr - base, (None, None))
        if reg is None:
            missing += 1
            reg, op = "?", "?"
        a = agg[reg]
        a[0] += int(r[ii]); a[1] += int(r[it]); a[2] += int(r[isamp]); a[3] += 1
        ops[reg][op.split(".")[0]] += int(r[ii])
    tot = max(sum(a[0] for a in agg.values()), 1)
    ts = max(sum(a[2] for a in agg.values()), 1)
    print(f"# {rep} [{ksub}]: executed warp instructions per traversal phase ({missing} SASS rows unmatched)")
    print(f"# {'phase':28s} {'%warp-inst':>10s} {'%samples':>9s} {'thr/inst':>9s} {'sass':>6s}  top opcodes")
    for reg, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        top = " ".join(f"{k}:{v / max(a[0], 1) * 100:.0f}%" for k, v in ops[reg].most_common(6))
        print(f"  {reg:28s} {a[0] / tot * 100:9.1f}% {a[2] / ts * 100:8.1f}% {a[1] / max(a[0], 1):9.1f} {a[3]:6d}  {top}")


if __name__ == "__main__":
    main()
