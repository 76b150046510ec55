#!/usr/bin/env python
"""compute-sanitizer logs (gpurun_out/r2_sanitizer_<tool>.log + _pytest.log) -> profiles/r2_sanitizer.md"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = os.path.join(ROOT, "gpurun_out")
out = ["# compute-sanitizer pass (round 2)\n",
       "`tools/gpu/sanitize.sh` on one B200: `compute-sanitizer --tool <tool> --target-processes all` around a\n"
       "reduced-size selection of the GPU tests that launches every kernel family (split-precision GEMM incl.\n"
       "transposed packs and the fused epilogues, loss pipeline, sim_topk + seed + re-rank + exact path, staged\n"
       "find_matches, metrics, Adam kernels).  Each tool is capped at 7 minutes of box time; a tool that did not\n"
       "finish reports how far the test session got.\n",
       "| tool | summary line(s) | pytest tail | distinct report kinds |", "|---|---|---|---|"]
for tool in ("memcheck", "synccheck", "racecheck"):
    log = os.path.join(src, f"r2_sanitizer_{tool}.log")
    py = os.path.join(src, f"r2_sanitizer_{tool}_pytest.log")
    if not os.path.exists(log):
        out.append(f"| {tool} | (no log) | | |")
        continue
    txt = open(log, errors="replace").read()
    summ = [l.strip("= ").strip() for l in txt.splitlines() if "SUMMARY" in l]
    kinds = {}
    for l in txt.splitlines():
        m = re.match(r"=========\s+(Invalid \S+ \S+ of size \d+|Race reported between .*? at|Barrier error.*|"
                     r"Uninitialized .*? of size \d+|Error: .*|Warning: .*|Program hit .*? on CUDA API call to \S+)", l)
        if m:
            k = re.sub(r"0x[0-9a-f]+", "0x..", m.group(1))[:90]
            kinds[k] = kinds.get(k, 0) + 1
    tail = ""
    if os.path.exists(py):
        lines = [l for l in open(py, errors="replace").read().splitlines() if l.strip()]
        tail = lines[-1][:120] if lines else ""
    out.append(f"| {tool} | {'; '.join(sorted(set(summ))) or '(none: run cut off)'} | `{tail}` | "
               f"{'; '.join(f'{k} x{v}' for k, v in sorted(kinds.items())) or 'none'} |")
open(os.path.join(ROOT, "profiles", "r2_sanitizer.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
