#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/r18_launches.csv profiles/r1_launches.md
  python tools/summarize_ncu.py full gpurun_out/r18_prof.ncu-rep profiles/r1_kernels.md
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict


def launches(src, dst):
    rows = [r for r in csv.reader(l for l in open(src) if not l.startswith("=="))]
    hdr = rows[0]
    ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = OrderedDict()
    order = []
    for r in rows[1:]:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        name = r[ik].split("(")[0].replace("void ", "").replace("mclst::", "")
        t = float(r[iv].replace(",", ""))
        unit = r[hdr.index("Metric Unit")]
        t_us = t / 1e3 if unit in ("ns", "nsecond") else (t if unit in ("us", "usecond") else t * 1e3)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t_us
        order.append((name, t_us))
    total = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({src})\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` "
                "(cold-cache, serialised: compare SHARES, not absolutes)\n\n"
                "| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {t:.1f} | {100 * t / total:.1f}% |\n")
        f.write(f"\ntotal {total / 1e3:.2f} ms over {len(order)} launches\n")
    print(open(dst).read())


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg"]


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summaries ({src})\n\nper launch; `--clock-control none`\n")
        for r in rows[2:]:
            f.write(f"\n## `{r[ik][:110]}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for h, u, v in zip(hdr, units, r):
                if h in WANT or any(h.endswith(w) and "TriageCompute" in h for w in WANT[4:5]):
                    f.write(f"| {h} | {v} | {u} |\n")
    print(open(dst).read()[:6000])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
