"""Print the per-kernel durations (ms) of the LAST of n repeated iterations in an ncu
`--metrics gpu__time_duration.sum --csv` log.  usage: launch_times.py log.csv [n_iters] [min_ms]"""
import csv, sys
path = sys.argv[1]; iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
min_ms = float(sys.argv[3]) if len(sys.argv) > 3 else 0.1
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = []
for x in csv.DictReader(lines):
    if x.get("Metric Name") == "gpu__time_duration.sum":
        scale = {"ns": 1e6, "us": 1e3, "ms": 1.0}[x["Metric Unit"]]
        rows.append((x["Kernel Name"][:48], x["Grid Size"], float(x["Metric Value"].replace(",", "")) / scale))
per = len(rows) // iters
for k in rows[-per:]:
    if k[2] >= min_ms:
        print(f"{k[2]:9.3f} ms  {k[1]:>16}  {k[0]}")
