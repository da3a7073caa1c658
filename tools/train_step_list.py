"""Per-kernel durations of the LAST of the three training steps in gpurun_out/launches_train_warm.csv (library kernels
listed one by one, PyTorch's summed)."""
import csv
import io
import os
import sys

src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "launches_train_warm.csv")
lines = [l for l in open(src) if not l.startswith("==")]
rows = [r for r in csv.DictReader(io.StringIO("".join(lines))) if r.get("Metric Name") == "gpu__time_duration.sum"]
n = len(rows) // 3
tot = lib = 0.0
agg = {}
for r in rows[2 * n:3 * n]:
    k = r["Kernel Name"].split("(")[0][:60]
    v = float(r["Metric Value"].replace(",", "")) / 1e3
    tot += v
    if "at::" in k or "unnamed" in k:
        continue
    lib += v
    a = agg.setdefault(k, [0, 0.0, []])
    a[0] += 1
    a[1] += v
    a[2].append(round(v, 1))
for k, (c, v, l) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:44s} x{c:2d} {v:8.1f} us  {l if c <= 4 else l[:3] + ['...'] + l[-2:]}")
print(f"step: {tot:.1f} us of kernels ({n} launches), library {lib:.1f} us")
