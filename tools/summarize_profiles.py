"""Turn gpurun_out/{launches.csv, prof_*.ncu-rep} into the committed summaries under profiles/.

    python tools/summarize_profiles.py r01
"""
import csv
import io
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
GP = os.path.join(ROOT, "gpurun_out")

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "smsp__inst_executed.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_uniform", "smsp__cycles_active.avg"]


def launches(tag, csv_name="launches.csv", out_name="launches_summary", command="python bench.py --views 1 --steps 1 --warmup 3 --no-cpu-baseline --no-config5"):
    src = os.path.join(GP, csv_name)
    if not os.path.exists(src):
        return
    rows = []
    with open(src) as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(io.StringIO("".join(lines))):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            unit = r.get("Metric Unit", "ns")
            v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1e-3)
            rows.append((r["Kernel Name"].split("(")[0], v))
    agg = defaultdict(lambda: [0, 0.0])
    for k, v in rows:
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v for _, v in rows)
    with open(os.path.join(OUT, f"{tag}_{out_name}.md"), "w") as f:
        f.write(f"# Launch list summary ({tag})\n\nncu --metrics gpu__time_duration.sum --clock-control none, command: "
                f"`{command}` (cold-cache, serialised: compare shares).\n\n"
                f"{len(rows)} launches, {tot / 1e3:.2f} ms total\n\n| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {k} | {n} | {t:.1f} | {100 * t / tot:.1f}% | {t / n:.1f} |\n")
    print("wrote launches summary:", len(rows), "launches")


def full(tag, name):
    rep = os.path.join(GP, f"{name}.ncu-rep")
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rd[0], rd[1], rd[2:]
    with open(os.path.join(OUT, f"{tag}_{name}_raw.md"), "w") as f:
        f.write(f"# ncu --set full --clock-control none ({tag}, {name})\n\n")
        for row in vals:
            f.write(f"## {row[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else ''}\n\n| metric | unit | value |\n|---|---|---|\n")
            for i, h in enumerate(hdr):
                if any(h == k or h.startswith(k) for k in KEYS) or "tensor" in h or "dram__bytes" in h or "smsp__average_warps_issue_stalled" in h \
                        or "smsp__warp_issue_stalled" in h or h.startswith("launch__"):
                    f.write(f"| {h} | {units[i]} | {row[i]} |\n")
            f.write("\n")
    print("wrote", name)


def traffic(name="prof_mlp_tc_fused_fine", rows=40000 * 192):
    """profiles/mlp_tc_traffic.json (read by bench.py: roofline.traffic) from the full capture of the fused fine launch."""
    import json
    rep = os.path.join(GP, f"{name}.ncu-rep")
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(raw)))
    hdr, units, row = rd[0], rd[1], rd[2]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

    def val(metric):
        i = hdr.index(metric)
        return float(row[i].replace(",", "")) * scale[units[i]]
    rd_b, wr_b = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    json.dump({"bytes_per_row": round((rd_b + wr_b) / rows, 3),
               "source": f"ncu --set full --clock-control none of the FUSED fine-pass k_mlp_tc<2,false> launch on 40 000 rays "
                         f"(profiles/r02_{name}_raw.md): ({rd_b / 1e6:.2f} MB dram read + {wr_b / 1e6:.2f} MB dram written) / {rows} rows; "
                         f"algorithmic 4 B (depth) + 44/192 B (ray) in, 52/192 B out per row"},
              open(os.path.join(OUT, "mlp_tc_traffic.json"), "w"))
    print("wrote mlp_tc_traffic.json")


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    os.makedirs(OUT, exist_ok=True)
    launches(tag)
    launches(tag, "launches_train.csv", "launches_train_summary", "python tests/tools/train_target.py  (3 SSR training steps: render + backward, 1024 rays)")
    launches(tag, "launches_train_warm.csv", "launches_train_warm_summary",
             "python tests/tools/train_target.py  (the same three steps under --cache-control none: L2 keeps what the previous kernel left)")
    for n in sorted(os.listdir(GP)):
        if n.endswith(".ncu-rep"):
            full(tag, n[:-8])
    traffic()
