"""Turns what a gpurun call left in gpurun_out/ into the tracked summaries under profiles/:

    python tools/summarize_profiles.py launches gpurun_out/launches_X.csv profiles/X_launch_summary.csv "title"
    python tools/summarize_profiles.py ncu gpurun_out/K.ncu-rep profiles/X_kernel   (writes _ncu_details.csv and _ncu_raw_selected.csv)
"""
import collections
import csv
import io
import subprocess
import sys

KEEP = ("duration", "pipe_fma", "pipe_alu", "pipe_xu", "pipe_fp64", "pipe_lsu", "pipe_uniform", "inst_executed.sum", "inst_issued",
        "ipc", "issue_active", "warp_issue_stalled", "dram__bytes", "dram__throughput", "lts__t_sector_hit_rate", "lts__throughput",
        "l1tex__data_bank_conflicts", "l1tex__throughput", "registers_per_thread", "shared_mem", "occupancy", "warps_active",
        "sm__throughput", "smsp__cycles_active", "launch__grid_size", "launch__block_size", "thread_inst_executed", "tensor")


def launches(src, dst, title):
    rows = [r for r in csv.reader(l for l in open(src) if not l.startswith("==")) if r]
    head = rows[0]
    ik, iv, im = head.index("Kernel Name"), head.index("Metric Value"), head.index("Metric Name")
    tot = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        k = r[ik][:60]
        n, t = tot.get(k, (0, 0.0))
        tot[k] = (n + 1, t + float(r[iv].replace(",", "")))
    s = sum(t for _, t in tot.values())
    with open(dst, "w") as f:
        f.write(f"# {title}\n# per-launch times are cold-cache and serialised: compare SHARES, not absolutes (units: ns)\n")
        f.write("kernel,launches,total_ns,share_pct\n")
        for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{k}\",{n},{t:.1f},{100 * t / s:.2f}\n")


def ncu(rep, prefix):
    det = subprocess.run(["ncu", "-i", rep, "--page", "details", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(det)))
    h = rows[0]
    with open(prefix + "_ncu_details.csv", "w") as f:
        w = csv.writer(f)
        w.writerow(["section", "metric", "unit", "value"])
        for r in rows[1:]:
            d = dict(zip(h, r))
            if d.get("Metric Name"):
                w.writerow([d.get("Section Name"), d["Metric Name"], d.get("Metric Unit"), d.get("Metric Value")])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    names, units, vals = rows[0], rows[1], rows[2]
    with open(prefix + "_ncu_raw_selected.csv", "w") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit", "value"])
        for n, u, v in zip(names, units, vals):
            if "breakdown" in n or "Triage" in n or n.startswith("device__") or "source__" in n or ".max" in n or ".min" in n:
                continue
            if n in ("Kernel Name", "Block Size", "Grid Size") or any(k in n for k in KEEP):
                w.writerow([n, u, v])


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        ncu(sys.argv[2], sys.argv[3])
