"""Condense ncu outputs brought back in gpurun_out/ into small text files for profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv  > profiles/<name>_launches.txt
    python tools/ncu_summary.py full gpurun_out/prof.ncu-rep       > profiles/<name>_full.txt
"""
import collections
import csv
import subprocess
import sys

FULL_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
]


def us(value, unit):
    v = float(value.replace(",", ""))
    return {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(unit, v)


def launches(path):
    with open(path) as f:
        rows = list(csv.DictReader(l for l in f if not l.startswith("==")))
    agg = collections.OrderedDict()
    for r in rows:
        agg.setdefault(r["Kernel Name"], []).append(us(r["Metric Value"], r["Metric Unit"]))
    tot = sum(sum(v) for v in agg.values())
    print(f"# {len(rows)} launches, {tot / 1e3:.2f} ms total device time (ncu: cold cache, serialised)")
    print(f"# {'total_us':>10s} {'share':>6s} {'n':>5s} {'avg_us':>9s}  kernel")
    for name, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"  {sum(v):10.1f} {100 * sum(v) / tot:5.1f}% {len(v):5d} {sum(v) / len(v):9.1f}  {name[:110]}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    for d in data:
        print("kernel:", d[col["Kernel Name"]][:120])
        for m in FULL_METRICS:
            if m in col:
                print(f"    {m:72s} {d[col[m]]:>18s} {units[col[m]]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
