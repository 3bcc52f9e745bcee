#!/usr/bin/env python
"""Summarise ncu outputs into small text files for profiles/ (run in the build container; needs `ncu` on PATH).

    python scripts/ncu_summarize.py launches <launches.csv>            # per-kernel totals / shares of one launch list
    python scripts/ncu_summarize.py full <report.ncu-rep> [metric ...]  # selected raw metrics per captured launch
    python scripts/ncu_summarize.py traffic <launches.csv> <out.json>  # per-kernel DRAM bytes per launch (dram__bytes_read/write.sum)
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

DEFAULT_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg",
]


def short(name):
    name = name.replace("lmrl::", "")
    return name.split("(")[0][:70]


def launches(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.DictReader(io.StringIO("".join(lines)))
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        rows.append((short(r["Kernel Name"]), v * scale))
    tot = sum(v for _, v in rows)
    by = defaultdict(lambda: [0.0, 0])
    for k, v in rows:
        by[k][0] += v
        by[k][1] += 1
    print(f"launches {len(rows)}  total kernel time {tot:.1f} ms (ncu: serialised, cold cache; compare SHARES)")
    for k, (ms, n) in sorted(by.items(), key=lambda kv: -kv[1][0]):
        print(f"{100 * ms / tot:6.2f}%  {ms:10.3f} ms  n={n:5d}  avg {1000 * ms / n:9.1f} us  {k}")


def full(path, metrics):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rd[0], rd[1], rd[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [m for m in metrics if m in idx]
    print("# " + path)
    for r in data:
        name = short(r[idx["Kernel Name"]])
        print(f"[{r[idx['ID']]}] {name}")
        for m in cols:
            print(f"      {m:75s} {r[idx[m]]:>18s} {units[idx[m]]}")


def traffic(path, out_path):
    """Launch list captured with --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -> per kernel family the
    mean DRAM bytes (read + write) and mean duration per launch; bench.py reads the JSON for roofline.traffic."""
    import json
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    by = defaultdict(lambda: dict(launches=0, read=0.0, write=0.0, ms=0.0))
    scale_b = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    scale_t = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
    for r in csv.DictReader(io.StringIO("".join(lines))):
        k, m = short(r["Kernel Name"]), r.get("Metric Name")
        v = float(r["Metric Value"].replace(",", ""))
        if m == "gpu__time_duration.sum":
            by[k]["ms"] += v * scale_t.get(r.get("Metric Unit", "ns"), 1e-6)
            by[k]["launches"] += 1
        elif m == "dram__bytes_read.sum":
            by[k]["read"] += v * scale_b.get(r.get("Metric Unit", "byte"), 1.0)
        elif m == "dram__bytes_write.sum":
            by[k]["write"] += v * scale_b.get(r.get("Metric Unit", "byte"), 1.0)
    out = {}
    for k, d in by.items():
        n = max(d["launches"], 1)
        out[k] = dict(launches=d["launches"], dram_bytes_per_launch=(d["read"] + d["write"]) / n, dram_read_per_launch=d["read"] / n,
                      dram_write_per_launch=d["write"] / n, us_per_launch=1000.0 * d["ms"] / n)
    fam = defaultdict(lambda: dict(launches=0, bytes=0.0))
    for k, d in out.items():
        f = "gemm_bf16_tcgen05_kernel" if "gemm_bf16_tcgen05" in k else ("attention_hpw_kernel" if "attention_hpw_kernel" in k else k)
        fam[f]["launches"] += d["launches"]
        fam[f]["bytes"] += d["dram_bytes_per_launch"] * d["launches"]
    summary = {f: dict(launches=d["launches"], dram_bytes_per_launch=d["bytes"] / max(d["launches"], 1)) for f, d in fam.items()}
    json.dump(dict(source=path, kernels=out, families=summary), open(out_path, "w"), indent=1)
    for f, d in sorted(summary.items(), key=lambda kv: -kv[1]["dram_bytes_per_launch"] * kv[1]["launches"]):
        print(f"{f:60s} n={d['launches']:5d}  {d['dram_bytes_per_launch'] / 1e6:10.2f} MB/launch")


if __name__ == "__main__":
    if len(sys.argv) < 3 or sys.argv[1] in ("-h", "--help"):
        print(__doc__)
        sys.exit(0 if len(sys.argv) > 1 and sys.argv[1] in ("-h", "--help") else 2)
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    elif sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3:] or DEFAULT_METRICS)
