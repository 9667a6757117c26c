#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into profiles/ (text, committed).

  python tools/ncu_summary.py launches <launches.csv> <out.md>     # per-kernel time shares
  python tools/ncu_summary.py full <file.ncu-rep> <out.md>         # key counters of a --set full capture
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_bytes.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fmaheavy.sum", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum"]


def launches(src, dst):
    rows = []
    with open(src, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(io.StringIO("".join(lines)))
    for r in rd:
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            unit = r.get("Metric Unit", "ns")
            scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}.get(unit, 1e-6)
            rows.append((r["Kernel Name"], v * scale))
    agg = defaultdict(lambda: [0, 0.0])
    for k, ms in rows:
        k = k.split("(")[0]
        agg[k][0] += 1
        agg[k][1] += ms
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({src}): {len(rows)} launches, {tot:.2f} ms total (cold-cache, serialised: compare SHARES)\n\n")
        f.write("| kernel | launches | total ms | avg ms | share |\n|---|---|---|---|---|\n")
        for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {ms:.3f} | {ms / n:.4f} | {100 * ms / tot:.1f} % |\n")
    print(open(dst).read())


def full(src, dst):
    """src: an .ncu-rep, or the `ncu -i <rep> --page raw --csv` text exported on the GPU box (*.csv)"""
    if src.endswith(".csv"):
        out = open(src).read()
    else:
        out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rd[0], rd[1], rd[2:]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary of {src}\n\n")
        for row in data:
            d = dict(zip(hdr, row))
            f.write(f"## {d.get('Kernel Name', '?')[:100]}  (id {d.get('ID')})\n\n| metric | value | unit |\n|---|---|---|\n")
            for k in KEYS:
                for hname in hdr:                       # newer ncu prefixes some columns with a section name
                    if hname == k or hname.endswith("." + k):
                        if d[hname] != "":
                            f.write(f"| {k} | {d[hname]} | {units[hdr.index(hname)]} |\n")
                        break
            f.write("\n")
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
