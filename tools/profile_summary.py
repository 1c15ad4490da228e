#!/usr/bin/env python
"""Turn the ncu outputs of a gpurun call into the text summaries kept under profiles/.

  python tools/profile_summary.py full   gpurun_out/prof_X.ncu-rep  "<header line>"  > profiles/rNN_..._full_summary.txt
  python tools/profile_summary.py list   gpurun_out/launches_X.csv  "<header line>"  > profiles/rNN_launch_summary.txt

`full` reads one kernel of an `ncu --set full` report (ncu -i ... --page raw --csv), `list` the
launch list of `ncu --metrics gpu__time_duration.sum --csv`.  CPU only (ncu -i needs no GPU)."""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active"]
STALL = "smsp__average_warps_issue_stalled_"


def full(rep, header):
    # either an .ncu-rep (exported here) or the CSV of `ncu -i X.ncu-rep --page raw --csv` made on the GPU box
    if rep.endswith(".csv"):
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    names, units, vals = rows[0], rows[1], rows[2]
    m = {n: (v, u) for n, u, v in zip(names, units, vals)}
    print(header)
    print("kernel %s" % m.get("Kernel Name", ("?", ""))[0])
    print()
    for k in KEEP:
        if k in m:
            print("%-85s %s %s" % (k, m[k][0], m[k][1]))
    print()
    print("warp stall reasons (%s*_per_issue_active.ratio)" % STALL)
    for n in sorted(m):
        if n.startswith(STALL) and n.endswith("_per_issue_active.ratio") and "not_issued" not in n:
            print("  %-28s %s" % (n[len(STALL):-len("_per_issue_active.ratio")], m[n][0]))
    rd = float(m["dram__bytes_read.sum"][0].replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[m["dram__bytes_read.sum"][1]]
    wr = float(m["dram__bytes_write.sum"][0].replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[m["dram__bytes_write.sum"][1]]
    print()
    print("dram bytes per launch (read + write): %d" % int(rd + wr))


def launches(path, header):
    lines = [l for l in open(path) if not l.startswith("==")]
    rd = csv.DictReader(io.StringIO("".join(lines)))
    agg, order = OrderedDict(), []
    total = 0.0
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r.get("Metric Unit", "ns"), 1e-6)
        name = r["Kernel Name"]
        short = name.split("(")[0][:70]
        c = agg.setdefault(short, [0, 0.0])
        c[0] += 1
        c[1] += ms
        total += ms
        if "lasso_fused" in name:
            order.append(ms)
    print(header)
    print()
    print("%-72s %6s %12s %8s" % ("kernel", "count", "total ms", "share"))
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        print("%-72s %6d %12.3f %8.4f" % (k, n, ms, ms / total))
    print("%-72s %6d %12.3f" % ("all launches", sum(v[0] for v in agg.values()), total))
    print()
    print("lasso_fused launches in order (ms): " + " ".join("%.3f" % v for v in order))


if __name__ == "__main__":
    {"full": full, "list": launches}[sys.argv[1]](sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
