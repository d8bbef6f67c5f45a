#!/usr/bin/env python
"""Summarise one kernel of an .ncu-rep (raw page) into the handful of numbers DESIGN.md quotes."""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    g = lambda k: d.get(k, "n/a")
    print("kernel:", g("Kernel Name")[:120])
    print("  grid x block        :", g("launch__grid_size"), "x", g("launch__block_size"), " regs/thread", g("launch__registers_per_thread"),
          " smem/block(dyn)", g("launch__shared_mem_per_block_dynamic"))
    print("  duration            :", g("gpu__time_duration.sum"), units[hdr.index("gpu__time_duration.sum")])
    print("  dram read / write   :", g("dram__bytes_read.sum"), units[hdr.index("dram__bytes_read.sum")], "/", g("dram__bytes_write.sum"),
          units[hdr.index("dram__bytes_write.sum")])
    print("  dram throughput     :", g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), "% of peak;",
          g("dram__bytes.sum.per_second"), units[hdr.index("dram__bytes.sum.per_second")] if "dram__bytes.sum.per_second" in hdr else "")
    print("  sm clock            :", g("sm__cycles_elapsed.avg.per_second"), units[hdr.index("sm__cycles_elapsed.avg.per_second")])
    print("  warps active        :", g("sm__warps_active.avg.pct_of_peak_sustained_active"), "% of peak")
    print("  issue active        :", g("smsp__issue_active.avg.pct_of_peak_sustained_active"), "%   eligible warps/cycle", g("smsp__warps_eligible.avg.per_cycle_active"))
    print("  warp instr executed :", g("smsp__inst_executed.sum"), "  threads/instr", g("smsp__thread_inst_executed_per_inst_executed.ratio"))
    for pipe in ("alu", "fma", "fp64", "xu", "lsu", "uniform", "adu", "cbu"):
        k = "sm__inst_executed_pipe_%s.avg.pct_of_peak_sustained_active" % pipe
        if k in d:
            print("  pipe %-8s       : %s %%" % (pipe, d[k]))
    st = []
    for k, v in d.items():
        if "pcsamp_warps_issue_stalled" in k and not k.endswith("_not_issued"):
            try:
                st.append((float(v.replace(",", "")), k.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError:
                pass
    tot = sum(v for v, _ in st) or 1.0
    print("  stall samples       :", ", ".join("%s %.1f%%" % (k, 100 * v / tot) for v, k in sorted(st, reverse=True)[:7]))
