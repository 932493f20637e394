#!/usr/bin/env python3
"""Compact summary of an .ncu-rep (raw page): python tools/ncu_summary.py report.ncu-rep [out.txt]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name','Block Size','Grid Size','gpu__time_duration.sum','launch__registers_per_thread','launch__occupancy_limit_registers',
 'sm__warps_active.avg.pct_of_peak_sustained_active','smsp__warps_active.avg.per_cycle_active','smsp__warps_eligible.avg.per_cycle_active',
 'smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__thread_inst_executed.sum',
 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','dram__bytes_read.sum','dram__bytes_write.sum',
 'l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed','sm__icc_request_hit_rate.pct']
out = []
for r in rows[2:]:
    for k in keys:
        if k in hdr:
            i = hdr.index(k); out.append("%-82s %-16s %s" % (k, units[i], r[i]))
    out.append("-- stalls (warps per issue-active cycle):")
    st = []
    for i, h in enumerate(hdr):
        if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio') and 'not_issued' not in h:
            try: st.append((float(r[i]), h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]))
            except ValueError: pass
    for v, n in sorted(st, reverse=True)[:10]:
        out.append("   %-28s %.3f" % (n, v))
    out.append("-- instruction mix (executed, top):")
    mix = []
    for i, h in enumerate(hdr):
        if h.startswith('sass__inst_executed_per_opcode') or h.startswith('smsp__sass_inst_executed_op'):
            pass
    out.append("")
txt = "\n".join(out)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt + "\n")
