#!/usr/bin/env python
"""Side-by-side ncu raw metrics of the codec kernels: python tools/ncu_cmp.py a.csv b.csv ...  (csv = `ncu -i X --page raw --csv`)"""
import csv, sys
WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_sectors.sum", "lts__t_sectors.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"] + [
        f"smsp__average_warps_issue_stalled_{k}_per_issue_active.ratio" for k in
        ("long_scoreboard", "short_scoreboard", "wait", "math_pipe_throttle", "lg_throttle", "no_instruction", "not_selected", "branch_resolving",
         "dispatch_stall", "mio_throttle", "barrier", "drain", "imc_miss", "sleeping", "selected", "membar", "tex_throttle")]
kern = sys.argv[1].split(",") if "," in sys.argv[1] or not sys.argv[1].endswith(".csv") else ["encode_kernel", "decode_kernel"]
files = [a for a in sys.argv[1:] if a.endswith(".csv")]
tabs = {}
for p in files:
    rows = list(csv.reader(open(p)))
    hdr, rows = rows[0], rows[2:]
    ik = hdr.index("Kernel Name")
    for r in rows:
        for k in kern:
            if k in r[ik] and (p, k) not in tabs:
                tabs[(p, k)] = {h: r[i] for i, h in enumerate(hdr)}
for k in kern:
    print("====", k, " | ".join(files))
    for w in WANT:
        vals = []
        for p in files:
            t = tabs.get((p, k), {})
            c = [h for h in t if h == w or h.endswith("." + w)]
            vals.append(t[c[0]] if c else "-")
        print(f"  {w[:78]:78s} " + " ".join(f"{v:>16s}" for v in vals))
