"""Key metrics of every kernel in an ncu report (raw page). usage: python tools/ncu_key.py rep.ncu-rep [kernel-substr]"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ki = hdr.index("Kernel Name")
sel = sys.argv[2] if len(sys.argv) > 2 else ""
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_ld.sum",
        "l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second",
        "launch__registers_per_thread", "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active"]
stalls = [h for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
for d in rows[2:]:
    if sel not in d[ki]:
        continue
    print("==", d[ki][:60])
    for w in want:
        if w in hdr:
            print(f"   {w:75s} {d[hdr.index(w)]:>16s} {units[hdr.index(w)]}")
    st = sorted(((float(d[hdr.index(h)] or 0), h[len('smsp__pcsamp_warps_issue_stalled_'):]) for h in stalls), reverse=True)
    tot = sum(v for v, _ in st) or 1
    print("   stalls: " + "  ".join(f"{n}={100*v/tot:.1f}%" for v, n in st[:10]))
