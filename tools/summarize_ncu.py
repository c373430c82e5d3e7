"""profiles/<tag>_ncu_summary.md + an entry in profiles/traffic_r2.json from an `ncu --set full` report.
usage: python tools/summarize_ncu.py gpurun_out/prof.ncu-rep <tag> <workload key: 4k|1080p|512|1080p_srgb8> <pairs_per_launch> ["description"]
The traffic entry is keyed to a hash of the kernel sources (turbo_metrics_b200/csrc/{ssimu2_kernels.cuh,exact_math.cuh}) AS THEY ARE when
this script runs -- run it right after the capture, before editing the kernels; bench.py flags a stale entry."""
import csv, hashlib, io, json, os, subprocess, sys
rep, tag, workload, ppl = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
M = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
     "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
     "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__occupancy_limit_registers",
     "launch__occupancy_limit_shared_mem", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
     "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
     "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
     "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg.per_second",
     "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "smsp__warps_eligible.avg.per_cycle_active",
     "smsp__warps_active.avg.per_cycle_active"]
ki = hdr.index("Kernel Name")
desc = sys.argv[5] if len(sys.argv) > 5 else workload
out = [f"# ncu --set full, {tag}: {desc}, {ppl:g} frame pairs per launch", "",
       f"source report: `{os.path.basename(rep)}` (gpurun scratch, not committed); command: "
       "`ncu --set full --clock-control none --import-source on -k regex:k_ -s <warm-up> -c 3 python tools/quick_time.py ...`", "",
       "| metric | unit | " + " | ".join(d[ki].split("(")[0].replace("void ", "") for d in data) + " |", "|---|---|" + "---|" * len(data)]
traffic = {}
for m in M:
    if m in hdr:
        i = hdr.index(m)
        out.append(f"| `{m}` | {units[i]} | " + " | ".join(d[i] for d in data) + " |")
for d in data:
    name = d[ki].split("(")[0].replace("void ", "").split("<")[0]
    r, w = float(d[hdr.index("dram__bytes_read.sum")]), float(d[hdr.index("dram__bytes_write.sum")])
    ur, uw = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_write.sum")]
    sc = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    traffic[name] = (r * sc[ur] + w * sc[uw]) / ppl
out += ["", "DRAM bytes per frame pair (read + write): " + ", ".join(f"`{k}` {v/1e6:.1f} MB" for k, v in traffic.items()),
        f"sum {sum(traffic.values())/1e6:.1f} MB"]
open(os.path.join(root, "profiles", f"{tag}_ncu_summary.md"), "w").write("\n".join(out) + "\n")
hh = hashlib.sha256()
for f in ("ssimu2_kernels.cuh", "exact_math.cuh"):
    hh.update(open(os.path.join(root, "turbo_metrics_b200", "csrc", f), "rb").read())
tf = os.path.join(root, "profiles", "traffic_r2.json")
allt = json.load(open(tf)) if os.path.exists(tf) else {}
allt[workload] = {"dram_bytes_per_pair": traffic, "pairs_per_launch": ppl, "report": os.path.basename(rep), "summary": f"{tag}_ncu_summary.md",
                  "source_hash": hh.hexdigest()[:16]}
json.dump(allt, open(tf, "w"), indent=1)
print("\n".join(out[-3:]))
