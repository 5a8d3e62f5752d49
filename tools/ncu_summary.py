"""Summarise an .ncu-rep (read here, no GPU):  python tools/ncu_summary.py rep [out.md]"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = [
 'gpu__time_duration.sum', 'sm__cycles_active.avg', 'smsp__inst_executed.sum',
 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
 'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__warps_active.avg.pct_of_peak_sustained_active',
 'launch__registers_per_thread', 'launch__shared_mem_per_block_static', 'launch__occupancy_limit_shared_mem',
 'launch__occupancy_limit_registers', 'launch__occupancy_limit_warps',
 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_xu.sum',
]
stalls = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
out = []
for r in rows[2:]:
    out.append(f"### {r[idx['Kernel Name']][:90]}  grid={r[idx['Grid Size']]} block={r[idx['Block Size']]}")
    for w in want:
        if w in idx:
            out.append(f"- {w}: {r[idx[w]]} {units[idx[w]]}")
    st = sorted(((float(r[idx[s]] or 0), s) for s in stalls), reverse=True)[:7]
    out.append("- top stalls (warps per issue-active): " + ", ".join(f"{s.split('stalled_')[1].split('_per_issue')[0]}={v:.2f}" for v, s in st))
    out.append("")
text = "\n".join(out)
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text + "\n")
