"""Selected raw metrics of one ncu report as text (profiling helper).  usage: tools_ncu_metrics.py <report.ncu-rep>"""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))   # repo root (bench.py, sbayes_b200/, oracle/)
import csv, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'sm__cycles_elapsed.avg ', 'dram__bytes_read.sum ', 'dram__bytes_write.sum ', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum ', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'smsp__inst_executed.sum ', 'smsp__issue_active.avg.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_tensor', 'smsp__average_warps_issue_stalled',
        'launch__registers_per_thread ', 'launch__occupancy_limit', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'smsp__inst_executed_op_shared_atom', 'sm__inst_executed_pipe_fp64', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum ', 'l1tex__t_sector_hit_rate']
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u = rows[0], rows[1]
for r in rows[2:]:
    print("== kernel:", r[h.index("Kernel Name")] if "Kernel Name" in h else "?")
    for i, k in enumerate(h):
        if any(x in k + ' ' for x in KEYS):
            try:
                if 'stalled' in k and float(r[i]) < 0.25: continue
            except ValueError:
                pass
            print(f"{k:95s} {r[i]:>18s} {u[i]}")
