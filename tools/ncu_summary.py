"""Summarise an .ncu-rep (first kernel): key counters + top stall sites.  Usage: ncu_summary.py rep [nstall]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; nst = int(sys.argv[2]) if len(sys.argv) > 2 else 14
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = ['Kernel Name', 'gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'sm__cycles_active.avg', 'gpc__cycles_elapsed.avg.per_second',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__pcsamp_warps_issue_stalled_long_scoreboard', 'smsp__pcsamp_warps_issue_stalled_barrier', 'smsp__pcsamp_warps_issue_stalled_wait',
        'smsp__pcsamp_warps_issue_stalled_short_scoreboard', 'smsp__pcsamp_warps_issue_stalled_math_pipe_throttle', 'smsp__pcsamp_warps_issue_stalled_not_selected',
        'smsp__pcsamp_warps_issue_stalled_selected', 'smsp__pcsamp_warps_issue_stalled_branch_resolving', 'smsp__pcsamp_warps_issue_stalled_dispatch_stall',
        'smsp__pcsamp_warps_issue_stalled_mio_throttle', 'smsp__pcsamp_warps_issue_stalled_lg_throttle', 'smsp__pcsamp_warps_issue_stalled_membar',
        'smsp__pcsamp_warps_issue_stalled_sleeping', 'smsp__pcsamp_warps_issue_stalled_no_instructions', 'smsp__pcsamp_warps_issue_stalled_tex_throttle',
        'smsp__pcsamp_warps_issue_stalled_imc_miss', 'smsp__pcsamp_warps_issue_stalled_misc', 'smsp__pcsamp_warps_issue_stalled_drain',
        'smsp__pcsamp_warps_issue_stalled_gmma', 'smsp__pcsamp_sample_buffer_full']
for k in keys:
    if k in d:
        print(f"{k:88s} {d[k][0]:>18s} {d[k][1]}")
if nst:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    if len(rows) > 3:
        data = [r for r in rows[2:] if len(r) > 5 and r[4].isdigit()]
        tot = sum(int(r[4]) for r in data) or 1
        print("total samples", tot)
        for r in sorted(data, key=lambda r: -int(r[4]))[:nst]:
            print(f"{int(r[4]):7d} {int(r[4]) * 100 / tot:5.1f}%  exec={r[5]:>9s}  {r[1].strip()[:100]}")
