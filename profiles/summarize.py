#!/usr/bin/env python
"""Summarise an ncu report (.ncu-rep) into a small text file that can be committed.
Usage: python profiles/summarize.py gpurun_out/prof.ncu-rep profiles/NAME.txt [launches.csv]"""
import csv
import io
import subprocess
import sys

KEYS = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
    'lts__t_sectors_op_read.sum', 'lts__t_bytes.sum',
    'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct',
    'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
    'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
    'l1tex__lsu_writeback_active_mem_lgds.sum.pct_of_peak_sustained_elapsed',
    'l1tex__m_xbar2l1tex_read_bytes.sum', 'lts__t_sectors_srcunit_tex_op_read.sum',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fp64.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
    'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
    'smsp__average_warp_latency_issue_stalled_long_scoreboard.pct',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu --set full summary of {rep}"]
    name_i = hdr.index('Kernel Name')
    for k, row in enumerate(rows[2:]):
        lines.append(f"\n## launch {k}: {row[name_i][:100]}")
        for key in KEYS:
            if key in hdr:
                i = hdr.index(key)
                lines.append(f"{key:85s} {row[i]:>18s} {units[i]}")
    if len(sys.argv) > 3:
        lines.append(f"\n# launch list ({sys.argv[3]}): kernel, gpu__time_duration.sum [ns]")
        lrows = [r for r in csv.reader(open(sys.argv[3])) if len(r) > 5]
        ik, iv = lrows[0].index('Kernel Name'), lrows[0].index('Metric Value')
        for r in lrows[1:]:
            lines.append(f"{r[ik][:80]:80s} {r[iv]:>12s}")
    open(out, 'w').write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == '__main__':
    main()
