#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into a small text file for profiles/.

    python tools/ncu_summary.py gpurun_out/prof_warp.ncu-rep profiles/r01_warp_full.txt [traffic_key]

Keeps the metrics B200_PROFILING.md names (dram bytes, dram %, duration, registers, warps
active, pipe use, stall reasons) for every captured launch; optionally records
dram__bytes_read+write per launch in profiles/traffic.json under `traffic_key` (bench.py
reports it as roofline.traffic)."""
import csv
import io
import json
import os
import subprocess
import sys

KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__cycles_active.avg',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__waves_per_multiprocessor',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio']

UNIT = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'Tbyte': 1e12}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    key = sys.argv[3] if len(sys.argv) > 3 else None
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = ['# ncu --set full --clock-control none summary of %s' % os.path.basename(rep),
             '# (per-launch values are cold-cache and serialised by ncu replay; see B200_PROFILING.md)']
    traffic = None
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')]
        lines.append('\nkernel: %s' % name)
        rd = wr = None
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                lines.append('  %-86s %s %s' % (k, r[i], units[i]))
                if k == 'dram__bytes_read.sum':
                    rd = float(r[i]) * UNIT.get(units[i], 1.0)
                if k == 'dram__bytes_write.sum':
                    wr = float(r[i]) * UNIT.get(units[i], 1.0)
        if rd is not None and wr is not None:
            lines.append('  %-86s %.0f byte' % ('dram traffic (read+write) per launch', rd + wr))
            if traffic is None:
                traffic = rd + wr
    os.makedirs(os.path.dirname(out) or '.', exist_ok=True)
    open(out, 'w').write('\n'.join(lines) + '\n')
    if key and traffic is not None:
        tj = os.path.join(os.path.dirname(out) or '.', 'traffic.json')
        d = json.load(open(tj)) if os.path.exists(tj) else {}
        d[key] = traffic
        json.dump(d, open(tj, 'w'), indent=1, sort_keys=True)
    print('\n'.join(lines))


if __name__ == '__main__':
    main()
