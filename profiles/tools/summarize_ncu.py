"""Turn gpurun_out/*.ncu-rep (ncu --set full) into the tracked text summaries under profiles/ and profiles/traffic.json.

    python profiles/tools/summarize_ncu.py r1 c2 c3 c4 c5     # reads gpurun_out/r1_<cfg>.ncu-rep
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
WANT = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
    'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
]


def to_bytes(value, unit):
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit)
    return float(value) * scale if scale else None


def to_us(value, unit):
    scale = {'ns': 1e-3, 'us': 1, 'ms': 1e3, 's': 1e6}.get(unit)
    return float(value) * scale if scale else None


def main():
    tag, cfgs = sys.argv[1], sys.argv[2:]
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}
    for cfg in cfgs:
        rep = os.path.join(ROOT, 'gpurun_out', f'{tag}_{cfg}.ncu-rep')
        out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units, data = rows[0], rows[1], rows[2:]
        ik = hdr.index('Kernel Name')
        lines = [f'# ncu --set full --clock-control none, {os.path.basename(rep)}: python bench.py --config {cfg} --steps 3 --warmup 3 '
                 f'--no-cpu-baseline --no-e2e (default batches_per_launch); per-launch values, cold-cache and serialised']
        dominant = None
        for r in data:
            lines.append(f'== {r[ik]}')
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    lines.append(f'   {w:86s} {r[i]:>16s} {units[i]}')
            rd = to_bytes(r[hdr.index('dram__bytes_read.sum')], units[hdr.index('dram__bytes_read.sum')])
            wr = to_bytes(r[hdr.index('dram__bytes_write.sum')], units[hdr.index('dram__bytes_write.sum')])
            us = to_us(r[hdr.index('gpu__time_duration.sum')], units[hdr.index('gpu__time_duration.sum')])
            lines.append(f'   -> DRAM traffic {(rd + wr) / 1e6:.1f} MB in {us:.1f} us = {(rd + wr) / us / 1e3:.0f} GB/s')
            if dominant is None or us > dominant[1]:
                dominant = (r[ik], us, rd + wr)
        open(os.path.join(ROOT, 'profiles', f'{tag}_{cfg}_ncu_summary.txt'), 'w').write('\n'.join(lines) + '\n')
        sys.path.insert(0, ROOT)
        import bench
        from ogbench_b200 import synthetic
        w = synthetic.WORKLOADS[cfg]
        traffic[cfg] = {'kernel': dominant[0], 'dram_bytes_per_launch': dominant[2], 'duration_us_under_ncu': dominant[1],
                        'transitions_per_launch': w.batch * bench.default_batches_per_launch(w),
                        'algorithmic_bytes_per_launch': w.bytes_per_transition * w.batch * bench.default_batches_per_launch(w),
                        'source': os.path.basename(rep)}
        print(cfg, dominant)
    json.dump(traffic, open(tpath, 'w'), indent=1)


if __name__ == '__main__':
    main()
