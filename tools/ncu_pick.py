"""Print the handful of ncu raw-page metrics that decide what bounds a kernel (one column per captured launch)."""
import csv
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed.sum', 'smsp__inst_executed.avg.per_cycle_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__grid_size',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warp_latency_per_inst_issued.ratio']


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    recs = [dict(zip(hdr, r)) for r in rows[2:]]
    u = dict(zip(hdr, units))
    print('%-84s' % 'metric', *['%18s' % r['Kernel Name'].split('(')[0][-18:] for r in recs])
    for k in KEYS:
        if k in hdr:
            print('%-84s' % (k + ' [' + u[k] + ']'), *['%18s' % r[k] for r in recs])


if __name__ == '__main__':
    main(sys.argv[1])
