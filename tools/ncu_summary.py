"""Print the handful of ncu metrics the roofline discussion uses from a .ncu-rep (raw page)."""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum ', 'dram__bytes_write.sum ', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread ', 'launch__occupancy_limit', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum ',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ',
        'l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg', 'lts__t_sectors.sum ', 'lts__throughput.avg.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__average_warps_issue_stalled', 'smsp__average_warp_latency_per_inst_issued', 'local', 'launch__shared_mem_per_block_dynamic',
        'sm__inst_executed_pipe_lsu', 'smsp__inst_executed_pipe_fp64', 'smsp__cycles_active.avg ']


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print("== kernel:", vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '')
        for h, u, v in zip(hdr, units, vals):
            if any(k.strip() in h for k in KEYS) and v not in ('', '0'):
                print(f"{h:86s} {u:14s} {v}")


if __name__ == '__main__':
    main(sys.argv[1])
