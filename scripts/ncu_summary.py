"""Text summary of one kernel launch out of an ncu report (the files committed under profiles/).

    python scripts/ncu_summary.py gpurun_out/r2_prof_msd.ncu-rep "<what was run>" > profiles/r2_ncu_....txt
"""
import csv
import subprocess
import sys

WANT = (
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
    'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes_read.sum.per_second',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
    'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
)


def main():
    rep, what = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ''
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print('# ncu --set full --clock-control none --import-source on ...  %s' % what)
        print('%-88s %s' % ('Kernel Name', vals[hdr.index('Kernel Name')]))
        for name in WANT:
            if name in hdr:
                i = hdr.index(name)
                print('%-88s %-16s %s' % (name, units[i], vals[i]))
        for i, name in enumerate(hdr):
            if 'issue_stalled' in name and name.endswith('per_issue_active.ratio'):
                print('%-88s %-16s %s' % (name, units[i], vals[i]))
        print()


if __name__ == '__main__':
    main()
