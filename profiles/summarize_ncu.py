"""Turns an .ncu-rep (brought back in gpurun_out/) into the small text summary
kept under profiles/: python profiles/summarize_ncu.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__cycles_active.avg']


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H, U = rows[0], rows[1]
    for V in rows[2:]:
        print('kernel:', V[H.index('Kernel Name')])
        for w in WANT:
            if w in H:
                i = H.index(w)
                print('  %-70s %s %s' % (w, V[i], U[i]))
        stalls = [(float(V[i]), H[i]) for i in range(len(H))
                  if H[i].startswith('smsp__average_warps_issue_stalled_') and H[i].endswith('_per_issue_active.ratio') and V[i]]
        for v, n in sorted(stalls, reverse=True)[:6]:
            print('  stall %-64s %.3f' % (n.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v))


if __name__ == '__main__':
    main(sys.argv[1])
