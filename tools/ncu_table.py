"""One line per launch of an `ncu --set full` report: time, grid, registers, occupancy, FP64 pipe, issue rate, DRAM bytes,
L2 hit rate, warp instructions and the three leading stall reasons (warps per issue-active cycle).
usage: python tools/ncu_table.py X.ncu-rep [kernel substring]"""
import csv, subprocess, sys
rep = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ''
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
ix = {k: i for i, k in enumerate(h)}
stall = [k for k in h if k.startswith('smsp__average_warps_issue_stalled_') and k.endswith('_per_issue_active.ratio') and 'not_issued' not in k]
print('# kernel | time | grid | regs | warps active % | FP64 pipe % of active cycles | issue active % | DRAM read | DRAM write | L2 hit % | warp instructions | top stalls')
for r in rows[2:]:
    name = r[ix['Kernel Name']]
    if want not in name:
        continue
    def g(k, fmt='%s'):
        return (fmt % float(r[ix[k]].replace(',', ''))) + ' ' + units[ix[k]] if k in ix and r[ix[k]] else '-'
    st = sorted(((float(r[ix[k]].replace(',', '') or 0), k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]) for k in stall), reverse=True)[:3]
    print(' | '.join([name.split('(')[0].replace('void ', '').replace('slv::', '')[:28].ljust(28), g('gpu__time_duration.sum', '%.1f'), r[ix['launch__grid_size']],
                      r[ix['launch__registers_per_thread']], g('sm__warps_active.avg.pct_of_peak_sustained_active', '%.1f').split()[0],
                      g('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', '%.1f').split()[0],
                      g('smsp__issue_active.avg.pct_of_peak_sustained_active', '%.1f').split()[0], g('dram__bytes_read.sum', '%.2f'), g('dram__bytes_write.sum', '%.2f'),
                      g('lts__t_sector_hit_rate.pct', '%.1f').split()[0], g('smsp__inst_executed.sum', '%.3g').split()[0],
                      ', '.join('%s %.2f' % (n, v) for v, n in st)]))
