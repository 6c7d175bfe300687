#!/usr/bin/env python
"""profiles/ncu_traffic.json from an `ncu --set full` capture: DRAM bytes per frame of a kernel.

  python tools/ncu_traffic.py <report.ncu-rep> <config> <kernel substring> <frames per launch> [note]

Reads dram__bytes_read.sum + dram__bytes_write.sum of every captured launch whose name contains the substring (averaged),
divides by the frames one launch processed and stores it under config<N>/<kernel> -- bench.py's roofline.traffic scales
the figure to its own launch.  Also prints a one-line metric summary for profiles/*.txt."""
import csv, json, os, subprocess, sys
rep, cfg, kern, frames = sys.argv[1], int(sys.argv[2]), sys.argv[3], float(sys.argv[4])
note = sys.argv[5] if len(sys.argv) > 5 else ''
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
ix = {k: i for i, k in enumerate(h)}
def val(r, k):
    v = float(r[ix[k]].replace(',', '')); u = units[ix[k]]
    return v * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'Tbyte': 1e12}.get(u, 1.0)
sel = [r for r in rows[2:] if kern in r[ix['Kernel Name']]]
assert sel, 'no launch of %s in %s' % (kern, rep)
tot = sum(val(r, 'dram__bytes_read.sum') + val(r, 'dram__bytes_write.sum') for r in sel) / len(sel)
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = os.path.join(root, 'profiles', 'ncu_traffic.json')
d = json.load(open(path)) if os.path.isfile(path) else {}
d.setdefault('config%d' % cfg, {})[kern] = dict(dram_bytes_per_frame=tot / frames, launches=len(sel), frames_per_launch=frames,
                                                source='ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum, %s%s'
                                                       % (os.path.basename(rep), (': ' + note) if note else ''))
json.dump(d, open(path, 'w'), indent=1)
keys = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct']
for r in sel[:3]:
    print(r[ix['Kernel Name']][:60], ', '.join('%s=%s %s' % (k.split('.')[0], r[ix[k]], units[ix[k]]) for k in keys if k in ix))
print('dram bytes per frame: %.0f' % (tot / frames))
