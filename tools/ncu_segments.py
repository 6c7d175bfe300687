"""Split the ncu source page (CSV, --print-source sass) of one kernel into barrier-delimited regions and print, per
region, its share of the stall samples, executed instructions, FP64 share and top stall reasons.
usage: ncu -i X.ncu-rep --page source --csv --print-source sass [--kernel-name K] > src.csv; python tools/ncu_segments.py src.csv [kernel-substring]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2] if len(sys.argv) > 2 else ''
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name' and want in r[1]]
k0 = starts[0]
end = next((i for i in range(k0 + 1, len(rows)) if rows[i] and rows[i][0] == 'Kernel Name'), len(rows))
hdr = rows[k0 + 1]; data = [r for r in rows[k0 + 2:end] if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
S = [int(r[ix['# Samples']]) for r in data]; ex = [int(r[ix['Instructions Executed']]) for r in data]
tot = sum(S); print('samples', tot, 'warp instructions %.3e' % sum(ex), 'sass lines', len(data))
st = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
def op(r):
    t = r[ix['Source']].split()
    return t[1] if t and t[0].startswith('@') else (t[0] if t else '')
seg0 = 0; segs = []
for i, r in enumerate(data):
    if 'BAR.SYNC' in r[ix['Source']]:
        segs.append((seg0, i)); seg0 = i + 1
segs.append((seg0, len(data) - 1))
for a, b in segs:
    s = sum(S[a:b + 1]); e = sum(ex[a:b + 1])
    if s > tot * 0.004:
        t = {h: sum(int(data[k][ix[h]]) for k in range(a, b + 1)) for h in st}
        n = sum(t.values()) or 1
        fp = sum(ex[k] for k in range(a, b + 1) if op(data[k]).startswith(('DFMA', 'DMUL', 'DADD')))
        print('sass %5d-%5d  samples %5.1f%%  instr %.2e  fp64 %3.0f%%  ' % (a, b, 100 * s / tot, e, 100 * fp / max(e, 1)),
              {h[6:]: round(100 * v / n) for h, v in sorted(t.items(), key=lambda x: -x[1])[:5]})
