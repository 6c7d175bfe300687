"""Per source line of one launch in an ncu report (--set full --import-source on): share of the stall samples and of the
executed warp instructions, with the two leading stall reasons.  usage: python tools/ncu_lines.py X.ncu-rep <launch index> [min share]"""
import csv, collections, sys, subprocess
rep, skip = sys.argv[1], sys.argv[2]
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.008
out = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--launch-skip',skip,'--launch-count','1','--print-source','sass,cuda'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=None; per=collections.Counter(); ex=collections.Counter(); src={}; st=collections.defaultdict(collections.Counter)
cur=None
for r in rows:
    if r and r[0]=="File Path": cur=r[1].split('/')[-1]; continue
    if r and r[0]=="Function Name": print(r[1][:100]); continue
    if r and r[0]=="Line No": hdr=r; isamp=hdr.index("# Samples"); iex=hdr.index("Instructions Executed"); sti=[(i,h) for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]; continue
    if hdr and len(r)>isamp and r[isamp].isdigit() and r[0].isdigit():
        key=(cur,int(r[0])); per[key]+=int(r[isamp]); ex[key]+=int(r[iex]); src[key]=r[1]
        for i,h in sti: st[key][h[6:]]+=int(r[i])
tot=sum(per.values()); totex=sum(ex.values())
print('samples',tot,'warp instr',totex)
allst=collections.Counter()
for k in st: allst.update(st[k])
n=sum(allst.values()); print({k:round(100*v/n,1) for k,v in allst.most_common(8)})
for k,v in sorted(per.items()):
    if v/tot>thr or ex[k]/totex>thr:
        top=', '.join('%s %d'%(a,100*b/max(1,sum(st[k].values()))) for a,b in st[k].most_common(2))
        print(f"{k[0]:14s}{k[1]:4d} samp {100*v/tot:5.1f}% instr {100*ex[k]/totex:5.1f}% [{top}] {src[k][:95]}")
