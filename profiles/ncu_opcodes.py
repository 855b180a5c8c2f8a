import csv,collections,re,sys,subprocess
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
d={h:(v,u) for h,u,v in zip(rows[0],rows[1],rows[2])}
for k in ['gpu__time_duration.sum','smsp__inst_executed.sum','sm__issue_active.avg.pct_of_peak_sustained_elapsed','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed','sm__cycles_elapsed.max']:
    print(k,d.get(k))
for k in sorted(d):
    if 'issue_stalled' in k and 'per_issue_active' in k and float(d[k][0])>0.15: print(' ',k.split('stalled_')[1].split('_per')[0],d[k][0])
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hdr=rows[1]; ix={h:i for i,h in enumerate(hdr)}
tot=collections.Counter(); inst=collections.Counter(); stall=collections.defaultdict(collections.Counter)
names=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
S=0
for r in rows[2:]:
    if len(r)<len(hdr): continue
    op=[t for t in r[ix['Source']].split() if not t.startswith('@')]
    o=op[0] if op else ''
    pre=o.split('.')[0]
    o=pre+('.'+o.split('.')[1] if pre in('MUFU','LDS','LDG','UTCHMMA','SYNCS','BAR','F2F','STS') and '.' in o else '')
    n=int(r[ix['# Samples']] or 0); e=int(r[ix['Instructions Executed']] or 0)
    tot[o]+=n; inst[o]+=e; S+=n
    for s in names: stall[o][s]+=int(r[ix[s]] or 0)
E=sum(inst.values())
print('samples',S,'inst',E)
for o,n in tot.most_common(int(sys.argv[2]) if len(sys.argv)>2 else 22):
    top=', '.join(f"{k[6:]}:{v}" for k,v in stall[o].most_common(3))
    print(f"{o:16s} samples {100*n/S:5.1f}%  inst {100*inst[o]/E:5.1f}%   {top}")
