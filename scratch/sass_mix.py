import csv,collections,re,sys
rows=list(csv.reader(open(sys.argv[1])))
nrows=float(sys.argv[2])
H=rows[1]
si=H.index('Source'); ei=H.index('Instructions Executed'); st=H.index('# Samples')
agg=collections.Counter(); samp=collections.Counter(); tot=0; stot=0
for r in rows[2:]:
    try: n=float(r[ei]); s=float(r[st])
    except: continue
    m=re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)', r[si])
    if not m: continue
    op=m.group(2)
    agg[op]+=n; tot+=n; samp[op]+=s; stot+=s
print("total warp-instr",tot, "per row", tot/nrows)
for op,n in agg.most_common(int(sys.argv[3]) if len(sys.argv)>3 else 45): print(f"{op:10s} {n/tot*100:6.2f}%  {n/nrows:8.1f}/row   stall-samples {samp[op]/stot*100:5.1f}%")
