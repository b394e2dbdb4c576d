import csv,collections,sys
rows=list(csv.reader(open(sys.argv[1])))
H=rows[1]
cols=[i for i,h in enumerate(H) if h.startswith('stall_') and 'Not Issued' not in h]
tot=collections.Counter()
for r in rows[2:]:
    for i in cols:
        try: tot[H[i]]+=float(r[i])
        except: pass
s=sum(tot.values())
print("  ".join(f"{k[6:]}={100*v/s:.1f}%" for k,v in tot.most_common(9)))
