"""Top stalled SASS instructions of one kernel from `ncu --page source --csv` output."""
import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
n=int(sys.argv[2]) if len(sys.argv)>2 else 40
hdr=rows[1]
ix={h:i for i,h in enumerate(hdr)}
data=[r for r in rows[2:] if len(r)==len(hdr) and r[ix['# Samples']].isdigit()]
tot=sum(int(r[ix['# Samples']]) for r in data)
print("total samples",tot, "instr", len(data))
keys=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg={k:sum(int(r[ix[k]]) for r in data) for k in keys}
print({k[6:]:v for k,v in sorted(agg.items(), key=lambda kv:-kv[1]) if v>0})
top=sorted(data,key=lambda r:-int(r[ix['# Samples']]))[:n]
for r in top:
    st={k[6:]:int(r[ix[k]]) for k in keys if int(r[ix[k]])>0}
    print(r[ix['Address']][-5:], r[ix['# Samples']], r[ix['Source']][:64], st)
