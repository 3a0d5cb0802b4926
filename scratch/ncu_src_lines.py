"""Per-source-line stall summary from `ncu --page source --csv --print-source cuda,sass`."""
import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
n=int(sys.argv[2]) if len(sys.argv)>2 else 40
# find header rows ("Line No", ...) ; multiple files
out=[]
hdr=None; fpath=None
for r in rows:
    if r and r[0]=="File Path": fpath=r[1]; continue
    if r and r[0]=="Line No": hdr=r; continue
    if hdr is None or len(r)!=len(hdr): continue
    if r[0]=="" : continue   # sass rows
    ix={h:i for i,h in enumerate(hdr)}
    try: s=int(r[hdr.index('# Samples')])
    except: continue
    keys=[i for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    st={hdr[i][6:]:int(r[i]) for i in keys if r[i].isdigit() and int(r[i])>0}
    inst=r[hdr.index('Instructions Executed')]
    out.append((s, fpath.split('/')[-1], r[0], r[1].strip()[:70], inst, st))
tot=sum(o[0] for o in out)
print("total", tot)
for o in sorted(out,key=lambda o:-o[0])[:n]:
    top=sorted(o[5].items(), key=lambda kv:-kv[1])[:4]
    print("%6d %5.1f%% %s:%s inst=%s | %s | %s"%(o[0],100*o[0]/tot,o[1],o[2],o[4],o[3],top))
