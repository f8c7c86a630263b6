"""mbarrier wait sites of a kernel in an ncu report (samples per SYNCS.TRYWAIT + branch) and stall samples per setmaxnreg role region: python tools/ncu_wait_sites.py rep.ncu-rep"""
import csv, sys, subprocess, bisect
rep=sys.argv[1]
out=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass,cuda"],capture_output=True,text=True).stdout
hdr=None; rows=[]
for row in csv.reader(out.splitlines()):
    if not row: continue
    if row[0]=="Line No": hdr=row; continue
    if hdr and len(row)==len(hdr): rows.append(row)
ia=hdr.index("Address"); isamp=hdr.index("# Samples")
isrc=[i for i,h in enumerate(hdr) if h=="Source"]
stall_cols=[(i,h) for i,h in enumerate(hdr) if h.startswith("stall_") and "(Not" not in h]
seen=set(); u=[]
for r in rows:
    if r[ia].startswith("0x") and r[ia] not in seen:
        seen.add(r[ia])
        st={h[6:]:int(r[i]) for i,h in stall_cols if r[i] not in ("0","","-")}
        try: s=int(r[isamp])
        except: s=0
        u.append((int(r[ia],16), r[isrc[1]], s, st))
u.sort(); base=u[0][0]
tot=sum(x[2] for x in u)
print("total samples",tot)
names={0x00:"a_full",0x20:"a_empty",0x40:"b_full",0x60:"b_empty",0x80:"acc_full",0x90:"acc_empty"}
for i,(a,t,s,st) in enumerate(u):
    if "TRYWAIT" in t:
        print(hex(a-base), s+u[i+1][2], t.strip()[:70])
marks=[a-base for a,t,s,st in u if "USETMAXREG" in t]
reg=[0]*(len(marks)+1); regst=[{} for _ in reg]
for a,t,s,st in u:
    k=bisect.bisect_right(marks,a-base); reg[k]+=s
    for kk,v in st.items(): regst[k][kk]=regst[k].get(kk,0)+v
for k in range(len(reg)):
    print("region",k,reg[k],sorted(regst[k].items(),key=lambda kv:-kv[1])[:5])
