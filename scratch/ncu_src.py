import csv, sys, subprocess
rep=sys.argv[1]; top=int(sys.argv[2]) if len(sys.argv)>2 else 30
out=subprocess.run(["ncu","-i",rep,"--page","source","--csv"]+sys.argv[3:],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
blocks=[]; cur=None
for r in rows:
    if r and r[0]=="Kernel Name": cur={"name":r[1],"rows":[]}; blocks.append(cur); continue
    if r and r[0]=="Address": cur["hdr"]=r; continue
    if cur is not None and r: cur["rows"].append(r)
for b in blocks:
    h={k:i for i,k in enumerate(b["hdr"])}
    c=h['Warp Stall Sampling (All Samples)']
    tot=sum(int(r[c] or 0) for r in b["rows"])
    print("==",b["name"],"samples",tot)
    stalls=[k for k in b["hdr"] if k.startswith("stall_")]
    agg={k:sum(int(r[h[k]] or 0) for r in b["rows"]) for k in stalls}
    print("  stall totals:",{k:v for k,v in sorted(agg.items(),key=lambda t:-t[1])[:8]})
    for r in sorted(b["rows"],key=lambda r:-int(r[c] or 0))[:top]:
        n=int(r[c] or 0)
        st=sorted(((k,int(r[h[k]] or 0)) for k in stalls),key=lambda t:-t[1])[:2]
        print("  %5.1f%% %8s  %-70s %s"%(100*n/max(tot,1), r[h['Instructions Executed']], r[h['Source']].strip()[:70], st))
