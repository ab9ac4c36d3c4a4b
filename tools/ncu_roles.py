import csv, io, re, subprocess, sys, collections
rep=sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hi]
ia, ie, isrc = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Source")
reasons = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data=[]
for r in rows[hi+1:]:
    if len(r) <= ie or not re.fullmatch(r"(0x)?[0-9a-fA-F]+", r[ia]): continue
    samples = sum(int(float(r[i] or 0)) for i,_ in reasons)
    data.append((int(r[ia],16), r[isrc], int(float(r[ie] or 0)), samples, {h[6:]:int(float(r[i] or 0)) for i,h in reasons if float(r[i] or 0)>0}))
data.sort()
base=data[0][0]
tot=sum(d[3] for d in data); toti=sum(d[2] for d in data)
# segments split at EXIT instructions
seg=[]; cur=[]
for d in data:
    cur.append(d)
    if d[1].strip().startswith("EXIT") or " EXIT" in d[1]:
        seg.append(cur); cur=[]
if cur: seg.append(cur)
for s in seg:
    ss=sum(d[3] for d in s); si=sum(d[2] for d in s)
    if ss < 0.002*tot: continue
    tw=[d for d in s if "TRYWAIT" in d[1]]; bar=[d for d in s if "BAR.SYNC" in d[1]]; ns=[d for d in s if "NANOSLEEP" in d[1]]
    reasons_c=collections.Counter()
    for d in s:
        for k,v in d[4].items(): reasons_c[k]+=v
    print(f"segment 0x{s[0][0]-base:05x}-0x{s[-1][0]-base:05x}: {100*ss/tot:5.1f}% samples {100*si/toti:5.1f}% inst | trywait n={len(tw)} samples {100*sum(d[3] for d in tw)/tot:4.1f}% | nanosleep {100*sum(d[3] for d in ns)/tot:4.1f}% | bar.sync {100*sum(d[3] for d in bar)/tot:4.1f}% | top:", ", ".join(f"{k}={100*v/ss:.0f}%" for k,v in reasons_c.most_common(5)))
