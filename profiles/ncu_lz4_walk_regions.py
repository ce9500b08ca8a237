"""profiles/ncu_lz4_walk_regions.py — instruction / sample share of the parts of the LZ4 chain search (probe, extension, the three
tiers, index build ...) from an ncu --set full --import-source on capture.  usage: python profiles/ncu_lz4_walk_regions.py rep.ncu-rep"""
import subprocess,csv,sys
rep=sys.argv[1]
txt = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass,cuda"],capture_output=True,text=True).stdout
cur=None;hdr=None;by={}
for r in csv.reader(txt.splitlines()):
    if not r: continue
    if r[0]=="File Path": cur=r[1].split("/")[-1]; continue
    if r[0]=="Line No": hdr=r; continue
    if hdr is None or not r[0].isdigit(): continue
    extra=len(r)-len(hdr)
    if extra>0: r=[r[0],",".join(r[1:2+extra])]+r[2+extra:]
    # duplicate "Source" header: build by index
    idx={h:i for i,h in enumerate(hdr)}
    try:
        n=float(r[idx["Instructions Executed"]] or 0); smp=float(r[idx["Warp Stall Sampling (All Samples)"]] or 0)
    except Exception: continue
    k=(cur,int(r[0])); a=by.get(k,(0,0)); by[k]=(a[0]+n,a[1]+smp)
tot=sum(v[0] for v in by.values()); ts=sum(v[1] for v in by.values())
def rng(f,a,b):
    n=sum(v[0] for k,v in by.items() if k[0]==f and a<=k[1]<=b); s=sum(v[1] for k,v in by.items() if k[0]==f and a<=k[1]<=b)
    return f"{100*n/tot:5.1f}% inst {100*s/ts:5.1f}% smp"
src=open('lz4-jpeg_b200/csrc/lz4_lazy.cuh').read().splitlines()
def find(s):
    for i,l in enumerate(src,1):
        if s in l: return i
    raise KeyError(s)
F='lz4_lazy.cuh'
print("probe  ", rng(F, find("auto probe ="), find("auto extend =")-1))
print("extend ", rng(F, find("auto extend ="), find("auto eval =")-1))
print("longcmp", rng(F, find("auto long_compare ="), find("for (uint32_t round = 0;;")-1))
print("index  ", rng(F, find("index: counting sort"), find("phase(3); // index")))
print("setup  ", rng(F, find("while (__any_sync(FULL, act))"), find("// (i) short lists: every lane its own")-1))
print("tier i ", rng(F, find("// (i) short lists: every lane its own"), find("// (ii) long lists: the whole warp")-1))
print("tier ii", rng(F, find("// (ii) long lists: the whole warp"), find("// (iii) the lists in between: teams")-1))
print("tieriii", rng(F, find("// (iii) the lists in between: teams"), find("// ---- record: the step")-1))
print("record+", rng(F, find("// ---- record: the step"), find("phase(1); // later rounds")))
for f in sorted(set(k[0] for k in by)):
    print(f, f"{100*sum(v[0] for k,v in by.items() if k[0]==f)/tot:.1f}% inst {100*sum(v[1] for k,v in by.items() if k[0]==f)/ts:.1f}% smp")
