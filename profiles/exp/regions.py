import csv, sys, collections
def num(x):
    try: return int(float(x))
    except ValueError: return 0
rows = list(csv.reader(open(sys.argv[1])))
cur=None; hdr=None
agg=collections.Counter(); samp=collections.Counter()
per_line={}
for r in rows:
    if not r: continue
    if r[0]=="File Path": cur=r[1].split("/")[-1]; continue
    if r[0]=="Function Name": continue
    if r[0]=="Line No": hdr=r; continue
    if hdr is None or r[0]=="": continue
    ie=hdr.index("Instructions Executed"); ss=hdr.index("# Samples")
    key=(cur,int(r[0]))
    a=per_line.setdefault(key,[0,0]); a[0]+=num(r[ie]); a[1]+=num(r[ss])
import re
src={}
def region(f,l):
    if f.startswith('kernels_probe2'):
        for name,lo,hi in REG2:
            if lo<=l<=hi: return name
        return 'p2:other'
    if f=='probe_common.cuh':
        for name,lo,hi in REGC:
            if lo<=l<=hi: return name
        return 'pc:other'
    if f=='common.cuh':
        if 90<=l<=112: return 'q15_mul/unpack (rerank compute)'
        if 113<=l<=140: return 'ndarray_dot (fp32 final)'
        return 'common:other'
    return f
REG2=[('anchors',120,205),('ranges_lockstep',206,242),('rerank_lookup',243,259),('rows_request',260,268),('rows_compute',269,320),('memo_zero',330,345),('locate',350,362),('load_idx/sketch/ham',363,398),('sweep consume',399,428),('topup+tail',429,460),('rerank drive',461,476),('insert/msd/stop',477,500),('cluster walk',505,640),('helpers(asm)',25,70)]
REGC=[('scan',44,55),('warp_max',56,64),('bitonic sort',65,108),('topk',109,147),('maxbuffer_filter',148,182),('maxbuffer_insert',183,219),('lcp/lead',224,236),('table_anchor',240,271),('table_range',272,330)]
ti=sum(v[0] for v in per_line.values()); ts=sum(v[1] for v in per_line.values())
for (f,l),(i,s) in per_line.items():
    rg=region(f,l); agg[rg]+=i; samp[rg]+=s
for k,v in agg.most_common(): print(f"{k:40s} inst {100*v/ti:5.1f}%  samples {100*samp[k]/ts:5.1f}%")
print("---- static SASS instructions per region (x16 B)")
static=collections.Counter()
cur=None; hdr=None
for r in rows:
    if not r: continue
    if r[0]=="File Path": cur=r[1].split("/")[-1]; continue
    if r[0]=="Function Name": continue
    if r[0]=="Line No": hdr=r; continue
    if hdr is None or r[0]=="": continue
    # count SASS lines: column 'Source' holds either cuda or sass; try to detect count column
    static[region(cur,int(r[0]))]+=1
print(hdr[:8])
for k,v in static.most_common(): print(f"{k:40s} rows {v}")
