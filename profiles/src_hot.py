"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line.
usage: python profiles/src_hot.py dump.csv [top_n]"""
import csv, sys, collections
def num(x):
    try: return int(float(x))
    except ValueError: return 0
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None
hdr = None
agg = collections.OrderedDict()
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or r[0] == "": continue
    ie = hdr.index("Instructions Executed"); ss = hdr.index("# Samples")
    names = hdr
    d = dict(zip(names[4:], r[4:]))
    key = (cur_file, int(r[0]))
    a = agg.setdefault(key, {"src": r[1].strip(), "inst": 0, "samples": 0, "long_sb": 0, "barrier": 0, "short_sb": 0, "wait": 0})
    a["inst"] += num(r[ie]); a["samples"] += num(r[ss])
    for k, col in (("long_sb", "stall_long_sb"), ("barrier", "stall_barrier"), ("short_sb", "stall_short_sb"), ("wait", "stall_wait")):
        if col in hdr: a[k] += num(r[hdr.index(col)])
ti = sum(a["inst"] for a in agg.values()); ts = sum(a["samples"] for a in agg.values())
print(f"total inst {ti}  total samples {ts}")
print("== by instructions")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1]["inst"])[:top]:
    print(f"{f}:{l:4d} inst {100*a['inst']/ti:5.1f}% samp {100*a['samples']/max(ts,1):5.1f}% | {a['src'][:110]}")
print("== by samples")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    print(f"{f}:{l:4d} samp {100*a['samples']/max(ts,1):5.1f}% (lsb {a['long_sb']} bar {a['barrier']} ssb {a['short_sb']} wait {a['wait']}) inst {100*a['inst']/ti:5.1f}% | {a['src'][:100]}")
