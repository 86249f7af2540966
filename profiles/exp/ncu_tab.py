import csv,sys
from collections import OrderedDict
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
hdr=rows[0]
iid=hdr.index('ID'); mn=hdr.index('Metric Name'); mv=hdr.index('Metric Value'); kn=hdr.index('Kernel Name'); gs=hdr.index('Grid Size'); bs=hdr.index('Block Size')
d=OrderedDict()
for r in rows[1:]:
    d.setdefault(r[iid],{'k':r[kn][:28],'grid':r[gs],'block':r[bs]})[r[mn].replace('.sum','').replace('smsp__','').replace('.avg.pct_of_peak_sustained_active','')]=r[mv]
for i,v in d.items(): print(i,v)
