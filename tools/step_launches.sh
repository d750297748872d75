#!/bin/bash
# ncu launch list (device time of every kernel, serialised) of the last eager citation2-shaped step: tools/step_launches.sh <tag>
tag=$1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv python tools/prof_step.py 6 ${2:-1024} 2>&1 | tail -1
python - <<PY
import csv,re
rows=[r for r in csv.reader(open('gpurun_out/${tag}_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
seq=[]
for r in rows[1:]:
    name=re.sub(r'\(.*','',r[ki]); name=re.sub(r'^void ','',name).split('::')[-1]
    v=float(r[vi].replace(',','')); u=r[ui]
    v = v/1e3 if u=='ns' else (v*1e3 if u=='ms' else v)
    seq.append((name,v))
idx=[i for i,(n,_) in enumerate(seq) if 'link_heads' in n and ', 0>' in n]      # the all-links heads launch opens a step
tot=0
for n,v in seq[idx[-1]:]:
    if 'elementwise' in n or 'at::' in n: continue
    print("%-50s %.1f us"%(n[:50],v)); tot+=v
print("sum %.1f us"%tot)
PY
