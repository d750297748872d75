#!/bin/bash
# ncu launch list (device time per kernel, serialised) of a short bench run of one of the other BASELINE shapes:
# tools/workload_launches.sh <tag> <workload> [steps]
tag=$1; wl=$2; steps=${3:-2}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_${wl}_launches.csv \
    python bench.py --workload $wl --steps $steps --warmup 1 --no-cpu-baseline --no-sampler > gpurun_out/${tag}_${wl}.log 2>&1
python - <<PY
import csv, re, collections
rows = [r for r in csv.reader(open('gpurun_out/${tag}_${wl}_launches.csv')) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit'); gi = hdr.index('Grid Size')
agg = collections.OrderedDict()
for r in rows[1:]:
    name = re.sub(r'\(.*', '', r[ki]); name = re.sub(r'^void ', '', name).split('::')[-1]
    v = float(r[vi].replace(',', '')); u = r[ui]
    v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
    k = (name[:60], r[gi])
    a = agg.setdefault(k, [0, 0.0, 0.0]); a[0] += 1; a[1] += v; a[2] = max(a[2], v)
tot = sum(a[1] for a in agg.values())
print("== ${wl}: %d launches, %.1f us in kernels" % (sum(a[0] for a in agg.values()), tot))
for (n, g), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    print("%-60s grid %-18s x%-4d total %9.1f us  avg %8.1f  max %8.1f" % (n, g, a[0], a[1], a[1] / a[0], a[2]))
PY
