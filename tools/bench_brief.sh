#!/bin/bash
# short bench lines of the named workloads with the per-entry-point breakdown: tools/bench_brief.sh <tag> <workload>...
tag=$1; shift
for w in "$@"; do
  timeout 200 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_$w.json 2> gpurun_out/${tag}_$w.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_$w.json"))
    print("$w", "ms_per_step", round(d["ms_per_step"], 3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"], 3), "roof", d["roofline"]["kernel"], round(d["roofline"]["frac"], 3))
    for k in d["kernels"][:9]:
        print("   %-60s %4d %9.1f us" % (k["name"][:60], k["calls"], 1e3 * k["total_ms"] / k["calls"]))
except Exception as e:
    print("$w failed", e); print(open("gpurun_out/${tag}_$w.err").read()[-1500:])
PY
done
