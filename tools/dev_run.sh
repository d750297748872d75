#!/bin/bash
# dev loop on the GPU box: selection parity tests, phase clocks, a short bench line.  usage: tools/dev_run.sh <tag> [pytest -k expr]
tag=$1; kexpr=${2:-"select_onepass or packed or plan"}
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "$kexpr" > gpurun_out/${tag}_tests.log 2>&1; tail -5 gpurun_out/${tag}_tests.log
timeout 300 python tools/select_clocks.py > gpurun_out/${tag}_clocks.log 2>&1; cat gpurun_out/${tag}_clocks.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_bench.json"))
    print("ms_per_step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "roof", d["roofline"]["frac"])
    for k in d["kernels"]:
        print("  %-90s %3d %.1f us" % (k["name"], k["calls"], 1e3 * k["total_ms"] / k["calls"]))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/${tag}_bench.err").read()[-3000:])
PY
