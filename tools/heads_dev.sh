#!/bin/bash
# dev loop for the heads kernel on the GPU box: parity tests, phase clocks, short bench.  usage: tools/heads_dev.sh <tag>
tag=$1
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -k "heads" 2>&1 | tail -8
timeout 200 python tools/heads_clocks.py 2>&1 | tail -26
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_bench.json"))
    print("ms_per_step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "roof", d["roofline"]["frac"], d["roofline"]["kernel"])
    for k in d["kernels"]:
        print("  %-90s %3d %.1f us" % (k["name"], k["calls"], 1e3 * k["total_ms"] / k["calls"]))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/${tag}_bench.err").read()[-3000:])
PY
