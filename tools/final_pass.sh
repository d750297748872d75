#!/bin/bash
# The measurement pass behind profiles/<tag>_*: tools/final_pass.sh <tag>   (one GPU, ~9 minutes)
tag=$1
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_gpu_tests.log 2>&1; tail -2 gpurun_out/${tag}_gpu_tests.log
timeout 200 python __graft_entry__.py smoke 2>&1 | grep smoke
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_bench.json"))
    print("ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"], "roof", d["roofline"]["frac"], d["roofline"]["kernel"], d["clocks"], d["cpu_baseline"]["value"])
    for k in d["kernels"]:
        print("  %-110s %3d %.1f us" % (k["name"], k["calls"], 1e3 * k["total_ms"] / k["calls"]))
    print({k: (round(v["frac"], 3), round(v["avg_launch_ms"], 4)) for k, v in d["roofline"].items() if isinstance(v, dict) and "frac" in v}, d.get("path_frac_of_peak"))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/${tag}_bench.err").read()[-3000:])
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_ref.err; tail -c 600 gpurun_out/${tag}_bench_reference_arm.json
tools/step_launches.sh ${tag} 2048
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:link_heads_f16|select_screen_packed|attend_kernel|select_resolve_packed|gemm_tc_kernel" -s 27 -c 18 -o gpurun_out/${tag}_kernels python tools/prof_step.py 6 2048 > gpurun_out/${tag}_ncu.log 2>&1; tail -2 gpurun_out/${tag}_ncu.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -c 1 -o gpurun_out/${tag}_gemm_ddi python tools/gemm_probe.py one > gpurun_out/${tag}_ncu_gemm.log 2>&1; tail -1 gpurun_out/${tag}_ncu_gemm.log
python tools/gemm_probe.py > gpurun_out/${tag}_gemm_probe.txt 2>&1; cat gpurun_out/${tag}_gemm_probe.txt
for w in collab ddi ppa cora; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_bench_$w.json"))
    print("$w", "ms_per_step", round(d["ms_per_step"], 3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "roof", d["roofline"]["kernel"], round(d["roofline"]["frac"] or 0, 3), "cpu", d["cpu_baseline"]["value"])
except Exception as e:
    print("$w failed", e); print(open("gpurun_out/${tag}_bench_$w.err").read()[-1500:])
PY
done
