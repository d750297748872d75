#!/bin/bash
# The measurement pass behind profiles/<tag>_*: tools/final_pass.sh <tag> [tests]   (one GPU, ~8 minutes; what it leaves in
# gpurun_out/ stays far below the 64 MiB that travel back: ncu reports are exported to csv on the box and removed)
tag=$1
if [ "$2" = "tests" ]; then
  timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_gpu_tests.log 2>&1; tail -2 gpurun_out/${tag}_gpu_tests.log
  timeout 200 python __graft_entry__.py smoke 2>&1 | grep smoke | tee -a gpurun_out/${tag}_gpu_tests.log
fi
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_bench.json"))
    print("ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["frac"], d["roofline"]["kernel"], d["clocks"], d["cpu_baseline"]["value"], d.get("path_frac_of_peak"))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/${tag}_bench.err").read()[-3000:])
PY
timeout 400 ncu --set full --clock-control none -k "regex:link_heads_f16|select_screen_packed|attend_kernel|select_resolve_packed|gemm_tc_kernel" -s 27 -c 9 -o /tmp/${tag}_kernels python tools/prof_step.py 6 2048 > gpurun_out/${tag}_ncu.log 2>&1; tail -1 gpurun_out/${tag}_ncu.log
ncu -i /tmp/${tag}_kernels.ncu-rep --page raw --csv > gpurun_out/${tag}_kernels_full_raw.csv 2>/dev/null
timeout 200 ncu --set full --clock-control none -k regex:gemm_tc_kernel -c 1 -o /tmp/${tag}_gemm_ddi python tools/gemm_probe.py one > /dev/null 2>&1
ncu -i /tmp/${tag}_gemm_ddi.ncu-rep --page raw --csv > gpurun_out/${tag}_gemm_ddi_full_raw.csv 2>/dev/null
tools/step_launches.sh ${tag} 2048 > gpurun_out/${tag}_launches_summary.txt; tail -3 gpurun_out/${tag}_launches_summary.txt
for w in ddi ppa collab cora; do
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
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_ref.err; tail -c 300 gpurun_out/${tag}_bench_reference_arm.json
python tools/gemm_probe.py > gpurun_out/${tag}_gemm_probe.txt 2>&1
du -sh gpurun_out
