for cfg in "--depth 4" "--depth 6" "--depth 8" "--queries 2048 --depth 4" "--queries 2048 --depth 6"; do
  timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu-baseline $cfg > gpurun_out/g7.json 2> gpurun_out/g7.err
  python - "$cfg" <<PY
import json, sys
try:
    d = json.load(open("gpurun_out/g7.json"))
    print(sys.argv[1], "| ms_per_step", round(d["ms_per_step"], 4), "value %.3f G" % (d["value"] / 1e9), "e2e %.3f G" % (d["e2e"]["value"] / 1e9))
except Exception as e:
    print(sys.argv[1], "failed", e); print(open("gpurun_out/g7.err").read()[-800:])
PY
done
tools/bench_brief.sh g7 ppa
