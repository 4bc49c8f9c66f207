#!/bin/bash
# developer tool: A/B several environment variants of the bench on one box
#   tools/ab.sh TAG "VAR=1 VAR2=x" "VAR=0" ...   -> gpurun_out/bench_TAG_<variant>.json + a one-line summary each
tag=$1; shift
for v in "$@"; do
  name=$(echo "$v" | tr " =" "__")
  env $v python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-ref-gpu --no-loss-leg ${BENCH_ARGS} > gpurun_out/bench_${tag}_${name}.json 2> gpurun_out/bench_${tag}_${name}.err
  python - "$v" gpurun_out/bench_${tag}_${name}.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    st = {k.replace("pxb_", ""): v["ms_avg"] for k, v in d.get("roofline_stages", {}).items()}
    print(sys.argv[1], "| value", d["value"], "ms", d["ms_per_step"], "median", d.get("step_ms_distribution", {}).get("median"),
          "e2e", d.get("e2e", {}).get("value"), "mpix", d.get("render_mpix_s"), st, d.get("errors"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
