#!/bin/bash
mkdir -p gpurun_out
for cfg in "10k 8 100" "10k 16 100" "10k 4 100" "40k 8 20" "40k 16 20" "40k 32 20"; do
  set -- $cfg
  EKF_LINE_SMS=$2 timeout 600 python bench.py --workload $1 --steps $3 --warmup 3 --no-cpu-baseline > gpurun_out/linesms_$1_$2.json 2> gpurun_out/linesms_$1_$2.err
  python - <<PY
import json
d=json.load(open("gpurun_out/linesms_$1_$2.json"))
r=d["roofline"]
print("$1 line_sms=$2: %.3f ms/step value %.1f  sweep %.3f ms (%.0f GB/s, share %.3f)  line stream %.3f ms  e2e %.1f" % (d["ms_per_step"], d["value"], r["launch_ms"], r["achieved"], r["sweep_share_of_step"], r["line_stream_ms_per_step"], d["e2e"]["value"]))
PY
done
