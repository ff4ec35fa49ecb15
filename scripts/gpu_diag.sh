#!/bin/bash
mkdir -p gpurun_out
echo "== MAXC=16 / 8 batched tests"
EKF_SWEEP_MAXC=16 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "batched_multi or large_line_groups" 2>&1 | tail -3
EKF_SWEEP_MAXC=8 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "batched_multi or large_line_groups" 2>&1 | tail -3
echo "== MAXC=16 m=64 at 10k under compute-sanitizer (memcheck), 3 steps"
EKF_SWEEP_MAXC=16 timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python scripts/probe_m64.py 2>&1 | tail -25
echo "== line timing 10k"
EKF_LIB=$PWD/scripts/libekfcuda_timing.so timeout 300 python scripts/line_timing.py 10000 2>&1 | tail -12
echo "== LINE_SMS sweep"
for sms in 8 10 12 16; do
  EKF_LINE_SMS=$sms timeout 600 python bench.py --workload 10k --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/q_linesms_$sms.json 2> gpurun_out/q_linesms_$sms.err
  python - <<PY
import json
d=json.load(open("gpurun_out/q_linesms_$sms.json"))
r=d["roofline"]
print("10k line_sms=$sms: %.3f ms/step value %.1f  sweep %.3f ms (%.0f GB/s)  line stream %.3f ms  e2e %.1f" % (d["ms_per_step"], d["value"], r["launch_ms"], r["achieved"], r["line_stream_ms_per_step"], d["e2e"]["value"]))
PY
done
echo "== ncu sweep quad"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep_quad -c 6 -o gpurun_out/prof_sweep_quad_r1 -f python scripts/ncu_sweep.py 10000 1,8,16,32 > gpurun_out/ncu_quad.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_quad.log
