#!/bin/bash
for w in 1k room; do for c in 16 8 4 2; do
EKF_CLUSTER=$c timeout 120 python bench.py --workload $w --steps 300 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$w cluster=$c: %.4f ms/step value %.1f e2e %.1f line %s' % (d['ms_per_step'], d['value'], d['e2e']['value'], r.get('line_stream_ms_per_step')))"
done; done
