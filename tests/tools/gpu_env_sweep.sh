#!/bin/bash
# Same-box sweep of one environment knob under bench.py: tests/tools/gpu_env_sweep.sh VAR "v1 v2 ..." [rounds] [steps]
set -u
VAR=$1; VALS=$2; R=${3:-2}; K=${4:-200}
for i in $(seq $R); do for v in $VALS; do
  env $VAR=$v timeout 120 python bench.py --steps $K --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); k=j['roofline']['kernels']
print('$VAR=$v', round(j['value']), round(j['e2e']['value']), round(k['k_match']['avg_launch_ms']*1e3,2), round(k['k_raycast']['avg_launch_ms']*1e3,2), j['e2e']['pose_checksum'])"
done; done
