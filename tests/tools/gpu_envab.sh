#!/bin/bash
# Same-box A/B of environment settings under bench.py: tests/tools/gpu_envab.sh "A=1 B=2|C=3|" [rounds] [steps]
# ("|" separates the settings of one run; an empty entry = defaults). Prints value, e2e, spread, k_match us, k_raycast us, checksum.
set -u
IFS='|' read -ra SETS <<< "$1"; R=${2:-2}; K=${3:-200}
for i in $(seq $R); do for s in "${SETS[@]}"; do
  env $s timeout 300 python bench.py --steps $K --warmup 5 --no-cpu-baseline --no-configs 2>/dev/null | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); k=j['roofline']['kernels']
print('[$s]', round(j['value']), round(j['e2e']['value']), round(j['e2e']['spread'],3), round(k['k_match']['avg_launch_ms']*1e3,2), round(k['k_raycast']['avg_launch_ms']*1e3,2), j['e2e']['pose_checksum'])"
done; done
