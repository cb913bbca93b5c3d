#!/bin/bash
mkdir -p gpurun_out
tests/tools/gpu_abn.sh "byte t384m32c4 t256m24c5 t256m32c4 t320m32c4 t512m32c3 t512m48c3" 2 200 | tee gpurun_out/s18_ab.txt
for s in "HSGPU_RAYCAST_BIG=1" "HSGPU_RAYCAST_BIG=0"; do env $s python bench.py --only-config c3 --steps 8 --warmup 3 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('c3 [$s]', round(j['value']), round(j['ms_per_step'],4), {n:round(r['avg_launch_ms'],4) for n,r in j['kernels'].items()})"; done | tee gpurun_out/s18_c3.txt
