#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/s4_pytest.log 2>&1; tail -3 gpurun_out/s4_pytest.log
for r in 1 2; do for s in "HSGPU_RAYCAST_BIG=0 HSGPU_MATCH2_THREADS=128" "HSGPU_RAYCAST_BIG=1 HSGPU_MATCH2_THREADS=128" "HSGPU_RAYCAST_BIG=1 HSGPU_MATCH2_THREADS=256" "HSGPU_RAYCAST_BIG=1 HSGPU_MATCH2_THREADS=512" "HSGPU_RAYCAST_BIG=0 HSGPU_MATCH2_THREADS=512"; do env $s python bench.py --only-config c3 --steps 8 --warmup 3 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('c3 [$s]', round(j['value']), round(j['ms_per_step'],4), {n:round(r['avg_launch_ms'],4) for n,r in j['kernels'].items()})"; done; done 2>&1 | tee gpurun_out/s4_c3.txt
tests/tools/gpu_envab.sh "|" 1 200 | tee gpurun_out/s4_c2.txt
