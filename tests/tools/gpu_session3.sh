#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/s3_pytest.log 2>&1; tail -3 gpurun_out/s3_pytest.log
( time HSGPU_RAYCAST_SPLIT=1 python -m pytest tests -m gpu -x -q ) > gpurun_out/s3_pytest_split.log 2>&1; tail -3 gpurun_out/s3_pytest_split.log
tests/tools/gpu_envab.sh "|HSGPU_MATCH_CT_PITCH=0|HSGPU_MATCH_SERIAL_SPREAD=1|HSGPU_RAYCAST_SPLIT=1|HSGPU_RAYCAST_SPLIT=1 HSGPU_MATCH_SERIAL_SPREAD=1" 2 200 > gpurun_out/s3_envab.txt 2>&1
cat gpurun_out/s3_envab.txt
for s in "" "HSGPU_RAYCAST_SPLIT=1"; do env $s python bench.py --only-config c3 --steps 8 --warmup 3 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('c3 [$s]', round(j['value']), j['ms_per_step'], {n:round(r['avg_launch_ms'],4) for n,r in j['kernels'].items()})"; done 2>&1 | tee gpurun_out/s3_c3.txt
for s in "" "HSGPU_RAYCAST_SPLIT=1"; do env $s python bench.py --only-config c5 --steps 10 --warmup 5 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('c5 [$s]', round(j['value']), j['ms_per_step'], {n:round(r['avg_launch_ms'],4) for n,r in j['kernels'].items()})"; done 2>&1 | tee gpurun_out/s3_c5.txt
