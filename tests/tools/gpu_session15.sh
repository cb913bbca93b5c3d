#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/s15_pytest.log 2>&1; tail -5 gpurun_out/s15_pytest.log
cp tfe-slam-ucl_b200/libhsgpu.so tmp_ab/libhsgpu_byte.so
tests/tools/gpu_abn.sh "i32 byte" 3 200 | tee gpurun_out/s15_ab.txt
for v in i32 byte; do cp tmp_ab/libhsgpu_$v.so tfe-slam-ucl_b200/libhsgpu.so; python bench.py --only-config c3 --steps 8 --warmup 3 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('c3 [$v]', round(j['value']), round(j['ms_per_step'],4), {n:round(r['avg_launch_ms'],4) for n,r in j['kernels'].items()})"; python bench.py --only-config c5 --steps 10 --warmup 5 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('c5 [$v]', round(j['value']), round(j['ms_per_step'],4), {n:round(r['avg_launch_ms'],4) for n,r in j['kernels'].items()})"; done | tee gpurun_out/s15_c3c5.txt
cp tmp_ab/libhsgpu_byte.so tfe-slam-ucl_b200/libhsgpu.so
