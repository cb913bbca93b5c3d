#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/s11_pytest.log 2>&1; tail -3 gpurun_out/s11_pytest.log
cp tfe-slam-ucl_b200/libhsgpu.so tmp_ab/libhsgpu_new.so
tests/tools/gpu_abn.sh "atom new" 3 200 | tee gpurun_out/s11_ab.txt
for v in atom new; do cp tmp_ab/libhsgpu_$v.so tfe-slam-ucl_b200/libhsgpu.so; python bench.py --only-config c3 --steps 8 --warmup 3 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('c3 [$v]', round(j['value']), round(j['ms_per_step'],4), {n:round(r['avg_launch_ms'],4) for n,r in j['kernels'].items()})"; done | tee gpurun_out/s11_c3.txt
cp tmp_ab/libhsgpu_new.so tfe-slam-ucl_b200/libhsgpu.so
