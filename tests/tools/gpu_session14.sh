#!/bin/bash
# k_raycast time budget with the timing switches of the experiments build (results are wrong when a switch is set)
mkdir -p gpurun_out
LIB=tfe-slam-ucl_b200/libhsgpu.so
cp $LIB tmp_ab/libhsgpu__intree.so
cp tmp_ab/libhsgpu_exp.so $LIB
for s in 0 8 64 2 1 4 3 256; do
  HSGPU_DEBUG_SKIP=$s python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-configs 2>/dev/null | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); k=j['roofline']['kernels']
print('skip=$s', round(j['value']), round(k['k_match']['avg_launch_ms']*1e3,2), round(k['k_raycast']['avg_launch_ms']*1e3,2))"
done | tee gpurun_out/s14_skip.txt
cp tmp_ab/libhsgpu__intree.so $LIB
