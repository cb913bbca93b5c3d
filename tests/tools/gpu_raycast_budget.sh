#!/bin/bash
# k_raycast time budget with the timing switches of the EXPERIMENTS build (results are wrong whenever a switch is set):
#   make -C tfe-slam-ucl_b200/csrc OUT=$PWD/tmp_ab/libhsgpu_exp.so LOG=/tmp/x.log EXPERIMENTS=1    (before calling this)
# HSGPU_DEBUG_SKIP bits: 1 no walk, 2 no sweep, 8 no free-marker stores; HSGPU_DEBUG_ONLY_LEVEL=l: one level's CTAs alone
mkdir -p gpurun_out
LIB=tfe-slam-ucl_b200/libhsgpu.so
cp $LIB tmp_ab/libhsgpu__intree.so
cp tmp_ab/libhsgpu_exp.so $LIB
for s in "HSGPU_DEBUG_SKIP=0" "HSGPU_DEBUG_SKIP=3" "HSGPU_DEBUG_SKIP=3 HSGPU_DEBUG_ONLY_LEVEL=0" "HSGPU_DEBUG_SKIP=3 HSGPU_DEBUG_ONLY_LEVEL=1" "HSGPU_DEBUG_SKIP=3 HSGPU_DEBUG_ONLY_LEVEL=2" "HSGPU_DEBUG_SKIP=2" "HSGPU_DEBUG_SKIP=1" "HSGPU_DEBUG_SKIP=8" "HSGPU_DEBUG_ONLY_LEVEL=0" "HSGPU_DEBUG_ONLY_LEVEL=1" "HSGPU_DEBUG_ONLY_LEVEL=2"; do
  env $s python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-configs 2>/dev/null | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); k=j['roofline']['kernels']
print('[$s]', round(j['value']), round(k['k_match']['avg_launch_ms']*1e3,2), round(k['k_raycast']['avg_launch_ms']*1e3,2))"
done | tee gpurun_out/s16_budget.txt
cp tmp_ab/libhsgpu__intree.so $LIB
