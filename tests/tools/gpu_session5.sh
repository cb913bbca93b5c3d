#!/bin/bash
# per-level CTA slot time of k_raycast: one level's CTAs alone (experiments build, wrong maps)
mkdir -p gpurun_out
LIB=tfe-slam-ucl_b200/libhsgpu.so
cp $LIB tmp_ab/libhsgpu__intree.so
cp tmp_ab/libhsgpu_exp.so $LIB
for l in "" 0 1 2; do
  if [ -z "$l" ]; then S="X=1"; else S="HSGPU_DEBUG_ONLY_LEVEL=$l"; fi
  env $S python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-configs 2>/dev/null | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); k=j['roofline']['kernels']
print('[$S]', round(j['value']), round(k['k_match']['avg_launch_ms']*1e3,2), round(k['k_raycast']['avg_launch_ms']*1e3,2))"
done | tee gpurun_out/s5_levels.txt
cp tmp_ab/libhsgpu__intree.so $LIB
