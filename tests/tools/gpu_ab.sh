#!/bin/bash
# Same-box A/B of two builds of libhsgpu.so: tmp_ab/libhsgpu_old.so (A) against the in-tree build (B), alternating.
# usage: tests/tools/gpu_ab.sh [rounds] [steps]      (prints value, e2e, k_match ms, k_raycast ms, pose checksum per run)
set -u
R=${1:-2}; K=${2:-200}
LIB=tfe-slam-ucl_b200/libhsgpu.so
mkdir -p tmp_ab   # tmp_ab/libhsgpu_old.so: build A with `make -C tfe-slam-ucl_b200/csrc OUT=$PWD/tmp_ab/libhsgpu_old.so` before editing
cp $LIB tmp_ab/libhsgpu_new.so
run() {
  cp tmp_ab/libhsgpu_$1.so $LIB
  timeout 120 python bench.py --steps $K --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); k=j['roofline']['kernels']
print('$1', round(j['value']), round(j['e2e']['value']), round(k['k_match']['avg_launch_ms']*1e3,2), round(k['k_raycast']['avg_launch_ms']*1e3,2), j['e2e']['pose_checksum'])"
}
for i in $(seq $R); do run old; run new; done
cp tmp_ab/libhsgpu_new.so $LIB
