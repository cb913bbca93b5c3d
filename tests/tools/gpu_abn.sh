#!/bin/bash
# Same-box A/B/... of several builds of libhsgpu.so: tmp_ab/libhsgpu_<name>.so for every <name> given, alternating.
# Build a variant with `make -C tfe-slam-ucl_b200/csrc OUT=$PWD/tmp_ab/libhsgpu_<name>.so LOG=/tmp/x.log EXTRA=-D...`.
# usage: tests/tools/gpu_abn.sh "name1 name2 ..." [rounds] [steps]   (prints value, e2e, k_match us, k_raycast us, pose checksum per run)
set -u
NAMES=$1; R=${2:-2}; K=${3:-200}
LIB=tfe-slam-ucl_b200/libhsgpu.so
cp $LIB tmp_ab/libhsgpu__intree.so
run() {
  cp tmp_ab/libhsgpu_$1.so $LIB
  timeout 300 python bench.py --steps $K --warmup 5 --no-cpu-baseline --no-configs 2>/dev/null | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); k=j['roofline']['kernels']
print('$1', round(j['value']), round(j['e2e']['value']), round(j['e2e']['spread'],3), round(k['k_match']['avg_launch_ms']*1e3,2), round(k['k_raycast']['avg_launch_ms']*1e3,2), j['e2e']['pose_checksum'])"
}
for i in $(seq $R); do for n in $NAMES; do run $n; done; done
cp tmp_ab/libhsgpu__intree.so $LIB
