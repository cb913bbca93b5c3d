#!/bin/bash
# round-2 GPU session 2: new tests, C4 chunk A/B, ncu captures (headline kernels, C3 kernels, C4 kernels)
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/s2_pytest.log 2>&1
tail -5 gpurun_out/s2_pytest.log
LIB=tfe-slam-ucl_b200/libhsgpu.so
cp $LIB tmp_ab/libhsgpu__intree.so
for r in 1 2; do for v in hd32 hd64 hd128; do cp tmp_ab/libhsgpu_$v.so $LIB; for p in 1 3; do HSGPU_HD_PPW=$p python bench.py --only-config c4 --steps 20 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read()); s=j['single_evaluation']
print('$v ppw $p', {m:(round(s[m]['device_entry']['kernel_ms']*1e3,2), round(s[m]['host_entry']['ms_per_call']*1e3,1)) for m in s}, 'full', round(j['full_match']['ms_per_call']*1e3,1), 'plane', round(j['probability_plane_build_ms']*1e3,1))"; done; done; done > gpurun_out/s2_c4_chunk.txt 2>&1
cp tmp_ab/libhsgpu__intree.so $LIB
cat gpurun_out/s2_c4_chunk.txt
# ncu: headline kernels
C1="python bench.py --steps 4 --warmup 3 --no-configs --no-cpu-baseline"
$C1 > gpurun_out/s2_plain1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_match|k_raycast" -s 6 -c 2 -f -o gpurun_out/r02_c2 $C1 > gpurun_out/s2_ncu1.log 2>&1
tail -3 gpurun_out/s2_ncu1.log
C3="python bench.py --only-config c3 --steps 2 --warmup 3"
$C3 > gpurun_out/s2_plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_match2|k_raycast" -s 6 -c 2 -f -o gpurun_out/r02_c3 $C3 > gpurun_out/s2_ncu3.log 2>&1
tail -3 gpurun_out/s2_ncu3.log
C4="python bench.py --only-config c4 --steps 3"
$C4 > gpurun_out/s2_plain4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_hessian_derivs|k_build_pblk" -s 2 -c 3 -f -o gpurun_out/r02_c4 $C4 > gpurun_out/s2_ncu4.log 2>&1
tail -3 gpurun_out/s2_ncu4.log
ls -la gpurun_out/*.ncu-rep
