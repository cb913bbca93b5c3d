#!/bin/bash
# evidence run (round 2): bench lines (both arms), launch list, ncu --set full captures of every hot kernel
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/s9_bench.err; tail -c 300 gpurun_out/s9_bench.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_n1.json 2>> gpurun_out/s9_bench.err
C0="python bench.py --steps 8 --warmup 3 --no-configs --no-cpu-baseline"
$C0 > gpurun_out/s9_plain0.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $C0 > gpurun_out/s9_ncu0.log 2>&1
C1="python bench.py --steps 4 --warmup 3 --no-configs --no-cpu-baseline"
$C1 > gpurun_out/s9_plain1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_match|k_raycast" -s 6 -c 2 -f -o gpurun_out/r02_c2 $C1 > gpurun_out/s9_ncu1.log 2>&1
C3="python bench.py --only-config c3 --steps 2 --warmup 3"
$C3 > gpurun_out/s9_plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_match2|k_raycast" -s 6 -c 2 -f -o gpurun_out/r02_c3 $C3 > gpurun_out/s9_ncu3.log 2>&1
C4="python bench.py --only-config c4 --steps 3"
$C4 > gpurun_out/s9_plain4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_hessian_derivs -s 2 -c 1 -f -o gpurun_out/r02_c4_hd $C4 > gpurun_out/s9_ncu4a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_build_pblk -c 1 -f -o gpurun_out/r02_c4_pb $C4 > gpurun_out/s9_ncu4b.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_match<.*int.2, .int.0>' -s 1 -c 1 -f -o gpurun_out/r02_c4_mh $C4 > gpurun_out/s9_ncu4c.log 2>&1
tail -2 gpurun_out/s9_ncu4c.log
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_*
