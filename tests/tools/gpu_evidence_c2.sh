#!/bin/bash
# evidence run, headline configuration only: bench lines (both arms), launch list, ncu --set full capture of k_match / k_raycast
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/s9_bench.err; tail -c 300 gpurun_out/s9_bench.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_n1.json 2>> gpurun_out/s9_bench.err
C0="python bench.py --steps 8 --warmup 3 --no-configs --no-cpu-baseline"
$C0 > gpurun_out/s9_plain0.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $C0 > gpurun_out/s9_ncu0.log 2>&1
C1="python bench.py --steps 4 --warmup 3 --no-configs --no-cpu-baseline"
$C1 > gpurun_out/s9_plain1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_match|k_raycast" -s 6 -c 2 -f -o gpurun_out/r02_c2 $C1 > gpurun_out/s9_ncu1.log 2>&1
ls -la gpurun_out/r02_c2.ncu-rep gpurun_out/r02_bench_n1.json
