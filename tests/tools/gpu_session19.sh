#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/s19_pytest.log 2>&1; tail -3 gpurun_out/s19_pytest.log
tests/tools/gpu_envab.sh "HSGPU_RAYCAST_ALLOC32=1|HSGPU_RAYCAST_ALLOC32=0" 3 200 | tee gpurun_out/s19_ab.txt
