#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/s6_pytest.log 2>&1; tail -3 gpurun_out/s6_pytest.log
cp tfe-slam-ucl_b200/libhsgpu.so tmp_ab/libhsgpu_new.so
tests/tools/gpu_abn.sh "base new" 3 200 | tee gpurun_out/s6_ab.txt
