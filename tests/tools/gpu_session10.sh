#!/bin/bash
mkdir -p gpurun_out
tests/tools/gpu_abn.sh "base atom defer ldsred" 3 200 | tee gpurun_out/s10_ab.txt
