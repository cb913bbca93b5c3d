#!/bin/bash
mkdir -p gpurun_out
tests/tools/gpu_abn.sh "new occ4 occ5" 3 200 | tee gpurun_out/s12_ab.txt
