#!/bin/bash
mkdir -p gpurun_out
tests/tools/gpu_abn.sh "byte occ4 occ5" 3 200 | tee gpurun_out/s17_ab.txt
