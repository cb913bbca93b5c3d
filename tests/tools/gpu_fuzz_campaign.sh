#!/bin/bash
# Longer randomised parity campaign on a GPU box: the fuzz test of tests/test_gpu_edge_cases.py with other seeds and more trials.
# usage: tests/tools/gpu_fuzz_campaign.sh "seed1 seed2 ..." [trials per seed]
set -u
for s in $1; do
  HSGPU_FUZZ_SEED=$s HSGPU_FUZZ_TRIALS=${2:-48} python -m pytest tests/test_gpu_edge_cases.py -m gpu -q -x -k fuzz_random_scans 2>&1 | tail -2 | sed "s/^/[seed $s] /"
done
