#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/s8_pytest.log 2>&1; tail -3 gpurun_out/s8_pytest.log
for r in 1 2; do for s in "HSGPU_HYP_PBLK=1" "HSGPU_HYP_PBLK=0"; do env $s python bench.py --only-config c4 --steps 20 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read()); s=j['single_evaluation']
print('[$s]', {m:(round(s[m]['device_entry']['kernel_ms']*1e3,2), round(s[m]['host_entry']['ms_per_call']*1e3,1)) for m in s}, 'full', round(j['full_match']['ms_per_call']*1e3,1), round(j['full_match']['kernel_ms']*1e3,1), 'plane', round(j['probability_plane_build_ms']*1e3,1), 'l2peak', round(j['l2_peak_gbs']))"; done; done 2>&1 | tee gpurun_out/s8_c4.txt
python tests/tools/gpu_overlap_experiment.py 1024 100 2>&1 | tee gpurun_out/s8_overlap.txt
python tests/tools/gpu_overlap_experiment.py 4096 30 2>&1 | tee -a gpurun_out/s8_overlap.txt
