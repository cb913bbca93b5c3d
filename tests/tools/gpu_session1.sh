#!/bin/bash
# round-2 GPU session 1: parity tests, k_match A/B, full bench with the config sub-records, C4 poses-per-warp sweep
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/s1_gpu.txt
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/s1_pytest.log 2>&1
tail -5 gpurun_out/s1_pytest.log
tests/tools/gpu_abn.sh "m0 m1 m2" 2 200 > gpurun_out/s1_ab_match.txt 2>&1
cat gpurun_out/s1_ab_match.txt
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/s1_bench.json 2> gpurun_out/s1_bench.err
tail -c 600 gpurun_out/s1_bench.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/s1_bench.json').read().strip().splitlines()[0])
print('value',j['value'],'e2e',j['e2e']['value'],j['e2e']['repeat_values'])
c=j['configs']
for k,v in c.items():
    if isinstance(v,dict): print(k, v.get('value'), v.get('unit'), v.get('ms_per_step'), {n:(round(r['avg_launch_ms'],4), round(r['frac'] or 0,3)) for n,r in v.get('kernels',{}).items()})
print(json.dumps(c.get('c4_hypotheses'),indent=1)[:3000])
PY
for p in 1 2 3; do HSGPU_HD_PPW=$p python bench.py --only-config c4 --steps 20 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read()); s=j['single_evaluation']
print('ppw $p', {m:(round(s[m]['device_entry']['kernel_ms']*1e3,2), round(s[m]['host_entry']['ms_per_call']*1e3,1)) for m in s}, 'full', round(j['full_match']['ms_per_call']*1e3,1))"; done > gpurun_out/s1_c4_ppw.txt 2>&1
cat gpurun_out/s1_c4_ppw.txt
