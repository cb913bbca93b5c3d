#!/bin/bash
# what the driver does at round end on a fresh box: GPU tests, smoke(), the reference arm, our arm
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/flow_pytest.log 2>&1; grep -E "passed|failed" gpurun_out/flow_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/flow_ref.json 2> gpurun_out/flow_ref.err; tail -3 gpurun_out/flow_ref.err
( time python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/flow_ours.json 2> gpurun_out/flow_ours.err; tail -3 gpurun_out/flow_ours.err
python - <<'PY'
import json
a=json.loads(open('gpurun_out/flow_ours.json').read().strip().splitlines()[0]); r=json.loads(open('gpurun_out/flow_ref.json').read().strip().splitlines()[0])
print('ours value', round(a['value']), 'e2e', round(a['e2e']['value']), 'ref', round(r['value']), 'ratio e2e/ref', round(a['e2e']['value']/r['value'],1), 'frac', round(a['roofline']['frac'],3), 'launches', a['gpu_launches'])
PY
