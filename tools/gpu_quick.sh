#!/bin/bash
# usage (on the GPU box): tools/gpu_quick.sh <tag>  -- GPU parity tests + opbench + short bench
tag=${1:-q}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/${tag}_pytest.log 2>&1
tail -5 gpurun_out/${tag}_pytest.log
timeout 300 python tools/opbench.py --ops sign,scalarmul,comb,x448 > gpurun_out/${tag}_opbench.txt 2>&1
cat gpurun_out/${tag}_opbench.txt
timeout 900 python bench.py --no-cpu --no-extra --steps 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -3 gpurun_out/${tag}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench.json'))
print('verify %.3f M/s e2e %.3f  kernels %s frac %.3f' % (d['value']/1e6, d['e2e']['value']/1e6, {k: round(v,2) for k,v in d['roofline']['kernel_ms'].items()}, d['roofline']['frac']))
PY
