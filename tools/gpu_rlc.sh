#!/bin/bash
# usage (on the GPU box, via gpurun): tools/gpu_rlc.sh <tag>  -- GPU parity tests (incl. the RLC path) + the full default bench line
tag=${1:-rlc}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${tag}_pytest.log 2>&1
tail -15 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -5 gpurun_out/${tag}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench.json'))
print('verify %.3f M/s e2e %.3f' % (d['value']/1e6, d['e2e']['value']/1e6))
for k,v in d['extra'].items(): print(k, json.dumps(v)[:900])
PY
