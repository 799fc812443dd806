#!/bin/bash
# round-2 late run: GPU tier, then the bench line (gathered single calls, half-size stand-alone verification)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02u_pytest.log
python bench.py --steps 5 --warmup 3 --no-strong > gpurun_out/r02u_bench.json 2> gpurun_out/r02u_bench.err
tail -c 1500 gpurun_out/r02u_pytest.log
python - <<'P'
import json
d = json.load(open('gpurun_out/r02u_bench.json'))
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline'])
x = d['extra']
for k in ('verify_distinct_keys', 'verify_64_per_key', 'verify_one_signer', 'verify_keyset_e2e'):
    print(k, json.dumps(x.get(k)))
print(json.dumps(x.get('rlc_sweep_distinct_keys')))
P
tail -5 gpurun_out/r02u_bench.err
