#!/bin/bash
# usage (on the GPU box): tools/gpu_tier_bench.sh <tag>  -- the GPU test tier, then a 5-step bench line without the strong-scaling leg;
# leaves gpurun_out/<tag>_pytest.log, <tag>_bench.json, <tag>_bench.err and prints the headline and the verification extras.
# (The builds r02k ... r02u of round 2 were measured with copies of this script, one per tag.)
tag=${1:-run}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${tag}_pytest.log
python bench.py --steps 5 --warmup 3 --no-strong > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -c 1500 gpurun_out/${tag}_pytest.log
python - "$tag" <<'P'
import json, sys
d = json.load(open('gpurun_out/%s_bench.json' % sys.argv[1]))
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline'])
x = d['extra']
for k in ('verify_distinct_keys', 'verify_64_per_key', 'verify_one_signer', 'verify_keyset_e2e', 'verify_keyset_compact_e2e', 'two_batches_in_flight'):
    print(k, json.dumps(x.get(k)))
print(json.dumps(x.get('rlc_sweep_distinct_keys')))
P
tail -5 gpurun_out/${tag}_bench.err
