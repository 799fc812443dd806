#!/bin/bash
# round-2 final run: native single-call tool, then the default bench line (strong leg included)
mkdir -p gpurun_out
tools/single_calls 1024 16 250 > gpurun_out/r02l_single_calls.json 2> gpurun_out/r02l_single_calls.err; echo "single_calls rc=$?"
cat gpurun_out/r02l_single_calls.json; tail -3 gpurun_out/r02l_single_calls.err
tools/single_calls 256 32 100 2>&1 | tail -2
python bench.py > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err; echo "bench rc=$?"
python - <<'P'
import json
d = json.load(open('gpurun_out/r02l_bench.json'))
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['steps'], d['roofline']['frac'])
print(json.dumps(d['extra']['single_calls_from_threads'])[:1500])
print(json.dumps(d['extra']['verify_distinct_keys']))
P
