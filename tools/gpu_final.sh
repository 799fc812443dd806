#!/bin/bash
# the round's closing run on one B200: smoke, GPU tier, default bench line, reference arm
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02v_pytest.log; tail -3 gpurun_out/r02v_pytest.log
python bench.py > gpurun_out/r02v_bench.json 2> gpurun_out/r02v_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02v_bench_ref.json 2> gpurun_out/r02v_bench_ref.err; echo "ref rc=$?"
python - <<'P'
import json
d = json.load(open('gpurun_out/r02v_bench.json')); r = json.load(open('gpurun_out/r02v_bench_ref.json'))
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['steps'], d['roofline']['frac'], d['gpu_launches'], d['clocks'])
print('reference', r['value'], r.get('cpu_baseline'))
print('ratio e2e/reference', d['e2e']['value'] / r['value'])
x = d['extra']
print({k: (round(v['value'] / 1e6, 2) if isinstance(v, dict) and 'value' in v else None) for k, v in x.items()})
print(json.dumps(x.get('strong'))[:1200])
P
