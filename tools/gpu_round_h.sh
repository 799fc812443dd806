#!/bin/bash
# usage (on the GPU box, via gpurun): tools/gpu_round_h.sh <tag>  -- tests + bench (+ reference arm) + RLC timing/timeline + RLC launch list
tag=${1:-run}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${tag}_pytest.log 2>&1
tail -3 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cut -c1-300 gpurun_out/${tag}_bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err
( timeout 200 python tools/rlcbench.py --timeline; timeout 200 python tools/rlcbench.py --per-key 1 --reps 2; timeout 200 python tools/rlcbench.py --per-key 4 --reps 2 | head -1; timeout 200 python tools/rlcbench.py --n 65536 | head -1 ) > gpurun_out/${tag}_rlcbench.txt 2>&1
grep "rlc n=" gpurun_out/${tag}_rlcbench.txt
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_rlc_launches.csv \
    python tools/rlcbench.py --reps 1 > /dev/null 2>&1
ls gpurun_out | grep ${tag}
