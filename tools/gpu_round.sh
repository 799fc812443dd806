#!/bin/bash
# usage (on the GPU box, via gpurun): tools/gpu_round.sh <tag>  -- tests + bench + launch list + ncu full captures into gpurun_out/<tag>_*
tag=${1:-run}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/${tag}_pytest.log 2>&1
tail -3 gpurun_out/${tag}_pytest.log
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cat gpurun_out/${tag}_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err
cat gpurun_out/${tag}_bench_ref.json
timeout 300 python tools/opbench.py --ops sign,scalarmul,elligator,comb,x448,gf_mul,point_add,point_double,decode,encode > gpurun_out/${tag}_opbench.txt 2>&1
cat gpurun_out/${tag}_opbench.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
tools/gpu_ncu.sh $tag finish ktab decode x448 comb ptadd
