#!/bin/bash
# usage (on the GPU box): tools/gpu_ncu.sh <tag> [finish chain columns launches ktab scalars decode x448 comb ptadd]  -- one `ncu --set full` capture per kernel into gpurun_out/<tag>_<kernel>.ncu-rep
tag=${1:-run}; shift
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
for k in "${@:-finish}"; do
  case $k in
    finish) timeout 600 $NCU -k regex:SlotEdVerifyFinishShared -s 1 -c 1 -f -o gpurun_out/${tag}_finish python bench.py --steps 1 --warmup 3 --no-cpu --no-extra --no-peak > gpurun_out/${tag}_ncu_finish.log 2>&1 ;;
    scalars) timeout 600 $NCU -k regex:LaneEdVerifyScalars -s 1 -c 1 -f -o gpurun_out/${tag}_scalars python bench.py --steps 1 --warmup 3 --no-cpu --no-extra --no-peak > gpurun_out/${tag}_ncu_scalars.log 2>&1 ;;
    chain)  timeout 600 $NCU -k regex:SlotKeyChain -s 1 -c 1 -f -o gpurun_out/${tag}_chain python bench.py --steps 1 --warmup 3 --no-cpu --no-extra --no-peak > gpurun_out/${tag}_ncu_chain.log 2>&1 ;;
    columns) timeout 600 $NCU -k regex:SlotKeyColumns -s 1 -c 1 -f -o gpurun_out/${tag}_columns python bench.py --steps 1 --warmup 3 --no-cpu --no-extra --no-peak > gpurun_out/${tag}_ncu_columns.log 2>&1 ;;
    launches) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extra --no-peak > gpurun_out/${tag}_ncu_launches.log 2>&1 ;;
    ktab)   timeout 600 $NCU -k regex:SlotKeysetTables -s 1 -c 1 -f -o gpurun_out/${tag}_ktab python bench.py --steps 1 --warmup 3 --no-cpu --no-extra --no-peak > gpurun_out/${tag}_ncu_ktab.log 2>&1 ;;
    ptadd)  timeout 600 $NCU -k regex:k_pt_staged -s 1 -c 1 -f -o gpurun_out/${tag}_ptadd python tools/opbench.py --ops point_add --reps 1 > gpurun_out/${tag}_ncu_ptadd.log 2>&1 ;;
    decode) timeout 600 $NCU -k regex:LaneEdVerifyDecode -s 1 -c 1 -f -o gpurun_out/${tag}_decode python bench.py --steps 1 --warmup 3 --no-cpu --no-extra --no-peak > gpurun_out/${tag}_ncu_decode.log 2>&1 ;;
    x448)   timeout 600 $NCU -k regex:SlotX448 -s 1 -c 1 -f -o gpurun_out/${tag}_x448 python tools/opbench.py --ops x448 --reps 1 > gpurun_out/${tag}_ncu_x448.log 2>&1 ;;
    comb)   timeout 600 $NCU -k regex:SlotComb -s 1 -c 1 -f -o gpurun_out/${tag}_comb python tools/opbench.py --ops comb --reps 1 > gpurun_out/${tag}_ncu_comb.log 2>&1 ;;
  esac
  tail -2 gpurun_out/${tag}_ncu_$k.log
done
ls -la gpurun_out | tail -8
