#!/bin/bash
# usage (on the GPU box): tools/gpu_rlcq.sh  -- RLC parity test + timing at 16 signatures per key and with distinct keys
timeout 600 python -m pytest tests -m gpu -x -q -k "eddsa_rlc or corner" 2>&1 | tail -3
timeout 300 python tools/rlcbench.py 2>&1 | tail -20
timeout 300 python tools/rlcbench.py --per-key 1 --reps 2 2>&1 | tail -20
for v in "$@"; do
  echo "== variant $v"
  GOLDILOCKS_B200_LIB=$PWD/libgoldilocks_b200/_exp_$v.so timeout 300 python tools/rlcbench.py 2>&1 | tail -20
done
