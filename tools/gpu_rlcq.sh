#!/bin/bash
# usage (on the GPU box): tools/gpu_rlcq.sh [r,k ...]  -- RLC parity test + timing (with a launch timeline); extra arguments = other
# resident-block splits of the two bucket kernels (GOLDILOCKS_B200_RLC_BLOCKS) to time
timeout 600 python -m pytest tests -m gpu -x -q -k "eddsa_rlc or corner" 2>&1 | tail -3
timeout 300 python tools/rlcbench.py --timeline 2>&1 | tail -36
timeout 300 python tools/rlcbench.py --per-key 1 --reps 2 2>&1 | head -1
for v in "$@"; do
  echo "== blocks $v"
  GOLDILOCKS_B200_RLC_BLOCKS=$v timeout 300 python tools/rlcbench.py 2>&1 | head -1
  GOLDILOCKS_B200_RLC_BLOCKS=$v timeout 300 python tools/rlcbench.py --per-key 1 --reps 2 2>&1 | head -1
done
