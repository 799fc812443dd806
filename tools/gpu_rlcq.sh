#!/bin/bash
# usage (on the GPU box): tools/gpu_rlcq.sh  -- RLC parity test + timing (with a launch timeline) at 16 signatures per key and with distinct keys
timeout 600 python -m pytest tests -m gpu -x -q -k "eddsa_rlc or corner" 2>&1 | tail -3
timeout 300 python tools/rlcbench.py --timeline 2>&1 | tail -50
timeout 300 python tools/rlcbench.py --per-key 1 --reps 2 2>&1 | tail -20
