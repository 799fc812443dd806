#!/bin/bash
# Decode / codec kernels at 3 and 4 resident blocks against the default (launch.cuh).  Build the variants HERE first:
#   tools/build_variant.sh d3 -DDECODE_MIN_BLOCKS=3 -DCODEC_MIN_BLOCKS=3 ; tools/build_variant.sh d4 -DDECODE_MIN_BLOCKS=4 -DCODEC_MIN_BLOCKS=4
# then on the GPU box: bash tools/gpu_variants_decode.sh   (libgoldilocks_b200/_exp_*.so travel with the snapshot)
for v in "" d3 d4; do
  if [ -n "$v" ]; then export GOLDILOCKS_B200_LIB=$PWD/libgoldilocks_b200/_exp_$v.so; else unset GOLDILOCKS_B200_LIB; fi
  [ -n "$v" ] && [ ! -f "$GOLDILOCKS_B200_LIB" ] && continue
  echo "== variant ${v:-default}"
  python tools/rlcbench.py 2>&1 | grep -E "rlc n=|LaneRlcDecode"
  python tools/opbench.py --ops decode,encode --reps 5 2>&1 | tail -2
done
