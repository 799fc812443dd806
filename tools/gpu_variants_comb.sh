#!/bin/bash
# usage (on the GPU box): tools/gpu_variants_comb.sh v1 v2 ... -- comb / window-scalarmul / sign timings for libgoldilocks_b200/_exp_<v>.so
for v in base "$@"; do
  if [ "$v" = base ]; then unset GOLDILOCKS_B200_LIB; else export GOLDILOCKS_B200_LIB=$PWD/libgoldilocks_b200/_exp_$v.so; fi
  echo "== $v"
  python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
  python tools/opbench.py --ops sign,scalarmul,comb 2>&1 | grep -v "^$" | cut -c1-160
done
