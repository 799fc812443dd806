#!/bin/bash
# usage (on the GPU box): tools/expbench.sh v1 v2 ...   -- opbench + verify bench for experimental builds libgoldilocks_b200/_exp_<v>.so
for v in base "$@"; do
  if [ "$v" = base ]; then unset GOLDILOCKS_B200_LIB; else export GOLDILOCKS_B200_LIB=$PWD/libgoldilocks_b200/_exp_$v.so; fi
  echo "== $v"
  python tools/opbench.py --ops comb,x448,decode 2>&1 | grep -v "^$"
  python bench.py --no-cpu --no-extra --steps 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('verify %.3f M/s  finish %.2f ms  decode %.2f ms' % (d['value']/1e6, d['roofline']['kernel_ms']['SlotEdVerifyFinish'], d['roofline']['kernel_ms']['LaneEdVerifyDecode']))"
done
