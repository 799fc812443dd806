#!/bin/bash
# usage (on the GPU box): tools/gpu_variants.sh v1 v2 ...  -- parity smoke + opbench + verify bench for libgoldilocks_b200/_exp_<v>.so (kernel experiments)
for v in base "$@"; do
  if [ "$v" = base ]; then unset GOLDILOCKS_B200_LIB; else export GOLDILOCKS_B200_LIB=$PWD/libgoldilocks_b200/_exp_$v.so; fi
  echo "== $v"
  python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
  python tools/opbench.py --ops scalarmul,comb,x448 2>&1 | grep -v "^$" | cut -c1-150
  python bench.py --no-cpu --no-extra --steps 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('verify %.3f M/s  e2e %.3f  finish %.2f ms  decode %.2f ms' % (d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['kernel_ms'].get('SlotEdVerifyFinishShared', d['roofline']['kernel_ms'].get('SlotEdVerifyFinish', 0)), d['roofline']['kernel_ms']['LaneEdVerifyDecode']))"
done
