#!/bin/bash
# usage (on the GPU box): tools/gpu_sanitize_rlc.sh  -- compute-sanitizer memcheck + racecheck over the RLC verification path (small batches)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
cat > /tmp/san_rlc.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import libgoldilocks_b200 as g
from util import stream_bytes
lib = g.load()
n = 300
rep = np.arange(n) % 23                    # 23 keys, ~13 signatures each
sk = stream_bytes("sanr/sk", 23 * 57).reshape(23, 57)[rep]
msgs = [bytes(stream_bytes("sanr/m%d" % i, i % 40)) for i in range(n)]
pk = lib.ed448_derive_public_key(sk)
sig = lib.ed448_sign(sk, pk, msgs)
st, fast = lib.ed448_verify_rlc(sig, pk, msgs)          # grouping, decodes, weights, both bucket classes, tree, verdict
assert (st == -1).all() and fast == 1
sig[7, :57] = 0; sig[7, 0] = 1                           # undecodable R: rejected up front
st, fast = lib.ed448_verify_rlc(sig, pk, msgs)
assert fast == 1 and st[7] == 0 and (st == -1).sum() == n - 1
sig[9, 70] ^= 1                                          # wrong S: equation fails, per-signature path
st, fast = lib.ed448_verify_rlc(sig, pk, msgs)
assert fast == 0 and st[7] == 0 and st[9] == 0 and (st == -1).sum() == n - 2
print("san rlc ok")
PY
for tool in memcheck racecheck; do
  timeout 400 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san_rlc.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|san rlc ok|Error|hazard" | head -12
done | tee gpurun_out/r01g_sanitizer_rlc.txt
