#!/bin/bash
# usage (on the GPU box): tools/gpu_sanitize_r02.sh  -- compute-sanitizer memcheck + racecheck over what round 2 added: the device-set
# sharding (worker threads, three context lanes), the chunked RLC localisation pass with its packed fallback, the register-resident sponge
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
cat > /tmp/san_r02.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import libgoldilocks_b200 as g
from util import stream_bytes
lib = g.load()
os.environ["GOLDILOCKS_B200_RLC_CHUNK"] = "64"
lib.rlc_policy(0)
n = 64 * 6 + 9
rep = np.arange(n) % 29
sk = stream_bytes("san2/sk", 29 * 57).reshape(29, 57)[rep]
msgs = [bytes(stream_bytes("san2/m%d" % i, i % 70)) for i in range(n)]
pk = lib.ed448_derive_public_key(sk)
sig = lib.ed448_sign(sk, pk, msgs, context=b"san")
st, fast = lib.ed448_verify_rlc(sig, pk, msgs, context=b"san")
assert (st == -1).all() and fast == 1
sig[70, 80] ^= 1; sig[n - 1, 90] ^= 2                     # two failing chunks (one of them the short last one): localisation + packed fallback
st, fast = lib.ed448_verify_rlc(sig, pk, msgs, context=b"san")
assert fast == 2 and st[70] == 0 and st[n - 1] == 0 and (st == -1).sum() == n - 2, (fast, (st == 0).sum())
want = lib.ed448_verify(sig, pk, msgs, context=b"san")
assert (want == st).all()
h = stream_bytes("san2/h", 9000 * 56).reshape(9000, 56)
base = lib.point_encode(lib.from_hash_nonuniform(h))
x0, s0 = lib.x448(h[:8192], h[::-1][:8192].copy())
lib.set_devices([0, 0, 0])                                # the partition and the worker threads, on one GPU
try:
    assert (lib.point_encode(lib.from_hash_nonuniform(h)) == base).all()
    x1, s1 = lib.x448(h[:8192], h[::-1][:8192].copy())
    assert (x0 == x1).all() and (s0 == s1).all()
    big = np.tile(sig, (12, 1)); bpk = np.tile(pk, (12, 1)); bm = msgs * 12
    assert (lib.ed448_verify(big, bpk, bm, context=b"san") == np.tile(want, 12)).all()
finally:
    lib.set_devices([])
print("san r02 ok")
PY
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san_r02.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|san r02 ok|Error|hazard|Assert|assert" | head -12
done | tee gpurun_out/r02f_sanitizer.txt
