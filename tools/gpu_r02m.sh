#!/bin/bash
# usage (on the GPU box): tools/gpu_r02m.sh -- `ncu --set full` of the finish kernel on stand-alone signatures (half-size multipliers) and of the
# half-gcd kernel, then compute-sanitizer (memcheck, racecheck) over the half-size paths (grouped and n < 64) and the gathered one-element calls
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
timeout 600 $NCU -k regex:SlotEdVerifyFinishShared -s 3 -c 1 -f -o gpurun_out/r02m_finish_half python tools/verify_timeline.py --n 131072 --per-key 1 > gpurun_out/r02m_ncu_finish_half.log 2>&1
timeout 600 $NCU -k regex:LaneVerifyHalf -s 3 -c 1 -f -o gpurun_out/r02m_halfgcd python tools/verify_timeline.py --n 131072 --per-key 1 > gpurun_out/r02m_ncu_halfgcd.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
cat > /tmp/san_r02m.py <<'PY'
import sys, os, threading
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import libgoldilocks_b200 as g
import util
lib = g.load(); chk = util.checker_lib()
for n in (40, 300):                                       # n < 64: every signature stand-alone, no grouping; 300: grouped path, all keys distinct
    sig, pk, msgs, kinds = util.verify_corpus(chk, "san4/v%d" % n, n, corrupt_every=3)
    want = chk.ed448_verify(sig, pk, msgs)
    assert (lib.ed448_verify(sig, pk, msgs) == want).all(), n
errs = []
lib.coalesce(300)
def w(t):
    try:
        for j in range(6):
            i = (t * 6 + j) % 300
            if lib.ed448_verify_one(sig[i], pk[i], msgs[i]) != want[i]: errs.append("verify %d" % i)
            o, st = lib.x448_one(sig[i, :56], pk[i, :56])
    except Exception as e:
        errs.append(repr(e))
ts = [threading.Thread(target=w, args=(t,)) for t in range(8)]
[t.start() for t in ts]; [t.join() for t in ts]
lib.coalesce(0)
assert not errs, errs
print("san r02m ok", lib.coalesce_stats())
PY
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san_r02m.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|san r02m ok|Error|hazard|Assert|assert" | head -12
done | tee gpurun_out/r02m_sanitizer.txt
