#!/bin/bash
# usage (on the GPU box): tools/gpu_ncu_r02h.sh  -- `ncu --set full` of the kernels whose occupancy changed late in round 2 (decode / codec at 3 / 4 resident
# blocks) and of the RLC kernels, plus compute-sanitizer over the a9 test entry point and the concurrent-thread path
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
timeout 600 $NCU -k regex:LaneRlcDecode -s 4 -c 1 -f -o gpurun_out/r02h_rlcdecode python tools/rlcbench.py --reps 1 > gpurun_out/r02h_ncu_rlcdecode.log 2>&1
timeout 600 $NCU -k regex:SlotRlcBucket -s 2 -c 1 -f -o gpurun_out/r02h_rlcbucket python tools/rlcbench.py --reps 1 > gpurun_out/r02h_ncu_rlcbucket.log 2>&1
timeout 600 $NCU -k regex:LanePtDecode -s 1 -c 1 -f -o gpurun_out/r02h_ptdecode python tools/opbench.py --ops decode --reps 1 > gpurun_out/r02h_ncu_ptdecode.log 2>&1
timeout 600 $NCU -k regex:SlotScalarmul -s 0 -c 1 -f -o gpurun_out/r02h_scalarmul python tools/opbench.py --ops scalarmul --n 262144 --reps 1 > gpurun_out/r02h_ncu_scalarmul.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
cat > /tmp/san_r02h.py <<'PY'
import sys, os, threading
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import libgoldilocks_b200 as g
import parity, util
lib = g.load(); chk = util.checker_lib()
parity.check_niels(lib, chk, 300)
n = 400
sig, pk, msgs, kinds = util.verify_corpus(chk, "san3/v", n)
want = lib.ed448_verify(sig, pk, msgs)
errs = []
def w(t):
    try:
        for _ in range(2):
            assert (lib.ed448_verify(sig, pk, msgs) == want).all()
            lib.point_encode(lib.from_hash_nonuniform(util.stream_bytes("san3/h%d" % t, 500 * 56).reshape(500, 56)))
    except Exception as e:
        errs.append(repr(e))
ts = [threading.Thread(target=w, args=(t,)) for t in range(4)]
[t.start() for t in ts]; [t.join() for t in ts]
assert not errs, errs
print("san r02h ok")
PY
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san_r02h.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|san r02h ok|Error|hazard|Assert|assert" | head -12
done | tee gpurun_out/r02h_sanitizer.txt
