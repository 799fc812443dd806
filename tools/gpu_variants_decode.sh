for v in "" d3 d4; do
  if [ -n "$v" ]; then export GOLDILOCKS_B200_LIB=$PWD/libgoldilocks_b200/_exp_$v.so; else unset GOLDILOCKS_B200_LIB; fi
  echo "== variant ${v:-default}"
  python tools/rlcbench.py 2>&1 | grep -E "rlc n=|LaneRlcDecode"
  python tools/opbench.py --ops decode,encode --reps 5 2>&1 | tail -2
  python - <<'PY'
import sys, os, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tests")
import numpy as np, ctypes as C
import libgoldilocks_b200 as g
from util import stream_bytes
lib = g.load()
n = 1 << 20
h = stream_bytes("v/h", n * 56).reshape(n, 56)
lib.from_hash_nonuniform(h[:1000])
L = lib.lib
L.goldilocks_b200_profile_read.restype = C.c_size_t
names = C.create_string_buffer(64 * 64); ms = (C.c_float * 64)()
L.goldilocks_b200_profile(C.c_int(1)); lib.from_hash_nonuniform(h); L.goldilocks_b200_profile(C.c_int(0))
cnt = L.goldilocks_b200_profile_read(names, ms, C.c_size_t(64))
print("elligator kernel ms", [round(ms[k], 3) for k in range(cnt)])
PY
done
