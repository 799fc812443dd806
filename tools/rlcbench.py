"""Times goldilocks_ed448_verify_rlc_batch (host pointers, pinned) on an all-valid corpus of the bench shape and prints
the per-kernel CUDA-event times of the library's profile hooks.  usage: python tools/rlcbench.py [--n N] [--per-key P]"""
import argparse
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import bench
    import libgoldilocks_b200 as g
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1 << 20)
    ap.add_argument("--per-key", type=int, default=16)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--timeline", action="store_true")
    ap.add_argument("--bad", type=int, default=0, help="corrupt this many signatures (spread evenly) before the timeline run: shows the localisation pass")
    a = ap.parse_args()
    lib = g.load()
    n = a.n
    lib.rlc_policy(0)
    sig, pk, arena, off, expect = bench.make_corpus(lib, n, "rlcbench", per=a.per_key, corrupt=False)
    pin = lambda x: torch.from_numpy(np.ascontiguousarray(x)).pin_memory()
    h_sig, h_pk, h_msg, h_off = pin(sig.reshape(-1)), pin(pk.reshape(-1)), pin(arena), pin(off.view(np.int64))
    h_st = torch.empty(n, dtype=torch.int32).pin_memory()
    fr = lib.lib.goldilocks_ed448_verify_rlc_batch
    fr.restype = C.c_int32
    fast = C.c_int(0)
    argr = [C.c_void_p(h_st.data_ptr()), C.c_void_p(h_sig.data_ptr()), C.c_void_p(h_pk.data_ptr()), C.c_void_p(h_msg.data_ptr()),
            C.c_void_p(h_off.data_ptr()), C.c_uint8(0), None, C.c_uint8(0), C.c_size_t(n), C.byref(fast)]
    for _ in range(2):
        assert fr(*argr) == -1
    assert (h_st.numpy() == -1).all() and fast.value == 1
    lib.lib.goldilocks_b200_profile(C.c_int(1))
    t0 = time.perf_counter()
    for _ in range(a.reps):
        assert fr(*argr) == -1
    t = (time.perf_counter() - t0) / a.reps
    lib.lib.goldilocks_b200_profile(C.c_int(0))
    names = C.create_string_buffer(64 * 4096)
    ms = (C.c_float * 4096)()
    lib.lib.goldilocks_b200_profile_read.restype = C.c_size_t
    cnt = lib.lib.goldilocks_b200_profile_read(names, ms, C.c_size_t(4096))
    k = {}
    for i in range(cnt):
        nm = names.raw[64 * i:64 * i + 64].split(b"\0")[0].decode()
        k.setdefault(nm, []).append(ms[i])
    print("rlc n=%d per_key=%d: %.2f ms/call = %.2f M verifies/s (host pointers)" % (n, a.per_key, t * 1e3, n / t / 1e6))
    for nm, v in k.items():
        per = len(v) // a.reps
        print("  %-22s %s ms" % (nm, " + ".join("%.3f" % (sum(v[j::per]) / a.reps) for j in range(per))))
    print("  sum of kernels %.2f ms" % (sum(sum(v) for v in k.values()) / a.reps))
    if a.timeline:
        if a.bad:
            idx = np.arange(n // (2 * a.bad), n, n // a.bad)[: a.bad]
            sig2 = sig.copy(); sig2[idx, 70] ^= 1
            h_sig.copy_(torch.from_numpy(sig2.reshape(-1)))
            assert fr(*argr) == -1       # warm-up of the localisation pass (grows the arena)
            print("with %d bad signatures: fast_path = %d, rejected = %d" % (a.bad, fast.value, int((h_st.numpy() == 0).sum())))
        lib.lib.goldilocks_b200_profile(C.c_int(1))
        t0 = time.perf_counter()
        assert fr(*argr) == -1
        wall = time.perf_counter() - t0
        lib.lib.goldilocks_b200_profile(C.c_int(0))
        s0, s1 = (C.c_float * 4096)(), (C.c_float * 4096)()
        lib.lib.goldilocks_b200_profile_timeline.restype = C.c_size_t
        cnt = lib.lib.goldilocks_b200_profile_timeline(names, s0, s1, C.c_size_t(4096))
        print("timeline of one call (wall %.2f ms; ms after the first kernel began):" % (wall * 1e3))
        for i in range(cnt):
            print("  %7.3f .. %7.3f  %s" % (s0[i], s1[i], names.raw[64 * i:64 * i + 64].split(b"\0")[0].decode()))


if __name__ == "__main__":
    main()
