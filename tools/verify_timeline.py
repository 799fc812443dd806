"""Kernel timeline of one goldilocks_ed448_verify_batch call (host pointers, pinned) at batch n: where a small shard of a
strong-scaling run spends its time.  usage: python tools/verify_timeline.py [--n N] [--per-key P]"""
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
    ap.add_argument("--n", type=int, default=1 << 17)
    ap.add_argument("--per-key", type=int, default=16)
    a = ap.parse_args()
    lib = g.load()
    for n in (a.n, a.n * 2, a.n * 4, a.n * 8):
        sig, pk, arena, off, expect = bench.make_corpus(lib, n, "vtl", per=a.per_key)
        pin = lambda x: torch.from_numpy(np.ascontiguousarray(x)).pin_memory()
        h_sig, h_pk, h_msg, h_off = pin(sig.reshape(-1)), pin(pk.reshape(-1)), pin(arena), pin(off.view(np.int64))
        h_st = torch.empty(n, dtype=torch.int32).pin_memory()
        fn = lib.lib.goldilocks_ed448_verify_batch
        fn.restype = C.c_int32
        argv = [C.c_void_p(h_st.data_ptr()), C.c_void_p(h_sig.data_ptr()), C.c_void_p(h_pk.data_ptr()), C.c_void_p(h_msg.data_ptr()),
                C.c_void_p(h_off.data_ptr()), C.c_uint8(0), None, C.c_uint8(0), C.c_size_t(n)]
        for _ in range(3):
            assert fn(*argv) == -1
        t0 = time.perf_counter()
        for _ in range(5):
            assert fn(*argv) == -1
        wall = (time.perf_counter() - t0) / 5
        lib.lib.goldilocks_b200_profile(C.c_int(1))
        assert fn(*argv) == -1
        lib.lib.goldilocks_b200_profile(C.c_int(0))
        names = C.create_string_buffer(64 * 256)
        s0, s1 = (C.c_float * 256)(), (C.c_float * 256)()
        lib.lib.goldilocks_b200_profile_timeline.restype = C.c_size_t
        cnt = lib.lib.goldilocks_b200_profile_timeline(names, s0, s1, C.c_size_t(256))
        print("n = %d: %.2f ms per call = %.2f M verifies/s" % (n, wall * 1e3, n / wall / 1e6))
        for i in range(cnt):
            print("  %7.3f .. %7.3f  %s" % (s0[i], s1[i], names.raw[64 * i:64 * i + 64].split(b"\0")[0].decode()))


if __name__ == "__main__":
    main()
