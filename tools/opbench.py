#!/usr/bin/env python3
"""Quick device-resident timing of single entry points (development aid; bench.py is the contract).
usage: tools/opbench.py [--n 1048576] [--ops comb,x448,gf_mul,point_add,point_double,decode,encode] [--reps 3]
Select another build of the CUDA library with GOLDILOCKS_B200_LIB=/path/to/lib.so."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from libgoldilocks_b200.engine import DeviceEngine
from util import stream_bytes
MAC = {"comb": 142656, "x448": 870208, "gf_mul": 192, "point_add": 1552, "point_double": 1536, "decode": 90064, "encode": 89696}
def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--n", type=int, default=1 << 20); ap.add_argument("--ops", default="comb,x448,gf_mul,point_add,point_double,decode,encode"); ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args(); n = a.n
    eng = DeviceEngine(); lib = eng.capi; dev = torch.device("cuda")
    rnd = lambda lab, w: torch.from_numpy(stream_bytes("opbench/" + lab, n * w)).to(dev)
    sc = torch.from_numpy(lib.scalar_decode_long(stream_bytes("opbench/sc", n * 56).reshape(n, 56), 56).reshape(-1)).to(dev)
    pts = torch.empty(n * 256, dtype=torch.uint8, device=dev); pts2 = torch.empty_like(pts); pts3 = torch.empty_like(pts)
    eng.precomputed_scalarmul(pts, sc); eng.point_double(pts2, pts)
    u, k = rnd("u", 56), rnd("k", 56); o56 = torch.empty(n * 56, dtype=torch.uint8, device=dev); st = torch.empty(n, dtype=torch.int32, device=dev)
    ser = torch.empty_like(o56); eng.point_encode(ser, pts)
    fns = {"comb": lambda: eng.precomputed_scalarmul(pts3, sc), "x448": lambda: eng.x448(o56, st, u, k), "gf_mul": lambda: eng.gf_mul(o56, u, k),
           "point_add": lambda: eng.point_add(pts3, pts, pts2), "point_double": lambda: eng.point_double(pts3, pts),
           "decode": lambda: eng.point_decode(pts3, st, ser), "encode": lambda: eng.point_encode(o56, pts)}
    def kernel_ms(fn):
        """per-kernel CUDA-event times of one host-API call (library profile hooks)"""
        import ctypes as C
        lib.lib.goldilocks_b200_profile(C.c_int(1)); fn(); lib.lib.goldilocks_b200_profile(C.c_int(0))
        names = C.create_string_buffer(64 * 256); ms = (C.c_float * 256)()
        lib.lib.goldilocks_b200_profile_read.restype = C.c_size_t
        cnt = lib.lib.goldilocks_b200_profile_read(names, ms, C.c_size_t(256))
        return [(names.raw[64 * k:64 * k + 64].split(b"\0")[0].decode(), round(ms[k], 3)) for k in range(cnt)]
    if "sign" in a.ops.split(","):
        m = min(n, 1 << 18)
        sk = stream_bytes("opbench/sk", m * 57).reshape(m, 57)
        arena = stream_bytes("opbench/msg", m * 32); off = np.arange(m + 1, dtype=np.uint64) * 32
        pk = lib.ed448_derive_public_key(sk)
        print("derive_pk n=%d kernels %s" % (m, kernel_ms(lambda: lib.ed448_derive_public_key(sk))))
        ks = kernel_ms(lambda: lib.ed448_sign(sk, pk, (arena, off)))
        tot = sum(t for _, t in ks)
        print("sign      n=%d kernels %s  -> %.2f M signatures/s (kernels only)" % (m, ks, m / tot / 1e3))
        a.ops = ",".join(o for o in a.ops.split(",") if o != "sign")
    if "elligator" in a.ops.split(","):
        m = min(n, 1 << 20)
        h = stream_bytes("opbench/h2c", m * 112).reshape(m, 112)
        which = (np.arange(m) % 8).astype(np.uint32)
        hp = lib.from_hash_uniform(h)
        peak_ = json.load(open(os.path.join(ROOT, "profiles", "r01_imad_peak.json")))["imad_wide_u32_gmac_s"]
        for name, fn, mac in (("from_hash_nonuniform", lambda: lib.from_hash_nonuniform(h[:, :56]), 90672), ("from_hash_uniform", lambda: lib.from_hash_uniform(h), 182896),
                              ("invert_elligator_nonuniform", lambda: lib.invert_elligator_nonuniform(hp, which), 2 * 90000),
                              ("invert_elligator_uniform", lambda: lib.invert_elligator_uniform(hp, h[:, 56:], which), 3 * 90000)):
            fn()
            ks = kernel_ms(fn)
            tot = sum(t for _, t in ks)
            print("%-30s n=%d kernels %s -> %.2f Mops/s imad_frac %.3f" % (name, m, ks, m / tot / 1e3, m * mac / (tot / 1e3) / 1e9 / peak_))
        a.ops = ",".join(o for o in a.ops.split(",") if o != "elligator")
    if "scalarmul" in a.ops.split(","):
        m = min(n, 1 << 18)
        hs = lib.scalar_decode_long(stream_bytes("opbench/vs", m * 56).reshape(m, 56), 56)
        hp = lib.precomputed_scalarmul(hs[::-1].copy())
        for name, fn, mac in (("point_scalarmul", lambda: lib.point_scalarmul(hp, hs), 760592),
                              ("base_double_scalarmul_non_secret", lambda: lib.base_double_scalarmul_non_secret(hs, hp, hs[::-1].copy()), 800300),
                              ("point_double_scalarmul", lambda: lib.point_double_scalarmul(hp, hs, hp[::-1].copy(), hs[::-1].copy()), 1500000)):
            fn()
            ks = kernel_ms(fn)
            tot = sum(t for _, t in ks)
            print("%-34s n=%d kernels %s -> %.2f Mops/s imad_frac %.3f" % (name, m, ks, m / tot / 1e3, m * mac / (tot / 1e3) / 1e9 / json.load(open(os.path.join(ROOT, "profiles", "r01_imad_peak.json")))["imad_wide_u32_gmac_s"]))
        a.ops = ",".join(o for o in a.ops.split(",") if o != "scalarmul")
    peak = json.load(open(os.path.join(ROOT, "profiles", "r01_imad_peak.json")))["imad_wide_u32_gmac_s"]
    for op in [o for o in a.ops.split(",") if o]:
        f = fns[op]; f(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps): f()
        e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 1e3 / a.reps
        print("%-13s n=%d  %9.3f ms  %8.2f Mops/s  imad_frac %.3f" % (op, n, t * 1e3, n / t / 1e6, n * MAC[op] / t / 1e9 / peak))
if __name__ == "__main__":
    main()
