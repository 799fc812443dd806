#!/usr/bin/env python3
"""Executed multiplier work of every entry point, counted -- not estimated.

The host simulator (tests/hostsim, the CUDA headers compiled for the host with -DGF_COUNT_OPS) runs the very functors
the kernels run and counts calls of gf_mul_body / gf_sqr_body / gf_mulw per functor.  One call = 193 / 110 / 16
IMAD.WIDE on the device (csrc/gf.cuh; `cuobjdump -sass` of the bodies, profiles/sass/).  The result,
profiles/executed_ops.json, is what bench.py divides by the measured IMAD.WIDE peak for `roofline.frac` and every
`extra.*.imad_frac_executed`.

Device-only differences (stated in the JSON): the final inversion of X448 and of the EdDSA/X448 encoders is shared by
four lanes of a block on the device (slot_algos.cuh s_block_invert4) while the host build inverts per lane, so those
kernels' device figure = host figure - 3/4 of one inversion (446 S + 13 M -> shared) + the 9 M of Montgomery's trick.

    python tools/count_ops.py            # writes profiles/executed_ops.json
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402

IMAD = {"mul": 193, "sqr": 110, "mulw": 16}
INV = (13, 447, 0)          # gf_invert = isr(x^2)^2 * x : 446 + 1 squarings (one more for x^2), 13 + ... see count below


def stage_counts(sim, reset=True):
    names = C.create_string_buffer(64 * 256)
    counts = (C.c_ulonglong * (4 * 256))()
    f = sim.lib.hostsim_stage_counts
    f.restype = C.c_size_t
    k = f(names, counts, C.c_size_t(256), C.c_int(1 if reset else 0))
    out = {}
    for i in range(k):
        nm = names.raw[64 * i:64 * i + 64].split(b"\0")[0].decode()
        m, s, w, n = (int(counts[4 * i + j]) for j in range(4))
        out[nm] = {"mul": m, "sqr": s, "mulw": w, "lanes": n}
    return out


def per_unit(st, units):
    out = {}
    for nm, c in st.items():
        m, s, w = c["mul"] / units, c["sqr"] / units, c["mulw"] / units
        out[nm] = {"mul": round(m, 3), "sqr": round(s, 3), "mulw": round(w, 3), "imad_wide": round(m * IMAD["mul"] + s * IMAD["sqr"] + w * IMAD["mulw"], 1),
                   "lanes_per_unit": round(c["lanes"] / units, 4)}
    out["_total_imad_wide"] = round(sum(v["imad_wide"] for v in out.values()), 1)
    return out


def main():
    sim = util.hostsim_lib()
    util.set_threads(sim, os.cpu_count() or 1)
    chk = util.checker_lib()
    res = {"_imad_wide_per_call": IMAD,
           "_note": "counted on the host simulator of the CUDA sources (tools/count_ops.py); per unit = per element of the batch"}

    def measure(name, units, fn):
        stage_counts(sim)
        fn()
        res[name] = per_unit(stage_counts(sim), units)
        print("%-28s %s" % (name, json.dumps(res[name])))

    n = 256
    sim.export_wnaf_table() if hasattr(sim, "export_wnaf_table") else sim.lib.goldilocks_b200_export_wnaf_table(np.zeros(32 * 192, np.uint8).ctypes.data_as(C.c_void_p))   # init-time tables built before anything is counted
    a, b = util.field_inputs("ops/f", n)
    measure("gf_mul", len(a), lambda: sim.gf_mul(a, b))
    measure("gf_sqr", len(a), lambda: sim.gf_sqr(a))
    measure("gf_invert", len(a), lambda: sim.gf_invert(a))
    p, q = util.random_points(chk, "ops/p", n), util.random_points(chk, "ops/q", n)
    measure("point_add", n, lambda: sim.point_add(p, q))
    measure("point_double", n, lambda: sim.point_double(p))
    sc = util.random_scalars(chk, "ops/s", n)
    sc2 = util.random_scalars(chk, "ops/t", n)
    measure("precomputed_scalarmul", n, lambda: sim.precomputed_scalarmul(sc))
    measure("point_scalarmul", n, lambda: sim.point_scalarmul(p, sc))
    measure("base_double_scalarmul_non_secret", n, lambda: sim.base_double_scalarmul_non_secret(sc, p, sc2))
    u = util.stream_bytes("ops/u", n * 56).reshape(n, 56)
    k = util.stream_bytes("ops/k", n * 56).reshape(n, 56)
    measure("x448", n, lambda: sim.x448(u, k))
    measure("x448_derive_public_key", n, lambda: sim.x448_derive_public_key(k))
    h = util.stream_bytes("ops/h", n * 112).reshape(n, 112)
    measure("from_hash_nonuniform", n, lambda: sim.from_hash_nonuniform(h[:, :56]))
    measure("from_hash_uniform", n, lambda: sim.from_hash_uniform(h))
    ser = chk.point_encode(p)
    measure("point_encode", n, lambda: sim.point_encode(p))
    measure("point_decode", n, lambda: sim.point_decode(ser))
    sk = util.stream_bytes("ops/sk", n * 57).reshape(n, 57)
    pk = chk.ed448_derive_public_key(sk)
    msgs = [bytes(util.stream_bytes("ops/m%d" % i, 32)) for i in range(n)]
    measure("ed448_derive_public_key", n, lambda: sim.ed448_derive_public_key(sk))
    measure("ed448_sign", n, lambda: sim.ed448_sign(sk, pk, msgs))
    sig = chk.ed448_sign(sk, pk, msgs)
    measure("ed448_verify_distinct_keys", n, lambda: sim.ed448_verify(sig, pk, msgs))
    # the bench corpus shape: 16 signatures per key
    per = 16
    nk = n // per
    sk16, pk16 = np.repeat(sk[:nk], per, axis=0), np.repeat(pk[:nk], per, axis=0)
    sig16 = chk.ed448_sign(sk16, pk16, msgs)
    measure("ed448_verify_16_per_key", n, lambda: sim.ed448_verify(sig16, pk16, msgs))
    # the other column shapes of the batch path (vsh_pick): 64 signatures per key -> 30 x 3, one signer -> 90 x 1 (additions only)
    for per2, name2 in ((64, "ed448_verify_64_per_key"), (n, "ed448_verify_one_signer")):
        skp, pkp = np.repeat(sk[:n // per2], per2, axis=0), np.repeat(pk[:n // per2], per2, axis=0)
        sigp = chk.ed448_sign(skp, pkp, msgs)
        measure(name2, n, lambda: sim.ed448_verify(sigp, pkp, msgs))
    handle = sim.keyset_create(pk[:nk])
    stage_counts(sim)
    measure("ed448_verify_keyset", n, lambda: sim.ed448_verify_keyset(handle, (np.arange(n) // per).astype(np.uint32), sig16, msgs))
    sim.keyset_destroy(handle)
    sim.keyset_policy(0)                                            # the compact layout (ten columns per key)
    handle = sim.keyset_create(pk[:nk])
    sim.keyset_policy(32 << 30)
    stage_counts(sim)
    measure("ed448_verify_keyset_compact", n, lambda: sim.ed448_verify_keyset(handle, (np.arange(n) // per).astype(np.uint32), sig16, msgs))
    sim.keyset_destroy(handle)
    inv = res["gf_invert"]["LaneGfILi6EE" if "LaneGfILi6EE" in res["gf_invert"] else next(k for k in res["gf_invert"] if not k.startswith("_"))]
    res["_device_only"] = {
        "block_shared_inversion": "SlotX448, SlotX448DerivePk, SlotEdDerivePk, SlotEdSignR: on the device one inversion serves four lanes "
                                  "(s_block_invert4); device imad_wide = host figure - 0.75 x %.0f + 9 x 193" % inv["imad_wide"],
        "gf_invert_imad_wide": inv["imad_wide"]}
    out = os.path.join(ROOT, "profiles", "executed_ops.json")
    with open(out, "w") as f:
        json.dump(res, f, indent=1, sort_keys=True)
    print("wrote", out)


if __name__ == "__main__":
    main()
