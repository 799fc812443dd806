"""CPU-tier checks of bench.py's bookkeeping: the executed-work table it divides by (profiles/executed_ops.json, written by
tools/count_ops.py from the host simulator) must carry every entry the bench line looks up, and the figures must be the ones the
host simulator counts NOW -- a kernel change that is not followed by `python tools/count_ops.py` would silently skew roofline.frac."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import util  # noqa: E402


def test_executed_ops_table_has_every_entry_the_bench_uses():
    import bench
    ops = bench.executed_ops()
    for key in ("gf_mul", "point_add", "point_double", "comb", "x448", "encode", "decode", "elligator", "elligator_uniform", "sign",
                "derive_public_key", "point_scalarmul", "bdsm", "verify_16_per_key", "verify_distinct", "verify_64_per_key",
                "verify_one_signer", "verify_keyset", "verify_keyset_compact", "finish_shared", "finish_alone", "key_table"):
        assert ops[key] > 0, key
    # orderings that follow from the algorithms: fewer doublings the more signatures share a key; flat key-set tables have none
    assert ops["verify_keyset"] < ops["verify_one_signer"] < ops["verify_64_per_key"] < ops["verify_16_per_key"] < ops["verify_distinct"]
    assert ops["verify_keyset"] < ops["verify_keyset_compact"]


def test_executed_ops_table_matches_the_host_simulator_now():
    sim = util.hostsim_lib()
    chk = util.checker_lib()
    util.set_threads(sim, os.cpu_count() or 1)
    with open(os.path.join(ROOT, "profiles", "executed_ops.json")) as f:
        table = json.load(f)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import count_ops
    sim.lib.goldilocks_b200_export_wnaf_table(np.zeros(32 * 192, np.uint8).ctypes.data_as(__import__("ctypes").c_void_p))   # tables first
    n, per = 256, 16
    sk = util.stream_bytes("ops/sk", n * 57).reshape(n, 57)
    pk = chk.ed448_derive_public_key(sk)
    msgs = [bytes(util.stream_bytes("ops/m%d" % i, 32)) for i in range(n)]
    nk = n // per
    sk16, pk16 = np.repeat(sk[:nk], per, axis=0), np.repeat(pk[:nk], per, axis=0)
    sig16 = chk.ed448_sign(sk16, pk16, msgs)
    count_ops.stage_counts(sim)
    assert (sim.ed448_verify(sig16, pk16, msgs) == -1).all()
    now = count_ops.per_unit(count_ops.stage_counts(sim), n)
    assert now["_total_imad_wide"] == table["ed448_verify_16_per_key"]["_total_imad_wide"], \
        "profiles/executed_ops.json is stale: run `python tools/count_ops.py`"
