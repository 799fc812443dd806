#!/usr/bin/env python3
"""bench.py -- BASELINE.json's metric on its own config: Ed448 verifies/s at batch 2^20 (config 4),
with X448 ops/s (config 3) and fixed-base comb scalarmuls/s (config 2) at 2^20 riding along in `extra`.

  python bench.py [--gpus N] [--steps K] [--warmup W]          our arm  (under torchrun for N > 1)
  python bench.py --impl reference [...]                        the reference's CPU path on all host cores

A step = one pass of the hot path (goldilocks_ed448_verify_batch) over one batch of 2^20 synthetic
signatures per GPU (weak scaling: the batch shards into independent contiguous ranges, one per rank,
no collective on the data path; torch.distributed is used only for the barrier and max-over-ranks time).

`value`  : device-resident -- inputs already in HBM, CUDA events on the launching stream.
`e2e`    : the same batch through the host-pointer C-ABI call with pinned host buffers; H2D + kernels +
           D2H inside the timed region.
`roofline`: the dominant kernel (SlotEdVerifyFinish = double scalar multiplication + point_eq), its
           launches timed with CUDA events inside the timed region; algorithmic MAC32 per signature from
           SURVEY.md 8(d); peak = measured IMAD.WIDE.U32 rate (profiles/r01_imad_peak.json).
`cpu_baseline`: the unmodified reference (oracle/_ref, arch_x86_64) on all host cores over a bounded sample.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

N_PER_GPU = 1 << 20
MSG_LEN = 32
MAC32 = {"verify": 981120, "verify_finish": 800300, "x448": 870208, "comb": 142656, "encode": 89696, "decode": 90064}  # the REFERENCE's algorithm
# IMAD.WIDE actually issued per signature by the finish kernel (multiply = 193, square = 110):
#   under a per-key table: 40 doublings (4S + 3M, +1M for T on every fifth), 90 x 8M + 30 x 7M additions (8 without T), the
#   square-root-free R comparison (3S + 8M);  stand-alone: 445 doublings + own window table + the same comparison
EXECUTED_MAC32 = {"finish_shared": 163 * 110 + 1058 * 193, "finish_alone": 1787 * 110 + 2396 * 193}
METRIC = "Ed448 verifies/s at batch 2^20 per GPU (X448 and comb ops/s in extra)"
UNIT = "verifies/s"


def imad_peak():
    try:
        with open(os.path.join(ROOT, "profiles", "r01_imad_peak.json")) as f:
            return float(json.load(f)["imad_wide_u32_gmac_s"]), "measured (tools/imad_peak.cu on this pool's B200)"
    except Exception:
        return 148 * 32 * 1.965, "nominal 148 SMs x 32 IMAD.WIDE/clk x 1.965 GHz"


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the committed
    `ncu --set full` capture of this same workload (profiles/r01f_finish_ncu.txt); None if absent."""
    try:
        tot, mult = 0.0, {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        with open(os.path.join(ROOT, "profiles", "r01f_finish_ncu.txt")) as f:
            for line in f:
                t = line.split()
                if len(t) >= 3 and t[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    tot += float(t[2]) * mult[t[1]]
        return tot or None
    except Exception:
        return None


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


# ------------------------------------------------------------------------------------------------
# workload (synthetic, seeded): 2^16 keys x 16 messages of 32 bytes, 1/8 corrupted
# ------------------------------------------------------------------------------------------------
def make_corpus(signer, n, label, per=16, varlen=False, corrupt=True):
    """varlen: message lengths uniform in [0, 256) (second run) instead of MSG_LEN bytes each"""
    from util import stream_bytes
    nk = max(1, n // per)
    sk = stream_bytes(label + "/sk", nk * 57).reshape(nk, 57)
    pk = signer.ed448_derive_public_key(sk)
    sk_all = np.repeat(sk, per, axis=0)[:n]
    pk_all = np.repeat(pk, per, axis=0)[:n].copy()
    if varlen:
        lens = stream_bytes(label + "/len", n).astype(np.uint64)
        off = np.concatenate([np.zeros(1, np.uint64), np.cumsum(lens, dtype=np.uint64)])
        arena = stream_bytes(label + "/msg", int(off[-1]) + 1)
    else:
        arena = stream_bytes(label + "/msg", n * MSG_LEN)
        off = np.arange(n + 1, dtype=np.uint64) * MSG_LEN
    sig = signer.ed448_sign(sk_all, pk_all, (arena, off))
    kinds = np.zeros(n, np.int32)
    if corrupt:
        kinds[::8] = 1 + (np.arange((n + 7) // 8) % 4)
    sel = stream_bytes(label + "/sel", n)
    i = np.flatnonzero(kinds == 1); sig[i, sel[i] % 57] ^= 1
    i = np.flatnonzero(kinds == 2); sig[i, 57 + sel[i] % 56] ^= 2
    i = np.flatnonzero(kinds == 3); pk_all[i, sel[i] % 57] ^= 4
    i = np.flatnonzero(kinds == 4)
    if varlen:
        i = i[off[i + 1] > off[i]]                       # an empty message has no byte to flip: left valid
        kinds[np.setdiff1d(np.flatnonzero(kinds == 4), i)] = 0
        arena[(off[i] + sel[i] % (off[i + 1] - off[i])).astype(np.int64)] ^= 8
    else:
        arena[i * MSG_LEN + sel[i] % MSG_LEN] ^= 8
    expect = np.where(kinds == 0, -1, 0).astype(np.int32)
    return sig, pk_all, arena, off, expect


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc, self.window = index, [], None, None

    def start(self):
        """started before the warm-up steps (nvidia-smi needs ~0.1 s to deliver its first sample); mark() brackets the timed region"""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def mark(self, t0, t1):
        self.window = (t0, t1)

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        rows, where = [r for _, r in self.rows], "warm-up + timed region (same load)"
        if self.window:
            inside = [r for t, r in self.rows if self.window[0] <= t <= self.window[1] + 0.02]
            if inside:
                rows, where = inside, "timed region"
        sm = [float(r[1]) for r in rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": where}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_reference_rate(n_sample, steps, warmup, corpus=None):
    """The reference's own goldilocks_ed448_verify over `n_sample` signatures of the bench corpus, all host cores."""
    import util
    ref = util.ref_lib()
    kind = "reference"
    if ref is None:
        ref, kind = util.oracle_lib(), "port"
    cores = host_cores()
    util.set_threads(ref, cores)
    if corpus is None:
        corpus = make_corpus(ref, n_sample, "bench/cpu")
    sig, pk, arena, off, expect = corpus
    sig, pk, arena, off, expect = sig[:n_sample], pk[:n_sample], arena[: n_sample * MSG_LEN], off[: n_sample + 1], expect[:n_sample]
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        st = ref.ed448_verify(sig, pk, (arena, off))
        dt = time.perf_counter() - t0
        assert (st == expect).all(), "reference disagrees with the corpus' expected accept bits"
        if it >= warmup:
            times.append(dt)
    t = sum(times) / len(times)
    return {"value": n_sample / t, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "%d signatures of the bench corpus (32-byte messages, 1/8 corrupted), %d timed passes, %s threads via pthreads" % (n_sample, len(times), cores),
            "lib": os.path.basename(ref.path)}, t


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    n_sample = min(N_PER_GPU, 1024 * cores)
    base, t = cpu_reference_rate(n_sample, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 limbs (448-bit integers)",
            "data": "synthetic", "config": config_dict(args.gpus, n_sample),
            "cpu_baseline": base, "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def config_dict(gpus, n_per_step, per_key=16):
    return {"workload": "BASELINE configs[3]: Ed448 batch verify (SHAKE256 + double scalarmul), 2^16 keys x 16 messages of 32 B, 1/8 corrupted"
                        + ("" if per_key == 16 else " -- NON-DEFAULT corpus: %d signatures per key" % per_key),
            "keys": "SURVEY 8(d) C4: 2^16 distinct keys x 16 signatures each; byte-identical keys of a batch share one per-key table "
                    "(extra.verify_distinct_keys = the same batch size with 2^20 distinct keys, no sharing possible)",
            "signatures_per_gpu_per_step": n_per_step, "parallelism": "independent shards x%d, no collective" % gpus,
            "l2": "inputs+scratch per step (>700 MB) exceed the 126 MB L2; no flush needed"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import libgoldilocks_b200 as g
    from libgoldilocks_b200.engine import DeviceEngine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    eng = DeviceEngine()
    lib = eng.capi
    n = args.n
    K, W = args.steps, args.warmup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- inputs (host, pinned) -----------------------------------------------------------------------
    sig, pk, arena, off, expect = make_corpus(lib, n, "bench/rank%d" % rank, per=args.per_key)

    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t

    h_sig, h_pk, h_msg = pinned(sig.reshape(-1)), pinned(pk.reshape(-1)), pinned(arena)
    h_off = pinned(off.view(np.int64))
    h_st = torch.empty(n, dtype=torch.int32).pin_memory()
    d_sig, d_pk, d_msg, d_off = (t.to(dev) for t in (h_sig, h_pk, h_msg, h_off))
    d_st = torch.empty(n, dtype=torch.int32, device=dev)
    d_scratch = torch.empty(eng.verify_scratch_bytes(n), dtype=torch.uint8, device=dev)

    def step():
        eng.ed448_verify(d_st, d_sig, d_pk, d_msg, d_off, d_scratch)

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(W):
        step()
    torch.cuda.synchronize()
    assert (d_st.cpu().numpy() == expect).all(), "device verify disagrees with the corpus' expected accept bits"

    # ---- device-resident timed region -------------------------------------------------------------------
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.time()
    l0 = eng.launch_count()
    lib.lib.goldilocks_b200_profile(C.c_int(1))
    e0.record()
    for _ in range(K):
        step()
    e1.record()
    barrier()
    sampler.mark(t_wall0, time.time())
    lib.lib.goldilocks_b200_profile(C.c_int(0))
    launches = eng.launch_count() - l0
    clocks = sampler.stop()
    t_step = max_over_ranks(e0.elapsed_time(e1) / 1e3 / K)
    names = C.create_string_buffer(64 * 4096)
    ms = (C.c_float * 4096)()
    lib.lib.goldilocks_b200_profile_read.restype = C.c_size_t
    cnt = lib.lib.goldilocks_b200_profile_read(names, ms, C.c_size_t(4096))
    per_kernel = {}
    for k in range(cnt):
        nm = names.raw[64 * k:64 * k + 64].split(b"\0")[0].decode()
        per_kernel.setdefault(nm, []).append(ms[k])
    kavg = {k: float(np.sum(v)) / K for k, v in per_kernel.items()}   # ms per step and kernel (a kernel may launch more than once per step)
    dominant = "SlotEdVerifyFinishShared" if "SlotEdVerifyFinishShared" in kavg else "SlotEdVerifyFinish"
    t_finish = kavg.get(dominant, 0.0) / 1e3
    peak, peak_how = imad_peak()
    achieved = n * MAC32["verify_finish"] / t_finish / 1e9 if t_finish > 0 else 0.0
    # signatures whose key occurs at least twice in the batch go through the per-key tables (what the device-side grouping finds)
    _, inv, cnt = np.unique(np.ascontiguousarray(pk).view(np.dtype((np.void, 57))).ravel(), return_inverse=True, return_counts=True)
    n_shared = int((cnt[inv.reshape(-1)] >= 2).sum())
    executed = n_shared * EXECUTED_MAC32["finish_shared"] + (n - n_shared) * EXECUTED_MAC32["finish_alone"]
    roofline = {"bound": "imad", "kernel": "k_slots_persist<%s>" % dominant, "achieved": achieved, "peak": peak, "unit": "GMAC32/s",
                "frac": achieved / peak, "traffic": ncu_traffic(), "traffic_unit": "bytes/launch (ncu --set full, profiles/r01f_finish_ncu.txt)",
                "algorithmic_bytes_per_launch": n * (512 + 112 + 8 + 4), "peak_source": peak_how,
                "algorithmic_mac32_per_signature": MAC32["verify_finish"],
                "executed_mac32_per_launch": executed, "signatures_under_a_shared_key_table": n_shared,
                "frac_executed": executed / t_finish / 1e9 / peak if t_finish > 0 else None,
                "kernel_ms": kavg,
                "hbm": (lambda tr, pk: None if not tr or t_finish <= 0 else {"achieved": tr / t_finish / 1e9, "peak": pk[0], "unit": "GB/s", "frac": tr / t_finish / 1e9 / pk[0],
                                                                               "peak_source": pk[1], "note": "informational: DRAM traffic of the ncu capture / this run's launch time; the kernel is multiplier-bound"})(ncu_traffic(), hbm_peak()),
                "kernel_share_of_step": t_finish / (e0.elapsed_time(e1) / 1e3 / K) if t_finish > 0 else None,
                "step_frac": n * MAC32["verify"] / t_step / 1e9 / peak,
                "note": "integer-multiply-pipe roofline (north_star). `achieved`/`frac` count the REFERENCE's algorithmic work per signature "
                        "(SURVEY 8(d): 800 300 MAC32 for the double-scalar multiplication), so sharing a per-key table pushes them above 1; "
                        "`frac_executed` counts the IMAD.WIDE this kernel really issues and is the pipe utilisation. DRAM traffic = key tables "
                        "(41 KB per key, re-read once per row because the keys resident at a time overflow L2) + wide fixed-base tables (30 MB): ~5 % of the HBM roof"}

    # ---- end to end through the host-pointer C ABI ---------------------------------------------------------
    fn = lib.lib.goldilocks_ed448_verify_batch
    fn.restype = C.c_int32
    argv = [C.c_void_p(h_st.data_ptr()), C.c_void_p(h_sig.data_ptr()), C.c_void_p(h_pk.data_ptr()), C.c_void_p(h_msg.data_ptr()),
            C.c_void_p(h_off.data_ptr()), C.c_uint8(0), None, C.c_uint8(0), C.c_size_t(n)]
    for _ in range(max(1, W - 1)):
        assert fn(*argv) == -1
    assert (h_st.numpy() == expect).all()
    ke = max(2, min(K, 5))
    barrier()
    t0 = time.perf_counter()
    for _ in range(ke):
        assert fn(*argv) == -1
    torch.cuda.synchronize()
    t_e2e = max_over_ranks((time.perf_counter() - t0) / ke)
    barrier()
    h2d = sig.nbytes + pk.nbytes + arena.nbytes + off.nbytes
    e2e = {"value": world * n / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(4 * n), "ms_per_step": t_e2e * 1e3,
           "api": "goldilocks_ed448_verify_batch (host pointers, pinned)"}

    # ---- extra: X448 (config 3) and comb (config 2), device-resident, 2^20 each ------------------------------
    extra = {}
    if not args.no_extra:
        from util import stream_bytes
        u = torch.from_numpy(stream_bytes("bench/x448/u%d" % rank, n * 56)).to(dev)
        kk = torch.from_numpy(stream_bytes("bench/x448/k%d" % rank, n * 56)).to(dev)
        xo = torch.empty(n * 56, dtype=torch.uint8, device=dev)
        xs = torch.empty(n, dtype=torch.int32, device=dev)
        sc = torch.from_numpy(lib.scalar_decode_long(stream_bytes("bench/comb/s%d" % rank, n * 56).reshape(n, 56), 56).reshape(-1)).to(dev)
        co = torch.empty(n * 256, dtype=torch.uint8, device=dev)
        for name, fnc, mac in (("x448", lambda: eng.x448(xo, xs, u, kk), MAC32["x448"]), ("comb", lambda: eng.precomputed_scalarmul(co, sc), MAC32["comb"])):
            for _ in range(2):
                fnc()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            kx = max(2, min(K, 3))
            barrier()
            a.record()
            for _ in range(kx):
                fnc()
            b.record()
            barrier()
            t = max_over_ranks(a.elapsed_time(b) / 1e3 / kx)
            extra[name] = {"value": world * n / t, "unit": "ops/s", "ms_per_step": t * 1e3, "batch_per_gpu": n,
                           "imad_frac": n * mac / t / 1e9 / peak, "algorithmic_mac32_per_op": mac}

    # ---- extra: decaf encode / decode (BASELINE configs[4]: 2^24 elements over 8 GPUs; here 2^20 per GPU like the other lines), on the
    #      comb outputs above (valid points).  Multiplier-bound like everything else (one inverse square root each); HBM GB/s is the
    #      informational figure BASELINE.json asks for on these paths ----------------------------------------------------------------------
    if not args.no_extra:
        enc = torch.empty(n * 56, dtype=torch.uint8, device=dev)
        dpts = torch.empty(n * 256, dtype=torch.uint8, device=dev)
        dstat = torch.empty(n, dtype=torch.int32, device=dev)
        hbm, _ = hbm_peak()
        for name, fnc, mac, nbytes in (("decaf_encode", lambda: eng.point_encode(enc, co), MAC32["encode"], 256 + 56),
                                       ("decaf_decode", lambda: eng.point_decode(dpts, dstat, enc, True), MAC32["decode"], 56 + 256 + 4)):
            for _ in range(2):
                fnc()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            kx = max(2, min(K, 3))
            barrier()
            a.record()
            for _ in range(kx):
                fnc()
            b.record()
            barrier()
            t = max_over_ranks(a.elapsed_time(b) / 1e3 / kx)
            extra[name] = {"value": world * n / t, "unit": "ops/s", "ms_per_step": t * 1e3, "batch_per_gpu": n, "imad_frac": n * mac / t / 1e9 / peak,
                           "algorithmic_mac32_per_op": mac, "hbm_gbs": n * nbytes / t / 1e9, "hbm_frac": n * nbytes / t / 1e9 / hbm}
        assert int((dstat == -1).sum().item()) == n, "decode(encode(P)) must succeed for every comb output"
        del enc, dpts, dstat

    # ---- extra: the same batch size (a) with 2^20 DISTINCT keys (no table can be shared), (b) with message lengths
    #      uniform in [0, 256) under the 2^16 x 16 keys (second run) --------------------------------------
    if not args.no_extra:
        for name, kw in (("verify_distinct_keys", {"per": 1}), ("verify_varlen_msgs", {"varlen": True})):
            sig1, pk1, arena1, off1, expect1 = make_corpus(lib, n, "bench/%s/rank%d" % (name, rank), **kw)
            t_sig, t_pk, t_msg, t_off = (torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (sig1.reshape(-1), pk1.reshape(-1), arena1, off1.view(np.int64)))
            for _ in range(2):
                eng.ed448_verify(d_st, t_sig, t_pk, t_msg, t_off, d_scratch)
            torch.cuda.synchronize()
            assert (d_st.cpu().numpy() == expect1).all(), "device verify (%s) disagrees with the expected accept bits" % name
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            kx = max(2, min(K, 3))
            barrier()
            a.record()
            for _ in range(kx):
                eng.ed448_verify(d_st, t_sig, t_pk, t_msg, t_off, d_scratch)
            b.record()
            barrier()
            t = max_over_ranks(a.elapsed_time(b) / 1e3 / kx)
            extra[name] = {"value": world * n / t, "unit": UNIT, "ms_per_step": t * 1e3, "batch_per_gpu": n,
                           "imad_frac": n * MAC32["verify"] / t / 1e9 / peak, "algorithmic_mac32_per_op": MAC32["verify"]}
            del t_sig, t_pk, t_msg, t_off

    # ---- extra: the same corpus against a KEY SET (tables of the 2^16 keys built once, outside the timed region), host pointers ----
    if not args.no_extra and args.per_key == 16:
        nk = n // 16
        keys = np.ascontiguousarray(pk.reshape(n, 57)[::16]).copy()
        idx8 = np.arange(0, n, 8)
        flipped = idx8[(np.arange(len(idx8)) % 4) == 2]          # make_corpus kind 3: the key bytes of these entries were changed
        keys[flipped[flipped % 16 == 0] // 16] = np.ascontiguousarray(pk.reshape(n, 57))[flipped[flipped % 16 == 0] + 1]
        expect_ks = expect.copy(); expect_ks[flipped] = -1        # under the key SET they verify against the intact key
        handle = lib.keyset_create(keys)
        h_idx = pinned((np.arange(n, dtype=np.uint32) // 16).astype(np.uint32))
        fk = lib.lib.goldilocks_ed448_verify_keyset_batch
        fk.restype = C.c_int32
        argk = [C.c_void_p(h_st.data_ptr()), handle, C.c_void_p(h_idx.data_ptr()), C.c_void_p(h_sig.data_ptr()), C.c_void_p(h_msg.data_ptr()),
                C.c_void_p(h_off.data_ptr()), C.c_uint8(0), None, C.c_uint8(0), C.c_size_t(n)]
        for _ in range(2):
            assert fk(*argk) == -1
        assert (h_st.numpy() == expect_ks).all(), "key-set verify disagrees with the expected accept bits"
        kx = max(2, min(K, 3))
        barrier()
        t0 = time.perf_counter()
        for _ in range(kx):
            assert fk(*argk) == -1
        torch.cuda.synchronize()
        t = max_over_ranks((time.perf_counter() - t0) / kx)
        barrier()
        lib.keyset_destroy(handle)
        extra["verify_keyset_e2e"] = {"value": world * n / t, "unit": UNIT, "ms_per_step": t * 1e3, "batch_per_gpu": n, "keys_in_set": int(nk),
                                      "api": "goldilocks_ed448_verify_keyset_batch (host pointers, pinned; tables of the key set built once, not timed)"}

    # ---- extra: random-linear-combination batch verification on an ALL-VALID corpus of the same shape, host
    #      pointers; and the price of the fallback when one signature of the batch is bad -----------------------------------------
    if not args.no_extra and lib.has("goldilocks_ed448_verify_rlc_batch"):
        sig2, pk2, arena2, off2, expect2 = make_corpus(lib, n, "bench/rlc/rank%d" % rank, per=args.per_key, corrupt=False)
        r_sig, r_pk, r_msg, r_off = pinned(sig2.reshape(-1)), pinned(pk2.reshape(-1)), pinned(arena2), pinned(off2.view(np.int64))
        fr = lib.lib.goldilocks_ed448_verify_rlc_batch
        fr.restype = C.c_int32
        fast = C.c_int(0)
        argr = [C.c_void_p(h_st.data_ptr()), C.c_void_p(r_sig.data_ptr()), C.c_void_p(r_pk.data_ptr()), C.c_void_p(r_msg.data_ptr()),
                C.c_void_p(r_off.data_ptr()), C.c_uint8(0), None, C.c_uint8(0), C.c_size_t(n), C.byref(fast)]
        for _ in range(2):
            assert fr(*argr) == -1
        assert (h_st.numpy() == -1).all() and fast.value == 1, "the batch equation must decide an all-valid batch"
        kx = max(2, min(K, 3))
        barrier()
        lib.lib.goldilocks_b200_profile(C.c_int(1))
        t0 = time.perf_counter()
        for _ in range(kx):
            assert fr(*argr) == -1
        torch.cuda.synchronize()
        t = max_over_ranks((time.perf_counter() - t0) / kx)
        lib.lib.goldilocks_b200_profile(C.c_int(0))
        cnt2 = lib.lib.goldilocks_b200_profile_read(names, ms, C.c_size_t(4096))
        krlc = {}
        for k in range(cnt2):
            nm = names.raw[64 * k:64 * k + 64].split(b"\0")[0].decode()
            krlc[nm] = krlc.get(nm, 0.0) + ms[k] / kx
        barrier()
        bad_i = 12345 % n
        r_sig[114 * bad_i + 70] ^= 1                                  # one wrong S: the equation fails, the ordinary path decides
        assert fr(*argr) == -1                                        # the first fallback also grows the arena by the ordinary path's scratch
        t0 = time.perf_counter()
        assert fr(*argr) == -1
        t_bad = time.perf_counter() - t0
        st_bad = h_st.numpy()
        assert fast.value == 0 and st_bad[bad_i] == 0 and (st_bad == -1).sum() == n - 1
        extra["verify_rlc_all_valid_e2e"] = {"value": world * n / t, "unit": UNIT, "ms_per_step": t * 1e3, "batch_per_gpu": n, "fast_path": 1,
                                             "kernel_ms": krlc, "ms_when_one_signature_is_bad": t_bad * 1e3,
                                             "api": "goldilocks_ed448_verify_rlc_batch (host pointers, pinned): one multi-scalar multiplication with secret 128-bit "
                                                    "weights decides the batch; per-signature fallback when it fails; corpus = the bench shape with NO corrupted entries"}
        del r_sig, r_pk, r_msg, r_off

    # ---- CPU baseline (rank 0, N = 1 only) ---------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = host_cores()
        n_cpu = min(n, 4096 * cores)       # ~0.6 s per pass on all cores, 1 warm-up + 2 timed = ~25 core-seconds
        cpu, _ = cpu_reference_rate(n_cpu, 2, 1, corpus=(sig, pk, arena, off, expect))
    if rank == 0:
        line = {"metric": METRIC, "value": world * n / t_step, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": t_step * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 limbs (448-bit integers)", "data": "synthetic",
                "config": config_dict(world, n, args.per_key), "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
                "cpu_baseline": cpu, "extra": extra}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def emit(line):
    """the ONE JSON line on the real stdout (library banners, e.g. NCCL's version line, are kept off it)"""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)          # anything a library prints to fd 1 from here on goes to stderr
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=N_PER_GPU, help="signatures per GPU per step (default 2^20, the BASELINE size)")
    ap.add_argument("--per-key", type=int, default=16, help="signatures per public key in the corpus (SURVEY 8(d) C4: 16); other values are for the "
                    "sensitivity table in DESIGN.md, not the headline")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
