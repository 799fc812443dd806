// tests/hostsim/hostsim.cpp -- TEST INFRASTRUCTURE ONLY (never shipped, never loaded by the product).
//
// Compiles the device math headers of libgoldilocks_b200/csrc (gf/sc/point/algos/lanes .cuh are
// __host__ __device__ clean) with the host compiler and runs the very same per-lane functors the
// CUDA kernels run, in a plain loop.  This lets the CPU-only test tier (`-m "not gpu"`) check the
// CUDA source's arithmetic, limb-bound discipline (-DGF_CHECK_BOUNDS aborts on any violation) and
// host-layout I/O against the oracle and the golden vectors in a container without a GPU.
// It exports the `*_batch` names of include/goldilocks_b200.h so the tests can drive it with the
// same ctypes bindings as the real library.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "../../libgoldilocks_b200/csrc/slot_lanes.cuh"

#define EXPORT extern "C" __attribute__((visibility("default")))

// multiplier-call counters (gf.cuh GF_COUNT): [0] multiplications, [1] squarings, [2] word multiplications since the last reset
EXPORT void hostsim_op_counts(unsigned long long out[3], int reset) {
    for (int k = 0; k < 3; k++) { out[k] = gf_op_counter(k).load(); if (reset) gf_op_counter(k).store(0); }
}
// the same counters per functor (= per kernel of the product): every run*() below adds what its functor executed
#include <mutex>
#include <typeinfo>
static std::mutex g_stage_mu;
static std::map<std::string, std::vector<unsigned long long>> g_stage;   /* name -> {mul, sqr, mulw, elements} */
struct StageScope {
    const char *name; size_t n; unsigned long long before[3];
    StageScope(const char *nm, size_t n_) : name(nm), n(n_) { for (int k = 0; k < 3; k++) before[k] = gf_op_counter(k).load(); }
    ~StageScope() {
        std::lock_guard<std::mutex> g(g_stage_mu);
        const char *nm = name;
        while (*nm >= '0' && *nm <= '9') nm++;
        auto &v = g_stage[nm];
        v.resize(4);
        for (int k = 0; k < 3; k++) v[k] += gf_op_counter(k).load() - before[k];
        v[3] += n;
    }
};
EXPORT size_t hostsim_stage_counts(char *names, unsigned long long *counts, size_t max, int reset) {
    std::lock_guard<std::mutex> g(g_stage_mu);
    size_t k = 0;
    for (auto &kv : g_stage) {
        if (k >= max) break;
        snprintf(names + 64 * k, 64, "%s", kv.first.c_str());
        for (int j = 0; j < 4; j++) counts[4 * k + j] = kv.second[j];
        k++;
    }
    if (reset) g_stage.clear();
    return k;
}
static int g_threads = 1;
EXPORT void hostsim_set_threads(int t) { g_threads = t < 1 ? 1 : t; }

template <class F>
static void run(const F &f, size_t n) {
    StageScope stage_(typeid(F).name(), n);
    int nt = g_threads;
    if ((size_t)nt > n) nt = n ? (int)n : 1;
    if (nt <= 1) { for (size_t i = 0; i < n; i++) f(i); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++)
        th.emplace_back([&f, n, nt, t]() { for (size_t i = n * t / nt; i < n * (t + 1) / nt; i++) f(i); });
    for (auto &x : th) x.join();
}
template <class F>
static void run_sm(const F &f, size_t n) {
    StageScope stage_(typeid(F).name(), n); /* slot machine: F::NSLOTS field elements of scratch per worker */
    int nt = g_threads;
    if ((size_t)nt > n) nt = n ? (int)n : 1;
    auto work = [&f](size_t lo, size_t hi) {
        std::vector<gf> slots(F::NSLOTS);
        sref base = {slots.data()};
        for (size_t i = lo; i < hi; i++) f(i, base, true);
    };
    if (nt <= 1) { work(0, n); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) th.emplace_back([&work, n, nt, t]() { work(n * t / nt, n * (t + 1) / nt); });
    for (auto &x : th) x.join();
}

template <class F>
static void run_smp(const F &f, size_t n) {
    StageScope stage_(typeid(F).name(), n); /* persistent slot machine: slots + one HBM scratch area per worker */
    int nt = g_threads;
    if ((size_t)nt > n) nt = n ? (int)n : 1;
    auto work = [&f](size_t lo, size_t hi, size_t slot) {
        std::vector<gf> slots(F::NSLOTS);
        sref base = {slots.data()};
        for (size_t i = lo; i < hi; i++) f(i, base, slot);
    };
    if (nt <= 1) { work(0, n, 0); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) th.emplace_back([&work, n, nt, t]() { work(n * t / nt, n * (t + 1) / nt, (size_t)t); });
    for (auto &x : th) x.join();
}

static fixed_tables *tables() {
    static fixed_tables *ft = nullptr;
    if (!ft) {
        ft = (fixed_tables *)aligned_alloc(64, sizeof(fixed_tables));
        for (int lane = 0; lane < TABLE_LANES; lane++) build_tables_lane(ft, lane);
    }
    return ft;
}
static niels *wide_table() {
    static niels *w = nullptr;
    if (!w) {
        w = (niels *)aligned_alloc(64, sizeof(niels) * WIDE_ENTRIES * WIDE_TABLES);
        std::vector<pniels> tmp(WIDE_ENTRIES * WIDE_TABLES);
        std::vector<gf> pre(WIDE_ENTRIES * WIDE_TABLES);
        LaneBuildWide fw = {w, tmp.data(), pre.data(), tables()};   /* 3.3 M entries: every core, whatever hostsim_set_threads says */
        const int keep = g_threads;
        g_threads = (int)std::max(1u, std::thread::hardware_concurrency());
        run(fw, (size_t)WIDE_LANES * WIDE_TABLES);
        g_threads = keep;
    }
    return w;
}
static void export_niels(uint8_t *out, const niels *t, int n) {
    uint64_t *o = (uint64_t *)out;
    for (int e = 0; e < n; e++) {
        const gf *g[3] = {&t[e].a, &t[e].b, &t[e].c};
        for (int j = 0; j < 3; j++)
            for (int l = 0; l < 8; l++) o[(e * 3 + j) * 8 + l] = (uint64_t)g[j]->v[2 * l] | ((uint64_t)g[j]->v[2 * l + 1] << 28);
    }
}
EXPORT int32_t goldilocks_b200_export_comb_table(uint8_t *out) { export_niels(out, tables()->comb, COMB_ENTRIES); return -1; }
EXPORT int32_t goldilocks_b200_export_wnaf_table(uint8_t *out) { export_niels(out, wide_table(), WNAF_FIXED_ENTRIES); return -1; }
EXPORT int32_t goldilocks_b200_init(void) { tables(); return -1; }

typedef abi_pt hpt;
typedef abi_sc hsc;
static std::vector<uint4> g_slots;
static uint4 *slots(size_t tables_per_thread) { /* window tables (wtab): workers rounded up to whole warps of 32 */
    g_slots.resize((size_t)((g_threads + 31) / 32 * 32) * tables_per_thread * WTAB_QUADS_PER_LANE);
    return g_slots.data();
}

#define GF2(NAME, OP) EXPORT int32_t NAME(uint8_t *o, const uint8_t *a, const uint8_t *b, size_t n) { LaneGf<OP> f = {o, nullptr, a, b, 0}; run(f, n); return -1; }
GF2(goldilocks_448_gf_mul_batch, GFOP_MUL)
GF2(goldilocks_448_gf_add_batch, GFOP_ADD)
GF2(goldilocks_448_gf_sub_batch, GFOP_SUB)
EXPORT int32_t goldilocks_448_gf_sqr_batch(uint8_t *o, const uint8_t *a, size_t n) { LaneGf<GFOP_SQR> f = {o, nullptr, a, nullptr, 0}; run(f, n); return -1; }
EXPORT int32_t goldilocks_448_gf_mulw_batch(uint8_t *o, const uint8_t *a, uint32_t w, size_t n) { LaneGf<GFOP_MULW> f = {o, nullptr, a, nullptr, w}; run(f, n); return -1; }
EXPORT int32_t goldilocks_448_gf_isr_batch(uint8_t *o, int32_t *st, const uint8_t *a, size_t n) { LaneGf<GFOP_ISR> f = {o, st, a, nullptr, 0}; run(f, n); return -1; }
EXPORT int32_t goldilocks_448_gf_invert_batch(uint8_t *o, const uint8_t *a, size_t n) { LaneGf<GFOP_INVERT> f = {o, nullptr, a, nullptr, 0}; run(f, n); return -1; }

EXPORT int32_t goldilocks_448_point_add_batch(hpt *o, const hpt *a, const hpt *b, size_t n) { LanePt<PTOP_ADD> f = {o, a, b}; run(f, n); return -1; }
EXPORT int32_t goldilocks_448_point_sub_batch(hpt *o, const hpt *a, const hpt *b, size_t n) { LanePt<PTOP_SUB> f = {o, a, b}; run(f, n); return -1; }
EXPORT int32_t goldilocks_448_point_double_batch(hpt *o, const hpt *a, size_t n) { LanePt<PTOP_DBL> f = {o, a, nullptr}; run(f, n); return -1; }
EXPORT int32_t goldilocks_448_point_negate_batch(hpt *o, const hpt *a, size_t n) { LanePt<PTOP_NEG> f = {o, a, nullptr}; run(f, n); return -1; }
EXPORT int32_t goldilocks_b200_debug_niels_batch(hpt *out, const hpt *p, const hpt *q, const uint32_t *which, uint32_t op, size_t n) {
    if (op > 9) return 0;
    std::vector<pt> recs(n + 1);
    if (op < 6) { LanePtNiels f = {out, recs.data(), p, q, tables()->comb, which, op}; run(f, n); return -1; }
    LanePtNiels f0 = {out, recs.data(), p, q, tables()->comb, which, 10};
    run(f0, n);
    SlotNielsDebug f = {out, p, recs.data(), tables()->comb, which, op};
    run_sm(f, n);
    return -1;
}
// sc_half_gcd (sc.cuh) on n scalars: u, |v| as 56-byte little-endian numbers, v_neg[i] = 1 when v is negative
EXPORT int32_t hostsim_half_gcd(hsc *u, hsc *v, uint32_t *v_neg, const hsc *c, size_t n) {
    for (size_t i = 0; i < n; i++) {
        sc cc, uu, vv;
        sc_from_abi(cc, c + i);
        v_neg[i] = sc_half_gcd(uu, vv, cc) ? 1u : 0u;
        sc_to_abi(u + i, uu);
        sc_to_abi(v + i, vv);
    }
    return -1;
}
EXPORT int32_t goldilocks_448_point_eq_batch(uint64_t *o, const hpt *a, const hpt *b, size_t n) { LanePtEq f = {o, a, b}; run(f, n); return -1; }
EXPORT int32_t goldilocks_448_point_valid_batch(uint64_t *o, const hpt *a, size_t n) { LanePtValid f = {o, a}; run(f, n); return -1; }
EXPORT int32_t goldilocks_448_point_encode_batch(uint8_t *o, const hpt *a, size_t n) { LanePtEncode f = {o, a}; run(f, n); return -1; }
EXPORT int32_t goldilocks_448_point_decode_batch(hpt *o, int32_t *st, const uint8_t *ser, uint64_t allow_identity, size_t n) { LanePtDecode f = {o, st, ser, allow_identity ? 1u : 0u}; run(f, n); return -1; }
EXPORT int32_t goldilocks_448_point_from_hash_nonuniform_batch(hpt *o, const uint8_t *h, size_t n) { LaneFromHash<false> f = {o, h}; run(f, n); return -1; }
EXPORT int32_t goldilocks_448_point_from_hash_uniform_batch(hpt *o, const uint8_t *h, size_t n) { LaneFromHash<true> f = {o, h}; run(f, n); return -1; }
EXPORT int32_t goldilocks_448_invert_elligator_nonuniform_batch(uint8_t *h, int32_t *st, const hpt *a, const uint32_t *which, size_t n) { LaneInvertElligator<false> f = {h, st, a, which}; run(f, n); return -1; }
EXPORT int32_t goldilocks_448_invert_elligator_uniform_batch(uint8_t *h, int32_t *st, const hpt *a, const uint32_t *which, size_t n) { LaneInvertElligator<true> f = {h, st, a, which}; run(f, n); return -1; }
EXPORT int32_t goldilocks_448_point_mul_by_ratio_and_encode_like_eddsa_batch(uint8_t *o, const hpt *a, size_t n) { LaneEncodeEddsa f = {o, a}; run(f, n); return -1; }
EXPORT int32_t goldilocks_448_point_decode_like_eddsa_and_mul_by_ratio_batch(hpt *o, int32_t *st, const uint8_t *enc, size_t n) { LaneDecodeEddsa f = {o, st, enc}; run(f, n); return -1; }
EXPORT int32_t goldilocks_448_point_mul_by_ratio_and_encode_like_x448_batch(uint8_t *o, const hpt *a, size_t n) { LaneEncodeX448 f = {o, a}; run(f, n); return -1; }

EXPORT int32_t goldilocks_448_precomputed_scalarmul_batch(hpt *o, const void *table, const hsc *s, size_t n) {
    if (!table) { SlotComb f = {o, s, tables()}; run_sm(f, n); return -1; }
    std::vector<niels> tab(COMB_ENTRIES);
    LaneNielsFromAbi cv = {tab.data(), (const abi_niels *)table};
    run(cv, COMB_ENTRIES);
    SlotCombTable f = {o, s, tab.data()};
    run_sm(f, n);
    return -1;
}
EXPORT int32_t goldilocks_448_point_scalarmul_batch(hpt *o, const hpt *b, const hsc *s, size_t n) { SlotScalarmul f = {o, b, s, slots(1)}; run_smp(f, n); return -1; }
EXPORT int32_t goldilocks_448_point_double_scalarmul_batch(hpt *o, const hpt *b1, const hsc *s1, const hpt *b2, const hsc *s2, size_t n) { SlotDoubleScalarmul f = {o, b1, s1, b2, s2, slots(2), (size_t)((g_threads + 31) / 32 * 32)}; run_smp(f, n); return -1; }
EXPORT int32_t goldilocks_448_point_dual_scalarmul_batch(hpt *o1, hpt *o2, const hpt *b, const hsc *s1, const hsc *s2, size_t n) { SlotDualScalarmul f = {o1, o2, b, s1, s2, slots(1)}; run_smp(f, n); return -1; }
EXPORT int32_t goldilocks_448_direct_scalarmul_batch(uint8_t *o, int32_t *st, const uint8_t *b, const hsc *s, uint64_t allow_identity, uint64_t short_circuit, size_t n) {
    SlotDirectScalarmul f = {o, st, b, s, allow_identity ? 1u : 0u, short_circuit ? 1u : 0u, tables(), slots(1)};
    run_smp(f, n);
    return -1;
}
EXPORT int32_t goldilocks_448_precompute_batch(void *tables_out, const hpt *pts, size_t n) {
    std::vector<niels> scratch(16 * COMB_N * n + 1);
    LanePrecompute f = {(abi_niels *)tables_out, pts, scratch.data()};
    run(f, COMB_N * n);
    return -1;
}
EXPORT int32_t goldilocks_448_point_debugging_torque_batch(hpt *o, const hpt *a, size_t n) { LanePt<PTOP_TORQUE> f = {o, a, nullptr}; run(f, n); return -1; }
EXPORT int32_t goldilocks_448_point_debugging_pscale_batch(hpt *o, const hpt *a, const uint8_t *fac, size_t n) { LanePtPscale f = {o, a, fac}; run(f, n); return -1; }
EXPORT int32_t goldilocks_448_scalar_invert_batch(hsc *o, int32_t *st, const hsc *a, size_t n) { LaneScInvert f = {o, st, a}; run(f, n); return -1; }
EXPORT int32_t goldilocks_ed448_convert_public_key_to_x448_batch(uint8_t *x, const uint8_t *ed, size_t n) { LaneEdPkToX448 f = {x, ed}; run(f, n); return -1; }
EXPORT int32_t goldilocks_ed448_convert_private_key_to_x448_batch(uint8_t *x, const uint8_t *ed, size_t n) { LaneEdSkToX448 f = {x, ed}; run(f, n); return -1; }
EXPORT int32_t goldilocks_448_base_double_scalarmul_non_secret_batch(hpt *o, const hsc *s1, const hpt *b2, const hsc *s2, size_t n) { SlotBaseDoubleScalarmul f = {o, s1, b2, s2, wide_table(), slots(1)}; run_smp(f, n); return -1; }

EXPORT int32_t goldilocks_448_scalar_add_batch(hsc *o, const hsc *a, const hsc *b, size_t n) { LaneSc<SCOP_ADD> f = {o, a, b}; run(f, n); return -1; }
EXPORT int32_t goldilocks_448_scalar_sub_batch(hsc *o, const hsc *a, const hsc *b, size_t n) { LaneSc<SCOP_SUB> f = {o, a, b}; run(f, n); return -1; }
EXPORT int32_t goldilocks_448_scalar_mul_batch(hsc *o, const hsc *a, const hsc *b, size_t n) { LaneSc<SCOP_MUL> f = {o, a, b}; run(f, n); return -1; }
EXPORT int32_t goldilocks_448_scalar_halve_batch(hsc *o, const hsc *a, size_t n) { LaneSc<SCOP_HALVE> f = {o, a, nullptr}; run(f, n); return -1; }
EXPORT int32_t goldilocks_448_scalar_decode_long_batch(hsc *o, const uint8_t *ser, size_t len, size_t n) { LaneScDecodeLong f = {o, ser, len}; run(f, n); return -1; }

EXPORT int32_t goldilocks_x448_batch(uint8_t *o, int32_t *st, const uint8_t *base, const uint8_t *sc, size_t n) { SlotX448 f = {o, st, base, sc}; run_sm(f, n); return -1; }
EXPORT int32_t goldilocks_x448_derive_public_key_batch(uint8_t *o, const uint8_t *sc, size_t n) { SlotX448DerivePk f = {o, sc, tables()}; run_sm(f, n); return -1; }
EXPORT int32_t goldilocks_shake256_hash_batch(uint8_t *o, size_t outlen, const uint8_t *in, const size_t *off, size_t n) { LaneShake256 f = {o, outlen, in, off}; run(f, n); return -1; }
EXPORT int32_t goldilocks_ed448_derive_public_key_batch(uint8_t *pk, const uint8_t *sk, size_t n) { SlotEdDerivePk f = {pk, sk, tables()}; run_sm(f, n); return -1; }
EXPORT int32_t goldilocks_ed448_sign_batch(uint8_t *sig, const uint8_t *sk, const uint8_t *pk, const uint8_t *msg, const size_t *off,
                                           uint8_t prehashed, const uint8_t *ctx, uint8_t ctx_len, size_t n) {
    std::vector<abi_sc> secret(n), nonce(n), nonce4(n);
    std::vector<uint8_t> seed(57 * n + 1);
    LaneEdSignExpand f0 = {secret.data(), seed.data(), sk};
    run(f0, n);
    LaneEdSignNonce f1 = {nonce.data(), nonce4.data(), seed.data(), msg, off, prehashed, ctx, ctx_len};
    run(f1, n);
    SlotEdSignR f2 = {sig, nonce4.data(), tables()};
    run_sm(f2, n);
    LaneEdSignFinish f3 = {sig, secret.data(), nonce.data(), pk, msg, off, prehashed, ctx, ctx_len};
    run(f3, n);
    return -1;
}
// key sets (include/goldilocks_b200.h): same functors, host memory
struct hostsim_keyset { size_t m; std::vector<uint8_t> pk; std::vector<int32_t> key_ok; std::vector<uint4> ktabs; bool flat; };
static unsigned long long g_keyset_flat_bytes = 32ull << 30;
EXPORT void goldilocks_b200_keyset_policy(unsigned long long max_table_bytes) { g_keyset_flat_bytes = max_table_bytes; }
EXPORT int32_t goldilocks_b200_keyset_create(hostsim_keyset **out, const uint8_t *pubkeys, size_t m) {
    const bool flat = (unsigned long long)(m ? m : 1) * KSET_QUADS * sizeof(uint4) <= g_keyset_flat_bytes;
    hostsim_keyset *ks = new hostsim_keyset{m, std::vector<uint8_t>(pubkeys, pubkeys + 57 * m), std::vector<int32_t>(m + 1),
                                            std::vector<uint4>((m + 1) * (size_t)(flat ? KSET_QUADS : KTAB_QUADS)), flat};
    std::vector<abi_pt> pts(m + 1);
    LaneDecodeEddsa fd = {pts.data(), ks->key_ok.data(), ks->pk.data()};
    run(fd, m);
    if (flat) {
        SlotKeysetChain fc = {pts.data(), ks->ktabs.data()};
        run_smp(fc, m);
        SlotKeysetColumns fcol = {ks->ktabs.data(), slots(1)};
        run_smp(fcol, m * KSET_COLS);
    } else {
        SlotKeysetTables ft = {pts.data(), ks->ktabs.data()};
        run_smp(ft, m);
    }
    LaneKeysetNormalize fn = {ks->ktabs.data(), flat ? (uint32_t)KSET_COLS : (uint32_t)VSH_CHUNKS, flat ? (uint32_t)KSET_QUADS : (uint32_t)KTAB_QUADS};
    run(fn, m * (flat ? KSET_COLS : VSH_CHUNKS));
    *out = ks;
    return -1;
}
EXPORT void goldilocks_b200_keyset_destroy(hostsim_keyset *ks) { delete ks; }
EXPORT size_t goldilocks_b200_keyset_size(const hostsim_keyset *ks) { return ks ? ks->m : 0; }
EXPORT int32_t goldilocks_ed448_verify_keyset_batch(int32_t *st, const hostsim_keyset *ks, const uint32_t *key_index, const uint8_t *sig, const uint8_t *msg,
                                                    const size_t *off, uint8_t prehashed, const uint8_t *ctx, uint8_t ctx_len, size_t n) {
    std::vector<abi_sc> chal(n + 1), resp(n + 1);
    LaneEdVerifyScalars f2 = {chal.data(), resp.data(), sig, ks->pk.data(), msg, off, prehashed, ctx, ctx_len, 0, key_index, (uint32_t)ks->m};
    run(f2, n);
    std::vector<verify_aux> aux(n + 1);
    if (ks->flat) {
        SlotEdVerifyFinishKeysetFlat f3 = {aux.data(), ks->key_ok.data(), chal.data(), resp.data(), wide_table(), ks->ktabs.data(), key_index, (uint32_t)ks->m, sig};
        run_smp(f3, n);
    } else {
        SlotEdVerifyFinishKeyset f3 = {aux.data(), ks->key_ok.data(), chal.data(), resp.data(), wide_table(), ks->ktabs.data(), key_index, (uint32_t)ks->m, sig};
        run_smp(f3, n);
    }
    LaneVerifySign fv = {st, aux.data(), 1, n};
    run(fv, (n + VSIGN_BATCH - 1) / VSIGN_BATCH);
    return -1;
}
EXPORT int32_t goldilocks_ed448_verify_batch(int32_t *st, const uint8_t *sig, const uint8_t *pk, const uint8_t *msg, const size_t *off,
                                             uint8_t prehashed, const uint8_t *ctx, uint8_t ctx_len, size_t n) {
    std::vector<abi_pt> pts(2 * n);
    std::vector<int32_t> ok(2 * n);
    std::vector<abi_sc> chal(n), resp(n);
    LaneEdVerifyScalars f2 = {chal.data(), resp.data(), sig, pk, msg, off, prehashed, ctx, ctx_len, 0};
    run(f2, n);
    if (n < 64) { /* abi.cu VERIFY_GROUP_MIN: no grouping pass, every signature stand-alone with half-size multipliers */
        const verify_plan none = {nullptr, nullptr, nullptr, nullptr, nullptr};
        LaneEdVerifyDecode f1 = {pts.data(), ok.data(), sig, pk, n, none, 0};
        run(f1, 2 * n);
        LaneVerifyHalf fh = {chal.data(), resp.data(), none};
        run(fh, n);
        SlotEdVerifyFinish f3 = {st, pts.data(), ok.data(), chal.data(), resp.data(), wide_table(), slots(2)};
        run_smp(f3, n);
        return -1;
    }
    /* the key-grouping pass of k_group.cu restated on the host: byte-identical keys that occur at least twice share a table */
    std::map<std::string, std::vector<uint32_t>> groups;
    for (size_t i = 0; i < n; i++) groups[std::string((const char *)pk + 57 * i, 57)].push_back((uint32_t)i);
    std::vector<uint32_t> shared_sig, shared_tab, unique_sig, tab_rep, counts(8);
    const size_t cap = n / 4 + 1;
    for (auto &g : groups) {
        if (g.second.size() >= 2 && tab_rep.size() < cap) {
            for (uint32_t i : g.second) { shared_sig.push_back(i); shared_tab.push_back((uint32_t)tab_rep.size()); }
            tab_rep.push_back(g.second[0]);
        } else {
            for (uint32_t i : g.second) unique_sig.push_back(i);
        }
    }
    counts[0] = (uint32_t)shared_sig.size(); counts[1] = (uint32_t)unique_sig.size(); counts[2] = (uint32_t)tab_rep.size();
    shared_sig.push_back(0); shared_tab.push_back(0); unique_sig.push_back(0); tab_rep.push_back(0); /* never empty */
    verify_plan plan = {shared_sig.data(), shared_tab.data(), unique_sig.data(), tab_rep.data(), counts.data()};
    std::vector<uint4> ktabs((size_t)(counts[2] + 1) * std::max<size_t>(KTAB_QUADS, vsh_pick(counts.data()).quads));   /* whichever column shape vsh_pick takes */
    LaneEdVerifyDecode f1 = {pts.data(), ok.data(), sig, pk, n, plan, n};   /* keys only: R is never decoded on this path */
    run(f1, n);
    LaneVerifyHalf fh = {chal.data(), resp.data(), plan};   /* stand-alone signatures: half-size multipliers, then their R */
    run(fh, n);
    LaneEdVerifyDecode fr = {pts.data(), ok.data(), sig, pk, n, plan, 2 * n};
    run(fr, n);
    SlotKeyChain fc = {pts.data(), ktabs.data(), plan};
    run_smp(fc, counts[2]);
    SlotKeyColumns ft = {ktabs.data(), slots(2), plan};
    run_smp(ft, (size_t)counts[2] * vsh_pick(counts.data()).chunks);
    SlotEdVerifyFinishShared fs = {pts.data(), ok.data(), chal.data(), resp.data(), wide_table(), ktabs.data(), slots(2), plan, sig};
    run_smp(fs, (size_t)counts[0] + counts[1]);
    LaneVerifySign fv = {st, (verify_aux *)(pts.data() + 1), 2, n};
    run(fv, (n + VSIGN_BATCH - 1) / VSIGN_BATCH);
    return -1;
}

// ---- random-linear-combination batch verification (csrc/rlc.cuh): the same lane functors, host orchestration ----------
// `force_c` > 0 overrides the digit width so small batches exercise several bucket shapes; `seed32` makes the weights
// reproducible in tests (the product draws them from getrandom()).
#include <algorithm>
#include "../../libgoldilocks_b200/csrc/rlc.cuh"
static int g_rlc_force_c = 0;
static size_t g_rlc_csize = 0; /* 0: one chunk */
static uint8_t g_rlc_seed[32] = {1, 2, 3};
EXPORT void hostsim_rlc_config(int force_c, const uint8_t *seed32) { g_rlc_force_c = force_c; if (seed32) memcpy(g_rlc_seed, seed32, 32); }
EXPORT void hostsim_rlc_chunk(size_t csize) { g_rlc_csize = csize; }
EXPORT int32_t goldilocks_ed448_verify_rlc_batch(int32_t *st, const uint8_t *sig, const uint8_t *pk, const uint8_t *msg, const size_t *off,
                                                 uint8_t prehashed, const uint8_t *ctx, uint8_t ctx_len, size_t n, int *fast_path) {
    if (fast_path) *fast_path = 0;
    if (n == 0) return -1;
    const size_t csize = g_rlc_csize && g_rlc_csize < n ? g_rlc_csize : n;
    const uint32_t nch = (uint32_t)((n + csize - 1) / csize);
    /* key groups per chunk, as k_group.cu leaves them when the chunk id is part of the key */
    std::map<std::pair<uint32_t, std::string>, std::vector<uint32_t>> groups;
    for (size_t i = 0; i < n; i++) groups[{(uint32_t)(i / csize), std::string((const char *)pk + 57 * i, 57)}].push_back((uint32_t)i);
    std::vector<uint32_t> order, gid, gstart, kchunk;
    for (auto &g : groups) {
        gstart.push_back((uint32_t)order.size());
        kchunk.push_back(g.first.first);
        for (uint32_t i : g.second) { order.push_back(i); gid.push_back((uint32_t)gstart.size()); }
    }
    const uint32_t m = (uint32_t)gstart.size();
    gstart.push_back((uint32_t)n);
    for (uint32_t ch = 0; ch < nch; ch++) kchunk.push_back(ch); /* the B of every chunk */
    const rlc_groups g = {order.data(), gid.data(), gstart.data(), m};
    rlc_shape sh_r = rlc_shape_for(csize, g_rlc_force_c, 0), sh_k = rlc_shape_for(((size_t)m + nch + nch - 1) / nch, g_rlc_force_c ? g_rlc_force_c + 1 : 0, 1);
    sh_r.nch = sh_k.nch = nch; sh_r.csize = sh_k.csize = (uint32_t)csize;
    const uint32_t cells = rlc_scells(sh_r);
    const size_t npts = n + m + nch;
    std::vector<pt> pts(npts);
    std::vector<int32_t> ok(npts), valid(n);
    std::vector<uint32_t> flags(1 + nch, 0), z(RLC_ZWORDS * (n + 8)), kscal(SC_WORDS * ((size_t)m + nch));
    std::vector<abi_sc> chal(n), resp(n);
    std::vector<unsigned long long> key_acc((size_t)RLC_ACC_WORDS * m, 0), s_acc((size_t)RLC_ACC_WORDS * cells * nch, 0);
    LaneRlcDecode fk = {pts.data(), ok.data(), flags.data(), sig, pk, n, g, n};
    run(fk, (size_t)m + nch);
    LaneRlcDecode f1 = {pts.data(), ok.data(), flags.data(), sig, pk, n, g, 0};
    run(f1, n);
    LaneEdVerifyScalars f2 = {chal.data(), resp.data(), sig, pk, msg, off, prehashed, ctx, ctx_len, 0};
    run(f2, n);
    LaneRlcZ f3 = {z.data(), g_rlc_seed, n, sh_r.zbits};
    run(f3, (n + RLC_Z_PER_LANE - 1) / RLC_Z_PER_LANE);
    const std::vector<uint32_t> z_for_digits = z; /* the product makes the R pair list before the decodes are in */
    /* the product's whole-batch pass: sums taken early (a signature counts if its key decodes), corrected when the R decodes are in;
     * its per-chunk pass takes them in one go.  Both orders here, by the parity of the chunk count, so that the CPU tier runs both. */
    uint32_t redo = 0;
    if (nch & 1) {
        LaneRlcWeights f4 = {z.data(), valid.data(), key_acc.data(), s_acc.data(), chal.data(), resp.data(), ok.data(), n, g, sh_r, 1u};
        run(f4, n);
        LaneRlcLate fl = {z.data(), valid.data(), key_acc.data(), s_acc.data(), chal.data(), resp.data(), ok.data(), n, g, sh_r, &redo};
        run(fl, n);
    } else {
        LaneRlcWeights f4 = {z.data(), valid.data(), key_acc.data(), s_acc.data(), chal.data(), resp.data(), ok.data(), n, g, sh_r, 0u};
        run(f4, n);
    }
    LaneRlcKeyScalars f5 = {kscal.data(), key_acc.data(), s_acc.data(), m, cells};
    run(f5, (size_t)m + nch);
    auto run_class = [&](const rlc_shape &sh, size_t count, const uint32_t *scal, uint32_t nwords, size_t p0, std::vector<pt> &total, bool subtract, const uint32_t *chunk_of) {
        const size_t npairs = count * sh.wn, nw = (size_t)sh.nch * sh.wn, nb = nw << sh.c;
        std::vector<uint32_t> keys(npairs), vals(npairs);
        std::vector<pt> buckets(nb), segsum(nw * sh.segs), nodesum(nw * sh.nodes), winsum(nw);
        total.resize(sh.nch);
        LaneRlcDigits f6 = {keys.data(), vals.data(), scal, nwords, p0, sh, chunk_of};
        run(f6, count);
        std::vector<std::pair<uint32_t, uint32_t>> pairs(npairs);
        for (size_t j = 0; j < npairs; j++) pairs[j] = {keys[j], vals[j]};
        std::stable_sort(pairs.begin(), pairs.end(), [](const std::pair<uint32_t, uint32_t> &a, const std::pair<uint32_t, uint32_t> &b) { return a.first < b.first; });
        for (size_t j = 0; j < npairs; j++) { keys[j] = pairs[j].first; vals[j] = pairs[j].second; }
        std::vector<uint32_t> bstart(nb), blen(nb), bid(nb), perm(nb);
        LaneRlcBucketRuns fr = {bstart.data(), blen.data(), bid.data(), keys.data(), npairs, sh};
        run(fr, nb);
        for (size_t b = 0; b < nb; b++) perm[b] = (uint32_t)b;
        std::stable_sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) { return blen[a] < blen[b]; });
        SlotRlcBucket f7 = {buckets.data(), keys.data(), vals.data(), npairs, pts.data(), sh, subtract ? ~0u : 0u, subtract ? valid.data() : nullptr, perm.data(), bstart.data()};
        run_sm(f7, nb);
        LaneRlcSegments f8 = {segsum.data(), buckets.data(), sh};
        run(f8, nw * sh.segs);
        LaneRlcNodes f9 = {nodesum.data(), segsum.data(), sh};
        run(f9, nw * sh.nodes);
        LaneRlcWindows f10 = {winsum.data(), nodesum.data(), sh};
        run(f10, nw);
        LaneRlcTotal f11 = {total.data(), winsum.data(), sh.wn, sh.nch > 1 ? sh.c : 0u};
        run(f11, sh.nch);
    };
    std::vector<pt> tot_r, tot_k;
    run_class(sh_k, (size_t)m + nch, kscal.data(), SC_WORDS, n, tot_k, false, kchunk.data());
    run_class(sh_r, n, z_for_digits.data(), RLC_ZWORDS, 0, tot_r, true, nullptr);
    LaneRlcVerdict f12 = {flags.data() + 1, tot_r.data(), tot_k.data(), flags.data()};
    run(f12, nch);
    /* chunks whose equation held take their statuses from `valid`; every other chunk goes through the ordinary path on its own */
    int all = 1;
    for (uint32_t ch = 0; ch < nch; ch++) {
        const size_t lo = (size_t)ch * csize, hi = lo + csize < n ? lo + csize : n;
        if (flags[1 + ch]) { for (size_t i = lo; i < hi; i++) st[i] = valid[i]; continue; }
        all = 0;
        goldilocks_ed448_verify_batch(st + lo, sig + 114 * lo, pk + 57 * lo, msg, off + lo, prehashed, ctx, ctx_len, hi - lo);
    }
    if (fast_path) *fast_path = all;
    return -1;
}
