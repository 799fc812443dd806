// algos.cuh -- per-lane scalar-multiplication algorithms and the CFRG glue built on point.cuh:
//   x448 Montgomery ladder, fixed-base signed comb, constant-time signed-window scalarmul,
//   warp-uniform double-scalar multiplication for verification, table construction, EdDSA pieces.
//
// "Constant time" here means per lane: no branch and no memory address depends on secret data;
// table lookups on secret indices scan the whole row and select with masks (the semantics of the
// reference's constant_time_lookup, src/include/constant_time.h:134-183).
#pragma once
#include "gf.cuh"
#include "sc.cuh"
#include "point.cuh"
#include "keccak.cuh"

#define COMB_N 5   /* reference goldilocks.c:25-27 */
#define COMB_T 5
#define COMB_S 18
#define COMB_ENTRIES (COMB_N << (COMB_T - 1)) /* 80 niels */
#define WNAF_FIXED_ENTRIES 32                 /* goldilocks.c:29: odd multiples 1B..63B */
#define WINDOW_BITS 5                         /* goldilocks.c:28 */
#define WINDOW_NTABLE 16
/* Verification-only fixed-base tables: the recoded scalar's 450 bits are WIDE_TABLES signed WIDE_BITS-bit digits, digit m looked up
 * in table m = odd multiples (2e+1) 2^(WIDE_BITS m) B as canonical affine niels, so scalar1 * B costs WIDE_TABLES additions and NO
 * doublings wherever it is added.  25 tables x 131072 entries x 192 B = 629 MB of HBM per device (the reads are random and miss the L2
 * either way: the per-key tables stream through it).  Public indices only -- never used on a secret scalar.
 * WIDE_BITS 15 / 30 digits (tables at the 45-bit column bases, digits added inside the rows) was the scheme up to build r02q. */
#ifndef WIDE_BITS
#define WIDE_BITS 18
#endif
#define WIDE_ENTRIES (1 << (WIDE_BITS - 1))
#define WIDE_TABLES ((450 + WIDE_BITS - 1) / WIDE_BITS)
#define WIDE_LANES 1024                        /* lanes that build one table at init */
#define WIDE_PER_LANE (WIDE_ENTRIES / WIDE_LANES)
/* Verification under a REPEATED public key splits both scalars into VSH_CHUNKS columns of VSH_ROWS
 * 5-bit digits: digit k = VSH_ROWS*c + r is served by tables of 2^(45c)*A (built once per key); the fixed base comes from the
 * doubling-free wide tables above, so one signature costs 8*5 doublings instead of 89*5.
 * Measured on the bench corpus (16 signatures per key), finish + key tables: 4 x 23 -> 50.7 + 5.3 ms, 6 x 15 -> 41.3 + 6.3,
 * 10 x 9 -> 36.0 + 7.8, 15 x 6 -> 34.0 + 9.6 (and 62 KB per key); 10 x 9 is never worse than 6 x 15 from 4 signatures per key on. */
#ifndef VSH_CHUNKS
#define VSH_CHUNKS 10
#define VSH_ROWS 9
#endif
#define VSH_SHIFT (VSH_ROWS * WINDOW_BITS)     /* bits between columns */

/* Doubling-free fixed-base table of the batched comb kernel: the reference's signed comb generalised to
 * (n, t, s) = (90, 5, 1) -- 90 rows of 16 canonical affine niels, row j = {(16 +- 8 +- 4 +- 2 +- 1) 2^(5j) B}.
 * Same 450-bit signed recoding as (5, 5, 18) (goldilocks.c:25-27, 842-843), but the 17 doublings are
 * gone: 90 constant-time additions.  270 KB, L2 resident. */
#define WIN_ROWS 90
#define TABLE_LANES (COMB_N + 2 + WIN_ROWS)    /* lanes that build the fixed tables at init */

// Device-resident fixed-base tables (built once per device by build_tables_lane()).
struct fixed_tables {
    niels comb[COMB_ENTRIES];        /* canonical affine niels, layout of goldilocks_448_precomputed_base */
    niels win[WIN_ROWS * 16];        /* doubling-free comb rows (see WIN_ROWS) */
    niels wnaf[WNAF_FIXED_ENTRIES];  /* canonical affine niels, layout of goldilocks_448_precomputed_wnaf_as_fe */
    pt base;                         /* decoded decaf base point */
};

// ---------------------------------------------------------------------------------------------
// Constant-time lookup: scan all `n` niels of a row, keep the one whose index matches.
// ---------------------------------------------------------------------------------------------
GD void niels_lookup_ct(niels &out, const niels *row, int n, uint32_t idx) {
    gf_set_zero(out.a); gf_set_zero(out.b); gf_set_zero(out.c);
#pragma unroll 1
    for (int e = 0; e < n; e++) {
        const gmask_t m = (gmask_t)(((uint64_t)((uint32_t)e ^ idx) - 1) >> 32); /* all-ones iff e == idx */
        gf_ld_or_masked<true>(out.a, &row[e].a, m);   /* fixed-base table: read-only */
        gf_ld_or_masked<true>(out.b, &row[e].b, m);
        gf_ld_or_masked<true>(out.c, &row[e].c, m);
    }
}
GD void pniels_lookup_ct(pniels &out, const pniels *row, int n, uint32_t idx) {
    gf_set_zero(out.n.a); gf_set_zero(out.n.b); gf_set_zero(out.n.c); gf_set_zero(out.z);
#pragma unroll 1
    for (int e = 0; e < n; e++) {
        const gmask_t m = (gmask_t)(((uint64_t)((uint32_t)e ^ idx) - 1) >> 32);
        gf_ld_or_masked<false>(out.n.a, &row[e].n.a, m); /* this lane's own table, written by this kernel */
        gf_ld_or_masked<false>(out.n.b, &row[e].n.b, m);
        gf_ld_or_masked<false>(out.n.c, &row[e].n.c, m);
        gf_ld_or_masked<false>(out.z, &row[e].z, m);
    }
}

// scalar1x = (scalar + adjustment) / 2 mod q  (goldilocks.c:420-421, 842-843)
GD void sc_recode_signed(sc &out, const sc &scalar) {
    sc adj, t;
    sc_set_adj(adj);
    sc_add(t, scalar, adj);
    sc_halve(out, t);
}

// ---------------------------------------------------------------------------------------------
// Fixed-base signed comb, constant time -- reference goldilocks.c:830-877.
// 18 rounds x 5 combs; 17 doublings + 90 mixed additions (T skipped before a doubling).
// `table` may live in shared, constant or global memory; every lane reads every entry.
// ---------------------------------------------------------------------------------------------
GD void comb_scalarmul(pt &out, const niels *table, const sc &scalar) {
    sc s1x;
    sc_recode_signed(s1x, scalar);
    niels ni;
#pragma unroll 1
    for (int i = COMB_S - 1; i >= 0; i--) {
        if (i != COMB_S - 1) pt_double(out, out, false);
#pragma unroll 1
        for (int j = 0; j < COMB_N; j++) {
            uint32_t tab = 0;
#pragma unroll
            for (int k = 0; k < COMB_T; k++) {
                const int bit = i + COMB_S * (k + j * COMB_T);
                if (bit < GOLDILOCKS_SCALAR_BITS_) tab |= sc_bit(s1x, bit) << k;
            }
            const gmask_t invert = (gmask_t)((int32_t)(tab >> (COMB_T - 1)) - 1);
            tab ^= invert;
            tab &= (1u << (COMB_T - 1)) - 1;
            niels_lookup_ct(ni, table + (j << (COMB_T - 1)), 1 << (COMB_T - 1), tab);
            niels_cond_neg(ni, invert);
            if (i != COMB_S - 1 || j != 0) pt_addsub_niels<false>(out, ni, j == COMB_N - 1 && i != 0);
            else niels_to_pt(out, ni);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Variable-base signed fixed-window table: odd multiples 1P,3P,...,(2n-1)P as pniels
// (goldilocks.c:382-403 prepare_fixed_window).
// ---------------------------------------------------------------------------------------------
GD void prepare_fixed_window(pniels *multiples, const pt &b, int ntable) {
    pt tmp;
    pniels pn;
    pt_double(tmp, b, false);
    pt_to_pniels(pn, tmp);
    pt_to_pniels(multiples[0], b);
    pt_copy(tmp, b);
#pragma unroll 1
    for (int i = 1; i < ntable; i++) {
        pt_addsub_pniels<false>(tmp, pn, false);
        pt_to_pniels(multiples[i], tmp);
    }
}

// 5-bit window starting at bit i of s (bits above 446+ are zero because s < q < 2^446)
GD uint32_t sc_window5(const sc &s, int i) { return sc_bits(s, i, WINDOW_BITS); }

// ---------------------------------------------------------------------------------------------
// Variable-base scalar multiplication, constant time -- reference goldilocks.c:405-465.
// `multiples` is this lane's scratch table of 16 pniels.
// ---------------------------------------------------------------------------------------------
GD void window_scalarmul(pt &a, const pt &b, const sc &scalar, pniels *multiples) {
    sc s1x;
    sc_recode_signed(s1x, scalar);
    prepare_fixed_window(multiples, b, WINDOW_NTABLE);
    pt tmp;
    pniels pn;
    bool first = true;
#pragma unroll 1
    for (int i = GOLDILOCKS_SCALAR_BITS_ - ((GOLDILOCKS_SCALAR_BITS_ - 1) % WINDOW_BITS) - 1; i >= 0; i -= WINDOW_BITS) {
        uint32_t bits = sc_window5(s1x, i);
        const gmask_t inv = (gmask_t)((int32_t)(bits >> (WINDOW_BITS - 1)) - 1);
        bits ^= inv;
        pniels_lookup_ct(pn, multiples, WINDOW_NTABLE, bits & (WINDOW_NTABLE - 1));
        niels_cond_neg(pn.n, inv);
        if (first) {
            pniels_to_pt(tmp, pn);
            first = false;
        } else {
#pragma unroll 1
            for (int j = 0; j < WINDOW_BITS - 1; j++) pt_double(tmp, tmp, true);
            pt_double(tmp, tmp, false);
            pt_addsub_pniels<false>(tmp, pn, i != 0);
        }
    }
    pt_copy(a, tmp);
}

// a = scalarb*b + scalarc*c, constant time -- reference goldilocks.c:467-541.
GD void window_double_scalarmul(pt &a, const pt &b, const sc &scalarb, const pt &c, const sc &scalarc,
                                pniels *multiples1, pniels *multiples2) {
    sc s1x, s2x;
    sc_recode_signed(s1x, scalarb);
    sc_recode_signed(s2x, scalarc);
    prepare_fixed_window(multiples1, b, WINDOW_NTABLE);
    prepare_fixed_window(multiples2, c, WINDOW_NTABLE);
    pt tmp;
    pniels pn;
    bool first = true;
#pragma unroll 1
    for (int i = GOLDILOCKS_SCALAR_BITS_ - ((GOLDILOCKS_SCALAR_BITS_ - 1) % WINDOW_BITS) - 1; i >= 0; i -= WINDOW_BITS) {
        uint32_t bits1 = sc_window5(s1x, i), bits2 = sc_window5(s2x, i);
        const gmask_t inv1 = (gmask_t)((int32_t)(bits1 >> (WINDOW_BITS - 1)) - 1);
        const gmask_t inv2 = (gmask_t)((int32_t)(bits2 >> (WINDOW_BITS - 1)) - 1);
        bits1 ^= inv1;
        bits2 ^= inv2;
        pniels_lookup_ct(pn, multiples1, WINDOW_NTABLE, bits1 & (WINDOW_NTABLE - 1));
        niels_cond_neg(pn.n, inv1);
        if (first) {
            pniels_to_pt(tmp, pn);
            first = false;
        } else {
#pragma unroll 1
            for (int j = 0; j < WINDOW_BITS - 1; j++) pt_double(tmp, tmp, true);
            pt_double(tmp, tmp, false);
            pt_addsub_pniels<false>(tmp, pn, false);
        }
        pniels_lookup_ct(pn, multiples2, WINDOW_NTABLE, bits2 & (WINDOW_NTABLE - 1));
        niels_cond_neg(pn.n, inv2);
        pt_addsub_pniels<false>(tmp, pn, i != 0);
    }
    pt_copy(a, tmp);
}

// ---------------------------------------------------------------------------------------------
// combo = scalar1*B + scalar2*base2 for PUBLIC inputs (signature verification).
//
// The reference (goldilocks.c:1260-1330) walks two wNAF expansions, so the positions of its
// additions depend on the scalars; executed per lane that diverges on almost every bit.  Only the
// resulting group element is observable (the projective representative is not part of the
// contract), so the GPU uses a warp-uniform schedule instead: both scalars are
// recoded with the constant-time paths' signed-digit recoding: scalar2 into 90 odd 5-bit digits,
// scalar1 into 25 odd 18-bit digits.  Every lane does, per 5-bit window, 5 doublings + one
// variable-base add (direct index into its own 16-pniels table) and, after the last window, 25
// fixed-base adds (direct indices into the init-time tables of odd multiples of 2^(18m) B): 445
// doublings + 90 + 25 additions, against the reference's expected 444 + 75 + 56.  Indices are
// public, so lookups are plain loads.
// ---------------------------------------------------------------------------------------------
// Implementation: s_base_double_scalarmul (slot_algos.cuh).

// ---------------------------------------------------------------------------------------------
// Table construction (runs once per device, a handful of lanes).
// ---------------------------------------------------------------------------------------------
// out[i] = 1/in[i] for i < n, one inversion (Montgomery's trick; reference goldilocks.c:703-726).
GD void gf_batch_invert(gf *out, const gf *in, int n) {
    gf acc, t;
    gf_copy(acc, in[0]);
    gf_copy(out[0], in[0]);
    for (int i = 1; i < n; i++) { gf_mul(acc, acc, in[i]); gf_copy(out[i], acc); } /* out[i] = prod_{k<=i} in[k] */
    gf_invert(acc, acc);
    for (int i = n - 1; i > 0; i--) {
        gf_mul(t, acc, out[i - 1]); /* 1/in[i] */
        gf_mul(acc, acc, in[i]);
        gf_copy(out[i], t);
    }
    gf_copy(out[0], acc);
}
// Normalise projective niels to canonical affine niels (goldilocks.c:728-755).
GD void normalize_niels(niels *table, const gf *zs, gf *zis, int n) {
    gf_batch_invert(zis, zs, n);
    for (int i = 0; i < n; i++) {
        gf_mul(table[i].a, table[i].a, zis[i]); gf_strong_reduce(table[i].a);
        gf_mul(table[i].b, table[i].b, zis[i]); gf_strong_reduce(table[i].b);
        gf_mul(table[i].c, table[i].c, zis[i]); gf_strong_reduce(table[i].c);
    }
}
// One comb (16 entries) of the signed fixed-base comb table:
//   entry[idx] = ( 2^(18*4) + sum_{k<4} (2*bit_k(idx) - 1) * 2^(18k) ) * 2^(90*comb) * B
// Same table as the reference's precompute() (goldilocks.c:757-818), built comb-by-comb so the
// five combs can be produced by five lanes in parallel.
GD void build_comb_row(niels *out16, const pt &base, int skip_doublings, int spacing) {
    pt working, start, doubles[COMB_T - 1];
    gf zs[16], zis[16];
    pt_copy(working, base);
    for (int k = 0; k < skip_doublings; k++) pt_double(working, working, false);
    for (int j = 0; j < COMB_T; j++) {
        if (j) pt_add(start, start, working);
        else pt_copy(start, working);
        if (j == COMB_T - 1) break;
        pt_double(working, working, false);
        pt_copy(doubles[j], working);          /* 2 * 2^(spacing j) * (row base) */
        for (int k = 0; k < spacing - 1; k++) pt_double(working, working, false);
    }
    /* start = all teeth positive = entry 15; walk a Gray code flipping one tooth at a time */
    for (int j = 0;; j++) {
        const int gray = j ^ (j >> 1);
        const int idx = 15 ^ gray;
        pniels pn;
        pt_to_pniels(pn, start);
        gf_copy(out16[idx].a, pn.n.a); gf_copy(out16[idx].b, pn.n.b); gf_copy(out16[idx].c, pn.n.c);
        gf_copy(zs[idx], pn.z);
        if (j >= 15) break;
        int delta = (j + 1) ^ ((j + 1) >> 1) ^ gray, k = 0;
        while (delta > 1) { delta >>= 1; k++; }
        if (gray & (1 << k)) pt_add(start, start, doubles[k]);
        else pt_sub(start, start, doubles[k]);
    }
    normalize_niels(out16, zs, zis, 16);
}
GD void build_comb(niels *out16, const pt &base, int comb) { build_comb_row(out16, base, COMB_S * COMB_T * comb, COMB_S); }
// Odd multiples 1B,3B,...,63B as canonical affine niels (goldilocks.c:1204-1258).
GD void build_wnaf_base(niels *out32, const pt &base) {
    gf zs[WNAF_FIXED_ENTRIES], zis[WNAF_FIXED_ENTRIES];
    pt tmp, twob;
    pniels pn;
    pt_double(twob, base, false);
    pt_copy(tmp, base);
    for (int i = 0; i < WNAF_FIXED_ENTRIES; i++) {
        if (i) pt_add(tmp, tmp, twob);
        pt_to_pniels(pn, tmp);
        gf_copy(out32[i].a, pn.n.a); gf_copy(out32[i].b, pn.n.b); gf_copy(out32[i].c, pn.n.c);
        gf_copy(zs[i], pn.z);
    }
    normalize_niels(out32, zs, zis, WNAF_FIXED_ENTRIES);
}
// One lane's share of the verification table: entries [per*lane, per*(lane+1)) = odd multiples
// (2e+1)B.  `tmp`/`pre` are global scratch (one pniels / one gf per entry).  Start point from the
// comb table, then repeated +2B; normalised with one inversion per lane (Montgomery's trick).
GD void build_wide_lane(niels *out, pniels *tmp, gf *pre, const niels *comb, int lane_all) {
    const int c = lane_all / WIDE_LANES, lane = lane_all % WIDE_LANES;   /* table c holds multiples of 2^(WIDE_BITS c) B */
    out += (size_t)c * WIDE_ENTRIES; tmp += (size_t)c * WIDE_ENTRIES; pre += (size_t)c * WIDE_ENTRIES;
    pt twob, p;
    sc start, two;
    {   /* (2 e0 + 1) 2^(WIDE_BITS c) and 2 * 2^(WIDE_BITS c) mod q: the top table's multiples pass q (2^450 > q) */
        sc odd, pw;
        sc_set_zero(odd);
        sc_set_zero(pw);
        odd.w[0] = 2u * (uint32_t)(WIDE_PER_LANE * lane) + 1u;
        const int sh = WIDE_BITS * c;                      /* <= 432: 2^sh itself is below q */
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int i = 0; i < SC_WORDS; i++) pw.w[i] = (i == sh / 32) ? (1u << (sh % 32)) : 0u;
        sc_mul(start, odd, pw);
        sc_add(two, pw, pw);
    }
    comb_scalarmul(twob, comb, two);
    comb_scalarmul(p, comb, start);
    const int e0 = WIDE_PER_LANE * lane;
    for (int i = 0; i < WIDE_PER_LANE; i++) {
        if (i) pt_add(p, p, twob);
        pt_to_pniels(tmp[e0 + i], p);
        if (i == 0) gf_copy(pre[e0], tmp[e0].z);
        else gf_mul(pre[e0 + i], pre[e0 + i - 1], tmp[e0 + i].z);
    }
    gf inv, zi;
    gf_invert(inv, pre[e0 + WIDE_PER_LANE - 1]);
    for (int i = WIDE_PER_LANE - 1; i >= 0; i--) {
        if (i) { gf_mul(zi, inv, pre[e0 + i - 1]); gf_mul(inv, inv, tmp[e0 + i].z); }
        else gf_copy(zi, inv);
        niels n;
        gf_mul(n.a, tmp[e0 + i].n.a, zi); gf_strong_reduce(n.a);
        gf_mul(n.b, tmp[e0 + i].n.b, zi); gf_strong_reduce(n.b);
        gf_mul(n.c, tmp[e0 + i].n.c, zi); gf_strong_reduce(n.c);
        gf_copy(out[e0 + i].a, n.a); gf_copy(out[e0 + i].b, n.b); gf_copy(out[e0 + i].c, n.c);
    }
}
GD void build_tables_lane(fixed_tables *ft, int lane) {
    const uint32_t bw[14] = GOLD_CONST_BASE_WORDS;
    pt base;
    (void)pt_decode(base, bw, 0);
    if (lane < COMB_N) build_comb(ft->comb + 16 * lane, base, lane);
    else if (lane == COMB_N) build_wnaf_base(ft->wnaf, base);
    else if (lane == COMB_N + 1) pt_copy(ft->base, base);
    else if (lane < TABLE_LANES) { const int row = lane - (COMB_N + 2); build_comb_row(ft->win + 16 * row, base, COMB_T * row, 1); }
}

// ---------------------------------------------------------------------------------------------
// EdDSA glue (reference src/eddsa.c)
// ---------------------------------------------------------------------------------------------
// clamp (eddsa.c:34-48) applied to the 57 hash bytes held as 15 words (byte 56 in word 14)
GD void ed448_clamp_words(uint32_t w[15]) {
    w[0] &= ~3u;
    w[14] &= ~0xffu;          /* byte 56 = 0 */
    w[13] |= 0x80000000u;     /* top bit of byte 55 */
}
// "SigEd448" || prehashed || ctx_len || ctx   (eddsa.c:51-74; dom byte = prehashed ? 1 : 0)
template <typename CtxAt>
GD void ed448_hash_init_with_dom(shake256_ctx &h, uint32_t prehashed, CtxAt ctx_at, uint32_t ctx_len) {
    shake256_init(h);
    const uint8_t dom[8] = {'S', 'i', 'g', 'E', 'd', '4', '4', '8'};
#pragma unroll 1
    for (int i = 0; i < 8; i++) shake256_absorb_byte(h, dom[i]);
    shake256_absorb_byte(h, prehashed ? 1 : 0);
    shake256_absorb_byte(h, (uint8_t)ctx_len);
#pragma unroll 1
    for (uint32_t i = 0; i < ctx_len; i++) shake256_absorb_byte(h, ctx_at(i));
}
