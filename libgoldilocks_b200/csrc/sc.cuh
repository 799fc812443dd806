// sc.cuh -- arithmetic mod the group order q (446 bits), one scalar per lane, 14 x 32-bit words.
//
// Replaces the reference's src/scalar.c (7 x 64-bit limbs, __uint128_t chains) with word-serial
// Montgomery multiplication on 32-bit words: every step is one IMAD.WIDE.U32 whose 64-bit result
// absorbs the running carry, so there are no carry flags.  All results are canonical (< q), which
// is what makes them byte-identical to the reference's outputs.
#pragma once
#include "gf.cuh"
#include "consts.cuh"

#define SC_WORDS 14
#define GOLDILOCKS_SCALAR_BITS_ 446 /* reference point_448.h:27 */
struct sc { uint32_t w[SC_WORDS]; };

// Constants are function-local constexpr tables: after inlining and unrolling every index is a
// compile-time constant, so the words become immediates of the IMAD/IADD3 instructions.
GD uint32_t sc_q(int i) { const uint32_t t[SC_WORDS] = GOLD_CONST_SC_Q; return t[i]; }
GD uint32_t sc_r2(int i) { const uint32_t t[SC_WORDS] = GOLD_CONST_SC_R2; return t[i]; }
GD uint32_t sc_adj(int i) { const uint32_t t[SC_WORDS] = GOLD_CONST_SC_ADJ; return t[i]; }

GD void sc_set_zero(sc &a) {
#pragma unroll
    for (int i = 0; i < SC_WORDS; i++) a.w[i] = 0;
}
GD void sc_copy(sc &o, const sc &a) {
#pragma unroll
    for (int i = 0; i < SC_WORDS; i++) o.w[i] = a.w[i];
}
GD void sc_set_q(sc &o) {
#pragma unroll
    for (int i = 0; i < SC_WORDS; i++) o.w[i] = sc_q(i);
}
GD void sc_set_r2(sc &o) {
#pragma unroll
    for (int i = 0; i < SC_WORDS; i++) o.w[i] = sc_r2(i);
}
GD void sc_set_adj(sc &o) {
#pragma unroll
    for (int i = 0; i < SC_WORDS; i++) o.w[i] = sc_adj(i);
}

// out = {extra, accum} - sub, then + q if that went negative.  (reference scalar.c:27-53 sc_subx)
GD void sc_subx(sc &out, const uint32_t accum[SC_WORDS], const uint32_t sub[SC_WORDS], uint32_t extra) {
    int64_t chain = 0;
#pragma unroll
    for (int i = 0; i < SC_WORDS; i++) {
        chain = chain + accum[i] - sub[i];
        out.w[i] = (uint32_t)chain;
        chain >>= 32;
    }
    uint32_t borrow = (uint32_t)chain + extra; /* 0 or 0xffffffff */
    uint64_t c2 = 0;
#pragma unroll
    for (int i = 0; i < SC_WORDS; i++) {
        c2 = c2 + out.w[i] + (sc_q(i) & borrow);
        out.w[i] = (uint32_t)c2;
        c2 >>= 32;
    }
}

GD void sc_add(sc &out, const sc &a, const sc &b) { /* reference scalar.c:176-189 */
    uint64_t chain = 0;
    uint32_t t[SC_WORDS];
#pragma unroll
    for (int i = 0; i < SC_WORDS; i++) {
        chain = chain + a.w[i] + b.w[i];
        t[i] = (uint32_t)chain;
        chain >>= 32;
    }
    sc q;
    sc_set_q(q);
    sc_subx(out, t, q.w, (uint32_t)chain);
}
GD void sc_sub(sc &out, const sc &a, const sc &b) { /* reference scalar.c:168-174 */
    sc_subx(out, a.w, b.w, 0);
}
GD void sc_neg(sc &out, const sc &a) {
    uint32_t z[SC_WORDS];
#pragma unroll
    for (int i = 0; i < SC_WORDS; i++) z[i] = 0;
    sc_subx(out, z, a.w, 0);
}

// out = a * b / 2^448 mod q.  a < 2^448, b < 2^448 with a*b < q*2^448 (true whenever one side is
// canonical).  (reference scalar.c:55-91 sc_montmul, word-serial CIOS)
GD void sc_montmul(sc &out, const sc &a, const sc &b) {
    uint32_t acc[SC_WORDS + 1];
#pragma unroll
    for (int i = 0; i <= SC_WORDS; i++) acc[i] = 0;
    uint32_t hi_carry = 0;
#pragma unroll 1
    for (int i = 0; i < SC_WORDS; i++) {
        uint32_t mand = a.w[i];
        uint64_t chain = 0;
#pragma unroll
        for (int j = 0; j < SC_WORDS; j++) {
            chain += (uint64_t)mand * b.w[j] + acc[j];
            acc[j] = (uint32_t)chain;
            chain >>= 32;
        }
        acc[SC_WORDS] = (uint32_t)chain;
        mand = acc[0] * GOLD_SC_MONT32;
        chain = 0;
#pragma unroll
        for (int j = 0; j < SC_WORDS; j++) {
            chain += (uint64_t)mand * sc_q(j) + acc[j];
            if (j) acc[j - 1] = (uint32_t)chain;
            chain >>= 32;
        }
        chain += acc[SC_WORDS];
        chain += hi_carry;
        acc[SC_WORDS - 1] = (uint32_t)chain;
        hi_carry = (uint32_t)(chain >> 32);
    }
    sc q;
    sc_set_q(q);
    sc_subx(out, acc, q.w, hi_carry);
}
GD void sc_mul(sc &out, const sc &a, const sc &b) { /* reference scalar.c:93-100 */
    sc t, r2;
    sc_set_r2(r2);
    sc_montmul(t, a, b);
    sc_montmul(out, t, r2);
}
// out = a^(q-2) mod q = 1/a (0 for a = 0).  The reference walks the same public exponent with a sliding
// window (scalar.c:107-166); the result is the unique inverse mod q, so a plain left-to-right
// square-and-multiply in the Montgomery domain gives identical bytes.  The exponent is public.
GD void sc_invert(sc &out, const sc &a) {
    sc am, acc, r2, one;
    sc_set_r2(r2);
    sc_montmul(am, a, r2);          /* a R */
    sc_copy(acc, am);               /* top bit of q - 2 (bit 445) is set */
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int i = GOLDILOCKS_SCALAR_BITS_ - 2; i >= 0; i--) {
        sc_montmul(acc, acc, acc);
        /* bit i of q - 2: q is odd and q = ...11 (low word 0xab5844f3), so only the low word differs from q */
        uint32_t w = sc_q(i >> 5);
        if ((i >> 5) == 0) w -= 2u;
        if ((w >> (i & 31)) & 1u) sc_montmul(acc, acc, am);
    }
    sc_set_zero(one);
    one.w[0] = 1;
    sc_montmul(out, acc, one);      /* leave the Montgomery domain */
}
GD void sc_reduce_short(sc &out, const sc &a) { /* "ham-handed reduce": a*1/R then *R^2/R (scalar.c:246) */
    sc one, t, r2;
    sc_set_zero(one);
    one.w[0] = 1;
    sc_set_r2(r2);
    sc_montmul(t, a, one);
    sc_montmul(out, t, r2);
}

GD void sc_halve(sc &out, const sc &a) { /* reference scalar.c:316-332 */
    uint32_t mask = (uint32_t)(-(int32_t)(a.w[0] & 1));
    uint64_t chain = 0;
#pragma unroll
    for (int i = 0; i < SC_WORDS; i++) {
        chain = chain + a.w[i] + (sc_q(i) & mask);
        out.w[i] = (uint32_t)chain;
        chain >>= 32;
    }
#pragma unroll
    for (int i = 0; i < SC_WORDS - 1; i++) out.w[i] = (out.w[i] >> 1) | (out.w[i + 1] << 31);
    out.w[SC_WORDS - 1] = (out.w[SC_WORDS - 1] >> 1) | ((uint32_t)chain << 31);
}

// Little-endian bytes of any length -> canonical scalar (value mod q).  `get(k)` returns byte k.
// (reference scalar.c:257-293 scalar_decode_long: Horner over 56-byte chunks from the top)
template <typename ByteAt>
GD void sc_decode_long(sc &out, ByteAt get, int len) {
    if (len == 0) { sc_set_zero(out); return; }
    int i = len - (len % 56);
    if (i == len) i -= 56;
    sc t1;
    sc_set_zero(t1);
    for (int k = 0; k < len - i; k++) t1.w[k >> 2] |= (uint32_t)get(i + k) << (8 * (k & 3));
    if (len == 56) { sc_reduce_short(out, t1); return; }
    if (i == 0) { /* fewer than 56 bytes: already < 2^440 < q */ sc_copy(out, t1); return; }
#pragma unroll 1
    while (i) {
        i -= 56;
        sc t2, t3, r2;
        sc_set_r2(r2);
        sc_montmul(t3, t1, r2); /* t1 * 2^448 mod q */
        sc_set_zero(t2);
        for (int k = 0; k < 56; k++) t2.w[k >> 2] |= (uint32_t)get(i + k) << (8 * (k & 3));
        sc_reduce_short(t2, t2);
        sc_add(t1, t3, t2);
    }
    sc_copy(out, t1);
}

// ---- wide reduction by folding (the EdDSA hot path: 114-byte SHAKE outputs and 57-byte S strings) -----------------------
// q = 2^446 - c with c of 224 bits, so 2^448 == 4c (mod q): a value of N > 14 words is its low 14 words plus its high words
// times 4c (8 words), until 15 words are left; then the bits from 446 up are folded with c and one conditional subtraction
// makes the result canonical.  Same value as sc_decode_long (reference scalar.c:257-293: the canonical residue), at 255
// word multiplications for 114 bytes instead of six Montgomery multiplications (2 352).  c = 2^446 - q, from GOLD_CONST_SC_Q.
#define GOLD_CONST_SC_C { 0x54a7bb0du, 0xdc873d6du, 0x723a70aau, 0xde933d8du, 0x5129c96fu, 0x3bb124b6u, 0x8335dc16u }
#define GOLD_CONST_SC_4C { 0x529eec34u, 0x721cf5b5u, 0xc8e9c2abu, 0x7a4cf635u, 0x44a725bfu, 0xeec492d9u, 0x0cd77058u, 0x00000002u }
// y (OUT words) = x[0..14) + x[14..N) * 4c; the caller guarantees the value fits OUT words.
template <int N, int OUT>
GD void sc_fold448(uint32_t (&y)[OUT], const uint32_t (&x)[N]) {
    const uint32_t c4[8] = GOLD_CONST_SC_4C;
#pragma unroll
    for (int i = 0; i < OUT; i++) y[i] = i < SC_WORDS ? x[i] : 0u;
    uint32_t pend = 0; /* carry out of the previous row, due at position i + 8 */
#pragma unroll
    for (int i = 0; i < N - SC_WORDS; i++) {
        const uint32_t h = x[SC_WORDS + i];
        uint64_t chain = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (i + j < OUT) {
                chain += (uint64_t)h * c4[j] + y[i + j];
                y[i + j] = (uint32_t)chain;
                chain >>= 32;
            }
        }
        if (i + 8 < OUT) {
            chain += (uint64_t)y[i + 8] + pend;
            y[i + 8] = (uint32_t)chain;
            pend = (uint32_t)(chain >> 32);
        }
    }
#pragma unroll
    for (int k = N - SC_WORDS + 8; k < OUT; k++) { /* the last row's carry ripples through the low words it lands in */
        const uint64_t t = (uint64_t)y[k] + pend;
        y[k] = (uint32_t)t;
        pend = (uint32_t)(t >> 32);
    }
}
// x < 2^448 + small (15 words, word 14 <= 1 after the folds) -> canonical residue
GD void sc_fold_finish(sc &out, const uint32_t (&x)[15]) {
    uint32_t y[15];
    sc_fold448<15, 15>(y, x);                /* word 14 is 0 now (see the bounds above) */
    const uint32_t c[7] = GOLD_CONST_SC_C;
    const uint32_t hi = y[13] >> 30;         /* bits 446, 447 */
    y[13] &= 0x3fffffffu;
    uint64_t chain = 0;
#pragma unroll
    for (int j = 0; j < SC_WORDS; j++) {
        chain += (uint64_t)y[j] + (j < 7 ? (uint64_t)hi * c[j] : 0u);
        y[j] = (uint32_t)chain;
        chain >>= 32;
    }                                        /* < 2^446 + 3 * 2^224 < 2q */
    uint32_t v[SC_WORDS];
#pragma unroll
    for (int j = 0; j < SC_WORDS; j++) v[j] = y[j];
    sc q;
    sc_set_q(q);
    sc_subx(out, v, q.w, 0);
}
// 114 little-endian bytes in 29 words (the upper half of word 28 must be zero)
GD void sc_reduce_114(sc &out, const uint32_t (&w)[29]) {
    uint32_t a[24], b[19], c[15];
    sc_fold448<29, 24>(a, w);   /* < 2^706 */
    sc_fold448<24, 19>(b, a);   /* < 2^485 */
    sc_fold448<19, 15>(c, b);   /* < 2^448 + 2^263 */
    sc_fold_finish(out, c);
}
// 57 little-endian bytes in 15 words (the upper three bytes of word 14 must be zero)
GD void sc_reduce_57(sc &out, const uint32_t (&w)[15]) {
    uint32_t c[15];
    sc_fold448<15, 15>(c, w);   /* < 2^448 + 2^234 */
    sc_fold_finish(out, c);
}

// bit `pos` of the scalar (0 for pos >= 448)
GD uint32_t sc_bit(const sc &a, int pos) {
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < SC_WORDS; i++) r |= (i == (pos >> 5)) ? a.w[i] : 0u; /* no dynamic register indexing */
    return (r >> (pos & 31)) & 1u;
}
// `nbits` (<= 31) bits starting at bit `pos`
GD uint32_t sc_bits(const sc &a, int pos, int nbits) {
    uint32_t lo = 0, hi = 0;
    const int wi = pos >> 5;
#pragma unroll
    for (int i = 0; i < SC_WORDS; i++) {
        lo |= (i == wi) ? a.w[i] : 0u;
        hi |= (i == wi + 1) ? a.w[i] : 0u;
    }
    uint64_t x = ((uint64_t)hi << 32) | lo;
    return (uint32_t)(x >> (pos & 31)) & ((1u << nbits) - 1u);
}

// ---------------------------------------------------------------------------------------------
// Half-size multipliers for a verification equation (Antipa, Brown, Gallant, Lambert, Struik: "Accelerated verification
// of ECDSA signatures", SAC 2005).  For a scalar c < q there are integers u, v with
//     v c == u (mod q),   0 <= u < 2^223,   0 < |v| < 2^223 :
// the extended Euclidean algorithm on (q, c), stopped at the first remainder r_i below 2^223; its cofactor t_i obeys
// |t_i| r_(i-1) <= q and r_(i-1) >= 2^223.  An equation  s B + c A == R  in a group of prime order q then holds iff
//     (v s) B + u A - v R == 0 ,
// which needs HALF the doublings: u and v are 223-bit multipliers and the full-size one sits on the fixed base.
// c = 0 gives u = 0, v = 1.  Public data only -- the number of steps depends on c.
// Quotients are taken by shift-and-subtract (they are small: 1.5 bits on average); everything is statically indexed.
// Returns all-ones when v is negative; u and |v| come back as 14-word scalars (upper half zero).
// ---------------------------------------------------------------------------------------------
#define HG_TW 8 /* words of a cofactor (< 2^224, one spare bit for the aligned divisor's companion) */
GD int hg_bitlen(const uint32_t (&a)[SC_WORDS]) {
    int len = 0;
#pragma unroll
    for (int i = 0; i < SC_WORDS; i++) {
#if defined(__CUDA_ARCH__)
        if (a[i]) len = 32 * i + 32 - __clz((int)a[i]);
#else
        if (a[i]) len = 32 * i + 32 - __builtin_clz(a[i]);
#endif
    }
    return len;
}
template <int N>
GD void hg_shl(uint32_t (&a)[N], int s) {
    const int ws = s >> 5, bs = s & 31;
#pragma unroll 1
    for (int k = 0; k < ws; k++) {
#pragma unroll
        for (int i = N - 1; i > 0; i--) a[i] = a[i - 1];
        a[0] = 0;
    }
    if (bs) {
#pragma unroll
        for (int i = N - 1; i > 0; i--) a[i] = (a[i] << bs) | (a[i - 1] >> (32 - bs));
        a[0] <<= bs;
    }
}
template <int N>
GD void hg_shr1(uint32_t (&a)[N]) {
#pragma unroll
    for (int i = 0; i < N - 1; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 31);
    a[N - 1] >>= 1;
}
// 62 bits of a starting at bit k (k + 62 <= 448 + 32 is fine: words past the end read as zero)
GD uint64_t hg_window62(const uint32_t (&a)[SC_WORDS], int k) {
    const int wi = k >> 5, bs = k & 31;
    uint32_t w0 = 0, w1 = 0, w2 = 0;
#pragma unroll
    for (int i = 0; i < SC_WORDS; i++) {
        w0 |= (i == wi) ? a[i] : 0u;
        w1 |= (i == wi + 1) ? a[i] : 0u;
        w2 |= (i == wi + 2) ? a[i] : 0u;
    }
    uint64_t lo = ((uint64_t)w1 << 32) | w0;
    if (bs) lo = (lo >> bs) | ((uint64_t)w2 << (64 - bs));
    return lo & 0x3fffffffffffffffull;
}
GD uint64_t hg_mulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}
// out = |m1 * x - m2 * y| for 32-bit multipliers, where the caller knows which of the two products is the larger one
// (y_minus_x: out = m2 y - m1 x).  The true result fits N words.
template <int N>
GD void hg_submul2(uint32_t (&out)[N], uint32_t m1, const uint32_t (&x)[N], uint32_t m2, const uint32_t (&y)[N], bool y_minus_x) {
    uint64_t cx = 0, cy = 0;
    int64_t borrow = 0;
#pragma unroll
    for (int i = 0; i < N; i++) {
        cx += (uint64_t)m1 * x[i];
        cy += (uint64_t)m2 * y[i];
        const uint32_t px = (uint32_t)cx, py = (uint32_t)cy;
        cx >>= 32; cy >>= 32;
        borrow += y_minus_x ? (int64_t)py - (int64_t)px : (int64_t)px - (int64_t)py;
        out[i] = (uint32_t)borrow;
        borrow >>= 32;
    }
}
// out = m1 * x + m2 * y (fits N words)
template <int N>
GD void hg_addmul2(uint32_t (&out)[N], uint32_t m1, const uint32_t (&x)[N], uint32_t m2, const uint32_t (&y)[N]) {
    uint64_t cx = 0, cy = 0, carry = 0;
#pragma unroll
    for (int i = 0; i < N; i++) {
        cx += (uint64_t)m1 * x[i];
        cy += (uint64_t)m2 * y[i];
        carry += (uint64_t)(uint32_t)cx + (uint32_t)cy;
        cx >>= 32; cy >>= 32;
        out[i] = (uint32_t)carry;
        carry >>= 32;
    }
}
// One exact division step: r0 <- r0 mod r1, t0 <- t0 + (r0 div r1) t1, by shift-and-subtract (quotients are small: 1.5 bits on
// average), then the roles swap.  Everything statically indexed.
GD void hg_exact_step(uint32_t (&r0)[SC_WORDS], uint32_t (&r1)[SC_WORDS], uint32_t (&t0)[HG_TW], uint32_t (&t1)[HG_TW]) {
    uint32_t d[SC_WORDS], td[HG_TW];
    int s = hg_bitlen(r0) - hg_bitlen(r1);                  /* r0 > r1 throughout, so s >= 0 */
#pragma unroll
    for (int i = 0; i < SC_WORDS; i++) d[i] = r1[i];
#pragma unroll
    for (int i = 0; i < HG_TW; i++) td[i] = t1[i];
    hg_shl(d, s);
    hg_shl(td, s);
#pragma unroll 1
    for (; s >= 0; s--) {
        uint32_t diff[SC_WORDS];
        int64_t chain = 0;
#pragma unroll
        for (int i = 0; i < SC_WORDS; i++) {
            chain = chain + r0[i] - d[i];
            diff[i] = (uint32_t)chain;
            chain >>= 32;
        }
        const uint32_t ge = ~(uint32_t)chain;               /* all-ones when r0 >= d */
#pragma unroll
        for (int i = 0; i < SC_WORDS; i++) r0[i] = (diff[i] & ge) | (r0[i] & ~ge);
        uint64_t carry = 0;
#pragma unroll
        for (int i = 0; i < HG_TW; i++) {
            carry = carry + t0[i] + (td[i] & ge);
            t0[i] = (uint32_t)carry;
            carry >>= 32;
        }
        hg_shr1(d);
        hg_shr1(td);
    }
#pragma unroll
    for (int i = 0; i < SC_WORDS; i++) { const uint32_t x = r0[i]; r0[i] = r1[i]; r1[i] = x; }
#pragma unroll
    for (int i = 0; i < HG_TW; i++) { const uint32_t x = t0[i]; t0[i] = t1[i]; t1[i] = x; }
}
// Lehmer's batching (Knuth, TAOCP 4.5.2 Algorithm L): run Euclid on the leading 62 bits (u^, v^) of (r0, r1) with cofactors A, B, C, D
// for as long as the quotient is the same at both ends of the interval the truncation leaves -- floor((u^ + A)/(v^ + C)) ==
// floor((u^ + B)/(v^ + D)) -- and the cofactors stay below 2^32, then apply the 2 x 2 matrix to the full numbers: ~18 division steps
// for four multiply-accumulate passes.  The cofactors are kept as magnitudes; their signs alternate with the parity of the step count
// (even: A, D > 0 >= B, C; odd: the other way round).  A step is refused when the new remainder could fall below 2^223 (the stopping
// point must not be jumped over): v' >= 2^k (v^' - |D'|) because the truncated tails weigh in with at most the cofactors.
// Returns the number of steps taken (0: the caller takes one exact step instead).
GD int hg_lehmer_batch(uint32_t (&r0)[SC_WORDS], uint32_t (&r1)[SC_WORDS], uint32_t (&t0)[HG_TW], uint32_t (&t1)[HG_TW]) {
    const int k = hg_bitlen(r0) - 62;                       /* r0 > r1 >= 2^223: k >= 162 */
    uint64_t uh = hg_window62(r0, k), vh = hg_window62(r1, k);
    const uint64_t vmin = k >= 223 ? 1ull : ((1ull << (223 - k)) + 1ull);
    uint64_t A = 1, B = 0, C = 0, D = 1;
    int steps = 0;
#pragma unroll 1
    for (;;) {
        const bool even = (steps & 1) == 0;
        const uint64_t n1 = even ? uh + A : uh - A, d1 = even ? vh - C : vh + C;
        const uint64_t n2 = even ? uh - B : uh + B, d2 = even ? vh + D : vh - D;
        if (d1 == 0 || d2 == 0) break;
        uint64_t q;
        if (n1 < 4 * d1) { q = 0; uint64_t x = n1; while (x >= d1) { x -= d1; q++; } }   /* two quotients in three are below 4 */
        else q = n1 / d1;
        if (q >= (1ull << 32)) break;
        if (hg_mulhi64(q, d2)) break;
        const uint64_t qd2 = q * d2;
        if (qd2 > n2 || n2 - qd2 >= d2) break;              /* floor(n2 / d2) != q */
        const uint64_t Cn = A + q * C, Dn = B + q * D;
        if (Cn >= (1ull << 32) || Dn >= (1ull << 32)) break;
        const uint64_t vn = uh - q * vh;
        if (vn < vmin + Dn) break;
        A = C; C = Cn; B = D; D = Dn; uh = vh; vh = vn;
        steps++;
    }
    if (steps == 0) return 0;
    const bool odd = (steps & 1) != 0;
    uint32_t n0[SC_WORDS], n1w[SC_WORDS], s0[HG_TW], s1[HG_TW];
    hg_submul2(n0, (uint32_t)A, r0, (uint32_t)B, r1, odd);   /* even: A r0 - B r1 ; odd: B r1 - A r0 */
    hg_submul2(n1w, (uint32_t)C, r0, (uint32_t)D, r1, !odd); /* even: D r1 - C r0 ; odd: C r0 - D r1 */
    hg_addmul2(s0, (uint32_t)A, t0, (uint32_t)B, t1);
    hg_addmul2(s1, (uint32_t)C, t0, (uint32_t)D, t1);
#pragma unroll
    for (int i = 0; i < SC_WORDS; i++) { r0[i] = n0[i]; r1[i] = n1w[i]; }
#pragma unroll
    for (int i = 0; i < HG_TW; i++) { t0[i] = s0[i]; t1[i] = s1[i]; }
    return steps;
}
GD gmask_t sc_half_gcd(sc &u, sc &v, const sc &c) {
    uint32_t r0[SC_WORDS], r1[SC_WORDS], t0[HG_TW], t1[HG_TW];
#pragma unroll
    for (int i = 0; i < SC_WORDS; i++) { r0[i] = sc_q(i); r1[i] = c.w[i]; }
#pragma unroll
    for (int i = 0; i < HG_TW; i++) { t0[i] = 0; t1[i] = i == 0 ? 1u : 0u; }
    gmask_t neg = 0;
#pragma unroll 1
    for (;;) {
        uint32_t high = r1[6] >> 31;                       /* r1 >= 2^223 ? */
#pragma unroll
        for (int i = 7; i < SC_WORDS; i++) high |= r1[i];
        if (!high) break;
        int steps = hg_lehmer_batch(r0, r1, t0, t1);
        if (steps == 0) { hg_exact_step(r0, r1, t0, t1); steps = 1; }
        if (steps & 1) neg = ~neg;
    }
#pragma unroll
    for (int i = 0; i < SC_WORDS; i++) { u.w[i] = r1[i]; v.w[i] = i < HG_TW ? t1[i] : 0u; }
    return neg;
}
// The 45 signed 5-bit digits of a multiplier m < 2^223 (values -16..15, zero allowed): digit k = window k of m + HALF_BIAS, minus 16,
// with HALF_BIAS = sum_k 16 * 32^k (no carries to propagate: one addition, then plain windows).  m + HALF_BIAS < 2^225.
#define HALF_WINDOWS 45
#define GOLD_CONST_HALF_BIAS { 0x21084210u, 0x08421084u, 0x42108421u, 0x10842108u, 0x84210842u, 0x21084210u, 0x08421084u, 0x00000001u }
GD void sc_half_bias(sc &out, const sc &m) {
    const uint32_t bias[8] = GOLD_CONST_HALF_BIAS;
    uint64_t chain = 0;
#pragma unroll
    for (int i = 0; i < SC_WORDS; i++) {
        chain = chain + m.w[i] + (i < 8 ? bias[i] : 0u);
        out.w[i] = (uint32_t)chain;
        chain >>= 32;
    }
}
