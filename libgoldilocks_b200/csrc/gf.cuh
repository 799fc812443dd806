// gf.cuh -- GF(p), p = 2^448 - 2^224 - 1, one field element per GPU lane.
//
// Replaces the reference's field layer (src/f_field.h:66-84 API, src/f_generic.c, src/f_arithmetic.c
// and the per-arch src/arch_*/f_impl.{c,h}) with a per-thread representation built for the sm_100a
// integer pipe:
//
//   * 16 limbs x 28 bits in 32-bit registers (radix 2^28, limb 8 sits at phi = 2^224).
//   * every 32x32->64 multiply-accumulate is one IMAD.WIDE(.U32) SASS instruction with a 64-bit
//     register-pair accumulator; there are no carry flags in the multiplier (measured on B200:
//     IMAD.WIDE.U32 issues at 31.5 lanes/clk/SM, the carry-chained .X form only at 25.5 -- see
//     profiles/r01_imad_peak.json -- which is why the unsaturated radix was chosen over 14x32).
//   * multiplication is Karatsuba over phi with the Solinas wrap phi^2 = phi + 1 (192 MACs),
//     squaring uses the symmetric half-products (108 MACs).
//
// Limb-bound discipline (checked on the host by tests/hostsim with GF_CHECK_BOUNDS):
//   TIGHT  limbs <= 2^28 + 2^11      output of mul/sqr/mulw/weak_reduce/sub/deserialize
//   LOOSE  limbs <  GF_LOOSE_MAX     e.g. sum of two TIGHT values; legal multiplier input
// Values are only defined mod p; canonical form is produced by gf_strong_reduce.
//
// All functions are __host__ __device__ so the very same source can be exercised on the CPU by the
// test-only host simulator (tests/hostsim); the product library only ever runs them on the GPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define GD __host__ __device__ __forceinline__
#define GDM __host__ __device__ __forceinline__ /* member functions */
#else
#define GD static inline
#define GDM inline
#endif

#define GF_NLIMBS 16
#define GF_MASK 0x0fffffffu
#define GF_TIGHT_MAX ((1u << 28) + (1u << 11))
#define GF_LOOSE_MAX 0x26000000u /* 1.1875 * 2^29 : 39 * B^2 + carry stays below 2^64 */

#ifdef GF_CHECK_BOUNDS
#include <assert.h>
#include <stdio.h>
#include <stdlib.h>
#define GF_ASSERT_LOOSE(x) do { for (int i_ = 0; i_ < 16; i_++) if ((x).v[i_] >= GF_LOOSE_MAX) { \
    fprintf(stderr, "gf bound violation (loose) %s:%d limb %d = %08x\n", __FILE__, __LINE__, i_, (x).v[i_]); abort(); } } while (0)
#define GF_ASSERT_TIGHT(x) do { for (int i_ = 0; i_ < 16; i_++) if ((x).v[i_] > GF_TIGHT_MAX) { \
    fprintf(stderr, "gf bound violation (tight) %s:%d limb %d = %08x\n", __FILE__, __LINE__, i_, (x).v[i_]); abort(); } } while (0)
#else
#define GF_ASSERT_LOOSE(x) do { } while (0)
#define GF_ASSERT_TIGHT(x) do { } while (0)
#endif

// Multiplier-call counters for the host simulator (tools/count_ops.py -> profiles/executed_ops.json: the executed
// IMAD.WIDE of every entry point = 193 x multiplications + 110 x squarings + 16 x word multiplications).  Host only.
#if defined(GF_COUNT_OPS) && !defined(__CUDA_ARCH__)
#include <atomic>
inline std::atomic<unsigned long long> &gf_op_counter(int k) { static std::atomic<unsigned long long> c[3]; return c[k]; }
#define GF_COUNT(k) gf_op_counter(k).fetch_add(1, std::memory_order_relaxed)
#else
#define GF_COUNT(k) do { } while (0)
#endif

// A zero the compiler cannot see through (constant bank, never written).  Adding it as a third
// operand keeps an addition a three-input IADD3 on the ALU pipe: ptxas otherwise turns about half of
// the two-input adds and register moves around the multiplier into IMAD.IADD / IMAD.MOV, which
// compete with IMAD.WIDE for the multiply pipe (profiles/r01_verify_finish_calls.txt).
#if defined(__CUDACC__)
static __constant__ uint32_t gf_opaque_zero;
#endif
#if defined(__CUDA_ARCH__)
#define GF_Z gf_opaque_zero
#else
#define GF_Z 0u
#endif
#define GF_ZS GF_Z /* measured: +2% on verify / X448 / comb (tools/expbench.sh, B200) */

struct alignas(16) gf { uint32_t v[GF_NLIMBS]; }; /* 16-byte aligned: table rows move as 128-bit loads */

typedef uint32_t gmask_t; /* all-ones / zero, like the reference's mask_t (word.h:263-278) */

GD void gf_set_zero(gf &a) {
#pragma unroll
    for (int i = 0; i < 16; i++) a.v[i] = 0;
}
GD void gf_set_ui(gf &a, uint32_t w) { /* w < 2^28 */
    gf_set_zero(a);
    a.v[0] = w;
}
GD void gf_copy(gf &o, const gf &a) {
#pragma unroll
    for (int i = 0; i < 16; i++) o.v[i] = a.v[i];
}

// Whole-element load from a 16-byte aligned table row: four 128-bit loads on the device.
// RO = true only for tables that are never written while the kernel runs (fixed-base tables): it
// takes the non-coherent path (LDG.CONSTANT).  Per-lane tables built by the same kernel use RO = false.
template <bool RO>
GD void gf_ld(gf &o, const gf *src) {
#if defined(__CUDA_ARCH__)
    const uint4 *p = reinterpret_cast<const uint4 *>(src);
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const uint4 x = RO ? __ldg(p + q) : p[q];
        o.v[4 * q] = x.x; o.v[4 * q + 1] = x.y; o.v[4 * q + 2] = x.z; o.v[4 * q + 3] = x.w;
    }
#else
    for (int i = 0; i < 16; i++) o.v[i] = src->v[i];
#endif
}
// o |= row & mask (masked full-row scan step of the constant-time lookups)
template <bool RO>
GD void gf_ld_or_masked(gf &o, const gf *src, uint32_t m) {
    gf t;
    gf_ld<RO>(t, src);
#pragma unroll
    for (int i = 0; i < 16; i++) o.v[i] |= t.v[i] & m;
}

// Carry-propagate once: TIGHT output for any input with limbs < 2^32.
// (reference arch_32/f_impl.h:32-39 gf_weak_reduce)
GD void gf_weak_reduce(gf &a) {
    uint32_t top = a.v[15] >> 28;
    uint32_t c[16];
#pragma unroll
    for (int i = 0; i < 15; i++) c[i] = a.v[i] >> 28;
#pragma unroll
    for (int i = 15; i > 0; i--) a.v[i] = (a.v[i] & GF_MASK) + c[i - 1];
    a.v[0] = (a.v[0] & GF_MASK) + top;
    a.v[8] += top;
}

// o = a + b, no reduction: TIGHT + TIGHT -> LOOSE.  (reference gf_add_nr / gf_add_RAW)
GD void gf_add_nr(gf &o, const gf &a, const gf &b) {
#pragma unroll
    for (int i = 0; i < 16; i++) o.v[i] = a.v[i] + b.v[i] + GF_ZS; /* three-input: stays an IADD3 (ALU pipe) */
}
// o = a + b, TIGHT output for any LOOSE inputs.  (reference f_generic.c:114-117 gf_add)
GD void gf_add(gf &o, const gf &a, const gf &b) {
    gf_add_nr(o, a, b);
    gf_weak_reduce(o);
}
// o = a - b, TIGHT output; b may be LOOSE (bias 4p keeps every limb non-negative).
// (reference f_generic.c:107-111 gf_sub, field.h:40-54 gf_sub_nr/gf_subx_nr)
GD void gf_sub(gf &o, const gf &a, const gf &b) {
    const uint32_t co1 = 4u * GF_MASK, co2 = co1 - 4u;
#pragma unroll
    for (int i = 0; i < 16; i++) o.v[i] = a.v[i] - b.v[i] + (i == 8 ? co2 : co1);
    gf_weak_reduce(o);
}
GD void gf_neg(gf &o, const gf &a) {
    const uint32_t co1 = 4u * GF_MASK, co2 = co1 - 4u;
#pragma unroll
    for (int i = 0; i < 16; i++) o.v[i] = (i == 8 ? co2 : co1) - a.v[i];
    gf_weak_reduce(o);
}

// Branch-free select / negate / swap (reference field.h:66-83, constant_time.h:134-362).
GD void gf_cond_sel(gf &o, const gf &y, const gf &z, gmask_t is_z) {
#pragma unroll
    for (int i = 0; i < 16; i++) o.v[i] = (y.v[i] & ~is_z) | (z.v[i] & is_z);
}
GD void gf_cond_neg(gf &x, gmask_t neg) {
    gf y;
    gf_neg(y, x);
    gf_cond_sel(x, x, y, neg);
}
GD void gf_cond_swap(gf &x, gf &y, gmask_t swap) {
#pragma unroll
    for (int i = 0; i < 16; i++) {
        uint32_t s = (x.v[i] ^ y.v[i]) & swap;
        x.v[i] ^= s;
        y.v[i] ^= s;
    }
}

// Final wrap of the two carry chains of a product (shared by mul/sqr/mulw).
GD void gf_fold_top(gf &c, uint64_t acc0, uint64_t acc1) {
    // acc0 = carry out of column 7 (goes to column 8); acc1 = carry out of column 15 (phi^2 = phi + 1)
    acc0 += acc1;
    acc0 += c.v[8];
    acc1 += c.v[0];
    c.v[8] = (uint32_t)acc0 & GF_MASK;
    c.v[0] = (uint32_t)acc1 & GF_MASK;
    c.v[9] += (uint32_t)(acc0 >> 28);
    c.v[1] += (uint32_t)(acc1 >> 28);
}

// c = a * b mod p.  LOOSE inputs, TIGHT output.  192 IMAD.WIDE.
// With a = a0 + a1*phi, b = b0 + b1*phi, P0 = a0*b0, P1 = a1*b1, PM = (a0+a1)(b0+b1):
//   a*b = (P0 + P1) + (PM - P0)*phi   (phi^2 = phi + 1), and columns 8..14 of each half wrap the same way:
//   low[j]  = P0[j] + P1[j] + PM[j+8] - P0[j+8]
//   high[j] = PM[j] - P0[j] + P1[j+8] + PM[j+8]
// (same identity as the reference's arch_32/f_impl.c:15-69; the subtracted P0[j+8] terms are
//  accumulated with a signed IMAD.WIDE on a pre-negated operand so they cost no extra instruction.)
GD void gf_mul_body(gf &c, const gf &a, const gf &b) {
    GF_COUNT(0);
    GF_ASSERT_LOOSE(a);
    GF_ASSERT_LOOSE(b);
    uint32_t aa[8], bb[8];
    int32_t nb[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        aa[i] = a.v[i] + a.v[i + 8] + GF_ZS;
        bb[i] = b.v[i] + b.v[i + 8] + GF_ZS;
        nb[i] = (int32_t)(GF_ZS - b.v[i]);
    }
    uint64_t acc0 = 0, acc1 = 0;
    gf r;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        uint64_t s = 0; /* P0[j] */
#pragma unroll
        for (int i = 0; i <= j; i++) {
            s += (uint64_t)a.v[j - i] * b.v[i];
            acc1 += (uint64_t)aa[j - i] * bb[i];            /* PM[j]   */
            acc0 += (uint64_t)a.v[8 + j - i] * b.v[8 + i];  /* P1[j]   */
        }
        uint64_t t = 0; /* PM[j+8] */
#pragma unroll
        for (int i = j + 1; i < 8; i++) {
            acc0 = (uint64_t)((int64_t)acc0 + (int64_t)(int32_t)a.v[8 + j - i] * (int64_t)nb[i]); /* -P0[j+8] */
            t += (uint64_t)aa[8 + j - i] * bb[i];
            acc1 += (uint64_t)a.v[16 + j - i] * b.v[8 + i]; /* P1[j+8] */
        }
        /* all four 64-bit combinations together, after the products: ptxas then emits three-input
         * IADD3 / IADD3.X pairs on the ALU pipe instead of IMAD.X on the multiply pipe */
        acc0 = acc0 + s + t;
        acc1 = acc1 + t - s;
        r.v[j] = (uint32_t)acc0 & GF_MASK;
        r.v[j + 8] = (uint32_t)acc1 & GF_MASK;
        acc0 >>= 28;
        acc1 >>= 28;
    }
    gf_fold_top(r, acc0, acc1);
    gf_copy(c, r);
}

// Half-size squaring column k of x[0..7]: sum_{i+l=k} x_i x_l, using the doubled operand x2 = 2x.
#define GF_SQR_COL(ACC, X, X2, K)                                                        \
    do {                                                                                 \
        _Pragma("unroll") for (int i_ = 0; i_ < 8; i_++) {                               \
            int l_ = (K) - i_;                                                           \
            if (l_ >= 0 && l_ < 8 && i_ < l_) ACC += (uint64_t)(X2)[i_] * (X)[l_];       \
            if (l_ == i_) ACC += (uint64_t)(X)[i_] * (X)[i_];                            \
        }                                                                                \
    } while (0)
#define GF_SQR_COL_NEG(ACC, NX, X2, X, K)                                                \
    do {                                                                                 \
        _Pragma("unroll") for (int i_ = 0; i_ < 8; i_++) {                               \
            int l_ = (K) - i_;                                                           \
            if (l_ >= 0 && l_ < 8 && i_ < l_) ACC = (uint64_t)((int64_t)ACC + (int64_t)(int32_t)(X2)[i_] * (int64_t)(NX)[l_]); \
            if (l_ == i_) ACC = (uint64_t)((int64_t)ACC + (int64_t)(int32_t)(X)[i_] * (int64_t)(NX)[i_]); \
        }                                                                                \
    } while (0)

// c = a^2 mod p.  LOOSE input, TIGHT output.  108 IMAD.WIDE (the reference's arch_32 has no
// dedicated squaring, arch_32/f_impl.c:98-100; arch_ref64/f_impl.c:151-301 does).
GD void gf_sqr_body(gf &c, const gf &a) {
    GF_COUNT(1);
    GF_ASSERT_LOOSE(a);
    uint32_t lo[8], hi[8], aa[8], lo2[8], hi2[8], aa2[8];
    int32_t nlo[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        lo[i] = a.v[i];
        hi[i] = a.v[i + 8];
        aa[i] = lo[i] + hi[i] + GF_ZS;
        lo2[i] = lo[i] + lo[i] + GF_ZS;
        hi2[i] = hi[i] + hi[i] + GF_ZS;
        aa2[i] = aa[i] + aa[i] + GF_ZS;
        nlo[i] = (int32_t)(GF_ZS - lo[i]);
    }
    uint64_t acc0 = 0, acc1 = 0;
    gf r;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        uint64_t s = 0;
        GF_SQR_COL(s, lo, lo2, j);        /* P0[j] */
        GF_SQR_COL(acc1, aa, aa2, j);     /* PM[j] */
        GF_SQR_COL(acc0, hi, hi2, j);     /* P1[j] */
        uint64_t t = 0;
        if (j < 7) {
            GF_SQR_COL(t, aa, aa2, j + 8);            /* PM[j+8] */
            GF_SQR_COL_NEG(acc0, nlo, lo2, lo, j + 8); /* -P0[j+8] */
            GF_SQR_COL(acc1, hi, hi2, j + 8);         /* P1[j+8] */
        }
        acc0 = acc0 + s + t;
        acc1 = acc1 + t - s;
        r.v[j] = (uint32_t)acc0 & GF_MASK;
        r.v[j + 8] = (uint32_t)acc1 & GF_MASK;
        acc0 >>= 28;
        acc1 >>= 28;
    }
    gf_fold_top(r, acc0, acc1);
    gf_copy(c, r);
}

// Call shape of the two heavy operations.  On the device they are real (non-inlined) functions taking
// and returning field elements BY VALUE: ptxas passes the 16-limb structs in registers (no local
// memory, checked in SASS), so a whole scalar multiplication holds exactly one multiplier body and one
// squaring body (~8 KB of SASS) instead of dozens of inlined copies.  Fully inlined, the verify loop
// was ~140 KB of code and spent 4 of every 8.5 cycles per issued instruction stalled on instruction
// fetch (ncu `no_instruction`, profiles/r01_verify_finish_inlined.txt); the instruction cache is
// 32 KB (L1.5).  Tiny kernels (one or two multiplications) define GF_INLINE_MUL and keep the inline form.
#if defined(__CUDA_ARCH__) && !defined(GF_INLINE_MUL)
static __device__ __noinline__ gf gf_mul_fn(gf a, gf b) { gf c; gf_mul_body(c, a, b); return c; }
static __device__ __noinline__ gf gf_sqr_fn(gf a) { gf c; gf_sqr_body(c, a); return c; }
static __device__ __noinline__ gf gf_sqrn_fn(gf a, int n) { /* n >= 1 squarings, loop inside the callee */
#pragma unroll 1
    for (int i = 0; i < n; i++) gf_sqr_body(a, a);
    return a;
}
GD void gf_mul(gf &c, const gf &a, const gf &b) { c = gf_mul_fn(a, b); }
GD void gf_sqr(gf &c, const gf &a) { c = gf_sqr_fn(a); }
GD void gf_sqrn(gf &y, const gf &x, int n) { y = gf_sqrn_fn(x, n); }
#else
GD void gf_mul(gf &c, const gf &a, const gf &b) { gf_mul_body(c, a, b); }
GD void gf_sqr(gf &c, const gf &a) { gf_sqr_body(c, a); }
GD void gf_sqrn(gf &y, const gf &x, int n) { /* reference field.h:19-38 */
    gf_sqr(y, x);
    for (int i = 1; i < n; i++) gf_sqr(y, y);
}
#endif

// c = a * w for a small unsigned w < 2^28.  LOOSE input, TIGHT output.  16 IMAD.WIDE.
// (reference arch_32/f_impl.c:71-96 gf_mulw_unsigned)
GD void gf_mulw(gf &c, const gf &a, uint32_t w) {
    GF_COUNT(2);
    GF_ASSERT_LOOSE(a);
    uint64_t acc0 = 0, acc1 = 0;
    gf r;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        acc0 += (uint64_t)a.v[j] * w;
        acc1 += (uint64_t)a.v[j + 8] * w;
        r.v[j] = (uint32_t)acc0 & GF_MASK;
        r.v[j + 8] = (uint32_t)acc1 & GF_MASK;
        acc0 >>= 28;
        acc1 >>= 28;
    }
    gf_fold_top(r, acc0, acc1);
    gf_copy(c, r);
}
// c = a * w for a signed compile-time-ish w (reference field.h:57-64 gf_mulw)
GD void gf_mulw_signed(gf &c, const gf &a, int32_t w) {
    if (w >= 0) {
        gf_mulw(c, a, (uint32_t)w);
    } else {
        gf_mulw(c, a, (uint32_t)(-w));
        gf_neg(c, c);
    }
}

// Canonical form: 0 <= value < p, every limb < 2^28.  (reference f_generic.c:71-104)
GD void gf_strong_reduce(gf &a) {
    gf_weak_reduce(a); /* value < 2p now */
    /* subtract p = (2^28-1 in every limb, 2^28-2 in limb 8) with a signed borrow chain */
    int32_t borrow = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        int32_t t = (int32_t)a.v[i] - (int32_t)(i == 8 ? GF_MASK - 1 : GF_MASK) + borrow; /* > -2^29 */
        a.v[i] = (uint32_t)t & GF_MASK;
        borrow = t >> 28; /* arithmetic: 0 or -1 (or +1 transiently when the limb had a carry bit) */
    }
    /* borrow == 0: value was >= p, difference is the answer.  borrow == -1: add p back. */
    uint32_t addback = (uint32_t)borrow; /* 0 or 0xffffffff */
    uint32_t carry = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        uint32_t t = a.v[i] + ((i == 8 ? GF_MASK - 1 : GF_MASK) & addback) + carry;
        a.v[i] = t & GF_MASK;
        carry = t >> 28;
    }
}

// 448-bit canonical value <-> 14 little-endian 32-bit words (the 56-byte wire form).
GD void gf_to_words(uint32_t w[14], const gf &a_in) { /* reference f_generic.c:19-37 gf_serialize */
    gf a;
    gf_copy(a, a_in);
    gf_strong_reduce(a);
#pragma unroll
    for (int k = 0; k < 14; k++) {
        /* word k covers bits [32k, 32k+32) ; limb i covers bits [28i, 28i+28) */
        const int lo_limb = (32 * k) / 28, sh = (32 * k) % 28;
        uint32_t x = a.v[lo_limb] >> sh;
        x |= a.v[lo_limb + 1] << (28 - sh);
        if (28 - sh + 28 < 32) x |= a.v[lo_limb + 2] << (56 - sh);
        w[k] = x;
    }
}
// Returns all-ones iff the 448-bit value is < p.  The element is always written (value mod p
// semantics: limbs hold the raw 448-bit value, which is < 2^448 < 2p).  (f_generic.c:48-68)
GD gmask_t gf_from_words(gf &a, const uint32_t w[14]) {
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const int lo_word = (28 * i) / 32, sh = (28 * i) % 32;
        uint32_t x = w[lo_word] >> sh;
        if (sh > 4 && lo_word + 1 < 14) x |= w[lo_word + 1] << (32 - sh);
        a.v[i] = x & GF_MASK;
    }
    /* value < p  <=>  subtracting p borrows */
    int32_t borrow = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        int32_t t = (int32_t)a.v[i] - (int32_t)(i == 8 ? GF_MASK - 1 : GF_MASK) + borrow;
        borrow = t >> 28;
    }
    return (gmask_t)borrow; /* -1 when < p */
}

GD gmask_t gf_is_zero(const gf &a_in) {
    gf a;
    gf_copy(a, a_in);
    gf_strong_reduce(a);
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) r |= a.v[i];
    return (gmask_t)(((uint64_t)r - 1) >> 32);
}
GD gmask_t gf_eq(const gf &a, const gf &b) { /* reference f_generic.c:120-131 */
    gf c;
    gf_sub(c, a, b);
    return gf_is_zero(c);
}
GD gmask_t gf_lobit(const gf &a_in) { /* reference f_generic.c:40-45 */
    gf a;
    gf_copy(a, a_in);
    gf_strong_reduce(a);
    return (gmask_t)(-(int32_t)(a.v[0] & 1));
}

// a = x^((p-3)/4) = +-1/sqrt(x); returns all-ones iff a^2 * x == 1 (so x = 0 and non-squares fail).
// Same exponent as the reference's addition chain (f_arithmetic.c:14-47: 446 S + 13 M) walked as a
// 12-step table so the GPU code holds one squaring loop and one multiply instead of 26 inlined bodies:
//   step: square `n` times, multiply by x or by the single saved power, optionally save.
// The schedule lives in two 60-bit immediates (ten bits per step), not in an array: a local array read in a loop
// is a dynamically addressed local load, which costs a memory access per step and which a static constant-time
// audit of the SASS (tools/ct_audit.py) cannot tell from secret-indexed data.
GD uint32_t gf_isr_step(int s) { /* n (bits 0-7), multiply-by-x flag (bit 8), save-after flag (bit 9) */
    const uint64_t lo = 0x4060980c03c0501ull, hi = 0x37d019be2509612ull;
    return (uint32_t)((s < 6 ? lo >> (10 * s) : hi >> (10 * (s - 6))) & 0x3ff);
}
GD gmask_t gf_isr(gf &a, const gf &x) {
    gf cur, saved;
    gf_copy(cur, x);
    gf_copy(saved, x);
    /* steps: {1|X, 1|X|S, 3, 3|S, 9|S, 1|X, 18|S, 37, 37|S, 111|S, 1|X, 223}  (X = multiply by x, S = save afterwards) */
#pragma unroll 1
    for (int s = 0; s < 12; s++) {
        const uint32_t step = gf_isr_step(s);
        const int n = (int)(step & 0xff);
        const gmask_t by_x = (step & 0x100) ? ~0u : 0u;
        gf_sqrn(cur, cur, n);
        gf m;
        gf_cond_sel(m, saved, x, by_x); /* schedule is public: not secret dependent */
        gf_mul(cur, cur, m);
        if (step & 0x200) gf_copy(saved, cur);
    }
    gf t0, t1;
    gf_sqr(t0, cur);
    gf_mul(t1, t0, x);
    gf_copy(a, cur);
    gf one;
    gf_set_ui(one, 1);
    return gf_eq(t1, one);
}

// y = 1/x (0 for x = 0).  (reference goldilocks.c:69-80: isr(x^2)^2 * x)
GD void gf_invert(gf &y, const gf &x) {
    gf t1, t2;
    gf_sqr(t1, x);
    (void)gf_isr(t2, t1);
    gf_sqr(t1, t2);
    gf_mul(t2, t1, x);
    gf_copy(y, t2);
}
