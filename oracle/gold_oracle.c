/*
 * oracle/gold_oracle.c -- TEST INFRASTRUCTURE ONLY.  Never linked, loaded or called by the product
 * library (libgoldilocks_b200.so); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg may use it, and only as the checker.
 *
 * A CPU restatement, in plain C, of the algorithms of otrv4/libgoldilocks' hot path.  It is NOT a
 * copy of the reference: arithmetic is done on fully reduced integers (7 x 64-bit words, always
 * canonical) instead of the reference's lazily reduced limb vectors, so every value it produces is
 * the canonical representative the reference would serialize.  Each function cites the reference
 * file:line whose behaviour it follows ("ref:" comments, paths relative to the reference root).
 *
 * Parity pinned: tests/test_oracle.py checks this file against every golden vector the reference's
 * own tests hold for the path (RFC 7748, RFC 8032 x 11, 16 base multiples, 16 Elligator pairs;
 * tests/golden/reference_vectors.json) and, in the build container, against the unmodified reference
 * compiled as oracle/_ref on random and edge inputs for every exported function.
 *
 * Exports the `*_batch` entry points of include/goldilocks_b200.h (same names and argument order)
 * so the tests drive it through the same ctypes binding as the CUDA library.
 */
#define _GNU_SOURCE 1
#include <pthread.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define EXPORT __attribute__((visibility("default")))
typedef unsigned __int128 u128;

/* =================================================================================================
 * Multi-word helpers (little-endian arrays of 64-bit words)
 * ================================================================================================= */
static uint64_t mw_add(uint64_t *r, const uint64_t *a, const uint64_t *b, int n) { /* r = a + b, returns carry */
    u128 c = 0;
    for (int i = 0; i < n; i++) { c += (u128)a[i] + b[i]; r[i] = (uint64_t)c; c >>= 64; }
    return (uint64_t)c;
}
static uint64_t mw_sub(uint64_t *r, const uint64_t *a, const uint64_t *b, int n) { /* r = a - b, returns borrow (0/1) */
    uint64_t br = 0;
    for (int i = 0; i < n; i++) {
        u128 d = (u128)a[i] - b[i] - br;
        r[i] = (uint64_t)d;
        br = (uint64_t)(d >> 64) & 1;
    }
    return br;
}
static int mw_cmp(const uint64_t *a, const uint64_t *b, int n) {
    for (int i = n - 1; i >= 0; i--) if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1;
    return 0;
}
static void mw_mul(uint64_t *r, const uint64_t *a, int na, const uint64_t *b, int nb) { /* r[na+nb] = a * b */
    memset(r, 0, sizeof(uint64_t) * (size_t)(na + nb));
    for (int i = 0; i < na; i++) {
        u128 c = 0;
        for (int j = 0; j < nb; j++) { c += (u128)a[i] * b[j] + r[i + j]; r[i + j] = (uint64_t)c; c >>= 64; }
        r[i + nb] = (uint64_t)c;
    }
}
static int mw_is_zero(const uint64_t *a, int n) { uint64_t x = 0; for (int i = 0; i < n; i++) x |= a[i]; return x == 0; }
/* r[n] = a[n] >> s (s may exceed 64) */
static void mw_shr(uint64_t *r, const uint64_t *a, int n, int s) {
    int ws = s / 64, bs = s % 64;
    for (int i = 0; i < n; i++) {
        uint64_t lo = (i + ws < n) ? a[i + ws] : 0, hi = (i + ws + 1 < n) ? a[i + ws + 1] : 0;
        r[i] = bs ? (lo >> bs) | (hi << (64 - bs)) : lo;
    }
}
static void mw_shl(uint64_t *r, const uint64_t *a, int n, int s) { /* r[n] = a[n] << s, truncating */
    int ws = s / 64, bs = s % 64;
    for (int i = n - 1; i >= 0; i--) {
        uint64_t hi = (i - ws >= 0) ? a[i - ws] : 0, lo = (i - ws - 1 >= 0) ? a[i - ws - 1] : 0;
        r[i] = bs ? (hi << bs) | (lo >> (64 - bs)) : hi;
    }
}

/* =================================================================================================
 * GF(p), p = 2^448 - 2^224 - 1       ref: src/f_field.h:66-84 (API), src/f_generic.c:14-16 (MODULUS)
 * ================================================================================================= */
typedef struct { uint64_t w[7]; } fe;
static const fe FE_P = {{~0ull, ~0ull, ~0ull, 0xfffffffeffffffffull, ~0ull, ~0ull, ~0ull}};
static const fe FE_ZERO = {{0}}, FE_ONE = {{1}};

static void fe_cond_sub_p(uint64_t *t8) { /* t8 (8 words) < 2p -> t8 mod p */
    uint64_t p8[8], d[8];
    memcpy(p8, FE_P.w, 56); p8[7] = 0;
    if (mw_cmp(t8, p8, 8) >= 0) { mw_sub(d, t8, p8, 8); memcpy(t8, d, 64); }
}
/* any 8-word value -> canonical: fold bits >= 448 with 2^448 = 2^224 + 1 until they vanish */
static void fe_from8(fe *r, uint64_t *t) {
    for (;;) {
        uint64_t hi = t[7];
        if (!hi) break;
        uint64_t add[8] = {hi, 0, 0, hi << 32, hi >> 32, 0, 0, 0};
        t[7] = 0;
        mw_add(t, t, add, 8);
    }
    fe_cond_sub_p(t);
    memcpy(r->w, t, 56);
}
static void fe_add(fe *r, const fe *a, const fe *b) { /* ref: f_generic.c:114-117 */
    uint64_t t[8];
    t[7] = mw_add(t, a->w, b->w, 7);
    fe_from8(r, t);
}
static void fe_sub(fe *r, const fe *a, const fe *b) { /* ref: f_generic.c:107-111 */
    uint64_t t[7];
    if (mw_sub(t, a->w, b->w, 7)) mw_add(t, t, FE_P.w, 7);
    memcpy(r->w, t, 56);
}
static void fe_neg(fe *r, const fe *a) { fe_sub(r, &FE_ZERO, a); }
/* 14-word product -> canonical (the Solinas reduction the reference interleaves with its Karatsuba,
 * ref: arch_ref64/f_impl.c:7-149; here done on the full product) */
static void fe_reduce_wide(fe *r, const uint64_t *x) {
    uint64_t L[8], H[8], Hhi[8], S[8], T[8];
    memcpy(L, x, 56); L[7] = 0;
    memcpy(H, x + 7, 56); H[7] = 0;
    mw_shr(Hhi, H, 8, 224);                       /* H >> 224 */
    uint64_t Hlo[8];
    memcpy(Hlo, H, 64); Hlo[3] &= 0xffffffffull; Hlo[4] = Hlo[5] = Hlo[6] = Hlo[7] = 0;
    mw_add(S, Hlo, Hhi, 8);                       /* S = H_lo + H_hi < 2^225 */
    mw_shl(S, S, 8, 224);                         /* S * 2^224 < 2^449 */
    mw_add(T, L, H, 8);
    mw_add(T, T, Hhi, 8);
    mw_add(T, T, S, 8);                           /* x = L + H + H_hi + (H_lo + H_hi) 2^224 (mod p) < 2^451 */
    fe_from8(r, T);
}
static void fe_mul(fe *r, const fe *a, const fe *b) { /* ref: f_field.h:66 gf_mul */
    uint64_t x[14];
    mw_mul(x, a->w, 7, b->w, 7);
    fe_reduce_wide(r, x);
}
static void fe_sqr(fe *r, const fe *a) { fe_mul(r, a, a); } /* ref: f_field.h:68 gf_sqr */
static void fe_mulw(fe *r, const fe *a, int64_t w) { /* ref: field.h:57-64 gf_mulw (signed small constant) */
    fe c = {{(uint64_t)(w < 0 ? -w : w)}};
    fe_mul(r, a, &c);
    if (w < 0) fe_neg(r, r);
}
static int fe_is_zero(const fe *a) { return mw_is_zero(a->w, 7); }
static int fe_eq(const fe *a, const fe *b) { return mw_cmp(a->w, b->w, 7) == 0; } /* ref: f_generic.c:120-131 */
static int fe_lobit(const fe *a) { return (int)(a->w[0] & 1); }                     /* ref: f_generic.c:40-45 */
static void fe_cond_neg(fe *a, int neg) { if (neg) fe_neg(a, a); }                  /* ref: field.h:72-77 */
/* 56 little-endian bytes -> element; returns 1 iff the encoded value is < p.  The element is always
 * set (value mod p), like the reference whose callers may ignore the result.  ref: f_generic.c:48-68 */
static int fe_deserialize(fe *r, const uint8_t *ser) {
    uint64_t t[8];
    memcpy(t, ser, 56); t[7] = 0;
    int ok = mw_cmp(t, FE_P.w, 7) < 0;
    fe_from8(r, t);
    return ok;
}
static void fe_serialize(uint8_t *ser, const fe *a) { memcpy(ser, a->w, 56); } /* ref: f_generic.c:19-37 */

/* r = x^((p-3)/4) = +-1/sqrt(x); returns 1 iff r^2 x == 1.  ref: f_arithmetic.c:14-47 (same exponent,
 * (p-3)/4 = 2^446 - 2^222 - 1, walked here by plain square-and-multiply from the top bit). */
static int fe_isr(fe *r, const fe *x) {
    fe acc = FE_ONE, t;
    for (int bit = 445; bit >= 0; bit--) {
        fe_sqr(&acc, &acc);
        if (bit != 222) fe_mul(&acc, &acc, x);   /* exponent bits: all ones except bit 222 */
    }
    *r = acc;
    fe_sqr(&t, &acc);
    fe_mul(&t, &t, x);
    return fe_eq(&t, &FE_ONE);
}
static void fe_invert(fe *y, const fe *x) { /* ref: goldilocks.c:69-80: isr(x^2)^2 * x */
    fe t1, t2;
    fe_sqr(&t1, x);
    (void)fe_isr(&t2, &t1);
    fe_sqr(&t1, &t2);
    fe_mul(y, &t1, x);
}

/* =================================================================================================
 * Scalars mod q                                           ref: src/scalar.c (sc_p at 18-19)
 * ================================================================================================= */
typedef struct { uint64_t w[7]; } scl;
static const scl SC_Q = {{0x2378c292ab5844f3ull, 0x216cc2728dc58f55ull, 0xc44edb49aed63690ull, 0xffffffff7cca23e9ull,
                          0xffffffffffffffffull, 0xffffffffffffffffull, 0x3fffffffffffffffull}};
/* c = 2^446 - q (224 bits) */
static const uint64_t SC_C[4] = {0xdc873d6d54a7bb0dull, 0xde933d8d723a70aaull, 0x3bb124b65129c96full, 0x000000008335dc16ull};

/* x[n] (n <= 16) -> x mod q, folding at bit 446 with 2^446 = c (mod q) */
static void sc_reduce_wide(scl *r, const uint64_t *x, int n) {
    uint64_t t[20], hi[20], prod[24];
    memset(t, 0, sizeof t);
    memcpy(t, x, 8 * (size_t)n);
    for (;;) {
        mw_shr(hi, t, 20, 446);
        if (mw_is_zero(hi, 20)) break;
        t[6] &= 0x3fffffffffffffffull;
        for (int i = 7; i < 20; i++) t[i] = 0;
        mw_mul(prod, hi, 16, SC_C, 4);
        mw_add(t, t, prod, 20);
    }
    if (mw_cmp(t, SC_Q.w, 7) >= 0) mw_sub(t, t, SC_Q.w, 7);
    memcpy(r->w, t, 56);
}
static void sc_add(scl *r, const scl *a, const scl *b) { /* ref: scalar.c:176-189 */
    uint64_t t[8];
    t[7] = mw_add(t, a->w, b->w, 7);
    sc_reduce_wide(r, t, 8);
}
static void sc_sub(scl *r, const scl *a, const scl *b) { /* ref: scalar.c:168-174 */
    uint64_t t[7];
    if (mw_sub(t, a->w, b->w, 7)) mw_add(t, t, SC_Q.w, 7);
    memcpy(r->w, t, 56);
    /* inputs may be >= q only through the raw ABI; keep the result canonical like sc_subx's single correction */
    if (mw_cmp(r->w, SC_Q.w, 7) >= 0) { uint64_t u[8]; memcpy(u, r->w, 56); u[7] = 0; sc_reduce_wide(r, u, 8); }
}
static void sc_mul(scl *r, const scl *a, const scl *b) { /* ref: scalar.c:93-100 (two Montgomery steps = plain product mod q) */
    uint64_t x[14];
    mw_mul(x, a->w, 7, b->w, 7);
    sc_reduce_wide(r, x, 14);
}
static void sc_halve(scl *r, const scl *a) { /* ref: scalar.c:316-332 */
    uint64_t t[8];
    memcpy(t, a->w, 56); t[7] = 0;
    if (t[0] & 1) t[7] = mw_add(t, t, SC_Q.w, 7);
    mw_shr(t, t, 8, 1);
    memcpy(r->w, t, 56);
}
/* little-endian bytes of any length -> value mod q.  ref: scalar.c:257-293 (Horner over 56-byte chunks) */
static void sc_decode_long(scl *r, const uint8_t *ser, size_t len) {
    scl acc = {{0}};
    if (len == 0) { *r = acc; return; }
    size_t i = len - (len % 56);
    if (i == len) i -= 56;
    uint64_t x[14];
    memset(x, 0, sizeof x);
    memcpy(x, ser + i, len - i);
    sc_reduce_wide(&acc, x, 7);
    while (i) {
        i -= 56;
        memcpy(x, ser + i, 56);          /* low 448 bits = this chunk */
        memcpy(x + 7, acc.w, 56);        /* acc * 2^448 */
        sc_reduce_wide(&acc, x, 14);
    }
    *r = acc;
}
static int sc_bit(const scl *s, unsigned bit) { return bit < 448 ? (int)((s->w[bit / 64] >> (bit % 64)) & 1) : 0; }
static unsigned sc_bits(const scl *s, unsigned pos, unsigned n) {
    unsigned v = 0;
    for (unsigned k = 0; k < n; k++) v |= (unsigned)sc_bit(s, pos + k) << k;
    return v;
}

/* =================================================================================================
 * Group: extended twisted Edwards, a = -1, d' = -39082       ref: src/goldilocks.c
 * ================================================================================================= */
#define EDWARDS_D (-39081)              /* ref: goldilocks.c:32 */
#define TWISTED_D (EDWARDS_D - 1)       /* ref: goldilocks.c:45 */
#define EFF_D 39082
#define SCALAR_BITS 446                 /* ref: point_448.h:27 */
#define COMBS_N 5
#define COMBS_T 5
#define COMBS_S 18                      /* ref: goldilocks.c:25-27 */
#define WINDOW 5                        /* ref: goldilocks.c:28 */
#define WNAF_FIXED_BITS 5
#define WNAF_VAR_BITS 3                 /* ref: goldilocks.c:29-30 */

typedef struct { fe x, y, z, t; } pt;
typedef struct { fe a, b, c; } niels;       /* y-x, y+x, 2 d' x y     ref: goldilocks.c:55 */
typedef struct { niels n; fe z; } pniels;   /* ref: goldilocks.c:56 */

static fe FACTOR;             /* ref: goldilocks.c:41-43 GOLDILOCKS_448_FACTOR, derived in oracle_init() */
static scl ADJUST;            /* ref: goldilocks.c:33-37 scalarmul adjustment = 2^450 - 1 mod q */
static pt BASE;
static niels COMB[COMBS_N << (COMBS_T - 1)];
static niels WNAF_BASE[1 << WNAF_FIXED_BITS];

static void pt_identity(pt *p) { p->x = FE_ZERO; p->y = FE_ONE; p->z = FE_ONE; p->t = FE_ZERO; } /* ref: goldilocks.c:83 */

/* ref: goldilocks.c:205-230 (add) / 178-203 (sub) */
static void pt_addsub(pt *p, const pt *q, const pt *r, int sub) {
    fe a, b, c, d;
    fe_sub(&b, &q->y, &q->x);
    if (!sub) { fe_sub(&c, &r->y, &r->x); fe_add(&d, &r->y, &r->x); }
    else      { fe_sub(&d, &r->y, &r->x); fe_add(&c, &r->y, &r->x); }
    fe_mul(&a, &c, &b);
    fe_add(&b, &q->y, &q->x);
    fe py, px;
    fe_mul(&py, &d, &b);
    fe_mul(&b, &r->t, &q->t);
    fe_mulw(&px, &b, 2 * EFF_D);
    fe_add(&b, &a, &py);
    fe_sub(&c, &py, &a);
    fe_mul(&a, &q->z, &r->z);
    fe_add(&a, &a, &a);
    if (!sub) { fe_add(&py, &a, &px); fe_sub(&a, &a, &px); }
    else      { fe_sub(&py, &a, &px); fe_add(&a, &a, &px); }
    fe_mul(&p->z, &a, &py);
    fe_mul(&p->x, &py, &c);
    fe_mul(&p->y, &a, &b);
    fe_mul(&p->t, &b, &c);
}
/* ref: goldilocks.c:232-254 point_double_internal */
static void pt_double(pt *p, const pt *q, int before_double) {
    fe a, b, c, d;
    fe_sqr(&c, &q->x);
    fe_sqr(&a, &q->y);
    fe_add(&d, &c, &a);
    fe t;
    fe_add(&t, &q->y, &q->x);
    fe_sqr(&b, &t);
    fe_sub(&b, &b, &d);
    fe_sub(&t, &a, &c);
    fe px;
    fe_sqr(&px, &q->z);
    fe pz;
    fe_add(&pz, &px, &px);
    fe_sub(&a, &pz, &t);
    fe_mul(&p->x, &a, &b);
    fe_mul(&p->z, &t, &a);
    fe_mul(&p->y, &t, &d);
    if (!before_double) fe_mul(&p->t, &b, &d);
}
static void pt_negate(pt *r, const pt *a) { fe_neg(&r->x, &a->x); r->y = a->y; r->z = a->z; fe_neg(&r->t, &a->t); } /* ref: goldilocks.c:260-268 */

static void niels_cond_neg(niels *n, int neg) { /* ref: goldilocks.c:271-278 */
    if (neg) { fe t = n->a; n->a = n->b; n->b = t; fe_neg(&n->c, &n->c); }
}
static void pt_to_pniels(pniels *b, const pt *a) { /* ref: goldilocks.c:280-288 */
    fe_sub(&b->n.a, &a->y, &a->x);
    fe_add(&b->n.b, &a->x, &a->y);
    fe_mulw(&b->n.c, &a->t, 2 * TWISTED_D);
    fe_add(&b->z, &a->z, &a->z);
}
static void pniels_to_pt(pt *e, const pniels *d) { /* ref: goldilocks.c:290-301 */
    fe eu, ey;
    fe_add(&eu, &d->n.b, &d->n.a);
    fe_sub(&ey, &d->n.b, &d->n.a);
    fe_mul(&e->t, &ey, &eu);
    fe_mul(&e->x, &d->z, &ey);
    fe_mul(&e->y, &d->z, &eu);
    fe_sqr(&e->z, &d->z);
}
static void niels_to_pt(pt *e, const niels *n) { /* ref: goldilocks.c:303-313 */
    fe_add(&e->y, &n->b, &n->a);
    fe_sub(&e->x, &n->b, &n->a);
    fe_mul(&e->t, &e->y, &e->x);
    e->z = FE_ONE;
}
/* ref: goldilocks.c:315-359 add_niels_to_pt / sub_niels_from_pt */
static void pt_addsub_niels(pt *d, const niels *e, int sub, int before_double) {
    fe a, b, c;
    fe_sub(&b, &d->y, &d->x);
    fe_mul(&a, sub ? &e->b : &e->a, &b);
    fe_add(&b, &d->x, &d->y);
    fe dy;
    fe_mul(&dy, sub ? &e->a : &e->b, &b);
    fe dx;
    fe_mul(&dx, &e->c, &d->t);
    fe_add(&c, &a, &dy);
    fe_sub(&b, &dy, &a);
    if (!sub) { fe_sub(&dy, &d->z, &dx); fe_add(&a, &dx, &d->z); }
    else      { fe_add(&dy, &d->z, &dx); fe_sub(&a, &d->z, &dx); }
    fe_mul(&d->z, &a, &dy);
    fe_mul(&d->x, &dy, &b);
    fe_mul(&d->y, &a, &c);
    if (!before_double) fe_mul(&d->t, &b, &c);
}
static void pt_addsub_pniels(pt *p, const pniels *pn, int sub, int before_double) { /* ref: goldilocks.c:361-380 */
    fe l0;
    fe_mul(&l0, &p->z, &pn->z);
    p->z = l0;
    pt_addsub_niels(p, &pn->n, sub, before_double);
}
static int pt_eq(const pt *p, const pt *q) { /* ref: goldilocks.c:644-653 */
    fe a, b;
    fe_mul(&a, &p->y, &q->x);
    fe_mul(&b, &q->y, &p->x);
    return fe_eq(&a, &b);
}
static int pt_valid(const pt *p) { /* ref: goldilocks.c:655-673 */
    fe a, b, c;
    fe_mul(&a, &p->x, &p->y);
    fe_mul(&b, &p->z, &p->t);
    int out = fe_eq(&a, &b);
    fe_sqr(&a, &p->x);
    fe_sqr(&b, &p->y);
    fe_sub(&a, &b, &a);
    fe_sqr(&b, &p->t);
    fe_mulw(&c, &b, TWISTED_D);
    fe_sqr(&b, &p->z);
    fe_add(&b, &b, &c);
    out &= fe_eq(&a, &b);
    out &= !fe_is_zero(&p->z);
    return out;
}

/* ---- decaf codec ------------------------------------------------------------------------------- */
static void pt_encode(uint8_t ser[56], const pt *p) { /* ref: goldilocks.c:98-140 deisogenize (toggles 0) + 136-140 */
    fe t1, t2, t3, t4, s;
    fe_add(&t1, &p->x, &p->t);
    fe_sub(&t2, &p->x, &p->t);
    fe_mul(&t3, &t1, &t2);
    fe_sqr(&t2, &p->x);
    fe_mul(&t1, &t2, &t3);
    fe_mulw(&t2, &t1, -1 - TWISTED_D);
    (void)fe_isr(&t1, &t2);
    fe_mul(&t2, &t1, &t3);
    fe_mul(&t4, &t2, &FACTOR);
    fe_cond_neg(&t2, fe_lobit(&t4));
    fe_mul(&t3, &t2, &p->z);
    fe_sub(&t3, &t3, &p->t);
    fe_mul(&t2, &t3, &p->x);
    fe_mulw(&t4, &t2, -1 - TWISTED_D);
    fe_mul(&s, &t4, &t1);
    fe_cond_neg(&s, fe_lobit(&s));
    fe_serialize(ser, &s);
}
static int pt_decode(pt *p, const uint8_t ser[56], int allow_identity) { /* ref: goldilocks.c:142-176 */
    fe s, s2, num, tmp, tmp2, ynum, isr, den;
    int succ = fe_deserialize(&s, ser);
    succ &= allow_identity | !fe_is_zero(&s);
    succ &= !fe_lobit(&s);
    fe_sqr(&s2, &s);
    fe_sub(&den, &FE_ONE, &s2);
    fe_add(&ynum, &FE_ONE, &s2);
    fe_mulw(&num, &s2, -4 * TWISTED_D);
    fe_sqr(&tmp, &den);
    fe_add(&num, &tmp, &num);
    fe_mul(&tmp2, &num, &tmp);
    succ &= fe_isr(&isr, &tmp2);
    fe_mul(&tmp, &isr, &den);
    fe_mul(&p->y, &tmp, &ynum);
    fe_mul(&tmp2, &tmp, &s);
    fe_add(&tmp2, &tmp2, &tmp2);
    fe_mul(&tmp, &tmp2, &isr);
    fe_mul(&p->x, &tmp, &num);
    fe_mul(&tmp, &tmp2, &FACTOR);
    fe_cond_neg(&p->x, fe_lobit(&tmp));
    p->z = FE_ONE;
    fe_mul(&p->t, &p->x, &p->y);
    return succ;
}

/* ---- RFC 8032 / RFC 7748 codecs ------------------------------------------------------------------- */
static void pt_encode_like_eddsa(uint8_t enc[57], const pt *q) { /* ref: goldilocks.c:905-946 */
    fe x, y, z, t, u;
    fe_sqr(&x, &q->x);
    fe_sqr(&t, &q->y);
    fe_add(&u, &x, &t);
    fe_add(&z, &q->y, &q->x);
    fe_sqr(&y, &z);
    fe_sub(&y, &y, &u);
    fe_sub(&z, &t, &x);
    fe_sqr(&x, &q->z);
    fe_add(&t, &x, &x);
    fe_sub(&t, &t, &z);
    fe_mul(&x, &t, &y);
    fe_mul(&y, &z, &u);
    fe_mul(&z, &u, &t);
    fe_invert(&z, &z);
    fe_mul(&t, &x, &z);
    fe_mul(&x, &y, &z);
    enc[56] = 0;
    fe_serialize(enc, &x);
    enc[56] |= fe_lobit(&t) ? 0x80 : 0;
}
static int pt_decode_like_eddsa(pt *p, const uint8_t enc[57]) { /* ref: goldilocks.c:949-1004 */
    fe a, b, c, d;
    int low = (enc[56] & 0x80) != 0;
    int succ = fe_deserialize(&p->y, enc);
    succ &= (enc[56] & 0x7f) == 0;
    fe_sqr(&p->x, &p->y);
    fe_sub(&p->z, &FE_ONE, &p->x);
    fe_mulw(&p->t, &p->x, EDWARDS_D);
    fe_sub(&p->t, &FE_ONE, &p->t);
    fe_mul(&p->x, &p->z, &p->t);
    succ &= fe_isr(&p->t, &p->x);
    fe_mul(&p->x, &p->t, &p->z);
    fe_cond_neg(&p->x, fe_lobit(&p->x) ^ low);
    p->z = FE_ONE;
    fe_sqr(&c, &p->x);
    fe_sqr(&a, &p->y);
    fe_add(&d, &c, &a);
    fe_add(&p->t, &p->y, &p->x);
    fe_sqr(&b, &p->t);
    fe_sub(&b, &b, &d);
    fe_sub(&p->t, &a, &c);
    fe_sqr(&p->x, &p->z);
    fe_add(&p->z, &p->x, &p->x);
    fe_sub(&a, &p->z, &d);
    fe_mul(&p->x, &a, &b);
    fe_mul(&p->z, &p->t, &a);
    fe_mul(&p->y, &p->t, &d);
    fe_mul(&p->t, &b, &d);
    return succ;
}
static void pt_encode_like_x448(uint8_t out[56], const pt *p) { /* ref: goldilocks.c:1104-1115 */
    fe t, z, y;
    fe_invert(&t, &p->x);
    fe_mul(&z, &t, &p->y);
    fe_sqr(&y, &z);
    fe_serialize(out, &y);
}

/* ---- Elligator 2 ------------------------------------------------------------------------------------ */
static void pt_from_hash_nonuniform(pt *p, const uint8_t ser[56]) { /* ref: elligator.c:32-84 */
    fe r0, r, a, b, c, N, e;
    (void)fe_deserialize(&r0, ser);
    fe_sqr(&a, &r0);
    fe_neg(&r, &a);                      /* gf_mul_qnr: qnr = -1 since p = 3 mod 4 (ref: field.h:84-98) */
    fe_sub(&a, &r, &FE_ONE);
    fe_mulw(&b, &a, EDWARDS_D);
    fe_add(&a, &b, &FE_ONE);
    fe_sub(&b, &b, &r);
    fe_mul(&c, &a, &b);
    fe_add(&a, &r, &FE_ONE);
    fe_mulw(&N, &a, 1 - 2 * EDWARDS_D);
    fe_mul(&a, &c, &N);
    int square = fe_isr(&b, &a);
    c = square ? FE_ONE : r0;
    fe_mul(&e, &b, &c);
    fe_mul(&a, &N, &e);
    fe_cond_neg(&a, fe_lobit(&a) ^ !square);
    fe_mulw(&c, &e, 1 - 2 * EDWARDS_D);
    fe_sqr(&b, &c);
    fe_sub(&e, &r, &FE_ONE);
    fe_mul(&c, &b, &e);
    fe_mul(&b, &c, &N);
    fe_cond_neg(&b, square);
    fe_sub(&b, &b, &FE_ONE);
    fe_sqr(&c, &a);
    fe_add(&a, &a, &a);
    fe_add(&e, &c, &FE_ONE);
    fe_mul(&p->t, &a, &e);
    fe_mul(&p->x, &a, &b);
    fe_sub(&a, &FE_ONE, &c);
    fe_mul(&p->y, &e, &a);
    fe_mul(&p->z, &a, &b);
}
static void pt_from_hash_uniform(pt *p, const uint8_t ser[112]) { /* ref: elligator.c:86-94 */
    pt p2;
    pt_from_hash_nonuniform(p, ser);
    pt_from_hash_nonuniform(&p2, ser + 56);
    pt_addsub(p, p, &p2, 0);
}

/* ---- Elligator inverse ------------------------------------------------------------------------------ */
static void pt_deisogenize_full(fe *s, fe *sum, fe *m1, const pt *p, int toggle_s, int toggle_altx) { /* ref: goldilocks.c:98-134 */
    fe t1, t2, t3, t4;
    fe_add(&t1, &p->x, &p->t);
    fe_sub(&t2, &p->x, &p->t);
    fe_mul(&t3, &t1, &t2);
    fe_sqr(&t2, &p->x);
    fe_mul(&t1, &t2, &t3);
    fe_mulw(&t2, &t1, -1 - TWISTED_D);
    (void)fe_isr(&t1, &t2);
    fe_mul(&t2, &t1, &t3);
    fe_mul(&t4, &t2, &FACTOR);
    int negx = fe_lobit(&t4) ^ toggle_altx;
    fe_cond_neg(&t2, negx);
    fe_mul(&t3, &t2, &p->z);
    fe_sub(sum, &t3, &p->t);
    fe_mul(&t2, sum, &p->x);
    fe_mulw(&t4, &t2, -1 - TWISTED_D);
    fe_mul(s, &t4, &t1);
    int lobs = fe_lobit(s);
    fe_cond_neg(s, lobs);
    *m1 = p->x;
    fe_cond_neg(m1, (!lobs) ^ negx ^ toggle_s);
    fe_add(m1, m1, &p->t);
}
static int pt_invert_elligator_nonuniform(uint8_t out[56], const pt *p, uint32_t hint) { /* ref: elligator.c:104-152 */
    int sgn_s = hint & 1, sgn_altx = (hint >> 1) & 1, sgn_r0 = (hint >> 2) & 1;
    fe a, b, c, t;
    pt_deisogenize_full(&a, &b, &c, p, sgn_s, sgn_altx);
    int is_identity = fe_is_zero(&p->t);
    if (is_identity && sgn_altx) b = FE_ONE;
    if (is_identity && sgn_s && !sgn_altx) c = FE_ONE;
    fe_mulw(&a, &b, EDWARDS_D - 1);
    fe_add(&b, &a, &b);
    fe_sub(&a, &a, &c);
    fe_add(&b, &b, &c);
    if (sgn_s) { t = a; a = b; b = t; }
    fe_neg(&c, &b);
    fe_mul(&b, &c, &a);
    int succ = fe_isr(&c, &b);
    succ |= fe_is_zero(&b);
    fe_mul(&b, &c, &a);
    fe_cond_neg(&b, sgn_r0 ^ fe_lobit(&b));
    succ &= !(fe_is_zero(&b) && (sgn_r0 | sgn_s));
    fe_serialize(out, &b);
    return succ;
}
static int pt_invert_elligator_uniform(uint8_t partial[112], const pt *p, uint32_t hint) { /* ref: elligator.c:154-164 */
    pt p2;
    pt_from_hash_nonuniform(&p2, partial + 56);
    pt_addsub(&p2, p, &p2, 1);
    return pt_invert_elligator_nonuniform(partial, &p2, hint);
}

/* ---- scalar multiplications --------------------------------------------------------------------------- */
static void sc_adjusted_half(scl *r, const scl *s) { /* ref: goldilocks.c:420-421, 842-843 */
    scl t;
    sc_add(&t, s, &ADJUST);
    sc_halve(r, &t);
}
static void prepare_fixed_window(pniels *multiples, const pt *b, int ntable) { /* ref: goldilocks.c:382-403 */
    pt tmp;
    pniels pn;
    pt_double(&tmp, b, 0);
    pt_to_pniels(&pn, &tmp);
    pt_to_pniels(&multiples[0], b);
    tmp = *b;
    for (int i = 1; i < ntable; i++) {
        pt_addsub_pniels(&tmp, &pn, 0, 0);
        pt_to_pniels(&multiples[i], &tmp);
    }
}
/* one signed window: returns table index, *neg = whether to negate.  ref: goldilocks.c:430-442 */
static unsigned signed_window(const scl *s1x, int i, int *neg) {
    unsigned bits = sc_bits(s1x, (unsigned)i, WINDOW);
    unsigned inv = (bits >> (WINDOW - 1)) - 1;      /* all-ones when the top bit is clear */
    bits ^= inv;
    *neg = inv != 0;
    return bits & ((1u << (WINDOW - 1)) - 1);
}
static void pt_scalarmul(pt *a, const pt *b, const scl *scalar) { /* ref: goldilocks.c:405-465 */
    enum { NTABLE = 1 << (WINDOW - 1) };
    scl s1x;
    pniels multiples[NTABLE], pn;
    pt tmp;
    sc_adjusted_half(&s1x, scalar);
    prepare_fixed_window(multiples, b, NTABLE);
    int first = 1;
    for (int i = SCALAR_BITS - ((SCALAR_BITS - 1) % WINDOW) - 1; i >= 0; i -= WINDOW) {
        int neg;
        pn = multiples[signed_window(&s1x, i, &neg)];
        niels_cond_neg(&pn.n, neg);
        if (first) { pniels_to_pt(&tmp, &pn); first = 0; }
        else {
            for (int j = 0; j < WINDOW - 1; j++) pt_double(&tmp, &tmp, 1);
            pt_double(&tmp, &tmp, 0);
            pt_addsub_pniels(&tmp, &pn, 0, i != 0);
        }
    }
    *a = tmp;
}
static void pt_double_scalarmul(pt *a, const pt *b, const scl *sb, const pt *c, const scl *sc_) { /* ref: goldilocks.c:467-541 */
    enum { NTABLE = 1 << (WINDOW - 1) };
    scl s1x, s2x;
    pniels m1[NTABLE], m2[NTABLE], pn;
    pt tmp;
    sc_adjusted_half(&s1x, sb);
    sc_adjusted_half(&s2x, sc_);
    prepare_fixed_window(m1, b, NTABLE);
    prepare_fixed_window(m2, c, NTABLE);
    int first = 1;
    for (int i = SCALAR_BITS - ((SCALAR_BITS - 1) % WINDOW) - 1; i >= 0; i -= WINDOW) {
        int neg;
        pn = m1[signed_window(&s1x, i, &neg)];
        niels_cond_neg(&pn.n, neg);
        if (first) { pniels_to_pt(&tmp, &pn); first = 0; }
        else {
            for (int j = 0; j < WINDOW - 1; j++) pt_double(&tmp, &tmp, 1);
            pt_double(&tmp, &tmp, 0);
            pt_addsub_pniels(&tmp, &pn, 0, 0);
        }
        pn = m2[signed_window(&s2x, i, &neg)];
        niels_cond_neg(&pn.n, neg);
        pt_addsub_pniels(&tmp, &pn, 0, i != 0);
    }
    *a = tmp;
}
/* affine-normalise projective niels: each coordinate times 1/z.  ref: goldilocks.c:728-755 (the
 * reference shares one inversion across the table; the normalised values are the same) */
static void normalize_niels(niels *table, const fe *zs, int n) {
    for (int i = 0; i < n; i++) {
        fe zi;
        fe_invert(&zi, &zs[i]);
        fe_mul(&table[i].a, &table[i].a, &zi);
        fe_mul(&table[i].b, &table[i].b, &zi);
        fe_mul(&table[i].c, &table[i].c, &zi);
    }
}
static void precompute_comb(niels *table, const pt *base) { /* ref: goldilocks.c:757-818 precompute */
    const unsigned n = COMBS_N, t = COMBS_T, s = COMBS_S;
    pt working = *base, start, doubles[COMBS_T - 1];
    pniels pn;
    fe zs[COMBS_N << (COMBS_T - 1)];
    for (unsigned i = 0; i < n; i++) {
        for (unsigned j = 0; j < t; j++) {
            if (j) pt_addsub(&start, &start, &working, 0); else start = working;
            if (j == t - 1 && i == n - 1) break;
            pt_double(&working, &working, 0);
            if (j < t - 1) doubles[j] = working;
            for (unsigned k = 0; k < s - 1; k++) pt_double(&working, &working, k < s - 2);
        }
        for (unsigned j = 0;; j++) {
            unsigned gray = j ^ (j >> 1);
            unsigned idx = (((i + 1) << (t - 1)) - 1) ^ gray;
            pt_to_pniels(&pn, &start);
            table[idx] = pn.n;
            zs[idx] = pn.z;
            if (j >= (1u << (t - 1)) - 1) break;
            unsigned delta = (j + 1) ^ ((j + 1) >> 1) ^ gray, k;
            for (k = 0; delta > 1; k++) delta >>= 1;
            pt_addsub(&start, &start, &doubles[k], !(gray & (1u << k)));
        }
    }
    normalize_niels(table, zs, (int)(n << (t - 1)));
}
static void prepare_wnaf_table(pniels *out, const pt *working, unsigned tbits) { /* ref: goldilocks.c:1204-1230 */
    pt tmp;
    pniels twop;
    pt_to_pniels(&out[0], working);
    if (tbits == 0) return;
    pt_double(&tmp, working, 0);
    pt_to_pniels(&twop, &tmp);
    pt_addsub_pniels(&tmp, &out[0], 0, 0);
    pt_to_pniels(&out[1], &tmp);
    for (int i = 2; i < 1 << tbits; i++) {
        pt_addsub_pniels(&tmp, &twop, 0, 0);
        pt_to_pniels(&out[i], &tmp);
    }
}
static void precompute_wnafs(niels *out, const pt *base) { /* ref: goldilocks.c:1236-1258 */
    pniels tmp[1 << WNAF_FIXED_BITS];
    fe zs[1 << WNAF_FIXED_BITS];
    prepare_wnaf_table(tmp, base, WNAF_FIXED_BITS);
    for (int i = 0; i < 1 << WNAF_FIXED_BITS; i++) { out[i] = tmp[i].n; zs[i] = tmp[i].z; }
    normalize_niels(out, zs, 1 << WNAF_FIXED_BITS);
}
static void pt_precomputed_scalarmul(pt *out, const niels *table, const scl *scalar) { /* ref: goldilocks.c:830-877 */
    const unsigned n = COMBS_N, t = COMBS_T, s = COMBS_S;
    scl s1x;
    sc_adjusted_half(&s1x, scalar);
    for (int i = (int)s - 1; i >= 0; i--) {
        if (i != (int)s - 1) pt_double(out, out, 0);
        for (unsigned j = 0; j < n; j++) {
            unsigned tab = 0;
            for (unsigned k = 0; k < t; k++) {
                unsigned bit = (unsigned)i + s * (k + j * t);
                if (bit < SCALAR_BITS) tab |= (unsigned)sc_bit(&s1x, bit) << k;
            }
            unsigned invert = (tab >> (t - 1)) - 1;
            tab ^= invert;
            tab &= (1u << (t - 1)) - 1;
            niels ni = table[(j << (t - 1)) + tab];
            niels_cond_neg(&ni, invert != 0);
            if (i != (int)s - 1 || j) pt_addsub_niels(out, &ni, 0, j == n - 1 && i);
            else niels_to_pt(out, &ni);
        }
    }
}
/* width-(tbits+1) NAF, most significant digit first.  ref: goldilocks.c:1151-1202 recode_wnaf
 * (restated on the whole integer instead of 16-bit refills: take the lowest set bit, emit the signed
 * odd digit of tbits+1 bits that clears it, continue) */
typedef struct { int power, addend; } wnaf_digit;
static int recode_wnaf(wnaf_digit *out, const scl *scalar, unsigned tbits) {
    uint64_t cur[8];
    wnaf_digit tmp[SCALAR_BITS + 8];
    int n = 0;
    memcpy(cur, scalar->w, 56); cur[7] = 0;
    const int64_t span = (int64_t)1 << (tbits + 1);
    for (int pos = 0; pos < 450 && !mw_is_zero(cur, 8); pos++) {
        if (!((cur[pos / 64] >> (pos % 64)) & 1)) continue;
        int64_t odd = 0;
        for (unsigned k = 0; k <= tbits + 1; k++) odd |= (int64_t)((cur[(pos + (int)k) / 64] >> ((pos + (int)k) % 64)) & 1) << k;
        int64_t delta = odd & (span - 1);
        if (odd & span) delta -= span;
        /* cur -= delta << pos */
        uint64_t mag[8] = {(uint64_t)(delta < 0 ? -delta : delta)}, sh[8];
        mw_shl(sh, mag, 8, pos);
        if (delta > 0) mw_sub(cur, cur, sh, 8); else mw_add(cur, cur, sh, 8);
        tmp[n].power = pos; tmp[n].addend = (int)delta; n++;
    }
    for (int i = 0; i < n; i++) out[i] = tmp[n - 1 - i];
    out[n].power = -1; out[n].addend = 0;
    return n;
}
/* ref: goldilocks.c:1260-1330, including its early return of the identity when scalar2 has no digits */
static void pt_base_double_scalarmul_non_secret(pt *combo, const scl *scalar1, const pt *base2, const scl *scalar2) {
    wnaf_digit cv[SCALAR_BITS + 8], cp[SCALAR_BITS + 8];
    pniels var[1 << WNAF_VAR_BITS];
    recode_wnaf(cp, scalar1, WNAF_FIXED_BITS);
    recode_wnaf(cv, scalar2, WNAF_VAR_BITS);
    prepare_wnaf_table(var, base2, WNAF_VAR_BITS);
    int contp = 0, contv = 0, i = cv[0].power;
    if (i < 0) { pt_identity(combo); return; }
    else if (i > cp[0].power) { pniels_to_pt(combo, &var[cv[0].addend >> 1]); contv++; }
    else if (i == cp[0].power) {
        pniels_to_pt(combo, &var[cv[0].addend >> 1]);
        pt_addsub_niels(combo, &WNAF_BASE[cp[0].addend >> 1], 0, i);
        contv++; contp++;
    } else { i = cp[0].power; niels_to_pt(combo, &WNAF_BASE[cp[0].addend >> 1]); contp++; }
    for (i--; i >= 0; i--) {
        int v = (i == cv[contv].power), p = (i == cp[contp].power);
        pt_double(combo, combo, i && !(v || p));
        if (v) {
            int ad = cv[contv].addend;
            pt_addsub_pniels(combo, &var[(ad > 0 ? ad : -ad) >> 1], ad < 0, i && !p);
            contv++;
        }
        if (p) {
            int ad = cp[contp].addend;
            pt_addsub_niels(combo, &WNAF_BASE[(ad > 0 ? ad : -ad) >> 1], ad < 0, i);
            contp++;
        }
    }
}

/* ---- X448 ------------------------------------------------------------------------------------------------ */
static int x448(uint8_t out[56], const uint8_t base[56], const uint8_t scalar[56]) { /* ref: goldilocks.c:1006-1076 */
    fe x1, x2 = FE_ONE, z2 = FE_ZERO, x3, z3 = FE_ONE, t1, t2;
    (void)fe_deserialize(&x1, base);
    x3 = x1;
    int swap = 0;
    for (int t = 447; t >= 0; t--) {
        uint8_t sb = scalar[t / 8];
        if (t / 8 == 0) sb &= 0xfc;             /* -(uint8_t)COFACTOR */
        else if (t == 447) sb = 0xff;
        int k_t = (sb >> (t % 8)) & 1;
        swap ^= k_t;
        if (swap) { fe s = x2; x2 = x3; x3 = s; s = z2; z2 = z3; z3 = s; }
        swap = k_t;
        fe_add(&t1, &x2, &z2);
        fe_sub(&t2, &x2, &z2);
        fe_sub(&z2, &x3, &z3);
        fe_mul(&x2, &t1, &z2);
        fe_add(&z2, &z3, &x3);
        fe_mul(&x3, &t2, &z2);
        fe_sub(&z3, &x2, &x3);
        fe_sqr(&z2, &z3);
        fe_mul(&z3, &x1, &z2);
        fe_add(&z2, &x2, &x3);
        fe_sqr(&x3, &z2);
        fe_sqr(&z2, &t1);
        fe_sqr(&t1, &t2);
        fe_mul(&x2, &z2, &t1);
        fe_sub(&t2, &z2, &t1);
        fe_mulw(&t1, &t2, -EDWARDS_D);
        fe_add(&t1, &t1, &z2);
        fe_mul(&z2, &t2, &t1);
    }
    if (swap) { fe s = x2; x2 = x3; x3 = s; s = z2; z2 = z3; z3 = s; }
    fe_invert(&z2, &z2);
    fe_mul(&x1, &x2, &z2);
    fe_serialize(out, &x1);
    return !fe_is_zero(&x1);
}
static void x448_derive_public_key(uint8_t out[56], const uint8_t scalar[56]) { /* ref: goldilocks.c:1117-1141 */
    uint8_t s2[56];
    scl s;
    pt p;
    memcpy(s2, scalar, 56);
    s2[0] &= 0xfc;
    s2[55] |= 0x80;                      /* X_PRIVATE_BITS = 448: the clear-mask is empty, bit 447 is set */
    sc_decode_long(&s, s2, 56);
    sc_halve(&s, &s);                    /* GOLDILOCKS_X448_ENCODE_RATIO = 2 (ref: point_448.h:57) */
    pt_precomputed_scalarmul(&p, COMB, &s);
    pt_encode_like_x448(out, &p);
}

/* =================================================================================================
 * SHAKE256                         ref: src/shake.c:60-162 (sponge), 211-213 (rate 136, pad 0x1f/0x80)
 * ================================================================================================= */
typedef struct { uint64_t a[25]; unsigned pos; int squeezing; } shake;
static uint64_t rol64(uint64_t x, unsigned s) { return s ? (x << s) | (x >> (64 - s)) : x; }
static void keccak_f(uint64_t a[25]) { /* FIPS 202 3.2-3.3; ref: shake.c:60-87 */
    uint64_t lfsr = 1;
    for (int round = 0; round < 24; round++) {
        uint64_t c[5], b[25];
        for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
        for (int x = 0; x < 5; x++) {
            uint64_t d = c[(x + 4) % 5] ^ rol64(c[(x + 1) % 5], 1);
            for (int y = 0; y < 25; y += 5) a[x + y] ^= d;
        }
        int x = 1, y = 0;
        b[0] = a[0];
        for (int t = 0; t < 24; t++) {               /* rho + pi */
            int nx = y, ny = (2 * x + 3 * y) % 5;
            b[nx + 5 * ny] = rol64(a[x + 5 * y], (unsigned)((t + 1) * (t + 2) / 2) % 64);
            x = nx; y = ny;
        }
        for (y = 0; y < 25; y += 5)
            for (x = 0; x < 5; x++) a[x + y] = b[x + y] ^ (~b[(x + 1) % 5 + y] & b[(x + 2) % 5 + y]);
        uint64_t rc = 0;                             /* iota: round constant from the degree-8 LFSR */
        for (int j = 0; j < 7; j++) {
            if (lfsr & 1) rc ^= 1ull << ((1u << j) - 1);
            lfsr = (lfsr & 0x80) ? ((lfsr << 1) ^ 0x171) : (lfsr << 1);
        }
        a[0] ^= rc;
    }
}
static void shake_init(shake *h) { memset(h, 0, sizeof *h); }
static void shake_update(shake *h, const uint8_t *in, size_t len) {
    for (size_t i = 0; i < len; i++) {
        h->a[h->pos / 8] ^= (uint64_t)in[i] << (8 * (h->pos % 8));
        if (++h->pos == 136) { keccak_f(h->a); h->pos = 0; }
    }
}
static void shake_output(shake *h, uint8_t *out, size_t len) {
    if (!h->squeezing) {
        h->a[h->pos / 8] ^= (uint64_t)0x1f << (8 * (h->pos % 8));
        h->a[16] ^= 0x80ull << 56;
        keccak_f(h->a);
        h->pos = 0;
        h->squeezing = 1;
    }
    for (size_t i = 0; i < len; i++) {
        if (h->pos == 136) { keccak_f(h->a); h->pos = 0; }
        out[i] = (uint8_t)(h->a[h->pos / 8] >> (8 * (h->pos % 8)));
        h->pos++;
    }
}

/* =================================================================================================
 * EdDSA                                                                        ref: src/eddsa.c
 * ================================================================================================= */
static void ed_clamp(uint8_t s[57]) { s[0] &= 0xfc; s[56] = 0; s[55] |= 0x80; } /* ref: eddsa.c:34-48 */
static void ed_hash_init_with_dom(shake *h, uint8_t prehashed, const uint8_t *ctx, uint8_t ctx_len) { /* ref: eddsa.c:51-74 */
    const uint8_t dom[2] = {(uint8_t)(prehashed ? 1 : 0), ctx_len};
    shake_init(h);
    shake_update(h, (const uint8_t *)"SigEd448", 8);
    shake_update(h, dom, 2);
    shake_update(h, ctx, ctx_len);
}
static void ed_secret_scalar(scl *secret, uint8_t seed[57], const uint8_t sk[57]) { /* ref: eddsa.c:98-117, 161-171 */
    uint8_t expanded[114];
    shake h;
    shake_init(&h);
    shake_update(&h, sk, 57);
    shake_output(&h, expanded, 114);
    ed_clamp(expanded);
    sc_decode_long(secret, expanded, 57);
    if (seed) memcpy(seed, expanded + 57, 57);
}
static void ed_derive_public_key(uint8_t pk[57], const uint8_t sk[57]) { /* ref: eddsa.c:129-144 */
    scl s;
    pt p;
    ed_secret_scalar(&s, NULL, sk);
    sc_halve(&s, &s);
    sc_halve(&s, &s);                    /* GOLDILOCKS_448_EDDSA_ENCODE_RATIO = 4 (ref: ed448.h:49) */
    pt_precomputed_scalarmul(&p, COMB, &s);
    pt_encode_like_eddsa(pk, &p);
}
static void ed_challenge(scl *c, const uint8_t r[57], const uint8_t pk[57], const uint8_t *msg, size_t len,
                         uint8_t prehashed, const uint8_t *ctx, uint8_t ctx_len) {
    shake h;
    uint8_t out[114];
    ed_hash_init_with_dom(&h, prehashed, ctx, ctx_len);
    shake_update(&h, r, 57);
    shake_update(&h, pk, 57);
    shake_update(&h, msg, len);
    shake_output(&h, out, 114);
    sc_decode_long(c, out, 114);
}
static void ed_sign(uint8_t sig[114], const uint8_t sk[57], const uint8_t pk[57], const uint8_t *msg, size_t len,
                    uint8_t prehashed, const uint8_t *ctx, uint8_t ctx_len) { /* ref: eddsa.c:146-230 */
    scl secret, nonce, n4, chal;
    uint8_t seed[57], out[114];
    shake h;
    pt p;
    ed_secret_scalar(&secret, seed, sk);
    ed_hash_init_with_dom(&h, prehashed, ctx, ctx_len);
    shake_update(&h, seed, 57);
    shake_update(&h, msg, len);
    shake_output(&h, out, 114);
    sc_decode_long(&nonce, out, 114);
    sc_halve(&n4, &nonce);
    sc_halve(&n4, &n4);
    pt_precomputed_scalarmul(&p, COMB, &n4);
    memset(sig, 0, 114);
    pt_encode_like_eddsa(sig, &p);
    ed_challenge(&chal, sig, pk, msg, len, prehashed, ctx, ctx_len);
    sc_mul(&chal, &chal, &secret);
    sc_add(&chal, &chal, &nonce);
    memcpy(sig + 57, chal.w, 56);
}
static int ed_verify(const uint8_t sig[114], const uint8_t pk[57], const uint8_t *msg, size_t len,
                     uint8_t prehashed, const uint8_t *ctx, uint8_t ctx_len) { /* ref: eddsa.c:253-306 */
    pt pkp, rp;
    scl chal, resp, zero = {{0}};
    if (!pt_decode_like_eddsa(&pkp, pk)) return 0;
    if (!pt_decode_like_eddsa(&rp, sig)) return 0;
    ed_challenge(&chal, sig, pk, msg, len, prehashed, ctx, ctx_len);
    sc_sub(&chal, &zero, &chal);
    sc_decode_long(&resp, sig + 57, 57);           /* reduced mod q, no range check */
    pt combo;                                       /* GOLDILOCKS_448_EDDSA_DECODE_RATIO = 1: no doubling */
    pt_base_double_scalarmul_non_secret(&combo, &resp, &pkp, &chal);
    return pt_eq(&combo, &rp);
}

/* =================================================================================================
 * One-time setup: constants and the two fixed-base tables      ref: src/goldilocks_gen_tables.c:59-124
 * ================================================================================================= */
static pthread_once_t g_once = PTHREAD_ONCE_INIT;
static void oracle_init_once(void) {
    fe d = {{39081}}, dinv;
    fe_neg(&d, &d);
    fe_invert(&dinv, &d);
    (void)fe_isr(&FACTOR, &dinv);                    /* gen_tables / goldilocks.c:41-43: isr(1/d) */
    uint64_t pow450[8] = {0, 0, 0, 0, 0, 0, 0, 4};  /* 2^450 */
    scl one = {{1}}, t;
    sc_reduce_wide(&t, pow450, 8);
    sc_sub(&ADJUST, &t, &one);
    uint8_t ser[56];
    memset(ser, 0x66, 28); memset(ser + 28, 0x33, 28); /* ref: goldilocks_gen_tables.c:21-23 */
    (void)pt_decode(&BASE, ser, 0);
    precompute_comb(COMB, &BASE);
    precompute_wnafs(WNAF_BASE, &BASE);
}
static void oracle_init(void) { pthread_once(&g_once, oracle_init_once); }

/* =================================================================================================
 * Host ABI conversion (ref: f_field.h:23-27 8 x u64 radix 2^56 limbs; point_448.h:66-86)
 * ================================================================================================= */
typedef struct { uint64_t limb[8]; } abi_gf;
typedef struct { abi_gf x, y, z, t; } abi_pt;
static void fe_from_abi(fe *r, const abi_gf *a) {
    uint64_t acc[9] = {0};
    for (int k = 0; k < 8; k++) {
        uint64_t one[9] = {a->limb[k]}, sh[9];
        mw_shl(sh, one, 9, 56 * k);
        mw_add(acc, acc, sh, 9);
    }
    /* acc < 2^(56*7+64+1): fold word 8 then the rest */
    uint64_t t[8];
    memcpy(t, acc, 64);
    if (acc[8]) { uint64_t hi[8] = {0, acc[8], 0, 0, acc[8] << 32, acc[8] >> 32, 0, 0}; /* 2^512 = 2^64 * (2^224 + 1) */
        uint64_t c = mw_add(t, t, hi, 8); (void)c; }
    fe_from8(r, t);
}
static void fe_to_abi(abi_gf *o, const fe *a) {
    for (int k = 0; k < 8; k++) {
        uint64_t sh[7];
        mw_shr(sh, a->w, 7, 56 * k);
        o->limb[k] = sh[0] & 0xffffffffffffffull;
    }
}
static void pt_from_abi(pt *p, const abi_pt *a) { fe_from_abi(&p->x, &a->x); fe_from_abi(&p->y, &a->y); fe_from_abi(&p->z, &a->z); fe_from_abi(&p->t, &a->t); }
static void pt_to_abi(abi_pt *o, const pt *p) { fe_to_abi(&o->x, &p->x); fe_to_abi(&o->y, &p->y); fe_to_abi(&o->z, &p->z); fe_to_abi(&o->t, &p->t); }

/* =================================================================================================
 * Batched exports (plain loops, optionally split over pthreads)
 * ================================================================================================= */
static int g_threads = 1;
EXPORT void oracle_set_threads(int t) { g_threads = t < 1 ? 1 : (t > 1024 ? 1024 : t); }
typedef void (*elem_fn)(size_t i, void *ctx);
typedef struct { elem_fn fn; void *ctx; size_t lo, hi; } job_t;
static void *job_main(void *p) { job_t *j = (job_t *)p; for (size_t i = j->lo; i < j->hi; i++) j->fn(i, j->ctx); return NULL; }
static int32_t pfor(elem_fn fn, void *ctx, size_t n) {
    oracle_init();
    int nt = g_threads;
    if ((size_t)nt > n) nt = n ? (int)n : 1;
    if (nt <= 1) { for (size_t i = 0; i < n; i++) fn(i, ctx); return -1; }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nt);
    job_t *jobs = (job_t *)malloc(sizeof(job_t) * (size_t)nt);
    for (int t = 0; t < nt; t++) {
        jobs[t] = (job_t){fn, ctx, n * (size_t)t / (size_t)nt, n * (size_t)(t + 1) / (size_t)nt};
        pthread_create(&th[t], NULL, job_main, &jobs[t]);
    }
    for (int t = 0; t < nt; t++) pthread_join(th[t], NULL);
    free(th); free(jobs);
    return -1;
}
typedef struct {
    int op;
    void *o0, *o1;
    const void *i0, *i1, *i2, *i3;
    const size_t *off;
    size_t len;
    uint64_t flag;
    const uint8_t *ctx; uint8_t ctx_len, prehashed;
} args_t;
enum { GF_MUL, GF_SQR, GF_ADD, GF_SUB, GF_MULW, GF_ISR, GF_INV, PT_ADD, PT_SUB, PT_DBL, PT_NEG, PT_EQ, PT_VALID, PT_ENC, PT_DEC,
       H2C_NU, H2C_U, PT_SMUL, PT_DSMUL, COMB_MUL, BDSM, ENC_ED, DEC_ED, ENC_X, SC_ADD, SC_SUB, SC_MUL, SC_HALVE, SC_DECODE_LONG,
       X448, X448_PK, ED_PK, ED_SIGN, ED_VERIFY, SHAKE, PT_COORDS, INV_ELL_NU, INV_ELL_U };
static const uint8_t EMPTY = 0;

static void elem(size_t i, void *vp) {
    const args_t *a = (const args_t *)vp;
    switch (a->op) {
    case GF_MUL: case GF_SQR: case GF_ADD: case GF_SUB: case GF_MULW: case GF_ISR: case GF_INV: {
        fe x, y, z;
        (void)fe_deserialize(&x, (const uint8_t *)a->i0 + 56 * i);
        if (a->i1) (void)fe_deserialize(&y, (const uint8_t *)a->i1 + 56 * i);
        if (a->op == GF_MUL) fe_mul(&z, &x, &y);
        else if (a->op == GF_SQR) fe_sqr(&z, &x);
        else if (a->op == GF_ADD) fe_add(&z, &x, &y);
        else if (a->op == GF_SUB) fe_sub(&z, &x, &y);
        else if (a->op == GF_MULW) fe_mulw(&z, &x, (int64_t)a->flag);
        else if (a->op == GF_ISR) ((int32_t *)a->o1)[i] = fe_isr(&z, &x) ? -1 : 0;
        else fe_invert(&z, &x);
        fe_serialize((uint8_t *)a->o0 + 56 * i, &z);
        break;
    }
    case PT_ADD: case PT_SUB: case PT_DBL: case PT_NEG: {
        pt p, q, r;
        pt_from_abi(&q, (const abi_pt *)a->i0 + i);
        if (a->i1) pt_from_abi(&r, (const abi_pt *)a->i1 + i);
        if (a->op == PT_ADD) pt_addsub(&p, &q, &r, 0);
        else if (a->op == PT_SUB) pt_addsub(&p, &q, &r, 1);
        else if (a->op == PT_DBL) pt_double(&p, &q, 0);
        else pt_negate(&p, &q);
        pt_to_abi((abi_pt *)a->o0 + i, &p);
        break;
    }
    case PT_EQ: case PT_VALID: {
        pt q, r;
        pt_from_abi(&q, (const abi_pt *)a->i0 + i);
        if (a->i1) pt_from_abi(&r, (const abi_pt *)a->i1 + i);
        ((uint64_t *)a->o0)[i] = (a->op == PT_EQ ? pt_eq(&q, &r) : pt_valid(&q)) ? ~0ull : 0;
        break;
    }
    case PT_ENC: { pt q; pt_from_abi(&q, (const abi_pt *)a->i0 + i); pt_encode((uint8_t *)a->o0 + 56 * i, &q); break; }
    case PT_DEC: {
        pt p;
        int ok = pt_decode(&p, (const uint8_t *)a->i0 + 56 * i, a->flag != 0);
        pt_to_abi((abi_pt *)a->o0 + i, &p);
        ((int32_t *)a->o1)[i] = ok ? -1 : 0;
        break;
    }
    case H2C_NU: { pt p; pt_from_hash_nonuniform(&p, (const uint8_t *)a->i0 + 56 * i); pt_to_abi((abi_pt *)a->o0 + i, &p); break; }
    case H2C_U: { pt p; pt_from_hash_uniform(&p, (const uint8_t *)a->i0 + 112 * i); pt_to_abi((abi_pt *)a->o0 + i, &p); break; }
    case INV_ELL_NU: case INV_ELL_U: {
        pt q; pt_from_abi(&q, (const abi_pt *)a->i0 + i);
        const uint32_t hint = ((const uint32_t *)a->i1)[i];
        int ok = a->op == INV_ELL_NU ? pt_invert_elligator_nonuniform((uint8_t *)a->o0 + 56 * i, &q, hint)
                                     : pt_invert_elligator_uniform((uint8_t *)a->o0 + 112 * i, &q, hint);
        ((int32_t *)a->o1)[i] = ok ? -1 : 0;
        break;
    }
    case PT_SMUL: {
        pt p, q; scl s;
        pt_from_abi(&q, (const abi_pt *)a->i0 + i); memcpy(s.w, (const uint8_t *)a->i1 + 56 * i, 56);
        pt_scalarmul(&p, &q, &s);
        pt_to_abi((abi_pt *)a->o0 + i, &p);
        break;
    }
    case PT_DSMUL: {
        pt p, q, r; scl s, t;
        pt_from_abi(&q, (const abi_pt *)a->i0 + i); memcpy(s.w, (const uint8_t *)a->i1 + 56 * i, 56);
        pt_from_abi(&r, (const abi_pt *)a->i2 + i); memcpy(t.w, (const uint8_t *)a->i3 + 56 * i, 56);
        pt_double_scalarmul(&p, &q, &s, &r, &t);
        pt_to_abi((abi_pt *)a->o0 + i, &p);
        break;
    }
    case COMB_MUL: {
        pt p; scl s; memcpy(s.w, (const uint8_t *)a->i0 + 56 * i, 56);
        pt_precomputed_scalarmul(&p, COMB, &s);
        pt_to_abi((abi_pt *)a->o0 + i, &p);
        break;
    }
    case BDSM: {
        pt p, q; scl s, t;
        memcpy(s.w, (const uint8_t *)a->i0 + 56 * i, 56); pt_from_abi(&q, (const abi_pt *)a->i1 + i); memcpy(t.w, (const uint8_t *)a->i2 + 56 * i, 56);
        pt_base_double_scalarmul_non_secret(&p, &s, &q, &t);
        pt_to_abi((abi_pt *)a->o0 + i, &p);
        break;
    }
    case ENC_ED: { pt q; pt_from_abi(&q, (const abi_pt *)a->i0 + i); pt_encode_like_eddsa((uint8_t *)a->o0 + 57 * i, &q); break; }
    case DEC_ED: {
        pt p;
        int ok = pt_decode_like_eddsa(&p, (const uint8_t *)a->i0 + 57 * i);
        pt_to_abi((abi_pt *)a->o0 + i, &p);
        ((int32_t *)a->o1)[i] = ok ? -1 : 0;
        break;
    }
    case ENC_X: { pt q; pt_from_abi(&q, (const abi_pt *)a->i0 + i); pt_encode_like_x448((uint8_t *)a->o0 + 56 * i, &q); break; }
    case PT_COORDS: {
        pt q; pt_from_abi(&q, (const abi_pt *)a->i0 + i);
        uint8_t *o = (uint8_t *)a->o0 + 224 * i;
        fe_serialize(o, &q.x); fe_serialize(o + 56, &q.y); fe_serialize(o + 112, &q.z); fe_serialize(o + 168, &q.t);
        break;
    }
    case SC_ADD: case SC_SUB: case SC_MUL: case SC_HALVE: {
        scl r, x, y;
        memcpy(x.w, (const uint8_t *)a->i0 + 56 * i, 56);
        if (a->i1) memcpy(y.w, (const uint8_t *)a->i1 + 56 * i, 56);
        if (a->op == SC_ADD) sc_add(&r, &x, &y);
        else if (a->op == SC_SUB) sc_sub(&r, &x, &y);
        else if (a->op == SC_MUL) sc_mul(&r, &x, &y);
        else sc_halve(&r, &x);
        memcpy((uint8_t *)a->o0 + 56 * i, r.w, 56);
        break;
    }
    case SC_DECODE_LONG: { scl r; sc_decode_long(&r, (const uint8_t *)a->i0 + a->len * i, a->len); memcpy((uint8_t *)a->o0 + 56 * i, r.w, 56); break; }
    case X448: ((int32_t *)a->o1)[i] = x448((uint8_t *)a->o0 + 56 * i, (const uint8_t *)a->i0 + 56 * i, (const uint8_t *)a->i1 + 56 * i) ? -1 : 0; break;
    case X448_PK: x448_derive_public_key((uint8_t *)a->o0 + 56 * i, (const uint8_t *)a->i0 + 56 * i); break;
    case ED_PK: ed_derive_public_key((uint8_t *)a->o0 + 57 * i, (const uint8_t *)a->i0 + 57 * i); break;
    case ED_SIGN:
        ed_sign((uint8_t *)a->o0 + 114 * i, (const uint8_t *)a->i0 + 57 * i, (const uint8_t *)a->i1 + 57 * i,
                (const uint8_t *)a->i2 + a->off[i], a->off[i + 1] - a->off[i], a->prehashed, a->ctx, a->ctx_len);
        break;
    case ED_VERIFY:
        ((int32_t *)a->o0)[i] = ed_verify((const uint8_t *)a->i0 + 114 * i, (const uint8_t *)a->i1 + 57 * i,
                (const uint8_t *)a->i2 + a->off[i], a->off[i + 1] - a->off[i], a->prehashed, a->ctx, a->ctx_len) ? -1 : 0;
        break;
    case SHAKE: {
        shake h;
        shake_init(&h);
        shake_update(&h, (const uint8_t *)a->i0 + a->off[i], a->off[i + 1] - a->off[i]);
        shake_output(&h, (uint8_t *)a->o0 + a->len * i, a->len);
        break;
    }
    }
}
#define A0(OP) args_t a; memset(&a, 0, sizeof a); a.op = OP
EXPORT int32_t goldilocks_448_gf_mul_batch(uint8_t *o, const uint8_t *x, const uint8_t *y, size_t n) { A0(GF_MUL); a.o0 = o; a.i0 = x; a.i1 = y; return pfor(elem, &a, n); }
EXPORT int32_t goldilocks_448_gf_sqr_batch(uint8_t *o, const uint8_t *x, size_t n) { A0(GF_SQR); a.o0 = o; a.i0 = x; return pfor(elem, &a, n); }
EXPORT int32_t goldilocks_448_gf_add_batch(uint8_t *o, const uint8_t *x, const uint8_t *y, size_t n) { A0(GF_ADD); a.o0 = o; a.i0 = x; a.i1 = y; return pfor(elem, &a, n); }
EXPORT int32_t goldilocks_448_gf_sub_batch(uint8_t *o, const uint8_t *x, const uint8_t *y, size_t n) { A0(GF_SUB); a.o0 = o; a.i0 = x; a.i1 = y; return pfor(elem, &a, n); }
EXPORT int32_t goldilocks_448_gf_mulw_batch(uint8_t *o, const uint8_t *x, uint32_t w, size_t n) { A0(GF_MULW); a.o0 = o; a.i0 = x; a.flag = w; return pfor(elem, &a, n); }
EXPORT int32_t goldilocks_448_gf_isr_batch(uint8_t *o, int32_t *st, const uint8_t *x, size_t n) { A0(GF_ISR); a.o0 = o; a.o1 = st; a.i0 = x; return pfor(elem, &a, n); }
EXPORT int32_t goldilocks_448_gf_invert_batch(uint8_t *o, const uint8_t *x, size_t n) { A0(GF_INV); a.o0 = o; a.i0 = x; return pfor(elem, &a, n); }
EXPORT int32_t goldilocks_448_point_add_batch(abi_pt *o, const abi_pt *x, const abi_pt *y, size_t n) { A0(PT_ADD); a.o0 = o; a.i0 = x; a.i1 = y; return pfor(elem, &a, n); }
EXPORT int32_t goldilocks_448_point_sub_batch(abi_pt *o, const abi_pt *x, const abi_pt *y, size_t n) { A0(PT_SUB); a.o0 = o; a.i0 = x; a.i1 = y; return pfor(elem, &a, n); }
EXPORT int32_t goldilocks_448_point_double_batch(abi_pt *o, const abi_pt *x, size_t n) { A0(PT_DBL); a.o0 = o; a.i0 = x; return pfor(elem, &a, n); }
EXPORT int32_t goldilocks_448_point_negate_batch(abi_pt *o, const abi_pt *x, size_t n) { A0(PT_NEG); a.o0 = o; a.i0 = x; return pfor(elem, &a, n); }
EXPORT int32_t goldilocks_448_point_eq_batch(uint64_t *o, const abi_pt *x, const abi_pt *y, size_t n) { A0(PT_EQ); a.o0 = o; a.i0 = x; a.i1 = y; return pfor(elem, &a, n); }
EXPORT int32_t goldilocks_448_point_valid_batch(uint64_t *o, const abi_pt *x, size_t n) { A0(PT_VALID); a.o0 = o; a.i0 = x; return pfor(elem, &a, n); }
EXPORT int32_t goldilocks_448_point_encode_batch(uint8_t *o, const abi_pt *x, size_t n) { A0(PT_ENC); a.o0 = o; a.i0 = x; return pfor(elem, &a, n); }
EXPORT int32_t goldilocks_448_point_decode_batch(abi_pt *o, int32_t *st, const uint8_t *ser, uint64_t allow_identity, size_t n) { A0(PT_DEC); a.o0 = o; a.o1 = st; a.i0 = ser; a.flag = allow_identity; return pfor(elem, &a, n); }
EXPORT int32_t goldilocks_448_point_from_hash_nonuniform_batch(abi_pt *o, const uint8_t *h, size_t n) { A0(H2C_NU); a.o0 = o; a.i0 = h; return pfor(elem, &a, n); }
EXPORT int32_t goldilocks_448_point_from_hash_uniform_batch(abi_pt *o, const uint8_t *h, size_t n) { A0(H2C_U); a.o0 = o; a.i0 = h; return pfor(elem, &a, n); }
EXPORT int32_t goldilocks_448_invert_elligator_nonuniform_batch(uint8_t *h, int32_t *st, const abi_pt *x, const uint32_t *hint, size_t n) { A0(INV_ELL_NU); a.o0 = h; a.o1 = st; a.i0 = x; a.i1 = hint; return pfor(elem, &a, n); }
EXPORT int32_t goldilocks_448_invert_elligator_uniform_batch(uint8_t *h, int32_t *st, const abi_pt *x, const uint32_t *hint, size_t n) { A0(INV_ELL_U); a.o0 = h; a.o1 = st; a.i0 = x; a.i1 = hint; return pfor(elem, &a, n); }
EXPORT int32_t goldilocks_448_point_scalarmul_batch(abi_pt *o, const abi_pt *b, const void *s, size_t n) { A0(PT_SMUL); a.o0 = o; a.i0 = b; a.i1 = s; return pfor(elem, &a, n); }
EXPORT int32_t goldilocks_448_point_double_scalarmul_batch(abi_pt *o, const abi_pt *b1, const void *s1, const abi_pt *b2, const void *s2, size_t n) { A0(PT_DSMUL); a.o0 = o; a.i0 = b1; a.i1 = s1; a.i2 = b2; a.i3 = s2; return pfor(elem, &a, n); }
EXPORT int32_t goldilocks_448_precomputed_scalarmul_batch(abi_pt *o, const void *table, const void *s, size_t n) { (void)table; A0(COMB_MUL); a.o0 = o; a.i0 = s; return pfor(elem, &a, n); }
EXPORT int32_t goldilocks_448_base_double_scalarmul_non_secret_batch(abi_pt *o, const void *s1, const abi_pt *b2, const void *s2, size_t n) { A0(BDSM); a.o0 = o; a.i0 = s1; a.i1 = b2; a.i2 = s2; return pfor(elem, &a, n); }
EXPORT int32_t goldilocks_448_point_mul_by_ratio_and_encode_like_eddsa_batch(uint8_t *o, const abi_pt *x, size_t n) { A0(ENC_ED); a.o0 = o; a.i0 = x; return pfor(elem, &a, n); }
EXPORT int32_t goldilocks_448_point_decode_like_eddsa_and_mul_by_ratio_batch(abi_pt *o, int32_t *st, const uint8_t *enc, size_t n) { A0(DEC_ED); a.o0 = o; a.o1 = st; a.i0 = enc; return pfor(elem, &a, n); }
EXPORT int32_t goldilocks_448_point_mul_by_ratio_and_encode_like_x448_batch(uint8_t *o, const abi_pt *x, size_t n) { A0(ENC_X); a.o0 = o; a.i0 = x; return pfor(elem, &a, n); }
EXPORT int32_t oracle_point_coords_batch(uint8_t *o, const abi_pt *x, size_t n) { A0(PT_COORDS); a.o0 = o; a.i0 = x; return pfor(elem, &a, n); }
EXPORT int32_t goldilocks_448_scalar_add_batch(void *o, const void *x, const void *y, size_t n) { A0(SC_ADD); a.o0 = o; a.i0 = x; a.i1 = y; return pfor(elem, &a, n); }
EXPORT int32_t goldilocks_448_scalar_sub_batch(void *o, const void *x, const void *y, size_t n) { A0(SC_SUB); a.o0 = o; a.i0 = x; a.i1 = y; return pfor(elem, &a, n); }
EXPORT int32_t goldilocks_448_scalar_mul_batch(void *o, const void *x, const void *y, size_t n) { A0(SC_MUL); a.o0 = o; a.i0 = x; a.i1 = y; return pfor(elem, &a, n); }
EXPORT int32_t goldilocks_448_scalar_halve_batch(void *o, const void *x, size_t n) { A0(SC_HALVE); a.o0 = o; a.i0 = x; return pfor(elem, &a, n); }
EXPORT int32_t goldilocks_448_scalar_decode_long_batch(void *o, const uint8_t *ser, size_t ser_len, size_t n) { A0(SC_DECODE_LONG); a.o0 = o; a.i0 = ser; a.len = ser_len; return pfor(elem, &a, n); }
EXPORT int32_t goldilocks_x448_batch(uint8_t *o, int32_t *st, const uint8_t *base, const uint8_t *sc, size_t n) { A0(X448); a.o0 = o; a.o1 = st; a.i0 = base; a.i1 = sc; return pfor(elem, &a, n); }
EXPORT int32_t goldilocks_x448_derive_public_key_batch(uint8_t *o, const uint8_t *sc, size_t n) { A0(X448_PK); a.o0 = o; a.i0 = sc; return pfor(elem, &a, n); }
EXPORT int32_t goldilocks_ed448_derive_public_key_batch(uint8_t *pk, const uint8_t *sk, size_t n) { A0(ED_PK); a.o0 = pk; a.i0 = sk; return pfor(elem, &a, n); }
EXPORT int32_t goldilocks_ed448_sign_batch(uint8_t *sig, const uint8_t *sk, const uint8_t *pk, const uint8_t *msg, const size_t *off,
                                           uint8_t prehashed, const uint8_t *ctx, uint8_t ctx_len, size_t n) {
    A0(ED_SIGN); a.o0 = sig; a.i0 = sk; a.i1 = pk; a.i2 = msg ? msg : &EMPTY; a.off = off; a.prehashed = prehashed;
    a.ctx = ctx ? ctx : &EMPTY; a.ctx_len = ctx_len; return pfor(elem, &a, n);
}
EXPORT int32_t goldilocks_ed448_verify_batch(int32_t *st, const uint8_t *sig, const uint8_t *pk, const uint8_t *msg, const size_t *off,
                                             uint8_t prehashed, const uint8_t *ctx, uint8_t ctx_len, size_t n) {
    A0(ED_VERIFY); a.o0 = st; a.i0 = sig; a.i1 = pk; a.i2 = msg ? msg : &EMPTY; a.off = off; a.prehashed = prehashed;
    a.ctx = ctx ? ctx : &EMPTY; a.ctx_len = ctx_len; return pfor(elem, &a, n);
}
EXPORT int32_t goldilocks_shake256_hash_batch(uint8_t *o, size_t outlen, const uint8_t *in, const size_t *off, size_t n) { A0(SHAKE); a.o0 = o; a.i0 = in ? in : &EMPTY; a.off = off; a.len = outlen; return pfor(elem, &a, n); }

/* canonical radix-2^56 limbs, the layout of the reference's generated tables (ref: goldilocks_gen_tables.c:95-121) */
static void export_niels(uint8_t *out, const niels *t, int n) {
    abi_gf *o = (abi_gf *)out;
    for (int e = 0; e < n; e++) { fe_to_abi(&o[3 * e], &t[e].a); fe_to_abi(&o[3 * e + 1], &t[e].b); fe_to_abi(&o[3 * e + 2], &t[e].c); }
}
EXPORT int32_t goldilocks_b200_export_comb_table(uint8_t out[15360]) { oracle_init(); export_niels(out, COMB, COMBS_N << (COMBS_T - 1)); return -1; }
EXPORT int32_t goldilocks_b200_export_wnaf_table(uint8_t out[6144]) { oracle_init(); export_niels(out, WNAF_BASE, 1 << WNAF_FIXED_BITS); return -1; }
EXPORT const char *oracle_name(void) { return "gold_oracle (CPU restatement)"; }
