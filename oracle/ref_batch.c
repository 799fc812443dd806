/*
 * oracle/ref_batch.c -- TEST INFRASTRUCTURE ONLY.
 *
 * The `*_batch` entry points of include/goldilocks_b200.h implemented as plain loops over the
 * UNMODIFIED reference's single-element functions.  Compiled together with the reference's own
 * sources (oracle/Makefile -> oracle/_ref/libgoldilocks_ref_<arch>.so), so the tests can hand the
 * same packed arrays to the reference and to the CUDA library and compare bytes, and bench.py can
 * time the reference on all host cores (`--impl reference`, `cpu_baseline.kind = "reference"`).
 *
 * This file is ours; it only *calls* the reference (public API from <goldilocks.h> plus the
 * hidden field API from the reference's field.h, which is visible because the oracle build does
 * not pass -fvisibility=hidden).
 */
#define _GNU_SOURCE 1
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "word.h"
#include "field.h"
#include <goldilocks.h>
#include <goldilocks/ed448.h>
#include <goldilocks/shake.h>

#define EXPORT __attribute__((visibility("default")))

/* ---------------------------------------------------------------- parallel-for over [0,n) */
static int g_threads = 1;
EXPORT void refb_set_threads(int t) { g_threads = t < 1 ? 1 : (t > 1024 ? 1024 : t); }
EXPORT int refb_get_threads(void) { return g_threads; }

typedef void (*range_fn)(size_t lo, size_t hi, void *ctx);
typedef struct { range_fn fn; void *ctx; size_t lo, hi; } job_t;
static void *job_main(void *p) { job_t *j = (job_t *)p; j->fn(j->lo, j->hi, j->ctx); return NULL; }

static void pfor(range_fn fn, void *ctx, size_t n) {
    int nt = g_threads;
    if ((size_t)nt > n) nt = n ? (int)n : 1;
    if (nt <= 1) { fn(0, n, ctx); return; }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nt);
    job_t *jobs = (job_t *)malloc(sizeof(job_t) * nt);
    for (int t = 0; t < nt; t++) {
        jobs[t].fn = fn; jobs[t].ctx = ctx;
        jobs[t].lo = n * (size_t)t / nt; jobs[t].hi = n * (size_t)(t + 1) / nt;
        pthread_create(&th[t], NULL, job_main, &jobs[t]);
    }
    for (int t = 0; t < nt; t++) pthread_join(th[t], NULL);
    free(th); free(jobs);
}

/* The reference needs 32-byte aligned point structs (AVX2 build uses aligned vector moves),
 * while test arrays are only packed: copy through aligned temporaries. */
typedef goldilocks_448_point_s pt_t;
typedef struct goldilocks_448_scalar_s sc_t;

typedef struct {
    int op;
    void *o0, *o1;
    const void *i0, *i1, *i2, *i3;
    const size_t *off;
    size_t len;
    uint64_t flag;
    const uint8_t *ctx; uint8_t ctx_len, prehashed;
} args_t;

enum {
    OP_GF_MUL, OP_GF_SQR, OP_GF_ADD, OP_GF_SUB, OP_GF_MULW, OP_GF_ISR, OP_GF_INVERT,
    OP_PT_ADD, OP_PT_SUB, OP_PT_DBL, OP_PT_NEG, OP_PT_EQ, OP_PT_VALID, OP_PT_ENC, OP_PT_DEC,
    OP_H2C_NU, OP_H2C_U, OP_PT_SMUL, OP_PT_DSMUL, OP_COMB, OP_BDSM, OP_ENC_EDDSA, OP_DEC_EDDSA, OP_ENC_X448,
    OP_SC_ADD, OP_SC_SUB, OP_SC_MUL, OP_SC_HALVE, OP_SC_DECODE_LONG,
    OP_X448, OP_X448_PK, OP_ED_PK, OP_ED_SIGN, OP_ED_VERIFY, OP_SHAKE256, OP_PT_COORDS,
    OP_SC_INVERT, OP_PT_DUAL, OP_DIRECT, OP_PRECOMPUTE, OP_COMB_TABLE, OP_TORQUE, OP_PSCALE, OP_PK_TO_X, OP_SK_TO_X, OP_INV_ELL_NU, OP_INV_ELL_U
};

static void run_range(size_t lo, size_t hi, void *vp) {
    const args_t *a = (const args_t *)vp;
    for (size_t i = lo; i < hi; i++) {
        switch (a->op) {
        case OP_GF_MUL: case OP_GF_SQR: case OP_GF_ADD: case OP_GF_SUB: case OP_GF_MULW:
        case OP_GF_ISR: case OP_GF_INVERT: {
            gf x, y, z;
            ignore_result(gf_deserialize(x, (const uint8_t *)a->i0 + 56 * i, 0));
            if (a->i1) ignore_result(gf_deserialize(y, (const uint8_t *)a->i1 + 56 * i, 0));
            switch (a->op) {
            case OP_GF_MUL: gf_mul(z, x, y); break;
            case OP_GF_SQR: gf_sqr(z, x); break;
            case OP_GF_ADD: gf_add(z, x, y); break;
            case OP_GF_SUB: gf_sub(z, x, y); break;
            case OP_GF_MULW: gf_mulw_unsigned(z, x, (uint32_t)a->flag); break;
            case OP_GF_ISR: { mask_t ok = gf_isr(z, x); ((int32_t *)a->o1)[i] = ok ? -1 : 0; break; }
            default: { /* goldilocks.c:69-80 gf_invert is static: same three steps through the field API */
                gf t1, t2; gf_sqr(t1, x); ignore_result(gf_isr(t2, t1)); gf_sqr(t1, t2); gf_mul(z, t1, x); break; }
            }
            gf_serialize((uint8_t *)a->o0 + 56 * i, z);
            break;
        }
        case OP_PT_ADD: case OP_PT_SUB: case OP_PT_DBL: case OP_PT_NEG: {
            pt_t p, q, r;
            memcpy(&q, (const pt_t *)a->i0 + i, sizeof(pt_t));
            if (a->i1) memcpy(&r, (const pt_t *)a->i1 + i, sizeof(pt_t));
            if (a->op == OP_PT_ADD) goldilocks_448_point_add(&p, &q, &r);
            else if (a->op == OP_PT_SUB) goldilocks_448_point_sub(&p, &q, &r);
            else if (a->op == OP_PT_DBL) goldilocks_448_point_double(&p, &q);
            else goldilocks_448_point_negate(&p, &q);
            memcpy((pt_t *)a->o0 + i, &p, sizeof(pt_t));
            break;
        }
        case OP_PT_EQ: case OP_PT_VALID: {
            pt_t q, r;
            memcpy(&q, (const pt_t *)a->i0 + i, sizeof(pt_t));
            if (a->i1) memcpy(&r, (const pt_t *)a->i1 + i, sizeof(pt_t));
            ((uint64_t *)a->o0)[i] = (a->op == OP_PT_EQ) ? goldilocks_448_point_eq(&q, &r) : goldilocks_448_point_valid(&q);
            break;
        }
        case OP_PT_ENC: {
            pt_t q; memcpy(&q, (const pt_t *)a->i0 + i, sizeof(pt_t));
            goldilocks_448_point_encode((uint8_t *)a->o0 + 56 * i, &q);
            break;
        }
        case OP_PT_DEC: {
            pt_t p;
            goldilocks_error_t e = goldilocks_448_point_decode(&p, (const uint8_t *)a->i0 + 56 * i, a->flag);
            memcpy((pt_t *)a->o0 + i, &p, sizeof(pt_t));
            ((int32_t *)a->o1)[i] = (int32_t)e;
            break;
        }
        case OP_H2C_NU: case OP_H2C_U: {
            pt_t p;
            if (a->op == OP_H2C_NU) goldilocks_448_point_from_hash_nonuniform(&p, (const uint8_t *)a->i0 + 56 * i);
            else goldilocks_448_point_from_hash_uniform(&p, (const uint8_t *)a->i0 + 112 * i);
            memcpy((pt_t *)a->o0 + i, &p, sizeof(pt_t));
            break;
        }
        case OP_PT_SMUL: {
            pt_t p, q; sc_t s;
            memcpy(&q, (const pt_t *)a->i0 + i, sizeof(pt_t)); memcpy(&s, (const sc_t *)a->i1 + i, sizeof(sc_t));
            goldilocks_448_point_scalarmul(&p, &q, &s);
            memcpy((pt_t *)a->o0 + i, &p, sizeof(pt_t));
            break;
        }
        case OP_PT_DSMUL: {
            pt_t p, q, r; sc_t s, t;
            memcpy(&q, (const pt_t *)a->i0 + i, sizeof(pt_t)); memcpy(&s, (const sc_t *)a->i1 + i, sizeof(sc_t));
            memcpy(&r, (const pt_t *)a->i2 + i, sizeof(pt_t)); memcpy(&t, (const sc_t *)a->i3 + i, sizeof(sc_t));
            goldilocks_448_point_double_scalarmul(&p, &q, &s, &r, &t);
            memcpy((pt_t *)a->o0 + i, &p, sizeof(pt_t));
            break;
        }
        case OP_COMB: {
            pt_t p; sc_t s; memcpy(&s, (const sc_t *)a->i0 + i, sizeof(sc_t));
            goldilocks_448_precomputed_scalarmul(&p, goldilocks_448_precomputed_base, &s);
            memcpy((pt_t *)a->o0 + i, &p, sizeof(pt_t));
            break;
        }
        case OP_BDSM: {
            pt_t p, q; sc_t s, t;
            memcpy(&s, (const sc_t *)a->i0 + i, sizeof(sc_t)); memcpy(&q, (const pt_t *)a->i1 + i, sizeof(pt_t));
            memcpy(&t, (const sc_t *)a->i2 + i, sizeof(sc_t));
            goldilocks_448_base_double_scalarmul_non_secret(&p, &s, &q, &t);
            memcpy((pt_t *)a->o0 + i, &p, sizeof(pt_t));
            break;
        }
        case OP_ENC_EDDSA: {
            pt_t q; memcpy(&q, (const pt_t *)a->i0 + i, sizeof(pt_t));
            goldilocks_448_point_mul_by_ratio_and_encode_like_eddsa((uint8_t *)a->o0 + 57 * i, &q);
            break;
        }
        case OP_DEC_EDDSA: {
            pt_t p;
            goldilocks_error_t e = goldilocks_448_point_decode_like_eddsa_and_mul_by_ratio(&p, (const uint8_t *)a->i0 + 57 * i);
            memcpy((pt_t *)a->o0 + i, &p, sizeof(pt_t));
            ((int32_t *)a->o1)[i] = (int32_t)e;
            break;
        }
        case OP_ENC_X448: {
            pt_t q; memcpy(&q, (const pt_t *)a->i0 + i, sizeof(pt_t));
            goldilocks_448_point_mul_by_ratio_and_encode_like_x448((uint8_t *)a->o0 + 56 * i, &q);
            break;
        }
        case OP_PT_COORDS: {   /* canonical bytes of X,Y,Z,T: the comparison form for BASELINE config 1 */
            pt_t q; memcpy(&q, (const pt_t *)a->i0 + i, sizeof(pt_t));
            uint8_t *o = (uint8_t *)a->o0 + 224 * i;
            gf_serialize(o, q.x); gf_serialize(o + 56, q.y);
            gf_serialize(o + 112, q.z); gf_serialize(o + 168, q.t);
            break;
        }
        case OP_SC_ADD: case OP_SC_SUB: case OP_SC_MUL: case OP_SC_HALVE: {
            sc_t r, x, y;
            memcpy(&x, (const sc_t *)a->i0 + i, sizeof(sc_t));
            if (a->i1) memcpy(&y, (const sc_t *)a->i1 + i, sizeof(sc_t));
            if (a->op == OP_SC_ADD) goldilocks_448_scalar_add(&r, &x, &y);
            else if (a->op == OP_SC_SUB) goldilocks_448_scalar_sub(&r, &x, &y);
            else if (a->op == OP_SC_MUL) goldilocks_448_scalar_mul(&r, &x, &y);
            else goldilocks_448_scalar_halve(&r, &x);
            memcpy((sc_t *)a->o0 + i, &r, sizeof(sc_t));
            break;
        }
        case OP_SC_DECODE_LONG: {
            sc_t r;
            goldilocks_448_scalar_decode_long(&r, (const uint8_t *)a->i0 + a->len * i, a->len);
            memcpy((sc_t *)a->o0 + i, &r, sizeof(sc_t));
            break;
        }
        case OP_X448: {
            goldilocks_error_t e = goldilocks_x448((uint8_t *)a->o0 + 56 * i, (const uint8_t *)a->i0 + 56 * i, (const uint8_t *)a->i1 + 56 * i);
            ((int32_t *)a->o1)[i] = (int32_t)e;
            break;
        }
        case OP_X448_PK:
            goldilocks_x448_derive_public_key((uint8_t *)a->o0 + 56 * i, (const uint8_t *)a->i0 + 56 * i);
            break;
        case OP_ED_PK:
            goldilocks_ed448_derive_public_key((uint8_t *)a->o0 + 57 * i, (const uint8_t *)a->i0 + 57 * i);
            break;
        case OP_ED_SIGN:
            goldilocks_ed448_sign((uint8_t *)a->o0 + 114 * i, (const uint8_t *)a->i0 + 57 * i, (const uint8_t *)a->i1 + 57 * i,
                                  (const uint8_t *)a->i2 + a->off[i], a->off[i + 1] - a->off[i], a->prehashed, a->ctx, a->ctx_len);
            break;
        case OP_ED_VERIFY: {
            goldilocks_error_t e = goldilocks_ed448_verify((const uint8_t *)a->i0 + 114 * i, (const uint8_t *)a->i1 + 57 * i,
                                  (const uint8_t *)a->i2 + a->off[i], a->off[i + 1] - a->off[i], a->prehashed, a->ctx, a->ctx_len);
            ((int32_t *)a->o0)[i] = (int32_t)e;
            break;
        }
        case OP_SC_INVERT: {
            sc_t r, x; memcpy(&x, (const sc_t *)a->i0 + i, sizeof(sc_t));
            goldilocks_error_t e = goldilocks_448_scalar_invert(&r, &x);
            memcpy((sc_t *)a->o0 + i, &r, sizeof(sc_t));
            ((int32_t *)a->o1)[i] = (int32_t)e;
            break;
        }
        case OP_PT_DUAL: {
            pt_t p1, p2, q; sc_t s, t;
            memcpy(&q, (const pt_t *)a->i0 + i, sizeof(pt_t)); memcpy(&s, (const sc_t *)a->i1 + i, sizeof(sc_t)); memcpy(&t, (const sc_t *)a->i2 + i, sizeof(sc_t));
            goldilocks_448_point_dual_scalarmul(&p1, &p2, &q, &s, &t);
            memcpy((pt_t *)a->o0 + i, &p1, sizeof(pt_t)); memcpy((pt_t *)a->o1 + i, &p2, sizeof(pt_t));
            break;
        }
        case OP_DIRECT: {
            sc_t s; memcpy(&s, (const sc_t *)a->i1 + i, sizeof(sc_t));
            goldilocks_error_t e = goldilocks_448_direct_scalarmul((uint8_t *)a->o0 + 56 * i, (const uint8_t *)a->i0 + 56 * i, &s, a->flag, a->len);
            ((int32_t *)a->o1)[i] = (int32_t)e;
            break;
        }
        case OP_PRECOMPUTE: {
            pt_t q; memcpy(&q, (const pt_t *)a->i0 + i, sizeof(pt_t));
            void *t = NULL;
            if (posix_memalign(&t, 32, goldilocks_448_sizeof_precomputed_s)) break;
            goldilocks_448_precompute((goldilocks_448_precomputed_s *)t, &q);
            memcpy((uint8_t *)a->o0 + goldilocks_448_sizeof_precomputed_s * i, t, goldilocks_448_sizeof_precomputed_s);
            free(t);
            break;
        }
        case OP_COMB_TABLE: {
            pt_t p; sc_t s; memcpy(&s, (const sc_t *)a->i0 + i, sizeof(sc_t));
            goldilocks_448_precomputed_scalarmul(&p, (const goldilocks_448_precomputed_s *)a->i1, &s);
            memcpy((pt_t *)a->o0 + i, &p, sizeof(pt_t));
            break;
        }
        case OP_TORQUE: case OP_PSCALE: {
            pt_t p, q; memcpy(&q, (const pt_t *)a->i0 + i, sizeof(pt_t));
            if (a->op == OP_TORQUE) goldilocks_448_point_debugging_torque(&p, &q);
            else goldilocks_448_point_debugging_pscale(&p, &q, (const uint8_t *)a->i1 + 56 * i);
            memcpy((pt_t *)a->o0 + i, &p, sizeof(pt_t));
            break;
        }
        case OP_INV_ELL_NU: case OP_INV_ELL_U: { /* elligator.c:104-164; hints are per element */
            pt_t q; memcpy(&q, (const pt_t *)a->i0 + i, sizeof(pt_t));
            const uint32_t hint = ((const uint32_t *)a->i1)[i];
            goldilocks_error_t e = a->op == OP_INV_ELL_NU
                ? goldilocks_448_invert_elligator_nonuniform((uint8_t *)a->o0 + 56 * i, &q, hint)
                : goldilocks_448_invert_elligator_uniform((uint8_t *)a->o0 + 112 * i, &q, hint);
            ((int32_t *)a->o1)[i] = (int32_t)e;
            break;
        }
        case OP_PK_TO_X:
            goldilocks_ed448_convert_public_key_to_x448((uint8_t *)a->o0 + 56 * i, (const uint8_t *)a->i0 + 57 * i);
            break;
        case OP_SK_TO_X:
            goldilocks_ed448_convert_private_key_to_x448((uint8_t *)a->o0 + 56 * i, (const uint8_t *)a->i0 + 57 * i);
            break;
        case OP_SHAKE256:
            goldilocks_shake256_hash((uint8_t *)a->o0 + a->len * i, a->len, (const uint8_t *)a->i0 + a->off[i], a->off[i + 1] - a->off[i]);
            break;
        }
    }
}

static int32_t go(args_t *a, size_t n) { pfor(run_range, a, n); return -1; }
static const uint8_t no_ctx = 0;

#define A0 args_t a; memset(&a, 0, sizeof a)
EXPORT int32_t goldilocks_448_gf_mul_batch(uint8_t *o, const uint8_t *x, const uint8_t *y, size_t n) { A0; a.op = OP_GF_MUL; a.o0 = o; a.i0 = x; a.i1 = y; return go(&a, n); }
EXPORT int32_t goldilocks_448_gf_sqr_batch(uint8_t *o, const uint8_t *x, size_t n) { A0; a.op = OP_GF_SQR; a.o0 = o; a.i0 = x; return go(&a, n); }
EXPORT int32_t goldilocks_448_gf_add_batch(uint8_t *o, const uint8_t *x, const uint8_t *y, size_t n) { A0; a.op = OP_GF_ADD; a.o0 = o; a.i0 = x; a.i1 = y; return go(&a, n); }
EXPORT int32_t goldilocks_448_gf_sub_batch(uint8_t *o, const uint8_t *x, const uint8_t *y, size_t n) { A0; a.op = OP_GF_SUB; a.o0 = o; a.i0 = x; a.i1 = y; return go(&a, n); }
EXPORT int32_t goldilocks_448_gf_mulw_batch(uint8_t *o, const uint8_t *x, uint32_t w, size_t n) { A0; a.op = OP_GF_MULW; a.o0 = o; a.i0 = x; a.flag = w; return go(&a, n); }
EXPORT int32_t goldilocks_448_gf_isr_batch(uint8_t *o, int32_t *st, const uint8_t *x, size_t n) { A0; a.op = OP_GF_ISR; a.o0 = o; a.o1 = st; a.i0 = x; return go(&a, n); }
EXPORT int32_t goldilocks_448_gf_invert_batch(uint8_t *o, const uint8_t *x, size_t n) { A0; a.op = OP_GF_INVERT; a.o0 = o; a.i0 = x; return go(&a, n); }

EXPORT int32_t goldilocks_448_point_add_batch(pt_t *o, const pt_t *x, const pt_t *y, size_t n) { A0; a.op = OP_PT_ADD; a.o0 = o; a.i0 = x; a.i1 = y; return go(&a, n); }
EXPORT int32_t goldilocks_448_point_sub_batch(pt_t *o, const pt_t *x, const pt_t *y, size_t n) { A0; a.op = OP_PT_SUB; a.o0 = o; a.i0 = x; a.i1 = y; return go(&a, n); }
EXPORT int32_t goldilocks_448_point_double_batch(pt_t *o, const pt_t *x, size_t n) { A0; a.op = OP_PT_DBL; a.o0 = o; a.i0 = x; return go(&a, n); }
EXPORT int32_t goldilocks_448_point_negate_batch(pt_t *o, const pt_t *x, size_t n) { A0; a.op = OP_PT_NEG; a.o0 = o; a.i0 = x; return go(&a, n); }
EXPORT int32_t goldilocks_448_point_eq_batch(uint64_t *o, const pt_t *x, const pt_t *y, size_t n) { A0; a.op = OP_PT_EQ; a.o0 = o; a.i0 = x; a.i1 = y; return go(&a, n); }
EXPORT int32_t goldilocks_448_point_valid_batch(uint64_t *o, const pt_t *x, size_t n) { A0; a.op = OP_PT_VALID; a.o0 = o; a.i0 = x; return go(&a, n); }
EXPORT int32_t goldilocks_448_point_encode_batch(uint8_t *o, const pt_t *x, size_t n) { A0; a.op = OP_PT_ENC; a.o0 = o; a.i0 = x; return go(&a, n); }
EXPORT int32_t goldilocks_448_point_decode_batch(pt_t *o, int32_t *st, const uint8_t *ser, uint64_t allow_identity, size_t n) { A0; a.op = OP_PT_DEC; a.o0 = o; a.o1 = st; a.i0 = ser; a.flag = allow_identity; return go(&a, n); }
EXPORT int32_t goldilocks_448_point_from_hash_nonuniform_batch(pt_t *o, const uint8_t *h, size_t n) { A0; a.op = OP_H2C_NU; a.o0 = o; a.i0 = h; return go(&a, n); }
EXPORT int32_t goldilocks_448_point_from_hash_uniform_batch(pt_t *o, const uint8_t *h, size_t n) { A0; a.op = OP_H2C_U; a.o0 = o; a.i0 = h; return go(&a, n); }
EXPORT int32_t goldilocks_448_point_scalarmul_batch(pt_t *o, const pt_t *b, const sc_t *s, size_t n) { A0; a.op = OP_PT_SMUL; a.o0 = o; a.i0 = b; a.i1 = s; return go(&a, n); }
EXPORT int32_t goldilocks_448_point_double_scalarmul_batch(pt_t *o, const pt_t *b1, const sc_t *s1, const pt_t *b2, const sc_t *s2, size_t n) { A0; a.op = OP_PT_DSMUL; a.o0 = o; a.i0 = b1; a.i1 = s1; a.i2 = b2; a.i3 = s2; return go(&a, n); }
EXPORT int32_t goldilocks_448_precomputed_scalarmul_batch(pt_t *o, const void *table, const sc_t *s, size_t n) {
    A0; a.o0 = o; a.i0 = s;
    if (table && table != (const void *)goldilocks_448_precomputed_base) { a.op = OP_COMB_TABLE; a.i1 = table; } else a.op = OP_COMB;
    return go(&a, n);
}
EXPORT int32_t goldilocks_448_precompute_batch(void *tables, const pt_t *pts, size_t n) { A0; a.op = OP_PRECOMPUTE; a.o0 = tables; a.i0 = pts; return go(&a, n); }
EXPORT int32_t goldilocks_448_point_dual_scalarmul_batch(pt_t *o1, pt_t *o2, const pt_t *b, const sc_t *s1, const sc_t *s2, size_t n) { A0; a.op = OP_PT_DUAL; a.o0 = o1; a.o1 = o2; a.i0 = b; a.i1 = s1; a.i2 = s2; return go(&a, n); }
EXPORT int32_t goldilocks_448_direct_scalarmul_batch(uint8_t *o, int32_t *st, const uint8_t *b, const sc_t *s, uint64_t allow_identity, uint64_t short_circuit, size_t n) {
    A0; a.op = OP_DIRECT; a.o0 = o; a.o1 = st; a.i0 = b; a.i1 = s; a.flag = allow_identity; a.len = short_circuit; return go(&a, n);
}
EXPORT int32_t goldilocks_448_point_debugging_torque_batch(pt_t *o, const pt_t *x, size_t n) { A0; a.op = OP_TORQUE; a.o0 = o; a.i0 = x; return go(&a, n); }
EXPORT int32_t goldilocks_448_point_debugging_pscale_batch(pt_t *o, const pt_t *x, const uint8_t *f, size_t n) { A0; a.op = OP_PSCALE; a.o0 = o; a.i0 = x; a.i1 = f; return go(&a, n); }
EXPORT int32_t goldilocks_448_invert_elligator_nonuniform_batch(uint8_t *h, int32_t *st, const pt_t *x, const uint32_t *hint, size_t n) { A0; a.op = OP_INV_ELL_NU; a.o0 = h; a.o1 = st; a.i0 = x; a.i1 = hint; return go(&a, n); }
EXPORT int32_t goldilocks_448_invert_elligator_uniform_batch(uint8_t *h, int32_t *st, const pt_t *x, const uint32_t *hint, size_t n) { A0; a.op = OP_INV_ELL_U; a.o0 = h; a.o1 = st; a.i0 = x; a.i1 = hint; return go(&a, n); }
EXPORT int32_t goldilocks_448_scalar_invert_batch(sc_t *o, int32_t *st, const sc_t *x, size_t n) { A0; a.op = OP_SC_INVERT; a.o0 = o; a.o1 = st; a.i0 = x; return go(&a, n); }
EXPORT int32_t goldilocks_ed448_convert_public_key_to_x448_batch(uint8_t *x, const uint8_t *ed, size_t n) { A0; a.op = OP_PK_TO_X; a.o0 = x; a.i0 = ed; return go(&a, n); }
EXPORT int32_t goldilocks_ed448_convert_private_key_to_x448_batch(uint8_t *x, const uint8_t *ed, size_t n) { A0; a.op = OP_SK_TO_X; a.o0 = x; a.i0 = ed; return go(&a, n); }
EXPORT int32_t goldilocks_448_base_double_scalarmul_non_secret_batch(pt_t *o, const sc_t *s1, const pt_t *b2, const sc_t *s2, size_t n) { A0; a.op = OP_BDSM; a.o0 = o; a.i0 = s1; a.i1 = b2; a.i2 = s2; return go(&a, n); }
EXPORT int32_t goldilocks_448_point_mul_by_ratio_and_encode_like_eddsa_batch(uint8_t *o, const pt_t *x, size_t n) { A0; a.op = OP_ENC_EDDSA; a.o0 = o; a.i0 = x; return go(&a, n); }
EXPORT int32_t goldilocks_448_point_decode_like_eddsa_and_mul_by_ratio_batch(pt_t *o, int32_t *st, const uint8_t *enc, size_t n) { A0; a.op = OP_DEC_EDDSA; a.o0 = o; a.o1 = st; a.i0 = enc; return go(&a, n); }
EXPORT int32_t goldilocks_448_point_mul_by_ratio_and_encode_like_x448_batch(uint8_t *o, const pt_t *x, size_t n) { A0; a.op = OP_ENC_X448; a.o0 = o; a.i0 = x; return go(&a, n); }
/* test helper: canonical X|Y|Z|T bytes (224 per point) */
EXPORT int32_t refb_point_coords_batch(uint8_t *o, const pt_t *x, size_t n) { A0; a.op = OP_PT_COORDS; a.o0 = o; a.i0 = x; return go(&a, n); }

EXPORT int32_t goldilocks_448_scalar_add_batch(sc_t *o, const sc_t *x, const sc_t *y, size_t n) { A0; a.op = OP_SC_ADD; a.o0 = o; a.i0 = x; a.i1 = y; return go(&a, n); }
EXPORT int32_t goldilocks_448_scalar_sub_batch(sc_t *o, const sc_t *x, const sc_t *y, size_t n) { A0; a.op = OP_SC_SUB; a.o0 = o; a.i0 = x; a.i1 = y; return go(&a, n); }
EXPORT int32_t goldilocks_448_scalar_mul_batch(sc_t *o, const sc_t *x, const sc_t *y, size_t n) { A0; a.op = OP_SC_MUL; a.o0 = o; a.i0 = x; a.i1 = y; return go(&a, n); }
EXPORT int32_t goldilocks_448_scalar_halve_batch(sc_t *o, const sc_t *x, size_t n) { A0; a.op = OP_SC_HALVE; a.o0 = o; a.i0 = x; return go(&a, n); }
EXPORT int32_t goldilocks_448_scalar_decode_long_batch(sc_t *o, const uint8_t *ser, size_t ser_len, size_t n) { A0; a.op = OP_SC_DECODE_LONG; a.o0 = o; a.i0 = ser; a.len = ser_len; return go(&a, n); }

EXPORT int32_t goldilocks_x448_batch(uint8_t *o, int32_t *st, const uint8_t *base, const uint8_t *sc, size_t n) { A0; a.op = OP_X448; a.o0 = o; a.o1 = st; a.i0 = base; a.i1 = sc; return go(&a, n); }
EXPORT int32_t goldilocks_x448_derive_public_key_batch(uint8_t *o, const uint8_t *sc, size_t n) { A0; a.op = OP_X448_PK; a.o0 = o; a.i0 = sc; return go(&a, n); }
EXPORT int32_t goldilocks_ed448_derive_public_key_batch(uint8_t *pk, const uint8_t *sk, size_t n) { A0; a.op = OP_ED_PK; a.o0 = pk; a.i0 = sk; return go(&a, n); }
EXPORT int32_t goldilocks_ed448_sign_batch(uint8_t *sig, const uint8_t *sk, const uint8_t *pk, const uint8_t *msg, const size_t *off,
                                           uint8_t prehashed, const uint8_t *ctx, uint8_t ctx_len, size_t n) {
    A0; a.op = OP_ED_SIGN; a.o0 = sig; a.i0 = sk; a.i1 = pk; a.i2 = msg ? msg : &no_ctx; a.off = off; a.prehashed = prehashed;
    a.ctx = ctx ? ctx : &no_ctx; a.ctx_len = ctx_len; return go(&a, n);
}
EXPORT int32_t goldilocks_ed448_verify_batch(int32_t *st, const uint8_t *sig, const uint8_t *pk, const uint8_t *msg, const size_t *off,
                                             uint8_t prehashed, const uint8_t *ctx, uint8_t ctx_len, size_t n) {
    A0; a.op = OP_ED_VERIFY; a.o0 = st; a.i0 = sig; a.i1 = pk; a.i2 = msg ? msg : &no_ctx; a.off = off; a.prehashed = prehashed;
    a.ctx = ctx ? ctx : &no_ctx; a.ctx_len = ctx_len; return go(&a, n);
}
EXPORT int32_t goldilocks_shake256_hash_batch(uint8_t *o, size_t outlen, const uint8_t *in, const size_t *off, size_t n) { A0; a.op = OP_SHAKE256; a.o0 = o; a.i0 = in ? in : &no_ctx; a.off = off; a.len = outlen; return go(&a, n); }

/* the reference's generated fixed-base tables, for comparison with the device-built ones */
extern const gf_448_s goldilocks_448_precomputed_base_as_fe[];
extern const gf_448_s goldilocks_448_precomputed_wnaf_as_fe[];
EXPORT void refb_export_comb_table(uint8_t out[15360]) { memcpy(out, goldilocks_448_precomputed_base_as_fe, 15360); }
EXPORT void refb_export_wnaf_table(uint8_t out[6144]) { memcpy(out, goldilocks_448_precomputed_wnaf_as_fe, 6144); }
EXPORT const char *refb_name(void) { return "libgoldilocks reference (unmodified sources)"; }
