// point.cuh -- Ed448-Goldilocks group layer, one point per GPU lane.
//
// Extended twisted-Edwards coordinates (X:Y:Z:T) on the internal curve -x^2 + y^2 = 1 + d' x^2 y^2,
// d' = -39082, and its decaf quotient.  Formula-for-formula the same group law as the reference's
// src/goldilocks.c (cited per function) so that X,Y,Z,T agree mod p with the reference, expressed
// in the TIGHT/LOOSE limb discipline of gf.cuh instead of the reference's GF_HEADROOM bookkeeping.
// Codecs (decaf, RFC 8032, RFC 7748) and Elligator follow goldilocks.c / elligator.c.
//
// Every function takes and returns TIGHT coordinates.
#pragma once
#include "gf.cuh"
#include "consts.cuh"

struct pt { gf x, y, z, t; };            /* reference point_448.h:66-70 */
struct niels { gf a, b, c; };            /* a = y-x, b = y+x, c = 2d'xy   (goldilocks.c:55) */
struct pniels { niels n; gf z; };        /* projective niels              (goldilocks.c:56) */

GD void gf_load_factor(gf &f) {
    const uint32_t t[16] = GOLD_CONST_GF_FACTOR;
#pragma unroll
    for (int i = 0; i < 16; i++) f.v[i] = t[i];
}

GD void pt_copy(pt &o, const pt &a) { gf_copy(o.x, a.x); gf_copy(o.y, a.y); gf_copy(o.z, a.z); gf_copy(o.t, a.t); }
GD void pt_set_identity(pt &p) { /* goldilocks.c:83 */
    gf_set_zero(p.x); gf_set_ui(p.y, 1); gf_set_ui(p.z, 1); gf_set_zero(p.t);
}

// p = q + r (SUB = false) or q - r (SUB = true).  8M + 1w.  (goldilocks.c:205-230 / 178-203)
template <bool SUB>
GD void pt_addsub(pt &p, const pt &q, const pt &r) {
    gf a, b, c, d, px, py;
    gf_sub(b, q.y, q.x);
    if (!SUB) { gf_sub(c, r.y, r.x); gf_add_nr(d, r.y, r.x); }
    else      { gf_sub(d, r.y, r.x); gf_add_nr(c, r.y, r.x); }
    gf_mul(a, c, b);
    gf_add_nr(b, q.y, q.x);
    gf_mul(py, d, b);
    gf_mul(b, r.t, q.t);
    gf_mulw(px, b, 2 * GOLD_EFF_D);
    gf_add_nr(b, a, py);
    gf_sub(c, py, a);
    gf_mul(a, q.z, r.z);
    gf_add(a, a, a);
    if (!SUB) { gf_add_nr(py, a, px); gf_sub(a, a, px); }
    else      { gf_sub(py, a, px); gf_add_nr(a, a, px); }
    gf_mul(p.z, a, py);
    gf_mul(p.x, py, c);
    gf_mul(p.y, a, b);
    gf_mul(p.t, b, c);
}
GD void pt_add(pt &p, const pt &q, const pt &r) { pt_addsub<false>(p, q, r); }
GD void pt_sub(pt &p, const pt &q, const pt &r) { pt_addsub<true>(p, q, r); }

// p = 2q.  4S + 4M, or 4S + 3M when T is not needed.  (goldilocks.c:232-254 point_double_internal)
GD void pt_double(pt &p, const pt &q, bool before_double) {
    gf a, b, c, d, e, f;
    gf_sqr(c, q.x);
    gf_sqr(a, q.y);
    gf_add_nr(d, c, a);
    gf_add_nr(e, q.y, q.x);
    gf_sqr(b, e);
    gf_sub(b, b, d);
    gf_sub(e, a, c);
    gf_sqr(f, q.z);
    gf_add_nr(f, f, f);
    gf_sub(a, f, e);
    gf_mul(p.x, a, b);
    gf_mul(p.z, e, a);
    gf_mul(p.y, e, d);
    if (!before_double) gf_mul(p.t, b, d);
}

GD void pt_negate(pt &o, const pt &a) { /* goldilocks.c:260-268 */
    gf_neg(o.x, a.x); gf_copy(o.y, a.y); gf_copy(o.z, a.z); gf_neg(o.t, a.t);
}

GD void niels_cond_neg(niels &n, gmask_t neg) { /* goldilocks.c:271-278 */
    gf_cond_swap(n.a, n.b, neg);
    gf_cond_neg(n.c, neg);
}
GD void pt_to_pniels(pniels &b, const pt &a) { /* goldilocks.c:280-288 */
    gf_sub(b.n.a, a.y, a.x);
    gf_add(b.n.b, a.x, a.y);
    gf_mulw_signed(b.n.c, a.t, 2 * GOLD_TWISTED_D);
    gf_add(b.z, a.z, a.z);
}
GD void pniels_to_pt(pt &e, const pniels &d) { /* goldilocks.c:290-301 */
    gf eu, ey;
    gf_add_nr(eu, d.n.b, d.n.a);
    gf_sub(ey, d.n.b, d.n.a);
    gf_mul(e.t, ey, eu);
    gf_mul(e.x, d.z, ey);
    gf_mul(e.y, d.z, eu);
    gf_sqr(e.z, d.z);
}
GD void niels_to_pt(pt &e, const niels &n) { /* goldilocks.c:303-313 */
    gf ey, ex;
    gf_add(ey, n.b, n.a);
    gf_sub(ex, n.b, n.a);
    gf_mul(e.t, ey, ex);
    gf_copy(e.y, ey);
    gf_copy(e.x, ex);
    gf_set_ui(e.z, 1);
}
// d += e (SUB = false) / d -= e (SUB = true) for an affine niels e.  7M (6M without T).
// (goldilocks.c:315-359 add_niels_to_pt / sub_niels_from_pt)
template <bool SUB>
GD void pt_addsub_niels(pt &d, const niels &e, bool before_double) {
    gf a, b, c, dy;
    gf_sub(b, d.y, d.x);
    gf_mul(a, SUB ? e.b : e.a, b);
    gf_add_nr(b, d.x, d.y);
    gf_mul(dy, SUB ? e.a : e.b, b);
    gf_mul(d.x, e.c, d.t);
    gf_add_nr(c, a, dy);
    gf_sub(b, dy, a);
    if (!SUB) { gf_sub(dy, d.z, d.x); gf_add_nr(a, d.x, d.z); }
    else      { gf_add_nr(dy, d.z, d.x); gf_sub(a, d.z, d.x); }
    gf_mul(d.z, a, dy);
    gf_mul(d.x, dy, b);
    gf_mul(d.y, a, c);
    if (!before_double) gf_mul(d.t, b, c);
}
template <bool SUB>
GD void pt_addsub_pniels(pt &p, const pniels &pn, bool before_double) { /* goldilocks.c:361-380 */
    gf l0;
    gf_mul(l0, p.z, pn.z);
    gf_copy(p.z, l0);
    pt_addsub_niels<SUB>(p, pn.n, before_double);
}
// Runtime-signed variants for the variable-time (public data) paths: one code body, sign folded
// into the niels operand selection instead of two inlined bodies.
GD void pt_add_niels_signed(pt &d, const niels &e, gmask_t neg, bool before_double) {
    niels m;
    gf_copy(m.a, e.a); gf_copy(m.b, e.b); gf_copy(m.c, e.c);
    niels_cond_neg(m, neg);
    pt_addsub_niels<false>(d, m, before_double);
}

// (goldilocks.c:644-653) equality on the quotient group: compares x/y, insensitive to 2-torsion.
GD gmask_t pt_eq(const pt &p, const pt &q) {
    gf a, b;
    gf_mul(a, p.y, q.x);
    gf_mul(b, q.y, p.x);
    return gf_eq(a, b);
}
// (goldilocks.c:655-673)
GD gmask_t pt_valid(const pt &p) {
    gf a, b, c;
    gf_mul(a, p.x, p.y);
    gf_mul(b, p.z, p.t);
    gmask_t out = gf_eq(a, b);
    gf_sqr(a, p.x);
    gf_sqr(b, p.y);
    gf_sub(a, b, a);
    gf_sqr(b, p.t);
    gf_mulw_signed(c, b, GOLD_TWISTED_D);
    gf_sqr(b, p.z);
    gf_add(b, b, c);
    out &= gf_eq(a, b);
    out &= ~gf_is_zero(p.z);
    return out;
}

// ---------------------------------------------------------------------------------------------
// Decaf codec (goldilocks.c:98-176)
// ---------------------------------------------------------------------------------------------
// s = canonical decaf encoding of p as a field element (goldilocks.c:98-140 with all toggles 0;
// the two Elligator-inverse by-products are not needed on the hot path).
GD void pt_deisogenize(gf &s, const pt &p) {
    gf t1, t2, t3, t4, factor;
    gf_add_nr(t1, p.x, p.t);
    gf_sub(t2, p.x, p.t);
    gf_mul(t3, t1, t2);                      /* num = x^2 - t^2 */
    gf_sqr(t2, p.x);
    gf_mul(t1, t2, t3);
    gf_mulw(t2, t1, (uint32_t)(-1 - GOLD_TWISTED_D)); /* x^2 * (a-d) * num, a-d = 39081 */
    (void)gf_isr(t1, t2);                    /* isr */
    gf_mul(t2, t1, t3);                      /* ratio */
    gf_load_factor(factor);
    gf_mul(t4, t2, factor);
    gmask_t negx = gf_lobit(t4);
    gf_cond_neg(t2, negx);
    gf_mul(t3, t2, p.z);
    gf_sub(t3, t3, p.t);
    gf_mul(t2, t3, p.x);
    gf_mulw(t4, t2, (uint32_t)(-1 - GOLD_TWISTED_D));
    gf_mul(s, t4, t1);
    gf_cond_neg(s, gf_lobit(s));
}
// Full deisogenize (goldilocks.c:98-140) with the Elligator-inverse by-products: `sum` = ratio*z - t and
// `m1` = +-x + t; toggle_s / toggle_altx pick one of the four preimage branches (elligator.c:107-118).
GD void pt_deisogenize_full(gf &s, gf &sum, gf &m1, const pt &p, gmask_t toggle_s, gmask_t toggle_altx) {
    gf t1, t2, t3, t4, factor;
    gf_add_nr(t1, p.x, p.t);
    gf_sub(t2, p.x, p.t);
    gf_mul(t3, t1, t2);
    gf_sqr(t2, p.x);
    gf_mul(t1, t2, t3);
    gf_mulw(t2, t1, (uint32_t)(-1 - GOLD_TWISTED_D));
    (void)gf_isr(t1, t2);
    gf_mul(t2, t1, t3);
    gf_load_factor(factor);
    gf_mul(t4, t2, factor);
    gmask_t negx = gf_lobit(t4) ^ toggle_altx;
    gf_cond_neg(t2, negx);
    gf_mul(t3, t2, p.z);
    gf_sub(sum, t3, p.t);
    gf_mul(t2, sum, p.x);
    gf_mulw(t4, t2, (uint32_t)(-1 - GOLD_TWISTED_D));
    gf_mul(s, t4, t1);
    gmask_t lobs = gf_lobit(s);
    gf_cond_neg(s, lobs);
    gf_copy(m1, p.x);
    gf_cond_neg(m1, ~lobs ^ negx ^ toggle_s);
    gf_add(m1, m1, p.t);
}
// Elligator inverse (elligator.c:104-152): one of up to 8 preimages of p selected by `hint`; returns the
// success mask, `out` (canonical words) is always written.
GD gmask_t pt_invert_elligator_nonuniform(uint32_t out[14], const pt &p, uint32_t hint) {
    const gmask_t sgn_s = 0u - (hint & 1u), sgn_altx = 0u - ((hint >> 1) & 1u), sgn_r0 = 0u - ((hint >> 2) & 1u);
    gf a, b, c, one, zero;
    gf_set_ui(one, 1);
    gf_set_zero(zero);
    pt_deisogenize_full(a, b, c, p, sgn_s, sgn_altx);
    const gmask_t is_identity = gf_is_zero(p.t);
    gf_cond_sel(b, b, one, is_identity & sgn_altx);
    gf_cond_sel(c, c, one, is_identity & sgn_s & ~sgn_altx);
    gf_mulw_signed(a, b, GOLD_EDWARDS_D - 1);
    gf_add(b, a, b);
    gf_sub(a, a, c);
    gf_add(b, b, c);
    gf_cond_swap(a, b, sgn_s);
    gf_neg(c, b);                               /* gf_mul_qnr, qnr = -1 (field.h:84-90) */
    gf_mul(b, c, a);
    gmask_t succ = gf_isr(c, b);
    succ |= gf_is_zero(b);
    gf t;
    gf_mul(t, c, a);
    gf_cond_neg(t, sgn_r0 ^ gf_lobit(t));
    succ &= ~(gf_is_zero(t) & (sgn_r0 | sgn_s)); /* duplicate preimages of the identity */
    gf_to_words(out, t);
    return succ;
}
// (goldilocks.c:142-176) returns success mask; p is always written.
GD gmask_t pt_decode(pt &p, const uint32_t ser[14], gmask_t allow_identity) {
    gf s, s2, num, tmp, tmp2, ynum, isr, den, factor;
    gmask_t succ = gf_from_words(s, ser);
    succ &= allow_identity | ~gf_is_zero(s);
    succ &= ~gf_lobit(s);
    gf one;
    gf_set_ui(one, 1);
    gf_sqr(s2, s);
    gf_sub(den, one, s2);
    gf_add(ynum, one, s2);
    gf_mulw(num, s2, (uint32_t)(-4 * GOLD_TWISTED_D));
    gf_sqr(tmp, den);
    gf_add(num, tmp, num);
    gf_mul(tmp2, num, tmp);
    succ &= gf_isr(isr, tmp2);
    gf_mul(tmp, isr, den);
    gf_mul(p.y, tmp, ynum);
    gf_mul(tmp2, tmp, s);
    gf_add(tmp2, tmp2, tmp2);
    gf_mul(tmp, tmp2, isr);
    gf_mul(p.x, tmp, num);
    gf_load_factor(factor);
    gf_mul(tmp, tmp2, factor);
    gf_cond_neg(p.x, gf_lobit(tmp));
    gf_set_ui(p.z, 1);
    gf_mul(p.t, p.x, p.y);
    return succ;
}

// ---------------------------------------------------------------------------------------------
// RFC 8032 / RFC 7748 point codecs (goldilocks.c:905-1004, 1104-1115)
// ---------------------------------------------------------------------------------------------
// y-coordinate words (canonical) + sign bit of x, after the 4-isogeny back to the untwisted curve.
GD void pt_encode_like_eddsa(uint32_t yw[14], uint32_t &xsign, const pt &q) {
    gf x, y, z, t, u;
    gf_sqr(x, q.x);
    gf_sqr(t, q.y);
    gf_add(u, x, t);
    gf_add_nr(z, q.y, q.x);
    gf_sqr(y, z);
    gf_sub(y, y, u);
    gf_sub(z, t, x);
    gf_sqr(x, q.z);
    gf_add_nr(t, x, x);
    gf_sub(t, t, z);
    gf_mul(x, t, y);
    gf_mul(y, z, u);
    gf_mul(z, u, t);
    gf_invert(z, z);
    gf_mul(t, x, z);
    gf_mul(x, y, z);
    gf_to_words(yw, x);
    xsign = gf_lobit(t) & 1u;
}
// enc = 14 words of y (bit 447.. are data) + the 57th byte; returns success mask.
GD gmask_t pt_decode_like_eddsa(pt &p, const uint32_t yw[14], uint32_t last_byte) {
    gmask_t low = (last_byte & 0x80) ? ~0u : 0u;
    gf a, b, c, d, one, px, py, pz, pt_;
    gmask_t succ = gf_from_words(py, yw);
    succ &= ((last_byte & 0x7f) == 0) ? ~0u : 0u;
    gf_set_ui(one, 1);
    gf_sqr(px, py);
    gf_sub(pz, one, px);                         /* num = 1 - y^2 */
    gf_mulw_signed(pt_, px, GOLD_EDWARDS_D);     /* d y^2 */
    gf_sub(pt_, one, pt_);                       /* denom = 1 - d y^2 */
    gf_mul(px, pz, pt_);
    succ &= gf_isr(pt_, px);                     /* 1/sqrt(num*denom) */
    gf_mul(px, pt_, pz);                         /* sqrt(num/denom) */
    gf_cond_neg(px, gf_lobit(px) ^ low);
    gf_set_ui(pz, 1);
    /* 4-isogeny 2xy/(y^2-ax^2), (y^2+ax^2)/(2-y^2-ax^2) */
    gf_sqr(c, px);
    gf_sqr(a, py);
    gf_add(d, c, a);
    gf_add_nr(pt_, py, px);
    gf_sqr(b, pt_);
    gf_sub(b, b, d);
    gf_sub(pt_, a, c);
    gf_sqr(px, pz);
    gf_add_nr(pz, px, px);
    gf_sub(a, pz, d);
    gf_mul(p.x, a, b);
    gf_mul(p.z, pt_, a);
    gf_mul(p.y, pt_, d);
    gf_mul(p.t, b, d);
    return succ;
}
GD void pt_encode_like_x448(uint32_t uw[14], const pt &p) { /* goldilocks.c:1104-1115 */
    gf t, z, y;
    gf_invert(t, p.x);
    gf_mul(z, t, p.y);
    gf_sqr(y, z);
    gf_to_words(uw, y);
}

// ---------------------------------------------------------------------------------------------
// Elligator 2 hash-to-curve (elligator.c:32-94)
// ---------------------------------------------------------------------------------------------
GD void pt_from_hash_nonuniform(pt &p, const uint32_t ser[14]) {
    gf r0, r, a, b, c, N, e, one;
    (void)gf_from_words(r0, ser);
    gf_strong_reduce(r0);
    gf_set_ui(one, 1);
    gf_sqr(a, r0);
    gf_neg(r, a);                                   /* r = qnr * r0^2, qnr = -1 */
    gf_sub(a, r, one);
    gf_mulw_signed(b, a, GOLD_EDWARDS_D);           /* dr - d */
    gf_add(a, b, one);
    gf_sub(b, b, r);
    gf_mul(c, a, b);                                /* D = (dr+1-d)(dr-r-d) */
    gf_add(a, r, one);
    gf_mulw(N, a, (uint32_t)(1 - 2 * GOLD_EDWARDS_D)); /* N = (r+1)(1-2d) */
    gf_mul(a, c, N);
    gmask_t square = gf_isr(b, a);
    gf_cond_sel(c, r0, one, square);
    gf_mul(e, b, c);
    gf_mul(a, N, e);
    gf_cond_neg(a, gf_lobit(a) ^ ~square);          /* s */
    gf_mulw(c, e, (uint32_t)(1 - 2 * GOLD_EDWARDS_D));
    gf_sqr(b, c);
    gf_sub(e, r, one);
    gf_mul(c, b, e);
    gf_mul(b, c, N);
    gf_cond_neg(b, square);
    gf_sub(b, b, one);                              /* t */
    gf_sqr(c, a);
    gf_add(a, a, a);
    gf_add(e, c, one);
    gf_mul(p.t, a, e);
    gf_mul(p.x, a, b);
    gf_sub(a, one, c);
    gf_mul(p.y, e, a);
    gf_mul(p.z, a, b);
}
