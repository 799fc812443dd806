// rlc.cuh -- random-linear-combination batch verification of Ed448 signatures: an OPTIONAL fast path
// with per-element fallback.  Results are the reference's (eddsa.c:253-306) except with probability < 2^-127 per call.
//
// The reference accepts signature i iff  resp_i*B + chal_i*A_i == R_i  on the internal curve, where A_i, R_i come out of
// `point_decode_like_eddsa_and_mul_by_ratio` (goldilocks.c:949-1004), chal_i = -SHAKE256(dom || R || A || M) mod q and
// resp_i = S mod q (eddsa.c:273-291), and `==` is `point_eq` (goldilocks.c:644-653).  Every decoded point is the image of
// a curve point under the 4-isogeny whose kernel is the whole 4-torsion of Ed448, so all of these points live in the
// subgroup of prime order q and `point_eq` on them is plain equality up to the 2-torsion point the isogeny never hits.
// With secret random odd weights z_i of at least 128 bits (drawn per call, after the signatures are fixed):
//
//     sum_i z_i*(-R_i)  +  sum_keys (sum_{i under key} z_i*chal_i mod q)*A_key  +  (sum_i z_i*resp_i mod q)*B  ==  identity
//
// holds if every signature is valid, and fails with probability >= 1 - 2^-127 if any is not (a non-zero element of a
// group of prime order times a uniform odd weight of >= 128 bits).  Signatures whose R or public key does not decode are
// rejected up front, exactly like the reference (eddsa.c:266-270), and take no part in the sum.  When the equation
// fails the caller runs the ordinary per-signature path over the batch, so statuses are always per element.
//
// The sum is a multi-scalar multiplication by the bucket method, in two classes that share nothing but the final addition:
// the n points R_i with their weights (135 bits for c = 15: whole windows only) and the distinct keys plus B with 446-bit
// scalars.  Per class: every scalar is cut into c-bit digits (c from the class's own size: about 32 resp. 64 points per
// bucket), the pairs (window, digit) -> point are radix-sorted, one lane adds up each bucket on the slot machine
// (slots.cuh; points are stored as projective-niels records, 8M per addition), and the buckets of a window are folded
// with running sums over segments of 32, a tree level of 32 and c*w doublings per window.  The short top window of the
// 446-bit class gets sub-buckets so that no bucket is longer than the rest.  A signature costs its R decode (one inverse
// square root), ceil(128 / c) = 9 point additions and a share of the per-key work, instead of the 90 + 30 additions and
// 40 doublings of the table path (slot_lanes.cuh SlotEdVerifyFinishShared).  Scheduling (two streams) is in abi.cu rlc_core.
//
// Everything below is a per-lane functor (host/device clean: tests/hostsim runs the same code on the CPU tier).
#pragma once
#include "lanes.cuh"
#include "slot_algos.cuh"

#define RLC_ZWORDS 5          /* weights: w1 * c bits, 128 <= bits < 143, so that every window of a weight is a full one */
#define RLC_ZBITS 128
#define RLC_Z_PER_LANE 6      /* weights per Keccak-f: 6 x 20 bytes of one 136-byte block */
#define RLC_SEG 32            /* buckets per running-sum segment, segments per tree node */
#define RLC_SCELLS 1024       /* accumulator cells of the response sum (spreads the atomics) */
#define RLC_ACC_WORDS 14      /* one 64-bit cell per 32-bit scalar word: sums of up to 2^32 words never overflow */
#define RLC_MAX_C 15

// Key groups of a batch (k_group.cu group_keys_all): order[j] = signature at sorted position j, gid[j] = 1-based group
// of that position, gstart[g] = first sorted position of group g.
struct rlc_groups { const uint32_t *order, *gid, *gstart; uint32_t ngroups; };

struct rlc_shape { /* one per class of points, chosen on the host from n (rlc_shape_for) */
    uint32_t c;        /* digit bits */
    uint32_t wn;       /* windows of this class: ceil(128 / c) for the weights of the R points, ceil(446 / c) for keys and B */
    uint32_t seg;      /* min(RLC_SEG, 2^c) buckets per segment */
    uint32_t segs;     /* segments per window = 2^c / seg */
    uint32_t nodes;    /* tree nodes per window = ceil(segs / RLC_SEG) */
    uint32_t zbits;    /* bits of a weight: ceil(128 / c) * c, so that every window of a weight is a full one */
    uint32_t top;      /* index of the one window with fewer than c bits (the last window of a 446-bit scalar), or ~0 */
    uint32_t top_shift; /* that window holds r = 446 - top * c bits: its 2^r digits get 2^(c - r) sub-buckets each (picked by the low
                         * bits of the point index), so that no bucket of it is longer than the others' */
    uint32_t nch;      /* chunks: every chunk of `csize` consecutive signatures has its own equation (own buckets, sums, verdict), so a
                        * bad signature sends only its chunk to the per-signature path.  Window ids run over nch * wn. */
    uint32_t csize;    /* signatures per chunk */
};
// is_key = 0: the n points -R_i with their weights; 1: the keys and B with 446-bit scalars.  Each class picks its digit
// width from its own number of points (about 32 points per bucket).
static inline rlc_shape rlc_shape_for(size_t count, int force_c, int is_key) {
    rlc_shape s;
    int c = 0;
    while (c < 31 && ((size_t)2 << c) <= count) c++; /* floor(log2 count) */
    c -= is_key ? 6 : 5;   /* about 32 points per bucket; 64 for the key class: half the lanes, one wave of the machine beside the R class */
    if (c < 2) c = 2;
    if (c > RLC_MAX_C) c = RLC_MAX_C;
    if (force_c > 0) c = force_c;
    s.c = (uint32_t)c;
    s.zbits = ((RLC_ZBITS + c - 1) / c) * c;
    s.wn = is_key ? (GOLDILOCKS_SCALAR_BITS_ + c - 1) / c : s.zbits / c;
    s.seg = (1u << c) < RLC_SEG ? (1u << c) : RLC_SEG;
    s.segs = (1u << c) / s.seg;
    s.nodes = (s.segs + RLC_SEG - 1) / RLC_SEG;
    s.top = is_key ? s.wn - 1 : ~0u;
    s.top_shift = is_key ? s.wn * s.c - GOLDILOCKS_SCALAR_BITS_ : 0u;
    s.nch = 1;
    s.csize = ~0u;
    return s;
}
#define RLC_SCELLS_CH 64 /* accumulator cells of the response sum per chunk when there is more than one chunk */
GD uint32_t rlc_scells(const rlc_shape &sh) { return sh.nch > 1 ? RLC_SCELLS_CH : RLC_SCELLS; }

GD void rlc_atomic_add(unsigned long long *p, unsigned long long v) {
#if defined(__CUDA_ARCH__)
    atomicAdd(p, v);
#else
    __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
#endif
}
GD void pt_ld(pt &o, const pt *p) { gf_ld<false>(o.x, &p->x); gf_ld<false>(o.y, &p->y); gf_ld<false>(o.z, &p->z); gf_ld<false>(o.t, &p->t); }
GD void pt_st(pt *p, const pt &a) { gf_copy(p->x, a.x); gf_copy(p->y, a.y); gf_copy(p->z, a.z); gf_copy(p->t, a.t); }
GD gmask_t gf_is_zero_mod_p(const gf &a) { gf z; gf_set_zero(z); return gf_eq(a, z); }

// 1) decode: lane j < n: R_j; lane n + k (k < ngroups): the key of group k; lane n + ngroups: the base point B.
//    (`lane0` = first lane of this launch: the R halves follow the copies, the keys go first.)
//    A decoded point with Z = 0 cannot occur for a point of the curve; if one ever shows up the whole call falls back.
struct LaneRlcDecode {
    pt *pts; int32_t *ok; uint32_t *force_fallback; const uint8_t *sig, *pk; size_t n; rlc_groups g; size_t lane0;
    GDM void operator()(size_t j0) const {
        const size_t j = j0 + lane0;
        pt p;
        gmask_t good;
        if (j >= n + g.ngroups) { /* one copy of B per chunk */
            const uint32_t bw[14] = GOLD_CONST_BASE_WORDS;
            good = pt_decode(p, bw, 0);
        } else {
            const uint8_t *enc = j < n ? sig + 114 * j : pk + 57 * (size_t)g.order[g.gstart[j - n]];
            uint32_t w[15];
            words_load_bytes(w, 15, enc, 57);
            good = pt_decode_like_eddsa(p, w, w[14] & 0xff);
        }
        if (good && gf_is_zero_mod_p(p.z)) *force_fallback = 1u;
        /* every point is stored as the projective-niels record the slot machine adds from (slot_algos.cuh
         * s_pt_to_pniels_negc_g): (y - x, y + x, -2 d' t, 2 z); the bucket kernel SUBTRACTS the R records */
        pt rec;
        gf_sub(rec.x, p.y, p.x);
        gf_add_nr(rec.y, p.y, p.x);
        gf_mulw(rec.z, p.t, (uint32_t)(-2 * GOLD_TWISTED_D));
        gf_add_nr(rec.t, p.z, p.z);
        pt_st(pts + j, rec);
        ok[j] = ST_OK(good);
    }
};

// 2) weights: one Keccak-f per 6 signatures: SHAKE256("b200-rlc" || seed32 || le64(l)) -> z_{6l} .. z_{6l+5}, cut to zbits, made odd.
struct LaneRlcZ {
    uint32_t *z; const uint8_t *seed32; size_t n; uint32_t zbits;
    GDM void operator()(size_t l) const {
        shake256_ctx h;
        shake256_init(h);
        const uint8_t tag[8] = {'b', '2', '0', '0', '-', 'r', 'l', 'c'};
        for (int k = 0; k < 8; k++) shake256_absorb_byte(h, tag[k]);
        for (int k = 0; k < 32; k++) shake256_absorb_byte(h, seed32[k]);
        for (int k = 0; k < 8; k++) shake256_absorb_byte(h, (uint8_t)((uint64_t)l >> (8 * k)));
        shake256_finish_absorb(h);
        uint32_t ow[RLC_Z_PER_LANE * RLC_ZWORDS]; /* one output block, read straight from the state words */
        shake256_out_words<RLC_Z_PER_LANE * RLC_ZWORDS>(h, ow);
#pragma unroll
        for (int e = 0; e < RLC_Z_PER_LANE; e++) {
            uint32_t w[RLC_ZWORDS];
#pragma unroll
            for (int k = 0; k < RLC_ZWORDS; k++) {
                uint32_t x = ow[e * RLC_ZWORDS + k];
                const int lo = 32 * k; /* keep bits [0, zbits) */
                if ((int)zbits <= lo) x = 0;
                else if ((int)zbits < lo + 32) x &= (1u << (zbits - lo)) - 1u;
                w[k] = x;
            }
            w[0] |= 1u;
            const size_t i = RLC_Z_PER_LANE * l + e;
            if (i < n) for (int k = 0; k < RLC_ZWORDS; k++) z[RLC_ZWORDS * i + k] = w[k];
        }
    }
};

// 3) products: lane j = sorted position.  Excluded signatures get valid = FAILURE (the bucket kernel skips them) and add nothing to the scalar sums.
//    `early` (the whole-batch pass): the sums are taken as soon as the KEY decodes and the challenge hashes are in -- a signature counts
//    if its key decodes -- so that the whole key class runs beside the R decodes; LaneRlcLate settles the rest when those are in.
struct LaneRlcWeights {
    uint32_t *z; int32_t *valid; unsigned long long *key_acc, *s_acc; const abi_sc *chal, *resp; const int32_t *ok; size_t n; rlc_groups g; rlc_shape sh; /* R-class shape: chunks */
    uint32_t early;
    GDM void operator()(size_t j) const {
        const size_t i = g.order[j], k = g.gid[j] - 1;
        const bool v = early ? ok[n + k] != 0 : (ok[i] && ok[n + k]);
        if (!early) {
            valid[i] = v ? -1 : 0;
            if (!v) { for (int q = 0; q < RLC_ZWORDS; q++) z[RLC_ZWORDS * i + q] = 0; return; }
        } else if (!v) return;
        sc zi, c, r, zc, zr;
        sc_set_zero(zi);
        for (int q = 0; q < RLC_ZWORDS; q++) zi.w[q] = z[RLC_ZWORDS * i + q];
        sc_from_abi(c, chal + i);
        sc_from_abi(r, resp + i);
        sc_mul(zc, zi, c);
        sc_mul(zr, zi, r);
        const uint32_t cells = rlc_scells(sh);
        unsigned long long *ka = key_acc + RLC_ACC_WORDS * k, *sa = s_acc + RLC_ACC_WORDS * ((i / sh.csize) * cells + j % cells);
        for (int q = 0; q < SC_WORDS; q++) { rlc_atomic_add(ka + q, zc.w[q]); rlc_atomic_add(sa + q, zr.w[q]); }
    }
};

// 3b) after the R decodes: the verdicts the bucket kernel and the caller read (valid = R and key both decode), and the correction of the
//     early sums -- a signature whose key decodes but whose R does not was counted by LaneRlcWeights(early) and is taken out again
//     (the cells are plain integers mod 2^64: subtracting is adding the complement).  Any such signature raises *redo: the key class that
//     ran on the early sums is then run again on the corrected ones (abi.cu rlc_core).  Never on honest traffic.
struct LaneRlcLate {
    uint32_t *z; int32_t *valid; unsigned long long *key_acc, *s_acc; const abi_sc *chal, *resp; const int32_t *ok; size_t n; rlc_groups g; rlc_shape sh; uint32_t *redo;
    GDM void operator()(size_t j) const {
        const size_t i = g.order[j], k = g.gid[j] - 1;
        const bool okk = ok[n + k] != 0, okr = ok[i] != 0;
        valid[i] = (okk && okr) ? -1 : 0;
        if (okk && okr) return;
        if (okk) {
            sc zi, c, r, zc, zr;
            sc_set_zero(zi);
            for (int q = 0; q < RLC_ZWORDS; q++) zi.w[q] = z[RLC_ZWORDS * i + q];
            sc_from_abi(c, chal + i);
            sc_from_abi(r, resp + i);
            sc_mul(zc, zi, c);
            sc_mul(zr, zi, r);
            const uint32_t cells = rlc_scells(sh);
            unsigned long long *ka = key_acc + RLC_ACC_WORDS * k, *sa = s_acc + RLC_ACC_WORDS * ((i / sh.csize) * cells + j % cells);
            for (int q = 0; q < SC_WORDS; q++) { rlc_atomic_add(ka + q, 0ull - (unsigned long long)zc.w[q]); rlc_atomic_add(sa + q, 0ull - (unsigned long long)zr.w[q]); }
            *redo = 1u;
        }
        for (int q = 0; q < RLC_ZWORDS; q++) z[RLC_ZWORDS * i + q] = 0;
    }
};

// 4) per-key scalars: carry-propagate the accumulator cells (an integer below 2^(448 + 32)) and reduce mod q.
//    lane k < ngroups: key k; lane ngroups: the response sum, scalar of B.
struct LaneRlcKeyScalars {
    uint32_t *kscal; const unsigned long long *key_acc, *s_acc; uint32_t ngroups; uint32_t cells; /* lane ngroups + ch: the response sum of chunk ch */
    GDM void operator()(size_t k) const {
        unsigned long long cell[RLC_ACC_WORDS];
        if (k < ngroups) {
            for (int q = 0; q < RLC_ACC_WORDS; q++) cell[q] = key_acc[RLC_ACC_WORDS * k + q];
        } else {
            const unsigned long long *base = s_acc + (size_t)RLC_ACC_WORDS * cells * (k - ngroups);
            for (int q = 0; q < RLC_ACC_WORDS; q++) cell[q] = 0;
            for (uint32_t s = 0; s < cells; s++)
                for (int q = 0; q < RLC_ACC_WORDS; q++) cell[q] += base[RLC_ACC_WORDS * s + q]; /* < 2^10 * 2^52 */
        }
        uint32_t w[28]; /* 112 bytes, two 56-byte chunks for sc_decode_long */
        unsigned long long carry = 0;
        for (int q = 0; q < 28; q++) {
            if (q < RLC_ACC_WORDS) carry += cell[q];
            w[q] = (uint32_t)carry;
            carry >>= 32;
        }
        sc out;
        ByteAtWords at = {w};
        sc_decode_long(out, at, 112);
        for (int q = 0; q < SC_WORDS; q++) kscal[SC_WORDS * k + q] = out.w[q];
    }
};

// 5) digits: lane p = point index ([0, n): -R with its 128-bit weight; [n, n + ngroups]: keys and B with 446-bit scalars).
//    Pair (window << c | digit) -> p; digit 0 goes to the sentinel key (sorted last, ignored).  One pair list per class.
GD uint32_t rlc_bits(const uint32_t *w, int nwords, uint32_t pos, uint32_t nbits) {
    const uint32_t wi = pos >> 5;
    const uint64_t lo = wi < (uint32_t)nwords ? w[wi] : 0u, hi = wi + 1 < (uint32_t)nwords ? w[wi + 1] : 0u;
    return (uint32_t)(((hi << 32) | lo) >> (pos & 31)) & ((1u << nbits) - 1u);
}
struct LaneRlcDigits { /* lane l = l-th point of the class: point index p0 + l, scalar = nwords words at scal + nwords * l */
    uint32_t *keys, *vals; const uint32_t *scal; uint32_t nwords; size_t p0; rlc_shape sh;
    const uint32_t *chunk_of; /* key class: chunk of the l-th point (null: l / csize, the R class) */
    GDM void operator()(size_t l) const {
        const uint32_t *w = scal + (size_t)nwords * l;
        const size_t off = l * sh.wn;
        const uint32_t sentinel = (sh.nch * sh.wn) << sh.c, p = (uint32_t)(p0 + l);
        const uint32_t w0 = (chunk_of ? chunk_of[l] : (uint32_t)(l / sh.csize)) * sh.wn;
        for (uint32_t k = 0; k < sh.wn; k++) {
            uint32_t d = rlc_bits(w, (int)nwords, k * sh.c, sh.c);
            if (d && k == sh.top) d = (d << sh.top_shift) | ((uint32_t)l & ((1u << sh.top_shift) - 1u)); /* sub-bucket */
            keys[off + k] = d ? (((w0 + k) << sh.c) | d) : sentinel;
            vals[off + k] = p;
        }
    }
};

GD uint32_t rlc_bucket_digit(size_t b, const rlc_shape &sh) {
    const uint32_t low = (uint32_t)b & ((1u << sh.c) - 1u);
    return (uint32_t)(b >> sh.c) % sh.wn == sh.top ? (low >> sh.top_shift) : low;
}
// 6) buckets: lane b = (window << c | digit) adds up (or, for the R class, subtracts) the points of its run in the sorted
//    pair list, on the slot machine (slots.cuh): accumulator in shared-memory slots, each point taken straight from its
//    projective-niels record in global memory (8M per point, no by-value calls).
// 5b) bucket runs: lane b finds the run of bucket b in the sorted pair list and files the bucket under its LENGTH.  The bucket kernel
//     then takes its buckets in order of length (a second, 8-bit radix sort of (255 - length, bucket)), so the 32 lanes of a warp add up
//     runs of nearly the same length: with one lane per bucket in bucket order a warp waited for its longest run -- Poisson around 32,
//     the longest of 32 is about 46 -- and the kernel sat at 52 % of the multiply pipe (profiles/r02h_rlcbucket_ncu.txt).
struct LaneRlcBucketRuns {
    uint32_t *start, *lenkey, *ident; const uint32_t *keys; size_t npairs; rlc_shape sh;
    GDM void operator()(size_t b) const {
        uint32_t lo = 0, len = 0;
        if (rlc_bucket_digit(b, sh)) {
            size_t l = 0, h = npairs;
            while (l < h) { const size_t mid = (l + h) >> 1; if (keys[mid] < (uint32_t)b) l = mid + 1; else h = mid; }
            size_t l2 = l;
            h = npairs;
            while (l2 < h) { const size_t mid = (l2 + h) >> 1; if (keys[mid] <= (uint32_t)b) l2 = mid + 1; else h = mid; }
            lo = (uint32_t)l; len = (uint32_t)(l2 - l);
        }
        start[b] = lo;
        lenkey[b] = 255u - (len > 255u ? 255u : len);
        ident[b] = (uint32_t)b;
    }
};
struct SlotRlcBucket {
    static constexpr int NSLOTS = 7;
    pt *buckets; const uint32_t *keys, *vals; size_t npairs; const pt *recs; rlc_shape sh; gmask_t subtract;
    const int32_t *valid; /* R class: the pair list is made before the decodes are in, so excluded signatures are skipped here (null: take all) */
    const uint32_t *perm, *start; /* lane l takes bucket perm[l], whose run begins at start[bucket] (LaneRlcBucketRuns) */
    GDM void operator()(size_t lane, sref sb, bool live) const {
        const spt p = {s_slot(sb, 0), s_slot(sb, 1), s_slot(sb, 2), s_slot(sb, 3)};
        const swk w = {s_slot(sb, 4), s_slot(sb, 5), s_slot(sb, 6)};
        s_pt_set_identity(p);
        if (!live) return;
        const size_t b = perm[lane];
        if (rlc_bucket_digit(b, sh)) {
            {
            for (size_t j = start[b]; j < npairs && keys[j] == (uint32_t)b; j++) {
                if (valid && !valid[vals[j]]) continue;
                wtab<1> t;
                t.base = reinterpret_cast<uint4 *>(const_cast<pt *>(recs + vals[j]));
                s_pt_add_pniels_g<1>(p, w, t, 0, subtract, ~subtract, false); /* minus the point: swap (a, b), keep the stored -c */
            }
            }
        }
        pt out;
        s_ld(out.x, p.x); s_ld(out.y, p.y); s_ld(out.z, p.z); s_ld(out.t, p.t);
        pt_st(buckets + b, out);
    }
};

GD void pt_mul_small(pt &out, const pt &p, uint32_t k) { /* 0 < k < 2^16, public; double-and-add from the top set bit */
    pt acc, t;
    pt_copy(acc, p);
    int bit = 15;
    while (bit > 0 && !((k >> bit) & 1u)) bit--;
    for (bit--; bit >= 0; bit--) {
        pt_double(t, acc, false); pt_copy(acc, t);
        if ((k >> bit) & 1u) { pt_add(t, acc, p); pt_copy(acc, t); }
    }
    pt_copy(out, acc);
}

// 7) segments: lane s covers buckets [s*seg, (s+1)*seg) of one window.  Bucket `low` of a window carries the digit
//    d(low) = low >> shift (shift = top_shift in the last window, else 0), so
//    out = sum_k d(base+k) * bucket_{base+k} = d(base) * sum_k bucket_{base+k} + sum over the k at which d steps up of
//    (sum_{k' >= k} bucket_{base+k'})  -- running sums from the top.
struct LaneRlcSegments {
    pt *segsum; const pt *buckets; rlc_shape sh;
    GDM void operator()(size_t s) const {
        const uint32_t base = (uint32_t)((s * sh.seg) & ((1u << sh.c) - 1u));
        const uint32_t shift = (uint32_t)((s * sh.seg) >> sh.c) % sh.wn == sh.top ? sh.top_shift : 0u;
        const pt *bk = buckets + s * sh.seg;
        pt run, acc, t, q;
        pt_set_identity(run);
        pt_set_identity(acc);
        for (uint32_t k = sh.seg; k-- > 0;) {
            pt_ld(q, bk + k);
            pt_add(t, run, q); pt_copy(run, t);
            if (k && ((base + k) >> shift) != ((base + k - 1) >> shift)) { pt_add(t, acc, run); pt_copy(acc, t); }
        }
        if (base >> shift) {
            pt_mul_small(q, run, base >> shift);
            pt_add(t, acc, q); pt_copy(acc, t);
        }
        pt_st(segsum + s, acc);
    }
};

// 8) tree: lane (w, node) adds up to RLC_SEG segment sums of window w.
struct LaneRlcNodes {
    pt *nodesum; const pt *segsum; rlc_shape sh;
    GDM void operator()(size_t l) const {
        const size_t w = l / sh.nodes, node = l % sh.nodes;
        const size_t lo = node * RLC_SEG, hi = lo + RLC_SEG < sh.segs ? lo + RLC_SEG : sh.segs;
        pt acc, t, q;
        pt_ld(acc, segsum + w * sh.segs + lo);
        for (size_t k = lo + 1; k < hi; k++) { pt_ld(q, segsum + w * sh.segs + k); pt_add(t, acc, q); pt_copy(acc, t); }
        pt_st(nodesum + l, acc);
    }
};
// 9) windows: lane w adds the nodes of its window and doubles c*w times.
struct LaneRlcWindows {
    pt *winsum; const pt *nodesum; rlc_shape sh;
    GDM void operator()(size_t w) const {
        pt acc, t, q;
        pt_ld(acc, nodesum + w * sh.nodes);
        for (size_t k = 1; k < sh.nodes; k++) { pt_ld(q, nodesum + w * sh.nodes + k); pt_add(t, acc, q); pt_copy(acc, t); }
        /* one equation: every window lane doubles its own sum into place (the lanes run side by side, the call waits for the
         * longest chain either way).  Many chunks: that would be c * wn^2 / 2 doublings per chunk -- LaneRlcTotal walks the windows
         * of a chunk from the top instead (Horner: 446 doublings per chunk in all). */
        const uint32_t dbl = sh.nch > 1 ? 0u : ((uint32_t)w % sh.wn) * sh.c; /* w = chunk * wn + window */
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (uint32_t k = 0; k < dbl; k++) { pt_double(t, acc, false); pt_copy(acc, t); }
        pt_st(winsum + w, acc);
    }
};
// 10) class totals: lane ch adds the window sums of chunk ch (the key class does this on the side stream, off the critical path).
struct LaneRlcTotal {
    pt *total; const pt *winsum; uint32_t wn; uint32_t horner_c; /* 0: the window sums are already in place; c: sum_w 2^(c w) winsum[w] from the top */
    GDM void operator()(size_t ch) const {
        const pt *ws = winsum + ch * wn;
        pt acc, t, q;
        if (horner_c) {
            pt_ld(acc, ws + (wn - 1));
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
            for (uint32_t w = wn - 1; w-- > 0;) {
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
                for (uint32_t k = 0; k < horner_c; k++) { pt_double(t, acc, false); pt_copy(acc, t); }
                pt_ld(q, ws + w); pt_add(t, acc, q); pt_copy(acc, t);
            }
        } else {
            pt_ld(acc, ws);
            for (uint32_t w = 1; w < wn; w++) { pt_ld(q, ws + w); pt_add(t, acc, q); pt_copy(acc, t); }
        }
        pt_st(total + ch, acc);
    }
};
// 11) verdicts: the equation of chunk ch holds iff its two class totals add up to the identity of the quotient group
//     (point_eq against (0, 1): X == 0, goldilocks.c:644-653) and nothing asked for the fallback.
struct LaneRlcVerdict {
    uint32_t *verdict; const pt *total_r, *total_k; const uint32_t *force_fallback;
    GDM void operator()(size_t ch) const {
        pt acc, a, b, id;
        pt_ld(a, total_r + ch);
        pt_ld(b, total_k + ch);
        pt_add(acc, a, b);
        pt_set_identity(id);
        const gmask_t same = pt_eq(acc, id) & ~gf_is_zero_mod_p(acc.z);
        verdict[ch] = (same && !*force_fallback) ? 1u : 0u;
    }
};
// 12) localisation: when the whole-batch equation fails, abi.cu rlc_core runs the equations again per CHUNK of consecutive
//     signatures (same R decodes, challenges and weights) and re-verifies only the chunks that fail, one signature at a time.
//     Their signatures are packed into one contiguous batch first (a batch of a few thousand signatures would leave the
//     machine nearly empty): fc[q] = q-th failed chunk in ascending order, so a short last chunk comes last.
struct LaneRlcPackPlan { /* one lane: where the messages of each failed chunk start in the packed arena; mbase[nf] = their total */
    size_t *mbase; const uint32_t *fc; uint32_t nf; const size_t *off; size_t n; uint32_t csize;
    GDM void operator()(size_t) const {
        size_t acc = 0;
        for (uint32_t q = 0; q < nf; q++) {
            const size_t lo = (size_t)fc[q] * csize, hi = lo + csize < n ? lo + csize : n;
            mbase[q] = acc;
            acc += off[hi] - off[lo];
        }
        mbase[nf] = acc;
    }
};
struct LaneRlcPack { /* lane j < np: packed signature j; lane np: the closing offset */
    uint8_t *psig, *ppk, *pmsg; size_t *poff; uint32_t *src;
    const uint8_t *sig, *pk, *msg; const size_t *off, *mbase; const uint32_t *fc; uint32_t nf; size_t np; uint32_t csize;
    GDM void operator()(size_t j) const {
        if (j == np) { poff[np] = mbase[nf]; return; }
        const uint32_t q = (uint32_t)(j / csize);
        const size_t lo = (size_t)fc[q] * csize, i = lo + j % csize;
        src[j] = (uint32_t)i;
        for (int b = 0; b < 114; b++) psig[114 * j + b] = sig[114 * i + b];
        for (int b = 0; b < 57; b++) ppk[57 * j + b] = pk[57 * i + b];
        const size_t mo = mbase[q] + (off[i] - off[lo]);
        poff[j] = mo;
        for (size_t b = off[i]; b < off[i + 1]; b++) pmsg[mo + (b - off[i])] = msg[b];
    }
};
struct LaneRlcUnpack { /* statuses of the packed batch back to their places */
    int32_t *dst; const int32_t *pst; const uint32_t *src;
    GDM void operator()(size_t j) const { dst[src[j]] = pst[j]; }
};
