// keccak.cuh -- per-lane SHAKE256 (Keccak-f[1600], rate 136, suffix 0x1f), one sponge per GPU lane.
//
// Replaces the one-shot use the reference's EdDSA makes of src/shake.c (keccakf 60-87, absorb
// 89-112, pad+squeeze 114-162, SHAKE256 parameters 211-213).  The 25 lanes live in registers as
// 64-bit values (50 x 32-bit registers); all loops are fully unrolled so rho offsets and the pi
// permutation are compile-time, and 64-bit rotations become funnel shifts (SHF).
#pragma once
#include <stdint.h>
#include "gf.cuh"

#define SHAKE256_RATE 136

struct keccak_state { uint64_t a[25]; };

GD uint64_t kc_rol(uint64_t x, int s) { return s == 0 ? x : ((x << s) | (x >> (64 - s))); }

// Round constants: constant memory on the device (indexed by the public round counter), a plain table on the host.  As a
// local array inside the function they were copied to the stack of every lane and re-read with dynamic local loads.
#define KC_RC_TABLE                                                                                     \
    {0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull,       \
     0x000000000000808bull, 0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull,       \
     0x000000000000008aull, 0x0000000000000088ull, 0x0000000080008009ull, 0x000000008000000aull,       \
     0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull, 0x8000000000008003ull,       \
     0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,       \
     0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull}
#if defined(__CUDACC__)
static __constant__ uint64_t kc_rc_dev[24] = KC_RC_TABLE;
#endif
GD uint64_t kc_rc(int r) {
#if defined(__CUDA_ARCH__)
    return kc_rc_dev[r];
#else
    static const uint64_t rc[24] = KC_RC_TABLE;
    return rc[r];
#endif
}

// One round, lanes indexed a[x + 5y] (FIPS 202 section 3.2).
GD void kc_round(uint64_t a[25], uint64_t rc) {
    uint64_t c[5], d[5], b[25];
#pragma unroll
    for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
#pragma unroll
    for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ kc_rol(c[(x + 1) % 5], 1);
    /* rho offsets r[x][y] = (t+1)(t+2)/2 along the (x,y) -> (y, 2x+3y) orbit; pi: B[y][2x+3y] = rot(A[x][y]) */
    const int rho[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
#pragma unroll
    for (int y = 0; y < 5; y++)
#pragma unroll
        for (int x = 0; x < 5; x++) {
            const int src = x + 5 * y;
            const int dst = y + 5 * ((2 * x + 3 * y) % 5);
            b[dst] = kc_rol(a[src] ^ d[x], rho[src]);
        }
#pragma unroll
    for (int y = 0; y < 5; y++)
#pragma unroll
        for (int x = 0; x < 5; x++) a[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
    a[0] ^= rc;
}

GD void keccak_f1600(keccak_state &st) {
#pragma unroll 1
    for (int r = 0; r < 24; r++) kc_round(st.a, kc_rc(r));
}

// Streaming byte-wise SHAKE256 sponge, register resident.  The 25 state lanes stay in registers; incoming bytes are
// gathered into one 64-bit accumulator and XORed into lane pos/8 by a compile-time unrolled select chain (17 SELs per
// eight bytes), so there is NO dynamically indexed local array: round 1's 136-byte word buffer lived in local memory
// (29.9 M local loads and 664 MB of DRAM writes per 2^20 challenge hashes, profiles/r01i_scalars_ncu.txt) and made the
// position counter look data-dependent to a static audit (tools/ct_audit.py).  Positions are public (lengths), data only
// ever flows into `acc` and the state.  Fixed-size digests are read straight from the state words (shake256_out_words).
struct shake256_ctx {
    keccak_state st;
    uint64_t acc;   /* absorbing: the bytes of lane pos/8 gathered so far; squeezing: the lane being handed out */
    uint32_t pos;   /* byte position inside the current block, 0 .. 135 */
};
GD void shake256_init(shake256_ctx &c) {
#pragma unroll
    for (int i = 0; i < 25; i++) c.st.a[i] = 0;
    c.acc = 0;
    c.pos = 0;
}
// a[li] ^= v for a run-time lane index below the rate (branch-free, register only)
GD void kc_xor_lane(keccak_state &st, uint32_t li, uint64_t v) {
#pragma unroll
    for (int i = 0; i < SHAKE256_RATE / 8; i++) st.a[i] ^= (li == (uint32_t)i) ? v : 0ull;
}
GD uint64_t kc_get_lane(const keccak_state &st, uint32_t li) {
    uint64_t v = 0;
#pragma unroll
    for (int i = 0; i < SHAKE256_RATE / 8; i++) v |= (li == (uint32_t)i) ? st.a[i] : 0ull;
    return v;
}
GD void shake256_absorb_byte(shake256_ctx &c, uint8_t v) { /* reference shake.c:89-112 */
    c.acc |= (uint64_t)v << (8 * (c.pos & 7));
    c.pos++;
    if ((c.pos & 7) == 0) {
        kc_xor_lane(c.st, (c.pos >> 3) - 1, c.acc);
        c.acc = 0;
        if (c.pos == SHAKE256_RATE) {
            keccak_f1600(c.st);
            c.pos = 0;
        }
    }
}
GD void shake256_finish_absorb(shake256_ctx &c) { /* reference shake.c:136-142: pad 0x1f ... 0x80, permute, switch to squeezing */
    c.acc ^= (uint64_t)0x1f << (8 * (c.pos & 7));
    kc_xor_lane(c.st, c.pos >> 3, c.acc);
    c.st.a[SHAKE256_RATE / 8 - 1] ^= 0x8000000000000000ull;
    keccak_f1600(c.st);
    c.acc = 0;
    c.pos = 0;
}
// First NW 32-bit words of the output, straight from the state (call right after shake256_finish_absorb; NW * 4 <= 136).
template <int NW>
GD void shake256_out_words(const shake256_ctx &c, uint32_t *w) {
    static_assert(NW * 4 <= SHAKE256_RATE, "one block");
#pragma unroll
    for (int k = 0; k < NW; k++) w[k] = (uint32_t)(c.st.a[k >> 1] >> (32 * (k & 1)));
}
GD uint8_t shake256_squeeze_byte(shake256_ctx &c) { /* generic output lengths (goldilocks_shake256_hash_batch); reference shake.c:114-162 */
    if (c.pos == SHAKE256_RATE) {
        keccak_f1600(c.st);
        c.pos = 0;
    }
    if ((c.pos & 7) == 0) c.acc = kc_get_lane(c.st, c.pos >> 3);
    const uint8_t r = (uint8_t)(c.acc >> (8 * (c.pos & 7)));
    c.pos++;
    return r;
}
