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

GD uint64_t kc_rc(int r) {
    const uint64_t rc[24] = {
        0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull,
        0x000000000000808bull, 0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull,
        0x000000000000008aull, 0x0000000000000088ull, 0x0000000080008009ull, 0x000000008000000aull,
        0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull, 0x8000000000008003ull,
        0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,
        0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
    return rc[r];
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

// Streaming byte-wise SHAKE256 sponge.  Bytes are staged in a 136-byte word buffer (dynamic indexing
// -> local memory, L1 resident) and whole blocks are XORed into the register-resident state with
// compile-time lane indices; squeezing reads the same buffer.  Hashing is <1% of an EdDSA verify.
struct shake256_ctx {
    keccak_state st;
    uint32_t buf[SHAKE256_RATE / 4];
    int pos;
};
GD void kc_buf_clear(shake256_ctx &c) {
#pragma unroll
    for (int i = 0; i < SHAKE256_RATE / 4; i++) c.buf[i] = 0;
}
GD void shake256_init(shake256_ctx &c) {
#pragma unroll
    for (int i = 0; i < 25; i++) c.st.a[i] = 0;
    kc_buf_clear(c);
    c.pos = 0;
}
// state ^= buffered block; permute (reference shake.c:89-112 absorb + dokeccak)
GD void kc_absorb_block(shake256_ctx &c) {
#pragma unroll
    for (int i = 0; i < SHAKE256_RATE / 8; i++) c.st.a[i] ^= (uint64_t)c.buf[2 * i] | ((uint64_t)c.buf[2 * i + 1] << 32);
    keccak_f1600(c.st);
}
GD void kc_fill_output(shake256_ctx &c) {
#pragma unroll
    for (int i = 0; i < SHAKE256_RATE / 8; i++) {
        c.buf[2 * i] = (uint32_t)c.st.a[i];
        c.buf[2 * i + 1] = (uint32_t)(c.st.a[i] >> 32);
    }
}
GD void shake256_absorb_byte(shake256_ctx &c, uint8_t v) {
    c.buf[c.pos >> 2] |= (uint32_t)v << (8 * (c.pos & 3));
    if (++c.pos == SHAKE256_RATE) {
        kc_absorb_block(c);
        kc_buf_clear(c);
        c.pos = 0;
    }
}
GD void shake256_finish_absorb(shake256_ctx &c) { /* reference shake.c:136-142: pad 0x1f ... 0x80 */
    c.buf[c.pos >> 2] ^= 0x1fu << (8 * (c.pos & 3));
    c.buf[SHAKE256_RATE / 4 - 1] ^= 0x80000000u;
    kc_absorb_block(c);
    kc_fill_output(c);
    c.pos = 0;
}
GD uint8_t shake256_squeeze_byte(shake256_ctx &c) {
    if (c.pos == SHAKE256_RATE) {
        keccak_f1600(c.st);
        kc_fill_output(c);
        c.pos = 0;
    }
    const uint8_t r = (uint8_t)(c.buf[c.pos >> 2] >> (8 * (c.pos & 3)));
    c.pos++;
    return r;
}
