// slots.cuh -- the "slot machine": field elements of the long-running scalar multiplications live in
// a per-lane register file in SHARED MEMORY, and every field operation is a small non-inlined
// device function that takes slot handles (d, a, b) instead of 16-limb structs.
//
// Why (measured on B200, profiles/r01h_*): with by-value calls (gf.cuh gf_mul_fn) a third of the
// caller's instructions were IMAD.MOV register shuffles in front of each call -- they run on the same
// multiply pipe as IMAD.WIDE -- and the 7..12 live field elements of a ladder/double-and-add step
// pinned the kernels at 255 registers + spills = 2 warps per scheduler.  With slots
//   * a call moves three 32-bit handles, the operands are fetched with 8 LDS.128 on the otherwise
//     idle load/store pipe and the result leaves with 4 STS.128;
//   * conditional swaps / conditional negations of table entries become per-lane HANDLE selections
//     (the data never moves), so the masked-swap ALU work of the ladder disappears;
//   * the kernels need ~100 registers, no local memory, and their whole code (one body per field
//     operation + a list of calls) stays inside the 32 KB instruction cache.
// Layout: slot s, quad q (4 limbs), lane t of a block of SLOT_BLOCK lanes sits at
// slot_mem[(4*s + q) * SLOT_BLOCK + t] (uint4), so every LDS.128/STS.128 of a warp is conflict-free.
// A handle is the absolute shared-state-space byte address of the lane's quad 0, so the four transfers
// of an element are ld/st.shared.v4 with immediate offsets and a call does no address arithmetic
// (measured: -1.3 % on every slot kernel against indexing the extern array).
//
// On the host (tests/hostsim) a handle is a plain pointer into a per-worker array of gf, so the same
// algorithm source is checked against the oracle on the CPU tier.
#pragma once
#include "gf.cuh"

#define SLOT_BLOCK 128

#if !defined(__CUDACC__)
struct alignas(16) uint4 { uint32_t x, y, z, w; }; /* host simulator only */
#endif

// Field elements in GLOBAL memory are four 128-bit quads, QS quads apart:
//   QS = 1  : contiguous 64 bytes (`gf` arrays: fixed-base tables, lane-contiguous window tables)
//   QS = 32 : warp-interleaved -- quad q of the 32 lanes of a warp is one 512-byte row, so a table
//             access in which all lanes touch the same entry is fully coalesced (constant-time scans).
template <bool RO, int QS>
GD void gq_ld(gf &o, const uint4 *p) { /* RO: never written while the kernel runs -> non-coherent path */
#pragma unroll
    for (int q = 0; q < 4; q++) {
#if defined(__CUDA_ARCH__)
        const uint4 x = RO ? __ldg(p + q * QS) : p[q * QS];
#else
        const uint4 x = p[q * QS];
#endif
        o.v[4 * q] = x.x; o.v[4 * q + 1] = x.y; o.v[4 * q + 2] = x.z; o.v[4 * q + 3] = x.w;
    }
}
template <int QS>
GD void gq_st(uint4 *p, const gf &x) {
#pragma unroll
    for (int q = 0; q < 4; q++) { uint4 v; v.x = x.v[4 * q]; v.y = x.v[4 * q + 1]; v.z = x.v[4 * q + 2]; v.w = x.v[4 * q + 3]; p[q * QS] = v; }
}
GD const uint4 *gq(const gf *g) { return reinterpret_cast<const uint4 *>(g); }

#if defined(__CUDA_ARCH__)
extern __shared__ uint4 slot_mem[];
#define SFN static __device__ __noinline__
// Handle = absolute shared-state-space BYTE address of this lane's quad 0: loads and stores are
// ld/st.shared with immediate offsets, no per-call address arithmetic.
struct sref { uint32_t a; };
#define SLOT_STRIDE (4 * SLOT_BLOCK * 16)
GD sref s_base_handle() { sref r = {(uint32_t)__cvta_generic_to_shared(slot_mem) + threadIdx.x * 16u}; return r; }
GD sref s_slot(sref base, int k) { sref r = {base.a + (uint32_t)k * SLOT_STRIDE}; return r; }
GD sref s_sel(sref y, sref z, gmask_t is_z) { sref r = {(y.a & ~is_z) | (z.a & is_z)}; return r; }
GD sref s_lane_shift(sref base, uint32_t lanes) { sref r = {base.a + lanes * 16u}; return r; } /* same slot of another lane of the block */
template <int OFF> GD void s_ldq(uint32_t *o, uint32_t a) {
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+%5];" : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3]) : "r"(a), "n"(OFF) : "memory");
}
template <int OFF> GD void s_stq(uint32_t a, const uint32_t *x) {
    asm volatile("st.shared.v4.u32 [%0+%1], {%2,%3,%4,%5};" :: "r"(a), "n"(OFF), "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]) : "memory");
}
GD void s_ld(gf &o, sref s) {
    s_ldq<0>(o.v, s.a); s_ldq<SLOT_BLOCK * 16>(o.v + 4, s.a); s_ldq<2 * SLOT_BLOCK * 16>(o.v + 8, s.a); s_ldq<3 * SLOT_BLOCK * 16>(o.v + 12, s.a);
}
GD void s_st(sref s, const gf &x) {
    s_stq<0>(s.a, x.v); s_stq<SLOT_BLOCK * 16>(s.a, x.v + 4); s_stq<2 * SLOT_BLOCK * 16>(s.a, x.v + 8); s_stq<3 * SLOT_BLOCK * 16>(s.a, x.v + 12);
}
#else
struct sref { gf *a; };
#define SLOT_STRIDE 1
#define SFN static inline
GD sref s_slot(sref base, int k) { sref r = {base.a + k}; return r; }
GD sref s_sel(sref y, sref z, gmask_t is_z) { return is_z ? z : y; }
GD void s_ld(gf &o, sref s) { gf_copy(o, *s.a); }
GD void s_st(sref s, const gf &x) { gf_copy(*s.a, x); }
#endif

// ---- the operation set (LOOSE inputs allowed wherever gf.cuh allows them) -------------------------
SFN void s_mul(sref d, sref a, sref b) { gf x, y, z; s_ld(x, a); s_ld(y, b); gf_mul_body(z, x, y); s_st(d, z); }
SFN void s_sqr(sref d, sref a) { gf x, z; s_ld(x, a); gf_sqr_body(z, x); s_st(d, z); }
SFN void s_sqrn(sref d, sref a, int n) { /* n >= 1 */
    gf x;
    s_ld(x, a);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int i = 0; i < n; i++) gf_sqr_body(x, x);
    s_st(d, x);
}
// d = a * g, g an element in global memory (table rows; never a secret-indexed address on the
// constant-time paths -- those scan with masks, see s_lookup_ct).
template <int QS>
SFN void s_mulg(sref d, sref a, const uint4 *g) { gf x, y, z; s_ld(x, a); gq_ld<false, QS>(y, g); gf_mul_body(z, x, y); s_st(d, z); }
SFN void s_add(sref d, sref a, sref b) { gf x, y; s_ld(x, a); s_ld(y, b); gf_add_nr(x, x, y); s_st(d, x); }   /* TIGHT+TIGHT -> LOOSE */
SFN void s_addr(sref d, sref a, sref b) { gf x, y; s_ld(x, a); s_ld(y, b); gf_add(x, x, y); s_st(d, x); }     /* reduced */
SFN void s_sub(sref d, sref a, sref b) { gf x, y; s_ld(x, a); s_ld(y, b); gf_sub(x, x, y); s_st(d, x); }
// sum = a + b (LOOSE), diff = a - b (TIGHT); one pass over the operands
SFN void s_addsub(sref sum, sref diff, sref a, sref b) {
    gf x, y, z;
    s_ld(x, a); s_ld(y, b);
    gf_sub(z, x, y);
    gf_add_nr(x, x, y);
    s_st(sum, x);
    s_st(diff, z);
}
// d = (a + b)^2
SFN void s_sqr_sum(sref d, sref a, sref b) { gf x, y, z; s_ld(x, a); s_ld(y, b); gf_add_nr(x, x, y); gf_sqr_body(z, x); s_st(d, z); }
// d = 2 a^2 - b
SFN void s_sqr2_sub(sref d, sref a, sref b) {
    gf x, y, z;
    s_ld(x, a);
    gf_sqr_body(z, x);
    s_ld(y, b);
    gf_add_nr(z, z, z);
    gf_sub(z, z, y);
    s_st(d, z);
}
// sum = a^2 + b^2 (LOOSE), diff = b^2 - a^2 (TIGHT), sq = a^2: the first half of a doubling in one pass
SFN void s_sqr2_addsub(sref sum, sref diff, sref a, sref b) {
    gf x, y, z;
    s_ld(x, a);
    gf_sqr_body(x, x);
    s_ld(y, b);
    gf_sqr_body(y, y);
    gf_sub(z, y, x);
    gf_add_nr(x, x, y);
    s_st(sum, x);
    s_st(diff, z);
}
SFN void s_neg(sref d, sref a) { gf x; s_ld(x, a); gf_neg(x, x); s_st(d, x); }
SFN void s_mulw(sref d, sref a, uint32_t w) { gf x, z; s_ld(x, a); gf_mulw(z, x, w); s_st(d, z); }
// d = a * w + c  (LOOSE)
SFN void s_mulw_add(sref d, sref a, uint32_t w, sref c) { gf x, y, z; s_ld(x, a); s_ld(y, c); gf_mulw(z, x, w); gf_add_nr(z, z, y); s_st(d, z); }
// Constant-time table lookup of ONE coordinate: d = entry[idx].<coordinate>, where `first` points at that
// coordinate of entry 0 and entries are `estride` quads apart.  Every lane reads every entry (the
// addresses do not depend on idx) and keeps the one whose index matches, by masks -- the semantics of
// the reference's constant_time_lookup (src/include/constant_time.h:134-183).  n <= 32 entries.
// Unrolling: 16 (the whole row in flight) for the fixed tables every lane reads at the same address,
// 4 for the per-lane window tables (measured: comb 23.48 -> 23.11 ms; window scalarmul 28.7 -> 29.4 ms with 16).
template <bool RO, int QS>
SFN void s_lookup_ct(sref d, const uint4 *first, int estride, int n, uint32_t idx) {
    gf o;
    gf_set_zero(o);
    if (RO) {
#if defined(__CUDA_ARCH__)
#pragma unroll 16
#endif
        for (int e = 0; e < n; e++) {
            const gmask_t m = (gmask_t)(((uint64_t)((uint32_t)e ^ idx) - 1) >> 32); /* all-ones iff e == idx */
            gf t;
            gq_ld<RO, QS>(t, first + (size_t)e * estride);
#pragma unroll
            for (int i = 0; i < 16; i++) o.v[i] |= t.v[i] & m;
        }
    } else {
#if defined(__CUDA_ARCH__)
#pragma unroll 4
#endif
        for (int e = 0; e < n; e++) {
            const gmask_t m = (gmask_t)(((uint64_t)((uint32_t)e ^ idx) - 1) >> 32);
            gf t;
            gq_ld<RO, QS>(t, first + (size_t)e * estride);
#pragma unroll
            for (int i = 0; i < 16; i++) o.v[i] |= t.v[i] & m;
        }
    }
    s_st(d, o);
}
SFN void s_copy(sref d, sref a) { gf x; s_ld(x, a); s_st(d, x); }
template <int QS>
SFN void s_ldg(sref d, const uint4 *g) { gf x; gq_ld<false, QS>(x, g); s_st(d, x); } /* global -> slot */
template <int QS>
SFN void s_stg(uint4 *g, sref a) { gf x; s_ld(x, a); gq_st<QS>(g, x); } /* slot -> global */

// a = x^((p-3)/4); returns all-ones iff a^2 x == 1.  Same addition chain as gf_isr (gf.cuh), walked
// on slots: `a` and `saved` are scratch/output slots distinct from x.
GD gmask_t s_isr(sref a, sref saved, sref x) {
    s_copy(a, x);
    s_copy(saved, x);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int s = 0; s < 12; s++) {
        const uint32_t step = gf_isr_step(s); /* the schedule of gf_isr, held in immediates (gf.cuh) */
        s_sqrn(a, a, (int)(step & 0xff));
        s_mul(a, a, (step & 0x100) ? x : saved); /* public schedule */
        if (step & 0x200) s_copy(saved, a);
    }
    s_sqr(saved, a);
    s_mul(saved, saved, x);
    gf t, one;
    s_ld(t, saved);
    gf_set_ui(one, 1);
    return gf_eq(t, one);
}
// y = 1/x (0 for x = 0); t1, t2 scratch slots, all four distinct.  (goldilocks.c:69-80)
GD void s_invert(sref y, sref x, sref t1, sref t2) {
    s_sqr(t1, x);
    (void)s_isr(y, t2, t1);
    s_sqr(t1, y);
    s_mul(y, t1, x);
}

// Kernel shape for slot functors: F::NSLOTS slots of dynamic shared memory per lane,
// `f(i, base)` with base = handle of this lane's slot 0.  (Host simulator: tests/hostsim run_slots.)
#if defined(__CUDACC__)
template <class F> struct slot_min_blocks { static constexpr int value = 1; };
template <class F>
__global__ void __launch_bounds__(SLOT_BLOCK, slot_min_blocks<F>::value) k_slots(F f, size_t n) {
    const size_t i = (size_t)blockIdx.x * SLOT_BLOCK + threadIdx.x;
#if defined(__CUDA_ARCH__)
    /* out-of-range lanes of the last block stay alive on a clamped index (block-wide barriers inside
     * the functors need all 128 lanes); `live` = false tells the functor not to store anything */
    f(i < n ? i : n - 1, s_base_handle(), i < n);
#endif
}
// Persistent shape for functors that also own a per-thread scratch area in HBM (`slot` = global thread index):
// f(i, base, slot).  Elements are handed out 32 at a time to whichever warp asks next (one atomicAdd per warp and
// chunk on `work_counter`, which the launcher leaves at zero): warps do not run at the same speed (4 vs 3 resident
// blocks per SM at the edges of the grid, different table hit rates) and elements may differ in cost, so a fixed
// grid-stride assignment left lanes idle behind slow neighbours (verify, 2^20 distinct keys: 8.13 -> 8.66 M/s).
// A null counter falls back to the grid-stride loop.
template <class F>
__global__ void __launch_bounds__(SLOT_BLOCK, slot_min_blocks<F>::value) k_slots_persist(F f, size_t n, unsigned *work_counter) {
    const size_t slot = (size_t)blockIdx.x * SLOT_BLOCK + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * SLOT_BLOCK;
#if defined(__CUDA_ARCH__)
    const sref base = s_base_handle();
    if (work_counter) {
        const unsigned lane = threadIdx.x & 31u;
        for (;;) {
            unsigned chunk = 0;
            if (lane == 0) chunk = atomicAdd(work_counter, 1u);
            chunk = __shfl_sync(0xffffffffu, chunk, 0);
            const size_t i = (size_t)chunk * 32u + lane;
            if ((size_t)chunk * 32u >= n) break;
            if (i < n) f(i, base, slot);
            __syncwarp();
        }
    } else {
        for (size_t i = slot; i < n; i += stride) f(i, base, slot);
    }
#endif
}
#endif
