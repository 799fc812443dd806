// staged.cuh -- the HBM-bound entry points (BASELINE config 1: gf_mul/gf_sqr/gf_add/..., point_add/sub/double
// over 2^20 elements) with their records staged through shared memory.
//
// The batch arrays keep the reference's packed host layout (56-byte field strings, 256-byte point structs).  A
// lane that reads its own record straight from HBM touches 14 (field) or 32 (point) different 128-byte lines per
// warp-level load, so the load/store unit, not DRAM, set the pace (gf_mul 3.7 TB/s, point_add 1.7 TB/s).  Here a
// block moves the contiguous records of its 128 lanes instead:
//   * field records (128 x 56 B = 7 168 B per array): one TMA bulk copy per array (cp.async.bulk ->
//     UBLKCP) completing on an mbarrier, results leave with one bulk store; lanes read their 7 words with
//     64-bit LDS at a stride of 7 words (odd -> conflict-free).  Ragged or unaligned blocks take a
//     cooperative, fully coalesced 64-bit copy instead.
//   * point records (128 x 256 B per array): cooperative 128-bit coalesced loads into rows padded to 264 B
//     (33 words of 8 bytes, odd -> the lanes' 64-bit LDS are conflict-free), results leave the same way.
// The arithmetic is the same gf.cuh / point.cuh code the plain functors (lanes.cuh) run; tests compare both
// shapes against the checker.
#pragma once
#include "lanes.cuh"

#if defined(__CUDACC__)
#define STAGE_BLOCK 128

__device__ __forceinline__ uint32_t st_smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void st_mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(st_smem_addr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void st_mbar_expect(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(st_smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void st_mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" :: "r"(st_smem_addr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void st_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(st_smem_addr(dst)), "l"(src), "r"(bytes), "r"(st_smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void st_bulk_s2g(void *dst, const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst), "r"(st_smem_addr(src)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void st_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- field records -------------------------------------------------------------------------------------
template <int OP>
struct StagedGf {
    uint8_t *out; int32_t *status; const uint8_t *a, *b; uint32_t w;
    static constexpr bool TWO = (OP == GFOP_MUL || OP == GFOP_ADD || OP == GFOP_SUB);
};
template <int OP>
__global__ void __launch_bounds__(STAGE_BLOCK) k_gf_staged(StagedGf<OP> f, size_t n) {
    constexpr int W = 7;                                   /* 64-bit words per record */
    constexpr uint32_t BYTES = STAGE_BLOCK * 56;
    __shared__ alignas(128) uint64_t sa[STAGE_BLOCK * W];
    __shared__ alignas(128) uint64_t sb[StagedGf<OP>::TWO ? STAGE_BLOCK * W : 2];
    __shared__ alignas(8) uint64_t bar;
    const int tid = threadIdx.x;
    const size_t base = (size_t)blockIdx.x * STAGE_BLOCK;
    const int cnt = (n - base) < (size_t)STAGE_BLOCK ? (int)(n - base) : STAGE_BLOCK;
    const uint8_t *ga = f.a + 56 * base, *gb = StagedGf<OP>::TWO ? f.b + 56 * base : nullptr;
    uint8_t *go = f.out + 56 * base;
    const bool bulk = cnt == STAGE_BLOCK && (((uintptr_t)ga | (uintptr_t)gb | (uintptr_t)go) & 15) == 0;
    if (bulk) {
        if (tid == 0) st_mbar_init(&bar, 1);
        __syncthreads();
        if (tid == 0) {
            st_mbar_expect(&bar, StagedGf<OP>::TWO ? 2 * BYTES : BYTES);
            st_bulk_g2s(sa, ga, BYTES, &bar);
            if (StagedGf<OP>::TWO) st_bulk_g2s(sb, gb, BYTES, &bar);
        }
        st_mbar_wait(&bar, 0);
    } else {
        const uint64_t *qa = (const uint64_t *)ga, *qb = (const uint64_t *)gb;
        for (int k = tid; k < cnt * W; k += STAGE_BLOCK) { sa[k] = qa[k]; if (StagedGf<OP>::TWO) sb[k] = qb[k]; }
        __syncthreads();
    }
    if (tid < cnt) {
        uint32_t wa[14], wb[14], wo[14];
        gf x, y, z;
#pragma unroll
        for (int k = 0; k < W; k++) { const uint64_t v = sa[tid * W + k]; wa[2 * k] = (uint32_t)v; wa[2 * k + 1] = (uint32_t)(v >> 32); }
        (void)gf_from_words(x, wa);
        if (StagedGf<OP>::TWO) {
#pragma unroll
            for (int k = 0; k < W; k++) { const uint64_t v = sb[tid * W + k]; wb[2 * k] = (uint32_t)v; wb[2 * k + 1] = (uint32_t)(v >> 32); }
            (void)gf_from_words(y, wb);
        }
        if (OP == GFOP_MUL) gf_mul(z, x, y);
        if (OP == GFOP_SQR) gf_sqr(z, x);
        if (OP == GFOP_ADD) gf_add(z, x, y);
        if (OP == GFOP_SUB) gf_sub(z, x, y);
        if (OP == GFOP_MULW) gf_mulw(z, x, f.w);
        if (OP == GFOP_ISR) f.status[base + tid] = ST_OK(gf_isr(z, x));
        if (OP == GFOP_INVERT) gf_invert(z, x);
        gf_to_words(wo, z);
#pragma unroll
        for (int k = 0; k < W; k++) sa[tid * W + k] = (uint64_t)wo[2 * k] | ((uint64_t)wo[2 * k + 1] << 32); /* own record only */
    }
    if (bulk) {
        st_fence_async();                                  /* generic-proxy writes -> visible to the bulk store */
        __syncthreads();
        if (tid == 0) st_bulk_s2g(go, sa, BYTES);
    } else {
        __syncthreads();
        uint64_t *qo = (uint64_t *)go;
        for (int k = tid; k < cnt * W; k += STAGE_BLOCK) qo[k] = sa[k];
    }
}
template <int OP>
cudaError_t launch_gf_staged(const StagedGf<OP> &f, size_t n, cudaStream_t s) {
    k_gf_staged<OP><<<(unsigned)((n + STAGE_BLOCK - 1) / STAGE_BLOCK), STAGE_BLOCK, 0, s>>>(f, n);
    return cudaGetLastError();
}

// ---- point records -------------------------------------------------------------------------------------
#define STAGE_PT_ROW 33 /* 64-bit words per padded row (32 of data) */
#ifndef STAGE_PT_MINB
#define STAGE_PT_MINB 2 /* resident blocks per SM asked of the register allocator */
#endif
template <int OP>
struct StagedPt {
    abi_pt *out; const abi_pt *a, *b;
    static constexpr bool TWO = (OP == PTOP_ADD || OP == PTOP_SUB);
    static constexpr int SMEM = (TWO ? 2 : 1) * STAGE_BLOCK * STAGE_PT_ROW * 8;
};
__device__ __forceinline__ void st_row_put(uint64_t *rows, int k, const uint4 &v) {
    uint64_t *d = rows + (k >> 4) * STAGE_PT_ROW + 2 * (k & 15);
    d[0] = (uint64_t)v.x | ((uint64_t)v.y << 32);
    d[1] = (uint64_t)v.z | ((uint64_t)v.w << 32);
}
__device__ __forceinline__ void st_rows_in(uint64_t *rows, const abi_pt *g, int cnt, int tid) {
    const uint4 *q = (const uint4 *)g;                     /* 16 quads per record, consecutive lanes -> consecutive quads */
    if (cnt == STAGE_BLOCK) {                              /* all 16 loads of a lane in flight before the first store */
        uint4 v[16];
#pragma unroll
        for (int j = 0; j < 16; j++) v[j] = q[tid + j * STAGE_BLOCK];
#pragma unroll
        for (int j = 0; j < 16; j++) st_row_put(rows, tid + j * STAGE_BLOCK, v[j]);
    } else {
        for (int k = tid; k < cnt * 16; k += STAGE_BLOCK) st_row_put(rows, k, q[k]);
    }
}
template <int OP>
__global__ void __launch_bounds__(STAGE_BLOCK, STAGE_PT_MINB) k_pt_staged(StagedPt<OP> f, size_t n) {
    extern __shared__ uint4 st_rows_q[];
    uint64_t *st_rows = reinterpret_cast<uint64_t *>(st_rows_q);
    uint64_t *ra = st_rows, *rb = st_rows + STAGE_BLOCK * STAGE_PT_ROW;
    const int tid = threadIdx.x;
    const size_t base = (size_t)blockIdx.x * STAGE_BLOCK;
    const int cnt = (n - base) < (size_t)STAGE_BLOCK ? (int)(n - base) : STAGE_BLOCK;
    st_rows_in(ra, f.a + base, cnt, tid);
    if (StagedPt<OP>::TWO) st_rows_in(rb, f.b + base, cnt, tid);
    __syncthreads();
    if (tid < cnt) {
        pt p, q, r;
        pt_from_abi(q, (const abi_pt *)(ra + tid * STAGE_PT_ROW));
        if (StagedPt<OP>::TWO) pt_from_abi(r, (const abi_pt *)(rb + tid * STAGE_PT_ROW));
        if (OP == PTOP_ADD) pt_add(p, q, r);
        if (OP == PTOP_SUB) pt_sub(p, q, r);
        if (OP == PTOP_DBL) pt_double(p, q, false);
        pt_to_abi((abi_pt *)(ra + tid * STAGE_PT_ROW), p);  /* own row only */
    }
    __syncthreads();
    uint4 *qo = (uint4 *)(f.out + base);
    for (int k = tid; k < cnt * 16; k += STAGE_BLOCK) {
        const uint64_t *s = ra + (k >> 4) * STAGE_PT_ROW + 2 * (k & 15);
        uint4 v;
        v.x = (uint32_t)s[0]; v.y = (uint32_t)(s[0] >> 32); v.z = (uint32_t)s[1]; v.w = (uint32_t)(s[1] >> 32);
        qo[k] = v;
    }
}
template <int OP>
cudaError_t launch_pt_staged(const StagedPt<OP> &f, size_t n, cudaStream_t s) {
    static bool configured_dev[64];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (!configured_dev[dev & 63]) {
        e = cudaFuncSetAttribute(k_pt_staged<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, StagedPt<OP>::SMEM);
        if (e != cudaSuccess) return e;
        configured_dev[dev & 63] = true;
    }
    k_pt_staged<OP><<<(unsigned)((n + STAGE_BLOCK - 1) / STAGE_BLOCK), STAGE_BLOCK, StagedPt<OP>::SMEM, s>>>(f, n);
    return cudaGetLastError();
}

#define STAGED_GF(X) X(GFOP_MUL) X(GFOP_SQR) X(GFOP_ADD) X(GFOP_SUB) X(GFOP_MULW) X(GFOP_ISR) X(GFOP_INVERT)
#define STAGED_PT(X) X(PTOP_ADD) X(PTOP_SUB) X(PTOP_DBL)
#define INSTANTIATE_STAGED_GF(OP) template cudaError_t launch_gf_staged<OP>(const StagedGf<OP> &, size_t, cudaStream_t);
#define INSTANTIATE_STAGED_PT(OP) template cudaError_t launch_pt_staged<OP>(const StagedPt<OP> &, size_t, cudaStream_t);
#define DECLARE_STAGED_GF(OP) extern INSTANTIATE_STAGED_GF(OP)
#define DECLARE_STAGED_PT(OP) extern INSTANTIATE_STAGED_PT(OP)
#endif
