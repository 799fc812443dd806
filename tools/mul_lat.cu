// tools/mul_lat.cu -- per-warp latency of the field multiplier and of IMAD.WIDE chains (development aid).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I libgoldilocks_b200/csrc -o tools/mul_lat tools/mul_lat.cu
// Prints cycles per dependent gf_mul / gf_sqr per warp for 1, 2, 4 warps per scheduler and the share of the
// multiply pipe that reaches (B200: mul 58 / 80 / 83 %, sqr 53 / 71 / 75 % -- the multiplier alone, without the
// slot machine around it, tops out at 83 % with the 16 warps/SM the kernels run).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define GF_INLINE_MUL 1
#include "slots.cuh"

template <int SQR>
__global__ void k_mul(uint32_t *out, long long *cyc, uint32_t seed, int iters) {
    gf x, y;
#pragma unroll
    for (int i = 0; i < 16; i++) { x.v[i] = (seed * (i + 1) + threadIdx.x) & GF_MASK; y.v[i] = (seed * (i + 7) + blockIdx.x) & GF_MASK; }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        if (SQR) gf_sqr_body(x, x); else gf_mul_body(x, x, y);
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s ^= x.v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// the same chain through the slot machine (operands in shared memory, non-inlined s_mul / s_sqr)
template <int SQR>
__global__ void k_slot_mul(uint32_t *out, long long *cyc, uint32_t seed, int iters) {
#if defined(__CUDA_ARCH__)
    const sref sb = s_base_handle();
    gf x, y;
#pragma unroll
    for (int i = 0; i < 16; i++) { x.v[i] = (seed * (i + 1) + threadIdx.x) & GF_MASK; y.v[i] = (seed * (i + 7) + blockIdx.x) & GF_MASK; }
    s_st(s_slot(sb, 0), x); s_st(s_slot(sb, 1), y);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        if (SQR) s_sqr(s_slot(sb, 0), s_slot(sb, 0)); else s_mul(s_slot(sb, 0), s_slot(sb, 0), s_slot(sb, 1));
    }
    long long t1 = clock64();
    s_ld(x, s_slot(sb, 0));
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s ^= x.v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
#endif
}
int main() {
    uint64_t *out; long long *cyc, h[1024];
    cudaMalloc(&out, 8 << 20); cudaMalloc(&cyc, 8 * 1024);
    int sms = 148;
    const int iters = 2000;
    for (int sq = 0; sq < 2; sq++)
        for (int threads : {32, 128, 256, 512}) {
            if (sq) k_mul<1><<<sms, threads>>>((uint32_t *)out, cyc, 12345, iters); else k_mul<0><<<sms, threads>>>((uint32_t *)out, cyc, 12345, iters);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            cudaMemcpy(h, cyc, 8 * sms, cudaMemcpyDeviceToHost);
            double per = (double)h[0] / iters;
            printf("%s: %4d threads/SM (%.1f warps/scheduler): %.0f cycles per dependent op per warp -> %.1f%% of the multiply pipe\n", sq ? "gf_sqr" : "gf_mul",
                   threads, threads / 128.0, per, 100.0 * (sq ? 110 : 193) * 4.0 * (threads / 128.0) / per);
        }
    for (int sq = 0; sq < 2; sq++)
        for (int blocks : {1, 2, 4}) {
            const int smem = 7 * 64 * SLOT_BLOCK;
            if (sq) { cudaFuncSetAttribute(k_slot_mul<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k_slot_mul<1><<<sms * blocks, SLOT_BLOCK, smem>>>((uint32_t *)out, cyc, 12345, iters); }
            else { cudaFuncSetAttribute(k_slot_mul<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k_slot_mul<0><<<sms * blocks, SLOT_BLOCK, smem>>>((uint32_t *)out, cyc, 12345, iters); }
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            cudaMemcpy(h, cyc, 8 * sms, cudaMemcpyDeviceToHost);
            double per = (double)h[0] / iters;
            printf("%s: %d blocks of 128 per SM (%d warps/scheduler): %.0f cycles per dependent op per warp -> %.1f%% of the multiply pipe\n", sq ? "s_sqr" : "s_mul",
                   blocks, blocks, per, 100.0 * (sq ? 110 : 193) * 4.0 * blocks / per);
        }
    return 0;
}
