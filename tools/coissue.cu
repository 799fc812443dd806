// tools/coissue.cu -- can anything run beside a saturated IMAD.WIDE stream on a B200 SM?  ("co-issue of an
// independent DFMA stream beside saturated IMAD.WIDE", and the interleaved-multiplier question.)
//
// Every thread keeps NI independent IMAD.WIDE.U32 chains and ND independent chains of a second instruction kind (DFMA, IMAD.LO,
// LOP3, IADD3) and issues them interleaved in one unrolled loop; a second mode gives whole warps one kind each (even warps
// IMAD.WIDE, odd warps the other).  Reported: lanes per clock per SM of each kind (clock64 of the slowest block) against the
// stand-alone rates.  If the second stream were free, IMAD.WIDE would stay at its 31.5/clk/SM.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/coissue tools/coissue.cu && tools/coissue
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

enum { K_DFMA = 0, K_LO = 1, K_LOP = 2, K_IADD = 3 };

template <int KIND>
__device__ __forceinline__ void other_op(double &d, uint32_t &u, double db, uint32_t ub) {
    if (KIND == K_DFMA) asm volatile("fma.rn.f64 %0, %0, %1, %0;" : "+d"(d) : "d"(db));
    if (KIND == K_LO) asm volatile("mad.lo.u32 %0, %0, %1, %0;" : "+r"(u) : "r"(ub));
    if (KIND == K_LOP) asm volatile("lop3.b32 %0, %0, %1, %1, 0x96;" : "+r"(u) : "r"(ub));
    if (KIND == K_IADD) asm volatile("add.u32 %0, %0, %1;" : "+r"(u) : "r"(ub));
}

// NI IMAD.WIDE chains and ND chains of KIND per thread; split = 1: even warps run only the IMAD.WIDE chains, odd warps only the others
template <int KIND, int NI, int ND>
__global__ void __launch_bounds__(1024) k_mix(uint64_t *out, long long *cycles, uint32_t seed, int iters, int split) {
    uint64_t acc[NI > 0 ? NI : 1];
    double d[ND > 0 ? ND : 1];
    uint32_t u[ND > 0 ? ND : 1];
    const uint32_t a = seed * (threadIdx.x | 1u), b = seed ^ (blockIdx.x * 2654435761u | 1u);
    const double db = (double)b * 1e-12;
#pragma unroll
    for (int i = 0; i < (NI > 0 ? NI : 1); i++) acc[i] = (uint64_t)(a + i) << 7;
#pragma unroll
    for (int i = 0; i < (ND > 0 ? ND : 1); i++) { d[i] = (double)(a + i) * 1e-9; u[i] = a + i; }
    const bool do_i = !split || ((threadIdx.x >> 5) & 1) == 0, do_o = !split || ((threadIdx.x >> 5) & 1) == 1;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int i = 0; i < (NI > ND ? NI : ND); i++) {
                if (i < NI && do_i) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"((uint32_t)acc[(i + 1) % (NI > 0 ? NI : 1)]), "r"(b));
                if (i < ND && do_o) other_op<KIND>(d[i], u[i], db, b);
            }
        }
    }
    const long long t1 = clock64();
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < (NI > 0 ? NI : 1); i++) s ^= acc[i];
#pragma unroll
    for (int i = 0; i < (ND > 0 ? ND : 1); i++) s ^= (uint64_t)__double_as_longlong(d[i]) ^ u[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int KIND, int NI, int ND>
static int run(const char *name, int sms, int split, uint64_t *out, long long *cyc) {
    const int threads = 1024, iters = 1500;
    k_mix<KIND, NI, ND><<<sms, threads>>>(out, cyc, 12345u, 20, split);
    CK(cudaDeviceSynchronize());
    k_mix<KIND, NI, ND><<<sms, threads>>>(out, cyc, 54321u, iters, split);
    CK(cudaDeviceSynchronize());
    std::vector<long long> h(sms);
    CK(cudaMemcpy(h.data(), cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
    const double clk = (double)*std::max_element(h.begin(), h.end());
    const double frac = split ? 0.5 : 1.0;  /* share of the threads that issue each kind */
    const double wide = NI ? threads * frac * iters * 8.0 * NI / clk : 0, other = ND ? threads * frac * iters * 8.0 * ND / clk : 0;
    printf("{\"second\": \"%s\", \"imad_wide_chains\": %d, \"second_chains\": %d, \"warps\": \"%s\", \"imad_wide_per_clk_sm\": %.2f, \"second_per_clk_sm\": %.2f},\n",
           name, NI, ND, split ? "split (even: IMAD.WIDE, odd: second)" : "mixed in every warp", wide, other);
    return 0;
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    uint64_t *out; long long *cyc;
    CK(cudaMalloc(&out, (size_t)sms * 1024 * sizeof(uint64_t)));
    CK(cudaMalloc(&cyc, sms * sizeof(long long)));
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"how\": \"tools/coissue.cu: 1024 threads per SM, independent chains, lanes per clock per SM by clock64\", \"runs\": [\n", p.name, sms);
    run<K_DFMA, 8, 0>("none", sms, 0, out, cyc);
    run<K_DFMA, 0, 8>("dfma", sms, 0, out, cyc);
    run<K_DFMA, 8, 8>("dfma", sms, 0, out, cyc);
    run<K_DFMA, 8, 4>("dfma", sms, 0, out, cyc);
    run<K_DFMA, 8, 2>("dfma", sms, 0, out, cyc);
    run<K_DFMA, 8, 8>("dfma", sms, 1, out, cyc);
    run<K_LO, 0, 8>("imad_lo", sms, 0, out, cyc);
    run<K_LO, 8, 8>("imad_lo", sms, 0, out, cyc);
    run<K_LOP, 0, 8>("lop3", sms, 0, out, cyc);
    run<K_LOP, 8, 8>("lop3", sms, 0, out, cyc);
    run<K_LOP, 8, 4>("lop3", sms, 0, out, cyc);
    run<K_LOP, 8, 8>("lop3", sms, 1, out, cyc);
    run<K_IADD, 8, 8>("iadd", sms, 0, out, cyc);
    printf("{}]}\n");
    return 0;
}
