// tools/imad_peak.cu -- measures the integer-multiply pipe peak of the GPU it runs on.
//
// the roofline denominator for this project is
//   peak_MAC32/s = SMs x (IMAD.WIDE.U32 issued per clk per SM) x sustained SM clock
// and it is not in MEASURED_PEAKS.json, so we measure it.  One MAC32 = one 32x32->64-bit
// multiply-accumulate = one IMAD.WIDE.U32 SASS instruction per lane.
//
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/imad_peak tools/imad_peak.cu
// Run  :  tools/imad_peak [out.json]
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

enum Kind { K_WIDE = 0, K_LO = 1, K_HI = 2, K_WIDE_ALU = 3, K_DFMA = 4, K_WIDE_CC = 5 };

template <int KIND, int ILP>
__global__ void __launch_bounds__(1024) k_peak(uint64_t *out, long long *cycles, uint32_t seed, int iters) {
    uint64_t acc[ILP];
    uint32_t lo32[ILP];
    double dacc[ILP];
    uint32_t a = seed * (threadIdx.x | 1u), b = seed ^ (blockIdx.x * 2654435761u | 1u);
    uint32_t alu0 = a, alu1 = b;
#pragma unroll
    for (int i = 0; i < ILP; i++) { acc[i] = (uint64_t)(a + i) << 7; lo32[i] = a + i; dacc[i] = (double)(a + i); }
    double da = (double)a * 1e-9, db = (double)b * 1e-9;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int i = 0; i < ILP; i++) {
                // the multiplicand is the low word of the neighbouring chain, so nothing is loop-invariant
                uint32_t m = (uint32_t)acc[(i + 1) % ILP];
                if (KIND == K_LO)
                    asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo32[i]) : "r"(lo32[(i + 1) % ILP]), "r"(b));
                if (KIND == K_HI)
                    asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(lo32[i]) : "r"(lo32[(i + 1) % ILP]), "r"(b));
                if (KIND == K_DFMA)
                    asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(dacc[i]) : "d"(dacc[(i + 1) % ILP]), "d"(db));
                if (KIND == K_WIDE_ALU) {   // two independent ALU-pipe ops per multiply: do they co-issue?
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(alu0) : "r"(a), "r"(b));
                    asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(alu1) : "r"(b));
                }
            }
            if (KIND == K_WIDE_CC) {       // carry-chained row: lo/hi pairs fused by ptxas into IMAD.WIDE.U32 + carry
                uint32_t m = (uint32_t)acc[0] | 1u;
#pragma unroll
                for (int i = 0; i < ILP; i++) {
                    uint32_t l = (uint32_t)acc[i], h = (uint32_t)(acc[i] >> 32);
                    uint32_t bi = (uint32_t)(acc[(i + 1) % ILP] >> 32);
                    if (i == 0) {
                        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;"
                                     : "+r"(l), "+r"(h) : "r"(m), "r"(bi));
                    } else {
                        asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;"
                                     : "+r"(l), "+r"(h) : "r"(m), "r"(bi));
                    }
                    acc[i] = ((uint64_t)h << 32) | l;
                }
            }
        }
    }
    long long t1 = clock64();
    uint64_t s = alu0 ^ alu1;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += acc[i] + lo32[i] + (uint64_t)dacc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}


// Realistic pattern: an 8x8 product-scanning block (the inner shape of a field multiplication):
// 64 IMAD.WIDE.U32 with 64-bit accumulate into 8 column accumulators, operands refreshed from the
// accumulators once per block (8 ALU ops per 64 multiplies) so nothing is loop-invariant.
template <int NALU>
__global__ void __launch_bounds__(1024) k_scan(uint64_t *out, long long *cycles, uint32_t seed, int iters) {
    uint32_t x[8], y[8];
    uint64_t acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = seed * (threadIdx.x + i) | 1u; y[i] = (seed ^ blockIdx.x) + 77u * i; acc[i] = i; }
    uint32_t alu0 = seed, alu1 = ~seed;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = 0; j < 8; j++) {
                acc[(i + j) & 7] += (uint64_t)x[i] * y[j];
                if (NALU >= 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(alu0) : "r"(x[i]), "r"(y[j]));
                if (NALU >= 2) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(alu1) : "r"(y[j]));
            }
#pragma unroll
        for (int i = 0; i < 8; i++) { x[i] ^= (uint32_t)(acc[i] >> 32); }
    }
    long long t1 = clock64();
    uint64_t s = alu0 ^ alu1;
#pragma unroll
    for (int i = 0; i < 8; i++) s += acc[i] ^ y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int NALU>
static struct Result run_scan(const char *name, int sms, int threads, int bps, uint64_t *out, long long *cyc);

struct Result { const char *name; int ilp, threads, blocks_per_sm; double gops, per_clk_sm, mhz; };

template <int KIND, int ILP>
static Result run(const char *name, int sms, int threads, int bps, uint64_t *out, long long *cyc) {
    int blocks = sms * bps;
    int iters = 2000;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_peak<KIND, ILP><<<blocks, threads>>>(out, cyc, 12345u, 50); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        CK(cudaEventRecord(e0));
        k_peak<KIND, ILP><<<blocks, threads>>>(out, cyc, 12345u + rep, iters);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = std::min(best, ms);
    }
    std::vector<long long> hc(blocks); CK(cudaMemcpy(hc.data(), cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost));
    long long mx = *std::max_element(hc.begin(), hc.end());
    double ops = (double)blocks * threads * (double)iters * 8.0 * ILP;      // lane-level multiply instructions
    Result r; r.name = name; r.ilp = ILP; r.threads = threads; r.blocks_per_sm = bps;
    r.gops = ops / (best * 1e-3) / 1e9;
    r.per_clk_sm = ops / sms / (double)mx;                                    // lanes per clk per SM (by clock64)
    r.mhz = (double)mx / (best * 1e-3) / 1e6;
    return r;
}

template <int NALU>
static Result run_scan(const char *name, int sms, int threads, int bps, uint64_t *out, long long *cyc) {
    int blocks = sms * bps;
    int iters = 2000;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_scan<NALU><<<blocks, threads>>>(out, cyc, 12345u, 50); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        CK(cudaEventRecord(e0));
        k_scan<NALU><<<blocks, threads>>>(out, cyc, 12345u + rep, iters);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = std::min(best, ms);
    }
    std::vector<long long> hc(blocks); CK(cudaMemcpy(hc.data(), cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost));
    long long mx = *std::max_element(hc.begin(), hc.end());
    double ops = (double)blocks * threads * (double)iters * 64.0;
    Result r; r.name = name; r.ilp = 8; r.threads = threads; r.blocks_per_sm = bps;
    r.gops = ops / (best * 1e-3) / 1e9;
    r.per_clk_sm = ops / sms / (double)mx;
    r.mhz = (double)mx / (best * 1e-3) / 1e6;
    return r;
}

int main(int argc, char **argv) {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    uint64_t *out; long long *cyc;
    CK(cudaMalloc(&out, (size_t)sms * 8 * 1024 * sizeof(uint64_t))); CK(cudaMalloc(&cyc, sms * 8 * sizeof(long long)));
    std::vector<Result> rs;
    const bool quick = argc > 1 && !strcmp(argv[1], "--quick");   /* bench.py: the three best IMAD.WIDE shapes only, JSON on stdout (~1 s) */
    if (quick) {
        rs.push_back(run_scan<0>("imad_wide_u32", sms, 1024, 1, out, cyc));
        rs.push_back(run_scan<0>("imad_wide_u32", sms, 512, 2, out, cyc));
        rs.push_back(run_scan<0>("imad_wide_u32", sms, 256, 4, out, cyc));
        argc = 1;
    } else {
    rs.push_back(run_scan<0>("imad_wide_u32", sms, 1024, 1, out, cyc));
    rs.push_back(run_scan<0>("imad_wide_u32", sms, 512, 1, out, cyc));
    rs.push_back(run_scan<0>("imad_wide_u32", sms, 256, 1, out, cyc));
    rs.push_back(run_scan<0>("imad_wide_u32", sms, 128, 1, out, cyc));
    rs.push_back(run_scan<0>("imad_wide_u32", sms, 512, 2, out, cyc));
    rs.push_back(run_scan<0>("imad_wide_u32", sms, 256, 4, out, cyc));
    rs.push_back(run_scan<1>("imad_wide_u32+1alu", sms, 1024, 1, out, cyc));
    rs.push_back(run_scan<2>("imad_wide_u32+2alu", sms, 1024, 1, out, cyc));
    rs.push_back(run_scan<2>("imad_wide_u32+2alu", sms, 512, 1, out, cyc));
    rs.push_back(run<K_LO, 8>("imad_lo", sms, 1024, 1, out, cyc));
    rs.push_back(run<K_LO, 8>("imad_lo", sms, 512, 1, out, cyc));
    rs.push_back(run<K_HI, 8>("imad_hi", sms, 1024, 1, out, cyc));
    rs.push_back(run<K_WIDE_CC, 8>("imad_wide_carry_chain", sms, 1024, 1, out, cyc));
    rs.push_back(run<K_WIDE_CC, 8>("imad_wide_carry_chain", sms, 512, 1, out, cyc));
    rs.push_back(run<K_DFMA, 8>("dfma", sms, 1024, 1, out, cyc));
    }
    FILE *f = argc > 1 ? fopen(argv[1], "w") : stdout;
    double best_wide = 0, best_wide_clk = 0, mhz = 0;
    for (auto &r : rs) if (!strcmp(r.name, "imad_wide_u32") && r.gops > best_wide) { best_wide = r.gops; best_wide_clk = r.per_clk_sm; mhz = r.mhz; }
    fprintf(f, "{\n \"gpu_name\": \"%s\", \"sms\": %d, \"clock_rate_khz\": %d,\n", p.name, sms, p.clockRate);
    fprintf(f, " \"imad_wide_u32_gmac_s\": %.1f, \"imad_wide_u32_per_clk_per_sm\": %.2f, \"sm_mhz_during\": %.0f,\n", best_wide, best_wide_clk, mhz);
    fprintf(f, " \"how\": \"tools/imad_peak.cu: ILP independent mad.wide.u32 chains per thread, 1 wave, best of 5, CUDA events; per-clk by clock64\",\n");
    fprintf(f, " \"runs\": [\n");
    for (size_t i = 0; i < rs.size(); i++)
        fprintf(f, "  {\"kind\": \"%s\", \"ilp\": %d, \"threads\": %d, \"blocks_per_sm\": %d, \"gops\": %.1f, \"lanes_per_clk_per_sm\": %.2f, \"sm_mhz\": %.0f}%s\n",
                rs[i].name, rs[i].ilp, rs[i].threads, rs[i].blocks_per_sm, rs[i].gops, rs[i].per_clk_sm, rs[i].mhz, i + 1 < rs.size() ? "," : "");
    fprintf(f, " ]\n}\n");
    if (f != stdout) fclose(f);
    return 0;
}
