// launch.cuh -- kernel shapes around the lane functors of lanes.cuh and their host-side launchers.
//
// Each heavy functor is compiled in its own translation unit (k_*.cu explicitly instantiates the
// launchers for its functors); abi.cu only sees `extern template` declarations, so the library
// builds in parallel and a change to one kernel recompiles one file.
#pragma once
#include <cuda_runtime.h>
#include "lanes.cuh"

#define BLOCK 128

// Resident blocks per SM the register allocator is asked to allow (default: whatever fits).
template <class F> struct lane_min_blocks { static constexpr int value = 1; };
// Measured on B200 at 2^20 (tools/opbench.py): comb 34.5 -> 38.5 Mops/s with 3 blocks (168 regs, no spills);
// X448 gains 1% at 3 blocks but spills, so it stays at 2.
template <> struct lane_min_blocks<LaneComb> { static constexpr int value = 3; };
template <> struct lane_min_blocks<LaneX448DerivePk> { static constexpr int value = 3; };

template <class F>
__global__ void __launch_bounds__(BLOCK, lane_min_blocks<F>::value) k_lanes(F f, size_t n) {
    const size_t i = (size_t)blockIdx.x * BLOCK + threadIdx.x;
    if (i < n) f(i);
}
// Persistent grid-stride shape for functors that own a per-thread scratch slot in HBM.
template <class F>
__global__ void __launch_bounds__(BLOCK) k_lanes_slot(F f, size_t n) {
    const size_t slot = (size_t)blockIdx.x * BLOCK + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * BLOCK;
    for (size_t i = slot; i < n; i += stride) f(i, slot);
}

template <class F>
cudaError_t launch_lanes(const F &f, size_t n, cudaStream_t s) {
    k_lanes<F><<<(unsigned)((n + BLOCK - 1) / BLOCK), BLOCK, 0, s>>>(f, n);
    return cudaGetLastError();
}
template <class F>
cudaError_t launch_lanes_slot(const F &f, size_t n, int grid, cudaStream_t s) {
    k_lanes_slot<F><<<grid, BLOCK, 0, s>>>(f, n);
    return cudaGetLastError();
}
// resident blocks per SM of the slot kernel of F
template <class F>
cudaError_t lanes_slot_occupancy(int *occ) {
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k_lanes_slot<F>, BLOCK, 0);
}

#define LANES_PLAIN(X)                                                                              \
    X(LaneGf<GFOP_MUL>) X(LaneGf<GFOP_SQR>) X(LaneGf<GFOP_ADD>) X(LaneGf<GFOP_SUB>)                 \
    X(LaneGf<GFOP_MULW>) X(LaneGf<GFOP_ISR>) X(LaneGf<GFOP_INVERT>)                                 \
    X(LanePt<PTOP_ADD>) X(LanePt<PTOP_SUB>) X(LanePt<PTOP_DBL>) X(LanePt<PTOP_NEG>)                 \
    X(LanePtEq) X(LanePtValid) X(LanePtEncode) X(LanePtDecode)                                      \
    X(LaneFromHash<false>) X(LaneFromHash<true>)                                                    \
    X(LaneEncodeEddsa) X(LaneDecodeEddsa) X(LaneEncodeX448)                                         \
    X(LaneComb) X(LaneX448DerivePk) X(LaneX448)                                                     \
    X(LaneSc<SCOP_ADD>) X(LaneSc<SCOP_SUB>) X(LaneSc<SCOP_MUL>) X(LaneSc<SCOP_HALVE>)               \
    X(LaneScDecodeLong) X(LaneShake256)                                                             \
    X(LaneEdDerivePk) X(LaneEdSecretScalar) X(LaneEdSignExpand) X(LaneEdSignNonce) X(LaneEdSignR) X(LaneEdSignFinish)   \
    X(LaneEdVerifyDecode) X(LaneEdVerifyScalars) X(LaneBuildTables) X(LaneBuildWide)
#define LANES_SLOT(X)                                                                               \
    X(LaneScalarmul) X(LaneDoubleScalarmul) X(LaneBaseDoubleScalarmul) X(LaneEdVerifyFinish)

#define INSTANTIATE_PLAIN(F) template cudaError_t launch_lanes<F>(const F &, size_t, cudaStream_t);
#define INSTANTIATE_SLOT(F)                                                                         \
    template cudaError_t launch_lanes_slot<F>(const F &, size_t, int, cudaStream_t);                \
    template cudaError_t lanes_slot_occupancy<F>(int *);
#define DECLARE_PLAIN(F) extern INSTANTIATE_PLAIN(F)
#define DECLARE_SLOT(F)                                                                             \
    extern template cudaError_t launch_lanes_slot<F>(const F &, size_t, int, cudaStream_t);         \
    extern template cudaError_t lanes_slot_occupancy<F>(int *);
