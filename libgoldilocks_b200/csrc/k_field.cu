// k_field.cu -- explicit kernel instantiations (see launch.cuh): field entry points, staged through shared memory
#define GF_INLINE_MUL 1 /* one or two multiplications per kernel: keep them inline */
#include "launch.cuh"
#include "staged.cuh"
STAGED_GF(INSTANTIATE_STAGED_GF)
