// k_field.cu -- explicit kernel instantiations (see launch.cuh)
#define GF_INLINE_MUL 1 /* one or two multiplications per kernel: keep them inline */
#include "launch.cuh"
INSTANTIATE_PLAIN(LaneGf<GFOP_MUL>)
INSTANTIATE_PLAIN(LaneGf<GFOP_SQR>)
INSTANTIATE_PLAIN(LaneGf<GFOP_ADD>)
INSTANTIATE_PLAIN(LaneGf<GFOP_SUB>)
INSTANTIATE_PLAIN(LaneGf<GFOP_MULW>)
INSTANTIATE_PLAIN(LaneGf<GFOP_ISR>)
INSTANTIATE_PLAIN(LaneGf<GFOP_INVERT>)
