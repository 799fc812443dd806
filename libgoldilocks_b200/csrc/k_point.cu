// k_point.cu -- explicit kernel instantiations (see launch.cuh)
#define GF_INLINE_MUL 1 /* one or two multiplications per kernel: keep them inline */
#include "launch.cuh"
#include "staged.cuh"
STAGED_PT(INSTANTIATE_STAGED_PT)
INSTANTIATE_PLAIN(LanePt<PTOP_NEG>)
INSTANTIATE_PLAIN(LanePtEq)
INSTANTIATE_PLAIN(LanePtValid)
INSTANTIATE_PLAIN(LanePt<PTOP_TORQUE>)
INSTANTIATE_PLAIN(LanePtPscale)
INSTANTIATE_PLAIN(LanePtNiels)
