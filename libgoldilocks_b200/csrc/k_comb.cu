// k_comb.cu -- explicit kernel instantiations (see launch.cuh)
#include "launch.cuh"
INSTANTIATE_PLAIN(LaneComb)
INSTANTIATE_PLAIN(LaneX448DerivePk)
