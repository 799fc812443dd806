// k_scalar.cu -- explicit kernel instantiations (see launch.cuh)
#include "launch.cuh"
INSTANTIATE_PLAIN(LaneSc<SCOP_ADD>)
INSTANTIATE_PLAIN(LaneSc<SCOP_SUB>)
INSTANTIATE_PLAIN(LaneSc<SCOP_MUL>)
INSTANTIATE_PLAIN(LaneSc<SCOP_HALVE>)
INSTANTIATE_PLAIN(LaneScDecodeLong)
INSTANTIATE_PLAIN(LaneShake256)
