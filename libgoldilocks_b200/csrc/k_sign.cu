// k_sign.cu -- explicit kernel instantiations (see launch.cuh)
#include "launch.cuh"
INSTANTIATE_PLAIN(LaneEdDerivePk)
INSTANTIATE_PLAIN(LaneEdSecretScalar)
INSTANTIATE_PLAIN(LaneEdSignExpand)
INSTANTIATE_PLAIN(LaneEdSignNonce)
INSTANTIATE_PLAIN(LaneEdSignR)
INSTANTIATE_PLAIN(LaneEdSignFinish)
