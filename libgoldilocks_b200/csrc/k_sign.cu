// k_sign.cu -- explicit kernel instantiations (see launch.cuh)
#include "launch.cuh"
INSTANTIATE_SM(SlotEdDerivePk)
INSTANTIATE_PLAIN(LaneEdSecretScalar)
INSTANTIATE_PLAIN(LaneEdSignExpand)
INSTANTIATE_PLAIN(LaneEdSignNonce)
INSTANTIATE_SM(SlotEdSignR)
INSTANTIATE_PLAIN(LaneEdSignFinish)
