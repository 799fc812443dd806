// k_comb.cu -- explicit kernel instantiations (see launch.cuh)
#include "launch.cuh"
INSTANTIATE_SM(SlotComb)
INSTANTIATE_SM(SlotX448DerivePk)
INSTANTIATE_SM(SlotCombTable)
INSTANTIATE_SM(SlotNielsDebug)
