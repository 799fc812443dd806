// k_verify.cu -- explicit kernel instantiations (see launch.cuh)
#include "launch.cuh"
INSTANTIATE_PLAIN(LaneEdVerifyDecode)
INSTANTIATE_PLAIN(LaneEdVerifyScalars)
INSTANTIATE_SMP(SlotEdVerifyFinish)
INSTANTIATE_SMP(SlotEdVerifyFinishShared)
INSTANTIATE_SMP(SlotKeyTables)
