// k_bdsm.cu -- explicit kernel instantiations (see launch.cuh)
#include "launch.cuh"
INSTANTIATE_SMP(SlotBaseDoubleScalarmul)
