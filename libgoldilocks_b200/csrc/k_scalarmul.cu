// k_scalarmul.cu -- explicit kernel instantiations (see launch.cuh)
#include "launch.cuh"
INSTANTIATE_SMP(SlotScalarmul)
INSTANTIATE_SMP(SlotDoubleScalarmul)
INSTANTIATE_SMP(SlotDualScalarmul)
INSTANTIATE_SMP(SlotDirectScalarmul)
