// k_scalarmul.cu -- explicit kernel instantiations (see launch.cuh)
#include "launch.cuh"
INSTANTIATE_SLOT(LaneScalarmul)
INSTANTIATE_SLOT(LaneDoubleScalarmul)
