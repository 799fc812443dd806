// k_elligator.cu -- explicit kernel instantiations (see launch.cuh)
#include "launch.cuh"
INSTANTIATE_PLAIN(LaneFromHash<false>)
INSTANTIATE_PLAIN(LaneFromHash<true>)
INSTANTIATE_PLAIN(LaneInvertElligator<false>)
INSTANTIATE_PLAIN(LaneInvertElligator<true>)
