// k_x448.cu -- explicit kernel instantiations (see launch.cuh)
#include "launch.cuh"
INSTANTIATE_SM(SlotX448)
