// k_tables.cu -- explicit kernel instantiations (see launch.cuh)
#include "launch.cuh"
INSTANTIATE_PLAIN(LaneBuildTables)
INSTANTIATE_PLAIN(LaneBuildWide)
INSTANTIATE_PLAIN(LanePrecompute)
INSTANTIATE_PLAIN(LaneNielsFromAbi)
