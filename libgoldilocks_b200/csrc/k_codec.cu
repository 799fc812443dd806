// k_codec.cu -- explicit kernel instantiations (see launch.cuh)
#include "launch.cuh"
INSTANTIATE_PLAIN(LanePtEncode)
INSTANTIATE_PLAIN(LanePtDecode)
INSTANTIATE_PLAIN(LaneEncodeEddsa)
INSTANTIATE_PLAIN(LaneDecodeEddsa)
INSTANTIATE_PLAIN(LaneEncodeX448)
INSTANTIATE_PLAIN(LaneEdPkToX448)
