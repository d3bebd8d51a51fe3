// link_stream_kernel instantiations for the 2PL link (own translation unit: compile time).
#include "vibo_stream_kernel.cuh"

namespace vibo {
VIBO_STREAM_LINK_INSTANTIATE(2, stream_link_run2)
}  // namespace vibo
