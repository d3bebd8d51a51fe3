// link_stream_kernel instantiations for the 3PL link (own translation unit: compile time).
#include "vibo_stream_kernel.cuh"

namespace vibo {
VIBO_STREAM_LINK_INSTANTIATE(3, stream_link_run3)
}  // namespace vibo
