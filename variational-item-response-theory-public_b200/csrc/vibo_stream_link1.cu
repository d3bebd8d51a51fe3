// link_stream_kernel instantiations for the 1PL link (own translation unit: compile time).
#include "vibo_stream_kernel.cuh"

namespace vibo {
VIBO_STREAM_LINK_INSTANTIATE(1, stream_link_run1)
}  // namespace vibo
