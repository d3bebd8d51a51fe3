// Host-side packing helpers (csrc/vibo_hostpack.cpp).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <functional>

namespace vibo {
// out[i] = mask[i] ? (resp[i] > 0.5 ? 1 : 0) : -1 for i < n, on the host pool's threads + the caller
void host_pack_parallel(const float* resp, const uint8_t* mask, int8_t* out, size_t n);
// the same on the calling thread only
void host_pack_range(const float* resp, const uint8_t* mask, int8_t* out, size_t n);
int host_pool_threads();   // hardware_concurrency (VIBO_HOST_THREADS overrides), at most 64
}  // namespace vibo
