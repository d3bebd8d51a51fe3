// Host-side packing of the reference's row layout (float32 response + uint8 mask, 5 B per cell,
// src/datasets.py:928-940) into the 1 B/cell transfer format (-1 missing, 0 / 1 observed response;
// src/config.py:14), on the caller's CPU cores.  Used (a) by vibo_pack_host, the one-off conversion
// at dataset load, and (b) inside vibo_fused_elbo_host, where part of every step's rows is packed
// by a pool of host threads WHILE the rest crosses PCIe in the reference layout: the two routes use
// different resources (CPU + host DRAM vs the PCIe link), so their throughputs add.
//
// Host-only translation unit (no device code): built with the host compiler so that GCC's function
// multiversioning picks an AVX-512 / AVX2 / baseline clone of the inner loop at load time.
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <mutex>
#include <thread>
#include <vector>

#include "vibo_hostpack.h"

namespace vibo {

#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#define VIBO_CLONES __attribute__((target_clones("arch=skylake-avx512", "avx2", "default")))
#else
#define VIBO_CLONES
#endif

VIBO_CLONES void host_pack_range(const float* __restrict__ resp, const uint8_t* __restrict__ mask,
                                 int8_t* __restrict__ out, size_t n) {
  // branch-free so that the loop vectorises: observed -> (response > 1/2), missing -> 0xFF
  for (size_t i = 0; i < n; ++i) {
    const uint8_t x = (uint8_t)(resp[i] > 0.5f);
    const uint8_t m = (uint8_t)(mask[i] != 0);
    out[i] = (int8_t)((uint8_t)(x & (uint8_t)(0u - m)) | (uint8_t)(m - 1u));
  }
}

// ---- a small persistent pool: `run(n_tasks, fn)` executes fn(task) for task < n_tasks on the pool's
// threads and the caller, and returns when all are done ----------------------------------------
class HostPool {
 public:
  static HostPool& get() {
    static HostPool pool;
    return pool;
  }
  int size() const { return (int)workers_.size() + 1; }
  void run(int n_tasks, const std::function<void(int)>& fn) {
    std::unique_lock<std::mutex> run_lock(run_mutex_);   // one parallel region at a time
    {
      std::lock_guard<std::mutex> lk(m_);
      fn_ = &fn;
      n_tasks_ = n_tasks;
      next_.store(0);
      pending_ = n_tasks;
      ++generation_;
    }
    cv_.notify_all();
    work();
    std::unique_lock<std::mutex> lk(m_);
    done_cv_.wait(lk, [&] { return pending_ == 0; });
    fn_ = nullptr;
  }

 private:
  HostPool() {
    unsigned hc = std::thread::hardware_concurrency();
    int n = hc == 0 ? 4 : (int)hc;
    if (const char* e = getenv("VIBO_HOST_THREADS")) n = atoi(e);
    if (n < 1) n = 1;
    if (n > 64) n = 64;
    for (int i = 1; i < n; ++i) workers_.emplace_back([this] { loop(); });
  }
  ~HostPool() {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  void work() {
    for (;;) {
      const int t = next_.fetch_add(1);
      if (t >= n_tasks_) break;
      (*fn_)(t);
      std::lock_guard<std::mutex> lk(m_);
      if (--pending_ == 0) done_cv_.notify_all();
    }
  }
  void loop() {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return stop_ || generation_ != seen; });
        if (stop_) return;
        seen = generation_;
      }
      work();
    }
  }
  std::vector<std::thread> workers_;
  std::mutex m_, run_mutex_;
  std::condition_variable cv_, done_cv_;
  const std::function<void(int)>* fn_ = nullptr;
  std::atomic<int> next_{0};
  int n_tasks_ = 0, pending_ = 0;
  uint64_t generation_ = 0;
  bool stop_ = false;
};

int host_pool_threads() { return HostPool::get().size(); }

void host_pack_parallel(const float* resp, const uint8_t* mask, int8_t* out, size_t n) {
  if (n == 0) return;
  HostPool& pool = HostPool::get();
  const size_t grain = 1u << 18;   // 256 K cells (1.25 MB read) per task
  const int n_tasks = (int)((n + grain - 1) / grain);
  if (n_tasks <= 1) {
    host_pack_range(resp, mask, out, n);
    return;
  }
  pool.run(n_tasks, [&](int t) {
    const size_t a = (size_t)t * grain, b = a + grain < n ? a + grain : n;
    host_pack_range(resp + a, mask + a, out + a, b - a);
  });
}

}  // namespace vibo
