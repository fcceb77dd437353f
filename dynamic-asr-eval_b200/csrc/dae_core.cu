// Library-level entry points of libdae.so: ABI version, error strings, launch counter.
#include <cstdlib>
#include <mutex>
#include "common.cuh"
#include "ctc_shared.cuh"

namespace dae {
std::atomic<int64_t> g_launches{0};
}

extern "C" int dae_abi_version(void) { return DAE_ABI_VERSION; }

extern "C" int64_t dae_launch_count(void) { return dae::g_launches.load(std::memory_order_relaxed); }

extern "C" void dae_ctc_configure(int blocked, int cluster, int pairs, int overlap) {
  dae::CtcConfig& c = dae::ctc_config();
  c.blocked.store(blocked < 0 ? -1 : (blocked ? 1 : 0));
  c.cluster.store(cluster);
  c.pairs.store(pairs);
  c.overlap.store(overlap < 0 ? -1 : overlap);
}

extern "C" const char* dae_error_string(int code) {
  switch (code) {
    case 0: return "ok";
    case DAE_E_BADARG: return "dae: bad argument (null pointer or negative size)";
    case DAE_E_TOOBIG: return "dae: a dimension exceeds what the kernel supports";
    case DAE_E_SCRATCH: return "dae: scratch buffer missing or too small";
    case DAE_E_ALIGN: return "dae: pointer/stride alignment requirement not met";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "dae: unknown error";
}

namespace dae {
CtcConfig& ctc_config() {
  static CtcConfig cfg;
  static std::once_flag once;
  std::call_once(once, [] {
    auto env_int = [](const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; };
    cfg.blocked.store(env_int("DAE_CTC_BLOCKED", -1));
    cfg.cluster.store(env_int("DAE_CTC_CLUSTER", 0));
    cfg.pairs.store(env_int("DAE_CTC_PAIRS", 0));
    cfg.overlap.store(env_int("DAE_CTC_OVERLAP", -1));
  });
  return cfg;
}

int sm_count() {
  static std::atomic<int> cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  int v = cache[dev].load(std::memory_order_relaxed);
  if (v == 0) {
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    cache[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}

cudaError_t ensure_dyn_smem_impl(const void* kern, int bytes) {
  struct Entry { const void* kern; int dev; int bytes; };
  static Entry table[128];
  static int used = 0;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const int n = used;
  for (int i = 0; i < n; ++i)
    if (table[i].kern == kern && table[i].dev == dev) {
      if (table[i].bytes >= bytes) return cudaSuccess;
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
      if (e == cudaSuccess) table[i].bytes = bytes;
      return e;
    }
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess && n < 128) {
    table[n] = Entry{kern, dev, bytes};
    used = n + 1;
  }
  return e;
}
}  // namespace dae
