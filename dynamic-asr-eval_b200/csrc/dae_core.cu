// Library-level entry points of libdae.so: ABI version, error strings, launch counter.
#include "common.cuh"

namespace dae {
std::atomic<int64_t> g_launches{0};
}

extern "C" int dae_abi_version(void) { return DAE_ABI_VERSION; }

extern "C" int64_t dae_launch_count(void) { return dae::g_launches.load(std::memory_order_relaxed); }

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
