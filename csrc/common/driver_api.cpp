#include "driver_api.h"

#include <cuda_runtime.h>

#include <mutex>
#include <sstream>
#include <stdexcept>

namespace hpcp {

namespace {

struct Loader {
  DriverApi api;
  bool ok = false;
  std::string why;

  template <typename Fn>
  bool resolve(const char* name, Fn& out) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    const cudaError_t e = cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || p == nullptr) {
      (void)cudaGetLastError();
      if (why.empty()) why = std::string("driver entry point not found: ") + name;
      return false;
    }
    out = reinterpret_cast<Fn>(p);
    return true;
  }

  Loader() {
    bool all = true;
#define HPCP_RESOLVE(sym) all = resolve(#sym, api.sym) && all
    HPCP_RESOLVE(cuGetErrorString);
    HPCP_RESOLVE(cuDeviceGet);
    HPCP_RESOLVE(cuDeviceGetAttribute);
    HPCP_RESOLVE(cuCtxGetCurrent);
    HPCP_RESOLVE(cuCtxSetCurrent);
    HPCP_RESOLVE(cuCtxGetDevice);
    HPCP_RESOLVE(cuDevicePrimaryCtxRetain);
    HPCP_RESOLVE(cuDevicePrimaryCtxRelease);
    HPCP_RESOLVE(cuStreamGetCtx);
    HPCP_RESOLVE(cuMemAlloc);
    HPCP_RESOLVE(cuMemFree);
    HPCP_RESOLVE(cuMemcpyDtoDAsync);
    HPCP_RESOLVE(cuMemGetAllocationGranularity);
    HPCP_RESOLVE(cuMemCreate);
    HPCP_RESOLVE(cuMemRelease);
    HPCP_RESOLVE(cuMemAddressReserve);
    HPCP_RESOLVE(cuMemAddressFree);
    HPCP_RESOLVE(cuMemMap);
    HPCP_RESOLVE(cuMemUnmap);
    HPCP_RESOLVE(cuMemSetAccess);
    HPCP_RESOLVE(cuMulticastCreate);
    HPCP_RESOLVE(cuMulticastAddDevice);
    HPCP_RESOLVE(cuMulticastBindMem);
    HPCP_RESOLVE(cuMulticastUnbind);
    HPCP_RESOLVE(cuMulticastGetGranularity);
#undef HPCP_RESOLVE
    ok = all;
  }
};

Loader& loader() {
  static Loader l;
  return l;
}

}  // namespace

bool DriverApi::available() { return loader().ok; }

const DriverApi& DriverApi::get() {
  Loader& l = loader();
  if (!l.ok) throw std::runtime_error("CUDA driver API unavailable: " + l.why);
  return l.api;
}

void DriverApi::check(CUresult r, const char* expr, const char* file, int line) const {
  if (r == CUDA_SUCCESS) return;
  const char* s = nullptr;
  if (cuGetErrorString != nullptr) cuGetErrorString(r, &s);
  std::ostringstream os;
  os << "CUDA driver error " << static_cast<int>(r) << " (" << (s ? s : "?") << ") at " << file
     << ":" << line << " in `" << expr << "`";
  throw std::runtime_error(os.str());
}

}  // namespace hpcp
