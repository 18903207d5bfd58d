// CUDA *driver* API entry points resolved at run time through the runtime
// (cudaGetDriverEntryPoint), so nothing links against libcuda.so: the same
// binaries / extension import fine on the GPU-less build box.
//
// Used for (a) VMM + NVSwitch multicast objects (peer_mem.cpp) and (b) the
// native-handle interop demo, which is the CUDA analogue of the reference
// pulling ze_driver/ze_context/ze_device handles out of an OpenMP interop object
// (sycl_omp_ze_interopt/interop_omp_ze_sycl.cpp:24-34).
#pragma once

#include <cuda.h>
#include <cudaTypedefs.h>

#include <string>

namespace hpcp {

struct DriverApi {
  PFN_cuGetErrorString cuGetErrorString = nullptr;
  PFN_cuDeviceGet cuDeviceGet = nullptr;
  PFN_cuDeviceGetAttribute cuDeviceGetAttribute = nullptr;
  PFN_cuCtxGetCurrent cuCtxGetCurrent = nullptr;
  PFN_cuCtxSetCurrent cuCtxSetCurrent = nullptr;
  PFN_cuCtxGetDevice cuCtxGetDevice = nullptr;
  PFN_cuDevicePrimaryCtxRetain cuDevicePrimaryCtxRetain = nullptr;
  PFN_cuDevicePrimaryCtxRelease cuDevicePrimaryCtxRelease = nullptr;
  PFN_cuStreamGetCtx cuStreamGetCtx = nullptr;
  PFN_cuMemAlloc cuMemAlloc = nullptr;
  PFN_cuMemFree cuMemFree = nullptr;
  PFN_cuMemcpyDtoDAsync cuMemcpyDtoDAsync = nullptr;
  PFN_cuMemGetAllocationGranularity cuMemGetAllocationGranularity = nullptr;
  PFN_cuMemCreate cuMemCreate = nullptr;
  PFN_cuMemRelease cuMemRelease = nullptr;
  PFN_cuMemAddressReserve cuMemAddressReserve = nullptr;
  PFN_cuMemAddressFree cuMemAddressFree = nullptr;
  PFN_cuMemMap cuMemMap = nullptr;
  PFN_cuMemUnmap cuMemUnmap = nullptr;
  PFN_cuMemSetAccess cuMemSetAccess = nullptr;
  PFN_cuMulticastCreate cuMulticastCreate = nullptr;
  PFN_cuMulticastAddDevice cuMulticastAddDevice = nullptr;
  PFN_cuMulticastBindMem cuMulticastBindMem = nullptr;
  PFN_cuMulticastUnbind cuMulticastUnbind = nullptr;
  PFN_cuMulticastGetGranularity cuMulticastGetGranularity = nullptr;

  // Throws std::runtime_error if no driver is present.
  static const DriverApi& get();
  // True iff a driver could be loaded (never throws).
  static bool available();

  void check(CUresult r, const char* expr, const char* file, int line) const;
};

}  // namespace hpcp

#define HPCP_CU(expr) ::hpcp::DriverApi::get().check((expr), #expr, __FILE__, __LINE__)
