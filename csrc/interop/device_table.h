// Cached per-device native-handle table shared by the interop demos.
//
// Library-like API kept from the reference's interop demos
// (sycl_omp_ze_interopt/interop_omp_ze_sycl.cpp:13-79: xomp_get_infos_devices() builds,
// once, a table {device -> sycl::context/sycl::device (+ze_context)} from the OpenMP
// interop object; xomp_get_device_info(n) indexes it).  CUDA analogue: for every
// runtime ordinal the driver-level CUdevice and the *primary* CUcontext that the
// runtime, PyTorch and any driver-API library all share.  Ownership is "keep":
// the table retains the primary context (ref-count) and never destroys it.
#pragma once

#include <cuda_runtime.h>

#include <vector>

#include "../common/cuda_check.h"
#include "../common/driver_api.h"

namespace hpcp {

struct DeviceNativeInfo {
  int ordinal = -1;          // CUDA runtime ordinal
  CUdevice cu_device = 0;    // driver handle
  CUcontext cu_context = nullptr;  // primary context (shared with the runtime / torch)
  std::vector<int> peers;    // ordinals reachable with P2P (the "same context group" analogue)
};

inline const std::vector<DeviceNativeInfo>& xcuda_get_infos_devices() {
  static std::vector<DeviceNativeInfo> table = [] {
    std::vector<DeviceNativeInfo> t;
    int n = 0;
    HPCP_CUDA(cudaGetDeviceCount(&n));
    const DriverApi& d = DriverApi::get();
    int prev = 0;
    HPCP_CUDA(cudaGetDevice(&prev));
    for (int i = 0; i < n; ++i) {
      DeviceNativeInfo info;
      info.ordinal = i;
      HPCP_CUDA(cudaSetDevice(i));
      HPCP_CUDA(cudaFree(nullptr));  // make the runtime create / bind the primary context
      HPCP_CU(d.cuDeviceGet(&info.cu_device, i));
      HPCP_CU(d.cuDevicePrimaryCtxRetain(&info.cu_context, info.cu_device));  // keep, never destroy
      for (int j = 0; j < n; ++j) {
        int ok = 0;
        if (i != j && cudaDeviceCanAccessPeer(&ok, i, j) == cudaSuccess && ok) info.peers.push_back(j);
      }
      t.push_back(info);
    }
    HPCP_CUDA(cudaSetDevice(prev));
    return t;
  }();
  return table;
}

inline const DeviceNativeInfo& xcuda_get_device_info(int n) { return xcuda_get_infos_devices().at(n); }

}  // namespace hpcp
