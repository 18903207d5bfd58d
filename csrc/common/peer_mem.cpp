#include "peer_mem.h"

#include <cstdlib>
#include <cstring>
#include <set>
#include <sstream>

#include "cuda_check.h"
#include "driver_api.h"
#include "signal_layout.h"

namespace hpcp {

namespace {

struct DeviceGuard {
  int prev = 0;
  explicit DeviceGuard(int d) {
    HPCP_CUDA(cudaGetDevice(&prev));
    if (prev != d) HPCP_CUDA(cudaSetDevice(d));
  }
  ~DeviceGuard() { (void)cudaSetDevice(prev); }
};

size_t round_up(size_t v, size_t g) { return (v + g - 1) / g * g; }

}  // namespace

AllocKind alloc_kind_from_letter(char c) {
  switch (c) {
    case 'D': return AllocKind::kDevice;
    case 'H': return AllocKind::kPinned;
    case 'S': return AllocKind::kManaged;
    case 'M': return AllocKind::kPageable;
    case 'R': return AllocKind::kMapped;
    default: HPCP_FAIL(std::string("unknown allocation letter '") + c + "'");
  }
}

const char* alloc_kind_name(AllocKind k) {
  switch (k) {
    case AllocKind::kDevice: return "device";
    case AllocKind::kPinned: return "pinned-host";
    case AllocKind::kManaged: return "managed";
    case AllocKind::kPageable: return "pageable-host";
    case AllocKind::kMapped: return "mapped-host-malloc";
  }
  return "?";
}

void* alloc_bytes(size_t bytes, AllocKind kind, int device, bool zero) {
  void* p = nullptr;
  const size_t n = bytes == 0 ? 16 : bytes;
  switch (kind) {
    case AllocKind::kDevice: {
      DeviceGuard g(device);
      HPCP_CUDA(cudaMalloc(&p, n));
      if (zero) HPCP_CUDA(cudaMemset(p, 0, n));
      break;
    }
    case AllocKind::kPinned: {
      DeviceGuard g(device);
      HPCP_CUDA(cudaHostAlloc(&p, n, cudaHostAllocPortable | cudaHostAllocMapped));
      if (zero) std::memset(p, 0, n);
      break;
    }
    case AllocKind::kManaged: {
      DeviceGuard g(device);
      HPCP_CUDA(cudaMallocManaged(&p, n, cudaMemAttachGlobal));
      if (zero) HPCP_CUDA(cudaMemset(p, 0, n));
      // Keep the pages resident on the owning GPU so peers reach them over NVLink.
      (void)cudaMemAdvise(p, n, cudaMemAdviseSetPreferredLocation, device);
      (void)cudaMemPrefetchAsync(p, n, device, 0);
      (void)cudaGetLastError();
      break;
    }
    case AllocKind::kPageable: {
      p = zero ? std::calloc(n, 1) : std::malloc(n);
      HPCP_REQUIRE(p != nullptr, "host allocation failed");
      break;
    }
    case AllocKind::kMapped: {
      DeviceGuard g(device);
      const size_t rounded = (n + 4095) / 4096 * 4096;
      void* host = std::aligned_alloc(4096, rounded);
      HPCP_REQUIRE(host != nullptr, "host allocation failed");
      if (zero) std::memset(host, 0, rounded);
      const cudaError_t e = cudaHostRegister(host, rounded, cudaHostRegisterPortable | cudaHostRegisterMapped);
      if (e != cudaSuccess) {
        std::free(host);
        HPCP_CUDA(e);
      }
      void* dev_alias = nullptr;
      HPCP_CUDA(cudaHostGetDevicePointer(&dev_alias, host, 0));
      // With unified addressing the alias equals the host pointer; the code below relies on it
      // to unregister/free through the same address.
      HPCP_REQUIRE(dev_alias == host, "mapped host memory has a distinct device alias on this platform");
      p = host;
      break;
    }
  }
  return p;
}

void free_bytes(void* p, AllocKind kind) {
  if (p == nullptr) return;
  switch (kind) {
    case AllocKind::kDevice:
    case AllocKind::kManaged: (void)cudaFree(p); break;
    case AllocKind::kPinned: (void)cudaFreeHost(p); break;
    case AllocKind::kPageable: std::free(p); break;
    case AllocKind::kMapped:
      (void)cudaHostUnregister(p);
      std::free(p);
      break;
  }
}

std::string peer_access_problem(const std::vector<int>& devices) {
  std::set<int> uniq(devices.begin(), devices.end());
  for (int a : uniq)
    for (int b : uniq) {
      if (a == b) continue;
      int ok = 0;
      HPCP_CUDA(cudaDeviceCanAccessPeer(&ok, a, b));
      if (!ok) {
        std::ostringstream os;
        os << "GPU " << a << " cannot peer-access GPU " << b
           << " (no NVLink/PCIe P2P path; check `topology` and CUDA_VISIBLE_DEVICES)";
        return os.str();
      }
    }
  return "";
}

void enable_peer_access(const std::vector<int>& devices) {
  const std::string problem = peer_access_problem(devices);
  HPCP_REQUIRE(problem.empty(), problem);
  std::set<int> uniq(devices.begin(), devices.end());
  for (int a : uniq) {
    DeviceGuard g(a);
    for (int b : uniq) {
      if (a == b) continue;
      const cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) {
        (void)cudaGetLastError();
        continue;
      }
      HPCP_CUDA(e);
    }
  }
}

NodeMemory::NodeMemory(std::vector<int> devices) : devices_(std::move(devices)) {
  HPCP_REQUIRE(!devices_.empty(), "NodeMemory: empty device list");
  enable_peer_access(devices_);
}

NodeMemory::~NodeMemory() = default;

SymmetricBuffer NodeMemory::alloc(size_t bytes, AllocKind kind, bool zero) {
  HPCP_REQUIRE(kind != AllocKind::kPageable,
               "pageable host memory cannot be a symmetric (kernel-visible) buffer");
  SymmetricBuffer b;
  b.bytes = bytes;
  b.kind = kind;
  try {
    for (int r = 0; r < world(); ++r) b.ptr.push_back(alloc_bytes(bytes, kind, devices_[r], zero));
  } catch (...) {
    free(b);  // do not leak the ranks that were already allocated
    throw;
  }
  return b;
}

void NodeMemory::free(SymmetricBuffer& b) {
  for (void* p : b.ptr) free_bytes(p, b.kind);
  b.ptr.clear();
  b.bytes = 0;
}

SymmetricBuffer NodeMemory::alloc_pads(size_t extra_words) {
  // The status word lives at word index kPadWords + extra_words (first word of the tail).
  return alloc((kPadWords + extra_words + kPadTailWords) * sizeof(uint32_t), AllocKind::kDevice, true);
}

bool NodeMemory::multicast_supported(int device) {
  if (!DriverApi::available()) return false;
  const DriverApi& d = DriverApi::get();
  CUdevice dev;
  if (d.cuDeviceGet(&dev, device) != CUDA_SUCCESS) return false;
  int v = 0;
  if (d.cuDeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev) != CUDA_SUCCESS)
    return false;
  return v != 0;
}

MulticastBuffer NodeMemory::alloc_multicast(size_t bytes) {
  const DriverApi& d = DriverApi::get();
  std::set<int> uniq(devices_.begin(), devices_.end());
  HPCP_REQUIRE(static_cast<int>(uniq.size()) == world(),
               "multicast needs one distinct GPU per rank");
  HPCP_REQUIRE(world() >= 2, "multicast needs at least 2 GPUs");
  for (int dev : devices_)
    HPCP_REQUIRE(multicast_supported(dev), "GPU does not support NVSwitch multicast (NVLS)");
  for (int dev : devices_) {  // make sure primary contexts exist
    DeviceGuard g(dev);
    HPCP_CUDA(cudaFree(nullptr));
  }

  MulticastBuffer mb;
  mb.bytes = bytes;

  try {
    CUmulticastObjectProp mprop{};
    mprop.numDevices = static_cast<unsigned>(world());
    mprop.handleTypes = 0;
    mprop.flags = 0;
    mprop.size = bytes;
    size_t mc_gran = 0;
    HPCP_CU(d.cuMulticastGetGranularity(&mc_gran, &mprop, CU_MULTICAST_GRANULARITY_RECOMMENDED));

    CUmemAllocationProp aprop{};
    aprop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    aprop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    aprop.location.id = devices_[0];
    aprop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_NONE;
    size_t mem_gran = 0;
    HPCP_CU(d.cuMemGetAllocationGranularity(&mem_gran, &aprop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
    const size_t gran = mc_gran > mem_gran ? mc_gran : mem_gran;
    mb.mapped = round_up(bytes == 0 ? 1 : bytes, gran);
    mprop.size = mb.mapped;

    CUmemGenericAllocationHandle mc_handle;
    HPCP_CU(d.cuMulticastCreate(&mc_handle, &mprop));
    mb.mc_handle = mc_handle;
    for (int dev : devices_) {
      CUdevice cudev;
      HPCP_CU(d.cuDeviceGet(&cudev, dev));
      HPCP_CU(d.cuMulticastAddDevice(mc_handle, cudev));
    }

    std::vector<CUmemAccessDesc> access;
    for (int dev : devices_) {
      CUmemAccessDesc a{};
      a.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
      a.location.id = dev;
      a.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
      access.push_back(a);
    }

    for (int r = 0; r < world(); ++r) {
      DeviceGuard g(devices_[r]);
      aprop.location.id = devices_[r];
      CUmemGenericAllocationHandle mem;
      HPCP_CU(d.cuMemCreate(&mem, mb.mapped, &aprop, 0));
      mb.mem_handles.push_back(mem);
      HPCP_CU(d.cuMulticastBindMem(mc_handle, 0, mem, 0, mb.mapped, 0));
      CUdeviceptr va = 0;
      HPCP_CU(d.cuMemAddressReserve(&va, mb.mapped, gran, 0, 0));
      mb.uc.push_back(reinterpret_cast<void*>(va));  // recorded first: free_multicast releases it on failure
      HPCP_CU(d.cuMemMap(va, mb.mapped, 0, mem, 0));
      HPCP_CU(d.cuMemSetAccess(va, mb.mapped, access.data(), access.size()));
      HPCP_CUDA(cudaMemset(reinterpret_cast<void*>(va), 0, mb.mapped));
      HPCP_CUDA(cudaDeviceSynchronize());
    }
    CUdeviceptr mc_va = 0;
    HPCP_CU(d.cuMemAddressReserve(&mc_va, mb.mapped, gran, 0, 0));
    mb.mc = reinterpret_cast<void*>(mc_va);
    HPCP_CU(d.cuMemMap(mc_va, mb.mapped, 0, mc_handle, 0));
    HPCP_CU(d.cuMemSetAccess(mc_va, mb.mapped, access.data(), access.size()));
  } catch (...) {
    free_multicast(mb);  // unmap / release whatever was created before the failure
    throw;
  }
  return mb;
}

void NodeMemory::free_multicast(MulticastBuffer& b) {
  if (!DriverApi::available()) return;
  const DriverApi& d = DriverApi::get();
  if (b.mc != nullptr) {
    (void)d.cuMemUnmap(reinterpret_cast<CUdeviceptr>(b.mc), b.mapped);
    (void)d.cuMemAddressFree(reinterpret_cast<CUdeviceptr>(b.mc), b.mapped);
  }
  for (size_t r = 0; r < b.uc.size(); ++r) {
    (void)d.cuMemUnmap(reinterpret_cast<CUdeviceptr>(b.uc[r]), b.mapped);
    (void)d.cuMemAddressFree(reinterpret_cast<CUdeviceptr>(b.uc[r]), b.mapped);
  }
  for (auto h : b.mem_handles) (void)d.cuMemRelease(h);
  if (b.mc_handle) (void)d.cuMemRelease(b.mc_handle);
  b = MulticastBuffer{};
}

void ipc_export(void* device_ptr, unsigned char out[kIpcHandleBytes]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == kIpcHandleBytes, "IPC handle size");
  cudaIpcMemHandle_t h;
  HPCP_CUDA(cudaIpcGetMemHandle(&h, device_ptr));
  std::memcpy(out, &h, kIpcHandleBytes);
}

void* ipc_open(const unsigned char handle[kIpcHandleBytes]) {
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, kIpcHandleBytes);
  void* p = nullptr;
  HPCP_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  return p;
}

void ipc_close(void* opened_ptr) {
  if (opened_ptr != nullptr) (void)cudaIpcCloseMemHandle(opened_ptr);
}

}  // namespace hpcp
