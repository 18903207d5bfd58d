// Peer-mapped ("symmetric") memory for one NVSwitch node.
//
// This is the memory side of what GPU-aware MPICH does for the reference
// (device pointers handed straight to MPI_Isend/MPI_Put/MPI_Send,
// p2p/peer2pear.cpp:34,79,121; allreduce-mpi-sycl.cpp:50-58): after setup,
// every rank holds a directly usable address for every other rank's buffer and
// kernels move the bytes themselves.
//
//  * single process (native CLIs): cudaDeviceEnablePeerAccess + UVA pointers;
//  * one process per GPU (torchrun / bench.py): CUDA IPC handles exported here,
//    exchanged by the Python front end over torch.distributed, opened here;
//  * NVSwitch multicast (NVLS): cuMulticastCreate + cuMemMap, single process;
//    multi-process multicast comes from torch's symmetric memory.
//
// Allocation kinds mirror the miniapp's -H/-D/-S switch
// (allreduce-mpi-sycl.cpp:112-121: usm host / device / shared).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <string>
#include <vector>

namespace hpcp {

enum class AllocKind : int {
  kDevice = 0,   // 'D' cudaMalloc
  kPinned = 1,   // 'H' cudaHostAlloc(portable|mapped)
  kManaged = 2,  // 'S' cudaMallocManaged
  kPageable = 3, // 'M' plain calloc (copy-engine commands only; kernels cannot touch it)
  kMapped = 4,   // 'R' host malloc *mapped* to the device: aligned_alloc + cudaHostRegister(mapped),
                 //     kernels use the device alias from cudaHostGetDevicePointer — the analogue of
                 //     OpenMP `target enter data map(alloc)` + `use_device_ptr`
                 //     (allreduce-map-mpi-omp-offload.cpp:113-115,38)
};

AllocKind alloc_kind_from_letter(char c);  // 'D','H','S','M','R'
const char* alloc_kind_name(AllocKind k);

void* alloc_bytes(size_t bytes, AllocKind kind, int device, bool zero);
void free_bytes(void* p, AllocKind kind);

// One buffer per rank, every copy addressable from every device of the group.
struct SymmetricBuffer {
  std::vector<void*> ptr;  // ptr[r] lives on devices[r]
  size_t bytes = 0;
  AllocKind kind = AllocKind::kDevice;
};

struct MulticastBuffer {
  std::vector<void*> uc;   // per-rank unicast mapping (peer-accessible)
  void* mc = nullptr;      // multicast mapping: st -> all ranks, ld_reduce -> sum over ranks
  size_t bytes = 0;        // usable bytes (mapping is rounded up to the granularity)
  size_t mapped = 0;
  std::vector<unsigned long long> mem_handles;
  unsigned long long mc_handle = 0;
};

// Owner of peer access + symmetric allocations of a single-process group.
class NodeMemory {
 public:
  // devices[r] = CUDA ordinal of rank r (duplicates allowed: oversubscription,
  // cf. aurora.mpich.miniapps/src/include/devices.hpp:46-47).
  explicit NodeMemory(std::vector<int> devices);
  ~NodeMemory();
  NodeMemory(const NodeMemory&) = delete;
  NodeMemory& operator=(const NodeMemory&) = delete;

  int world() const { return static_cast<int>(devices_.size()); }
  int device(int rank) const { return devices_[rank]; }
  const std::vector<int>& devices() const { return devices_; }

  SymmetricBuffer alloc(size_t bytes, AllocKind kind = AllocKind::kDevice, bool zero = true);
  void free(SymmetricBuffer& b);

  // Zeroed signal pads (kPadWords + extra_words 32-bit words per rank) and
  // one zeroed status word per rank (inside the same allocation, after the pad).
  SymmetricBuffer alloc_pads(size_t extra_words = 0);

  static bool multicast_supported(int device);
  MulticastBuffer alloc_multicast(size_t bytes);
  void free_multicast(MulticastBuffer& b);

 private:
  std::vector<int> devices_;
};

// Checks cudaDeviceCanAccessPeer for every distinct pair; returns a description
// of the first failing pair or an empty string.
std::string peer_access_problem(const std::vector<int>& devices);
// Enables access between every distinct pair (idempotent).
void enable_peer_access(const std::vector<int>& devices);

// ---- CUDA IPC (one process per GPU) ----
constexpr size_t kIpcHandleBytes = 64;
void ipc_export(void* device_ptr, unsigned char out[kIpcHandleBytes]);
void* ipc_open(const unsigned char handle[kIpcHandleBytes]);
void ipc_close(void* opened_ptr);

}  // namespace hpcp
