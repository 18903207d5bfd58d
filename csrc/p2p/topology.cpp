// topology — NVLink / NVSwitch fabric discovery for one node.
//
//   topology            one line per connectivity plane: space-separated GPU ordinals
//   topology <k>        the k-th GPU of the flattened plane list (the device local
//                       rank k should bind to; consumed by scripts/tile_mapping.sh)
//   topology --matrix   P2P matrix: access / native atomics / performance rank per pair
//   topology --json     everything as one JSON object
//   topology --policy compact|spread|compact_plan --rank R [--ndev N]   rank -> device
//
// Capability parity: p2p/topology.cpp:28-107 of the reference (Level-Zero fabric
// ports -> planes).  Sources here: NVML NVLink state + remote device type, and
// the CUDA runtime's cudaDeviceCanAccessPeer / cudaDeviceGetP2PAttribute matrix.
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>

#include "topology_core.hpp"

namespace hpcp {
namespace topo {

// CUDA-runtime view: fills the pairwise attributes; if NVML was unavailable it
// also synthesises links from P2P accessibility (every accessible pair = a link).
static bool augment_with_cuda(Fabric* f, bool synthesize_links, std::string* why) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    (void)cudaGetLastError();
    if (why) *why = "CUDA runtime reports no device";
    return false;
  }
  if (f->gpus.empty()) {
    for (int g = 0; g < n; ++g) {
      GpuInfo gi;
      gi.index = g;
      cudaDeviceProp p{};
      if (cudaGetDeviceProperties(&p, g) == cudaSuccess) gi.name = p.name;
      char bus[32] = {0};
      if (cudaDeviceGetPCIBusId(bus, sizeof bus, g) == cudaSuccess) gi.pci_bus_id = bus;
      f->gpus.push_back(gi);
    }
    f->source = "cuda-p2p";
  }
  const int m = std::min<int>(n, static_cast<int>(f->gpus.size()));
  for (int a = 0; a < m; ++a)
    for (int b = 0; b < m; ++b) {
      if (a == b) continue;
      int acc = 0, atom = 0, rank = 0;
      (void)cudaDeviceGetP2PAttribute(&acc, cudaDevP2PAttrAccessSupported, a, b);
      (void)cudaDeviceGetP2PAttribute(&atom, cudaDevP2PAttrNativeAtomicSupported, a, b);
      (void)cudaDeviceGetP2PAttribute(&rank, cudaDevP2PAttrPerformanceRank, a, b);
      f->p2p_access[{a, b}] = acc;
      f->p2p_atomics[{a, b}] = atom;
      f->p2p_perf_rank[{a, b}] = rank;
      // Native atomics over the link distinguish NVLink from PCIe P2P.
      if (synthesize_links && acc && atom && a < b)
        f->links.push_back({std::to_string(a), std::to_string(b)});
    }
  (void)cudaGetLastError();
  return true;
}

bool discover_fabric(Fabric* out, std::string* why) {
  if (const char* fake = std::getenv("HPCP_FAKE_TOPOLOGY")) return fabric_from_fake(fake, out, why);
  std::string why_nvml, why_cuda;
  const bool have_nvml = fabric_from_nvml(out, &why_nvml);
  const bool have_cuda = augment_with_cuda(out, /*synthesize_links=*/!have_nvml, &why_cuda);
  if (!have_nvml && !have_cuda) {
    if (why) *why = why_nvml + "; " + why_cuda;
    return false;
  }
  return true;
}

}  // namespace topo
}  // namespace hpcp

int main(int argc, char** argv) {
  using namespace hpcp::topo;
  bool matrix = false, json = false;
  std::string policy;
  int rank = -1, ndev = -1, index = -1;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if (a == "--matrix") {
      matrix = true;
    } else if (a == "--json") {
      json = true;
    } else if (a == "--policy" && i + 1 < argc) {
      policy = argv[++i];
    } else if (a == "--rank" && i + 1 < argc) {
      rank = std::atoi(argv[++i]);
    } else if (a == "--ndev" && i + 1 < argc) {
      ndev = std::atoi(argv[++i]);
    } else if (a == "-h" || a == "--help") {
      std::cout << "Usage: topology [k] [--matrix] [--json] [--policy compact|spread|compact_plan "
                   "--rank R [--ndev N]]\n"
                   "  no argument : one connectivity plane per line (GPU ordinals)\n"
                   "  k           : k-th GPU of the flattened planes (device for local rank k)\n"
                   "Environment: HPCP_FAKE_TOPOLOGY='<n>:switch' | '<n>:0-1,2-3' (testing)\n";
      return 0;
    } else if (!a.empty() && (std::isdigit(static_cast<unsigned char>(a[0])))) {
      index = std::atoi(a.c_str());
    } else {
      std::cerr << "topology: unknown argument '" << a << "'" << std::endl;
      return 1;
    }
  }

  Fabric fabric;
  std::string why;
  if (!discover_fabric(&fabric, &why)) {
    std::cerr << "topology: no fabric information: " << why << std::endl;
    return 1;
  }
  const int n = static_cast<int>(fabric.gpus.size());
  const auto planes = merge_planes(n, fabric.links);

  try {
    if (!policy.empty()) {
      if (rank < 0) {
        if (const char* lr = std::getenv("LOCAL_RANK")) rank = std::atoi(lr);
      }
      if (rank < 0) {
        std::cerr << "topology: --policy needs --rank R (or $LOCAL_RANK)" << std::endl;
        return 1;
      }
      std::cout << device_for_rank(policy, rank, ndev > 0 ? ndev : n, planes) << std::endl;
      return 0;
    }
    if (index >= 0) {
      const auto flat = flatten(planes);
      if (index >= static_cast<int>(flat.size())) {
        std::cerr << "topology: index " << index << " out of range (" << flat.size() << " GPUs)"
                  << std::endl;
        return 1;
      }
      std::cout << flat[index] << std::endl;
      return 0;
    }
  } catch (const std::exception& e) {
    std::cerr << "topology: " << e.what() << std::endl;
    return 1;
  }

  if (json) {
    std::cout << to_json(fabric, planes) << std::endl;
    return 0;
  }
  for (const auto& p : planes) {
    for (int g : p) std::cout << g << " ";
    std::cout << std::endl;
  }
  if (matrix) {
    std::cout << "# source: " << fabric.source << std::endl;
    for (const auto& g : fabric.gpus)
      std::cout << "# GPU " << g.index << ": " << g.name << " pci=" << g.pci_bus_id
                << " numa=" << g.numa_node << " nvlinks=" << g.nvlinks_active
                << " (to switch: " << g.nvlinks_to_switch << ")" << std::endl;
    std::cout << "# P2P matrix (access/atomics/perf-rank), row = src, col = dst" << std::endl;
    for (int a = 0; a < n; ++a) {
      std::cout << "#  " << a << ":";
      for (int b = 0; b < n; ++b) {
        if (a == b) {
          std::cout << "   -   ";
          continue;
        }
        auto acc = fabric.p2p_access.find({a, b});
        if (acc == fabric.p2p_access.end()) {
          std::cout << "   ?   ";
          continue;
        }
        std::cout << " " << acc->second << "/" << fabric.p2p_atomics[{a, b}] << "/"
                  << fabric.p2p_perf_rank[{a, b}] << " ";
      }
      std::cout << std::endl;
    }
  }
  return 0;
}
