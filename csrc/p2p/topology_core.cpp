#include "topology_core.hpp"

#include <dlfcn.h>
#include <nvml.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <numeric>
#include <sstream>
#include <stdexcept>

namespace hpcp {
namespace topo {

namespace {

bool is_gpu_id(const std::string& s) {
  return !s.empty() && std::all_of(s.begin(), s.end(), [](char c) { return c >= '0' && c <= '9'; });
}

struct UnionFind {
  std::vector<int> parent;
  explicit UnionFind(int n) : parent(n) { std::iota(parent.begin(), parent.end(), 0); }
  int find(int x) {
    while (parent[x] != x) x = parent[x] = parent[parent[x]];
    return x;
  }
  void unite(int a, int b) { parent[find(a)] = find(b); }
};

std::string lower(std::string s) {
  for (auto& c : s) c = static_cast<char>(std::tolower(static_cast<unsigned char>(c)));
  return s;
}

}  // namespace

std::vector<std::vector<int>> merge_planes(int n_gpus, const std::vector<LinkSet>& links) {
  // Node ids: GPUs 0..n-1, then one id per distinct fabric-node name.
  std::map<std::string, int> fabric_id;
  auto node_of = [&](const std::string& name) -> int {
    if (is_gpu_id(name)) {
      const int g = std::atoi(name.c_str());
      return g < n_gpus ? g : -1;
    }
    auto it = fabric_id.find(name);
    if (it == fabric_id.end()) it = fabric_id.emplace(name, static_cast<int>(fabric_id.size())).first;
    return n_gpus + it->second;
  };
  // First pass: count fabric nodes so the union-find is large enough.
  for (const auto& l : links)
    for (const auto& e : l) (void)node_of(e);
  UnionFind uf(n_gpus + static_cast<int>(fabric_id.size()));
  for (const auto& l : links) {
    int first = -1;
    for (const auto& e : l) {
      const int id = node_of(e);
      if (id < 0) continue;
      if (first < 0)
        first = id;
      else
        uf.unite(first, id);
    }
  }
  std::map<int, std::vector<int>> by_root;
  for (int g = 0; g < n_gpus; ++g) by_root[uf.find(g)].push_back(g);
  std::vector<std::vector<int>> planes;
  for (auto& kv : by_root) {
    std::sort(kv.second.begin(), kv.second.end());
    planes.push_back(kv.second);
  }
  std::sort(planes.begin(), planes.end(),
            [](const std::vector<int>& a, const std::vector<int>& b) { return a.front() < b.front(); });
  return planes;
}

std::vector<int> flatten(const std::vector<std::vector<int>>& planes) {
  std::vector<int> out;
  for (const auto& p : planes) out.insert(out.end(), p.begin(), p.end());
  return out;
}

int device_for_rank(const std::string& policy, int local_rank, int n_devices,
                    const std::vector<std::vector<int>>& planes, int n_domains) {
  if (n_devices <= 0) throw std::invalid_argument("device_for_rank: no devices");
  if (local_rank < 0) throw std::invalid_argument("device_for_rank: negative rank");
  if (policy == "compact") return local_rank % n_devices;
  if (policy == "spread") {
    // Round-robin over the domains (contiguous blocks; the first n % d hold one more GPU), skipping exhausted ones:
    // every GPU is used for any count; equal blocks give (r % d) * per + r / d.  Same rule in
    // parallel/tile_mapping.py and scripts/tile_mapping.sh.
    const int d = std::max(1, std::min(n_domains, n_devices));
    const int base = n_devices / d, extra = n_devices % d;
    const int r = local_rank % n_devices;
    int seen = 0;
    for (int level = 0;; ++level)
      for (int k = 0; k < d; ++k)
        if (level < base + (k < extra ? 1 : 0)) {
          if (seen == r) return k * base + std::min(k, extra) + level;
          ++seen;
        }
  }
  if (policy == "compact_plan") {
    const std::vector<int> flat = flatten(planes);
    if (flat.empty()) return local_rank % n_devices;
    return flat[local_rank % static_cast<int>(flat.size())];
  }
  throw std::invalid_argument("unknown mapping policy '" + policy +
                              "' (compact | spread | compact_plan)");
}

bool fabric_from_fake(const std::string& spec, Fabric* out, std::string* why) {
  // "<n>:<a>-<b>,<c>-<d>,..."  explicit GPU-GPU links, or "<n>:switch" = all on NVSwitch,
  // or "<n>:switch:0-3;4-7" = two switch planes.
  const auto colon = spec.find(':');
  if (colon == std::string::npos) {
    if (why) *why = "fake topology must look like '<n_gpus>:<links>'";
    return false;
  }
  const int n = std::atoi(spec.substr(0, colon).c_str());
  if (n <= 0 || n > 1024) {
    if (why) *why = "fake topology: bad GPU count";
    return false;
  }
  Fabric f;
  f.source = "fake";
  for (int g = 0; g < n; ++g) {
    GpuInfo gi;
    gi.index = g;
    gi.name = "FakeGPU";
    gi.uuid = "GPU-fake-" + std::to_string(g);
    gi.pci_bus_id = "0000:" + std::to_string(g) + ":00.0";
    gi.numa_node = g < n / 2 ? 0 : 1;
    f.gpus.push_back(gi);
  }
  std::string rest = spec.substr(colon + 1);
  if (rest.rfind("switch", 0) == 0) {
    std::vector<std::vector<int>> groups;
    const auto c2 = rest.find(':');
    if (c2 == std::string::npos) {
      groups.emplace_back();
      for (int g = 0; g < n; ++g) groups.back().push_back(g);
    } else {
      std::stringstream ss(rest.substr(c2 + 1));
      std::string grp;
      while (std::getline(ss, grp, ';')) {
        const auto dash = grp.find('-');
        if (dash == std::string::npos) continue;
        const int a = std::atoi(grp.substr(0, dash).c_str()), b = std::atoi(grp.substr(dash + 1).c_str());
        groups.emplace_back();
        for (int g = a; g <= b && g < n; ++g) groups.back().push_back(g);
      }
    }
    for (size_t k = 0; k < groups.size(); ++k)
      for (int g : groups[k]) {
        f.links.push_back({std::to_string(g), "nvswitch" + std::to_string(k)});
        f.gpus[g].nvlinks_active += 18;
        f.gpus[g].nvlinks_to_switch += 18;
      }
  } else {
    std::stringstream ss(rest);
    std::string item;
    while (std::getline(ss, item, ',')) {
      const auto dash = item.find('-');
      if (dash == std::string::npos) continue;
      const int a = std::atoi(item.substr(0, dash).c_str()), b = std::atoi(item.substr(dash + 1).c_str());
      if (a < 0 || b < 0 || a >= n || b >= n) continue;
      f.links.push_back({std::to_string(a), std::to_string(b)});
      f.gpus[a].nvlinks_active++;
      f.gpus[b].nvlinks_active++;
    }
  }
  *out = f;
  return true;
}

// ------------------------------------------------------------------ NVML ----
namespace {

struct Nvml {
  void* lib = nullptr;
  nvmlReturn_t (*Init)() = nullptr;
  nvmlReturn_t (*Shutdown)() = nullptr;
  nvmlReturn_t (*DeviceGetCount)(unsigned*) = nullptr;
  nvmlReturn_t (*DeviceGetHandleByIndex)(unsigned, nvmlDevice_t*) = nullptr;
  nvmlReturn_t (*DeviceGetName)(nvmlDevice_t, char*, unsigned) = nullptr;
  nvmlReturn_t (*DeviceGetUUID)(nvmlDevice_t, char*, unsigned) = nullptr;
  nvmlReturn_t (*DeviceGetPciInfo)(nvmlDevice_t, nvmlPciInfo_t*) = nullptr;
  nvmlReturn_t (*DeviceGetNvLinkState)(nvmlDevice_t, unsigned, nvmlEnableState_t*) = nullptr;
  nvmlReturn_t (*DeviceGetNvLinkRemotePciInfo)(nvmlDevice_t, unsigned, nvmlPciInfo_t*) = nullptr;
  nvmlReturn_t (*DeviceGetNvLinkRemoteDeviceType)(nvmlDevice_t, unsigned,
                                                  nvmlIntNvLinkDeviceType_t*) = nullptr;
  nvmlReturn_t (*DeviceGetNumaNodeId)(nvmlDevice_t, unsigned*) = nullptr;

  template <typename Fn>
  bool sym(Fn& fn, const char* name, bool required = true) {
    fn = reinterpret_cast<Fn>(dlsym(lib, name));
    return fn != nullptr || !required;
  }
  bool load(std::string* why) {
    for (const char* n : {"libnvidia-ml.so.1", "libnvidia-ml.so"}) {
      lib = dlopen(n, RTLD_LAZY | RTLD_LOCAL);
      if (lib) break;
    }
    if (!lib) {
      if (why) *why = "libnvidia-ml.so.1 not found (no NVIDIA driver on this machine)";
      return false;
    }
    bool ok = sym(Init, "nvmlInit_v2") && sym(Shutdown, "nvmlShutdown") &&
              sym(DeviceGetCount, "nvmlDeviceGetCount_v2") &&
              sym(DeviceGetHandleByIndex, "nvmlDeviceGetHandleByIndex_v2") &&
              sym(DeviceGetName, "nvmlDeviceGetName") && sym(DeviceGetUUID, "nvmlDeviceGetUUID") &&
              sym(DeviceGetPciInfo, "nvmlDeviceGetPciInfo_v3") &&
              sym(DeviceGetNvLinkState, "nvmlDeviceGetNvLinkState") &&
              sym(DeviceGetNvLinkRemotePciInfo, "nvmlDeviceGetNvLinkRemotePciInfo_v2") &&
              sym(DeviceGetNvLinkRemoteDeviceType, "nvmlDeviceGetNvLinkRemoteDeviceType", false) &&
              sym(DeviceGetNumaNodeId, "nvmlDeviceGetNumaNodeId", false);
    if (!ok && why) *why = "NVML library lacks required symbols";
    return ok;
  }
  ~Nvml() {
    if (lib) dlclose(lib);
  }
};

}  // namespace

bool fabric_from_nvml(Fabric* out, std::string* why) {
  Nvml nv;
  if (!nv.load(why)) return false;
  if (nv.Init() != NVML_SUCCESS) {
    if (why) *why = "nvmlInit failed";
    return false;
  }
  Fabric f;
  f.source = "nvml";
  unsigned n = 0;
  if (nv.DeviceGetCount(&n) != NVML_SUCCESS || n == 0) {
    nv.Shutdown();
    if (why) *why = "NVML reports no GPU";
    return false;
  }
  std::vector<nvmlDevice_t> devs(n);
  std::map<std::string, int> by_bus;
  for (unsigned i = 0; i < n; ++i) {
    GpuInfo g;
    g.index = static_cast<int>(i);
    if (nv.DeviceGetHandleByIndex(i, &devs[i]) != NVML_SUCCESS) continue;
    char buf[128] = {0};
    if (nv.DeviceGetName(devs[i], buf, sizeof buf) == NVML_SUCCESS) g.name = buf;
    if (nv.DeviceGetUUID(devs[i], buf, sizeof buf) == NVML_SUCCESS) g.uuid = buf;
    nvmlPciInfo_t pci{};
    if (nv.DeviceGetPciInfo(devs[i], &pci) == NVML_SUCCESS) g.pci_bus_id = lower(pci.busId);
    unsigned numa = 0;
    if (nv.DeviceGetNumaNodeId && nv.DeviceGetNumaNodeId(devs[i], &numa) == NVML_SUCCESS)
      g.numa_node = static_cast<int>(numa);
    by_bus[g.pci_bus_id] = g.index;
    f.gpus.push_back(g);
  }
  for (unsigned i = 0; i < n; ++i) {
    for (unsigned link = 0; link < NVML_NVLINK_MAX_LINKS; ++link) {
      nvmlEnableState_t st = NVML_FEATURE_DISABLED;
      if (nv.DeviceGetNvLinkState(devs[i], link, &st) != NVML_SUCCESS) continue;
      if (st != NVML_FEATURE_ENABLED) continue;
      f.gpus[i].nvlinks_active++;
      bool to_switch = false;
      if (nv.DeviceGetNvLinkRemoteDeviceType) {
        nvmlIntNvLinkDeviceType_t t = NVML_NVLINK_DEVICE_TYPE_UNKNOWN;
        if (nv.DeviceGetNvLinkRemoteDeviceType(devs[i], link, &t) == NVML_SUCCESS)
          to_switch = (t == NVML_NVLINK_DEVICE_TYPE_SWITCH);
      }
      nvmlPciInfo_t rpci{};
      const bool have_remote =
          nv.DeviceGetNvLinkRemotePciInfo(devs[i], link, &rpci) == NVML_SUCCESS;
      const std::string rbus = have_remote ? lower(rpci.busId) : "";
      auto it = by_bus.find(rbus);
      if (!to_switch && it != by_bus.end()) {
        f.links.push_back({std::to_string(i), std::to_string(it->second)});
      } else {
        // NVSwitch (or an endpoint that is not a visible GPU): one shared fabric node —
        // all NVSwitches of an HGX baseboard form a single non-blocking plane.
        f.gpus[i].nvlinks_to_switch++;
        f.links.push_back({std::to_string(i), "nvswitch"});
      }
    }
  }
  nv.Shutdown();
  *out = f;
  return true;
}

#if !defined(HPCP_TOPOLOGY_WITH_CUDA)
bool discover_fabric(Fabric* out, std::string* why) {
  if (const char* fake = std::getenv("HPCP_FAKE_TOPOLOGY")) return fabric_from_fake(fake, out, why);
  return fabric_from_nvml(out, why);
}
#endif

std::string to_json(const Fabric& f, const std::vector<std::vector<int>>& planes) {
  std::ostringstream os;
  os << "{\"source\":\"" << f.source << "\",\"gpus\":[";
  for (size_t i = 0; i < f.gpus.size(); ++i) {
    const auto& g = f.gpus[i];
    os << (i ? "," : "") << "{\"index\":" << g.index << ",\"name\":\"" << g.name << "\",\"uuid\":\""
       << g.uuid << "\",\"pci\":\"" << g.pci_bus_id << "\",\"numa\":" << g.numa_node
       << ",\"nvlinks_active\":" << g.nvlinks_active
       << ",\"nvlinks_to_switch\":" << g.nvlinks_to_switch
       << ",\"multicast\":" << (g.multicast ? "true" : "false") << "}";
  }
  os << "],\"planes\":[";
  for (size_t p = 0; p < planes.size(); ++p) {
    os << (p ? "," : "") << "[";
    for (size_t k = 0; k < planes[p].size(); ++k) os << (k ? "," : "") << planes[p][k];
    os << "]";
  }
  os << "],\"p2p\":[";
  bool first = true;
  for (const auto& kv : f.p2p_access) {
    const auto key = kv.first;
    os << (first ? "" : ",") << "{\"src\":" << key.first << ",\"dst\":" << key.second
       << ",\"access\":" << kv.second;
    auto a = f.p2p_atomics.find(key);
    if (a != f.p2p_atomics.end()) os << ",\"atomics\":" << a->second;
    auto r = f.p2p_perf_rank.find(key);
    if (r != f.p2p_perf_rank.end()) os << ",\"perf_rank\":" << r->second;
    os << "}";
    first = false;
  }
  os << "]}";
  return os.str();
}

}  // namespace topo
}  // namespace hpcp
