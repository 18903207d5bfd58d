// Host-only self-test of the native runtime pieces that need no GPU: the thread-per-rank runtime,
// the device-subset policy, the dtype trait table, plane clustering and rank->device policies, and
// the concurrency driver's pure functions.  Built as bin/native_selftest; run by tests/test_native_selftest.py
// (the reference has no unit tests at all — SURVEY.md §4).
#include <atomic>
#include <cmath>
#include <cstdio>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../common/dtype_traits.h"
#include "../common/rank_runtime.h"
#include "../concurency/driver.hpp"
#include "../miniapps/devices.hpp"
#include "../p2p/topology_core.hpp"

namespace {

int failures = 0;
#define CHECK(cond)                                                              \
  do {                                                                           \
    if (!(cond)) {                                                               \
      ++failures;                                                                \
      std::fprintf(stderr, "CHECK failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__); \
    }                                                                            \
  } while (0)

void test_rank_runtime() {
  using namespace hpcp;
  for (int world : {1, 2, 8}) {
    std::atomic<int> visited{0};
    std::vector<double> maxes(world), mins(world), sums(world);
    run_ranks(world, [&](RankCtx& ctx) {
      CHECK(ctx.world == world);
      for (int it = 0; it < 50; ++it) {
        ctx.barrier();
        maxes[ctx.rank] = ctx.max(static_cast<double>(ctx.rank * 10 + it));
        mins[ctx.rank] = ctx.min(static_cast<double>(ctx.rank + it));
        sums[ctx.rank] = ctx.sum(1.0);
      }
      visited.fetch_add(1);
    });
    CHECK(visited.load() == world);
    for (int r = 0; r < world; ++r) {
      CHECK(maxes[r] == (world - 1) * 10 + 49);
      CHECK(mins[r] == 49);
      CHECK(sums[r] == world);
    }
  }
  // A failing rank must not leave the others stuck in a barrier; the root cause is rethrown.
  bool threw = false;
  try {
    run_ranks(4, [&](RankCtx& ctx) {
      if (ctx.rank == 2) throw std::runtime_error("rank 2 exploded");
      ctx.barrier();
      ctx.barrier();
    });
  } catch (const std::exception& e) {
    threw = std::string(e.what()) == "rank 2 exploded";
  }
  CHECK(threw);
}

void test_devices_and_dtypes() {
  using namespace hpcp;
  CHECK(device_subset(3, 8, 8) == std::vector<int>({3}));
  CHECK(device_subset(1, 4, 8) == std::vector<int>({2, 3}));        // contiguous block of ndev/size
  CHECK(device_subset(5, 8, 2) == std::vector<int>({1}));           // oversubscription: round-robin
  CHECK(device_subset(0, 4, 0).empty());
  CHECK(primary_device(5, 8, 2) == 1);
  CHECK(primary_device(1, 4, 8) == 3);                               // subset {2,3}[1 % 2]
  CHECK(primary_device(0, 1, 0) == -1);
  CHECK(std::string(get_dtype<float>().name) == "float" && get_dtype<float>().reducible);
  CHECK(std::string(get_dtype<int>().name) == "int" && get_dtype<int>().elem == ElemType::kInt);
  CHECK(!get_dtype<double>().reducible && get_dtype<double>().size == 8);
  struct Opaque { char x[24]; };
  CHECK(std::string(get_dtype<Opaque>().name) == "bytes" && get_dtype<Opaque>().size == 24);  // MPI_BYTE analogue
  ElemType t;
  CHECK(elem_type_from_name("int32", &t) && t == ElemType::kInt);
  CHECK(!elem_type_from_name("complex", &t));
}

void test_topology() {
  using namespace hpcp::topo;
  Fabric f;
  std::string why;
  CHECK(fabric_from_fake("8:switch", &f, &why));
  auto planes = merge_planes(8, f.links);
  CHECK(planes.size() == 1 && planes[0].size() == 8);
  CHECK(fabric_from_fake("6:0-2,2-4,1-3,3-5", &f, &why));
  planes = merge_planes(6, f.links);
  CHECK(planes.size() == 2 && planes[0] == std::vector<int>({0, 2, 4}) && planes[1] == std::vector<int>({1, 3, 5}));
  CHECK(flatten(planes) == std::vector<int>({0, 2, 4, 1, 3, 5}));
  CHECK(device_for_rank("compact", 9, 8, planes) == 1);
  CHECK(device_for_rank("spread", 1, 8, planes) == 4);
  CHECK(device_for_rank("compact_plan", 3, 6, planes) == 1);
  bool threw = false;
  try {
    device_for_rank("diagonal", 0, 8, planes);
  } catch (const std::invalid_argument&) {
    threw = true;
  }
  CHECK(threw);
  CHECK(!fabric_from_fake("nonsense", &f, &why));
  CHECK(to_json(f, planes).find("\"planes\":[[0,2,4],[1,3,5]]") != std::string::npos);
}

void test_driver_pure_functions() {
  using namespace hpcp::con;
  CHECK(strip_twos("M2D") == "MD" && strip_twos("22") == "");
  CHECK(is_compute_command("C") && is_compute_command("T") && !is_compute_command("MD"));
  CHECK(tuned_parameter_of("C") == "tripcount_C" && tuned_parameter_of("T") == "tripcount_T" &&
        tuned_parameter_of("DP") == "globalsize_DP");
  Params p{{"globalsize_MD", 250000000}, {"globalsize_C", 1}};
  CHECK(bytes_moved({"C", "MD"}, p, 4) == 1000000000ull);
  CHECK(bytes_moved({"MD", "MD", "MD", "MD", "MD"}, p, 4) == 5000000000ull);   // > 2^32: no wrap
  CHECK(time_with_bandwidth(1000, 0) == "1000us");
  CHECK(time_with_bandwidth(1000, 1000000) == "1000us (1 GBytes/s)");
  CHECK(judge(2.0, 1.9, 10, -1, 100) == Verdict::kSuccess);
  CHECK(judge(2.0, 1.5, 10, -1, 100) == Verdict::kFarFromTheoretical);          // 2.0 >= 1.3 * 1.5
  CHECK(judge(2.0, 2.0, 10, 50, 100) == Verdict::kBandwidthFloor);
  CHECK(verdict_line("fused", {"C", "DP"}, Verdict::kSuccess) ==
        "## fused | C DP | SUCCESS: Close from Theoretical Speedup");
  auto fake = make_fake_backend("C=0.01,MD=0.0005,overlap=1.0");
  bool usage = false;
  try {
    parse_arguments({"in_order", "--commands", "Q"}, *fake);
  } catch (const UsageError&) {
    usage = true;
  }
  CHECK(usage);
  const Options opt = parse_arguments({"fused", "--queues", "3", "--commands", "C", "M2D", "--commands", "D2P"}, *fake);
  CHECK(opt.mode == "fused" && opt.n_queues == 3 && opt.groups.size() == 2 && opt.groups[0][1] == "MD");
  Params params = resolve_parameters(opt);
  CHECK(params.at("tripcount_C") == 40000 && params.at("globalsize_C") == 1 &&
        params.at("globalsize_MD") == 250000000);
}

}  // namespace

int main() {
  test_rank_runtime();
  test_devices_and_dtypes();
  test_topology();
  test_driver_pure_functions();
  if (failures == 0) {
    std::cout << "native selftest: OK" << std::endl;
    return 0;
  }
  std::cout << "native selftest: " << failures << " failure(s)" << std::endl;
  return 1;
}
