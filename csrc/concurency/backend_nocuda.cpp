// Stub used by the host-only build (`make omp_con`): no CUDA backend linked in.
#include "bench.hpp"

namespace hpcp {
namespace con {
std::unique_ptr<Backend> make_cuda_backend(std::string* why_not) {
  if (why_not) *why_not = "binary built without CUDA";
  return nullptr;
}
}  // namespace con
}  // namespace hpcp
