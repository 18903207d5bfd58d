// concurency — "does running independent GPU commands concurrently beat running
// them one after the other?"  Self-judging benchmark with autotuning and a
// SUCCESS/FAILURE verdict per command group.
//
// Entry point only: picks the execution backend and hands argv to the host-pure
// driver (driver.cpp).  Backends: `cuda` (streams / graph / threads / fused
// persistent kernel, backend_cuda.cu), `cpu` (OpenMP host threads / tasks,
// backend_cpu.cpp), `fake:<spec>` (deterministic, for tests).
//   concurency <mode> [--backend cuda|cpu|fake:SPEC] [flags] --commands C M2D --commands ...
// Capability parity: concurency/main.cpp:115-322 of the reference.
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include "bench.hpp"
#include "driver.hpp"

int main(int argc, char** argv) {
  // CUDA backend: load every kernel at start-up — a first launch next to a running kernel would wait for it
  // (lazy module loading; see csrc/common/cuda_check.h::prefer_eager_module_loading).  Harmless for the CPU backend.
  (void)setenv("CUDA_MODULE_LOADING", "EAGER", /*overwrite=*/0);
  using namespace hpcp::con;
  std::vector<std::string> args(argv + 1, argv + argc);

  std::string choice;
  if (const char* env = std::getenv("HPCP_BACKEND")) choice = env;
  for (size_t i = 0; i < args.size();) {
    if (args[i] == "--backend" && i + 1 < args.size()) {
      choice = args[i + 1];
      args.erase(args.begin() + static_cast<long>(i), args.begin() + static_cast<long>(i) + 2);
    } else {
      ++i;
    }
  }

  std::unique_ptr<Backend> backend;
  try {
    if (choice.rfind("fake:", 0) == 0) {
      backend = make_fake_backend(choice.substr(5));
    } else if (choice == "cpu") {
      backend = make_cpu_backend();
    } else {
      std::string why;
      backend = make_cuda_backend(&why);
      if (!backend) {
        if (choice == "cuda") {
          std::cerr << "ERROR: CUDA backend unavailable: " << why << std::endl;
          return 1;
        }
        std::cerr << "# CUDA backend unavailable (" << why << "); using the CPU/OpenMP backend"
                  << std::endl;
        backend = make_cpu_backend();
      }
    }
    const std::string program = argc > 0 ? argv[0] : "concurency";
    return run(args, *backend, std::cout, std::cerr, program);
  } catch (const std::exception& e) {
    std::cerr << "ERROR: " << e.what() << std::endl;
    return 1;
  }
}
