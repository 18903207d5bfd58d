// CPU / OpenMP backend of the concurrency benchmark.
//
// Role: the "no GPU" plumbing configuration (BASELINE.json config #1).  The
// reference's OpenMP backend is target-offload code that falls back to the host
// when no device exists (concurency/bench_omp.cpp:65-117, modes `host_threads`
// and `nowait` chosen by -DHOST_THREADS / -DNOWAIT at build time,
// run_omp.sh:6-7).  This backend runs the same command set on host threads
// with the mode chosen at run time:
//   host_threads : one OpenMP thread per command (↔ `omp parallel for` over commands)
//   nowait       : one producer creates a deferred task per command, the team
//                  executes them, `taskwait` joins (↔ `target ... nowait` + taskwait)
// Every memory letter maps to host memory (D = 64-byte aligned allocation).
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <limits>
#include <numeric>
#include <stdexcept>

#if defined(_OPENMP)
#include <omp.h>
#endif

#include "bench.hpp"
#include "driver.hpp"

namespace hpcp {
namespace con {

namespace {

using Clock = std::chrono::steady_clock;

long elapsed_us(Clock::time_point a, Clock::time_point b) {
  return static_cast<long>(std::chrono::duration_cast<std::chrono::microseconds>(b - a).count());
}

struct HostCommand {
  std::string name;
  size_t n = 0;            // elements / work-items
  size_t tripcount = 0;    // C only
  float* buf[3] = {nullptr, nullptr, nullptr};

  void allocate() {
    auto make = [&](size_t elems) {
      void* p = nullptr;
      if (posix_memalign(&p, 64, std::max<size_t>(elems, 1) * sizeof(float)) != 0)
        throw std::runtime_error("host allocation failed");
      std::memset(p, 0, std::max<size_t>(elems, 1) * sizeof(float));
      return static_cast<float*>(p);
    };
    if (name == "C") {
      buf[0] = make(n);
    } else if (name == "A") {
      buf[0] = make(n);
      buf[1] = make(n);
      buf[2] = make(n);
    } else {
      buf[0] = make(n);
      buf[1] = make(n);
    }
  }
  void release() {
    for (auto& b : buf) {
      std::free(b);
      b = nullptr;
    }
  }
  void execute() const {
    if (name == "C") {
      for (size_t j = 0; j < n; ++j) buf[0][j] = busy_wait<float>(tripcount, static_cast<float>(j));
    } else if (name == "A") {
      float* __restrict__ a = buf[0];
      const float* __restrict__ b = buf[1];
      const float* __restrict__ c = buf[2];
      for (size_t j = 0; j < n; ++j) a[j] = b[j] + 3.0f * c[j];
    } else {
      std::memcpy(buf[1], buf[0], n * sizeof(float));
    }
  }
};

class CpuBackend final : public Backend {
 public:
  std::string name() const override {
#if defined(_OPENMP)
    return "cpu-openmp";
#else
    return "cpu-serial(no OpenMP)";
#endif
  }
  std::vector<std::string> modes() const override { return {"nowait", "host_threads"}; }  // order of the reference usage line (bench_omp.cpp:12)
  std::string memory_letters() const override { return "MDHS"; }
  std::string compute_letters() const override { return "CA"; }

  BenchResult run(const BenchRequest& req) override {
    const size_t nc = req.commands.size();
    int n_queues = req.n_queues;
    // -1: one host thread per command (deferred tasks also need a team to run on).
    if (n_queues == -1) n_queues = req.mode == "serial" ? 1 : static_cast<int>(nc);
    if (req.verbose) std::cout << "#n_host_threads used: " << n_queues << std::endl;

    std::vector<HostCommand> cmds(nc);
    try {
      for (size_t i = 0; i < nc; ++i) {
        cmds[i].name = req.commands[i];
        cmds[i].n = require_param(req, "globalsize_" + req.commands[i]);
        if (cmds[i].name == "C") cmds[i].tripcount = require_param(req, "tripcount_C");
        cmds[i].allocate();
      }
    } catch (...) {
      for (auto& c : cmds) c.release();  // a failed allocation / missing parameter must not leak the earlier buffers
      throw;
    }

    BenchResult res;
    res.total_us = std::numeric_limits<long>::max();
    const bool serial = req.mode == "serial";
    if (serial) res.per_command_us.assign(nc, std::numeric_limits<long>::max());

    for (int r = 0; r < req.n_repetitions; ++r) {
      const auto t0 = Clock::now();
      if (serial) {
        for (size_t i = 0; i < nc; ++i) {
          const auto s = Clock::now();
          cmds[i].execute();
          res.per_command_us[i] = std::min(res.per_command_us[i], elapsed_us(s, Clock::now()));
        }
      } else if (req.mode == "host_threads") {
#if defined(_OPENMP)
#pragma omp parallel for num_threads(n_queues) schedule(static, 1)
#endif
        for (long i = 0; i < static_cast<long>(nc); ++i) cmds[i].execute();
      } else if (req.mode == "nowait") {
#if defined(_OPENMP)
#pragma omp parallel num_threads(n_queues)
#pragma omp single
        {
          for (size_t i = 0; i < nc; ++i) {
            const HostCommand* c = &cmds[i];
#pragma omp task firstprivate(c)
            c->execute();
          }
#pragma omp taskwait
        }
#else
        for (size_t i = 0; i < nc; ++i) cmds[i].execute();
#endif
      } else {
        for (auto& c : cmds) c.release();
        throw std::runtime_error("cpu backend: unknown mode '" + req.mode + "'");
      }
      const long t = elapsed_us(t0, Clock::now());
      if (req.verbose) std::cout << "#repetition " << r << ": " << t << " us" << std::endl;
      res.total_us = std::min(res.total_us, t);
    }
    // Best theoretical serial: never charge the serial run more than the sum of
    // its commands' best times (same rule as bench_omp.cpp:119-121).
    if (serial)
      res.total_us = std::min(
          res.total_us, std::accumulate(res.per_command_us.begin(), res.per_command_us.end(), 0L));

    for (auto& c : cmds) c.release();
    return res;
  }
};

}  // namespace

std::unique_ptr<Backend> make_cpu_backend() { return std::make_unique<CpuBackend>(); }

}  // namespace con
}  // namespace hpcp
