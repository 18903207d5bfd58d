// Thread-per-rank runtime for the native CLIs.
//
// The reference is launched as `mpirun -n N` with one MPI rank per GPU tile
// (p2p/run.sh:17, aurora.mpich.miniapps/src/CMakeLists.txt:41,49) and uses MPI
// for bootstrap (MPI_Init/Comm_rank/Comm_size), MPI_Barrier, and host-scalar
// reductions of timestamps (p2p/peer2pear.cpp:49-50, allreduce-mpi-sycl.cpp:189).
// An NVSwitch node is one address space away from that: a single process can
// peer-map all 8 GPUs, so a "rank" here is a host thread bound to one device.
// The rank body sees the same primitives (rank, size, barrier, min/max/sum
// reductions); data never moves through this layer — only GPU kernels move data.
#pragma once

#include <condition_variable>
#include <exception>
#include <functional>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace hpcp {

class RankGroup {
 public:
  explicit RankGroup(int world) : world_(world), slots_(world, 0.0) {}
  int world() const { return world_; }

  // Sense-reversing barrier; returns early (throwing) if any rank aborted.
  void barrier() {
    std::unique_lock<std::mutex> lk(mu_);
    if (aborted_) throw std::runtime_error("rank group aborted");
    const unsigned long gen = generation_;
    if (++arrived_ == world_) {
      arrived_ = 0;
      ++generation_;
      cv_.notify_all();
      return;
    }
    cv_.wait(lk, [&] { return generation_ != gen || aborted_; });
    if (aborted_ && generation_ == gen) throw std::runtime_error("rank group aborted");
  }

  enum class Op { kMin, kMax, kSum };

  double allreduce(int rank, double v, Op op) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      slots_[rank] = v;
    }
    barrier();
    double r = slots_[0];
    for (int i = 1; i < world_; ++i) {
      const double x = slots_[i];
      r = op == Op::kMin ? (x < r ? x : r) : op == Op::kMax ? (x > r ? x : r) : r + x;
    }
    barrier();  // nobody overwrites a slot before everyone has read
    return r;
  }
  double allreduce_max(int rank, double v) { return allreduce(rank, v, Op::kMax); }
  double allreduce_min(int rank, double v) { return allreduce(rank, v, Op::kMin); }
  double allreduce_sum(int rank, double v) { return allreduce(rank, v, Op::kSum); }

  void abort() {
    std::lock_guard<std::mutex> lk(mu_);
    aborted_ = true;
    cv_.notify_all();
  }

 private:
  const int world_;
  std::mutex mu_;
  std::condition_variable cv_;
  int arrived_ = 0;
  unsigned long generation_ = 0;
  bool aborted_ = false;
  std::vector<double> slots_;
};

struct RankCtx {
  int rank = 0;
  int world = 1;
  RankGroup* group = nullptr;
  void barrier() const { group->barrier(); }
  double max(double v) const { return group->allreduce_max(rank, v); }
  double min(double v) const { return group->allreduce_min(rank, v); }
  double sum(double v) const { return group->allreduce_sum(rank, v); }
};

// Runs body(ctx) on `world` threads; rethrows the first exception on the caller.
inline void run_ranks(int world, const std::function<void(RankCtx&)>& body) {
  RankGroup group(world);
  std::vector<std::thread> threads;
  std::vector<std::exception_ptr> errors(world);
  for (int r = 0; r < world; ++r) {
    threads.emplace_back([&, r] {
      RankCtx ctx;
      ctx.rank = r;
      ctx.world = world;
      ctx.group = &group;
      try {
        body(ctx);
      } catch (...) {
        errors[r] = std::current_exception();
        group.abort();
      }
    });
  }
  for (auto& t : threads) t.join();
  // Prefer a root-cause error over the secondary "rank group aborted" ones.
  std::exception_ptr first;
  for (auto& e : errors) {
    if (!e) continue;
    try {
      std::rethrow_exception(e);
    } catch (const std::exception& ex) {
      if (std::string(ex.what()) != "rank group aborted") std::rethrow_exception(e);
      if (!first) first = e;
    }
  }
  if (first) std::rethrow_exception(first);
}

}  // namespace hpcp
