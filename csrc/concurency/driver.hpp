// Host-pure half of the concurrency benchmark: argument grammar, default
// parameters, autotuning, speed-up analysis, verdict lines and JSON rows.
// No CUDA, no OpenMP: everything here is unit-testable on a GPU-less box with
// the deterministic FakeBackend.
//
// Behavioural contract (what must stay compatible with the reference's
// concurency/main.cpp because scripts parse it — concurency/parse.py:20-26):
//   argv[1] = mode; flags --enable_profiling --verbose --queues N
//   --repetitions N --min_bandwidth F --tripcount_C N --globalsize_<CMD> N
//   --globalsize_default_memory N; repeated `--commands A B ...` groups
//   (main.cpp:143-196); every '2' in a command token is dropped (:14-19);
//   verdict line `## <mode> | <cmds> | SUCCESS|FAILURE: ...` (:310-319);
//   exit status 1 if any group failed (:321).
#pragma once

#include <iosfwd>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "bench.hpp"

namespace hpcp {
namespace con {

constexpr double kSpeedupTolerance = 0.3;   // main.cpp:12 TOL_SPEEDUP
constexpr double kUnbalanceWarning = 1.5;   // main.cpp:295

struct UsageError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

struct Options {
  std::string program = "concurency";
  std::string mode;
  bool enable_profiling = false;
  bool verbose = false;
  int n_queues = -1;
  int n_repetitions = 10;
  double min_bandwidth = -1.0;             // GB/s; <0 = no floor
  std::map<std::string, long> cli_params;  // -1 = autotune / default
  std::vector<std::vector<std::string>> groups;
  std::string json_path;                   // optional: append one JSON row per group
};

enum class Verdict { kSuccess, kFarFromTheoretical, kBandwidthFloor };

struct GroupReport {
  std::vector<std::string> commands;
  long serial_total_us = 0;
  std::vector<long> serial_command_us;
  long concurrent_total_us = 0;
  double max_speedup = 0;
  double speedup = 0;
  unsigned long long bytes = 0;
  double concurrent_gbytes_per_s = 0;
  double device_total_us = -1;
  Verdict verdict = Verdict::kSuccess;
};

std::string strip_twos(const std::string& token);                 // "M2D" -> "MD"
bool is_compute_command(const std::string& cmd);                  // "C", "A", "T"
std::string tuned_parameter_of(const std::string& cmd);           // C -> tripcount_C, else globalsize_<cmd>
std::string usage_text(const std::string& program, const Backend& backend);

// Throws UsageError (message may be empty = "just print help").
Options parse_arguments(const std::vector<std::string>& argv_tail, const Backend& backend,
                        const std::string& program = "concurency");

// Fills every parameter the groups need; -1 entries get their default value
// (globalsize_C=1, tripcount_C=40000, copies = globalsize_default_memory or 1e9/sizeof(float)).
Params resolve_parameters(const Options& opt);

// If any tunable parameter was left at -1 and there is more than one distinct
// command: measure each command alone at its default size and rescale the
// tunables linearly so that every command lasts as long as the fastest copy.
// Returns true if tuning ran.
bool autotune(const Options& opt, Backend& backend, Params& params, std::ostream& out);

unsigned long long bytes_moved(const std::vector<std::string>& commands, const Params& params,
                               size_t elem_size);
// "<t>us" or "<t>us (<bw> GBytes/s)" (bw = 1e-3 * bytes / us).
std::string time_with_bandwidth(long us, unsigned long long bytes);

Verdict judge(double max_speedup, double speedup, double gbytes_per_s, double min_bandwidth,
              unsigned long long bytes);
std::string verdict_text(Verdict v);
std::string verdict_line(const std::string& mode, const std::vector<std::string>& commands,
                         Verdict v);

GroupReport run_group(const Options& opt, Backend& backend, const Params& params,
                      const std::vector<std::string>& commands, std::ostream& out,
                      std::ostream& err);

// Whole program minus backend selection.  Returns the process exit status.
int run(const std::vector<std::string>& argv_tail, Backend& backend, std::ostream& out,
        std::ostream& err, const std::string& program = "concurency");

// Deterministic backend for tests: time(cmd) = coef[cmd] * tuned_parameter,
// concurrent = max + (1 - overlap) * (sum - max).  Spec: "C=0.01,MD=0.0005,overlap=0.9".
std::unique_ptr<Backend> make_fake_backend(const std::string& spec);

}  // namespace con
}  // namespace hpcp
