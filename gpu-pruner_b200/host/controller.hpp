// controller.hpp — the tick: window -> GPU decision -> gates -> owner walk -> scale requests.
//
// C++ mirror of run_query_and_scale and the two tasks around it
// (/root/reference/gpu-pruner/src/main.rs:286-367, 390-570).  The aggregation that the reference
// delegates to Prometheus is done by libgpr (include/gpr.h); everything after it follows the
// reference step by step, with the Kubernetes API behind the KubeApi interface.
#pragma once
#include <cstdint>
#include <functional>
#include <memory>
#include <string>
#include <unordered_set>
#include <vector>

#include "../../include/gpr.h"
#include "cli.hpp"
#include "ingest.hpp"
#include "kube.hpp"

namespace gph {

// lib.rs:136-145
struct PodMetricData {
  std::string name, ns, container, node_type, gpu_model;
  double value = 0;
};
// lib.rs:131-134
struct QueryResponse {
  size_t num_pods = 0;         // series returned by the query, pre-dedup (main.rs:418)
  size_t shutdown_events = 0;  // distinct root objects (main.rs:536)
};

class Logger {
 public:
  Logger(LogFormat f, FILE* out) : fmt_(f), out_(out) {}
  void log(const char* level, const std::string& msg,
           const std::vector<std::pair<std::string, std::string>>& fields = {}) const;
  void info(const std::string& m) const { log("INFO", m); }
  void warn(const std::string& m) const { log("WARN", m); }
  void error(const std::string& m) const { log("ERROR", m); }
  void counter(const char* level, const std::string& name, uint64_t v, const std::string& msg) const;

 private:
  LogFormat fmt_;
  FILE* out_;
};

struct TickResult {
  bool ok = false;
  std::string error;
  QueryResponse qr;
  std::vector<PodMetricData> unique_pods;     // after the ANY-GPU dedup (main.rs:416-437)
  std::vector<ScaleKind> shutdown;            // distinct roots, insertion order
  std::vector<Request> requests;              // what scale-down mode would send
  double kernel_ms = 0;
  uint64_t n_candidates = 0, n_decisions = 0;
};

// Where the window comes from, selected by the scheme of --prometheus-url.
class WindowSource {
 public:
  virtual ~WindowSource() = default;
  virtual Window fetch(const Cli& args) = 0;   // throws std::runtime_error on failure
};
std::unique_ptr<WindowSource> make_window_source(const std::string& url);

class Controller {
 public:
  Controller(const Cli& args, KubeApi* kube, Logger log, Clock clock);
  ~Controller();
  Controller(const Controller&) = delete;

  // one pass of main.rs:390-570 on an already-fetched window
  TickResult run_query_and_scale(const Window& w);
  // main.rs:286-330: one-shot or daemon loop with the consecutive-failure budget; returns exit code
  int run(WindowSource& src);

  const std::string& engine_error() const { return engine_error_; }
  uint64_t query_successes = 0, query_failures = 0, scale_successes = 0, scale_failures = 0;

 private:
  bool ensure_engine(const Window& w);
  Cli args_;
  KubeApi* kube_;
  Logger log_;
  Clock clock_;
  gpr_ctx* ctx_ = nullptr;
  uint64_t cap_cells_ = 0;
  bool cap_power_ = false;
  std::string engine_error_;
  uint8_t enabled_;
};

}  // namespace gph
