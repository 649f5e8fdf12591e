// gpr_engine.cpp — the product VerdictEngine: libgpr.so (include/gpr.h), CUDA on sm_100a.
// There is no other implementation in the product; without a CUDA device every tick fails and is
// counted like a failed Prometheus query in the reference (main.rs:310-321).
#include <cstring>

#include "../../include/gpr.h"
#include "controller.hpp"

namespace gph {
namespace {

class GprVerdictEngine : public VerdictEngine {
 public:
  ~GprVerdictEngine() override {
    if (ctx_) gpr_destroy(ctx_);
  }

  bool decide(const VerdictRequest& rq, Verdict* out, std::string* error) override {
    const Window& w = *rq.window;
    if (!ensure(w, rq.power_on, rq.gpu_device, error)) return false;
    const uint32_t W = (w.P + 31) / 32;
    out->decision_bits.assign(W, 0), out->candidate_bits.assign(W, 0);
    out->series_max.assign((size_t)w.P * w.G, 0.f);
    gpr_window win;
    memset(&win, 0, sizeof win);
    win.struct_size = sizeof win;
    win.mem_kind = GPR_MEM_HOST;
    win.util = w.util.data();
    win.power = rq.power_on ? w.power.data() : nullptr;
    win.power_threshold = rq.power_on ? rq.power_threshold : 0.0;
    win.eligible = rq.eligible;
    win.created_ts = rq.created_ts;
    win.cutoff_ts = rq.cutoff_ts;
    win.n_pods = w.P, win.n_gpus = w.G, win.n_samples = w.T;
    gpr_result res;
    memset(&res, 0, sizeof res);
    res.struct_size = sizeof res;
    res.out_mem_kind = GPR_MEM_HOST;
    res.decision_bits = out->decision_bits.data();
    res.candidate_bits = out->candidate_bits.data();
    res.series_max = out->series_max.data();
    const int rc = gpr_decide(ctx_, &win, &res);
    if (rc != GPR_OK) {
      *error = "idle engine (" + std::to_string(rc) + "): " + gpr_last_error(ctx_);
      return false;
    }
    out->n_series = res.n_series, out->n_candidates = res.n_candidates, out->n_decisions = res.n_decisions;
    out->kernel_ms = res.kernel_ms;
    return true;
  }

 private:
  bool ensure(const Window& w, bool need_power, int device, std::string* error) {
    const uint64_t cells = (uint64_t)w.P * w.G * w.T;
    if (ctx_ && cells <= cap_cells_ && (!need_power || cap_power_)) return true;
    if (ctx_) gpr_destroy(ctx_), ctx_ = nullptr;
    gpr_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.struct_size = sizeof cfg;
    cfg.device = device;
    // head-room so that a growing cluster does not re-create the context every tick
    cfg.max_pods = w.P + w.P / 4 > 64 ? w.P + w.P / 4 : 64;
    cfg.max_gpus = w.G ? w.G : 1;
    cfg.max_samples = w.T ? w.T : 1;
    cfg.flags = need_power ? GPR_F_POWER_PLANE : 0;
    const int rc = gpr_create(&cfg, &ctx_);
    if (rc != GPR_OK) {
      *error = std::string("idle engine unavailable (") + std::to_string(rc) + "): " + gpr_last_error(nullptr);
      ctx_ = nullptr;
      return false;
    }
    cap_cells_ = (uint64_t)cfg.max_pods * cfg.max_gpus * cfg.max_samples;
    cap_power_ = need_power;
    return true;
  }
  gpr_ctx* ctx_ = nullptr;
  uint64_t cap_cells_ = 0;
  bool cap_power_ = false;
};

}  // namespace

std::unique_ptr<VerdictEngine> make_gpr_engine() { return std::make_unique<GprVerdictEngine>(); }

}  // namespace gph
