// promql.hpp — text of the queries the controller sends.
//
// render_query() reproduces, byte for byte, what the reference renders from
// /root/reference/gpu-pruner/src/query.promql.j2 with minijinja (main.rs:280-281): the legacy
// server-side aggregation, kept for users who still want Prometheus to do the arithmetic and as
// the specification the GPU path is checked against.  The selector forms are what the engine
// needs instead: the same series selection, but as raw range vectors (matrix results) that are
// laid out as the dense tensor.
#pragma once
#include <string>

#include "cli.hpp"

namespace gph {

std::string format_float(double v);  // minijinja / Python repr style: 150.0, 120.5, 1e-07

std::string render_query(const Cli& args);

struct Selectors {
  std::string prof;   // DCGM_FI_PROF_GR_ENGINE_ACTIVE{...}[Nm]   (query.promql.j2:10-14)
  std::string util;   // DCGM_FI_DEV_GPU_UTIL{...}[Nm]            (query.promql.j2:16-20)
  std::string power;  // DCGM_FI_DEV_POWER_USAGE{...}[Nm] or ""   (query.promql.j2:39-42)
  std::string dmi;    // node_dmi_info, an instant query             (query.promql.j2:25-30)
};
// the same selectors over the last `seconds` only (daemon mode: a tick asks for what was scraped since
// the previous one; the 30-minute window itself stays resident in HBM)
Selectors render_selectors(const Cli& args, int64_t seconds);
Selectors render_selectors(const Cli& args);

struct LabelNames {  // query.promql.j2:5-7
  const char* pod;
  const char* ns;
  const char* container;
};
LabelNames label_names(bool honor_labels);

}  // namespace gph
