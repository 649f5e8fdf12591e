// ingest.hpp — Prometheus range-vector (matrix) wire format -> dense (pod x gpu x t) f32 tensor.
//
// The reference never receives raw samples: it asks Prometheus for the already-aggregated instant
// vector (/root/reference/gpu-pruner/src/main.rs:397-409).  The only reference code that walks a
// matrix result is the debug tool gpu-pruner/src/bin/querytest.rs:41-53
// (series = label map + [(timestamp f64, value)]); that is the shape ingested here.  Label
// precedence follows PodMetricData::try_from (gpu-pruner/src/lib.rs:153-187): exported_* first,
// then the bare name; node_type defaults to "unknown"; modelName is mandatory.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <string>
#include <vector>

#include "json.hpp"

namespace gph {

struct GpuSlot {            // one (Hostname, container, pod, namespace, gpu, modelName) group
  std::string hostname, container, gpu, model, node_type;
  bool from_prof = false;   // cell holds DCGM_FI_PROF_GR_ENGINE_ACTIVE (0..1) rather than GPU_UTIL (0..100)
};

struct PodEntry {
  std::string name, ns;
  std::vector<GpuSlot> slots;        // util / prof groups, in order of first appearance
  uint32_t power_slots = 0;
};

struct IngestStats {
  uint64_t series_in = 0, series_skipped = 0, samples_in = 0, samples_out_of_window = 0,
           duplicates_merged = 0, tiny_values_clamped = 0;
  std::vector<std::string> warnings;
};

struct Window {
  uint32_t P = 0, G = 0, T = 0;
  int64_t t_end = 0, step = 1;        // seconds; column c covers timestamp t_end - (T-1-c)*step
  std::vector<PodEntry> pods;
  std::vector<float> util;            // [P][G][T], NaN = no sample
  std::vector<float> power;           // empty, or [P][G][T]
  // device-resident planes (ingest_matrix_device): util / power above are empty then
  const float* d_util = nullptr;
  const float* d_power = nullptr;
  IngestStats stats;
};

struct IngestOptions {
  int64_t duration_min = 30;          // window length, --duration
  int64_t step = 0;                   // seconds; 0 = infer (smallest positive timestamp delta)
  int64_t t_end = 0;                  // 0 = newest timestamp in the response
};

// `util` is required; `prof` and `power` may be null pointers.  Each is a full Prometheus HTTP API
// response ({"status":"success","data":{"resultType":"matrix","result":[...]}}) or just the
// "result" array.  Throws std::runtime_error on malformed input / non-matrix result types.
Window ingest_matrix(const Json& util, const Json* prof, const Json* power, const IngestOptions& opt);

// Same result from the raw response TEXT, without building a DOM for the samples: series boundaries
// are found with one memmem per series, label sets go through the small DOM parser, and the sample
// arrays — >99 % of the bytes — are parsed by n_threads workers (0 = hardware concurrency) straight
// into the tensor rows.  This is the path the controller uses; ingest_matrix() is the reference
// implementation it is tested against (tests/test_host.py).
Window ingest_matrix_text(const std::string& util, const std::string* prof, const std::string* power,
                          const IngestOptions& opt, int n_threads = 0);

}  // namespace gph
