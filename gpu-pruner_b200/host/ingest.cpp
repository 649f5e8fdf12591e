#include "ingest.hpp"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <limits>
#include <stdexcept>
#include <unordered_map>

namespace gph {

namespace {

const Json& result_array(const Json& resp) {
  if (resp.is_array()) return resp;
  if (resp.is_object()) {
    const Json& st = resp["status"];
    if (st.is_string() && st.as_string() != "success")
      throw std::runtime_error("prometheus response status: " + st.as_string() + " " +
                               resp["error"].as_string());
    const Json& data = resp["data"];
    const Json& rt = data["resultType"];
    if (rt.is_string() && rt.as_string() != "matrix")
      throw std::runtime_error("expected matrix response from prometheus, got " + rt.as_string());
    if (data["result"].is_array()) return data["result"];
  }
  throw std::runtime_error("not a Prometheus matrix response");
}

// exported_<x> first, then <x> (lib.rs:158-175)
const std::string* label(const Json& metric, const char* exported, const char* bare) {
  const Json* j = metric.find(exported);
  if (j && j->is_string()) return &j->as_string();
  j = metric.find(bare);
  if (j && j->is_string()) return &j->as_string();
  return nullptr;
}

double sample_value(const Json& v) {
  if (v.is_string()) return strtod(v.as_string().c_str(), nullptr);  // "NaN", "+Inf", "0.37"
  return v.as_number(std::numeric_limits<double>::quiet_NaN());
}

struct RawSeries {
  uint32_t pod, slot;
  const Json* values;
  bool prof;
};

float to_f32(double x, IngestStats& st) {
  float f = (float)x;
  if (x != 0.0 && f == 0.0f && !std::isnan(x)) {  // below the f32 denormal range: keep it non-zero
    f = std::copysign(std::numeric_limits<float>::denorm_min(), (float)(x < 0 ? -1.0 : 1.0));
    ++st.tiny_values_clamped;
  }
  return f;
}

}  // namespace

Window ingest_matrix(const Json& util, const Json* prof, const Json* power, const IngestOptions& opt) {
  Window w;
  std::map<std::pair<std::string, std::string>, uint32_t> pod_index;
  std::vector<std::map<std::string, uint32_t>> slot_index;       // per pod: group key -> util slot
  std::vector<std::map<std::string, uint32_t>> pslot_index;      // per pod: group key -> power slot
  std::map<std::pair<uint32_t, uint32_t>, std::vector<std::string>> prof_sigs;  // (pod, slot) -> PROF label sets
  std::vector<RawSeries> useries, pseries;
  int64_t newest = std::numeric_limits<int64_t>::min();
  int64_t min_step = std::numeric_limits<int64_t>::max();

  auto scan = [&](const Json& resp, bool is_power, bool is_prof) {
    for (const Json& s : result_array(resp).items()) {
      ++w.stats.series_in;
      const Json& m = s["metric"];
      const std::string* pod = label(m, "exported_pod", "pod");
      const std::string* ns = label(m, "exported_namespace", "namespace");
      const std::string* ctr = label(m, "exported_container", "container");
      const Json* model = m.find("modelName");
      // the selector demands pod != "" (query.promql.j2:11,17,40); a series that cannot be turned
      // into PodMetricData is skipped with a log line (main.rs:423-428)
      if (!pod || pod->empty() || !ns || (!is_power && (!ctr || !model || !model->is_string()))) {
        ++w.stats.series_skipped;
        continue;
      }
      const Json& vals = s["values"];
      if (!vals.is_array() || vals.size() == 0) continue;  // no sample in range: no element
      auto key = std::make_pair(*pod, *ns);
      auto it = pod_index.find(key);
      uint32_t p;
      if (it == pod_index.end()) {
        p = (uint32_t)w.pods.size();
        pod_index[key] = p;
        w.pods.push_back(PodEntry{*pod, *ns, {}, 0});
        slot_index.emplace_back();
        pslot_index.emplace_back();
      } else {
        p = it->second;
      }
      const std::string host = m["Hostname"].as_string(), gpu = m["gpu"].as_string();
      const std::string mdl = model && model->is_string() ? model->as_string() : std::string();
      // `sum by (Hostname, container, pod, namespace, gpu, modelName)` groups (query.promql.j2:9)
      const std::string gkey = host + "\x1f" + (ctr ? *ctr : std::string()) + "\x1f" + gpu + "\x1f" + mdl;
      uint32_t slot;
      if (is_power) {
        auto& idx = pslot_index[p];
        auto f = idx.find(gkey);
        if (f == idx.end()) slot = idx[gkey] = w.pods[p].power_slots++;
        else slot = f->second, ++w.stats.duplicates_merged;
        pseries.push_back(RawSeries{p, slot, &vals, false});
      } else {
        auto& idx = slot_index[p];
        auto f = idx.find(gkey);
        if (f == idx.end()) {
          slot = idx[gkey] = (uint32_t)w.pods[p].slots.size();
          GpuSlot g;
          g.hostname = host, g.container = *ctr, g.gpu = gpu, g.model = mdl;
          const Json* nt = m.find("node_type");
          g.node_type = nt && nt->is_string() ? nt->as_string() : "unknown";  // lib.rs:176-179
          g.from_prof = is_prof;
          w.pods[p].slots.push_back(g);
        } else {
          slot = f->second;
          ++w.stats.duplicates_merged;
        }
        // `A or B` (query.promql.j2:10-20) matches on the FULL label set: a UTIL element is dropped
        // only if a PROF element with identical labels exists; series that differ in any other label
        // both survive the `or` and are then folded together by `sum by`
        std::vector<std::string> parts;
        for (const Json::Member& kv : m.members())
          if (kv.first != "__name__") parts.push_back(kv.first + "\x1f" + kv.second.as_string());
        std::sort(parts.begin(), parts.end());
        std::string sig;
        for (const std::string& x : parts) sig += x + "\x1e";
        std::vector<std::string>& ps = prof_sigs[std::make_pair(p, slot)];
        if (is_prof) {
          ps.push_back(sig);
          w.pods[p].slots[slot].from_prof = true;
        } else {
          bool shadowed = false;
          for (const std::string& x : ps) shadowed |= (x == sig);
          if (shadowed) {
            --w.stats.duplicates_merged;
            continue;
          }
        }
        useries.push_back(RawSeries{p, slot, &vals, is_prof});
      }
      int64_t prev = std::numeric_limits<int64_t>::min();
      for (const Json& tv : vals.items()) {
        const int64_t ts = (int64_t)std::llround(tv[0].as_number());
        newest = std::max(newest, ts);
        if (prev != std::numeric_limits<int64_t>::min() && ts > prev) min_step = std::min(min_step, ts - prev);
        prev = ts;
      }
    }
  };
  if (prof) scan(*prof, false, true);   // PROF first so it wins the `or`
  scan(util, false, false);
  if (power) scan(*power, true, false);

  w.P = (uint32_t)w.pods.size();
  uint32_t G = 1;
  for (const PodEntry& pe : w.pods) G = std::max<uint32_t>(G, std::max<uint32_t>((uint32_t)pe.slots.size(), pe.power_slots));
  w.G = G;
  w.step = opt.step > 0 ? opt.step : (min_step == std::numeric_limits<int64_t>::max() ? 1 : min_step);
  w.t_end = opt.t_end > 0 ? opt.t_end : (newest == std::numeric_limits<int64_t>::min() ? 0 : newest);
  const int64_t span = opt.duration_min * 60;
  w.T = (uint32_t)std::max<int64_t>(1, span / w.step);   // (t_end - N, t_end] sampled every `step`
  const size_t cells = (size_t)w.P * w.G * w.T;
  const float nan = std::numeric_limits<float>::quiet_NaN();
  w.util.assign(cells, nan);
  if (power) w.power.assign(cells, nan);

  auto place = [&](const std::vector<RawSeries>& list, std::vector<float>& plane) {
    for (const RawSeries& rs : list) {
      float* row = plane.data() + ((size_t)rs.pod * w.G + rs.slot) * w.T;
      for (const Json& tv : rs.values->items()) {
        ++w.stats.samples_in;
        const int64_t ts = (int64_t)std::llround(tv[0].as_number());
        const int64_t back = (w.t_end - ts + w.step / 2) / w.step;   // 0 = newest column
        if (ts > w.t_end || back < 0 || back >= (int64_t)w.T) {
          ++w.stats.samples_out_of_window;
          continue;
        }
        const float v = to_f32(sample_value(tv[1]), w.stats);
        float& cell = row[w.T - 1 - (size_t)back];
        // duplicates of one group: per-step max.  For the non-negative DCGM metrics the window max
        // of that is 0 exactly when every member's max is 0, i.e. when `sum by` of the maxima is 0.
        cell = std::isnan(cell) ? v : (std::isnan(v) ? cell : std::max(cell, v));
      }
    }
  };
  place(useries, w.util);
  if (power) place(pseries, w.power);
  return w;
}

}  // namespace gph
