#include "ingest.hpp"

#include <thread>

#include "ingest_internal.hpp"

namespace gph {

using namespace detail;

namespace {

// =====================================================================================================
// DOM path
// =====================================================================================================
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

double sample_value(const Json& v) {
  if (v.is_string()) return strtod(v.as_string().c_str(), nullptr);  // "NaN", "+Inf", "0.37"
  return v.as_number(std::numeric_limits<double>::quiet_NaN());
}

struct RawSeries {
  uint32_t pod, slot;
  const Json* values;
};

}  // namespace


Window ingest_matrix(const Json& util, const Json* prof, const Json* power, const IngestOptions& opt) {
  Window w;
  Assigner asg(w);
  std::vector<RawSeries> useries, pseries;
  int64_t newest = kNoTs;
  int64_t min_step = std::numeric_limits<int64_t>::max();

  auto scan = [&](const Json& resp, bool is_power, bool is_prof) {
    for (const Json& s : result_array(resp).items()) {
      ++w.stats.series_in;
      const Json& vals = s["values"];
      uint32_t p, slot;
      if (!vals.is_array() || vals.size() == 0) {  // no sample in range: no element
        // still counts as skipped if it could not have been converted
        continue;
      }
      if (asg.assign(s["metric"], is_power, is_prof, &p, &slot) != Assigner::Placed) continue;
      (is_power ? pseries : useries).push_back(RawSeries{p, slot, &vals});
      int64_t prev = kNoTs;
      for (const Json& tv : vals.items()) {
        // a sample is exactly [ <unix time>, "<value>" ]; anything else is a malformed response
        // (both paths reject it, the tick then counts as a query failure, main.rs:310-321)
        if (!tv.is_array() || tv.size() != 2 || !tv[0].is_number())
          throw std::runtime_error("matrix response: sample is not [time, value]");
        const int64_t ts = ts_seconds(tv[0].as_number());
        if (ts == kBadTs) continue;
        newest = std::max(newest, ts);
        if (prev != kNoTs && ts > prev) min_step = std::min(min_step, ts - prev);
        prev = ts;
      }
    }
  };
  if (prof) scan(*prof, false, true);  // PROF first so it wins the `or`
  scan(util, false, false);
  if (power) scan(*power, true, false);
  finish_shape(w, opt, newest, min_step, power != nullptr);

  auto place = [&](const std::vector<RawSeries>& list, std::vector<float>& plane) {
    for (const RawSeries& rs : list) {
      float* row = plane.data() + ((size_t)rs.pod * w.G + rs.slot) * w.T;
      for (const Json& tv : rs.values->items()) {
        ++w.stats.samples_in;
        const int64_t col = column_of(w, ts_seconds(tv[0].as_number()));
        if (col < 0) {
          ++w.stats.samples_out_of_window;
          continue;
        }
        merge_cell(row[col], to_f32(sample_value(tv[1]), &w.stats.tiny_values_clamped));
      }
    }
  };
  place(useries, w.util);
  if (power) place(pseries, w.power);
  return w;
}

// =====================================================================================================
// text path: no DOM for the samples.  The response is scanned once for series boundaries (the
// `values` array of a series contains no nested brackets beyond its [ts,"v"] pairs and no escapes,
// so its end is the first "]]"), label sets are parsed with the small DOM parser, and the sample
// arrays — the bulk of the bytes — are parsed by worker threads straight into the tensor rows.

Window ingest_matrix_text(const std::string& util, const std::string* prof, const std::string* power,
                          const IngestOptions& opt, int n_threads) {
  Window w;
  Assigner asg(w);
  std::vector<TextSeries> useries, pseries;
  if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());

  FlatLabels flat;
  auto scan = [&](const std::string& text, bool is_power, bool is_prof) {
    for (const Span& s : series_spans(text)) {
      ++w.stats.series_in;
      if (!s.values_b || !s.metric_b) continue;
      const char* v = skip_ws(s.values_b + 1, s.values_e);
      if (v < s.values_e && *v == ']') continue;  // no sample in range: no element
      uint32_t p, slot;
      if (flat.parse(s.metric_b, s.metric_e)) {  // Prometheus' own shape: read in place
        if (asg.assign(flat, is_power, is_prof, &p, &slot) != Assigner::Placed) continue;
      } else {
        const Json metric = Json::parse(std::string(s.metric_b, s.metric_e));
        if (asg.assign(metric, is_power, is_prof, &p, &slot) != Assigner::Placed) continue;
      }
      (is_power ? pseries : useries).push_back(TextSeries{p, slot, s.values_b, s.values_e, true});
    }
  };
  if (prof) scan(*prof, false, true);
  scan(util, false, false);
  if (power) scan(*power, true, false);

  // pre-pass only when the caller did not say where the window ends / what the step is
  int64_t newest = kNoTs, min_step = std::numeric_limits<int64_t>::max();
  if (opt.t_end <= 0 || opt.step <= 0) {
    auto pre = [&](const std::vector<TextSeries>& list) {
      for (const TextSeries& ts : list) {
        int64_t prev = kNoTs;
        for_each_sample(ts.vb, ts.ve, [&](double t, double) {
          const int64_t ti = ts_seconds(t);
          if (ti == kBadTs) return;
          newest = std::max(newest, ti);
          if (prev != kNoTs && ti > prev) min_step = std::min(min_step, ti - prev);
          prev = ti;
        });
      }
    };
    pre(useries);
    pre(pseries);
  }
  finish_shape(w, opt, newest, min_step, power != nullptr);

  auto place = [&](std::vector<TextSeries>& list, std::vector<float>& plane) {
    // rows written by exactly one series can be filled concurrently; shared rows are merged afterwards
    std::vector<uint32_t> writers((size_t)w.P * w.G, 0);
    for (const TextSeries& ts : list) ++writers[(size_t)ts.pod * w.G + ts.slot];
    for (TextSeries& ts : list) ts.sole = writers[(size_t)ts.pod * w.G + ts.slot] == 1;
    std::vector<IngestStats> st((size_t)n_threads);
    auto work = [&](int tid, bool sole_pass) {
      // thread-local counters (adjacent IngestStats would false-share a cache line per sample)
      uint64_t n_in = 0, n_out = 0, n_tiny = 0;
      for (size_t i = (size_t)tid; i < list.size(); i += (size_t)(sole_pass ? n_threads : 1)) {
        const TextSeries& ts = list[i];
        if (ts.sole != sole_pass) continue;
        float* row = plane.data() + ((size_t)ts.pod * w.G + ts.slot) * w.T;
        for_each_sample(ts.vb, ts.ve, [&](double t, double v) {
          ++n_in;
          const int64_t col = column_of(w, ts_seconds(t));
          if (col < 0) {
            ++n_out;
            return;
          }
          merge_cell(row[col], to_f32(v, &n_tiny));
        });
      }
      IngestStats& s = st[(size_t)tid];
      s.samples_in += n_in, s.samples_out_of_window += n_out, s.tiny_values_clamped += n_tiny;
    };
    // a malformed sample array must surface as an exception on the caller's thread, not terminate()
    std::vector<std::string> errors((size_t)n_threads);
    auto guarded = [&](int tid, bool sole_pass) {
      try {
        work(tid, sole_pass);
      } catch (const std::exception& e) {
        errors[(size_t)tid] = e.what();
      }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < n_threads; ++t) th.emplace_back(guarded, t, true);
    guarded(0, true);
    for (std::thread& t : th) t.join();
    for (const std::string& e : errors)
      if (!e.empty()) throw std::runtime_error(e);
    work(0, false);  // duplicates of one group: sequential, order-independent merge
    for (const IngestStats& s : st) {
      w.stats.samples_in += s.samples_in;
      w.stats.samples_out_of_window += s.samples_out_of_window;
      w.stats.tiny_values_clamped += s.tiny_values_clamped;
    }
  };
  place(useries, w.util);
  if (power) place(pseries, w.power);
  return w;
}

}  // namespace gph
