// ingest_device.cpp — host half of the device-side ingest (see ingest_device.hpp and
// gpu-pruner_b200/csrc/gpr_text.cuh).  The wire shape is the matrix result that
// /root/reference/gpu-pruner/src/bin/querytest.rs:41-53 walks; label precedence and defaults follow
// PodMetricData::try_from (gpu-pruner/src/lib.rs:153-187) through the Assigner shared with ingest.cpp.
// The device reports where sample lists open and close; this file walks the series with those offsets,
// turns every label map into a tensor row, lets the device parse the samples, and re-parses on the CPU
// the few rows the strict device parser declined.
#include "ingest_device.hpp"

#include <algorithm>
#include <chrono>
#include <cstring>
#include <memory>

#include "ingest_internal.hpp"

namespace gph {

using namespace detail;

namespace {

struct NotCompact {  // the response is not in the one encoding the device scan understands
  std::string why;
};

struct DevSeries {
  uint32_t pod, slot;
  uint64_t begin, end;  // gpr_text_span.begin / .end
};

struct TextPlan {
  int slot = 0;                    // resident text slot on the device
  const std::string* text = nullptr;
  std::vector<DevSeries> series;   // placed series, in text order
};

const char kHead[] = "{\"status\":\"success\",\"data\":{\"resultType\":\"matrix\",\"result\":[";
const char kMetric[] = "{\"metric\":";

bool starts_with(const std::string& s, size_t at, const char* lit) {
  const size_t n = strlen(lit);
  return at + n <= s.size() && memcmp(s.data() + at, lit, n) == 0;
}

size_t skip_trailing_ws(const std::string& s, size_t at) {
  while (at < s.size() && (s[at] == ' ' || s[at] == '\n' || s[at] == '\r' || s[at] == '\t')) ++at;
  return at;
}

double ms_since(std::chrono::steady_clock::time_point t0) {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

struct SeriesLoc {
  size_t mb, close_brace, vb, list_close;  // label map [mb, close_brace], list [vb, list_close]
};

// Walk the series of one response with the device's marker lists, as they arrive: the upload + scan runs as a
// pipeline (TextDevice::scan_begin / scan_next) and every series whose markers are in is validated, its label map
// parsed and its row assigned through `asg` while later chunks of the text are still on their way to the GPU.
void plan_text(TextDevice& dev, TextPlan& plan, Assigner& asg, Window& w, bool is_power, bool is_prof,
               DeviceIngestReport& rep, bool remember) {
  const std::string& t = *plan.text;
  size_t at;
  bool bare;
  if (starts_with(t, 0, kHead)) at = sizeof kHead - 1, bare = false;
  else if (starts_with(t, 0, "[")) at = 1, bare = true;
  else throw NotCompact{"response does not start with the compact success/matrix header"};

  std::vector<uint64_t> opens, closes, co, cc;  // all markers so far (sorted: chunks come in text order)
  size_t oi = 0, ci = 0, walked = 0;
  bool finished = false;  // the closing ']' of the result array has been reached
  if (at < t.size() && t[at] == ']') ++at, finished = true;  // empty result
  FlatLabels flat;

  // consume every series that is completely covered by the markers delivered so far
  auto drain = [&](bool final) {
    while (!finished) {
      if (!starts_with(t, at, kMetric) || at + sizeof kMetric - 1 >= t.size() || t[at + sizeof kMetric - 1] != '{')
        throw NotCompact{"series does not start with {\"metric\":{"};
      const size_t mb = at + sizeof kMetric - 1;  // '{' of the label map
      while (oi < opens.size() && opens[oi] < mb) ++oi;
      if (oi == opens.size()) {
        if (final) throw NotCompact{"label map without a following \"values\" list"};
        return;
      }
      const size_t close_brace = (size_t)opens[oi];  // '}' of the label map
      const size_t vb = close_brace + 12;            // first byte after `},"values":[`
      if (vb >= t.size()) throw NotCompact{"truncated values list"};
      size_t list_close;  // offset of the ']' closing the list
      if (t[vb] == ']') {
        list_close = vb;
      } else {
        if (t[vb] != '[') throw NotCompact{"values list does not start with a sample"};
        while (ci < closes.size() && closes[ci] < vb) ++ci;
        if (ci == closes.size()) {
          if (final) throw NotCompact{"unterminated values list"};
          return;
        }
        list_close = (size_t)closes[ci] + 2;
      }
      if (list_close + 1 >= t.size() || t[list_close + 1] != '}')
        throw NotCompact{"series object has members after \"values\""};
      ++oi, ++walked;
      // The label map is parsed for EVERY series (also the ones whose list is empty): a complete JSON object
      // ending exactly at the marker's '}' proves that [mb, close_brace] is the whole map and that no series
      // without a "values" member was jumped over.  (A `},"values":[` inside a label value is impossible: a raw
      // '"' ends a JSON string.)  Prometheus' own shape — string values, no escapes — is read in place without
      // allocating (FlatLabels); anything else goes through the DOM parser.
      const char* b = t.data() + mb;
      const char* e = t.data() + close_brace + 1;
      const bool element = list_close != vb;  // an empty list is no element (as in the CPU paths)
      uint32_t p = 0, slot = 0;
      Assigner::Result r = Assigner::Skipped;
      const std::string_view raw(b, (size_t)(e - b));
      if (remember && element && asg.lookup_known(raw, is_power, is_prof, &r, &p, &slot)) {
        ++w.stats.series_in;  // a series of an earlier tick: same bytes, same row, nothing to parse
      } else if (flat.parse(b, e)) {
        ++w.stats.series_in;
        if (element) {
          r = asg.assign(flat, is_power, is_prof, &p, &slot);
          if (remember) asg.remember(r, p, slot);
        }
      } else {
        Json metric;
        try {
          metric = Json::parse(std::string(b, e));
        } catch (const std::exception& ex) {
          throw NotCompact{std::string("label map: ") + ex.what()};
        }
        if (!metric.is_object()) throw NotCompact{"label map is not an object"};
        ++w.stats.series_in;
        if (element) {
          r = asg.assign(metric, is_power, is_prof, &p, &slot);
          if (remember) asg.remember(r, p, slot);
        }
      }
      if (element && r == Assigner::Placed)
        plan.series.push_back(DevSeries{p, slot, (uint64_t)vb, (uint64_t)list_close});
      at = list_close + 2;  // past '}'
      if (at < t.size() && t[at] == ',') {
        ++at;
        continue;
      }
      if (at < t.size() && t[at] == ']') {
        ++at;
        finished = true;
        break;
      }
      throw NotCompact{"expected ',' or ']' after a series"};
    }
  };

  auto t0 = std::chrono::steady_clock::now();
  dev.scan_begin(plan.slot, t.data(), t.size());
  bool more = true;
  try {
    while (more) {
      uint64_t ready = 0;
      more = dev.scan_next(&co, &cc, &ready);
      rep.scan_ms += ms_since(t0);  // time spent waiting for the upload / scan
      t0 = std::chrono::steady_clock::now();
      opens.insert(opens.end(), co.begin(), co.end());
      closes.insert(closes.end(), cc.begin(), cc.end());
      drain(!more);
      rep.assign_ms += ms_since(t0);  // series walk + label maps -> rows, overlapped with the upload
      t0 = std::chrono::steady_clock::now();
    }
  } catch (...) {
    while (more) {  // let the pipeline run to its end: the device slot must not be left half written
      uint64_t ready = 0;
      try {
        more = dev.scan_next(&co, &cc, &ready);
      } catch (...) {
        break;
      }
    }
    throw;
  }
  if (walked != opens.size()) throw NotCompact{"values markers outside the series walk"};
  if (!bare) {
    if (!starts_with(t, at, "}}")) throw NotCompact{"response has members after \"result\""};
    at += 2;
  }
  if (skip_trailing_ws(t, at) != t.size()) throw NotCompact{"trailing bytes after the response"};
}

}  // namespace

struct DeviceIngestSession::State {
  Window w;                        // skeleton: pods / slots / shape, no planes
  std::unique_ptr<Assigner> asg;   // bound to `w`; lives as long as the rows do
  bool valid = false;              // the resident ring holds the window ending at w.t_end
  uint32_t pods_cap = 0;           // pods the ring has rows for
  bool with_power = false;
  std::vector<std::pair<uint32_t, uint32_t>> prof_rows;  // (pod, slot) fed by PROF series, sorted
};

DeviceIngestSession::DeviceIngestSession(TextDevice& dev) : dev_(dev), st_(new State()) {}
DeviceIngestSession::~DeviceIngestSession() { delete st_; }
int64_t DeviceIngestSession::resident_t_end() const { return st_->valid ? st_->w.t_end : 0; }
void DeviceIngestSession::invalidate() { st_->valid = false; }

Window DeviceIngestSession::ingest(const std::string& util, const std::string* prof, const std::string* power,
                                   const IngestOptions& opt, DeviceIngestReport* report) {
  DeviceIngestReport local;
  DeviceIngestReport& rep = report ? *report : local;
  rep = DeviceIngestReport{};
  State& st = *st_;
  const bool delta = opt.slice_seconds > 0;
  auto cpu = [&](const std::string& why) {
    st.valid = false;
    if (delta) throw NeedFullWindow("device ingest not possible for the tick's slice: " + why);
    rep.on_device = false, rep.reason = why;
    Window w = ingest_matrix_text(util, prof, power, opt);
    w.stats.warnings.push_back("device ingest not used: " + why);
    return w;
  };
  if (opt.t_end <= 0 || opt.step <= 0) return cpu("window end / step not given (query.json)");

  if (delta) {
    // what must hold for the ring to take this tick: same grid, contiguous with what is resident
    const char* why = nullptr;
    if (!st.valid) why = "nothing resident";
    else if (opt.step != st.w.step || opt.duration_min * 60 != st.w.span) why = "step / window length changed";
    else if (opt.t_end - opt.slice_seconds != st.w.t_end) why = "the slice does not start where the resident window ends";
    else if (opt.slice_seconds % opt.step != 0) why = "the slice is not a whole number of steps";
    else if (opt.slice_seconds / opt.step >= (int64_t)st.w.T) why = "the slice is as long as the window";
    else if ((power != nullptr) != st.with_power) why = "power plane appeared / disappeared";
    if (why) {
      st.valid = false;
      throw NeedFullWindow(why);
    }
  } else {
    st.w = Window();
    st.asg.reset(new Assigner(st.w));
    st.valid = false;
    st.prof_rows.clear();
  }
  Window& w = st.w;
  w.stats = IngestStats();
  Assigner& asg = *st.asg;

  TextPlan plans[3];  // prof, util, power — the order the CPU paths assign rows in
  int n_plans = 0;
  auto add = [&](const std::string* text, int slot) -> TextPlan* {
    if (!text) return nullptr;
    plans[n_plans].slot = slot, plans[n_plans].text = text;
    return &plans[n_plans++];
  };
  TextPlan* pl_prof = add(prof, 0);
  TextPlan* pl_util = add(&util, 1);
  TextPlan* pl_power = add(power, 2);
  try {
    const bool remember = delta || opt.resident;
    if (pl_prof) plan_text(dev_, *pl_prof, asg, w, false, true, rep, remember);
    plan_text(dev_, *pl_util, asg, w, false, false, rep, remember);
    if (pl_power) plan_text(dev_, *pl_power, asg, w, true, false, rep, remember);
  } catch (const NotCompact& e) {
    return cpu(e.why);
  }
  std::vector<std::pair<uint32_t, uint32_t>> prof_rows;
  if (pl_prof)
    for (const DevSeries& s : pl_prof->series) prof_rows.emplace_back(s.pod, s.slot);
  std::sort(prof_rows.begin(), prof_rows.end());

  uint32_t n_new = 0;  // buckets this call opens (delta) — 0: the whole window
  if (!delta) {
    finish_shape(w, opt, 0, 1, power != nullptr, /*allocate=*/false);  // t_end / step given: nothing to infer
    st.with_power = power != nullptr;
    st.prof_rows = prof_rows;
    if (opt.resident) {
      st.pods_cap = w.P + w.P / 4 + 64;   // head-room: new pods get rows without a rebuild
      dev_.resident_init(st.pods_cap, w.G, w.T, st.with_power);
    }
  } else {
    // the shape is the ring's: anything that does not fit needs the full window again
    uint32_t g_now = 1;
    for (const PodEntry& pe : w.pods) g_now = std::max<uint32_t>(g_now, std::max<uint32_t>((uint32_t)pe.slots.size(), pe.power_slots));
    const char* why = nullptr;
    if (w.pods.size() > st.pods_cap) why = "more pods than the resident window has rows for";
    else if (g_now > w.G) why = "a pod gained a GPU slot beyond the resident window's shape";
    // `A or B` (query.promql.j2:10-20) is resolved per tick at assignment time; if the set of PROF-fed rows
    // changes, UTIL samples that were (not) shadowed earlier in the window no longer match a fresh query
    else if (prof_rows != st.prof_rows) why = "the set of DCGM_FI_PROF_GR_ENGINE_ACTIVE series changed";
    if (why) {
      st.valid = false;
      throw NeedFullWindow(why);
    }
    w.P = (uint32_t)w.pods.size();
    w.t_end = opt.t_end;
    n_new = (uint32_t)(opt.slice_seconds / opt.step);
    dev_.resident_advance(n_new);
  }
  const bool resident = delta || opt.resident;
  const uint32_t n_rows = (resident ? st.pods_cap : w.P) * w.G;
  auto result = [&]() {
    Window out = w;  // pods / shape / stats
    out.resident = resident;
    if (resident) out.resident_pods = st.pods_cap, out.resident_power = st.with_power, st.valid = true;
    return out;
  };
  if (w.P == 0) {
    rep.on_device = true;
    return result();
  }

  // window the parse accepts: the whole range, or only the tick's slice
  const int64_t parse_span = delta ? opt.slice_seconds : w.span;
  const uint32_t patch_cols = delta ? n_new : w.T;
  auto run_plane = [&](std::vector<TextPlan*> texts, int plane) {
    std::vector<uint32_t> writers(n_rows, 0);
    for (TextPlan* tp : texts)
      for (const DevSeries& s : tp->series) ++writers[(size_t)s.pod * w.G + s.slot];
    std::vector<std::vector<gpr_text_span>> spans(texts.size());
    bool fill = !resident;
    for (size_t k = 0; k < texts.size(); ++k) {
      for (const DevSeries& s : texts[k]->series) {
        gpr_text_span sp;
        memset(&sp, 0, sizeof sp);
        sp.begin = s.begin, sp.end = s.end, sp.row = s.pod * w.G + s.slot;
        sp.flags = writers[sp.row] > 1 ? GPR_SPAN_SHARED : 0u;
        spans[k].push_back(sp);
      }
      const auto tp = std::chrono::steady_clock::now();
      TextDevice::TextGrid grid;
      grid.t_end = w.t_end, grid.span = parse_span, grid.step = w.step, grid.T = w.T, grid.n_rows = n_rows;
      grid.fill = fill, grid.resident = resident;
      dev_.parse(texts[k]->slot, spans[k], grid, plane);
      rep.parse_ms += ms_since(tp);
      fill = false;
      rep.spans += spans[k].size();
    }
    // rows the device gave up on: re-parse every series feeding them with the CPU walker
    std::vector<uint8_t> dirty(n_rows, 0);
    bool any_dirty = false;
    for (const auto& list : spans)
      for (const gpr_text_span& sp : list)
        if (sp.flags & GPR_SPAN_HARD) dirty[sp.row] = 1, any_dirty = true, ++rep.hard_spans;
    std::vector<std::vector<float>> rows;
    std::vector<uint32_t> row_ids;
    std::vector<int64_t> row_slot(any_dirty ? n_rows : 0, -1);
    Window bucket;  // the grid of the patched columns: the newest `patch_cols` buckets
    bucket.t_end = w.t_end, bucket.step = w.step, bucket.span = parse_span, bucket.T = patch_cols;
    for (size_t k = 0; k < texts.size(); ++k) {
      const std::string& t = *texts[k]->text;
      for (const gpr_text_span& sp : spans[k]) {
        if (!dirty[sp.row]) {
          w.stats.samples_in += sp.n_in;
          w.stats.samples_out_of_window += sp.n_oow;
          w.stats.tiny_values_clamped += sp.n_tiny;
          continue;
        }
        if (row_slot[sp.row] < 0) {
          row_slot[sp.row] = (int64_t)rows.size();
          rows.emplace_back(patch_cols, std::numeric_limits<float>::quiet_NaN());
          row_ids.push_back(sp.row);
        }
        float* row = rows[(size_t)row_slot[sp.row]].data();
        for_each_sample(t.data() + sp.begin - 1, t.data() + sp.end + 1, [&](double ts, double v) {
          ++w.stats.samples_in;
          const int64_t col = column_of(bucket, ts_millis(ts));
          if (col < 0) {
            ++w.stats.samples_out_of_window;
            return;
          }
          merge_cell(row[col], to_f32(v, &w.stats.tiny_values_clamped));
        });
      }
    }
    for (size_t i = 0; i < rows.size(); ++i) dev_.patch_row(plane, row_ids[i], w.T, rows[i].data(), patch_cols, resident);
    rep.rows_patched += rows.size();
  };
  std::vector<TextPlan*> util_texts;
  if (pl_prof) util_texts.push_back(pl_prof);
  util_texts.push_back(pl_util);
  run_plane(util_texts, 0);
  if (pl_power) run_plane({pl_power}, 1);
  rep.on_device = true;
  Window out = result();
  if (!resident) {
    out.d_util = dev_.plane(0);
    if (pl_power) out.d_power = dev_.plane(1);
  }
  return out;
}

Window ingest_matrix_device(TextDevice& dev, const std::string& util, const std::string* prof,
                            const std::string* power, const IngestOptions& opt, DeviceIngestReport* report) {
  DeviceIngestSession once(dev);
  IngestOptions o = opt;
  o.slice_seconds = 0, o.resident = false;
  return once.ingest(util, prof, power, o, report);
}

}  // namespace gph
