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
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>

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

// A few worker threads for the per-series label work (parse the label map, hash it, probe the table of known
// series): independent per series, and on the critical path of a tick while the text is crossing PCIe.
class Workers {
 public:
  explicit Workers(int n) : n_(std::max(1, n)) {
    for (int i = 1; i < n_; ++i) th_.emplace_back([this, i] { loop(i); });
  }
  ~Workers() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      quit_ = true;
    }
    cv_.notify_all();
    for (std::thread& t : th_) t.join();
  }
  // fn(begin, end) over [0, n) in contiguous shares; returns when all shares are done.  Exceptions are carried over.
  void run(size_t n, const std::function<void(size_t, size_t)>& fn) {
    if (n_ == 1 || n < 256) {
      fn(0, n);
      return;
    }
    {
      std::lock_guard<std::mutex> lk(mu_);
      fn_ = &fn, total_ = n, pending_ = n_ - 1, ++gen_, error_.clear();
    }
    cv_.notify_all();
    share(0);
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [this] { return pending_ == 0; });
    fn_ = nullptr;
    if (!error_.empty()) throw std::runtime_error(error_);
  }

 private:
  void share(int i) {
    const size_t b = total_ * (size_t)i / (size_t)n_, e = total_ * (size_t)(i + 1) / (size_t)n_;
    try {
      if (b < e) (*fn_)(b, e);
    } catch (const std::exception& ex) {
      std::lock_guard<std::mutex> lk(mu_);
      error_ = ex.what();
    }
  }
  void loop(int i) {
    unsigned long seen = 0;
    std::unique_lock<std::mutex> lk(mu_);
    while (true) {
      cv_.wait(lk, [&] { return quit_ || gen_ != seen; });
      if (quit_) return;
      seen = gen_;
      lk.unlock();
      share(i);
      lk.lock();
      if (--pending_ == 0) done_.notify_one();
    }
  }
  int n_;
  std::vector<std::thread> th_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  const std::function<void(size_t, size_t)>* fn_ = nullptr;
  size_t total_ = 0;
  int pending_ = 0;
  unsigned long gen_ = 0;
  bool quit_ = false;
  std::string error_;
};

Workers& label_workers() {
  static Workers w([] {
    const char* v = getenv("GPR_LABEL_THREADS");
    const int n = v && *v ? atoi(v) : (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency() / 2));
    return std::max(1, std::min(64, n));
  }());
  return w;
}

// [[ts,"v"],[ts,"v"],...] exactly as Prometheus prints it: digits with an optional fraction, a quoted number
// (strict_sample_value), no white space.  (Anything else is for the CPU parser to judge.)
bool samples_are_compact(const char* p, const char* e) {
  if (p >= e || *p++ != '[') return false;
  if (p < e && *p == ']') return p + 1 == e;
  while (true) {
    if (p >= e || *p++ != '[') return false;
    const char* d = p;
    while (p < e && *p >= '0' && *p <= '9') ++p;
    if (p == d) return false;
    if (p < e && *p == '.') {
      d = ++p;
      while (p < e && *p >= '0' && *p <= '9') ++p;
      if (p == d) return false;
    }
    if (e - p < 2 || p[0] != ',' || p[1] != '"') return false;
    p += 2;
    d = p;
    while (p < e && *p != '"') ++p;
    double v;
    if (e - p < 2 || p[1] != ']' || !detail::strict_sample_value(d, p, &v)) return false;
    p += 2;
    if (p >= e) return false;
    if (*p == ']') return p + 1 == e;
    if (*p++ != ',') return false;
  }
}

// Walk the series of one response with the device's marker lists, as they arrive: the upload + scan runs as a
// pipeline (TextDevice::scan_begin / scan_next).  Every batch of series whose markers are in goes through three steps
// while later chunks of the text are still on their way to the GPU:
//   walk    (this thread, markers only)   series i <-> i-th `},"values":[`; its list ends at the first `"]]` behind
//                                         it, or is empty when that lies beyond the next series' marker
//   labels  (worker threads, per series)  check the framing bytes around the markers, then either recognise the
//                                         series by the hash of its label bytes (daemon mode) or parse the label
//                                         map in place and pull out the six labels the assignment needs
//   assign  (this thread, in text order)  label set -> (pod, slot) through `asg`
void plan_text(TextDevice& dev, TextPlan& plan, Assigner& asg, Window& w, bool is_power, bool is_prof,
               DeviceIngestReport& rep, bool remember) {
  const std::string& t = *plan.text;
  size_t at0;
  bool bare;
  if (starts_with(t, 0, kHead)) at0 = sizeof kHead - 1, bare = false;
  else if (starts_with(t, 0, "[")) at0 = 1, bare = true;
  else throw NotCompact{"response does not start with the compact success/matrix header"};

  struct Loc {
    size_t at, close_brace, list_close;  // series object starts at `at`; '}' of the label map; ']' closing the list
    bool empty;                          // "values":[]
  };
  struct Rec {
    const char* err = nullptr;  // framing violation (a string literal)
    bool known = false, flat = false;
    Assigner::Result r = Assigner::Skipped;
    uint32_t pod = 0, slot = 0;
    uint64_t h1 = 0, h2 = 0;
    LabelFields f;
  };
  std::vector<uint64_t> opens, closes, co, cc;  // all markers so far (sorted: chunks come in text order)
  std::vector<Loc> locs;
  std::vector<Rec> recs;
  size_t oi = 0, ci = 0;     // next series' open marker; cursor into closes
  size_t at = at0;           // where the next series object starts
  const size_t kM = sizeof kMetric - 1;

  // consume every series that is completely covered by the markers delivered so far
  auto drain = [&](bool final) {
    // ---- walk: markers only ----------------------------------------------------------------------------------------
    const auto tw = std::chrono::steady_clock::now();
    locs.clear();
    while (oi < opens.size()) {
      if (!final && oi + 1 >= opens.size()) break;  // emptiness of list i is decided by where series i + 1 begins
      const size_t close_brace = (size_t)opens[oi], vb = close_brace + 12;
      if (close_brace < at + kM) throw NotCompact{"values marker inside the framing of a series"};
      while (ci < closes.size() && closes[ci] < vb) ++ci;
      const uint64_t next_open = oi + 1 < opens.size() ? opens[oi + 1] : (uint64_t)t.size();
      Loc l;
      l.at = at, l.close_brace = close_brace;
      l.empty = ci == closes.size() || closes[ci] > next_open;
      l.list_close = l.empty ? vb : (size_t)closes[ci] + 2;
      if (l.list_close + 3 > t.size()) throw NotCompact{"truncated values list"};
      locs.push_back(l);
      at = l.list_close + 3;  // past `]},`
      ++oi;
    }
    // ---- labels: per series, in parallel --------------------------------------------------------------------------
    recs.resize(locs.size());  // (every worker resets the records of its share)
    label_workers().run(locs.size(), [&](size_t b, size_t e) {
      FlatLabels flat;
      for (size_t i = b; i < e; ++i) {
        const Loc& l = locs[i];
        Rec& r = recs[i];
        r = Rec();
        const size_t mb = l.at + kM, vb = l.close_brace + 12;
        if (!starts_with(t, l.at, kMetric) || t[mb] != '{') r.err = "series does not start with {\"metric\":{";
        else if (t[vb] != (l.empty ? ']' : '[')) r.err = l.empty ? "unterminated values list" : "values list does not start with a sample";
        else if (t[l.list_close + 1] != '}') r.err = "series object has members after \"values\"";
        if (r.err) continue;
        const char* lb = t.data() + mb;
        const char* le = t.data() + l.close_brace + 1;
        if (remember && !l.empty) {
          Assigner::series_identity(std::string_view(lb, (size_t)(le - lb)), is_power, is_prof, &r.h1, &r.h2);
          r.known = asg.find_known(r.h1, r.h2, &r.r, &r.pod, &r.slot);
          if (r.known) continue;  // same bytes as in an earlier tick: validated then, same row now
        }
        // The label map is parsed for EVERY series (also the ones whose list is empty): a complete JSON object ending
        // exactly at the marker's '}' proves that [mb, close_brace] is the whole map and that no series without a
        // "values" member was jumped over.  (A `},"values":[` inside a label value is impossible: a raw '"' ends a
        // JSON string.)  Prometheus' own shape — string values, no escapes — is read in place without allocating.
        r.flat = flat.parse(lb, le);
        if (r.flat && !l.empty) r.f = extract_fields(flat);
      }
    });
    rep.labels_ms += ms_since(tw);  // walk + parallel label phase
    // ---- assign: in text order ----------------------------------------------------------------------------------------
    const auto ta = std::chrono::steady_clock::now();
    FlatLabels flat;
    for (size_t i = 0; i < locs.size(); ++i) {
      const Loc& l = locs[i];
      Rec& r = recs[i];
      if (r.err) throw NotCompact{r.err};
      // between series exactly one ',' ; the last one is followed by the ']' of the result array
      const char sep = t[l.list_close + 2];
      const bool last = final && oi == opens.size() && i + 1 == locs.size();
      if (sep != (last ? ']' : ',')) throw NotCompact{"expected ',' or ']' after a series"};
      ++w.stats.series_in;
      if (l.empty) {
        if (!r.flat) {  // still has to be a label map
          try {
            if (!Json::parse(std::string(t.data() + l.at + kM, t.data() + l.close_brace + 1)).is_object())
              throw NotCompact{"label map is not an object"};
          } catch (const std::exception& ex) {
            throw NotCompact{std::string("label map: ") + ex.what()};
          }
        }
        continue;  // an empty list is no element (as in the CPU paths)
      }
      const char* lb = t.data() + l.at + kM;
      const char* le = t.data() + l.close_brace + 1;
      if (r.known) {
        if (r.r == Assigner::Skipped) asg.count_skipped();
      } else if (r.flat) {
        r.r = asg.assign_fields(r.f, is_power, is_prof, [&]() {
          flat.parse(lb, le);
          return label_signature(flat);
        }, &r.pod, &r.slot);
        if (remember) asg.insert_known(r.h1, r.h2, r.r, r.pod, r.slot);
      } else {
        Json metric;
        try {
          metric = Json::parse(std::string(lb, le));
        } catch (const std::exception& ex) {
          throw NotCompact{std::string("label map: ") + ex.what()};
        }
        if (!metric.is_object()) throw NotCompact{"label map is not an object"};
        r.r = asg.assign(metric, is_power, is_prof, &r.pod, &r.slot);
        if (remember) asg.insert_known(r.h1, r.h2, r.r, r.pod, r.slot);
      }
      if (r.r == Assigner::Placed)
        plan.series.push_back(DevSeries{r.pod, r.slot, (uint64_t)l.close_brace + 12, (uint64_t)l.list_close});
      // A series that gets no row (no workload-pod label, lib.rs:161-175) is never seen by the device parser, so its
      // samples are checked here: a response that is malformed there is malformed, and the CPU parser — which reads
      // every sample — has to be the one to say so.  Such series are rare by construction (the selector asks for a
      // non-empty pod label).  Series shadowed by a PROF series of the same label set are NOT walked: there can be as
      // many of them as there are series, and nothing in them can reach the verdict.
      else if (r.r == Assigner::Skipped && !samples_are_compact(t.data() + l.close_brace + 11, t.data() + l.list_close + 1))
        throw NotCompact{"values list of a skipped series is not in the compact encoding"};
    }
    rep.assign_ms += ms_since(ta);
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
      // batches of a few thousand series keep the workers' hand-over cost negligible
      if (!more || opens.size() - oi >= 4096) drain(!more);  // (timed inside: labels_ms / assign_ms)
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
  // the result array closes right behind the last series (or at once when there is none)
  if (opens.empty()) {
    if (at0 >= t.size() || t[at0] != ']') throw NotCompact{"label map without a following \"values\" list"};
    at = at0 + 1;
  }
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
    // From here on the session's rows and the ring are being changed: whatever interrupts this tick (a device error
    // in the parse, a malformed slice) must not leave a ring that claims to hold the window ending at this tick —
    // a slice that never arrived reads as "no samples", and a busy GPU would look idle.  result() sets it again.
    st.valid = false;
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
  } catch (const DeviceDeclined& e) {
    return cpu(e.what());
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
