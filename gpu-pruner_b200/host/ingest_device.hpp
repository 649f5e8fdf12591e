// ingest_device.hpp — the matrix wire format parsed ON THE GPU: the host keeps only what needs hashing
// and strings (label maps -> tensor rows, ~1 % of the bytes); the sample lists go to the device as
// text and land in the dense tensor in HBM (include/gpr.h, gpr_text_scan / gpr_text_parse).  The f32
// window never exists on the host.  Result: the same Window as ingest_matrix_text() — shape, pods,
// statistics — with `d_util` / `d_power` device planes instead of the host vectors, bit-identical
// cell for cell (tests/test_text_device_cpu.py on an emulated device, tests/test_gpu_text.py on the GPU).
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/gpr.h"
#include "ingest.hpp"

namespace gph {

// What the orchestration needs from the device; implemented over libgpr.so in gpr_engine.cpp.
// Thrown by a TextDevice that cannot take a response although nothing is wrong with the response (more series
// markers in one upload chunk than the scan has room for — label sets a tenth of DCGM's size): the session hands the
// response to the CPU parser instead of failing the tick.
struct DeviceDeclined : std::runtime_error {
  using std::runtime_error::runtime_error;
};

class TextDevice {
 public:
  virtual ~TextDevice() = default;
  // Upload `text` into resident slot `slot` (0..2) and scan it for `},"values":[` ('}' position) and `"]]`
  // ('"' position), as a pipeline: scan_begin starts the upload, every scan_next blocks until the next chunk of
  // the text has been scanned on the device and returns that chunk's markers (sorted, absolute offsets) and how
  // much of the text is covered; false with the last chunk.  The host works on the series it already has
  // markers for while later chunks are still crossing PCIe.
  virtual void scan_begin(int slot, const char* text, size_t n) = 0;
  virtual bool scan_next(std::vector<uint64_t>* opens, std::vector<uint64_t>* closes, uint64_t* bytes_done) = 0;
  // parse the samples of `spans` (sorted by begin) of the text in `slot` into plane 0 (util) / 1 (power):
  // samples with grid.t_end - grid.span < ts <= grid.t_end, bucket (t_end - ts) / step, NaN-aware max merge
  struct TextGrid {
    int64_t t_end = 0, span = 0, step = 1;
    uint32_t T = 0, n_rows = 0;
    bool fill = true;        // start from an all-"no sample" plane
    bool resident = false;   // destination = the resident ring of daemon mode instead of the context plane
  };
  virtual void parse(int slot, std::vector<gpr_text_span>& spans, const TextGrid& grid, int plane) = 0;
  // overwrite the newest `n_newest` buckets of `row` (chronological order in `data`; n_newest == T: the whole
  // row) — rows the strict device parser declined, re-parsed on the CPU
  virtual void patch_row(int plane, uint32_t row, uint32_t T, const float* data, uint32_t n_newest, bool resident) = 0;
  virtual const float* plane(int plane) = 0;
  // daemon mode: (re)create the resident ring [rows][T] (all "no sample"), and open the next n_new buckets
  virtual void resident_init(uint32_t pods, uint32_t G, uint32_t T, bool with_power) = 0;
  virtual void resident_advance(uint32_t n_new) = 0;
};

struct DeviceIngestReport {
  bool on_device = false;        // false: the CPU text path produced the window (reason says why)
  std::string reason;
  uint64_t spans = 0, hard_spans = 0, rows_patched = 0;
  // where the time went: device scan (text upload + marker scan), series walk over the markers,
  // label maps -> rows, device parse (NaN fill + sample parse)
  double scan_ms = 0, labels_ms = 0, assign_ms = 0, parse_ms = 0;
};

// One-shot: needs opt.t_end > 0 and opt.step > 0 (the caller issued the range query, so it knows both);
// otherwise, and for any response that is not in Prometheus' compact encoding, the CPU text path
// runs instead and the returned Window carries host vectors as usual.
Window ingest_matrix_device(TextDevice& dev, const std::string& util, const std::string* prof,
                            const std::string* power, const IngestOptions& opt,
                            DeviceIngestReport* report = nullptr);

// Daemon mode (main.rs:286-330): the row assignment and the window survive between ticks.  The first tick (and
// any tick after NeedFullWindow) ingests the full range query into the engine's resident ring; later ticks ingest
// only what was scraped since (opt.slice_seconds), appended to the ring.  Pods and slots keep their rows for the
// life of the session, so a pod whose series stop reporting simply ages out of the window.
class DeviceIngestSession {
 public:
  explicit DeviceIngestSession(TextDevice& dev);
  ~DeviceIngestSession();
  // opt.slice_seconds == 0: full window (re)build; > 0: delta — throws NeedFullWindow if it cannot be absorbed
  Window ingest(const std::string& util, const std::string* prof, const std::string* power, const IngestOptions& opt,
                DeviceIngestReport* report = nullptr);
  // newest second the resident window holds (0 = nothing resident): the next delta must start right after it
  int64_t resident_t_end() const;
  void invalidate();

 private:
  struct State;
  TextDevice& dev_;
  State* st_;
};

}  // namespace gph
