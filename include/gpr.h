/*
 * gpr.h — C ABI of the B200 idle-decision engine (libgpr.so).
 *
 * This is the drop-in boundary for gpu-pruner's one data-parallel path: the per-pod
 * windowed aggregation of DCGM GPU-utilisation samples into an idle/active verdict.
 * In the reference that arithmetic is a PromQL expression evaluated by a remote
 * Prometheus server; the seam this ABI replaces is, in the reference tree,
 *
 *     gpu-pruner/src/main.rs:397      client.query(query).get().await        (send PromQL)
 *     gpu-pruner/src/main.rs:405-409  response.data().into_vector()          (decode result)
 *     gpu-pruner/src/main.rs:416-437  HashSet<(pod, namespace)> dedup        (ANY-GPU fold)
 *     gpu-pruner/src/main.rs:494,508  create_time >= now - lookback => skip  (age gate)
 *
 * i.e. "obtain window matrix -> gpr_decide() -> expand set bits to PodMetricData".
 * Everything after main.rs:444 (Kubernetes lookups, owner walk, scale patches) is
 * unchanged host logic.
 *
 * Conventions:
 *   - every entry point returns int: 0 = GPR_OK, negative = GPR_E_*; the message for the
 *     last failure on a context is gpr_last_error(ctx) (gpr_last_error(NULL) for a failed
 *     gpr_create).  Nothing aborts, exits or throws across this boundary, so the caller's
 *     failure accounting (main.rs:310-321, QUERY_FAILURES) keeps working.
 *   - the caller owns every in/out buffer; the library owns only what is behind gpr_ctx*.
 *   - a context is NOT re-entrant (one call at a time, matching the single caller at
 *     main.rs:297) but it is thread-agnostic: every entry point selects its device itself
 *     and keeps no thread-local state, because tokio may migrate the caller between ticks.
 *   - plain pointers and sizes only; no torch / C++ types.
 *
 *   - device buffers handed to the library must be complete when the call is made: the context's
 *     own stream is not ordered with the caller's streams.  Either synchronise first or give the
 *     context the caller's stream (gpr_config.stream), in which case stream order is enough.
 *
 * There is no CPU fallback: without a CUDA device gpr_create fails with GPR_E_CUDA.
 */
#ifndef GPR_H_
#define GPR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define GPR_API __attribute__((visibility("default")))
#else
#define GPR_API
#endif

#define GPR_VERSION_MAJOR 0
#define GPR_VERSION_MINOR 2
#define GPR_VERSION_PATCH 0

/* ---- status codes ------------------------------------------------------------------- */
enum {
  GPR_OK = 0,
  GPR_E_INVALID = -1,   /* bad argument / shape / struct_size                           */
  GPR_E_CUDA = -2,      /* CUDA runtime or driver error (message has the CUDA string)    */
  GPR_E_NOMEM = -3,     /* allocation failed (host or device)                            */
  GPR_E_CAPACITY = -4,  /* window larger than the capacity given at gpr_create           */
  GPR_E_STATE = -5,     /* call not valid in this state (e.g. resident window not set)   */
  GPR_E_NCCL = -6,      /* NCCL error or NCCL library not loadable                       */
  GPR_E_UNSUPPORTED = -7
};

/* ---- where a buffer lives ------------------------------------------------------------ */
enum {
  GPR_MEM_HOST = 0,   /* host memory (pinned via gpr_host_alloc for full PCIe speed)     */
  GPR_MEM_DEVICE = 1  /* device memory on the context's GPU                              */
};

/* ---- kernel selection (both variants are always built) ------- */
enum {
  GPR_KERNEL_AUTO = 0,
  GPR_KERNEL_LDG = 1,  /* 128-bit ld.global.nc streaming loads, warp per series          */
  GPR_KERNEL_TMA = 2   /* cp.async.bulk (TMA) rows into an mbarrier-guarded smem ring     */
};

/* ---- sample format of gpr_window.util (gpr_window.util_format) ----------------------------
 * DCGM_FI_DEV_GPU_UTIL is an integer percentage, so a caller that parses the range query itself
 * can hand the window over at one byte per sample: a quarter of the PCIe / HBM bytes of the f32
 * layout and the same verdict.  The power plane and the resident ring stay f32.               */
enum {
  GPR_FMT_F32 = 0,  /* f32, NaN = no sample                                               */
  GPR_FMT_U8B = 1   /* biased u8: 0 = no sample, b in 1..255 = sample value b - 1 (0..254);
                       a zero-filled buffer is an all-absent window, like a NaN-filled f32 one;
                       `util` then points at bytes and row_stride counts bytes               */
};

/* ---- gpr_config.flags ---------------------------------------------------------------- */
#define GPR_F_POWER_PLANE 0x1u /* reserve staging for the power plane (host windows)      */
#define GPR_F_BLOCK_INDEX 0x2u /* gpr_resident_init: keep the max of every 64-sample block of the
                                  resident rows up to date in gpr_append and decide on that index —
                                  identical verdict and series_max, 1/64 of the bytes per tick      */

typedef struct gpr_ctx gpr_ctx;

/* Creation-time configuration.  Set struct_size = sizeof(gpr_config).                    */
typedef struct gpr_config {
  uint32_t struct_size;
  int32_t device;          /* CUDA device ordinal                                        */
  uint32_t max_pods;       /* capacity for HOST windows (staging in HBM); 0 = none        */
  uint32_t max_gpus;
  uint32_t max_samples;
  uint32_t flags;          /* GPR_F_*                                                    */
  int32_t kernel_variant;  /* GPR_KERNEL_*                                               */
  int32_t reserved0;
  void *stream;            /* optional caller-owned cudaStream_t all work is ordered on;
                              NULL = the context creates its own non-blocking stream      */
} gpr_config;

/*
 * One window = the range-vector result laid out densely.
 *
 *   n_gpus          series slots per pod, 1..256 (GPR_E_UNSUPPORTED above).
 *   capacity        for HOST windows only the number of cells counts: n_pods * n_gpus * n_samples must not
 *                   exceed max_pods * max_gpus * max_samples of gpr_create (staging is dense).
 *   util[p][g][t]   f32, t fastest; NaN = "no sample" (stale / absent / scrape gap); or biased
 *                   bytes when util_format = GPR_FMT_U8B (cast the pointer).
 *                   Restates DCGM_FI_DEV_GPU_UTIL{pod != ""}[Nm]   (query.promql.j2:16-20)
 *   power[p][g][t]  f32 or NULL.  Restates DCGM_FI_DEV_POWER_USAGE{...}[Nm]
 *                   (query.promql.j2:39-42).  Used only if power_threshold is "truthy".
 *   row_stride      elements between consecutive (p,g) rows; 0 means n_samples.
 *   power_threshold watts; the veto clause exists iff power != NULL and the threshold is
 *                   neither 0.0 nor NaN (Jinja truthiness of `args.power_threshold`,
 *                   query.promql.j2:36).  veto(p) = any g: max_t power[p][g][:] >= threshold.
 *   eligible[p]     u8 or NULL: 0 = pod fails the Pending / missing-timestamp gates
 *                   (main.rs:473-492).  NULL = all eligible.
 *   created_ts[p]   i64 or NULL: pod creation time in caller-chosen ticks; the pod is skipped
 *                   iff created_ts[p] >= cutoff_ts   (main.rs:494,508-510, cutoff = now -
 *                   (duration*60 + grace_period)).  INT64_MAX = "no creationTimestamp".
 */
typedef struct gpr_window {
  uint32_t struct_size;
  int32_t mem_kind;        /* GPR_MEM_*: applies to util, power, eligible, created_ts     */
  const float *util;
  const float *power;
  const uint8_t *eligible;
  const int64_t *created_ts;
  int64_t cutoff_ts;
  uint32_t n_pods;
  uint32_t n_gpus;
  uint32_t n_samples;
  uint32_t util_format;    /* GPR_FMT_*; ignored by gpr_decide_resident (the ring is f32)   */
  uint64_t row_stride;
  double power_threshold;
} gpr_window;

/*
 * Result.  Bitmaps are packed little-endian within a word: pod p is bit (p & 31) of word
 * (p >> 5); padding bits above n_pods are zero.  With a communicator attached
 * (gpr_comm_init) the bitmaps cover all ranks: world * ceil(n_pods/32) words, rank-major,
 * and n_pods must be a multiple of 32 and identical on every rank.
 *
 *   decision_bits   required.  candidate(p) && eligible(p)         (after main.rs:473-510)
 *   candidate_bits  optional.  (any g: max_t util == 0) && !veto(p) (what Prometheus + the
 *                   dedup at main.rs:416-437 return)
 *   series_max      optional, n_pods * n_gpus f32: window max per series, NaN if no sample
 *                   (the reference's reported `value` is this / 100; lib.rs:184,
 *                   query.promql.j2:20).
 *   veto_bits       optional, ceil(n_pods/32) words, THIS rank's pods only (never exchanged): pods with a
 *                   power series at or above the threshold (the `unless on (pod, namespace)` clause,
 *                   query.promql.j2:36-44).  With series_max it lets a caller re-derive a pod's verdict
 *                   (exact `sum by` of duplicate series, gpu-pruner_b200/host/ingest.cpp).
 *   out_mem_kind    where the buffers above live.
 *   n_series        number of idle series in non-vetoed pods = QueryResponse.num_pods
 *                   (main.rs:418; a series count despite the name).
 *   n_candidates / n_decisions   popcounts of the two bitmaps (this rank's pods).
 *   kernel_ms       device time of the decision kernel(s) for this call (CUDA events);
 *                   0 from the _async entry point.
 */
typedef struct gpr_result {
  uint32_t struct_size;
  int32_t out_mem_kind;
  uint32_t *decision_bits;
  uint32_t *candidate_bits;
  float *series_max;
  uint32_t *veto_bits;
  uint64_t n_series;
  uint64_t n_candidates;
  uint64_t n_decisions;
  double kernel_ms;
} gpr_result;

/* ---- lifecycle ----------------------------------------------------------------------- */
GPR_API int gpr_version(void); /* major*10000 + minor*100 + patch */
GPR_API int gpr_create(const gpr_config *cfg, gpr_ctx **out);
GPR_API void gpr_destroy(gpr_ctx *ctx);
GPR_API const char *gpr_last_error(const gpr_ctx *ctx);

/* ---- the hot path -------------------------------------------------------------------- */
/* Blocking: on return the result buffers and counters are complete.                      */
GPR_API int gpr_decide(gpr_ctx *ctx, const gpr_window *win, gpr_result *res);
/* Enqueue only (device or pinned-host buffers); counters are filled by gpr_sync().       */
GPR_API int gpr_decide_async(gpr_ctx *ctx, const gpr_window *win, gpr_result *res);
GPR_API int gpr_sync(gpr_ctx *ctx);
/* Enqueue n independent decisions (windows[i] -> results[i]) in one call: the same as n calls of
 * gpr_decide_async, without n trips through the caller's FFI.  At most 256 results may be
 * outstanding between two gpr_sync calls.  Stops at the first failing window and returns its code. */
GPR_API int gpr_decide_batch_async(gpr_ctx *ctx, const gpr_window *windows, gpr_result *results,
                                   uint32_t n);

/* ---- resident window for daemon mode (--daemon-mode / --check-interval, main.rs:286-330)
 * The window lives in HBM as a ring over the time axis; each tick appends the columns that
 * arrived since the previous tick and rescans.  max is order-independent so ring order is
 * irrelevant to the verdict.                                                              */
GPR_API int gpr_resident_init(gpr_ctx *ctx, uint32_t n_pods, uint32_t n_gpus, uint32_t n_samples,
                      uint32_t flags /* GPR_F_POWER_PLANE | GPR_F_BLOCK_INDEX */);
/* new columns laid out [p][g][n_new] (row_stride 0 = n_new); power_cols may be NULL.      */
GPR_API int gpr_append(gpr_ctx *ctx, const float *util_cols, const float *power_cols, uint32_t n_new,
               uint64_t row_stride, int32_t mem_kind);
/* Open the next n_new buckets of the ring without data: their columns become "no sample" in every row of
 * every resident plane and the ring head moves on.  The tick's samples are then merged in by
 * gpr_text_parse(GPR_TEXT_RESIDENT) (device-side ingest of the tick's range-query slice).  With
 * GPR_F_BLOCK_INDEX call gpr_resident_reindex after the parse.                                      */
GPR_API int gpr_resident_advance(gpr_ctx *ctx, uint32_t n_new);
/* rebuild the GPR_F_BLOCK_INDEX index after writing the resident planes directly
 * (gpr_resident_planes); a no-op without an index                                          */
GPR_API int gpr_resident_reindex(gpr_ctx *ctx);
/* win->util / win->power are ignored (resident planes are used); gates come from win.     */
GPR_API int gpr_decide_resident(gpr_ctx *ctx, const gpr_window *win, gpr_result *res);
/* device pointers of the resident planes (for generators / inspection); power may be NULL */
GPR_API int gpr_resident_planes(gpr_ctx *ctx, float **util, float **power, uint64_t *row_stride);
/* ring position the next appended bucket goes to; the newest bucket is at (head + n_samples - 1) % n_samples  */
GPR_API int gpr_resident_head(gpr_ctx *ctx, uint32_t *head);

/* ---- multi-GPU: one process per GPU, pods sharded by rank, one allgather of the bitmap - */
#define GPR_UNIQUE_ID_BYTES 128
GPR_API int gpr_comm_unique_id(void *id128);                       /* rank 0; ship to the others */
GPR_API int gpr_comm_init(gpr_ctx *ctx, const void *id128, int rank, int world);
GPR_API int gpr_comm_destroy(gpr_ctx *ctx);

/* ---- multi-GPU, fused: the decision kernel itself exchanges the bitmap over NVLink peer memory.
 * Each rank: gpr_p2p_init -> ship the 64-byte handle to every rank (any side channel) ->
 * gpr_p2p_attach(all handles, rank-major).  Afterwards gpr_decide behaves as with a communicator
 * (global rank-major bitmaps, n_pods a multiple of 32 and <= max_pods_per_rank, all ranks calling
 * in lock-step) but launches no collective: the fold kernel's CTAs store this rank's words into every peer's
 * buffer as 64-bit {step tag, word} slots the moment they exist (no fence, no flag), and its last CTA reads the
 * peers' slots until their tags match and assembles the result.  (GPR_EXCHANGE=flags selects the older protocol:
 * words, one system-scope fence, one flag per peer; GPR_EXCHANGE=pipelined lets a decision's exchange overlap
 * with its predecessor's — only the write of the caller's outputs stays ordered.)  At most 8 ranks.  A peer that never arrives does not hang the
 * GPU: the wait gives up after 20 s and the next gpr_sync / blocking call returns GPR_E_STATE.                  */
#define GPR_P2P_HANDLE_BYTES 64
GPR_API int gpr_p2p_init(gpr_ctx *ctx, int rank, int world, uint32_t max_pods_per_rank,
                         void *handle64);
GPR_API int gpr_p2p_attach(gpr_ctx *ctx, const void *handles /* world * 64 bytes */);
/* Timing switch for attributing the cost of the fused exchange (bench.py's breakdown); all ranks must
 * switch together.  0 = normal; 1 = push the words and flags but do not wait for the peers; 2 = no push
 * at all.  In modes 1 and 2 the returned bitmaps are NOT global.                                    */
GPR_API int gpr_p2p_debug(gpr_ctx *ctx, int32_t mode);

/* ---- memory helpers ------------------------------------------------------------------ */
GPR_API int gpr_host_alloc(gpr_ctx *ctx, size_t bytes, void **out); /* pinned host memory         */
GPR_API int gpr_host_free(gpr_ctx *ctx, void *p);
GPR_API int gpr_device_alloc(gpr_ctx *ctx, size_t bytes, void **out);
GPR_API int gpr_device_free(gpr_ctx *ctx, void *p);
GPR_API int gpr_memcpy(gpr_ctx *ctx, void *dst, const void *src, size_t bytes,
               int32_t dst_kind, int32_t src_kind);          /* blocking                   */

/* ---- measurement support ------------------------------------------------------------- */
/* CUDA events on the context's stream: begin; ...enqueue...; end -> elapsed ms.
 * gpr_timer_begin first enqueues a device-side rendezvous: with a fused exchange attached
 * (gpr_p2p_attach) every rank's stream waits, on the GPU, until all ranks have reached their
 * gpr_timer_begin, so the timed regions of all ranks start within an NVLink round trip of each other
 * whatever the skew between the host threads (with gpr_comm_init only, a one-word ncclAllGather plays
 * that role).  It is therefore COLLECTIVE when world > 1: all ranks must call it in lock-step.     */
GPR_API int gpr_timer_begin(gpr_ctx *ctx);
GPR_API int gpr_timer_end(gpr_ctx *ctx, double *ms);
/* Per-decision completion times: ns[i] = the device's %globaltimer (nanoseconds) at which the i-th of
 * the decisions retired by the most recent gpr_sync / blocking call finished (bitmap complete, exchange
 * included); *begin_ns = the same clock at the release of the last gpr_timer_begin.  *n = number of
 * decisions retired (may exceed cap).  Differences of consecutive stamps are the per-step device times
 * SURVEY.md §8(d) asks the median of.                                                              */
GPR_API int gpr_step_stamps(gpr_ctx *ctx, uint64_t *ns, uint32_t cap, uint32_t *n, uint64_t *begin_ns);
/* Four more stamps per retired decision, for attributing the time of the fold / exchange kernel: the fold kernel's
 * start (the reduce has completed and the previous fold is done), fold finished, peer flags raised (stores and
 * system-scope fence done; 0 without an exchange), all peers' words arrived (0 without an exchange).           */
GPR_API int gpr_phase_stamps(gpr_ctx *ctx, uint64_t *ns, uint32_t cap, uint32_t *n);
/* writes > L2-size bytes so the next launch starts with a cold L2                          */
GPR_API int gpr_flush_l2(gpr_ctx *ctx);
/* number of kernels this context has launched since creation                               */
GPR_API int gpr_launch_count(const gpr_ctx *ctx, uint64_t *n);
/* device facts: sm_count, l2 bytes, total HBM bytes, cc major/minor                        */
typedef struct gpr_device_info {
  uint32_t struct_size;
  int32_t sm_count;
  int32_t cc_major, cc_minor;
  uint64_t l2_bytes;
  uint64_t hbm_bytes;
  char name[64];
} gpr_device_info;
GPR_API int gpr_get_device_info(gpr_ctx *ctx, gpr_device_info *info);

/* ---- synthetic DCGM windows; counter-based so any implementation can
 * regenerate any cell.  plane: 0 = util, 1 = power.  Fills rows for pods
 * [pod_offset, pod_offset + n_pods) of a seeded (.., n_gpus, n_samples) universe into dst
 * (device memory, row_stride 0 = n_samples).  elig/created are optional device outputs.    */
GPR_API int gpr_synth_fill(gpr_ctx *ctx, uint64_t seed, int32_t plane, float *dst, uint64_t pod_offset,
                   uint32_t n_pods, uint32_t n_gpus, uint32_t n_samples, uint64_t row_stride);
GPR_API int gpr_synth_eligible(gpr_ctx *ctx, uint64_t seed, uint8_t *dst, uint64_t pod_offset,
                       uint32_t n_pods);

/* ---- device-side ingest of the range-query response TEXT ---------------------------------------
 * The step before the hot path: Prometheus' matrix JSON (the shape gpu-pruner/src/bin/querytest.rs:41-53
 * walks; series = label map + [[<unix time>, "<value>"], ...]) is parsed on the GPU straight into the
 * dense tensor in HBM, so the f32 window never exists on the host.  Division of labour:
 *   gpr_text_scan    uploads the text and reports where every sample list opens (`},"values":[`)
 *                    and closes (`"]]`);
 *   the caller       parses the label maps (~1 % of the bytes) and assigns every series its tensor
 *                    row — label precedence of lib.rs:153-187, `sum by` groups of query.promql.j2:9
 *                    (gpu-pruner_b200/host/ingest_device.cpp does this);
 *   gpr_text_parse   parses all samples of the given spans into a plane [n_rows][n_samples] f32
 *                    (0xFFFFFFFF, a NaN = no sample) — a context-owned plane, or the resident ring of
 *                    daemon mode.
 * Where a sample goes (the same rule as the CPU ingest, gpu-pruner_b200/host/ingest_internal.hpp):
 *   inside the window iff  t_end - window_seconds < ts <= t_end      (PromQL [Nm] at t_end, left-open)
 *   bucket back = (t_end - ts) / step  (0 = newest), column n_samples - 1 - back; several samples of a
 *   row in one bucket are merged with a NaN-aware max — which is what max_over_time over the row
 *   (query.promql.j2:10,16) computes anyway, so collisions, sample order and duplicate series need no
 *   special handling.
 * Numbers are converted exactly like strtod + (float): Clinger's fast path or Eisel-Lemire (17-digit
 * DCGM_FI_PROF_GR_ENGINE_ACTIVE ratios included).  The device parser is strict: anything but
 * `[digits[.digits],"<decimal, at most 19 significant digits>|NaN|+Inf|-Inf"]` sets GPR_SPAN_HARD on the span;
 * the caller re-parses the rows of hard spans on the CPU and overwrites them with gpr_memcpy, so the
 * tensor equals a CPU ingest for every input.
 */
typedef struct gpr_text_span {
  uint64_t begin;    /* offset of the first byte after `"values":[` (a '[')                    */
  uint64_t end;      /* offset of the ']' that closes the sample list                          */
  uint32_t row;      /* destination row = pod * n_gpus + slot                                  */
  uint32_t flags;    /* GPR_SPAN_SHARED in; GPR_SPAN_HARD out                                  */
  uint32_t n_in;     /* out: samples parsed                                                    */
  uint32_t n_oow;    /* out: samples outside the window                                        */
  uint32_t n_tiny;   /* out: non-zero values below the f32 denormal range, kept non-zero       */
  uint32_t reserved;
} gpr_text_span;
#define GPR_SPAN_SHARED 1u /* several series feed this row (informational; every merge is atomic)   */
#define GPR_SPAN_HARD 2u   /* the device parser gave up on this span: re-parse its row on the CPU */

typedef struct gpr_text_grid {
  uint32_t struct_size;
  uint32_t flags;          /* GPR_TEXT_*                                                            */
  int64_t t_end;           /* newest second of the window (inclusive)                               */
  int64_t window_seconds;  /* samples with t_end - window_seconds < ts <= t_end are inside          */
  int64_t step;            /* seconds per column, > 0                                               */
  uint32_t n_samples;      /* columns; >= ceil(window_seconds / step)                               */
  uint32_t n_rows;
} gpr_text_grid;
#define GPR_TEXT_FILL 1u     /* fill the destination plane with "no sample" first (context planes)    */
#define GPR_TEXT_RESIDENT 2u /* destination = the resident ring (gpr_resident_init): n_samples must be its
                                n_samples, n_rows <= its rows; the newest bucket is the ring's newest
                                column (call gpr_resident_advance first to open the tick's buckets)  */

/* Copy `n_bytes` of response text to the device (pinned host memory from gpr_host_alloc moves at
 * full PCIe speed; ordinary memory is staged through a pinned ring by a few host threads, and every
 * chunk is scanned as it lands) and scan it.  Up to `cap` offsets are written to each of opens[]
 * (position of the '}' of `},"values":[`) and closes[] (position of the '"' of `"]]`), UNSORTED; the
 * true counts are returned in *n_opens / *n_closes (GPR_E_CAPACITY if either exceeds cap).  The text
 * stays resident in the context's slot `slot` (0..2: a tick has up to three responses — PROF, UTIL,
 * POWER) for gpr_text_parse until the next gpr_text_scan of that slot.  Blocking.                   */
GPR_API int gpr_text_scan(gpr_ctx *ctx, int32_t slot, const char *text, uint64_t n_bytes,
                          int32_t mem_kind, uint64_t *opens, uint64_t *closes, uint64_t cap,
                          uint64_t *n_opens, uint64_t *n_closes);
/* The same, as a pipeline the caller can work alongside: gpr_text_scan_begin starts the upload (producer threads
 * owned by the library) and returns; every gpr_text_scan_next blocks until the next chunk of the text (2 MB; 16 MB
 * for pinned text) has landed and been scanned and returns that chunk's markers (sorted; room for `cap` of each
 * kind — 16,384 always suffices) together with *bytes_done = how much of the text is covered so far.  *more = 0 with the last chunk
 * (or at once for an empty text); the scan is then complete and the text ready for gpr_text_parse.  Between the
 * calls the caller can already walk the series whose markers it has (gpu-pruner_b200/host/ingest_device.cpp turns
 * label maps into tensor rows while later chunks are still crossing PCIe).  Dropped by the next
 * gpr_text_scan_begin / gpr_text_scan / gpr_destroy if not run to the end.  `text` must stay valid until then. */
GPR_API int gpr_text_scan_begin(gpr_ctx *ctx, int32_t slot, const char *text, uint64_t n_bytes, int32_t mem_kind);
GPR_API int gpr_text_scan_next(gpr_ctx *ctx, uint64_t *opens, uint64_t *closes, uint64_t cap, uint64_t *n_opens,
                               uint64_t *n_closes, uint64_t *bytes_done, int32_t *more);
/* Parse the samples of spans[0..n_spans) (host array, sorted by begin, non-overlapping) of the text in
 * `slot` into plane `plane` (0 = util, 1 = power).  Out-fields of the spans are filled.  Blocking.  */
GPR_API int gpr_text_parse(gpr_ctx *ctx, int32_t slot, gpr_text_span *spans, uint32_t n_spans,
                           const gpr_text_grid *grid, int32_t plane);
/* Device pointers of the context planes (NULL if never parsed); valid until the next gpr_text_parse
 * that has to grow them, or gpr_destroy.  Hand them to gpr_decide with mem_kind = GPR_MEM_DEVICE.   */
GPR_API int gpr_text_planes(gpr_ctx *ctx, float **util, float **power);

#ifdef __cplusplus
}
#endif
#endif /* GPR_H_ */
