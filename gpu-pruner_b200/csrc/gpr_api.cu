// gpr_api.cu — the C ABI of include/gpr.h over the sm_100a kernels.
//
// Seam replaced in the reference (paths relative to /root/reference):
//   gpu-pruner/src/main.rs:397-437   send PromQL, decode instant vector, dedup by (pod, ns)
//   gpu-pruner/src/main.rs:494,508   age gate
// The library never falls back to a CPU path: every entry point that computes needs a CUDA
// device and reports GPR_E_CUDA otherwise.
#include "../../include/gpr.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <nvtx3/nvToolsExt.h>  // header-only; ranges show up in nsys / ncu --nvtx timelines

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "gpr_kernels.cuh"
#include "gpr_synth.cuh"
#include "gpr_text_kernels.cuh"

namespace {

constexpr int kLdgWarps = 16;
constexpr int kLdgUnroll = 8;
constexpr size_t kTmaSmemBudget = 208 * 1024;
constexpr int kSlots = 256;      // outstanding async results
constexpr int kMaxChunkEvents = 64;
constexpr int kExchangeDepth = 4;  // exchange buffer sets of the fused multi-GPU path = 2 x scratch sets

std::mutex g_err_mu;
char g_create_err[512] = "";

// ---- NCCL, loaded on demand so the library itself has no hard dependency on it ----------
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;

bool load_nccl(char* err, size_t n) {
  std::lock_guard<std::mutex> lk(g_nccl_mu);
  if (g_nccl.ok) return true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    g_nccl.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.handle) break;
  }
  if (!g_nccl.handle) {
    snprintf(err, n, "NCCL not loadable: %s", dlerror());
    return false;
  }
#define GPR_SYM(field, name)                                              \
  g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(g_nccl.handle, name)); \
  if (!g_nccl.field) {                                                    \
    snprintf(err, n, "NCCL symbol %s missing", name);                     \
    return false;                                                         \
  }
  GPR_SYM(GetUniqueId, "ncclGetUniqueId")
  GPR_SYM(CommInitRank, "ncclCommInitRank")
  GPR_SYM(CommDestroy, "ncclCommDestroy")
  GPR_SYM(AllGather, "ncclAllGather")
  GPR_SYM(GetErrorString, "ncclGetErrorString")
#undef GPR_SYM
  g_nccl.ok = true;
  return true;
}

struct Pending {
  gpr_result* res;
  int slot;
};

}  // namespace

struct gpr_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
  cudaEvent_t ev_join = nullptr;
  cudaEvent_t ev_chunk[kMaxChunkEvents] = {};
  int sm_count = 0;
  size_t l2_bytes = 0, hbm_bytes = 0;
  int cc_major = 0, cc_minor = 0;
  char name[64] = "";
  int variant = GPR_KERNEL_AUTO;
  int ldg_ctas_per_sm = 2;
  int tma_depth_max = 3;
  int tma_warps = 16;
  int tma_chunk_bytes = 8192;
  size_t chunk_bytes = 32u << 20;
  int parse_ctas_per_sm = 6;   // k_text_parse: 4 warps and 37 KB of shared memory per CTA (GPR_PARSE_CTAS)

  // capacity for host windows
  uint32_t max_pods = 0, max_gpus = 0, max_samples = 0;
  bool cap_power = false;
  float* d_util_stage = nullptr;
  float* d_power_stage = nullptr;
  uint8_t* d_elig_stage = nullptr;
  int64_t* d_created_stage = nullptr;
  size_t gate_cap = 0;

  // scratch (grown on demand)
  // Two scratch sets used alternately by successive single-launch decisions so that a launch may
  // start (programmatic dependent launch) while its predecessor is still folding.
  uint32_t* d_masks[2] = {nullptr, nullptr};  // each [idle P | veto P], all-zero between uses
  size_t masks_cap[2] = {0, 0};
  bool masks_dirty = false;     // a failed call may have left bits behind
  unsigned int* d_tickets = nullptr;          // [2] fold-grid tickets
  unsigned long long* d_acc = nullptr;        // [2][3] fold-grid count accumulators
  unsigned long long* d_done = nullptr;       // [2] completed folds per scratch set
  unsigned long long uses[2] = {0, 0};        // folds issued per scratch set
  unsigned parity = 0;
  bool pdl_enabled = true;      // GPR_PDL=0 disables programmatic dependent launch
  bool last_was_reduce = false; // the newest op on the stream is one of our fold kernels
  uint32_t* d_bits = nullptr;  // [dbits W | cbits W]
  size_t bits_cap = 0;
  uint32_t* d_gather = nullptr;  // [world][2W]
  size_t gather_cap = 0;
  float* d_smax = nullptr;
  size_t smax_cap = 0;
  unsigned long long* h_counts = nullptr;  // pinned [kSlots][4]: n_series, n_candidates, n_decisions,
                                           // %globaltimer at completion; then [mark, error word]
  unsigned long long* h_mark = nullptr;    // %globaltimer written by the last gpr_timer_begin
  unsigned int* h_err = nullptr;           // raised by a kernel whose peer wait timed out
  std::vector<Pending> pending;
  std::vector<uint64_t> stamps;            // completion stamps of the decisions retired by the last gpr_sync
  std::vector<uint64_t> phase_stamps;      // 4 per decision: fold start, folded, flags raised, peers arrived
  unsigned long long rdv_seq = 0;          // rendezvous sequence number (same on all ranks)

  void* d_flush = nullptr;
  size_t flush_bytes = 0;

  // resident window (daemon mode)
  float* d_res_util = nullptr;
  float* d_res_power = nullptr;
  uint32_t res_P = 0, res_G = 0, res_T = 0, res_head = 0;
  // optional index (GPR_F_BLOCK_INDEX): max of every 64-sample block of every resident row, kept up to
  // date by gpr_append.  max-of-block-maxima == max-of-samples (NaN = block without a sample), so the
  // index is itself a window tensor with ceil(T/64) "samples" per series and the same kernels decide on it
  float* d_idx_util = nullptr;
  float* d_idx_power = nullptr;
  uint32_t idx_ld = 0;      // padded to a multiple of 4 (TMA-able rows), padding stays NaN
  float* d_cols = nullptr;
  size_t cols_cap = 0;

  // device-side ingest of response text (gpr_text_scan / gpr_text_parse)
  uint8_t* d_text[3] = {nullptr, nullptr, nullptr};
  size_t text_cap[3] = {0, 0, 0};
  uint64_t text_n[3] = {0, 0, 0};
  uint64_t* d_marks = nullptr;  // [2][marks_cap]
  size_t marks_cap = 0;         // total entries (both lists)
  unsigned long long* d_mark_counts = nullptr;
  gpr::text::Span* d_spans = nullptr;
  size_t spans_cap = 0;
  float* d_tplane[2] = {nullptr, nullptr};
  size_t tplane_cap[2] = {0, 0};
  // The text goes up in chunks through a few producer threads (pageable text is staged through a small pinned
  // ring) and every chunk is scanned as soon as it has landed; the markers of a chunk are written by the scan
  // kernel straight into a block of mapped pinned memory (ScanPipe, gpr_text_scan_begin / _next).
  static constexpr int kUpThreads = 16, kUpSlots = 2;
  // chunk size.  Pageable text: GPR_TEXT_CHUNK_MB (1..16), default 2 — the pinned staging ring is
  // up_threads x 2 x chunk, and page-locking it is paid by the first scan of a process (measured on the box:
  // 1, 2 and 4 MB chunks all reach ~41 GB/s with 8 producers).  Pinned / device text needs no staging and goes
  // in 16 MB pieces (49 GB/s).
  static constexpr size_t kPinnedChunk = 16u << 20;
  static constexpr int kMarkBlocks = 64;                   // ring of marker blocks (chunks in flight ahead of the consumer)
  static constexpr uint32_t kMarkCap = 16384;              // markers of one kind per chunk (2 MB: one per 128 B)
  size_t up_chunk = 2u << 20;
  size_t up_slot_bytes = 0;            // up_chunk + a page (the 16 bytes of overlap, page aligned)
  unsigned char* h_up_ring = nullptr;  // [up_threads][kUpSlots][up_slot_bytes], pinned
  uint32_t* h_mark_blocks = nullptr;   // [kMarkBlocks][2 + 2 * kMarkCap], pinned + device-mapped
  uint32_t* d_mark_blocks = nullptr;   // the same in device memory: the scan appends here (atomics), then publishes
  cudaStream_t up_stream[kUpThreads] = {};
  cudaEvent_t up_event[kUpThreads][kUpSlots] = {};
  cudaEvent_t mark_event[kMarkBlocks] = {};
  int up_threads = 8;                  // GPR_TEXT_UPLOAD_THREADS (1..16)
  struct ScanPipe* pipe = nullptr;     // the scan in progress (gpr_text_scan_begin .. last gpr_text_scan_next)

  // multi-GPU
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  // fused exchange over peer memory (gpr_p2p_init / gpr_p2p_attach)
  unsigned char* p2p_block = nullptr;            // [flags u64 x kMaxPeers | rendezvous flags | gather[4][world][stride] | slots[4][world][stride]]
  unsigned char* p2p_peer[gpr::kMaxPeers] = {};  // peer-mapped base of every rank's block (self = local)
  // Exchange buffers are kExchangeDepth deep (step % depth), twice the number of scratch sets: with the late
  // output ordering a rank may push step n + 4 only after every peer has consumed step n (see k_fold)
  size_t p2p_gather_off[kExchangeDepth] = {};
  size_t p2p_ll_off[kExchangeDepth] = {};        // tagged 64-bit slot arrays [world][stride], one per step % depth
  bool exchange_ll = true;                       // tagged 64-bit slots (default); GPR_EXCHANGE=flags: data + fence + flag
  bool exchange_late = false;                    // GPR_EXCHANGE=pipelined: tagged slots, and a fold waits for its
                                                 // predecessor only before it writes the caller's outputs
  int fold_threads = 256;                        // GPR_FOLD_THREADS: 64 / 128 / 256 threads per fold CTA
  uint32_t p2p_stride = 0;                       // words per rank slot = 2 * W_max
  bool p2p_ready = false;
  int exchange_debug = 0;                        // GPR_DEBUG_EXCHANGE (developer timing switch)
  unsigned int poll_ns = 4000;                   // GPR_POLL_NS: longest pause between polls of the peers' flags
  unsigned long long p2p_step = 0;

  uint64_t launches = 0;
  char err[512] = "";
};

void scan_pipe_abort(gpr_ctx* ctx);   // defined with the text ingest below

namespace {

int fail(gpr_ctx* c, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) {
    snprintf(c->err, sizeof c->err, "%s", buf);
  } else {
    std::lock_guard<std::mutex> lk(g_err_mu);
    snprintf(g_create_err, sizeof g_create_err, "%s", buf);
  }
  return code;
}

#define CU(call)                                                                          \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess)                                                                \
      return fail(ctx, e_ == cudaErrorMemoryAllocation ? GPR_E_NOMEM : GPR_E_CUDA,        \
                  "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);   \
  } while (0)

#define NC(call)                                                                          \
  do {                                                                                    \
    ncclResult_t r_ = (call);                                                             \
    if (r_ != ncclSuccess)                                                                \
      return fail(ctx, GPR_E_NCCL, "%s: %s", #call, g_nccl.GetErrorString(r_));           \
  } while (0)

template <typename T>
int grow(gpr_ctx* ctx, T** p, size_t* cap, size_t need) {
  if (need <= *cap) return GPR_OK;
  if (*p) {
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaFree(*p));
    *p = nullptr;
    *cap = 0;
  }
  size_t want = need + need / 4 + 64;
  CU(cudaMalloc(reinterpret_cast<void**>(p), want * sizeof(T)));
  *cap = want;
  return GPR_OK;
}

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// smallest f32 >= thr, so that (m >= thr_f) in f32 equals ((double)m >= thr) for every f32 m
float threshold_f32(double thr) {
  float t = (float)thr;
  if ((double)t < thr) t = nextafterf(t, INFINITY);
  return t;
}

bool power_truthy(double thr) { return thr != 0.0 && !std::isnan(thr); }

gpr::TmaLayout tma_layout(const gpr_ctx* ctx, uint32_t T, int nw) {
  gpr::TmaLayout L;
  const uint32_t row_bytes = T * 4u;
  const uint32_t max_chunk = (uint32_t)ctx->tma_chunk_bytes;
  L.n_chunks = (row_bytes + max_chunk - 1) / max_chunk;
  uint32_t ce = (T + L.n_chunks - 1) / L.n_chunks;
  ce = (ce + 3u) & ~3u;
  L.chunk_elems = ce;
  L.n_chunks = (T + ce - 1) / ce;
  L.stage_bytes = (ce * 4u + 127u) & ~127u;
  uint32_t d = (uint32_t)((kTmaSmemBudget - 1024) / ((size_t)L.stage_bytes * nw));
  d = std::min<uint32_t>(d, (uint32_t)ctx->tma_depth_max);
  L.depth = std::max<uint32_t>(d, 1u);
  return L;
}

size_t tma_smem_bytes(const gpr::TmaLayout& L, int nw) {
  return (size_t)nw * L.depth * L.stage_bytes + (size_t)nw * L.depth * sizeof(uint64_t);
}

// One launch helper for both variants; `pdl` adds the programmatic-stream-serialization
// attribute so the kernel may begin while the previous reduce kernel on the stream drains.
template <typename Kernel, typename... Args>
cudaError_t launch_ex(Kernel k, uint32_t grid, uint32_t block, size_t smem, cudaStream_t st, bool pdl,
                      Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid), cfg.blockDim = dim3(block), cfg.dynamicSmemBytes = smem, cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, k, args...);
}

template <int NW>
cudaError_t launch_tma(gpr_ctx* ctx, const gpr::ReduceParams& rp, uint32_t grid, bool pdl) {
  const gpr::TmaLayout L = tma_layout(ctx, rp.T, NW);
  return launch_ex(gpr::k_reduce_tma<NW>, grid, NW * 32, tma_smem_bytes(L, NW), ctx->stream, pdl, rp, L);
}

// launch one reduce pass over the rows described by rp
int launch_reduce(gpr_ctx* ctx, gpr::ReduceParams& rp, bool tma_ok, bool pdl) {
  if (rp.total_rows == 0) return GPR_OK;
  // AUTO = the TMA pipeline (measured winner on B200 at C2 and C3, profiles/README.md); rows that are
  // not 16-byte aligned or have T % 4 != 0 cannot be bulk-copied and take the LDG kernel
  int variant = ctx->variant == GPR_KERNEL_AUTO ? GPR_KERNEL_TMA : ctx->variant;
  if (variant == GPR_KERNEL_TMA && !tma_ok) variant = GPR_KERNEL_LDG;
  if (variant == GPR_KERNEL_TMA &&
      tma_smem_bytes(tma_layout(ctx, rp.T, ctx->tma_warps), ctx->tma_warps) > kTmaSmemBudget)
    variant = GPR_KERNEL_LDG;  // a tuning override (GPR_TMA_WARPS / GPR_TMA_CHUNK) that does not fit
  cudaError_t e;
  if (rp.util_u8) {  // biased-byte util plane: one kernel, any alignment (the power plane stays f32)
    uint32_t grid = (uint32_t)(ctx->sm_count * ctx->ldg_ctas_per_sm);
    const uint32_t need = (rp.total_rows + kLdgWarps - 1) / kLdgWarps;
    grid = std::max<uint32_t>(1u, std::min<uint32_t>(grid, need));
    e = launch_ex(gpr::k_reduce_u8<kLdgWarps, 4>, grid, kLdgWarps * 32, 0, ctx->stream, pdl, rp);
  } else if (variant == GPR_KERNEL_TMA) {
    const int nw = ctx->tma_warps;
    uint32_t grid = (uint32_t)ctx->sm_count;
    grid = std::max<uint32_t>(1u, std::min<uint32_t>(grid, (rp.total_rows + nw - 1) / nw));
    if (nw == 4) e = launch_tma<4>(ctx, rp, grid, pdl);
    else if (nw == 16) e = launch_tma<16>(ctx, rp, grid, pdl);
    else if (nw == 32) e = launch_tma<32>(ctx, rp, grid, pdl);
    else e = launch_tma<8>(ctx, rp, grid, pdl);
  } else {
    uint32_t grid = (uint32_t)(ctx->sm_count * ctx->ldg_ctas_per_sm);
    const uint32_t need = (rp.total_rows + kLdgWarps - 1) / kLdgWarps;
    grid = std::max<uint32_t>(1u, std::min<uint32_t>(grid, need));
    e = launch_ex(gpr::k_reduce_ldg<kLdgWarps, kLdgUnroll>, grid, kLdgWarps * 32, 0, ctx->stream, pdl, rp);
  }
  ctx->launches++;
  CU(e);
  CU(cudaGetLastError());
  return GPR_OK;
}

int copy_rows_h2d(gpr_ctx* ctx, void* dst, const void* src, size_t n_rows, uint32_t T,
                  uint64_t ld, size_t esize, cudaStream_t s) {
  if (n_rows == 0) return GPR_OK;
  if (ld == T) {
    CU(cudaMemcpyAsync(dst, src, n_rows * (size_t)T * esize, cudaMemcpyHostToDevice, s));
  } else {
    CU(cudaMemcpy2DAsync(dst, (size_t)T * esize, src, (size_t)ld * esize, (size_t)T * esize, n_rows,
                         cudaMemcpyHostToDevice, s));
  }
  return GPR_OK;
}

struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

int decide_impl(gpr_ctx* ctx, const gpr_window* win, gpr_result* res, bool resident, bool async) {
  if (!ctx) return GPR_E_INVALID;
  NvtxRange nvtx_range(resident ? "gpr_decide_resident" : "gpr_decide");
  if (!win || !res) return fail(ctx, GPR_E_INVALID, "window/result is NULL");
  if (win->struct_size != sizeof(gpr_window) || res->struct_size != sizeof(gpr_result))
    return fail(ctx, GPR_E_INVALID, "struct_size mismatch (window %u/%zu result %u/%zu)",
                win->struct_size, sizeof(gpr_window), res->struct_size, sizeof(gpr_result));
  CU(cudaSetDevice(ctx->device));

  uint32_t P = win->n_pods, G = win->n_gpus, T = win->n_samples;
  uint64_t ld = win->row_stride ? win->row_stride : T;
  const float* util = win->util;
  const float* power = win->power;
  int in_kind = win->mem_kind;
  const bool u8 = !resident && win->util_format == GPR_FMT_U8B;
  if (!resident && win->util_format != GPR_FMT_F32 && win->util_format != GPR_FMT_U8B)
    return fail(ctx, GPR_E_INVALID, "bad util_format %u", win->util_format);
  if (resident) {
    if (!ctx->d_res_util) return fail(ctx, GPR_E_STATE, "no resident window (gpr_resident_init)");
    P = ctx->res_P, G = ctx->res_G, T = ctx->res_T, ld = T;
    util = ctx->d_res_util;
    power = ctx->d_res_power;
    if (ctx->d_idx_util) {  // decide on the block-maxima index: same verdict, T/64 of the bytes
      T = ctx->idx_ld, ld = ctx->idx_ld;
      util = ctx->d_idx_util;
      power = ctx->d_idx_power;
    }
  }
  if (in_kind != GPR_MEM_HOST && in_kind != GPR_MEM_DEVICE)
    return fail(ctx, GPR_E_INVALID, "bad mem_kind %d", in_kind);
  if (res->out_mem_kind != GPR_MEM_HOST && res->out_mem_kind != GPR_MEM_DEVICE)
    return fail(ctx, GPR_E_INVALID, "bad out_mem_kind %d", res->out_mem_kind);
  if (G == 0 || T == 0) return fail(ctx, GPR_E_INVALID, "n_gpus and n_samples must be > 0");
  if (G > 256) return fail(ctx, GPR_E_UNSUPPORTED, "n_gpus %u > 256 series slots per pod is not supported", G);
  const uint32_t MW = (G + 31u) / 32u;  // mask words per pod
  if (ld < T) return fail(ctx, GPR_E_INVALID, "row_stride %llu < n_samples %u",
                          (unsigned long long)ld, T);
  if (P > 0 && !util) return fail(ctx, GPR_E_INVALID, "util is NULL");
  if (!res->decision_bits && P > 0) return fail(ctx, GPR_E_INVALID, "decision_bits is NULL");
  const uint64_t S64 = (uint64_t)P * G;
  if (S64 > 0x7fffffffull) return fail(ctx, GPR_E_INVALID, "too many series (%llu)",
                                       (unsigned long long)S64);
  const uint32_t S = (uint32_t)S64;
  const uint32_t W = (P + 31u) / 32u;
  const bool use_power = power != nullptr && power_truthy(win->power_threshold);
  const bool host_in = !resident && in_kind == GPR_MEM_HOST;
  const bool gates_host = in_kind == GPR_MEM_HOST;
  const bool host_out = res->out_mem_kind == GPR_MEM_HOST;
  const bool fused = ctx->p2p_ready && ctx->world > 1;
  const bool comm = (ctx->comm != nullptr || fused) && ctx->world > 1;
  if (fused && 2u * W > ctx->p2p_stride)
    return fail(ctx, GPR_E_CAPACITY, "n_pods %u exceeds the p2p exchange capacity (%u pods per rank)", P,
                ctx->p2p_stride * 16u);
  if (comm && (P % 32u) != 0)
    return fail(ctx, GPR_E_INVALID, "with a communicator n_pods must be a multiple of 32 (got %u)", P);
  if ((int)ctx->pending.size() >= kSlots)
    return fail(ctx, GPR_E_STATE, "too many outstanding async results; call gpr_sync");

  if (host_in) {
    // staging is dense, so only the number of cells matters: a window with one more GPU slot or a few more
    // samples than the shape given at gpr_create still fits as long as the product does
    if ((uint64_t)P * G * T > (uint64_t)ctx->max_pods * ctx->max_gpus * ctx->max_samples)
      return fail(ctx, GPR_E_CAPACITY, "host window %ux%ux%u exceeds the staging capacity of %ux%ux%u cells", P, G, T,
                  ctx->max_pods, ctx->max_gpus, ctx->max_samples);
    if (use_power && !ctx->cap_power)
      return fail(ctx, GPR_E_CAPACITY, "power plane not reserved (GPR_F_POWER_PLANE)");
  }

  // ---- scratch ---------------------------------------------------------------------------
  int rc;
  const unsigned sset = ctx->parity;  // scratch set of this call; successive calls alternate
  ctx->parity ^= 1u;
  if (ctx->masks_dirty) {  // a failed launch may also have left ticket / accumulators behind
    CU(cudaMemsetAsync(ctx->d_tickets, 0, 2 * sizeof(unsigned int), ctx->stream));
    CU(cudaMemsetAsync(ctx->d_acc, 0, 6 * sizeof(unsigned long long), ctx->stream));
    ctx->last_was_reduce = false;
  }
  for (int k = 0; k < 2; ++k) {
    const size_t cap_before = ctx->masks_cap[k];
    if ((rc = grow(ctx, &ctx->d_masks[k], &ctx->masks_cap[k], (size_t)2 * P * MW + 16)) != GPR_OK) return rc;
    if (ctx->masks_cap[k] != cap_before || ctx->masks_dirty) {
      CU(cudaMemsetAsync(ctx->d_masks[k], 0, ctx->masks_cap[k] * sizeof(uint32_t), ctx->stream));
      ctx->last_was_reduce = false;
    }
  }
  ctx->masks_dirty = false;
  uint32_t* const masks = ctx->d_masks[sset];
  if ((rc = grow(ctx, &ctx->d_bits, &ctx->bits_cap, (size_t)3 * W + 2)) != GPR_OK) return rc;
  if (comm && !fused &&
      (rc = grow(ctx, &ctx->d_gather, &ctx->gather_cap, (size_t)ctx->world * 2 * W + 2)) != GPR_OK)
    return rc;
  const bool want_smax = res->series_max != nullptr;
  if (want_smax && host_out &&
      (rc = grow(ctx, &ctx->d_smax, &ctx->smax_cap, (size_t)S + 4)) != GPR_OK)
    return rc;

  // ---- gates -----------------------------------------------------------------------------
  const uint8_t* d_elig = win->eligible;
  const int64_t* d_created = win->created_ts;
  if (gates_host && (win->eligible || win->created_ts)) {
    if (P > ctx->gate_cap) {
      if (ctx->d_elig_stage) {
        CU(cudaStreamSynchronize(ctx->stream));
        CU(cudaFree(ctx->d_elig_stage));
        CU(cudaFree(ctx->d_created_stage));
        ctx->d_elig_stage = nullptr, ctx->d_created_stage = nullptr, ctx->gate_cap = 0;
      }
      const size_t cap = (size_t)P + P / 4 + 64;
      CU(cudaMalloc(reinterpret_cast<void**>(&ctx->d_elig_stage), cap));
      CU(cudaMalloc(reinterpret_cast<void**>(&ctx->d_created_stage), cap * sizeof(int64_t)));
      ctx->gate_cap = cap;
    }
    ctx->last_was_reduce = false;
    if (win->eligible) {
      CU(cudaMemcpyAsync(ctx->d_elig_stage, win->eligible, P, cudaMemcpyHostToDevice, ctx->stream));
      d_elig = ctx->d_elig_stage;
    }
    if (win->created_ts) {
      CU(cudaMemcpyAsync(ctx->d_created_stage, win->created_ts, (size_t)P * sizeof(int64_t),
                         cudaMemcpyHostToDevice, ctx->stream));
      d_created = ctx->d_created_stage;
    }
  }

  // ---- output targets --------------------------------------------------------------------
  const bool direct_bits = !host_out && !comm;
  uint32_t* dbits_dev = direct_bits ? res->decision_bits : ctx->d_bits;
  uint32_t* cbits_dev = direct_bits ? res->candidate_bits
                                    : ((res->candidate_bits || comm) ? ctx->d_bits + W : nullptr);
  // pods vetoed by the power clause: never exchanged (this rank's pods only)
  uint32_t* vbits_dev = res->veto_bits ? (host_out ? ctx->d_bits + 2 * (size_t)W : res->veto_bits) : nullptr;
  uint32_t* my_gather = nullptr;  // local gather buffer of this call (fused exchange)
  const unsigned xset = (unsigned)((ctx->p2p_step + 1ull) % kExchangeDepth);  // exchange buffers of this step
  if (fused) {
    my_gather = reinterpret_cast<uint32_t*>(ctx->p2p_block + ctx->p2p_gather_off[xset]);
    dbits_dev = my_gather + (size_t)ctx->rank * ctx->p2p_stride;   // this rank's slot: [decision | candidate]
    cbits_dev = dbits_dev + W;
  }
  float* smax_dev = want_smax ? (host_out ? ctx->d_smax : res->series_max) : nullptr;

  gpr::FoldParams fp;
  fp.idle_mask = masks;
  fp.veto_mask = use_power ? masks + (size_t)P * MW : nullptr;
  fp.eligible = d_elig;
  fp.created = d_created;
  fp.cutoff = win->cutoff_ts;
  fp.dbits = dbits_dev;
  fp.cbits = cbits_dev;
  fp.vbits = vbits_dev;
  // single-launch path: the last CTA stores the three counters straight into this call's
  // pinned (device-mapped, UVA) host slot, so no copy operation separates back-to-back steps
  const int slot = (int)ctx->pending.size();
  unsigned long long* h_slot = ctx->h_counts + (size_t)slot * 8;
  fp.counts = h_slot;
  fp.stamp = h_slot + 3;
  fp.err = ctx->h_err;
  fp.acc = ctx->d_acc + 3 * sset;
  fp.ticket = ctx->d_tickets + sset;
  fp.done = ctx->d_done + sset;
  fp.need = ctx->uses[sset];
  fp.prev_done = ctx->d_done + (sset ^ 1u);
  fp.prev_need = ctx->uses[sset ^ 1u];
  fp.P = P;
  fp.G = G;
  fp.mw = MW;
  fp.world = 1, fp.rank = 0;
  fp.exchange_debug = 0;
  fp.poll_ns = ctx->poll_ns;
  fp.my_ll = nullptr;
  fp.late_order = 0;
  for (int r = 0; r < gpr::kMaxPeers; ++r) fp.peer_ll[r] = nullptr;
  if (fused) {
    fp.exchange_debug = ctx->exchange_debug;
    fp.world = ctx->world, fp.rank = ctx->rank;
    fp.rank_stride = ctx->p2p_stride;
    for (int r = 0; r < ctx->world; ++r) {
      fp.peer_gather[r] = reinterpret_cast<uint32_t*>(ctx->p2p_peer[r] + ctx->p2p_gather_off[xset]);
      fp.peer_flag[r] = reinterpret_cast<unsigned long long*>(ctx->p2p_peer[r]) + ctx->rank;
    }
    fp.my_flags = reinterpret_cast<const unsigned long long*>(ctx->p2p_block);
    if (ctx->exchange_ll) {
      for (int r = 0; r < ctx->world; ++r)
        fp.peer_ll[r] = reinterpret_cast<unsigned long long*>(ctx->p2p_peer[r] + ctx->p2p_ll_off[xset]);
      fp.my_ll = reinterpret_cast<unsigned long long*>(ctx->p2p_block + ctx->p2p_ll_off[xset]);
      // (veto bits go straight to the caller's buffer from every fold CTA, so that call keeps the early wait)
      fp.late_order = ctx->exchange_late && vbits_dev == nullptr ? 1 : 0;
    }
    fp.step = ++ctx->p2p_step;
    fp.out_dbits = host_out ? nullptr : res->decision_bits;
    fp.out_cbits = host_out ? nullptr : res->candidate_bits;
  }

  gpr::ReduceParams rp;
  memset(&rp, 0, sizeof rp);
  rp.T = T;
  rp.G = G;
  rp.mw = MW;
  rp.thr = threshold_f32(win->power_threshold);
  rp.done = fp.done;
  rp.need = fp.need;
  const bool can_pdl = ctx->pdl_enabled && ctx->own_stream;
  // the fold grid: 32 bitmap words per CTA and round; a handful of CTAs even at millions of pods
  // one bitmap word per warp and round, 4 words per warp in flight (fold_words<4>): small CTAs spread the fold's
  // loads over many SMs — each SM's path to L2 is busy with the next decision's reduce CTA
  const uint32_t fold_threads = (uint32_t)ctx->fold_threads, fold_warps = fold_threads / 32u;
  const uint32_t fold_grid = std::max<uint32_t>(1u, std::min<uint32_t>((W + 4u * fold_warps - 1u) / (4u * fold_warps), 148u));

  if (!async) {
    CU(cudaEventRecord(ctx->ev_k0, ctx->stream));
    ctx->last_was_reduce = false;
  }

  if (!host_in) {
    // ---- device-resident window: one reduce launch + the PDL-chained fold -------------------
    rp.ld = ld;
    rp.seg[0] = gpr::Segment{util, masks, smax_dev, S, 0u};
    rp.seg[1] = gpr::Segment{power, masks + (size_t)P * MW, nullptr, use_power ? S : 0u, 1u};
    rp.total_rows = S + (use_power ? S : 0u);
    rp.util_u8 = u8 ? 1u : 0u;
    const bool tma_ok = (T % 4u) == 0 && (ld % 4u) == 0 && aligned16(util) &&
                        (!use_power || aligned16(power));
    if (P > 0) {
      // The reduce grid may start while the previous decision's fold is still running, but only
      // when that is provably safe: our own stream (no foreign producer kernels between), the
      // newest op on it is one of our fold kernels, and this launch writes nothing but its own
      // scratch set (series_max would go straight to the caller's buffer).
      const bool pdl = can_pdl && ctx->last_was_reduce && !want_smax;
      if ((rc = launch_reduce(ctx, rp, tma_ok, pdl)) != GPR_OK) return rc;
      CU(fused ? launch_ex(gpr::k_fold<true>, fold_grid, fold_threads, 0, ctx->stream, can_pdl, fp)
               : launch_ex(gpr::k_fold<false>, fold_grid, fold_threads, 0, ctx->stream, can_pdl, fp));
      ctx->launches++;
      ctx->uses[sset]++;
      ctx->last_was_reduce = true;
    }
  } else {
    // ---- host window: pod chunks, H2D on the copy stream overlapped with the reduce --------
    ctx->last_was_reduce = false;
    CU(cudaEventRecord(ctx->ev_join, ctx->stream));
    CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_join, 0));
    const size_t usize = u8 ? 1u : 4u;  // bytes per util sample on the wire and in the staging plane
    const size_t pod_bytes = (size_t)G * T * (usize + (use_power ? 4u : 0u));
    uint32_t chunk_pods = (uint32_t)std::max<size_t>(1, ctx->chunk_bytes / std::max<size_t>(pod_bytes, 1));
    uint32_t n_chunks = P ? (P + chunk_pods - 1) / chunk_pods : 0;
    rp.ld = T;  // staging is dense
    rp.util_u8 = u8 ? 1u : 0u;
    const bool tma_ok = (T % 4u) == 0;
    const char* util_bytes = reinterpret_cast<const char*>(util);
    char* stage_bytes = reinterpret_cast<char*>(ctx->d_util_stage);
    for (uint32_t c = 0; c < n_chunks; ++c) {
      const uint32_t p0 = c * chunk_pods, p1 = std::min(P, p0 + chunk_pods);
      const size_t row0 = (size_t)p0 * G, n_rows = (size_t)(p1 - p0) * G;
      float* du = reinterpret_cast<float*>(stage_bytes + row0 * T * usize);
      if ((rc = copy_rows_h2d(ctx, du, util_bytes + row0 * ld * usize, n_rows, T, ld, usize,
                              ctx->copy_stream)) != GPR_OK)
        return rc;
      float* dp = nullptr;
      if (use_power) {
        dp = ctx->d_power_stage + row0 * T;
        if ((rc = copy_rows_h2d(ctx, dp, power + row0 * ld, n_rows, T, ld, 4u, ctx->copy_stream)) !=
            GPR_OK)
          return rc;
      }
      cudaEvent_t ev = ctx->ev_chunk[c % kMaxChunkEvents];
      CU(cudaEventRecord(ev, ctx->copy_stream));
      CU(cudaStreamWaitEvent(ctx->stream, ev, 0));
      rp.seg[0] = gpr::Segment{du, masks + (size_t)p0 * MW, smax_dev ? smax_dev + row0 : nullptr,
                               (uint32_t)n_rows, 0u};
      rp.seg[1] = gpr::Segment{dp, masks + ((size_t)P + p0) * MW, nullptr,
                               use_power ? (uint32_t)n_rows : 0u, 1u};
      rp.total_rows = (uint32_t)n_rows * (use_power ? 2u : 1u);
      if ((rc = launch_reduce(ctx, rp, tma_ok, false)) != GPR_OK) return rc;
    }
    if (P > 0) {
      CU(fused ? launch_ex(gpr::k_fold<true>, fold_grid, fold_threads, 0, ctx->stream, false, fp)
               : launch_ex(gpr::k_fold<false>, fold_grid, fold_threads, 0, ctx->stream, false, fp));
      ctx->launches++;
      ctx->uses[sset]++;
    }
  }
  if (P == 0) memset(h_slot, 0, 8 * sizeof(unsigned long long));  // slot is not in flight

  // ---- the one collective: allgather of the packed bitmap over NVLink ----------------------
  if ((comm && !fused) || host_out || !async) ctx->last_was_reduce = false;  // something follows
  if (comm && !fused && W > 0) {
    NC(g_nccl.AllGather(ctx->d_bits, ctx->d_gather, (size_t)2 * W, ncclUint32, ctx->comm,
                        ctx->stream));
  }
  if (!async) CU(cudaEventRecord(ctx->ev_k1, ctx->stream));

  // ---- deliver -----------------------------------------------------------------------------
  const cudaMemcpyKind out_kind = host_out ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  if (W > 0) {
    if (fused) {
      if (host_out) {  // device outputs were assembled by the folding CTA itself
        const size_t pitch = (size_t)ctx->p2p_stride * 4u;
        CU(cudaMemcpy2DAsync(res->decision_bits, (size_t)W * 4u, my_gather, pitch, (size_t)W * 4u,
                             ctx->world, out_kind, ctx->stream));
        if (res->candidate_bits)
          CU(cudaMemcpy2DAsync(res->candidate_bits, (size_t)W * 4u, my_gather + W, pitch,
                               (size_t)W * 4u, ctx->world, out_kind, ctx->stream));
      }
    } else if (comm) {
      CU(cudaMemcpy2DAsync(res->decision_bits, (size_t)W * 4u, ctx->d_gather, (size_t)2 * W * 4u,
                           (size_t)W * 4u, ctx->world, out_kind, ctx->stream));
      if (res->candidate_bits)
        CU(cudaMemcpy2DAsync(res->candidate_bits, (size_t)W * 4u, ctx->d_gather + W,
                             (size_t)2 * W * 4u, (size_t)W * 4u, ctx->world, out_kind, ctx->stream));
    } else if (host_out) {
      CU(cudaMemcpyAsync(res->decision_bits, ctx->d_bits, (size_t)W * 4u, out_kind, ctx->stream));
      if (res->candidate_bits)
        CU(cudaMemcpyAsync(res->candidate_bits, ctx->d_bits + W, (size_t)W * 4u, out_kind,
                           ctx->stream));
    }
  }
  if (res->veto_bits && host_out && W > 0)
    CU(cudaMemcpyAsync(res->veto_bits, ctx->d_bits + 2 * (size_t)W, (size_t)W * 4u, cudaMemcpyDeviceToHost, ctx->stream));
  if (want_smax && host_out && S > 0)
    CU(cudaMemcpyAsync(res->series_max, ctx->d_smax, (size_t)S * 4u, cudaMemcpyDeviceToHost,
                       ctx->stream));
  ctx->pending.push_back(Pending{res, slot});
  res->kernel_ms = 0.0;
  return GPR_OK;
}

int sync_impl(gpr_ctx* ctx) {
  CU(cudaSetDevice(ctx->device));
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) {
    ctx->pending.clear();
    ctx->masks_dirty = true;
    return fail(ctx, GPR_E_CUDA, "cudaStreamSynchronize: %s", cudaGetErrorString(e));
  }
  ctx->stamps.clear();
  ctx->phase_stamps.clear();
  for (const Pending& p : ctx->pending) {
    const unsigned long long* c = ctx->h_counts + (size_t)p.slot * 8;
    p.res->n_series = c[0];
    p.res->n_candidates = c[1];
    p.res->n_decisions = c[2];
    ctx->stamps.push_back(c[3]);
    for (int k = 4; k < 8; ++k) ctx->phase_stamps.push_back(c[k]);
  }
  ctx->pending.clear();
  if (*ctx->h_err) {
    *ctx->h_err = 0;
    ctx->masks_dirty = true;
    return fail(ctx, GPR_E_STATE, "a peer rank never arrived at the bitmap exchange / rendezvous (waited %llu s); "
                "the results of this batch are not global", gpr::kPeerTimeoutNs / 1000000000ull);
  }
  return GPR_OK;
}

constexpr uint32_t kIdxBlock = 64;

// NaN-skipping max of block b (samples [64 b, 64 b + 64) of one resident row), one warp
__device__ __forceinline__ float block_max_warp(const float* row, uint32_t T, uint32_t b, int lane) {
  const uint32_t t0 = b * kIdxBlock;
  float m = gpr::nan_f();
  for (uint32_t t = t0 + lane; t < min(T, t0 + kIdxBlock); t += 32) m = fmaxf(m, row[t]);
  return gpr::warp_max(m);
}

// Scatter n_new columns of every row into the time ring; with an index, recompute the maxima of the
// blocks the new columns landed in (the overwritten samples may have been the old maximum).
__global__ void __launch_bounds__(128) k_append(float* __restrict__ dst, const float* __restrict__ src,
                                               uint32_t n_rows, uint32_t T, uint32_t head, uint32_t n_new,
                                               uint64_t ld_src, float* __restrict__ idx, uint32_t idx_ld) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  for (uint32_t r = blockIdx.x; r < n_rows; r += gridDim.x) {
    float* out = dst + (size_t)r * T;
    const float* in = src + (size_t)r * ld_src;
    for (uint32_t j = threadIdx.x; j < n_new; j += blockDim.x) {
      uint32_t t = head + j;
      if (t >= T) t -= T;
      out[t] = in[j];
    }
    if (idx) {
      __syncthreads();  // this CTA's column stores are visible to its own warps
      // touched blocks: those of [head, head + n_new) modulo T — at most two contiguous runs
      const uint32_t n_blocks = (T + kIdxBlock - 1) / kIdxBlock;
      const uint32_t first = head / kIdxBlock;
      const uint32_t span_end = head + n_new;  // exclusive, may exceed T (wraps)
      const uint32_t last = (min(span_end, T) - 1) / kIdxBlock;
      for (uint32_t b = first + warp; b <= last; b += n_warps) {
        const float m = block_max_warp(out, T, b, lane);
        if (lane == 0) idx[(size_t)r * idx_ld + b] = m;
      }
      if (span_end > T) {  // wrapped part [0, span_end - T)
        const uint32_t wlast = min((span_end - T - 1) / kIdxBlock, n_blocks - 1);
        for (uint32_t b = warp; b <= wlast; b += n_warps) {
          const float m = block_max_warp(out, T, b, lane);
          if (lane == 0) idx[(size_t)r * idx_ld + b] = m;
        }
      }
      __syncthreads();
    }
  }
}

// full rebuild of the index (after the caller wrote the resident planes directly)
__global__ void __launch_bounds__(128) k_reindex(const float* __restrict__ plane, uint32_t n_rows, uint32_t T,
                                                float* __restrict__ idx, uint32_t idx_ld) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  const uint32_t n_blocks = (T + kIdxBlock - 1) / kIdxBlock;
  for (uint32_t r = blockIdx.x; r < n_rows; r += gridDim.x)
    for (uint32_t b = warp; b < n_blocks; b += n_warps) {
      const float m = block_max_warp(plane + (size_t)r * T, T, b, lane);
      if (lane == 0) idx[(size_t)r * idx_ld + b] = m;
    }
}

}  // namespace

// ---- upload + scan pipeline ------------------------------------------------------------------------------------
// The text is cut into chunks.  A few producer threads each take every nt-th chunk: copy it into their pinned
// double buffer (pageable sources; a plain cudaMemcpy would go through the driver's single bounce buffer at
// ~10 GB/s), enqueue the H2D copy on their own stream, and right behind it the scan kernel for that chunk, which
// appends the chunk's markers to a block of mapped pinned memory.  A chunk is copied with 16 bytes of overlap
// into the next one (identical bytes written twice), so its scan never needs another stream's data.  The
// consumer (gpr_text_scan_next) takes the chunks in text order while later ones are still in flight.
struct ScanPipe {
  int slot = 0;
  const char* src = nullptr;
  uint8_t* dst = nullptr;
  uint64_t n = 0, n_chunks = 0, chunk = 0;
  int nt = 1;
  int src_kind = GPR_MEM_HOST;
  bool staged = false;  // source is pageable host memory: goes through the pinned ring
  std::vector<std::thread> th;
  std::atomic<uint64_t> consumed{0};                      // chunks handed to the caller
  std::atomic<uint64_t> recorded[gpr_ctx::kMarkBlocks];   // chunk + 1 whose event has been recorded in this block
  std::atomic<int> error{0};                              // first cudaError_t of a producer
  std::atomic<bool> stop{false};
  uint64_t next = 0;                                      // next chunk the consumer returns
};

namespace {

constexpr uint32_t kBlockWords = 2 + 2 * gpr_ctx::kMarkCap;

// markers of one chunk, 32-bit offsets relative to the chunk: [n_open, n_close, opens[cap], closes[cap]]
struct ChunkSink {
  uint32_t* block;
  uint64_t base;
  __device__ __forceinline__ void values_open(uint64_t p) {
    const uint32_t i = atomicAdd(block + 0, 1u);
    if (i < gpr_ctx::kMarkCap) block[2 + i] = (uint32_t)(p - base);
  }
  __device__ __forceinline__ void values_close(uint64_t p) {
    const uint32_t i = atomicAdd(block + 1, 1u);
    if (i < gpr_ctx::kMarkCap) block[2 + gpr_ctx::kMarkCap + i] = (uint32_t)(p - base);
  }
};

// copies a chunk's markers from the device block to the host-mapped one with plain coalesced stores (the scan's
// atomic appends must not go to host memory: an atomic across PCIe costs microseconds)
__global__ void __launch_bounds__(256) k_publish_marks(const uint32_t* __restrict__ d_block, uint32_t* __restrict__ h_block) {
  const uint32_t no = min(d_block[0], gpr_ctx::kMarkCap), nc = min(d_block[1], gpr_ctx::kMarkCap);
  for (uint32_t i = threadIdx.x; i < no; i += blockDim.x) h_block[2 + i] = d_block[2 + i];
  for (uint32_t i = threadIdx.x; i < nc; i += blockDim.x)
    h_block[2 + gpr_ctx::kMarkCap + i] = d_block[2 + gpr_ctx::kMarkCap + i];
  if (threadIdx.x == 0) h_block[0] = d_block[0], h_block[1] = d_block[1];
}

__global__ void __launch_bounds__(256) k_text_scan_chunk(const uint8_t* __restrict__ t, uint64_t n, uint64_t slice_begin,
                                                         uint64_t slice_end, uint32_t* block) {
  ChunkSink sink{block, slice_begin * gpr::text::kScanBytes};
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t slice = slice_begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; slice < slice_end; slice += stride)
    gpr::text::scan_slice(t, n, slice, sink);
}

void scan_producer(gpr_ctx* ctx, ScanPipe* sp, int k) {
  constexpr int NS = gpr_ctx::kUpSlots, NB = gpr_ctx::kMarkBlocks;
  const size_t SLOT = ctx->up_slot_bytes;
  cudaError_t e = cudaSetDevice(ctx->device);
  bool used[NS] = {};
  int slot = 0;
  cudaStream_t st = ctx->up_stream[k];
  for (uint64_t c = (uint64_t)k; c < sp->n_chunks && e == cudaSuccess && !sp->stop.load(); c += (uint64_t)sp->nt) {
    // the marker block of this chunk is free once the consumer has taken chunk c - NB
    for (int spins = 0; c >= sp->consumed.load(std::memory_order_acquire) + NB && !sp->stop.load(); ++spins) {
      if (spins < 64) std::this_thread::yield();
      else std::this_thread::sleep_for(std::chrono::microseconds(50));  // a slow consumer: do not burn the core
    }
    if (sp->stop.load()) break;
    const uint64_t off = c * sp->chunk;
    const uint64_t len = std::min<uint64_t>(sp->chunk, sp->n - off);
    const uint64_t len_ov = std::min<uint64_t>(len + 16, sp->n - off);  // overlap into the next chunk
    const void* from = sp->src + off;
    if (sp->staged) {
      unsigned char* buf = ctx->h_up_ring + ((size_t)k * NS + slot) * SLOT;
      if (used[slot]) e = cudaEventSynchronize(ctx->up_event[k][slot]);  // its previous DMA has drained
      if (e != cudaSuccess) break;
      memcpy(buf, from, len_ov);
      from = buf;
    }
    e = cudaMemcpyAsync(sp->dst + off, from, len_ov,
                        sp->src_kind == GPR_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && sp->staged) e = cudaEventRecord(ctx->up_event[k][slot], st), used[slot] = true;
    slot = (slot + 1) % NS;
    if (e != cudaSuccess) break;
    // (the kernels that last used block c % NB have completed: their chunk was consumed)
    uint32_t* h_block = ctx->h_mark_blocks + (size_t)(c % NB) * kBlockWords;
    uint32_t* d_block = ctx->d_mark_blocks + (size_t)(c % NB) * kBlockWords;
    e = cudaMemsetAsync(d_block, 0, 2 * sizeof(uint32_t), st);
    if (e != cudaSuccess) break;
    const uint64_t s0 = off / gpr::text::kScanBytes;
    const uint64_t s1 = (off + len + gpr::text::kScanBytes - 1) / gpr::text::kScanBytes;
    const uint32_t grid = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>((s1 - s0 + 255) / 256, (uint64_t)ctx->sm_count * 8));
    k_text_scan_chunk<<<grid, 256, 0, st>>>(sp->dst, sp->n, s0, s1, d_block);
    k_publish_marks<<<1, 256, 0, st>>>(d_block, h_block);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaEventRecord(ctx->mark_event[c % NB], st);
    if (e != cudaSuccess) break;
    sp->recorded[c % NB].store(c + 1, std::memory_order_release);
  }
  if (e != cudaSuccess) {
    int zero = 0;
    sp->error.compare_exchange_strong(zero, (int)e);
    sp->stop.store(true);
  }
}

int scan_pipe_finish(gpr_ctx* ctx, bool ok) {
  ScanPipe* sp = ctx->pipe;
  if (!sp) return GPR_OK;
  if (!ok) sp->stop.store(true);
  for (std::thread& t : sp->th) t.join();
  cudaError_t e = (cudaError_t)sp->error.load();
  // work queued on the context's stream afterwards (the parse kernels) must see the whole text
  for (int k = 0; k < sp->nt && e == cudaSuccess; ++k) {
    e = cudaEventRecord(ctx->up_event[k][0], ctx->up_stream[k]);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, ctx->up_event[k][0], 0);
  }
  if (!ok)
    for (int k = 0; k < sp->nt; ++k) cudaStreamSynchronize(ctx->up_stream[k]);
  delete sp;
  ctx->pipe = nullptr;
  if (e != cudaSuccess) return fail(ctx, GPR_E_CUDA, "text upload / scan: %s", cudaGetErrorString(e));
  return GPR_OK;
}

}  // namespace

void scan_pipe_abort(gpr_ctx* ctx) {
  if (ctx && ctx->pipe) (void)scan_pipe_finish(ctx, false);
}


// ===========================================================================================
// extern "C"
// ===========================================================================================
#define GPR_TRY try {
#define GPR_CATCH(ctxp)                                                     \
  }                                                                         \
  catch (const std::bad_alloc&) {                                           \
    return fail(ctxp, GPR_E_NOMEM, "host allocation failed");               \
  }                                                                         \
  catch (...) {                                                             \
    return fail(ctxp, GPR_E_INVALID, "unexpected C++ exception");           \
  }

extern "C" {

int gpr_version(void) {
  return GPR_VERSION_MAJOR * 10000 + GPR_VERSION_MINOR * 100 + GPR_VERSION_PATCH;
}

const char* gpr_last_error(const gpr_ctx* ctx) { return ctx ? ctx->err : g_create_err; }

void gpr_destroy(gpr_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->comm && g_nccl.ok) g_nccl.CommDestroy(ctx->comm);
  for (int r = 0; r < gpr::kMaxPeers; ++r)
    if (ctx->p2p_peer[r] && ctx->p2p_peer[r] != ctx->p2p_block) cudaIpcCloseMemHandle(ctx->p2p_peer[r]);
  if (ctx->p2p_block) cudaFree(ctx->p2p_block);
  void* dev[] = {ctx->d_util_stage, ctx->d_power_stage, ctx->d_elig_stage, ctx->d_created_stage,
                 ctx->d_masks[0],   ctx->d_masks[1],    ctx->d_bits,       ctx->d_gather,
                 ctx->d_smax,       ctx->d_acc,         ctx->d_tickets,    ctx->d_done,
                 ctx->d_flush,      ctx->d_res_util,    ctx->d_res_power,  ctx->d_cols,
                 ctx->d_idx_util,   ctx->d_idx_power,   ctx->d_text[0],    ctx->d_text[1],
                 ctx->d_text[2],    ctx->d_marks,       ctx->d_mark_counts, ctx->d_spans,
                 ctx->d_tplane[0],  ctx->d_tplane[1]};
  for (void* p : dev)
    if (p) cudaFree(p);
  if (ctx->h_counts) cudaFreeHost(ctx->h_counts);
  scan_pipe_abort(ctx);
  if (ctx->h_up_ring) cudaFreeHost(ctx->h_up_ring);
  if (ctx->h_mark_blocks) cudaFreeHost(ctx->h_mark_blocks);
  if (ctx->d_mark_blocks) cudaFree(ctx->d_mark_blocks);
  for (cudaEvent_t e : ctx->mark_event)
    if (e) cudaEventDestroy(e);
  for (int k = 0; k < gpr_ctx::kUpThreads; ++k) {
    if (ctx->up_stream[k]) cudaStreamDestroy(ctx->up_stream[k]);
    for (cudaEvent_t e : ctx->up_event[k])
      if (e) cudaEventDestroy(e);
  }
  cudaEvent_t evs[] = {ctx->ev_k0, ctx->ev_k1, ctx->ev_t0, ctx->ev_t1, ctx->ev_join};
  for (cudaEvent_t e : evs)
    if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : ctx->ev_chunk)
    if (e) cudaEventDestroy(e);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int gpr_create(const gpr_config* cfg, gpr_ctx** out) {
  gpr_ctx* ctx = nullptr;  // errors before allocation go to the global slot
  GPR_TRY
  if (!cfg || !out) return fail(nullptr, GPR_E_INVALID, "config/out is NULL");
  *out = nullptr;
  if (cfg->struct_size != sizeof(gpr_config))
    return fail(nullptr, GPR_E_INVALID, "gpr_config.struct_size %u != %zu", cfg->struct_size,
                sizeof(gpr_config));
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0)
    return fail(nullptr, GPR_E_CUDA, "no CUDA device (%s); this engine has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  if (cfg->device < 0 || cfg->device >= n_dev)
    return fail(nullptr, GPR_E_INVALID, "device %d out of range [0,%d)", cfg->device, n_dev);
  gpr_ctx* c = new gpr_ctx();
  c->device = cfg->device;
  // from here on failures are reported through the global slot too, then the ctx is freed
  auto bail = [&](int code) {
    {
      std::lock_guard<std::mutex> lk(g_err_mu);
      snprintf(g_create_err, sizeof g_create_err, "%s", c->err);
    }
    gpr_destroy(c);
    return code;
  };
  ctx = c;
  auto body = [&]() -> int {
    CU(cudaSetDevice(c->device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, c->device));
    c->sm_count = prop.multiProcessorCount;
    c->l2_bytes = (size_t)prop.l2CacheSize;
    c->hbm_bytes = prop.totalGlobalMem;
    c->cc_major = prop.major, c->cc_minor = prop.minor;
    memcpy(c->name, prop.name, sizeof c->name - 1);
    if (prop.major < 10)
      return fail(c, GPR_E_UNSUPPORTED, "device %s is sm_%d%d; this library is built for sm_100a",
                  prop.name, prop.major, prop.minor);
    if (cfg->stream) {
      c->stream = static_cast<cudaStream_t>(cfg->stream);
    } else {
      CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
      c->own_stream = true;
    }
    CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CU(cudaEventCreate(&c->ev_k0));
    CU(cudaEventCreate(&c->ev_k1));
    CU(cudaEventCreate(&c->ev_t0));
    CU(cudaEventCreate(&c->ev_t1));
    CU(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    for (cudaEvent_t& ev : c->ev_chunk) CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CU(cudaMalloc(reinterpret_cast<void**>(&c->d_acc), 6 * sizeof(unsigned long long)));
    CU(cudaMalloc(reinterpret_cast<void**>(&c->d_tickets), 2 * sizeof(unsigned int)));
    CU(cudaMalloc(reinterpret_cast<void**>(&c->d_done), 2 * sizeof(unsigned long long)));
    CU(cudaMemset(c->d_acc, 0, 6 * sizeof(unsigned long long)));
    CU(cudaMemset(c->d_tickets, 0, 2 * sizeof(unsigned int)));
    CU(cudaMemset(c->d_done, 0, 2 * sizeof(unsigned long long)));
    c->pdl_enabled = env_int("GPR_PDL", 1) != 0;
    c->up_threads = std::max(1, std::min((int)gpr_ctx::kUpThreads, env_int("GPR_TEXT_UPLOAD_THREADS", 8)));
    c->up_chunk = (size_t)std::max(1, std::min(16, env_int("GPR_TEXT_CHUNK_MB", 2))) << 20;
    c->up_slot_bytes = c->up_chunk + 4096;
    c->exchange_debug = env_int("GPR_DEBUG_EXCHANGE", 0);
    c->poll_ns = (unsigned int)std::max(100, std::min(100000, env_int("GPR_POLL_NS", 4000)));
    if (const char* x = getenv("GPR_EXCHANGE")) {
      c->exchange_ll = strcmp(x, "flags") != 0;
      c->exchange_late = strcmp(x, "pipelined") == 0;
    }
    c->fold_threads = env_int("GPR_FOLD_THREADS", 256);
    if (c->fold_threads != 64 && c->fold_threads != 128 && c->fold_threads != 256) c->fold_threads = 256;
    CU(cudaMallocHost(reinterpret_cast<void**>(&c->h_counts),
                      ((size_t)kSlots * 8 + 2) * sizeof(unsigned long long)));
    memset(c->h_counts, 0, ((size_t)kSlots * 8 + 2) * sizeof(unsigned long long));
    c->h_mark = c->h_counts + (size_t)kSlots * 8;
    c->h_err = reinterpret_cast<unsigned int*>(c->h_mark + 1);
    c->pending.reserve(kSlots);
    c->stamps.reserve(kSlots);

    c->variant = cfg->kernel_variant;
    if (const char* k = getenv("GPR_KERNEL")) {
      if (!strcmp(k, "ldg")) c->variant = GPR_KERNEL_LDG;
      else if (!strcmp(k, "tma")) c->variant = GPR_KERNEL_TMA;
    }
    if (c->variant < GPR_KERNEL_AUTO || c->variant > GPR_KERNEL_TMA)
      return fail(c, GPR_E_INVALID, "bad kernel_variant %d", c->variant);
    c->ldg_ctas_per_sm = std::max(1, env_int("GPR_LDG_CTAS", 2));
    c->tma_depth_max = std::max(1, env_int("GPR_TMA_DEPTH", 3));
    c->tma_warps = env_int("GPR_TMA_WARPS", 16);
    if (c->tma_warps != 4 && c->tma_warps != 8 && c->tma_warps != 16 && c->tma_warps != 32)
      c->tma_warps = 16;
    c->tma_chunk_bytes = std::min(65536, std::max(512, env_int("GPR_TMA_CHUNK", 8192))) & ~15;
    c->chunk_bytes = (size_t)std::max(1, env_int("GPR_CHUNK_MB", 32)) << 20;  // sweep: profiles/README.md
    c->parse_ctas_per_sm = std::max(1, std::min(8, env_int("GPR_PARSE_CTAS", 6)));
    CU(cudaFuncSetAttribute(gpr::k_reduce_tma<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)kTmaSmemBudget));
    CU(cudaFuncSetAttribute(gpr::k_reduce_tma<8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)kTmaSmemBudget));
    CU(cudaFuncSetAttribute(gpr::k_reduce_tma<16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)kTmaSmemBudget));
    CU(cudaFuncSetAttribute(gpr::k_reduce_tma<32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)kTmaSmemBudget));

    c->max_pods = cfg->max_pods, c->max_gpus = cfg->max_gpus, c->max_samples = cfg->max_samples;
    c->cap_power = (cfg->flags & GPR_F_POWER_PLANE) != 0;
    const size_t cells = (size_t)cfg->max_pods * cfg->max_gpus * cfg->max_samples;
    if (cells) {
      CU(cudaMalloc(reinterpret_cast<void**>(&c->d_util_stage), cells * sizeof(float) + 256));
      if (c->cap_power)
        CU(cudaMalloc(reinterpret_cast<void**>(&c->d_power_stage), cells * sizeof(float) + 256));
    }
    return GPR_OK;
  };
  int rc = body();
  if (rc != GPR_OK) return bail(rc);
  *out = c;
  return GPR_OK;
  GPR_CATCH(nullptr)
}

int gpr_decide_async(gpr_ctx* ctx, const gpr_window* win, gpr_result* res) {
  GPR_TRY
  const int rc = decide_impl(ctx, win, res, false, true);
  if (rc != GPR_OK && ctx) ctx->masks_dirty = true;
  return rc;
  GPR_CATCH(ctx)
}

int gpr_decide_batch_async(gpr_ctx* ctx, const gpr_window* wins, gpr_result* results, uint32_t n) {
  if (!ctx) return GPR_E_INVALID;
  GPR_TRY
  if (n && (!wins || !results)) return fail(ctx, GPR_E_INVALID, "windows/results is NULL");
  for (uint32_t i = 0; i < n; ++i) {
    const int rc = decide_impl(ctx, &wins[i], &results[i], false, true);
    if (rc != GPR_OK) {
      ctx->masks_dirty = true;
      return rc;
    }
  }
  return GPR_OK;
  GPR_CATCH(ctx)
}

int gpr_sync(gpr_ctx* ctx) {
  if (!ctx) return GPR_E_INVALID;
  GPR_TRY
  return sync_impl(ctx);
  GPR_CATCH(ctx)
}

static int decide_blocking(gpr_ctx* ctx, const gpr_window* win, gpr_result* res, bool resident) {
  int rc = decide_impl(ctx, win, res, resident, false);
  if (rc != GPR_OK) {
    if (ctx) {
      cudaStreamSynchronize(ctx->stream);
      ctx->pending.clear();
      ctx->masks_dirty = true;
    }
    return rc;
  }
  rc = sync_impl(ctx);
  if (rc != GPR_OK) return rc;
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, ctx->ev_k0, ctx->ev_k1));
  res->kernel_ms = ms;
  return GPR_OK;
}

int gpr_decide(gpr_ctx* ctx, const gpr_window* win, gpr_result* res) {
  GPR_TRY
  return decide_blocking(ctx, win, res, false);
  GPR_CATCH(ctx)
}

int gpr_decide_resident(gpr_ctx* ctx, const gpr_window* win, gpr_result* res) {
  GPR_TRY
  return decide_blocking(ctx, win, res, true);
  GPR_CATCH(ctx)
}

// ---- resident window -----------------------------------------------------------------------
int gpr_resident_init(gpr_ctx* ctx, uint32_t P, uint32_t G, uint32_t T, uint32_t flags) {
  if (!ctx) return GPR_E_INVALID;
  GPR_TRY
  if (P == 0 || G == 0 || T == 0) return fail(ctx, GPR_E_INVALID, "empty resident window");
  if ((uint64_t)P * G > 0x7fffffffull) return fail(ctx, GPR_E_INVALID, "too many series");
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  if (ctx->d_res_util) CU(cudaFree(ctx->d_res_util));
  if (ctx->d_res_power) CU(cudaFree(ctx->d_res_power));
  ctx->d_res_util = ctx->d_res_power = nullptr;
  const size_t bytes = (size_t)P * G * T * sizeof(float);
  CU(cudaMalloc(reinterpret_cast<void**>(&ctx->d_res_util), bytes));
  // 0xFFFFFFFF is a NaN: every step starts out "no sample"
  CU(cudaMemsetAsync(ctx->d_res_util, 0xFF, bytes, ctx->stream));
  if (flags & GPR_F_POWER_PLANE) {
    CU(cudaMalloc(reinterpret_cast<void**>(&ctx->d_res_power), bytes));
    CU(cudaMemsetAsync(ctx->d_res_power, 0xFF, bytes, ctx->stream));
  }
  if (ctx->d_idx_util) CU(cudaFree(ctx->d_idx_util));
  if (ctx->d_idx_power) CU(cudaFree(ctx->d_idx_power));
  ctx->d_idx_util = ctx->d_idx_power = nullptr, ctx->idx_ld = 0;
  if (flags & GPR_F_BLOCK_INDEX) {
    ctx->idx_ld = (((T + kIdxBlock - 1) / kIdxBlock) + 3u) & ~3u;
    const size_t ib = (size_t)P * G * ctx->idx_ld * sizeof(float);
    CU(cudaMalloc(reinterpret_cast<void**>(&ctx->d_idx_util), ib));
    CU(cudaMemsetAsync(ctx->d_idx_util, 0xFF, ib, ctx->stream));
    if (flags & GPR_F_POWER_PLANE) {
      CU(cudaMalloc(reinterpret_cast<void**>(&ctx->d_idx_power), ib));
      CU(cudaMemsetAsync(ctx->d_idx_power, 0xFF, ib, ctx->stream));
    }
  }
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->res_P = P, ctx->res_G = G, ctx->res_T = T, ctx->res_head = 0;
  return GPR_OK;
  GPR_CATCH(ctx)
}

int gpr_resident_reindex(gpr_ctx* ctx) {
  if (!ctx) return GPR_E_INVALID;
  GPR_TRY
  if (!ctx->d_res_util) return fail(ctx, GPR_E_STATE, "no resident window (gpr_resident_init)");
  if (!ctx->d_idx_util) return GPR_OK;  // no index to maintain
  CU(cudaSetDevice(ctx->device));
  ctx->last_was_reduce = false;
  const size_t rows = (size_t)ctx->res_P * ctx->res_G;
  const uint32_t grid = (uint32_t)std::min<size_t>(rows, (size_t)ctx->sm_count * 16);
  k_reindex<<<grid, 128, 0, ctx->stream>>>(ctx->d_res_util, (uint32_t)rows, ctx->res_T, ctx->d_idx_util,
                                           ctx->idx_ld);
  if (ctx->d_res_power)
    k_reindex<<<grid, 128, 0, ctx->stream>>>(ctx->d_res_power, (uint32_t)rows, ctx->res_T, ctx->d_idx_power,
                                             ctx->idx_ld);
  ctx->launches += ctx->d_res_power ? 2 : 1;
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(ctx->stream));
  return GPR_OK;
  GPR_CATCH(ctx)
}

int gpr_append(gpr_ctx* ctx, const float* util_cols, const float* power_cols, uint32_t n_new,
               uint64_t row_stride, int32_t mem_kind) {
  if (!ctx) return GPR_E_INVALID;
  GPR_TRY
  if (!ctx->d_res_util) return fail(ctx, GPR_E_STATE, "no resident window (gpr_resident_init)");
  if (n_new == 0) return GPR_OK;
  if (!util_cols) return fail(ctx, GPR_E_INVALID, "util_cols is NULL");
  ctx->last_was_reduce = false;
  if (mem_kind != GPR_MEM_HOST && mem_kind != GPR_MEM_DEVICE)
    return fail(ctx, GPR_E_INVALID, "bad mem_kind %d", mem_kind);
  CU(cudaSetDevice(ctx->device));
  const uint32_t T = ctx->res_T;
  const size_t rows = (size_t)ctx->res_P * ctx->res_G;
  uint64_t ld = row_stride ? row_stride : n_new;
  if (ld < n_new) return fail(ctx, GPR_E_INVALID, "row_stride < n_new");
  // only the newest T columns can survive in a ring of T
  uint32_t skip = n_new > T ? n_new - T : 0;
  const uint32_t n_eff = n_new - skip;
  const float* planes_in[2] = {util_cols, ctx->d_res_power ? power_cols : nullptr};
  float* planes_out[2] = {ctx->d_res_util, ctx->d_res_power};
  float* planes_idx[2] = {ctx->d_idx_util, ctx->d_idx_power};
  for (int pl = 0; pl < 2; ++pl) {
    if (!planes_in[pl]) continue;
    const float* src = planes_in[pl] + skip;
    uint64_t ld_dev = ld;
    if (mem_kind == GPR_MEM_HOST) {
      int rc = grow(ctx, &ctx->d_cols, &ctx->cols_cap, rows * n_eff + 4);
      if (rc != GPR_OK) return rc;
      if (ld == n_eff)  // dense block: one linear copy (a 2-D copy of 720-byte rows runs at ~6 GB/s)
        CU(cudaMemcpyAsync(ctx->d_cols, src, rows * (size_t)n_eff * 4u, cudaMemcpyHostToDevice,
                           ctx->stream));
      else
        CU(cudaMemcpy2DAsync(ctx->d_cols, (size_t)n_eff * 4u, src, (size_t)ld * 4u,
                             (size_t)n_eff * 4u, rows, cudaMemcpyHostToDevice, ctx->stream));
      src = ctx->d_cols;
      ld_dev = n_eff;
    }
    const uint32_t grid = (uint32_t)std::min<size_t>(rows, (size_t)ctx->sm_count * 16);
    k_append<<<grid, 128, 0, ctx->stream>>>(planes_out[pl], src, (uint32_t)rows, T,
                                            (ctx->res_head + skip) % T, n_eff, ld_dev, planes_idx[pl],
                                            ctx->idx_ld);
    ctx->launches++;
    CU(cudaGetLastError());
  }
  ctx->res_head = (uint32_t)(((uint64_t)ctx->res_head + n_new) % T);
  CU(cudaStreamSynchronize(ctx->stream));
  return GPR_OK;
  GPR_CATCH(ctx)
}

int gpr_resident_planes(gpr_ctx* ctx, float** util, float** power, uint64_t* row_stride) {
  if (!ctx) return GPR_E_INVALID;
  if (!ctx->d_res_util) return fail(ctx, GPR_E_STATE, "no resident window");
  if (util) *util = ctx->d_res_util;
  if (power) *power = ctx->d_res_power;
  if (row_stride) *row_stride = ctx->res_T;
  return GPR_OK;
}

int gpr_resident_head(gpr_ctx* ctx, uint32_t* head) {
  if (!ctx || !head) return GPR_E_INVALID;
  if (!ctx->d_res_util) return fail(ctx, GPR_E_STATE, "no resident window");
  *head = ctx->res_head;
  return GPR_OK;
}

int gpr_resident_advance(gpr_ctx* ctx, uint32_t n_new) {
  if (!ctx) return GPR_E_INVALID;
  GPR_TRY
  if (!ctx->d_res_util) return fail(ctx, GPR_E_STATE, "no resident window (gpr_resident_init)");
  if (n_new == 0) return GPR_OK;
  CU(cudaSetDevice(ctx->device));
  ctx->last_was_reduce = false;
  const uint32_t T = ctx->res_T;
  const size_t rows = (size_t)ctx->res_P * ctx->res_G;
  float* planes[2] = {ctx->d_res_util, ctx->d_res_power};
  for (float* pl : planes) {
    if (!pl) continue;
    if (n_new >= T) {
      CU(cudaMemsetAsync(pl, 0xFF, rows * (size_t)T * sizeof(float), ctx->stream));
    } else {
      const uint64_t total = (uint64_t)rows * n_new;
      const uint32_t grid = (uint32_t)std::min<uint64_t>((total + 255) / 256, (uint64_t)ctx->sm_count * 16);
      gpr::text::k_fill_columns<<<grid, 256, 0, ctx->stream>>>(pl, (uint32_t)rows, T, T, ctx->res_head, n_new);
      ctx->launches++;
      CU(cudaGetLastError());
    }
  }
  ctx->res_head = (uint32_t)(((uint64_t)ctx->res_head + n_new) % T);
  CU(cudaStreamSynchronize(ctx->stream));
  return GPR_OK;
  GPR_CATCH(ctx)
}

// ---- multi-GPU -------------------------------------------------------------------------------
int gpr_comm_unique_id(void* id128) {
  gpr_ctx* ctx = nullptr;
  GPR_TRY
  if (!id128) return fail(nullptr, GPR_E_INVALID, "id buffer is NULL");
  char err[256];
  if (!load_nccl(err, sizeof err)) return fail(nullptr, GPR_E_NCCL, "%s", err);
  static_assert(sizeof(ncclUniqueId) == GPR_UNIQUE_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId id;
  NC(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, sizeof id);
  return GPR_OK;
  GPR_CATCH(nullptr)
}

int gpr_comm_init(gpr_ctx* ctx, const void* id128, int rank, int world) {
  if (!ctx) return GPR_E_INVALID;
  GPR_TRY
  if (!id128 || world < 1 || rank < 0 || rank >= world)
    return fail(ctx, GPR_E_INVALID, "bad communicator arguments (rank %d world %d)", rank, world);
  if (ctx->comm) return fail(ctx, GPR_E_STATE, "communicator already attached");
  char err[256];
  if (!load_nccl(err, sizeof err)) return fail(ctx, GPR_E_NCCL, "%s", err);
  CU(cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, id128, sizeof id);
  NC(g_nccl.CommInitRank(&ctx->comm, world, id, rank));
  ctx->rank = rank, ctx->world = world;
  return GPR_OK;
  GPR_CATCH(ctx)
}

int gpr_comm_destroy(gpr_ctx* ctx) {
  if (!ctx) return GPR_E_INVALID;
  GPR_TRY
  if (ctx->comm) {
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    NC(g_nccl.CommDestroy(ctx->comm));
    ctx->comm = nullptr;
  }
  ctx->rank = 0, ctx->world = 1;
  return GPR_OK;
  GPR_CATCH(ctx)
}

// Fused exchange over NVLink peer memory: every rank allocates one exchange block, publishes its
// CUDA IPC handle, and maps everybody else's.  From then on gpr_decide needs no collective launch.
int gpr_p2p_init(gpr_ctx* ctx, int rank, int world, uint32_t max_pods_per_rank, void* handle64) {
  if (!ctx) return GPR_E_INVALID;
  GPR_TRY
  if (!handle64 || world < 2 || world > gpr::kMaxPeers || rank < 0 || rank >= world)
    return fail(ctx, GPR_E_INVALID, "bad p2p arguments (rank %d world %d, at most %d ranks)", rank, world,
                gpr::kMaxPeers);
  if (ctx->p2p_block) return fail(ctx, GPR_E_STATE, "p2p exchange already initialised");
  if (ctx->comm && (ctx->rank != rank || ctx->world != world))
    return fail(ctx, GPR_E_INVALID, "rank/world differ from the NCCL communicator's");
  static_assert(sizeof(cudaIpcMemHandle_t) == GPR_P2P_HANDLE_BYTES, "cudaIpcMemHandle_t size");
  CU(cudaSetDevice(ctx->device));
  const uint32_t w_max = (max_pods_per_rank + 31u) / 32u;
  ctx->p2p_stride = 2u * std::max<uint32_t>(w_max, 1u);
  const size_t gather_bytes = ((size_t)world * ctx->p2p_stride * 4u + 255u) & ~(size_t)255u;
  const size_t ll_bytes = ((size_t)world * ctx->p2p_stride * 8u + 255u) & ~(size_t)255u;
  for (int k = 0; k < kExchangeDepth; ++k) {
    ctx->p2p_gather_off[k] = 256 + (size_t)k * gather_bytes;
    ctx->p2p_ll_off[k] = 256 + (size_t)kExchangeDepth * gather_bytes + (size_t)k * ll_bytes;
  }
  const size_t total = 256 + (size_t)kExchangeDepth * (gather_bytes + ll_bytes);
  CU(cudaMalloc(reinterpret_cast<void**>(&ctx->p2p_block), total));
  CU(cudaMemset(ctx->p2p_block, 0, total));
  cudaIpcMemHandle_t h;
  CU(cudaIpcGetMemHandle(&h, ctx->p2p_block));
  memcpy(handle64, &h, sizeof h);
  ctx->rank = rank, ctx->world = world;
  return GPR_OK;
  GPR_CATCH(ctx)
}

int gpr_p2p_attach(gpr_ctx* ctx, const void* handles) {
  if (!ctx) return GPR_E_INVALID;
  GPR_TRY
  if (!handles) return fail(ctx, GPR_E_INVALID, "handles is NULL");
  if (!ctx->p2p_block) return fail(ctx, GPR_E_STATE, "call gpr_p2p_init first");
  if (ctx->p2p_ready) return fail(ctx, GPR_E_STATE, "p2p exchange already attached");
  CU(cudaSetDevice(ctx->device));
  for (int r = 0; r < ctx->world; ++r) {
    if (r == ctx->rank) {
      ctx->p2p_peer[r] = ctx->p2p_block;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const unsigned char*>(handles) + (size_t)r * GPR_P2P_HANDLE_BYTES, sizeof h);
    void* p = nullptr;
    CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    ctx->p2p_peer[r] = static_cast<unsigned char*>(p);
  }
  ctx->p2p_ready = true;
  return GPR_OK;
  GPR_CATCH(ctx)
}

// ---- memory helpers -----------------------------------------------------------------------------
int gpr_host_alloc(gpr_ctx* ctx, size_t bytes, void** out) {
  if (!ctx || !out) return GPR_E_INVALID;
  CU(cudaSetDevice(ctx->device));
  CU(cudaMallocHost(out, bytes ? bytes : 1));
  return GPR_OK;
}
int gpr_host_free(gpr_ctx* ctx, void* p) {
  if (!ctx) return GPR_E_INVALID;
  if (p) CU(cudaFreeHost(p));
  return GPR_OK;
}
int gpr_device_alloc(gpr_ctx* ctx, size_t bytes, void** out) {
  if (!ctx || !out) return GPR_E_INVALID;
  CU(cudaSetDevice(ctx->device));
  CU(cudaMalloc(out, bytes ? bytes : 1));
  return GPR_OK;
}
int gpr_device_free(gpr_ctx* ctx, void* p) {
  if (!ctx) return GPR_E_INVALID;
  CU(cudaSetDevice(ctx->device));
  if (p) CU(cudaFree(p));
  return GPR_OK;
}
int gpr_memcpy(gpr_ctx* ctx, void* dst, const void* src, size_t bytes, int32_t dst_kind,
               int32_t src_kind) {
  if (!ctx) return GPR_E_INVALID;
  if (bytes == 0) return GPR_OK;
  if (!dst || !src) return fail(ctx, GPR_E_INVALID, "NULL pointer in gpr_memcpy");
  ctx->last_was_reduce = false;
  CU(cudaSetDevice(ctx->device));
  cudaMemcpyKind k = dst_kind == GPR_MEM_DEVICE
                         ? (src_kind == GPR_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice)
                         : (src_kind == GPR_MEM_DEVICE ? cudaMemcpyDeviceToHost : cudaMemcpyHostToHost);
  CU(cudaMemcpyAsync(dst, src, bytes, k, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return GPR_OK;
}

// ---- measurement support --------------------------------------------------------------------------
int gpr_timer_begin(gpr_ctx* ctx) {
  if (!ctx) return GPR_E_INVALID;
  ctx->last_was_reduce = false;
  CU(cudaSetDevice(ctx->device));
  // rendezvous first, then the start event: with an exchange attached the timed regions of all ranks
  // begin within an NVLink round trip of each other, whatever the skew between their host threads
  gpr::RendezvousParams q;
  memset(&q, 0, sizeof q);
  q.world = 1, q.rank = 0;
  q.stamp = ctx->h_mark;
  q.err = ctx->h_err;
  if (ctx->p2p_ready && ctx->world > 1) {
    q.world = ctx->world, q.rank = ctx->rank;
    q.seq = ++ctx->rdv_seq;
    for (int r = 0; r < ctx->world; ++r)
      q.peer_flag[r] = reinterpret_cast<unsigned long long*>(ctx->p2p_peer[r]) + gpr::kMaxPeers + ctx->rank;
    q.my_flags = reinterpret_cast<const unsigned long long*>(ctx->p2p_block) + gpr::kMaxPeers;
  } else if (ctx->comm && ctx->world > 1) {
    // NCCL only: a one-word allgather is the rendezvous
    int rc = grow(ctx, &ctx->d_gather, &ctx->gather_cap, (size_t)ctx->world + 2);
    if (rc != GPR_OK) return rc;
    rc = grow(ctx, &ctx->d_bits, &ctx->bits_cap, 4);
    if (rc != GPR_OK) return rc;
    NC(g_nccl.AllGather(ctx->d_bits, ctx->d_gather, 1, ncclUint32, ctx->comm, ctx->stream));
  }
  gpr::k_rendezvous<<<1, 32, 0, ctx->stream>>>(q);
  ctx->launches++;
  CU(cudaGetLastError());
  CU(cudaEventRecord(ctx->ev_t0, ctx->stream));
  return GPR_OK;
}
int gpr_step_stamps(gpr_ctx* ctx, uint64_t* ns, uint32_t cap, uint32_t* n, uint64_t* begin_ns) {
  if (!ctx || !n) return GPR_E_INVALID;
  *n = (uint32_t)ctx->stamps.size();
  if (begin_ns) *begin_ns = *ctx->h_mark;
  if (ns)
    for (uint32_t i = 0; i < cap && i < *n; ++i) ns[i] = ctx->stamps[i];
  return GPR_OK;
}
int gpr_phase_stamps(gpr_ctx* ctx, uint64_t* ns, uint32_t cap, uint32_t* n) {
  if (!ctx || !n) return GPR_E_INVALID;
  *n = (uint32_t)ctx->phase_stamps.size();
  if (ns)
    for (uint32_t i = 0; i < cap && i < *n; ++i) ns[i] = ctx->phase_stamps[i];
  return GPR_OK;
}
int gpr_p2p_debug(gpr_ctx* ctx, int32_t mode) {
  if (!ctx) return GPR_E_INVALID;
  if (mode < 0 || mode > 2) return fail(ctx, GPR_E_INVALID, "exchange debug mode %d (0..2)", mode);
  ctx->exchange_debug = mode;
  return GPR_OK;
}
int gpr_timer_end(gpr_ctx* ctx, double* ms) {
  if (!ctx || !ms) return GPR_E_INVALID;
  ctx->last_was_reduce = false;
  CU(cudaSetDevice(ctx->device));
  CU(cudaEventRecord(ctx->ev_t1, ctx->stream));
  CU(cudaEventSynchronize(ctx->ev_t1));
  float f = 0.f;
  CU(cudaEventElapsedTime(&f, ctx->ev_t0, ctx->ev_t1));
  *ms = f;
  return GPR_OK;
}
int gpr_flush_l2(gpr_ctx* ctx) {
  if (!ctx) return GPR_E_INVALID;
  ctx->last_was_reduce = false;
  CU(cudaSetDevice(ctx->device));
  if (!ctx->d_flush) {
    ctx->flush_bytes = std::max<size_t>(ctx->l2_bytes * 2, (size_t)256 << 20);
    CU(cudaMalloc(&ctx->d_flush, ctx->flush_bytes));
  }
  CU(cudaMemsetAsync(ctx->d_flush, 0x5a, ctx->flush_bytes, ctx->stream));
  return GPR_OK;
}
int gpr_launch_count(const gpr_ctx* ctx, uint64_t* n) {
  if (!ctx || !n) return GPR_E_INVALID;
  *n = ctx->launches;
  return GPR_OK;
}
int gpr_get_device_info(gpr_ctx* ctx, gpr_device_info* info) {
  if (!ctx || !info) return GPR_E_INVALID;
  if (info->struct_size != sizeof(gpr_device_info))
    return fail(ctx, GPR_E_INVALID, "gpr_device_info.struct_size mismatch");
  info->sm_count = ctx->sm_count;
  info->cc_major = ctx->cc_major, info->cc_minor = ctx->cc_minor;
  info->l2_bytes = ctx->l2_bytes, info->hbm_bytes = ctx->hbm_bytes;
  memcpy(info->name, ctx->name, sizeof info->name);
  return GPR_OK;
}

// ---- synthetic windows -----------------------------------------------------------------------------
// ---- device-side ingest of the response text -------------------------------------------------------
static_assert(sizeof(gpr_text_span) == sizeof(gpr::text::Span) && offsetof(gpr_text_span, row) == offsetof(gpr::text::Span, row) &&
                  offsetof(gpr_text_span, n_tiny) == offsetof(gpr::text::Span, n_tiny),
              "gpr_text_span mirrors gpr::text::Span");
static_assert(GPR_SPAN_SHARED == gpr::text::kSpanShared && GPR_SPAN_HARD == gpr::text::kSpanHard, "span flags");

int gpr_text_scan_begin(gpr_ctx* ctx, int32_t slot, const char* text, uint64_t n_bytes, int32_t mem_kind) {
  if (!ctx) return GPR_E_INVALID;
  GPR_TRY
  NvtxRange nvtx_range("gpr_text_scan_begin");
  if (slot < 0 || slot > 2) return fail(ctx, GPR_E_INVALID, "text slot %d (0..2)", slot);
  if (!text && n_bytes) return fail(ctx, GPR_E_INVALID, "text is NULL");
  if (mem_kind != GPR_MEM_HOST && mem_kind != GPR_MEM_DEVICE) return fail(ctx, GPR_E_INVALID, "bad mem_kind %d", mem_kind);
  scan_pipe_abort(ctx);  // an unfinished scan is dropped
  CU(cudaSetDevice(ctx->device));
  int rc;
  if ((rc = grow(ctx, &ctx->d_text[slot], &ctx->text_cap[slot], (size_t)n_bytes + gpr::text::kTextPad)) != GPR_OK)
    return rc;
  constexpr int NT = gpr_ctx::kUpThreads, NS = gpr_ctx::kUpSlots, NB = gpr_ctx::kMarkBlocks;
  if (!ctx->h_mark_blocks) {
    CU(cudaHostAlloc(reinterpret_cast<void**>(&ctx->h_mark_blocks), (size_t)NB * kBlockWords * sizeof(uint32_t),
                     cudaHostAllocMapped));
    CU(cudaMalloc(reinterpret_cast<void**>(&ctx->d_mark_blocks), (size_t)NB * kBlockWords * sizeof(uint32_t)));
    for (int b = 0; b < NB; ++b) CU(cudaEventCreateWithFlags(&ctx->mark_event[b], cudaEventDisableTiming));
    for (int k = 0; k < NT; ++k) {
      CU(cudaStreamCreateWithFlags(&ctx->up_stream[k], cudaStreamNonBlocking));
      for (int s = 0; s < NS; ++s) CU(cudaEventCreateWithFlags(&ctx->up_event[k][s], cudaEventDisableTiming));
    }
  }
  uint8_t* d = ctx->d_text[slot];
  ctx->text_n[slot] = n_bytes;
  ctx->last_was_reduce = false;
  // the zero pad behind the text, and everything earlier on the context's stream that may still read the buffer
  CU(cudaMemsetAsync(d + n_bytes, 0, gpr::text::kTextPad, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  ScanPipe* sp = new ScanPipe();
  sp->slot = slot, sp->src = text, sp->dst = d, sp->n = n_bytes, sp->src_kind = mem_kind;
  for (auto& r : sp->recorded) r.store(0);
  if (mem_kind == GPR_MEM_HOST && n_bytes) {
    cudaPointerAttributes at;
    const bool pinned = cudaPointerGetAttributes(&at, text) == cudaSuccess && at.type == cudaMemoryTypeHost;
    (void)cudaGetLastError();  // an unregistered pointer may leave an error code behind
    sp->staged = !pinned;
  }
  sp->chunk = sp->staged ? ctx->up_chunk : gpr_ctx::kPinnedChunk;
  sp->n_chunks = (n_bytes + sp->chunk - 1) / sp->chunk;
  // pinned and device sources need no staging copy: one producer keeps the DMA engine busy
  sp->nt = sp->staged ? (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)ctx->up_threads, sp->n_chunks)) : 1;
  if (sp->staged && !ctx->h_up_ring)
    CU(cudaHostAlloc(reinterpret_cast<void**>(&ctx->h_up_ring), (size_t)ctx->up_threads * NS * ctx->up_slot_bytes,
                     cudaHostAllocDefault));
  ctx->pipe = sp;
  ctx->launches += 2 * sp->n_chunks;
  try {
    for (int k = 0; k < sp->nt; ++k) sp->th.emplace_back(scan_producer, ctx, sp, k);
  } catch (...) {
    if (sp->th.empty()) {  // no thread to be had at all: produce on this thread (blocks run ahead by at most NB chunks)
      (void)scan_pipe_finish(ctx, false);
      return fail(ctx, GPR_E_NOMEM, "could not start an upload thread");
    }
    sp->nt = (int)sp->th.size();  // fewer producers than planned would skip chunks: restart with what we have
    (void)scan_pipe_finish(ctx, false);
    return fail(ctx, GPR_E_NOMEM, "could not start the upload threads");
  }
  return GPR_OK;
  GPR_CATCH(ctx)
}

int gpr_text_scan_next(gpr_ctx* ctx, uint64_t* opens, uint64_t* closes, uint64_t cap, uint64_t* n_opens, uint64_t* n_closes,
                       uint64_t* bytes_done, int32_t* more) {
  if (!ctx) return GPR_E_INVALID;
  GPR_TRY
  if (!n_opens || !n_closes || !bytes_done || !more || (cap && (!opens || !closes)))
    return fail(ctx, GPR_E_INVALID, "output pointers are NULL");
  ScanPipe* sp = ctx->pipe;
  if (!sp) return fail(ctx, GPR_E_STATE, "no scan in progress (gpr_text_scan_begin)");
  *n_opens = *n_closes = 0;
  if (sp->next >= sp->n_chunks) {  // everything delivered (also: empty text)
    *bytes_done = sp->n, *more = 0;
    return scan_pipe_finish(ctx, true);
  }
  constexpr int NB = gpr_ctx::kMarkBlocks;
  const uint64_t c = sp->next;
  while (sp->recorded[c % NB].load(std::memory_order_acquire) != c + 1) {
    if (sp->error.load() || sp->stop.load()) return scan_pipe_finish(ctx, false) != GPR_OK ? GPR_E_CUDA : fail(ctx, GPR_E_CUDA, "text upload stopped");
    std::this_thread::yield();
  }
  CU(cudaSetDevice(ctx->device));
  cudaError_t e = cudaEventSynchronize(ctx->mark_event[c % NB]);
  if (e != cudaSuccess) {
    (void)scan_pipe_finish(ctx, false);
    return fail(ctx, GPR_E_CUDA, "text scan: %s", cudaGetErrorString(e));
  }
  const uint32_t* block = ctx->h_mark_blocks + (size_t)(c % NB) * kBlockWords;
  const uint64_t no = block[0], nc = block[1], base = c * sp->chunk;
  if (no > gpr_ctx::kMarkCap || nc > gpr_ctx::kMarkCap || no > cap || nc > cap) {
    const bool caller = no <= gpr_ctx::kMarkCap && nc <= gpr_ctx::kMarkCap;
    *n_opens = no, *n_closes = nc;
    if (!caller) (void)scan_pipe_finish(ctx, false);  // (a too small `cap` may be retried with a larger one)
    return fail(ctx, GPR_E_CAPACITY, "%llu / %llu markers in one %llu-byte chunk, room for %llu", (unsigned long long)no,
                (unsigned long long)nc, (unsigned long long)sp->chunk, (unsigned long long)(caller ? cap : gpr_ctx::kMarkCap));
  }
  for (uint64_t i = 0; i < no; ++i) opens[i] = base + block[2 + i];
  for (uint64_t i = 0; i < nc; ++i) closes[i] = base + block[2 + gpr_ctx::kMarkCap + i];
  std::sort(opens, opens + no);
  std::sort(closes, closes + nc);
  *n_opens = no, *n_closes = nc;
  sp->next = c + 1;
  sp->consumed.store(c + 1, std::memory_order_release);
  *bytes_done = std::min<uint64_t>(sp->n, (c + 1) * sp->chunk);
  *more = 1;
  if (sp->next >= sp->n_chunks) {
    *more = 0;
    return scan_pipe_finish(ctx, true);
  }
  return GPR_OK;
  GPR_CATCH(ctx)
}

int gpr_text_scan(gpr_ctx* ctx, int32_t slot, const char* text, uint64_t n_bytes, int32_t mem_kind,
                  uint64_t* opens, uint64_t* closes, uint64_t cap, uint64_t* n_opens, uint64_t* n_closes) {
  if (!ctx) return GPR_E_INVALID;
  GPR_TRY
  NvtxRange nvtx_range("gpr_text_scan");
  if (!n_opens || !n_closes || (cap && (!opens || !closes))) return fail(ctx, GPR_E_INVALID, "output arrays are NULL");
  int rc = gpr_text_scan_begin(ctx, slot, text, n_bytes, mem_kind);
  if (rc != GPR_OK) return rc;
  std::vector<uint64_t> co(gpr_ctx::kMarkCap), cc(gpr_ctx::kMarkCap);
  uint64_t no = 0, nc = 0;
  int32_t more = 1;
  while (more) {
    uint64_t a = 0, b = 0, done = 0;
    rc = gpr_text_scan_next(ctx, co.data(), cc.data(), gpr_ctx::kMarkCap, &a, &b, &done, &more);
    if (rc != GPR_OK) {
      scan_pipe_abort(ctx);
      return rc;
    }
    for (uint64_t i = 0; i < a; ++i, ++no)
      if (no < cap) opens[no] = co[i];
    for (uint64_t i = 0; i < b; ++i, ++nc)
      if (nc < cap) closes[nc] = cc[i];
  }
  CU(cudaStreamSynchronize(ctx->stream));
  *n_opens = no, *n_closes = nc;
  if (no > cap || nc > cap)
    return fail(ctx, GPR_E_CAPACITY, "%llu / %llu markers, room for %llu: call again with a larger cap", (unsigned long long)no,
                (unsigned long long)nc, (unsigned long long)cap);
  return GPR_OK;
  GPR_CATCH(ctx)
}

int gpr_text_parse(gpr_ctx* ctx, int32_t slot, gpr_text_span* spans, uint32_t n_spans, const gpr_text_grid* grid,
                   int32_t plane) {
  if (!ctx) return GPR_E_INVALID;
  NvtxRange nvtx_range("gpr_text_parse");
  if (!grid || grid->struct_size != sizeof(gpr_text_grid)) return fail(ctx, GPR_E_INVALID, "grid is NULL / struct_size mismatch");
  if (slot < 0 || slot > 2 || plane < 0 || plane > 1) return fail(ctx, GPR_E_INVALID, "bad slot %d / plane %d", slot, plane);
  if (!ctx->d_text[slot]) return fail(ctx, GPR_E_STATE, "no text in slot %d (gpr_text_scan)", slot);
  if (n_spans && !spans) return fail(ctx, GPR_E_INVALID, "spans is NULL");
  const uint32_t n_samples = grid->n_samples, n_rows = grid->n_rows;
  if (grid->step <= 0 || grid->step > 4000000ll || n_samples == 0 || grid->window_seconds <= 0 ||
      grid->window_seconds > 4000000000ll)
    return fail(ctx, GPR_E_INVALID, "step (<= 4e6 s), window_seconds and n_samples must be > 0");
  if ((grid->window_seconds + grid->step - 1) / grid->step > (int64_t)n_samples)
    return fail(ctx, GPR_E_INVALID, "window of %lld s needs more than %u columns of %lld s", (long long)grid->window_seconds,
                n_samples, (long long)grid->step);
  const bool resident = (grid->flags & GPR_TEXT_RESIDENT) != 0;
  const uint64_t n = ctx->text_n[slot];
  for (uint32_t i = 0; i < n_spans; ++i) {
    if (spans[i].begin > spans[i].end || spans[i].end > n || spans[i].row >= n_rows ||
        (i && spans[i].begin < spans[i - 1].end))
      return fail(ctx, GPR_E_INVALID, "span %u is out of order, out of the text or out of the plane", i);
    spans[i].flags &= GPR_SPAN_SHARED;
    spans[i].n_in = spans[i].n_oow = spans[i].n_tiny = 0;
  }
  CU(cudaSetDevice(ctx->device));
  int rc;
  float* pl = nullptr;
  gpr::text::Grid g;
  memset(&g, 0, sizeof g);
  // the device works in milliseconds, the resolution of Prometheus timestamps
  g.t_end = grid->t_end * 1000, g.t_lo = (grid->t_end - grid->window_seconds) * 1000;
  g.step = (uint32_t)(grid->step * 1000), g.T = n_samples;
  if (resident) {
    if (!ctx->d_res_util) return fail(ctx, GPR_E_STATE, "no resident window (gpr_resident_init)");
    if (n_samples != ctx->res_T || (uint64_t)n_rows > (uint64_t)ctx->res_P * ctx->res_G)
      return fail(ctx, GPR_E_INVALID, "grid %u rows x %u does not match the resident window (%u x %u)", n_rows, n_samples,
                  ctx->res_P * ctx->res_G, ctx->res_T);
    pl = plane == 0 ? ctx->d_res_util : ctx->d_res_power;
    if (!pl) return fail(ctx, GPR_E_STATE, "the resident window has no power plane");
    g.ld = ctx->res_T;
    g.col_end = (ctx->res_head + ctx->res_T - 1) % ctx->res_T;  // the newest bucket sits just before the head
  } else {
    const size_t cells = (size_t)n_rows * n_samples;
    const size_t cap_before = ctx->tplane_cap[plane];
    if ((rc = grow(ctx, &ctx->d_tplane[plane], &ctx->tplane_cap[plane], cells + 4)) != GPR_OK) return rc;
    if (ctx->tplane_cap[plane] != cap_before && !(grid->flags & GPR_TEXT_FILL))
      return fail(ctx, GPR_E_STATE, "plane %d had to grow: the first parse of a window must pass GPR_TEXT_FILL", plane);
    pl = ctx->d_tplane[plane];
    g.ld = n_samples, g.col_end = n_samples - 1;
    // 0xFFFFFFFF: a NaN, and -1 as an int — below every non-negative sample for the integer atomicMax merge
    if ((grid->flags & GPR_TEXT_FILL) && cells) CU(cudaMemsetAsync(pl, 0xFF, cells * sizeof(float), ctx->stream));
  }
  if ((rc = grow(ctx, &ctx->d_spans, &ctx->spans_cap, (size_t)n_spans + 1)) != GPR_OK) return rc;
  ctx->last_was_reduce = false;
  if (n_spans && n) {
    CU(cudaMemcpyAsync(ctx->d_spans, spans, (size_t)n_spans * sizeof(gpr_text_span), cudaMemcpyHostToDevice,
                       ctx->stream));
    constexpr int kWarps = 4;
    const uint64_t tiles = (n + gpr::text::kTileBytes - 1) / gpr::text::kTileBytes;
    const uint32_t blocks = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>((tiles + kWarps - 1) / kWarps,
                                                                                (uint64_t)ctx->sm_count * ctx->parse_ctas_per_sm));
    gpr::text::k_text_parse<kWarps><<<blocks, kWarps * 32, gpr::text::text_parse_smem<kWarps>(), ctx->stream>>>(
        ctx->d_text[slot], n, n + gpr::text::kTextPad, ctx->d_spans, n_spans, g, pl);
    ctx->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(spans, ctx->d_spans, (size_t)n_spans * sizeof(gpr_text_span), cudaMemcpyDeviceToHost,
                       ctx->stream));
  }
  CU(cudaStreamSynchronize(ctx->stream));
  return GPR_OK;
}

int gpr_text_planes(gpr_ctx* ctx, float** util, float** power) {
  if (!ctx) return GPR_E_INVALID;
  if (util) *util = ctx->d_tplane[0];
  if (power) *power = ctx->d_tplane[1];
  return GPR_OK;
}

int gpr_synth_fill(gpr_ctx* ctx, uint64_t seed, int32_t plane, float* dst, uint64_t pod_offset,
                   uint32_t n_pods, uint32_t n_gpus, uint32_t n_samples, uint64_t row_stride) {
  if (!ctx) return GPR_E_INVALID;
  if (!dst || n_gpus == 0 || n_samples == 0 || (plane != 0 && plane != 1))
    return fail(ctx, GPR_E_INVALID, "bad gpr_synth_fill arguments");
  const uint64_t rows = (uint64_t)n_pods * n_gpus;
  if (rows == 0) return GPR_OK;
  ctx->last_was_reduce = false;
  if (rows > 0x7fffffffull) return fail(ctx, GPR_E_INVALID, "too many series");
  CU(cudaSetDevice(ctx->device));
  const uint64_t ld = row_stride ? row_stride : n_samples;
  const uint32_t grid = (uint32_t)std::min<uint64_t>(rows, (uint64_t)ctx->sm_count * 32);
  gpr::k_synth_fill<<<grid, 256, 0, ctx->stream>>>(dst, seed, plane, pod_offset * n_gpus,
                                                   (uint32_t)rows, n_samples, ld);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(ctx->stream));
  return GPR_OK;
}

int gpr_synth_eligible(gpr_ctx* ctx, uint64_t seed, uint8_t* dst, uint64_t pod_offset,
                       uint32_t n_pods) {
  if (!ctx) return GPR_E_INVALID;
  if (!dst) return fail(ctx, GPR_E_INVALID, "dst is NULL");
  if (n_pods == 0) return GPR_OK;
  ctx->last_was_reduce = false;
  CU(cudaSetDevice(ctx->device));
  const uint32_t grid = std::min<uint32_t>((n_pods + 255u) / 256u, (uint32_t)ctx->sm_count * 8u);
  gpr::k_synth_eligible<<<grid, 256, 0, ctx->stream>>>(dst, seed, pod_offset, n_pods);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(ctx->stream));
  return GPR_OK;
}

}  // extern "C"

#ifdef GPR_TIMELINE
extern "C" GPR_API int gpr_debug_timeline(gpr_ctx* ctx, unsigned long long* out, int n_ctas) {
  if (!ctx || !out) return GPR_E_INVALID;
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  CU(cudaMemcpyFromSymbol(out, gpr::g_timeline, sizeof(unsigned long long) * 4 * (size_t)n_ctas));
  return GPR_OK;
}
#endif
