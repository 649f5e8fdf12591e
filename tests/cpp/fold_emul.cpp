// Host emulation of the fold kernel and its fused multi-GPU exchange.
//
// tests/test_fold_exchange_emul.py cuts fold_words, block_counts, exchange_bitmaps, exchange_bitmaps_ll and k_fold out
// of gpu-pruner_b200/csrc/gpr_kernels.cuh verbatim (one inline-PTX store and the two __shared__ declarations are
// rewritten to shim calls) into fold_extract.inc.  This file supplies the CUDA names that text uses on top of
// std::thread — a CTA is a group of threads with a barrier, a warp 32 of them with an emulated ballot / shuffle —
// and drives several "ranks" through a sequence of decisions the way gpr_api.cu's decide_impl does: two scratch
// sets, exchange buffers four deep, a reduce that may run while the previous decision's fold is still exchanging
// and that waits for its scratch set before publishing, folds launched without waiting for each other.
//
// Checked for every protocol (tagged slots in order / pipelined, flags): every decision's counters, every
// decision's gathered rank-major bitmaps, the order of writes into a shared output buffer, and that all scratch
// state is back to zero — under random delays that let ranks drift steps apart.  It validates the SOURCE logic
// (indexing, waits, ordering, absence of deadlock), not the generated machine code; tests/test_gpu_multi.py does that
// on real GPUs.
#include <algorithm>
#include <atomic>
#include <barrier>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <random>
#include <thread>
#include <vector>

using std::min;

// ---- the CUDA names the extracted text uses -----------------------------------------------------------------
#define __device__
#define __global__
#define __forceinline__ inline
#define __launch_bounds__(x)

struct Dim3 { unsigned x; };
struct WarpCtx {
  std::barrier<> bar{32};
  uint32_t slot[32];
};
struct CtaCtx {
  explicit CtaCtx(unsigned threads) : bar((std::ptrdiff_t)threads), warps(threads / 32) {}
  std::barrier<> bar;
  std::vector<WarpCtx> warps;
  unsigned long long s_cnt[3] = {0, 0, 0};
  unsigned int s_last = 0;
};
static thread_local Dim3 threadIdx, blockIdx, blockDim, gridDim;
static thread_local CtaCtx* tl_cta = nullptr;
static thread_local const std::atomic<int>* tl_reduce_complete = nullptr;
static thread_local std::minstd_rand tl_rng;
static thread_local bool tl_slow = false;   // a GPU whose collecting CTA keeps being descheduled
static std::atomic<bool> g_deadline_hit{false};

static inline void jitter() {   // let the OS reorder things now and then
  if (tl_slow) {
    if ((tl_rng() & 3u) == 0) std::this_thread::sleep_for(std::chrono::milliseconds(5 + tl_rng() % 20));
  } else if ((tl_rng() & 255u) == 0) {
    std::this_thread::sleep_for(std::chrono::microseconds(tl_rng() % 300));
  }
}
static inline void __syncthreads() { tl_cta->bar.arrive_and_wait(); }
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __nanosleep(unsigned) { std::this_thread::yield(); }
static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline unsigned long long gtime() {
  return (unsigned long long)std::chrono::duration_cast<std::chrono::nanoseconds>(
             std::chrono::steady_clock::now().time_since_epoch()).count();
}
template <class T> static inline T __ldcg(const T* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
static inline uint32_t __ballot_sync(unsigned, bool pred) {
  WarpCtx& w = tl_cta->warps[threadIdx.x >> 5];
  w.slot[threadIdx.x & 31] = pred ? 1u : 0u;
  w.bar.arrive_and_wait();
  uint32_t r = 0;
  for (int l = 0; l < 32; ++l) r |= w.slot[l] << l;
  w.bar.arrive_and_wait();
  return r;
}
static inline uint32_t __shfl_xor_sync(unsigned, uint32_t v, int o) {
  WarpCtx& w = tl_cta->warps[threadIdx.x >> 5];
  w.slot[threadIdx.x & 31] = v;
  w.bar.arrive_and_wait();
  const uint32_t r = w.slot[(threadIdx.x & 31) ^ o];
  w.bar.arrive_and_wait();
  return r;
}
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
  return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST);
}
static inline unsigned int atomicAdd(unsigned int* p, unsigned int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline void spin(const unsigned long long* p, unsigned long long want, unsigned int* err) {
  const auto t0 = std::chrono::steady_clock::now();
  unsigned n = 0;
  while (__atomic_load_n(p, __ATOMIC_ACQUIRE) < want) {
    std::this_thread::yield();
    if ((++n & 4095u) == 0 && std::chrono::steady_clock::now() - t0 > std::chrono::seconds(60)) {
      g_deadline_hit = true;
      if (err) *err = 1u;
      return;
    }
  }
}
static inline void spin_until_gpu(const unsigned long long* p, unsigned long long want) { spin(p, want, nullptr); }
static inline void spin_until_sys(const unsigned long long* p, unsigned long long want, unsigned int* err, unsigned) {
  spin(p, want, err);
}
static inline unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
  jitter();
  return __atomic_load_n(p, __ATOMIC_RELAXED);
}
static inline void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) {
  jitter();
  __atomic_store_n(p, v, __ATOMIC_RELAXED);
}
static inline void st_release_sys_u64(unsigned long long* p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
static inline void st_release_u64(unsigned long long* p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
static inline void pdl_launch_dependents() {}
static inline void pdl_wait_prior_grids() {   // the reduce grid of this decision has completed
  while (tl_reduce_complete->load(std::memory_order_acquire) == 0) std::this_thread::yield();
}

namespace gpr {
#include "fold_extract.inc"
}

// ---- one emulated GPU ------------------------------------------------------------------------------------------
constexpr int kDepth = 4;   // exchange buffer sets (gpr_api.cu kExchangeDepth)
struct Rank {
  int rank = 0;
  uint32_t P = 0, W = 0, stride = 0;
  std::vector<uint32_t> masks[2];                 // [idle P | veto P]
  std::vector<uint8_t> eligible;
  unsigned long long acc[6] = {0, 0, 0, 0, 0, 0};
  unsigned int tickets[2] = {0, 0};
  unsigned long long done[2] = {0, 0}, uses[2] = {0, 0};
  std::vector<unsigned long long> flags;          // [world]
  std::vector<uint32_t> gather[kDepth];           // [world][stride]
  std::vector<unsigned long long> ll[kDepth];     // [world][stride]
  unsigned int err = 0;
  std::vector<std::thread> gpu_threads;
};

struct Scenario {
  int world;
  uint32_t words_per_rank;
  unsigned fold_threads;
  int steps;
  int protocol;       // 0 tagged slots in order, 1 tagged slots pipelined, 2 flags
  int shared_output;  // 1: every decision writes the same output buffer (the last one must stay)
  int slow_rank;      // >= 0: that rank's loads of peer-written slots are slow, so its collector lags steps behind
  int with_veto;      // 1: a power plane vetoes pods and the caller asks for veto_bits (decide_impl then keeps the
                      //    early wait even under the pipelined protocol: veto words go straight to the caller)
};

static uint32_t veto_bits_of(uint64_t seed, int rank, int step, uint32_t pod) {
  uint64_t x = (seed * 0x9e3779b97f4a7c15ull) ^ ((uint64_t)rank << 40) ^ ((uint64_t)step << 24) ^ ((uint64_t)pod * 0x100000001b3ull);
  x ^= x >> 29, x *= 0xbf58476d1ce4e5b9ull, x ^= x >> 32;
  return (x % 10u) == 0 ? 1u + (uint32_t)(x >> 40) % 15u : 0u;   // every tenth pod has a GPU drawing power
}
static uint32_t idle_bits(uint64_t seed, int rank, int step, uint32_t pod) {
  uint64_t x = seed ^ ((uint64_t)rank << 48) ^ ((uint64_t)step << 32) ^ pod;
  x ^= x >> 33, x *= 0xff51afd7ed558ccdull, x ^= x >> 33, x *= 0xc4ceb9fe1a85ec53ull, x ^= x >> 33;
  return (x & 3u) == 0 ? 0u : (uint32_t)(x >> 8) & 0xfu;   // G = 4 series per pod, a quarter of the pods busy
}

static int run(const Scenario& sc, uint64_t seed) {
  const int world = sc.world;
  const uint32_t W = sc.words_per_rank, P = W * 32u, stride = 2u * W + 2u;
  std::vector<std::unique_ptr<Rank>> ranks;
  std::mt19937_64 rng(seed);
  for (int r = 0; r < world; ++r) {
    auto R = std::make_unique<Rank>();
    R->rank = r, R->P = P, R->W = W, R->stride = stride;
    for (auto& m : R->masks) m.assign(2 * (size_t)P, 0u);
    R->eligible.resize(P);
    for (auto& e : R->eligible) e = (rng() % 10) != 0;
    R->flags.assign(world, 0ull);
    for (auto& g : R->gather) g.assign((size_t)world * stride, 0xabababab);
    for (auto& l : R->ll) l.assign((size_t)world * stride, 0ull);
    ranks.push_back(std::move(R));
  }
  const int K = sc.steps;
  // outputs: [rank][step or 0][world * W] decision and candidate words, counters per step
  const int n_out = sc.shared_output ? 1 : K + 1;
  std::vector<std::vector<uint32_t>> out_d(world), out_c(world);
  std::vector<std::vector<unsigned long long>> counts(world), stamps(world);
  std::vector<std::vector<uint32_t>> out_v(world);   // veto words: this rank's pods only, one buffer per decision
  std::vector<std::vector<std::atomic<int>>> reduce_complete(world);
  for (int r = 0; r < world; ++r) {
    out_d[r].assign((size_t)n_out * world * W, 0xdeadbeefu);
    out_c[r].assign((size_t)n_out * world * W, 0xdeadbeefu);
    out_v[r].assign((size_t)(K + 1) * W, 0xdeadbeefu);
    counts[r].assign((size_t)(K + 1) * 3, ~0ull);
    stamps[r].assign((size_t)(K + 1) * 5, 0ull);
    reduce_complete[r] = std::vector<std::atomic<int>>(K + 1);
    for (auto& a : reduce_complete[r]) a = 0;
  }
  const uint32_t warps = sc.fold_threads / 32u;
  const uint32_t grid = std::max<uint32_t>(1u, (W + 4u * warps - 1u) / (4u * warps));

  auto stream = [&](int r) {   // what decide_impl + the reduce kernel do on rank r, decision after decision
    Rank& me = *ranks[r];
    std::minstd_rand srng((unsigned)(seed * 31 + r));
    std::vector<std::unique_ptr<CtaCtx>> ctas;
    for (int n = 1; n <= K; ++n) {
      const unsigned sset = (unsigned)(n - 1) & 1u, xset = (unsigned)n % kDepth;
      gpr::FoldParams fp;
      memset(&fp, 0, sizeof fp);
      fp.idle_mask = me.masks[sset].data();
      fp.eligible = me.eligible.data();
      if (sc.with_veto) fp.veto_mask = me.masks[sset].data() + P, fp.vbits = out_v[r].data() + (size_t)n * W;
      fp.dbits = me.gather[xset].data() + (size_t)r * stride;
      fp.cbits = fp.dbits + W;
      fp.counts = &counts[r][(size_t)n * 3];
      fp.stamp = &stamps[r][(size_t)n * 5];
      fp.err = &me.err;
      fp.acc = me.acc + 3 * sset, fp.ticket = me.tickets + sset;
      fp.done = me.done + sset, fp.need = me.uses[sset];
      fp.prev_done = me.done + (sset ^ 1u), fp.prev_need = me.uses[sset ^ 1u];
      fp.P = P, fp.G = 4, fp.mw = 1;
      fp.world = world, fp.rank = r, fp.rank_stride = stride;
      fp.poll_ns = 200;
      for (int q = 0; q < world; ++q) {
        fp.peer_gather[q] = ranks[q]->gather[xset].data();
        fp.peer_flag[q] = ranks[q]->flags.data() + r;
        fp.peer_ll[q] = sc.protocol == 2 ? nullptr : ranks[q]->ll[xset].data();
      }
      fp.my_flags = me.flags.data();
      fp.my_ll = sc.protocol == 2 ? nullptr : me.ll[xset].data();
      fp.late_order = sc.protocol == 1 && fp.vbits == nullptr;
      fp.step = (unsigned long long)n;
      const size_t o = sc.shared_output ? 0 : (size_t)n * world * W;
      fp.out_dbits = out_d[r].data() + o, fp.out_cbits = out_c[r].data() + o;

      // ---- reduce n: streams for a while, waits for its scratch set before the first publish, then completes
      if ((srng() & 7u) == (unsigned)r % 8u) std::this_thread::sleep_for(std::chrono::microseconds(srng() % 4000));
      spin(me.done + sset, me.uses[sset], &me.err);   // wait_scratch_free
      for (uint32_t pod = 0; pod < P; ++pod) {
        const uint32_t b = idle_bits(seed, r, n, pod);
        if (b) __atomic_fetch_or(&me.masks[sset][pod], b, __ATOMIC_RELAXED);
        const uint32_t v = sc.with_veto ? veto_bits_of(seed, r, n, pod) : 0u;
        if (v) __atomic_fetch_or(&me.masks[sset][P + pod], v, __ATOMIC_RELAXED);
      }
      // ---- fold n: resident already, runs once the reduce has completed; nothing orders it behind fold n - 1
      for (uint32_t c = 0; c < grid; ++c) {
        ctas.push_back(std::make_unique<CtaCtx>(sc.fold_threads));
        CtaCtx* cta = ctas.back().get();
        for (unsigned t = 0; t < sc.fold_threads; ++t)
          me.gpu_threads.emplace_back([&, fp, cta, c, t, n, r] {
            threadIdx.x = t, blockIdx.x = c, blockDim.x = sc.fold_threads, gridDim.x = grid;
            tl_cta = cta, tl_reduce_complete = &reduce_complete[r][n];
            tl_slow = r == sc.slow_rank;
            tl_rng.seed((unsigned)(seed + 977u * (unsigned)n + 131u * c + t + 7u * (unsigned)r));
            if (world > 1) gpr::k_fold<true>(fp);
            else gpr::k_fold<false>(fp);
          });
      }
      reduce_complete[r][n].store(1, std::memory_order_release);
      me.uses[sset]++;
    }
    for (auto& t : me.gpu_threads) t.join();
  };
  std::vector<std::thread> streams;
  for (int r = 0; r < world; ++r) streams.emplace_back(stream, r);
  for (auto& t : streams) t.join();

  // ---- what every rank must hold now ---------------------------------------------------------------------------
  int bad = 0;
  if (g_deadline_hit) bad++;
  for (int n = 1; n <= K; ++n) {
    std::vector<uint32_t> want_d((size_t)world * W), want_c((size_t)world * W);
    std::vector<unsigned long long> cnt((size_t)world * 3, 0ull);
    for (int q = 0; q < world; ++q)
      for (uint32_t pod = 0; pod < P; ++pod) {
        const uint32_t b = idle_bits(seed, q, n, pod);
        const bool veto = sc.with_veto && veto_bits_of(seed, q, n, pod) != 0;
        if (veto && (out_v[q][(size_t)n * W + pod / 32] >> (pod & 31) & 1u) == 0) bad++;
        if (!veto && sc.with_veto && (out_v[q][(size_t)n * W + pod / 32] >> (pod & 31) & 1u) != 0) bad++;
        const bool cand = b != 0 && !veto, dec = cand && ranks[q]->eligible[pod];
        if (cand) want_c[(size_t)q * W + pod / 32] |= 1u << (pod & 31), cnt[q * 3 + 0] += __builtin_popcount(b), cnt[q * 3 + 1]++;
        if (dec) want_d[(size_t)q * W + pod / 32] |= 1u << (pod & 31), cnt[q * 3 + 2]++;
      }
    for (int r = 0; r < world; ++r) {
      for (int k = 0; k < 3; ++k)
        if (counts[r][(size_t)n * 3 + k] != cnt[(size_t)r * 3 + k]) bad++;
      if (sc.shared_output && n != K) continue;
      const size_t o = sc.shared_output ? 0 : (size_t)n * world * W;
      if (world > 1) {
        if (memcmp(out_d[r].data() + o, want_d.data(), want_d.size() * 4)) bad++;
        if (memcmp(out_c[r].data() + o, want_c.data(), want_c.size() * 4)) bad++;
      }
    }
  }
  for (int r = 0; r < world; ++r) {
    Rank& me = *ranks[r];
    if (me.err) bad++;
    for (int s = 0; s < 2; ++s) {
      if (me.done[s] != me.uses[s] || me.tickets[s] != 0) bad++;
      for (uint32_t m : me.masks[s]) if (m) { bad++; break; }   // idle and veto planes
    }
    for (unsigned long long a : me.acc) if (a) bad++;
    // completion stamps are in launch order: outputs of decision n are written after those of decision n - 1
    for (int n = 2; n <= K; ++n)
      if (stamps[r][(size_t)n * 5] < stamps[r][(size_t)(n - 1) * 5]) bad++;
  }
  return bad;
}

int main(int argc, char** argv) {
  const Scenario all[] = {
      {3, 12, 64, 10, 0, 0, -1, 0}, {3, 12, 64, 10, 1, 0, -1, 0}, {3, 12, 64, 10, 2, 0, -1, 0},   // the three protocols
      {3, 12, 64, 11, 0, 1, -1, 0}, {3, 12, 64, 11, 1, 1, -1, 0},   // shared output buffer: the last decision stays
      {2, 40, 128, 8, 1, 0, -1, 0}, {8, 4, 64, 8, 0, 0, -1, 0},  {8, 4, 64, 9, 1, 1, -1, 0},      // 2 and 8 ranks
      {4, 9, 32, 8, 1, 0, -1, 0},   {1, 20, 64, 6, 0, 0, -1, 0},    // odd word count; a single GPU (no exchange)
      // one rank's collector lags: the others run ahead as far as the scratch sets let them and overwrite its
      // slots — which is only safe because the exchange buffers are four deep (with two, this case times out)
      {3, 6, 64, 14, 1, 0, 1, 0},   {3, 6, 64, 14, 0, 0, 2, 0},
      // power veto + veto_bits requested: both tagged-slot protocols and one GPU
      {3, 12, 64, 9, 1, 1, -1, 1},  {3, 12, 64, 9, 0, 0, -1, 1}, {1, 20, 64, 6, 0, 0, -1, 1},
  };
  const int only = argc > 1 ? atoi(argv[1]) : -1;
  int bad = 0, i = 0;
  for (const Scenario& sc : all) {
    if (only >= 0 && only != i++) continue;
    const int b = run(sc, 0x5EED0000ull + 17u * (unsigned)sc.world + (unsigned)sc.protocol);
    printf("world %d words %u threads %u steps %d protocol %d shared %d slow rank %d: %s\n", sc.world, sc.words_per_rank,
           sc.fold_threads, sc.steps, sc.protocol, sc.shared_output, sc.slow_rank, b ? "FAIL" : "ok");
    if (sc.with_veto) printf("  (with power veto)\n");
    fflush(stdout);
    bad += b;
  }
  printf("%s\n", bad ? "FAIL" : "ALL OK");
  return bad ? 1 : 0;
}
