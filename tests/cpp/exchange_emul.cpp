// Host emulation of the collecting side of the fused bitmap exchange.
//
// The test (tests/test_exchange_collect_emul.py) cuts `struct FoldParams` and `exchange_bitmaps_ll` out of
// gpu-pruner_b200/csrc/gpr_kernels.cuh verbatim into exchange_extract.inc; this file supplies the few CUDA names
// that text uses (threadIdx / blockDim, __syncthreads, the system-scope load, %globaltimer) on top of std::thread and
// std::barrier, runs one "CTA" of real threads against peers that fill the tagged slots late and in random order,
// and checks what the caller would see.  It validates the SOURCE logic of the collector (indexing, stale tags,
// re-polling, ordering behind the previous fold) — not the generated machine code.
#include <atomic>
#include <barrier>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <thread>
#include <vector>

#define __device__
#define __forceinline__ inline
struct Dim3 { unsigned x; };
static thread_local Dim3 threadIdx;
static Dim3 blockDim;
static std::barrier<>* g_barrier = nullptr;
static inline void __syncthreads() { g_barrier->arrive_and_wait(); }
static inline unsigned long long gtime() {
  return (unsigned long long)std::chrono::duration_cast<std::chrono::nanoseconds>(
             std::chrono::steady_clock::now().time_since_epoch()).count();
}
static inline unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
  return __atomic_load_n(p, __ATOMIC_RELAXED);
}
static inline void __nanosleep(unsigned) { std::this_thread::yield(); }
static std::atomic<int> g_prev_waits{0};
static inline void spin_until_gpu(const unsigned long long* p, unsigned long long want) {
  g_prev_waits++;
  while (__atomic_load_n(p, __ATOMIC_ACQUIRE) < want) std::this_thread::yield();
}

namespace gpr {
#include "exchange_extract.inc"
}

struct Case { int world, rank; uint32_t n_words, pad; unsigned threads; int late; int want_c; };

static int run_case(const Case& c, uint64_t seed) {
  std::mt19937_64 rng(seed);
  const uint32_t span = 2 * c.n_words, stride = span + c.pad;
  const size_t slots = (size_t)c.world * stride;
  std::vector<unsigned long long> ll(slots);
  std::vector<uint32_t> gather(slots), out_d((size_t)c.world * c.n_words), out_c(out_d.size());
  unsigned long long stamp[5] = {0, 0, 0, 0, 0};
  unsigned int err = 0;
  unsigned long long prev_done = 0;
  int bad = 0;
  for (unsigned long long step = 5; step < 5 + 6; ++step) {   // slots still hold step - 4 (or junk) when a step begins
    std::vector<uint32_t> words(slots);
    for (auto& w : words) w = (uint32_t)rng();
    for (size_t i = 0; i < slots; ++i)
      ll[i] = (((step - 4) & 0xffffffffull) << 32) | (uint32_t)rng();   // stale tag, junk word
    for (auto& g : gather) g = (uint32_t)rng();
    for (uint32_t w = 0; w < span; ++w)   // this rank's own words were stored by the other fold CTAs
      gather[(size_t)c.rank * stride + w] = words[(size_t)c.rank * stride + w];
    std::fill(out_d.begin(), out_d.end(), 0xdeadbeefu);
    std::fill(out_c.begin(), out_c.end(), 0xdeadbeefu);
    prev_done = step - 1;   // the previous fold is NOT done yet (needs `step`)

    gpr::FoldParams f;
    memset(&f, 0, sizeof f);
    f.world = c.world, f.rank = c.rank, f.rank_stride = stride, f.step = step;
    f.P = c.n_words * 32u;
    for (int r = 0; r < c.world; ++r) f.peer_gather[r] = nullptr;
    f.peer_gather[c.rank] = gather.data();
    f.my_ll = ll.data();
    f.out_dbits = out_d.data();
    f.out_cbits = c.want_c ? out_c.data() : nullptr;
    f.stamp = stamp, f.err = &err;
    f.prev_done = &prev_done, f.prev_need = step;
    f.late_order = c.late;
    f.exchange_debug = 0;

    std::atomic<unsigned> arrived{0};
    std::vector<std::thread> peers;
    for (int r = 0; r < c.world; ++r) {
      if (r == c.rank) continue;
      peers.emplace_back([&, r, s = rng()] {
        std::mt19937_64 prng(s);
        std::vector<uint32_t> order(span);
        for (uint32_t w = 0; w < span; ++w) order[w] = w;
        std::shuffle(order.begin(), order.end(), prng);
        std::this_thread::sleep_for(std::chrono::microseconds(prng() % 3000));
        for (uint32_t k = 0; k < span; ++k) {
          const size_t i = (size_t)r * stride + order[k];
          __atomic_store_n(&ll[i], ((step & 0xffffffffull) << 32) | words[i], __ATOMIC_RELAXED);
          if ((prng() & 63) == 0) std::this_thread::sleep_for(std::chrono::microseconds(prng() % 200));
        }
        arrived++;
      });
    }
    // the previous fold finishes some time after every peer has delivered (late ordering must still hold back)
    std::thread prev([&] {
      while (arrived.load() < (unsigned)(c.world - 1)) std::this_thread::yield();
      std::this_thread::sleep_for(std::chrono::milliseconds(2));
      // nothing of this step may be in the caller's buffers before the previous fold is done
      if (c.late)
        for (uint32_t v : out_d) if (v != 0xdeadbeefu) { bad++; break; }
      __atomic_store_n(&prev_done, step, __ATOMIC_RELEASE);
    });
    if (!c.late) __atomic_store_n(&prev_done, step, __ATOMIC_RELEASE);   // in-order form: waited for before the fold

    blockDim.x = c.threads;
    std::barrier<> bar((std::ptrdiff_t)c.threads);
    g_barrier = &bar;
    g_prev_waits = 0;
    std::vector<std::thread> cta;
    for (unsigned t = 0; t < c.threads; ++t)
      cta.emplace_back([&, t] {
        threadIdx.x = t;
        gpr::exchange_bitmaps_ll(f, c.n_words, nullptr);
      });
    for (auto& t : cta) t.join();
    for (auto& t : peers) t.join();
    prev.join();

    if (err) bad++;
    if (g_prev_waits.load() != (c.late ? 1 : 0)) bad++;
    for (int r = 0; r < c.world; ++r)
      for (uint32_t w = 0; w < c.n_words; ++w) {
        const uint32_t d = words[(size_t)r * stride + w], cc = words[(size_t)r * stride + c.n_words + w];
        if (out_d[(size_t)r * c.n_words + w] != d) bad++;
        if (c.want_c ? out_c[(size_t)r * c.n_words + w] != cc : out_c[(size_t)r * c.n_words + w] != 0xdeadbeefu) bad++;
        if (gather[(size_t)r * stride + w] != d || gather[(size_t)r * stride + c.n_words + w] != cc) bad++;
      }
    if (stamp[4] < stamp[3]) bad++;
  }
  return bad;
}

int main() {
  const Case cases[] = {
      {2, 0, 313, 0, 256, 0, 1},  {2, 1, 313, 6, 256, 1, 1},   {4, 2, 313, 0, 256, 0, 1},  {8, 0, 313, 0, 256, 0, 1},
      {8, 7, 313, 2, 256, 1, 1},  {8, 3, 313, 0, 128, 1, 0},   {8, 5, 1, 0, 256, 0, 1},    {8, 4, 7, 5, 64, 1, 1},
      {3, 1, 2001, 0, 256, 0, 1}, {8, 6, 1250, 0, 256, 1, 1},  {5, 0, 33, 1, 64, 0, 0},
  };
  int bad = 0, n = 0;
  for (const Case& c : cases) {
    const int b = run_case(c, 0x5EED0000ull + (unsigned)n);
    printf("world %d rank %d words %u pad %u threads %u late %d cbits %d: %s\n", c.world, c.rank, c.n_words, c.pad,
           c.threads, c.late, c.want_c, b ? "FAIL" : "ok");
    bad += b, ++n;
  }
  printf("%s\n", bad ? "FAIL" : "ALL OK");
  return bad ? 1 : 0;
}
