// gpr_text_kernels.cuh — the launch side of the device ingest: kernels around the host/device parser
// core in gpr_text.cuh (which tests/cpp/text_emul.cpp runs on the CPU, byte for byte the same code).
// Byte/integer work bound by instruction issue, not HBM: the text is read once (scan: one 128-bit load
// per 16 bytes; parse: TMA bulk copies of 4 KB tiles into shared memory), the tensor written once.
#pragma once
#include <cuda_runtime.h>

#include "gpr_kernels.cuh"  // mbarrier / cp.async.bulk helpers shared with the reduce kernel
#include "gpr_text.cuh"

namespace gpr {
namespace text {

struct ScanSink {
  uint64_t* opens;
  uint64_t* closes;
  unsigned long long* counts;  // [2]
  uint64_t cap;
  __device__ __forceinline__ void values_open(uint64_t p) {
    const unsigned long long i = atomicAdd(counts + 0, 1ull);
    if (i < cap) opens[i] = p;
  }
  __device__ __forceinline__ void values_close(uint64_t p) {
    const unsigned long long i = atomicAdd(counts + 1, 1ull);
    if (i < cap) closes[i] = p;
  }
};

// scans the 16-byte slices [slice_begin, slice_end) of the text
__global__ void __launch_bounds__(256) k_text_scan(const uint8_t* __restrict__ t, uint64_t n, uint64_t slice_begin,
                                                   uint64_t slice_end, uint64_t* opens, uint64_t* closes,
                                                   unsigned long long* counts, uint64_t cap) {
  ScanSink sink{opens, closes, counts, cap};
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t slice = slice_begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; slice < slice_end;
       slice += stride)
    scan_slice(t, n, slice, sink);
}

// NaN-aware max into a cell that starts as kFillBits (0xFFFFFFFF).  Non-negative values (sign bit clear:
// every DCGM sample, +Inf, the clamped denormal) order like their bit patterns as signed ints and the
// fill is -1, so the merge is a single fire-and-forget integer max at L2.  Negative values take the
// compare-and-swap loop.
__device__ __forceinline__ void atomic_merge(float* cell, float v) {
  if (v != v) return;  // a NaN sample never replaces anything
  const unsigned int bits = __float_as_uint(v);
  if ((bits >> 31) == 0u) {
    atomicMax(reinterpret_cast<int*>(cell), (int)bits);
    return;
  }
  unsigned int* a = reinterpret_cast<unsigned int*>(cell);
  unsigned int old = *a;
  while (true) {
    const float c = __uint_as_float(old);
    if (!(c != c) && !(c < v)) return;  // present and not smaller: keep
    const unsigned int seen = atomicCAS(a, old, bits);
    if (seen == old) return;
    old = seen;
  }
}

// shared-memory tile as a byte source for the parser core
struct SmemBytes {
  const uint8_t* p;
  __device__ __forceinline__ uint32_t operator[](uint32_t i) const { return p[i]; }
};

struct ParseSink {
  float* plane;
  Span* spans;
  uint32_t cur, c_in, c_oow, c_tiny;  // counts of the span this lane is in (flushed when it changes)
  __device__ __forceinline__ void put(uint64_t cell, float v) { atomic_merge(plane + cell, v); }
  __device__ __forceinline__ void hard(uint32_t s) { atomicOr(&spans[s].flags, kSpanHard); }
  __device__ __forceinline__ void flush() {
    if (cur != 0xffffffffu) {
      if (c_in) atomicAdd(&spans[cur].n_in, c_in);
      if (c_oow) atomicAdd(&spans[cur].n_oow, c_oow);
      if (c_tiny) atomicAdd(&spans[cur].n_tiny, c_tiny);
    }
    c_in = c_oow = c_tiny = 0;
  }
  __device__ __forceinline__ void count(uint32_t s, uint32_t a, uint32_t b, uint32_t c) {
    if (s != cur) flush(), cur = s;
    c_in += a, c_oow += b, c_tiny += c;
  }
};

constexpr uint32_t kStageBytes = kTileBytes + kTileHalo;  // 4192, a multiple of 16
constexpr uint32_t kListCap = 512;                        // '[' offsets listed per pass (a 4 KB tile of samples has ~240)

// One warp = one private two-stage TMA pipeline over tiles w, w + n_warps, ...  Per tile:
//   1. every lane inspects 16-byte pieces (conflict-free LDS.128) and the warp compacts the offsets of all
//      '[' bytes into a shared list (prefix sum by shuffles);
//   2. rounds of 32: lane l parses candidate 32 r + l — one sample per lane, all lanes in the same digit
//      loops at the same time;
//   3. the stage is re-armed with the tile two steps ahead.
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_text_parse(const uint8_t* __restrict__ t, uint64_t n, uint64_t n_copyable,
                                                           Span* spans, uint32_t n_spans, Grid g, float* plane) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31;
  const uint32_t w = threadIdx.x >> 5;
  unsigned char* stage0 = smem + (size_t)w * 2 * kStageBytes;
  uint16_t* list = reinterpret_cast<uint16_t*>(smem + (size_t)WARPS * 2 * kStageBytes) + (size_t)w * kListCap;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)WARPS * 2 * kStageBytes + (size_t)WARPS * kListCap * 2) + w * 2;

  const uint64_t n_tiles = (n + kTileBytes - 1) / kTileBytes;
  const uint64_t n_warps = (uint64_t)gridDim.x * WARPS;
  const uint64_t first = (uint64_t)blockIdx.x * WARPS + w;
  if (lane == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const uint64_t pol = l2_evict_first_policy();
  auto issue = [&](uint64_t tile, uint32_t st) {
    if (lane == 0 && tile < n_tiles) {
      const uint64_t off = tile * kTileBytes;
      const uint64_t room = (n_copyable - off) & ~15ull;
      const uint32_t bytes = room < kStageBytes ? (uint32_t)room : kStageBytes;
      mbar_expect_tx(&full[st], bytes);
      tma_load_1d(stage0 + (size_t)st * kStageBytes, t + off, bytes, &full[st], pol);
    }
  };
  issue(first, 0);
  issue(first + n_warps, 1);

  ParseSink sink{plane, spans, 0xffffffffu, 0u, 0u, 0u};
  uint32_t st = 0, ph = 0;
  for (uint64_t tile = first; tile < n_tiles; tile += n_warps) {
    mbar_wait(&full[st], ph);
    const unsigned char* base = stage0 + (size_t)st * kStageBytes;
    const uint64_t tile_off = tile * kTileBytes;
    const uint32_t tile_n = n - tile_off < kTileBytes ? (uint32_t)(n - tile_off) : kTileBytes;
    uint32_t s0 = 0;
    if (lane == 0) s0 = find_span(spans, n_spans, tile_off);
    s0 = __shfl_sync(0xffffffffu, s0, 0);
    const bool any_span = s0 < n_spans && __ldg(&spans[s0].begin) < tile_off + tile_n;
    if (any_span) {
      // ---- 1. the '[' bytes of the tile -------------------------------------------------------------------
      uint32_t masks[kTileBytes / 512];  // piece j of this lane = bytes (32 j + lane) * 16 ...: bit k = '[' at byte k
      uint32_t total = 0;
#pragma unroll
      for (uint32_t j = 0; j < kTileBytes / 512; ++j) {
        const uint32_t off = (j * 32u + lane) * 16u;
        const uint4 x = *reinterpret_cast<const uint4*>(base + off);
        uint32_t m = 0;
        const uint32_t ws[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t eq = __vcmpeq4(ws[k], 0x5B5B5B5Bu);  // 0xFF per byte equal to '['
          m |= ((eq & 1u) | ((eq >> 7) & 2u) | ((eq >> 14) & 4u) | ((eq >> 21) & 8u)) << (4 * k);
        }
        if (off + 16u > tile_n) m &= off < tile_n ? (1u << (tile_n - off)) - 1u : 0u;
        masks[j] = m;
        total += __reduce_add_sync(0xffffffffu, __popc(m));
      }
      // ---- 2. passes of at most kListCap candidates, rounds of 32 ---------------------------------------------
      uint32_t s = s0;
      for (uint32_t pass0 = 0; pass0 < total; pass0 += kListCap) {
        uint32_t ord = 0;  // ordinal of the next candidate of piece j, warp-wide
#pragma unroll
        for (uint32_t j = 0; j < kTileBytes / 512; ++j) {
          uint32_t m = masks[j];
          const uint32_t cnt = __popc(m);
          uint32_t incl = cnt;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += up;
          }
          uint32_t k = ord + incl - cnt;
          ord += __shfl_sync(0xffffffffu, incl, 31);
          while (m) {
            const uint32_t b = __ffs(m) - 1;
            m &= m - 1;
            if (k >= pass0 && k - pass0 < kListCap) list[k - pass0] = (uint16_t)((j * 32u + lane) * 16u + b);
            ++k;
          }
        }
        __syncwarp();
        const uint32_t n_list = min(kListCap, total - pass0);
        const SmemBytes src{base};
        for (uint32_t r = lane; r < n_list; r += 32) parse_candidate(src, tile_off, (uint32_t)list[r], spans, n_spans, s, g, sink);
        __syncwarp();  // the list is rewritten by the next pass (candidates are listed in offset order, so the
                       // span cursor only ever moves forward)
      }
    }
    __syncwarp();  // every lane is done with the stage: it may be overwritten
    issue(tile + 2 * n_warps, st);
    if (++st == 2) st = 0, ph ^= 1u;
  }
  // per-span statistics: one atomic per warp when all lanes ended in the same span
  const uint32_t c0 = __shfl_sync(0xffffffffu, sink.cur, 0);
  if (__all_sync(0xffffffffu, sink.cur == c0)) {
    const uint32_t a = __reduce_add_sync(0xffffffffu, sink.c_in);
    const uint32_t b = __reduce_add_sync(0xffffffffu, sink.c_oow);
    const uint32_t c = __reduce_add_sync(0xffffffffu, sink.c_tiny);
    if (lane == 0 && c0 != 0xffffffffu) {
      if (a) atomicAdd(&spans[c0].n_in, a);
      if (b) atomicAdd(&spans[c0].n_oow, b);
      if (c) atomicAdd(&spans[c0].n_tiny, c);
    }
  } else {
    sink.flush();
  }
}

template <int WARPS>
constexpr size_t text_parse_smem() {
  return (size_t)WARPS * 2 * kStageBytes + (size_t)WARPS * kListCap * 2 + (size_t)WARPS * 2 * sizeof(uint64_t);
}

// NaN-fill `n_cols` columns starting at ring position `col0` (wrapping at T) of every row: the new buckets
// of a resident window before a tick's samples are merged in
__global__ void __launch_bounds__(256) k_fill_columns(float* __restrict__ plane, uint32_t n_rows, uint32_t T,
                                                      uint64_t ld, uint32_t col0, uint32_t n_cols) {
  const uint64_t total = (uint64_t)n_rows * n_cols;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const uint64_t r = i / n_cols;
    uint32_t c = col0 + (uint32_t)(i - r * n_cols);
    if (c >= T) c -= T;
    reinterpret_cast<uint32_t*>(plane)[r * ld + c] = kFillBits;
  }
}

}  // namespace text
}  // namespace gpr
