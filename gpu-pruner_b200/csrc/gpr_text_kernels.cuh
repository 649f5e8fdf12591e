// gpr_text_kernels.cuh — the launch side of the device ingest: thin kernels around the host/device
// parser core in gpr_text.cuh (which tests/cpp/text_emul.cpp runs on the CPU, byte for byte the same
// code).  Both passes are flat over the text: thread i owns bytes [i*S, (i+1)*S).  Byte/integer work
// bound by instruction issue and L1/L2, not HBM: the text is read once (scan: one 128-bit load per
// 16 bytes; parse: byte loads through L1), the tensor written once.
#pragma once
#include <cuda_runtime.h>

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

__global__ void __launch_bounds__(256) k_text_scan(const uint8_t* __restrict__ t, uint64_t n,
                                                   uint64_t* opens, uint64_t* closes,
                                                   unsigned long long* counts, uint64_t cap) {
  ScanSink sink{opens, closes, counts, cap};
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t slice = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; slice * kScanBytes < n; slice += stride)
    scan_slice(t, n, slice, sink);
}

// NaN-aware max, the merge of duplicate series of one `sum by` group (ingest.cpp merge_cell)
__device__ __forceinline__ void atomic_merge(float* cell, float v) {
  if (v != v) return;  // a NaN never replaces anything (the cell starts as NaN)
  unsigned int* a = reinterpret_cast<unsigned int*>(cell);
  unsigned int old = *a;
  while (true) {
    const float c = __uint_as_float(old);
    if (!(c != c) && !(c < v)) return;  // present and not smaller: keep
    const unsigned int seen = atomicCAS(a, old, __float_as_uint(v));
    if (seen == old) return;
    old = seen;
  }
}

struct ParseSink {
  float* plane;
  Span* spans;
  uint32_t T;
  // counts of the first span this thread touched stay in registers (warp-aggregated by the kernel)
  uint32_t first, c_in, c_oow, c_tiny;
  __device__ __forceinline__ void store(uint32_t row, uint32_t col, float v) {
    plane[(size_t)row * T + col] = v;
  }
  __device__ __forceinline__ void merge(uint32_t row, uint32_t col, float v) {
    atomic_merge(plane + (size_t)row * T + col, v);
  }
  __device__ __forceinline__ void hard(uint32_t s) { atomicOr(&spans[s].flags, kSpanHard); }
  __device__ __forceinline__ void count(uint32_t s, uint32_t a, uint32_t b, uint32_t c) {
    if (first == 0xffffffffu) {
      first = s, c_in = a, c_oow = b, c_tiny = c;
      return;
    }
    if (a) atomicAdd(&spans[s].n_in, a);
    if (b) atomicAdd(&spans[s].n_oow, b);
    if (c) atomicAdd(&spans[s].n_tiny, c);
  }
};

__global__ void __launch_bounds__(256) k_text_parse(const uint8_t* __restrict__ t, uint64_t n, Span* spans,
                                                    uint32_t n_spans, Grid g, float* plane) {
  // every thread of the grid reaches the warp collectives below (no early return)
  const uint64_t slice = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  ParseSink sink{plane, spans, g.T, 0xffffffffu, 0u, 0u, 0u};
  // one binary search per warp; lanes walk on from there (a warp covers 4 KB of text)
  uint32_t hint = 0;
  if (lane == 0) hint = find_span(spans, n_spans, (slice - lane) * kParseBytes);
  hint = __shfl_sync(0xffffffffu, hint, 0);
  parse_slice(t, n, spans, n_spans, hint, slice, g, sink);
  __syncwarp();
  const uint32_t s0 = __shfl_sync(0xffffffffu, sink.first, 0);
  const bool uniform = __all_sync(0xffffffffu, sink.first == s0);
  if (uniform) {
    if (s0 != 0xffffffffu) {
      const uint32_t a = __reduce_add_sync(0xffffffffu, sink.c_in);
      const uint32_t b = __reduce_add_sync(0xffffffffu, sink.c_oow);
      const uint32_t c = __reduce_add_sync(0xffffffffu, sink.c_tiny);
      if (lane == 0) {
        if (a) atomicAdd(&spans[s0].n_in, a);
        if (b) atomicAdd(&spans[s0].n_oow, b);
        if (c) atomicAdd(&spans[s0].n_tiny, c);
      }
    }
  } else if (sink.first != 0xffffffffu) {
    if (sink.c_in) atomicAdd(&spans[sink.first].n_in, sink.c_in);
    if (sink.c_oow) atomicAdd(&spans[sink.first].n_oow, sink.c_oow);
    if (sink.c_tiny) atomicAdd(&spans[sink.first].n_tiny, sink.c_tiny);
  }
}

__global__ void __launch_bounds__(256) k_text_fill_nan(uint4* __restrict__ p, uint64_t n_vec) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint4 v = make_uint4(0x7fc00000u, 0x7fc00000u, 0x7fc00000u, 0x7fc00000u);
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) p[i] = v;
}

}  // namespace text
}  // namespace gpr
