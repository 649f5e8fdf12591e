// gpr_text.cuh — device-side parse of a Prometheus range-query response (matrix JSON) straight into
// the dense (pod x gpu x t) f32 tensor in HBM.
//
// Wire shape: the one /root/reference/gpu-pruner/src/bin/querytest.rs:41-53 walks — per series a label
// map and a list of [<unix time>, "<value>"] pairs:
//     {"metric":{...},"values":[[1700000000,"0"],[1700000001,"37"],...]}
// More than 99 % of the bytes are the sample lists; the label maps (hashing, string work, ~1 % of the
// bytes) stay on the host (gpu-pruner_b200/host/ingest_device.cpp), which turns them into one
// `Span` per series: where its sample list sits in the text and which tensor row it feeds.
//
// Two passes over the text, both flat over bytes (a thread owns a fixed slice of the text, not a series,
// so the work is balanced whatever the series lengths are):
//   scan  : report the offsets of `},"values":[` and `"]]` — the two byte patterns that delimit a
//           sample list.  Neither can occur inside a JSON string (a raw '"' ends the string), so in the
//           compact encoding Prometheus emits they are exact; the host cross-checks every series and
//           falls back to the CPU parser for anything else (pretty-printed JSON, histograms, ...).
//   parse : every '[' inside a span starts a sample; the thread whose slice holds the '[' parses the
//           sample (reading past its slice if need be) and stores the value at (row, column(ts)).
//
// Strictness instead of generality: a sample that is not exactly `[digits[.digits],"number|NaN|±Inf"]`
// with a number the exact decimal fast path can convert (<= 2^53 mantissa, |exp10| <= 22 — every DCGM
// integer and every short decimal), two samples of one series that land in the same column, or
// timestamps that run backwards, mark the whole SPAN `hard`; the host re-parses the rows of hard spans
// with the CPU text parser and overwrites them.  So the tensor is bit-identical to the CPU ingest
// (gpu-pruner_b200/host/ingest.cpp) for every input, and malformed input raises the same errors.
//
// Everything that decides a byte's meaning is in GPR_HD functions that also compile as plain C++:
// tests/cpp/text_emul.cpp runs the very same code thread by thread on the CPU against the CPU ingest.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define GPR_HD __host__ __device__ __forceinline__
#else
#define GPR_HD inline
#endif

namespace gpr {
namespace text {

constexpr uint32_t kSpanShared = 1u;  // in : several series feed this row -> merge, do not store
constexpr uint32_t kSpanHard = 2u;    // out: the host must re-parse this span's row
constexpr uint32_t kScanBytes = 16;   // bytes per thread, scan pass
constexpr uint32_t kParseBytes = 128; // bytes per thread, parse pass
constexpr uint32_t kTextPad = 256;    // zero bytes the caller guarantees after the text

struct Span {        // mirrors gpr_text_span (include/gpr.h)
  uint64_t begin;    // offset of the first byte after `"values":[`
  uint64_t end;      // offset of the ']' that closes the list
  uint32_t row;      // destination row = pod * G + slot
  uint32_t flags;    // kSpan*
  uint32_t n_in;     // out: samples parsed
  uint32_t n_oow;    // out: samples outside (t_end - N, t_end]
  uint32_t n_tiny;   // out: values below the f32 denormal range, clamped to +-denorm_min
  uint32_t reserved;
};

struct Grid {        // time axis of the window (gpu-pruner_b200/host/ingest.cpp column_of)
  int64_t t_end;     // seconds; newest column
  int64_t step;      // seconds per column, > 0
  uint32_t T;        // columns
  uint32_t pad;
};

// ---- numbers -------------------------------------------------------------------------------------------
// exact powers of ten: 10^0 .. 10^22 are representable in binary64
GPR_HD double pow10_exact(int e) {
  // a switch (not a table) keeps this usable from host and device without a __constant__ copy
  switch (e) {
    case 0: return 1e0;   case 1: return 1e1;   case 2: return 1e2;   case 3: return 1e3;
    case 4: return 1e4;   case 5: return 1e5;   case 6: return 1e6;   case 7: return 1e7;
    case 8: return 1e8;   case 9: return 1e9;   case 10: return 1e10; case 11: return 1e11;
    case 12: return 1e12; case 13: return 1e13; case 14: return 1e14; case 15: return 1e15;
    case 16: return 1e16; case 17: return 1e17; case 18: return 1e18; case 19: return 1e19;
    case 20: return 1e20; case 21: return 1e21; default: return 1e22;
  }
}

// [-+]digits[.digits][(e|E)[-+]digits] at t[p...]; on success *v is the correctly rounded binary64
// (one exact int->double conversion, one IEEE multiply or divide: Clinger's fast path) and the
// position after the number is returned.  0 = not convertible this way (caller marks the span hard).
GPR_HD uint64_t parse_decimal(const uint8_t* __restrict__ t, uint64_t p, double* v) {
  bool neg = false;
  if (t[p] == '-' || t[p] == '+') neg = t[p] == '-', ++p;
  uint64_t m = 0;
  int nd = 0, frac = 0;
  bool any = false;
  for (; t[p] >= '0' && t[p] <= '9'; ++p) {
    any = true;
    if (m > 900719925474099ull) return 0;  // next digit could pass 2^53
    m = m * 10 + (uint64_t)(t[p] - '0'), ++nd;
  }
  if (!any) return 0;
  if (t[p] == '.') {
    ++p;
    bool anyf = false;
    for (; t[p] >= '0' && t[p] <= '9'; ++p) {
      anyf = true;
      if (m > 900719925474099ull) return 0;
      m = m * 10 + (uint64_t)(t[p] - '0'), ++frac;
      if (frac > 22) return 0;
    }
    if (!anyf) return 0;
  }
  int e10 = -frac;
  if (t[p] == 'e' || t[p] == 'E') {
    ++p;
    bool eneg = false;
    if (t[p] == '-' || t[p] == '+') eneg = t[p] == '-', ++p;
    int ex = 0;
    bool anye = false;
    for (; t[p] >= '0' && t[p] <= '9'; ++p) {
      anye = true;
      ex = ex * 10 + (int)(t[p] - '0');
      if (ex > 400) return 0;
    }
    if (!anye) return 0;
    e10 += eneg ? -ex : ex;
  }
  if (m > 9007199254740992ull) return 0;
  double d = (double)m;  // exact
  if (m != 0) {
    if (e10 < -22 || e10 > 22) return 0;
    if (e10 < 0) d = d / pow10_exact(-e10);
    else if (e10 > 0) d = d * pow10_exact(e10);
  }
  *v = neg ? -d : d;
  return p;
}

// what gph::to_f32 does (ingest.cpp): a non-zero value that rounds to 0 in f32 stays non-zero
GPR_HD float to_f32(double x, uint32_t* tiny) {
  float f = (float)x;
  if (x != 0.0 && f == 0.0f && x == x) {
#if defined(__CUDA_ARCH__)
    f = __int_as_float(x < 0 ? 0x80000001 : 0x00000001);
#else
    union { uint32_t u; float f; } c;
    c.u = x < 0 ? 0x80000001u : 0x00000001u;
    f = c.f;
#endif
    ++*tiny;
  }
  return f;
}

GPR_HD float quiet_nan_f32() {
#if defined(__CUDA_ARCH__)
  return __int_as_float(0x7fc00000);
#else
  union { uint32_t u; float f; } c;
  c.u = 0x7fc00000u;
  return c.f;
#endif
}

GPR_HD float inf_f32(bool neg) {
#if defined(__CUDA_ARCH__)
  return __int_as_float(neg ? 0xff800000 : 0x7f800000);
#else
  union { uint32_t u; float f; } c;
  c.u = neg ? 0xff800000u : 0x7f800000u;
  return c.f;
#endif
}

// timestamp -> whole seconds (ingest.cpp ts_seconds): garbage maps far outside any window
constexpr int64_t kBadTs = INT64_MIN / 4;
GPR_HD int64_t ts_seconds(double t) {
  if (!(t > -4e12 && t < 4e12)) return kBadTs;
  return (int64_t)llround(t);
}

// column of ts, or -1 outside (t_end - N, t_end]   (ingest.cpp column_of)
GPR_HD int64_t column_of(const Grid& g, int64_t ts) {
  if (ts > g.t_end || ts < g.t_end - (int64_t)g.T * g.step - g.step) return -1;
  const int64_t back = (g.t_end - ts + g.step / 2) / g.step;
  if (back < 0 || back >= (int64_t)g.T) return -1;
  return (int64_t)g.T - 1 - back;
}

// ---- one sample ------------------------------------------------------------------------------------------
// t[p] == '['.  Strict compact form only.  Returns the offset after the closing ']' or 0 (hard).
GPR_HD uint64_t parse_timestamp(const uint8_t* __restrict__ t, uint64_t p, int64_t* ts) {
  double d;
  const uint64_t q = parse_decimal(t, p + 1, &d);
  if (q == 0 || t[p + 1] == '+' || t[q] != ',') return 0;
  *ts = ts_seconds(d);
  return q;
}

GPR_HD uint64_t parse_sample(const uint8_t* __restrict__ t, uint64_t p, int64_t* ts, float* val,
                             uint32_t* tiny) {
  uint64_t q = parse_timestamp(t, p, ts);
  if (q == 0 || t[q + 1] != '"') return 0;
  q += 2;  // past ,"
  if (t[q] == 'N') {
    if (t[q + 1] != 'a' || t[q + 2] != 'N') return 0;
    *val = quiet_nan_f32();
    q += 3;
  } else if (t[q] == 'I' || ((t[q] == '+' || t[q] == '-') && t[q + 1] == 'I')) {
    const bool neg = t[q] == '-';
    if (t[q] != 'I') ++q;
    if (t[q + 1] != 'n' || t[q + 2] != 'f') return 0;
    *val = inf_f32(neg);
    q += 3;
  } else {
    double d;
    const uint64_t r = parse_decimal(t, q, &d);
    if (r == 0) return 0;
    *val = to_f32(d, tiny);
    q = r;
  }
  if (t[q] != '"' || t[q + 1] != ']') return 0;
  return q + 2;
}

// ---- spans -------------------------------------------------------------------------------------------------
// index of the first span whose end lies beyond `pos` (spans sorted by begin, non-overlapping)
GPR_HD uint32_t find_span(const Span* __restrict__ spans, uint32_t n, uint64_t pos) {
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (spans[mid].end > pos) hi = mid;
    else lo = mid + 1;
  }
  return lo;
}

// ---- the scan pass: one slice of kScanBytes ----------------------------------------------------------------
// Sink: void values_open(uint64_t pos_of_closing_brace), void values_close(uint64_t pos_of_quote)
template <typename Sink>
GPR_HD void scan_slice(const uint8_t* __restrict__ t, uint64_t n, uint64_t slice, Sink& sink) {
  const uint64_t b = slice * kScanBytes;
  if (b >= n) return;
  // the slice's 16 bytes plus 4 of look-ahead live in registers (the text is padded, so reading
  // past n is safe; `t` is 16-byte aligned and slices are 16 bytes, so the vector load is aligned)
  uint32_t w[5];
#if defined(__CUDA_ARCH__)
  const uint4 x = *reinterpret_cast<const uint4*>(t + b);
  w[0] = x.x, w[1] = x.y, w[2] = x.z, w[3] = x.w;
  w[4] = *reinterpret_cast<const uint32_t*>(t + b + 16);
#else
  memcpy(w, t + b, sizeof w);
#endif
#define GPR_TEXT_BYTE(k) ((w[(k) >> 2] >> (((k)&3) * 8)) & 0xffu)
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < (int)kScanBytes; ++k) {
    const uint64_t p = b + (uint64_t)k;
    if (p >= n) break;
    const uint32_t c = GPR_TEXT_BYTE(k);
    if (c == '"') {
      if (GPR_TEXT_BYTE(k + 1) == ']' && GPR_TEXT_BYTE(k + 2) == ']') sink.values_close(p);
    } else if (c == '}') {
      // },"values":[   (rare: once per series, plus the '}' that end label maps)
      if (t[p + 1] == ',' && t[p + 2] == '"' && t[p + 3] == 'v' && t[p + 4] == 'a' && t[p + 5] == 'l' &&
          t[p + 6] == 'u' && t[p + 7] == 'e' && t[p + 8] == 's' && t[p + 9] == '"' && t[p + 10] == ':' &&
          t[p + 11] == '[')
        sink.values_open(p);
    }
  }
#undef GPR_TEXT_BYTE
}

// ---- the parse pass: one slice of kParseBytes ----------------------------------------------------------------
// Sink: void store(uint32_t row, uint32_t col, float v)      sole writer of the row
//       void merge(uint32_t row, uint32_t col, float v)      row shared by several series
//       void hard(uint32_t span)
//       void count(uint32_t span, uint32_t n_in, uint32_t n_oow, uint32_t n_tiny)
// `s` = find_span(spans, n_spans, slice begin) (the kernel computes it once per warp and walks on).
template <typename Sink>
GPR_HD void parse_slice(const uint8_t* __restrict__ t, uint64_t n, const Span* __restrict__ spans,
                        uint32_t n_spans, uint32_t s, uint64_t slice, const Grid& g, Sink& sink) {
  const uint64_t c0 = slice * kParseBytes;
  if (c0 >= n) return;
  const uint64_t c1 = c0 + kParseBytes < n ? c0 + kParseBytes : n;
  while (s < n_spans && spans[s].end <= c0) ++s;
  for (; s < n_spans && spans[s].begin < c1; ++s) {
    const uint64_t sb = spans[s].begin, se = spans[s].end;
    const uint32_t row = spans[s].row;
    const bool shared = (spans[s].flags & kSpanShared) != 0;
    uint64_t p = sb > c0 ? sb : c0;
    const uint64_t pe = se < c1 ? se : c1;
    uint32_t n_in = 0, n_oow = 0, n_tiny = 0;
    bool is_hard = false;
    int64_t prev_ts = kBadTs, prev_col = -1;
    bool have_prev = false;
    for (; p < pe; ++p) {
      if (t[p] != '[') continue;
      if (!have_prev && p > sb) {
        // the sample before this slice's first one (owned by another thread): needed to see
        // column collisions and backwards time across the slice boundary
        uint64_t q = p - 1;
        const uint64_t stop = (p - sb > 96) ? p - 96 : sb;
        while (q > stop && t[q] != '[') --q;
        if (t[q] == '[' && q >= sb && parse_timestamp(t, q, &prev_ts) != 0) {
          prev_col = column_of(g, prev_ts);
        } else {
          is_hard = true;
        }
      }
      have_prev = true;
      int64_t ts;
      float v;
      const uint64_t q = parse_sample(t, p, &ts, &v, &n_tiny);
      if (q == 0 || q > se) {
        is_hard = true;
        break;
      }
      // between samples exactly one ',' ; after the last one the list closes at `se`
      if (!((t[q] == ',' && t[q + 1] == '[') || q == se)) {
        is_hard = true;
        break;
      }
      ++n_in;
      const int64_t col = column_of(g, ts);
      if (prev_ts != kBadTs && ts < prev_ts) is_hard = true;            // time runs backwards
      if (col >= 0 && col == prev_col) is_hard = true;                  // two samples, one cell: merge on the host
      if (ts != kBadTs) prev_ts = ts;
      if (col < 0) {
        ++n_oow;
      } else {
        prev_col = col;
        if (shared) sink.merge(row, (uint32_t)col, v);
        else sink.store(row, (uint32_t)col, v);
      }
      p = q - 1;  // continue after this sample (the for's ++p lands on ',' or the closing ']')
    }
    if (is_hard) sink.hard(s);
    sink.count(s, n_in, n_oow, n_tiny);
  }
}

}  // namespace text
}  // namespace gpr
