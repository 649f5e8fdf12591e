// Host emulation of the whole decision path: the three reduce kernels (k_reduce_ldg, k_reduce_tma, k_reduce_u8) and
// the fold kernel, compiled from the SOURCE TEXT of gpu-pruner_b200/csrc/gpr_kernels.cuh.
//
// tests/test_hotpath_emul.py cuts the body of `namespace gpr` out of the kernel header, removes the small helper
// functions that are nothing but inline PTX (loads with cache hints, mbarrier / bulk-copy instructions, scoped
// atomics, %globaltimer) and rewrites the three inline-PTX statements and the __shared__ declarations inside the
// kernels to shim calls -> hotpath_extract.inc.  This file supplies those helpers and the CUDA built-ins on top of
// std::thread: a CTA is a group of threads with a barrier, a warp 32 of them with emulated shuffles / ballots /
// reductions, a bulk copy is a memcpy that completes an emulated mbarrier phase.
//
// For every case directory given on the command line (files written by the test: util.f32, [power.f32], [elig.u8],
// [created.i64], [util.u8], params.txt) it runs each applicable reduce variant + the fold, twice on the same
// scratch set, and prints one line per variant:   <dir> <variant> <dbits hex> <cbits hex> <vbits hex> <n_series>
// <n_cand> <n_dec> <smax hex...>.  The test compares them with the hand-derived known answers and with the oracle.
// It validates the SOURCE logic of the kernels (row tails, NaN rules, ANY-GPU fold, veto, gates, counters, scratch
// reuse) — not the generated machine code; tests/test_gpu_parity.py does that on a B200.
#include "cuda_shim.hpp"

#define __host__
namespace gpr {
#include "hotpath_extract.inc"
#include "synth_extract.inc"   // gpr_synth.cuh: the device generator of the synthetic windows (no PTX in it)
}

// ---- driver --------------------------------------------------------------------------------------------------------
template <class T>
static bool slurp(const std::string& path, std::vector<T>* out) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  f.seekg(0, std::ios::end);
  const size_t n = (size_t)f.tellg();
  f.seekg(0);
  out->resize(n / sizeof(T));
  f.read(reinterpret_cast<char*>(out->data()), (std::streamsize)(out->size() * sizeof(T)));
  return true;
}

struct Case {
  uint32_t P = 0, G = 0, T = 0;
  uint64_t ld = 0;
  int use_power = 0;
  uint32_t thr_bits = 0;
  int64_t cutoff = 0;
  std::vector<float> util, power;
  std::vector<uint8_t> elig, util_u8;
  std::vector<int64_t> created;
};

static void print_words(const std::vector<uint32_t>& w) {
  for (uint32_t x : w) printf("%08x", x);
  if (w.empty()) printf("-");
}

static void run_variant(const std::string& dir, const char* name, const Case& c, int variant, size_t shift_floats) {
  const uint32_t P = c.P, G = c.G, T = c.T, S = P * G, MW = (G + 31) / 32, W = (P + 31) / 32;
  // the planes, 16-byte aligned plus an optional shift (the vectorised kernels peel to alignment themselves)
  std::vector<float> ubuf((size_t)S * c.ld + 16 + shift_floats), pbuf(c.use_power ? (size_t)S * c.ld + 16 + shift_floats : 0);
  auto aligned = [&](std::vector<float>& b) {
    float* p = b.data();
    while (reinterpret_cast<uintptr_t>(p) % 16u) ++p;
    return p + shift_floats;
  };
  float* util = aligned(ubuf);
  memcpy(util, c.util.data(), c.util.size() * 4);
  float* power = nullptr;
  if (c.use_power) power = aligned(pbuf), memcpy(power, c.power.data(), c.power.size() * 4);
  std::vector<uint32_t> masks((size_t)2 * P * MW + 16, 0u);
  unsigned long long acc[3] = {0, 0, 0}, done = 0, other_done = 0;
  unsigned int ticket = 0, err = 0;
  for (int rep = 0; rep < 2; ++rep) {   // the second decision reuses the scratch set the first one must have zeroed
    std::vector<uint32_t> dbits(W, 0xdeadbeefu), cbits(W, 0xdeadbeefu), vbits(W, 0xdeadbeefu);
    std::vector<float> smax(S, -12345.f);
    unsigned long long counts[3] = {~0ull, ~0ull, ~0ull};
    gpr::ReduceParams rp;
    memset(&rp, 0, sizeof rp);
    rp.seg[0] = gpr::Segment{variant == 2 ? reinterpret_cast<const float*>(c.util_u8.data()) : util, masks.data(),
                             smax.data(), S, 0u};
    rp.seg[1] = gpr::Segment{power, masks.data() + (size_t)P * MW, nullptr, c.use_power ? S : 0u, 1u};
    rp.ld = c.ld, rp.T = T, rp.G = G, rp.mw = MW;
    rp.total_rows = S + (c.use_power ? S : 0u);
    memcpy(&rp.thr, &c.thr_bits, 4);
    rp.done = &done, rp.need = (unsigned long long)rep;
    rp.util_u8 = variant == 2 ? 1u : 0u;
    const unsigned grid_r = 3;
    if (variant == 0) {
      launch(grid_r, 2 * 32, 0, [&] { gpr::k_reduce_ldg<2, 2>(rp); });
    } else if (variant == 1) {
      gpr::TmaLayout L;
      L.chunk_elems = 64, L.stage_bytes = 256, L.depth = 2, L.n_chunks = (T + 63) / 64;   // several chunks per row
      launch(grid_r, 4 * 32, (size_t)4 * L.depth * L.stage_bytes + 4 * L.depth * 8, [&] { gpr::k_reduce_tma<4>(rp, L); });
    } else {
      launch(grid_r, 2 * 32, 0, [&] { gpr::k_reduce_u8<2, 2>(rp); });
    }
    gpr::FoldParams fp;
    memset(&fp, 0, sizeof fp);
    fp.idle_mask = masks.data();
    fp.veto_mask = c.use_power ? masks.data() + (size_t)P * MW : nullptr;
    fp.eligible = c.elig.empty() ? nullptr : c.elig.data();
    fp.created = c.created.empty() ? nullptr : c.created.data();
    fp.cutoff = c.cutoff;
    fp.dbits = dbits.data(), fp.cbits = cbits.data(), fp.vbits = vbits.data();
    fp.counts = counts, fp.acc = acc, fp.ticket = &ticket;
    fp.done = &done, fp.need = (unsigned long long)rep;
    fp.prev_done = &other_done, fp.prev_need = 0;
    fp.P = P, fp.G = G, fp.mw = MW;
    fp.world = 1, fp.rank = 0;
    fp.err = &err;
    const unsigned fold_threads = 64, fold_warps = 2;
    launch(std::max(1u, (W + 4 * fold_warps - 1) / (4 * fold_warps)), fold_threads, 0, [&] { gpr::k_fold<false>(fp); });
    bool clean = done == (unsigned long long)rep + 1 && ticket == 0 && !acc[0] && !acc[1] && !acc[2];
    for (uint32_t m : masks) clean = clean && m == 0;
    printf("%s %s%s %s ", dir.c_str(), name, rep ? "#2" : "", clean ? "clean" : "DIRTY");
    print_words(dbits), printf(" "), print_words(cbits), printf(" "), print_words(vbits);
    printf(" %llu %llu %llu ", counts[0], counts[1], counts[2]);
    for (float v : smax) printf("%08x", f2u(v));
    printf("\n");
  }
}

int main(int argc, char** argv) {
  // --synth SEED P G T OUT_PREFIX : run the device generator's source, write <prefix>.util.f32 / .power.f32 / .elig.u8
  if (argc == 7 && std::string(argv[1]) == "--synth") {
    const uint64_t seed = strtoull(argv[2], nullptr, 0);
    const uint32_t P = (uint32_t)atoi(argv[3]), G = (uint32_t)atoi(argv[4]), T = (uint32_t)atoi(argv[5]);
    const std::string prefix = argv[6];
    std::vector<float> buf((size_t)P * G * T);
    for (int plane = 0; plane < 2; ++plane) {
      launch(5, 64, 0, [&] { gpr::k_synth_fill(buf.data(), seed, plane, 0, P * G, T, T); });
      std::ofstream(prefix + (plane ? ".power.f32" : ".util.f32"), std::ios::binary)
          .write(reinterpret_cast<const char*>(buf.data()), (std::streamsize)(buf.size() * 4));
    }
    std::vector<uint8_t> e(P);
    launch(2, 64, 0, [&] { gpr::k_synth_eligible(e.data(), seed, 0, P); });
    std::ofstream(prefix + ".elig.u8", std::ios::binary).write(reinterpret_cast<const char*>(e.data()), (std::streamsize)e.size());
    return 0;
  }
  for (int a = 1; a < argc; ++a) {
    const std::string dir = argv[a];
    Case c;
    {
      std::ifstream f(dir + "/params.txt");
      unsigned long long ld;
      f >> c.P >> c.G >> c.T >> ld >> c.use_power >> c.thr_bits >> c.cutoff;
      c.ld = ld;
    }
    slurp(dir + "/util.f32", &c.util);
    if (c.use_power) slurp(dir + "/power.f32", &c.power);
    slurp(dir + "/elig.u8", &c.elig);
    slurp(dir + "/created.i64", &c.created);
    const bool has_u8 = slurp(dir + "/util.u8", &c.util_u8);
    if (c.P == 0) continue;
    run_variant(dir, "ldg", c, 0, 0);
    run_variant(dir, "ldg+1", c, 0, 1);   // rows start 4 bytes off 16-byte alignment
    if (c.T % 4 == 0 && c.ld % 4 == 0) run_variant(dir, "tma", c, 1, 0);
    if (has_u8) run_variant(dir, "u8", c, 2, 0);
    fflush(stdout);
  }
  return 0;
}
