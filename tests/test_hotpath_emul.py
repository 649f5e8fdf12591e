"""The whole decision path run on the CPU from the kernels' SOURCE TEXT: k_reduce_ldg, k_reduce_tma, k_reduce_u8 and
k_fold of gpu-pruner_b200/csrc/gpr_kernels.cuh, compiled under a host shim (tests/cpp/hotpath_emul.cpp: CTAs, warps
with emulated shuffles, bulk copies that complete emulated mbarrier phases) and compared with
  * the hand-derived known answers K1..K14 of tests/kat.py (each traced to a line of query.promql.j2 / main.rs), and
  * the numpy oracle on random windows (ragged shapes, row strides, > 32 series per pod, power veto, gates).
The -m gpu suite checks the compiled kernels on a B200; this keeps the arithmetic of SURVEY.md §8(a) a2, a7, a8,
a10, a11, a12 under test on a machine without a GPU.  The product still has no CPU path: this shim lives in tests/."""
import os
import re
import subprocess

import numpy as np
import pytest

import kat as KAT

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "gpu-pruner_b200", "csrc", "gpr_kernels.cuh")

# helper functions of the header that are nothing but inline PTX: the shim provides them
PTX_HELPERS = ["gtime", "pdl_launch_dependents", "pdl_wait_prior_grids", "ld_acquire_u64", "spin_until_gpu",
               "st_release_u64", "ld_acquire_sys_u64", "st_release_sys_u64", "ld_relaxed_sys_u64", "spin_until_sys",
               "ldg_stream", "ldg_stream_u4", "smem_u32", "mbar_init", "mbar_expect_tx", "mbar_arrive", "mbar_wait",
               "tma_load_1d", "l2_evict_first_policy"]
REWRITES = [
    ('asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(v) : "memory");', "st_relaxed_sys_u64(dst, v);", 1),
    ('asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");', ";", 1),
    ("__shared__ unsigned long long s_cnt[3];", "unsigned long long* s_cnt = tl_cta->s_cnt;", 1),
    ("__shared__ unsigned int s_last;", "unsigned int& s_last = tl_cta->s_last;", 1),
    ("__shared__ unsigned int s_next;", "unsigned int& s_next = tl_cta->s_next;", 2),
    ("extern __shared__ __align__(128) unsigned char smem[];", "unsigned char* smem = tl_cta->smem;", 1),
]


def _drop_function(src, name):
    m = re.search(r"^__device__ __forceinline__ [^\n(]*\b%s\(" % re.escape(name), src, flags=re.M)
    assert m, name
    i = src.index("{", m.end())
    depth = 0
    while True:
        depth += {"{": 1, "}": -1}.get(src[i], 0)
        i += 1
        if depth == 0:
            break
    return src[:m.start()] + src[i:]


def _extract():
    src = open(HDR).read()
    body = src[src.index("namespace gpr {") + len("namespace gpr {"):src.rindex("}  // namespace gpr")]
    # (the developer-only timeline instrumentation is compiled out; drop its text, it is inline PTX too)
    body = re.sub(r"#ifdef GPR_TIMELINE.*?#else\n(#define TL_MARK\(slot\)\n)#endif", r"\1", body, flags=re.S)
    for name in PTX_HELPERS:
        body = _drop_function(body, name)
    for old, new, count in REWRITES:
        assert body.count(old) == count, (old, body.count(old))
        body = body.replace(old, new)
    assert "asm" not in body and "__shared__" not in body
    return body


def _extract_synth():
    src = open(os.path.join(ROOT, "gpu-pruner_b200", "csrc", "gpr_synth.cuh")).read()
    body = src[src.index("namespace gpr {") + len("namespace gpr {"):src.rindex("}  // namespace gpr")]
    assert "asm" not in body
    return body


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    d = tmp_path_factory.mktemp("hotpath")
    (d / "hotpath_extract.inc").write_text(_extract())
    (d / "synth_extract.inc").write_text(_extract_synth())
    exe = d / "hotpath_emul"
    subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-Wall", "-Wno-unknown-pragmas", "-Wno-unused-function",
                    "-I", str(d), os.path.join(ROOT, "tests", "cpp", "hotpath_emul.cpp"), "-o", str(exe)],
                   check=True, capture_output=True, text=True)
    return str(exe)


def _thr_bits(thr):
    t = np.float32(thr)
    if float(t) < float(thr):
        t = np.nextafter(t, np.float32(np.inf))
    return int(np.array([t], np.float32).view(np.uint32)[0])


def _write_case(d, util, power=None, thr=0.0, eligible=None, created=None, cutoff=0, ld=None):
    util = np.asarray(util, np.float32)
    P, G, T = util.shape
    ld = ld or T
    def strided(x, fill):
        out = np.full((P * G, ld), fill, np.float32)
        out[:, :T] = x.reshape(P * G, T)
        return out
    os.makedirs(d, exist_ok=True)
    strided(util, 77.0).tofile(os.path.join(d, "util.f32"))        # the padding between rows must never be read
    use_power = power is not None and thr is not None and thr != 0.0 and not np.isnan(thr)
    if use_power:
        strided(np.asarray(power, np.float32), 1e9).tofile(os.path.join(d, "power.f32"))
    if eligible is not None:
        np.asarray(eligible, np.uint8).tofile(os.path.join(d, "elig.u8"))
    if created is not None:
        np.asarray(created, np.int64).tofile(os.path.join(d, "created.i64"))
    present = ~np.isnan(util)
    v = np.where(present, util, 0.0)
    if np.all((v >= 0) & (v <= 254) & (v == np.floor(v))):           # representable in GPR_FMT_U8B
        b = np.full((P * G, ld), 9, np.uint8)
        b[:, :T] = np.where(present, v + 1, 0).astype(np.uint8).reshape(P * G, T)
        b.tofile(os.path.join(d, "util.u8"))
    with open(os.path.join(d, "params.txt"), "w") as f:
        f.write(f"{P} {G} {T} {ld} {int(use_power)} {_thr_bits(thr) if use_power else 0} {int(cutoff)}\n")


def _parse(lines):
    out = {}
    for l in lines:
        f = l.split()
        words = lambda h: np.array([int(h[i:i + 8], 16) for i in range(0, len(h), 8)], np.uint32) if h != "-" else np.zeros(0, np.uint32)
        out.setdefault(f[0], []).append({"variant": f[1], "clean": f[2] == "clean", "d": words(f[3]), "c": words(f[4]),
                                          "v": words(f[5]), "counts": tuple(int(x) for x in f[6:9]),
                                          "smax": words(f[9]).view(np.float32) if len(f) > 9 else np.zeros(0, np.float32)})
    return out


def _run(emul, dirs):
    r = subprocess.run([emul] + [str(d) for d in dirs], capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stderr[-2000:]
    return _parse(r.stdout.splitlines())


def test_known_answers_through_every_kernel_variant(emul, tmp_path):
    kats = [k for k in KAT.all_kats() if k.util.shape[0] > 0]
    dirs = []
    for k in kats:
        d = tmp_path / k.name
        _write_case(str(d), k.util, k.power, k.power_threshold, k.eligible, k.created_ts, k.cutoff_ts)
        dirs.append(d)
    res = _run(emul, dirs)
    n_variants = 0
    for k, d in zip(kats, dirs):
        runs = res[str(d)]
        assert {r["variant"] for r in runs} >= {"ldg", "ldg#2", "ldg+1"}, k.name
        for r in runs:
            n_variants += 1
            assert r["clean"], (k.name, r["variant"])                       # scratch zeroed for the next decision
            assert np.array_equal(r["c"], KAT.expected_bits(k.candidate)), (k.name, r["variant"], k.why)
            assert np.array_equal(r["d"], KAT.expected_bits(k.decision)), (k.name, r["variant"], k.why)
            if k.series_max is not None:
                assert KAT.smax_equal(r["smax"].reshape(k.series_max.shape), k.series_max), (k.name, r["variant"])
            if k.n_series is not None:
                assert r["counts"][0] == k.n_series, (k.name, r["variant"])
            assert r["counts"][1:] == (int(np.sum(k.candidate)), int(np.sum(k.decision))), (k.name, r["variant"])
    assert n_variants >= 4 * len(kats)
    assert any(r["variant"] == "tma" for rs in res.values() for r in rs)
    assert any(r["variant"] == "u8" for rs in res.values() for r in rs)


def test_random_windows_equal_the_oracle(emul, tmp_path, oracle_np):
    rng = np.random.default_rng(20260921)
    cases = []
    for i in range(36):
        P = int(rng.integers(1, 75))
        G = int(rng.choice([1, 2, 4, 8, 33, 40]))                 # 33 / 40: more than one mask word per pod
        T = int(rng.choice([1, 3, 4, 36, 37, 64, 128, 130, 200, 452]))
        if G >= 33:
            P = min(P, 20)
        kind = rng.random((P, G, 1))
        util = np.where(kind < 0.35, 0.0, rng.integers(0, 101, (P, G, T)) * (rng.random((P, G, T)) < 0.5)).astype(np.float32)
        util[rng.random((P, G, T)) < 0.02] = np.nan                 # scrape gaps
        util[rng.random((P, G)) < 0.08] = np.nan                    # series without a sample in the window
        if i % 4 == 0:
            util[rng.random((P, G, T)) < 0.01] = -1.0               # not representable in the byte format: f32 only
        power = thr = None
        if i % 3 == 0:
            power = np.where(rng.random((P, G, T)) < 0.9, rng.uniform(40, 149, (P, G, T)), rng.uniform(150, 700, (P, G, T))).astype(np.float32)
            power[rng.random((P, G)) < 0.5] = 55.5
            thr = float(rng.choice([150.0, 149.99999, 100.5, 0.0]))
        elig = (rng.random(P) > 0.1).astype(np.uint8) if i % 2 else None
        created = rng.integers(1_700_000_000 - 5000, 1_700_000_000, P) if i % 5 < 2 else None
        cutoff = 1_700_000_000 - 2500
        ld = T if i % 2 else T + int(rng.integers(1, 9)) * 4
        if i % 6 == 5:
            ld = T + 1                                                # odd stride: rows at every alignment
        d = tmp_path / f"r{i}"
        _write_case(str(d), util, power, thr, elig, created, cutoff, ld)
        cases.append((d, util, power, thr, elig, created, cutoff))
    res = _run(emul, [c[0] for c in cases])
    seen = set()
    for d, util, power, thr, elig, created, cutoff in cases:
        want = oracle_np.decide(util, power, elig, created, cutoff, thr if thr is not None else 0.0)
        for r in res[str(d)]:
            seen.add(r["variant"].split("#")[0])
            tag = (str(d), r["variant"])
            assert r["clean"], tag
            assert np.array_equal(r["d"], want["decision_bits"]), tag
            assert np.array_equal(r["c"], want["candidate_bits"]), tag
            assert np.array_equal(r["v"], want["veto_bits"]), tag
            assert r["counts"] == (want["n_series"], want["n_candidates"], want["n_decisions"]), tag
            assert KAT.smax_equal(r["smax"].reshape(want["series_max"].shape), want["series_max"]), tag
    assert seen == {"ldg", "ldg+1", "tma", "u8"}


def test_no_data_race_under_thread_sanitizer(tmp_path):
    """reduce kernels + fold under ThreadSanitizer (see tests/test_fold_exchange_emul.py): publishing rows with atomics
    into the pod masks, the scratch-set guard, the warp-private bulk-copy rings, the fold's ticket"""
    (tmp_path / "hotpath_extract.inc").write_text(_extract())
    (tmp_path / "synth_extract.inc").write_text(_extract_synth())
    exe = tmp_path / "hotpath_emul_tsan"
    subprocess.run(["g++", "-std=c++20", "-O1", "-g", "-pthread", "-fsanitize=thread", "-Wno-unknown-pragmas",
                    "-Wno-unused-function", "-I", str(tmp_path), os.path.join(ROOT, "tests", "cpp", "hotpath_emul.cpp"),
                    "-o", str(exe)], check=True, capture_output=True, text=True)
    kats = [k for k in KAT.all_kats() if k.util.shape[0] > 0]
    pick = [k for k in kats if k.name in ("K1_all_zero", "K9_any_gpu", "K10_power_veto", "K11_age_phase_gate", "K12_pack_P65_G4")]
    pick += [k for k in kats if k.name in ("K2_single_one_T37", "K2_single_one_T128", "K2_single_one_T450")]
    assert len(pick) == 8, [k.name for k in kats]
    dirs = []
    for k in pick:
        d = tmp_path / k.name
        _write_case(str(d), k.util, k.power, k.power_threshold, k.eligible, k.created_ts, k.cutoff_ts)
        dirs.append(str(d))
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=1")
    r = subprocess.run([str(exe)] + dirs, capture_output=True, text=True, timeout=1200, env=env)
    assert r.returncode == 0 and "ThreadSanitizer" not in r.stderr, r.stderr[-3000:]
    assert len(r.stdout.splitlines()) >= 6 * len(dirs)


def test_config_1_fixture_through_the_kernel_source(emul, tmp_path, oracle_np):
    """BASELINE config #1 (100 pods x 4 GPUs x 1800 samples, the reference's own CPU-runnable case): the checked-in
    idle set tests/golden/c1_idle_set.npz, reproduced by every kernel variant's source, with and without power."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "c1_idle_set.npz"))
    seed, P, G, T = int(g["seed"]), int(g["P"]), int(g["G"]), int(g["T"])
    u = oracle_np.synth_fill(seed, 0, 0, P, G, T)
    w = oracle_np.synth_fill(seed, 1, 0, P, G, T)
    e = oracle_np.synth_eligible(seed, 0, P)
    _write_case(str(tmp_path / "c1"), u, None, 0.0, e)
    _write_case(str(tmp_path / "c1p"), u, w, float(g["power_threshold"]), e)
    res = _run(emul, [tmp_path / "c1", tmp_path / "c1p"])
    for name, dkey, nkey in (("c1", "decision_bits", "n_series"), ("c1p", "decision_bits_power", "n_series_power")):
        runs = res[str(tmp_path / name)]
        assert {r["variant"] for r in runs} >= {"ldg", "ldg+1", "tma", "u8", "tma#2"}
        for r in runs:
            assert r["clean"] and np.array_equal(r["d"], g[dkey]) and r["counts"][0] == int(g[nkey]), (name, r["variant"])
            if name == "c1":
                assert np.array_equal(r["c"], g["candidate_bits"]), r["variant"]


def test_device_generator_source_equals_the_oracle_generators(emul, tmp_path, oracle_np, oracle_c):
    """bench.py's parity check compares the decision on device-generated windows with the oracle's decision on
    oracle-generated windows: the generators (gpr_synth.cuh, oracle/gpr_oracle.c, oracle/oracle_np.py) must produce
    identical cells.  Checked on the GPU by tests/test_gpu_parity.py; here from the device generator's source."""
    for seed, P, G, T in ((0x5EED0001, 100, 4, 1800), (0x5EED0002, 37, 8, 61), (7, 5, 1, 1)):
        prefix = str(tmp_path / f"s{seed:x}")
        subprocess.run([emul, "--synth", str(seed), str(P), str(G), str(T), prefix], check=True, timeout=600)
        for plane, name in ((0, "util"), (1, "power")):
            got = np.fromfile(f"{prefix}.{name}.f32", np.float32).reshape(P, G, T)
            for orc in (oracle_np, oracle_c):
                want = orc.synth_fill(seed, plane, 0, P, G, T)
                assert np.array_equal(got.view(np.uint32) & 0x7fffffff >= 0x7f800001, np.isnan(want)), (seed, name)
                assert np.array_equal(np.nan_to_num(got, nan=-1.0), np.nan_to_num(want, nan=-1.0)), (seed, name)
        assert np.array_equal(np.fromfile(prefix + ".elig.u8", np.uint8), oracle_np.synth_eligible(seed, 0, P).astype(np.uint8))
