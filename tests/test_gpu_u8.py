"""GPU parity of the biased-byte window format (GPR_FMT_U8B: 0 = no sample, b = value + 1).

The u8 window is decoded to f32 on the CPU and handed to the same oracles as every other test:
bitmaps, counts and series_max must be identical to what the f32 window gives.
"""
import numpy as np
import pytest

import kat

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def eng():
    import gpu_pruner_b200 as g
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device; the engine has no CPU fallback")
    e = g.IdleEngine(device=0, max_pods=12000, max_gpus=8, max_samples=2048, power_plane=True)
    yield e
    e.close()


def _random_u8(rng, P, G, T):
    """biased bytes: mostly idle/absent, some active, the extremes 1 (value 0) and 255 (value 254)"""
    b = rng.choice(np.array([1, 0, 2, 51, 101, 255], np.uint8), size=(P, G, T), p=[.85, .09, .02, .02, .01, .01])
    idle = rng.random((P, G)) < 0.5
    b[idle] = np.where(rng.random((int(idle.sum()), T)) < 0.05, 0, 1).astype(np.uint8)
    absent = rng.random((P, G)) < 0.05
    b[absent] = 0
    burst = np.flatnonzero(rng.random(P) < 0.2)
    b[burst, rng.integers(0, G, burst.size), rng.integers(0, T, burst.size)] = 2
    return b


def _gates(rng, P):
    return dict(eligible=(rng.random(P) < 0.9).astype(np.uint8),
                created_ts=rng.integers(1000, 2000, P).astype(np.int64), cutoff_ts=1500)


def _power(rng, P, G, T):
    w = rng.choice(np.array([40, 60, 149.99, np.nan], np.float32), size=(P, G, T), p=[.5, .44, .02, .04])
    hot = np.flatnonzero(rng.random(P) < 0.4)
    w[hot, rng.integers(0, G, hot.size), rng.integers(0, T, hot.size)] = rng.choice(
        np.array([150, 150.01, 400], np.float32), size=hot.size)
    return w


def _check(bits, cbits, counts, exp, smax=None):
    assert np.array_equal(bits, exp["decision_bits"]), "decision bitmap differs from oracle"
    assert np.array_equal(cbits, exp["candidate_bits"]), "candidate bitmap differs from oracle"
    assert counts == (exp["n_series"], exp["n_candidates"], exp["n_decisions"])
    if smax is not None:
        assert kat.smax_equal(smax, exp["series_max"])


def _device_decide(eng, b_t, P, G, T, power=None, gates=None, thr=0.0, stride=0, want_smax=True):
    from gpu_pruner_b200 import ffi
    dev = "cuda:0"
    gates = gates or {}
    w_t = torch.from_numpy(np.ascontiguousarray(power)).to(dev) if power is not None else None
    e_t = torch.from_numpy(gates["eligible"]).to(dev) if "eligible" in gates else None
    c_t = torch.from_numpy(gates["created_ts"]).to(dev) if "created_ts" in gates else None
    W = max((P + 31) // 32, 1)
    db = torch.full((W,), 0x7BADBEEF, dtype=torch.int32, device=dev)
    cb = torch.full((W,), 0x7BADBEEF, dtype=torch.int32, device=dev)
    sm = torch.full((max(P * G, 1),), -777.0, dtype=torch.float32, device=dev) if want_smax else None
    torch.cuda.synchronize()
    r = eng.decide_ptr(b_t, P, G, T, db, power=w_t, eligible=e_t, created_ts=c_t,
                       cutoff_ts=gates.get("cutoff_ts", 0), power_threshold=thr, candidate_bits=cb,
                       series_max=sm, row_stride=stride, util_format=ffi.GPR_FMT_U8B)
    W = (P + 31) // 32
    return (db.cpu().numpy().view(np.uint32)[:W], cb.cpu().numpy().view(np.uint32)[:W],
            (r.n_series, r.n_candidates, r.n_decisions),
            sm.cpu().numpy()[: P * G].reshape(P, G) if want_smax else None)


SHAPES = [(1, 1, 1), (3, 2, 5), (31, 4, 15), (31, 4, 33), (64, 1, 450), (257, 8, 100), (1000, 4, 180),
          (999, 3, 1801), (4097, 4, 64), (50, 4, 7200), (20, 2, 9001), (6, 1, 20000)]


@pytest.mark.parametrize("P,G,T", SHAPES)
@pytest.mark.parametrize("opts", [(False, True), (True, True), (True, False)])
def test_random_device_window(P, G, T, opts, eng, oracle_c, oracle_np):
    from gpu_pruner_b200 import from_biased_u8
    with_power, want_smax = opts
    rng = np.random.default_rng(P * 31 + G * 7 + T + 5)
    b = _random_u8(rng, P, G, T)
    u = from_biased_u8(b)
    kw = _gates(rng, P) if with_power else {}
    power = _power(rng, P, G, T) if with_power else None
    okw = dict(kw)
    if with_power:
        okw.update(power=power, power_threshold=150.0)
    exp = oracle_c.decide(u, **okw)
    assert np.array_equal(exp["decision_bits"], oracle_np.decide(u, **okw)["decision_bits"])
    bits, cbits, counts, smax = _device_decide(eng, torch.from_numpy(b).to("cuda:0"), P, G, T, power, kw,
                                               150.0 if with_power else 0.0, want_smax=want_smax)
    _check(bits, cbits, counts, exp, smax)


@pytest.mark.parametrize("P,G,T", [(3, 2, 5), (257, 8, 100), (1000, 4, 180), (999, 3, 1801), (300, 4, 2048)])
def test_random_host_window(P, G, T, eng, oracle_c):
    from gpu_pruner_b200 import from_biased_u8
    rng = np.random.default_rng(P + T + 11)
    b = _random_u8(rng, P, G, T)
    kw = _gates(rng, P)
    power = _power(rng, P, G, T)
    exp = oracle_c.decide(from_biased_u8(b), power=power, power_threshold=150.0, **kw)
    d = eng.decide(b, power, kw["eligible"], kw["created_ts"], kw["cutoff_ts"], 150.0, want_series_max=True)
    _check(d.decision_bits, d.candidate_bits, (d.n_series, d.n_candidates, d.n_decisions), exp, d.series_max)
    # and with no power plane / no series_max (the OR-only fast path)
    exp = oracle_c.decide(from_biased_u8(b))
    d = eng.decide(b)
    _check(d.decision_bits, d.candidate_bits, (d.n_series, d.n_candidates, d.n_decisions), exp)


@pytest.mark.parametrize("T,stride,offset", [(100, 104, 0), (100, 101, 0), (97, 97, 1), (64, 64, 3), (15, 15, 5),
                                             (16, 17, 15), (1800, 1800, 2), (1800, 1816, 9), (33, 40, 1),
                                             (4097, 4099, 7)])
@pytest.mark.parametrize("want_smax", [False, True])
def test_strided_and_misaligned_rows(T, stride, offset, want_smax, eng, oracle_c):
    """rows starting at any byte, row_stride > T: every sample read exactly once, no neighbour's"""
    from gpu_pruner_b200 import from_biased_u8
    P, G = 130, 4
    rng = np.random.default_rng(T * 7 + stride + offset)
    b = _random_u8(rng, P, G, T)
    buf = np.full(offset + P * G * stride + 32, 200, np.uint8)   # poison: reading it flips verdicts
    view = buf[offset: offset + P * G * stride].reshape(P * G, stride)
    view[:, :T] = b.reshape(P * G, T)
    t = torch.from_numpy(buf).to("cuda:0")
    exp = oracle_c.decide(from_biased_u8(b))
    bits, cbits, counts, smax = _device_decide(eng, t.data_ptr() + offset, P, G, T, stride=stride,
                                               want_smax=want_smax)
    _check(bits, cbits, counts, exp, smax)


def test_kats_in_both_formats(eng):
    """every known-answer vector that is representable as bytes gives the same answer in both formats"""
    from gpu_pruner_b200 import to_biased_u8
    n = 0
    for k in kat.all_kats():
        try:
            b = to_biased_u8(k.util)
        except ValueError:
            continue
        n += 1
        d = eng.decide(b, k.power, k.eligible, k.created_ts, k.cutoff_ts, k.power_threshold,
                       want_series_max=True)
        assert np.array_equal(d.candidate_bits, kat.expected_bits(k.candidate)), k.why
        assert np.array_equal(d.decision_bits, kat.expected_bits(k.decision)), k.why
        if k.series_max is not None:
            assert kat.smax_equal(d.series_max, k.series_max)
    assert n >= 5


@pytest.mark.parametrize("power", [False, True])
def test_config_c2_full_parity(power, eng, oracle_c):
    """BASELINE config #2 at full size, window re-encoded as bytes on the device"""
    seed, P, G, T = 0x5EED0002, 10000, 4, 1800
    u = torch.empty((P, G, T), dtype=torch.float32, device="cuda:0")
    eng.synth_fill(seed, 0, u, 0, P, G, T)
    w = None
    if power:
        w = torch.empty((P, G, T), dtype=torch.float32, device="cuda:0")
        eng.synth_fill(seed, 1, w, 0, P, G, T)
    e = torch.empty(P, dtype=torch.uint8, device="cuda:0")
    eng.synth_eligible(seed, e, 0, P)
    torch.cuda.synchronize()
    b = torch.where(torch.isnan(u), torch.zeros_like(u), u + 1).to(torch.uint8)
    exp = oracle_c.decide_synth(seed, 0, P, G, T, use_power=power, power_threshold=150.0, use_elig=True)
    bits, cbits, counts, _ = _device_decide(eng, b, P, G, T, None if w is None else w.cpu().numpy(),
                                            {"eligible": e.cpu().numpy()}, 150.0 if power else 0.0,
                                            want_smax=False)
    _check(bits, cbits, counts, exp)
    d = eng.decide(b.cpu().numpy(), None if w is None else w.cpu().numpy(), e.cpu().numpy(),
                   power_threshold=150.0 if power else 0.0)
    _check(d.decision_bits, d.candidate_bits, (d.n_series, d.n_candidates, d.n_decisions), exp)


def test_formats_interleaved_in_one_batch(eng, oracle_c):
    """f32 and u8 decisions back to back in one enqueue: the launch chain (PDL, alternating scratch
    sets) must hold across the two reduce kernels"""
    from gpu_pruner_b200 import ffi, from_biased_u8
    P, G, T = 2048, 4, 512
    rng = np.random.default_rng(77)
    dev = "cuda:0"
    calls, exps, keep = [], [], []
    for i in range(12):
        b = _random_u8(rng, P, G, T)
        u = from_biased_u8(b)
        exps.append(oracle_c.decide(u))
        t = torch.from_numpy(b if i % 2 else u).to(dev)
        db = torch.zeros((P + 31) // 32, dtype=torch.int32, device=dev)
        cb = torch.zeros((P + 31) // 32, dtype=torch.int32, device=dev)
        keep.append((t, db, cb))
        calls.append(dict(util=t, P=P, G=G, T=T, decision_bits=db, candidate_bits=cb,
                          util_format=ffi.GPR_FMT_U8B if i % 2 else ffi.GPR_FMT_F32))
    torch.cuda.synchronize()
    batch = eng.make_batch(calls)
    for _ in range(3):
        ress = eng.decide_batch_async(batch)
        eng.sync()
        for (t, db, cb), r, exp in zip(keep, ress, exps):
            _check(db.cpu().numpy().view(np.uint32), cb.cpu().numpy().view(np.uint32),
                   (r.n_series, r.n_candidates, r.n_decisions), exp)


def test_bad_format_is_rejected(eng):
    import gpu_pruner_b200 as g
    t = torch.zeros(64, dtype=torch.uint8, device="cuda:0")
    db = torch.zeros(1, dtype=torch.int32, device="cuda:0")
    torch.cuda.synchronize()
    with pytest.raises(g.GprError) as ei:
        eng.decide_ptr(t, 4, 4, 4, db, util_format=7)
    assert ei.value.code == g.ffi.GPR_E_INVALID
