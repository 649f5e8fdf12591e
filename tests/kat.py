"""Known-answer vectors K1..K14 for the idle decision.

The reference pins no results for this path (its tests assert on query text only,
/root/reference/gpu-pruner/src/main.rs:572-740), so each vector below is derived by hand from
one line of the PromQL template /root/reference/gpu-pruner/src/query.promql.j2 or of the Rust
post-processing in main.rs, and the expected verdict is written out explicitly — it is NOT
computed by either oracle.  Both oracles and the CUDA path must reproduce every one.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np

NAN = np.float32(np.nan)
T0 = 37  # deliberately not a multiple of 4 or 32


@dataclass
class Kat:
    name: str
    why: str                      # the template / source line the expectation comes from
    util: np.ndarray              # [P, G, T] float32
    candidate: list               # expected candidate(p)
    decision: list                # expected decision(p)
    power: Optional[np.ndarray] = None
    power_threshold: Optional[float] = 0.0
    eligible: Optional[np.ndarray] = None
    created_ts: Optional[np.ndarray] = None
    cutoff_ts: int = 0
    series_max: Optional[np.ndarray] = None   # expected [P, G] (NaN-aware) when given
    n_series: Optional[int] = None
    extra: dict = field(default_factory=dict)


def _rows(*rows, T=T0):
    """one pod per row, G = 1"""
    out = np.zeros((len(rows), 1, T), dtype=np.float32)
    for i, r in enumerate(rows):
        out[i, 0, :] = r
    return out


def _row(T=T0, fill=0.0, **at):
    r = np.full(T, fill, dtype=np.float32)
    for k, v in at.items():
        r[int(k[1:])] = v
    return r


def all_kats() -> list:
    K = []

    # K1: all-zero row => idle  (`== 0`, query.promql.j2:35)
    K.append(Kat("K1_all_zero", "query.promql.j2:35", _rows(_row()), [1], [1],
                 series_max=np.array([[0.0]], np.float32), n_series=1))

    # K2: a single 1 anywhere => active: every element must be read, incl. head/tail lanes
    for T in (1, 2, 3, 4, 5, 31, 32, 33, 37, 127, 128, 129, 450, 1800, 1801, 1803):
        idxs = sorted({0, T - 1, T // 2, min(T - 1, 33), min(T - 1, 5)})
        rows = [_row(T)] + [_row(T, **{f"i{i}": 1.0}) for i in idxs]
        K.append(Kat(f"K2_single_one_T{T}", "max_over_time reads every sample (query.promql.j2:16)",
                     _rows(*rows, T=T), [1] + [0] * len(idxs), [1] + [0] * len(idxs)))

    # K3: -0.0 everywhere => idle (IEEE -0.0 == 0)
    K.append(Kat("K3_negative_zero", "query.promql.j2:35", _rows(_row(fill=-0.0)), [1], [1]))

    # K4: zeros + missing steps => idle (missing steps are not samples)
    r = _row()
    r[::3] = NAN
    K.append(Kat("K4_zeros_and_gaps", "max_over_time over present samples", _rows(r), [1], [1],
                 series_max=np.array([[0.0]], np.float32)))

    # K5: no sample in the window => series absent => never idle
    K.append(Kat("K5_all_missing", "empty range vector yields no element", _rows(_row(fill=NAN)),
                 [0], [0], series_max=np.array([[np.nan]], np.float32), n_series=0))

    # K6: gap then positive / positive then gap => active
    a = _row(fill=NAN, i20=7.0)
    b = _row(fill=0.0, i0=3.0)
    b[1:] = NAN
    K.append(Kat("K6_gap_and_positive", "max_over_time", _rows(a, b), [0, 0], [0, 0],
                 series_max=np.array([[7.0], [3.0]], np.float32)))

    # K7: smallest normal and a denormal are != 0 (no flush-to-zero)
    K.append(Kat("K7_denormal", "SURVEY §7: denormals must not be flushed",
                 _rows(_row(i5=np.float32(1.17549435e-38)), _row(i36=np.float32(1e-45))),
                 [0, 0], [0, 0]))

    # K8: negative-only row: max < 0, `== 0` fails
    K.append(Kat("K8_negative_only", "query.promql.j2:35", _rows(_row(fill=-1.0), _row(fill=-0.5, i3=0.0)),
                 [0, 1], [0, 1], series_max=np.array([[-1.0], [0.0]], np.float32)))

    # K9: ANY-GPU fold (main.rs:416-437): one idle GPU of four is enough
    T = T0
    idle, act, absent = _row(T), _row(T, i9=50.0), _row(T, fill=NAN)
    u = np.stack([
        np.stack([idle, act, act, act]),
        np.stack([act, act, act, act]),
        np.stack([absent, absent, absent, idle]),
        np.stack([absent, absent, absent, absent]),
        np.stack([idle, idle, idle, idle]),
    ]).astype(np.float32)
    K.append(Kat("K9_any_gpu", "main.rs:430-435 HashSet dedup keeps a pod with >= 1 series", u,
                 [1, 0, 1, 0, 1], [1, 0, 1, 0, 1], n_series=1 + 1 + 4))

    # K10: power veto, `>=` and pod-wide `unless on (pod, namespace)` (query.promql.j2:36-44)
    thr = 150.0
    below = np.nextafter(np.float32(150.0), np.float32(0.0))
    above = np.nextafter(np.float32(150.0), np.float32(1e9))
    u = np.zeros((6, 2, T), np.float32)               # every series idle
    w = np.full((6, 2, T), 50.0, np.float32)
    w[0, 0, 7] = below                                  # T - eps      -> candidate
    w[1, 0, 7] = 150.0                                  # == T         -> veto (>=)
    w[2, 0, 7] = above                                  # T + eps      -> veto
    w[3, 1, 11] = 400.0                                 # other GPU of the same pod -> veto pod
    w[4, :, :] = NAN                                    # power series absent -> no veto
    u[5, 1, :] = 30.0                                   # pod 5: gpu1 active, gpu0 idle, low power
    K.append(Kat("K10_power_veto", "query.promql.j2:37,42", u, [1, 0, 0, 0, 1, 1], [1, 0, 0, 0, 1, 1],
                 power=w, power_threshold=thr, n_series=2 + 2 + 1))
    # threshold unset / 0.0 => clause absent (Jinja truthiness, query.promql.j2:36)
    K.append(Kat("K10b_power_threshold_zero", "query.promql.j2:36", u, [1] * 6, [1] * 6,
                 power=w, power_threshold=0.0))
    K.append(Kat("K10c_power_threshold_none", "query.promql.j2:36", u, [1] * 6, [1] * 6,
                 power=w, power_threshold=None))
    # a fractional threshold that is not representable in f32
    w2 = np.full((3, 1, T), 10.0, np.float32)
    t_frac = 100.1
    w2[0, 0, 0] = np.float32(100.1)                       # f32(100.1) = 100.09999847 < 100.1
    w2[1, 0, 0] = np.nextafter(np.float32(100.1), np.float32(1e9))  # first f32 above
    K.append(Kat("K10d_power_fraction", "f64 compare semantics of `>= T`", np.zeros((3, 1, T), np.float32),
                 [1, 0, 1], [1, 0, 1], power=w2, power_threshold=t_frac))

    # K11: age / phase gate (main.rs:473-510): created >= cutoff => skip; Pending => skip
    cutoff = 1_700_000_000_000_000_000
    created = np.array([cutoff, cutoff - 1, cutoff + 5, cutoff - 10**9, np.iinfo(np.int64).max],
                       dtype=np.int64)
    elig = np.array([1, 1, 1, 0, 1], dtype=np.uint8)    # pod 3 is Pending
    K.append(Kat("K11_age_phase_gate", "main.rs:508 `create_time >= lookback_start`",
                 np.zeros((5, 1, T), np.float32), [1] * 5, [0, 1, 0, 0, 0],
                 eligible=elig, created_ts=created, cutoff_ts=cutoff))

    # K12: shapes around the packing: P < 32, P not multiple of 32, G in {1, 4, 8}; padding bits 0
    for P, G in ((1, 1), (31, 4), (33, 8), (64, 1), (65, 4), (100, 4)):
        u = np.zeros((P, G, 8), np.float32)
        u[::2] = 5.0                                     # even pods fully active
        cand = [int(p % 2 == 1) for p in range(P)]
        K.append(Kat(f"K12_pack_P{P}_G{G}", "bitmap layout", u, cand, cand))

    # K13: reported value: UTIL max 37 -> 0.37 after `/ 100` (query.promql.j2:20)
    K.append(Kat("K13_value_report", "query.promql.j2:20", _rows(_row(i4=37.0, i30=12.0)), [0], [0],
                 series_max=np.array([[37.0]], np.float32), extra={"value": 0.37}))

    # +inf sample: max = inf, not idle; vetoes for any finite threshold
    K.append(Kat("Kx_infinity", "IEEE compare", _rows(_row(i2=np.float32(np.inf))), [0], [0]))
    return K


def expected_bits(flags) -> np.ndarray:
    flags = np.asarray(flags, dtype=bool)
    W = (flags.size + 31) // 32
    out = np.zeros(W, dtype=np.uint32)
    for p, f in enumerate(flags):
        if f:
            out[p >> 5] |= np.uint32(1 << (p & 31))
    return out


def smax_equal(a, b) -> bool:
    """numeric equality (so -0.0 == +0.0) with NaN matching NaN"""
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    return a.shape == b.shape and bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))
