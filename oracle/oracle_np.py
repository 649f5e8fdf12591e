"""numpy restatement of gpu-pruner's idle decision.  TEST INFRASTRUCTURE ONLY.

Second, independent restatement of the same semantics as ``gpr_oracle.c`` (see its header for
the line-by-line map onto ``/root/reference/gpu-pruner/src/query.promql.j2:1-44`` and
``/root/reference/gpu-pruner/src/main.rs:416-437,473-510``).  PARITY UNPINNED: the reference
holds no golden vectors for this path; the two restatements are written in
different styles (sequential first-sample fold in C, NaN-ignoring ``fmax`` reduction here)
and must agree with each other and with the hand-derived known-answer vectors before any
GPU result is trusted.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this module; the product path never does.
"""
from __future__ import annotations

import numpy as np

_U64 = np.uint64
_MASK = (1 << 64) - 1

TAG_SERIES = 0x5345524945530001
TAG_CELL = 0x43454C4C00000002
TAG_POWER = 0x504F574552000003
TAG_ELIG = 0x454C494700000004


# --------------------------------------------------------------------------------------------
# decision
# --------------------------------------------------------------------------------------------
def window_max(x: np.ndarray) -> np.ndarray:
    """max_over_time along the last axis, float64, NaN = missing step.

    Prometheus ``max_over_time`` (query.promql.j2:16) skips nothing explicitly, but a step with
    no sample simply is not in the range vector; a window with no sample yields no element
    (restated as NaN).  ``np.fmax`` ignores NaN operands and returns NaN only if all are NaN.
    """
    x64 = np.asarray(x, dtype=np.float64)
    if x64.shape[-1] == 0:
        return np.full(x64.shape[:-1], np.nan)
    with np.errstate(invalid="ignore"):
        return np.fmax.reduce(x64, axis=-1)


def power_clause_enabled(power, thr) -> bool:
    """Jinja truthiness of ``args.power_threshold`` (query.promql.j2:36): None / 0.0 => absent."""
    return power is not None and thr is not None and thr != 0.0 and not np.isnan(thr)


def pack_bits(flags: np.ndarray) -> np.ndarray:
    """bool[P] -> uint32[ceil(P/32)], pod p = bit (p & 31) of word (p >> 5), padding zero."""
    flags = np.asarray(flags, dtype=bool)
    n_words = (flags.size + 31) // 32
    padded = np.zeros(n_words * 32, dtype=np.uint8)
    padded[: flags.size] = flags
    return np.packbits(padded, bitorder="little").view("<u4").copy()


def unpack_bits(words: np.ndarray, n: int) -> np.ndarray:
    b = np.unpackbits(np.ascontiguousarray(words, dtype="<u4").view(np.uint8), bitorder="little")
    return b[:n].astype(bool)


def decide(util, power=None, eligible=None, created_ts=None, cutoff_ts=0, power_threshold=0.0):
    """util/power: float32[P, G, T].  Returns dict with decision/candidate bool[P], bitmaps,
    series_max float32[P, G] and the three counts (idle series in non-vetoed pods, candidates,
    decisions)."""
    util = np.asarray(util, dtype=np.float32)
    P, G, _ = util.shape
    smax = window_max(util)                              # [P, G] float64
    idle_s = smax == 0.0                                 # `== 0` (query.promql.j2:35); NaN -> False
    veto = np.zeros(P, dtype=bool)
    if power_clause_enabled(power, power_threshold):
        wmax = window_max(np.asarray(power, dtype=np.float32))
        with np.errstate(invalid="ignore"):
            veto = (wmax >= float(power_threshold)).any(axis=1)  # unless on (pod, namespace)
    candidate = idle_s.any(axis=1) & ~veto               # main.rs:416-437 ANY-GPU dedup
    elig = np.ones(P, dtype=bool)
    if eligible is not None:
        elig &= np.asarray(eligible).astype(bool)        # Pending / no timestamp, main.rs:473-492
    if created_ts is not None:
        elig &= ~(np.asarray(created_ts, dtype=np.int64) >= np.int64(cutoff_ts))  # main.rs:508
    decision = candidate & elig
    n_series = int(idle_s[candidate].sum())
    return {
        "decision": decision,
        "candidate": candidate,
        "decision_bits": pack_bits(decision),
        "candidate_bits": pack_bits(candidate),
        "series_max": smax.astype(np.float32),
        "veto": veto,
        "veto_bits": pack_bits(veto),
        "n_series": n_series,
        "n_candidates": int(candidate.sum()),
        "n_decisions": int(decision.sum()),
    }


# --------------------------------------------------------------------------------------------
# synthetic universe
# --------------------------------------------------------------------------------------------
def mix64(x):
    """splitmix64 finaliser on uint64 arrays (wrapping arithmetic)."""
    x = np.asarray(x, dtype=_U64)
    with np.errstate(over="ignore"):
        x = x + _U64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> _U64(30))) * _U64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> _U64(27))) * _U64(0x94D049BB133111EB)
        return x ^ (x >> _U64(31))


def _scalar_mix(x: int) -> int:
    return int(mix64(np.array([x & _MASK], dtype=_U64))[0])


def synth_eligible(seed: int, pod_offset: int, n_pods: int) -> np.ndarray:
    k = _U64(_scalar_mix(seed ^ TAG_ELIG))
    pods = np.arange(pod_offset, pod_offset + n_pods, dtype=_U64)
    return ((mix64(k ^ pods) % _U64(100)) >= _U64(5)).astype(np.uint8)


def synth_fill(seed: int, plane: int, pod_offset: int, n_pods: int, n_gpus: int, n_samples: int):
    """float32[n_pods, n_gpus, n_samples] of plane 0 (util) or 1 (power)."""
    P, G, T = n_pods, n_gpus, n_samples
    s = (np.arange(pod_offset * G, (pod_offset + P) * G, dtype=_U64)).reshape(P * G, 1)
    hs = mix64(_U64(_scalar_mix(seed ^ TAG_SERIES)) ^ s)           # [S, 1]
    c = hs % _U64(100)
    a = (hs >> _U64(8)) % _U64(T)
    b = hs >> _U64(40)
    idle = c < _U64(30)
    burst = (c >= _U64(30)) & (c < _U64(40))
    active = (c >= _U64(40)) & (c < _U64(95))
    gappy = c >= _U64(95)
    tail_active = (b & _U64(1)) == _U64(1)
    t = np.arange(T, dtype=_U64).reshape(1, T)
    with np.errstate(over="ignore"):
        idx = s * _U64(T) + t                                        # [S, T]
    out = np.zeros((P * G, T), dtype=np.float32)
    if plane == 0:
        hc = mix64(_U64(_scalar_mix(seed ^ TAG_CELL)) ^ idx)
        v_active = np.where(((hc >> _U64(10)) & _U64(1)) == _U64(1),
                            (_U64(1) + (hc >> _U64(11)) % _U64(100)).astype(np.float32),
                            np.float32(0))
        burst_val = (_U64(1) + b % _U64(100)).astype(np.float32)
        out = np.where(burst & (t == a), burst_val, out)
        out = np.where(active | (gappy & tail_active), v_active, out)
        out = np.where(gappy & (t <= a), np.float32(np.nan), out)
        out = np.where(hc % _U64(1000) == _U64(0), np.float32(np.nan), out)
    else:
        hp = mix64(_U64(_scalar_mix(seed ^ TAG_POWER)) ^ idx)
        low = idle | (gappy & ~tail_active)
        r = hp >> _U64(10)
        out = np.where(low, (_U64(40) + r % _U64(31)).astype(np.float32),
                       (_U64(70) + r % _U64(631)).astype(np.float32))
        out = np.where(hp % _U64(1000) == _U64(0), np.float32(np.nan), out)
    return out.astype(np.float32).reshape(P, G, T)
