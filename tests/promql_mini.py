"""A tiny label-set evaluator for the ONE PromQL expression gpu-pruner sends
(/root/reference/gpu-pruner/src/query.promql.j2:1-44).  TEST INFRASTRUCTURE.

Third, structurally different restatement: where the oracles work on a dense (pod x gpu x t)
tensor, this one works the way Prometheus does — on instant vectors of labelled elements, with
vector matching (`on`, `group_left`), set operators (`or`, `unless`), aggregation (`sum by`),
`label_replace`, scalar arithmetic and comparison filters — and evaluates the expression tree of
the template node by node.  Semantics restated from the Prometheus documentation (querying/operators,
querying/functions); nothing under /root/reference implements them (the reference ships the text to
a server, main.rs:397).  Used by tests/test_promql_semantics.py to check ingest + oracle end to end.
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass


@dataclass(frozen=True)
class Series:
    labels: tuple          # sorted (k, v) pairs, includes __name__
    samples: tuple         # ((t, v), ...), t ascending


def series(name, labels, samples):
    d = dict(labels)
    d["__name__"] = name
    return Series(tuple(sorted(d.items())), tuple(samples))


def _sig(labels: dict, names) -> tuple:
    return tuple((n, labels.get(n, "")) for n in names)


# ---- selectors ----------------------------------------------------------------------------------
def select(db, name, matchers):
    """matchers: list of (label, op, value) with op in {'!=', '=~'}; regexes are fully anchored"""
    out = []
    for s in db:
        lab = dict(s.labels)
        if lab.get("__name__") != name:
            continue
        ok = True
        for k, op, v in matchers:
            have = lab.get(k, "")
            if op == "!=":
                ok &= have != v
            elif op == "=~":
                ok &= re.fullmatch(v, have) is not None
            elif op == "=":
                ok &= have == v
        if ok:
            out.append(s)
    return out


def max_over_time(selected, t_eval, range_s):
    """instant vector {labels-without-name: value}; a series with no sample in (t-range, t] yields
    no element; start from the first sample, replace on `cur > max or isnan(max)`"""
    vec = {}
    for s in selected:
        vals = [v for (t, v) in s.samples if t_eval - range_s < t <= t_eval]
        if not vals:
            continue
        m = vals[0]
        for v in vals:
            if v > m or math.isnan(m):
                m = v
        lab = tuple((k, v) for k, v in s.labels if k != "__name__")
        assert lab not in vec, "duplicate label set in one instant vector"
        vec[lab] = m
    return vec


def instant(selected, t_eval, lookback=300):
    vec = {}
    for s in selected:
        vals = [v for (t, v) in s.samples if t_eval - lookback < t <= t_eval]
        if vals:
            vec[tuple((k, v) for k, v in s.labels if k != "__name__")] = vals[-1]
    return vec


# ---- operators ------------------------------------------------------------------------------------
def scalar_div(vec, k):
    return {lab: v / k for lab, v in vec.items()}


def v_or(a, b):
    """`a or b`: all of a, plus elements of b whose full label set is not in a"""
    out = dict(a)
    for lab, v in b.items():
        if lab not in out:
            out[lab] = v
    return out


def v_or_on(a, b, on):
    out = dict(a)
    have = {_sig(dict(lab), on) for lab in a}
    for lab, v in b.items():
        if _sig(dict(lab), on) not in have:
            out[lab] = v
    return out


def v_unless_on(a, b, on):
    drop = {_sig(dict(lab), on) for lab in b}
    return {lab: v for lab, v in a.items() if _sig(dict(lab), on) not in drop}


def sum_by(vec, by):
    groups = {}
    for lab, v in vec.items():
        key = tuple((n, dict(lab)[n]) for n in by if n in dict(lab))
        groups.setdefault(key, []).append(v)
    return {k: math.fsum(vs) if not any(math.isnan(x) for x in vs) else float("nan")
            for k, vs in groups.items()}


def label_replace(vec, dst, repl, src, regex):
    out = {}
    for lab, v in vec.items():
        d = dict(lab)
        m = re.fullmatch(regex, d.get(src, ""))
        if m:
            d[dst] = m.expand(repl.replace("$1", r"\1"))
        out[tuple(sorted(d.items()))] = v
    return out


def mul_on_group_left(lhs, rhs, on, extra):
    """lhs * on(on) group_left(extra) rhs  — many-to-one; duplicate rhs signatures are a query error"""
    idx = {}
    for lab, v in rhs.items():
        s = _sig(dict(lab), on)
        if s in idx:
            raise ValueError("many-to-many matching not allowed: duplicate series on the right side")
        idx[s] = (dict(lab), v)
    out = {}
    for lab, v in lhs.items():
        hit = idx.get(_sig(dict(lab), on))
        if hit is None:
            continue
        d = dict(lab)
        for e in extra:
            if e in hit[0]:
                d[e] = hit[0][e]
        out[tuple(sorted(d.items()))] = v * hit[1]
    return out


def filt(vec, pred):
    return {lab: v for lab, v in vec.items() if pred(v)}


# ---- the template, node by node ----------------------------------------------------------------------
def evaluate_template(db, t_eval, duration_min=30, namespace=None, model_name=None, power_threshold=None,
                      honor_labels=False):
    pl, nl, cl = ("pod", "namespace", "container") if honor_labels else (
        "exported_pod", "exported_namespace", "exported_container")
    rng = duration_min * 60
    m_compute = [(pl, "!=", "")]
    if namespace:
        m_compute.append((nl, "=~", namespace))
    m_power = list(m_compute)
    if model_name:
        m_compute.append(("modelName", "=~", model_name))
    by = ["Hostname", cl, pl, nl, "gpu", "modelName"]

    def idle_gpus():                                                        # j2:8-22
        prof = max_over_time(select(db, "DCGM_FI_PROF_GR_ENGINE_ACTIVE", m_compute), t_eval, rng)
        util = scalar_div(max_over_time(select(db, "DCGM_FI_DEV_GPU_UTIL", m_compute), t_eval, rng), 100.0)
        return sum_by(v_or(prof, util), by)

    dmi = instant(select(db, "node_dmi_info", []), t_eval)
    dmi = label_replace(label_replace(dmi, "Hostname", "$1", "instance", "(.+)"),
                        "node_type", "$1", "product_name", "(.+)")             # j2:25-30
    enriched = mul_on_group_left(idle_gpus(), dmi, ["Hostname"], ["node_type"])  # j2:24
    combined = v_or_on(enriched, idle_gpus(), by)                              # j2:32-33
    result = filt(combined, lambda v: v == 0)                                  # j2:35
    if power_threshold:                                                        # j2:36 (Jinja truthiness)
        hot = filt(max_over_time(select(db, "DCGM_FI_DEV_POWER_USAGE", m_power), t_eval, rng),
                   lambda v: v >= power_threshold)
        result = v_unless_on(result, hot, [pl, nl])                            # j2:37-43
    return result


def unique_pods(result, honor_labels=False):
    """the Rust side: exported_* first then bare (lib.rs:158-175), dedup by (pod, namespace)
    (main.rs:416-437).  Returns (n_series, ordered unique (pod, ns) list)"""
    seen, order, n = set(), [], 0
    for lab in result:
        d = dict(lab)
        pod = d.get("exported_pod", d.get("pod"))
        ns = d.get("exported_namespace", d.get("namespace"))
        ctr = d.get("exported_container", d.get("container"))
        if pod is None or ns is None or ctr is None or "modelName" not in d:
            continue                                       # PodConvertError -> skipped (main.rs:423-428)
        n += 1
        if (pod, ns) not in seen:
            seen.add((pod, ns))
            order.append((pod, ns))
    return n, order
