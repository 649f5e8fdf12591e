"""Recorded daemon-mode ticks for the resident-window tests (TEST INFRASTRUCTURE).

A tiny time-series store — {series: (metric, labels, [(ts, value)])} — answers the two questions a daemon-mode tick can
ask a Prometheus server (main.rs:286-330 runs the same query every --check-interval seconds):
  full  : `METRIC{...}[Nm]` evaluated at t_k          -> every sample with t_k - N < ts <= t_k
  delta : `METRIC{...}[(t_k - t_{k-1})s]` at t_k       -> every sample with t_{k-1} < ts <= t_k
and write_ticks() lays the answers out as file://DIR/tick-%04d/{full,delta}/ fixtures (controller.cpp FileSource)."""
import json
import os


def _fmt_value(v):
    if isinstance(v, str):
        return v
    if v != v:
        return "NaN"
    if float(v).is_integer() and abs(v) < 1e15:
        return str(int(v))
    return repr(float(v))


def _fmt_ts(ts):
    return str(int(ts)) if float(ts).is_integer() else ("%.3f" % ts)


def response(store, metric, lo, hi):
    """compact matrix response of `metric[...]` over (lo, hi]; series without a sample in range are absent"""
    parts = []
    for name, labels, samples in store:
        if name != metric:
            continue
        vals = ",".join('[%s,"%s"]' % (_fmt_ts(t), _fmt_value(v)) for t, v in samples if lo < t <= hi)
        if vals:
            parts.append('{"metric":%s,"values":[%s]}' % (json.dumps(dict(labels, __name__=metric), separators=(",", ":")), vals))
    return '{"status":"success","data":{"resultType":"matrix","result":[' + ",".join(parts) + "]}}"


def write_ticks(root, store_at, tick_times, window_s, step, with_power=False, with_prof=True, skip_delta=()):
    """store_at(k) -> the store as of tick k (lets a scenario add / drop series over time).  Returns root."""
    prev = None
    for k, t in enumerate(tick_times):
        store = store_at(k)
        base = os.path.join(root, "tick-%04d" % k)
        for kind, lo in (("full", t - window_s), ("delta", prev)):
            if lo is None or (kind == "delta" and k in skip_delta):
                continue
            d = os.path.join(base, kind)
            os.makedirs(d, exist_ok=True)
            with open(os.path.join(d, "util.json"), "w") as f:
                f.write(response(store, "DCGM_FI_DEV_GPU_UTIL", lo, t))
            if with_prof and any(s[0] == "DCGM_FI_PROF_GR_ENGINE_ACTIVE" for s in store):
                with open(os.path.join(d, "prof.json"), "w") as f:
                    f.write(response(store, "DCGM_FI_PROF_GR_ENGINE_ACTIVE", lo, t))
            if with_power:
                with open(os.path.join(d, "power.json"), "w") as f:
                    f.write(response(store, "DCGM_FI_DEV_POWER_USAGE", lo, t))
            q = {"end": t, "step": step}
            if kind == "delta":
                q["start"] = lo
            with open(os.path.join(d, "query.json"), "w") as f:
                json.dump(q, f)
        prev = t
    return root


def labels(pod, gpu, ns="ml", host=None, **extra):
    d = {"Hostname": host or f"node-{hash(pod) % 5}", "gpu": str(gpu), "modelName": "NVIDIA B200", "exported_pod": pod,
         "exported_namespace": ns, "exported_container": "main", "UUID": f"GPU-{pod}-{gpu}"}
    d.update(extra)
    return d
