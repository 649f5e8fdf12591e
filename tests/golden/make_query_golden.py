"""Generates tests/golden/query_render.json: the reference's PromQL template
(/root/reference/gpu-pruner/src/query.promql.j2) rendered by jinja2 for a grid of CLI argument
sets.  The reference renders it with minijinja (main.rs:280-281), whose whitespace-control and
float formatting rules match jinja2 for this template.  Only the RENDERED TEXT is committed (the
template itself is not copied); /root/reference is needed only when regenerating.

    python tests/golden/make_query_golden.py
"""
import itertools
import json
import os

import jinja2

TEMPLATE = "/root/reference/gpu-pruner/src/query.promql.j2"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "query_render.json")


def main():
    tpl = jinja2.Environment().from_string(open(TEMPLATE).read())
    cases = []
    for duration, ns, model, power, honor in itertools.product(
            (30, 45, 15), (None, "ml-team"), (None, "NVIDIA A100"), (None, 150.0, 100.0, 120.5, 0.0),
            (False, True)):
        args = {"duration": duration, "namespace": ns, "model_name": model,
                "power_threshold": power, "honor_labels": honor}
        cases.append({"args": args, "text": tpl.render(args=args)})
    # a second, sparse grid: strings are inserted into the PromQL verbatim (no escaping in the template), regex
    # alternations and dots included; a few more float spellings of the threshold; extreme look-backs
    odd_ns = ("ml-team|prod", "team.with.dots", "ns-[0-9]+", 'we"ird', "x" * 70)
    odd_model = ("NVIDIA A100-SXM4-80GB", ".*H100.*", "Tesla V100|Tesla T4", "A\\d+")
    odd_power = (99.9, 0.5, 250.25, 1000.0, 1.0)
    for i, (ns, model) in enumerate(itertools.product(odd_ns, odd_model)):
        args = {"duration": (1, 5, 60, 1440)[i % 4], "namespace": ns, "model_name": model if i % 3 else None,
                "power_threshold": odd_power[i % len(odd_power)] if i % 2 else None, "honor_labels": bool(i % 5 == 0)}
        cases.append({"args": args, "text": tpl.render(args=args)})
    json.dump(cases, open(OUT, "w"), indent=1)
    print(f"wrote {len(cases)} renderings to {OUT}")
    print(cases[0]["text"])


if __name__ == "__main__":
    main()
