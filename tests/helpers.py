"""Shared test helpers: golden loading, strict equality (float bits, key order), a CPU stand-in device."""
import json
import math
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        return json.load(f)["cases"]


def same(a, b) -> bool:
    """Equality the way the parity bar means it: same types, same dict key ORDER, same float bits (NaN == NaN)."""
    if isinstance(a, float) and isinstance(b, float):
        return (math.isnan(a) and math.isnan(b)) or (a == b and math.copysign(1, a) == math.copysign(1, b))
    if type(a) is not type(b):
        return False
    if isinstance(a, dict):
        return list(a) == list(b) and all(same(a[k], b[k]) for k in a)
    if isinstance(a, (list, tuple)):
        return len(a) == len(b) and all(same(x, y) for x, y in zip(a, b))
    return a == b


def raising_embeddings(texts):
    raise RuntimeError("no network in tests")


def oracle_run(plan):
    """Stand-in for Plan.run(): evaluate the recorded groups with the columnar C ORACLE instead of the GPU.
    Used by CPU tests of the host prologue/epilogue only."""
    from oracle import columnar as OC
    out = {}
    if plan.vote_rows:
        _, meta = OC.vote(np.asarray(plan.vote_rows, dtype=np.int32), None)
        out["vote_meta"] = meta
    if plan.num_rows:
        value, meta = OC.numeric(np.asarray(plan.num_rows, dtype=np.float64), plan.rel_eps, plan.abs_eps)
        out["num_value"], out["num_meta"] = value, meta
    return out
