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
    if plan.medoid_groups:
        out["medoid_idx"], out["medoid_avg"] = OC.medoid(plan.medoid_groups)
    return out


_WORDS = ("invoice total due amount net gross payment bank transfer within thirty days from receipt of goods and services "
          "the a an of to acme corp ltd gmbh street road avenue suite floor new york london paris berlin 2024 2025 q1 q2 "
          "ref no id number 000123 77 ab-12 x y z").split()


def random_string_groups(rng, n_groups, max_k=16, long_frac=0.1):
    """Groups of 2..max_k multi-word strings: noisy copies of a base phrase (the shape of LLM string fields)."""
    groups = []
    for _ in range(n_groups):
        k = int(rng.integers(2, max_k + 1))
        base = [_WORDS[int(i)] for i in rng.integers(0, len(_WORDS), int(rng.integers(3, 9)))]
        grp = []
        for _c in range(k):
            words = list(base)
            r = rng.random()
            if r < 0.35:
                pass
            elif r < 0.6:
                words[int(rng.integers(0, len(words)))] = _WORDS[int(rng.integers(0, len(_WORDS)))]
            elif r < 0.75:
                words = words[: max(1, len(words) - int(rng.integers(1, 3)))]
            elif r < 0.9:
                words = words + [_WORDS[int(i)] for i in rng.integers(0, len(_WORDS), int(rng.integers(1, 4)))]
            else:
                words = [w.upper() if rng.random() < 0.5 else w + "," for w in words]
            s = " ".join(words)
            if rng.random() < 0.03:
                s = ""
            grp.append(s)
        if rng.random() < long_frac:  # one long member is still inside K4's contract (pattern = the shorter string)
            grp[int(rng.integers(0, k))] = " ".join(_WORDS[int(i)] for i in rng.integers(0, len(_WORDS), 40))
        groups.append(grp)
    return groups
