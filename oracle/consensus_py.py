"""Object-level CPU oracle: a restatement of the reference's consolidation hot path.

TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this module; the product
(`k_llms_b200`) never does, and fails loudly when its CUDA library is missing.

It restates, in plain Python + numpy (the same numpy the reference calls), what
`/root/reference/k_llms/utils/consensus_utils.py` (abbrev. `cu`) computes for ONE record:

    consensus_values      cu:1376-1454   dispatcher (a1)
    consensus_dict        cu:1269-1306   field by field (a2)
    consensus_list        cu:1309-1352   element by element (a2')
    voting_consensus      cu:936-982     str/bool mode, first-seen ties (a3)
    sanitize_value        cu:925-933
    consensus_as_primitive cu:1075-1237  numeric 1-D clustering (a4) + similarity medoid (a5)
    generic_similarity & friends cu:660-917
    recursive_list_alignments, dict/flat part only  cu:458-548 (a7)

Parity pin: the reference ships no tests or golden vectors, so this oracle is
pinned against OUTPUTS OF THE REFERENCE ITSELF, run in the build container behind three import
stubs (`oracle/ref_loader.py`): `oracle/gen_golden.py` writes `tests/golden/*.json`, and
`tests/test_oracle_golden.py` replays them here and on the GPU box.  Non-ASCII strings are
rejected (the reference needs `unidecode`, absent from this image: unpinned).
"""
from __future__ import annotations

import math
from typing import Any, Callable, Optional

import numpy as np

SIM_FLOOR = 1e-8  # cu:78
REL_EPS = 0.03  # cu:63
ABS_EPS = 1e-6  # cu:64
SKIPPED_KEY_MARKERS = ("reasoning___", "source___")  # cu:1287


class OracleSettings:
    """The live fields of ConsensusSettings (cu:53-69)."""

    def __init__(self, allow_none_as_candidate: bool = False, rel_eps: float = REL_EPS, abs_eps: float = ABS_EPS,
                 string_similarity_method: str = "embeddings", min_support_ratio: float = 0.51):
        self.allow_none_as_candidate = allow_none_as_candidate
        self.rel_eps = rel_eps
        self.abs_eps = abs_eps
        self.string_similarity_method = string_similarity_method
        self.min_support_ratio = min_support_ratio


DEFAULTS = OracleSettings()

# ----------------------------------------------------------------------------- a3: vote


def sanitize(v: Any) -> str:
    """cu:925-933: str() -> lower -> drop spaces -> unidecode -> keep [a-zA-Z0-9]."""
    s = str(v).lower().replace(" ", "")
    if not s.isascii():
        raise ValueError("oracle: non-ASCII input; unidecode behaviour is unpinned (SURVEY §8c)")
    return "".join(ch for ch in s if ch.isalnum())  # ascii-only here, so isalnum == [a-zA-Z0-9]


def vote(values: list, settings: OracleSettings = DEFAULTS, pvf: float = 1.0):
    """cu:936-982.  Ties go to the first-seen class (Counter insertion order + max())."""
    present = len(values)
    first = next((v for v in values if v is not None), None)
    if first is None:
        return None, pvf  # cu:947-948
    if isinstance(first, bool):
        keys = [v or False for v in values]  # cu:956 (None and every falsy value -> False)
        originals = keys
    else:
        originals = values if settings.allow_none_as_candidate else [v for v in values if v is not None]
        keys = [None if v is None else sanitize(v) for v in originals]  # cu:966
    tally: dict = {}
    for k in keys:
        tally[k] = tally.get(k, 0) + 1
    best_key, best_count = None, 0
    for k, c in tally.items():  # insertion order; strict '>' keeps the first-seen maximum
        if c > best_count:
            best_key, best_count = k, c
    if isinstance(first, bool):
        winner = next(k for k in tally if k == best_key)  # the Counter key itself (cu:958)
    else:
        winner = originals[keys.index(best_key)]  # first original with that sanitised form (cu:971)
    return winner, round(pvf * (best_count / present), 5)  # cu:973,982


# ----------------------------------------------------------------------------- a4: numeric


def _close(a: float, b: float, rel_eps: float, abs_eps: float) -> bool:
    """cu:1130-1133 / 1146-1148."""
    return abs(a - b) <= max(abs_eps, rel_eps * max(abs(a), abs(b), 1.0))


def _close_pow10(a: float, b: float, rel_eps: float, abs_eps: float) -> bool:
    """cu:1153-1160: b rescaled by 10^k, k in [-6, 6]."""
    if a == 0.0 or b == 0.0:
        return _close(a, b, rel_eps, abs_eps)
    return any(_close(a, b * (10.0 ** k), rel_eps, abs_eps) for k in range(-6, 7))


def numeric(values: list, settings: OracleSettings = DEFAULTS, pvf: float = 1.0):
    """Numeric branch of consensus_as_primitive (cu:1098-1219) for a list WITHOUT Nones
    (the dispatcher strips them, cu:1444-1448, so none_count == 0 and pvf is never applied)."""
    total = len(values)
    xs = sorted(float(v) for v in values
                if isinstance(v, (int, float)) and not isinstance(v, bool) and _finite(v))
    if not xs:
        return None, pvf  # cu:1115-1116
    rel_eps, abs_eps = settings.rel_eps, settings.abs_eps
    clusters = [[xs[0]]]
    for a, b in zip(xs, xs[1:]):  # chain adjacent values (cu:1127-1144)
        if _close(a, b, rel_eps, abs_eps):
            clusters[-1].append(b)
        else:
            clusters.append([b])
    sizes = [len(c) for c in clusters]
    top = max(sizes)
    if sizes.count(top) == 1:  # covers both cu:1171-1178 and cu:1180-1187
        return float(np.mean(clusters[sizes.index(top)])), round(top / total, 5)
    # tie between equally large clusters (cu:1189-1219)
    centers = [float(np.median(c)) for c in clusters]
    spreads = [float(np.std(c)) if len(c) > 1 else 0.0 for c in clusters]
    ranked = []
    for ci, c in enumerate(clusters):
        if len(c) != top:
            continue
        support = top
        for oi, other in enumerate(clusters):
            if oi != ci and len(other) < top and (
                _close(centers[ci], centers[oi], rel_eps, abs_eps)
                or _close(abs(centers[ci]), abs(centers[oi]), rel_eps, abs_eps)
                or _close_pow10(centers[ci], centers[oi], rel_eps, abs_eps)
            ):
                support += len(other)
        ranked.append((-support, spreads[ci], -abs(centers[ci]), ci))  # ci last == stable order
    ranked.sort()
    neg_support, _, _, best = ranked[0]
    return float(np.mean(clusters[best])), round(-neg_support / total, 5)


def _finite(v) -> bool:
    try:
        return math.isfinite(float(v))
    except OverflowError:  # huge ints (cu:1109-1114)
        return False


# ----------------------------------------------------------------------------- a5: medoid


def edit_distance(a: str, b: str) -> int:
    if a == b:
        return 0
    if len(a) < len(b):
        a, b = b, a
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


def _alnum_lower(text: str) -> str:
    """cu:660-673 (regex is ASCII-only, so non-ASCII letters are dropped, not transliterated)."""
    return "".join(ch for ch in text if ch.isascii() and ch.isalnum()).lower() if text else ""


def string_similarity(s1: str, s2: str, method: str, embed: Optional[Callable]) -> float:
    """cu:797-824 without the TTL cache (the cache is keyed symmetrically and is value-neutral)."""
    a, b = _alnum_lower(s1), _alnum_lower(s2)
    if method == "jaccard":  # cu:720-742
        sa, sb = set(a), set(b)
        return 1.0 if not (sa | sb) else max(SIM_FLOOR, len(sa & sb) / len(sa | sb))
    if method == "hamming":  # cu:676-717
        longest = max(len(a), len(b))
        if longest == 0:
            return 1.0
        mism = sum(1 for i in range(longest) if (a[i] if i < len(a) else " ") != (b[i] if i < len(b) else " "))
        return max(SIM_FLOOR, 1 - mism / longest)
    if method == "embeddings" and len(s1) > 50 and len(s2) > 50 and embed is not None:  # cu:813-817
        try:
            v1, v2 = np.array(embed([s1])[0]), np.array(embed([s2])[0])
            n1, n2 = np.linalg.norm(v1), np.linalg.norm(v2)
            if n1 == 0 or n2 == 0:
                return SIM_FLOOR
            return np.clip(0.5 * (np.dot(v1, v2) / (n1 * n2) + 1.0), SIM_FLOOR, 1.0)
        except Exception:
            pass  # reference logs and falls through to Levenshtein
    longest = max(len(a), len(b))  # cu:745-761
    if longest == 0:
        return 1.0
    return max(SIM_FLOOR, 1 - edit_distance(a, b) / longest)


def similarity(v1: Any, v2: Any, method: str, embed: Optional[Callable]) -> float:
    """generic_similarity, cu:892-917."""
    if not bool(v1) and not bool(v2):
        return 1.0
    if v1 is None or v2 is None:
        return SIM_FLOOR
    if isinstance(v1, str) and isinstance(v2, str):
        return string_similarity(v1, v2, method, embed)
    if isinstance(v1, (int, float)) and isinstance(v2, (int, float)):  # cu:827-841
        if isinstance(v1, bool) and isinstance(v2, bool):
            return 1.0 if v1 == v2 else SIM_FLOOR
        if math.isclose(v1, v2, rel_tol=0.01):
            return 1.0
        return 1.0 if v1 == v2 else SIM_FLOOR
    if isinstance(v1, dict) and isinstance(v2, dict):  # cu:844-869
        keys = [k for k in set(v1) | set(v2) if not k.startswith(SKIPPED_KEY_MARKERS)]  # re.match == prefix
        if not keys:
            return 1.0
        return sum(similarity(v1.get(k), v2.get(k), method, embed) for k in keys) / len(keys)
    if isinstance(v1, (list, tuple)) and isinstance(v2, (list, tuple)):  # cu:872-889
        longest = max(len(v1), len(v2))
        if longest == 0:
            return 1.0
        acc = 0.0
        for i in range(longest):
            acc += similarity(v1[i] if i < len(v1) else None, v2[i] if i < len(v2) else None, method, embed)
        return acc / longest
    return SIM_FLOOR


def medoid(values: list, settings: OracleSettings = DEFAULTS, pvf: float = 1.0, embed: Optional[Callable] = None):
    """cu:1221-1237 for n >= 2 non-None values."""
    n = len(values)
    sims = np.zeros((n, n), dtype=float)
    for i in range(n):
        for j in range(i + 1, n):
            sims[i, j] = sims[j, i] = similarity(values[i], values[j], settings.string_similarity_method, embed)
        sims[i, i] = np.nan
    avg = np.nanmean(sims, axis=1)
    best = int(np.argmax(avg))
    return values[best], round(pvf * float(avg[best]), 5)


def primitive(values: list, settings: OracleSettings = DEFAULTS, pvf: float = 1.0, embed: Optional[Callable] = None):
    """consensus_as_primitive (cu:1075-1237) as reached from the dispatcher (values hold no None).
    The `llm-consensus` branch (cu:1090-1096) is a network call and is out of scope."""
    if len(values) == 0:
        return None, pvf
    if len(values) == 1:
        return values[0], pvf * (1 / 1)  # cu:1085-1086: unrounded, original object
    first_type = type(values[0])
    try:
        numeric_like = isinstance(first_type(), (int, float))  # cu:1099
    except Exception:
        numeric_like = False
    if numeric_like or all(isinstance(v, (int, float)) for v in values):
        return numeric(values, settings, pvf)
    return medoid(values, settings, pvf, embed)


# ----------------------------------------------------------------------------- a1/a2/a2'


def consensus(values: list, settings: OracleSettings = DEFAULTS, pvf: float = 1.0, embed: Optional[Callable] = None):
    """consensus_values, cu:1376-1454."""
    if not values:
        return None, pvf
    live = [v for v in values if v is not None]
    if not live:
        return None, 0.0
    head = live[0]
    if isinstance(head, (str, bool)) and all(len(str(v).strip().split()) < 3 for v in live):  # cu:1405-1411
        return vote(values, settings, pvf)
    if isinstance(head, dict):  # cu:1414-1426
        dicts = [v for v in values if isinstance(v, dict)]
        sub = pvf * (len(dicts) / len(values))
        keys: dict = {}
        for d in dicts:  # first-seen key order (cu:1281-1282)
            for k in d:
                keys.setdefault(k, None)
        out, conf = {}, {}
        for k in keys:
            if any(m in k for m in SKIPPED_KEY_MARKERS):  # cu:1292-1294
                continue
            out[k], conf[k] = consensus([d.get(k) for d in dicts], settings, sub, embed)
        return out, conf
    if isinstance(head, list):  # cu:1429-1441
        lists = [v for v in values if isinstance(v, list)]
        sub = pvf * (len(lists) / len(values))
        longest = max(len(l) for l in lists)
        if longest == 0:
            return [], []
        out_l, conf_l = [], []
        for i in range(longest):  # shorter lists contribute None (cu:1341)
            v, c = consensus([l[i] if i < len(l) else None for l in lists], settings, sub, embed)
            out_l.append(v)
            conf_l.append(c)
        return out_l, conf_l
    sub = pvf * (len(live) / len(values))  # cu:1444
    return primitive(live, settings, sub, embed)


def align_flat(values: list) -> list:
    """The dict / scalar part of recursive_list_alignments (cu:458-548): every dict gets every key,
    keys sorted, missing -> None, recursively.  List alignment (cu:550-613) is NOT restated here
    ('next'); lists are passed through untouched."""
    if not values or all(v is None for v in values):
        return values
    live = [v for v in values if v is not None]
    if not all(isinstance(x, type(live[0])) for x in live) or type(live[0]) is not dict:
        return list(values)
    dicts = [dict(d) if isinstance(d, dict) else {} for d in values]
    keys = sorted({k for d in dicts for k in d})
    for k in keys:
        col = align_flat([d.get(k) for d in dicts])
        for d, v in zip(dicts, col):
            d[k] = v
    return [{k: d.get(k) for k in keys} for d in dicts]


def client_order(values: list, settings: OracleSettings = DEFAULTS, embed: Optional[Callable] = None):
    """consolidation.py:339-354 for list-free payloads: align, coerce to dicts, consensus."""
    aligned = [(d if isinstance(d, dict) else {}) for d in align_flat(values)] if len(values) >= 2 else values
    return consensus(aligned, settings, 1.0, embed)
