"""Host-side similarity functions and the similarity medoid (reference consensus_utils.py:626-917, 1221-1237).

These serve multi-word strings and mixed payloads, which the reference sends to a pairwise-similarity
"centroid" instead of a vote.  SURVEY.md §8f-2 ranks a GPU batched edit distance as the next row after the
scalar hot path; until then this stays host Python, like the reference.
"""
from __future__ import annotations

import logging
import math
from threading import Lock
from typing import Any, Callable, Optional

import numpy as np

logger = logging.getLogger(__name__)

SIMILARITY_SCORE_LOWER_BOUND = 1e-8  # cu:78
IGNORED_KEY_PREFIXES = ("reasoning___", "source___")  # cu:38-43 (re.match == prefix test)

try:  # optional accelerator, same name the reference imports (cu:15)
    from Levenshtein import distance as _lev  # type: ignore
except Exception:  # pragma: no cover - depends on the image
    _lev = None


def levenshtein_distance(a: str, b: str) -> int:
    if _lev is not None:
        return _lev(a, b)
    if a.isascii() and b.isascii():  # always true for normalize_string() output: native C++ DP
        from .. import _native
        return _native.levenshtein(a, b)
    if a == b:
        return 0
    if len(a) < len(b):
        a, b = b, a
    if not b:
        return len(a)
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


def normalize_string(text: str) -> str:
    """cu:660-673: keep ASCII alphanumerics, lower-case."""
    if not text:
        return ""
    return "".join(ch for ch in text if ch.isascii() and ch.isalnum()).lower()


def levenshtein_similarity(s1: str, s2: str) -> float:  # cu:745-761
    a, b = normalize_string(s1), normalize_string(s2)
    longest = max(len(a), len(b))
    if longest == 0:
        return 1.0
    return max(SIMILARITY_SCORE_LOWER_BOUND, 1 - (levenshtein_distance(a, b) / longest))


def jaccard_similarity(s1: str, s2: str) -> float:  # cu:720-742
    sa, sb = set(normalize_string(s1)), set(normalize_string(s2))
    if not (sa | sb):
        return 1.0
    return max(SIMILARITY_SCORE_LOWER_BOUND, len(sa & sb) / len(sa | sb))


def hamming_similarity(s1: str, s2: str) -> float:  # cu:676-717
    a, b = normalize_string(s1), normalize_string(s2)
    longest = max(len(a), len(b))
    if longest == 0:
        return 1.0
    dist = sum(1 for i in range(longest) if (a[i] if i < len(a) else " ") != (b[i] if i < len(b) else " "))
    return max(SIMILARITY_SCORE_LOWER_BOUND, 1 - (dist / longest))


def cosine_similarity(v1, v2) -> float:  # cu:626-649
    a, b = np.array(v1), np.array(v2)
    if a.shape != b.shape:
        raise ValueError("Vectors must have the same shape for cosine similarity")
    n1, n2 = np.linalg.norm(a), np.linalg.norm(b)
    if n1 == 0 or n2 == 0:
        return SIMILARITY_SCORE_LOWER_BOUND
    return np.clip(0.5 * (np.dot(a, b) / (n1 * n2) + 1.0), SIMILARITY_SCORE_LOWER_BOUND, 1.0)


# symmetric memo of string similarities (the reference keeps a TTLCache(1024, 300 s), cu:620-623, 780-794;
# a bounded dict is value-equivalent: entries are pure functions of their key)
_cache: dict = {}
_cache_lock = Lock()
_CACHE_MAX = 4096


def string_similarity(s1: str, s2: str, method: str, embed: Optional[Callable]) -> float:  # cu:797-824
    key = (min(s1, s2), max(s1, s2), method)
    with _cache_lock:
        hit = _cache.get(key)
    if hit is not None:
        return hit
    result = None
    if method == "jaccard":
        result = jaccard_similarity(s1, s2)
    elif method == "hamming":
        result = hamming_similarity(s1, s2)
    elif method == "embeddings" and len(s1) > 50 and len(s2) > 50:  # cu:813: short strings are not worth a request
        try:
            result = cosine_similarity(embed([s1])[0], embed([s2])[0])
        except Exception as exc:  # cu:816-817: logged, then Levenshtein
            logger.error("Error getting embeddings for %r and %r", s1, s2, exc_info=exc)
    if result is None:
        result = levenshtein_similarity(s1, s2)
    with _cache_lock:
        if len(_cache) >= _CACHE_MAX:
            _cache.clear()
        _cache[key] = result
    return result


def numerical_similarity(v1, v2) -> float:  # cu:827-841
    if isinstance(v1, bool) and isinstance(v2, bool):
        return 1.0 if v1 == v2 else SIMILARITY_SCORE_LOWER_BOUND
    if isinstance(v1, (int, float)) and isinstance(v2, (int, float)) and math.isclose(v1, v2, rel_tol=0.01):
        return 1.0
    return 1.0 if v1 == v2 else SIMILARITY_SCORE_LOWER_BOUND


def generic_similarity(v1: Any, v2: Any, method: str, embed: Optional[Callable]) -> float:  # cu:892-917
    if not bool(v1) and not bool(v2):
        return 1.0
    if v1 is None or v2 is None:
        return SIMILARITY_SCORE_LOWER_BOUND
    if isinstance(v1, str) and isinstance(v2, str):
        return string_similarity(v1, v2, method, embed)
    if isinstance(v1, (int, float)) and isinstance(v2, (int, float)):
        return numerical_similarity(v1, v2)
    if isinstance(v1, dict) and isinstance(v2, dict):  # cu:844-869
        keys = [k for k in set(v1) | set(v2) if not k.startswith(IGNORED_KEY_PREFIXES)]
        if not keys:
            return 1.0
        total = 0.0
        for k in keys:
            total += generic_similarity(v1.get(k), v2.get(k), method, embed)
        return total / len(keys)
    if isinstance(v1, (list, tuple)) and isinstance(v2, (list, tuple)):  # cu:872-889
        longest = max(len(v1), len(v2))
        if longest == 0:
            return 1.0
        total = 0.0
        for i in range(longest):
            total += generic_similarity(v1[i] if i < len(v1) else None, v2[i] if i < len(v2) else None, method, embed)
        return total / longest
    return SIMILARITY_SCORE_LOWER_BOUND


def medoid(values: list, method: str, embed: Optional[Callable], parent_valid_frac: float):
    """cu:1221-1237: the value with the highest mean similarity to the others (first on ties)."""
    n = len(values)
    if n == 0:
        return None, 0.0
    if n == 1:
        return values[0], parent_valid_frac
    sims = np.zeros((n, n), dtype=float)
    for i in range(n):
        for j in range(i + 1, n):
            sims[i, j] = sims[j, i] = generic_similarity(values[i], values[j], method, embed)
        sims[i, i] = np.nan
    avg = np.nanmean(sims, axis=1)
    best = int(np.argmax(avg))
    return values[best], round(parent_valid_frac * float(avg[best]), 5)
