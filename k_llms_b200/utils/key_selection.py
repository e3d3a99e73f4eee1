"""Join-key selection for the key-based list aligner (host logic, as in the reference).

Mirrors the public surface of the reference's `k_llms/utils/key_selection.py` (same names, arguments, results and
error messages) so that `from k_llms_b200.utils.key_selection import select_best_keys, CascadeConfig` is a drop-in.
Upstream nothing on the client path imports this module; results are pinned against the
reference's own outputs in `tests/golden/key_alignment.json`.

Own structure: every extraction's records are resolved ONCE into a `_Table`; each candidate path's value column is read
from it once and the metrics of single and composite keys come from the same `_score` routine (the reference walks the
JSON again for every evaluation and evaluates every single twice).

What is computed (reference `key_selection.py`):
  * a candidate is a dot path that reaches a scalar somewhere in a record (`:100-122`); list-valued keys never are;
  * a value is compared after `normalize_scalar` (`:24-30`): strings stripped, lower-cased, whitespace runs -> " ";
  * per candidate: coverage and uniqueness inside each extraction, Jaccard between every pair of extractions, how many
    values occur in all / all-but-one / at least two extractions, size of the union (`:156-214`);
  * a four-stage funnel over the single keys (`:290-350`) and a greedy + exhaustive search for a composite of up to
    `max_k` of the funnel's survivors (`:383-445`).
Python `set` / `dict` equality decides which values are "the same" (so 1, 1.0 and True are one value, as upstream).
"""
from __future__ import annotations

import math
import re
from collections import Counter
from itertools import combinations
from typing import Any, Dict, List, Optional, Sequence, Set, Tuple

from pydantic import BaseModel, ConfigDict

JSONScalar = Any
JSONPath = str

# Keys tried, in order, when no `list_key` is given (reference `:36`); module-level so callers can override it.
RECORD_LIST_KEYS: List[str] = ["products"]

_WS = re.compile(r"\s+")
_MISSING = object()


def normalize_scalar(value: Any) -> Any:
    """Strings: strip, lower-case, collapse whitespace runs; everything else unchanged (reference `:24-30`)."""
    if isinstance(value, str):
        return _WS.sub(" ", value.strip().lower())
    return value


def _dict_items_of(seq: Any) -> List[Dict[str, Any]]:
    return [item for item in seq if isinstance(item, dict)] if isinstance(seq, list) else []


def iter_records(extraction: Dict[str, Any], list_key: Optional[str] = None) -> List[Dict[str, Any]]:
    """The record dicts of one extraction (reference `:38-77`): under `list_key` when given; otherwise under the
    `RECORD_LIST_KEYS`, and if those hold none, the dict items of EVERY list-valued top-level entry."""
    if list_key is not None:
        return _dict_items_of(extraction.get(list_key))
    found: List[Dict[str, Any]] = []
    for key in RECORD_LIST_KEYS:
        found.extend(_dict_items_of(extraction.get(key)))
    if found:
        return found
    for value in extraction.values():
        found.extend(_dict_items_of(value))
    return found


def _resolve(record: Any, parts: Sequence[str]) -> Any:
    """The scalar at `parts` inside `record`, or `_MISSING` (unreachable, None, dict or list)."""
    node = record
    for token in parts:
        if not (isinstance(node, dict) and token in node):
            return _MISSING
        node = node[token]
    if node is None or isinstance(node, (dict, list)):
        return _MISSING
    return node


def values_for_path(extraction: Dict[str, Any], path: JSONPath, list_key: Optional[str] = None) -> List[JSONScalar]:
    """Normalised scalar values of `path` over the records of one extraction (reference `:79-97`)."""
    parts = path.split(".")
    hits = (_resolve(record, parts) for record in iter_records(extraction, list_key=list_key))
    return [normalize_scalar(v) for v in hits if v is not _MISSING]


def tuple_values_for_paths(extraction: Dict[str, Any], paths: List[JSONPath], list_key: Optional[str] = None) -> List[Tuple[Any, ...]]:
    """Composite values: one tuple per record in which EVERY path resolves (reference `:237-262`)."""
    split = [p.split(".") for p in paths]
    out: List[Tuple[Any, ...]] = []
    for record in iter_records(extraction, list_key=list_key):
        parts_values = [_resolve(record, parts) for parts in split]
        if parts_values and all(v is not _MISSING for v in parts_values):
            out.append(tuple(normalize_scalar(v) for v in parts_values))
    return out


def discover_scalar_paths(extractions: List[Dict[str, Any]], list_key: Optional[str] = None) -> List[JSONPath]:
    """Sorted dot paths that hold a non-dict, non-list value (None included) in some record (reference `:100-122`)."""
    found: Set[str] = set()

    def walk(prefix: str, node: Dict[str, Any]) -> None:
        for key, value in node.items():
            path = f"{prefix}.{key}" if prefix else key
            if isinstance(value, dict):
                walk(path, value)
            elif not isinstance(value, list):
                found.add(path)

    for extraction in extractions:
        for record in iter_records(extraction, list_key=list_key):
            walk("", record)
    return sorted(found)


def jaccard(a: Set[Any], b: Set[Any]) -> float:
    """|a & b| / |a | b|; two empty sets agree (1.0), one empty set does not (0.0) (reference `:127-134`)."""
    if not a and not b:
        return 1.0
    if not a or not b:
        return 0.0
    union = len(a | b)
    return len(a & b) / union if union else 1.0


class KeyMetrics(BaseModel):
    """Reference `:137-152`."""
    model_config = ConfigDict(frozen=True)

    path: Tuple[str, ...]            # one path for a single key, several for a composite
    coverage_min: float
    coverage_mean: float
    uniqueness_min: float
    uniqueness_mean: float
    jaccard_min: float
    jaccard_mean: float
    I_E: int                         # values present in all extractions
    I_E_minus_1: int                 # ... in all but one
    I_ge_2: int                      # ... in at least two
    union_size: int
    score_tuple: Tuple               # lexicographic rank, larger is better


def _score(path: Tuple[str, ...], columns: List[List[Any]], totals: List[int], depth: int) -> KeyMetrics:
    """Metrics of one (single or composite) key from its value column per extraction (reference `:156-214`).
    `totals[e]` is the record count of extraction e; `depth` the number of dots over the key's paths."""
    n_ext = len(columns)
    sets = [set(col) for col in columns]

    coverage: List[float] = []
    uniqueness: List[float] = []
    for col, total in zip(columns, totals):
        present = len(col)
        coverage.append(present / max(1, total))
        once = sum(1 for c in Counter(col).values() if c == 1)
        uniqueness.append(once / max(1, present) if present else 0.0)

    pairs = [jaccard(sets[i], sets[j]) for i in range(n_ext) for j in range(i + 1, n_ext)]
    j_min = min(pairs) if pairs else 1.0
    j_mean = sum(pairs) / len(pairs) if pairs else 1.0

    support: Counter = Counter()
    for s in sets:
        support.update(s)
    by_support = Counter(support.values())
    in_all = by_support.get(n_ext, 0)
    in_all_but_one = by_support.get(n_ext - 1, 0) if n_ext >= 2 else 0
    in_two_plus = sum(c for sup, c in by_support.items() if sup >= 2)
    union_size = len(support)

    rank = (round(j_min, 6), in_all, in_all_but_one, round(j_mean, 6), round(min(uniqueness), 6), round(min(coverage), 6),
            -union_size, depth, -len(path))
    return KeyMetrics(
        path=path,
        coverage_min=min(coverage) if coverage else 0.0,
        coverage_mean=sum(coverage) / len(coverage) if coverage else 0.0,
        uniqueness_min=min(uniqueness) if uniqueness else 0.0,
        uniqueness_mean=sum(uniqueness) / len(uniqueness) if uniqueness else 0.0,
        jaccard_min=j_min, jaccard_mean=j_mean,
        I_E=in_all, I_E_minus_1=in_all_but_one, I_ge_2=in_two_plus, union_size=union_size, score_tuple=rank)


class _Table:
    """The records of every extraction, resolved once; value columns per path on demand."""

    def __init__(self, extractions: List[Dict[str, Any]], list_key: Optional[str]):
        self.records = [iter_records(e, list_key=list_key) for e in extractions]
        self.totals = [len(r) for r in self.records]
        self._raw: Dict[str, List[List[Any]]] = {}

    def raw(self, path: str) -> List[List[Any]]:
        """Per extraction, one entry per record: the normalised scalar at `path` or `_MISSING`."""
        got = self._raw.get(path)
        if got is None:
            parts = path.split(".")
            got = []
            for records in self.records:
                hits = [_resolve(record, parts) for record in records]
                got.append([v if v is _MISSING else normalize_scalar(v) for v in hits])
            self._raw[path] = got
        return got

    def single(self, path: str, canon=None) -> KeyMetrics:
        columns = [[v for v in col if v is not _MISSING] for col in self.raw(path)]
        if canon is not None:
            columns = [[canon(v) for v in col] for col in columns]
        return _score((path,), columns, self.totals, path.count("."))

    def composite(self, paths: Sequence[str]) -> KeyMetrics:
        per_path = [self.raw(p) for p in paths]
        columns: List[List[Any]] = []
        for e in range(len(self.records)):
            rows = zip(*(col[e] for col in per_path))
            columns.append([row for row in rows if all(v is not _MISSING for v in row)])
        return _score(tuple(paths), columns, self.totals, sum(p.count(".") for p in paths))


def evaluate_single_key(extractions: List[Dict[str, Any]], path: JSONPath, list_key: Optional[str] = None) -> KeyMetrics:
    """Reference `:217-234`."""
    return _Table(extractions, list_key).single(path)


def evaluate_composite_key(extractions: List[Dict[str, Any]], paths: List[JSONPath], list_key: Optional[str] = None) -> KeyMetrics:
    """Reference `:265-285`."""
    return _Table(extractions, list_key).composite(list(paths))


class CascadeConfig(BaseModel):
    """Reference `:290-298`."""
    model_config = ConfigDict(frozen=True)

    min_coverage: float = 0.0
    min_uniqueness: float = 0.0
    topk_stage1: int = 30
    topk_stage2: int = 12
    topk_stage3: int = 6


class CascadeReport(BaseModel):
    """Reference `:301-309`."""
    model_config = ConfigDict(frozen=True)

    stage0_kept: List[KeyMetrics]
    stage1_kept: List[KeyMetrics]
    stage2_kept: List[KeyMetrics]
    stage3_kept: List[KeyMetrics]
    final_best: KeyMetrics


def _dots(m: KeyMetrics) -> int:
    return sum(p.count(".") for p in m.path)


def _funnel(singles: List[KeyMetrics], config: CascadeConfig, empty_message: str) -> CascadeReport:
    """Gate, then three stable sorts with a cut after each, then the tie-break (reference `:312-372`).  Ties keep the
    candidates' (sorted-path) order at every stage: Python's sort is stable also with reverse=True."""
    gate = [m for m in singles
            if m.I_ge_2 > 0 and m.jaccard_min > 0.0 and m.coverage_min >= config.min_coverage and m.uniqueness_min >= config.min_uniqueness]
    if not gate:
        raise ValueError(empty_message)
    stable = sorted(gate, key=lambda m: (m.I_E, m.I_E_minus_1, round(m.jaccard_min, 6), round(m.jaccard_mean, 6)),
                    reverse=True)[:config.topk_stage1]
    clean = sorted(stable, key=lambda m: (round(m.uniqueness_min, 6), round(m.coverage_min, 6)), reverse=True)[:config.topk_stage2]
    small = sorted(clean, key=lambda m: (m.union_size,))[:config.topk_stage3]
    best = sorted(small, key=lambda m: (_dots(m), -len(m.path)), reverse=True)[0]
    return CascadeReport(stage0_kept=gate, stage1_kept=stable, stage2_kept=clean, stage3_kept=small, final_best=best)


_STAGE0_MESSAGE = "No keys pass Stage 0 (require I_ge_2>0, jaccard_min>0, and coverage)."


def cascade_select_keys(extractions: List[Dict[str, Any]], candidates: List[str], config: CascadeConfig = CascadeConfig(),
                        list_key: Optional[str] = None) -> CascadeReport:
    """Reference `:312-372`."""
    table = _Table(extractions, list_key)
    return _funnel([table.single(p) for p in candidates], config, _STAGE0_MESSAGE)


class KeySelectionResult(BaseModel):
    """Reference `:377-385`."""
    model_config = ConfigDict(frozen=True)

    best_single: KeyMetrics
    best_composite: Optional[KeyMetrics]
    candidate_table: List[KeyMetrics]
    min_support_for_autolock: int
    cascade_report: CascadeReport


def _stability(m: KeyMetrics) -> Tuple:
    return (round(m.jaccard_min, 6), m.I_E, m.I_E_minus_1, round(m.jaccard_mean, 6))


def _select(table: _Table, candidates: List[str], n_ext: int, max_candidates_for_composite: int, max_k: int,
            min_support_ratio_for_autolock: float, cascade_cfg: CascadeConfig) -> KeySelectionResult:
    singles = [table.single(p) for p in candidates]
    report = _funnel(singles, cascade_cfg, _STAGE0_MESSAGE)

    listing = [m for m in singles if m.I_ge_2 > 0 and m.jaccard_min > 0.0]
    listing.sort(key=lambda m: (round(m.jaccard_min, 6), m.I_E, m.I_E_minus_1, round(m.jaccard_mean, 6),
                                round(m.uniqueness_min, 6), round(m.coverage_min, 6), -m.union_size), reverse=True)

    pool = [m.path[0] for m in report.stage3_kept][:max_candidates_for_composite]
    best: Optional[KeyMetrics] = None
    if pool:
        # greedy growth from the funnel's first survivor (reference `:419-431`): a path joins when both the full rank and
        # the stability rank rise; `max_k` is only looked at between passes, so one pass may add several paths
        chosen = [pool[0]]
        best = table.composite(chosen)
        grew = True
        while grew and len(chosen) < max_k:
            grew = False
            for path in pool:
                if path in chosen:
                    continue
                trial = table.composite(chosen + [path])
                if trial.score_tuple > best.score_tuple and _stability(trial) > _stability(best):
                    best, grew = trial, True
                    chosen.append(path)
        # every 2..max_k subset of the pool (reference `:433-439`): taken when either rank rises
        for size in range(2, min(max_k, len(pool)) + 1):
            for combo in combinations(pool, size):
                trial = table.composite(list(combo))
                if _stability(trial) > _stability(best) or trial.score_tuple > best.score_tuple:
                    best = trial

    return KeySelectionResult(best_single=report.final_best, best_composite=best, candidate_table=listing,
                              min_support_for_autolock=max(2, math.ceil(min_support_ratio_for_autolock * n_ext)),
                              cascade_report=report)


def select_best_keys(extractions: List[Dict[str, Any]], max_candidates_for_composite: int = 20, max_k: int = 3,
                     min_support_ratio_for_autolock: float = 0.75, cascade_cfg: CascadeConfig = CascadeConfig(),
                     list_key: Optional[str] = None) -> KeySelectionResult:
    """Best single key (the funnel's winner) and best composite key over the extractions' records (reference `:388-445`).
    Raises ValueError when there are no extractions, no scalar paths, or no key passes the gate."""
    if not extractions:
        raise ValueError("No extractions provided.")
    candidates = discover_scalar_paths(extractions, list_key=list_key)
    if not candidates:
        raise ValueError("No scalar candidate paths discovered.")
    return _select(_Table(extractions, list_key), candidates, len(extractions), max_candidates_for_composite, max_k,
                   min_support_ratio_for_autolock, cascade_cfg)
