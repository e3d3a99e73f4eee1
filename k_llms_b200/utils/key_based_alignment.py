"""Key-based list aligner (host logic, as in the reference).

The alternative to the similarity aligner (`consensus_utils.recursive_list_alignments`): lists of dicts are JOINED on a
key chosen by `key_selection` / `fuzzy_key_selection` instead of being matched by pairwise similarity.  Mirrors the
reference's `k_llms/utils/key_based_alignment.py` — `recursive_align` has the similarity aligner's signature and
result shape (per-source aligned values, mapping aligned path -> per-source original path) — including the upstream
behaviours a caller would see, each pinned by `tests/golden/key_alignment.json` (reference outputs):

  * the key is selected on NORMALISED values but the join compares RAW values (`:50-70` vs `key_selection.py:24`), so
    "X 1" and "x  1" select a key together and then do not join;
  * rows follow the longest source's order (first of the longest), then the remaining keys in sorted order (`:108-131`) —
    a TypeError when those keys do not compare (str vs int), as upstream;
  * lists of scalars (or no usable key) are zipped by position (`:318-337`); there a source that is None while its path
    is known raises TypeError, as upstream (`len(None)`, `:325`);
  * a LIST root is materialised under "items" with the mapping keys re-prefixed once per list-rooted source (`:384-390`),
    so such sources read the aligned (first non-null) leaves, and the mapping keys carry one "items." per source;
  * with a `current_path` the mapped paths are prefixed before the lookup in the un-prefixed source, which then finds
    nothing (`:403-419`).
Upstream nothing on the client path imports this module; neither does this package's client path.

Own structure: one `_Joiner` walks the candidates once, sharing a single record table per list between the standard and
the fuzzy selection (the reference runs the standard selection twice); paths are joined by one helper.
"""
from __future__ import annotations

import logging
from copy import deepcopy
from typing import Any, Dict, List, Optional, Sequence, Tuple

from .fuzzy_key_selection import _bucketing
from .key_selection import CascadeConfig, KeySelectionResult, _funnel, _select, _stability, _Table, discover_scalar_paths

JSONScalar = str | int | float | bool | None
JSONLike = Dict[str, Any] | List[Any] | JSONScalar
Mapping = Dict[str, List[Optional[str]]]

VERBOSE: bool = False            # print the selection trace (the reference's CLI switch, `:27`)
LOG_FILE: Optional[str] = None   # and / or append it to this file
_logger = logging.getLogger(__name__)


def _log(msg: str) -> None:
    _logger.debug(msg)
    if VERBOSE:
        print(msg)
    if LOG_FILE:
        try:
            with open(LOG_FILE, "a", encoding="utf-8") as fh:
                fh.write(msg + "\n")
        except OSError:
            pass


def _join(prefix: Optional[str], token: Any) -> str:
    """`prefix.token`, or just the token at the root ("" / None prefix)."""
    return f"{prefix}.{token}" if prefix else str(token)


def _get_key_tuple(obj: Dict[str, Any], paths: Tuple[str, ...]) -> Optional[Tuple[Any, ...]]:
    """RAW scalar values of `paths` in one item; None if any is unreachable, None, a dict or a list (reference `:47-67`)."""
    out = []
    for path in paths:
        node: Any = obj
        for token in path.split("."):
            if not (isinstance(node, dict) and token in node):
                return None
            node = node[token]
        if node is None or isinstance(node, (dict, list)):
            return None
        out.append(node)
    return tuple(out)


def _align_lists_by_key(lists_to_align: Sequence[Optional[List[Dict[str, Any]]]], key_paths: Tuple[str, ...],
                        ) -> Tuple[List[List[Optional[Dict[str, Any]]]], List[List[Optional[int]]]]:
    """Outer join of the sources' lists on `key_paths` (reference `:70-150`): per row, each source's FIRST item with that
    key (or None) and that item's position in its list."""
    if not any(lists_to_align):
        return [], []
    first_at: List[Dict[Tuple[Any, ...], int]] = []
    in_order: List[List[Tuple[Any, ...]]] = []
    for source in lists_to_align:
        where: Dict[Tuple[Any, ...], int] = {}
        if isinstance(source, list):
            for pos, item in enumerate(source):
                key = _get_key_tuple(item, key_paths) if isinstance(item, dict) else None
                if key is not None:
                    where.setdefault(key, pos)
        first_at.append(where)
        in_order.append(list(where))          # dicts keep insertion order: first occurrences in list order

    lengths = [len(s) if isinstance(s, list) else 0 for s in lists_to_align]
    lead = lengths.index(max(lengths))        # the first of the longest sources sets the row order
    keys = in_order[lead]
    seen = set(keys)
    keys = keys + sorted(set().union(*first_at) - seen)

    rows = [[(src[where[k]] if k in where else None) for src, where in zip(lists_to_align, first_at)] for k in keys]
    positions = [[where.get(k) for where in first_at] for k in keys]
    return rows, positions


class _Joiner:
    def __init__(self, cascade_cfg: CascadeConfig):
        self.cfg = cascade_cfg

    # ---- which key joins these lists (reference `:217-300`) ----
    def join_key(self, lists: List[List[Any]]) -> Optional[Tuple[str, ...]]:
        extractions = [{"items": lst} for lst in lists]
        table = _Table(extractions, "items")
        candidates = discover_scalar_paths(extractions, list_key="items")
        if not candidates:
            _log("[KEY-SELECT] no key found (no scalar paths)")
            return None
        standard: Optional[KeySelectionResult]
        try:
            standard = _select(table, candidates, len(extractions), 20, 3, 0.75, self.cfg)
        except ValueError:
            standard = None
        canon = _bucketing(2)
        try:
            fuzzy = _funnel([table.single(p, canon) for p in candidates], self.cfg, "No keys pass Stage 0 (fuzzy)").final_best
        except ValueError:
            fuzzy = None

        if standard is None:
            if fuzzy is None:
                _log("[KEY-SELECT] no key found (standard failed, fuzzy failed)")
                return None
            _log(f"[KEY-SELECT fuzzy-only] path={list(fuzzy.path)}")
            return fuzzy.path
        single, combo = standard.best_single, standard.best_composite
        picked = combo if (combo and combo.score_tuple > single.score_tuple) else single
        # the fuzzy winner is weighed against the best SINGLE key, also when a composite was picked (reference `:243-252`)
        if fuzzy is not None and _stability(fuzzy) > _stability(single):
            _log(f"[KEY-SELECT] chosen=fuzzy path={list(fuzzy.path)} jaccard_min={round(fuzzy.jaccard_min, 6)}")
            return fuzzy.path
        _log(f"[KEY-SELECT] chosen=standard path={list(picked.path)} jaccard_min={round(picked.jaccard_min, 6)}")
        return picked.path

    # ---- the recursion (reference `:155-339`) ----
    def node(self, values: Sequence[Any], paths: Sequence[Optional[str]]) -> Tuple[Any, Mapping]:
        present = [v for v in values if v is not None]
        if not present:
            return None, {}
        kind = type(present[0])
        if kind not in (dict, list) or not all(isinstance(v, kind) for v in present):
            return deepcopy(present[0]), {"": list(paths)}       # leaf: the first non-null value stands for the row
        if kind is dict:
            return self.dicts([v if isinstance(v, dict) else {} for v in values], paths)
        return self.lists(values, [v if isinstance(v, list) else [] for v in values], paths)

    @staticmethod
    def _lift(into: Mapping, token: Any, sub: Mapping) -> None:
        for key, where in sub.items():
            into[f"{token}.{key}" if key else str(token)] = where

    def dicts(self, dicts: List[Dict[str, Any]], paths: Sequence[Optional[str]]) -> Tuple[Dict[str, Any], Mapping]:
        out: Dict[str, Any] = {}
        mapping: Mapping = {}
        for key in sorted({k for d in dicts for k in d}):
            below = [None if p is None else _join(p, key) for p in paths]
            out[key], sub = self.node([d.get(key) for d in dicts], below)
            self._lift(mapping, key, sub)
        return out, mapping

    def lists(self, values: Sequence[Any], lists: List[List[Any]], paths: Sequence[Optional[str]]) -> Tuple[List[Any], Mapping]:
        out: List[Any] = []
        mapping: Mapping = {}
        key_paths = None
        if all(isinstance(item, dict) for lst in lists for item in lst):
            key_paths = self.join_key(lists)
        if key_paths:
            rows, positions = _align_lists_by_key(lists, key_paths)
            for r, (row, pos) in enumerate(zip(rows, positions)):
                below = [None if (p is None or at is None) else _join(p, at) for p, at in zip(paths, pos)]
                item, sub = self.node(row, below)
                out.append(item)
                self._lift(mapping, r, sub)
            return out, mapping
        _log("[ALIGN] lists zipped by position (scalars or no key)")
        for r in range(max(len(lst) for lst in lists) if lists else 0):
            # len(values[j]) and not len(lists[j]): a None source with a known path is a TypeError upstream (`:325`)
            below = [None if p is None else (_join(p, r) if r < len(values[j]) else None) for j, p in enumerate(paths)]
            item, sub = self.node([lst[r] if r < len(lst) else None for lst in lists], below)
            out.append(item)
            self._lift(mapping, r, sub)
        return out, mapping


def _compute_key_aligned_structure(values: Sequence[Any], original_paths: Sequence[Optional[str]], cascade_cfg: CascadeConfig,
                                   ) -> Tuple[Any, Mapping]:
    """One aligned structure for all sources plus, per aligned leaf path, each source's original path (reference `:155-339`)."""
    if not values:
        return None, {}
    return _Joiner(cascade_cfg).node(values, original_paths)


def _get_value_by_path(obj: Any, path: Optional[str]) -> Any:
    """Follow a dot path; a token that parses as an int indexes a LIST (and yields None on anything else, a dict with
    that key included); "" is the root; empty tokens are skipped (reference `:428-462`)."""
    if path is None:
        return None
    node = obj
    for token in path.split("."):
        if token == "":
            continue
        try:
            index = int(token)
        except ValueError:
            if not (isinstance(node, dict) and token in node):
                return None
            node = node[token]
            continue
        if not (isinstance(node, list) and 0 <= index < len(node)):
            return None
        node = node[index]
    return node


def _materialize_source_view(aligned_node: Any, key_mappings: Mapping, source_idx: int, current_path: str = "",
                             source_root: Optional[Dict[str, Any]] = None) -> Any:
    """The aligned structure filled with ONE source's values: leaves are fetched through the mapping; a leaf without a
    mapping entry keeps the aligned value (reference `:465-515`)."""
    if source_root is None:
        raise ValueError("source_root must be provided at the top-level call.")
    if isinstance(aligned_node, dict):
        return {k: _materialize_source_view(v, key_mappings, source_idx, _join(current_path, k), source_root) for k, v in aligned_node.items()}
    if isinstance(aligned_node, list):
        return [_materialize_source_view(v, key_mappings, source_idx, _join(current_path, i), source_root) for i, v in enumerate(aligned_node)]
    where = key_mappings.get(current_path)
    if where is not None and 0 <= source_idx < len(where):
        return _get_value_by_path(source_root, where[source_idx])
    return aligned_node


def recursive_align(values: Sequence[JSONLike], string_similarity_method: str, min_support_ratio: float = 0.5,
                    max_novelty_ratio: float = 0.25, current_path: str = "", reference_idx: Optional[int] = None,
                    min_uniqueness: Optional[float] = None, min_coverage: Optional[float] = None,
                    ) -> tuple[Sequence[JSONLike], dict[str, list[str | None]]]:
    """Key-based counterpart of `recursive_list_alignments` (reference `:342-421`).  `string_similarity_method`,
    `max_novelty_ratio` and `reference_idx` are accepted for signature parity and unused, as upstream;
    `min_support_ratio` is the coverage gate unless `min_coverage` is given; the uniqueness gate defaults to 0.5."""
    if not values:
        return list(values), {}
    if all(v is None for v in values):
        return list(values), {current_path: [current_path for _ in values]}

    cfg = CascadeConfig(min_coverage=min_coverage if min_coverage is not None else min_support_ratio,
                        min_uniqueness=min_uniqueness if min_uniqueness is not None else 0.5)
    aligned, mapping = _compute_key_aligned_structure(values, [current_path for _ in values], cfg)

    views: List[JSONLike] = []
    for i, source in enumerate(values):
        if isinstance(source, dict):
            root: Dict[str, Any] = source
        elif isinstance(source, list):
            root = {"items": source}
            if mapping:   # re-keyed once per list-rooted source, the later sources (and the caller) see the result (`:384-390`)
                mapping = {(f"items.{k}" if k else "items"): v for k, v in mapping.items()}
        else:
            root = {}
        views.append(_materialize_source_view(aligned, mapping, i, "", root))

    if current_path:
        mapping = {(f"{current_path}.{key}" if key else current_path):
                   [current_path if (p is None or p == "") else f"{current_path}.{p}" for p in where]
                   for key, where in mapping.items()}
    return views, mapping
