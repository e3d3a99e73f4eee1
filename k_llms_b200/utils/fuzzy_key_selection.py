"""Join-key selection with a fuzzy second opinion (host logic, as in the reference).

Mirrors the reference's `k_llms/utils/fuzzy_key_selection.py`: run the standard selection (`key_selection.py`), then the
same funnel over single keys whose values were first bucketed — numbers rounded to `fuzzy_numeric_round_decimals`
(1.29 and 1.294 meet), strings lower-cased and whitespace-collapsed — and prefer the fuzzy winner only when its
stability rank (worst-pair Jaccard, values in all, values in all but one, mean Jaccard) is strictly higher than the
standard best SINGLE key's.  Pinned by `tests/golden/key_alignment.json` (reference outputs).
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional

from pydantic import BaseModel, ConfigDict

from .key_selection import (CascadeConfig, KeyMetrics, _funnel, _stability, _Table, discover_scalar_paths,
                            select_best_keys as select_best_keys_standard)


def _normalize_string(value: str) -> str:
    """Reference `:29-34`."""
    return " ".join(value.strip().lower().split())


def _canonicalize_scalar(value: Any, numeric_round_decimals: int) -> Any:
    """Bucket one scalar (reference `:37-52`): numbers (not bools) -> round(float(v), decimals), kept as they are when
    that fails (an int too large for a float); strings normalised; the rest unchanged."""
    if isinstance(value, (int, float)) and not isinstance(value, bool):
        try:
            return round(float(value), numeric_round_decimals)
        except Exception:
            return value
    if isinstance(value, str):
        return _normalize_string(value)
    return value


def _bucketing(decimals: int) -> Callable[[Any], Any]:
    return lambda v: _canonicalize_scalar(v, decimals)


def _evaluate_single_key_fuzzy(extractions: List[Dict[str, Any]], path: str, list_key: Optional[str], numeric_round_decimals: int) -> KeyMetrics:
    """Reference `:55-94`."""
    return _Table(extractions, list_key).single(path, _bucketing(numeric_round_decimals))


def _cascade_select_keys_fuzzy(extractions: List[Dict[str, Any]], candidates: List[str], config: CascadeConfig,
                               list_key: Optional[str], numeric_round_decimals: int) -> KeyMetrics:
    """The funnel of `key_selection._funnel` over bucketed single keys (reference `:101-160`)."""
    table = _Table(extractions, list_key)
    canon = _bucketing(numeric_round_decimals)
    return _funnel([table.single(p, canon) for p in candidates], config, "No keys pass Stage 0 (fuzzy)").final_best


_stability_tuple = _stability


class SelectionComparison(BaseModel):
    """What both selections found and which one is used: "normal" | "fuzzy" (reference `:163-176`)."""
    model_config = ConfigDict(frozen=True)

    normal_best: Optional[KeyMetrics]
    fuzzy_best: Optional[KeyMetrics]
    chosen: str


def select_best_keys_with_fuzzy_fallback(extractions: List[Dict[str, Any]], cascade_cfg: CascadeConfig = CascadeConfig(),
                                         list_key: Optional[str] = None, fuzzy_numeric_round_decimals: int = 2,
                                         enable_fuzzy_fallback: bool = True, prefer_fuzzy_if_better: bool = True) -> SelectionComparison:
    """Reference `:179-235`.  ValueError when neither selection finds a key."""
    try:
        normal = select_best_keys_standard(extractions, cascade_cfg=cascade_cfg, list_key=list_key).best_single
    except ValueError:
        normal = None

    fuzzy = None
    if enable_fuzzy_fallback:
        candidates = discover_scalar_paths(extractions, list_key=list_key)
        if candidates:
            try:
                fuzzy = _cascade_select_keys_fuzzy(extractions, candidates, cascade_cfg, list_key, fuzzy_numeric_round_decimals)
            except ValueError:
                fuzzy = None

    if normal is None and fuzzy is None:
        raise ValueError("No keys pass Stage 0 (normal or fuzzy)")
    if fuzzy is None:
        return SelectionComparison(normal_best=normal, fuzzy_best=None, chosen="normal")
    if normal is None:
        return SelectionComparison(normal_best=None, fuzzy_best=fuzzy, chosen="fuzzy")
    better = prefer_fuzzy_if_better and _stability(fuzzy) > _stability(normal)
    return SelectionComparison(normal_best=normal, fuzzy_best=fuzzy, chosen="fuzzy" if better else "normal")
