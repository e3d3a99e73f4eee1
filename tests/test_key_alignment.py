"""The key-based aligner (k_llms_b200/utils/{key_based_alignment,key_selection,fuzzy_key_selection}.py) against
the reference's own outputs (tests/golden/key_alignment.json, made by oracle/gen_golden_key_alignment.py) and against the
properties of a join that need no reference.  Host logic: CPU only."""
import copy
import json
import random

import pytest

from tests.helpers import load_golden

from k_llms_b200.utils import fuzzy_key_selection as FZ
from k_llms_b200.utils import key_based_alignment as KA
from k_llms_b200.utils import key_selection as KS


def _metrics(m):
    return None if m is None else {"path": list(m.path), **{k: getattr(m, k) for k in (
        "coverage_min", "coverage_mean", "uniqueness_min", "uniqueness_mean", "jaccard_min", "jaccard_mean", "I_E", "I_E_minus_1",
        "I_ge_2", "union_size")}, "score_tuple": list(m.score_tuple)}


def _same(a, b):
    return json.dumps(a, sort_keys=False) == json.dumps(b, sort_keys=False)   # key order, 1 vs 1.0 vs true all count


CASES = load_golden("key_alignment")


def test_golden_file_covers_every_kind():
    kinds = {}
    for c in CASES:
        k = c["kind"] + ("/raises" if "raises" in c else "")
        kinds[k] = kinds.get(k, 0) + 1
    assert kinds["align"] > 200 and kinds["align/raises"] > 20 and kinds["select"] > 80 and kinds["fuzzy"] > 80, kinds
    assert sum(1 for c in CASES if c.get("chosen") == "fuzzy") >= 5
    assert sum(1 for c in CASES if c["kind"] == "select" and c.get("best_composite") and len(c["best_composite"]["path"]) > 1) >= 10


def test_recursive_align_equals_the_reference():
    for c in (c for c in CASES if c["kind"] == "align"):
        values = copy.deepcopy(c["values"])
        if "raises" in c:
            with pytest.raises(Exception) as info:
                KA.recursive_align(values, "levenshtein", **c["kwargs"])
            assert type(info.value).__name__ == c["raises"], c["values"]
            continue
        views, mapping = KA.recursive_align(values, "levenshtein", **c["kwargs"])
        assert _same(views, c["views"]), (c["values"], c["kwargs"])
        assert _same(mapping, c["mapping"]), (c["values"], c["kwargs"])
        assert _same(values, c["values"])       # inputs are not modified


def test_select_best_keys_equals_the_reference():
    for c in (c for c in CASES if c["kind"] == "select"):
        args = dict(cascade_cfg=KS.CascadeConfig(**c["cfg"]), list_key=c["list_key"], **c["kwargs"])
        if "raises" in c:
            with pytest.raises(ValueError) as info:
                KS.select_best_keys(c["extractions"], **args)
            assert str(info.value) == c["raises"]
            continue
        res = KS.select_best_keys(c["extractions"], **args)
        assert _same(_metrics(res.best_single), c["best_single"]), c["extractions"]
        assert _same(_metrics(res.best_composite), c["best_composite"]), c["extractions"]
        assert [list(m.path) for m in res.candidate_table] == c["table"]
        assert res.min_support_for_autolock == c["autolock"]
        rep = res.cascade_report
        assert [[list(m.path) for m in st] for st in (rep.stage0_kept, rep.stage1_kept, rep.stage2_kept, rep.stage3_kept)] == c["stages"]
        assert rep.final_best == res.best_single
        assert KS.discover_scalar_paths(c["extractions"], list_key=c["list_key"]) == c["candidates"]
        # the public single / composite evaluators agree with what the selection reported
        assert _same(_metrics(KS.evaluate_single_key(c["extractions"], res.best_single.path[0], list_key=c["list_key"])), c["best_single"])
        if res.best_composite is not None:
            again = KS.evaluate_composite_key(c["extractions"], list(res.best_composite.path), list_key=c["list_key"])
            assert _same(_metrics(again), c["best_composite"])
            assert _same(_metrics(KS.cascade_select_keys(c["extractions"], c["candidates"], KS.CascadeConfig(**c["cfg"]),
                                                         list_key=c["list_key"]).final_best), c["best_single"])


def test_fuzzy_fallback_equals_the_reference():
    for c in (c for c in CASES if c["kind"] == "fuzzy"):
        args = dict(cascade_cfg=KS.CascadeConfig(**c["cfg"]), list_key=c["list_key"], **c["kwargs"])
        if "raises" in c:
            with pytest.raises(ValueError) as info:
                FZ.select_best_keys_with_fuzzy_fallback(c["extractions"], **args)
            assert str(info.value) == c["raises"]
            continue
        comp = FZ.select_best_keys_with_fuzzy_fallback(c["extractions"], **args)
        assert comp.chosen == c["chosen"], c["extractions"]
        assert _same(_metrics(comp.normal_best), c["normal"]) and _same(_metrics(comp.fuzzy_best), c["fuzzy_best"])


def test_value_access_helpers():
    ext = {"products": [{"a": " X  y ", "b": {"c": 1}}, {"a": None, "b": {"c": [1]}}, "junk", {"a": "x Y"}], "other": [{"a": "z"}]}
    assert KS.normalize_scalar("  A \t\n B ") == "a b" and KS.normalize_scalar(3) == 3
    assert len(KS.iter_records(ext)) == 3 and KS.iter_records(ext, list_key="other") == [{"a": "z"}]
    assert KS.iter_records({"u": [{"k": 1}], "v": 3, "w": [{"k": 2}, 5]}) == [{"k": 1}, {"k": 2}]      # auto-detect: every list
    assert KS.iter_records(ext, list_key="missing") == [] and KS.iter_records({"products": "no list"}) == []
    assert KS.values_for_path(ext, "a") == ["x y", "x y"] and KS.values_for_path(ext, "b.c") == [1]
    assert KS.tuple_values_for_paths(ext, ["a", "b.c"]) == [("x y", 1)]
    assert KS.discover_scalar_paths([ext]) == ["a", "b.c"]                                            # None counts, a list does not
    assert KS.jaccard(set(), set()) == 1.0 and KS.jaccard({1}, set()) == 0.0 and KS.jaccard({1, 2}, {2, 3}) == 1 / 3
    assert FZ._canonicalize_scalar(1.294, 2) == 1.29 and FZ._canonicalize_scalar(True, 2) is True
    assert FZ._canonicalize_scalar(10 ** 400, 2) == 10 ** 400 and FZ._canonicalize_scalar("  A  b ", 2) == "a b"
    assert KA._get_value_by_path({"a": [{"b": 5}], "0": 1}, "a.0.b") == 5 and KA._get_value_by_path({"0": 1}, "0") is None
    assert KA._get_value_by_path({"a": 1}, "") == {"a": 1} and KA._get_value_by_path({"a": 1}, None) is None
    with pytest.raises(ValueError):
        KS.select_best_keys([])
    with pytest.raises(ValueError):
        KA._materialize_source_view({}, {}, 0)


def test_join_properties_on_fresh_random_sources():
    """What a join on the selected key must satisfy whatever the key: per source, an item appears in at most one row and at its
    own key's row; the first of the longest lists keeps its order; an aligned row's sources all share the row's key."""
    from oracle.gen_golden_key_alignment import random_sources
    rng = random.Random(77)
    joined = 0
    for _ in range(300):
        lists = [s["items"] for s in random_sources(rng, "dict") if s is not None]
        if len(lists) < 2 or not all(isinstance(i, dict) for lst in lists for i in lst):
            continue
        key = KA._Joiner(KS.CascadeConfig(min_coverage=0.5, min_uniqueness=0.5)).join_key(lists)
        if not key:
            continue
        try:
            rows, positions = KA._align_lists_by_key(lists, key)
        except TypeError:      # leftover keys of different types do not sort, as upstream
            continue
        joined += 1
        lead = max(range(len(lists)), key=lambda i: len(lists[i]))
        lead_positions = [pos[lead] for pos in positions if pos[lead] is not None]
        assert lead_positions == sorted(lead_positions)
        for s, lst in enumerate(lists):
            used = [pos[s] for pos in positions if pos[s] is not None]
            assert len(used) == len(set(used))
            first_keys = {}
            for i, item in enumerate(lst):
                k = KA._get_key_tuple(item, key)
                if k is not None:
                    first_keys.setdefault(k, i)
            assert sorted(used) == sorted(first_keys.values())
        for row, pos in zip(rows, positions):
            keys = {KA._get_key_tuple(item, key) for item in row if item is not None}
            assert len(keys) == 1
            for s, item in enumerate(row):
                assert (item is None) == (pos[s] is None) and (item is None or item is lists[s][pos[s]])
    assert joined > 100


def test_trace_switches(tmp_path, capsys):
    KA.VERBOSE, KA.LOG_FILE = True, str(tmp_path / "trace.log")
    try:
        KA.recursive_align([{"a": [{"id": 1}, {"id": 2}]}, {"a": [{"id": 2}, {"id": 1}]}], "levenshtein")
    finally:
        KA.VERBOSE, KA.LOG_FILE = False, None
    assert "[KEY-SELECT]" in capsys.readouterr().out and "[KEY-SELECT]" in (tmp_path / "trace.log").read_text()
