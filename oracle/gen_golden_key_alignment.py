"""Golden vectors of the reference's KEY-BASED aligner (TEST INFRASTRUCTURE; run in the build container, where
/root/reference exists):  python -m oracle.gen_golden_key_alignment  ->  tests/golden/key_alignment.json

The reference ships no tests for `key_based_alignment.py` / `key_selection.py` / `fuzzy_key_selection.py` (SURVEY §2 row 14:
nothing on its client path imports them), so its own outputs on these inputs are the pin for
`k_llms_b200/utils/{key_based_alignment,key_selection,fuzzy_key_selection}.py`.  Three kinds of case:

  align   recursive_align(values, "levenshtein", **kwargs)        -> [per-source views, mapping]  or the exception's type
  select  select_best_keys(extractions, list_key=..., cascade_cfg) -> paths, metrics, funnel stages or ValueError text
  fuzzy   select_best_keys_with_fuzzy_fallback(...)               -> chosen, both winners          or ValueError text

`random_sources` / `HAND_ALIGN` are imported by tests/test_key_alignment.py, which draws fresh random cases for the properties
(row order, first-occurrence join) that need no reference."""
from __future__ import annotations

import copy
import importlib
import json
import os
import random

from oracle.gen_golden import GOLDEN_DIR

NAMES = ["Widget", "Gadget", "Sprocket", "Flange", "Gasket", "Bolt M8", "Bolt M10", "Washer", "Bearing 6204", "O-Ring"]


def _truth_items(rng: random.Random, count: int) -> list:
    names = rng.sample(NAMES, min(count, len(NAMES)))
    items = []
    for i in range(count):
        item = {"id": 100 + i if rng.random() < 0.8 else rng.choice([100, 101]), "name": names[i % len(names)],
                "price": round(rng.uniform(1, 50), 2), "unit": rng.choice(["kg", "pc", "m"]), "qty": rng.randrange(1, 5)}
        if rng.random() < 0.6:
            item["meta"] = {"sku": f"SKU-{rng.randrange(10, 99)}", "origin": {"code": rng.choice(["FR", "DE", "US"])}}
        if rng.random() < 0.3:
            item["tags"] = rng.sample(["a", "b", "c"], rng.randrange(0, 3))
        if rng.random() < 0.25:
            item["parts"] = [{"pn": f"P{j}", "n": rng.randrange(1, 4)} for j in range(rng.randrange(1, 4))]
        items.append(item)
    return items


def _noisy(rng: random.Random, item: dict, level: float) -> dict:
    out = copy.deepcopy(item)
    for key in list(out):
        r = rng.random()
        if r > level:
            continue
        v = out[key]
        if key == "name":
            out[key] = rng.choice([v.upper(), v.lower(), f" {v}  ", v.replace(" ", "  "), v + "s", None])
        elif key == "price":
            out[key] = rng.choice([round(v + 0.004, 3), round(v + 0.01, 2), v + 1, str(v), None])
        elif key == "id":
            out[key] = rng.choice([float(v), str(v), v + 50, None])
        elif key == "qty":
            out[key] = rng.choice([v, v + 1, True])
        elif key == "meta":
            if rng.random() < 0.5:
                v["sku"] = v["sku"].lower()
            else:
                del out[key]
        elif key == "parts":
            rng.shuffle(v)
        elif key == "tags":
            if rng.random() < 0.1:
                del out[key]        # a list a source lacks: the positional zip then raises upstream (len(None))
        elif rng.random() < 0.3:
            del out[key]
    return out


def random_sources(rng: random.Random, shape: str | None = None) -> list:
    """n candidate outputs holding noisy, reordered, truncated or extended copies of one list of items."""
    n = rng.choice([2, 3, 3, 4, 5])
    truth = _truth_items(rng, rng.randrange(0, 7))
    level = rng.choice([0.0, 0.05, 0.15, 0.4])
    shape = shape or rng.choice(["dict", "dict", "dict", "list", "nested"])
    sources = []
    for _ in range(n):
        items = [_noisy(rng, it, level) for it in truth]
        r = rng.random()
        if r < 0.35:
            rng.shuffle(items)
        elif r < 0.5 and items:
            items.pop(rng.randrange(len(items)))
        elif r < 0.6:
            items.append(_noisy(rng, _truth_items(rng, 1)[0], level))
        elif r < 0.65 and items:
            items.append(copy.deepcopy(items[0]))
        if shape == "list":
            src = items
        elif shape == "nested":
            src = {"doc": {"lines": items, "total": round(sum((i.get("qty") or 0) for i in items if isinstance(i.get("qty"), int)), 2)},
                   "vendor": rng.choice(["ACME", "Acme", "acme corp"])}
        else:
            src = {"items": items, "count": len(items), "currency": rng.choice(["EUR", "eur", "USD"])}
            if rng.random() < 0.2:
                src["notes"] = rng.sample(["late", "fragile", "paid"], rng.randrange(0, 3))
        if rng.random() < 0.03:
            src = None
        sources.append(src)
    return sources


def keyed_sources(rng: random.Random, mode: str) -> list:
    """Sources whose only usable join key needs the fuzzy buckets ("price": floats that differ in the third decimal, strings in
    case / spacing) or a composite ("pair": first and last names, neither unique alone)."""
    n = rng.choice([2, 3, 4])
    count = rng.randrange(2, 7)
    if mode == "price":
        base = rng.sample(range(100, 4000), count)
        truth = [{"price": b / 100.0, "note": "n"} for b in base]
    else:
        firsts, lasts = rng.sample(["ada", "alan", "grace", "edsger"], 2), rng.sample(["king", "hopper", "turing", "lovelace"], 3)
        pairs = rng.sample([(f, l) for f in firsts for l in lasts], min(count, 6))
        truth = [{"first": f, "last": l, "born": rng.choice([1906, 1912, 1930])} for f, l in pairs]
    sources = []
    for _ in range(n):
        items = copy.deepcopy(truth)
        for it in items:
            if mode == "price":
                it["note"] = rng.choice(["ok", "check", "see invoice", "n/a", "-"])       # free text: no stable key
                if rng.random() < 0.5:
                    it["price"] = round(it["price"] + rng.choice([0.001, 0.002, 0.003, -0.002, 0.004]), 3)
            if mode == "pair" and rng.random() < 0.1:
                it["last"] = it["last"].upper()
        r = rng.random()
        if r < 0.5:
            rng.shuffle(items)
        elif r < 0.65 and len(items) > 1:
            items.pop(rng.randrange(len(items)))
        sources.append({"items": items})
    return sources


HAND_ALIGN = [
    ([{"a": [{"id": 1, "v": 2}, {"id": 2, "v": 3}]}, {"a": [{"id": 2, "v": 3}, {"id": 1, "v": 2.5}]}], {}),
    ([[{"id": 1, "v": 2}, {"id": 2, "v": 3}], [{"id": 2, "v": 3}, {"id": 1, "v": 2.5}]], {}),
    ([[{"id": 1, "v": 2}, {"id": 2, "v": 3}], [{"id": 2, "v": 3}, {"id": 1, "v": 2.5}], [{"id": 2, "v": 9}]], {}),
    ([{"a": [1, 2]}, {"a": [1]}], {}), ([{"a": [1, 2]}, {"b": [1]}], {}), ([{"a": [1, 2]}, None], {}),
    ([None, None], {"current_path": "x"}), ([None, None], {}), ([], {}), ([1, 2], {}), (["a", None], {}), ([{}, {}], {}), ([[], []], {}),
    ([{"a": {"b": 1}}, {"a": {"b": 2, "c": None}}], {"current_path": "root"}),
    ([{"a": {"b": 1}}, {"a": {"b": 2, "c": None}}], {}),
    ([{"a": [{"id": "X 1", "v": 2}, {"id": "y", "v": 3}]}, {"a": [{"id": "x  1", "v": 3}, {"id": "Y", "v": 2.5}]}], {}),
    ([{"a": [{"id": 1, "v": "p"}, {"id": "b", "v": "q"}]}, {"a": [{"id": "b", "v": "q"}, {"id": 1, "v": "p"}, {"id": 7, "v": "r"}]},
      {"a": [{"id": 1, "v": "p"}, {"id": "zz", "v": "s"}]}], {}),                                       # unsortable leftovers
    ([{"a": [{"id": 1, "v": "p"}, {"id": 2, "v": "q"}]}, {"a": [{"id": 1.0, "v": "p"}, {"id": True, "v": "z"}, {"id": 2, "v": "q"}]}], {}),
    ([{"a": [{"k": {"x": 1}, "v": 1}, {"k": {"x": 2}, "v": 2}]}, {"a": [{"k": {"x": 2}, "v": 2}, {"k": {"x": 1}, "v": 1}]}], {}),
    ([{"a": [{"p": 1.29, "n": "u"}, {"p": 2.5, "n": "u"}]}, {"a": [{"p": 2.5, "n": "u"}, {"p": 1.294, "n": "u"}]}], {}),  # fuzzy only
    ([{"a": [{"f": "x", "l": "1"}, {"f": "x", "l": "2"}, {"f": "y", "l": "1"}]},
      {"a": [{"f": "y", "l": "1"}, {"f": "x", "l": "2"}, {"f": "x", "l": "1"}]}], {"min_uniqueness": 0.0}),               # composite
    ([{"a": [{"f": "x", "l": "1"}, {"f": "x", "l": "2"}, {"f": "y", "l": "1"}]},
      {"a": [{"f": "y", "l": "1"}, {"f": "x", "l": "2"}, {"f": "x", "l": "1"}]}], {}),
    ([{"a": [{"id": 1}, {"id": 2}, {"id": 3}]}, {"a": [{"id": 3}]}, {"a": [{"id": 2}, {"id": 1}]}], {"min_support_ratio": 0.2}),
    ([{"a": [{"id": 1}, {"id": 2}, {"id": 3}]}, {"a": [{"id": 3}]}, {"a": [{"id": 2}, {"id": 1}]}], {"min_coverage": 0.9}),
    ([{"a": [{"id": 1}, {"id": 1}, {"id": 2}]}, {"a": [{"id": 2}, {"id": 1}]}], {"min_uniqueness": 0.0}),
    ([{"0": {"1": "x"}, "l": [{"0": 1}]}, {"0": {"1": "y"}, "l": [{"0": 1}]}], {}),                      # int-looking dict keys
    ([{"a.b": 1, "a": {"b": 2}}, {"a.b": 3, "a": {"b": 4}}], {}),                                        # dotted key names
    ([{"": 1, "x": {"": 2}}, {"": 3, "x": {"": 4}}], {}),                                                # empty key names
    ([{"a": [[{"id": 1}], [{"id": 2}]]}, {"a": [[{"id": 2}], [{"id": 1}]]}], {}),                        # list of lists: zipped
    ([{"a": [{"id": 1}, "s"]}, {"a": [{"id": 1}]}], {}),                                                 # mixed list: zipped
    ([{"a": 1}, {"a": [1]}, {"a": {"b": 1}}], {}), ([{"a": True}, {"a": 1}], {}), ([{"a": [1]}, {"a": "s"}], {}),
    ([{"a": [{"id": 10 ** 400, "n": "x"}, {"id": 5, "n": "y"}]}, {"a": [{"id": 5, "n": "y"}, {"id": 10 ** 400, "n": "x"}]}], {}),
]


def _metrics(m) -> dict | None:
    return None if m is None else {"path": list(m.path), **{k: getattr(m, k) for k in (
        "coverage_min", "coverage_mean", "uniqueness_min", "uniqueness_mean", "jaccard_min", "jaccard_mean", "I_E", "I_E_minus_1",
        "I_ge_2", "union_size")}, "score_tuple": list(m.score_tuple)}


def selection_extractions(rng: random.Random):
    """(extractions, list_key) in the shapes select_best_keys accepts: a named list key, the default 'products', auto-detect."""
    r = rng.random()
    sources = keyed_sources(rng, "price") if r < 0.15 else keyed_sources(rng, "pair") if r < 0.3 else \
        [s for s in random_sources(rng, "dict") if s is not None]
    mode = rng.choice(["items", "products", "auto"])
    if mode == "items":
        return sources, "items"
    if mode == "products":
        return [{"products": s["items"], "other": [{"z": 1}]} for s in sources], None
    return [{"first": s["items"], "second": [{"name": "extra"}], "n": 3} for s in sources], None


def main() -> None:
    from oracle.ref_loader import load_reference
    load_reference()
    ka = importlib.import_module("k_llms.utils.key_based_alignment")
    ks = importlib.import_module("k_llms.utils.key_selection")
    fz = importlib.import_module("k_llms.utils.fuzzy_key_selection")

    cases = []
    rng = random.Random(20260921)
    align_inputs = list(HAND_ALIGN)
    for _ in range(260):
        kwargs = rng.choice([{}, {}, {}, {"min_uniqueness": 0.0}, {"min_support_ratio": 0.8}, {"current_path": "payload"},
                             {"min_coverage": 0.3, "min_uniqueness": 0.3}])
        align_inputs.append((random_sources(rng), kwargs))
    for i in range(60):
        align_inputs.append((keyed_sources(rng, "price" if i % 2 else "pair"), rng.choice([{}, {"min_uniqueness": 0.0}])))
    for values, kwargs in align_inputs:
        case = {"kind": "align", "values": values, "kwargs": kwargs}
        try:
            views, mapping = ka.recursive_align(copy.deepcopy(values), "levenshtein", **kwargs)
            case["views"], case["mapping"] = views, mapping
        except Exception as exc:
            case["raises"] = type(exc).__name__
        cases.append(case)

    for _ in range(120):
        extractions, list_key = selection_extractions(rng)
        cfg = rng.choice([{}, {"min_coverage": 0.5, "min_uniqueness": 0.5}, {"min_uniqueness": 0.3, "topk_stage1": 3, "topk_stage2": 2, "topk_stage3": 2}])
        kw = rng.choice([{}, {}, {"max_k": 2}, {"max_candidates_for_composite": 2}])
        case = {"kind": "select", "extractions": extractions, "list_key": list_key, "cfg": cfg, "kwargs": kw}
        try:
            res = ks.select_best_keys(copy.deepcopy(extractions), cascade_cfg=ks.CascadeConfig(**cfg), list_key=list_key, **kw)
            rep = res.cascade_report
            case.update(best_single=_metrics(res.best_single), best_composite=_metrics(res.best_composite),
                        table=[list(m.path) for m in res.candidate_table], autolock=res.min_support_for_autolock,
                        stages=[[list(m.path) for m in st] for st in (rep.stage0_kept, rep.stage1_kept, rep.stage2_kept, rep.stage3_kept)],
                        candidates=ks.discover_scalar_paths(extractions, list_key=list_key))
        except ValueError as exc:
            case["raises"] = str(exc)
        cases.append(case)
        fcase = {"kind": "fuzzy", "extractions": extractions, "list_key": list_key, "cfg": cfg,
                 "kwargs": rng.choice([{}, {}, {"fuzzy_numeric_round_decimals": 0}, {"enable_fuzzy_fallback": False}, {"prefer_fuzzy_if_better": False}])}
        try:
            comp = fz.select_best_keys_with_fuzzy_fallback(copy.deepcopy(extractions), cascade_cfg=ks.CascadeConfig(**cfg), list_key=list_key,
                                                           **fcase["kwargs"])
            fcase.update(chosen=comp.chosen, normal=_metrics(comp.normal_best), fuzzy_best=_metrics(comp.fuzzy_best))
        except ValueError as exc:
            fcase["raises"] = str(exc)
        cases.append(fcase)

    meta = {"generator": "oracle/gen_golden_key_alignment.py", "reference": "retab-dev/k-LLMs @ 089dba9 behind 3 import stubs",
            "entry": "key_based_alignment.recursive_align, key_selection.select_best_keys, fuzzy_key_selection.select_best_keys_with_fuzzy_fallback"}
    with open(os.path.join(GOLDEN_DIR, "key_alignment.json"), "w") as f:
        json.dump({"meta": meta, "cases": cases}, f, separators=(",", ":"))
    by_kind = {}
    for c in cases:
        k = c["kind"] + ("/raises" if "raises" in c else "")
        by_kind[k] = by_kind.get(k, 0) + 1
    print(f"wrote key_alignment.json: {len(cases)} cases {by_kind}")


def fuzz(count: int, seed: int = 1) -> None:
    """Differential run: the product modules against the reference on `count` fresh random inputs (build container only)."""
    from oracle.ref_loader import load_reference
    load_reference()
    ref = importlib.import_module("k_llms.utils.key_based_alignment")
    from k_llms_b200.utils import key_based_alignment as mine
    rng = random.Random(seed)

    def outcome(fn, values, kwargs):
        try:
            return json.dumps(fn(copy.deepcopy(values), "levenshtein", **kwargs))
        except Exception as exc:
            return "raises " + type(exc).__name__

    bad = 0
    for i in range(count):
        values = random_sources(rng) if i % 4 else keyed_sources(rng, "price" if i % 8 else "pair")
        kwargs = rng.choice([{}, {"min_uniqueness": 0.0}, {"min_support_ratio": 0.8}, {"current_path": "p"}, {"min_coverage": 0.2}])
        if outcome(ref.recursive_align, values, kwargs) != outcome(mine.recursive_align, values, kwargs):
            bad += 1
            print("MISMATCH", json.dumps(values), kwargs)
    print(f"fuzz: {count} inputs, {bad} mismatches")


if __name__ == "__main__":
    import sys
    if len(sys.argv) > 2 and sys.argv[1] == "--fuzz":
        fuzz(int(sys.argv[2]))
    else:
        main()
