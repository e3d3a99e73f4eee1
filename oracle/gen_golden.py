"""Generate the golden vectors under tests/golden/ by RUNNING THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference; see oracle/ref_loader.py for the three
import stubs).  The reference ships no tests, so these reference-produced
outputs are the pin for the oracle and for the CUDA path:

    python oracle/gen_golden.py            # rewrites tests/golden/*.json
    python oracle/gen_golden.py --fuzz N   # also differential-fuzz oracle/consensus_py.py vs the reference

Every case is `{"values": [...], "pvf": 1.0, "value": ..., "conf": ...}` with JSON-native payloads
(ASCII strings, bools, ints, floats incl. NaN/Infinity tokens, null, dicts, lists).
"""
from __future__ import annotations

import argparse
import json
import logging
import math
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import consensus_py as O  # noqa: E402
from oracle.ref_loader import load_reference, raising_embeddings, ref_client_order, ref_consensus_values  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# SURVEY.md §8c known-answer inputs (outputs are re-derived from the reference below, not typed in).
KNOWN_INPUTS = [
    [{"name": "John", "age": 30, "active": True, "city": "Paris"},
     {"name": "John", "age": 30, "active": True, "city": "paris"},
     {"name": "Jon", "age": 31, "active": False, "city": "Paris"}],
    ["A", "B"], ["B", "A"], ["b", "a", "a", "b"], ["Paris!", "paris", "PARIS ", "Lyon"],
    ["x", None, None, None], [None, None], [], ["a", "b", "c"],
    ["the big cat", "the big cat", "the big dog"],
    [True, True, False], [True, None, None], [True, False], [False, True],
    [5, 5, 5], [30, 30, 31], [100, 101, 102], [100, 102.9, 105.8], [1, 2], [1, 1, 2, 2],
    [5, 5, 50, 50, 7, 7], [1, None, None], [1, None], [float("nan"), 1.0, 1.0], [1, True, 1],
    ["30", "30", "31"], [30, 30, 31, None],
    [{"a": "x"}, {"a": "x"}, None, None], [{"a": 1}, {"a": 1}, None, None],
    [[1, 2, 3], [1, 2], [1, 9, 3]], [{"b": "x"}, {"a": "y", "b": "x"}],
    # extra edge cases beyond the survey table
    [0.0, -0.0, 0.0], [-5, 5, -5, 5], [1e-7, 2e-7, 5.0, 5.0], [10, 100, 1000, 1000, 10],
    [3, 3, 30, 30, 300], [1.0, 1.02, 1.04, 1.06, 1.08, 1.10, 7, 7], [float("inf"), 2, 2], [float("inf"), float("nan")],
    [1, "a"], [2 ** 60, 2 ** 60 + 1, 5], [True, 1, 1.0, None], [False, 0, None, ""], ["true", True, "TRUE"],
    ["", " ", "a"], ["a b", "A  B", "a-b", "c"], ["a b c", "a b c", "a b d", None],
    [{"reasoning___x": "r1", "v": "a"}, {"reasoning___x": "r2", "v": "a"}],
    [{"k": {"z": [1, 2]}}, {"k": {"z": [1, 3]}}, {"k": None}, "junk"],
    [[{"a": 1}, {"a": 2}], [{"a": 1}], []], [[], []], [[], None],
    [9] * 9 + [1] * 7, list(range(1, 17)), [1000 + i * 25 for i in range(16)],
    [0.1 * i for i in range(1, 33)], [7.25] * 3 + [7.3] * 5 + [100.0] * 8,
]

CLIENT_ORDER_INPUTS = [
    [{"b": "x", "a": 1}, {"b": "x"}, {"a": 1, "b": "y"}],
    [{"z": {"q": True}, "a": "k"}, {"a": "k", "z": {"q": False, "r": 1}}, {"a": "K!"}],
    # lists that arrive reordered / with missing or extra elements (alignment pre-pass, cu:550-613)
    [{"items": [{"name": "apple", "qty": 3}, {"name": "banana", "qty": 5}]},
     {"items": [{"name": "banana", "qty": 5}, {"name": "apple", "qty": 3}]},
     {"items": [{"name": "apple", "qty": 3}, {"name": "banana", "qty": 6}, {"name": "cherry", "qty": 1}]}],
    [{"tags": ["red", "green", "blue"]}, {"tags": ["green", "red", "blue"]}, {"tags": ["red", "blue"]}, {"tags": []}],
    [{"rows": [[1, 2], [3, 4]]}, {"rows": [[1, 2], [3, 5]]}, {"rows": [[3, 4], [1, 2]]}],
    [{"l": [1, 2, 3]}, {"l": [1, 2, 3]}, {"l": None}],
    [{"people": [{"first": "Ada", "last": "Lovelace"}, {"first": "Alan", "last": "Turing"}], "n": 2},
     {"people": [{"first": "Alan", "last": "Turing"}, {"first": "Ada", "last": "Lovelace"}], "n": 2},
     {"people": [{"first": "Ada", "last": "Lovelace"}], "n": 1},
     {"people": [{"first": "Grace", "last": "Hopper"}, {"first": "Ada", "last": "Lovelase"}, {"first": "Alan", "last": "Turing"}], "n": 3}],
]


def random_list_records(seed: int, count: int) -> list:
    """Candidate dicts whose list fields are shuffled / truncated / extended copies of one truth list."""
    rng = random.Random(seed)
    out = []
    for _ in range(count):
        n = rng.choice([2, 3, 4, 5, 6])
        kind = rng.choice(["str", "dict", "int"])
        length = rng.randrange(0, 6)
        if kind == "str":
            truth = [rng.choice(WORDS) + str(i) for i in range(length)]
        elif kind == "int":
            truth = [rng.randrange(1, 1000) * 10 ** i for i in range(length)]
        else:
            truth = [{"name": rng.choice(WORDS) + str(i), "qty": rng.randrange(1, 50), "ok": rng.random() < 0.5} for i in range(length)]
        cands = []
        for _c in range(n):
            lst = [json.loads(json.dumps(x)) for x in truth]
            if rng.random() < 0.4:
                rng.shuffle(lst)
            if lst and rng.random() < 0.3:
                lst.pop(rng.randrange(len(lst)))
            if rng.random() < 0.2:
                lst.append({"name": "extra", "qty": 1, "ok": True} if kind == "dict" else (rng.choice(WORDS) if kind == "str" else 7))
            if lst and kind == "dict" and rng.random() < 0.3:
                lst[rng.randrange(len(lst))]["qty"] = rng.randrange(1, 50)
            cands.append({"id": rng.choice(["A1", "A1", "a1", "B2"]), "items": lst if rng.random() > 0.05 else None})
        out.append(cands)
    return out

WORDS = ["alpha", "Bravo", "charlie", "DELTA", "echo", "fox-trot", "golf", "Hotel"]


def _variant(rng: random.Random, w: str) -> str:
    r = rng.random()
    if r < 0.15:
        return w.upper()
    if r < 0.3:
        return w.lower() + "!"
    if r < 0.4:
        return " " + w + " "
    return w


def _scalar_group(rng: random.Random, kind: str, n: int, p_agree: float, p_none: float) -> list:
    def draw():
        if kind == "str":
            return _variant(rng, rng.choice(WORDS))
        if kind == "bool":
            return rng.random() < 0.5
        if kind == "int":
            return rng.randrange(1, 10 ** rng.randrange(1, 7))
        if kind == "float":
            return rng.uniform(1.0, 1e4)
        if kind == "near":  # numbers that straddle the 3% tolerance, negative and tiny values
            return rng.choice([1.0, -1.0, 100.0, 1e-7, 0.0]) * (1.0 + rng.choice([-0.05, -0.029, 0, 0.015, 0.031]))
        if kind == "pow10":  # decimal-shift / sign-flip mistakes
            return rng.choice([12.5, 125.0, 1250.0, -12.5, 0.125])
        if kind == "phrase":
            return " ".join(rng.choice(WORDS) for _ in range(rng.randrange(1, 5)))
        raise ValueError(kind)
    truth = draw()
    vals = []
    for _ in range(n):
        v = truth if rng.random() < p_agree else draw()
        vals.append(None if rng.random() < p_none else v)
    return vals


def _record_candidates(rng: random.Random, n: int, depth: int) -> list:
    """n candidate dicts sharing one random schema (nested dicts / position-aligned lists)."""
    def schema(d):
        s = {}
        for i in range(rng.randrange(2, 6)):
            r = rng.random()
            if d > 0 and r < 0.2:
                s[f"o{i}"] = ("dict", schema(d - 1))
            elif d > 0 and r < 0.4:
                s[f"l{i}"] = ("list", rng.choice(["str", "int", "float", "bool"]), rng.randrange(0, 5))
            else:
                s[f"f{i}"] = ("leaf", rng.choice(["str", "bool", "int", "float", "near", "pow10"]))
        return s

    def fill(s, cols):
        out = {}
        for k, spec in s.items():
            if spec[0] == "leaf":
                out[k] = cols[("leaf", id(spec))].pop()
            elif spec[0] == "dict":
                out[k] = fill(spec[1], cols) if rng.random() > 0.08 else None
            else:
                length = max(0, spec[2] + rng.choice([0, 0, 0, -1, 1]))
                out[k] = [cols[("list", id(spec), j)].pop() if ("list", id(spec), j) in cols and cols[("list", id(spec), j)] else None
                          for j in range(length)]
            if rng.random() < 0.05:
                out.pop(k)
        return out

    def columns(s, cols):
        for spec in s.values():
            if spec[0] == "leaf":
                cols[("leaf", id(spec))] = _scalar_group(rng, spec[1], n, 0.75, 0.08)
            elif spec[0] == "dict":
                columns(spec[1], cols)
            else:
                for j in range(spec[2] + 1):
                    cols[("list", id(spec), j)] = _scalar_group(rng, spec[1], n, 0.75, 0.05)
        return cols

    s = schema(depth)
    cols = columns(s, {})
    return [fill(s, cols) for _ in range(n)]


def random_cases(seed: int, count: int) -> list:
    rng = random.Random(seed)
    cases = []
    for i in range(count):
        n = rng.choice([2, 3, 4, 5, 8, 16, 32])
        r = i % 10
        if r < 6:
            kind = ["str", "bool", "int", "float", "near", "pow10"][r]
            cases.append(_scalar_group(rng, kind, n, rng.choice([0.3, 0.6, 0.8, 0.95]), rng.choice([0.0, 0.05, 0.3])))
        elif r == 6:
            cases.append(_scalar_group(rng, "phrase", min(n, 8), 0.6, 0.1))
        else:
            cases.append(_record_candidates(rng, min(n, 8), depth=rng.choice([0, 1, 3])))
    return cases


PHRASE_WORDS = ("invoice total due amount net gross payment bank transfer within thirty days from receipt of goods and "
                "services the a an of to acme corp ltd gmbh street road avenue suite floor new york london paris berlin "
                "2024 2025 q1 q2 ref no id number 000123 77 ab-12 x y z").split()


def phrase_groups(seed: int, count: int) -> list:
    """Groups of 2..20 multi-word strings (noisy copies of one phrase): inputs of the similarity medoid, cu:1221-1237."""
    rng = random.Random(seed)
    out = []
    for _ in range(count):
        k = rng.randint(2, 20)
        base = [rng.choice(PHRASE_WORDS) for _ in range(rng.randint(3, 8))]
        grp = []
        for _c in range(k):
            words, r = list(base), rng.random()
            if r < 0.35:
                pass
            elif r < 0.6:
                words[rng.randrange(len(words))] = rng.choice(PHRASE_WORDS)
            elif r < 0.75:
                words = words[: max(1, len(words) - rng.randint(1, 2))]
            elif r < 0.9:
                words = words + [rng.choice(PHRASE_WORDS) for _ in range(rng.randint(1, 3))]
            else:
                words = [w.upper() if rng.random() < 0.5 else w + "," for w in words]
            grp.append("" if rng.random() < 0.03 else " ".join(words))
        if rng.random() < 0.1:
            grp[rng.randrange(k)] = " ".join(rng.choice(PHRASE_WORDS) for _ in range(40))
        if rng.random() < 0.15:
            grp[rng.randrange(k)] = None
        out.append(grp)
    return out


def _same(a, b) -> bool:
    if isinstance(a, float) and isinstance(b, float):
        return (math.isnan(a) and math.isnan(b)) or (a == b and math.copysign(1, a) == math.copysign(1, b))
    if type(a) is not type(b):
        return False
    if isinstance(a, dict):
        return list(a) == list(b) and all(_same(a[k], b[k]) for k in a)
    if isinstance(a, (list, tuple)):
        return len(a) == len(b) and all(_same(x, y) for x, y in zip(a, b))
    return a == b


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--fuzz", type=int, default=0)
    args = ap.parse_args()
    logging.disable(logging.CRITICAL)
    load_reference()
    os.makedirs(GOLDEN_DIR, exist_ok=True)

    known = []
    for vals in KNOWN_INPUTS:
        v, c = ref_consensus_values(vals)
        known.append({"values": vals, "value": v, "conf": c})
    client = []
    for vals in CLIENT_ORDER_INPUTS:
        v, c = ref_client_order(vals)
        client.append({"values": vals, "value": v, "conf": c})
    rnd = []
    for vals in random_cases(20260921, 600):
        v, c = ref_consensus_values(vals)
        rnd.append({"values": vals, "value": v, "conf": c})
    for vals in random_list_records(99, 120):
        v, c = ref_client_order(vals)
        client.append({"values": vals, "value": v, "conf": c})
    cu = load_reference()
    alignment = []
    for case in client:
        aligned, mapping = cu.recursive_list_alignments(case["values"], "embeddings", raising_embeddings, None, 0.51)
        alignment.append({"values": case["values"], "aligned": aligned, "key_mappings": mapping})
    medoid = []
    for grp in phrase_groups(4242, 250):
        for method in ("levenshtein", "embeddings"):
            settings = cu.ConsensusSettings(string_similarity_method=method)
            try:
                v, c = cu.consensus_values(grp, settings, raising_embeddings, None, 0.8)
            except Exception:  # two strings longer than 50 characters: the reference asks the embeddings service (cu:813)
                continue
            medoid.append({"values": grp, "method": method, "pvf": 0.8, "value": v, "conf": c})
    meta = {"generator": "oracle/gen_golden.py", "reference": "retab-dev/k-LLMs @ 089dba9 behind 3 import stubs",
            "entry": "consensus_values(values, ConsensusSettings(), raising_embeddings, client=None)"}
    for name, payload in (("known_answers", known), ("client_order", client), ("random_cases", rnd), ("alignment", alignment),
                          ("medoid", medoid)):
        with open(os.path.join(GOLDEN_DIR, name + ".json"), "w") as f:
            json.dump({"meta": meta, "cases": payload}, f, separators=(",", ":"))
        print(f"wrote {name}.json: {len(payload)} cases")

    # oracle restatement vs the goldens just produced
    bad = 0
    for case in known + rnd:
        got = O.consensus(case["values"], embed=raising_embeddings)
        if not (_same(got[0], case["value"]) and _same(got[1], case["conf"])):
            bad += 1
            print("MISMATCH", case["values"], "ref=", (case["value"], case["conf"]), "oracle=", got)
    for case in client:
        if any(isinstance(x, list) for d in case["values"] if isinstance(d, dict) for x in d.values()):
            continue  # the oracle restates only the dict part of the pre-pass (lists: tests/golden pins the product directly)
        got = O.client_order(case["values"], embed=raising_embeddings)
        if not (_same(got[0], case["value"]) and _same(got[1], case["conf"])):
            bad += 1
            print("MISMATCH(client)", case["values"], (case["value"], case["conf"]), got)
    print("oracle vs golden mismatches:", bad)

    if args.fuzz:
        bad = 0
        for vals in random_cases(777, args.fuzz):
            ref = ref_consensus_values(vals)
            got = O.consensus(vals, embed=raising_embeddings)
            if not (_same(got[0], ref[0]) and _same(got[1], ref[1])):
                bad += 1
                if bad < 10:
                    print("FUZZ MISMATCH", vals, ref, got)
        print(f"fuzz: {args.fuzz} cases, {bad} mismatches")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
