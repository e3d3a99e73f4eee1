"""CPU: the object-level oracle (oracle/consensus_py.py) against the reference-generated golden vectors."""
import pytest

from oracle import consensus_py as O
from tests.helpers import load_golden, raising_embeddings, same


@pytest.mark.parametrize("name", ["known_answers", "random_cases"])
def test_oracle_matches_reference_outputs(name):
    cases = load_golden(name)
    assert len(cases) > 50
    for case in cases:
        got = O.consensus(case["values"], embed=raising_embeddings)
        assert same(got[0], case["value"]) and same(got[1], case["conf"]), (case["values"], got, case["value"], case["conf"])


def test_oracle_client_order_dict_part():
    n = 0
    for case in load_golden("client_order"):
        if any(isinstance(x, list) for d in case["values"] if isinstance(d, dict) for x in d.values()):
            continue  # list alignment is pinned on the product directly (tests/test_alignment_golden.py)
        got = O.client_order(case["values"], embed=raising_embeddings)
        assert same(got[0], case["value"]) and same(got[1], case["conf"])
        n += 1
    assert n >= 2


def test_survey_table_spot_checks():
    """A few rows of SURVEY.md §8c typed in by hand, as a guard against a stale golden file."""
    c = lambda v: O.consensus(v, embed=raising_embeddings)  # noqa: E731
    assert c(["b", "a", "a", "b"]) == ("b", 0.5)
    assert c(["Paris!", "paris", "PARIS ", "Lyon"]) == ("Paris!", 0.75)
    assert c([True, None, None]) == (False, 0.66667)
    assert c([5, 5, 50, 50, 7, 7]) == (50.0, 0.33333)
    assert c([100, 102.9, 105.8]) == (102.89999999999999, 1.0)
    assert c([1, None, None]) == (1, 0.3333333333333333)
    assert c([30, 30, 31, None]) == (30.0, 0.66667)
    assert c([{"a": "x"}, {"a": "x"}, None, None]) == ({"a": "x"}, {"a": 0.5})
    assert c([]) == (None, 1.0) and c([None, None]) == (None, 0.0)


def test_host_similarity_medoid_matches_reference_for_jaccard_and_hamming():
    """k_llms_b200.utils.similarity (the host side of the medoid, used for what K4 does not take) against the reference's
    outputs under string_similarity_method 'jaccard' / 'hamming' (tests/golden/medoid_methods.json)."""
    from k_llms_b200.utils import similarity as S
    from tests.helpers import load_golden
    cases = load_golden("medoid_methods")
    assert len(cases) > 400
    for c in cases:
        live = [v for v in c["values"] if v is not None]
        if len(live) < 2:
            continue
        got = S.medoid(live, c["method"], None, c["pvf"] * len(live) / len(c["values"]))
        assert got[0] == c["value"] and got[1] == c["conf"], (c, got)
