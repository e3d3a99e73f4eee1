"""CPU: the columnar C oracle — against numpy (the library the reference calls) and, through the product's host
prologue/epilogue, against the reference-generated golden vectors.  No GPU, no CUDA library involved."""
import numpy as np
import pytest

from oracle import columnar as OC
from tests.helpers import load_golden, oracle_run, raising_embeddings, same


def test_numpy_reductions_restated_exactly():
    L = OC.lib()
    rng = np.random.default_rng(0)
    for t in range(20000):
        n = int(rng.integers(1, 65))
        a = np.sort([rng.uniform(-1e4, 1e4, n), rng.uniform(1, 1e6, n).round(2), rng.standard_normal(n) * 10.0 ** rng.integers(-8, 8)][t % 3])
        lst = a.tolist()
        assert L.ko_np_mean(a.ctypes.data, n) == float(np.mean(lst))
        assert L.ko_np_std(a.ctypes.data, n) == float(np.std(lst))
        assert L.ko_np_median_sorted(a.ctypes.data, n) == float(np.median(lst))


def test_vote_oracle_basics():
    codes = np.array([[3, 7, 7, 3], [-1, -1, -1, -1], [-2, 5, -1, 5], [-1, 1, 1, -1]], dtype=np.int32)
    win, meta = OC.vote(codes, None)
    f = OC.meta_fields(meta)
    assert win.tolist() == [3, -1, 5, 1]
    assert f["idx"].tolist()[0] == 0 and f["support"].tolist() == [2, 0, 2, 2]
    assert f["present"].tolist() == [4, 4, 3, 4] and f["nn"].tolist() == [4, 0, 2, 2]
    assert (f["flags"][0] & 4) and not (f["flags"][2] & 4)  # tie flag
    win, meta = OC.vote(codes, np.array([0], dtype=np.int32))  # None votes as code 0
    assert win.tolist() == [3, 0, 5, 0] and OC.meta_fields(meta)["support"].tolist() == [2, 4, 2, 2]


@pytest.mark.parametrize("name", ["known_answers", "random_cases"])
def test_host_plan_and_epilogue_with_oracle_kernels(name):
    """Host logic of the product (planner, encoder, decoder) with the C oracle standing in for K1/K2."""
    from k_llms_b200.utils.consensus_utils import ConsensusSettings, _plan_for
    for case in load_golden(name):
        values = case["values"]
        if len(values) > 64:
            continue
        plan = _plan_for(len(values), ConsensusSettings())
        root = plan.add(values, 1.0, raising_embeddings)
        got = plan.materialise(root, oracle_run(plan))
        assert same(got[0], case["value"]) and same(got[1], case["conf"]), (values, got, case["value"], case["conf"])


def test_logprob_sum_order_is_the_documented_one():
    rng = np.random.default_rng(1)
    lens = rng.integers(0, 130, 200)
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    lp = (-rng.exponential(1.0, offsets[-1])).astype(np.float32)
    got = OC.logprob_sum(lp, offsets)
    for s in range(len(lens)):
        seg = lp[offsets[s]:offsets[s + 1]]
        lanes = [np.float32(0)] * 32
        for l in range(32):
            acc = np.float32(0)
            for x in seg[l::32]:
                acc = np.float32(acc + x)
            lanes[l] = acc
        stride = 16
        while stride:
            lanes = [np.float32(lanes[l] + lanes[l ^ stride]) for l in range(32)]
            stride //= 2
        assert got[s] == lanes[0]


def test_medoid_oracle_matches_host_medoid():
    """ko_medoid_str (the K4 checker) == similarity.medoid (cu:1221-1237 restated on numpy) on random phrase groups."""
    from k_llms_b200.columnar import _normalize
    from k_llms_b200.utils import similarity
    from tests.helpers import random_string_groups

    rng = np.random.default_rng(5)
    groups = random_string_groups(rng, 400, max_k=20)
    idx, avg = OC.medoid([[_normalize(s) for s in g] for g in groups])
    for g, i, a in zip(groups, idx, avg):
        value, conf = similarity.medoid(g, "levenshtein", None, 0.7)
        assert value == g[int(i)] and g.index(value) == int(i) or g[int(i)] == value
        assert conf == round(0.7 * float(a), 5)


def test_medoid_goldens_through_host_plan_with_oracle_k4():
    """Planner routes phrase groups to K4 (C oracle standing in); results == the reference's (tests/golden/medoid.json)."""
    from k_llms_b200.utils.consensus_utils import ConsensusSettings, _plan_for
    on_device = 0
    for case in load_golden("medoid"):
        st = ConsensusSettings(string_similarity_method=case["method"])
        plan = _plan_for(len(case["values"]), st)
        root = plan.add(case["values"], case["pvf"], raising_embeddings)
        on_device += len(plan.medoid_groups)
        got = plan.materialise(root, oracle_run(plan))
        assert same(got[0], case["value"]) and same(got[1], case["conf"]), (case, got)
    assert on_device > 300  # the fixture really exercises the K4 route, not the host fallback
