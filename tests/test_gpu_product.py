"""GPU parity, product level: the drop-in Python surface (consensus_values, batch entry, consolidation) against the
reference-generated golden vectors and the object-level oracle.  Every call goes host prologue -> C ABI -> CUDA."""
import json
import logging
import random

import pytest

from oracle import consensus_py as O
from tests.helpers import load_golden, raising_embeddings, same

pytestmark = pytest.mark.gpu


def _settings():
    from k_llms_b200.utils.consensus_utils import ConsensusSettings
    return ConsensusSettings()


@pytest.mark.parametrize("name", ["known_answers", "random_cases"])
def test_consensus_values_matches_reference_goldens(name):
    from k_llms_b200.utils.consensus_utils import consensus_values
    for case in load_golden(name):
        got = consensus_values(case["values"], _settings(), raising_embeddings, None)
        assert same(got[0], case["value"]) and same(got[1], case["conf"]), (case["values"], got, case["value"], case["conf"])


def test_medoid_goldens_on_gpu():
    """Multi-word string fields: K4 (kc_medoid_str) behind consensus_values == the reference (tests/golden/medoid.json)."""
    from k_llms_b200.utils.consensus_utils import ConsensusSettings, consensus_values, consensus_values_batch
    cases = load_golden("medoid")
    for case in cases:
        st = ConsensusSettings(string_similarity_method=case["method"])
        got = consensus_values(case["values"], st, raising_embeddings, None, case["pvf"])
        assert same(got[0], case["value"]) and same(got[1], case["conf"]), (case, got)
    lev = [c for c in cases if c["method"] == "levenshtein"]
    outs = consensus_values_batch([[{"k": v} for v in c["values"]] for c in lev],
                                  ConsensusSettings(string_similarity_method="levenshtein"), raising_embeddings)  # one K4 launch
    assert [o[0]["k"] for o in outs] == [c["value"] for c in lev]


def test_medoid_other_similarity_methods_on_gpu(monkeypatch):
    """string_similarity_method 'jaccard' / 'hamming': K4 computes those pair similarities too (kc_medoid_str_method) — the
    reference's outputs (tests/golden/medoid_methods.json), with a guard that the groups really went to the device."""
    from k_llms_b200 import columnar
    from k_llms_b200.utils.consensus_utils import ConsensusSettings, consensus_values
    on_device = []
    real = columnar.Plan._medoid_on_device
    monkeypatch.setattr(columnar.Plan, "_medoid_on_device", lambda self, live: on_device.append(real(self, live)) or on_device[-1])
    cases = load_golden("medoid_methods")
    assert {c["method"] for c in cases} == {"jaccard", "hamming"}
    for case in cases:
        st = ConsensusSettings(string_similarity_method=case["method"])
        got = consensus_values(case["values"], st, raising_embeddings, None, case["pvf"])
        assert same(got[0], case["value"]) and same(got[1], case["conf"]), (case, got)
    assert sum(on_device) > 0.9 * len(on_device) > 100


def test_batch_entry_equals_per_record_and_goldens():
    from k_llms_b200.utils.consensus_utils import consensus_values_batch
    cases = [c for c in load_golden("random_cases") + load_golden("known_answers") if len(c["values"]) <= 64]
    outs = consensus_values_batch([c["values"] for c in cases], _settings(), raising_embeddings)
    assert len(outs) == len(cases)
    for case, got in zip(cases, outs):
        assert same(got[0], case["value"]) and same(got[1], case["conf"]), case["values"]


def test_client_order_goldens_through_consolidation_helpers():
    """align (host) + consensus (GPU) in the order consolidation.py runs them, incl. reordered lists."""
    from k_llms_b200.utils.consolidation import _consensus_sync
    logging.disable(logging.CRITICAL)
    try:
        for case in load_golden("client_order"):
            got = _consensus_sync(json.loads(json.dumps(case["values"])), _settings(), raising_embeddings, None)
            assert same(got[0], case["value"]) and same(got[1], case["conf"]), (case["values"], got, case["value"], case["conf"])
    finally:
        logging.disable(logging.NOTSET)


def test_other_settings_against_oracle():
    from k_llms_b200.utils.consensus_utils import ConsensusSettings, consensus_values
    from oracle.gen_golden import random_cases
    for vals in random_cases(31337, 300):
        for kw in ({"allow_none_as_candidate": True}, {"rel_eps": 0.1, "abs_eps": 0.5}, {"string_similarity_method": "jaccard"}):
            st = ConsensusSettings(**kw)
            os_ = O.OracleSettings(allow_none_as_candidate=st.allow_none_as_candidate, rel_eps=st.rel_eps, abs_eps=st.abs_eps,
                                   string_similarity_method=st.string_similarity_method)
            got = consensus_values(vals, st, raising_embeddings, None)
            exp = O.consensus(vals, os_, embed=raising_embeddings)
            assert same(got[0], exp[0]) and same(got[1], exp[1]), (kw, vals, got, exp)


def test_nested_depth4_lists_n8_config3():
    """BASELINE config 3: nested JSON (depth 4) with list fields, n=8, element-wise merge."""
    from k_llms_b200.utils.consensus_utils import consensus_values_batch
    from oracle.gen_golden import _record_candidates
    rng = random.Random(8)
    records = [_record_candidates(rng, 8, depth=3) for _ in range(400)]
    outs = consensus_values_batch(records, _settings(), raising_embeddings)
    for cands, got in zip(records, outs):
        exp = O.consensus(cands, embed=raising_embeddings)
        assert same(got[0], exp[0]) and same(got[1], exp[1])


def test_consolidate_chat_completions_end_to_end():
    """The public consolidation entry with synthetic ChatCompletion objects (no network)."""
    from openai.types.chat import ChatCompletion
    from k_llms_b200.utils.consolidation import consolidate_chat_completions
    payloads = [{"name": "John", "age": 30, "active": True, "city": "Paris"},
                {"name": "John", "age": 30, "active": True, "city": "paris"},
                {"name": "Jon", "age": 31, "active": False, "city": "Paris"}]
    completion = ChatCompletion.model_validate({
        "id": "x", "object": "chat.completion", "created": 0, "model": "m",
        "choices": [{"index": i, "finish_reason": "stop", "message": {"role": "assistant", "content": json.dumps(p)}}
                    for i, p in enumerate(payloads)]})
    out = consolidate_chat_completions(completion, raising_embeddings, client=None)
    assert len(out.choices) == 4 and out.choices[0].index == 0
    assert json.loads(out.choices[0].message.content) == {"active": True, "age": 30.0, "city": "Paris", "name": "John"}
    assert out.likelihoods == {"active": 0.66667, "age": 0.66667, "city": 1.0, "name": 0.66667}
    assert [c.index for c in out.choices[1:]] == [1, 2, 3]
    # free text goes through the {"text": ...} wrapper and back (consolidation.py:25-60)
    texts = ["Yes", "yes", "No"]
    completion = ChatCompletion.model_validate({
        "id": "x", "object": "chat.completion", "created": 0, "model": "m",
        "choices": [{"index": i, "finish_reason": "stop", "message": {"role": "assistant", "content": t}} for i, t in enumerate(texts)]})
    out = consolidate_chat_completions(completion, raising_embeddings, client=None)
    assert out.choices[0].message.content == "Yes" and out.likelihoods == {"text": 0.66667}


def test_host_buffer_entry_matches_device_entry():
    import numpy as np
    import torch
    from k_llms_b200 import _native as K
    from k_llms_b200 import synth
    codes, none_code, vals = synth.s32_numpy(70_001, 16, 5)
    res = K.consensus_host(codes, none_code, vals)
    win, meta = K.vote(torch.from_numpy(codes.reshape(-1, 16)).cuda(), torch.from_numpy(none_code).cuda())
    value, nmeta = K.numeric(torch.from_numpy(vals.reshape(-1, 16)).cuda())
    assert np.array_equal(res["win_code"].reshape(-1), win.cpu().numpy())
    assert np.array_equal(res["vote_meta"].reshape(-1), meta.cpu().numpy().view(np.uint32))
    assert np.array_equal(res["value"].reshape(-1).view(np.uint64), value.cpu().numpy().view(np.uint64))
    assert np.array_equal(res["num_meta"].reshape(-1), nmeta.cpu().numpy().view(np.uint32))
    assert res["device_ms"] > 0
