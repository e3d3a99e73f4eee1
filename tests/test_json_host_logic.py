"""Host logic of the native JSON path (H1: parse, key merge, planning, encoding, decoding, json.dumps formatting) on a machine
without a GPU: kc_json_plan -> the C oracle in the kernels' place -> kc_json_emit, against the reference's client order
(the object-level oracle).  The same generators as tests/test_gpu_json.py, where the real kernels sit in the middle."""
import random

from tests.helpers import consolidate_json_with_oracle
from tests.test_gpu_json import _expected, _random_nested_record, _random_record


def test_two_phase_native_json_matches_client_order_cpu():
    rng = random.Random(11)
    by_n = {}
    for _ in range(500):
        n = rng.choice([2, 3, 5, 8, 16])
        by_n.setdefault(n, []).append(_random_record(rng, n))
    for _ in range(300):
        n = rng.choice([2, 3, 5, 8])
        by_n.setdefault(n, []).append(_random_nested_record(rng, n))
    native = 0
    for _n, recs in by_n.items():
        for texts, got in zip(recs, consolidate_json_with_oracle(recs)):
            if got is None:
                continue
            native += 1
            assert got == _expected(texts), texts
    assert native > 600


def test_two_phase_native_json_list_records_cpu():
    """Records with list fields through the native path (H2 alignment on the parsed tree + element-wise merge), the oracle in
    the kernels' place: the reference's client-order goldens and random list records."""
    import json

    from oracle.gen_golden import random_list_records
    from tests.helpers import load_golden
    from tests.test_gpu_json import _expected_with_lists
    from k_llms_b200.utils.consolidation import _format_consensus_content
    by_n, native = {}, 0
    for case in load_golden("client_order"):
        if len(case["values"]) >= 2:
            by_n.setdefault(len(case["values"]), []).append(([json.dumps(v) for v in case["values"]], case))
    for _n, items in by_n.items():
        for (texts, case), got in zip(items, consolidate_json_with_oracle([t for t, _ in items])):
            assert got is not None, texts
            native += 1
            assert got == (_format_consensus_content(case["value"]), json.dumps(case["conf"])), (texts, got)
    assert native > 100
    recs = [[json.dumps(v) for v in r] for r in random_list_records(78, 200)]
    by_n = {}
    for r in recs:
        by_n.setdefault(len(r), []).append(r)
    for _n, rs in by_n.items():
        for texts, got in zip(rs, consolidate_json_with_oracle(rs)):
            assert got is not None and got == _expected_with_lists(texts), texts
