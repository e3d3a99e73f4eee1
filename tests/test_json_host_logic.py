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
