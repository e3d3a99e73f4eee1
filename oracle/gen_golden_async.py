"""Golden vectors of the reference's ASYNC dispatcher (TEST INFRASTRUCTURE; run in the build container, where /root/reference
exists):  python -m oracle.gen_golden_async  ->  tests/golden/async_cases.json

The reference's async primitive has no numeric clustering (consensus_utils.py:1638-1688): numeric fields that are not
unanimous take the similarity medoid, so AsyncKLLMs results differ from KLLMs results.  The product reproduces that; these
vectors pin it.  Inputs: the SURVEY §8c known-answer list, random scalar / nested cases and client-order cases (the async
alignment pre-pass first) from oracle/gen_golden.py's generators."""
from __future__ import annotations

import asyncio
import json
import logging
import os

from oracle.gen_golden import CLIENT_ORDER_INPUTS, GOLDEN_DIR, KNOWN_INPUTS, random_cases, random_list_records
from oracle.ref_loader import load_reference


async def _raising(texts):
    raise RuntimeError("network embeddings are not available in the oracle")


def main() -> None:
    logging.disable(logging.CRITICAL)
    cu = load_reference()
    settings = cu.ConsensusSettings()

    async def consensus(vals):
        return await cu.async_consensus_values(vals, settings, _raising, client=None)

    async def client_order(vals):
        aligned, _ = await cu.async_recursive_list_alignments(vals, settings.string_similarity_method, _raising, None, settings.min_support_ratio)
        aligned = [(d if isinstance(d, dict) else {}) for d in aligned]
        return await cu.async_consensus_values(aligned, settings, _raising, client=None)

    cases = []
    for vals in list(KNOWN_INPUTS) + random_cases(31337, 300):
        try:
            v, c = asyncio.run(consensus(vals))
        except Exception as exc:  # e.g. the medoid of values generic_similarity cannot compare
            cases.append({"kind": "consensus", "values": vals, "raises": type(exc).__name__})
            continue
        cases.append({"kind": "consensus", "values": vals, "value": v, "conf": c})
    for vals in list(CLIENT_ORDER_INPUTS) + random_list_records(5150, 60):
        try:
            v, c = asyncio.run(client_order(vals))
        except Exception as exc:
            cases.append({"kind": "client_order", "values": vals, "raises": type(exc).__name__})
            continue
        cases.append({"kind": "client_order", "values": vals, "value": v, "conf": c})
    meta = {"generator": "oracle/gen_golden_async.py", "reference": "retab-dev/k-LLMs @ 089dba9 behind 3 import stubs",
            "entry": "async_consensus_values / async_recursive_list_alignments with a raising embeddings coroutine"}
    with open(os.path.join(GOLDEN_DIR, "async_cases.json"), "w") as f:
        json.dump({"meta": meta, "cases": cases}, f, separators=(",", ":"))
    n_raise = sum(1 for c in cases if "raises" in c)
    print(f"wrote async_cases.json: {len(cases)} cases ({n_raise} raise in the reference)")


if __name__ == "__main__":
    main()
