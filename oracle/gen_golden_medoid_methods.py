"""Golden vectors of the similarity medoid under the reference's OTHER string similarity methods (TEST INFRASTRUCTURE; run in the
build container):  python -m oracle.gen_golden_medoid_methods  ->  tests/golden/medoid_methods.json

consensus_values(group, ConsensusSettings(string_similarity_method=m), ...) of the running reference for m in {jaccard, hamming}
on the phrase groups of oracle/gen_golden.py (multi-word strings: the medoid fallback, consensus_utils.py:1221-1237 with
jaccard_similarity :720-742 and hamming_similarity :676-717)."""
from __future__ import annotations

import json
import logging
import os

from oracle.gen_golden import GOLDEN_DIR, phrase_groups
from oracle.ref_loader import load_reference, raising_embeddings


def main() -> None:
    logging.disable(logging.CRITICAL)
    cu = load_reference()
    cases = []
    for grp in phrase_groups(777, 220):
        for method in ("jaccard", "hamming"):
            settings = cu.ConsensusSettings(string_similarity_method=method)
            v, c = cu.consensus_values(grp, settings, raising_embeddings, None, 0.9)
            cases.append({"values": grp, "method": method, "pvf": 0.9, "value": v, "conf": c})
    meta = {"generator": "oracle/gen_golden_medoid_methods.py", "reference": "retab-dev/k-LLMs @ 089dba9 behind 3 import stubs",
            "entry": "consensus_values(values, ConsensusSettings(string_similarity_method=m), raising_embeddings, None, 0.9)"}
    with open(os.path.join(GOLDEN_DIR, "medoid_methods.json"), "w") as f:
        json.dump({"meta": meta, "cases": cases}, f, separators=(",", ":"))
    print(f"wrote medoid_methods.json: {len(cases)} cases")


if __name__ == "__main__":
    main()
