"""Load the UNMODIFIED reference (`/root/reference/k_llms`) behind three import stubs.

TEST INFRASTRUCTURE ONLY.  This module is used in the build container to (a) validate the
oracle restatement and (b) generate the golden vectors committed under `tests/golden/`
(see `oracle/gen_golden.py`).  `/root/reference` does not exist on the GPU box, so nothing
in the `-m gpu` tests, `smoke()` or `bench.py` imports this file.

The reference cannot be imported as shipped: three third-party modules
that are absent from this image are imported at module scope of
`k_llms/utils/consensus_utils.py`:

* `Levenshtein.distance`   (consensus_utils.py:15)  -> textbook unit-cost edit distance
* `retab.types.documents.extract.RetabParsedChatCompletion` (consensus_utils.py:21)
                                                     -> empty class (annotation only, dead code)
* `unidecode.unidecode`    (consensus_utils.py:22)  -> identity, ASSERTING ascii input
  (the real library is the identity on ASCII; non-ASCII parity is unpinned)

Nothing else is altered: the reference's own functions run as they are.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("KLLMS_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "k_llms", "utils", "consensus_utils.py"))


def _edit_distance(a: str, b: str) -> int:
    """Unit-cost Levenshtein distance (what python-Levenshtein's `distance` returns)."""
    if a == b:
        return 0
    if len(a) < len(b):
        a, b = b, a
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


def _ascii_unidecode(s: str) -> str:
    assert s.isascii(), "unidecode stub: non-ASCII input — parity is unpinned for it (SURVEY §8c)"
    return s


def _install_stubs() -> None:
    if "Levenshtein" not in sys.modules:
        m = types.ModuleType("Levenshtein")
        m.distance = _edit_distance
        sys.modules["Levenshtein"] = m
    if "unidecode" not in sys.modules:
        m = types.ModuleType("unidecode")
        m.unidecode = _ascii_unidecode
        sys.modules["unidecode"] = m
    if "retab" not in sys.modules:
        names = ["retab", "retab.types", "retab.types.documents", "retab.types.documents.extract"]
        mods = {n: types.ModuleType(n) for n in names}
        for n in names[:-1]:
            mods[n].__path__ = []  # mark as packages
        mods[names[-1]].RetabParsedChatCompletion = type("RetabParsedChatCompletion", (), {})
        sys.modules.update(mods)


_cached = None


def load_reference():
    """Return the reference's `k_llms.utils.consensus_utils` module (unmodified code)."""
    global _cached
    if _cached is not None:
        return _cached
    if not reference_available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT} (expected on the GPU box)")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # `k_llms` in /root/reference, NOT this repo's `k_llms_b200`.
    _cached = importlib.import_module("k_llms.utils.consensus_utils")
    assert _cached.__file__.startswith(REFERENCE_ROOT), _cached.__file__
    return _cached


def raising_embeddings(texts):
    """Embeddings callable that makes any accidental network use loud (reference then falls
    back to Levenshtein, consensus_utils.py:816-820)."""
    raise RuntimeError("network embeddings are not available in the oracle")


def ref_consensus_values(values, settings=None, parent_valid_frac: float = 1.0):
    cu = load_reference()
    settings = settings or cu.ConsensusSettings()
    return cu.consensus_values(values, settings, raising_embeddings, client=None, parent_valid_frac=parent_valid_frac)


def ref_client_order(values, settings=None):
    """align (consolidation.py:340-347) then consensus (consolidation.py:349)."""
    cu = load_reference()
    settings = settings or cu.ConsensusSettings()
    aligned, _ = cu.recursive_list_alignments(values, settings.string_similarity_method, raising_embeddings, None, settings.min_support_ratio)
    aligned = [(d if isinstance(d, dict) else {}) for d in aligned]
    return cu.consensus_values(aligned, settings, raising_embeddings, client=None)
