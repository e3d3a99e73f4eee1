"""k_llms_b200 — B200-native n-way consensus consolidator behind the k_llms client surface.

    from k_llms_b200 import KLLMs, AsyncKLLMs     # same surface as `from k_llms import KLLMs, AsyncKLLMs`

The client / resources / consolidation layers are host Python; the consolidation hot path (str/bool votes and
numeric cluster consensus) runs as sm_100a CUDA behind the C ABI in include/kllms_b200.h.
"""


def __getattr__(name):  # lazy: importing the package must not require openai / torch
    if name in ("KLLMs", "AsyncKLLMs"):
        from . import client
        return getattr(client, name)
    raise AttributeError(name)


__all__ = ["KLLMs", "AsyncKLLMs"]
