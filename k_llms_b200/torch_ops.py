"""torch.ops.kllms_b200.{vote, numeric, logprob_sum} — the three kernels registered as PyTorch custom operators (SURVEY.md
§8b: "same three registered as torch.ops.kllms_b200.*").  Thin: each op forwards to the C ABI through k_llms_b200._native on
the tensors' device and current stream; there is no CPU implementation (a CPU tensor raises NotImplementedError from the
dispatcher).  Import this module to register them:

    import k_llms_b200.torch_ops
    win, meta = torch.ops.kllms_b200.vote(codes, none_code)        # int32 [G, n], int32 [F] or None -> int32 [G], int32 [G]
    value, meta = torch.ops.kllms_b200.numeric(vals, 0.03, 1e-6)   # float64 [G, n] -> float64 [G], int32 [G]
    sums = torch.ops.kllms_b200.logprob_sum(logprobs, offsets)     # float32 [T], int64 [S+1] -> float32 [S]
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _native

_lib = torch.library.Library("kllms_b200", "DEF")
_lib.define("vote(Tensor codes, Tensor? none_code=None) -> (Tensor, Tensor)")
_lib.define("numeric(Tensor vals, float rel_eps=0.03, float abs_eps=1e-6) -> (Tensor, Tensor)")
_lib.define("logprob_sum(Tensor logprobs, Tensor offsets) -> Tensor")


def _vote(codes: torch.Tensor, none_code: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    return _native.vote(codes, none_code)


def _numeric(vals: torch.Tensor, rel_eps: float = 0.03, abs_eps: float = 1e-6) -> Tuple[torch.Tensor, torch.Tensor]:
    return _native.numeric(vals, rel_eps, abs_eps)


def _logprob_sum(logprobs: torch.Tensor, offsets: torch.Tensor) -> torch.Tensor:
    return _native.logprob_sum(logprobs, offsets)


_lib.impl("vote", _vote, "CUDA")
_lib.impl("numeric", _numeric, "CUDA")
_lib.impl("logprob_sum", _logprob_sum, "CUDA")


def _meta_pair(first_dtype):
    def fn(x, *args, **kwargs):
        g = x.shape[0]
        return x.new_empty((g,), dtype=first_dtype), x.new_empty((g,), dtype=torch.int32)
    return fn


_lib.impl("vote", _meta_pair(torch.int32), "Meta")
_lib.impl("numeric", _meta_pair(torch.float64), "Meta")
_lib.impl("logprob_sum", lambda lp, off: lp.new_empty((off.shape[0] - 1,), dtype=torch.float32), "Meta")
