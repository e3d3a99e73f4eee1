"""torch.ops.kllms_b200.* (k_llms_b200/torch_ops.py) call the same kernels as the C ABI bindings."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_registered_ops_match_the_bindings():
    import k_llms_b200.torch_ops  # noqa: F401
    from k_llms_b200 import _native as K
    from k_llms_b200 import synth
    codes, none_code, vals = synth.s32_numpy(2000, 16, 5)
    c = torch.from_numpy(codes.reshape(-1, 16)).cuda()
    nc = torch.from_numpy(none_code).cuda()
    v = torch.from_numpy(vals.reshape(-1, 16)).cuda()
    w0, m0 = K.vote(c, nc)
    w1, m1 = torch.ops.kllms_b200.vote(c, nc)
    assert torch.equal(w0, w1) and torch.equal(m0, m1)
    x0, n0 = K.numeric(v)
    x1, n1 = torch.ops.kllms_b200.numeric(v, 0.03, 1e-6)
    assert torch.equal(x0.view(torch.int64), x1.view(torch.int64)) and torch.equal(n0, n1)
    off = torch.from_numpy(np.arange(0, 33 * 50, 33, dtype=np.int64)).cuda()
    lp = -torch.rand(int(off[-1]), device="cuda")
    assert torch.equal(K.logprob_sum(lp, off), torch.ops.kllms_b200.logprob_sum(lp, off))
    with pytest.raises(NotImplementedError):
        torch.ops.kllms_b200.vote(c.cpu(), None)
