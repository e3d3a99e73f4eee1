"""Seeded synthetic candidate sets in the columnar device encoding.

Schema S32: f00-f15 string-enum (8-token vocabulary; the code of a cell is the id of its SANITISED token, so
case/punctuation variants of a token share a code), f16-f23 bool (code 0 = False, 1 = True, None votes as
False), f24-f29 int in [1, 1e6), f30-f31 float U(1, 1e4).  Per (record, field): draw a truth; each of the
n candidates copies it w.p. p_agree, else draws independently; then becomes None w.p. p_none.
"""
from __future__ import annotations

import numpy as np

F64_NONE = np.array([0x7FF8C0DE00000000], dtype=np.uint64).view(np.float64)[0]
F64_ABSENT = np.array([0x7FF8C0DF00000000], dtype=np.uint64).view(np.float64)[0]

S32_VOTE_FIELDS = 24  # 16 string-enum + 8 bool
S32_NUM_FIELDS = 8    # 6 int + 2 float
S32_NONE_CODE = np.array([-1] * 16 + [0] * 8, dtype=np.int32)


def s32_bytes_per_record(n: int) -> int:
    """Algorithmic bytes per record: every cell read once + winning value + 4-byte result word."""
    return S32_VOTE_FIELDS * (4 * n + 8) + S32_NUM_FIELDS * (8 * n + 12)


def s32_numpy(n_records: int, n: int, seed: int, p_agree: float = 0.8, p_none: float = 0.05):
    """(codes int32 [N,24,n], none_code int32 [24], vals float64 [N,8,n]) with numpy PCG64(seed)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    N = n_records
    vocab = np.array([8] * 16 + [2] * 8)
    truth = (rng.random((N, 24, 1)) * vocab[None, :, None]).astype(np.int32)
    draw = (rng.random((N, 24, n)) * vocab[None, :, None]).astype(np.int32)
    codes = np.where(rng.random((N, 24, n)) < p_agree, truth, draw).astype(np.int32)
    codes[rng.random((N, 24, n)) < p_none] = -1

    def num_draw(shape):
        ints = np.floor(1 + rng.random(shape[:1] + (6,) + shape[2:]) * (1e6 - 1))
        flts = 1.0 + rng.random(shape[:1] + (2,) + shape[2:]) * (1e4 - 1.0)
        return np.concatenate([ints, flts], axis=1)

    t = num_draw((N, 8, 1))
    d = num_draw((N, 8, n))
    vals = np.where(rng.random((N, 8, n)) < p_agree, t, d)
    vals[rng.random((N, 8, n)) < p_none] = F64_NONE
    return np.ascontiguousarray(codes), S32_NONE_CODE.copy(), np.ascontiguousarray(vals)


def s32_torch(n_records: int, n: int, seed: int, device, p_agree: float = 0.8, p_none: float = 0.05):
    """Same distribution generated on the GPU with torch's generator (fast path for bench.py)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    N = n_records
    vocab = torch.tensor([8] * 16 + [2] * 8, device=device, dtype=torch.float32).view(1, 24, 1)
    codes = torch.empty((N, 24, n), dtype=torch.int32, device=device)
    vals = torch.empty((N, 8, n), dtype=torch.float64, device=device)
    step = 1 << 16
    none_val = torch.tensor([0x7FF8C0DE00000000], dtype=torch.int64, device=device).view(torch.float64)
    for r0 in range(0, N, step):
        r1 = min(N, r0 + step)
        m = r1 - r0
        truth = (torch.rand((m, 24, 1), generator=g, device=device) * vocab).to(torch.int32)
        draw = (torch.rand((m, 24, n), generator=g, device=device) * vocab).to(torch.int32)
        c = torch.where(torch.rand((m, 24, n), generator=g, device=device) < p_agree, truth, draw)
        c = torch.where(torch.rand((m, 24, n), generator=g, device=device) < p_none, torch.full_like(c, -1), c)
        codes[r0:r1] = c

        def num_draw(k):
            ints = torch.floor(1 + torch.rand((m, 6, k), generator=g, device=device, dtype=torch.float64) * (1e6 - 1))
            flts = 1.0 + torch.rand((m, 2, k), generator=g, device=device, dtype=torch.float64) * (1e4 - 1.0)
            return torch.cat([ints, flts], dim=1)

        t, d = num_draw(1), num_draw(n)
        v = torch.where(torch.rand((m, 8, n), generator=g, device=device) < p_agree, t.expand(-1, -1, n), d)
        v = torch.where(torch.rand((m, 8, n), generator=g, device=device) < p_none, none_val.expand_as(v), v)
        vals[r0:r1] = v
    none_code = torch.tensor(S32_NONE_CODE, device=device)
    return codes, none_code, vals


def phrase_groups_numpy(n_groups: int, k: int, seed: int, min_len: int = 16, max_len: int = 48, p_agree: float = 0.5):
    """Synthetic input of K4 (similarity medoid): n_groups groups of k normalised strings ([a-z0-9], what
    normalize_string (consensus_utils.py:660-673) leaves of a multi-word field).  Every group has one base string of
    min_len..max_len characters; a candidate repeats it with probability p_agree, otherwise carries 1..4 substituted
    characters and may lose up to 4 trailing ones.  Returns (chars uint8, str_off int32 [G*k+1], grp_off int32 [G+1])."""
    rng = np.random.default_rng(seed)
    alphabet = np.frombuffer(b"abcdefghijklmnopqrstuvwxyz0123456789", dtype=np.uint8)
    base_len = rng.integers(min_len, max_len + 1, n_groups)
    base = alphabet[rng.integers(0, 36, (n_groups, max_len))]
    cand = np.repeat(base[:, None, :], k, axis=1)  # [G, k, L]
    mutate = rng.random((n_groups, k)) >= p_agree
    n_sub = rng.integers(1, 5, (n_groups, k)) * mutate
    for s in range(4):
        pos = (rng.random((n_groups, k)) * base_len[:, None]).astype(np.int64)
        new = alphabet[rng.integers(0, 36, (n_groups, k))]
        g, c = np.nonzero(n_sub > s)
        cand[g, c, pos[g, c]] = new[g, c]
    lens = base_len[:, None] - rng.integers(0, 5, (n_groups, k)) * (mutate & (rng.random((n_groups, k)) < 0.3))
    lens = np.maximum(lens, 1).astype(np.int64)
    keep = np.arange(max_len)[None, None, :] < lens[:, :, None]
    chars = np.ascontiguousarray(cand[keep])
    str_off = np.zeros(n_groups * k + 1, dtype=np.int32)
    np.cumsum(lens.reshape(-1), out=str_off[1:])
    grp_off = (np.arange(n_groups + 1, dtype=np.int64) * k).astype(np.int32)
    return chars, str_off, grp_off
