"""ctypes front-end of the columnar C oracle (oracle/consensus_oracle.c).

TEST INFRASTRUCTURE — only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import it.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libkllms_oracle.so")

NONE_CODE = -1
ABSENT_CODE = -2
F64_NONE_BITS = 0x7FF8C0DE00000000
F64_ABSENT_BITS = 0x7FF8C0DF00000000
F64_NONE = np.array([F64_NONE_BITS], dtype=np.uint64).view(np.float64)[0]
F64_ABSENT = np.array([F64_ABSENT_BITS], dtype=np.uint64).view(np.float64)[0]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "consensus_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        c = ctypes
        _lib.ko_vote_i32.argtypes = [c.c_void_p, c.c_int64, c.c_int32, c.c_void_p, c.c_int32, c.c_void_p, c.c_void_p]
        _lib.ko_vote_i32.restype = None
        _lib.ko_numeric_f64.argtypes = [c.c_void_p, c.c_int64, c.c_int32, c.c_double, c.c_double, c.c_void_p, c.c_void_p]
        _lib.ko_numeric_f64.restype = None
        _lib.ko_logprob_sum_f32.argtypes = [c.c_void_p, c.c_void_p, c.c_int64, c.c_void_p]
        _lib.ko_logprob_sum_f32.restype = None
        _lib.ko_weighted_vote_i32.argtypes = [c.c_void_p, c.c_void_p, c.c_int64, c.c_int32, c.c_int32, c.c_void_p, c.c_void_p,
                                              c.c_void_p, c.c_void_p]
        _lib.ko_weighted_vote_i32.restype = None
        _lib.ko_medoid_str.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p, c.c_int64, c.c_void_p, c.c_void_p]
        _lib.ko_medoid_str.restype = None
        for name in ("ko_np_sum", "ko_np_mean", "ko_np_median_sorted", "ko_np_std"):
            f = getattr(_lib, name)
            f.argtypes = [c.c_void_p, c.c_int]
            f.restype = c.c_double
    return _lib


def _ptr(a: np.ndarray) -> int:
    return a.ctypes.data


def vote(codes: np.ndarray, none_code: np.ndarray | None = None):
    """codes int32 [G, n] -> (win_code int32 [G], meta uint32 [G])."""
    codes = np.ascontiguousarray(codes, dtype=np.int32)
    G, n = codes.shape
    win = np.empty(G, dtype=np.int32)
    meta = np.empty(G, dtype=np.uint32)
    if none_code is not None:
        none_code = np.ascontiguousarray(none_code, dtype=np.int32)
    lib().ko_vote_i32(_ptr(codes), G, n, _ptr(none_code) if none_code is not None else None,
                      len(none_code) if none_code is not None else 0, _ptr(win), _ptr(meta))
    return win, meta


def numeric(vals: np.ndarray, rel_eps: float = 0.03, abs_eps: float = 1e-6):
    """vals float64 [G, n] -> (value float64 [G], meta uint32 [G])."""
    vals = np.ascontiguousarray(vals, dtype=np.float64)
    G, n = vals.shape
    value = np.empty(G, dtype=np.float64)
    meta = np.empty(G, dtype=np.uint32)
    lib().ko_numeric_f64(_ptr(vals), G, n, rel_eps, abs_eps, _ptr(value), _ptr(meta))
    return value, meta


def logprob_sum(logprobs: np.ndarray, offsets: np.ndarray) -> np.ndarray:
    logprobs = np.ascontiguousarray(logprobs, dtype=np.float32)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    out = np.empty(len(offsets) - 1, dtype=np.float32)
    lib().ko_logprob_sum_f32(_ptr(logprobs), _ptr(offsets), len(out), _ptr(out))
    return out


def weighted_vote(codes: np.ndarray, seq_logprob: np.ndarray, none_code: np.ndarray | None = None):
    """codes int32 [R, F, n], seq_logprob float32 [R, n] -> (win int32 [R*F], meta uint32 [R*F], weight float32 [R*F])."""
    codes = np.ascontiguousarray(codes, dtype=np.int32)
    seq_logprob = np.ascontiguousarray(seq_logprob, dtype=np.float32)
    R, F, n = codes.shape
    win = np.empty(R * F, dtype=np.int32)
    meta = np.empty(R * F, dtype=np.uint32)
    weight = np.empty(R * F, dtype=np.float32)
    if none_code is not None:
        none_code = np.ascontiguousarray(none_code, dtype=np.int32)
    lib().ko_weighted_vote_i32(_ptr(codes), _ptr(seq_logprob), R, F, n, _ptr(none_code) if none_code is not None else None,
                               _ptr(win), _ptr(meta), _ptr(weight))
    return win, meta, weight


def pack_string_groups(groups):
    """groups: list of lists of normalised (ASCII) strings -> (chars uint8, str_off int32, grp_off int32)."""
    blobs, str_off, grp_off = [], [0], [0]
    for grp in groups:
        for s in grp:
            b = s.encode("ascii")
            blobs.append(b)
            str_off.append(str_off[-1] + len(b))
        grp_off.append(grp_off[-1] + len(grp))
    chars = np.frombuffer(b"".join(blobs) or b"\0", dtype=np.uint8).copy()
    return chars, np.asarray(str_off, dtype=np.int32), np.asarray(grp_off, dtype=np.int32)


def medoid(groups):
    """(best index int32 [G], mean similarity float64 [G]) of groups of normalised strings."""
    chars, str_off, grp_off = pack_string_groups(groups)
    G = len(groups)
    idx = np.empty(G, dtype=np.int32)
    avg = np.empty(G, dtype=np.float64)
    lib().ko_medoid_str(_ptr(chars), _ptr(str_off), _ptr(grp_off), G, _ptr(idx), _ptr(avg))
    return idx, avg


def meta_fields(meta: np.ndarray):
    m = meta.astype(np.uint32)
    return {"idx": m & 0x3F, "support": (m >> 6) & 0x7F, "nn": (m >> 13) & 0x7F, "present": (m >> 20) & 0x7F,
            "flags": (m >> 27) & 0x1F}
