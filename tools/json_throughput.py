"""Throughput of H1 (kc_consolidate_json): n JSON texts per record in -> consensus JSON + likelihoods out, versus the
Python port of the reference's client order (json.loads -> align -> consensus -> json.dumps) on one core."""
import argparse
import ctypes
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from k_llms_b200 import _native as K  # noqa: E402
from k_llms_b200 import synth  # noqa: E402


def make_texts(R, n, seed):
    codes, none_code, vals = synth.s32_numpy(R, n, seed)
    vocab = ["alpha", "Bravo", "charlie", "DELTA", "echo", "foxtrot", "golf", "Hotel"]
    variants = [lambda w: w, lambda w: w.upper(), lambda w: w.lower() + "!", lambda w: " " + w]
    out = []
    for r in range(R):
        rec = []
        for c in range(n):
            d = {}
            for f in range(16):
                k = int(codes[r, f, c])
                d[f"f{f:02d}"] = None if k < 0 else variants[(r + c + f) % 4](vocab[k])
            for f in range(16, 24):
                k = int(codes[r, f, c])
                d[f"f{f:02d}"] = None if k < 0 else bool(k)
            for f in range(8):
                v = vals[r, f, c]
                d[f"f{24 + f:02d}"] = None if v != v else (int(v) if f < 6 else float(v))
            rec.append(json.dumps(d))
        out.append(rec)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=20000)
    ap.add_argument("--n", type=int, default=16)
    ap.add_argument("--py-records", type=int, default=300)
    ap.add_argument("--threads", type=int, default=0)
    args = ap.parse_args()
    recs = make_texts(args.records, args.n, 11)
    nbytes = sum(len(t) for r in recs for t in r)
    lib = K.load()
    R, n = len(recs), args.n
    blobs = [t.encode() for r in recs for t in r]
    texts = (ctypes.c_char_p * (R * n))(*blobs)
    lens = (ctypes.c_int64 * (R * n))(*[len(b) for b in blobs])
    out_c, out_l, status = (ctypes.c_void_p * R)(), (ctypes.c_void_p * R)(), (ctypes.c_uint8 * R)()
    cast = lambda a: ctypes.cast(a, ctypes.c_void_p)  # noqa: E731
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        K.check(lib.kc_consolidate_json(cast(texts), cast(lens), R, n, 0.03, 1e-6, 0, args.threads, cast(out_c), cast(out_l), cast(status)))
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        first = ctypes.string_at(out_c[0]).decode()
        lib.kc_free_strings(cast(out_c), R)
        lib.kc_free_strings(cast(out_l), R)
    native_ok = sum(1 for s in status if s == 0)

    from oracle import consensus_py as O
    from k_llms_b200.utils.consolidation import _format_consensus_content, _safe_parse_content
    embed = lambda t: [[0.0] for _ in t]  # noqa: E731
    sub = recs[: args.py_records]
    t0 = time.perf_counter()
    for r in sub:
        value, conf = O.client_order([_safe_parse_content(t) for t in r], embed=embed)
        _format_consensus_content(value), json.dumps(conf)
    py_dt = time.perf_counter() - t0
    print(json.dumps({"records": R, "n": n, "json_MB": round(nbytes / 1e6, 1), "native_s": round(best, 4),
                      "native_records_per_s": round(R / best), "native_GBps_of_json": round(nbytes / best / 1e9, 2),
                      "native_handled": native_ok, "host_threads": os.cpu_count(),
                      "python_port_records_per_s_1core": round(len(sub) / py_dt, 1),
                      "example": first[:120]}))


if __name__ == "__main__":
    main()
