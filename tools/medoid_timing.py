"""Device-resident timing of K4 (kc_medoid_str) on synthetic phrase groups + the C oracle on a sample (development aid)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from k_llms_b200 import _native as K  # noqa: E402
from k_llms_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--groups", type=int, default=200_000)
    ap.add_argument("--k", type=int, nargs="+", default=[5, 16, 32])
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--cpu-sample", type=int, default=2000)
    args = ap.parse_args()
    for k in args.k:
        chars, str_off, grp_off = synth.phrase_groups_numpy(args.groups, k, 7)
        d = [torch.from_numpy(a).cuda() for a in (chars, str_off, grp_off)]
        for _ in range(3):
            K.medoid_str(*d, max_group=k)
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.iters)]
        for a, b in evs:
            a.record()
            idx, avg = K.medoid_str(*d, max_group=k)
            b.record()
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in evs)[args.iters // 2]
        pairs = args.groups * k * (k - 1) // 2
        lens = np.diff(str_off).astype(np.int64).reshape(args.groups, k)
        cells = 0  # DP cells the textbook algorithm fills = sum over pairs of len_i * len_j
        s1, s2 = lens.sum(1), (lens * lens).sum(1)
        cells = int(((s1 * s1 - s2) // 2).sum())
        out = {"k": k, "groups": args.groups, "mean_len": round(float(lens.mean()), 1), "ms": round(ms, 3),
               "groups_per_s": round(args.groups / ms * 1e3), "pairs_per_s": round(pairs / ms * 1e3),
               "dp_cells_per_s": round(cells / ms * 1e3)}
        if args.cpu_sample:
            from oracle import columnar as OC
            S = min(args.cpu_sample, args.groups)
            sub = (chars[: str_off[S * k]], str_off[: S * k + 1], grp_off[: S + 1])
            ei, ea = np.empty(S, np.int32), np.empty(S, np.float64)
            t0 = time.perf_counter()
            OC.lib().ko_medoid_str(*[OC._ptr(a) for a in sub], S, OC._ptr(ei), OC._ptr(ea))
            dt = time.perf_counter() - t0
            out["cpu_groups_per_s_1core"] = round(S / dt)
            out["parity_sample"] = bool(np.array_equal(idx[:S].cpu().numpy(), ei) and np.array_equal(avg[:S].cpu().numpy(), ea))
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
