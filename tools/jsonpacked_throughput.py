"""Throughput of H1g (kc_consolidate_json_packed): S32 candidate texts in pinned host memory -> consensus / likelihoods texts,
wall clock around the C-ABI call, with the per-stage device times the call reports.  One JSON line per configuration."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from k_llms_b200 import _native as K  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=262144)
    ap.add_argument("--n", type=int, default=16)
    ap.add_argument("--reps", type=int, default=4)
    ap.add_argument("--chunk-mb", default="64")
    ap.add_argument("--streams", default="3")
    ap.add_argument("--pageable", action="store_true", help="input blob in ordinary (not page-locked) memory")
    args = ap.parse_args()
    t0 = time.perf_counter()
    blob, off = K.s32_texts_packed(args.records, args.n, 11, pinned=not args.pageable)
    gen_s = time.perf_counter() - t0
    for chunk in args.chunk_mb.split(","):
        for streams in args.streams.split(","):
            os.environ["KC_JSON_CHUNK_MB"], os.environ["KC_JSON_STREAMS"] = chunk, streams
            walls, stats = [], None
            for i in range(args.reps + 1):
                t0 = time.perf_counter()
                res = K.consolidate_json_packed(blob, off, args.n)
                dt = time.perf_counter() - t0
                stats = res.stats.as_dict()
                first = res.content(0)
                res.close()
                if i:
                    walls.append(dt)
            best = min(walls)
            print(json.dumps({"records": args.records, "n": args.n, "chunk_mb": int(chunk), "streams": int(streams),
                              "pinned_input": not args.pageable, "json_GB": round(stats["input_bytes"] / 1e9, 3),
                              "best_s": round(best, 4), "mean_s": round(sum(walls) / len(walls), 4),
                              "records_per_s": round(args.records / best), "json_GBps": round(stats["input_bytes"] / best / 1e9, 2),
                              "stats": {k: (round(v, 2) if isinstance(v, float) else v) for k, v in stats.items()},
                              "generate_s": round(gen_s, 1), "example": first[:80]}), flush=True)


if __name__ == "__main__":
    main()
