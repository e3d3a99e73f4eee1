"""Throughput of H1g (kc_consolidate_json_packed): S32 candidate texts in pinned host memory -> consensus / likelihoods texts,
wall clock around the C-ABI call, with the per-stage device times the call reports.  One JSON line per configuration."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from k_llms_b200 import _native as K  # noqa: E402


def invoice_texts(records, n, seed, nested=False):
    """An extraction-like schema with FREE-TEXT fields (multi-word strings -> similarity medoid, K4) next to enums, bools and
    numbers: 4 phrases, 3 enums, 2 bools, 3 numbers per record; every candidate copies the record's truth with probability 0.8
    per field, otherwise a variant (case / punctuation / one word changed / another value), None with probability 0.05."""
    import random
    rng = random.Random(seed)
    vendors = ["Acme Industrial Supply Co", "Globex Logistics and Freight", "Initech Software Services Ltd", "Umbrella Medical Devices Inc"]
    streets = ["12 Rue de la Paix 75002 Paris", "221B Baker Street London NW1", "1600 Amphitheatre Parkway Mountain View", "5 Avenue Anatole France Paris"]
    terms = ["net 30 days from invoice date", "payment due on receipt", "2 percent 10 net 30", "net 60 days end of month"]
    notes = ["deliver to the rear loading dock", "fragile handle with care", "partial shipment remaining items to follow", "customer will collect in person"]

    def vary(p):
        r = rng.random()
        if r < 0.3:
            return p.upper()
        if r < 0.6:
            return p.replace(" ", ", ", 1) + "."
        words = p.split()
        words[rng.randrange(len(words))] = rng.choice(["north", "30", "depot", "ltd"])
        return " ".join(words)

    out = []
    for _ in range(records):
        truth = {"vendor": rng.choice(vendors), "address": rng.choice(streets), "terms": rng.choice(terms), "note": rng.choice(notes),
                 "currency": rng.choice(["EUR", "USD", "GBP"]), "status": rng.choice(["paid", "open", "overdue"]), "kind": rng.choice(["invoice", "credit note"]),
                 "taxable": rng.random() < 0.5, "signed": rng.random() < 0.5,
                 "total": round(rng.uniform(10, 9000), 2), "tax": round(rng.uniform(1, 900), 2), "items": rng.randrange(1, 40)}
        cands = []
        for _c in range(n):
            d = {}
            for k, v in truth.items():
                r = rng.random()
                if r < 0.05:
                    v = None
                elif r < 0.25:
                    if isinstance(v, bool):
                        v = not v
                    elif isinstance(v, str):
                        v = vary(v) if " " in v and len(v) > 12 else v.upper()
                    elif isinstance(v, float):
                        v = round(v * rng.choice([1.0, 1.01, 10.0]), 2)
                    else:
                        v = v + rng.choice([0, 1])
                d[k] = v
            if nested:  # the same fields as an extraction schema would nest them (depth 3)
                d = {"vendor": {"name": d["vendor"], "location": {"address": d["address"], "currency": d["currency"]}},
                     "payment": {"terms": d["terms"], "status": d["status"], "signed": d["signed"]},
                     "amounts": {"total": d["total"], "tax": d["tax"], "taxable": d["taxable"]},
                     "kind": d["kind"], "note": d["note"], "items": d["items"]}
            cands.append(json.dumps(d))
        out.append(cands)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", choices=["s32", "invoice", "invoice_nested"], default="s32",
                    help="s32: the bench schema (enum / bool / number fields); invoice: 12 fields, 4 of them free text (medoid, K4); "
                         "invoice_nested: the same fields in nested objects (depth 3)")
    ap.add_argument("--records", type=int, default=262144)
    ap.add_argument("--n", type=int, default=16)
    ap.add_argument("--reps", type=int, default=4)
    ap.add_argument("--chunk-mb", default="64")
    ap.add_argument("--streams", default="3")
    ap.add_argument("--pageable", action="store_true", help="input blob in ordinary (not page-locked) memory")
    args = ap.parse_args()
    t0 = time.perf_counter()
    if args.workload != "s32":
        blob, off, _n = K.pack_texts(invoice_texts(args.records, args.n, 11, nested=args.workload == "invoice_nested"), pinned=not args.pageable)
    else:
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
            print(json.dumps({"workload": args.workload, "records": args.records, "n": args.n, "chunk_mb": int(chunk), "streams": int(streams),
                              "pinned_input": not args.pageable, "json_GB": round(stats["input_bytes"] / 1e9, 3),
                              "best_s": round(best, 4), "mean_s": round(sum(walls) / len(walls), 4),
                              "records_per_s": round(args.records / best), "json_GBps": round(stats["input_bytes"] / best / 1e9, 2),
                              "stats": {k: (round(v, 2) if isinstance(v, float) else v) for k, v in stats.items()},
                              "generate_s": round(gen_s, 1), "example": first[:80]}), flush=True)


if __name__ == "__main__":
    main()
