"""BASELINE config 3 on one GPU: nested JSON (depth 4) with list fields, n = 8, element-wise merge through
consensus_values_batch (host planner in Python -> one K1 / K2 / K4 launch for the whole batch -> host epilogue).
Prints one JSON line; the split shows where the time goes (development aid)."""
import argparse
import json
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=4000)
    ap.add_argument("--n", type=int, default=8)
    args = ap.parse_args()
    import torch
    from k_llms_b200.utils import consensus_utils as CU
    from oracle.gen_golden import _record_candidates  # the generator of the parity test (tests/test_gpu_product.py)

    rng = random.Random(8)
    records = [_record_candidates(rng, args.n, depth=3) for _ in range(args.records)]
    embed = lambda texts: [[0.0] for _ in texts]  # noqa: E731
    st = CU.ConsensusSettings()
    CU.consensus_values_batch(records[:50], st, embed)  # warm-up (library load, allocator)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    plan = CU._plan_for(args.n, st)
    roots = [plan.add(r, 1.0, embed) for r in records]
    t1 = time.perf_counter()
    res = plan.run()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    out = [plan.materialise(root, res) for root in roots]
    t3 = time.perf_counter()
    print(json.dumps({"config": f"config 3: {args.records} nested records (depth 4, list fields), n={args.n}",
                      "groups": {"vote": len(plan.vote_rows), "numeric": len(plan.num_rows), "medoid": len(plan.medoid_groups)},
                      "plan_s": round(t1 - t0, 4), "gpu_s": round(t2 - t1, 4), "materialise_s": round(t3 - t2, 4),
                      "records_per_s": round(args.records / (t3 - t0)), "host_threads_used": 1, "outputs": len(out)}))


if __name__ == "__main__":
    main()
