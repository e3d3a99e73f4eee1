"""Per-request latency and concurrency of the drop-in client path: consolidate_parsed_chat_completions on synthetic
32-field ParsedChatCompletions (schema S32), n in {3, 16}: p50 / p99 of the whole function (pydantic assembly included), of its
consensus part alone (candidate texts -> consensus + likelihoods: the device JSON path), and of the same part on the CPU with
the oracle port of the reference's client order (json.loads -> align -> consensus -> json.dumps); then T threads calling
concurrently.  One JSON line."""
import argparse
import json
import os
import statistics
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pct(xs, p):
    xs = sorted(xs)
    return xs[min(len(xs) - 1, int(p * len(xs)))]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--requests", type=int, default=400)
    ap.add_argument("--threads", type=int, default=64)
    args = ap.parse_args()
    from openai.types.chat import ParsedChatCompletion
    from k_llms_b200 import _native as K
    from k_llms_b200.utils import consolidation as C
    from oracle import consensus_py as O
    embed = lambda t: [[0.0] for _ in t]  # noqa: E731
    out = {"requests": args.requests}
    for n in (3, 16):
        blob, off = K.s32_texts_packed(args.requests, n, 77 + n, pinned=False)
        text = blob.tobytes()
        recs = [[text[off[r * n + c]:off[r * n + c + 1]].decode() for c in range(n)] for r in range(args.requests)]
        comps = [ParsedChatCompletion.model_validate({"id": "x", "object": "chat.completion", "created": 0, "model": "m",
                                                      "choices": [{"index": i, "finish_reason": "stop", "message": {"role": "assistant", "content": t}}
                                                                  for i, t in enumerate(r)]}) for r in recs]
        for c in comps[:20]:  # warm-up: pools, streams, kernels
            C.consolidate_parsed_chat_completions(c, embed, None)
        whole, part, cpu = [], [], []
        for c, r in zip(comps, recs):
            t0 = time.perf_counter()
            res = C.consolidate_parsed_chat_completions(c, embed, None)
            whole.append((time.perf_counter() - t0) * 1e3)
            t0 = time.perf_counter()
            pair = C._combiner.run(r, 0.03, 1e-6)
            part.append((time.perf_counter() - t0) * 1e3)
            t0 = time.perf_counter()
            v, conf = O.client_order([json.loads(t) for t in r], embed=embed)
            exp = (json.dumps(v), json.dumps(conf))
            cpu.append((time.perf_counter() - t0) * 1e3)
            assert pair == exp and res.choices[0].message.content == exp[0], "device path differs from the oracle"
        # concurrency: T threads, each its share of the requests
        def worker(lo, hi):
            for i in range(lo, hi):
                C._combiner.run(recs[i], 0.03, 1e-6)  # what the per-request client path calls: concurrent requests share launches
        per = max(1, args.requests // args.threads)
        ths = [threading.Thread(target=worker, args=(i * per, min(args.requests, (i + 1) * per))) for i in range(args.threads)]
        t0 = time.perf_counter()
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        conc_s = time.perf_counter() - t0
        done = min(args.requests, per * args.threads)
        # the same requests as ONE batched call (what a server would do with queued requests)
        C._native_consolidate(recs, 0.03, 1e-6)  # warm-up: the first large call grows the pooled buffers
        t0 = time.perf_counter()
        C._native_consolidate(recs, 0.03, 1e-6)
        batch_s = time.perf_counter() - t0
        out[f"n{n}"] = {"function_ms": {"p50": round(statistics.median(whole), 3), "p99": round(pct(whole, 0.99), 3)},
                        "consensus_part_gpu_ms": {"p50": round(statistics.median(part), 3), "p99": round(pct(part, 0.99), 3)},
                        "consensus_part_cpu_port_ms": {"p50": round(statistics.median(cpu), 3), "p99": round(pct(cpu, 0.99), 3)},
                        "concurrent": {"threads": args.threads, "requests_per_s": round(done / conc_s), "sequential_requests_per_s": round(1e3 / statistics.mean(part))},
                        "one_batched_call_requests_per_s": round(args.requests / batch_s)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
