"""Cross-rank correctness of every multi-GPU reassembly route (run under torchrun, >= 2 GPUs; tests/test_gpu_multi.py spawns it):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_reassembly.py [--records 200000] [--n 16]

On EVERY rank, EVERY rank's slot of the gathered buffer must equal, bit for bit,
  * route push (narrow / wide wire words, 1 and several chunks): the packing of the results the NCCL all-gather delivers;
  * routes peers, peers-packed, multimem (stores fused into K1 / K2): the NCCL all-gather itself (packed: its packing);
and the NCCL-gathered results of a sample of groups must equal the C oracle.  A wrong peer offset fails here.
Prints one line per route (timings: device events, max over ranks) and exits non-zero on any mismatch."""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from k_llms_b200 import _native as K  # noqa: E402
from k_llms_b200 import synth  # noqa: E402
from k_llms_b200.distributed import (FusedShardedConsensus, OutputLayout, PipelinedShardedConsensus, ShardedConsensus,  # noqa: E402
                                     wire_confidences, wire_pack_num, wire_pack_votes)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=200_000)
    ap.add_argument("--cands", "--n", dest="n", type=int, default=16)
    ap.add_argument("--time", type=int, default=1)
    ap.add_argument("--short", type=int, default=0, help="sweep only the promising configurations")
    ap.add_argument("--sweep", type=int, default=0, help="time the push route over modes / chunks / CTA counts instead of checking every route")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    N, n = args.records, args.n
    codes, none_code, vals = synth.s32_torch(N, n, 555 + rank, dev)
    c2, v2 = codes.view(N * 24, n), vals.view(N * 8, n)
    lib = K.load()
    K.check(lib.kc_set_device(local))
    sp = int(torch.cuda.current_stream().cuda_stream)

    def say(*a):
        if rank == 0:
            print(*a, flush=True)

    def all_true(flag: bool, what: str):
        t = torch.tensor([1 if flag else 0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        say(f"{what}: {'ok' if t.item() else 'MISMATCH'}")
        if not t.item():
            raise SystemExit(f"rank {rank}: {what} failed")

    def timed(fn, reps=10):
        for _ in range(3):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # --- the independent route: NCCL all-gather of the full results
    ref = ShardedConsensus(OutputLayout(N, 24, 8), dev, chunks=1)

    def compute_full(c, views):
        win, vmeta, value, nmeta = views
        K.check(lib.kc_vote_i32(c2.data_ptr(), N * 24, n, none_code.data_ptr(), 24, win.data_ptr(), vmeta.data_ptr(), sp))
        K.check(lib.kc_numeric_f64(v2.data_ptr(), N * 8, n, 0.03, 1e-6, value.data_ptr(), nmeta.data_ptr(), sp))

    ref.step(compute_full)
    torch.cuda.synchronize()
    # the NCCL-gathered results of this rank against the C oracle (a sample; the kernels' own tests cover the rest)
    from oracle import columnar as OC
    S = 2000
    w, m, v, x = [t.cpu().numpy() for t in ref.rank_views(rank)]
    ew, em = OC.vote(codes.view(N * 24, n)[:S * 24].cpu().numpy(), none_code.cpu().numpy())
    ev, en = OC.numeric(vals.view(N * 8, n)[:S * 8].cpu().numpy())
    all_true(bool(np.array_equal(w[:S * 24], ew) and np.array_equal(m[:S * 24].view(np.uint32), em)
                  and np.array_equal(v[:S * 8].view(np.uint64), ev.view(np.uint64)) and np.array_equal(x[:S * 8].view(np.uint32), en)),
             "NCCL-gathered results == C oracle (sample)")
    if args.time:
        say(f"nccl all-gather (1 chunk): {timed(lambda: ref.step(compute_full)):.3f} ms/step, {N} records/rank x {world} ranks, n={n}")

    if args.sweep:
        wide = n > 31
        for mode, chunks, ctas in [("wire", 4, 0), ("wire", 4, 148), ("wire", 4, 64), ("wire", 4, 32), ("wire", 4, 16), ("wire", 2, 0), ("wire", 2, 32), ("wire", 8, 32),
                                   ("wire", 1, 32), ("pack", 4, 32), ("dma", 4, 0), ("dma", 2, 0), ("dma", 8, 0), ("hybrid", 1, 0), ("hybrid", 2, 0), ("hybrid", 4, 0), ("hybrid", 8, 0)]:
            if N % (chunks * 8):
                continue
            if args.short and (mode, chunks, ctas) not in (("wire", 4, 0), ("wire", 2, 0), ("hybrid", 1, 0), ("hybrid", 2, 0)):
                continue
            pipe = PipelinedShardedConsensus(N, 24, 8, dev, chunks=chunks, wide=wide, mode=mode, push_ctas=ctas)
            R = N // chunks

            def compute(c, views, pipe=pipe, R=R):
                pipe.vote(c, c2[c * R * 24:(c + 1) * R * 24].data_ptr(), R * 24, n, none_code.data_ptr(), 24, sp)
                pipe.numeric(c, v2[c * R * 8:(c + 1) * R * 8].data_ptr(), R * 8, n, 0.03, 1e-6, sp)

            def cnum(c, views, pipe=pipe, R=R):
                pipe.numeric(c, v2[c * R * 8:(c + 1) * R * 8].data_ptr(), R * 8, n, 0.03, 1e-6, sp)

            def cvote(c, views, pipe=pipe, R=R):
                pipe.vote(c, c2[c * R * 24:(c + 1) * R * 24].data_ptr(), R * 24, n, none_code.data_ptr(), 24, sp)

            t_all = timed(lambda: pipe.step(compute, compute_numeric=cnum, compute_vote=cvote), 20)
            t_c = timed(lambda: pipe.step(compute, gather=False, compute_numeric=cnum, compute_vote=cvote), 20)
            nop = lambda c, v: None  # noqa: E731
            t_p = timed(lambda: pipe.step(nop, compute_numeric=nop, compute_vote=nop), 20)
            say(f"sweep mode={mode} chunks={chunks} ctas={ctas}: step {t_all:.3f} ms, compute alone {t_c:.3f}, push alone {t_p:.3f} "
                f"(floor {(world - 1) * pipe.layout.nbytes / 770e9 * 1e3:.3f}), {world} ranks")
            del pipe
        dist.destroy_process_group()
        return

    # --- route push: narrow and wide words, one and several chunks
    for wide, chunks, mode in [(w, c, m) for w in ([False, True] if n <= 31 else [True]) for c, m in ((1, "wire"), (4, "wire"), (4, "pack"), (4, "dma"), (4, "hybrid"), (1, "hybrid"))]:
        if True:
            pipe = PipelinedShardedConsensus(N, 24, 8, dev, chunks=chunks, wide=wide, mode=mode)
            pipe.flat.zero_()
            dist.barrier()
            R = N // chunks

            def compute(c, views, pipe=pipe, R=R):
                pipe.vote(c, c2[c * R * 24:(c + 1) * R * 24].data_ptr(), R * 24, n, none_code.data_ptr(), 24, sp)
                pipe.numeric(c, v2[c * R * 8:(c + 1) * R * 8].data_ptr(), R * 8, n, 0.03, 1e-6, sp)

            def cnum(c, views, pipe=pipe, R=R):
                pipe.numeric(c, v2[c * R * 8:(c + 1) * R * 8].data_ptr(), R * 8, n, 0.03, 1e-6, sp)

            def cvote(c, views, pipe=pipe, R=R):
                pipe.vote(c, c2[c * R * 24:(c + 1) * R * 24].data_ptr(), R * 24, n, none_code.data_ptr(), 24, sp)

            pipe.step(compute, compute_numeric=cnum, compute_vote=cvote)
            torch.cuda.synchronize()
            same = not pipe.overflowed()
            for r in range(world):
                rw, rm, rv, rn = [t.cpu().numpy() for t in ref.rank_views(r)]
                pv, pval, pn = [t.cpu().numpy() for t in pipe.rank_wire_views(r)]
                ut = np.uint32 if wide else np.uint16
                same = same and np.array_equal(pv.view(ut), wire_pack_votes(rw, rm.view(np.uint32), wide))
                same = same and np.array_equal(pval.view(np.uint64), rv.view(np.uint64))
                same = same and np.array_equal(pn.view(ut), wire_pack_num(rn.view(np.uint32), wide))
                if r == (rank + 1) % world:  # what a remote consumer derives from the words == the library's confidences
                    vc, nc = wire_confidences(pv.view(ut), pn.view(ut), wide)
                    cv = K.confidence(ref.rank_views(r)[1], False).cpu().numpy()
                    cn = K.confidence(ref.rank_views(r)[3], True).cpu().numpy()
                    same = same and np.array_equal(vc, cv) and np.array_equal(nc, cn)
            lw, lm = pipe.win.cpu().numpy(), pipe.vmeta.cpu().numpy()  # the owner keeps its full K1 results
            rw, rm, _, _ = [t.cpu().numpy() for t in ref.rank_views(rank)]
            same = same and np.array_equal(lw, rw) and np.array_equal(lm, rm)
            all_true(bool(same), f"push {mode} {'wide' if wide else 'narrow'} x{chunks}: every slot on every rank == packing of the NCCL-gathered results")
            if args.time:
                say(f"push {mode} {'wide' if wide else 'narrow'} x{chunks}: {timed(lambda: pipe.step(compute, compute_numeric=cnum, compute_vote=cvote)):.3f} ms/step "
                    f"({pipe.layout.nbytes / N:.0f} B/record on the wire)")
            del pipe

    # --- stores fused into K1 / K2
    for route in ("peers", "multimem", "peers-packed"):
        packed = route == "peers-packed"
        fused = FusedShardedConsensus(OutputLayout(N, 24, 8, packed_votes=packed), dev, route="peers" if packed else route)
        if not fused.available():
            say(f"route {route}: not available on this box")
            continue
        fused.flat.zero_()
        dist.barrier()

        def launch(f):
            f.vote(c2.data_ptr(), N * 24, n, none_code.data_ptr(), 24, sp)
            f.numeric(v2.data_ptr(), N * 8, n, 0.03, 1e-6, sp)

        fused.step(launch)
        torch.cuda.synchronize()
        if packed:
            same = not fused.packed_overflowed()
            for r in range(world):
                rw, rm, rv, rn = ref.rank_views(r)
                pw, _, pv, pn = fused.rank_views(r)
                expect = (rw & 0x3FFFF) | (((rm >> 6) & 0x7F) << 18) | (((rm >> 20) & 0x7F) << 25)
                same = same and torch.equal(pw, expect) and torch.equal(pv.view(torch.int64), rv.view(torch.int64)) and torch.equal(pn, rn)
            rw, rm, _, _ = ref.rank_views(rank)
            same = same and torch.equal(fused.local_win, rw) and torch.equal(fused.local_vmeta, rm)
        else:
            same = torch.equal(fused.gathered, ref.gathered[0])
        all_true(bool(same), f"fused {route}: every slot on every rank == NCCL all-gather")
        if args.time:
            say(f"fused {route}: {timed(lambda: fused.step(launch)):.3f} ms/step")
        del fused
    say("ALL ROUTES OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
