"""Multi-GPU check (run under torchrun, >= 2 GPUs): the multicast-fused reassembly must produce, on every rank, exactly
the batch the NCCL all-gather path produces, and both must equal the oracle on a sample.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_fused_gather.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from k_llms_b200 import _native as K  # noqa: E402
from k_llms_b200 import synth  # noqa: E402
from k_llms_b200.distributed import FusedShardedConsensus, OutputLayout, ShardedConsensus  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    N, n = 200_000, 16
    codes, none_code, vals = synth.s32_torch(N, n, 555 + rank, dev)
    c2, v2 = codes.view(N * 24, n), vals.view(N * 8, n)
    layout = OutputLayout(N, 24, 8)
    lib = K.load()
    K.check(lib.kc_set_device(local))
    sp = int(torch.cuda.current_stream().cuda_stream)

    ref = ShardedConsensus(layout, dev, chunks=1)

    def compute(c, views):
        win, vmeta, value, nmeta = views
        K.check(lib.kc_vote_i32(c2.data_ptr(), N * 24, n, none_code.data_ptr(), 24, win.data_ptr(), vmeta.data_ptr(), sp))
        K.check(lib.kc_numeric_f64(v2.data_ptr(), N * 8, n, 0.03, 1e-6, value.data_ptr(), nmeta.data_ptr(), sp))

    ref.step(compute)
    torch.cuda.synchronize()

    for route in ("peers", "multimem", "peers-packed"):
        packed = route == "peers-packed"
        fused = FusedShardedConsensus(OutputLayout(N, 24, 8, packed_votes=packed), dev, route="peers" if packed else route)
        ok = fused.available()
        if rank == 0:
            print(f"route {route}: available:", ok, flush=True)
        if ok:
            fused.flat.zero_()
            dist.barrier()

            def launch(f):
                f.vote(c2.data_ptr(), N * 24, n, none_code.data_ptr(), 24, sp)
                f.numeric(v2.data_ptr(), N * 8, n, 0.03, 1e-6, sp)

            fused.step(launch)
            torch.cuda.synchronize()
            if packed:  # every rank's slot must hold the packing of that rank's NCCL-gathered full results
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
            t = torch.tensor([1 if same else 0], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            if rank == 0:
                print("fused == nccl on every rank:", bool(t.item()), flush=True)
            assert bool(t.item())
            # timing: fused vs nccl, device events, max over ranks
            for name, fn in (("nccl", lambda: ref.step(compute)), ("fused", lambda: fused.step(launch))):
                for _ in range(3):
                    fn()
                dist.barrier()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                ms = torch.tensor([e0.elapsed_time(e1) / 10], device=dev)
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
                if rank == 0:
                    print(f"{name}: {ms.item():.3f} ms/step for {N} records/rank x {world} ranks", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
