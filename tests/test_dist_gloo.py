"""CPU, world_size 2, gloo: the multi-GPU sharding + all-gather reassembly logic (k_llms_b200/distributed.py) with the
columnar C oracle standing in for the kernels.  The assembled batch must equal a single-process run."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_records, n, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from k_llms_b200 import synth
        from k_llms_b200.distributed import OutputLayout, ShardedConsensus, shard_range
        from oracle import columnar as OC
        codes, none_code, vals = synth.s32_numpy(n_records, n, 77)  # the same global batch on every rank
        lo, hi = shard_range(n_records, world, rank)
        per = n_records // world
        assert hi - lo == per
        layout = OutputLayout(per, 24, 8)
        chunks = 4
        sharded = ShardedConsensus(layout, torch.device("cpu"), chunks=chunks)
        step = per // chunks

        def compute(c, views):
            win, vmeta, value, nmeta = views
            a, b = lo + c * step, lo + (c + 1) * step
            w, m = OC.vote(codes[a:b].reshape(-1, n), none_code)
            v, nm = OC.numeric(vals[a:b].reshape(-1, n))
            win.copy_(torch.from_numpy(w))
            vmeta.copy_(torch.from_numpy(m.view(np.int32)))
            value.copy_(torch.from_numpy(v))
            nmeta.copy_(torch.from_numpy(nm.view(np.int32)))

        sharded.step(compute)
        parts = [sharded.rank_views(r, c) for r in range(world) for c in range(chunks)]  # global record order
        full = [torch.cat([p[k] for p in parts]).numpy() for k in range(4)]
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), win=full[0], vmeta=full[1], value=full[2], nmeta=full[3])
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_and_all_gather(tmp_path):
    from k_llms_b200 import synth
    from k_llms_b200.distributed import shard_range
    from oracle import columnar as OC
    assert [shard_range(10, 3, r) for r in range(3)] == [(0, 4), (4, 7), (7, 10)]
    n_records, n, world = 512, 8, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_records, n, str(tmp_path)), nprocs=world, join=True)
    codes, none_code, vals = synth.s32_numpy(n_records, n, 77)
    w, m = OC.vote(codes.reshape(-1, n), none_code)
    v, nm = OC.numeric(vals.reshape(-1, n))
    for rank in range(world):
        got = np.load(os.path.join(tmp_path, f"rank{rank}.npz"))
        assert np.array_equal(got["win"], w)
        assert np.array_equal(got["vmeta"].view(np.uint32), m)
        assert np.array_equal(got["value"].view(np.uint64), v.view(np.uint64))
        assert np.array_equal(got["nmeta"].view(np.uint32), nm)


def test_output_layout_packed_votes_views():
    """Packed gathered vote results: one word per vote group; value / result-word columns follow at the packed offset."""
    import torch
    from k_llms_b200.distributed import OutputLayout
    full, packed = OutputLayout(10, 24, 8), OutputLayout(10, 24, 8, packed_votes=True)
    assert full.nbytes == (240 * 8 + 80 * 12 + 15) // 16 * 16 and packed.nbytes == (240 * 4 + 80 * 12 + 15) // 16 * 16
    buf = torch.arange(packed.nbytes, dtype=torch.int64).to(torch.uint8)
    words, vmeta, value, nmeta = packed.views(buf)
    assert vmeta is None and words.numel() == 240 and value.numel() == 80 and nmeta.numel() == 80
    assert value.data_ptr() - buf.data_ptr() == 240 * 4 and nmeta.data_ptr() - buf.data_ptr() == 240 * 4 + 80 * 8
    w, m, v, nm = full.views(torch.zeros(full.nbytes, dtype=torch.uint8))
    assert m.numel() == 240 and v.data_ptr() - w.data_ptr() == 240 * 8
