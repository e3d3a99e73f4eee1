"""Multi-GPU sharding of a batch of independent records: one process per GPU, contiguous record
ranges, no collective during compute, ONE all-gather of the packed output columns to reassemble the batch.

Output packing per rank (bytes, in this order): win_code int32[G_v] | vote_meta uint32[G_v] | value f64[G_x] |
num_meta uint32[G_x], padded to 16 B — so a single NCCL all-gather moves everything, and the kernels write
straight into the rank's slot of the gathered buffer (in-place all-gather, send = recv + rank*chunk).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional, Tuple


def shard_range(n_records: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced split: the first (n % world) ranks get one extra record."""
    base, extra = divmod(n_records, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


@dataclass
class OutputLayout:
    n_records: int  # records per rank (equal on every rank: pad the batch if needed)
    n_vote_fields: int
    n_num_fields: int
    packed_votes: bool = False  # gathered vote results as ONE word code:18|support:7|present:7 (include/kllms_b200.h)

    @property
    def gv(self) -> int:
        return self.n_records * self.n_vote_fields

    @property
    def gx(self) -> int:
        return self.n_records * self.n_num_fields

    @property
    def vote_bytes(self) -> int:
        return self.gv * (4 if self.packed_votes else 8)

    @property
    def nbytes(self) -> int:
        raw = self.vote_bytes + self.gx * 12
        return (raw + 15) // 16 * 16

    def views(self, buf):
        """Typed views into one rank's slot (a 1-D uint8 tensor of nbytes): (win, vote_meta, value, num_meta), or with
        packed votes (packed, None, value, num_meta)."""
        import torch
        o = 0
        win = buf[o:o + self.gv * 4].view(torch.int32); o += self.gv * 4
        vmeta = None
        if not self.packed_votes:
            vmeta = buf[o:o + self.gv * 4].view(torch.int32); o += self.gv * 4
        value = buf[o:o + self.gx * 8].view(torch.float64); o += self.gx * 8
        nmeta = buf[o:o + self.gx * 4].view(torch.int32)
        return win, vmeta, value, nmeta


class ShardedConsensus:
    """Per-rank consensus + all-gather reassembly, pipelined: the shard is cut into `chunks` record ranges; the packed
    outputs of chunk c are all-gathered (asynchronously, on NCCL's stream) while chunk c+1 is being computed.

    gathered[c][r] holds rank r's outputs for its chunk c.  `compute(c, views)` fills this rank's slot of chunk c; on
    the GPU it is the two kernel launches, in the gloo/CPU tests it is a stand-in."""

    def __init__(self, layout: OutputLayout, device, group=None, chunks: int = 1):
        import torch
        import torch.distributed as dist
        assert layout.n_records % chunks == 0, "records per rank must divide into equal chunks"
        self.layout, self.device, self.chunks = layout, device, chunks
        self.chunk_layout = OutputLayout(layout.n_records // chunks, layout.n_vote_fields, layout.n_num_fields)
        self.dist = dist if (dist.is_available() and dist.is_initialized()) else None
        self.group = group
        self.world = self.dist.get_world_size(group) if self.dist else 1
        self.rank = self.dist.get_rank(group) if self.dist else 0
        self.gathered = torch.empty((chunks, self.world, self.chunk_layout.nbytes), dtype=torch.uint8, device=device)

    def my_views(self, c: int = 0):
        return self.chunk_layout.views(self.gathered[c, self.rank])

    def rank_views(self, r: int, c: int = 0):
        return self.chunk_layout.views(self.gathered[c, r])

    def step(self, compute: Callable, gather: bool = True):
        works = []
        for c in range(self.chunks):
            compute(c, self.my_views(c))
            if gather and self.dist and self.world > 1:
                # in-place all-gather (input = this rank's slice of the output); async: it waits for the kernels just
                # enqueued on the current stream, then runs on the communicator's stream beside the next chunk's kernels
                works.append(self.dist.all_gather_into_tensor(self.gathered[c].view(-1), self.gathered[c, self.rank],
                                                              group=self.group, async_op=True))
        for w in works:
            w.wait()
        return self.gathered


class FusedShardedConsensus:
    """Fused compute + reassembly over NVLink (no NCCL collective on the data path).

    The gathered buffer [world][nbytes] is torch symmetric memory: every GPU maps every rank's copy, and an NVSwitch
    multicast mapping covers all of them.  The kernels store each result into all copies while they compute the next
    groups, either way:

    * route "peers" (default): one local store plus one P2P store per peer (kc_*_peers, KC_OUT_PEERS) — every GPU
      receives world-1 shares and sends as many;
    * route "multimem": one `multimem.st` to the multicast address (kc_*_ex, KC_OUT_MULTIMEM); the switch replicates it
      into all copies, the sender's included — every GPU receives `world` shares and sends one.

    One cross-GPU barrier (symmetric-memory signal pads) closes the step.  `available()` says whether the mappings
    exist — callers fall back to ShardedConsensus (NCCL) otherwise.
    """

    def __init__(self, layout: OutputLayout, device, group=None, route: str = "peers"):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        assert dist.is_initialized(), "FusedShardedConsensus needs an initialised process group"
        assert route in ("peers", "multimem")
        self.layout, self.device, self.route = layout, device, route
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.flat = symm_mem.empty(self.world * layout.nbytes, dtype=torch.uint8, device=device)
        self.handle = symm_mem.rendezvous(self.flat, self.group)
        self.gathered = self.flat.view(self.world, layout.nbytes)
        self.mc_ptr = int(self.handle.multicast_ptr or 0)
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        import ctypes
        self.n_peers = self.world - 1
        self.peer_deltas = (ctypes.c_int64 * max(self.n_peers, 1))(*[ptrs[p] - ptrs[self.rank] for p in range(self.world) if p != self.rank])
        if layout.packed_votes:
            assert route == "peers", "packed vote results travel as P2P stores"
            # the full K1 results of this rank's shard stay local (the owner's decoder needs the first-seen index)
            self.local_win = torch.empty(layout.gv, dtype=torch.int32, device=device)
            self.local_vmeta = torch.empty(layout.gv, dtype=torch.int32, device=device)
            self.overflow = torch.zeros(1, dtype=torch.int32, device=device)

    def available(self) -> bool:
        if self.route == "multimem":
            return self.mc_ptr != 0
        return 1 <= self.n_peers <= 7

    def slot_pointers(self, multicast: bool = True):
        """(win or packed, vote_meta, value, num_meta) raw addresses of THIS rank's slot, in the multicast or the local
        mapping."""
        base = (self.mc_ptr if multicast else self.flat.data_ptr()) + self.rank * self.layout.nbytes
        gv, gx, vb = self.layout.gv, self.layout.gx, self.layout.vote_bytes
        return base, base + gv * 4, base + vb, base + vb + gx * 8

    def rank_views(self, r: int):
        return self.layout.views(self.gathered[r])

    def vote(self, codes_ptr: int, n_groups: int, n: int, none_ptr: int, n_fields: int, stream_ptr: int) -> None:
        """K1 over this rank's shard, results to every GPU's copy of this rank's slot (enqueued on the stream)."""
        import ctypes
        from . import _native as K
        lib = K.load()
        if self.route == "multimem":
            win, vmeta, _, _ = self.slot_pointers(multicast=True)
            K.check(lib.kc_vote_i32_ex(codes_ptr, n_groups, n, none_ptr, n_fields, win, vmeta, K.OUT_MULTIMEM, stream_ptr))
        elif self.layout.packed_votes:
            packed, _, _, _ = self.slot_pointers(multicast=False)
            K.check(lib.kc_vote_i32_peers_packed(codes_ptr, n_groups, n, none_ptr, n_fields, self.local_win.data_ptr(),
                                                 self.local_vmeta.data_ptr(), packed, self.n_peers,
                                                 ctypes.addressof(self.peer_deltas), self.overflow.data_ptr(), stream_ptr))
        else:
            win, vmeta, _, _ = self.slot_pointers(multicast=False)
            K.check(lib.kc_vote_i32_peers(codes_ptr, n_groups, n, none_ptr, n_fields, win, vmeta, self.n_peers,
                                          ctypes.addressof(self.peer_deltas), stream_ptr))

    def numeric(self, vals_ptr: int, n_groups: int, n: int, rel_eps: float, abs_eps: float, stream_ptr: int) -> None:
        import ctypes
        from . import _native as K
        lib = K.load()
        if self.route == "multimem":
            _, _, value, nmeta = self.slot_pointers(multicast=True)
            K.check(lib.kc_numeric_f64_ex(vals_ptr, n_groups, n, rel_eps, abs_eps, value, nmeta, K.OUT_MULTIMEM, stream_ptr))
        else:
            _, _, value, nmeta = self.slot_pointers(multicast=False)
            K.check(lib.kc_numeric_f64_peers(vals_ptr, n_groups, n, rel_eps, abs_eps, value, nmeta, self.n_peers,
                                             ctypes.addressof(self.peer_deltas), stream_ptr))

    def packed_overflowed(self) -> bool:
        """True if a winning code did not fit 18 bits in some step since construction (the gathered words are then unusable:
        rebuild with packed_votes=False).  Synchronises."""
        return bool(self.layout.packed_votes and int(self.overflow.item()) != 0)

    def step(self, launch: Callable):
        """launch(self) enqueues K1/K2 through self.vote / self.numeric on the current stream; the barrier afterwards makes
        every rank's stores visible everywhere."""
        launch(self)
        self.handle.barrier(channel=0)
        return self.gathered
