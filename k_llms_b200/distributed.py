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


# ---------------------------------------------------------------------------------------------------------------------
# Wire format + pipelined push (round 2): K1 / K2 keep their full results local; a hand-written copy kernel
# (kc_push_results, csrc/kc_push.cuh) packs chunk c and stores it into every peer's copy of the gathered buffer with
# 16-byte vectors on a second stream while chunk c + 1 is computed.


@dataclass
class WireLayout:
    """One rank's slot of the gathered buffer in the wire format: vote words | numeric values | numeric words.
    narrow (n <= 31): u16 vote word code:6|support:5|present:5, f64 value, u16 numeric word kind:2|payload:10 — 128 B per
    S32 record; wide: u32 vote word code:18|support:7|present:7, f64 value, the u32 result word — 224 B per S32 record."""
    n_records: int
    n_vote_fields: int
    n_num_fields: int
    wide: bool = False

    @property
    def gv(self) -> int:
        return self.n_records * self.n_vote_fields

    @property
    def gx(self) -> int:
        return self.n_records * self.n_num_fields

    @property
    def word_bytes(self) -> int:
        return 4 if self.wide else 2

    @property
    def off_value(self) -> int:
        return (self.gv * self.word_bytes + 15) // 16 * 16

    @property
    def off_nmeta(self) -> int:
        return self.off_value + self.gx * 8

    @property
    def nbytes(self) -> int:
        return (self.off_nmeta + self.gx * self.word_bytes + 15) // 16 * 16

    def views(self, buf):
        """(vote words, values, numeric words) of one slot (a 1-D uint8 tensor of nbytes)."""
        import torch
        wt = torch.int32 if self.wide else torch.int16
        votes = buf[0:self.gv * self.word_bytes].view(wt)
        value = buf[self.off_value:self.off_value + self.gx * 8].view(torch.float64)
        nwords = buf[self.off_nmeta:self.off_nmeta + self.gx * self.word_bytes].view(wt)
        return votes, value, nwords


def wire_pack_votes(win, vmeta, wide: bool):
    """The wire vote words of full K1 results (numpy int32 win, uint32 meta): what kc_push_results writes."""
    import numpy as np
    win = np.asarray(win).astype(np.uint32)
    m = np.asarray(vmeta).astype(np.uint32)
    support, present = (m >> 6) & 0x7F, (m >> 20) & 0x7F
    if wide:
        return ((win & 0x3FFFF) | (support << 18) | (present << 25)).astype(np.uint32)
    return ((win & 63) | ((support & 31) << 6) | ((present & 31) << 11)).astype(np.uint16)


def wire_pack_num(nmeta, wide: bool):
    import numpy as np
    m = np.asarray(nmeta).astype(np.uint32)
    if wide:
        return m
    support, nn, present, flags = (m >> 6) & 0x7F, (m >> 13) & 0x7F, (m >> 20) & 0x7F, m >> 27
    has, single, nofin = (flags & 1) != 0, (flags & 2) != 0, (flags & 8) != 0
    w = np.where(has & single, (1 << 10) | (present & 31),
                 np.where(has, (support & 31) | ((nn & 31) << 5),
                          np.where(nofin, (2 << 10) | (nn & 31) | ((present & 31) << 5), (3 << 10) | (present & 31))))
    return w.astype(np.uint16)


def wire_confidences(vote_words, num_words, wide: bool):
    """Confidences a remote consumer derives from the wire words (numpy; pvf = 1): votes round(support / present, 5);
    numeric by kind (kc_push.cuh).  Returns (vote_conf, num_conf) float64 arrays."""
    import numpy as np
    v = np.asarray(vote_words).astype(np.uint32)
    if wide:
        vs, vp = (v >> 18) & 0x7F, (v >> 25) & 0x7F
    else:
        vs, vp = (v >> 6) & 31, (v >> 11) & 31
    with np.errstate(divide="ignore", invalid="ignore"):
        vconf = np.where(vs > 0, np.round(vs / np.maximum(vp, 1), 5), np.where(vp == 0, 1.0, 0.0))
    w = np.asarray(num_words).astype(np.uint32)
    if wide:
        support, nn, present, flags = (w >> 6) & 0x7F, (w >> 13) & 0x7F, (w >> 20) & 0x7F, w >> 27
        kind = np.where((flags & 1) != 0, np.where((flags & 2) != 0, 1, 0), np.where((flags & 8) != 0, 2, 3))
    else:
        kind = (w >> 10) & 3
        support, nn = w & 31, (w >> 5) & 31
        present = np.where(kind == 2, (w >> 5) & 31, w & 31)
        nn = np.where(kind == 2, w & 31, nn)
    with np.errstate(divide="ignore", invalid="ignore"):
        nconf = np.where(kind == 0, np.round(support / np.maximum(nn, 1), 5),
                         np.where(kind == 1, 1.0 / np.maximum(present, 1),
                                  np.where(kind == 2, nn / np.maximum(present, 1), np.where(present == 0, 1.0, 0.0))))
    return vconf, nconf


class PipelinedShardedConsensus:
    """Default multi-GPU route of round 2.  Per rank: the shard is cut into `chunks` record ranges; K1 / K2 write chunk c's FULL
    results into plain local arrays (the fast kernels, no routing in their epilogues); a push kernel on a side stream then
    packs chunk c into the wire format and stores it into this rank's slot of EVERY GPU's copy of the gathered buffer
    (torch symmetric memory; 16-byte P2P stores over NVLink) while the main stream computes chunk c + 1.  One cross-GPU
    barrier closes the step.  No NCCL collective on the data path.

    compute(c, views) launches K1 / K2 for chunk c on the current stream; views = (win, vote_meta, value, num_meta) of the
    chunk (local, full results).  After step(), rank_wire_views(r) gives any rank's wire words on this GPU."""

    def __init__(self, n_records: int, n_vote_fields: int, n_num_fields: int, device, group=None, chunks: int = 4,
                 wide: bool = False, push_ctas: int = 0, mode: str = "wire"):
        import ctypes
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        assert dist.is_initialized(), "PipelinedShardedConsensus needs an initialised process group"
        assert n_records % chunks == 0 and (n_records // chunks) % 8 == 0 or chunks == 1, "chunks must cut the shard into multiples of 8 records"
        self.device, self.chunks, self.push_ctas = device, chunks, push_ctas
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.layout = WireLayout(n_records, n_vote_fields, n_num_fields, wide)
        self.full = OutputLayout(n_records, n_vote_fields, n_num_fields)
        self.chunk_records = n_records // chunks
        self.flat = symm_mem.empty(self.world * self.layout.nbytes, dtype=torch.uint8, device=device)
        self.handle = symm_mem.rendezvous(self.flat, self.group)
        self.gathered = self.flat.view(self.world, self.layout.nbytes)
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self.n_peers = self.world - 1
        self.peer_deltas = (ctypes.c_int64 * max(self.n_peers, 1))(*[ptrs[p] - ptrs[self.rank] for p in range(self.world) if p != self.rank])
        assert mode in ("wire", "pack", "dma", "hybrid")
        if mode in ("dma", "hybrid"):  # every peer's copy of the gathered buffer as a tensor, for copy-engine transfers
            self.peer_bufs = {p: self.handle.get_buffer(p, (self.world, self.layout.nbytes), torch.uint8) for p in range(self.world) if p != self.rank}
        gv, gx = self.full.gv, self.full.gx
        self.win = torch.empty(gv, dtype=torch.int32, device=device)
        self.vmeta = torch.empty(gv, dtype=torch.int32, device=device)
        # K2 writes its values straight into this rank's slot (they travel as they are); K1 writes wire words there too
        self.value = self.layout.views(self.gathered[self.rank])[1]
        self.nmeta = torch.empty(gx, dtype=torch.int32, device=device)
        self.overflow = torch.zeros(1, dtype=torch.int32, device=device)
        # high priority: the push kernel's CTAs take the SM slots the compute kernels' first wave frees, ahead of their second wave
        self.side = torch.cuda.Stream(device=device, priority=-1)
        self.done = [torch.cuda.Event() for _ in range(chunks)]
        self.pushed = torch.cuda.Event()
        self._gather = True
        self.mode = mode  # "wire": K1 writes the wire words itself; "pack": the push kernel packs the full K1 results; "dma": copy engines

    def available(self) -> bool:
        return 1 <= self.n_peers <= 7

    def my_views(self, c: int = 0):
        r = self.chunk_records
        fv, fx = self.full.n_vote_fields, self.full.n_num_fields
        return (self.win[c * r * fv:(c + 1) * r * fv], self.vmeta[c * r * fv:(c + 1) * r * fv],
                self.value[c * r * fx:(c + 1) * r * fx], self.nmeta[c * r * fx:(c + 1) * r * fx])

    def rank_wire_views(self, r: int):
        return self.layout.views(self.gathered[r])

    def _slot_ptrs(self, c: int):
        L = self.layout
        base = self.flat.data_ptr() + self.rank * L.nbytes
        r = self.chunk_records
        g0v, g0x = c * r * L.n_vote_fields, c * r * L.n_num_fields
        return base + g0v * L.word_bytes, base + L.off_value + g0x * 8, base + L.off_nmeta + g0x * L.word_bytes

    def vote(self, c: int, codes_ptr: int, n_groups: int, n: int, none_ptr: int, n_fields: int, stream_ptr: int) -> None:
        """K1 for chunk c: full results into the local arrays (+ the wire words into the slot in mode "wire" / "dma")."""
        from . import _native as K
        win, vmeta, _, _ = self.my_views(c)
        assert win.numel() == n_groups
        if self.mode == "pack":
            K.check(K.load().kc_vote_i32(codes_ptr, n_groups, n, none_ptr, n_fields, win.data_ptr(), vmeta.data_ptr(), stream_ptr))
        else:
            import ctypes
            fused = self.mode == "hybrid" and self._gather  # K1 stores its wire words into the peers' copies itself
            K.check(K.load().kc_vote_i32_wire(codes_ptr, n_groups, n, none_ptr, n_fields, win.data_ptr(), vmeta.data_ptr(), self._slot_ptrs(c)[0],
                                              1 if self.layout.wide else 0, self.n_peers if fused else 0,
                                              ctypes.addressof(self.peer_deltas) if fused else None, self.overflow.data_ptr(), stream_ptr))

    def numeric(self, c: int, vals_ptr: int, n_groups: int, n: int, rel_eps: float, abs_eps: float, stream_ptr: int) -> None:
        """K2 for chunk c: values straight into the slot, result words into the local array."""
        from . import _native as K
        _, _, value, nmeta = self.my_views(c)
        assert value.numel() == n_groups
        K.check(K.load().kc_numeric_f64(vals_ptr, n_groups, n, rel_eps, abs_eps, value.data_ptr(), nmeta.data_ptr(), stream_ptr))

    def push(self, c: int, stream_ptr: int) -> None:
        """Enqueue the replication of chunk c's part of the slot into every peer's copy on the given stream."""
        import ctypes
        from . import _native as K
        win, vmeta, value, nmeta = self.my_views(c)
        wv, wx, wm = self._slot_ptrs(c)
        K.check(K.load().kc_push_results(win.data_ptr() if self.mode == "pack" else None, vmeta.data_ptr() if self.mode == "pack" else None,
                                         win.numel(), value.data_ptr(), nmeta.data_ptr(), value.numel(), wv, wx, wm,
                                         1 if self.layout.wide else 0, 0 if self.mode in ("dma", "hybrid") else self.n_peers,
                                         ctypes.addressof(self.peer_deltas), self.overflow.data_ptr(), self.push_ctas, stream_ptr))

    def _dma(self, c: int) -> None:
        """The copy engines replicate chunk c's slot segments into every peer's copy (no SM work): all three in mode "dma",
        the two numeric ones in mode "hybrid" (K1 stores its words remotely itself)."""
        import torch
        L = self.layout
        r = self.chunk_records
        segs = [(c * r * L.n_vote_fields * L.word_bytes, r * L.n_vote_fields * L.word_bytes),
                (L.off_value + c * r * L.n_num_fields * 8, r * L.n_num_fields * 8),
                (L.off_nmeta + c * r * L.n_num_fields * L.word_bytes, r * L.n_num_fields * L.word_bytes)]
        if self.mode == "hybrid":
            segs = segs[1:]
        mine = self.gathered[self.rank]
        for p in range(self.world):
            if p == self.rank:
                continue
            dst = self.peer_bufs[p][self.rank]
            for o, nb in segs:
                dst[o:o + nb].copy_(mine[o:o + nb], non_blocking=True)

    def step(self, compute: Callable, gather: bool = True, compute_numeric: Optional[Callable] = None, compute_vote: Optional[Callable] = None):
        """compute(c, views) launches K1 + K2 of chunk c.  Mode "hybrid" wants them separately (compute_numeric, compute_vote):
        K2 of chunk c first, then the copy engines ship its values / words while K1 of chunk c (which stores its own words into
        the peers' copies) and K2 of chunk c + 1 run."""
        import torch
        main = torch.cuda.current_stream()
        self._gather = gather  # gather=False: the same launches with no remote stores, no pushes, no barrier (compute only)
        if self.mode == "hybrid" and compute_numeric is not None and compute_vote is not None:
            for c in range(self.chunks):
                compute_numeric(c, self.my_views(c))
                if gather:
                    self.push(c, int(main.cuda_stream))  # packs the numeric words into the slot (no peers: a few microseconds)
                    self.done[c].record(main)
                    self.side.wait_event(self.done[c])
                    with torch.cuda.stream(self.side):
                        self._dma(c)
                compute_vote(c, self.my_views(c))
            if gather:
                self.pushed.record(self.side)
                main.wait_event(self.pushed)
                self.handle.barrier(channel=0)
            return self.gathered
        for c in range(self.chunks):
            compute(c, self.my_views(c))
            if gather:
                self.done[c].record(main)
                self.side.wait_event(self.done[c])
                self.push(c, int(self.side.cuda_stream))  # mode "dma": packs the numeric words only (n_peers = 0)
                if self.mode == "dma":
                    with torch.cuda.stream(self.side):
                        self._dma(c)
        if gather:
            self.pushed.record(self.side)
            main.wait_event(self.pushed)
            self.handle.barrier(channel=0)
        return self.gathered

    def overflowed(self) -> bool:
        """True if some result did not fit the narrow wire words since construction (rebuild with wide=True).  Synchronises."""
        return int(self.overflow.item()) != 0
