"""Multi-GPU: agents sharded by contiguous id block, graph replicated, one exchange per tick.

An agent on rank r entering link l must see the events of all lower-index agents of the same
tick, i.e. of lower ranks (a sum-only all-reduce of volumes gives wrong entry
speeds from tick 4).  So the ranks exchange their *link events* (16 B each, a few thousand per
tick) instead of a dense per-link vector (4*L bytes): every rank then resolves the global event
set -- occupancy stays replicated and identical everywhere, speeds are written for local agents
only.  The collective is an all-gather over ``torch.distributed`` (NCCL on GPUs; gloo in the CPU
tests of this module's host logic).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

EVENT_WORDS = 4  # one event = 4 x int32: link, agent (global id), kind, pos


def shard_bounds(n_total, rank, world):
    """Contiguous block of agent ids owned by ``rank`` (lower rank <=> lower agent index)."""
    base = n_total // world
    rem = n_total % world
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class _DevPtr:
    """Expose a raw device pointer to torch through the CUDA array interface (no copy)."""

    def __init__(self, ptr, n_words):
        self.__cuda_array_interface__ = dict(shape=(n_words,), typestr="<i4", data=(ptr, False), version=2)


def exchange_events(local: torch.Tensor, n_local: int, group=None):
    """All-gather variable-length event lists.  ``local``: int32 tensor [cap, 4] whose first
    ``n_local`` rows are valid.  Returns (events [n_total, 4] in rank order, counts list)."""
    world = dist.get_world_size(group)
    cnt = torch.tensor([n_local], dtype=torch.int64, device=local.device)
    counts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(counts, cnt, group=group)
    counts = [int(c.item()) for c in counts]
    m = max(counts)
    if m == 0:
        return local[:0], counts
    if local.shape[0] >= m:
        send = local[:m].contiguous()  # rows past n_local are padding
    else:
        send = torch.zeros((m, EVENT_WORDS), dtype=local.dtype, device=local.device)
        send[:n_local] = local[:n_local]
    recv = torch.empty((world, m, EVENT_WORDS), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(recv.view(-1), send.view(-1), group=group) if local.is_cuda else \
        dist.all_gather(list(recv.unbind(0)), send, group=group)
    parts = [recv[r, :counts[r]] for r in range(world) if counts[r]]
    return torch.cat(parts, dim=0), counts


class ShardedEngine:
    """One rank of a sharded run: wraps a local ``Engine`` and performs the per-tick exchange.

    The engine runs on torch's current CUDA stream; a tick is: advance + pack (engine kernels) ->
    ``all_gather_into_tensor`` of fixed-size send buffers (NCCL, same stream) -> link + resolve
    (engine kernels).  Nothing waits on the host; the buffer capacity follows the largest event
    count seen in the previous block of ticks with a 3x margin (an overflow raises)."""

    def __init__(self, lg, n_total, alpha=0.15, beta=4.0, device=0, route_arena_links=0, group=None,
                 exchange_capacity=0):
        from .engine import Engine

        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.lo, self.hi = shard_bounds(n_total, self.rank, self.world)
        self.n_total = n_total
        self.device = torch.device("cuda", device)
        self.engine = Engine(lg, self.hi - self.lo, alpha=alpha, beta=beta, device=device, agent_base=self.lo,
                             route_arena_links=route_arena_links, total_agents=n_total)
        self.engine.set_stream(torch.cuda.current_stream(self.device).cuda_stream)
        self.cap = 0
        self.p2p = False
        self._fixed_cap = exchange_capacity > 0
        self._set_cap(exchange_capacity if exchange_capacity > 0 else max(4096, (self.hi - self.lo) // 4))

    def enable_p2p(self, cap=0):
        """Move the exchange into the tick kernel (NVLink peer memory, ``k_tick_loop_p2p``): links are
        owned round-robin, events go straight into the owner's inbox, entry speeds straight back into
        the home rank's agent arrays.  One process per GPU on one node; call on all ranks before the
        first tick.  ``cap`` = events one rank may send to one owner per tick."""
        n_local = self.hi - self.lo
        cap = int(cap) if cap > 0 else max(8192, 2 * n_local // self.world + 4096)
        mine = self.engine.p2p_export(self.rank, self.world, self.n_total, cap)
        send = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(self.device)
        recv = torch.empty(self.world * len(mine), dtype=torch.uint8, device=self.device)
        dist.all_gather_into_tensor(recv, send, group=self.group)
        self.engine.p2p_connect(recv.cpu().numpy().tobytes())
        dist.barrier(group=self.group)
        self.p2p = True
        self.p2p_cap = cap

    def occupancy(self):
        """Link volumes, identical on every rank.  With the peer-memory exchange a link's volume lives
        on its owner only (0 elsewhere): sum over ranks."""
        occ = self.engine.occupancy()
        if not self.p2p:
            return occ
        t = torch.from_numpy(occ).to(self.device)
        dist.all_reduce(t, group=self.group)
        return t.cpu().numpy()

    def dense_volume_allreduce(self):
        """North_star's literal exchange ("report both"): every rank recounts the volumes of ITS agents
        (histogram of cur_link) and the dense int32[L] vector is summed with one ``all_reduce`` (NCCL).  Returns the
        global volume vector as a device tensor (valid until the next call).  This is a throughput / validation
        mode: the sum alone cannot order the entries of one tick across ranks, so the tick itself
        exchanges link events instead; the result equals ``occupancy()``."""
        ptr, n = self.engine.device_ptr("volume_recount")
        self.engine.recount_device()
        t = torch.as_tensor(_DevPtr(ptr, n), device=self.device)
        dist.all_reduce(t, group=self.group)
        return t

    def update_travel_times(self):
        """Extension (congestion-aware rerouting) on sharded agents: the link volumes of all ranks are summed with the
        dense all-reduce, every rank then computes the same BPR travel times (K4) and re-plans its own agents."""
        vol = self.dense_volume_allreduce()
        self.engine.update_travel_times_with(vol.data_ptr())

    def _set_cap(self, cap):
        cap = int(cap)
        if cap == self.cap:
            return
        self.cap = cap
        self._send = torch.zeros((cap + 1, EVENT_WORDS), dtype=torch.int32, device=self.device)
        self._recv = torch.zeros((self.world, cap + 1, EVENT_WORDS), dtype=torch.int32, device=self.device)

    def _tick(self, delta, t):
        self.engine.tick_begin_async(delta, t, self._send.data_ptr(), self.cap)
        dist.all_gather_into_tensor(self._recv.view(-1), self._send.view(-1), group=self.group)
        self.engine.tick_finish_gathered(self._recv.data_ptr(), self.world, self.cap)

    def _after_block(self):
        """One host sync per block of ticks: overflow check and capacity for the next block."""
        m = self.engine.max_events(reset=True)  # raises on overflow
        if not self._fixed_cap:
            mt = torch.tensor([m], dtype=torch.int64, device=self.device)
            dist.all_reduce(mt, op=dist.ReduceOp.MAX, group=self.group)
            want = max(4096, 3 * int(mt.item()))
            if want > self.cap or want < self.cap // 4:
                self._set_cap(want)

    def step(self, delta):
        """Trace / host-driven mode: pending departures were submitted to the local engine.  With the
        peer-memory exchange the tick runs inside ``k_tick_loop_p2p`` (all ranks must call this together);
        link volumes then live on the owner only -- read them with ``occupancy()``."""
        if self.p2p:
            self.engine.step(1, delta)
            self.engine.max_events(reset=True)  # raises on overflow
            return
        self._tick(delta, 0.0)
        self._after_block()

    def run(self, n_ticks, delta, t0):
        """Device-driven demand, ``n_ticks`` ticks starting at clock ``t0``."""
        t = t0
        if self.p2p:  # the whole block is a few cooperative launches per rank, no collective call
            self.engine.run(n_ticks, delta, t0)
            for _ in range(n_ticks):
                t += delta
            self.engine.max_events(reset=True)  # raises on overflow
            return t
        for _ in range(n_ticks):
            t += delta
            self._tick(delta, t)
        self._after_block()
        return t

    def close(self):
        if self.p2p:
            self.engine.sync()
            dist.barrier(group=self.group)  # nobody unmaps a buffer a peer's kernel may still write
        self.engine.close()
