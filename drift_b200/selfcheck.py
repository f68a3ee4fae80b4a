"""N-rank parity self-check: a device-demand run sharded over all ranks of the current process group
must equal the same run on ONE engine over all agents, bit for bit (SURVEY.md 8e acceptance: "identical
occupancy vectors, speeds and trip records for P in {1,2,4,8}").

Used by ``bench.py`` (prints ``parity_check`` in the JSON line of every multi-GPU run and exits non-zero
on a mismatch) and by ``tests/test_gpu_dist.py`` (worlds 2/3/4/8).  The single-engine run happens on
rank 0 and is broadcast; every rank compares its own shard.
"""
from __future__ import annotations

import numpy as np

FIELDS = ("state", "dist", "speed", "route_idx", "trip_count")


def _drive(run, eng, occupancy, lg, start, first, cands, hourly, ticks, seed, router, hier=None, reroute=None):
    if hier is not None:
        eng.set_hierarchy(hier)
    eng.set_demand(seed, hourly, start, chain_capacity=4, router=router)
    eng.set_option("route_window", ticks)
    eng.push_trips(first)
    t, occs = 0.0, []
    for b in range(len(cands)):
        eng.push_trips(cands[b])
        if reroute is not None and b > 0:  # extension: BPR travel times from the global volumes, re-plan en-route agents
            reroute()
            eng.reroute_moving(router)
            eng.invalidate_prefetch(t, 1.0, ticks)
        t = run(ticks, 1.0, t)
        occs.append(occupancy().copy())
    arr = eng.drain_arrivals()
    return occs, eng.agents(fields=FIELDS), arr


def sharded_vs_single(device, agents=200_000, nodes=20_000, ticks=100, blocks=3, p2p=True, seed=77, group=None,
                      labels=True, reroute=False):
    """Returns a dict ``{world, agents, nodes, links, ticks, exchange, ok, detail}``; ``ok`` is the AND over
    all ranks.  Needs an initialised ``torch.distributed`` process group with one CUDA device per rank."""
    import torch
    import torch.distributed as dist

    from .dist import ShardedEngine
    from .engine import Engine, ROUTER_BATCHED
    from .synth import road_graph

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = torch.device("cuda", device)
    lg = road_graph(nodes, seed=4)
    N = int(agents)
    rng = np.random.default_rng(3)
    start = rng.integers(0, lg.V, size=N).astype(np.int32)
    cands = rng.integers(0, lg.V, size=(blocks, N, 2)).astype(np.int32)
    first = rng.integers(0, lg.V, size=(N, 4)).astype(np.int32)
    hourly = np.full(24, 0.9)
    arena = N * 300
    hier = None
    if labels:  # route by hub labels (what bench.py does on road graphs); every rank contracts the same graph
        from . import hierarchy

        hier = hierarchy.build(lg)

    se = ShardedEngine(lg, N, device=device, route_arena_links=arena, group=group)
    if p2p:
        se.enable_p2p()
    lo, hi = se.lo, se.hi
    occ_s, ag_s, arr_s = _drive(se.run, se.engine, se.occupancy, lg, start[lo:hi], first[lo:hi], cands[:, lo:hi],
                                hourly, ticks, seed, ROUTER_BATCHED, hier,
                                reroute=se.update_travel_times if reroute else None)
    se.close()

    # the same run on one engine (rank 0), broadcast to everybody
    occ_1 = np.zeros((blocks, lg.L), dtype=np.int32)
    ag_1 = {"state": np.zeros(N, np.uint8), "dist": np.zeros(N), "speed": np.zeros(N),
            "route_idx": np.zeros(N, np.int32), "trip_count": np.zeros(N, np.int32)}
    n_arr_1 = np.zeros(1, dtype=np.int64)
    if rank == 0:
        eng = Engine(lg, N, device=device, route_arena_links=arena)

        def run1(n, delta, t):
            eng.run(n, delta, t)
            return t + n * delta

        o1, a1, arr1 = _drive(run1, eng, eng.occupancy, lg, start, first, cands, hourly, ticks, seed, ROUTER_BATCHED, hier,
                              reroute=eng.update_travel_times if reroute else None)
        eng.close()
        occ_1[:] = np.stack(o1)
        for k in FIELDS:
            ag_1[k][:] = a1[k]
        n_arr_1[0] = len(arr1[0])
    bufs = [occ_1, n_arr_1] + [ag_1[k] for k in FIELDS]
    for b in bufs:
        t = torch.from_numpy(b).to(dev)
        dist.broadcast(t, src=0, group=group)
        b[...] = t.cpu().numpy()

    bad = []
    for b in range(blocks):
        if not np.array_equal(occ_s[b], occ_1[b]):
            bad.append("link volumes after block %d" % b)
    for k in FIELDS:
        x, y = ag_s[k], ag_1[k][lo:hi]
        if k in ("dist", "speed"):  # only meaningful while moving (a waiting agent keeps stale values on both sides)
            mv = ag_1["state"][lo:hi] == 1
            x, y = x[mv], y[mv]
        if not np.array_equal(x, y):
            bad.append("%s of agents [%d, %d)" % (k, lo, hi))
    n_arr = torch.tensor([len(arr_s[0])], dtype=torch.int64, device=dev)
    dist.all_reduce(n_arr, group=group)
    if int(n_arr.item()) != int(n_arr_1[0]):
        bad.append("completed trips: %d sharded vs %d single" % (int(n_arr.item()), int(n_arr_1[0])))
    ok = torch.tensor([0 if bad else 1], dtype=torch.int64, device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    moved = int(occ_1[-1].sum())
    return dict(world=world, agents=N, nodes=lg.V, links=lg.L, ticks=ticks * blocks, blocks=blocks,
                exchange="p2p" if p2p else "allgather", router="hub labels" if labels else "trees", reroute=bool(reroute), completed_trips=int(n_arr_1[0]), on_roads_at_end=moved,
                compared=["link volumes at every block boundary"] + list(FIELDS) + ["completed trips"],
                ok=bool(ok.item()) and moved > 0, detail=bad)
