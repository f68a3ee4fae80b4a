"""World-size-2 gloo tests (CPU) of the multi-rank host logic: shard bounds, the variable-length
event exchange, and the rank-prefix semantics of the sharded tick using the oracle's event model."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker_exchange(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from drift_b200.dist import exchange_events, shard_bounds

    try:
        assert shard_bounds(10, 0, 3) == (0, 4) and shard_bounds(10, 1, 3) == (4, 7) and shard_bounds(10, 2, 3) == (7, 10)
        for trial, n_local in enumerate([3 + 2 * rank, 0 if rank == 0 else 5, 0]):
            local = torch.full((16, 4), -1, dtype=torch.int32)
            for i in range(n_local):
                local[i] = torch.tensor([100 * rank + i, rank, 1, trial], dtype=torch.int32)
            ev, counts = exchange_events(local, n_local)
            want = []
            for r in range(world):
                n_r = [3 + 2 * r, 0 if r == 0 else 5, 0][trial]
                want += [[100 * r + i, r, 1, trial] for i in range(n_r)]
            assert counts == [[3 + 2 * r, 0 if r == 0 else 5, 0][trial] for r in range(world)]
            assert ev.tolist() == want
        out.put((rank, "ok"))
    except Exception as exc:  # pragma: no cover
        out.put((rank, repr(exc)))
    finally:
        dist.destroy_process_group()


def test_event_exchange_gloo_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_exchange, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")]


def _worker_sharded_port(rank, world, port, out):
    """Each rank advances its own agent block with the oracle's event model and resolves the
    all-gathered events; the result must equal the sequential single-process port."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from drift_b200.dist import exchange_events, shard_bounds
        from oracle import golden
        from oracle.event_model import EventModel
        from oracle.port import PortSim

        case = golden.GoldenCase("g40_congested_bpr")
        pg = case.plain_graph()
        trace = case.departure_trace()
        ticks = 300
        lo, hi = shard_bounds(case.N, rank, world)
        em = EventModel(pg, case.N, lo, hi, alpha=case.alpha, beta=case.beta)
        ref = PortSim(pg, case.N, hourly=case.hourly, alpha=case.alpha, beta=case.beta, accel=case.accel,
                      hours=case.hours, departure_trace=trace, init_nodes=case.init_nodes)
        for k in range(1, ticks + 1):
            deps = {a: v for a, v in trace.get(k, {}).items() if lo <= a < hi}
            local = em.tick_begin(1.0, deps)
            buf = torch.zeros((max(len(local), 1), 4), dtype=torch.int32)
            if len(local):
                buf[:len(local)] = torch.from_numpy(local)
            ev, counts = exchange_events(buf, len(local))
            em.tick_finish(ev.numpy())
            ref.tick()
            assert np.array_equal(em.occ, np.asarray(ref.occ)), "volumes differ at tick %d" % k
        st, d, sp, pi = ref.snapshot()
        mv = st[lo:hi] == 1
        assert np.array_equal(em.state, st[lo:hi])
        assert np.array_equal(em.dist[mv], d[lo:hi][mv]) and np.array_equal(em.speed[mv], sp[lo:hi][mv])
        out.put((rank, "ok"))
    except Exception as exc:  # pragma: no cover
        import traceback

        out.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_tick_semantics_gloo(world):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_sharded_port, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, "ok") for r in range(world)], res


def test_home_rank_arithmetic_matches_shard_bounds():
    """k_tick_loop_p2p finds the home rank of a global agent id with two divisions (p2p_home in
    csrc/tick_kernels.cuh); it must agree with the contiguous blocks ShardedEngine hands out."""
    from drift_b200.dist import shard_bounds

    def home(agent, n_total, world):  # transliteration of p2p_home
        base, rem = n_total // world, n_total % world
        big = (base + 1) * rem
        if agent < big:
            h = agent // (base + 1)
            return h, agent - h * (base + 1)
        r = agent - big
        q = r // base
        return rem + q, r - q * base

    for n_total, world in [(10, 3), (1000, 8), (1001, 8), (7, 7), (40_000, 2), (999_999, 16), (16, 16)]:
        owner = {}
        for r in range(world):
            lo, hi = shard_bounds(n_total, r, world)
            for a in (lo, hi - 1, (lo + hi) // 2):
                if lo <= a < hi:
                    owner[a] = (r, a - lo)
        for a, want in owner.items():
            assert home(a, n_total, world) == want, (n_total, world, a)


def _worker_link_owner(rank, world, port, out):
    """The link-owner protocol of the peer-memory exchange, host-level: every rank sends each event to the
    owner of its link (link % world), the owner orders the events of its links and returns the entry
    speeds to the agents' home ranks; owned volumes summed over the ranks and all agent state must
    equal the sequential port."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from drift_b200.dist import shard_bounds
        from oracle import golden
        from oracle.event_model import EventModel
        from oracle.port import PortSim

        case = golden.GoldenCase("g40_congested_bpr")
        pg = case.plain_graph()
        trace = case.departure_trace()
        ticks = 300
        lo, hi = shard_bounds(case.N, rank, world)
        bounds = [shard_bounds(case.N, r, world) for r in range(world)]
        em = EventModel(pg, case.N, lo, hi, alpha=case.alpha, beta=case.beta)
        ref = PortSim(pg, case.N, hourly=case.hourly, alpha=case.alpha, beta=case.beta, accel=case.accel,
                      hours=case.hours, departure_trace=trace, init_nodes=case.init_nodes)
        for k in range(1, ticks + 1):
            deps = {a: v for a, v in trace.get(k, {}).items() if lo <= a < hi}
            local = em.tick_begin(1.0, deps)
            # exchange 1: events to the owners of their links
            to_owner = [local[local[:, 0] % world == o].tolist() if len(local) else [] for o in range(world)]
            # (gloo has no all-to-all for objects: gather everything, keep what is mine)
            gathered = [None] * world
            dist.all_gather_object(gathered, to_owner)
            mine = [e for src in gathered for e in src[rank]]
            results = em.resolve_owned(np.array(mine, dtype=np.int64).reshape(-1, 4), rank, world)
            # exchange 2: entry speeds to the home ranks
            to_home = [[(a, sp) for a, sp in results if bounds[h][0] <= a < bounds[h][1]] for h in range(world)]
            gathered = [None] * world
            dist.all_gather_object(gathered, to_home)
            em.apply_speeds([x for src in gathered for x in src[rank]])
            ref.tick()
            occ = torch.from_numpy(em.occ.copy())
            owned = torch.from_numpy((np.arange(pg.L) % world == rank).astype(np.int64))
            assert bool(torch.all(occ * (1 - owned) == 0)), "a rank holds volume of a link it does not own"
            dist.all_reduce(occ)
            assert np.array_equal(occ.numpy(), np.asarray(ref.occ)), "volumes differ at tick %d" % k
        st, d, sp, pi = ref.snapshot()
        mv = st[lo:hi] == 1
        assert np.array_equal(em.state, st[lo:hi])
        assert np.array_equal(em.dist[mv], d[lo:hi][mv]) and np.array_equal(em.speed[mv], sp[lo:hi][mv])
        out.put((rank, "ok"))
    except Exception:  # pragma: no cover
        import traceback

        out.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_link_owner_protocol_gloo(world):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_link_owner, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, "ok") for r in range(world)], res
