"""2-rank NCCL run of the sharded tick (needs >= 2 GPUs; run with `gpurun --gpus 2`): every rank
replays its block of a recorded reference run; volumes must equal the single-process reference at
every tick and fp64 agent state must match."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, out, p2p=False):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from drift_b200.dist import ShardedEngine
        from drift_b200.graph import lower_graph
        from oracle import golden

        case = golden.GoldenCase(name)
        lg = lower_graph(case.nx_graph())
        occ_ref = case.occupancy()
        snaps = case.snapshots()
        trace = case.departure_trace()
        se = ShardedEngine(lg, case.N, alpha=case.alpha, beta=case.beta, device=rank)
        if p2p:  # the exchange inside k_tick_loop_p2p: link volumes live on the owner, occupancy() sums them
            se.enable_p2p()
        lo, hi = se.lo, se.hi
        delta = 0.1 * case.accel
        ticks = min(case.ticks, 1200)
        arr = []
        for k in range(1, ticks + 1):
            deps = {a: v for a, v in trace.get(k, {}).items() if lo <= a < hi}
            if deps:
                agents = sorted(deps)
                st, nl, links = se.engine.route([deps[a][0] for a in agents], [deps[a][1] for a in agents])
                for a, ln in zip(agents, links):
                    if deps[a][2] is not None and len(deps[a][2]) > 1:
                        assert lg.links_to_nodes(deps[a][0], ln) == deps[a][2]
                se.engine.commit_departures([a - lo for a in agents])
            se.step(delta)
            assert np.array_equal(se.occupancy(), occ_ref[k - 1]), "rank %d: volumes differ at tick %d" % (rank, k)
            if k in snaps:
                a = se.engine.agents()
                st, d, sp, pi = snaps[k]
                st, d, sp, pi = st[lo:hi], d[lo:hi], sp[lo:hi], pi[lo:hi]
                mv = st == 1
                assert np.array_equal(a["state"], st)
                assert np.array_equal(a["dist"][mv], d[mv]) and np.array_equal(a["speed"][mv], sp[mv])
                aa, at = se.engine.drain_arrivals()
                arr.extend(zip(at.tolist(), (aa + lo).tolist()))
        aa, at = se.engine.drain_arrivals()
        arr.extend(zip(at.tolist(), (aa + lo).tolist()))
        want = [(t, a) for t, a in case.arrivals() if t <= ticks and lo <= a < hi]
        assert sorted(arr) == want
        se.close()
        out.put((rank, "ok"))
    except Exception:
        import traceback

        out.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def _spawn(target, world, *args):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    ctx = mp.get_context("spawn")
    res = []
    for attempt in range(2):  # a rendezvous port can be taken between _free_port() and the store's bind: one retry
        out = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=target, args=(r, world, port) + args + (out,)) for r in range(world)]
        for p in procs:
            p.start()
        res = []
        try:
            for _ in procs:
                res.append(out.get(timeout=600))
                if res[-1][1] != "ok":  # fail fast: the other ranks would wait for this one until their timeout
                    break
        finally:
            done = len(res) == world and all(r[1] == "ok" for r in res)
            for p in procs:
                if not done and p.is_alive():
                    p.terminate()
                p.join(timeout=60)
        if done or not any("EADDRINUSE" in str(r[1]) or "address already in use" in str(r[1]) for r in res):
            break
    assert sorted(res) == [(r, "ok") for r in range(world)], res


def _replay(rank, world, port, name, p2p, out):
    _worker(rank, world, port, name, out, p2p)


@pytest.mark.parametrize("world,p2p", [(2, False), (2, True), (3, True), (4, True), (8, True)])
@pytest.mark.parametrize("name", ["g40_congested", "g100_random_accel50"])
def test_sharded_replay_matches_reference(name, world, p2p):
    """Recorded reference runs replayed over `world` ranks: link volumes at every tick, fp64 agent state
    at the snapshots and the arrival ticks equal the single-process reference -- with the per-tick
    all-gather (p2p=False) and with the peer-memory exchange inside the tick kernel (p2p=True)."""
    _spawn(_replay, world, name, p2p)


def _selfcheck(rank, world, port, p2p, agents, reroute, out):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from drift_b200.selfcheck import sharded_vs_single

        r = sharded_vs_single(rank, agents=agents, nodes=5000, ticks=60, blocks=4, p2p=p2p, reroute=reroute)
        assert r["ok"], r
        assert r["completed_trips"] > 100 and r["on_roads_at_end"] > 1000, r
        out.put((rank, "ok"))
    except Exception:
        import traceback

        out.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,p2p", [(2, False), (2, True), (3, True), (4, True), (8, True)])
def test_sharded_device_demand_equals_single_engine(world, p2p):
    """Device-driven demand sharded over `world` ranks == the same run on one engine, bit for bit (link
    volumes at every block boundary, fp64 positions / speeds, route indices, trip counts, completed trips).
    p2p=False: per-tick all-gather of the link events (NCCL); p2p=True: the exchange inside the persistent
    tick kernel over NVLink peer memory (k_tick_loop_p2p, link-owner resolution) -- what bench.py uses."""
    _spawn(_selfcheck, world, p2p, 40_000 + world, False)  # + world: uneven shards


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_reroute_equals_single_engine(world):
    """The extension on sharded agents: every rank recounts the volumes of its agents, the dense int32[L] vector is
    all-reduced (NCCL), every rank computes the same BPR travel times and re-plans its own en-route agents -- the run
    stays bit-identical to one engine doing the same on all agents."""
    _spawn(_selfcheck, world, True, 30_000 + world, True)
