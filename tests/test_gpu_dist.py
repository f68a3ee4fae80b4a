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
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, out):
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
            assert np.array_equal(se.engine.occupancy(), occ_ref[k - 1]), "rank %d: volumes differ at tick %d" % (rank, k)
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


@pytest.mark.parametrize("name", ["g40_congested", "g100_random_accel50"])
def test_two_rank_replay_matches_reference(name):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def _worker_demand(rank, world, port, out, p2p=False):
    """Device-driven demand sharded over 2 ranks == the same run on one engine (bit for bit)."""
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from drift_b200.dist import ShardedEngine
        from drift_b200.engine import Engine, ROUTER_BATCHED
        from drift_b200.synth import road_graph

        lg = road_graph(5000, seed=4)
        N, blocks, ticks = 40_000, 4, 60
        rng = np.random.default_rng(3)
        start = rng.integers(0, lg.V, size=N).astype(np.int32)
        cands = rng.integers(0, lg.V, size=(blocks, N, 2)).astype(np.int32)
        first = rng.integers(0, lg.V, size=(N, 4)).astype(np.int32)
        hourly = np.full(24, 0.9)

        def drive(eng_like, eng, lo, hi):
            eng.set_demand(77, hourly, start[lo:hi], chain_capacity=4, router=ROUTER_BATCHED)
            eng.set_option("route_window", ticks)
            eng.push_trips(first[lo:hi])
            t, occs = 0.0, []
            for b in range(blocks):
                eng.push_trips(cands[b, lo:hi])
                if eng_like is eng:
                    eng.run(ticks, 1.0, t)
                    t += ticks
                else:
                    t = eng_like.run(ticks, 1.0, t)
                occs.append((eng if eng_like is eng else eng_like).occupancy().copy())
            return occs, eng.agents(fields=("state", "dist", "speed", "route_idx", "trip_count"))

        se = ShardedEngine(lg, N, device=rank, route_arena_links=N * 300)
        if p2p:
            se.enable_p2p()
        occ_s, ag_s = drive(se, se.engine, se.lo, se.hi)
        se.close()
        if rank == 0:
            eng = Engine(lg, N, device=0, route_arena_links=N * 300)
            occ_1, ag_1 = drive(eng, eng, 0, N)
            eng.close()
            for a, b in zip(occ_s, occ_1):
                assert np.array_equal(a, b), "sharded volumes differ from the single-engine run"
            assert occ_1[-1].sum() > 1000
            for k in ag_s:
                assert np.array_equal(ag_s[k], ag_1[k][se.lo:se.hi]), k
        dist.barrier()
        out.put((rank, "ok"))
    except Exception:
        import traceback

        out.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("p2p", [False, True])
def test_two_rank_device_demand_equals_single_engine(p2p):
    """p2p=False: per-tick all-gather of the link events (NCCL); p2p=True: the exchange inside the
    persistent tick kernel over NVLink peer memory (k_tick_loop_p2p, link-owner resolution)."""
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_demand, args=(r, 2, port, out, p2p)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
