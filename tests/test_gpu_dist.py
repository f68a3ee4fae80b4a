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
