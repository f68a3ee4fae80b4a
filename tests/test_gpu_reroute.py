"""Extension: periodic congestion-aware rerouting (no reference behaviour -- the reference never
updates travel_time).  The oracle is composed from reference pieces in
oracle/port.py (BPR factor -> link travel time, NetworkX-order bidirectional Dijkstra, splice after
the current link); the CUDA path must match it bit for bit."""
import os
import random

import numpy as np
import pytest

from oracle import golden
from oracle.port import PortSim, RandomSelector

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,period", [("g40_congested", 60), ("g40_congested_bpr", 25), ("multi_random_s0", 120)])
def test_periodic_reroute_matches_composed_oracle(name, period, tmp_path):
    from drift_b200.simulation_thread import SimulationThread

    case = golden.GoldenCase(name)
    G = case.nx_graph()
    pg = case.plain_graph()
    N, ticks, seed = 1200, 700, 5
    hourly = {h: 0.9 for h in range(24)}
    # oracle
    random.seed(seed)
    sim = PortSim(pg, N, selector=RandomSelector(pg.nodes), hourly=hourly, alpha=case.alpha, beta=case.beta,
                  accel=10, hours=1)
    occ_ref, last, n_rer = [], 0.0, 0
    tt_ref = None
    for _ in range(ticks):
        sim.tick()
        if sim.t - last >= period:
            sim.update_travel_times()
            n_rer += sim.reroute_moving()
            tt_ref = np.array(sim.tt)
            last = sim.t
        occ_ref.append(np.asarray(sim.occ, dtype=np.int32).copy())
    sim.finish()
    # engine through the drop-in thread
    random.seed(seed)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        th = SimulationThread(G, N, "random", duration_hours=1)
        th.apply_settings(dict(bpr_alpha=case.alpha, bpr_beta=case.beta, hourly_probabilities=hourly))
        th.reroute_period = period
        th.running = True
        assert th.initialize_simulation()
        th.update_frequency = 10 ** 9
        for k in range(ticks):
            assert th.update_simulation_step()
            assert np.array_equal(th.engine.occupancy(), occ_ref[k]), "volumes differ at tick %d" % (k + 1)
        assert th.reroutes == n_rer and n_rer > 50
        assert np.array_equal(th.engine.travel_times(), tt_ref)  # fp64 ==: same table lookup, same division
        assert not np.array_equal(tt_ref, case.link_tt)          # congestion did change the weights
        th.running = True
        th.finish_simulation()
        got = th.data_logger.trips_data
        assert len(got) == len(sim.trips)
        id2idx = {"A%06d" % id(a): a.index for a in th.agents}
        for g, w in zip(got, sim.trips):
            g = dict(g)
            g["agent_index"] = id2idx[g.pop("agent_id")]
            assert g == w
        th.close()
    finally:
        os.chdir(cwd)
