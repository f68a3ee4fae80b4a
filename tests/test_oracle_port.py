"""The sequential port (oracle/port.py) against the fixtures recorded from the live reference."""
import random

import numpy as np
import pytest

from oracle import golden
from oracle.port import PortSim, RandomSelector, bpr_factor, link_capacity

CASES = golden.case_names()


def _run(case, sim):
    occ_ref = case.occupancy()
    snaps = case.snapshots()
    k = 0
    while sim.tick():
        if case.params.get("max_ticks") and sim.step > case.params["max_ticks"]:
            break
        assert sim.t == case.times[k], (k, sim.t)
        assert np.array_equal(np.asarray(sim.occ, dtype=np.int32), occ_ref[k]), "occupancy differs at tick %d" % (k + 1)
        if sim.step in snaps:
            st, d, sp, pi = sim.snapshot()
            rs = snaps[sim.step]
            assert np.array_equal(st, rs[0])
            mv = st == 1
            assert np.array_equal(d[mv], rs[1][mv]) and np.array_equal(sp[mv], rs[2][mv])
            assert np.array_equal(pi[mv], rs[3][mv])
        k += 1
        if case.params.get("max_ticks") and sim.step >= case.params["max_ticks"]:
            break
    assert sim.step == case.ticks
    assert sim.entries == case.entries()
    assert sim.arrivals == case.arrivals()


def _check_trips(case, sim):
    if case.finish_error:
        with pytest.raises(TypeError):
            sim.finish()
        assert len(sim.trips) >= case.trips_before_finish
        got = sim.trips[:case.trips_before_finish]
        want = case.trips[:case.trips_before_finish]
    else:
        sim.finish()
        got, want = sim.trips, case.trips
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert a == b


@pytest.mark.parametrize("name", CASES)
def test_port_replay_matches_reference(name):
    case = golden.GoldenCase(name)
    pg = case.plain_graph()
    assert np.array_equal(pg.link_cap, case.capacities)
    sim = PortSim(pg, case.N, hourly=case.hourly, alpha=case.alpha, beta=case.beta, accel=case.accel,
                  hours=case.hours, departure_trace=case.departure_trace(), init_nodes=case.init_nodes,
                  agent_types=case.agent_types0, record_entries=True)
    _run(case, sim)
    _check_trips(case, sim)


@pytest.mark.parametrize("name", [n for n in CASES if "_random_" in n or "congested" in n])
def test_port_live_rng_matches_reference(name):
    """Own MT19937 stream + own router: same departures, same (s,t), same NX paths."""
    case = golden.GoldenCase(name)
    pg = case.plain_graph()
    random.seed(case.params["seed"])
    np.random.seed(case.params["seed"])
    sim = PortSim(pg, case.N, selector=RandomSelector(pg.nodes), hourly=case.hourly, alpha=case.alpha,
                  beta=case.beta, accel=case.accel, hours=case.hours, record_entries=True)
    assert np.array_equal(np.array([a.source for a in sim.agents]), case.init_nodes)
    _run(case, sim)
    assert sim.departures == case.departures()
    _check_trips(case, sim)


def test_capacity_formula_matches_recorded_capacities():
    for name in CASES:
        case = golden.GoldenCase(name)
        caps = [link_capacity(float(l), float(v), float(n)) for l, v, n in
                zip(case.link_lanes, case.link_v0, case.link_len)]
        assert np.array_equal(np.array(caps), case.capacities)


def test_bpr_known_answers():
    # SURVEY.md section 0-6 (probe P9): count = cap = 532 on a default link
    f = bpr_factor(532, 532, 0.15, 4.0)
    speed = (30.0 * f) * (1000 / 3600)
    assert speed == 7.246376811594204
    assert 1000 / speed == 137.99999999999997
    assert bpr_factor(0, 532, 0.15, 4.0) == 1.0
    assert bpr_factor(10 ** 6, 532, 0.15, 4.0) == 0.1
    d, k = 0.0, 0
    while d < 1000.0:
        d += speed * 1.0
        k += 1
    assert k == 138 and d == 1000.0000000000027
