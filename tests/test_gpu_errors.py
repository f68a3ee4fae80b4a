"""Error behaviour of the C ABI: bad arguments and call-sequence errors come back as negative codes
with a message (no exceptions, no crashes), ring overflows are reported."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _engine(n=16, **kw):
    from drift_b200.engine import Engine
    from drift_b200.synth import grid_graph

    return Engine(grid_graph(6, 6, seed=1), n, **kw)


def test_bad_arguments_and_sequence_errors():
    from drift_b200 import _capi
    from drift_b200._capi import DriftError

    lib = _capi.load()
    with _engine() as eng:
        with pytest.raises(DriftError) as ei:
            eng.commit_departures([3])  # nothing staged
        assert ei.value.code == -1
        with pytest.raises(DriftError):
            eng.submit_departures([99], [[0, 1]])  # agent out of range
        with pytest.raises(DriftError):
            eng.submit_departures([1], [[10 ** 6]])  # link id out of range
        with pytest.raises(DriftError) as ei:
            eng.run(5, 1.0, 0.0)  # no demand configured
        assert ei.value.code == -5 and "drift_set_demand" in str(ei.value)
        with pytest.raises(DriftError):
            eng.set_option("no_such_option", 1)
        with pytest.raises(DriftError):
            eng.device_ptr("nope")
        assert lib.drift_step(None, 1, C.c_double(1.0)) == -1
        assert b"drift_step" in lib.drift_last_error()
        # out-of-range nodes are "no path", like NodeNotFound in the reference (lib/agent.py:87-99)
        st, nl, links = eng.route([0, -1, 5, 7], [35, 3, 10 ** 6, 7])
        assert st.tolist() == [0, 2, 2, 1] and nl[0] > 0 and nl[1] == nl[2] == nl[3] == 0
        eng.step(3, 1.0)  # still usable
        assert eng.occupancy().sum() == 0


def test_arrival_ring_overflow_is_reported():
    from drift_b200._capi import DriftError
    from drift_b200.engine import ROUTER_BATCHED
    from drift_b200.synth import grid_graph

    from drift_b200.engine import Engine

    lg = grid_graph(4, 4, seed=2, len_lo=20.0, len_hi=30.0)  # 2-second links: trips of a few ticks
    N = 3000
    rng = np.random.default_rng(0)
    with Engine(lg, N) as eng:
        eng.set_demand(1, np.full(24, 900.0), rng.integers(0, lg.V, N).astype(np.int32), chain_capacity=8,
                       router=ROUTER_BATCHED)
        eng.set_option("route_window", 50)
        eng.push_trips(rng.integers(0, lg.V, size=(N, 8)).astype(np.int32))
        eng.run(200, 1.0, 0.0)  # every agent makes several trips: more arrivals than the ring holds
        with pytest.raises(DriftError) as ei:
            eng.drain_arrivals()
        assert ei.value.code == -4 and ("arrivals" in str(ei.value) or "ring" in str(ei.value))
