"""Vectorised s-t selectors (drift_b200/fast_selection.py) against the reference's classes:
same seed -> the same (source, target) sequence, bit for bit.

* fixtures: ``tests/golden_selectors/selectors.npz`` was recorded from the unmodified reference classes by
  ``oracle/make_golden_selectors.py`` (graphs are rebuilt deterministically by that module);
* live: where the reference tree is present (build container) the two are also run side by side,
  including their public state (hubs, zones, major nodes, size / attraction variables).
"""
import os
import random
import sys

import numpy as np
import pytest

from drift_b200 import fast_selection
from oracle import make_golden_selectors as mg

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_selectors", "selectors.npz")
FAST = {"hub": fast_selection.HubAndSpokeSelection, "zone": fast_selection.ZoneBasedSelection,
        "gravity": fast_selection.GravitySelection, "activity": fast_selection.ActivityBasedSelection}
_graphs = {}


def _graph(kind):
    if kind not in _graphs:
        _graphs[kind] = mg.make_graph(kind)
    return _graphs[kind]


@pytest.mark.parametrize("name", sorted(mg.CASES))
def test_sequences_match_recorded_reference(name):
    kind, model, kwargs, seed, agents, calls = mg.CASES[name]
    g = _graph(kind)
    want = np.load(GOLDEN)[name]
    if model in mg.SEED_BEFORE_BUILD:  # the activity model draws its node pools in the constructor
        random.seed(seed)
    sel = FAST[model](g, **kwargs)
    if model not in mg.SEED_BEFORE_BUILD:
        random.seed(seed)
    seq = mg.ask(sel, agents, calls)
    idx = {n: i for i, n in enumerate(g.nodes())}
    got = np.array([(idx[s], idx[t]) for s, t in seq], dtype=np.int32)
    assert np.array_equal(got, want)
    # the draws consumed are the reference's too: the generator is left in the same state
    assert all(s != t for s, t in seq) or model in ("gravity", "activity")  # commuters may stay (lib/st_selection.py:941-954)


def _have_reference():
    return os.path.isdir(os.path.join(mg.REF, "lib"))


@pytest.mark.skipif(not _have_reference(), reason="reference tree not present")
@pytest.mark.parametrize("name", ["hub_small", "hub_large", "zone_small", "zone_sparse", "gravity_hops_beta",
                                  "activity_small", "activity_large", "activity_tiny"])
def test_live_against_reference_classes(name):
    kind, model, kwargs, seed, agents, calls = mg.CASES[name]
    g = _graph(kind)
    random.seed(seed + 7)
    ref = mg.reference_selector(model, g, kwargs)
    built_ref = random.getstate()
    random.seed(seed + 7)
    fast = FAST[model](g, **kwargs)
    assert random.getstate() == built_ref  # the constructors consume the same draws
    if model == "activity":
        assert ref.get_activity_nodes() == fast.get_activity_nodes()
        assert ref.get_agent_types_distribution() == fast.get_agent_types_distribution()
    elif model == "hub":
        assert ref.hubs == fast.hubs and ref.non_hubs == fast.non_hubs and ref.node_centrality == fast.node_centrality
    elif model == "zone":
        assert ref.zones == fast.zones and ref.node_to_zone == fast.node_to_zone
        assert ref.major_nodes_by_zone == fast.major_nodes_by_zone and ref.get_zone_info() == fast.get_zone_info()
    else:
        assert ref.size_variables == fast.size_variables and ref.attraction_variables == fast.attraction_variables
    random.seed(seed + 100)
    a = mg.ask(ref, agents, 1500)
    state_ref = random.getstate()
    random.seed(seed + 100)
    b = mg.ask(fast, agents, 1500)
    assert a == b and random.getstate() == state_ref
    if model == "hub":
        assert ref.last_trip_type == fast.last_trip_type


def test_gravity_row_cache_and_parameters():
    g = _graph("geo300")
    sel = fast_selection.GravitySelection(g, row_cache=8)
    random.seed(3)
    first = mg.ask(sel, 20, 400)
    assert len(sel._rows) <= 8
    random.seed(3)
    assert mg.ask(sel, 20, 400) == first  # evicted rows are recomputed identically
    sel.set_parameters(beta=3.0)
    assert not sel._rows and sel.get_model_info()["beta"] == 3.0
    random.seed(3)
    assert mg.ask(sel, 20, 400) != first


def test_host_pow_is_cpython_pow():
    rng = np.random.default_rng(0)
    x = rng.uniform(0.01, 1e6, size=20000)
    for y in (0.5, 2.0, 1.5, 2.5):
        assert np.array_equal(fast_selection._host_pow(x, y), np.array([v ** y for v in x.tolist()]))


def test_make_selector_fast_switch():
    from drift_b200.st_selection import make_selector

    g = _graph("geo300")
    st = {"hub_trip_probability": 0.4, "hub_percentage": 0.2, "zone_intra_probability": 0.6,
          "gravity_alpha": 1.5, "gravity_beta": 1.8}
    hub = make_selector("hub", g, st, fast=True)
    assert type(hub) is fast_selection.HubAndSpokeSelection and hub.__class__.__name__ == "HubAndSpokeSelection"
    assert hub.hub_trip_probability == 0.4 and len(hub.hubs) == 60
    zone = make_selector("zone", g, st, fast=True)
    assert zone.__class__.__name__ == "ZoneBasedSelection" and zone.intra_zone_probability == 0.6  # lib/agent.py:57,61
    grav = make_selector("gravity", g, st, fast=True)
    assert type(grav) is fast_selection.GravitySelection and grav.alpha == 1.5 and grav.beta == 1.8
    act = make_selector("activity", g, {"agent_type_distributions": {"commuter": 1.0}}, fast=True)
    assert type(act) is fast_selection.ActivityBasedSelection and act.get_agent_types_distribution() == {"commuter": 1.0}
    assert list(fast_selection.ActivityBasedSelection(g).get_agent_types_distribution()) == \
        ["commuter", "leisure", "business", "delivery"]  # key order = config.py:326-331 (random.choices order)
    # models without a vectorised restatement fall through to the reference / built-in ones
    assert make_selector("random", g, None, fast=True).__class__.__name__ == "RandomSelection"
