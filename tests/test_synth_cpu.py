"""Synthetic workload generators (CPU): the graphs bench.py builds for the large BASELINE.json configs
must be well-formed lowered graphs -- same invariants as a graph lowered from the reference loader."""
import numpy as np
import pytest

from drift_b200 import demand
from drift_b200.synth import grid_graph, rgg_graph, road_graph


def _check_lowered(lg):
    V, L = lg.V, lg.L
    assert lg.fwd_rowptr[0] == 0 and lg.fwd_rowptr[-1] == L and np.all(np.diff(lg.fwd_rowptr) >= 0)
    assert lg.bwd_rowptr[0] == 0 and lg.bwd_rowptr[-1] == L and np.all(np.diff(lg.bwd_rowptr) >= 0)
    # forward entry e of node u is link e (no parallel edges in synthetic graphs): tail u, head fwd_col[e]
    tails = np.repeat(np.arange(V), np.diff(lg.fwd_rowptr))
    assert np.array_equal(tails, lg.link_u[lg.fwd_link]) and np.array_equal(lg.fwd_col, lg.link_v[lg.fwd_link])
    heads = np.repeat(np.arange(V), np.diff(lg.bwd_rowptr))
    assert np.array_equal(heads, lg.link_v[lg.bwd_link]) and np.array_equal(lg.bwd_col, lg.link_u[lg.bwd_link])
    assert np.array_equal(np.sort(lg.bwd_link), np.arange(L))
    # travel_time = length / speed (lib/graph_loader.py:759-766), weights carried per CSR entry
    t0 = ((lg.link_len / 1000.0) / lg.link_v0) * 3600
    assert np.array_equal(lg.link_t0, t0) and np.array_equal(lg.fwd_w, t0[lg.fwd_link])
    assert np.array_equal(lg.bwd_w, t0[lg.bwd_link])
    assert lg.link_cap.min() >= 1 and lg.pos.shape == (V, 2)


def test_grid_graph_zorder_is_the_same_graph():
    a = grid_graph(40, 30, seed=3)
    b = grid_graph(40, 30, seed=3, morton=True)
    _check_lowered(a)
    _check_lowered(b)
    assert a.V == b.V == 1200 and a.L == b.L == 2 * (39 * 30 + 40 * 29)
    assert np.array_equal(np.sort(a.link_len), np.sort(b.link_len))
    # node p of the row-major grid is the node of b with the same position; links keep their lengths
    key = {(int(x), int(y)): i for i, (x, y) in enumerate(b.pos)}
    remap = np.array([key[(int(x), int(y))] for x, y in a.pos])
    la = {(int(remap[u]), int(remap[v])): w for u, v, w in zip(a.link_u, a.link_v, a.link_len)}
    lb = {(int(u), int(v)): w for u, v, w in zip(b.link_u, b.link_v, b.link_len)}
    assert la == lb


def test_rgg_graph_is_symmetric_and_connected():
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components

    g = rgg_graph(4000, seed=1)
    _check_lowered(g)
    assert 5.0 < g.L / g.V < 7.0 and abs(g.link_len.mean() - 120.0) < 1e-6
    pairs = set(zip(g.link_u.tolist(), g.link_v.tolist()))
    assert len(pairs) == g.L and all((v, u) in pairs for u, v in pairs)
    n, _ = connected_components(coo_matrix((np.ones(g.L), (g.link_u, g.link_v)), shape=(g.V, g.V)), directed=True,
                                connection="strong")
    assert n == 1


def test_road_graph_invariants():
    _check_lowered(road_graph(3000, seed=2))


@pytest.mark.parametrize("model", sorted(demand.MODELS))
def test_demand_chains_are_valid_targets(model):
    lg = road_graph(2000, seed=5)
    rng = np.random.default_rng(0)
    start = rng.integers(0, lg.V, size=500).astype(np.int32)
    off, dst = demand.MODELS[model](rng, lg, start, 6)
    assert off[0] == 0 and off[-1] == len(dst) == 500 * 6 and np.all(np.diff(off) == 6)
    assert dst.min() >= 0 and dst.max() < lg.V
    if model in ("random", "hub"):  # these models redraw a target equal to the trip's source (st_selection.py:47-49)
        d = dst.reshape(500, 6)
        assert np.all(d[:, 0] != start) and np.all(d[:, 1:] != d[:, :-1])


def test_activity_law_matches_reference_shares():
    """The vectorised activity demand follows the reference's trip law per agent type
    (lib/st_selection.py:926-1050): share of targets drawn from the type's pool."""
    lg = road_graph(6000, seed=3)
    start = np.random.default_rng(1).integers(0, lg.V, size=40000).astype(np.int32)
    rng = np.random.default_rng(9)
    P = demand._activity_pools(np.random.default_rng(9), lg)  # same first draws as inside activity_chains
    _, dst = demand.activity_chains(rng, lg, start, 6)
    d = dst.reshape(-1, 6)
    pools = np.unique(np.concatenate(list(P.values())))
    share = np.isin(d, pools).mean()
    # .4 commuter (always a pool) + .3 * .6 + .2 * .8 + .1 * .7 = 0.81, plus arbitrary nodes that happen to be pool nodes
    assert 0.80 < share < 0.83
    cold_cold = (~np.isin(d[:, 1:], pools) & ~np.isin(d[:, :-1], pools)).mean()
    assert 0.05 < cold_cold < 0.08
    _, dp = demand.activity_pools_chains(np.random.default_rng(9), lg, start, 6)
    assert np.isin(dp.reshape(-1, 6)[:, 1::2], pools).all()  # return legs always end in a pool


def test_hub_and_zone_laws():
    """Vectorised hub / zone demand against the reference's rules (lib/st_selection.py:551-606, 1162-1231)."""
    lg = road_graph(8000, seed=2)
    start = np.random.default_rng(1).integers(0, lg.V, size=30000).astype(np.int32)
    _, d = demand.hub_chains(np.random.default_rng(3), lg, start, 6)
    d = d.reshape(-1, 6)
    deg = np.diff(lg.fwd_rowptr) + np.diff(lg.bwd_rowptr)
    is_hub = np.zeros(lg.V, dtype=bool)
    is_hub[np.argsort(-deg, kind="stable")[:int(lg.V * 0.1)]] = True
    src = np.concatenate([start[:, None], d[:, :-1]], axis=1)
    assert not np.any(src == d)
    # a trip ending at a hub from a non-hub source is a "hub trip" (30 %); non-hub trips (70 %) end at non-hubs
    from_non = ~is_hub[src]
    assert abs(is_hub[d][from_non].mean() - 0.3) < 0.02
    # from a hub: 30 % anywhere (10 % of the nodes are hubs), 70 % non-hub
    assert abs(is_hub[d][~from_non].mean() - 0.3 * 0.1) < 0.02

    _, z = demand.zone_chains(np.random.default_rng(3), lg, start, 6)
    z = z.reshape(-1, 6)
    srcz = np.concatenate([start[:, None], z[:, :-1]], axis=1)
    assert not np.any(srcz == z)
    lo, hi = lg.pos.min(axis=0), lg.pos.max(axis=0)
    cell = np.minimum(((lg.pos - lo) / (hi - lo) * 3).astype(np.int64), 2)
    zone = cell[:, 0] * 3 + cell[:, 1]
    intra = zone[z] == zone[srcz]
    assert abs(intra.mean() - 0.7) < 0.02
    # inter-zone trips prefer the destination zone's major nodes (top third by degree): 70 % + a third of the rest
    major = np.zeros(lg.V, dtype=bool)
    for zz in range(9):
        members = np.nonzero(zone == zz)[0]
        top = members[np.lexsort((members, -deg[members]))][:max(1, len(members) // 3)]
        major[top] = True
    assert abs(major[z][~intra].mean() - (0.7 + 0.3 / 3)) < 0.03
