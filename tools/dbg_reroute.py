import os, random, sys
sys.path.insert(0, '/root/repo')
import numpy as np
from oracle import golden
from oracle.port import PortSim, RandomSelector
from drift_b200.simulation_thread import SimulationThread
case = golden.GoldenCase("g40_congested"); G = case.nx_graph(); pg = case.plain_graph()
N, seed, period = 1200, 5, 60
hourly = {h: 0.9 for h in range(24)}
random.seed(seed)
sim = PortSim(pg, N, selector=RandomSelector(pg.nodes), hourly=hourly, alpha=case.alpha, beta=case.beta, accel=10, hours=1)
for _ in range(60): sim.tick()
paths_before = {a.idx: list(a.path) for a in sim.agents if a.state == 1}
pidx = {a.idx: a.pidx for a in sim.agents if a.state == 1}
sim.update_travel_times(); sim.reroute_moving()
random.seed(seed)
os.chdir('/tmp')
th = SimulationThread(G, N, "random", duration_hours=1)
th.apply_settings(dict(bpr_alpha=case.alpha, bpr_beta=case.beta, hourly_probabilities=hourly))
th.running = True; th.initialize_simulation(); th.update_frequency = 10**9
for k in range(60): th.update_simulation_step()
assert np.array_equal(th.engine.occupancy(), np.asarray(sim.occ))
th._reroute(th.simulation_time)
tt = th.engine.travel_times()
print('tt equal', np.array_equal(tt, np.array(sim.tt)))
idx = th.lowered.node_index
bad = 0
for a in sim.agents:
    if a.state != 1: continue
    b = th.agents[a.idx]
    hp = [idx[n] for n in b.path]
    if hp != a.path:
        bad += 1
        if bad < 4:
            print('agent', a.idx, 'pidx', pidx[a.idx], 'host pidx', b.path_index)
            print(' before', paths_before[a.idx]); print(' oracle', a.path); print(' device', hp)
            links = th.engine.agent_route(a.idx).tolist(); print(' dev links->nodes', th.lowered.links_to_nodes(th.lowered.link_u[links[0]], links))
print('mismatching agents', bad, 'of', sum(1 for a in sim.agents if a.state == 1))
