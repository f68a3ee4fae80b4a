"""Throughput of the drop-in SimulationThread (exact mode: host MT19937 draws, device step) at the
reference's own scale, next to the oracle port on the same host."""
import os, random, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import golden
from oracle.port import PortSim, RandomSelector
from drift_b200.headless import run_headless

for name, agents, hours in (("g100_random_s0", 300, 1), ("liz_random_nopath", 10_000, 0.25), ("g40_congested", 2500, 1)):
    case = golden.GoldenCase(name)
    G = case.nx_graph(); pg = case.plain_graph()
    tmp = tempfile.mkdtemp()
    th, wall = run_headless(G, agents, "random", hours=hours, accel=10, seed=0, out_dir=tmp)
    n = agents * th.step
    th.close()
    random.seed(0)
    sim = PortSim(pg, agents, selector=RandomSelector(pg.nodes), accel=10, hours=hours)
    t0 = time.perf_counter()
    while sim.tick():
        pass
    wp = time.perf_counter() - t0
    print("%-20s agents=%5d ticks=%d  drop-in %.3g agent-steps/s (%.1f us/tick)   port %.3g agent-steps/s" % (
        name, agents, th.step, n / wall, 1e6 * wall / th.step, agents * sim.step / wp))
