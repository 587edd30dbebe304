#!/usr/bin/env python
"""profiles/straggler_curve.py — how fast do electrons leave their bin's window, and what does it cost the mover?
Never re-sorts after the first sort; prints per step the straggler fraction and the mover time (run under gpurun)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import physical_normalisation  # noqa: E402
from picsp_b200 import ELECTRON, ION, Params, Simulation  # noqa: E402

nm = physical_normalisation()
cells, n = 1024, 100_000_000
with Simulation(Params(cells, cells, nm["dx"], nm["dt"], nm["mass_i"], n, n, solverType=1)) as sim:
    sim.set_sort_period(ELECTRON, 10_000); sim.set_sort_period(ION, 10_000)
    sim.fill_synthetic(ION, n, seed=1, vth=nm["vth_i"])
    sim.fill_synthetic(ELECTRON, n, seed=2, vth=1.0, xdrift=nm["drift_e"])
    sim.bootstrap()
    sim.profile_enable(True)
    print("step  electron stragglers  mover ms (both species)")
    for st in range(1, 33):
        sim.profile_reset()
        sim.step(1)
        ms = sim.profile()["push"][0]
        print(f"{st:4d}  {sim.straggler_count(ELECTRON) / n:8.4%}            {ms:7.3f}", flush=True)
