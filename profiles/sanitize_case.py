import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import physical_normalisation
from picsp_b200 import ELECTRON, ION, Params, Simulation
nm = physical_normalisation()
for solver, numx, numy, n in ((1, 48, 40, 30001), (2, 70, 33, 20011)):
    with Simulation(Params(numx, numy, nm["dx"], nm["dt"], nm["mass_i"], n, n, solverType=solver)) as sim:
        sim.set_sort_period(ELECTRON, 2)
        sim.fill_synthetic(ION, n, seed=1, vth=nm["vth_i"])
        sim.fill_synthetic(ELECTRON, n, seed=2, vth=2.0, xdrift=nm["drift_e"])
        sim.bootstrap(); sim.step(5)
        x, y, vx, vy = sim.get_species(ELECTRON)
        print(solver, sim.computeKE(ELECTRON), sim.delta_phi(), float(x.max()), sim.straggler_count(ELECTRON))
# re-binning on consecutive steps (k_tile_mover<3>), upload/download through host buffers (asynchronous first
# binning, un-permute overlapped with the copies), a grid with fewer than 3 bins per side, the stand-alone re-sort
rng = np.random.default_rng(4)
for numx, numy, n, period, flags in ((40, 72, 20003, 1, 0), (32, 20, 9001, 3, 0), (64, 48, 15000, 2, 16)):
    with Simulation(Params(numx, numy, nm["dx"], nm["dt"], nm["mass_i"], n, n, solverType=1, flags=flags)) as sim:
        sim.set_sort_period(ELECTRON, period); sim.set_sort_period(ION, period)
        for s, vth in ((ION, nm["vth_i"]), (ELECTRON, 2.5)):
            sim.set_species(s, rng.random(n) * numx * nm["dx"], rng.random(n) * numy * nm["dx"],
                            vth * rng.standard_normal(n), vth * rng.standard_normal(n))
        sim.bootstrap(); sim.step(7)
        x, y, vx, vy = sim.get_species(ELECTRON)
        print(numx, numy, period, flags, sim.computeKE(ELECTRON), float(x.max()), sim.straggler_count(ELECTRON), sim.repush_count(ELECTRON))
