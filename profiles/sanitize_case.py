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
