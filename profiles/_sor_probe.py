import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import physical_normalisation
from picsp_b200 import Params, Simulation
nm = physical_normalisation()
numx = 512
rng = np.random.default_rng(1)
rho = np.zeros((numx + 1, numx + 1)); rho[1:-1, 1:-1] = rng.standard_normal((numx - 1, numx - 1))
with Simulation(Params(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], 8, 8, solverType=2)) as sim:
    sim.set_grid("rho", rho)
    for _ in range(5):
        sim.solvePotential()
    print(sim.last_sweeps, sim.last_l2)
