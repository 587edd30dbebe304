#!/usr/bin/env python
"""profiles/sor_only.py — repeated solvePotential calls on a BASELINE-size grid (for ncu / timing of the SOR kernels)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import physical_normalisation
from picsp_b200 import Params, Simulation
nm = physical_normalisation()
numx = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
rng = np.random.default_rng(0)
nix = numx + 1
rho = np.zeros((nix, nix)); rho[1:-1, 1:-1] = rng.standard_normal((nix - 2, nix - 2))
with Simulation(Params(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], 8, 8, solverType=2)) as sim:
    sim.set_grid("rho", rho)
    sim.solve(); sim.sync()
    sim.profile_enable(True); sim.profile_reset()
    for _ in range(reps):
        sim.solve()
    sim.sync()
    ms = sim.profile()["solve"][0] / reps
    print(f"SOR {nix}^2: {ms * 1e3:.1f} us per solve, {ms * 1e6 / (2 * nix - 1):.0f} ns per anti-diagonal")
