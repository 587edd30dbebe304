#!/usr/bin/env python
"""profiles/fft_only.py — repeated spectralPotentialSolver calls: the library's own shared-memory DFT against cuFFT
(for timing and ncu).  usage: fft_only.py [cells ...]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import physical_normalisation
from picsp_b200 import Params, Simulation
from picsp_b200.sim import FLAG_CUFFT_ONLY, FLAG_OWN_FFT
nm = physical_normalisation()
sizes = [int(a) for a in sys.argv[1:]] or [256, 512, 1024, 2048]
reps = 50
for numx in sizes:
    rng = np.random.default_rng(0)
    nix = numx + 1
    rho = np.zeros((nix, nix)); rho[1:-1, 1:-1] = rng.standard_normal((nix - 2, nix - 2))
    out = {}
    for name, flags in (("cufft", FLAG_CUFFT_ONLY), ("own", FLAG_OWN_FFT)):
        with Simulation(Params(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], 8, 8, solverType=1, flags=flags)) as sim:
            sim.set_grid("rho", rho)
            sim.spectralPotentialSolver(); sim.sync()
            sim.profile_enable(True); sim.profile_reset()
            for _ in range(reps):
                sim.spectralPotentialSolver()
            sim.sync()
            out[name] = sim.profile()["solve"][0] / reps
    print(f"spectral solve {nix}^2 nodes: cuFFT {out['cufft'] * 1e3:.1f} us, own {out['own'] * 1e3:.1f} us  ({out['cufft'] / out['own']:.2f}x)", flush=True)
