import sys; sys.path.insert(0,'.')
import numpy as np
from oracle.oracle import ELECTRON, ION, Oracle, normalise
from picsp_b200 import Params, Simulation
from tests.helpers import GRIDS, relerr, interior
nm = normalise()
numx, numy, n = 130, 33, 30000
for flags in (0, 8, 2):
    o = Oracle(numx, numy, nm["dx"], nm["dt"], nm["mass_i"], n, n, vth_i=nm["vth_i"], solver=2)
    o.seed(5); o.init(ION, 1); o.init(ELECTRON, 1)
    with Simulation(Params(numx, numy, nm["dx"], nm["dt"], nm["mass_i"], n, n, solverType=2, flags=flags)) as sim:
        for s in (ION, ELECTRON):
            sim.set_species(s, *o.get_species(s))
        o.bootstrap(); sim.bootstrap()
        print("flags", flags, "boot", {g: "%.2e" % relerr(interior(sim.grid(g), sim.nix, sim.niy), interior(o.grid(g), sim.nix, sim.niy)) for g in GRIDS})
        for st in range(3):
            o.step(1); sim.step(1)
            print("  step", st, {g: "%.2e" % relerr(interior(sim.grid(g), sim.nix, sim.niy), interior(o.grid(g), sim.nix, sim.niy)) for g in GRIDS},
                  ["%.1e" % relerr(a, b) for a, b in zip(sim.get_species(ELECTRON), o.get_species(ELECTRON))], sim.last_sweeps if hasattr(sim,'last_sweeps') else '')
