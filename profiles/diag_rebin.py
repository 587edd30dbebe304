import sys; sys.path.insert(0, '.')
import numpy as np
from oracle.oracle import ELECTRON, ION, Oracle, normalise
from picsp_b200 import Params, Simulation
from picsp_b200.sim import FLAG_SEPARATE_SORT
nm = normalise()
for numx, numy, period in [(32, 48, 1), (64, 64, 1)]:
    n = 40000
    o = Oracle(numx, numy, nm["dx"], nm["dt"], nm["mass_i"], n, n, vth_i=nm["vth_i"], solver=1)
    o.seed(33); o.init(ION, 1); o.init(ELECTRON, 1)
    x, y, vx, vy = o.get_species(ELECTRON)
    o.set_species(ELECTRON, x, y, vx * 2.5, vy * 2.5)
    for nsteps in (1, 2, 3, 9):
        runs = []
        for flags in (0, FLAG_SEPARATE_SORT):
            with Simulation(Params(numx, numy, nm["dx"], nm["dt"], nm["mass_i"], n, n, solverType=1, flags=flags)) as sim:
                sim.set_sort_period(ION, period); sim.set_sort_period(ELECTRON, period)
                for s in (ION, ELECTRON):
                    sim.set_species(s, *o.get_species(s))
                sim.bootstrap(); sim.step(nsteps)
                runs.append((np.stack(sim.get_species(ELECTRON)), sim.grid("den_e"), sim.straggler_count(ELECTRON), sim.repush_count(ELECTRON)))
        a, b = runs
        d = a[0] != b[0]
        rel = np.abs(a[0] - b[0]) / np.maximum(np.abs(b[0]), 1e-300)
        print(numx, numy, "steps", nsteps, "differing per row", d.sum(axis=1), "max rel", rel.max(), "den equal", np.array_equal(a[1], b[1]),
              "stragglers", a[2], b[2], "repush", a[3], b[3])
