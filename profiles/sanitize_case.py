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
# round 2: cell order inside a bin + warp-aggregated deposit (forced on), a clustered load (REDUX groups, two
# interleaved groups), the asynchronous dump, the walls extension (cooperative red-black SOR, absorption)
for numx, numy, n, cell, agg in ((48, 40, 30001, 2, -1), (64, 64, 40000, 0, 1)):
    with Simulation(Params(numx, numy, nm["dx"], nm["dt"], nm["mass_i"], n, n, solverType=1)) as sim:
        sim.set_sort_period(ELECTRON, 3); sim.set_cell_sort_period(ION, cell); sim.set_cell_sort_period(ELECTRON, cell)
        sim.set_deposit_aggregation(ION, agg); sim.set_deposit_aggregation(ELECTRON, agg)
        x = rng.random(n) * numx * nm["dx"]; y = rng.random(n) * numy * nm["dx"]
        x[: n // 2] = (7 + (np.arange(n // 2) & 1) + rng.random(n // 2)) * nm["dx"]; y[: n // 2] = (9 + rng.random(n // 2)) * nm["dx"]
        sim.set_species(ION, x, y, 0 * x, 0 * x); sim.set_species(ELECTRON, x, y, rng.standard_normal(n), rng.standard_normal(n))
        sim.bootstrap(); sim.step(6)
        d = sim.dump()
        print("r2", numx, cell, agg, float(d["ke"][1]), float(d["rows_e"][:, 0].max()), sim.computeKE(ION))
with Simulation(Params(40, 56, nm["dx"], nm["dt"], nm["mass_i"], 20000, 20000, solverType=2, flags=64 | 1)) as sim:
    sim.fill_synthetic(ION, 20000, seed=5, vth=nm["vth_i"]); sim.fill_synthetic(ELECTRON, 20000, seed=6, vth=5.0)
    sim.bootstrap(); sim.step(5)
    print("walls", sim.solve_status(), sim.repush_count(ELECTRON), int(np.isnan(sim.get_species(ELECTRON)[0]).sum()))
# round 2, second session: bank order inside the chunks for both species (k_bank_order, after every re-binning), a
# species split into 3 parts sharing one spare (uneven last part; upload, re-binning mover, stand-alone re-sort, dumps),
# the library's own DFT (prime-factor split 3 x 11 / plain Bluestein 49 / even length 48 / two different plans)
for numx, numy, n, parts, flags in ((48, 40, 30001, 1, 0), (64, 56, 20011, 3, 0), (40, 40, 15001, 3, 16)):
    with Simulation(Params(numx, numy, nm["dx"], nm["dt"], nm["mass_i"], n, n, solverType=1, flags=flags, parts=parts)) as sim:
        sim.set_sort_period(ELECTRON, 2); sim.set_sort_period(ION, 3)
        sim.set_bank_order(ION, 1); sim.set_bank_order(ELECTRON, 1)
        for s, vth in ((ION, nm["vth_i"]), (ELECTRON, 2.0)):
            sim.set_species(s, rng.random(n) * numx * nm["dx"], rng.random(n) * numy * nm["dx"],
                            vth * rng.standard_normal(n), vth * rng.standard_normal(n))
        sim.bootstrap(); sim.step(7)
        d = sim.dump()
        print("r2b", numx, parts, flags, sim.parts(), float(d["ke"][1]), float(sim.get_species_rows(ELECTRON)[:, 0].max()), sim.computeKE(ION))
for numx, numy in ((32, 32), (48, 48), (47, 47), (64, 130)):
    with Simulation(Params(numx, numy, nm["dx"], nm["dt"], nm["mass_i"], 5000, 5000, solverType=1, flags=512)) as sim:
        sim.fill_synthetic(ION, 5000, seed=7, vth=nm["vth_i"]); sim.fill_synthetic(ELECTRON, 5000, seed=8, vth=1.0)
        sim.bootstrap(); sim.step(2)
        print("fft", numx, numy, sim.spectral_engine(), sim.delta_phi())
# far movers (more than one tile per step): global gather / deposit, individual slots at the re-binning, the far-mover counter
n = 20000
x = rng.random(n) * 128 * nm["dx"]; y = rng.random(n) * 128 * nm["dx"]
vx = 0.3 * rng.standard_normal(n); vy = 0.3 * rng.standard_normal(n)
vx[:7] = 40 * nm["dx"] / nm["dt"]; vy[7:12] = -70 * nm["dx"] / nm["dt"]
with Simulation(Params(128, 128, nm["dx"], nm["dt"], nm["mass_i"], n, n, solverType=1)) as sim:
    sim.set_sort_period(ELECTRON, 2)
    sim.set_species(ION, x, y, 0 * x, 0 * x); sim.set_species(ELECTRON, x, y, vx, vy)
    sim.bootstrap(); sim.step(5); sim.sync()
    print("far", sim.computeKE(ELECTRON), sim.straggler_count(ELECTRON))
