"""Wall-clock per step of the small BASELINE configurations with and without the CUDA-graph replay of step pairs
(no profiling: the per-phase events would disable the graph path).  Run under gpurun."""
import sys, time; sys.path.insert(0, '.')
from bench import physical_normalisation
from picsp_b200 import ELECTRON, ION, Params, Simulation
from picsp_b200.sim import FLAG_NO_GRAPH
nm = physical_normalisation()
for name, cells, n, solver, steps in (("config1 64^2 1e4+1e4 SOR", 64, 10_000, 2, 2000), ("config2 256^2 6.55e6/species spectral", 256, 6_553_600, 1, 400),
                                      ("256^2 1e6/species spectral", 256, 1_000_000, 1, 1000)):
    for rep in range(2):
        for flags, tag in ((FLAG_NO_GRAPH, "plain"), (0, "graph")):
            with Simulation(Params(cells, cells, nm["dx"], nm["dt"], nm["mass_i"], n, n, solverType=solver, flags=flags)) as sim:
                sim.fill_synthetic(ION, n, seed=1, vth=nm["vth_i"]); sim.fill_synthetic(ELECTRON, n, seed=2, vth=1.0, xdrift=nm["drift_e"])
                sim.bootstrap(); sim.step(9); sim.sync()
                t0 = time.perf_counter(); sim.step(steps); sim.sync(); dt = time.perf_counter() - t0
                print(f"{name:40s} {tag:5s} {1e3 * dt / steps:8.4f} ms/step  {2 * n * steps / dt:.3e} particle-steps/s", flush=True)
