"""Where the time of the FIRST step after an upload goes (initial binning of an arbitrary load + stand-alone deposit)."""
import sys, time; sys.path.insert(0, '.')
import ctypes as C
import numpy as np, torch
from oracle.oracle import normalise
from picsp_b200 import Params, Simulation, ION, ELECTRON
from picsp_b200.lib import check
nm = normalise()
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 500_000_000
sim = Simulation(Params(1024, 1024, nm["dx"], nm["dt"], nm["mass_i"], n, n, solverType=1, capacity=(n, n)))
sim.fill_synthetic(ION, n, seed=1, vth=nm["vth_i"]); sim.fill_synthetic(ELECTRON, n, seed=2, vth=1.0, xdrift=nm["drift_e"])
sim.bootstrap(); sim.step(3); sim.sync()
dp = C.POINTER(C.c_double)
host = [[torch.empty(n, dtype=torch.float64, pin_memory=True) for _ in range(4)] for _ in range(2)]
ptr = lambda t: C.cast(t.data_ptr(), dp)
for s in range(2):
    t0 = time.perf_counter(); check(sim.L.picsp_species_download(sim.ctx, s, *(ptr(t) for t in host[s]))); print("download", s, time.perf_counter() - t0)
for s in range(2):
    t0 = time.perf_counter(); check(sim.L.picsp_species_upload(sim.ctx, s, *(ptr(t) for t in host[s]), n)); print("upload", s, time.perf_counter() - t0)
t0 = time.perf_counter(); sim.sync(); print("first binning still running after the last upload returned: ms", (time.perf_counter() - t0) * 1e3)
sim.profile_enable(True)
for k in range(3):
    sim.profile_reset(); t0 = time.perf_counter(); sim.step(1); sim.sync(); dt = time.perf_counter() - t0
    print("step", k, "wall ms", dt * 1e3, {a: round(b[0], 3) for a, b in sim.profile().items()})
for s in range(2):
    t0 = time.perf_counter(); ke = sim.computeKE(s); print("KE", s, time.perf_counter() - t0)
