"""Row-layout dump (picsp_species_download_rows) against the four-array download, 1e8 particles, pinned host buffers."""
import sys, time; sys.path.insert(0, '.')
import ctypes as C
import torch
from oracle.oracle import normalise
from picsp_b200 import Params, Simulation, ION, ELECTRON
from picsp_b200.lib import check
nm = normalise(); n = 100_000_000
sim = Simulation(Params(1024, 1024, nm["dx"], nm["dt"], nm["mass_i"], n, n, solverType=1, capacity=(n, n)))
sim.fill_synthetic(ION, n, seed=1, vth=nm["vth_i"]); sim.fill_synthetic(ELECTRON, n, seed=2, vth=1.0, xdrift=nm["drift_e"])
sim.bootstrap(); sim.step(2); sim.sync()
dp = C.POINTER(C.c_double)
rows = torch.empty(4 * n, dtype=torch.float64, pin_memory=True)
cols = [torch.empty(n, dtype=torch.float64, pin_memory=True) for _ in range(4)]
ptr = lambda t: C.cast(t.data_ptr(), dp)
for rep in range(2):
    t0 = time.perf_counter(); check(sim.L.picsp_species_download_rows(sim.ctx, ELECTRON, ptr(rows))); t1 = time.perf_counter()
    check(sim.L.picsp_species_download(sim.ctx, ELECTRON, *(ptr(t) for t in cols))); t2 = time.perf_counter()
    print("rows %.3f s (%.1f GB/s)   arrays %.3f s (%.1f GB/s)" % (t1 - t0, 32e-9 * n / (t1 - t0), t2 - t1, 32e-9 * n / (t2 - t1)))
print("equal:", bool((rows.view(n, 4)[:, 2] == cols[2]).all()))
