#!/usr/bin/env python
"""profiles/configs_bench.py — per-phase device timings of BASELINE.json's other configurations
(they are parity-test cases, not bench lines; this records where their time goes).  Run under gpurun."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import physical_normalisation  # noqa: E402
from picsp_b200 import ELECTRON, ION, Params, Simulation  # noqa: E402

nm = physical_normalisation()
CONFIGS = [
    ("config1: shipped input.ini shape, 64^2 cells, 1e4+1e4, SOR", 64, 10_000, 2, 200),
    ("config2: 256^2 cells, 100 ppc/species, spectral", 256, 6_553_600, 1, 50),
    ("config3: 512^2 cells, 200 ppc/species, periodic SOR", 512, 52_428_800, 2, 20),
    ("config4: 1024^2 cells, 5e8/species, spectral", 1024, 500_000_000, 1, 10),
]
out = []
for name, cells, n, solver, steps in CONFIGS:
    if len(sys.argv) > 1 and not any(a in name for a in sys.argv[1:]):
        continue
    with Simulation(Params(cells, cells, nm["dx"], nm["dt"], nm["mass_i"], n, n, solverType=solver)) as sim:
        sim.fill_synthetic(ION, n, seed=1, vth=nm["vth_i"])
        sim.fill_synthetic(ELECTRON, n, seed=2, vth=1.0, xdrift=nm["drift_e"])
        sim.bootstrap()
        sim.step(3); sim.sync()
        sim.profile_enable(True); sim.profile_reset()
        l0 = sim.kernel_launches()
        sim.step(steps); sim.sync()
        prof = sim.profile()
        ms = prof["step"][0] / steps
        rec = {"config": name, "ms_per_step": ms, "particle_steps_per_s": 2 * n / (ms * 1e-3),
               "phases_ms_per_step": {k: v[0] / steps for k, v in prof.items()},
               "launches_per_step": (sim.kernel_launches() - l0) / steps}
        print(json.dumps(rec), flush=True)
        out.append(rec)
json.dump(out, open("gpurun_out/configs_bench.json", "w"), indent=1)
