"""GPU: a species split into several stores ("parts", picsp_params::parts) that share ONE spare buffer set — the layout
that lets BASELINE config 5 (4e9 particles) run on a single 180 GB device.  All parts deposit into the species' one
integer accumulator grid with one fixed-point scale, so every result must be BIT-IDENTICAL to the one-part run
(the kinetic energy is summed part by part: equal to rounding)."""
import numpy as np
import pytest

from oracle.oracle import ELECTRON, ION, Oracle, normalise
from picsp_b200 import Params, Simulation
from picsp_b200.lib import PicspError
from tests.helpers import GRIDS, RTOL, assert_grid_close, relerr

pytestmark = pytest.mark.gpu


def run(parts, solver, numx, n, flags=0, upload=None, steps=(9, 8)):
    nm = normalise()
    out = {}
    with Simulation(Params(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n, solverType=solver, flags=flags, parts=parts)) as sim:
        assert sim.parts() == max(parts, 1)
        sim.set_sort_period(ION, 7); sim.set_sort_period(ELECTRON, 4)
        if upload is None:
            sim.fill_synthetic(ION, n, seed=41, vth=nm["vth_i"])
            sim.fill_synthetic(ELECTRON, n, seed=42, vth=1.0, xdrift=nm["drift_e"])
        else:
            for s in (ION, ELECTRON):
                sim.set_species(s, *upload[s])
        sim.bootstrap()
        for k in steps:
            sim.step(k)
        out.update({g: sim.grid(g) for g in GRIDS})
        out["pi"] = np.stack(sim.get_species(ION)); out["pe"] = np.stack(sim.get_species(ELECTRON))
        out["rows_e"] = sim.get_species_rows(ELECTRON)
        out["ke"] = np.array([sim.computeKE(ION), sim.computeKE(ELECTRON)])
        out["count"] = np.array([sim.count(ION), sim.count(ELECTRON)])
        out["repush"] = np.array([sim.repush_count(ELECTRON)])
    return out


@pytest.mark.parametrize("solver,numx,n,flags", [(1, 128, 300_001, 0), (2, 80, 150_000, 0), (1, 96, 200_000, 16), (1, 256, 1_500_001, 0)],
                         ids=["spectral", "sor", "separate-sort", "two-pass-first-binning-own-fft"])
def test_parts_do_not_change_the_result(solver, numx, n, flags):
    """1, 3 and 4 parts (uneven split: the last part is short), device loader, several re-binnings of both species
    through the shared spare (re-binning mover, or the stand-alone re-sort with flag 16)."""
    ref = run(1, solver, numx, n, flags)
    for parts in (3, 4):
        got = run(parts, solver, numx, n, flags)
        for k in ref:
            if k == "ke":
                assert relerr(got[k], ref[k]) <= 1e-13, (parts, k)
            else:
                assert np.array_equal(got[k], ref[k]), f"parts={parts}: {k} differs from the one-part run"


def test_parts_uploaded_load_matches_the_oracle():
    """Host upload split over 5 parts, downloads (arrays and rows) in list order, three steps against the oracle."""
    nm = normalise()
    numx, n = 128, 250_003
    o = Oracle(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n, vth_i=nm["vth_i"], solver=1)
    o.seed(3); o.init(ION, 1); o.init(ELECTRON, 1)
    with Simulation(Params(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n, solverType=1, parts=5)) as sim:
        sim.set_sort_period(ION, 2); sim.set_sort_period(ELECTRON, 2)
        for s in (ION, ELECTRON):
            sim.set_species(s, *o.get_species(s))
        for s in (ION, ELECTRON):
            got, want = sim.get_species(s), o.get_species(s)
            for k in range(4):
                assert np.array_equal(got[k], want[k]), "download after the first binning differs from the upload"
            rows = sim.get_species_rows(s)
            assert np.array_equal(rows, np.stack(want, axis=1))
        o.bootstrap(); sim.bootstrap()
        for st in range(3):
            o.step(1); sim.step(1)
            for name in GRIDS:
                assert_grid_close(sim.grid(name), o.grid(name), sim.nix, sim.niy, 10 * RTOL, f"step{st}/{name}")
            for s in (ION, ELECTRON):
                got, want = sim.get_species(s), o.get_species(s)
                for k in range(4):
                    assert relerr(got[k], want[k]) <= 10 * RTOL
            for s in (ION, ELECTRON):
                assert relerr(np.array([sim.computeKE(s)]), np.array([o.computeKE(s)])) <= 1e-12


def test_parts_dump_equals_the_downloads():
    """picsp_dump_begin / _wait with several parts: rows of every part land at their place in list order."""
    nm = normalise()
    numx, n = 96, 120_000
    with Simulation(Params(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n, solverType=1, parts=3)) as sim:
        sim.fill_synthetic(ION, n, seed=1, vth=nm["vth_i"]); sim.fill_synthetic(ELECTRON, n, seed=2, vth=1.0, xdrift=nm["drift_e"])
        sim.bootstrap(); sim.step(5)
        d = sim.dump()
        for s, key in ((ION, "rows_i"), (ELECTRON, "rows_e")):
            assert np.array_equal(d[key], sim.get_species_rows(s))
        assert np.array_equal(d["den_i"], sim.grid("den_i")) and np.array_equal(d["phi"], sim.grid("phi"))
        assert relerr(d["ke"], np.array([sim.computeKE(ION), sim.computeKE(ELECTRON)])) <= 1e-13


def test_parts_need_the_tiled_store():
    nm = normalise()
    with pytest.raises(PicspError):
        Simulation(Params(32, 32, nm["dx"], nm["dt"], nm["mass_i"], 1000, 1000, flags=2, parts=2))
