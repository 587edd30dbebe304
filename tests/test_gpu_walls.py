"""GPU: the PICSP_FLAG_WALLS extension (BASELINE.json config 3: bounded domain, wall boundaries, Gauss-Seidel solver).

NO REFERENCE ORACLE: the reference is periodic-only (absorbing walls and a Dirichlet solver are commented-out
sketches, main.cpp:826-843, :1064-1108).  The checker here is the repository's own CPU restatement
(oracle/walls_check.c, pinned to textbook facts by tests/test_walls_checker_cpu.py)."""
import numpy as np
import pytest

import picsp_b200
from oracle.oracle import ELECTRON, ION, WallsChecker, normalise
from picsp_b200 import Params, Simulation
from picsp_b200.sim import FLAG_CLEAR_DENSITY, FLAG_NO_SORT, FLAG_WALLS
from tests.helpers import GRIDS, relerr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("numx,numy", [(64, 64), (200, 120), (512, 512)])
def test_red_black_sor_is_bit_identical_to_the_checker_and_converged(numx, numy):
    nm = normalise()
    rng = np.random.default_rng(numx)
    nix, niy = numx + 1, numy + 1
    rho = np.zeros((nix, niy)); rho[1:-1, 1:-1] = rng.standard_normal((nix - 2, niy - 2))
    w = WallsChecker(numx, numy, nm["dx"], nm["dt"], nm["mass_i"], 8, 8)
    with Simulation(Params(numx, numy, nm["dx"], nm["dt"], nm["mass_i"], 8, 8, solverType=2, flags=FLAG_WALLS)) as sim:
        sim.set_grid("rho", rho)
        sim.set_grid("phi", rng.standard_normal(nix * niy))          # a dirty warm start, walls included
        w.phi[:] = sim.grid("phi")
        sim.solve()
        sweeps, l2 = sim.solve_status()
        assert sweeps > 0 and l2 < WallsChecker.TOL, (sweeps, l2)
        w.rho[:] = rho.reshape(-1)
        w.solve(fixed_sweeps=sweeps)
        phi = sim.grid("phi")
        assert np.array_equal(phi, w.phi), f"phi after {sweeps} sweeps differs from the checker: {relerr(phi, w.phi):.2e}"
        assert abs(w.last_l2 - l2) <= 1e-6 * l2
        g = phi.reshape(nix, niy)
        assert np.all(g[0] == 0) and np.all(g[-1] == 0) and np.all(g[:, 0] == 0) and np.all(g[:, -1] == 0)
        # the checker's own stopping decision lands on the same batch
        w2 = WallsChecker(numx, numy, nm["dx"], nm["dt"], nm["mass_i"], 8, 8)
        w2.rho[:] = rho.reshape(-1); w2.phi[:] = w.phi * 0 + sim.grid("phi") * 0
        sim.computeEF(); w.computeEF()
        assert np.array_equal(sim.grid("efx"), w.efx) and np.array_equal(sim.grid("efy"), w.efy)
        print(f"walls SOR {nix}x{niy}: {sweeps} sweeps, L2 {l2:.2e}")


@pytest.mark.parametrize("clear", [True, False], ids=["clear-density", "accumulate"])
def test_bounded_plasma_steps_match_the_checker(clear):
    """bootstrap + 4 steps of a thermal plasma with a hot electron tail that reaches the walls: densities, fields and
    the phase space (absorbed particles = NaN rows, same particles on both sides) against the checker; the checker is
    given the sweep count the GPU's residual test arrived at, so the comparison is iterate against same iterate."""
    nm = normalise()
    numx, n = 96, 200_000
    rng = np.random.default_rng(7)
    xl = numx * nm["dx"]
    xe, ye = rng.random(n) * xl, rng.random(n) * xl
    vxe, vye = rng.standard_normal(n), rng.standard_normal(n)
    hot = rng.random(n) < 0.05
    vxe[hot] *= 8; vye[hot] *= 8                            # ~2.7 cells per step per sigma (max ~12 < 16): many absorbed in 4 steps
    xi, yi = rng.random(n) * xl, rng.random(n) * xl
    vi = 0.02 * rng.standard_normal(n)
    flags = FLAG_WALLS | (FLAG_CLEAR_DENSITY if clear else 0)
    w = WallsChecker(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n, clear=clear)
    w.set_species(ION, xi, yi, vi, vi); w.set_species(ELECTRON, xe, ye, vxe, vye)
    with Simulation(Params(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n, solverType=2, flags=flags)) as sim:
        sim.set_species(ION, xi, yi, vi, vi); sim.set_species(ELECTRON, xe, ye, vxe, vye)
        absorbed_total = 0
        for st in range(5):
            for s in (ION, ELECTRON):
                sim.scatterSpecies(s); w.scatterSpecies(s)
            sim.computeRho(); w.computeRho()
            sim.solve(); sweeps, l2 = sim.solve_status()
            assert sweeps > 0 and l2 < WallsChecker.TOL
            w.solve(fixed_sweeps=sweeps)
            sim.computeEF(); w.computeEF()
            for s in (ION, ELECTRON):
                if st == 0:
                    sim.rewindSpecies(s); w.rewindSpecies(s)
                else:
                    sim.pushSpecies(s); w.pushSpecies(s)
                    assert sim.repush_count(s) == w.absorbed[s], (st, s, sim.repush_count(s), w.absorbed[s])
                    absorbed_total += w.absorbed[s]
            worst = 0.0
            for name in GRIDS:
                e = relerr(sim.grid(name), w.grid(name))
                assert e <= (1e-12 if name.startswith("den") else 1e-10), f"step{st}/{name}: {e:.2e}"
                worst = max(worst, e)
            for s in (ION, ELECTRON):
                for a, b in zip(sim.get_species(s), w.get_species(s)):
                    e = relerr(a, b)                         # also asserts that the NaN (absorbed) pattern is the same
                    assert e <= 1e-11, (st, s, e)
                    worst = max(worst, e)
            print(f"walls step {st}: {sweeps} sweeps, worst rel err vs the repo's CPU checker {worst:.2e}, absorbed so far {absorbed_total}")
        assert absorbed_total > 1000
        ke = sim.computeKE(ELECTRON)
        _, _, vx, vy = w.get_species(ELECTRON)
        assert abs(ke - 0.5 * sim.p.spwt[1] - np.sum(vx * vx + vy * vy)) <= 1e-10 * ke      # absorbed particles carry no energy


def test_fused_step_equals_the_per_function_sequence_and_rebinning_keeps_absorbed_particles_inert():
    nm = normalise()
    numx, n = 64, 120_000
    flags = FLAG_WALLS | FLAG_CLEAR_DENSITY
    runs = []
    for fused in (False, True):
        with Simulation(Params(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n, solverType=2, flags=flags)) as sim:
            sim.set_sort_period(ELECTRON, 3); sim.set_sort_period(ION, 5)
            sim.fill_synthetic(ION, n, seed=1, vth=nm["vth_i"])
            sim.fill_synthetic(ELECTRON, n, seed=2, vth=6.0, xdrift=nm["drift_e"])
            sim.bootstrap()
            if fused:
                sim.step(9)
            else:
                for _ in range(9):
                    sim.scatterSpecies(ION); sim.scatterSpecies(ELECTRON); sim.computeRho(); sim.solve(); sim.computeEF()
                    sim.pushSpecies(ION); sim.pushSpecies(ELECTRON)
            runs.append({g: sim.grid(g) for g in GRIDS} | {"pe": np.stack(sim.get_species(ELECTRON)), "pi": np.stack(sim.get_species(ION))})
    a, b = runs
    for k in a:
        assert np.array_equal(a[k], b[k], equal_nan=True), k
    dead = np.isnan(a["pe"][0])
    assert 1000 < dead.sum() < n and np.array_equal(dead, np.isnan(a["pe"][1])) and np.all(a["pe"][2][dead] == 0)


def test_walls_needs_the_tiled_store():
    nm = normalise()
    with pytest.raises(picsp_b200.PicspError) as ei:
        Simulation(Params(32, 32, nm["dx"], nm["dt"], nm["mass_i"], 10, 10, flags=FLAG_WALLS | FLAG_NO_SORT))
    assert ei.value.code == -1
