"""CPU: the repository's own checker of the PICSP_FLAG_WALLS extension (oracle/walls_check.c — NOT the reference, which
has no bounded-domain semantics) is pinned to textbook facts, since there is nothing else to pin it to: the Dirichlet
solve reproduces the discrete eigenfunction solution, walls stay at phi = 0, the deposit conserves charge without a
fold, and particles that leave the box are absorbed for good."""
import numpy as np

from oracle.oracle import ELECTRON, ION, WallsChecker, normalise


def test_dirichlet_solve_matches_the_discrete_eigenfunction():
    numx, dx = 48, 0.05
    w = WallsChecker(numx, numx, dx, 0.01, 1836.0, 10, 10)
    i = np.arange(numx + 1)
    sx = np.sin(np.pi * i / numx)
    mode = np.outer(sx, sx)
    # 5-point Laplacian eigenvalue of sin(pi i/N) sin(pi j/N): (4/dx^2) * 2 * sin^2(pi/(2N))
    lam = (4.0 / dx ** 2) * 2.0 * np.sin(np.pi / (2 * numx)) ** 2
    w.rho[:] = (lam * mode).reshape(-1)
    w.rho.reshape(numx + 1, -1)[[0, -1], :] = 0; w.rho.reshape(numx + 1, -1)[:, [0, -1]] = 0
    w.solve()
    phi = w.phi.reshape(numx + 1, numx + 1)
    assert w.last_sweeps > 0 and w.last_l2 < 1e-12
    assert np.abs(phi - mode).max() < 1e-8
    assert np.all(phi[0] == 0) and np.all(phi[-1] == 0) and np.all(phi[:, 0] == 0) and np.all(phi[:, -1] == 0)
    w.computeEF()
    ex = w.efx.reshape(numx + 1, -1)
    assert abs(ex[0, numx // 2] + phi[1, numx // 2] / dx) < 1e-12          # full one-sided difference on the wall


def test_deposit_without_fold_conserves_charge_and_absorption_is_final():
    nm = normalise()
    numx, n = 32, 20000
    rng = np.random.default_rng(1)
    xl = numx * nm["dx"]
    w = WallsChecker(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n)
    x, y = rng.random(n) * xl, rng.random(n) * xl
    vx = rng.standard_normal(n) * 30; vy = rng.standard_normal(n) * 30      # fast: many reach a wall
    w.set_species(ELECTRON, x, y, vx, vy); w.set_species(ION, x, y, 0 * vx, 0 * vy)
    w.scatterSpecies(ELECTRON)
    assert abs(w.den[1].sum() - n * w.spwt[1] / nm["dx"] ** 2) < 1e-9 * n * w.spwt[1] / nm["dx"] ** 2
    w.bootstrap()
    gone = 0
    for _ in range(5):
        w.step()
        gone += w.absorbed[ELECTRON]
        xe = w.part[ELECTRON][0]
        assert np.isnan(xe).sum() == gone
        alive = ~np.isnan(xe)
        assert (xe[alive] >= 0).all() and (xe[alive] < xl).all()
    assert gone > 100
    assert abs(w.den[1].sum() - (n - gone + w.absorbed[ELECTRON]) * w.spwt[1] / nm["dx"] ** 2) < 1e-6 * n * w.spwt[1] / nm["dx"] ** 2
