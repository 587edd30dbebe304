"""Shared helpers for the parity tests (test infrastructure)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GRIDS = ("den_i", "den_e", "rho", "phi", "efx", "efy")

# north_star: single-step density, potential and phase space within 1e-12 relative (max-norm)
RTOL = 1e-12


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def relerr(a, b):
    """max|a-b| / max|b| (max-norm relative error; 0 if both are identically zero)."""
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    assert a.shape == b.shape, (a.shape, b.shape)
    fin = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), fin), "non-finite pattern differs"
    assert np.array_equal(a[~fin], b[~fin], equal_nan=True) or not (~fin).any()      # same infinities / NaNs in the same places
    if not fin.any():
        return 0.0
    d = np.abs(a[fin] - b[fin]).max()
    m = np.abs(b[fin]).max()
    return 0.0 if d == 0.0 else d / (m if m > 0 else 1.0)


def interior(a, nix, niy):
    return np.asarray(a).reshape(nix, niy)[1:-1, 1:-1]


def edges(a, nix, niy):
    g = np.asarray(a).reshape(nix, niy)
    return np.concatenate([g[0, :], g[-1, :], g[:, 0], g[:, -1]])


def assert_grid_close(a, b, nix, niy, tol=RTOL, what=""):
    """Interior and edge nodes are normalised separately: under the reference's
    accumulate-and-fold semantics edge values grow exponentially and
    would otherwise hide interior errors behind a huge max-norm."""
    ei = relerr(interior(a, nix, niy), interior(b, nix, niy))
    ee = relerr(edges(a, nix, niy), edges(b, nix, niy))
    assert ei <= tol, f"{what} interior rel err {ei:.3e} > {tol:g}"
    assert ee <= tol, f"{what} edge rel err {ee:.3e} > {tol:g}"
    return max(ei, ee)
