"""GPU: the library's own shared-memory DFT (picsp_b200/csrc/fft_kernels.cuh: prime-factor split + Bluestein on a
power-of-two length, one transform per CTA) against cuFFT and against the oracle's long-double DFT — the
spectralPotentialSolver of src/main.cpp:960-1058 on node counts of every shape."""
import numpy as np
import pytest

from oracle.oracle import ELECTRON, ION, Oracle, normalise
from picsp_b200 import Params, Simulation
from picsp_b200.sim import FLAG_CUFFT_ONLY, FLAG_OWN_FFT
from tests.helpers import GRIDS, RTOL, assert_grid_close, relerr

pytestmark = pytest.mark.gpu


def solve(numx, numy, flags, rho):
    nm = normalise()
    with Simulation(Params(numx, numy, nm["dx"], nm["dt"], nm["mass_i"], 8, 8, solverType=1, flags=flags)) as sim:
        sim.set_grid("rho", rho)
        sim.spectralPotentialSolver()
        return sim.grid("phi").reshape(numx + 1, numy + 1)


# nodes: 33 = 3*11, 49 = 7^2 (no coprime split: plain Bluestein), 48 (even: Nyquist bin), 65 = 5*13, 129 = 3*43, 257 prime,
# 513 = 27*19, rectangular grids with two different plans, 1025 = 25*41
@pytest.mark.parametrize("numx,numy", [(32, 32), (48, 48), (47, 47), (64, 64), (128, 128), (256, 256), (512, 512),
                                       (130, 33), (47, 64), (96, 255), (1024, 1024)])
def test_own_fft_equals_cufft(numx, numy):
    rng = np.random.default_rng(numx * 1000 + numy)
    rho = np.zeros((numx + 1, numy + 1)); rho[1:-1, 1:-1] = rng.standard_normal((numx - 1, numy - 1))
    a = solve(numx, numy, FLAG_OWN_FFT, rho)
    b = solve(numx, numy, FLAG_CUFFT_ONLY, rho)
    err = relerr(a, b)
    print(f"{numx + 1} x {numy + 1} nodes: own FFT vs cuFFT {err:.2e}")
    assert err <= 1e-13


@pytest.mark.parametrize("numx", [32, 47, 256, 2048])
def test_own_fft_solve_against_the_oracle(numx):
    """Against the oracle's DFT (long-double direct DFT for small lengths, cached Bluestein above)."""
    nm = normalise()
    rng = np.random.default_rng(numx)
    nix = numx + 1
    rho = np.zeros((nix, nix)); rho[1:-1, 1:-1] = rng.standard_normal((nix - 2, nix - 2))
    Oracle.lib().oracle_set_fft_mode(3 if numx > 128 else 0)
    try:
        o = Oracle(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], 8, 8, solver=1)
        o.set_grid("rho", rho)
        o.spectralPotentialSolver()
    finally:
        Oracle.lib().oracle_set_fft_mode(0)
    phi = solve(numx, numx, FLAG_OWN_FFT, rho)
    err = relerr(phi.reshape(-1), o.phi)
    print(f"{nix}^2 nodes: own FFT vs oracle {err:.2e}")
    assert err <= RTOL


def test_own_fft_in_the_time_loop():
    """257^2 nodes select the own transform automatically (257 is prime): bootstrap + 3 steps against the oracle, and
    bit-for-bit repeatable."""
    nm = normalise()
    numx, n = 256, 200_000
    Oracle.lib().oracle_set_fft_mode(3)
    try:
        o = Oracle(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n, vth_i=nm["vth_i"], solver=1)
        o.seed(11); o.init(ION, 1); o.init(ELECTRON, 1)
        up = {s: o.get_species(s) for s in (ION, ELECTRON)}
        o.bootstrap(); o.step(3)
    finally:
        Oracle.lib().oracle_set_fft_mode(0)
    runs = []
    for rep in range(2):
        with Simulation(Params(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n, solverType=1)) as sim:
            for s in (ION, ELECTRON):
                sim.set_species(s, *up[s])
            sim.bootstrap(); sim.step(3)
            for name in GRIDS:
                assert_grid_close(sim.grid(name), o.grid(name), sim.nix, sim.niy, 10 * RTOL, name)
            runs.append(sim.grid("phi"))
    assert np.array_equal(runs[0], runs[1])


@pytest.mark.parametrize("numx", [32, 64, 512, 1024])
def test_bluestein_path_on_lengths_that_normally_split_directly(numx, monkeypatch):
    """33 = 3*11, 65 = 5*13, 513 = 19*27 and 1025 = 25*41 normally take the two-direct-factor transform; PICSP_FFT_NO_DIRECT
    forces the prime-factor + Bluestein transform on them (25 rows of length 41 on 128 points ...): both against cuFFT."""
    rng = np.random.default_rng(numx + 7)
    rho = np.zeros((numx + 1, numx + 1)); rho[1:-1, 1:-1] = rng.standard_normal((numx - 1, numx - 1))
    ref = solve(numx, numx, FLAG_CUFFT_ONLY, rho)
    direct = solve(numx, numx, FLAG_OWN_FFT, rho)
    monkeypatch.setenv("PICSP_FFT_NO_DIRECT", "1")
    blue = solve(numx, numx, FLAG_OWN_FFT, rho)
    assert relerr(direct, ref) <= 1e-13 and relerr(blue, ref) <= 1e-13
    assert not np.array_equal(direct, blue), "the switch did not select another transform"
