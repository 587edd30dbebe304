"""CPU: the C restatement (oracle/picsp_oracle.c) against the golden vectors produced by the
unmodified reference translation unit (tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest

from oracle.oracle import ELECTRON, ION, Oracle
from tests.helpers import GOLDEN, GRIDS, load_golden

LOOPS = ["loop_sor_65_load2_O0", "loop_sor_33_load1", "loop_spectral_33_load1", "loop_spectral_48x_load1"]


def make(g):
    numx, n, solver, load_type, nsteps = (int(v) for v in g["meta"])
    dx, dt, mass_i, vth_i, vth_e, drift_e = g["params"]
    o = Oracle(numx, numx, dx, dt, mass_i, n, n, vth_i=vth_i, vth_e=vth_e, solver=solver)
    return o, load_type, nsteps, drift_e


def check(o, g, tag, particles=True):
    for name in GRIDS:
        assert np.array_equal(o.grid(name), g[f"{tag}/{name}"]), f"{tag}/{name}"
    if particles:
        for s, nm in ((ION, "i"), (ELECTRON, "e")):
            assert np.array_equal(np.stack(o.get_species(s)), g[f"{tag}/part_{nm}"]), f"{tag}/part_{nm}"
        assert np.array_equal(np.array([o.computeKE(ION), o.computeKE(ELECTRON)]), g[f"{tag}/ke"])


@pytest.mark.parametrize("name", LOOPS)
def test_loop_bitwise(name):
    """Loader, bootstrap (phase by phase) and the time loop are bit-identical to the reference."""
    g = load_golden(name)
    o, load_type, nsteps, drift_e = make(g)
    o.seed(0)
    o.init(ION, load_type, 0.0, 0.0)
    o.init(ELECTRON, load_type, drift_e, 0.0)
    check(o, g, "loaded")
    o.scatterSpecies(ION); o.scatterSpecies(ELECTRON); check(o, g, "boot_deposit", False)
    o.computeRho(); check(o, g, "boot_rho", False)
    o.solve(); check(o, g, "boot_solve", False)
    o.computeEF(); check(o, g, "boot_ef", False)
    o.rewindSpecies(ION); o.rewindSpecies(ELECTRON); check(o, g, "boot_rewind")
    for st in range(nsteps):
        o.step(1)
        check(o, g, f"step{st}")


def test_edge_push_and_rewind_bitwise():
    """Wrap / re-push chain, corner crossings, fast particles and guard-band gathers."""
    g = load_golden("edge_push")
    numx, n = (int(v) for v in g["meta"])
    dx, dt, mass_i = g["params"]
    o = Oracle(numx, numx, dx, dt, mass_i, n, n, solver=2)
    for name in GRIDS:
        o.set_grid(name, g["field/" + name])
    cin = g["in"]
    total_extra = 0
    for s, tag in ((ION, "i"), (ELECTRON, "e")):
        o.set_species(s, *cin)
        total_extra += o.pushSpecies(s)
        assert np.array_equal(np.stack(o.get_species(s)), g["push_" + tag])
        o.set_species(s, *cin)
        o.rewindSpecies(s)
        assert np.array_equal(np.stack(o.get_species(s)), g["rewind_" + tag])
    assert total_extra > 100, "fixture must exercise the re-push path"


def test_rng_stream():
    o = Oracle(8, 8, 0.01, 0.005, 1836.0, 8, 8)
    o.seed(0)
    want = np.load(os.path.join(GOLDEN, "rng_mt19937_seed0.npy"))
    got = np.array([o.rnd() for _ in range(len(want))])
    assert np.array_equal(got, want)


def test_spectral_against_independent_fft():
    """The oracle's long-double DFT stands in for FFTW (absent here): cross-check the whole
    spectral solve against numpy's pocketfft, an independent FFT."""
    rng = np.random.default_rng(1)
    for numx in (32, 47, 64):
        o = Oracle(numx, numx, 0.017, 0.005, 1836.0, 8, 8, solver=1)
        nix = o.nix
        rho = np.zeros((nix, nix)); rho[1:-1, 1:-1] = rng.standard_normal((nix - 2, nix - 2))
        o.set_grid("rho", rho)
        o.spectralPotentialSolver()
        R = np.fft.rfft2(rho)
        PI = 3.14159265359
        L = numx * 0.017
        i = np.arange(nix)
        kx = np.where(i < nix // 2, 2.0 * PI * i / L, 2.0 * PI * (nix - i) / L)
        ky = 2.0 * PI * np.arange(nix // 2 + 1) / L
        with np.errstate(divide="ignore", invalid="ignore"):
            P = R / (kx[:, None] ** 2 + ky[None, :] ** 2)
        P[nix // 2, :] = 0.0
        P[0, 0] = 0.0
        phi = np.fft.irfft2(P, s=(nix, nix))
        err = np.abs(phi.reshape(-1) - o.phi).max() / np.abs(phi).max()
        assert err < 1e-12, (numx, err)


def test_bluestein_matches_direct_dft():
    L = Oracle.lib()
    rng = np.random.default_rng(2)
    numx = 40
    rho = np.zeros((numx + 1, numx + 1)); rho[1:-1, 1:-1] = rng.standard_normal((numx - 1, numx - 1))
    res = []
    for mode in (1, 2):
        L.oracle_set_fft_mode(mode)
        o = Oracle(numx, numx, 0.017, 0.005, 1836.0, 8, 8, solver=1)
        o.set_grid("rho", rho)
        o.spectralPotentialSolver()
        res.append(o.phi.copy())
    L.oracle_set_fft_mode(0)
    assert np.abs(res[0] - res[1]).max() / np.abs(res[0]).max() < 1e-15


def test_reference_quirks_are_restated():
    """accumulate-not-clear (Q1), rho boundary zero (Q3), one SOR sweep (Q6)."""
    o = Oracle(16, 16, 0.017, 0.005, 1836.0, 500, 500, solver=2)
    o.seed(0); o.init(ION, 1); o.init(ELECTRON, 1)
    o.scatterSpecies(ION)
    d1 = o.den[0].copy()
    o.scatterSpecies(ION)
    assert o.den[0].reshape(17, 17)[1:-1, 1:-1].sum() > 1.9 * d1.reshape(17, 17)[1:-1, 1:-1].sum()
    o.scatterSpecies(ELECTRON); o.computeRho()
    r = o.rho.reshape(17, 17)
    assert not r[0].any() and not r[-1].any() and not r[:, 0].any() and not r[:, -1].any()
    assert o.solvePotential() and o.last_sweeps == 1


def test_ini_golden_matches_shipped_values():
    with open(os.path.join(GOLDEN, "input_ini_parsed.json")) as f:
        cfg = json.load(f)
    assert cfg["numxCells"] == 64 and cfg["nParticlesI"] == 10000 and cfg["solverType"] == 2 and cfg["loadType"] == 2
    assert abs(cfg["timeStep"] - 0.005640957083153478) < 1e-18


def test_chaos_envelope_is_anchored_on_the_reference_main():
    """tests/golden/chaos_envelope_input_ini.npz (the reference re-run with one-ulp perturbations, the yardstick of
    the long-run statistical test): member 0 is the unperturbed run and equals the reference's own main() trace."""
    env, whole = load_golden("chaos_envelope_input_ini"), load_golden("whole_run_input_ini")
    assert int(env["members"][0]) == -1
    assert np.array_equal(env["energy"][0], whole["energy"]) and np.array_equal(env["momentum"][0], whole["momentum"])
    assert env["energy"].shape[1:] == (201, 2) and env["momentum"].shape[1:] == (201, 4)
    distinct = sum(not np.array_equal(env["momentum"][k], env["momentum"][0]) for k in range(1, len(env["members"])))
    assert distinct >= 5
    # before the instability has amplified the perturbation the members agree to round-off ...
    early = np.abs(env["energy"][:, :6] - env["energy"][0, :6]) / env["energy"][0, :6]
    assert early.max() < 1e-11
    # ... afterwards they do not: this is why the long-run comparison is statistical
    late = np.abs(env["energy"][:, -1] - env["energy"][0, -1]) / env["energy"][0, -1]
    assert late.max() > 1e-3
