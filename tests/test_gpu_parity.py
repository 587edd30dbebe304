"""GPU: parity of the CUDA path (through the C ABI) with the reference.

Checkers: the golden vectors produced by the unmodified reference TU (tests/golden/) and the
C restatement (oracle/), which is bit-identical to that TU.  Tolerance: the north_star's
1e-12 relative (max-norm) for density, potential, fields and phase space in double precision;
interior and edge nodes are normalised separately (see tests/helpers.assert_grid_close).
"""
import numpy as np
import pytest

import picsp_b200
from oracle.oracle import ELECTRON, ION, Oracle, normalise
from picsp_b200 import Params, Simulation
from picsp_b200.sim import FLAG_NO_GRAPH, FLAG_SEPARATE_SORT
from tests.helpers import GRIDS, RTOL, assert_grid_close, load_golden, relerr

pytestmark = pytest.mark.gpu

LOOPS = ["loop_sor_65_load2_O0", "loop_sor_33_load1", "loop_spectral_33_load1", "loop_spectral_48x_load1"]
PART = {ION: "i", ELECTRON: "e"}


def sim_from_golden(g, flags=0):
    numx, n, solver, load_type, nsteps = (int(v) for v in g["meta"])
    dx, dt, mass_i = g["params"][:3]
    return Simulation(Params(numx, numx, float(dx), float(dt), float(mass_i), n, n, solverType=solver, flags=flags)), nsteps


def put_state(sim, g, tag, particles_tag=None):
    for name in GRIDS:
        sim.set_grid(name, g[f"{tag}/{name}"])
    pt = particles_tag or tag
    for s in (ION, ELECTRON):
        sim.set_species(s, *g[f"{pt}/part_{PART[s]}"])


def check_grids(sim, g, tag, names=GRIDS, tol=RTOL):
    for name in names:
        assert_grid_close(sim.grid(name), g[f"{tag}/{name}"], sim.nix, sim.niy, tol, f"{tag}/{name}")


def check_particles(sim, g, tag, tol=RTOL):
    for s in (ION, ELECTRON):
        got = sim.get_species(s)
        want = g[f"{tag}/part_{PART[s]}"]
        for k, nm in enumerate("x y vx vy".split()):
            e = relerr(got[k], want[k])
            assert e <= tol, f"{tag} species {s} {nm}: rel err {e:.3e}"


@pytest.mark.parametrize("name", LOOPS)
def test_single_ops_against_reference_golden(name):
    """Each hot-path function on the reference's own input state for that phase."""
    g = load_golden(name)
    sim, _ = sim_from_golden(g)
    with sim:
        put_state(sim, g, "loaded")
        sim.scatterSpecies(ION); sim.scatterSpecies(ELECTRON)
        check_grids(sim, g, "boot_deposit", ("den_i", "den_e"))
        put_state(sim, g, "boot_deposit", "loaded"); sim.computeRho()
        check_grids(sim, g, "boot_rho", ("rho",))
        put_state(sim, g, "boot_rho", "loaded"); sim.solve()
        check_grids(sim, g, "boot_solve", ("phi",))
        put_state(sim, g, "boot_solve", "loaded"); sim.computeEF()
        check_grids(sim, g, "boot_ef", ("efx", "efy"))
        put_state(sim, g, "boot_ef", "loaded"); sim.rewindSpecies(ION); sim.rewindSpecies(ELECTRON)
        check_particles(sim, g, "boot_rewind")
        # one full loop body from the reference's state after the bootstrap
        put_state(sim, g, "boot_rewind")
        sim.step(1)
        check_grids(sim, g, "step0"); check_particles(sim, g, "step0")
        for s in (ION, ELECTRON):
            assert abs(sim.computeKE(s) - g["step0/ke"][s]) <= RTOL * abs(g["step0/ke"][s])


@pytest.mark.parametrize("name", LOOPS)
@pytest.mark.parametrize("flags", [0, 4, 2, 6], ids=["tiled-fused", "tiled-nofuse", "unsorted-fused", "unsorted-nofuse"])
def test_chained_loop_against_reference_golden(name, flags):
    """bootstrap + several steps with no resynchronisation (fused and unfused mover)."""
    g = load_golden(name)
    sim, nsteps = sim_from_golden(g, flags)
    with sim:
        put_state(sim, g, "loaded")
        sim.bootstrap()
        check_grids(sim, g, "boot_rewind"); check_particles(sim, g, "boot_rewind")
        for st in range(nsteps):
            sim.step(1)
            check_grids(sim, g, f"step{st}", tol=10 * RTOL)
            check_particles(sim, g, f"step{st}", tol=10 * RTOL)


def test_edge_push_and_rewind():
    """Wrap / re-push chain (main.cpp:807-845), corners, fast particles, guard-band gathers."""
    g = load_golden("edge_push")
    numx, n = (int(v) for v in g["meta"])
    dx, dt, mass_i = (float(v) for v in g["params"])
    cin = g["in"]
    with Simulation(Params(numx, numx, dx, dt, mass_i, n, n, solverType=2, capacity=(cin.shape[1],) * 2)) as sim:
        for name in GRIDS:
            sim.set_grid(name, g["field/" + name])
        for s in (ION, ELECTRON):
            sim.set_species(s, *cin)
            sim.pushSpecies(s)
            got = sim.get_species(s)
            want = g["push_" + PART[s]]
            for k in range(4):
                assert relerr(got[k], want[k]) <= RTOL, (s, k)
            assert sim.repush_count(s) > 50
            sim.set_species(s, *cin)
            sim.rewindSpecies(s)
            got = sim.get_species(s)
            want = g["rewind_" + PART[s]]
            for k in range(4):
                assert relerr(got[k], want[k]) <= RTOL, (s, k)


@pytest.mark.parametrize("solver,numx,n", [(1, 128, 300_000), (2, 128, 300_000), (1, 256, 400_000), (2, 100, 50_000)])
def test_step_against_oracle_midsize(solver, numx, n):
    """Seeded Maxwellian plasma, bootstrap + 2 steps, CUDA path vs the C restatement."""
    nm = normalise()
    o = Oracle(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n, vth_i=nm["vth_i"], solver=solver)
    o.seed(3); o.init(ION, 1); o.init(ELECTRON, 1)
    with Simulation(Params(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n, solverType=solver)) as sim:
        for s in (ION, ELECTRON):
            sim.set_species(s, *o.get_species(s))
        o.bootstrap(); sim.bootstrap()
        for st in range(2):
            o.step(1); sim.step(1)
            for name in GRIDS:
                assert_grid_close(sim.grid(name), o.grid(name), sim.nix, sim.niy, 10 * RTOL, f"step{st}/{name}")
            for s in (ION, ELECTRON):
                got, want = sim.get_species(s), o.get_species(s)
                for k in range(4):
                    assert relerr(got[k], want[k]) <= 10 * RTOL


@pytest.mark.parametrize("period", [1, 3, 1000], ids=["sort-every-step", "sort-every-3", "never-resort"])
def test_sort_period_does_not_change_results(period):
    """The tile sort only reorders storage: results after 12 steps are identical to 1e-12 whatever the
    sort cadence, including 'never re-sort' where fast electrons leave their window (straggler path)."""
    nm = normalise()
    numx, n, solver = 64, 60_000, 1
    o = Oracle(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n, vth_i=nm["vth_i"], solver=solver)
    o.seed(21); o.init(ION, 1); o.init(ELECTRON, 1)
    x, y, vx, vy = o.get_species(ELECTRON)
    o.set_species(ELECTRON, x, y, vx * 2.5, vy * 2.5)          # up to ~2 cells per step
    with Simulation(Params(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n, solverType=solver)) as sim:
        sim.set_sort_period(ION, period); sim.set_sort_period(ELECTRON, period)
        for s in (ION, ELECTRON):
            sim.set_species(s, *o.get_species(s))
        o.bootstrap(); sim.bootstrap()
        o.step(12); sim.step(12)
        if period == 1000:
            assert sim.straggler_count(ELECTRON) > 0, "fixture must exercise the straggler path"
        for name in GRIDS:
            assert_grid_close(sim.grid(name), o.grid(name), sim.nix, sim.niy, 100 * RTOL, name)
        for s in (ION, ELECTRON):
            got, want = sim.get_species(s), o.get_species(s)
            for k in range(4):
                assert relerr(got[k], want[k]) <= 100 * RTOL


@pytest.mark.parametrize("numx,numy,period", [(64, 64, 1), (32, 48, 1), (100, 72, 2), (16, 16, 1), (64, 64, 3), (32, 48, 2), (48, 32, 3),
                                              (16, 16, 2)])
def test_rebinning_mover_equals_separate_sort(numx, numy, period):
    """A due re-sort rides on the mover itself (the pushed particle is written straight into the new binned layout):
    k_tile_mover<4> with ranges reserved from the previous launch's per-chunk counts (period >= 2), k_tile_mover<3>
    with per-slice reservations when there are none (period 1: re-sort on consecutive steps).  Storage order is the only thing that may differ from the stand-alone re-sort
    (PICSP_FLAG_SEPARATE_SORT): grids and the phase space in upload order must be bit-identical, and both must
    match the oracle.  Grids with fewer than 3 tiles per side take the individual-slot path."""
    nm = normalise()
    n = 40_000
    o = Oracle(numx, numy, nm["dx"], nm["dt"], nm["mass_i"], n, n, vth_i=nm["vth_i"], solver=1)
    o.seed(33); o.init(ION, 1); o.init(ELECTRON, 1)
    x, y, vx, vy = o.get_species(ELECTRON)
    o.set_species(ELECTRON, x, y, vx * 2.5, vy * 2.5)          # fast electrons: many bin changes and periodic wraps
    runs = []
    for flags in (0, FLAG_SEPARATE_SORT):
        with Simulation(Params(numx, numy, nm["dx"], nm["dt"], nm["mass_i"], n, n, solverType=1, flags=flags)) as sim:
            sim.set_sort_period(ION, period); sim.set_sort_period(ELECTRON, period)
            for s in (ION, ELECTRON):
                sim.set_species(s, *o.get_species(s))
            sim.bootstrap(); sim.step(9)
            runs.append({g: sim.grid(g) for g in GRIDS} | {"pi": np.stack(sim.get_species(ION)),
                                                            "pe": np.stack(sim.get_species(ELECTRON))})
            if flags == 0:
                assert sim.repush_count(ELECTRON) > 0, "fixture must exercise periodic wraps"
    a, b = runs
    for k in a:
        assert np.array_equal(a[k], b[k]), f"{k}: re-binning mover differs from the stand-alone sort"
    o.bootstrap(); o.step(9)
    for name in GRIDS:
        assert_grid_close(a[name], o.grid(name), numx + 1, numy + 1, 100 * RTOL, name)
    for s, key in ((ION, "pi"), (ELECTRON, "pe")):
        want = o.get_species(s)
        for k in range(4):
            assert relerr(a[key][k], want[k]) <= 100 * RTOL


@pytest.mark.parametrize("solver", [1, 2])
@pytest.mark.parametrize("numx,numy", [(40, 72), (130, 33), (200, 64)])
def test_rectangular_grids(solver, numx, numy):
    """numxCells != numyCells (the reference allows it): bootstrap + 3 steps vs the oracle."""
    nm = normalise()
    n = 30_000
    o = Oracle(numx, numy, nm["dx"], nm["dt"], nm["mass_i"], n, n, vth_i=nm["vth_i"], solver=solver)
    o.seed(5); o.init(ION, 1); o.init(ELECTRON, 1)
    with Simulation(Params(numx, numy, nm["dx"], nm["dt"], nm["mass_i"], n, n, solverType=solver)) as sim:
        for s in (ION, ELECTRON):
            sim.set_species(s, *o.get_species(s))
        o.bootstrap(); sim.bootstrap()
        o.step(3); sim.step(3)
        for name in GRIDS:
            assert_grid_close(sim.grid(name), o.grid(name), sim.nix, sim.niy, 10 * RTOL, name)
        for s in (ION, ELECTRON):
            got, want = sim.get_species(s), o.get_species(s)
            for k in range(4):
                assert relerr(got[k], want[k]) <= 10 * RTOL


@pytest.mark.parametrize("numx,numy", [(24, 24), (64, 64), (65, 40), (130, 33), (150, 70), (197, 9), (512, 512)])
def test_sor_pipelined_equals_single_cta_and_oracle(numx, numy):
    """One SOR call (solvePotential, main.cpp:904-957) from a warm-start phi: the pipelined multi-CTA sweep,
    the single-CTA anti-diagonal sweep and the oracle's lexicographic loop give the same iterate."""
    nm = normalise()
    rng = np.random.default_rng(12)
    nix, niy = numx + 1, numy + 1
    rho = np.zeros((nix, niy)); rho[1:-1, 1:-1] = rng.standard_normal((nix - 2, niy - 2))
    phi0 = 1e-3 * rng.standard_normal((nix, niy))
    o = Oracle(numx, numy, nm["dx"], nm["dt"], nm["mass_i"], 8, 8, solver=2)
    o.set_grid("rho", rho); o.set_grid("phi", phi0)
    assert o.solvePotential() and o.last_sweeps == 1
    res = []
    for flags in (0, 8):
        with Simulation(Params(numx, numy, nm["dx"], nm["dt"], nm["mass_i"], 8, 8, solverType=2, flags=flags)) as sim:
            sim.set_grid("rho", rho); sim.set_grid("phi", phi0)
            sim.solvePotential()
            assert sim.last_sweeps == 1
            assert abs(sim.last_l2 - o.last_l2) <= 1e-9 * o.last_l2
            res.append(sim.grid("phi"))
    assert relerr(res[0], o.phi) <= RTOL and relerr(res[1], o.phi) <= RTOL
    assert np.array_equal(res[0], res[1]), "pipelined and single-CTA sweeps must be the same arithmetic"


def test_sor_multiple_sweeps_when_first_test_fails():
    """A source large enough that the reference's L2 < 1e-2 test fails after sweep 0: 101 sweeps, then converged."""
    nm = normalise()
    numx = 20
    nix = numx + 1
    rng = np.random.default_rng(3)
    base = rng.standard_normal((nix - 2, nix - 2)); base -= base.mean()
    rho = np.zeros((nix, nix)); rho[1:-1, 1:-1] = 3e4 * base
    o = Oracle(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], 8, 8, solver=2)
    o.set_grid("rho", rho)
    assert o.solvePotential()
    assert o.last_sweeps == 101, o.last_sweeps
    for flags in (0, 8):
        with Simulation(Params(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], 8, 8, solverType=2, flags=flags)) as sim:
            sim.set_grid("rho", rho)
            sim.solvePotential()
            assert sim.last_sweeps == o.last_sweeps
            assert relerr(sim.grid("phi"), o.phi) <= 1e-10


@pytest.mark.parametrize("flags", [0, 2], ids=["tiled", "unsorted"])
def test_density_is_deterministic_and_order_independent(flags):
    """Fixed-point accumulation: bit-identical density for any particle order and on repeat
    (within one code path; the tiled and unsorted paths round weights at different scales)."""
    nm = normalise()
    numx, n = 96, 200_000
    rng = np.random.default_rng(5)
    xl = numx * nm["dx"]
    x, y = rng.random(n) * xl, rng.random(n) * xl
    v = np.zeros(n)
    outs = []
    for trial in range(3):
        perm = np.arange(n) if trial == 0 else rng.permutation(n)
        with Simulation(Params(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n, flags=flags)) as sim:
            sim.set_species(ION, x[perm], y[perm], v, v)
            sim.scatterSpecies(ION)
            outs.append(sim.grid("den_i"))
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])


def test_fused_mover_density_is_deterministic():
    """Same state pushed twice in two contexts: the density accumulated by the fused mover and
    the phase space are bit-identical (independent of chunk scheduling and atomic arrival order)."""
    nm = normalise()
    numx, n = 96, 150_000
    res = []
    for trial in range(2):
        with Simulation(Params(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n)) as sim:
            sim.fill_synthetic(ION, n, seed=3, vth=nm["vth_i"])
            sim.fill_synthetic(ELECTRON, n, seed=4, vth=1.0, xdrift=nm["drift_e"])
            sim.bootstrap(); sim.step(3)
            res.append([sim.grid(g) for g in GRIDS] + list(sim.get_species(ELECTRON)))
    for a, b in zip(*res):
        assert np.array_equal(a, b)


def test_charge_conservation_large():
    """Size-independent property at a BASELINE-scale grid: the deposited weights sum to N*spwt/dx^2."""
    nm = normalise()
    numx, n = 1024, 20_000_000
    with Simulation(Params(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n)) as sim:
        sim.fill_synthetic(ELECTRON, n, seed=11)
        sim.scatterSpecies(ELECTRON)
        den = sim.grid("den_e").reshape(numx + 1, numx + 1)
        total = den[:-1, :-1].sum()            # unique periodic nodes after the fold
        want = n * sim.p.spwt[1] / nm["dx"] ** 2
        assert abs(total - want) <= 1e-12 * want
        # the mover keeps every particle inside the box and conserves the count
        sim.scatterSpecies(ION); sim.computeRho(); sim.solve(); sim.computeEF()
        sim.pushSpecies(ELECTRON)
        x, y, _, _ = sim.get_species(ELECTRON)
        xl = numx * nm["dx"]
        assert x.min() >= 0 and x.max() < xl and y.min() >= 0 and y.max() < xl


@pytest.mark.parametrize("numx", [256, 1024, 2048])
def test_spectral_solve_at_baseline_grid_sizes(numx):
    """spectralPotentialSolver (main.cpp:960-1058) on the node counts of BASELINE configs 2, 4 and 5
    (257 is prime, 1025 = 5^2*41, 2049 = 3*683: cuFFT takes its mixed-radix / Bluestein paths) against the
    oracle's DFT (cached double-precision Bluestein, itself checked against the long-double engine)."""
    nm = normalise()
    rng = np.random.default_rng(numx)
    nix = numx + 1
    rho = np.zeros((nix, nix)); rho[1:-1, 1:-1] = rng.standard_normal((nix - 2, nix - 2))
    Oracle.lib().oracle_set_fft_mode(3)
    try:
        o = Oracle(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], 8, 8, solver=1)
        o.set_grid("rho", rho)
        o.spectralPotentialSolver()
    finally:
        Oracle.lib().oracle_set_fft_mode(0)
    with Simulation(Params(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], 8, 8, solverType=1)) as sim:
        sim.set_grid("rho", rho)
        sim.spectralPotentialSolver()
        err = relerr(sim.grid("phi"), o.phi)
        assert err <= RTOL, f"phi rel err {err:.3e} at {nix}^2 nodes"
        sim.computeEF()
        o.computeEF()
        assert relerr(sim.grid("efx"), o.efx) <= RTOL and relerr(sim.grid("efy"), o.efy) <= RTOL


def test_full_size_cross_implementation_and_determinism():
    """BASELINE-scale grid (1024^2 cells, spectral), 4e7 particles: the tiled/fused path (TMA windows, shared-memory
    fixed point, periodic sort) and the unsorted path (global gathers, global 64-bit REDs) are two independent
    implementations of the same step; after bootstrap + 4 steps they agree to 1e-12 on every grid and on the phase
    space, and a repeat of the tiled run is bit-identical."""
    nm = normalise()
    numx, n = 1024, 20_000_000
    runs = []
    for flags in (0, 0, 2):
        with Simulation(Params(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n, flags=flags)) as sim:
            sim.set_sort_period(ELECTRON, 2)          # several re-sorts inside the window
            sim.fill_synthetic(ION, n, seed=21, vth=nm["vth_i"])
            sim.fill_synthetic(ELECTRON, n, seed=22, vth=1.0, xdrift=nm["drift_e"])
            sim.bootstrap(); sim.step(4)
            runs.append({g: sim.grid(g) for g in GRIDS} | {"pe": np.stack(sim.get_species(ELECTRON))}
                        | {"ke": np.array([sim.computeKE(ION), sim.computeKE(ELECTRON)])})
    a, b, c = runs
    for k in a:
        assert np.array_equal(a[k], b[k]), f"{k}: repeat of the tiled run is not bit-identical"
    for g in GRIDS:
        assert_grid_close(a[g], c[g], numx + 1, numx + 1, 10 * RTOL, f"tiled vs unsorted {g}")
    for k in range(4):
        assert relerr(a["pe"][k], c["pe"][k]) <= 10 * RTOL
    assert np.allclose(a["ke"], c["ke"], rtol=1e-12, atol=0)


def test_config5_grid_cross_implementation():
    """BASELINE config 5's grid (2048^2 cells: 16384 particle bins, the 64 KB shared-memory histogram, 2049^2-node
    FFT) with a thin plasma: tiled/fused path vs the unsorted path after bootstrap + 3 steps, and exact charge
    conservation of the deposit (sum of den * dx^2 / spwt == particle count, to round-off)."""
    nm = normalise()
    numx, n = 2048, 3_000_000
    runs = []
    for flags in (0, 2):
        with Simulation(Params(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n, flags=flags | 1)) as sim:   # CLEAR_DENSITY: den is one deposit
            sim.fill_synthetic(ION, n, seed=5, vth=nm["vth_i"])
            sim.fill_synthetic(ELECTRON, n, seed=6, vth=1.0, xdrift=nm["drift_e"])
            sim.bootstrap(); sim.step(3)
            runs.append({g: sim.grid(g) for g in GRIDS} | {"pe": np.stack(sim.get_species(ELECTRON))})
    a, c = runs
    for g in GRIDS:
        assert_grid_close(a[g], c[g], numx + 1, numx + 1, 10 * RTOL, f"tiled vs unsorted {g}")
    for k in range(4):
        assert relerr(a["pe"][k], c["pe"][k]) <= 10 * RTOL
    # den_e after the periodic fold: the unique periodic nodes carry every particle exactly once
    with Simulation(Params(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n)) as sim:
        want = n * sim.p.spwt[1] / nm["dx"] ** 2
    for r in runs:
        total = r["den_e"].reshape(numx + 1, numx + 1)[:-1, :-1].sum()
        assert abs(total - want) <= 1e-12 * want


def test_ke_after_a_dump_uses_the_staged_velocities_and_is_bit_identical():
    """computeKE right after a download reduces the upload-order velocities the download left in the staging
    buffers instead of re-ordering the terms again: same reduction tree, so the value is bit-identical to the
    stand-alone path, before and after, and any push invalidates the shortcut."""
    nm = normalise()
    numx, n = 96, 300_000
    with Simulation(Params(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n)) as sim:
        sim.fill_synthetic(ION, n, seed=3, vth=nm["vth_i"])
        sim.fill_synthetic(ELECTRON, n, seed=4, vth=1.0, xdrift=nm["drift_e"])
        sim.bootstrap(); sim.step(3)
        for s in (ION, ELECTRON):
            before = sim.computeKE(s)
            _, _, vx, vy = sim.get_species(s)
            after = sim.computeKE(s)
            assert before == after
            mass = sim.p.massI if s == ION else 1.0          # KE = sum(v^2) + 0.5*spwt*m (main.cpp:1190-1203, Q10)
            assert abs(after - 0.5 * sim.p.spwt[s] * mass - np.sum(vx * vx + vy * vy)) <= 1e-10 * after
        sim.step(1)
        ke1 = sim.computeKE(ELECTRON)
        _, _, vx, vy = sim.get_species(ELECTRON)
        assert ke1 == sim.computeKE(ELECTRON) and ke1 != after


@pytest.mark.parametrize("numx,shape", [(512, "corner"), (1024, "uniform"), (256, "stripe")])
def test_two_pass_first_binning(numx, shape):
    """First binning of a large arbitrary load (k_sort_pass: coarse pass, fine pass).  'corner' puts 90 % of the
    particles into one corner and spreads the rest thinly, so that a 4096-particle slice of the coarse-ordered store
    spans many coarse ranges (individual-slot path); the binned store must hold exactly the uploaded particles
    (download == upload, bit for bit, in upload order) and give the same step as the unsorted implementation."""
    nm = normalise()
    n = 250_000
    rng = np.random.default_rng(17)
    xl = numx * nm["dx"]
    if shape == "corner":
        x = np.where(rng.random(n) < 0.9, rng.random(n) * xl / 40, rng.random(n) * xl)
        y = np.where(rng.random(n) < 0.9, rng.random(n) * xl / 40, rng.random(n) * xl)
    elif shape == "stripe":
        x = rng.random(n) * xl; y = (0.45 + 0.1 * rng.random(n)) * xl
    else:
        x = rng.random(n) * xl; y = rng.random(n) * xl
    vx, vy = rng.standard_normal(n), rng.standard_normal(n)
    runs = []
    for flags in (0, 2):
        with Simulation(Params(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n, flags=flags)) as sim:
            sim.set_species(ELECTRON, x, y, vx, vy)
            sim.set_species(ION, x[::-1].copy(), y[::-1].copy(), 0.02 * vx, 0.02 * vy)
            sim.scatterSpecies(ELECTRON)                       # forces the binning, leaves the particles alone
            got = sim.get_species(ELECTRON)
            for a, b in zip(got, (x, y, vx, vy)):
                assert np.array_equal(a, b)
            sim.bootstrap(); sim.step(2)
            runs.append({g: sim.grid(g) for g in ("rho", "phi", "efx", "efy")} | {"pe": np.stack(sim.get_species(ELECTRON)),
                                                                                   "pi": np.stack(sim.get_species(ION))})
    a, b = runs
    for g in ("rho", "phi", "efx", "efy"):
        assert_grid_close(a[g], b[g], numx + 1, numx + 1, 10 * RTOL, f"tiled vs unsorted {g}")
    for key in ("pe", "pi"):
        for k in range(4):
            assert relerr(a[key][k], b[key][k]) <= 10 * RTOL


@pytest.mark.parametrize("flags", [0, 2], ids=["tiled", "unsorted"])
def test_row_layout_dump_equals_array_download(flags):
    """picsp_species_download_rows (the [n][4] layout writeSpecies dumps, main.cpp:1152-1162) is built on the device
    in the idle buffer set; it must equal the four-array download, and must not disturb a later KE or download."""
    nm = normalise()
    numx, n = 80, 123_457
    with Simulation(Params(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n, flags=flags)) as sim:
        sim.fill_synthetic(ION, n, seed=8, vth=nm["vth_i"])
        sim.fill_synthetic(ELECTRON, n, seed=9, vth=1.0, xdrift=nm["drift_e"])
        sim.bootstrap(); sim.step(3)
        for s in (ION, ELECTRON):
            cols = np.stack(sim.get_species(s))
            ke = sim.computeKE(s)                      # staged-velocity shortcut (tiled store)
            rows = sim.get_species_rows(s)
            assert rows.shape == (n, 4) and np.array_equal(rows.T, cols)
            assert sim.computeKE(s) == ke              # the row dump overwrote the staging block: full path, same value
            assert np.array_equal(np.stack(sim.get_species(s)), cols)


@pytest.mark.parametrize("solver", [1, 2])
def test_graph_replay_of_step_pairs_is_bit_identical(solver):
    """Launch-bound populations replay a captured CUDA graph of two consecutive steps between re-binnings.  The graph
    contains exactly the launches of the plain path, so 37 steps (odd count, several re-binnings of both species, a
    second picsp_step call that reuses the graph, a single picsp_push in between that flips the histogram buffers
    and so invalidates it) must give bit-identical grids, phase space, KE and launch count with and without it."""
    nm = normalise()
    numx, n = 64, 30_000
    runs = []
    for flags in (0, FLAG_NO_GRAPH):
        with Simulation(Params(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n, solverType=solver, flags=flags)) as sim:
            sim.set_sort_period(ION, 11); sim.set_sort_period(ELECTRON, 5)
            sim.fill_synthetic(ION, n, seed=31, vth=nm["vth_i"])
            sim.fill_synthetic(ELECTRON, n, seed=32, vth=1.5, xdrift=nm["drift_e"])
            sim.bootstrap()
            l0 = sim.kernel_launches()
            sim.step(21); sim.step(16)
            launches = sim.kernel_launches() - l0
            sim.scatterSpecies(ION); sim.scatterSpecies(ELECTRON); sim.computeRho(); sim.solve(); sim.computeEF()
            sim.pushSpecies(ION); sim.pushSpecies(ELECTRON)          # one step through the per-function calls
            sim.step(6)
            runs.append({g: sim.grid(g) for g in GRIDS} | {"pi": np.stack(sim.get_species(ION)), "pe": np.stack(sim.get_species(ELECTRON)),
                                                            "ke": np.array([sim.computeKE(ION), sim.computeKE(ELECTRON)]),
                                                            "launches": np.array([launches])})
    a, b = runs
    for k in a:
        assert np.array_equal(a[k], b[k]), f"{k}: graph replay differs from the plain path"


def test_clear_density_extension_and_accumulate_default():
    nm = normalise()
    numx, n = 32, 5000
    rng = np.random.default_rng(9)
    xl = numx * nm["dx"]
    x, y, v = rng.random(n) * xl, rng.random(n) * xl, np.zeros(n)
    res = {}
    for flags in (0, 1):
        with Simulation(Params(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n, flags=flags)) as sim:
            sim.set_species(ION, x, y, v, v)
            sim.scatterSpecies(ION); a = sim.grid("den_i")
            sim.scatterSpecies(ION); b = sim.grid("den_i")
            res[flags] = (a, b)
    a, b = res[0]
    assert b.reshape(numx + 1, -1)[1:-1, 1:-1].sum() > 1.99 * a.reshape(numx + 1, -1)[1:-1, 1:-1].sum()   # Q1: accumulates
    a, b = res[1]
    assert np.array_equal(a, b)


@pytest.mark.parametrize("solver", [1, 2])
@pytest.mark.parametrize("numx,numy", [(3, 3), (4, 7), (9, 5), (16, 16), (17, 31)])
def test_tiny_grids(solver, numx, numy):
    """Grids smaller than one particle tile / one SOR band, odd and even node counts."""
    nm = normalise()
    n = 700
    o = Oracle(numx, numy, nm["dx"], nm["dt"], nm["mass_i"], n, n, vth_i=nm["vth_i"], solver=solver)
    o.seed(17); o.init(ION, 1); o.init(ELECTRON, 1)
    with Simulation(Params(numx, numy, nm["dx"], nm["dt"], nm["mass_i"], n, n, solverType=solver)) as sim:
        for s in (ION, ELECTRON):
            sim.set_species(s, *o.get_species(s))
        o.bootstrap(); sim.bootstrap()
        o.step(3); sim.step(3)
        for name in GRIDS:
            assert_grid_close(sim.grid(name), o.grid(name), sim.nix, sim.niy, 10 * RTOL, name)
        for s in (ION, ELECTRON):
            got, want = sim.get_species(s), o.get_species(s)
            for k in range(4):
                assert relerr(got[k], want[k]) <= 10 * RTOL


def test_empty_species_and_single_particle():
    nm = normalise()
    numx = 32
    with Simulation(Params(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], 1, 1, solverType=1, capacity=(4, 4))) as sim:
        sim.set_species(ION, *(np.zeros(0),) * 4)                      # no ions at all
        sim.set_species(ELECTRON, [0.2], [0.3], [0.5], [-0.25])
        sim.bootstrap(); sim.step(2)
        assert sim.count(ION) == 0 and sim.count(ELECTRON) == 1
        o = Oracle(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], 1, 1, solver=1)
        o.set_species(ION, *(np.zeros(0),) * 4); o.set_species(ELECTRON, [0.2], [0.3], [0.5], [-0.25])
        o.bootstrap(); o.step(2)
        for k in range(4):
            assert relerr(sim.get_species(ELECTRON)[k], o.get_species(ELECTRON)[k]) <= RTOL
        assert_grid_close(sim.grid("phi"), o.grid("phi"), numx + 1, numx + 1, RTOL, "phi")


def test_displacement_guard_reports_an_error():
    """The fixed-point scale of the fused deposit bounds the particles that cross at most one 16-cell tile per step; the
    ones that cross more are counted, and only when there are so many of them that a node could overflow (here
    2^(62 - 49) = 8192 of 20000) does the launch report PICSP_ERR_DISPLACEMENT at the next synchronising call."""
    nm = normalise()
    numx, n = 128, 20000
    rng = np.random.default_rng(2)
    xl = numx * nm["dx"]
    x, y = rng.random(n) * xl, rng.random(n) * xl
    vx = np.zeros(n); vx[:12000] = 40 * nm["dx"] / nm["dt"]            # 40 cells in one step
    with Simulation(Params(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n)) as sim:
        sim.set_species(ION, x, y, np.zeros(n), np.zeros(n)); sim.set_species(ELECTRON, x, y, vx, np.zeros(n))
        sim.bootstrap(); sim.step(1)
        with pytest.raises(picsp_b200.PicspError) as ei:
            sim.sync()
        assert ei.value.code == -7


@pytest.mark.parametrize("flags", [0, 2, 4], ids=["tiled-fused", "unsorted", "tiled-unfused"])
def test_fast_particles_are_handled_like_in_the_reference(flags):
    """A few particles that cross SEVERAL tiles per step (40 and 70 cells, one of them through the periodic boundary): the
    reference has no speed limit, and neither has the mover — global gather, global deposit, individual slots at the
    re-binnings.  Four steps against the oracle."""
    nm = normalise()
    numx, n = 128, 30000
    rng = np.random.default_rng(8)
    xl = numx * nm["dx"]
    x, y = rng.random(n) * xl, rng.random(n) * xl
    vx = 0.3 * rng.standard_normal(n); vy = 0.3 * rng.standard_normal(n)
    vx[:5] = 40 * nm["dx"] / nm["dt"]; vy[5:9] = -70 * nm["dx"] / nm["dt"]; x[0] = 0.95 * xl
    o = Oracle(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n, vth_i=nm["vth_i"], solver=1)
    o.set_species(ION, x, y, np.zeros(n), np.zeros(n)); o.set_species(ELECTRON, x, y, vx, vy)
    with Simulation(Params(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n, flags=flags)) as sim:
        sim.set_sort_period(ELECTRON, 2)
        sim.set_species(ION, x, y, np.zeros(n), np.zeros(n)); sim.set_species(ELECTRON, x, y, vx, vy)
        o.bootstrap(); sim.bootstrap()
        for st in range(4):
            o.step(1); sim.step(1)
            for name in GRIDS:
                assert_grid_close(sim.grid(name), o.grid(name), sim.nix, sim.niy, 10 * RTOL, f"step{st}/{name}")
            got, want = sim.get_species(ELECTRON), o.get_species(ELECTRON)
            for k in range(4):
                assert relerr(got[k], want[k]) <= 10 * RTOL
        sim.sync()


def test_error_codes():
    nm = normalise()
    with Simulation(Params(16, 16, nm["dx"], nm["dt"], nm["mass_i"], 10, 10, solverType=2)) as sim:
        with pytest.raises(picsp_b200.PicspError) as ei:
            sim.scatterSpecies(2)
        assert ei.value.code == -1
        with pytest.raises(picsp_b200.PicspError):
            sim.set_species(ION, *(np.zeros(11),) * 4)       # exceeds capacity
        with pytest.raises(picsp_b200.PicspError) as ei:
            sim.spectralPotentialSolver()                     # SOR context has no FFT plans
        assert ei.value.code == -6
    with pytest.raises(picsp_b200.PicspError):
        Simulation(Params(16, 16, nm["dx"], nm["dt"], nm["mass_i"], 10, 10, solverType=3))
