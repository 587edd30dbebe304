"""GPU: round-2 features through the C ABI — cell order inside a bin, warp-aggregated deposit, asynchronous dumps,
error reporting from picsp_step, the last-cell edge case of the deposit."""
import time

import numpy as np
import pytest

import picsp_b200
from oracle.oracle import ELECTRON, ION, Oracle, normalise
from picsp_b200 import Params, Simulation
from tests.helpers import GRIDS, RTOL, assert_grid_close, relerr

pytestmark = pytest.mark.gpu


def snapshot(sim):
    return ({g: sim.grid(g) for g in GRIDS} | {"pi": np.stack(sim.get_species(ION)), "pe": np.stack(sim.get_species(ELECTRON))}
            | {"ke": np.array([sim.computeKE(ION), sim.computeKE(ELECTRON)])})


@pytest.mark.parametrize("solver,numx,n", [(1, 96, 400_000), (2, 64, 150_000), (1, 256, 3_000_000)])
def test_cell_order_inside_a_bin_changes_storage_only(solver, numx, n):
    """The cell ordering (k_cell_count / k_cell_scan / k_cell_permute) permutes particles inside their bin's range and
    switches the mover's deposit to the warp-aggregated commit; integer accumulation makes the result independent of
    both: grids, phase space (in upload order) and KE are BIT-IDENTICAL with the ordering every 2 steps for both
    species, with the aggregation forced on without any ordering, forced off with it, and with everything off (the
    default), across several re-binnings."""
    nm = normalise()
    runs = []
    for cell_i, cell_e, agg in ((0, 0, -1), (2, 2, -1), (0, 0, 1), (1, 3, 0), (64, 0, -1)):
        with Simulation(Params(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n, solverType=solver)) as sim:
            sim.set_sort_period(ION, 7); sim.set_sort_period(ELECTRON, 4)
            sim.set_cell_sort_period(ION, cell_i); sim.set_cell_sort_period(ELECTRON, cell_e)
            sim.set_deposit_aggregation(ION, agg); sim.set_deposit_aggregation(ELECTRON, agg)
            sim.fill_synthetic(ION, n, seed=41, vth=nm["vth_i"])
            sim.fill_synthetic(ELECTRON, n, seed=42, vth=1.0, xdrift=nm["drift_e"])
            sim.bootstrap(); sim.step(9); sim.step(8)
            runs.append(snapshot(sim))
    for r in runs[1:]:
        for k in runs[0]:
            assert np.array_equal(runs[0][k], r[k]), f"{k}: cell ordering changed the result"

@pytest.mark.gpu
@pytest.mark.parametrize("solver,numx,n", [(1, 128, 300_000), (2, 80, 150_000)])
def test_bank_order_inside_a_chunk_changes_storage_only(solver, numx, n):
    """k_bank_order permutes the particles of every chunk (in place) after each (re-)binning so that the lanes of a warp
    sit on distinct shared-memory banks; integer accumulation makes the result independent of the order: grids, phase
    space (in upload order) and KE are BIT-IDENTICAL with the order on for both species, on for the ions only (the
    default), and off, across several re-binnings — and with the stand-alone re-sort instead of the re-binning mover."""
    nm = normalise()
    runs = []
    for bank_i, bank_e, flags in ((0, 0, 0), (1, 1, 0), (-1, -1, 0), (1, 0, 16), (1, 1, 16)):
        with Simulation(Params(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n, solverType=solver, flags=flags)) as sim:
            sim.set_sort_period(ION, 7); sim.set_sort_period(ELECTRON, 4)
            sim.set_bank_order(ION, bank_i); sim.set_bank_order(ELECTRON, bank_e)
            sim.fill_synthetic(ION, n, seed=41, vth=nm["vth_i"])
            sim.fill_synthetic(ELECTRON, n, seed=42, vth=1.0, xdrift=nm["drift_e"])
            sim.bootstrap(); sim.step(9); sim.step(8)
            runs.append(snapshot(sim))
    for r in runs[1:]:
        for k in runs[0]:
            assert np.array_equal(runs[0][k], r[k]), f"{k}: bank ordering changed the result"


@pytest.mark.gpu
def test_bank_ordered_upload_round_trips_and_matches_the_oracle():
    """An uploaded load (first binning + bank order of both species) downloads unchanged bit for bit, in upload order,
    and three steps on it equal the oracle."""
    nm = normalise()
    numx, n = 128, 400_000
    o = Oracle(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n, vth_i=nm["vth_i"], solver=1)
    o.seed(7); o.init(ION, 1); o.init(ELECTRON, 1)
    with Simulation(Params(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n, solverType=1)) as sim:
        sim.set_bank_order(ION, 1); sim.set_bank_order(ELECTRON, 1)
        sim.set_sort_period(ION, 2); sim.set_sort_period(ELECTRON, 2)
        for s in (ION, ELECTRON):
            sim.set_species(s, *o.get_species(s))
        for s in (ION, ELECTRON):         # the upload enqueued the first binning and the ordering
            got, want = sim.get_species(s), o.get_species(s)
            for k in range(4):
                assert np.array_equal(got[k], want[k]), "download after the first binning differs from the upload"
    with Simulation(Params(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n, solverType=1)) as sim:
        sim.set_bank_order(ION, 1); sim.set_bank_order(ELECTRON, 1)
        sim.set_sort_period(ION, 2); sim.set_sort_period(ELECTRON, 2)
        for s in (ION, ELECTRON):
            sim.set_species(s, *o.get_species(s))
        o.bootstrap(); sim.bootstrap()
        for st in range(3):
            o.step(1); sim.step(1)
            for name in GRIDS:
                assert_grid_close(sim.grid(name), o.grid(name), sim.nix, sim.niy, 10 * RTOL, f"step{st}/{name}")
            for s in (ION, ELECTRON):
                got, want = sim.get_species(s), o.get_species(s)
                for k in range(4):
                    assert relerr(got[k], want[k]) <= 10 * RTOL


def test_cell_ordered_store_is_actually_ordered_and_complete():
    """After the ordering pass the download (upload order) is unchanged bit for bit, and one more step equals the oracle."""
    nm = normalise()
    numx, n = 128, 500_000
    o = Oracle(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n, vth_i=nm["vth_i"], solver=1)
    o.seed(5); o.init(ION, 1); o.init(ELECTRON, 1)
    with Simulation(Params(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n, solverType=1)) as sim:
        sim.set_cell_sort_period(ION, 1); sim.set_cell_sort_period(ELECTRON, 1)
        for s in (ION, ELECTRON):
            sim.set_species(s, *o.get_species(s))
        o.bootstrap(); sim.bootstrap()
        for st in range(3):
            o.step(1); sim.step(1)
            for name in GRIDS:
                assert_grid_close(sim.grid(name), o.grid(name), sim.nix, sim.niy, 10 * RTOL, f"step{st}/{name}")
            for s in (ION, ELECTRON):
                got, want = sim.get_species(s), o.get_species(s)
                for k in range(4):
                    assert relerr(got[k], want[k]) <= 10 * RTOL


@pytest.mark.parametrize("flags", [0, 2], ids=["tiled-aggregated", "unsorted"])
def test_deposit_of_a_clustered_load_matches_the_oracle(flags):
    """Hundreds of thousands of particles in a handful of cells (what the reference's loadType 2 does at scale, SURVEY
    Q12): whole warps share a cell and go through the REDUX-aggregated commit; also warps with two and three cells, and
    a thin uniform background (singletons).  Density vs the oracle at 1e-12 and exact charge conservation."""
    nm = normalise()
    numx, n = 64, 600_000
    rng = np.random.default_rng(12)
    dx = nm["dx"]; xl = numx * dx
    x = rng.random(n) * xl; y = rng.random(n) * xl
    # 300k particles into 3 cells (contiguous runs), 100k alternating between 2 cells lane by lane, 50k cycling over 5 cells
    x[:100_000] = (10 + rng.random(100_000)) * dx; y[:100_000] = (20 + rng.random(100_000)) * dx
    x[100_000:200_000] = (10 + rng.random(100_000)) * dx; y[100_000:200_000] = (21 + rng.random(100_000)) * dx
    x[200_000:300_000] = (47 + rng.random(100_000)) * dx; y[200_000:300_000] = (63 + rng.random(100_000)) * dx   # last cell row: periodic fold
    k = np.arange(100_000)
    x[300_000:400_000] = (30 + (k & 1) + rng.random(100_000)) * dx; y[300_000:400_000] = (5 + rng.random(100_000)) * dx
    k = np.arange(50_000)
    x[400_000:450_000] = (3 + rng.random(50_000)) * dx; y[400_000:450_000] = (40 + (k % 5) + rng.random(50_000)) * dx
    v = np.zeros(n)
    o = Oracle(numx, numx, dx, nm["dt"], nm["mass_i"], n, n, solver=1)
    o.set_species(ELECTRON, x, y, v, v); o.set_species(ION, x[::-1].copy(), y[::-1].copy(), v, v)
    with Simulation(Params(numx, numx, dx, nm["dt"], nm["mass_i"], n, n, flags=flags)) as sim:
        sim.set_deposit_aggregation(ELECTRON, 1)          # forced on for the electrons, automatic for the ions
        sim.set_species(ELECTRON, x, y, v, v); sim.set_species(ION, x[::-1].copy(), y[::-1].copy(), v, v)
        for s, name in ((ELECTRON, "den_e"), (ION, "den_i")):
            o.scatterSpecies(s); sim.scatterSpecies(s)
            e = assert_grid_close(sim.grid(name), o.grid(name), sim.nix, sim.niy, RTOL, name)
            total = sim.grid(name).reshape(numx + 1, numx + 1)[:-1, :-1].sum()
            want = n * sim.p.spwt[s] / dx ** 2
            assert abs(total - want) <= 1e-12 * want
            print(f"clustered deposit {name} (flags {flags}): rel err vs oracle {e:.2e}")
        # and through the fused mover: bootstrap + 2 steps
        o.den[0][:] = 0; o.den[1][:] = 0
    o2 = Oracle(numx, numx, dx, nm["dt"], nm["mass_i"], n, n, solver=1)
    o2.set_species(ELECTRON, x, y, v + 0.2, v); o2.set_species(ION, x, y, v, v)
    with Simulation(Params(numx, numx, dx, nm["dt"], nm["mass_i"], n, n, flags=flags)) as sim:
        sim.set_deposit_aggregation(ELECTRON, 1); sim.set_deposit_aggregation(ION, 1)
        sim.set_species(ELECTRON, x, y, v + 0.2, v); sim.set_species(ION, x, y, v, v)
        o2.bootstrap(); sim.bootstrap(); o2.step(2); sim.step(2)
        for name in GRIDS:
            assert_grid_close(sim.grid(name), o2.grid(name), sim.nix, sim.niy, 10 * RTOL, name)
        for s in (ION, ELECTRON):
            for a, b in zip(sim.get_species(s), o2.get_species(s)):
                assert relerr(a, b) <= 10 * RTOL


@pytest.mark.parametrize("flags", [0, 2], ids=["tiled", "unsorted"])
def test_position_one_ulp_below_the_box_edge_keeps_its_charge(flags):
    """x = nextafter(xl, 0): x/dx rounds to numxCells, the reference deposits the whole weight on the last node row
    (di == 0, main.cpp:657-667), the fold adds it to row 0.  The weight must not be dropped."""
    nm = normalise()
    numx, n = 32, 4096
    dx = nm["dx"]; xl = numx * dx
    rng = np.random.default_rng(4)
    x = rng.random(n) * xl; y = rng.random(n) * xl
    x[:64] = np.nextafter(xl, 0.0); y[64:128] = np.nextafter(xl, 0.0)
    x[128:160] = np.nextafter(xl, 0.0); y[128:160] = np.nextafter(xl, 0.0)
    v = np.zeros(n)
    o = Oracle(numx, numx, dx, nm["dt"], nm["mass_i"], n, n, solver=1)
    o.set_species(ELECTRON, x, y, v, v); o.scatterSpecies(ELECTRON)
    with Simulation(Params(numx, numx, dx, nm["dt"], nm["mass_i"], n, n, flags=flags)) as sim:
        sim.set_species(ELECTRON, x, y, v, v)
        sim.scatterSpecies(ELECTRON)
        den = sim.grid("den_e")
        total = den.reshape(numx + 1, numx + 1)[:-1, :-1].sum()
        want = n * sim.p.spwt[1] / dx ** 2
        assert abs(total - want) <= 1e-12 * want, (total, want)
        assert_grid_close(den, o.grid("den_e"), sim.nix, sim.niy, RTOL, "den_e")


def test_async_dump_equals_the_synchronous_downloads_while_the_loop_goes_on():
    """picsp_dump_begin snapshots what the reference dumps (rows, den, phi, KE); steps enqueued before picsp_dump_wait
    must not leak into it."""
    nm = normalise()
    numx, n = 96, 300_000
    with Simulation(Params(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n)) as sim:
        sim.fill_synthetic(ION, n, seed=3, vth=nm["vth_i"])
        sim.fill_synthetic(ELECTRON, n, seed=4, vth=1.0, xdrift=nm["drift_e"])
        sim.bootstrap(); sim.step(5)
        want = {"rows_i": sim.get_species_rows(ION), "rows_e": sim.get_species_rows(ELECTRON), "den_i": sim.grid("den_i"),
                "den_e": sim.grid("den_e"), "phi": sim.grid("phi"), "ke": np.array([sim.computeKE(ION), sim.computeKE(ELECTRON)])}
        nn = sim.nix * sim.niy
        got = {"rows_i": np.empty((n, 4)), "rows_e": np.empty((n, 4)), "den_i": np.empty(nn), "den_e": np.empty(nn),
               "phi": np.empty(nn), "ke": np.empty(2)}
        sim.dump_begin(got["rows_i"], got["rows_e"], got["den_i"], got["den_e"], got["phi"], got["ke"])
        sim.step(9)                      # includes re-binnings of the electrons: the snapshot must be unaffected
        sim.dump_wait()
        for k in want:
            assert np.array_equal(got[k], want[k]), k
        after = sim.dump()               # a second dump, begin + wait back to back, sees the advanced state
        assert not np.array_equal(after["rows_e"], want["rows_e"])
        assert np.array_equal(after["rows_e"], sim.get_species_rows(ELECTRON)) and np.array_equal(after["phi"], sim.grid("phi"))
        assert after["ke"][1] == sim.computeKE(ELECTRON)
        # partial dumps: any pointer may be NULL
        ke = np.empty(2)
        sim.dump_begin(ke=ke); sim.dump_wait()
        assert ke[0] == sim.computeKE(ION)


def test_picsp_step_reports_a_violation_of_an_earlier_call():
    """A displacement violation is flagged on the device by the mover; the grid phase of the next step mirrors the flag
    into mapped host memory and the next picsp_step call returns PICSP_ERR_DISPLACEMENT without an explicit sync."""
    nm = normalise()
    numx, n = 128, 20000
    rng = np.random.default_rng(2)
    xl = numx * nm["dx"]
    x, y = rng.random(n) * xl, rng.random(n) * xl
    vx = np.zeros(n); vx[:12000] = 40 * nm["dx"] / nm["dt"]      # more far movers than the fixed-point deposit has room for
    with Simulation(Params(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n)) as sim:
        sim.set_species(ION, x, y, np.zeros(n), np.zeros(n)); sim.set_species(ELECTRON, x, y, vx, np.zeros(n))
        sim.bootstrap(); sim.step(1)
        sim.step(1)                      # its grid phase publishes the flag
        code = None
        for _ in range(500):
            try:
                sim.step(0)              # enqueues nothing; only looks at the mirrored flag
            except picsp_b200.PicspError as e:
                code = e.code
                break
            time.sleep(0.01)
        assert code == -7


@pytest.mark.parametrize("numx,numy", [(64, 64), (32, 100), (100, 47)])
def test_small_grid_sor_sweep_in_shared_memory_is_bit_identical(numx, numy, monkeypatch):
    """Grids that fit shared memory sweep in one CTA (k_sor_sweep_smem); PICSP_SOR_NO_SMEM sends them through the
    pipelined bands instead and PICSP_FLAG_SOR_SINGLE_CTA through the single-CTA global-memory kernel: the three are
    the same lexicographic iterate with the same arithmetic — bit-identical phi after bootstrap + 3 steps."""
    nm = normalise()
    n = 20_000

    def run(flags):
        with Simulation(Params(numx, numy, nm["dx"], nm["dt"], nm["mass_i"], n, n, solverType=2, flags=flags)) as sim:
            sim.fill_synthetic(ION, n, seed=3, vth=nm["vth_i"]); sim.fill_synthetic(ELECTRON, n, seed=4, vth=1.0, xdrift=nm["drift_e"])
            sim.bootstrap(); sim.step(3)
            return sim.grid("phi"), sim.grid("rho")

    smem = run(0)
    single = run(8)
    monkeypatch.setenv("PICSP_SOR_NO_SMEM", "1")
    pipelined = run(0)
    for a, b in ((smem, single), (smem, pipelined)):
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
