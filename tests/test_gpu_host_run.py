"""GPU: the whole drop-in program (`picsp_host_run` == `picsp_b200_run <input.ini>`) on BASELINE config 1,
the shipped input.ini, against the reference's own main() on the same file (tests/golden/whole_run_input_ini.npz,
produced at the reference's -O0 build): HDF5 names/shapes/attribute types, first dumps, energy rows."""
import os
import subprocess

import numpy as np
import pytest

from picsp_b200 import host
from picsp_b200.lib import PKG
from tests import h5mini
from tests.helpers import GOLDEN, load_golden, relerr

pytestmark = pytest.mark.gpu
INI = os.path.join(GOLDEN, "input_ini_shipped.ini")


def test_whole_program_first_100_steps(tmp_path):
    g = load_golden("whole_run_input_ini")
    out = str(tmp_path / "data.h5")
    host.run(INI, out, max_steps=100, quiet=True)
    f = h5mini.File(out)
    # layout contract: root attributes with the reference's types, six groups
    a = f.attrs()
    for k in ("Lx", "Ly"):
        assert a[k].dtype == np.float64 and a[k] == g["attr_" + k][0]
    for k in ("dp", "Nt", "Nx", "Ny"):
        assert a[k].dtype == np.int32 and a[k] == g["attr_" + k][0]
    assert sorted("/" + n for n in f.groups()) == sorted(g["groups"].tolist())
    assert f.datasets("phi") == sorted(["0", "50", "100"])
    assert f.read("/particle.e/0").shape == (10000, 4) and f.read("/den.i/50").shape == (65, 65)
    assert f.read("/timedata/energy").shape == g["energy"].shape
    # numbers.  The shipped load puts ions and electrons on (almost) the same positions, so at ts = 0 rho — and phi,
    # its solve — is a cancellation residue of two O(1) densities: phi_0 inherits the densities' last-bit
    # summation-order differences (1e-15) amplified by that ratio (measured 5.0e-11 on B200, hence 1e-9 for phi_0 only);
    # everything else is held to the north_star's 1e-12 / 1e-11 (measured: den 2e-15 .. 8e-15, phi_50 1.5e-14).
    errs = {
        "particle_e_0": (relerr(f.read("/particle.e/0"), g["particle_e_0"]), 1e-12),
        "particle_i_0": (relerr(f.read("/particle.i/0"), g["particle_i_0"]), 1e-12),
        "den_e_0": (relerr(f.read("/den.e/0")[1:-1, 1:-1], g["den_e_0"][1:-1, 1:-1]), 1e-12),
        "den_i_50": (relerr(f.read("/den.i/50")[1:-1, 1:-1], g["den_i_50"][1:-1, 1:-1]), 1e-12),
        "phi_0": (relerr(f.read("/phi/0"), g["phi_0"]), 1e-9),
        "phi_50": (relerr(f.read("/phi/50"), g["phi_50"]), 1e-11),
    }
    print("whole program, measured relative errors vs the reference's main():", {k: f"{v[0]:.2e}" for k, v in errs.items()})
    bad = {k: v for k, v in errs.items() if not v[0] <= v[1]}
    assert not bad, f"{bad} (all: {errs})"
    e = f.read("/timedata/energy")
    assert np.allclose(e[:3], g["energy"][:3], rtol=1e-11, atol=0)


def trace_stats(e, m, g, gm):
    """Distances between two whole runs (KE traces e/g: 201 x 2, momentum traces m/gm: 201 x 4)."""
    rel = np.abs(e - g) / np.abs(g)
    pk_a, pk_b = int(np.argmax(e[:30, 1])), int(np.argmax(g[:30, 1]))
    tr_a, tr_b = pk_a + int(np.argmin(e[pk_a:60, 1])), pk_b + int(np.argmin(g[pk_b:60, 1]))
    return {
        "ke_early": rel[:6].max(), "ke_all": rel.max(), "ke_final": rel[-1].max(),
        "ke_logcorr": min(np.corrcoef(np.log(e[:, k]), np.log(g[:, k]))[0, 1] for k in (0, 1)),
        "peak_shift": abs(pk_a - pk_b), "peak_value": abs(e[pk_a, 1] - g[pk_b, 1]) / g[pk_b, 1],
        "trough_shift": abs(tr_a - tr_b), "trough_value": abs(e[tr_a, 1] - g[tr_b, 1]) / g[tr_b, 1],
        "mom_early": (np.abs(m[:6] - gm[:6]).max(axis=1) / np.abs(gm[:6]).max(axis=1)).max(),
        "mom_corr": np.array([np.corrcoef(m[:, k], gm[:, k])[0, 1] for k in range(4)]),
        "mom_dev": np.array([np.abs(m[:, k] - gm[:, k]).max() / np.abs(gm[:, k]).max() for k in range(4)]),
    }


def test_long_run_energy_trace_statistics(tmp_path):
    """north_star: 'energy and momentum traces over long runs must agree within a stated statistical tolerance,
    since chaotic trajectories diverge'.  The full 10001-step run of the shipped input.ini (BASELINE config 1)
    against the reference's own main().

    The yardstick is the reference itself: tests/golden/chaos_envelope_input_ini.npz holds the traces of the
    unmodified reference TU re-run with ONE particle coordinate moved by ONE ulp (8 members, member 0 = main()).
    Round-off grows to O(1) differences after ts ~ 450; the spread between members is what 'the same run' means
    from then on.  Stated tolerances (envelope = worst value over all pairs of distinct members):
      * dumps up to ts = 250, before the instability amplifies round-off: KE 1e-9 relative, momentum 1e-7 of the
        largest component                                       [envelope 9e-13; measured here 2.6e-13 and 2e-9]
      * first KE_e maximum (ts ~ 650): position +-1 dump, value 2 %;  following minimum: +-2 dumps, 5 %
      * KE at every dump: 2 x envelope (envelope 1.3 % ions, 5.7 % electrons);  final KE: 2 x envelope
      * correlation of the log KE traces: 1 - corr <= 3 x envelope (envelope 1 - 0.99959)
      * momentum (sum of velocities per species and component; signed sums, diverge fastest): correlation of each
        trace with the reference's >= envelope minimum - 0.15 (envelope 0.99, 0.95, 0.82, 0.72); largest deviation
        <= 2 x envelope (envelope 0.17, 1.29, 0.50, 0.39 of the trace's own maximum)."""
    gold = load_golden("whole_run_input_ini")
    env = load_golden("chaos_envelope_input_ini")
    g, gm = gold["energy"], gold["momentum"]
    EK, EM = env["energy"], env["momentum"]
    assert np.array_equal(EK[0], g) and np.array_equal(EM[0], gm), "member 0 of the envelope is the reference's main()"
    pairs = [trace_stats(EK[a], EM[a], EK[b], EM[b]) for a in range(len(EK)) for b in range(len(EK))
             if a != b and not np.array_equal(EM[a], EM[b])]
    assert len(pairs) >= 20
    worst = {k: (np.min([p[k] for p in pairs], axis=0) if k in ("ke_logcorr", "mom_corr") else np.max([p[k] for p in pairs], axis=0))
             for k in pairs[0]}
    assert worst["ke_early"] < 1e-9 and worst["peak_shift"] <= 1 and worst["trough_shift"] <= 2   # the fixed bounds hold for the reference itself

    out = str(tmp_path / "full.h5")
    host.run(INI, out, max_steps=-1, quiet=True)
    f = h5mini.File(out)
    e = f.read("/timedata/energy")
    m = np.array([[f.read(f"/particle.i/{ts}")[:, 2].sum(), f.read(f"/particle.i/{ts}")[:, 3].sum(),
                   f.read(f"/particle.e/{ts}")[:, 2].sum(), f.read(f"/particle.e/{ts}")[:, 3].sum()]
                  for ts in gold["momentum_ts"]])
    assert e.shape == g.shape == (201, 2) and m.shape == gm.shape == (201, 4)
    s = trace_stats(e, m, g, gm)
    report = {k: (s[k], worst[k]) for k in s}
    assert s["ke_early"] < 1e-9 and s["mom_early"] <= 1e-7, report
    assert s["peak_shift"] <= 1 and s["peak_value"] <= 0.02, report
    assert s["trough_shift"] <= 2 and s["trough_value"] <= 0.05, report
    assert s["ke_all"] <= 2 * worst["ke_all"] and s["ke_final"] <= 2 * worst["ke_final"], report
    assert 1 - s["ke_logcorr"] <= 3 * (1 - worst["ke_logcorr"]), report
    assert (s["mom_corr"] >= worst["mom_corr"] - 0.15).all(), report
    assert (s["mom_dev"] <= 2 * worst["mom_dev"]).all(), report
    print("long-run statistics (ours vs reference, envelope):", report)


def test_cli_executable_prints_the_reference_banner(tmp_path):
    exe = os.path.join(PKG, "picsp_b200_run")
    out = str(tmp_path / "cli.h5")
    r = subprocess.run([exe, INI, "--out", out, "--steps", "0"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    for line in ("********** IMPORTANT PLASMA QUANTITIES ***********", "STATUS, Input parameters are compatible.",
                 "Ion mass: 1836.65 charge: 1 spwt: 0.000118562 Num of particles: 10000", "vdriftE: 0.222222 vdriftI: 0",
                 "Nx: 64 Ny: 64", "Total timesteps: 10000", "TS: 0 \t delta_phi:", "Total time taken by PICSP:"):
        assert line in r.stdout, line
    assert subprocess.run([exe], capture_output=True, text=True).returncode != 0
