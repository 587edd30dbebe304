#!/usr/bin/env python
"""tests/golden/make_golden.py — regenerates the golden vectors in this directory.

Every vector is an output of the UNMODIFIED reference translation unit
(/root/reference/src/main.cpp) compiled as oracle/_ref/libpicsp_ref*.so
(`make -C oracle ref`, needs /root/reference).  The reference ships no golden
vectors of its own, so these pin the oracle and the CUDA path to what
the reference's code computes in this container.

Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import ELECTRON, ION, REF_O0_SO, Reference, normalise  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
GRIDS = ("den_i", "den_e", "rho", "phi", "efx", "efy")


def snapshot(r, out, tag, particles=True):
    for g in GRIDS:
        out[f"{tag}/{g}"] = r.grid(g).copy()
    if not particles:
        return
    for s, nm in ((ION, "i"), (ELECTRON, "e")):
        x, y, vx, vy = r.get_species(s)
        out[f"{tag}/part_{nm}"] = np.stack([x, y, vx, vy])
    out[f"{tag}/ke"] = np.array([r.computeKE(ION), r.computeKE(ELECTRON)])


def loop_case(name, numx, n, solver, load_type, nsteps, lib=None, drift_e=0.0):
    """Loader -> bootstrap -> nsteps bodies of the time loop, state after each phase of step 0
    and after every step."""
    nm = normalise()
    kw = {} if lib is None else dict(lib_path=lib)
    r = Reference(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n, vth_i=nm["vth_i"], vth_e=nm["vth_e"],
                  solver=solver, **kw)
    r.seed(0)
    r.init_both(load_type, 0.0, drift_e)
    out = {"meta": np.array([numx, n, solver, load_type, nsteps], dtype=np.int64),
           "params": np.array([nm["dx"], nm["dt"], nm["mass_i"], nm["vth_i"], nm["vth_e"], drift_e])}
    snapshot(r, out, "loaded")
    # bootstrap, phase by phase (main.cpp:453-472)
    r.scatterSpecies(ION); r.scatterSpecies(ELECTRON); snapshot(r, out, "boot_deposit", particles=False)
    r.computeRho(); snapshot(r, out, "boot_rho", particles=False)
    r.solve(); snapshot(r, out, "boot_solve", particles=False)
    r.computeEF(); snapshot(r, out, "boot_ef", particles=False)
    r.rewindSpecies(ION); r.rewindSpecies(ELECTRON); snapshot(r, out, "boot_rewind")
    for st in range(nsteps):
        r.step(1)
        snapshot(r, out, f"step{st}")
    r.close()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name)


def edge_push_case():
    """Hand-placed particles that exercise the wrap / re-push chain (main.cpp:807-845),
    corner crossings, x == xmax landing, guard-band gathers and fast particles, on a
    non-trivial E field produced by the reference's own deposit/solve/EF."""
    nm = normalise()
    numx, n = 16, 4096
    r = Reference(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n, vth_i=nm["vth_i"], vth_e=nm["vth_e"], solver=2)
    r.seed(0); r.init_both(1)
    r.bootstrap()
    for _ in range(3):
        r.step(1)
    xl = numx * nm["dx"]; dt = nm["dt"]; eps = 1e-9
    cases = []
    for vx in (-8.0, -3.0, -1.0, -1e-3, 1e-3, 1.0, 3.0, 8.0):
        for vy in (-8.0, -3.0, -1e-3, 0.0, 1e-3, 3.0, 8.0):
            for x0 in (eps, xl / 3, xl - eps, 0.0):
                for y0 in (eps, xl / 2, xl - eps, 0.0):
                    cases.append((x0, y0, vx, vy))
    # a particle that lands exactly on xmax after the drift is hard to place analytically;
    # cover the >= branch with positions one ulp inside instead
    cases.append((np.nextafter(xl, 0), np.nextafter(xl, 0), 1e-30, 1e-30))
    c = np.array(cases)
    out = {"meta": np.array([numx, n], dtype=np.int64), "params": np.array([nm["dx"], nm["dt"], nm["mass_i"]])}
    for g in GRIDS:
        out["field/" + g] = r.grid(g).copy()
    out["in"] = c.T.copy()
    for s, tag in ((ION, "i"), (ELECTRON, "e")):
        r.set_species(s, c[:, 0], c[:, 1], c[:, 2], c[:, 3])
        r.pushSpecies(s)
        out["push_" + tag] = np.stack(r.get_species(s))
        r.set_species(s, c[:, 0], c[:, 1], c[:, 2], c[:, 3])
        r.rewindSpecies(s)
        out["rewind_" + tag] = np.stack(r.get_species(s))
    r.close()
    np.savez_compressed(os.path.join(OUT, "edge_push.npz"), **out)
    print("wrote edge_push", c.shape)


def rng_and_ini():
    r = Reference(8, 8, 0.01, 0.005, 1836.0, 8, 8)
    r.seed(0)
    vals = np.array([r.rnd() for _ in range(256)])
    r.close()
    np.save(os.path.join(OUT, "rng_mt19937_seed0.npy"), vals)
    ini = os.path.join(OUT, "input_ini_shipped.ini")
    cfg = Reference.parse_ini(ini)
    with open(os.path.join(OUT, "input_ini_parsed.json"), "w") as f:
        json.dump(cfg, f, indent=1, sort_keys=True)
    print("wrote rng + ini")


def whole_run():
    """The reference's real main() on its shipped input.ini, at the reference's own
    optimisation level (-O0, makefile:29): /timedata/energy and a few snapshots."""
    ini = os.path.join(OUT, "input_ini_shipped.ini")
    res = Reference.run_main(ini, lib_path=REF_O0_SO)
    keep = {"energy": res["/timedata/energy"]}
    for k in ("@Lx", "@Ly", "@dp", "@Nt", "@Nx", "@Ny"):
        keep["attr_" + k[1:]] = res[k]
    for ts in (0, 50, 500, 10000):
        keep[f"phi_{ts}"] = res[f"/phi/{ts}"]
    keep["den_e_0"] = res["/den.e/0"]; keep["den_i_50"] = res["/den.i/50"]
    keep["particle_e_0"] = res["/particle.e/0"]; keep["particle_i_0"] = res["/particle.i/0"]
    # momentum trace: sum of velocities of each species at every dump (201 dumps), from the reference's own
    # /particle.{i,e}/<ts> datasets — the north_star asks for energy AND momentum traces
    ts_list = sorted(int(k.split("/")[-1]) for k in res if k.startswith("/particle.e/"))
    keep["momentum_ts"] = np.array(ts_list)
    keep["momentum"] = np.array([[res[f"/particle.i/{ts}"][:, 2].sum(), res[f"/particle.i/{ts}"][:, 3].sum(),
                                  res[f"/particle.e/{ts}"][:, 2].sum(), res[f"/particle.e/{ts}"][:, 3].sum()] for ts in ts_list])
    names = sorted(k for k in res if not k.startswith("#"))
    keep["dataset_names"] = np.array(names)
    keep["groups"] = np.array(res["#groups"])
    np.savez_compressed(os.path.join(OUT, "whole_run_input_ini.npz"), **keep)
    print("wrote whole_run; energy[0], energy[-1] =", keep["energy"][0], keep["energy"][-1])


def _envelope_member(perturb):
    """One whole run of the shipped input.ini through the unmodified reference TU (-O0, like whole_run), with the x
    coordinate of electron `perturb` moved by ONE ulp after loading (perturb < 0: untouched).  Same loop and dump
    cadence as the reference's main() (main.cpp:453-531): traces sampled after the push of every 50th step."""
    nm = normalise()
    r = Reference(64, 64, nm["dx"], nm["dt"], nm["mass_i"], 10000, 10000, vth_i=nm["vth_i"], vth_e=nm["vth_e"],
                  solver=2, lib_path=REF_O0_SO)
    r.seed(0); r.init_both(2, 0.0, nm["drift_e"])
    if perturb >= 0:
        x, y, vx, vy = [a.copy() for a in r.get_species(ELECTRON)]
        x[perturb] = np.nextafter(x[perturb], np.inf)
        r.set_species(ELECTRON, x, y, vx, vy)
    r.bootstrap()
    mom, ke = [], []
    for ts in range(10001):
        r.step(1)
        if ts % 50 == 0:
            pi, pe = r.get_species(ION), r.get_species(ELECTRON)
            mom.append([pi[2].sum(), pi[3].sum(), pe[2].sum(), pe[3].sum()])
            ke.append([r.computeKE(ION), r.computeKE(ELECTRON)])
    r.close()
    return np.array(mom), np.array(ke)


def chaos_envelope(members=(-1, 2500, 5001, 7777, 9998, 1234, 42, 6100)):
    """How far the REFERENCE'S OWN long-run traces move under a one-ulp change of one particle coordinate: the
    two-stream run of input.ini is chaotic, so this ensemble is the yardstick for 'agrees within a statistical
    tolerance' (tests/test_gpu_host_run.py).  Member -1 is the unperturbed run and must reproduce whole_run()."""
    import multiprocessing as mp
    with mp.get_context("spawn").Pool(min(len(members), os.cpu_count() or 1)) as pool:
        res = pool.map(_envelope_member, members)
    whole = np.load(os.path.join(OUT, "whole_run_input_ini.npz"))
    assert np.array_equal(res[0][0], whole["momentum"]) and np.array_equal(res[0][1], whole["energy"]), \
        "the unperturbed member must equal the reference's main()"
    np.savez_compressed(os.path.join(OUT, "chaos_envelope_input_ini.npz"), members=np.array(members),
                        momentum=np.array([m for m, _ in res]), energy=np.array([k for _, k in res]))
    print("wrote chaos_envelope", len(members), "members")


if __name__ == "__main__":
    which = sys.argv[1:] or ["loops", "edge", "rng", "whole", "envelope"]
    if "loops" in which:
        loop_case("loop_sor_65_load2_O0", 64, 1500, 2, 2, 4, lib=REF_O0_SO, drift_e=normalise()["drift_e"])
        loop_case("loop_sor_33_load1", 32, 1500, 2, 1, 4)
        loop_case("loop_spectral_33_load1", 32, 1500, 1, 1, 4)
        loop_case("loop_spectral_48x_load1", 47, 1500, 1, 1, 2)   # even node count (48): Nyquist bin path
    if "edge" in which:
        edge_push_case()
    if "rng" in which:
        rng_and_ini()
    if "whole" in which:
        whole_run()
    if "envelope" in which:
        chaos_envelope()
