#!/usr/bin/env python
"""oracle/gpu_patch.py — TEST INFRASTRUCTURE ONLY: proves the drop-in boundary by compiling it.

Applies the maintainer's patch of INTEGRATION.md section 2 to the reference's own translation unit,
/root/reference/src/main.cpp, and writes the result to oracle/_ref/main_gpu.cpp (git-ignored: a derived
artefact, never committed).  Everything the reference does around the hot path stays the reference's own code —
parse_ini_file, init (loader + RNG), the banner, writeSpecies / writePot / writeKE, the diagnostics cadence — and
every hot-path call inside main() (main.cpp:453-504) is replaced, one for one, by the C-ABI entry point that
include/picsp_b200.h declares for it:

    scatterSpecies(&ions)                 -> picsp_deposit(gpu, 0)
    scatterSpeciesVel(...)                -> dropped (outputs never consumed)
    computeRho(rho, &ions, &electrons)    -> picsp_compute_rho(gpu)
    spectralPotentialSolver(phi, rho)     -> picsp_solve_spectral(gpu)
    solvePotential(phi, rho)              -> picsp_solve_sor(gpu, NULL, NULL)
    computeEF(phi, efx, efy)              -> picsp_compute_ef(gpu)
    pushSpecies(&ions, efx, efy)          -> picsp_push(gpu, 0)
    rewindSpecies(&ions, efx, efy)        -> picsp_rewind(gpu, 0)
    computeKE(&ions)                      -> picsp_compute_ke(gpu, 0, &ke)

With -DPICSP_B200_FUSED_STEP the seven calls of one loop body collapse into picsp_step(gpu, 1), the form
INTEGRATION.md recommends.  Each edit is anchored on the reference's text and the script fails loudly if an anchor
is missing or ambiguous, so a change of the reference cannot silently produce a half-patched program.

    python oracle/gpu_patch.py [--reference /root/reference] [--out oracle/_ref/main_gpu.cpp]
"""
from __future__ import annotations

import argparse
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))

PROLOGUE = r'''
/* ---- inserted by oracle/gpu_patch.py (INTEGRATION.md section 2) ---------------------------------- */
#include "picsp_b200.h"
#define GPU(call) do { if ((call) != PICSP_OK) { fprintf(stderr, "picsp_b200: %s\n", picsp_last_error()); exit(EXIT_FAILURE); } } while (0)
static picsp_ctx *gpu = nullptr;
/* Species::part_list -> SoA upload (main.cpp:140) */
static void picsp_b200_push_species(int s, Species *sp) {
    std::vector<double> x, y, vx, vy;
    for (auto &p : sp->part_list) { x.push_back(p.xpos); y.push_back(p.ypos); vx.push_back(p.xvel); vy.push_back(p.yvel); }
    GPU(picsp_species_upload(gpu, s, x.data(), y.data(), vx.data(), vy.data(), (int64_t)x.size()));
}
/* what writeSpecies reads (main.cpp:1152-1172): the particle list in list order and species->den */
static void picsp_b200_pull_species(int s, Species *sp) {
    std::vector<double> rows(4 * sp->part_list.size());
    GPU(picsp_species_download_rows(gpu, s, rows.data()));
    size_t k = 0;
    for (auto &p : sp->part_list) { p.xpos = rows[4 * k]; p.ypos = rows[4 * k + 1]; p.xvel = rows[4 * k + 2]; p.yvel = rows[4 * k + 3]; k++; }
    GPU(picsp_grid_download(gpu, s == 0 ? PICSP_DEN_I : PICSP_DEN_E, sp->den));
}
/* ---------------------------------------------------------------------------------------------------- */
'''

CREATE = r'''
    /* ---- inserted by oracle/gpu_patch.py: context + upload (INTEGRATION.md section 2) ---- */
    {
        picsp_params prm = {};
        prm.numxCells = numxCells; prm.numyCells = numyCells;
        prm.stepSize = stepSize; prm.timeStep = timeStep;
        prm.solverType = solverType;
        prm.charge[0] = chargeE; prm.charge[1] = -chargeE;
        prm.mass[0] = massI; prm.mass[1] = massE;
        prm.spwt[0] = ion_spwt; prm.spwt[1] = electron_spwt;
        prm.capacity[0] = nParticlesI; prm.capacity[1] = nParticlesE;
        prm.device = 0;
        GPU(picsp_create(&prm, &gpu));
        picsp_b200_push_species(0, &ions); picsp_b200_push_species(1, &electrons);
    }
'''

DUMP = r'''
          /* ---- inserted by oracle/gpu_patch.py: bring back what the diagnostics below read ---- */
          GPU(picsp_grid_download(gpu, PICSP_PHI, phi));
          picsp_b200_pull_species(0, &ions); picsp_b200_pull_species(1, &electrons);
'''

# (regex on one source line inside main(), replacement, expected number of hits)
CALLS = [
    (r"^(\s*)scatterSpecies\(&ions\);", r"\1GPU(picsp_deposit(gpu, 0));", 2),
    (r"^(\s*)scatterSpecies\(&electrons\);", r"\1GPU(picsp_deposit(gpu, 1));", 2),
    (r"^(\s*)scatterSpeciesVel\(&(ions|electrons)\);", r"\1/* scatterSpeciesVel(&\2): dead work, dropped */", 2),
    (r"^(\s*)computeRho\(rho, &ions, &electrons\);", r"\1GPU(picsp_compute_rho(gpu));", 2),
    (r"^(\s*)spectralPotentialSolver\(phi, rho\);", r"\1GPU(picsp_solve_spectral(gpu));", 2),
    (r"^(\s*)solvePotential\(phi, rho\);", r"\1GPU(picsp_solve_sor(gpu, nullptr, nullptr));", 2),
    (r"^(\s*)computeEF\(phi,\s*efx,\s*efy\);", r"\1GPU(picsp_compute_ef(gpu));", 2),
    (r"^(\s*)pushSpecies\(&ions, efx, efy\);", r"\1GPU(picsp_push(gpu, 0));", 1),
    (r"^(\s*)pushSpecies\(&electrons, efx, efy\);", r"\1GPU(picsp_push(gpu, 1));", 1),
    (r"^(\s*)rewindSpecies\(&ions,efx,efy\);", r"\1GPU(picsp_rewind(gpu, 0));", 1),
    (r"^(\s*)rewindSpecies\(&electrons,efx,efy\);", r"\1GPU(picsp_rewind(gpu, 1));", 1),
    (r"^(\s*)energy\[ti\]\[0\] = computeKE\(&ions\);", r"\1GPU(picsp_compute_ke(gpu, 0, &energy[ti][0]));", 1),
    (r"^(\s*)energy\[ti\]\[1\] = computeKE\(&electrons\);", r"\1GPU(picsp_compute_ke(gpu, 1, &energy[ti][1]));", 1),
]


def patch(src: str) -> str:
    lines = src.split("\n")
    main_at = [i for i, l in enumerate(lines) if re.match(r"^int main\(int argc, char \*argv\[\]\)", l)]
    assert len(main_at) == 1, "anchor `int main(int argc, char *argv[])` not found exactly once"
    m0 = main_at[0]
    # main() ends at the first line that is exactly "}" after it
    m1 = next(i for i in range(m0 + 1, len(lines)) if lines[i] == "}")
    body = lines[m0:m1 + 1]

    for rx, rep, want in CALLS:
        hits = 0
        for i, l in enumerate(body):
            if l.lstrip().startswith("//"):
                continue
            new, n = re.subn(rx, rep, l)
            if n:
                body[i] = new
                hits += n
        assert hits == want, f"anchor {rx!r}: {hits} hits inside main(), expected {want}"

    def insert_after(pattern, text, what):
        at = [i for i, l in enumerate(body) if re.match(pattern, l)]
        assert len(at) == 1, f"anchor for {what}: {len(at)} hits"
        body[at[0] + 1:at[0] + 1] = text.strip("\n").split("\n")
        return at[0]

    insert_after(r"^\s*init\(&electrons,driftE,0\);", CREATE, "context creation")
    # the diagnostics block: `if(ts%50== 0)` followed by its opening brace
    at = [i for i, l in enumerate(body) if re.match(r"^\s*if\(ts%50== 0\)", l)]
    assert len(at) == 1 and body[at[0] + 1].strip() == "{", "anchor for the diagnostics block"
    body[at[0] + 2:at[0] + 2] = DUMP.strip("\n").split("\n")
    at = [i for i, l in enumerate(body) if re.match(r"^\s*writeKE\(energy\);", l)]
    assert len(at) == 1, "anchor writeKE(energy)"
    body[at[0] + 1:at[0] + 1] = ["    picsp_destroy(gpu); gpu = nullptr;   /* inserted by oracle/gpu_patch.py */"]

    # optional fused form: the seven calls of the loop body -> picsp_step(gpu, 1)
    loop = [i for i, l in enumerate(body) if re.match(r"^\s*for \(int ts=0; ts<nTimeSteps\+1; ts\+\+\)", l)]
    assert len(loop) == 1, "anchor for the time loop"
    first = next(i for i in range(loop[0], len(body)) if "GPU(picsp_deposit(gpu, 0));" in body[i])
    last = next(i for i in range(first, len(body)) if "GPU(picsp_push(gpu, 1));" in body[i])
    body[first:first] = ["#ifdef PICSP_B200_FUSED_STEP", "      GPU(picsp_step(gpu, 1));", "#else"]
    body[last + 4:last + 4] = ["#endif"]

    out = lines[:m0] + PROLOGUE.strip("\n").split("\n") + body + lines[m1 + 1:]
    return "\n".join(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default=os.environ.get("PICSP_REFERENCE_ROOT", "/root/reference"))
    ap.add_argument("--out", default=os.path.join(HERE, "_ref", "main_gpu.cpp"))
    a = ap.parse_args()
    with open(os.path.join(a.reference, "src", "main.cpp")) as f:
        src = f.read()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        f.write(patch(src))
    print(a.out)


if __name__ == "__main__":
    sys.exit(main())
