/* oracle/ref_harness.cpp — TEST INFRASTRUCTURE ONLY (builds oracle/_ref/libpicsp_ref*.so).
 *
 * The reference's own translation unit, /root/reference/src/main.cpp, is
 * #included UNMODIFIED (by absolute path, from where it lies; nothing is copied
 * into this repository) with `main` renamed, against the two shim headers in
 * oracle/shims/ (FFTW3 and HDF5 are absent from this image).  This file only adds
 * a C API that sets the reference's globals the way its `main` does
 * (main.cpp:363-433) and calls the reference's hot-path functions on
 * caller-supplied state, so tests can use the real reference as the oracle and
 * bench.py can time it (`--impl reference`, `cpu_baseline.kind = "reference"`).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load
 * the resulting library.
 */
#define main picsp_ref_main_entry
#include PICSP_REFERENCE_MAIN_CPP   /* -DPICSP_REFERENCE_MAIN_CPP="\"/root/reference/src/main.cpp\"" */
#undef main

#include <time.h>

namespace {

Species *g_species[2] = {nullptr, nullptr};   /* 0 = ions, 1 = electrons (main.cpp:407-412) */
double *g_guard_efx = nullptr, *g_guard_efy = nullptr;
size_t g_guard = 0;
double g_phase_seconds[8] = {0, 0, 0, 0, 0, 0, 0, 0};

double now_s() {
    timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

void free_state() {
    for (int s = 0; s < 2; s++) {
        if (g_species[s]) {
            delete[] g_species[s]->den; delete[] g_species[s]->xvel; delete[] g_species[s]->yvel;
            delete g_species[s]; g_species[s] = nullptr;
        }
    }
    if (g_guard_efx) {
        delete[] domain.phi; delete[] domain.rho;
        delete[] g_guard_efx; delete[] g_guard_efy;
        g_guard_efx = g_guard_efy = nullptr;
    }
}

}  // namespace

extern "C" {

/* Normalised parameters in, reference globals out — the statements of
 * main.cpp:363-433 with caller-chosen values.  efx/efy live inside zeroed guard
 * bands (>= niy+2 doubles each side): the reference's gather can index one row
 * outside the array on a re-push; the guard turns that undefined
 * read into a defined zero, and the CUDA path makes the same choice. */
int picsp_ref_setup(int numx, int numy, double stepSizeN, double timeStepN, double massIN,
                    int nI, int nE, double vthIN, double vthEN, int solver) {
    free_state();
    numxCells = numx; numyCells = numy; stepSize = stepSizeN; timeStep = timeStepN;
    EPS = 1.0; chargeE = 1.0; massI = massIN; massE = 1.0; density = 1.0;
    nParticlesI = nI; nParticlesE = nE; vthI = vthIN; vthE = vthEN;
    solverType = (short)solver;

    domain.nix = numxCells + 1; domain.dx = stepSize; domain.x0 = 0;
    domain.xl = (domain.nix - 1) * domain.dx; domain.xmax = domain.x0 + domain.xl;
    domain.niy = numyCells + 1; domain.dy = stepSize; domain.y0 = 0;
    domain.yl = (domain.niy - 1) * domain.dy; domain.ymax = domain.y0 + domain.yl;

    size_t n = (size_t)domain.nix * domain.niy;
    g_guard = 4 * (size_t)domain.niy + 8;
    domain.phi = new double[n]; domain.rho = new double[n];
    g_guard_efx = new double[n + 2 * g_guard]; g_guard_efy = new double[n + 2 * g_guard];
    memset(g_guard_efx, 0, sizeof(double) * (n + 2 * g_guard));
    memset(g_guard_efy, 0, sizeof(double) * (n + 2 * g_guard));
    domain.efx = g_guard_efx + g_guard; domain.efy = g_guard_efy + g_guard;
    memset(domain.phi, 0, sizeof(double) * n); memset(domain.rho, 0, sizeof(double) * n);

    ion_spwt = (density * numxCells * numyCells * stepSize * stepSize) / (nParticlesI);
    electron_spwt = (density * numxCells * numyCells * stepSize * stepSize) / (nParticlesE);
    g_species[0] = new Species("Ion", massI, chargeE, ion_spwt, nParticlesI, vthI);
    g_species[1] = new Species("Electrons", massE, -chargeE, electron_spwt, nParticlesE, vthE);
    for (int s = 0; s < 2; s++) {
        g_species[s]->den = new double[n]; g_species[s]->xvel = new double[n]; g_species[s]->yvel = new double[n];
        memset(g_species[s]->den, 0, sizeof(double) * n);
        memset(g_species[s]->xvel, 0, sizeof(double) * n);
        memset(g_species[s]->yvel, 0, sizeof(double) * n);
    }
    return 0;
}

void picsp_ref_teardown(void) { free_state(); }

double picsp_ref_spwt(int s) { return g_species[s]->spwt; }

/* which: 0 den_i, 1 den_e, 2 rho, 3 phi, 4 efx, 5 efy */
double *picsp_ref_grid_ptr(int which) {
    switch (which) {
        case 0: return g_species[0]->den;
        case 1: return g_species[1]->den;
        case 2: return domain.rho;
        case 3: return domain.phi;
        case 4: return domain.efx;
        case 5: return domain.efy;
    }
    return nullptr;
}

long picsp_ref_species_count(int s) { return (long)g_species[s]->part_list.size(); }

void picsp_ref_species_set(int s, const double *x, const double *y, const double *vx, const double *vy, long n) {
    g_species[s]->part_list.clear();
    for (long p = 0; p < n; p++) g_species[s]->add(Particle(x[p], y[p], vx[p], vy[p]));
}

void picsp_ref_species_get(int s, double *x, double *y, double *vx, double *vy) {
    long p = 0;
    for (auto &part : g_species[s]->part_list) {
        x[p] = part.xpos; y[p] = part.ypos; vx[p] = part.xvel; vy[p] = part.yvel; p++;
    }
}

/* --- the reference's own functions, one call each --- */
void picsp_ref_scatterSpecies(int s)    { scatterSpecies(g_species[s]); }
void picsp_ref_scatterSpeciesVel(int s) { scatterSpeciesVel(g_species[s]); }
void picsp_ref_computeRho(void)         { computeRho(domain.rho, g_species[0], g_species[1]); }
int  picsp_ref_solvePotential(void)     { return solvePotential(domain.phi, domain.rho) ? 1 : 0; }
int  picsp_ref_spectralPotentialSolver(void) { return spectralPotentialSolver(domain.phi, domain.rho) ? 1 : 0; }
void picsp_ref_computeEF(void)          { computeEF(domain.phi, domain.efx, domain.efy); }
void picsp_ref_pushSpecies(int s)       { pushSpecies(g_species[s], domain.efx, domain.efy); }
void picsp_ref_rewindSpecies(int s)     { rewindSpecies(g_species[s], domain.efx, domain.efy); }
double picsp_ref_computeKE(int s)       { return computeKE(g_species[s]); }

static void solve_selected() {
    if (solverType == 1) spectralPotentialSolver(domain.phi, domain.rho);
    else if (solverType == 2) solvePotential(domain.phi, domain.rho);
}

/* main.cpp:453-472 */
void picsp_ref_bootstrap(void) {
    scatterSpecies(g_species[0]); scatterSpecies(g_species[1]);
    computeRho(domain.rho, g_species[0], g_species[1]);
    solve_selected();
    computeEF(domain.phi, domain.efx, domain.efy);
    rewindSpecies(g_species[0], domain.efx, domain.efy);
    rewindSpecies(g_species[1], domain.efx, domain.efy);
}

/* One body of the time loop, main.cpp:481-504, timed per phase.
 * phase slots: 0 deposit, 1 dead velocity deposit, 2 rho, 3 solve, 4 EF, 5 push.
 * with_dead_vel=0 skips scatterSpeciesVel (its outputs are never consumed). */
void picsp_ref_step(int nsteps, int with_dead_vel) {
    for (int it = 0; it < nsteps; it++) {
        double t0 = now_s();
        scatterSpecies(g_species[0]); scatterSpecies(g_species[1]);
        double t1 = now_s();
        if (with_dead_vel) { scatterSpeciesVel(g_species[0]); scatterSpeciesVel(g_species[1]); }
        double t2 = now_s();
        computeRho(domain.rho, g_species[0], g_species[1]);
        double t3 = now_s();
        solve_selected();
        double t4 = now_s();
        computeEF(domain.phi, domain.efx, domain.efy);
        double t5 = now_s();
        pushSpecies(g_species[0], domain.efx, domain.efy);
        pushSpecies(g_species[1], domain.efx, domain.efy);
        double t6 = now_s();
        g_phase_seconds[0] += t1 - t0; g_phase_seconds[1] += t2 - t1; g_phase_seconds[2] += t3 - t2;
        g_phase_seconds[3] += t4 - t3; g_phase_seconds[4] += t5 - t4; g_phase_seconds[5] += t6 - t5;
    }
}
void picsp_ref_phase_seconds(double *out6, int reset) {
    for (int i = 0; i < 6; i++) { out6[i] = g_phase_seconds[i]; if (reset) g_phase_seconds[i] = 0; }
}

/* --- loader / RNG (main.cpp:49-54, 567-640) --- */
void picsp_ref_seed(unsigned seed) { mt_gen.seed(seed); rnd_dist.reset(); }
double picsp_ref_rnd(void) { return rnd(); }
void picsp_ref_init(int s, int load, double xdrift, double ydrift) {
    loadType = (short)load;
    g_species[s]->part_list.clear();
    init(g_species[s], xdrift, ydrift);
}

/* main.cpp:437-438: the two loads back to back, as `main` issues them (matters for
 * loadType 2, whose self-referencing initialiser reads a stale stack slot) */
void picsp_ref_init_both(int load, double driftI_, double driftE_) {
    loadType = (short)load;
    g_species[0]->part_list.clear(); g_species[1]->part_list.clear();
    init(g_species[0], driftI_, 0);
    init(g_species[1], driftE_, 0);
}

/* --- config (main.cpp:240-331); NOTE the reference exit()s on a failed sanity check --- */
int picsp_ref_parse_ini(const char *path, double *out /* 20 */) {
    std::string p(path);
    int rc = parse_ini_file(&p[0]);
    if (rc != 0) return rc;
    out[0] = nTimeSteps; out[1] = timeStep; out[2] = stepSize; out[3] = numxCells; out[4] = numyCells;
    out[5] = nParticlesI; out[6] = nParticlesE; out[7] = massI; out[8] = massE; out[9] = chargeE;
    out[10] = density; out[11] = vthE; out[12] = vthI; out[13] = driftE; out[14] = driftI;
    out[15] = dumpPeriod; out[16] = solverType; out[17] = loadType; out[18] = ion_spwt; out[19] = electron_spwt;
    return 0;
}

/* --- the reference's real main(), outputs captured by the H5 shim --- */
int picsp_ref_main(const char *ini_path) {
    /* the reference's main() deletes `file` and the six groups on exit and owns its
     * own field arrays; re-run its static initialisers (main.cpp:29-36) first */
    free_state();
    file = new H5File(FILE_NAME, H5F_ACC_TRUNC);
    groupE = new Group(file->createGroup("/particle.e"));
    groupI = new Group(file->createGroup("/particle.i"));
    groupT = new Group(file->createGroup("/timedata"));
    groupP = new Group(file->createGroup("/phi"));
    groupDE = new Group(file->createGroup("/den.e"));
    groupDI = new Group(file->createGroup("/den.i"));
    mt_gen.seed(0); rnd_dist.reset();
    std::string prog("picsp"), p(ini_path);
    char *argv[3] = {&prog[0], &p[0], nullptr};
    return picsp_ref_main_entry(2, argv);
}

long picsp_ref_h5_count(void) { return (long)picsp_shim_h5_registry().size(); }
const char *picsp_ref_h5_name(long i) { return picsp_shim_h5_registry()[i].name.c_str(); }
/* meta: is_attr, elem (0 f64 / 1 i32), rank, dim0, dim1, nbytes */
void picsp_ref_h5_meta(long i, long long *meta6) {
    const picsp_shim_h5_record &r = picsp_shim_h5_registry()[i];
    meta6[0] = r.is_attr; meta6[1] = r.elem; meta6[2] = r.rank;
    meta6[3] = (long long)r.dims[0]; meta6[4] = (long long)r.dims[1]; meta6[5] = (long long)r.bytes.size();
}
const void *picsp_ref_h5_data(long i) { return picsp_shim_h5_registry()[i].bytes.data(); }
long picsp_ref_h5_group_count(void) { return (long)picsp_shim_h5_groups().size(); }
const char *picsp_ref_h5_group_name(long i) { return picsp_shim_h5_groups()[i].c_str(); }

void picsp_ref_set_fft_mode(int mode) { oracle_dft_mode = mode; }

}  /* extern "C" */
