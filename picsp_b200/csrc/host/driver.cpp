// The reference's `main` around the hot path (src/main.cpp:336-561), on top of the kernel ABI.
#include <chrono>
#include <cstdio>
#include <iostream>
#include <sys/stat.h>

#include "host.hpp"

namespace picsp_host {

namespace {
struct Guard {
    picsp_ctx *c = nullptr;
    ~Guard() { if (c) picsp_destroy(c); }
};
// page-locked host buffer (picsp_host_alloc): device->host copies into it do not block the caller
struct Pinned {
    double *p = nullptr;
    explicit Pinned(size_t n) : p((double *)picsp_host_alloc(sizeof(double) * (n ? n : 1))) {}
    ~Pinned() { picsp_host_free(p); }
    Pinned(const Pinned &) = delete;
    Pinned &operator=(const Pinned &) = delete;
};
#define HOST_CHECK(call)                                                          \
    do {                                                                          \
        int rc__ = (call);                                                        \
        if (rc__ != PICSP_OK) { if (err) *err = picsp_last_error(); return rc__; } \
    } while (0)
}  // namespace

int run(const std::string &ini_path, const std::string &out_path, int max_steps, bool quiet, int device, std::string *err) {
    auto start = std::chrono::steady_clock::now();
    picsp_run_config cfg;
    int rc = parse_run_config(ini_path, cfg, !quiet, err);
    if (rc != PICSP_OK) return rc;

    // the reference creates output/data.h5 and its six groups in static initialisers (main.cpp:29-36)
    std::string path = out_path.empty() ? "output/data.h5" : out_path;
    if (out_path.empty()) mkdir("output", 0777);
    H5Writer h5;
    if (!h5.open(path, err)) return PICSP_ERR_INVALID;
    for (const char *g : {"/particle.e", "/particle.i", "/timedata", "/phi", "/den.e", "/den.i"}) h5.create_group(g);
    // root attributes, main.cpp:348-353
    h5.write_attr_f64("Lx", cfg.numxCells * cfg.stepSize);
    h5.write_attr_f64("Ly", cfg.numyCells * cfg.stepSize);
    h5.write_attr_i32("dp", cfg.dumpPeriod);
    h5.write_attr_i32("Nt", cfg.nTimeSteps);
    h5.write_attr_i32("Nx", cfg.numxCells + 1);
    h5.write_attr_i32("Ny", cfg.numyCells + 1);

    const int64_t nI = cfg.nParticlesI, nE = cfg.nParticlesE;
    const int nix = cfg.numxCells + 1, niy = cfg.numyCells + 1;
    picsp_params prm = {};
    prm.numxCells = cfg.numxCells; prm.numyCells = cfg.numyCells;
    prm.stepSize = cfg.stepSize; prm.timeStep = cfg.timeStep;
    prm.solverType = cfg.solverType; prm.flags = 0;
    prm.charge[0] = cfg.chargeE; prm.charge[1] = -cfg.chargeE;      // main.cpp:407-408
    prm.mass[0] = cfg.massI; prm.mass[1] = cfg.massE;
    prm.spwt[0] = cfg.ion_spwt; prm.spwt[1] = cfg.electron_spwt;
    prm.capacity[0] = nI; prm.capacity[1] = nE;
    prm.device = device;
    Guard g;
    HOST_CHECK(picsp_create(&prm, &g.c));

    {   // init(&ions, driftI, 0); init(&electrons, driftE, 0)  (main.cpp:437-438); RNG seeded 0 (main.cpp:49)
        Loader ld(0);
        for (int s = 0; s < 2; s++) {
            const int64_t n = s == 0 ? nI : nE;
            std::vector<double> x(n), y(n), vx(n), vy(n);
            ld.fill(cfg, s, x.data(), y.data(), vx.data(), vy.data());
            HOST_CHECK(picsp_species_upload(g.c, s, x.data(), y.data(), vx.data(), vy.data(), n));
        }
    }
    if (!quiet) {   // main.cpp:440-450
        std::cout << "*********** Normalized Parameters ***********" << std::endl;
        std::cout << "Ion mass: " << cfg.massI << " charge: " << cfg.chargeE << " spwt: " << cfg.ion_spwt
                  << " Num of particles: " << cfg.nParticlesI << std::endl;
        std::cout << "Electrons mass: " << cfg.massE << " charge: " << -cfg.chargeE << " spwt: " << cfg.electron_spwt
                  << " Num of particles: " << cfg.nParticlesE << std::endl;
        std::cout << "vdriftE: " << cfg.driftE << " vdriftI: " << cfg.driftI << std::endl;
        std::cout << "density: " << cfg.density << std::endl;
        std::cout << "************ Simulation Parameters **********" << std::endl;
        std::cout << "Nx: " << cfg.numxCells << " Ny: " << cfg.numyCells << std::endl;
        std::cout << "Total timesteps: " << cfg.nTimeSteps << std::endl;
        std::cout << "timeStep: " << cfg.timeStep << " stepSize: " << cfg.stepSize << std::endl;
        std::cout << "********** Beginning of Simulation  **********" << std::endl;
    }

    HOST_CHECK(picsp_bootstrap(g.c));                                   // main.cpp:453-472

    // energy[int(nTimeSteps/dumpPeriod)+1][2] (main.cpp:360); rows are filled at ts % 50 == 0 (main.cpp:507),
    // heap-allocated and bounds-checked here (the reference's stack array overflows when dumpPeriod > 50)
    const int dp = cfg.dumpPeriod > 0 ? cfg.dumpPeriod : 1;
    const size_t nT = (size_t)(cfg.nTimeSteps / dp) + 1;
    std::vector<double> energy(2 * nT, 0.0);
    const int last = max_steps >= 0 && max_steps < cfg.nTimeSteps ? max_steps : cfg.nTimeSteps;

    // One set of page-locked dump buffers: a dump is snapshot on the device and copied out while the steps of the next
    // period run (picsp_dump_begin / picsp_dump_wait); it is written to the file when the next dump is due.
    const size_t nn = (size_t)nix * niy;
    Pinned rows_i(4 * (size_t)nI), rows_e(4 * (size_t)nE), den_i(nn), den_e(nn), phi(nn), ke(2);
    if (!rows_i.p || !rows_e.p || !den_i.p || !den_e.p || !phi.p || !ke.p) {
        if (err) *err = "cannot allocate page-locked dump buffers";
        return PICSP_ERR_CUDA;
    }
    int pending_ts = -1, ti = 0;
    auto flush_pending = [&]() -> int {           // writeSpecies x2, writePot, energy row (main.cpp:518-525)
        if (pending_ts < 0) return PICSP_OK;
        int rc = picsp_dump_wait(g.c);
        if (rc != PICSP_OK) return rc;
        const std::string t = std::to_string(pending_ts);
        h5.write_dataset_f64("/particle.i/" + t, rows_i.p, nI, 4);
        h5.write_dataset_f64("/den.i/" + t, den_i.p, nix, niy);
        h5.write_dataset_f64("/particle.e/" + t, rows_e.p, nE, 4);
        h5.write_dataset_f64("/den.e/" + t, den_e.p, nix, niy);
        h5.write_dataset_f64("/phi/" + t, phi.p, nix, niy);
        if ((size_t)ti < nT) { energy[2 * ti] = ke.p[0]; energy[2 * ti + 1] = ke.p[1]; }
        ti++;
        pending_ts = -1;
        return PICSP_OK;
    };
    // for (ts = 0; ts < nTimeSteps+1; ts++) { body; if (ts % 50 == 0) diagnostics }  (main.cpp:479-531): the bodies
    // between two diagnostics are enqueued by ONE picsp_step call (small runs replay CUDA graphs of step pairs)
    for (int ts = 0; ts <= last;) {
        const int next_diag = ((ts + 49) / 50) * 50;
        const int upto = next_diag <= last ? next_diag : last;        // index of the last body of this batch
        HOST_CHECK(picsp_step(g.c, upto - ts + 1));
        ts = upto + 1;
        if (upto % 50 == 0) {
            double max_phi = 0, phi0 = 0;
            HOST_CHECK(picsp_delta_phi(g.c, &max_phi, &phi0));
            if (!quiet) std::printf("TS: %i \t delta_phi: %.3g\n", upto, max_phi - phi0);
            HOST_CHECK(flush_pending());
            HOST_CHECK(picsp_dump_begin(g.c, rows_i.p, rows_e.p, den_i.p, den_e.p, phi.p, ke.p));
            pending_ts = upto;
        }
    }
    HOST_CHECK(flush_pending());
    h5.write_dataset_f64("/timedata/energy", energy.data(), nT, 2);     // writeKE, main.cpp:1179-1188
    if (!h5.close(err)) return PICSP_ERR_INVALID;
    if (!quiet) {
        std::chrono::duration<double> diff = std::chrono::steady_clock::now() - start;
        std::cout << "Total time taken by PICSP: " << diff.count() << " s" << std::endl;   // main.cpp:558
    }
    return PICSP_OK;
}

}  // namespace picsp_host
