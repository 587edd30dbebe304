// The reference's `main` around the hot path (src/main.cpp:336-561), on top of the kernel ABI.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>

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

// The NCCL unique id of a sharded run travels through a small file next to the output (rank 0 writes it atomically,
// the others poll): the host program has no launcher of its own, any launcher that sets RANK / WORLD_SIZE /
// LOCAL_RANK (torchrun --no-python, srun, a shell loop) will do.
static int exchange_unique_id(const std::string &id_path, int rank, char id[128], std::string *err) {
    if (rank == 0) {
        int rc = picsp_comm_unique_id(id);
        if (rc != PICSP_OK) { if (err) *err = picsp_last_error(); return rc; }
        const std::string tmp = id_path + ".tmp";
        FILE *f = std::fopen(tmp.c_str(), "wb");
        if (!f || std::fwrite(id, 1, 128, f) != 128) { if (f) std::fclose(f); if (err) *err = "cannot write " + tmp; return PICSP_ERR_INVALID; }
        std::fclose(f);
        if (std::rename(tmp.c_str(), id_path.c_str()) != 0) { if (err) *err = "cannot publish " + id_path; return PICSP_ERR_INVALID; }
        return PICSP_OK;
    }
    for (int tries = 0; tries < 6000; tries++) {          // up to 60 s
        FILE *f = std::fopen(id_path.c_str(), "rb");
        if (f) {
            const size_t n = std::fread(id, 1, 128, f);
            std::fclose(f);
            if (n == 128) return PICSP_OK;
        }
        std::this_thread::sleep_for(std::chrono::milliseconds(10));
    }
    if (err) *err = "timed out waiting for rank 0's communicator id at " + id_path;
    return PICSP_ERR_NCCL;
}

// One process per GPU.  rank r of nranks owns the particles [N*r/nranks, N*(r+1)/nranks) of each species in loader
// order; every rank runs the loader over ALL particles (the RNG stream and the loadType-2 recurrence are
// sequential) and keeps its range.  Rank 0 owns the file: it writes the grids (den.i / den.e arrive reduced over the
// ranks), the energies and the metadata; every rank writes its own rows of the /particle.i and /particle.e datasets into the block rank 0
// reserved (all ranks track the same offsets).
int run(const std::string &ini_path, const std::string &out_path, int max_steps, bool quiet, int device, std::string *err,
        int rank, int nranks) {
    auto start = std::chrono::steady_clock::now();
    if (rank != 0) quiet = true;
    picsp_run_config cfg;
    int rc = parse_run_config(ini_path, cfg, !quiet, err);
    if (rc != PICSP_OK) return rc;

    // the reference creates output/data.h5 and its six groups in static initialisers (main.cpp:29-36)
    std::string path = out_path.empty() ? "output/data.h5" : out_path;
    if (out_path.empty()) mkdir("output", 0777);
    H5Writer h5;
    if (rank == 0) {          // the other ranks open it (for their row slices only) once rank 0 has created it
        if (!h5.open(path, err)) return PICSP_ERR_INVALID;
        for (const char *g : {"/particle.e", "/particle.i", "/timedata", "/phi", "/den.e", "/den.i"}) h5.create_group(g);
        // root attributes, main.cpp:348-353
        h5.write_attr_f64("Lx", cfg.numxCells * cfg.stepSize);
        h5.write_attr_f64("Ly", cfg.numyCells * cfg.stepSize);
        h5.write_attr_i32("dp", cfg.dumpPeriod);
        h5.write_attr_i32("Nt", cfg.nTimeSteps);
        h5.write_attr_i32("Nx", cfg.numxCells + 1);
        h5.write_attr_i32("Ny", cfg.numyCells + 1);
    }

    const int64_t nI = cfg.nParticlesI, nE = cfg.nParticlesE;
    const int64_t loI = nI * rank / nranks, hiI = nI * (rank + 1) / nranks, loE = nE * rank / nranks, hiE = nE * (rank + 1) / nranks;
    const int64_t mI = hiI - loI, mE = hiE - loE;                   // this rank's particles
    const int nix = cfg.numxCells + 1, niy = cfg.numyCells + 1;
    picsp_params prm = {};
    prm.numxCells = cfg.numxCells; prm.numyCells = cfg.numyCells;
    prm.stepSize = cfg.stepSize; prm.timeStep = cfg.timeStep;
    prm.solverType = cfg.solverType; prm.flags = 0;
    prm.charge[0] = cfg.chargeE; prm.charge[1] = -cfg.chargeE;      // main.cpp:407-408
    prm.mass[0] = cfg.massI; prm.mass[1] = cfg.massE;
    prm.spwt[0] = cfg.ion_spwt; prm.spwt[1] = cfg.electron_spwt;
    prm.capacity[0] = mI; prm.capacity[1] = mE;
    prm.device = device;
    Guard g;
    HOST_CHECK(picsp_create(&prm, &g.c));

    {   // init(&ions, driftI, 0); init(&electrons, driftE, 0)  (main.cpp:437-438); RNG seeded 0 (main.cpp:49)
        Loader ld(0);
        for (int s = 0; s < 2; s++) {
            const int64_t n = s == 0 ? nI : nE;
            std::vector<double> x(n), y(n), vx(n), vy(n);
            ld.fill(cfg, s, x.data(), y.data(), vx.data(), vy.data());
            const int64_t lo = s == 0 ? loI : loE, m = s == 0 ? mI : mE;
            HOST_CHECK(picsp_species_upload(g.c, s, x.data() + lo, y.data() + lo, vx.data() + lo, vy.data() + lo, m));
        }
    }
    const std::string id_path = path + ".ncclid";
    if (nranks > 1) {
        char id[128];
        rc = exchange_unique_id(id_path, rank, id, err);
        if (rc != PICSP_OK) return rc;
        HOST_CHECK(picsp_comm_attach(g.c, id, rank, nranks));
        HOST_CHECK(picsp_comm_barrier(g.c));                        // rank 0 has created the file
        if (rank == 0) std::remove(id_path.c_str());
        else if (!h5.open(path, err, true)) return PICSP_ERR_INVALID;
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
    Pinned rows_i(4 * (size_t)mI), rows_e(4 * (size_t)mE), den_i(nn), den_e(nn), phi(nn), ke(2);
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
        uint64_t at = h5.reserve_dataset_f64("/particle.i/" + t, nI, 4);          // same offset on every rank
        if (!h5.write_rows(at, loI, mI, 4, rows_i.p)) return PICSP_ERR_INVALID;
        h5.write_dataset_f64("/den.i/" + t, den_i.p, nix, niy);                    // rank 0 (others only advance the offset)
        at = h5.reserve_dataset_f64("/particle.e/" + t, nE, 4);
        if (!h5.write_rows(at, loE, mE, 4, rows_e.p)) return PICSP_ERR_INVALID;
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
            HOST_CHECK(picsp_dump_begin(g.c, rows_i.p, rows_e.p, rank == 0 ? den_i.p : nullptr, rank == 0 ? den_e.p : nullptr,
                                        rank == 0 ? phi.p : nullptr, ke.p));
            pending_ts = upto;
        }
    }
    HOST_CHECK(flush_pending());
    if (nranks > 1) {
        if (rank != 0 && !h5.close(err)) return PICSP_ERR_INVALID;    // this rank's rows are on disk
        HOST_CHECK(picsp_comm_barrier(g.c));
        if (rank != 0) return PICSP_OK;
    }
    h5.write_dataset_f64("/timedata/energy", energy.data(), nT, 2);     // writeKE, main.cpp:1179-1188
    if (!h5.close(err)) return PICSP_ERR_INVALID;
    if (!quiet) {
        std::chrono::duration<double> diff = std::chrono::steady_clock::now() - start;
        std::cout << "Total time taken by PICSP: " << diff.count() << " s" << std::endl;   // main.cpp:558
    }
    return PICSP_OK;
}

}  // namespace picsp_host
