/* picsp_b200_host.h — C entry points of the C++ host driver that sits above the kernel ABI
 * (picsp_b200.h).  These restate the parts of the reference's `main` that surround the hot
 * path, so that `picsp_b200_run <input.ini>` is a drop-in for `./picsp <input.ini>`:
 *   - INI reading + unit normalisation + sanity gates + banner   src/main.cpp:240-331, 440-450
 *   - particle loader and RNG                                    src/main.cpp:49-54, 567-640
 *   - time loop, diagnostics cadence, HDF5 output layout         src/main.cpp:336-561, 1142-1247
 */
#ifndef PICSP_B200_HOST_H
#define PICSP_B200_HOST_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* The reference's globals after parse_ini_file (src/main.cpp:66-92, 252-294). */
typedef struct picsp_run_config {
    int32_t nTimeSteps, numxCells, numyCells, nParticlesI, nParticlesE, dumpPeriod, solverType, loadType;
    double timeStep, stepSize;            /* normalised (main.cpp:288-289) */
    double massI, massE, chargeE, density, vthE, vthI, driftE, driftI;   /* normalised (main.cpp:282-291) */
    double ion_spwt, electron_spwt;       /* main.cpp:293-294 */
    double omega_pe, Lambda_D;            /* main.cpp:280-281 */
} picsp_run_config;

/* Reads an ini file with iniparser's rules (lower-cased "section:key", value cut at the first
 * ';' or '#', ints via strtol base 0, doubles via atof, missing key -> -1) and normalises.
 * print_banner != 0 prints the reference's stdout lines (main.cpp:296-322).
 * Returns PICSP_ERR_INVALID where the reference would exit(EXIT_FAILURE) (main.cpp:323-326). */
int picsp_host_parse_ini(const char *path, picsp_run_config *out, int print_banner);

/* Every "section:key" <TAB> value <NEWLINE> pair the reader finds in a file (keys lower-cased, sorted), for checking the
 * reader against iniparser's own fixtures.  Returns the number of bytes needed (including the terminating NUL) or a
 * negative error code; writes at most buflen bytes. */
int64_t picsp_host_ini_dump(const char *path, char *buf, int64_t buflen);

/* The loader (main.cpp:567-617) with the reference's RNG (std::mt19937(seed) +
 * uniform_real_distribution<double>(0,1), main.cpp:49-54).  One loader carries the RNG state and the
 * loadType-2 recurrence across species; fill ions (species 0) first, then electrons, as main does. */
typedef struct picsp_loader picsp_loader;
picsp_loader *picsp_host_loader_create(uint32_t seed);
void picsp_host_loader_destroy(picsp_loader *ld);
int picsp_host_loader_fill(picsp_loader *ld, const picsp_run_config *cfg, int species,
                           double *x, double *y, double *vx, double *vy);

/* The whole program: parse, banner, load, bootstrap, time loop with the reference's diagnostics
 * (every 50 steps, main.cpp:507), HDF5 file with the reference's layout written to out_path
 * (NULL: "output/data.h5").  max_steps >= 0 truncates the loop (ts = 0 .. min(nTimeSteps, max_steps)).
 * quiet != 0 suppresses stdout. */
int picsp_host_run(const char *ini_path, const char *out_path, int max_steps, int quiet, int device);

/* The same program as one process per GPU: rank r of nranks owns the particles [N*r/nranks, N*(r+1)/nranks) of
 * each species in loader order, the grid is replicated, rank 0 owns the output file (den.i / den.e reduced over the ranks,
 * phi, energies, metadata) and every rank writes its own rows of the /particle.i and /particle.e datasets into it.  The NCCL unique id is passed
 * from rank 0 to the others through the file <out_path>.ncclid; all ranks must see the same out_path (one node). */
int picsp_host_run_ranked(const char *ini_path, const char *out_path, int max_steps, int quiet, int device, int rank, int nranks);

/* The HDF5 writer the driver uses, exposed for tests and tools: the subset of the format that
 * picsp's output needs (root attributes, first-level groups, contiguous f64 rank-2 datasets). */
typedef struct picsp_h5 picsp_h5;
picsp_h5 *picsp_host_h5_open(const char *path);
int picsp_host_h5_group(picsp_h5 *h, const char *abs_name);
int picsp_host_h5_dataset_f64(picsp_h5 *h, const char *abs_name, const double *data, uint64_t d0, uint64_t d1);
int picsp_host_h5_attr_f64(picsp_h5 *h, const char *name, double v);
int picsp_host_h5_attr_i32(picsp_h5 *h, const char *name, int32_t v);
int picsp_host_h5_close(picsp_h5 *h);   /* writes the metadata, closes and frees h */

#ifdef __cplusplus
}
#endif
#endif
