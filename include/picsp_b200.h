/* picsp_b200.h — the C ABI of the B200-native PICSP particle loop.
 *
 * This is the drop-in boundary for PICSP's per-timestep hot path.  The reference
 * (sayanadhikari/picsp) has no plugin/FFI layer: its hot path is the set of free
 * functions declared at src/main.cpp:202-231, which take raw `double*` grids and
 * `Species*`, read a handful of globals (`domain`, `timeStep`, `EPS`), return
 * void/bool and never throw.  Each entry point below replaces one of them and
 * cites it.  A maintainer of the reference binds this header directly from C++
 * (see INTEGRATION.md for the patch to src/main.cpp).
 *
 * Conventions
 *  - plain C, no CUDA/torch types in any signature; `extern "C"` linkage.
 *  - one opaque context per GPU rank; the library owns all device memory and one
 *    CUDA stream; the caller owns every host buffer it passes in.
 *  - every call returns PICSP_OK (0) or a negative error code; nothing throws.
 *    picsp_last_error() gives a human-readable message for the last failure.
 *  - calls are stream-ordered; the ones that fill host memory synchronise.
 *  - grids are nix*niy doubles, F[i][j] = F[i*niy + j] (y contiguous),
 *    nix = numxCells+1, niy = numyCells+1 (src/main.cpp:363,371,378-381).
 *  - species index: 0 = ions, 1 = electrons (src/main.cpp:407-412).
 *  - there is NO CPU fallback: without a CUDA device picsp_create fails with
 *    PICSP_ERR_NO_DEVICE.
 */
#ifndef PICSP_B200_H
#define PICSP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PICSP_B200_ABI_VERSION 1

typedef struct picsp_ctx picsp_ctx;

enum {
    PICSP_OK = 0,
    PICSP_ERR_INVALID = -1,     /* bad argument */
    PICSP_ERR_NO_DEVICE = -2,   /* no CUDA device / driver: there is no CPU path */
    PICSP_ERR_CUDA = -3,
    PICSP_ERR_CUFFT = -4,
    PICSP_ERR_NCCL = -5,
    PICSP_ERR_STATE = -6,       /* call sequence error (e.g. download before upload) */
    PICSP_ERR_DISPLACEMENT = -7,/* so many particles moved more than one particle tile (16 cells) in ONE step that the fixed-point
                                 * deposit could overflow (>= 2^(62 - fraction bits): 2048 for the sparsest load, millions for a
                                 * dense one), or a particle needed more than 64 consecutive re-pushes.  Fewer fast particles are
                                 * handled like any other, as in the reference. */
    PICSP_ERR_NOT_CONVERGED = -8/* SOR hit the reference's 200000-sweep cap (main.cpp:955) */
};

/* Grid selectors for picsp_grid_upload / picsp_grid_download. */
enum {
    PICSP_DEN_I = 0,  /* ions.den       (src/main.cpp:417) */
    PICSP_DEN_E = 1,  /* electrons.den  (src/main.cpp:420) */
    PICSP_RHO   = 2,  /* domain.rho     (src/main.cpp:381) */
    PICSP_PHI   = 3,  /* domain.phi     (src/main.cpp:378) */
    PICSP_EFX   = 4,  /* domain.efx     (src/main.cpp:379) */
    PICSP_EFY   = 5   /* domain.efy     (src/main.cpp:380) */
};

enum {
    PICSP_SOLVER_SPECTRAL = 1,  /* solverType 1: spectralPotentialSolver (src/main.cpp:960) */
    PICSP_SOLVER_SOR = 2        /* solverType 2: solvePotential          (src/main.cpp:904) */
};

/* Flags.  0 = the reference's behaviour, bug for bug. */
enum {
    PICSP_FLAG_CLEAR_DENSITY = 1 << 0,  /* extension: zero den before each deposit (the reference never does, main.cpp:692) */
    PICSP_FLAG_NO_SORT       = 1 << 1,  /* keep particles in upload order (no periodic tile sort) */
    PICSP_FLAG_NO_FUSE       = 1 << 2,  /* push does not pre-accumulate the next step's deposit */
    PICSP_FLAG_SOR_SINGLE_CTA = 1 << 3, /* SOR: use the single-CTA anti-diagonal kernel for every sweep (cross-check path) */
    PICSP_FLAG_SEPARATE_SORT  = 1 << 4, /* periodic re-binning as a stand-alone pass instead of inside the mover (cross-check path) */
    PICSP_FLAG_NO_GRAPH       = 1 << 5, /* picsp_step never replays captured CUDA graphs (cross-check path) */
    /* EXTENSION WITHOUT REFERENCE SEMANTICS (BASELINE.json config 3 "bounded domain with wall boundaries"): the reference
     * is periodic-only; its README says "bounded" and main.cpp:826-843 / :1064-1108 are commented-out sketches of
     * absorbing walls and a Dirichlet solver.  With this flag: no periodic fold of the densities; phi = 0 on the four
     * walls and a red-black Gauss-Seidel/SOR iteration on the interior, run to an L2 residual of 1e-12 (solverType is
     * ignored); E from central differences inside and full one-sided differences on the walls; a particle that leaves
     * the box is absorbed (position NaN, velocity 0 from then on: skip NaN rows in a download).  Checked against the
     * repo's own CPU restatement oracle/walls_check.c, which is NOT the reference.  Needs the tiled store
     * (not combinable with PICSP_FLAG_NO_SORT); usually combined with PICSP_FLAG_CLEAR_DENSITY. */
    PICSP_FLAG_WALLS          = 1 << 6,
    PICSP_FLAG_NCCL_ONLY      = 1 << 7, /* sharded runs: sum the partial rho with ncclAllReduce instead of the library's own peer-memory kernels (cross-check path) */
    PICSP_FLAG_CUFFT_ONLY     = 1 << 8, /* spectral solver: always cuFFT (cross-check path; by default node counts with a prime factor > 127, e.g. 2049 = 3 * 683, use the library's own shared-memory transform) */
    PICSP_FLAG_OWN_FFT        = 1 << 9  /* spectral solver: always the library's own transform, whatever the node count (cross-check path for small grids) */
};

/* Normalised quantities, i.e. the reference's globals after parse_ini_file
 * (src/main.cpp:279-294). */
typedef struct picsp_params {
    int32_t numxCells;        /* grid:numxCells  */
    int32_t numyCells;        /* grid:numyCells  */
    double  stepSize;         /* normalised dx = dy (main.cpp:289,364,372) */
    double  timeStep;         /* normalised dt (main.cpp:288) */
    int32_t solverType;       /* PICSP_SOLVER_* */
    int32_t flags;            /* PICSP_FLAG_* */
    double  charge[2];        /* +chargeE, -chargeE (main.cpp:407-408) */
    double  mass[2];          /* massI, massE */
    double  spwt[2];          /* ion_spwt, electron_spwt (main.cpp:401-402) */
    int64_t capacity[2];      /* max particles of each species held by THIS rank */
    int32_t device;           /* CUDA device ordinal */
    int32_t parts;            /* 0 (default): automatic.  Parts a species' store is split into (contiguous ranges of its
                               * particles, each binned on its own, all depositing into the species' one grid) so that the
                               * second buffer set of the re-binning is ONE part-sized spare instead of a copy of the whole
                               * state: automatic = 1 unless the device cannot hold state + copy (e.g. 4e9 particles on one
                               * 180 GB GPU -> 8).  Results do not depend on it (integer accumulation: bit-identical grids
                               * and phase space).  A part holds at most 2^32-1 particles. */
} picsp_params;

/* ---- lifetime -------------------------------------------------------------- */
int  picsp_abi_version(void);
int  picsp_create(const picsp_params *params, picsp_ctx **out);
void picsp_destroy(picsp_ctx *ctx);
const char *picsp_last_error(void);
int  picsp_sync(picsp_ctx *ctx);

/* ---- state exchange (host buffers; these are what parity tests use) --------- */
/* Replaces filling Species::part_list (src/main.cpp:140,594,613): n particles in list order.  The host buffers are
 * free again when the call returns; the first binning of the new load is left running on the device, so that it
 * overlaps the upload of the other species (every later call is ordered after it). */
int picsp_species_upload(picsp_ctx *ctx, int species, const double *x, const double *y,
                         const double *vx, const double *vy, int64_t n);
/* Particles come back in upload (list) order regardless of any internal re-binning (any of the four outputs may
 * be NULL).  Host-synchronous. */
int picsp_species_download(picsp_ctx *ctx, int species, double *x, double *y, double *vx, double *vy);
int picsp_species_count(picsp_ctx *ctx, int species, int64_t *n);
/* Same, [n][4] rows of {x, y, vx, vy}: the layout writeSpecies dumps (src/main.cpp:1152-1162). */
int picsp_species_download_rows(picsp_ctx *ctx, int species, double *rows);
int picsp_grid_upload(picsp_ctx *ctx, int which, const double *host);
/* With a communicator attached, PICSP_DEN_I / PICSP_DEN_E are the SUM over ranks of the per-rank partial densities
 * (each rank accumulates the deposit of its own particles): the call is then a collective, every rank must make it and
 * every rank receives the global density.  rho, phi, efx, efy are global on every rank already. */
int picsp_grid_download(picsp_ctx *ctx, int which, double *host);

/* Asynchronous dump: everything the reference writes at a diagnostics step (writeSpecies x2, writePot, computeKE x2;
 * src/main.cpp:507-527) without stalling the time loop.  picsp_dump_begin snapshots, stream-ordered after the steps
 * enqueued so far, the phase space of both species as [n][4] rows {x, y, vx, vy} in list order, den.i, den.e, phi and
 * the two kinetic energies on the device, and starts copying them into the caller's buffers on a second stream; it
 * returns at once, and steps enqueued afterwards run concurrently with the copies.  The buffers must stay valid (and
 * unread) until picsp_dump_wait returns; pass page-locked memory (picsp_host_alloc) or the copies will block the call.
 * Any pointer may be NULL (that item is skipped).  With a communicator attached both calls are collectives: den.i /
 * den.e are the SUM over ranks and arrive on rank 0 only (other ranks may pass NULL), phi is written on rank 0 only,
 * ke2 is the global value on every rank, rows are this rank's particles.  If the device has no room for the snapshot
 * (4*capacity doubles per species) the dump is done synchronously inside picsp_dump_begin. */
int picsp_dump_begin(picsp_ctx *ctx, double *rows_i, double *rows_e, double *den_i, double *den_e, double *phi, double *ke2);
int picsp_dump_wait(picsp_ctx *ctx);
/* Page-locked host memory for the asynchronous paths (NULL on failure). */
void *picsp_host_alloc(size_t bytes);
void picsp_host_free(void *p);

/* ---- the hot path, one entry per reference function -------------------------- */
int picsp_deposit(picsp_ctx *ctx, int species);     /* scatterSpecies            src/main.cpp:684-721 (+scatter :655-668) */
int picsp_compute_rho(picsp_ctx *ctx);              /* computeRho                src/main.cpp:869-901; all-reduces when a communicator is attached */
int picsp_solve(picsp_ctx *ctx);                    /* the solverType switch     src/main.cpp:492-497 */
int picsp_solve_spectral(picsp_ctx *ctx);           /* spectralPotentialSolver   src/main.cpp:960-1058 */
int picsp_solve_sor(picsp_ctx *ctx, int64_t *sweeps, double *l2); /* solvePotential src/main.cpp:904-957 (outputs may be NULL) */
/* Sweeps and residual of the last iterative solve (periodic SOR or the PICSP_FLAG_WALLS red-black SOR; negative sweeps:
 * the cap was hit).  Host-synchronous. */
int picsp_solve_status(picsp_ctx *ctx, int64_t *sweeps, double *l2);
int picsp_compute_ef(picsp_ctx *ctx);               /* computeEF                 src/main.cpp:1111-1139 */
int picsp_push(picsp_ctx *ctx, int species);        /* pushSpecies (+gather)     src/main.cpp:772-847, :671-681 */
int picsp_rewind(picsp_ctx *ctx, int species);      /* rewindSpecies             src/main.cpp:850-866 */
int picsp_bootstrap(picsp_ctx *ctx);                /* pre-loop sequence         src/main.cpp:453-472 */
int picsp_step(picsp_ctx *ctx, int nsteps);         /* nsteps bodies of the loop src/main.cpp:481-504 (dead scatterSpeciesVel omitted) */

/* ---- diagnostics ------------------------------------------------------------- */
int picsp_compute_ke(picsp_ctx *ctx, int species, double *ke);        /* computeKE  src/main.cpp:1190-1203 (summed over ranks when a communicator is attached) */
int picsp_delta_phi(picsp_ctx *ctx, double *max_phi, double *phi0);   /* max(phi), phi[0]  src/main.cpp:509-516 */
int picsp_repush_count(picsp_ctx *ctx, int species, int64_t *n);      /* extra pushes done by the last picsp_push (main.cpp:807-845); with PICSP_FLAG_WALLS: particles it absorbed */
int picsp_straggler_count(picsp_ctx *ctx, int species, int64_t *n);   /* particles of the last push/deposit that fell outside their tile window */
/* Steps between two tile sorts of a species (default: electrons 8, ions 96; <= 0 restores the default). */
int picsp_set_sort_period(picsp_ctx *ctx, int species, int period);
/* Steps between two orderings of the particles by CELL inside their tile (makes the mover's field gathers broadcast and
 * its deposits warp-aggregated; storage order only, results are bit-identical).  0 = never (the default: measured
 * on B200 the ordered store halves the shared-memory traffic of the mover but the aggregation costs as many issue slots
 * as it saves, profiles/r02_mover_aggregation.md). */
int picsp_set_cell_sort_period(picsp_ctx *ctx, int species, int period);
/* Bank order inside the mover's work items (chunks of <= 4096 particles of one tile): after every (re-)binning of the
 * species its particles are permuted, in place, so that the lanes of a warp sit on distinct shared-memory banks when
 * they gather E and deposit (no bank conflicts, no extra instruction in the mover).  -1 automatic (default: species that
 * are re-binned every >= 32 steps, i.e. ions, whose particles keep their cells in between), 0 off, 1 on.  Storage order
 * only: results are bit-identical. */
int picsp_set_bank_order(picsp_ctx *ctx, int species, int mode);
/* Warp-aggregated deposit (lanes of a warp whose particles share a cell combine their weights with REDUX before touching
 * shared memory): -1 automatic (default: on for a cell-ordered store and for loads concentrated on few bins, such as the
 * reference's diagonal loadType 2), 0 off, 1 on.  Speed only: integer accumulation makes the result bit-identical. */
int picsp_set_deposit_aggregation(picsp_ctx *ctx, int species, int mode);

/* ---- multi-GPU: particles sharded by index range, grid replicated ------------ */
/* One communicator per rank; id is an NCCL unique id (128 bytes) created by rank 0
 * with picsp_comm_unique_id and distributed by the caller (e.g. torch.distributed).
 * When all ranks sit on one NVLink node, picsp_comm_attach also maps the ranks' rho buffers into each other (CUDA IPC) and
 * picsp_step sums the partial densities with the library's own reduce-scatter + all-gather kernel over peer memory;
 * otherwise, and for the per-function calls, NCCL does it. */
int picsp_comm_unique_id(void *id128);
int picsp_comm_attach(picsp_ctx *ctx, const void *id128, int rank, int nranks);
/* Collective: returns when every rank's library stream has reached this point (a plain stream synchronise without a
 * communicator). */
int picsp_comm_barrier(picsp_ctx *ctx);
/* 1 when picsp_step sums the partial rho with the library's own peer-memory kernels, 0 when NCCL does it. */
int picsp_comm_peer_reduction(picsp_ctx *ctx);

/* ---- bench-only synthetic loader (NOT reference behaviour) ------------------- */
/* Fills n particles of a species on the device: positions uniform in the box,
 * velocities vth*sqrt(2)*(r1+r2+r3-1.5) (+/- drift alternating in x), counter-based
 * RNG keyed by (seed, global particle index = first_index + p). */
int picsp_species_fill_synthetic(picsp_ctx *ctx, int species, int64_t n, int64_t first_index,
                                 uint64_t seed, double vth, double xdrift);

/* ---- instrumentation ---------------------------------------------------------- */
enum {
    PICSP_PHASE_DEPOSIT = 0,   /* standalone deposit kernels + finalize/fold */
    PICSP_PHASE_RHO = 1,
    PICSP_PHASE_ALLREDUCE = 2,
    PICSP_PHASE_SOLVE = 3,
    PICSP_PHASE_EF = 4,
    PICSP_PHASE_PUSH = 5,      /* the mover (fused with the next deposit unless NO_FUSE) */
    PICSP_PHASE_SORT = 6,
    PICSP_PHASE_STEP = 7,      /* one whole picsp_step() call, first launch to last */
    PICSP_PHASE_PUSH_IONS = 8, /* the mover launches of species 0 alone (also counted in PICSP_PHASE_PUSH) */
    PICSP_PHASE_PUSH_ELECTRONS = 9,
    PICSP_PHASE_COUNT = 10
};
int picsp_profile_enable(picsp_ctx *ctx, int on);                       /* CUDA events around every phase on the library's stream */
int picsp_profile_get(picsp_ctx *ctx, int phase, double *ms, int64_t *calls); /* synchronises; accumulates since last reset */
int picsp_profile_reset(picsp_ctx *ctx);
/* How the library's own DFT would split a transform of length M (no device needed): out5 = {kind, P, Q, L, largest prime
 * factor of M}; kind 1 = two direct coprime factors P x Q (both <= 64), kind 0 = P direct (<= 32) x Q by Bluestein on the
 * power-of-two length L, kind -1 = does not fit shared memory (cuFFT is used). */
int picsp_fft_plan_query(int M, int32_t *out5);
/* Transform behind picsp_solve_spectral: 1 = the library's own shared-memory DFT (node counts with a prime factor > 127 such
 * as 2049 = 3 * 683, and small grids), 0 = cuFFT. */
int picsp_spectral_engine(picsp_ctx *ctx, int *own);
/* Parts every species' store is split into (picsp_params::parts resolved; 1 = the whole species in one store). */
int picsp_parts(picsp_ctx *ctx, int *parts);
int picsp_kernel_launches(picsp_ctx *ctx, int64_t *n);                  /* number of this library's kernels launched so far */

#ifdef __cplusplus
}
#endif
#endif /* PICSP_B200_H */
