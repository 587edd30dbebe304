// picsp_b200/csrc/ctx.cuh — internal context and helpers (not part of the ABI).
#pragma once

#include <cuda_runtime.h>
#include <cufft.h>
#include <nccl.h>      // types only; the library is dlopen'ed in comm.cuh
#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges are no-ops unless a profiler injects itself

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/picsp_b200.h"

namespace picsp {

// ---------------------------------------------------------------------------
// error plumbing: internal code throws, the extern "C" layer converts to codes
// ---------------------------------------------------------------------------
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

#define PICSP_CUDA(expr)                                                                      \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess)                                                               \
            throw ::picsp::Error(PICSP_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
    } while (0)

#define PICSP_CUFFT(expr)                                                                     \
    do {                                                                                      \
        cufftResult r__ = (expr);                                                             \
        if (r__ != CUFFT_SUCCESS)                                                             \
            throw ::picsp::Error(PICSP_ERR_CUFFT, std::string(#expr) + ": cufft error " + std::to_string((int)r__)); \
    } while (0)

#define PICSP_REQUIRE(cond, code, msg)                                  \
    do {                                                                \
        if (!(cond)) throw ::picsp::Error((code), (msg));               \
    } while (0)

// ---------------------------------------------------------------------------
// geometry (reference globals `domain`, `timeStep`; src/main.cpp:363-375)
// ---------------------------------------------------------------------------
constexpr int TILE = 16;   // particle-tile edge in cells (binning, fixed-point bound, smem windows)

struct Geom {
    int nix, niy;          // nodes
    int ncx, ncy;          // cells
    int ntx, nty;          // particle tiles
    long long nn;          // nix*niy
    long long guard;       // zero guard band (elements) either side of E, reference UB -> defined zero
    double dx;             // == dy
    double dt;
    double xl, yl;         // == xmax, ymax since x0 = y0 = 0
};

// ---------------------------------------------------------------------------
// per-species device store: structure of arrays, double precision
// ---------------------------------------------------------------------------
// A species may be split into several PARTS (contiguous ranges of its particles in upload order), each a store of its
// own — bins, chunk table, histogram, slot map — that deposit into the species' one accumulator grid.  Parts exist so
// that the second buffer set of the re-binning (the sort destination) can be ONE part-sized spare shared by all
// parts of both species instead of a full copy of the state: 4e9 particles (144 GB with their slot maps) then fit one
// 180 GB device.  Every part is a Species object; the species-level members (q, m, spwt, den, acc, frac, ...) of the
// parts >= 1 alias those of part 0 (= picsp_ctx::sp[s]).  One part (the normal case) is exactly the old layout.
struct Species {
    int s = 0, part = 0;           // species index, part index
    int64_t first = 0;             // index (in the species' upload order) of this part's first particle
    bool shares_spare = false;     // x2.. / id2 are borrowed from picsp_ctx::spare (several parts) instead of owned
    double *x = nullptr, *y = nullptr, *vx = nullptr, *vy = nullptr;
    uint32_t *id = nullptr;        // slot -> index in upload order; nullptr == identity
    int64_t n = 0, cap = 0;        // particles / capacity of THIS part
    double q = 0, m = 0, spwt = 0;
    double *den = nullptr;         // nn, accumulating
    long long *acc = nullptr;      // nn, fixed-point deposit accumulator (order-independent => deterministic)
    int *frac = nullptr;           // device [2]: fixed-point fraction bits acc was filled with; warp-aggregated deposit on/off
    unsigned long long *frac_scratch = nullptr;   // [0] running max neighbourhood population, [1] CTA ticket, [2] running max bin population (k_frac_from_hist)
    unsigned int *hist = nullptr;      // particles per tile at the positions currently stored
    unsigned int *hist_next = nullptr; // filled by the mover for the positions it writes
    bool hist_valid = false;
    bool acc_valid = false;        // acc holds the deposit of the stored positions (filled by the fused mover)
    unsigned long long *counters = nullptr; // device: [0] extra pushes of the last push, [1] out-of-window deposits

    // tile binning (fast path)
    double *x2 = nullptr, *y2 = nullptr, *vx2 = nullptr, *vy2 = nullptr;   // sort destination (lazily allocated)
    uint32_t *id2 = nullptr;
    bool has_perm = false;         // id[] holds a permutation (otherwise identity)
    bool sorted = false;           // particles are binned by tile and chunks/tile_off describe them
    int steps_since_sort = 0;
    int sort_period = 8;
    long long *tile_off = nullptr; // ntiles+1
    void *chunks = nullptr;        // picsp::Chunk[max_chunks]
    int *nchunks = nullptr;        // device scalar
    void *chunks2 = nullptr;       // second chunk table (a re-sort reads the current one while the next is built)
    int *nchunks2 = nullptr;
    unsigned int *cursor = nullptr;// ntiles
    long long max_chunks = 0;
    int chunk = 4096;              // particles per CTA work item of the current binning (pick_chunk)
    int chunk2 = 4096;             // ... of the binning under construction
    unsigned int *chunk_cnt = nullptr;   // [max_chunks][9] per-chunk neighbour-bin populations left by the last fused mover launch
    unsigned int *chunk_base = nullptr;  // [max_chunks][9] ranges reserved from them for a re-binning launch
    bool cnt_valid = false;        // chunk_cnt describes the stored positions under the current chunk table
    bool staged_v_valid = false;   // vx2/vy2 hold the current velocities in upload order (left there by a download)
    // cell order inside a bin (k_cell_count / k_cell_scan / k_cell_permute)
    int cell_period = 0;           // steps between two cell orderings (0: never)
    int aggregate = -1;            // warp-aggregated deposit: -1 automatic (cell-ordered store or concentrated load), 0 off, 1 on
    int steps_since_cellsort = 0;
    int bank_order = -1;           // bank order inside the chunks after every (re-)binning (k_bank_order): -1 automatic (slow species), 0 off, 1 on
    unsigned int *cell_cnt = nullptr;    // [chunks][CELLKEYS] populations, then first slots
    long long cell_cnt_chunks = 0;
    int *tile_chunk0 = nullptr;          // [ntiles] first chunk of every bin
    int *scan_chunk0 = nullptr;          // [ntiles] the same for the chunk table under construction (k_scan_tiles -> k_fill_chunks)
    int ntiles = 0;
};

struct PhaseTimer {
    std::vector<cudaEvent_t> pool;
    size_t used = 0;
    double ms = 0.0;
    int64_t calls = 0;
};

}  // namespace picsp

struct picsp_ctx {
    picsp_params prm;
    picsp::Geom g;
    cudaStream_t stream = nullptr;
    picsp::Species sp[2];                        // part 0 of every species (the whole species when nparts == 1)
    int nparts = 1;                              // parts per species (picsp_params::reserved, 0 = as many as the device memory asks for)
    std::vector<picsp::Species> more[2];         // parts 1 .. nparts-1
    int64_t n_total[2] = {0, 0}, cap_total[2] = {0, 0};
    int64_t part_cap = 0;                        // capacity of every part (all parts of both species are the same size, so any set can be the spare)
    struct Spare { double *x = nullptr, *y = nullptr, *vx = nullptr, *vy = nullptr; uint32_t *id = nullptr; } spare;   // nparts > 1: THE second buffer set
    unsigned int *hist_sum[2] = {nullptr, nullptr};   // nparts > 1: sum of the parts' tile histograms (fixed-point scale)
    double *d_part_sums = nullptr;               // nparts > 1: per-part partial results of reductions

    double *rho = nullptr, *phi = nullptr;
    double2 *E_alloc = nullptr, *E = nullptr;     // interleaved {efx, efy}; E = E_alloc + guard
    cufftHandle plan_fwd = 0, plan_inv = 0;
    bool have_plans = false;
    cufftDoubleComplex *rhok = nullptr, *phik = nullptr;
    // own shared-memory DFT for node counts with a large prime factor (fft_kernels.cuh); plan 0: length niy (rows),
    // plan 1: length nix (columns; the same tables when nix == niy)
    struct OwnFft {
        bool on = false;
        int kind[2] = {0, 0}, M[2] = {0, 0}, P[2] = {0, 0}, Q[2] = {0, 0}, L[2] = {0, 0}, logL[2] = {0, 0};
        void *chirp[2] = {}, *bhat[2] = {}, *tw[2] = {}, *in_pos[2] = {}, *out_idx[2] = {}, *wp[2] = {}, *rootP[2] = {}, *rootQ[2] = {};
        size_t smem[2] = {0, 0};
    } fft;

    // small device scratch
    double *d_red = nullptr;        // reduction partials
    double *d_scalars = nullptr;    // [0..7] results (ke, maxphi, phi0, sor l2, ...)
    long long *d_sor_status = nullptr;
    int *d_sor_progress = nullptr;  // per-band column progress of the pipelined SOR sweep
    double *d_walls_partial = nullptr;   // PICSP_FLAG_WALLS: per-CTA residual partials of k_rb_sor
    int walls_grid = 0;             // its cooperative grid (co-resident CTAs)
    double walls_omega = 1.0;
    int *d_error = nullptr;         // sticky device-side error flag
    double *h_pinned = nullptr;     // small pinned staging for scalar read-backs

    // TMA descriptor of E viewed as [nix][2*niy] doubles (128 bytes, CUtensorMap)
    alignas(64) unsigned char tmapE[128];
    bool have_tmap = false;
    bool smem_opted_in = false;
    bool hist_smem_opted_in = false;
    bool sort2_opted_in = false;
    bool cellsort_opted_in = false;
    bool bankorder_opted_in = false;
    bool sor_smem_opted_in = false;

    // staging for grid component uploads/downloads
    double *stage = nullptr; int64_t stage_cap = 0;
    // particle downloads: device->host copies run on their own stream so that the un-permute of the next array
    // overlaps the copy of the previous one
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_ready[4] = {nullptr, nullptr, nullptr, nullptr};
    // busy[s]: kernels that may touch species s have been enqueued on `stream` since it was last synchronised
    // (set conservatively by every launch; an upload of species s only has to wait when busy[s])
    bool busy[2] = {false, false};

    // Small populations are launch-bound (~20 launches per step, 0.15-0.35 ms per step): between re-binnings two
    // consecutive steps (the histogram buffers alternate, so a PAIR returns to the same pointers) are captured once
    // into a CUDA graph and replayed.  The key lists everything the captured launches have baked in.
    // A re-binning swaps buffer sets and chunk tables, so the same few keys recur: the instantiated graphs are cached.
    struct StepGraphKey { const void *ptr[14]; long long n[2]; int chunk[2]; };
    struct StepGraph { StepGraphKey key; cudaGraphExec_t exec; int64_t launches; };
    static constexpr int STEP_GRAPH_CACHE = 8;
    StepGraph step_graphs[STEP_GRAPH_CACHE] = {};
    int step_graph_count = 0, step_graph_next = 0;
    std::vector<cudaGraphExec_t> retired_graphs;    // evicted while possibly still in flight: destroyed at the next synchronisation

    // multi-GPU
    ncclComm_t comm = nullptr; int rank = 0, nranks = 1;
    // peer-memory reduction of the partial rho (peer_kernels.cuh): one cudaMalloc block per rank, mapped by every
    // other rank of the node through CUDA IPC:  [ partial rho (nn) | summed rho (nn) | arrival flags ]
    bool peer_ok = false;
    double *peer_block = nullptr;                // this rank's block
    void *peer_mapped[8] = {};                   // the others' blocks as mapped here (nullptr for this rank)
    double *peer_part[8] = {}, *peer_full[8] = {};
    unsigned long long *peer_flags[8] = {};
    unsigned long long peer_epoch = 0;
    double *rho_owned = nullptr;                 // the context's own rho allocation (c->rho points into peer_block while peer_ok)

    // asynchronous dumps (picsp_dump_begin / picsp_dump_wait): device-side snapshot of what the reference dumps
    // (writeSpecies / writePot, src/main.cpp:1142-1216), copied out on the copy stream while the time loop goes on
    double *snap_rows[2] = {nullptr, nullptr};   // [n][4] rows in upload order
    int64_t snap_rows_cap[2] = {0, 0};
    double *snap_grids = nullptr;                // den_i | den_e | phi, nn each
    double *snap_ke = nullptr;                   // [2] sum(vx^2+vy^2) per species (summed over ranks)
    cudaEvent_t ev_snap = nullptr, ev_dump_done = nullptr, ev_dump_done2 = nullptr;
    cudaStream_t copy_stream2 = nullptr;         // second copy engine for the electrons' rows
    bool dump_in_flight = false;
    double *dump_ke_host = nullptr;              // caller's [2]: the Q10 constant is added in picsp_dump_wait
    bool snapshot_unavailable = false;           // not enough memory for a snapshot: dumps are synchronous

    // device-side error flag mirrored into mapped host memory at the start of every step, so that picsp_step can
    // report a violation of the previous steps without synchronising
    int *h_error_mapped = nullptr;
    bool graphs_disabled = false;                // a capture / instantiation failed once: plain launches from then on

    // instrumentation
    bool profiling = false;
    picsp::PhaseTimer timers[PICSP_PHASE_COUNT];
    int64_t launches = 0;
    int num_sms = 148;
};

namespace picsp {

// RAII phase scope: CUDA events on the library's stream when profiling is on.
inline const char *phase_name(int ph) {
    static const char *names[PICSP_PHASE_COUNT] = {"picsp:deposit", "picsp:rho", "picsp:allreduce", "picsp:solve", "picsp:ef", "picsp:push",
                                                    "picsp:sort", "picsp:step", "picsp:push_ions", "picsp:push_electrons"};
    return ph >= 0 && ph < PICSP_PHASE_COUNT ? names[ph] : "picsp";
}

struct PhaseScope {
    picsp_ctx *c; int phase; cudaEvent_t e0 = nullptr, e1 = nullptr;
    PhaseScope(picsp_ctx *ctx, int ph) : c(ctx), phase(ph) {
        nvtxRangePushA(phase_name(ph));          // NVTX range per phase (nsys / ncu --nvtx timelines)
        if (!c->profiling) return;
        PhaseTimer &t = c->timers[phase];
        while (t.pool.size() < t.used + 2) {
            cudaEvent_t e; PICSP_CUDA(cudaEventCreate(&e)); t.pool.push_back(e);
        }
        e0 = t.pool[t.used]; e1 = t.pool[t.used + 1]; t.used += 2;
        PICSP_CUDA(cudaEventRecord(e0, c->stream));
    }
    ~PhaseScope() {
        if (e1) cudaEventRecord(e1, c->stream);
        if (c->profiling) c->timers[phase].calls++;
        nvtxRangePop();
    }
};

inline void profile_collect(picsp_ctx *c) {
    for (int p = 0; p < PICSP_PHASE_COUNT; p++) {
        PhaseTimer &t = c->timers[p];
        for (size_t i = 0; i + 1 < t.used; i += 2) {
            float ms = 0.f;
            PICSP_CUDA(cudaEventSynchronize(t.pool[i + 1]));
            PICSP_CUDA(cudaEventElapsedTime(&ms, t.pool[i], t.pool[i + 1]));
            t.ms += ms;
        }
        t.used = 0;
    }
}

inline int blocks_for(long long n, int threads, int max_blocks) {
    long long b = (n + threads - 1) / threads;
    if (b < 1) b = 1;
    if (b > max_blocks) b = max_blocks;
    return (int)b;
}

#define PICSP_LAUNCH(ctx, kernel, grid, block, smem, ...)                                   \
    do {                                                                                    \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                    \
        (ctx)->launches++;                                                                  \
        (ctx)->busy[0] = (ctx)->busy[1] = true;                                             \
        PICSP_CUDA(cudaGetLastError());                                                     \
    } while (0)

}  // namespace picsp
