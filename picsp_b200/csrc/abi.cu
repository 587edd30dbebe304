// picsp_b200/csrc/abi.cu — the extern "C" boundary declared in include/picsp_b200.h.
//
// Host-side orchestration only: allocation, launch order, stream/event plumbing,
// NCCL.  All arithmetic is in the kernels (grid_kernels.cuh, particle_kernels.cuh).
// There is no CPU implementation of any operation in this library.
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <mutex>

#include "ctx.cuh"
#include "grid_kernels.cuh"
#include "particle_kernels.cuh"
#include "peer_kernels.cuh"
#include "tile_kernels.cuh"
#include "fft_kernels.cuh"

using namespace picsp;

namespace {

thread_local std::string g_last_error;

int fail(const Error &e) { g_last_error = e.what(); return e.code; }
int fail(int code, const std::string &m) { g_last_error = m; return code; }

#define PICSP_API_BEGIN try {
#define PICSP_API_END                                                                  \
    return PICSP_OK; }                                                                 \
    catch (const Error &e) { return fail(e); }                                         \
    catch (const std::exception &e) { return fail(PICSP_ERR_INVALID, e.what()); }      \
    catch (...) { return fail(PICSP_ERR_INVALID, "unknown exception"); }

void check_ctx(picsp_ctx *c) { PICSP_REQUIRE(c != nullptr, PICSP_ERR_INVALID, "null context"); }
void check_species(int s) { PICSP_REQUIRE(s == 0 || s == 1, PICSP_ERR_INVALID, "species must be 0 (ions) or 1 (electrons)"); }

PushConst push_const(const picsp_ctx *c, int s) {
    const Geom &g = c->g;
    const Species &sp = c->sp[s];
    PushConst pc;
    const double qm = sp.q / sp.m;                 // src/main.cpp:775
    pc.dx = g.dx; pc.dt = g.dt;
    pc.dtqm = g.dt * qm;                           // src/main.cpp:793
    pc.hdtqm = 0.5 * g.dt * qm;                    // src/main.cpp:863
    pc.xl = g.xl; pc.yl = g.yl;
    pc.nix = g.nix; pc.niy = g.niy; pc.ntx = g.ntx; pc.nty = g.nty;
    pc.nn = g.nn; pc.guard = g.guard;
    pc.walls = (c->prm.flags & PICSP_FLAG_WALLS) ? 1 : 0;
    pc.far_shift = 0;
    while ((1 << pc.far_shift) < c->nparts) pc.far_shift++;
    return pc;
}

template <class T> void dalloc(T **p, size_t count) {
    PICSP_CUDA(cudaMalloc((void **)p, std::max<size_t>(count, 1) * sizeof(T)));
}

// -- parts of a species (ctx.cuh) -----------------------------------------------------------
Species &part_of(picsp_ctx *c, int s, int p) { return p == 0 ? c->sp[s] : c->more[s][(size_t)p - 1]; }
template <class F> void for_parts(picsp_ctx *c, int s, F f) {
    for (int p = 0; p < c->nparts; p++) f(part_of(c, s, p));
}
// the shared spare (several parts): a part borrows it as its second buffer set and, after a sort that left the result
// there, keeps it and hands its old primary set back
void borrow_spare(picsp_ctx *c, Species &sp) {
    if (!sp.shares_spare) return;
    sp.x2 = c->spare.x; sp.y2 = c->spare.y; sp.vx2 = c->spare.vx; sp.vy2 = c->spare.vy; sp.id2 = c->spare.id;
}
void return_spare(picsp_ctx *c, Species &sp) {
    if (!sp.shares_spare) return;
    c->spare.x = sp.x2; c->spare.y = sp.y2; c->spare.vx = sp.vx2; c->spare.vy = sp.vy2; c->spare.id = sp.id2;
}

// One buffer set of a species = ONE allocation holding x, y, vx, vy back to back (x is its base).  +2: bulk slices
// are widened to even indices; the stride is a multiple of 32 doubles so every array stays 256-byte aligned.  The
// idle set doubles as a contiguous staging block of 4*cap doubles (row-layout dumps).
void alloc_particle_set(double **x, double **y, double **vx, double **vy, int64_t cap) {
    const size_t stride = (((size_t)cap + 2 + 31) / 32) * 32;
    dalloc(x, 4 * stride);
    *y = *x + stride; *vx = *x + 2 * stride; *vy = *x + 3 * stride;
}

int particle_blocks(const picsp_ctx *c, long long n, int threads) {
    return blocks_for(n, threads, c->num_sms * 16);
}

// -- device error flag ------------------------------------------------------------
void check_device_error(picsp_ctx *c) {
    int *h = reinterpret_cast<int *>(c->h_pinned + 8);
    PICSP_CUDA(cudaMemcpyAsync(h, c->d_error, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    PICSP_CUDA(cudaStreamSynchronize(c->stream));
    for (cudaGraphExec_t g : c->retired_graphs) cudaGraphExecDestroy(g);     // nothing is in flight any more
    c->retired_graphs.clear();
    if (*h) {
        int v = *h;
        PICSP_CUDA(cudaMemsetAsync(c->d_error, 0, sizeof(int), c->stream));
        *c->h_error_mapped = 0;
        if (v & ERR_BIT_PEER)
            throw Error(PICSP_ERR_NCCL, "a rank did not reach the peer-memory barrier of the rho reduction within 2 s");
        if (v & ERR_BIT_REBIN)
            throw Error(PICSP_ERR_STATE, "internal: a re-binning mover overflowed a bin (histogram and bin function disagree)");
        if (v & ERR_BIT_RUNAWAY)
            throw Error(PICSP_ERR_DISPLACEMENT, "a particle needed more than 64 consecutive re-pushes (non-finite or absurd velocity)");
        throw Error(PICSP_ERR_DISPLACEMENT, "too many particles moved by more than one particle tile (16 cells) in a single step: "
                                            "the fixed-point deposit could overflow (see far_mover() in particle_kernels.cuh)");
    }
}

// -- histogram / fixed-point scale --------------------------------------------------
void ensure_hist(picsp_ctx *c, Species &sp) {
    const int s = sp.s;
    if (sp.hist_valid) return;
    const int nt = c->g.ntx * c->g.nty;
    PICSP_CUDA(cudaMemsetAsync(sp.hist, 0, sizeof(unsigned int) * nt, c->stream));
    if (sp.n > 0) {
        // bins in shared memory up to 16384 of them (64 KB: 2048^2 cells), else global atomics;
        // few, fat CTAs so that the flush (one atomic per non-zero bin and CTA) stays small beside the counting
        const int nsmem = nt <= 16384 ? nt : 0;
        if (nsmem > 12288 && !c->hist_smem_opted_in) {
            PICSP_CUDA(cudaFuncSetAttribute(k_tile_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
            c->hist_smem_opted_in = true;
        }
        const int blocks = nsmem ? (int)std::min<long long>(c->num_sms * 2, (sp.n + 16383) / 16384) : particle_blocks(c, sp.n, 256);
        PICSP_LAUNCH(c, k_tile_hist, std::max(blocks, 1), nsmem ? 1024 : 256, sizeof(unsigned int) * nsmem, sp.x, sp.y,
                     (long long)sp.n, push_const(c, s), sp.hist, nsmem);
    }
    sp.hist_valid = true;
}
bool tiled(const picsp_ctx *c) { return !(c->prm.flags & PICSP_FLAG_NO_SORT); }
bool walls(const picsp_ctx *c) { return (c->prm.flags & PICSP_FLAG_WALLS) != 0; }

// species-level: the scale bounds the sums of ALL parts (they deposit into one accumulator grid), so with several
// parts it is taken from the sum of their histograms (every part's histogram must be valid: ensure_hist)
void compute_frac(picsp_ctx *c, int s) {
    Species &sp = c->sp[s];
    // the shared-memory limbs of the tiled path carry at most MAX_FRAC_TILED fraction bits
    const int nt = c->g.ntx * c->g.nty;
    const unsigned int *hist = sp.hist;
    if (c->nparts > 1) {
        PICSP_CUDA(cudaMemsetAsync(c->hist_sum[s], 0, sizeof(unsigned int) * nt, c->stream));
        for_parts(c, s, [&](Species &q) {
            PICSP_LAUNCH(c, k_hist_add, (nt + 255) / 256, 256, 0, c->hist_sum[s], q.hist, nt);
        });
        hist = c->hist_sum[s];
    }
    PICSP_LAUNCH(c, k_frac_from_hist, (nt + 255) / 256, 256, 0, hist, c->g.ntx, c->g.nty, sp.frac,
                 tiled(c) ? MAX_FRAC_TILED : 60, sp.frac_scratch, (long long)c->n_total[s],
                 sp.aggregate == 0 ? -1 : ((sp.aggregate > 0 || sp.cell_period > 0) ? 1 : 0));
}

// -- tile binning -------------------------------------------------------------------------
int mover_grid(const Species &sp);
bool bank_order_on(const Species &sp);
void op_bank_order(picsp_ctx *c, Species &sp);

// Particles per CTA work item: CHUNK (4096) for big populations; smaller when there are too few particles to give
// every SM several waves of CTAs (tail effect), never below 512, always a multiple of the slice size.
int pick_chunk(const picsp_ctx *c, int64_t n) {
    const int64_t want_ctas = (int64_t)c->num_sms * MOVER_MIN_CTAS * 6;
    int64_t ch = (n + want_ctas - 1) / want_ctas;
    ch = ((ch + MOVER_THREADS - 1) / MOVER_THREADS) * MOVER_THREADS;
    return (int)std::min<int64_t>(CHUNK, std::max<int64_t>(512, ch));
}

// allocates the second buffer set, scans the histogram of the stored positions into the NEW bin offsets and
// chunk table (second table: the kernels that move the particles still walk the current one), zeroes the cursors
void sort_prepare(picsp_ctx *c, Species &sp) {
    const Geom &g = c->g;
    const int nt = g.ntx * g.nty;
    ensure_hist(c, sp);
    borrow_spare(c, sp);
    if (!sp.x2) alloc_particle_set(&sp.x2, &sp.y2, &sp.vx2, &sp.vy2, sp.cap);
    if (!sp.id) dalloc(&sp.id, sp.cap + 8);        // +8: bulk slices of ids are widened to multiples of 4
    if (!sp.id2) dalloc(&sp.id2, sp.cap + 8);
    if (!sp.chunks2) {
        dalloc((Chunk **)&sp.chunks2, (size_t)sp.max_chunks); dalloc(&sp.nchunks2, 1);
        dalloc(&sp.chunk_cnt, (size_t)sp.max_chunks * 9); dalloc(&sp.chunk_base, (size_t)sp.max_chunks * 9);
    }
    sp.chunk2 = pick_chunk(c, sp.n);
    if (!sp.scan_chunk0) dalloc(&sp.scan_chunk0, (size_t)nt);
    PICSP_LAUNCH(c, k_scan_tiles, 1, 1024, 0, sp.hist, nt, sp.tile_off, sp.scan_chunk0, sp.nchunks2, sp.cursor, sp.chunk2);
    PICSP_LAUNCH(c, k_fill_chunks, (nt * 32 + 255) / 256, 256, 0, sp.hist, nt, sp.tile_off, sp.scan_chunk0, (Chunk *)sp.chunks2, sp.chunk2);
}
void sort_finish(picsp_ctx *c, Species &sp, bool result_in_second_set = true) {
    std::swap(sp.chunks, sp.chunks2); std::swap(sp.nchunks, sp.nchunks2); sp.chunk = sp.chunk2;
    if (result_in_second_set) {
        std::swap(sp.x, sp.x2); std::swap(sp.y, sp.y2); std::swap(sp.vx, sp.vx2); std::swap(sp.vy, sp.vy2);
        std::swap(sp.id, sp.id2);
        return_spare(c, sp);       // several parts: the old primary set is the spare of whoever sorts next
    }
    sp.has_perm = true; sp.sorted = true; sp.steps_since_sort = 0;
    sp.cnt_valid = false;          // new chunk table
    sp.staged_v_valid = false;     // the staging buffers are now the live ones
}

void op_sort(picsp_ctx *c, Species &sp) {
    PhaseScope ph(c, PICSP_PHASE_SORT);
    const int s = sp.s;
    sort_prepare(c, sp);
    const uint32_t *ids = sp.has_perm ? sp.id : (const uint32_t *)nullptr;
    const int nt = c->g.ntx * c->g.nty;
    if (!sp.sorted && sp.n >= 100000 && nt >= 256) {
        // first binning of a large arbitrary load: coarse pass into the second set, fine pass back (k_sort_pass)
        int shift = 0;
        while ((((nt - 1) >> shift) + 1) > 64) shift++;
        const int blocks = (int)((sp.n + SORT2_SLICE - 1) / SORT2_SLICE);
        if (!c->sort2_opted_in) {
            PICSP_CUDA(cudaFuncSetAttribute(k_sort_pass<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SORT2_SMEM_BYTES));
            PICSP_CUDA(cudaFuncSetAttribute(k_sort_pass<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SORT2_SMEM_BYTES));
            c->sort2_opted_in = true;
        }
        PICSP_LAUNCH(c, (k_sort_pass<true>), blocks, SORT2_THREADS, SORT2_SMEM_BYTES, sp.x, sp.y, sp.vx, sp.vy, ids, (long long)sp.n,
                     push_const(c, s), shift, sp.tile_off, sp.cursor, sp.x2, sp.y2, sp.vx2, sp.vy2, sp.id2);
        PICSP_CUDA(cudaMemsetAsync(sp.cursor, 0, sizeof(unsigned int) * nt, c->stream));
        PICSP_LAUNCH(c, (k_sort_pass<false>), blocks, SORT2_THREADS, SORT2_SMEM_BYTES, sp.x2, sp.y2, sp.vx2, sp.vy2, sp.id2, (long long)sp.n,
                     push_const(c, s), shift, sp.tile_off, sp.cursor, sp.x, sp.y, sp.vx, sp.vy, sp.id);
        sort_finish(c, sp, false);
        if (bank_order_on(sp)) op_bank_order(c, sp);
        return;
    }
    if (sp.n > 0) {
        if (sp.sorted)
            PICSP_LAUNCH(c, k_resort_chunks, mover_grid(sp), RESORT_THREADS, 0, sp.x, sp.y, sp.vx, sp.vy, ids,
                         (const Chunk *)sp.chunks, sp.nchunks, push_const(c, s), sp.tile_off, sp.cursor,
                         sp.x2, sp.y2, sp.vx2, sp.vy2, sp.id2);
        else
            PICSP_LAUNCH(c, k_sort_scatter, particle_blocks(c, sp.n, 256), 256, 0, sp.x, sp.y, sp.vx, sp.vy, ids,
                         (long long)sp.n, push_const(c, s), sp.tile_off, sp.cursor, sp.x2, sp.y2, sp.vx2, sp.vy2, sp.id2);
    }
    sort_finish(c, sp);
    if (bank_order_on(sp)) op_bank_order(c, sp);
}

// Cell order inside every bin (stand-alone; see tile_kernels.cuh).  Particles stay in their bin's range; the result is
// left in the second buffer set, which becomes the live one.
void op_cell_sort(picsp_ctx *c, Species &sp) {
    PhaseScope ph(c, PICSP_PHASE_SORT);
    const int s = sp.s;
    borrow_spare(c, sp);
    if (!sp.sorted || sp.n <= 0 || !sp.x2) return;
    const int grid = mover_grid(sp);
    if (sp.cell_cnt_chunks < grid) {
        cudaFree(sp.cell_cnt); sp.cell_cnt = nullptr;
        dalloc(&sp.cell_cnt, (size_t)grid * CELLKEYS);
        sp.cell_cnt_chunks = grid;
    }
    if (!sp.tile_chunk0) dalloc(&sp.tile_chunk0, (size_t)sp.ntiles);
    if (!c->cellsort_opted_in) {
        PICSP_CUDA(cudaFuncSetAttribute(k_cell_permute, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CELLSORT_SMEM_BYTES));
        c->cellsort_opted_in = true;
    }
    const PushConst pc = push_const(c, s);
    PICSP_LAUNCH(c, k_cell_count, grid, 256, 0, sp.x, sp.y, (const Chunk *)sp.chunks, sp.nchunks, pc, sp.cell_cnt, sp.tile_chunk0);
    PICSP_LAUNCH(c, k_cell_scan, sp.ntiles, CELLKEYS, 0, sp.tile_off, sp.tile_chunk0, sp.chunk, sp.cell_cnt);
    PICSP_LAUNCH(c, k_cell_permute, grid, SORT2_THREADS, CELLSORT_SMEM_BYTES, sp.x, sp.y, sp.vx, sp.vy,
                 sp.has_perm ? sp.id : (const uint32_t *)nullptr, (const Chunk *)sp.chunks, sp.nchunks, pc, sp.tile_off, sp.cell_cnt,
                 sp.x2, sp.y2, sp.vx2, sp.vy2, sp.id2);
    std::swap(sp.x, sp.x2); std::swap(sp.y, sp.y2); std::swap(sp.vx, sp.vx2); std::swap(sp.vy, sp.vy2);
    std::swap(sp.id, sp.id2);
    return_spare(c, sp);
    sp.has_perm = true;
    sp.cnt_valid = false;          // the per-chunk neighbour counts described the old order
    sp.staged_v_valid = false;
    sp.steps_since_cellsort = 0;
}

int mover_grid(const Species &sp) {
    long long b = sp.n / sp.chunk + sp.ntiles + 1;   // upper bound on the number of chunks
    return (int)std::max<long long>(1, std::min<long long>(b, sp.max_chunks));
}

// Bank order inside every chunk (see tile_kernels.cuh): in place, chunk table and counts untouched.  Automatic mode:
// species that are re-binned rarely (ions: every 96 steps), whose particles keep their cells between two re-binnings.
bool bank_order_on(const Species &sp) { return sp.bank_order < 0 ? (sp.sort_period >= 32 && sp.cell_period == 0) : sp.bank_order > 0; }
void op_bank_order(picsp_ctx *c, Species &sp) {
    const int s = sp.s;
    if (!BULK_PIPE || !sp.sorted || !sp.has_perm || sp.n <= 0) return;       // (callers open the PICSP_PHASE_SORT scope)
    if (!c->bankorder_opted_in) {
        PICSP_CUDA(cudaFuncSetAttribute(k_bank_order, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BANKORDER_SMEM_BYTES));
        c->bankorder_opted_in = true;
    }
    PICSP_LAUNCH(c, k_bank_order, mover_grid(sp), SORT2_THREADS, BANKORDER_SMEM_BYTES, sp.x, sp.y, sp.vx, sp.vy, sp.id,
                 (const Chunk *)sp.chunks, sp.nchunks, push_const(c, s), sp.frac);
}

template <int MODE> void launch_tile_mover(picsp_ctx *c, Species &sp) {
    const int s = sp.s;
    RebinArgs rb = {};
    if (MODE == 3 || MODE == 4) {
        PICSP_REQUIRE(sp.has_perm, PICSP_ERR_STATE, "internal: re-binning mover on a store without a slot map");
        rb.id = sp.id;
        rb.tile_off = sp.tile_off; rb.cursor = sp.cursor;
        rb.x2 = sp.x2; rb.y2 = sp.y2; rb.vx2 = sp.vx2; rb.vy2 = sp.vy2; rb.id2 = sp.id2;
        rb.chunk_base = sp.chunk_base;
    }
    if (MODE == 0) rb.chunk_cnt = sp.chunk_cnt;      // nullptr until the first sort allocated it
    CUtensorMap tm;
    memcpy(&tm, c->tmapE, sizeof(tm));
    if (!c->smem_opted_in) {               // dynamic shared memory above 48 KB needs a per-function opt-in (per device)
        PICSP_CUDA(cudaFuncSetAttribute(k_tile_mover<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MOVER_SMEM_BYTES));
        PICSP_CUDA(cudaFuncSetAttribute(k_tile_mover<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MOVER_SMEM_BYTES));
        PICSP_CUDA(cudaFuncSetAttribute(k_tile_mover<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MOVER_SMEM_BYTES));
        PICSP_CUDA(cudaFuncSetAttribute(k_tile_mover<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MOVER_SMEM_BYTES));
        PICSP_CUDA(cudaFuncSetAttribute(k_tile_mover<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MOVER_SMEM_BYTES));
        c->smem_opted_in = true;
    }
    PICSP_LAUNCH(c, (k_tile_mover<MODE>), mover_grid(sp), MOVER_THREADS, MOVER_SMEM_BYTES, tm, sp.x, sp.y, sp.vx, sp.vy,
                 (const Chunk *)sp.chunks, sp.nchunks, push_const(c, s), c->E, sp.acc, sp.frac, sp.hist_next,
                 sp.counters, c->d_error, rb);
}

// -- TMA descriptor of the E field -------------------------------------------------------
// E is [nix][niy] of {efx, efy}; seen by TMA as a 2-D tensor of doubles, inner extent 2*niy.
// The encoder lives in libcuda; it is fetched through the runtime so that the library has
// no link-time dependency on the driver (it must load on a box without one).
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
void make_tensor_map(picsp_ctx *c) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    PICSP_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    PICSP_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, PICSP_ERR_CUDA, "cuTensorMapEncodeTiled unavailable");
    const Geom &g = c->g;
    cuuint64_t dims[2] = {(cuuint64_t)(2 * g.niy), (cuuint64_t)g.nix};
    cuuint64_t strides[1] = {(cuuint64_t)(2 * g.niy) * sizeof(double)};
    cuuint32_t box[2] = {(cuuint32_t)(2 * WPITCH), (cuuint32_t)WIN};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMap tm;
    CUresult r = ((EncodeTiledFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)c->E, dims, strides, box, estr,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PICSP_REQUIRE(r == CUDA_SUCCESS, PICSP_ERR_CUDA, "cuTensorMapEncodeTiled failed: " + std::to_string((int)r));
    static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap is 128 bytes");
    memcpy(c->tmapE, &tm, sizeof(tm));
    c->have_tmap = true;
}

// -- operations ---------------------------------------------------------------------
void op_allreduce_rho(picsp_ctx *c);   // comm section below
void op_peer_reduce_rho(picsp_ctx *c);

// makes sure acc_s holds the fixed-point deposit of the stored positions (scatter loop of scatterSpecies)
void ensure_acc(picsp_ctx *c, int s) {
    if (c->sp[s].acc_valid) return;        // (a species-level flag, kept equal on all parts)
    for_parts(c, s, [&](Species &sp) {
        if (tiled(c) && !sp.sorted) op_sort(c, sp);
        ensure_hist(c, sp);
    });
    compute_frac(c, s);
    for_parts(c, s, [&](Species &sp) {
        PICSP_CUDA(cudaMemsetAsync(sp.counters, 0, 4 * sizeof(unsigned long long), c->stream));
        if (sp.n > 0) {
            if (tiled(c))
                launch_tile_mover<1>(c, sp);
            else
                PICSP_LAUNCH(c, k_deposit, particle_blocks(c, sp.n, 256), 256, 0, sp.x, sp.y, (long long)sp.n,
                             push_const(c, s), sp.acc, sp.frac);
        }
        sp.acc_valid = true;
    });
}

void op_deposit(picsp_ctx *c, int s) {
    PhaseScope ph(c, PICSP_PHASE_DEPOSIT);
    Species &sp = c->sp[s];
    const Geom &g = c->g;
    ensure_acc(c, s);
    const double weight = sp.spwt / (g.dx * g.dx);   // value/dxdy, src/main.cpp:657,664
    const int clear = (c->prm.flags & PICSP_FLAG_CLEAR_DENSITY) ? 1 : 0;
    PICSP_LAUNCH(c, k_deposit_finalize, blocks_for(g.nn, 256, c->num_sms * 8), 256, 0, sp.den, sp.acc, sp.frac, weight, g.nn, clear);
    if (!walls(c)) PICSP_LAUNCH(c, k_fold_periodic, 1, 1024, 0, sp.den, g.nix, g.niy);
    for_parts(c, s, [](Species &q) { q.acc_valid = false; });
}

// picsp_step's grid phase: scatterSpecies x2 (finalize + fold) and computeRho in ONE launch
void op_grid_phase(picsp_ctx *c) {
    const Geom &g = c->g;
    {
        PhaseScope ph(c, PICSP_PHASE_DEPOSIT);
        ensure_acc(c, 0); ensure_acc(c, 1);
    }
    {
        PhaseScope ph(c, PICSP_PHASE_RHO);
        GridPhaseSpecies gs[2];
        for (int s = 0; s < 2; s++) {
            Species &sp = c->sp[s];
            gs[s].den = sp.den; gs[s].acc = sp.acc; gs[s].frac = sp.frac;
            gs[s].weight = sp.spwt / (g.dx * g.dx); gs[s].q = sp.q;
            for_parts(c, s, [](Species &q) { q.acc_valid = false; });
        }
        const int clear = (c->prm.flags & PICSP_FLAG_CLEAR_DENSITY) ? 1 : 0;
        double *rho_out = c->peer_ok ? c->peer_part[c->rank] : c->rho;      // sharded: the partial goes where the peers can read it
        if (walls(c))
            PICSP_LAUNCH(c, k_grid_phase_walls, blocks_for(g.nn, 256, c->num_sms * 32), 256, 0, gs[0], gs[1], rho_out, g.nix, g.niy, clear,
                         c->d_error, c->h_error_mapped);
        else
            PICSP_LAUNCH(c, k_grid_phase, blocks_for(g.nn, 256, c->num_sms * 32), 256, 0, gs[0], gs[1], rho_out, g.nix, g.niy, clear,
                         c->d_error, c->h_error_mapped);   // one node per thread up to 1.2M nodes: latency-bound otherwise
    }
    if (c->peer_ok) op_peer_reduce_rho(c);
    else if (c->comm) op_allreduce_rho(c);    // the folds are linear: folding the partial rho first commutes with the sum
}


void op_compute_rho(picsp_ctx *c) {
    const Geom &g = c->g;
    {
        PhaseScope ph(c, PICSP_PHASE_RHO);
        PICSP_LAUNCH(c, k_compute_rho, blocks_for((long long)(g.nix - 2) * (g.niy - 2), 256, c->num_sms * 8), 256, 0,
                     c->rho, c->sp[0].den, c->sp[1].den, c->sp[0].q, c->sp[1].q, g.nix, g.niy);
    }
    if (c->comm) op_allreduce_rho(c);
    {
        PhaseScope ph(c, PICSP_PHASE_RHO);
        if (walls(c)) PICSP_LAUNCH(c, k_zero_walls, 8, 256, 0, c->rho, g.nix, g.niy);
        else PICSP_LAUNCH(c, k_fold_periodic, 1, 1024, 0, c->rho, g.nix, g.niy);
    }
}

// -- own shared-memory DFT (fft_kernels.cuh) -----------------------------------------------------
int largest_prime_factor(int m) {
    int best = 1;
    for (int f = 2; (long long)f * f <= m; f++) while (m % f == 0) { best = std::max(best, f); m /= f; }
    return std::max(best, m);
}
int gcd_int(int a, int b) { while (b) { const int t = a % b; a = b; b = t; } return a; }
constexpr size_t OWN_FFT_MAX_SMEM = 200 * 1024;

// M = P * Q with gcd(P, Q) = 1 and P <= 32 that needs the least shared memory (P * L complex, L = 2^k >= 2Q - 1)
bool choose_fft_split(int M, int *P, int *Q, int *L, int *logL) {
    long long best = -1;
    for (int p = 1; p <= 32 && p <= M; p++) {
        if (M % p) continue;
        const int q = M / p;
        if (gcd_int(p, q) != 1) continue;
        int l = 2, ll = 1;
        while (l < 2 * q - 1) { l <<= 1; ll++; }
        const long long cost = (long long)p * l;
        if (best < 0 || cost < best) { best = cost; *P = p; *Q = q; *L = l; *logL = ll; }
    }
    return best > 0 && fft_padded((size_t)best) * sizeof(double2) <= OWN_FFT_MAX_SMEM;
}

// M = P * Q, coprime, both <= 64 (1025 = 25 * 41, 513 = 19 * 27, 65 = 5 * 13 ...): both factors as direct DFTs (kind 1)
bool choose_direct_split(int M, int *P, int *Q) {
    int best = -1;
    for (int p = 2; p <= 64 && (long long)p * p <= M; p++) {
        if (M % p) continue;
        const int q = M / p;
        if (q > 64 || gcd_int(p, q) != 1) continue;
        if (best < 0 || p + q < best) { best = p + q; *P = p; *Q = q; }
    }
    return best > 0;
}

template <class T> void upload_table(void **dev, const std::vector<T> &h, cudaStream_t st) {
    PICSP_CUDA(cudaMalloc(dev, std::max<size_t>(h.size(), 1) * sizeof(T)));
    PICSP_CUDA(cudaMemcpyAsync(*dev, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, st));
    PICSP_CUDA(cudaStreamSynchronize(st));       // the host vector dies with the caller
}

void build_direct_fft_plan(picsp_ctx *c, int which, int M, int P, int Q) {
    picsp_ctx::OwnFft &f = c->fft;
    f.kind[which] = 1; f.M[which] = M; f.P[which] = P; f.Q[which] = Q; f.L[which] = 0; f.logL[which] = 0;
    const long double PI = 3.141592653589793238462643383279502884L;
    std::vector<double2> rp((size_t)P), rq((size_t)Q);
    for (int m = 0; m < P; m++) { const long double a = 2.0L * PI * m / P; rp[(size_t)m] = make_double2((double)cosl(a), (double)-sinl(a)); }
    for (int m = 0; m < Q; m++) { const long double a = 2.0L * PI * m / Q; rq[(size_t)m] = make_double2((double)cosl(a), (double)-sinl(a)); }
    std::vector<int> in_pos((size_t)M), out_idx((size_t)M);
    for (int n1 = 0; n1 < P; n1++)
        for (int n2 = 0; n2 < Q; n2++) in_pos[(size_t)(((long long)n1 * Q + (long long)n2 * P) % M)] = n1 * Q + n2;
    for (int k = 0; k < M; k++) out_idx[(size_t)(k % P) * Q + (k % Q)] = k;
    upload_table(&f.rootP[which], rp, c->stream); upload_table(&f.rootQ[which], rq, c->stream);
    upload_table(&f.in_pos[which], in_pos, c->stream); upload_table(&f.out_idx[which], out_idx, c->stream);
    f.smem[which] = (size_t)(2 * M + P + Q) * sizeof(double2);
}

void build_own_fft_plan(picsp_ctx *c, int which, int M) {
    picsp_ctx::OwnFft &f = c->fft;
    int dp = 0, dq = 0;
    if (!getenv("PICSP_FFT_NO_DIRECT") && choose_direct_split(M, &dp, &dq)) { build_direct_fft_plan(c, which, M, dp, dq); return; }
    int P = 1, Q = M, L = 2, logL = 1;
    PICSP_REQUIRE(choose_fft_split(M, &P, &Q, &L, &logL), PICSP_ERR_INVALID, "own FFT: the transform does not fit shared memory");
    f.M[which] = M; f.P[which] = P; f.Q[which] = Q; f.L[which] = L; f.logL[which] = logL;
    const long double PI = 3.141592653589793238462643383279502884L;
    std::vector<double2> chirp((size_t)Q), b((size_t)L, make_double2(0.0, 0.0)), tw((size_t)L / 2), wp((size_t)P * P);
    for (int n = 0; n < Q; n++) {                      // c[n] = exp(-i pi n^2 / Q), the argument reduced exactly
        const long long r = ((long long)n * n) % (2ll * Q);
        const long double a = PI * (long double)r / (long double)Q;
        chirp[(size_t)n] = make_double2((double)cosl(a), (double)-sinl(a));
    }
    for (int m = 0; m < Q; m++) {                      // b[m] = conj(c[|m|]), m = -(Q-1) .. Q-1, wrapped to length L
        const double2 v = make_double2(chirp[(size_t)m].x, -chirp[(size_t)m].y);
        b[(size_t)m] = v;
        if (m > 0) b[(size_t)(L - m)] = v;
    }
    for (int k = 0; k < L / 2; k++) {
        const long double a = 2.0L * PI * (long double)k / (long double)L;
        tw[(size_t)k] = make_double2((double)cosl(a), (double)-sinl(a));
    }
    for (int k1 = 0; k1 < P; k1++)
        for (int n1 = 0; n1 < P; n1++) {
            const long double a = 2.0L * PI * (long double)((k1 * n1) % P) / (long double)P;
            wp[(size_t)k1 * P + n1] = make_double2((double)cosl(a), (double)-sinl(a));
        }
    std::vector<int> in_pos((size_t)M), out_idx((size_t)M);
    for (int n1 = 0; n1 < P; n1++)
        for (int n2 = 0; n2 < Q; n2++) in_pos[(size_t)(((long long)n1 * Q + (long long)n2 * P) % M)] = n1 * L + n2;
    for (int k = 0; k < M; k++) out_idx[(size_t)(k % P) * Q + (k % Q)] = k;
    upload_table(&f.chirp[which], chirp, c->stream); upload_table(&f.tw[which], tw, c->stream);
    upload_table(&f.wp[which], wp, c->stream);
    upload_table(&f.in_pos[which], in_pos, c->stream); upload_table(&f.out_idx[which], out_idx, c->stream);
    void *b_dev = nullptr;
    upload_table(&b_dev, b, c->stream);
    PICSP_CUDA(cudaMalloc(&f.bhat[which], (size_t)L * sizeof(double2)));
    const size_t bsm = fft_padded((size_t)L) * sizeof(double2);
    PICSP_CUDA(cudaFuncSetAttribute(k_blue_bhat, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OWN_FFT_MAX_SMEM));
    PICSP_LAUNCH(c, k_blue_bhat, 1, FFT_THREADS, bsm, (const double2 *)b_dev, (double2 *)f.bhat[which], L, logL, (const double2 *)f.tw[which]);
    PICSP_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(b_dev);
    f.smem[which] = fft_padded((size_t)P * L) * sizeof(double2);
}

// threads per transform: the radix-8 groups of a pass (P * L / 8) spread evenly, two or three per thread
int fft_threads(const picsp_ctx *c, int which) {
    if (const char *e = getenv("PICSP_FFT_THREADS")) return std::max(32, std::min(BLUE_THREADS, atoi(e)));
    if (c->fft.kind[which] == 1) {       // one round per stage: threads = work items of the larger stage
        const int P = c->fft.P[which], Q = c->fft.Q[which];
        const int i1 = P * ((Q / 2 + 1 + PFA2_KP - 1) / PFA2_KP), i2 = Q * ((P / 2 + 1 + PFA2_KP - 1) / PFA2_KP);
        return std::max(64, std::min(PFA2_THREADS, ((std::max(i1, i2) + 31) / 32) * 32));
    }
    const int ng = std::max(1, c->fft.P[which] * c->fft.L[which] / 8);
    return (ng % 384 == 0) ? 384 : 256;          // measured at 2049^2: 128 / 192 / 256 / 384 / 512 threads -> 633 / 549 / 513 / 490 / 499 us
}

BluePlanDev own_fft_plan(const picsp_ctx *c, int which) {
    const picsp_ctx::OwnFft &f = c->fft;
    BluePlanDev pl;
    pl.kind = f.kind[which];
    pl.rootP = (const double2 *)f.rootP[which]; pl.rootQ = (const double2 *)f.rootQ[which];
    pl.M = f.M[which]; pl.P = f.P[which]; pl.Q = f.Q[which]; pl.L = f.L[which]; pl.logL = f.logL[which];
    pl.chirp = (const double2 *)f.chirp[which]; pl.bhat = (const double2 *)f.bhat[which]; pl.tw = (const double2 *)f.tw[which];
    pl.in_pos = (const int *)f.in_pos[which]; pl.out_idx = (const int *)f.out_idx[which]; pl.wp = (const double2 *)f.wp[which];
    return pl;
}

// cuFFT unless a node count has a prime factor > 127 (cuFFT then runs Bluestein through global memory: 2049^2 takes
// 1.5-1.7 ms, the shared-memory transform a fraction of that; measured in profiles/r02b_own_fft.md)
void setup_own_fft(picsp_ctx *c) {
    const Geom &g = c->g;
    const int flags = c->prm.flags;
    if (flags & PICSP_FLAG_CUFFT_ONLY) return;
    // ... and the small grids where cuFFT's launches cost more than the transform (65^2 .. 257^2 nodes: 1.2-1.5x, same file)
    // ... and node counts that split into two coprime factors <= 64, both done as direct DFTs (1025 = 25 * 41: 1.28x cuFFT)
    const int big = std::max(g.nix, g.niy);
    int dp, dq;
    const bool direct_both = choose_direct_split(g.nix, &dp, &dq) && choose_direct_split(g.niy, &dp, &dq);
    const bool wanted = (flags & PICSP_FLAG_OWN_FFT) || largest_prime_factor(g.nix) > 127 || largest_prime_factor(g.niy) > 127 ||
                        (big >= 65 && (big <= 257 || direct_both));
    if (!wanted) return;
    int P, Q, L, logL;
    auto fits = [&](int M) { return choose_direct_split(M, &P, &Q) || choose_fft_split(M, &P, &Q, &L, &logL); };
    if (!fits(g.niy) || !fits(g.nix)) {
        PICSP_REQUIRE(!(flags & PICSP_FLAG_OWN_FFT), PICSP_ERR_INVALID, "PICSP_FLAG_OWN_FFT: a transform of this length does not fit shared memory");
        return;                                        // too long for shared memory: cuFFT
    }
    build_own_fft_plan(c, 0, g.niy);
    build_own_fft_plan(c, 1, g.nix);
    picsp_ctx::OwnFft &f = c->fft;
    const int big_smem = (int)OWN_FFT_MAX_SMEM;       // (an upper bound: each launch passes its own size)
    PICSP_CUDA(cudaFuncSetAttribute(k_fft_rows_fwd<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, big_smem));
    PICSP_CUDA(cudaFuncSetAttribute(k_fft_rows_inv<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, big_smem));
    PICSP_CUDA(cudaFuncSetAttribute((k_fft_cols<0, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, big_smem));
    PICSP_CUDA(cudaFuncSetAttribute((k_fft_cols<0, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, big_smem));
    PICSP_CUDA(cudaFuncSetAttribute(k_fft_rows_fwd<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big_smem));
    PICSP_CUDA(cudaFuncSetAttribute(k_fft_rows_inv<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big_smem));
    PICSP_CUDA(cudaFuncSetAttribute((k_fft_cols<1, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, big_smem));
    PICSP_CUDA(cudaFuncSetAttribute((k_fft_cols<1, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, big_smem));
    f.on = true;
}

void op_solve_spectral(picsp_ctx *c) {
    PhaseScope ph(c, PICSP_PHASE_SOLVE);
    const Geom &g = c->g;
    const int Nh = g.niy / 2 + 1;
    if (c->fft.on) {
        // D2Z = rows forward (two real rows per transform) + columns forward; Z2D = columns inverse + rows inverse
        const BluePlanDev py = own_fft_plan(c, 0), px = own_fft_plan(c, 1);
        double2 *rhok = reinterpret_cast<double2 *>(c->rhok), *phik = reinterpret_cast<double2 *>(c->phik);
        if (py.kind == 1) PICSP_LAUNCH(c, k_fft_rows_fwd<1>, (g.nix + 1) / 2, fft_threads(c, 0), c->fft.smem[0], py, c->rho, rhok, g.nix);
        else PICSP_LAUNCH(c, k_fft_rows_fwd<0>, (g.nix + 1) / 2, fft_threads(c, 0), c->fft.smem[0], py, c->rho, rhok, g.nix);
        if (px.kind == 1) PICSP_LAUNCH(c, (k_fft_cols<1, false>), Nh, fft_threads(c, 1), c->fft.smem[1], px, rhok, Nh);
        else PICSP_LAUNCH(c, (k_fft_cols<0, false>), Nh, fft_threads(c, 1), c->fft.smem[1], px, rhok, Nh);
        PICSP_LAUNCH(c, k_kspace_green, blocks_for((long long)g.nix * Nh, 256, c->num_sms * 8), 256, 0, c->rhok, c->phik,
                     g.nix, g.niy, g.xl, g.yl);
        if (px.kind == 1) PICSP_LAUNCH(c, (k_fft_cols<1, true>), Nh, fft_threads(c, 1), c->fft.smem[1], px, phik, Nh);
        else PICSP_LAUNCH(c, (k_fft_cols<0, true>), Nh, fft_threads(c, 1), c->fft.smem[1], px, phik, Nh);
        if (py.kind == 1) PICSP_LAUNCH(c, k_fft_rows_inv<1>, (g.nix + 1) / 2, fft_threads(c, 0), c->fft.smem[0], py, (const double2 *)phik, c->phi, g.nix);
        else PICSP_LAUNCH(c, k_fft_rows_inv<0>, (g.nix + 1) / 2, fft_threads(c, 0), c->fft.smem[0], py, (const double2 *)phik, c->phi, g.nix);
        return;
    }
    PICSP_REQUIRE(c->have_plans, PICSP_ERR_STATE, "cuFFT plans missing");
    PICSP_CUFFT(cufftExecD2Z(c->plan_fwd, c->rho, c->rhok));
    PICSP_LAUNCH(c, k_kspace_green, blocks_for((long long)g.nix * Nh, 256, c->num_sms * 8), 256, 0, c->rhok, c->phik,
                 g.nix, g.niy, g.xl, g.yl);
    PICSP_CUFFT(cufftExecZ2D(c->plan_inv, c->phik, c->phi));
}

void op_solve_sor(picsp_ctx *c) {
    PhaseScope ph(c, PICSP_PHASE_SOLVE);
    const Geom &g = c->g;
    const int bands = (g.nix + SOR_ROWS - 1) / SOR_ROWS;
    const size_t small_bytes = sor_smem_bytes(g.nix, g.niy);
    if (g.nix >= 3 && g.nix <= 1024 && small_bytes <= 200 * 1024 && !(c->prm.flags & PICSP_FLAG_SOR_SINGLE_CTA) && !getenv("PICSP_SOR_NO_SMEM")) {
        // small grid (the shipped input.ini: 65^2 nodes, 68 KB): sweep 0 entirely in shared memory, then the same test and fall-through
        if (!c->sor_smem_opted_in) {
            PICSP_CUDA(cudaFuncSetAttribute(k_sor_sweep_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            c->sor_smem_opted_in = true;
        }
        PICSP_LAUNCH(c, k_sor_sweep_smem, 1, ((g.nix + 31) / 32) * 32, small_bytes, c->phi, c->rho, g.nix, g.niy, g.dx, g.dx);
        PICSP_LAUNCH(c, k_sor_residual_partial, RED_BLOCKS, RED_THREADS, 0, c->phi, c->rho, g.nix, g.niy, g.dx, c->d_red);
        PICSP_LAUNCH(c, k_sor_residual_final, 1, 1024, 0, c->d_red, RED_BLOCKS, g.nix, g.niy, c->d_sor_status, c->d_scalars + 3);
        PICSP_LAUNCH(c, k_sor_solve, 1, 1024, 0, c->phi, c->rho, g.nix, g.niy, g.dx, g.dx, c->d_sor_status, c->d_scalars + 3, 200000, 1);
    } else if (bands <= c->num_sms && !(c->prm.flags & PICSP_FLAG_SOR_SINGLE_CTA)) {
        // sweep 0 pipelined over co-resident bands, then the reference's convergence test; further sweeps
        // (never needed in practice) fall through to the single-CTA kernel
        PICSP_CUDA(cudaMemsetAsync(c->d_sor_progress, 0, sizeof(int) * bands, c->stream));
        PICSP_LAUNCH(c, k_sor_sweep_pipelined, bands, SOR_ROWS, 0, c->phi, c->rho, g.nix, g.niy, g.dx, g.dx, c->d_sor_progress);
        PICSP_LAUNCH(c, k_sor_residual_partial, RED_BLOCKS, RED_THREADS, 0, c->phi, c->rho, g.nix, g.niy, g.dx, c->d_red);
        PICSP_LAUNCH(c, k_sor_residual_final, 1, 1024, 0, c->d_red, RED_BLOCKS, g.nix, g.niy, c->d_sor_status, c->d_scalars + 3);
        PICSP_LAUNCH(c, k_sor_solve, 1, 1024, 0, c->phi, c->rho, g.nix, g.niy, g.dx, g.dx, c->d_sor_status, c->d_scalars + 3, 200000, 1);
    } else {
        PICSP_LAUNCH(c, k_sor_solve, 1, 1024, 0, c->phi, c->rho, g.nix, g.niy, g.dx, g.dx, c->d_sor_status, c->d_scalars + 3, 200000, 0);
    }
}

// PICSP_FLAG_WALLS: Dirichlet red-black SOR to a residual of WALLS_TOL, one cooperative launch (extension, no reference)
constexpr double WALLS_TOL = 1e-12;
constexpr int WALLS_BATCH = 16, WALLS_MAX_SWEEPS = 100000;
void op_solve_walls(picsp_ctx *c) {
    PhaseScope ph(c, PICSP_PHASE_SOLVE);
    const Geom &g = c->g;
    if (c->walls_grid == 0) {
        int per_sm = 0;
        PICSP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_rb_sor, 256, 0));
        PICSP_REQUIRE(per_sm > 0, PICSP_ERR_CUDA, "k_rb_sor cannot be made resident");
        const long long want = ((long long)(g.nix - 2) * ((g.niy - 1) / 2) + 255) / 256;      // one node of a colour per thread
        c->walls_grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)std::min(per_sm, 4) * c->num_sms));
        dalloc(&c->d_walls_partial, (size_t)c->walls_grid);
    }
    double *phi = c->phi; const double *rho = c->rho;
    int nix = g.nix, niy = g.niy, max_sweeps = WALLS_MAX_SWEEPS, batch = WALLS_BATCH;
    double dx = g.dx, omega = c->walls_omega, tol = WALLS_TOL;
    long long *status = c->d_sor_status; double *l2 = c->d_scalars + 3, *partial = c->d_walls_partial;
    void *args[] = {&phi, &rho, &nix, &niy, &dx, &omega, &tol, &max_sweeps, &batch, &status, &l2, &partial};
    PICSP_CUDA(cudaLaunchCooperativeKernel((const void *)k_rb_sor, dim3(c->walls_grid), dim3(256), args, 0, c->stream));
    c->launches++;
    c->busy[0] = c->busy[1] = true;
}

void op_solve(picsp_ctx *c) {
    if (walls(c)) op_solve_walls(c);
    else if (c->prm.solverType == PICSP_SOLVER_SPECTRAL) op_solve_spectral(c);
    else op_solve_sor(c);
}

void op_compute_ef(picsp_ctx *c) {
    PhaseScope ph(c, PICSP_PHASE_EF);
    const Geom &g = c->g;
    if (walls(c)) PICSP_LAUNCH(c, k_compute_ef_walls, blocks_for(g.nn, 256, c->num_sms * 8), 256, 0, c->phi, c->E, g.nix, g.niy, g.dx);
    else PICSP_LAUNCH(c, k_compute_ef, blocks_for(g.nn, 256, c->num_sms * 8), 256, 0, c->phi, c->E, g.nix, g.niy, g.dx, g.dx);
}

// pushSpecies for one species: all its parts, one after the other on the library stream
void op_push(picsp_ctx *c, int s) {
    const bool fuse = !(c->prm.flags & PICSP_FLAG_NO_FUSE);
    const bool tile = tiled(c);
    const int nt = c->g.ntx * c->g.nty;
    // every part: first binning of a new load / stand-alone sort when one is due; histogram of the positions about to be
    // pushed.  The fixed-point scale of the fused deposit (species-level) needs ALL the histograms before the first launch.
    for_parts(c, s, [&](Species &sp) {
        sp.staged_v_valid = false;
        const bool due = tile && sp.sorted && sp.steps_since_sort >= sp.sort_period;
        // a periodic re-bin rides on the mover itself (MODE 3 / 4) when the fused bulk-pipeline mover is in use;
        // the first binning of an arbitrary load, and the unfused mover, use the stand-alone sort
        const bool rebin_in_mover = due && fuse && BULK_PIPE && sp.n > 0 && !(c->prm.flags & PICSP_FLAG_SEPARATE_SORT);
        if (tile && (!sp.sorted || (due && !rebin_in_mover))) op_sort(c, sp);
        // cell order inside the bins: a pass of its own on the steps where no re-binning is due
        if (tile && sp.sorted && !due && sp.cell_period > 0 && sp.steps_since_cellsort >= sp.cell_period && BULK_PIPE) op_cell_sort(c, sp);
        if (fuse || tile) ensure_hist(c, sp);   // histogram of the positions about to be pushed -> bound for acc
    });
    if (fuse) compute_frac(c, s);
    if (fuse && c->sp[s].acc_valid)   // a previous fused push was never consumed by a deposit: drop it
        PICSP_CUDA(cudaMemsetAsync(c->sp[s].acc, 0, sizeof(long long) * c->g.nn, c->stream));
    for_parts(c, s, [&](Species &sp) {
        // (a stand-alone sort above has reset steps_since_sort, so `due` here means: re-binning inside the mover)
        const bool rebin_in_mover = tile && sp.sorted && sp.steps_since_sort >= sp.sort_period && fuse && BULK_PIPE && sp.n > 0 &&
                                    !(c->prm.flags & PICSP_FLAG_SEPARATE_SORT);
        if (rebin_in_mover) { PhaseScope phs(c, PICSP_PHASE_SORT); sort_prepare(c, sp); }
        bool rebinned = false;
        {
            PhaseScope ph(c, PICSP_PHASE_PUSH);
            PhaseScope phs(c, s == 0 ? PICSP_PHASE_PUSH_IONS : PICSP_PHASE_PUSH_ELECTRONS);
            PICSP_CUDA(cudaMemsetAsync(sp.counters, 0, 4 * sizeof(unsigned long long), c->stream));
            if (fuse || tile) PICSP_CUDA(cudaMemsetAsync(sp.hist_next, 0, sizeof(unsigned int) * nt, c->stream));
            if (sp.n > 0) {
                if (tile) {
                    if (rebin_in_mover && sp.cnt_valid) {
                        // the previous launch counted, per chunk, where its particles went: reserve the ranges up front
                        const long long q = 9ll * mover_grid(sp);
                        PICSP_LAUNCH(c, k_rebin_bases, (int)((q + 255) / 256), 256, 0, (const Chunk *)sp.chunks, sp.nchunks, c->g.ntx, c->g.nty,
                                     sp.chunk_cnt, sp.tile_off, sp.cursor, sp.chunk_base, c->d_error);
                        launch_tile_mover<4>(c, sp); sort_finish(c, sp); rebinned = true;
                    } else if (rebin_in_mover) { launch_tile_mover<3>(c, sp); sort_finish(c, sp); rebinned = true; }
                    else if (fuse) { launch_tile_mover<0>(c, sp); sp.cnt_valid = sp.chunk_cnt != nullptr; }
                    else launch_tile_mover<2>(c, sp);
                } else {
                    const int blocks = particle_blocks(c, sp.n, 256);
                    if (fuse)
                        PICSP_LAUNCH(c, (k_push<true>), blocks, 256, 0, sp.x, sp.y, sp.vx, sp.vy, (long long)sp.n, push_const(c, s),
                                     c->E, sp.acc, sp.frac, sp.hist_next, sp.counters, c->d_error);
                    else
                        PICSP_LAUNCH(c, (k_push<false>), blocks, 256, 0, sp.x, sp.y, sp.vx, sp.vy, (long long)sp.n, push_const(c, s),
                                     c->E, sp.acc, sp.frac, sp.hist_next, sp.counters, c->d_error);
                }
            }
        }
        if (rebinned && bank_order_on(sp)) { PhaseScope phs(c, PICSP_PHASE_SORT); op_bank_order(c, sp); }    // the new layout, in bank order
        if (fuse || tile) {
            std::swap(sp.hist, sp.hist_next);
            sp.hist_valid = true;
        } else {
            sp.hist_valid = false;
        }
        sp.acc_valid = fuse;
        sp.steps_since_sort++;
        sp.steps_since_cellsort++;
    });
}

void op_rewind(picsp_ctx *c, int s) {
    PhaseScope ph(c, PICSP_PHASE_PUSH);
    for_parts(c, s, [&](Species &sp) {
        sp.staged_v_valid = false;
        if (sp.n > 0)
            PICSP_LAUNCH(c, k_rewind, particle_blocks(c, sp.n, 256), 256, 0, sp.x, sp.y, sp.vx, sp.vy, (long long)sp.n,
                         push_const(c, s), c->E);
    });
}

double read_scalar(picsp_ctx *c, const double *dptr) {
    PICSP_CUDA(cudaMemcpyAsync(c->h_pinned, dptr, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    PICSP_CUDA(cudaStreamSynchronize(c->stream));
    return c->h_pinned[0];
}

// -- NCCL, loaded lazily so the library has no link-time dependency on it -------------
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Reduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi &nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        // if the process (e.g. PyTorch) already loaded an NCCL, this resolves to that same copy
        api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!api.handle) return;
        api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.handle, "ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.handle, "ncclCommInitRank");
        api.AllReduce = (decltype(api.AllReduce))dlsym(api.handle, "ncclAllReduce");
        api.Reduce = (decltype(api.Reduce))dlsym(api.handle, "ncclReduce");
        api.AllGather = (decltype(api.AllGather))dlsym(api.handle, "ncclAllGather");
        api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
        api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
    });
    PICSP_REQUIRE(api.handle && api.GetUniqueId && api.CommInitRank && api.AllReduce && api.Reduce && api.CommDestroy,
                  PICSP_ERR_NCCL, "libnccl.so.2 could not be loaded");
    return api;
}
#define PICSP_NCCL(expr)                                                                        \
    do {                                                                                        \
        ncclResult_t r__ = (expr);                                                              \
        if (r__ != ncclSuccess)                                                                 \
            throw Error(PICSP_ERR_NCCL, std::string(#expr) + ": " +                             \
                        (nccl().GetErrorString ? nccl().GetErrorString(r__) : "nccl error"));   \
    } while (0)

void op_allreduce_rho(picsp_ctx *c) {
    // every rank holds the partial interior rho of its own particles; boundary nodes are 0 (Q3)
    PhaseScope ph(c, PICSP_PHASE_ALLREDUCE);
    PICSP_NCCL(nccl().AllReduce(c->rho, c->rho, (size_t)c->g.nn, ncclFloat64, ncclSum, c->comm, c->stream));
}

// reduce-scatter + all-gather of the partial rho over peer memory (peer_kernels.cuh): barrier, one kernel, barrier
void op_peer_reduce_rho(picsp_ctx *c) {
    PhaseScope ph(c, PICSP_PHASE_ALLREDUCE);
    PeerPtrs pp;
    for (int k = 0; k < PEER_MAX_RANKS; k++) { pp.part[k] = c->peer_part[k]; pp.full[k] = c->peer_full[k]; pp.flags[k] = c->peer_flags[k]; }
    const long long slice = (c->g.nn + c->nranks - 1) / c->nranks;
    c->peer_epoch++;
    PICSP_LAUNCH(c, k_peer_barrier, 1, 32, 0, pp, c->rank, c->nranks, c->peer_epoch, 0, c->d_error);       // every partial is complete
    PICSP_LAUNCH(c, k_peer_reduce, blocks_for(slice, 256, c->num_sms * 8), 256, 0, pp, c->rank, c->nranks, c->g.nn);
    PICSP_LAUNCH(c, k_peer_barrier, 1, 32, 0, pp, c->rank, c->nranks, c->peer_epoch, 1, c->d_error);       // every slice has landed everywhere
}

// Maps the ranks' blocks into each other (CUDA IPC handles exchanged with one ncclAllGather).  Any failure on any rank
// (ranks on different nodes, no peer access, IPC refused) leaves ALL ranks on the NCCL all-reduce.
void peer_setup(picsp_ctx *c) {
    const Geom &g = c->g;
    if (c->nranks < 2 || c->nranks > PEER_MAX_RANKS || (c->prm.flags & PICSP_FLAG_NCCL_ONLY) || !nccl().AllGather) return;
    const size_t nnp = ((size_t)g.nn + 31) & ~(size_t)31;          // every sub-buffer 256-byte aligned (cuFFT reads rho from here)
    const size_t words = 2 * nnp + 2 * PEER_MAX_RANKS + 16;
    int ok = 1;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    if (cudaMalloc((void **)&c->peer_block, words * sizeof(double)) != cudaSuccess) { cudaGetLastError(); c->peer_block = nullptr; ok = 0; }
    if (ok && cudaMemsetAsync(c->peer_block, 0, words * sizeof(double), c->stream) != cudaSuccess) ok = 0;
    if (ok && cudaIpcGetMemHandle(&mine, c->peer_block) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    // exchange {ok, handle}: 128 bytes per rank
    struct Slot { int ok; int pad[15]; cudaIpcMemHandle_t h; };
    static_assert(sizeof(cudaIpcMemHandle_t) == 64 && sizeof(Slot) == 128, "IPC handle layout");
    Slot my; memset(&my, 0, sizeof(my)); my.ok = ok; my.h = mine;
    Slot *d_all = nullptr;
    PICSP_CUDA(cudaMalloc((void **)&d_all, sizeof(Slot) * c->nranks));
    PICSP_CUDA(cudaMemcpyAsync(d_all + c->rank, &my, sizeof(Slot), cudaMemcpyHostToDevice, c->stream));
    PICSP_NCCL(nccl().AllGather(d_all + c->rank, d_all, sizeof(Slot), ncclInt8, c->comm, c->stream));
    std::vector<Slot> all(c->nranks);
    PICSP_CUDA(cudaMemcpyAsync(all.data(), d_all, sizeof(Slot) * c->nranks, cudaMemcpyDeviceToHost, c->stream));
    PICSP_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_all);
    for (int k = 0; k < c->nranks; k++) ok = ok && all[k].ok;
    for (int k = 0; ok && k < c->nranks; k++) {
        if (k == c->rank) continue;
        if (cudaIpcOpenMemHandle(&c->peer_mapped[k], all[k].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); c->peer_mapped[k] = nullptr; ok = 0; }
    }
    // second round: did every rank manage to map every block?
    double *d_flag = c->d_scalars + 5;
    const double mine_ok = ok ? 0.0 : 1.0;
    PICSP_CUDA(cudaMemcpyAsync(d_flag, &mine_ok, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    PICSP_NCCL(nccl().AllReduce(d_flag, d_flag, 1, ncclFloat64, ncclSum, c->comm, c->stream));
    const double failures = read_scalar(c, d_flag);
    if (failures != 0.0) {
        for (int k = 0; k < c->nranks; k++) if (c->peer_mapped[k]) { cudaIpcCloseMemHandle(c->peer_mapped[k]); c->peer_mapped[k] = nullptr; }
        cudaFree(c->peer_block); c->peer_block = nullptr;
        return;
    }
    for (int k = 0; k < c->nranks; k++) {
        double *base = k == c->rank ? c->peer_block : (double *)c->peer_mapped[k];
        c->peer_part[k] = base; c->peer_full[k] = base + nnp;
        c->peer_flags[k] = reinterpret_cast<unsigned long long *>(base + 2 * nnp);
    }
    // the summed rho now lives in this rank's block: every consumer of c->rho (solve, downloads, per-function calls) follows
    PICSP_CUDA(cudaMemcpyAsync(c->peer_full[c->rank], c->rho, sizeof(double) * g.nn, cudaMemcpyDeviceToDevice, c->stream));
    PICSP_CUDA(cudaStreamSynchronize(c->stream));
    c->rho_owned = c->rho;
    c->rho = c->peer_full[c->rank];
    c->peer_ok = true;
}

}  // namespace

// ====================================================================================
// extern "C"
// ====================================================================================
extern "C" {

int picsp_abi_version(void) { return PICSP_B200_ABI_VERSION; }
const char *picsp_last_error(void) { return g_last_error.c_str(); }

int picsp_create(const picsp_params *p, picsp_ctx **out) {
    picsp_ctx *c = nullptr;
    try {
        PICSP_REQUIRE(p && out, PICSP_ERR_INVALID, "null argument");
        *out = nullptr;
        PICSP_REQUIRE(p->numxCells >= 2 && p->numyCells >= 2, PICSP_ERR_INVALID, "numxCells/numyCells must be >= 2");
        PICSP_REQUIRE(p->stepSize > 0 && p->timeStep > 0, PICSP_ERR_INVALID, "stepSize and timeStep must be positive");
        PICSP_REQUIRE(p->solverType == PICSP_SOLVER_SPECTRAL || p->solverType == PICSP_SOLVER_SOR, PICSP_ERR_INVALID,
                      "solverType must be 1 (spectral) or 2 (SOR)");   // src/main.cpp:310
        PICSP_REQUIRE(p->capacity[0] >= 0 && p->capacity[1] >= 0, PICSP_ERR_INVALID, "negative capacity");
        PICSP_REQUIRE(!((p->flags & PICSP_FLAG_WALLS) && (p->flags & PICSP_FLAG_NO_SORT)), PICSP_ERR_INVALID,
                      "PICSP_FLAG_WALLS needs the tiled store (not combinable with PICSP_FLAG_NO_SORT)");
        PICSP_REQUIRE(p->parts >= 0 && p->parts <= 64, PICSP_ERR_INVALID, "parts must be 0 (automatic) .. 64");
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0) {
            cudaGetLastError();
            return fail(PICSP_ERR_NO_DEVICE, "no CUDA device available: picsp_b200 has no CPU path");
        }
        PICSP_REQUIRE(p->device >= 0 && p->device < ndev, PICSP_ERR_INVALID, "bad device ordinal");
        PICSP_CUDA(cudaSetDevice(p->device));

        c = new picsp_ctx();
        c->prm = *p;
        Geom &g = c->g;
        g.ncx = p->numxCells; g.ncy = p->numyCells;
        g.nix = g.ncx + 1; g.niy = g.ncy + 1;                   // src/main.cpp:363,371
        g.dx = p->stepSize; g.dt = p->timeStep;
        g.xl = (g.nix - 1) * g.dx; g.yl = (g.niy - 1) * g.dx;   // src/main.cpp:366,374
        g.nn = (long long)g.nix * g.niy;
        g.guard = 4ll * g.niy + 8;
        g.ntx = (g.ncx + TILE - 1) / TILE; g.nty = (g.ncy + TILE - 1) / TILE;
        // optimal over-relaxation of the 5-point Dirichlet problem on the larger grid extent
        c->walls_omega = 2.0 / (1.0 + std::sin(3.14159265358979323846 / (double)std::max(g.ncx, g.ncy)));

        cudaDeviceProp prop;
        PICSP_CUDA(cudaGetDeviceProperties(&prop, p->device));
        c->num_sms = prop.multiProcessorCount;
        PICSP_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));

        // Parts per species.  One part = each species with a full second buffer set of its own (allocated at its first
        // binning).  When the device cannot hold that (state + as much again), the species are split so that ONE part-sized
        // spare serves all parts of both species: 2 * 36 B * capacity * (1 + 1 / (2 * parts)).
        const int64_t cap_max = std::max(p->capacity[0], p->capacity[1]);
        int nparts = p->parts;
        if (nparts == 0) {
            nparts = 1;
            if (!(p->flags & PICSP_FLAG_NO_SORT)) {
                size_t free_b = 0, total_b = 0;
                PICSP_CUDA(cudaMemGetInfo(&free_b, &total_b));
                const double state = 36.0 * ((double)p->capacity[0] + (double)p->capacity[1]);
                const double grids = 16.0 * 8.0 * (double)g.nn + 3e9;                  // grids, FFT work area, tables, head room
                while (nparts < 64 && state * (nparts == 1 ? 2.0 : 1.0 + 0.5 / nparts) + grids > (double)free_b) nparts *= 2;
            }
        }
        PICSP_REQUIRE(nparts == 1 || !(p->flags & PICSP_FLAG_NO_SORT), PICSP_ERR_INVALID, "parts > 1 need the tiled store");
        c->nparts = nparts;
        c->part_cap = nparts == 1 ? 0 : (cap_max + nparts - 1) / nparts;
        PICSP_REQUIRE((nparts == 1 ? cap_max : c->part_cap) <= 0xFFFFFFFFll, PICSP_ERR_INVALID,
                      "a part of a species is limited to 2^32-1 particles (slot map): raise picsp_params::parts");
        const size_t ntl = (size_t)g.ntx * g.nty;
        for (int s = 0; s < 2; s++) {
            c->cap_total[s] = p->capacity[s];
            c->more[s].resize((size_t)nparts - 1);
            for (int q = 0; q < nparts; q++) {
                Species &sp = part_of(c, s, q);
                sp.s = s; sp.part = q;
                sp.q = p->charge[s]; sp.m = p->mass[s]; sp.spwt = p->spwt[s];
                sp.shares_spare = nparts > 1;
                if (nparts == 1) { sp.first = 0; sp.cap = p->capacity[s]; }
                else {
                    sp.first = (int64_t)q * c->part_cap;
                    sp.cap = c->part_cap;      // every set the same size: any of them can become the spare (the LAST part of a species may hold fewer particles)
                }
                alloc_particle_set(&sp.x, &sp.y, &sp.vx, &sp.vy, sp.cap);
                if (nparts > 1) dalloc(&sp.id, sp.cap + 8);
                if (q == 0) {
                    dalloc(&sp.den, g.nn); dalloc(&sp.acc, g.nn); dalloc(&sp.frac, 2); dalloc(&sp.frac_scratch, 3);
                    PICSP_CUDA(cudaMemsetAsync(sp.den, 0, sizeof(double) * g.nn, c->stream));      // src/main.cpp:427,431
                    PICSP_CUDA(cudaMemsetAsync(sp.acc, 0, sizeof(long long) * g.nn, c->stream));
                    PICSP_CUDA(cudaMemsetAsync(sp.frac, 0, 2 * sizeof(int), c->stream));
                    PICSP_CUDA(cudaMemsetAsync(sp.frac_scratch, 0, 3 * sizeof(unsigned long long), c->stream));
                } else {                       // species-level members: aliases of part 0's
                    const Species &p0 = c->sp[s];
                    sp.den = p0.den; sp.acc = p0.acc; sp.frac = p0.frac; sp.frac_scratch = p0.frac_scratch;
                }
                dalloc(&sp.hist, ntl); dalloc(&sp.hist_next, ntl);
                dalloc(&sp.counters, 4);
                sp.sort_period = (s == 0) ? 96 : 8;     // steps between re-binnings (ions barely move; electrons: profiles/r01_sweeps.md)
                sp.cell_period = 0;                     // steps between cell orderings inside the bins: off (profiles/r02_mover_aggregation.md)
                sp.steps_since_cellsort = sp.cell_period;   // the first push after a load orders it
                sp.ntiles = g.ntx * g.nty;
                sp.max_chunks = sp.cap / 512 + (long long)ntl + 1;   // 512 = smallest chunk pick_chunk() returns
                dalloc(&sp.tile_off, ntl + 1);
                dalloc((Chunk **)&sp.chunks, (size_t)sp.max_chunks);
                dalloc(&sp.nchunks, 1);
                dalloc(&sp.cursor, ntl);
                PICSP_CUDA(cudaMemsetAsync(sp.nchunks, 0, sizeof(int), c->stream));
                PICSP_CUDA(cudaMemsetAsync(sp.counters, 0, 4 * sizeof(unsigned long long), c->stream));
            }
            if (nparts > 1) dalloc(&c->hist_sum[s], ntl);
        }
        if (nparts > 1) {
            alloc_particle_set(&c->spare.x, &c->spare.y, &c->spare.vx, &c->spare.vy, c->part_cap);
            dalloc(&c->spare.id, (size_t)c->part_cap + 8);
            dalloc(&c->d_part_sums, (size_t)nparts);
        }
        dalloc(&c->rho, g.nn); dalloc(&c->phi, g.nn);
        dalloc(&c->E_alloc, (size_t)(g.nn + 2 * g.guard));
        c->E = c->E_alloc + g.guard;
        PICSP_CUDA(cudaMemsetAsync(c->rho, 0, sizeof(double) * g.nn, c->stream));          // src/main.cpp:392-395
        PICSP_CUDA(cudaMemsetAsync(c->phi, 0, sizeof(double) * g.nn, c->stream));
        PICSP_CUDA(cudaMemsetAsync(c->E_alloc, 0, sizeof(double2) * (g.nn + 2 * g.guard), c->stream));
        make_tensor_map(c);
        dalloc(&c->d_red, RED_BLOCKS); dalloc(&c->d_scalars, 8); dalloc(&c->d_sor_status, 2); dalloc(&c->d_sor_progress, (size_t)g.nix + 1); dalloc(&c->d_error, 1);
        PICSP_CUDA(cudaMemsetAsync(c->d_scalars, 0, sizeof(double) * 8, c->stream));
        PICSP_CUDA(cudaMemsetAsync(c->d_sor_status, 0, sizeof(long long) * 2, c->stream));
        PICSP_CUDA(cudaMemsetAsync(c->d_error, 0, sizeof(int), c->stream));
        PICSP_CUDA(cudaMallocHost((void **)&c->h_pinned, 64 * sizeof(double)));
        PICSP_CUDA(cudaHostAlloc((void **)&c->h_error_mapped, sizeof(int), cudaHostAllocMapped));   // same address on the device (UVA)
        *c->h_error_mapped = 0;

        if (p->solverType == PICSP_SOLVER_SPECTRAL) {
            const size_t nk = (size_t)g.nix * (g.niy / 2 + 1);
            dalloc(&c->rhok, nk); dalloc(&c->phik, nk);
            setup_own_fft(c);
            if (!c->fft.on) {
                PICSP_CUFFT(cufftPlan2d(&c->plan_fwd, g.nix, g.niy, CUFFT_D2Z));
                PICSP_CUFFT(cufftPlan2d(&c->plan_inv, g.nix, g.niy, CUFFT_Z2D));
                PICSP_CUFFT(cufftSetStream(c->plan_fwd, c->stream));
                PICSP_CUFFT(cufftSetStream(c->plan_inv, c->stream));
                c->have_plans = true;
            }
        }
        PICSP_CUDA(cudaStreamSynchronize(c->stream));
        *out = c;
        return PICSP_OK;
    } catch (const Error &e) {
        if (c) picsp_destroy(c);
        return fail(e);
    } catch (const std::exception &e) {
        if (c) picsp_destroy(c);
        return fail(PICSP_ERR_INVALID, e.what());
    }
}

void picsp_destroy(picsp_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->prm.device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    if (c->peer_ok) {
        // nobody may unmap or free a block a peer kernel could still touch: meet once more, then tear down
        try { if (c->comm) { nccl().AllReduce(c->d_scalars + 4, c->d_scalars + 4, 1, ncclFloat64, ncclSum, c->comm, c->stream); cudaStreamSynchronize(c->stream); } } catch (...) {}
        for (int k = 0; k < c->nranks; k++) if (c->peer_mapped[k]) cudaIpcCloseMemHandle(c->peer_mapped[k]);
        c->rho = c->rho_owned;
        cudaFree(c->peer_block);
    }
    if (c->comm) { try { nccl().CommDestroy(c->comm); } catch (...) {} }
    if (c->have_plans) { cufftDestroy(c->plan_fwd); cufftDestroy(c->plan_inv); }
    for (int w = 0; w < 2; w++) {
        cudaFree(c->fft.rootP[w]); cudaFree(c->fft.rootQ[w]);
        cudaFree(c->fft.chirp[w]); cudaFree(c->fft.bhat[w]); cudaFree(c->fft.tw[w]); cudaFree(c->fft.in_pos[w]); cudaFree(c->fft.out_idx[w]); cudaFree(c->fft.wp[w]);
    }
    for (int s = 0; s < 2; s++) {
        for (int q = 0; q < 1 + (int)c->more[s].size(); q++) {
            Species &sp = part_of(c, s, q);
            cudaFree(sp.x); cudaFree(sp.id);            // x is the base of the set's single allocation
            if (q == 0) { cudaFree(sp.den); cudaFree(sp.acc); cudaFree(sp.frac); cudaFree(sp.frac_scratch); }   // parts >= 1 alias these
            cudaFree(sp.hist); cudaFree(sp.hist_next); cudaFree(sp.counters);
            if (!sp.shares_spare) { cudaFree(sp.x2); cudaFree(sp.id2); }     // (borrowed from the spare otherwise)
            cudaFree(sp.tile_off); cudaFree(sp.chunks); cudaFree(sp.nchunks); cudaFree(sp.cursor);
            cudaFree(sp.chunks2); cudaFree(sp.nchunks2); cudaFree(sp.chunk_cnt); cudaFree(sp.chunk_base);
            cudaFree(sp.cell_cnt); cudaFree(sp.tile_chunk0); cudaFree(sp.scan_chunk0);
        }
        cudaFree(c->hist_sum[s]);
    }
    cudaFree(c->spare.x); cudaFree(c->spare.id); cudaFree(c->d_part_sums);
    cudaFree(c->rho); cudaFree(c->phi); cudaFree(c->E_alloc); cudaFree(c->rhok); cudaFree(c->phik);
    cudaFree(c->d_walls_partial);
    cudaFree(c->d_red); cudaFree(c->d_scalars); cudaFree(c->d_sor_status); cudaFree(c->d_sor_progress); cudaFree(c->d_error); cudaFree(c->stage);
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    if (c->h_error_mapped) cudaFreeHost(c->h_error_mapped);
    cudaFree(c->snap_rows[0]); cudaFree(c->snap_rows[1]); cudaFree(c->snap_grids); cudaFree(c->snap_ke);
    if (c->ev_snap) cudaEventDestroy(c->ev_snap);
    if (c->ev_dump_done) cudaEventDestroy(c->ev_dump_done);
    if (c->ev_dump_done2) cudaEventDestroy(c->ev_dump_done2);
    if (c->copy_stream2) { cudaStreamSynchronize(c->copy_stream2); cudaStreamDestroy(c->copy_stream2); }
    for (auto &t : c->timers) for (auto e : t.pool) cudaEventDestroy(e);
    for (int g = 0; g < c->step_graph_count; g++) cudaGraphExecDestroy(c->step_graphs[g].exec);
    for (cudaGraphExec_t g : c->retired_graphs) cudaGraphExecDestroy(g);
    if (c->copy_stream) {
        cudaStreamDestroy(c->copy_stream);
        for (int k = 0; k < 4; k++) cudaEventDestroy(c->ev_ready[k]);
    }
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int picsp_sync(picsp_ctx *c) {
    PICSP_API_BEGIN
    check_ctx(c);
    PICSP_CUDA(cudaSetDevice(c->prm.device));
    check_device_error(c);
    PICSP_API_END
}

// ---- state exchange -------------------------------------------------------------------
static void ensure_copy_stream(picsp_ctx *c) {
    if (c->copy_stream) return;
    int lo_prio = 0, hi_prio = 0;      // highest priority: a copy kernel's few CTAs are placed before the mover's pending ones
    PICSP_CUDA(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
    PICSP_CUDA(cudaStreamCreateWithPriority(&c->copy_stream, cudaStreamNonBlocking, hi_prio));
    for (int k = 0; k < 4; k++) PICSP_CUDA(cudaEventCreateWithFlags(&c->ev_ready[k], cudaEventDisableTiming));
}

// particles of part `sp` when the species holds n in all: parts are filled in order, only the last one may be short
static int64_t part_count(const Species &sp, int64_t n) { return std::max<int64_t>(0, std::min<int64_t>(sp.cap, n - sp.first)); }

// a new load has been written into the part's arrays: every derived state is stale
static void reset_part_state(picsp_ctx *c, Species &sp, int64_t n_part) {
    sp.n = n_part; sp.hist_valid = false; sp.acc_valid = false;
    sp.has_perm = false; sp.sorted = false; sp.steps_since_sort = 0; sp.cnt_valid = false; sp.staged_v_valid = false;
    sp.steps_since_cellsort = sp.cell_period;
}

int picsp_species_upload(picsp_ctx *c, int s, const double *x, const double *y, const double *vx, const double *vy, int64_t n) {
    PICSP_API_BEGIN
    check_ctx(c); check_species(s);
    PICSP_CUDA(cudaSetDevice(c->prm.device));
    PICSP_REQUIRE(n >= 0 && n <= c->cap_total[s], PICSP_ERR_INVALID, "particle count exceeds the capacity given to picsp_create");
    PICSP_REQUIRE(n == 0 || (x && y && vx && vy), PICSP_ERR_INVALID, "null particle array");
    // The copies run on the copy stream; the library stream is only waited for when work that may touch THIS species
    // is still queued on it.  The first binning of the new load is enqueued before returning and not waited for, so
    // it runs while the caller uploads the other species (a 5e8-particle binning is ~70 ms, an upload ~290 ms).
    if (c->busy[s]) {
        PICSP_CUDA(cudaStreamSynchronize(c->stream));
        c->busy[0] = c->busy[1] = false;
    }
    ensure_copy_stream(c);
    for_parts(c, s, [&](Species &sp) {
        const int64_t np = part_count(sp, n);
        const size_t bytes = sizeof(double) * (size_t)np;
        if (np == 0) return;
        PICSP_CUDA(cudaMemcpyAsync(sp.x, x + sp.first, bytes, cudaMemcpyHostToDevice, c->copy_stream));
        PICSP_CUDA(cudaMemcpyAsync(sp.y, y + sp.first, bytes, cudaMemcpyHostToDevice, c->copy_stream));
        PICSP_CUDA(cudaMemcpyAsync(sp.vx, vx + sp.first, bytes, cudaMemcpyHostToDevice, c->copy_stream));
        PICSP_CUDA(cudaMemcpyAsync(sp.vy, vy + sp.first, bytes, cudaMemcpyHostToDevice, c->copy_stream));
    });
    PICSP_CUDA(cudaStreamSynchronize(c->copy_stream));        // the caller's buffers are free again
    const bool other_busy = c->busy[1 - s];
    if (c->sp[s].acc_valid) PICSP_CUDA(cudaMemsetAsync(c->sp[s].acc, 0, sizeof(long long) * c->g.nn, c->stream));
    c->n_total[s] = n;
    for_parts(c, s, [&](Species &sp) {
        reset_part_state(c, sp, part_count(sp, n));
        if (tiled(c) && sp.n > 0) op_sort(c, sp);
    });
    c->busy[s] = true; c->busy[1 - s] = other_busy;           // what was enqueued here touches species s only
    PICSP_API_END
}

static void ensure_stage(picsp_ctx *c, int64_t n) {
    if (c->stage_cap >= n) return;
    cudaFree(c->stage); c->stage = nullptr; c->stage_cap = 0;
    dalloc(&c->stage, (size_t)n);
    c->stage_cap = n;
}

int picsp_species_download(picsp_ctx *c, int s, double *x, double *y, double *vx, double *vy) {
    PICSP_API_BEGIN
    check_ctx(c); check_species(s);
    PICSP_CUDA(cudaSetDevice(c->prm.device));
    double *dst[4] = {x, y, vx, vy};
    for_parts(c, s, [&](Species &sp) {
        const size_t bytes = sizeof(double) * (size_t)sp.n;
        double *src[4] = {sp.x, sp.y, sp.vx, sp.vy};
        borrow_spare(c, sp);
        if (sp.has_perm && sp.n > 0) {
            // Binned store: bring every array back to upload order on the device (out[id[slot]] = in[slot]) and copy it
            // out.  The idle half of the sort's ping-pong buffers is the staging area (its contents are dead between
            // sorts), one buffer per array, so the four un-permutes run back to back on the library stream while the
            // device->host copies follow them on a second stream: only the first un-permute is exposed.
            ensure_copy_stream(c);
            double *stage[4] = {sp.x2, sp.y2, sp.vx2, sp.vy2};
            for (int k = 0; k < 4; k++) {
                if (!dst[k]) continue;
                PICSP_LAUNCH(c, k_unpermute, particle_blocks(c, sp.n, 256), 256, 0, src[k], sp.id, stage[k], (long long)sp.n);
                PICSP_CUDA(cudaEventRecord(c->ev_ready[k], c->stream));
                PICSP_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_ready[k], 0));
                PICSP_CUDA(cudaMemcpyAsync(dst[k] + sp.first, stage[k], bytes, cudaMemcpyDeviceToHost, c->copy_stream));
            }
            PICSP_CUDA(cudaStreamSynchronize(c->copy_stream));     // (several parts share the staging area: one part at a time)
            sp.staged_v_valid = !sp.shares_spare && dst[2] && dst[3];     // a KE request right after the dump can reduce these directly
        } else if (sp.n > 0) {
            for (int k = 0; k < 4; k++)
                if (dst[k]) PICSP_CUDA(cudaMemcpyAsync(dst[k] + sp.first, src[k], bytes, cudaMemcpyDeviceToHost, c->stream));
        }
    });
    check_device_error(c);
    PICSP_API_END
}

// writeSpecies row layout {x, y, vx, vy} (src/main.cpp:1156-1159) of one part, built on the device: every particle's
// 32 bytes go to row id[slot] of `part_rows` (one whole sector per particle)
static void rows_of_part(picsp_ctx *c, Species &sp, double *part_rows) {
    if (sp.n <= 0) return;
    PICSP_LAUNCH(c, k_rows_unpermute, particle_blocks(c, sp.n, 256), 256, 0, sp.x, sp.y, sp.vx, sp.vy,
                 sp.has_perm ? sp.id : (const uint32_t *)nullptr, (long long)sp.n, reinterpret_cast<double2 *>(part_rows));
}

int picsp_species_download_rows(picsp_ctx *c, int s, double *rows) {
    PICSP_API_BEGIN
    check_ctx(c); check_species(s);
    PICSP_REQUIRE(rows != nullptr, PICSP_ERR_INVALID, "null rows");
    PICSP_CUDA(cudaSetDevice(c->prm.device));
    bool host_interleave = false;
    for_parts(c, s, [&](Species &sp) {
        borrow_spare(c, sp);
        if (sp.n > 0 && sp.x2) {
            // built in the idle buffer set (a contiguous block of 4 * cap doubles), then one device->host copy
            rows_of_part(c, sp, sp.x2);
            sp.staged_v_valid = false;                  // the staging block has been overwritten
            PICSP_CUDA(cudaMemcpyAsync(rows + 4 * sp.first, sp.x2, sizeof(double) * 4 * (size_t)sp.n, cudaMemcpyDeviceToHost, c->stream));
            if (sp.shares_spare) PICSP_CUDA(cudaStreamSynchronize(c->stream));
        } else if (sp.n > 0) host_interleave = true;
    });
    if (host_interleave) {                              // store without a second buffer set (PICSP_FLAG_NO_SORT): interleave on the host
        const int64_t n = c->n_total[s];
        std::vector<double> tmp((size_t)n * 4);
        double *a = tmp.data();
        int rc = picsp_species_download(c, s, a, a + n, a + 2 * n, a + 3 * n);
        if (rc != PICSP_OK) return rc;
        for (int64_t p = 0; p < n; p++) {
            rows[4 * p + 0] = a[p]; rows[4 * p + 1] = a[n + p];
            rows[4 * p + 2] = a[2 * n + p]; rows[4 * p + 3] = a[3 * n + p];
        }
    }
    check_device_error(c);
    PICSP_API_END
}

int picsp_species_count(picsp_ctx *c, int s, int64_t *n) {
    PICSP_API_BEGIN
    check_ctx(c); check_species(s);
    PICSP_REQUIRE(n != nullptr, PICSP_ERR_INVALID, "null output");
    *n = c->n_total[s];
    PICSP_API_END
}

int picsp_grid_upload(picsp_ctx *c, int which, const double *host) {
    PICSP_API_BEGIN
    check_ctx(c);
    PICSP_REQUIRE(host != nullptr, PICSP_ERR_INVALID, "null host buffer");
    PICSP_CUDA(cudaSetDevice(c->prm.device));
    const Geom &g = c->g;
    const size_t bytes = sizeof(double) * (size_t)g.nn;
    double *dst = nullptr;
    switch (which) {
        case PICSP_DEN_I: dst = c->sp[0].den; break;
        case PICSP_DEN_E: dst = c->sp[1].den; break;
        case PICSP_RHO: dst = c->rho; break;
        case PICSP_PHI: dst = c->phi; break;
        case PICSP_EFX: case PICSP_EFY: break;
        default: throw Error(PICSP_ERR_INVALID, "bad grid selector");
    }
    if (dst) {
        PICSP_CUDA(cudaMemcpyAsync(dst, host, bytes, cudaMemcpyHostToDevice, c->stream));
    } else {
        ensure_stage(c, g.nn);
        PICSP_CUDA(cudaMemcpyAsync(c->stage, host, bytes, cudaMemcpyHostToDevice, c->stream));
        PICSP_LAUNCH(c, k_ef_set_component, blocks_for(g.nn, 256, c->num_sms * 8), 256, 0, c->E, c->stage, g.nn,
                     which == PICSP_EFY ? 1 : 0);
    }
    PICSP_CUDA(cudaStreamSynchronize(c->stream));
    PICSP_API_END
}

int picsp_grid_download(picsp_ctx *c, int which, double *host) {
    PICSP_API_BEGIN
    check_ctx(c);
    PICSP_REQUIRE(host != nullptr, PICSP_ERR_INVALID, "null host buffer");
    PICSP_CUDA(cudaSetDevice(c->prm.device));
    const Geom &g = c->g;
    const size_t bytes = sizeof(double) * (size_t)g.nn;
    const double *src = nullptr;
    switch (which) {
        case PICSP_DEN_I: src = c->sp[0].den; break;
        case PICSP_DEN_E: src = c->sp[1].den; break;
        case PICSP_RHO: src = c->rho; break;
        case PICSP_PHI: src = c->phi; break;
        case PICSP_EFX: case PICSP_EFY: break;
        default: throw Error(PICSP_ERR_INVALID, "bad grid selector");
    }
    if (!src) {
        ensure_stage(c, g.nn);
        PICSP_LAUNCH(c, k_ef_get_component, blocks_for(g.nn, 256, c->num_sms * 8), 256, 0, c->E, c->stage, g.nn,
                     which == PICSP_EFY ? 1 : 0);
        src = c->stage;
    } else if (c->comm && (which == PICSP_DEN_I || which == PICSP_DEN_E)) {
        // Sharded run: every rank accumulates the density of ITS particles only (deposit and fold are linear, so the
        // per-rank accumulation commutes with the sum).  What the reference dumps is the sum over ranks:
        // the download is a collective and every rank receives the global density.
        ensure_stage(c, g.nn);
        PICSP_NCCL(nccl().AllReduce(src, c->stage, (size_t)g.nn, ncclFloat64, ncclSum, c->comm, c->stream));
        src = c->stage;
    }
    PICSP_CUDA(cudaMemcpyAsync(host, src, bytes, cudaMemcpyDeviceToHost, c->stream));
    check_device_error(c);
    PICSP_API_END
}

// ---- hot path ---------------------------------------------------------------------------
#define PICSP_OP(body)                                   \
    PICSP_API_BEGIN                                      \
    check_ctx(c);                                        \
    PICSP_CUDA(cudaSetDevice(c->prm.device));            \
    body;                                                \
    PICSP_API_END

int picsp_deposit(picsp_ctx *c, int s) { PICSP_OP(check_species(s); op_deposit(c, s)) }
int picsp_compute_rho(picsp_ctx *c) { PICSP_OP(op_compute_rho(c)) }
int picsp_solve(picsp_ctx *c) { PICSP_OP(op_solve(c)) }
int picsp_solve_spectral(picsp_ctx *c) {
    PICSP_OP(PICSP_REQUIRE(c->have_plans || c->fft.on, PICSP_ERR_STATE, "context was created with solverType 2: no FFT plans"); op_solve_spectral(c))
}
int picsp_solve_sor(picsp_ctx *c, int64_t *sweeps, double *l2) {
    PICSP_API_BEGIN
    check_ctx(c);
    PICSP_CUDA(cudaSetDevice(c->prm.device));
    op_solve_sor(c);
    if (sweeps || l2) {
        long long *h = reinterpret_cast<long long *>(c->h_pinned + 16);
        PICSP_CUDA(cudaMemcpyAsync(h, c->d_sor_status, sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
        double v = read_scalar(c, c->d_scalars + 3);
        if (sweeps) *sweeps = *h;
        if (l2) *l2 = v;
        if (*h < 0) throw Error(PICSP_ERR_NOT_CONVERGED, "Gauss-Seidel solver failed to converge");   // src/main.cpp:955
    }
    PICSP_API_END
}
int picsp_solve_status(picsp_ctx *c, int64_t *sweeps, double *l2) {
    PICSP_API_BEGIN
    check_ctx(c);
    PICSP_CUDA(cudaSetDevice(c->prm.device));
    long long *h = reinterpret_cast<long long *>(c->h_pinned + 16);
    PICSP_CUDA(cudaMemcpyAsync(h, c->d_sor_status, sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    const double v = read_scalar(c, c->d_scalars + 3);
    if (sweeps) *sweeps = *h;
    if (l2) *l2 = v;
    PICSP_API_END
}
int picsp_compute_ef(picsp_ctx *c) { PICSP_OP(op_compute_ef(c)) }
int picsp_push(picsp_ctx *c, int s) { PICSP_OP(check_species(s); op_push(c, s)) }
int picsp_rewind(picsp_ctx *c, int s) { PICSP_OP(check_species(s); op_rewind(c, s)) }

int picsp_bootstrap(picsp_ctx *c) {   // src/main.cpp:453-472
    PICSP_OP(op_deposit(c, 0); op_deposit(c, 1); op_compute_rho(c); op_solve(c); op_compute_ef(c);
             op_rewind(c, 0); op_rewind(c, 1))
}

static void one_step(picsp_ctx *c) {
    op_grid_phase(c);           // scatterSpecies x2 + computeRho (src/main.cpp:482-490)
    op_solve(c);
    op_compute_ef(c);
    op_push(c, 0); op_push(c, 1);
}

// -- CUDA graph of a pair of steps (launch-bound populations only) ----------------------------------
constexpr long long STEP_GRAPH_MAX_PARTICLES = 1ll << 26;   // above this a step is >= 1 ms of kernels and the launches hide behind them

static bool step_graph_eligible(const picsp_ctx *c) {
    if (c->profiling || c->comm || c->graphs_disabled || c->nparts > 1 || (c->prm.flags & (PICSP_FLAG_NO_GRAPH | PICSP_FLAG_NO_FUSE | PICSP_FLAG_NO_SORT | PICSP_FLAG_WALLS))) return false;
    if (c->sp[0].n + c->sp[1].n > STEP_GRAPH_MAX_PARTICLES) return false;
    for (int s = 0; s < 2; s++) {
        const Species &sp = c->sp[s];
        // steady state of the fused loop, and no re-binning due in either of the two steps
        if (!sp.sorted || !sp.acc_valid || !sp.hist_valid || sp.n <= 0 || !sp.chunk_cnt) return false;
        if (sp.steps_since_sort + 1 >= sp.sort_period) return false;
        if (sp.cell_period > 0 && sp.steps_since_cellsort + 1 >= sp.cell_period) return false;
    }
    return c->smem_opted_in;    // the per-function attributes must not be set during a capture
}

static picsp_ctx::StepGraphKey step_graph_key(const picsp_ctx *c) {
    picsp_ctx::StepGraphKey k;
    memset(&k, 0, sizeof(k));
    int q = 0;
    for (int s = 0; s < 2; s++) {
        const Species &sp = c->sp[s];
        k.ptr[q++] = sp.x; k.ptr[q++] = sp.hist; k.ptr[q++] = sp.hist_next; k.ptr[q++] = sp.chunks; k.ptr[q++] = sp.nchunks;
        k.ptr[q++] = sp.chunk_cnt; k.ptr[q++] = sp.acc;
        k.n[s] = sp.n; k.chunk[s] = sp.chunk;
    }
    return k;
}

static void step_pair_graph(picsp_ctx *c) {
    const picsp_ctx::StepGraphKey key = step_graph_key(c);
    for (int g = 0; g < c->step_graph_count; g++) {
        picsp_ctx::StepGraph &sg = c->step_graphs[g];
        if (memcmp(&key, &sg.key, sizeof(key)) != 0) continue;
        PICSP_CUDA(cudaGraphLaunch(sg.exec, c->stream));
        for (int s = 0; s < 2; s++) {       // the host-side bookkeeping of two fused steps (the buffer swaps cancel out)
            Species &sp = c->sp[s];
            sp.steps_since_sort += 2;
            sp.steps_since_cellsort += 2;
            sp.staged_v_valid = false;
            sp.cnt_valid = true;            // the replayed movers left their per-chunk counts like any other launch
            sp.acc_valid = true; sp.hist_valid = true;
        }
        c->launches += sg.launches;
        c->busy[0] = c->busy[1] = true;
        return;
    }
    // not cached: record the launches of two ordinary steps (nothing runs yet; the host state advances), instantiate, launch.
    // Recording advances the host-side bookkeeping (buffer swaps, validity flags, counters) of a device that has not
    // run anything yet: if the capture or the instantiation fails, that bookkeeping is put back, graphs are switched
    // off for this context and the two steps are enqueued the plain way.
    const int64_t launches0 = c->launches;
    const Species saved_sp[2] = {c->sp[0], c->sp[1]};
    const bool saved_busy[2] = {c->busy[0], c->busy[1]};
    auto give_up = [&]() {
        c->sp[0] = saved_sp[0]; c->sp[1] = saved_sp[1];
        c->busy[0] = saved_busy[0]; c->busy[1] = saved_busy[1];
        c->launches = launches0;
        c->graphs_disabled = true;
        cudaGetLastError();
        one_step(c); one_step(c);
    };
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { give_up(); return; }
    bool recorded = true;
    try {
        one_step(c); one_step(c);
    } catch (...) {
        recorded = false;
    }
    const cudaError_t ec = cudaStreamEndCapture(c->stream, &graph);
    cudaGraphExec_t exec = nullptr;
    if (!recorded || ec != cudaSuccess || !graph || cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
        if (graph) cudaGraphDestroy(graph);
        give_up();
        return;
    }
    cudaGraphDestroy(graph);
    int slot;
    if (c->step_graph_count < picsp_ctx::STEP_GRAPH_CACHE) slot = c->step_graph_count++;
    else {                                 // evict round robin; the old graph may still be running
        slot = c->step_graph_next; c->step_graph_next = (slot + 1) % picsp_ctx::STEP_GRAPH_CACHE;
        c->retired_graphs.push_back(c->step_graphs[slot].exec);
    }
    c->step_graphs[slot].key = key;        // a pair leaves every pointer of the key where it was
    c->step_graphs[slot].exec = exec;
    c->step_graphs[slot].launches = c->launches - launches0;
    PICSP_CUDA(cudaGraphLaunch(exec, c->stream));
}

int picsp_step(picsp_ctx *c, int nsteps) {   // src/main.cpp:481-504
    PICSP_API_BEGIN
    check_ctx(c);
    PICSP_REQUIRE(nsteps >= 0, PICSP_ERR_INVALID, "negative step count");
    PICSP_CUDA(cudaSetDevice(c->prm.device));
    // a violation flagged by the movers of an EARLIER call (mirrored into mapped host memory by the grid phase of
    // the step that followed it) is reported here, before anything else is enqueued
    if (*(volatile int *)c->h_error_mapped) check_device_error(c);
    PhaseScope whole(c, PICSP_PHASE_STEP);
    for (int it = 0; it < nsteps;) {
        if (nsteps - it >= 2 && step_graph_eligible(c)) { step_pair_graph(c); it += 2; }
        else { one_step(c); it++; }
    }
    PICSP_API_END
}

// ---- asynchronous dumps -------------------------------------------------------------------------
// What the reference writes every 50 steps (writeSpecies x2, writePot, computeKE x2; src/main.cpp:507-527) is first
// SNAPSHOT on the device, stream-ordered with the time loop (un-permute into [n][4] rows in list order, copies of
// den / phi, the KE sums), and then copied to the caller's buffers on the copy stream while the next steps run.
static bool ensure_snapshot(picsp_ctx *c) {
    if (c->snapshot_unavailable) return false;
    const Geom &g = c->g;
    if (!c->snap_grids) {
        if (cudaMalloc((void **)&c->snap_grids, sizeof(double) * 3 * (size_t)g.nn) != cudaSuccess ||
            cudaMalloc((void **)&c->snap_ke, sizeof(double) * 2) != cudaSuccess) {
            cudaGetLastError(); c->snapshot_unavailable = true; return false;
        }
    }
    for (int s = 0; s < 2; s++) {
        if (c->snap_rows_cap[s] >= c->n_total[s]) continue;
        cudaFree(c->snap_rows[s]); c->snap_rows[s] = nullptr; c->snap_rows_cap[s] = 0;
        const int64_t cap = std::max<int64_t>(c->cap_total[s], 1);
        if (cudaMalloc((void **)&c->snap_rows[s], sizeof(double) * 4 * (size_t)cap) != cudaSuccess) {
            cudaGetLastError(); c->snapshot_unavailable = true; return false;   // e.g. 4e9 particles on one GPU: synchronous dumps
        }
        c->snap_rows_cap[s] = cap;
    }
    if (!c->ev_snap) {
        PICSP_CUDA(cudaEventCreateWithFlags(&c->ev_snap, cudaEventDisableTiming));
        PICSP_CUDA(cudaEventCreateWithFlags(&c->ev_dump_done, cudaEventDisableTiming));
        PICSP_CUDA(cudaEventCreateWithFlags(&c->ev_dump_done2, cudaEventDisableTiming));
        int lo_prio = 0, hi_prio = 0;
        PICSP_CUDA(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
        PICSP_CUDA(cudaStreamCreateWithPriority(&c->copy_stream2, cudaStreamNonBlocking, hi_prio));
    }
    return true;
}

int picsp_dump_wait(picsp_ctx *c) {
    PICSP_API_BEGIN
    check_ctx(c);
    if (!c->dump_in_flight) return PICSP_OK;
    PICSP_CUDA(cudaSetDevice(c->prm.device));
    PICSP_CUDA(cudaEventSynchronize(c->ev_dump_done));
    PICSP_CUDA(cudaEventSynchronize(c->ev_dump_done2));
    c->dump_in_flight = false;
    if (c->dump_ke_host)
        for (int s = 0; s < 2; s++)      // src/main.cpp:1198: 0.5*spwt*mass is ADDED (Q10); chargeE == 1 (:1201)
            c->dump_ke_host[s] = (c->h_pinned[32 + s] + 0.5 * (c->sp[s].spwt * c->sp[s].m)) / 1.0;
    c->dump_ke_host = nullptr;
    PICSP_API_END
}

int picsp_dump_begin(picsp_ctx *c, double *rows_i, double *rows_e, double *den_i, double *den_e, double *phi, double *ke2) {
    PICSP_API_BEGIN
    check_ctx(c);
    PICSP_CUDA(cudaSetDevice(c->prm.device));
    if (c->dump_in_flight) { int rc = picsp_dump_wait(c); if (rc != PICSP_OK) return rc; }
    const Geom &g = c->g;
    double *rows[2] = {rows_i, rows_e};
    double *den[2] = {den_i, den_e};
    const bool root = !c->comm || c->rank == 0;
    if (!ensure_snapshot(c)) {
        // no room for a snapshot: the same dump, synchronously, through the ordinary downloads
        for (int s = 0; s < 2; s++) {
            if (rows[s]) { int rc = picsp_species_download_rows(c, s, rows[s]); if (rc != PICSP_OK) return rc; }
            if (den[s] || c->comm) {
                ensure_stage(c, g.nn);
                std::vector<double> tmp;
                double *dst = den[s];
                if (!dst) { tmp.resize((size_t)g.nn); dst = tmp.data(); }     // the density download is a collective
                int rc = picsp_grid_download(c, s == 0 ? PICSP_DEN_I : PICSP_DEN_E, dst); if (rc != PICSP_OK) return rc;
            }
            if (ke2) { int rc = picsp_compute_ke(c, s, &ke2[s]); if (rc != PICSP_OK) return rc; }
        }
        if (phi) { int rc = picsp_grid_download(c, PICSP_PHI, phi); if (rc != PICSP_OK) return rc; }
        return PICSP_OK;
    }
    ensure_copy_stream(c);
    const size_t gbytes = sizeof(double) * (size_t)g.nn;
    for (int s = 0; s < 2; s++) {
        Species &sp = c->sp[s];
        if (rows[s] || ke2)      // (the parts' rows are contiguous in the snapshot: only the last part may be short)
            for_parts(c, s, [&](Species &q) { rows_of_part(c, q, c->snap_rows[s] + 4 * q.first); });
        if (ke2) {
            PICSP_LAUNCH(c, k_ke_rows_partial, RED_BLOCKS, RED_THREADS, 0, c->snap_rows[s], (long long)c->n_total[s], c->d_red);
            PICSP_LAUNCH(c, k_sum_final, 1, 1024, 0, c->d_red, RED_BLOCKS, c->snap_ke + s);
        }
        // density: the sum over ranks lands on rank 0 only ("at dump steps only: reduce den_i, den_e to rank 0")
        if (c->comm)
            PICSP_NCCL(nccl().Reduce(sp.den, c->snap_grids + (size_t)s * g.nn, (size_t)g.nn, ncclFloat64, ncclSum, 0, c->comm, c->stream));
        else if (den[s])
            PICSP_CUDA(cudaMemcpyAsync(c->snap_grids + (size_t)s * g.nn, sp.den, gbytes, cudaMemcpyDeviceToDevice, c->stream));
    }
    if (ke2 && c->comm) PICSP_NCCL(nccl().AllReduce(c->snap_ke, c->snap_ke, 2, ncclFloat64, ncclSum, c->comm, c->stream));
    if (phi && root) PICSP_CUDA(cudaMemcpyAsync(c->snap_grids + 2 * (size_t)g.nn, c->phi, gbytes, cudaMemcpyDeviceToDevice, c->stream));
    c->busy[0] = c->busy[1] = true;
    PICSP_CUDA(cudaEventRecord(c->ev_snap, c->stream));
    PICSP_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_snap, 0));
    // small things first, so that a caller polling the grids does not wait behind 32 GB of phase space
    if (ke2) PICSP_CUDA(cudaMemcpyAsync(c->h_pinned + 32, c->snap_ke, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->copy_stream));
    if (root) {
        for (int s = 0; s < 2; s++)
            if (den[s]) PICSP_CUDA(cudaMemcpyAsync(den[s], c->snap_grids + (size_t)s * g.nn, gbytes, cudaMemcpyDeviceToHost, c->copy_stream));
        if (phi) PICSP_CUDA(cudaMemcpyAsync(phi, c->snap_grids + 2 * (size_t)g.nn, gbytes, cudaMemcpyDeviceToHost, c->copy_stream));
    }
    // the two species' rows go out on two streams (two copy engines): under a memory-bound mover one engine alone
    // does not fill the PCIe link
    PICSP_CUDA(cudaStreamWaitEvent(c->copy_stream2, c->ev_snap, 0));
    // (a copy KERNEL on a high-priority stream instead of the copy engines was tried: it takes SM slots from the mover and
    // is no faster over PCIe — e2e 5.1e10 -> 4.9e10 / 4.3e10 / 3.5e10 with 8 / 32 / 128 CTAs, profiles/r02_e2e_dumps.md)
    for (int s = 0; s < 2; s++)
        if (rows[s] && c->n_total[s] > 0)
            PICSP_CUDA(cudaMemcpyAsync(rows[s], c->snap_rows[s], sizeof(double) * 4 * (size_t)c->n_total[s], cudaMemcpyDeviceToHost,
                                       s == 0 ? c->copy_stream : c->copy_stream2));
    PICSP_CUDA(cudaEventRecord(c->ev_dump_done, c->copy_stream));
    PICSP_CUDA(cudaEventRecord(c->ev_dump_done2, c->copy_stream2));
    c->dump_in_flight = true;
    c->dump_ke_host = ke2;
    PICSP_API_END
}

void *picsp_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void picsp_host_free(void *p) { if (p) cudaFreeHost(p); }

// ---- diagnostics -------------------------------------------------------------------------
int picsp_compute_ke(picsp_ctx *c, int s, double *ke) {
    PICSP_API_BEGIN
    check_ctx(c); check_species(s);
    PICSP_REQUIRE(ke != nullptr, PICSP_ERR_INVALID, "null output");
    PICSP_CUDA(cudaSetDevice(c->prm.device));
    Species &sp = c->sp[s];
    for_parts(c, s, [&](Species &q) {
        borrow_spare(c, q);
        if (q.has_perm && q.n > 0 && q.staged_v_valid) {
            // the download that preceded this call left the velocities in upload order in the staging buffers
            PICSP_LAUNCH(c, k_ke_partial, RED_BLOCKS, RED_THREADS, 0, q.vx2, q.vy2, (long long)q.n, c->d_red);
        } else if (q.has_perm && q.n > 0) {
            // sorted store: reduce in upload order so the sum is reproducible regardless of the storage order
            // (staging = the idle half of the sort's ping-pong buffers, dead between sorts)
            PICSP_LAUNCH(c, k_ke_terms, particle_blocks(c, q.n, 256), 256, 0, q.vx, q.vy, q.id, (long long)q.n, q.x2);
            PICSP_LAUNCH(c, k_sum_partial, RED_BLOCKS, RED_THREADS, 0, q.x2, (long long)q.n, c->d_red);
        } else {
            PICSP_LAUNCH(c, k_ke_partial, RED_BLOCKS, RED_THREADS, 0, q.vx, q.vy, (long long)q.n, c->d_red);
        }
        // one part: the species' sum; several: per-part sums, added up (in part order) below
        PICSP_LAUNCH(c, k_sum_final, 1, 1024, 0, c->d_red, RED_BLOCKS, c->nparts > 1 ? c->d_part_sums + q.part : c->d_scalars + 0);
    });
    if (c->nparts > 1) PICSP_LAUNCH(c, k_sum_final, 1, 1024, 0, c->d_part_sums, c->nparts, c->d_scalars + 0);
    if (c->comm) PICSP_NCCL(nccl().AllReduce(c->d_scalars, c->d_scalars, 1, ncclFloat64, ncclSum, c->comm, c->stream));
    double sum = read_scalar(c, c->d_scalars + 0);
    sum += 0.5 * (sp.spwt * sp.m);   // src/main.cpp:1198: added, not multiplied (Q10)
    sum /= 1.0;                      // chargeE == 1 after normalisation, src/main.cpp:1201
    *ke = sum;
    PICSP_API_END
}

int picsp_delta_phi(picsp_ctx *c, double *max_phi, double *phi0) {
    PICSP_API_BEGIN
    check_ctx(c);
    PICSP_CUDA(cudaSetDevice(c->prm.device));
    PICSP_LAUNCH(c, k_max_partial, RED_BLOCKS, RED_THREADS, 0, c->phi, c->g.nn, c->d_red);
    PICSP_LAUNCH(c, k_max_final, 1, 1024, 0, c->d_red, RED_BLOCKS, c->phi, c->d_scalars + 1, c->d_scalars + 2);
    PICSP_CUDA(cudaMemcpyAsync(c->h_pinned, c->d_scalars + 1, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    PICSP_CUDA(cudaStreamSynchronize(c->stream));
    if (max_phi) *max_phi = c->h_pinned[0];
    if (phi0) *phi0 = c->h_pinned[1];
    PICSP_API_END
}

int picsp_repush_count(picsp_ctx *c, int s, int64_t *n) {
    PICSP_API_BEGIN
    check_ctx(c); check_species(s);
    PICSP_REQUIRE(n != nullptr, PICSP_ERR_INVALID, "null output");
    PICSP_CUDA(cudaSetDevice(c->prm.device));
    unsigned long long *h = reinterpret_cast<unsigned long long *>(c->h_pinned + 24);
    *n = 0;
    for_parts(c, s, [&](Species &q) {
        PICSP_CUDA(cudaMemcpyAsync(h, q.counters, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
        PICSP_CUDA(cudaStreamSynchronize(c->stream));
        *n += (int64_t)*h;
    });
    PICSP_API_END
}

int picsp_straggler_count(picsp_ctx *c, int s, int64_t *n) {
    PICSP_API_BEGIN
    check_ctx(c); check_species(s);
    PICSP_REQUIRE(n != nullptr, PICSP_ERR_INVALID, "null output");
    PICSP_CUDA(cudaSetDevice(c->prm.device));
    unsigned long long *h = reinterpret_cast<unsigned long long *>(c->h_pinned + 24);
    *n = 0;
    for_parts(c, s, [&](Species &q) {
        PICSP_CUDA(cudaMemcpyAsync(h, q.counters + 1, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
        PICSP_CUDA(cudaStreamSynchronize(c->stream));
        *n += (int64_t)*h;
    });
    PICSP_API_END
}

int picsp_set_sort_period(picsp_ctx *c, int s, int period) {
    PICSP_API_BEGIN
    check_ctx(c); check_species(s);
    for_parts(c, s, [&](Species &q) { q.sort_period = period > 0 ? period : (s == 0 ? 96 : 8); });
    PICSP_API_END
}

int picsp_set_cell_sort_period(picsp_ctx *c, int s, int period) {
    PICSP_API_BEGIN
    check_ctx(c); check_species(s);
    for_parts(c, s, [&](Species &q) {
        q.cell_period = period > 0 ? period : 0;
        q.steps_since_cellsort = q.cell_period;       // due at the next push
    });
    PICSP_API_END
}

int picsp_set_bank_order(picsp_ctx *c, int s, int mode) {
    PICSP_API_BEGIN
    check_ctx(c); check_species(s);
    PICSP_REQUIRE(mode >= -1 && mode <= 1, PICSP_ERR_INVALID, "bank order mode must be -1 (automatic), 0 (off) or 1 (on)");
    for_parts(c, s, [&](Species &q) { q.bank_order = mode; });
    PICSP_API_END
}

int picsp_set_deposit_aggregation(picsp_ctx *c, int s, int mode) {
    PICSP_API_BEGIN
    check_ctx(c); check_species(s);
    PICSP_REQUIRE(mode >= -1 && mode <= 1, PICSP_ERR_INVALID, "aggregation mode must be -1 (automatic), 0 (off) or 1 (on)");
    for_parts(c, s, [&](Species &q) { q.aggregate = mode; });
    PICSP_API_END
}

// ---- multi-GPU ---------------------------------------------------------------------------
int picsp_comm_unique_id(void *id128) {
    PICSP_API_BEGIN
    PICSP_REQUIRE(id128 != nullptr, PICSP_ERR_INVALID, "null id");
    static_assert(sizeof(ncclUniqueId) == 128, "NCCL unique id is 128 bytes");
    ncclUniqueId id;
    PICSP_NCCL(nccl().GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    PICSP_API_END
}

int picsp_comm_attach(picsp_ctx *c, const void *id128, int rank, int nranks) {
    PICSP_API_BEGIN
    check_ctx(c);
    PICSP_REQUIRE(id128 && nranks >= 1 && rank >= 0 && rank < nranks, PICSP_ERR_INVALID, "bad communicator arguments");
    PICSP_REQUIRE(c->comm == nullptr, PICSP_ERR_STATE, "communicator already attached");
    PICSP_CUDA(cudaSetDevice(c->prm.device));
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    PICSP_NCCL(nccl().CommInitRank(&c->comm, nranks, id, rank));
    c->rank = rank; c->nranks = nranks;
    peer_setup(c);
    PICSP_API_END
}

int picsp_comm_peer_reduction(picsp_ctx *c) { return (c && c->peer_ok) ? 1 : 0; }

int picsp_comm_barrier(picsp_ctx *c) {
    PICSP_API_BEGIN
    check_ctx(c);
    PICSP_CUDA(cudaSetDevice(c->prm.device));
    if (c->comm) PICSP_NCCL(nccl().AllReduce(c->d_scalars + 4, c->d_scalars + 4, 1, ncclFloat64, ncclSum, c->comm, c->stream));
    PICSP_CUDA(cudaStreamSynchronize(c->stream));
    PICSP_API_END
}

// ---- bench-only loader ---------------------------------------------------------------------
int picsp_species_fill_synthetic(picsp_ctx *c, int s, int64_t n, int64_t first_index, uint64_t seed, double vth, double xdrift) {
    PICSP_API_BEGIN
    check_ctx(c); check_species(s);
    PICSP_CUDA(cudaSetDevice(c->prm.device));
    PICSP_REQUIRE(n >= 0 && n <= c->cap_total[s], PICSP_ERR_INVALID, "particle count exceeds capacity");
    if (c->sp[s].acc_valid) PICSP_CUDA(cudaMemsetAsync(c->sp[s].acc, 0, sizeof(long long) * c->g.nn, c->stream));
    c->n_total[s] = n;
    for_parts(c, s, [&](Species &sp) {
        const int64_t np = part_count(sp, n);
        if (np > 0)
            PICSP_LAUNCH(c, k_fill_synthetic, particle_blocks(c, np, 256), 256, 0, sp.x, sp.y, sp.vx, sp.vy, (long long)np,
                         (long long)(first_index + sp.first), seed, c->g.xl, c->g.yl, vth, xdrift);
        reset_part_state(c, sp, np);
    });
    PICSP_CUDA(cudaStreamSynchronize(c->stream));
    PICSP_API_END
}

// ---- instrumentation -------------------------------------------------------------------------
int picsp_profile_enable(picsp_ctx *c, int on) {
    PICSP_API_BEGIN
    check_ctx(c);
    if (c->profiling && !on) profile_collect(c);
    c->profiling = on != 0;
    PICSP_API_END
}
int picsp_profile_get(picsp_ctx *c, int phase, double *ms, int64_t *calls) {
    PICSP_API_BEGIN
    check_ctx(c);
    PICSP_REQUIRE(phase >= 0 && phase < PICSP_PHASE_COUNT, PICSP_ERR_INVALID, "bad phase");
    PICSP_CUDA(cudaSetDevice(c->prm.device));
    PICSP_CUDA(cudaStreamSynchronize(c->stream));
    profile_collect(c);
    if (ms) *ms = c->timers[phase].ms;
    if (calls) *calls = c->timers[phase].calls;
    PICSP_API_END
}
int picsp_profile_reset(picsp_ctx *c) {
    PICSP_API_BEGIN
    check_ctx(c);
    PICSP_CUDA(cudaStreamSynchronize(c->stream));
    profile_collect(c);
    for (auto &t : c->timers) { t.ms = 0.0; t.calls = 0; }
    PICSP_API_END
}
int picsp_fft_plan_query(int M, int32_t *out5) {
    PICSP_API_BEGIN
    PICSP_REQUIRE(M >= 2 && out5 != nullptr, PICSP_ERR_INVALID, "bad arguments");
    int P = 0, Q = 0, L = 0, logL = 0;
    if (choose_direct_split(M, &P, &Q)) { out5[0] = 1; out5[1] = P; out5[2] = Q; out5[3] = 0; }
    else if (choose_fft_split(M, &P, &Q, &L, &logL)) { out5[0] = 0; out5[1] = P; out5[2] = Q; out5[3] = L; }
    else { out5[0] = -1; out5[1] = out5[2] = out5[3] = 0; }
    out5[4] = largest_prime_factor(M);
    PICSP_API_END
}
int picsp_spectral_engine(picsp_ctx *c, int *own) {
    PICSP_API_BEGIN
    check_ctx(c);
    PICSP_REQUIRE(own != nullptr, PICSP_ERR_INVALID, "null output");
    *own = c->fft.on ? 1 : 0;
    PICSP_API_END
}
int picsp_parts(picsp_ctx *c, int *parts) {
    PICSP_API_BEGIN
    check_ctx(c);
    PICSP_REQUIRE(parts != nullptr, PICSP_ERR_INVALID, "null output");
    *parts = c->nparts;
    PICSP_API_END
}
int picsp_kernel_launches(picsp_ctx *c, int64_t *n) {
    PICSP_API_BEGIN
    check_ctx(c);
    PICSP_REQUIRE(n != nullptr, PICSP_ERR_INVALID, "null output");
    *n = c->launches;
    PICSP_API_END
}

}  // extern "C"
