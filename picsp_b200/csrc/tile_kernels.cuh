// picsp_b200/csrc/tile_kernels.cuh — the tile-binned fast path of the particle loop.
//
// Particles are kept binned by TILE x TILE-cell tile (periodic counting sort, every few
// steps).  One CTA processes one CHUNK (<= 4096 consecutive particles of one bin):
//   * the E field of the tile plus a HALO-cell ring is staged in shared memory by ONE TMA
//     tensor copy (cp.async.bulk.tensor.2d, zero-filled outside the grid),
//   * the mover (pushSpecies + gather, src/main.cpp:772-847, :671-681) gathers from that
//     window, writes the particle back in place (32 B read + 32 B written),
//   * the next step's CIC deposit (scatterSpecies, src/main.cpp:684-700) of the position just
//     written goes into a shared-memory window of 64-bit FIXED-POINT accumulators held as two
//     32-bit limbs and updated with native 32-bit shared atomics plus an exact carry
//     (64-bit shared atomics are CAS loops on sm_100a; see add64_limbs),
//   * the window is flushed to the global int64 accumulator grid with native 64-bit
//     integer REDs.  Integer addition commutes, so the density is bit-identical for any
//     particle order, chunking or launch order.
// Particles that drifted out of their bin's window since the last sort ("stragglers"), or
// that wrap through the periodic boundary, take the global path (same arithmetic).
#pragma once
#include <cuda.h>   // CUtensorMap (type only; the encoder is fetched at run time)

#include "ctx.cuh"
#include "particle_kernels.cuh"

namespace picsp {

#ifndef PICSP_REPL
#define PICSP_REPL 1          // bank-staggered copies of the shared-memory window (1 = plain window, 4 = experiment)
#endif
#ifndef PICSP_BULK_PIPE
#define PICSP_BULK_PIPE 1     // particle data streamed through shared memory by TMA bulk copies (0: per-thread register prefetch)
#endif
#ifndef PICSP_STAGES
#define PICSP_STAGES 2
#endif
#ifndef PICSP_HALO
#define PICSP_HALO 4
#endif
#ifndef PICSP_CHUNK
#define PICSP_CHUNK 4096
#endif

constexpr int HALO = PICSP_HALO;              // cells of drift a window tolerates on each side
constexpr int WIN = TILE + 1 + 2 * HALO;      // window edge in nodes (25)
// Bank-conflict avoidance by replication.  Particle positions inside a tile are random, so the 16-byte E
// loads of a quarter-warp (8 lanes over 8 bank groups) and the 4-byte accumulator atomics of a warp
// (32 lanes over 32 banks) collide like birthdays: 2.7 and 3.7 wavefronts per access measured.  The window
// is therefore kept REPL times, copy r shifted by 2r bank groups (E) / 8r banks (accumulators); every lane
// computes its rank among the lanes that would collide with it (one VOTE / one MATCH) and takes the copy
// that lands it on a bank of its own.  All copies hold the same E and the accumulator copies are summed at
// the flush, so ANY choice is correct: the choice only decides speed.
constexpr int REPL = PICSP_REPL;
// Row pitch of the window in nodes.  With pitch 26 the 16 x 16 cells of a tile spread EVENLY over the 32 residues of
// (row * pitch + column) mod 32 (8 cells each; with pitch 25 it is 6..10), which suits k_bank_order — but the unordered
// electrons run 2 % slower with it (same-box A/B, profiles/r02b_mover_bank_order.md), so the pitch stays WIN.
#ifndef PICSP_WIN_PAD
#define PICSP_WIN_PAD 0       // 1: pitch 26 (even classes for k_bank_order) — measured 2 % slower for unordered particles, so 25 it is
#endif
constexpr int WPITCH = (REPL > 1) ? WIN + 2 * (REPL - 1) : WIN + PICSP_WIN_PAD;    // E row pitch (replicated: copy r starts 2r nodes in)
constexpr int APITCH = (REPL > 1) ? WIN : WIN + PICSP_WIN_PAD;                     // accumulator row pitch in words
constexpr int E_COPY = ((WIN * WPITCH + 7) / 8) * 8;                   // nodes per E copy: multiple of 8 bank groups
constexpr int ACC_COPY = (REPL > 1) ? (((WIN * WIN + 23) / 32) * 32 + 8) : WIN * APITCH;   // words per accumulator copy (replicated: == 8 (mod 32))
static_assert(REPL == 1 || REPL == 4, "REPL must be 1 or 4");
static_assert(REPL > 1 || APITCH == WPITCH, "k_bank_order assumes one pitch for the E window and the accumulators");
static_assert(REPL == 1 || (ACC_COPY % 32 == 8 && ACC_COPY >= WIN * WIN), "accumulator copy stride must be 8 mod 32");
constexpr size_t MOVER_WINDOW_BYTES = ((sizeof(double2) * (size_t)REPL * E_COPY + 2 * sizeof(unsigned) * (size_t)REPL * ACC_COPY + 127) / 128) * 128;
constexpr int CHUNK = PICSP_CHUNK;            // particles per CTA work item
constexpr int MAX_FRAC_TILED = 51;            // w*2^frac <= 2^51 keeps (w*2^frac + 2^52) below 2^53: the magic-number conversion is exact
#ifndef PICSP_MOVER_THREADS
#define PICSP_MOVER_THREADS (PICSP_REPL > 1 ? 512 : 128)
#endif
constexpr int MOVER_THREADS = PICSP_MOVER_THREADS;
#ifndef PICSP_MOVER_MIN_CTAS
#define PICSP_MOVER_MIN_CTAS (PICSP_REPL > 1 ? 2 : 8)     // 8 x 128 threads: small CTAs interleave their load / compute / flush phases best (profiles/r01_sweeps.md)
#endif
constexpr int MOVER_MIN_CTAS = PICSP_MOVER_MIN_CTAS;
// Particle pipeline: the chunk is consumed in slices of MOVER_THREADS particles.  One elected thread keeps
// STAGES-1 slices in flight with TMA bulk copies (cp.async.bulk, completion on an mbarrier per stage); the
// threads read their particle from shared memory.  Global-load latency is then covered by the copy engine,
// not by registers or occupancy (the register-prefetch version spent 38 % of its stall samples on it).
constexpr bool BULK_PIPE = PICSP_BULK_PIPE != 0;
constexpr int STAGES = PICSP_STAGES;
constexpr int STAGE_W = MOVER_THREADS + 2;          // doubles per array per stage (+2: a slice may start at an odd index)
constexpr int STAGE_IDW = MOVER_THREADS + 8;        // u32 ids per stage (+8: the slice is widened to a multiple of 4 ids at both ends)
constexpr size_t MOVER_PIPE_BYTES = sizeof(double) * (size_t)STAGES * 4 * STAGE_W;
constexpr size_t MOVER_SMEM_BYTES = MOVER_WINDOW_BYTES + (BULK_PIPE ? MOVER_PIPE_BYTES + sizeof(uint32_t) * (size_t)STAGES * STAGE_IDW : 0) + 64;

struct __align__(16) Chunk {
    long long start;   // first particle (index into the species arrays)
    int count;         // <= CHUNK
    int tile;          // bin (tx * nty + ty)
};

// ---------------------------------------------------------------------------
// binning: exclusive scan of the tile histogram -> tile offsets + chunk table
// k_scan_tiles (one CTA; ntiles <= 128*128): block-wide scan of the populations and of the chunk counts ->
// tile_off[], first chunk of every bin, number of chunks, zeroed cursors.  k_fill_chunks (one warp per bin) then writes
// the bin's chunk entries in parallel.  (Round 1 did both in one kernel with a serial prefix on thread 0 and serial
// per-thread chunk writes: 178 us per re-binning at 4096 bins with ~30 chunks each; now ~10 us.)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024, 1)
k_scan_tiles(const unsigned int *__restrict__ hist, int ntiles, long long *__restrict__ tile_off,
             int *__restrict__ chunk0, int *__restrict__ nchunks, unsigned int *__restrict__ cursor, int chunk) {
    __shared__ long long s_part[32];
    __shared__ int s_chunk[32];
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, w = tid >> 5;
    const int per = (ntiles + nt - 1) / nt;
    const int lo = tid * per, hi = min(lo + per, ntiles);
    long long sum = 0; int csum = 0;
    for (int t = lo; t < hi; t++) { unsigned int h = hist[t]; sum += h; csum += (int)((h + chunk - 1) / chunk); }
    // inclusive scan over the threads: inside the warps, then over the warp totals
    long long isum = sum; int icsum = csum;
    for (int o = 1; o < 32; o <<= 1) {
        const long long v = __shfl_up_sync(0xffffffffu, isum, o); const int u = __shfl_up_sync(0xffffffffu, icsum, o);
        if (lane >= o) { isum += v; icsum += u; }
    }
    if (lane == 31) { s_part[w] = isum; s_chunk[w] = icsum; }
    __syncthreads();
    if (w == 0) {
        long long v = lane < nt / 32 ? s_part[lane] : 0, iv = v; int u = lane < nt / 32 ? s_chunk[lane] : 0, iu = u;
        for (int o = 1; o < 32; o <<= 1) {
            const long long a = __shfl_up_sync(0xffffffffu, iv, o); const int b = __shfl_up_sync(0xffffffffu, iu, o);
            if (lane >= o) { iv += a; iu += b; }
        }
        s_part[lane] = iv - v; s_chunk[lane] = iu - u;            // exclusive warp offsets
        if (lane == 31) { tile_off[ntiles] = iv; *nchunks = iu; }
    }
    __syncthreads();
    long long off = s_part[w] + isum - sum; int coff = s_chunk[w] + icsum - csum;
    for (int t = lo; t < hi; t++) {
        const unsigned int h = hist[t];
        tile_off[t] = off;
        chunk0[t] = coff;
        cursor[t] = 0u;
        off += h; coff += (int)((h + chunk - 1) / chunk);
    }
}

// one warp per bin: its chunk entries {first particle, count <= chunk, bin}
__global__ void __launch_bounds__(256)
k_fill_chunks(const unsigned int *__restrict__ hist, int ntiles, const long long *__restrict__ tile_off,
              const int *__restrict__ chunk0, Chunk *__restrict__ chunks, int chunk) {
    const int t = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (t >= ntiles) return;
    const unsigned int h = hist[t];
    const int nch = (int)((h + chunk - 1) / chunk), c0 = chunk0[t];
    const long long off = tile_off[t];
    for (int k = lane; k < nch; k += 32) {
        Chunk ck; ck.start = off + (long long)k * chunk; ck.count = (int)min((unsigned int)chunk, h - (unsigned int)k * (unsigned int)chunk); ck.tile = t;
        chunks[c0 + k] = ck;
    }
}

// out-of-place scatter into the bins.  Slots inside a bin are handed out by an atomic
// cursor (warp-aggregated per destination tile), so the order inside a bin is arbitrary —
// harmless: the mover is per-particle and the deposit is order-independent.
__global__ void __launch_bounds__(256)
k_sort_scatter(const double *__restrict__ x, const double *__restrict__ y, const double *__restrict__ vx,
               const double *__restrict__ vy, const uint32_t *__restrict__ id, long long n, PushConst c,
               const long long *__restrict__ tile_off, unsigned int *__restrict__ cursor,
               double *__restrict__ x2, double *__restrict__ y2, double *__restrict__ vx2, double *__restrict__ vy2,
               uint32_t *__restrict__ id2) {
    const int lane = threadIdx.x & 31;
    // Each CTA takes one CONTIGUOUS slice of the array.  With a grid-stride loop all resident CTAs would
    // work on neighbouring particles, i.e. on the same handful of bins, and their cursor atomics would
    // serialise on a few L2 addresses; contiguous slices put every CTA in a different bin.
    const long long per_cta = ((n + gridDim.x - 1) / gridDim.x + 31) & ~31ll;
    const long long lo = blockIdx.x * per_cta, hi = min(n, lo + per_cta);
    // software pipeline: all five loads of the NEXT particle are in flight while the current one takes its slot
    long long base = lo + (threadIdx.x - lane);
    double px = 0, py = 0, pvx = 0, pvy = 0;
    uint32_t pid = 0;
    if (base + lane < hi) {
        const long long p = base + lane;
        px = x[p]; py = y[p]; pvx = vx[p]; pvy = vy[p]; pid = id ? id[p] : (uint32_t)p;
    }
    for (; base < hi; base += blockDim.x) {
        const long long p = base + lane;
        const bool live = p < hi;
        const long long pn = p + blockDim.x;
        double nx = 0, ny = 0, nvx = 0, nvy = 0;
        uint32_t nid = 0;
        if (pn < hi) { nx = x[pn]; ny = y[pn]; nvx = vx[pn]; nvy = vy[pn]; nid = id ? id[pn] : (uint32_t)pn; }
        const int t = live ? tile_of(px, py, c) : -1;
        const unsigned int peers = __match_any_sync(0xffffffffu, t);
        const int leader = __ffs(peers) - 1;
        const int rank = __popc(peers & ((1u << lane) - 1u));
        unsigned int first = 0;
        if (live && lane == leader) first = atomicAdd(&cursor[t], (unsigned int)__popc(peers));
        first = __shfl_sync(0xffffffffu, first, leader);
        if (live) {
            const long long dst = tile_off[t] + first + rank;
            x2[dst] = px; y2[dst] = py; vx2[dst] = pvx; vy2[dst] = pvy; id2[dst] = pid;
        }
        px = nx; py = ny; pvx = nvx; pvy = nvy; pid = nid;
    }
}

// First binning of a LARGE arbitrary load in two passes.  A one-pass scatter of randomly ordered particles sends
// every 8-byte store of a warp to a different bin: 2.5e9 single-sector writes for 5e8 particles (56 ms, 0.64 TB/s).
// Here pass A groups the particles by COARSE range (up to 64 contiguous ranges of bins; tpc = bins per range, a
// power of two) and pass B orders each range by bin.  In both passes one CTA takes 4096 consecutive source
// particles, which fall into few classes (<= 64 coarse ranges; the <= 2*tpc bins of the one or two ranges the slice
// touches); it ranks them per class (warp-aggregated shared-memory counters), reserves ONE contiguous range per
// class with a single global atomic, and then moves each array through shared memory IN CLASS ORDER, so that
// consecutive threads store to consecutive addresses (runs of ~64 particles).  Pass A goes from the primary arrays
// to the second set, pass B back: the result is in the primary arrays.
constexpr int SORT2_THREADS = 512;
constexpr int SORT2_SLICE = 4096;
constexpr int SORT2_MAXCLS = 1024;            // shared-memory counters; classes beyond take an individual slot
constexpr size_t SORT2_SMEM_BYTES = sizeof(double) * SORT2_SLICE + sizeof(unsigned) * (2 * SORT2_SLICE + 3 * SORT2_MAXCLS);

template <bool COARSE>
__global__ void __launch_bounds__(SORT2_THREADS, 2)
k_sort_pass(const double *__restrict__ x, const double *__restrict__ y, const double *__restrict__ vx,
            const double *__restrict__ vy, const uint32_t *__restrict__ id, long long n, PushConst c, int tpc_shift,
            const long long *__restrict__ tile_off, unsigned int *__restrict__ cursor,
            double *__restrict__ x2, double *__restrict__ y2, double *__restrict__ vx2, double *__restrict__ vy2,
            uint32_t *__restrict__ id2) {
    extern __shared__ __align__(16) unsigned char sort_smem[];
    double *s_val = reinterpret_cast<double *>(sort_smem);                  // [SLICE] one array of the slice, in class order
    unsigned *s_dst = reinterpret_cast<unsigned *>(s_val + SORT2_SLICE);     // [SLICE] destination slot of class-ordered element q
    unsigned *s_pos = s_dst + SORT2_SLICE;                                   // [SLICE] class | rank, later class-ordered position, of source element k
    unsigned *s_cnt = s_pos + SORT2_SLICE;                                   // [MAXCLS]
    unsigned *s_base = s_cnt + SORT2_MAXCLS;                                 // [MAXCLS] first slot reserved in the destination bin / range
    unsigned *s_pref = s_base + SORT2_MAXCLS;                                // [MAXCLS] exclusive prefix of s_cnt
    __shared__ int s_first, s_total;
    const int tid = threadIdx.x, lane = tid & 31;
    const int nt = c.ntx * c.nty;
    const long long lo = (long long)blockIdx.x * SORT2_SLICE;
    const int count = (int)min((long long)SORT2_SLICE, n - lo);
    if (count <= 0) return;
    const int ncls = COARSE ? min(SORT2_MAXCLS, ((nt - 1) >> tpc_shift) + 1) : min(SORT2_MAXCLS, 2 << tpc_shift);
    for (int k = tid; k < ncls; k += SORT2_THREADS) s_cnt[k] = 0u;
    // pass B: classes are bins relative to the first bin of the coarse range the slice starts in (the source is
    // ordered by coarse range, so every particle of the slice is in that range or a later one)
    if (!COARSE && tid == 0) s_first = (tile_of(x[lo], y[lo], c) >> tpc_shift) << tpc_shift;
    __syncthreads();
    const int first = COARSE ? 0 : s_first;

    // 1. class and rank of every particle
    for (int k0 = 0; k0 < count; k0 += SORT2_THREADS) {          // warp-uniform trip count
        const int k = k0 + tid;
        const bool live = k < count;
        int cls = -1;
        if (live) {
            const long long p = lo + k;
            const int t = tile_of(x[p], y[p], c);
            cls = COARSE ? (t >> tpc_shift) : (t - first);
            if (cls >= ncls) {            // more classes than counters (tiny coarse ranges of a very non-uniform load): individual slot
                const int u = COARSE ? (t >> tpc_shift) : t;          // cursor index: coarse range or bin
                const long long base = COARSE ? tile_off[min(u << tpc_shift, nt)] : tile_off[t];
                const long long dst = base + atomicAdd(&cursor[u], 1u);
                x2[dst] = x[p]; y2[dst] = y[p]; vx2[dst] = vx[p]; vy2[dst] = vy[p];
                id2[dst] = id ? id[p] : (uint32_t)p;
                cls = -1;
            }
        }
        const unsigned peers = __match_any_sync(0xffffffffu, cls);
        const int leader = __ffs(peers) - 1;
        unsigned r0 = 0;
        if (cls >= 0 && lane == leader) r0 = atomicAdd(&s_cnt[cls], (unsigned)__popc(peers));
        r0 = __shfl_sync(0xffffffffu, r0, leader);
        if (live) s_pos[k] = cls < 0 ? 0xFFFFFFFFu : ((unsigned)cls << 16) | (r0 + (unsigned)__popc(peers & ((1u << lane) - 1u)));   // rank < 4096 < 2^16
    }
    __syncthreads();
    // 2. one reservation per class; exclusive prefix of the class populations (warp 0: 32 classes per lane)
    for (int k = tid; k < ncls; k += SORT2_THREADS)
        if (s_cnt[k]) s_base[k] = atomicAdd(&cursor[COARSE ? k : first + k], s_cnt[k]);
    if (tid < 32) {
        const int per = (ncls + 31) / 32, k0 = tid * per, k1 = min(k0 + per, ncls);
        unsigned sum = 0;
        for (int k = k0; k < k1; k++) sum += s_cnt[k];
        unsigned incl = sum;
        for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        unsigned run = incl - sum;
        for (int k = k0; k < k1; k++) { s_pref[k] = run; run += s_cnt[k]; }
        if (tid == 31) s_total = (int)incl;
    }
    __syncthreads();
    // 3. class-ordered position of every source element and the destination slot of every position
    for (int k = tid; k < count; k += SORT2_THREADS) {
        const unsigned code = s_pos[k];
        if (code == 0xFFFFFFFFu) continue;
        const int cls = (int)(code >> 16);
        const unsigned rank = code & 0xFFFFu;
        const long long base = COARSE ? tile_off[min(cls << tpc_shift, nt)] : tile_off[first + cls];
        const unsigned q = s_pref[cls] + rank;
        s_pos[k] = q;
        s_dst[q] = (unsigned)(base + s_base[cls] + rank);       // < 2^32: per-rank species capacity
    }
    __syncthreads();
    // 4. each array: coalesced read -> shared memory in class order -> runs of consecutive stores
    const int total = s_total;
    auto move = [&](const double *__restrict__ src, double *__restrict__ dstp) {
        for (int k = tid; k < count; k += SORT2_THREADS) { const unsigned q = s_pos[k]; if (q != 0xFFFFFFFFu) s_val[q] = src[lo + k]; }
        __syncthreads();
        for (int q = tid; q < total; q += SORT2_THREADS) dstp[s_dst[q]] = s_val[q];
        __syncthreads();
    };
    move(x, x2); move(y, y2); move(vx, vx2); move(vy, vy2);
    unsigned *s_ival = reinterpret_cast<unsigned *>(s_val);
    for (int k = tid; k < count; k += SORT2_THREADS) { const unsigned q = s_pos[k]; if (q != 0xFFFFFFFFu) s_ival[q] = id ? id[lo + k] : (uint32_t)(lo + k); }
    __syncthreads();
    for (int q = tid; q < total; q += SORT2_THREADS) id2[s_dst[q]] = s_ival[q];
}

// Re-sort of an already binned store: one CTA per source chunk.  Nearly every particle stays in its bin
// or moves to one of the 8 neighbours, so the CTA first ranks its particles per destination class in shared
// memory (warp-aggregated), reserves ONE contiguous range per destination bin with a single global atomic,
// and then writes: every CTA emits a few long sequential runs instead of isolated 256-byte pieces, and the
// global cursor sees ~10 atomics per 4096 particles.  Far movers (rare) take a cursor slot individually.
constexpr int RESORT_THREADS = 256;
static_assert(CHUNK <= 4096, "k_resort_chunks packs a rank < 4096 into 12 bits");

__global__ void __launch_bounds__(RESORT_THREADS, 6)
k_resort_chunks(const double *__restrict__ x, const double *__restrict__ y, const double *__restrict__ vx,
                const double *__restrict__ vy, const uint32_t *__restrict__ id, const Chunk *__restrict__ chunks,
                const int *__restrict__ nchunks, PushConst c, const long long *__restrict__ tile_off,
                unsigned int *__restrict__ cursor, double *__restrict__ x2, double *__restrict__ y2,
                double *__restrict__ vx2, double *__restrict__ vy2, uint32_t *__restrict__ id2) {
    __shared__ unsigned s_cnt[9];
    __shared__ unsigned s_base[9];
    __shared__ unsigned short s_code[CHUNK];     // class (4 bits) | rank inside the CTA and class (12 bits); kept out of registers
    if ((int)blockIdx.x >= *nchunks) return;
    const Chunk ck = chunks[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31;
    const int tx = ck.tile / c.nty, ty = ck.tile - tx * c.nty;
    const bool nbr_ok = (c.ntx >= 3 && c.nty >= 3);
    if (tid < 9) s_cnt[tid] = 0u;
    __syncthreads();

    for (int k0 = 0; k0 < ck.count; k0 += RESORT_THREADS) {       // warp-uniform trip count
        const int k = k0 + tid;
        const bool live = k < ck.count;
        int cls = 15;                   // 15: nothing to do in pass B
        if (live) {
            const long long p = ck.start + k;
            const int t = tile_of(x[p], y[p], c);
            int ddx = t / c.nty - tx, ddy = t % c.nty - ty;
            if (ddx > 1) ddx -= c.ntx; else if (ddx < -1) ddx += c.ntx;
            if (ddy > 1) ddy -= c.nty; else if (ddy < -1) ddy += c.nty;
            if (nbr_ok && ddx >= -1 && ddx <= 1 && ddy >= -1 && ddy <= 1) {
                cls = (ddx + 1) * 3 + (ddy + 1);
            } else {                    // far mover (or a grid with fewer than 3 tiles per side): individual slot
                const long long dst = tile_off[t] + atomicAdd(&cursor[t], 1u);
                x2[dst] = x[p]; y2[dst] = y[p]; vx2[dst] = vx[p]; vy2[dst] = vy[p];
                id2[dst] = id ? id[p] : (uint32_t)p;
            }
        }
        const unsigned peers = __match_any_sync(0xffffffffu, cls);
        unsigned first = 0;
        const int leader = __ffs(peers) - 1;
        if (cls < 9 && lane == leader) first = atomicAdd(&s_cnt[cls], (unsigned)__popc(peers));
        first = __shfl_sync(0xffffffffu, first, leader);
        if (live) s_code[k] = (unsigned short)((unsigned)cls | ((first + (unsigned)__popc(peers & ((1u << lane) - 1u))) << 4));
    }
    __syncthreads();
    if (tid < 9 && s_cnt[tid]) {
        const int ux = (tx + tid / 3 - 1 + c.ntx) % c.ntx, uy = (ty + tid % 3 - 1 + c.nty) % c.nty;
        s_base[tid] = atomicAdd(&cursor[ux * c.nty + uy], s_cnt[tid]);
    }
    __syncthreads();
    for (int k = tid; k < ck.count; k += RESORT_THREADS) {
        const unsigned code = s_code[k];
        const int cls = (int)(code & 15u);
        if (cls < 9) {
            const long long p = ck.start + k;
            const int ux = (tx + cls / 3 - 1 + c.ntx) % c.ntx, uy = (ty + cls % 3 - 1 + c.nty) % c.nty;
            const long long dst = tile_off[ux * c.nty + uy] + s_base[cls] + (code >> 4);
            x2[dst] = x[p]; y2[dst] = y[p]; vx2[dst] = vx[p]; vy2[dst] = vy[p];
            id2[dst] = id ? id[p] : (uint32_t)p;
        }
    }
}

// ---------------------------------------------------------------------------
// Cell order inside a bin (periodic, stand-alone; slow species only).
//
// The mover's shared-memory traffic is dominated by bank conflicts of RANDOM in-tile addresses (E gather 2.7-way,
// accumulator atomics 3.7-way: profiles/r01_v5_rebinning_mover.md).  When the particles of a bin are stored in cell
// order, the lanes of a warp sit in the same cell: the four corner loads broadcast and deposit_commit() combines the
// warp's weights before it touches the accumulators.  Ions move ~0.007 cells per step, so the order survives tens
// of steps; thermal electrons (0.23 cells per step) lose it in two or three, so it is off for them by default.
//
// The bin's range [tile_off[t], tile_off[t+1]) keeps its particles; they are permuted inside it by a counting
// sort on the cell of the bin's WINDOW (stragglers outside the bin are clamped into the ring), in three kernels:
//   k_cell_count   per chunk: population per window cell                       (reads x, y)
//   k_cell_scan    per bin:   first slot of every (chunk, cell) run in the bin (cells major, chunks minor)
//   k_cell_permute per chunk: rank inside (chunk, cell), arrays staged through shared memory in cell order so that
//                  consecutive threads store consecutive slots, written to the second buffer set
// Chunk table, bin offsets and histogram are unchanged.  Any order inside a bin is correct: the step's result is
// bit-identical with and without this pass (tested).
// ---------------------------------------------------------------------------
constexpr int CELLW = TILE + 2 * PICSP_HALO;                 // window edge in cells (24)
constexpr int CELLKEYS = CELLW * CELLW;                      // 576
static_assert(CELLKEYS <= 1024, "cell keys must fit the staging kernel's class counters");

__device__ __forceinline__ int cell_key(double x, double y, const PushConst &c, int tx, int ty) {
    if (!in_box(x, y, c)) return 0;
    const double inv_dx = 1.0 / c.dx;
    double fi, fj;
    const int ci = floor_nonneg(to_logical_fast(x, c.dx, inv_dx), fi);
    const int cj = floor_nonneg(to_logical_fast(y, c.dx, inv_dx), fj);
    const int li = min(max(ci - (tx * TILE - PICSP_HALO), 0), CELLW - 1);
    const int lj = min(max(cj - (ty * TILE - PICSP_HALO), 0), CELLW - 1);
    return li * CELLW + lj;
}

__global__ void __launch_bounds__(256)
k_cell_count(const double *__restrict__ x, const double *__restrict__ y, const Chunk *__restrict__ chunks,
             const int *__restrict__ nchunks, PushConst c, unsigned int *__restrict__ cnt, int *__restrict__ tile_chunk0) {
    __shared__ unsigned s_cnt[CELLKEYS];
    const int b = blockIdx.x;
    if (b >= *nchunks) return;
    const Chunk ck = chunks[b];
    if (threadIdx.x == 0 && (b == 0 || chunks[b - 1].tile != ck.tile)) tile_chunk0[ck.tile] = b;   // chunks are listed bin by bin
    for (int k = threadIdx.x; k < CELLKEYS; k += blockDim.x) s_cnt[k] = 0u;
    __syncthreads();
    const int tx = ck.tile / c.nty, ty = ck.tile - tx * c.nty;
    for (int k = threadIdx.x; k < ck.count; k += blockDim.x)
        atomicAdd(&s_cnt[cell_key(x[ck.start + k], y[ck.start + k], c, tx, ty)], 1u);
    __syncthreads();
    for (int k = threadIdx.x; k < CELLKEYS; k += blockDim.x) cnt[(size_t)b * CELLKEYS + k] = s_cnt[k];
}

// one CTA per bin; thread = cell key.  cnt[chunk][key] -> first slot (relative to the bin) of that run
__global__ void __launch_bounds__(CELLKEYS)
k_cell_scan(const long long *__restrict__ tile_off, const int *__restrict__ tile_chunk0, int chunk, unsigned int *__restrict__ cnt) {
    __shared__ unsigned s_warp[CELLKEYS / 32];
    const int t = blockIdx.x, key = threadIdx.x, lane = key & 31, w = key >> 5;
    const long long pop = tile_off[t + 1] - tile_off[t];
    if (pop <= 0) return;
    const int nch = (int)((pop + chunk - 1) / chunk);
    unsigned int *p = cnt + (size_t)tile_chunk0[t] * CELLKEYS + key;
    unsigned total = 0;
    for (int q = 0; q < nch; q++) total += p[(size_t)q * CELLKEYS];
    // exclusive scan of the per-key totals over the CELLKEYS threads
    unsigned incl = total;
    for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) s_warp[w] = incl;
    __syncthreads();
    if (w == 0) {
        unsigned v = lane < CELLKEYS / 32 ? s_warp[lane] : 0u, inc2 = v;
        for (int o = 1; o < 32; o <<= 1) { const unsigned u = __shfl_up_sync(0xffffffffu, inc2, o); if (lane >= o) inc2 += u; }
        if (lane < CELLKEYS / 32) s_warp[lane] = inc2 - v;
    }
    __syncthreads();
    unsigned run = s_warp[w] + incl - total;
    for (int q = 0; q < nch; q++) { const unsigned v = p[(size_t)q * CELLKEYS]; p[(size_t)q * CELLKEYS] = run; run += v; }
}

constexpr size_t CELLSORT_SMEM_BYTES = sizeof(double) * CHUNK + sizeof(unsigned) * (2 * CHUNK + 2 * CELLKEYS);

__global__ void __launch_bounds__(SORT2_THREADS, 2)
k_cell_permute(const double *__restrict__ x, const double *__restrict__ y, const double *__restrict__ vx,
               const double *__restrict__ vy, const uint32_t *__restrict__ id, const Chunk *__restrict__ chunks,
               const int *__restrict__ nchunks, PushConst c, const long long *__restrict__ tile_off,
               const unsigned int *__restrict__ base, double *__restrict__ x2, double *__restrict__ y2,
               double *__restrict__ vx2, double *__restrict__ vy2, uint32_t *__restrict__ id2) {
    extern __shared__ __align__(16) unsigned char cs_smem[];
    double *s_val = reinterpret_cast<double *>(cs_smem);                    // [CHUNK] one array of the chunk, in cell order
    unsigned *s_dst = reinterpret_cast<unsigned *>(s_val + CHUNK);           // [CHUNK] destination slot (relative to the bin) of cell-ordered element q
    unsigned *s_pos = s_dst + CHUNK;                                         // [CHUNK] key | rank, later the cell-ordered position, of source element k
    unsigned *s_cnt = s_pos + CHUNK;                                         // [CELLKEYS]
    unsigned *s_pref = s_cnt + CELLKEYS;                                     // [CELLKEYS] exclusive prefix of s_cnt
    const int b = blockIdx.x;
    if (b >= *nchunks) return;
    const Chunk ck = chunks[b];
    const int tid = threadIdx.x, lane = tid & 31;
    const int tx = ck.tile / c.nty, ty = ck.tile - tx * c.nty;
    for (int k = tid; k < CELLKEYS; k += SORT2_THREADS) s_cnt[k] = 0u;
    __syncthreads();
    // 1. key and rank (inside chunk and key) of every particle; warp-aggregated counters
    for (int k0 = 0; k0 < ck.count; k0 += SORT2_THREADS) {        // warp-uniform trip count
        const int k = k0 + tid;
        const bool live = k < ck.count;
        const int key = live ? cell_key(x[ck.start + k], y[ck.start + k], c, tx, ty) : -1;
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        const int leader = __ffs(peers) - 1;
        unsigned r0 = 0;
        if (live && lane == leader) r0 = atomicAdd(&s_cnt[key], (unsigned)__popc(peers));
        r0 = __shfl_sync(0xffffffffu, r0, leader);
        if (live) s_pos[k] = ((unsigned)key << 16) | (r0 + (unsigned)__popc(peers & ((1u << lane) - 1u)));   // rank < 4096 < 2^16
    }
    __syncthreads();
    // 2. exclusive prefix of the key populations (warp 0: CELLKEYS / 32 keys per lane)
    if (tid < 32) {
        const int per = (CELLKEYS + 31) / 32, k0 = tid * per, k1 = min(k0 + per, CELLKEYS);
        unsigned sum = 0;
        for (int k = k0; k < k1; k++) sum += s_cnt[k];
        unsigned incl = sum;
        for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        unsigned run = incl - sum;
        for (int k = k0; k < k1; k++) { s_pref[k] = run; run += s_cnt[k]; }
    }
    __syncthreads();
    // 3. cell-ordered position of every source element, destination slot of every position
    const unsigned int *mybase = base + (size_t)b * CELLKEYS;
    for (int k = tid; k < ck.count; k += SORT2_THREADS) {
        const unsigned code = s_pos[k];
        const int key = (int)(code >> 16);
        const unsigned rank = code & 0xFFFFu;
        const unsigned q = s_pref[key] + rank;
        s_pos[k] = q;
        s_dst[q] = mybase[key] + rank;
    }
    __syncthreads();
    // 4. each array: coalesced read -> shared memory in cell order -> runs of consecutive stores
    const long long t0 = tile_off[ck.tile];
    auto move = [&](const double *__restrict__ src, double *__restrict__ dstp) {
        for (int k = tid; k < ck.count; k += SORT2_THREADS) s_val[s_pos[k]] = src[ck.start + k];
        __syncthreads();
        for (int q = tid; q < ck.count; q += SORT2_THREADS) dstp[t0 + s_dst[q]] = s_val[q];
        __syncthreads();
    };
    move(x, x2); move(y, y2); move(vx, vx2); move(vy, vy2);
    unsigned *s_ival = reinterpret_cast<unsigned *>(s_val);
    for (int k = tid; k < ck.count; k += SORT2_THREADS) s_ival[s_pos[k]] = id ? id[ck.start + k] : (uint32_t)(ck.start + k);
    __syncthreads();
    for (int q = tid; q < ck.count; q += SORT2_THREADS) id2[t0 + s_dst[q]] = s_ival[q];
}

// ---------------------------------------------------------------------------
// Bank order inside a chunk (slow species; after every (re-)binning).
//
// The mover's lane l of every warp processes particle (32 q + l) of its chunk.  Its shared-memory accesses are the
// four 16-byte corner loads at window node e = li * pitch + lj (bank group e mod 8, one quarter-warp per wavefront)
// and the eight 4-byte accumulator atomics at words e, e + 1, e + pitch, e + pitch + 1 of two planes (bank e mod 32).
// With particles in arbitrary order those addresses collide like birthdays (2.7 and 3.7 wavefronts per access: 73
// of the mover's 85 wavefronts per 32 particles).  Here the particles of a chunk are permuted, in place, so that
// consecutive particles sit on consecutive values of e mod 32: the chunk is laid out in rows, row r holding the r-th
// particle of every class (class = e mod 32, ascending) that has more than r particles.  While all 32 classes last
// a row is exactly one warp: 32 distinct banks, 8 distinct bank groups per quarter-warp, every access one wavefront —
// without a single extra instruction in the mover.  The shorter rows at the end of the chunk still keep the lanes of
// a warp on mostly distinct banks.
// Measured (profiles/r02b_mover_bank_order.md): the ion launch right after the ordering has 1.74e8 instead of 2.73e8
// shared-memory wavefronts per 1e8 particles and is 7 % faster (0.89 of the measured HBM peak under ncu).  The order
// is fragile, though: a warp instruction costs as many wavefronts as its WORST bank, so a single lane whose particle
// has left its cell doubles it.  Ions move ~0.007 cells per step: the gain is 7 % for the first ~10 steps, ~4 % after
// 30 and ~1 % after 60-80; thermal electrons (0.23 cells per step) lose the order within two steps, so it is off for
// them.  Any order inside a chunk is correct (integer accumulation): results are bit-identical (tested).
// One CTA per chunk; the chunk's four arrays and its slot map pass through shared memory one after the other.
// ---------------------------------------------------------------------------
constexpr int BANK_CLASSES = 32;
constexpr size_t BANKORDER_SMEM_BYTES = sizeof(double) * CHUNK + sizeof(unsigned short) * CHUNK + sizeof(unsigned) * BANK_CLASSES;

__global__ void __launch_bounds__(SORT2_THREADS, 2)
k_bank_order(double *__restrict__ x, double *__restrict__ y, double *__restrict__ vx, double *__restrict__ vy,
             uint32_t *__restrict__ id, const Chunk *__restrict__ chunks, const int *__restrict__ nchunks, PushConst c,
             const int *__restrict__ frac) {
    extern __shared__ __align__(16) unsigned char bo_smem[];
    double *s_val = reinterpret_cast<double *>(bo_smem);                               // [CHUNK] one array of the chunk, in the new order
    unsigned short *s_pos = reinterpret_cast<unsigned short *>(s_val + CHUNK);          // [CHUNK] new position of source element k
    unsigned *s_cnt = reinterpret_cast<unsigned *>(s_pos + CHUNK);                      // [32] class populations
    const int b = blockIdx.x;
    if (b >= *nchunks) return;
    if (frac[1] != 0) return;      // the load is concentrated on few cells and the deposit warp-aggregated: same-cell neighbours are wanted there
    const Chunk ck = chunks[b];
    const int tid = threadIdx.x, lane = tid & 31;
    const int tx = ck.tile / c.nty, ty = ck.tile - tx * c.nty;
    const int wx0 = tx * TILE - HALO, wy0 = ty * TILE - HALO;
    const double inv_dx = 1.0 / c.dx;
    if (tid < BANK_CLASSES) s_cnt[tid] = 0u;
    __syncthreads();
    // 1. class and rank (inside the chunk and class) of every particle; warp-aggregated counters.  The codes stay in
    //    registers: each thread owns CHUNK / SORT2_THREADS = 8 particles.
    constexpr int PER = (CHUNK + SORT2_THREADS - 1) / SORT2_THREADS;
    unsigned code[PER];
#pragma unroll
    for (int u = 0; u < PER; u++) {
        const int k = u * SORT2_THREADS + tid;
        const bool live = k < ck.count;                      // warp-uniform trip count: every lane takes part in the match
        int cls = -1;
        if (live) {
            const double px = x[ck.start + k], py = y[ck.start + k];
            cls = 0;
            if (in_box(px, py, c)) {
                double fi, fj;
                const int ci = floor_nonneg(to_logical_fast(px, c.dx, inv_dx), fi);
                const int cj = floor_nonneg(to_logical_fast(py, c.dx, inv_dx), fj);
                cls = ((ci - wx0) * APITCH + (cj - wy0)) & (BANK_CLASSES - 1);      // stragglers outside the window: any class will do
            }
        }
        const unsigned peers = __match_any_sync(0xffffffffu, cls);
        const int leader = __ffs(peers) - 1;
        unsigned r0 = 0;
        if (live && lane == leader) r0 = atomicAdd(&s_cnt[cls], (unsigned)__popc(peers));
        r0 = __shfl_sync(0xffffffffu, r0, leader);
        code[u] = live ? (((unsigned)cls << 16) | (r0 + (unsigned)__popc(peers & ((1u << lane) - 1u)))) : 0xFFFFFFFFu;
    }
    __syncthreads();
    // 2. position of (class, rank) in the row layout: all earlier rows (sum over classes of min(population, rank)) plus
    //    the classes below mine that reach into my row
#pragma unroll
    for (int u = 0; u < PER; u++) {
        if (code[u] == 0xFFFFFFFFu) continue;
        const unsigned cls = code[u] >> 16, rank = code[u] & 0xFFFFu;
        unsigned q = 0;
#pragma unroll 8
        for (unsigned k = 0; k < BANK_CLASSES; k++) {
            const unsigned n = s_cnt[k];
            q += min(n, rank) + ((k < cls && n > rank) ? 1u : 0u);
        }
        s_pos[u * SORT2_THREADS + tid] = (unsigned short)q;
    }
    __syncthreads();
    // 3. each array: coalesced read -> shared memory in the new order -> coalesced write back to the same range
    auto move = [&](double *__restrict__ a) {
        for (int k = tid; k < ck.count; k += SORT2_THREADS) s_val[s_pos[k]] = a[ck.start + k];
        __syncthreads();
        for (int q = tid; q < ck.count; q += SORT2_THREADS) a[ck.start + q] = s_val[q];
        __syncthreads();
    };
    move(x); move(y); move(vx); move(vy);
    unsigned *s_ival = reinterpret_cast<unsigned *>(s_val);
    for (int k = tid; k < ck.count; k += SORT2_THREADS) s_ival[s_pos[k]] = id[ck.start + k];
    __syncthreads();
    for (int q = tid; q < ck.count; q += SORT2_THREADS) id[ck.start + q] = s_ival[q];
}

// ---------------------------------------------------------------------------
// PTX helpers: mbarrier + TMA tensor load
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 2-D tiled TMA load: box (c0 fastest, c1) of the tensor described by tmap -> dst, completes on bar
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *tmap, int c0, int c1, unsigned long long *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

// 1-D bulk copy global -> shared (16-byte aligned, size a multiple of 16), completes on bar
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct TileCtx {
    int wx0, wy0;        // window origin (node indices, may be negative)
    int ilo, jlo;        // cells whose four nodes lie inside window AND grid: [ilo, ilo+ispan] x [jlo, jlo+jspan]
    unsigned ispan, jspan;
    int tx, ty;          // this chunk's bin
    int ntx1, nty1;      // ntx-1, nty-1
    unsigned long long xl_bits, yl_bits;
};

// 0 <= v < limit for IEEE doubles, as ONE unsigned 64-bit integer compare: non-negative
// doubles order like their bit patterns; negatives (sign bit), NaN and +inf compare high.
// (-0.0 reports "outside"; the callers' slow paths treat it exactly like the reference.)
__device__ __forceinline__ bool in_range_bits(double v, unsigned long long limit_bits) {
    return (unsigned long long)__double_as_longlong(v) < limit_bits;
}
// (Round 2 measured two "instruction diets" here, same box: the test on the high words alone (4 instead of 8 integer
// instructions per test) changed nothing, and a branch of its own for particles that start and end in the un-wrapped
// 3 x 3 block of bins (20 instead of 50 instructions for most electron warps) was 2 % SLOWER with 62 instead of 56
// registers — profiles/r02b_mover_bank_order.md.  Both were removed again.)

// bilinear interpolation of E from the four corners of a cell (gather, src/main.cpp:671-681).  Written with
// explicit roundings so that the window path and the straggler path are the same arithmetic bit for bit whatever the
// compiler would contract: a particle's result must not depend on which bin it happens to be stored in.
__device__ __forceinline__ double2 interp_E(double2 f00, double2 f01, double2 f10, double2 f11, double di, double dj) {
    const double a = 1 - di, b = 1 - dj;
    const double w00 = __dmul_rn(a, b), w10 = __dmul_rn(di, b), w01 = __dmul_rn(a, dj), w11 = __dmul_rn(di, dj);
    double2 e;
    e.x = __fma_rn(f11.x, w11, __fma_rn(f01.x, w01, __fma_rn(f10.x, w10, __dmul_rn(f00.x, w00))));
    e.y = __fma_rn(f11.y, w11, __fma_rn(f01.y, w01, __fma_rn(f10.y, w10, __dmul_rn(f00.y, w00))));
    return e;
}

// E at the position of a particle whose cell is outside its bin's window (it drifted since the last sort, or it
// wrapped through the periodic boundary and is pushed again): same nodes from global memory, same arithmetic as the
// window path.  i/j: the particle's cell, or -1 for a position outside the box (caller-supplied garbage; the mover
// keeps particles inside), which takes the reference's flat-index gather with the zero guard band.
struct Straggler { double ex, ey; int i, j; };     // returned by value: reference outputs would cost a stack frame
#ifndef PICSP_STRAGGLER_ATTR
#define PICSP_STRAGGLER_ATTR __forceinline__   // a real call costs the hot loop 6 registers and 2.5 % (profiles/r01_sweeps.md)
#endif
__device__ PICSP_STRAGGLER_ATTR Straggler gather_straggler(const double2 *__restrict__ E, double px, double py, double dx, double inv_dx,
                                                   double xl, double yl, int nix, int niy, long long nn, long long guard) {
    // scalars, not the PushConst: taking the address of the kernel parameter would give every thread a stack copy
    PushConst c;
    c.dx = dx; c.xl = xl; c.yl = yl; c.nix = nix; c.niy = niy; c.nn = nn; c.guard = guard;
    Straggler r;
    if (!in_box(px, py, c)) {
        const double2 e = gather_E(E, c, to_logical(px, dx), to_logical(py, dx));
        r.ex = e.x; r.ey = e.y; r.i = r.j = -1;
        return r;
    }
    const double lx = to_logical_fast(px, dx, inv_dx), ly = to_logical_fast(py, dx, inv_dx);
    double fi, fj;
    r.i = floor_nonneg(lx, fi); r.j = floor_nonneg(ly, fj);    // -0.0 -> cell 0, as in tile_of()
    const double2 *w = E + ((long long)r.i * niy + r.j);       // 0 <= pos < xl: all four corners are real nodes
    const double2 e = interp_E(w[0], w[1], w[niy], w[niy + 1], lx - fi, ly - fj);
    r.ex = e.x; r.ey = e.y;
    return r;
}

// mover body for one particle (pushSpecies, src/main.cpp:779-846); returns the number of extra pushes.
// oi/oj: cell the push starts from (-1 if outside the box).
__device__ __forceinline__ int push_one(double &px, double &py, double &pvx, double &pvy, const PushConst &c,
                                        double inv_dx, const TileCtx &tc, const double2 *sE,
                                        const double2 *__restrict__ E, int *err, int &oi, int &oj) {
    int extra = 0;
    oi = oj = -1;
    for (int iter = 0;; iter++) {
        double2 e;
        bool done_fast = false;
        if (in_range_bits(px, tc.xl_bits) && in_range_bits(py, tc.yl_bits)) {
            double lx = to_logical_fast(px, c.dx, inv_dx), ly = to_logical_fast(py, c.dx, inv_dx);
            double fi, fj;
            int i = floor_nonneg(lx, fi), j = floor_nonneg(ly, fj);
            if (iter == 0) { oi = i; oj = j; }
            if ((unsigned)(i - tc.ilo) <= tc.ispan && (unsigned)(j - tc.jlo) <= tc.jspan) {
                double di = lx - fi, dj = ly - fj;
                const int li = i - tc.wx0, lj = j - tc.wy0;
                int e0 = li * WPITCH + lj;                       // node offset inside copy 0
                if (REPL > 1) {
                    // lanes of my quarter-warp with my bank-group parity; the r-th of them takes copy r
                    const unsigned act = __activemask();
                    const unsigned lane = threadIdx.x & 31u;
                    const unsigned odd = __ballot_sync(act, e0 & 1);
                    const unsigned mine = ((e0 & 1) ? odd : ~odd) & act & (0xFFu << (lane & 24u));
                    const int rank = __popc(mine & ((1u << lane) - 1u));
                    const int r = (((e0 & 1) + 2 * (rank & 3) - e0) >> 1) & 3;   // bank group (e0 + 2r) mod 8 == parity + 2*rank
                    e0 += r * (E_COPY + 2);
                }
                const double2 *w = sE + e0;
                e = interp_E(w[0], w[1], w[WPITCH], w[WPITCH + 1], di, dj);
                done_fast = true;
            }
        }
        if (!done_fast) { // straggler, wrapped or out-of-box position
            if (c.walls && px != px) break;      // walls extension: an absorbed particle (position NaN) stays where it is
            const Straggler r = gather_straggler(E, px, py, c.dx, inv_dx, c.xl, c.yl, c.nix, c.niy, c.nn, c.guard);
            e.x = r.ex; e.y = r.ey;
            if (iter == 0) { oi = r.i; oj = r.j; }
        }
        pvx += c.dtqm * e.x;
        pvy += c.dtqm * e.y;
        px += c.dt * pvx;
        py += c.dt * pvy;
        // src/main.cpp:807-845: exactly one wrap per push, then push again.  Common case first:
        // both coordinates inside the box means none of the four tests fires.
        if (in_range_bits(px, tc.xl_bits) && in_range_bits(py, tc.yl_bits)) break;
        if (c.walls) {
            // Extension (PICSP_FLAG_WALLS; the reference is periodic-only, its absorbing wall is a commented-out sketch at
            // src/main.cpp:826-843): a particle that leaves the box is absorbed — position NaN, velocity 0: it never
            // deposits, never gathers, never moves again and adds nothing to the kinetic energy.
            if (!in_box(px, py, c)) { px = py = __longlong_as_double(0x7FF8000000000000ll); pvx = pvy = 0.0; extra++; }
            break;
        }
        if (px < 0.0) px += c.xl;
        else if (px >= c.xl) px -= c.xl;
        else if (py < 0.0) py += c.yl;
        else if (py >= c.yl) py -= c.yl;
        else break;                       // -0.0 or NaN: no test fires in the reference either
        extra++;
        if (iter >= 64) { atomicOr(err, ERR_BIT_RUNAWAY); break; }
    }
    return extra;
}

// Exact 64-bit accumulation out of two native 32-bit shared atomics: the low-limb add
// returns the previous value, which tells THIS add whether it wrapped the limb; the wrap is
// then carried into the high limb together with the high half of the addend.  Every carry is
// accounted exactly once, by the add that produced it, so the (hi, lo) pair equals the
// 64-bit sum for any interleaving of adds (64-bit shared atomics are CAS loops on sm_100a).
__device__ __forceinline__ void add64_limbs(unsigned *sLo, unsigned *sHi, int k, unsigned long long w) {
    const unsigned lo = (unsigned)w, hi = (unsigned)(w >> 32);
    const unsigned old = atomicAdd(&sLo[k], lo);
    const unsigned carry = (old + lo) < lo ? 1u : 0u;
    atomicAdd(&sHi[k], hi + carry);
}

// fixed-point CIC deposit of one particle into the window limbs, or the global grid, in two halves:
// deposit_prepare() is per-particle arithmetic, deposit_commit() is warp-cooperative (every live lane of the warp
// must call it, with the same `act` mask).
struct Deposit {
    unsigned long long w00, w10, w01, w11;   // fixed-point corner weights (< 2^52)
    int k;                                   // window node index of the cell's (i, j) corner
    int i, j;                                // cell (global path)
    int mode;                                // 0: nothing (position outside the box), 1: window, 2: global grid
};

// ci/cj return the particle's cell (or -1 when the position is outside the box: skipped).
__device__ __forceinline__ void deposit_prepare(double px, double py, const PushConst &c, double inv_dx, const TileCtx &tc,
                                                double scale, Deposit &dp, int &ci, int &cj) {
    ci = cj = -1;
    dp.mode = 0; dp.k = -1; dp.i = dp.j = 0; dp.w00 = dp.w10 = dp.w01 = dp.w11 = 0ull;
    if (!(in_range_bits(px, tc.xl_bits) && in_range_bits(py, tc.yl_bits))) {
        if (!in_box(px, py, c)) return;          // -0.0 is inside the box; everything else here is not
    }
    double lx = to_logical_fast(px, c.dx, inv_dx), ly = to_logical_fast(py, c.dx, inv_dx);
    double fi, fj;
    int i = floor_nonneg(lx, fi), j = floor_nonneg(ly, fj);
    ci = i; cj = j;
    double di = lx - fi, dj = ly - fj;
    double a = (1 - di) * scale, d = di * scale, b = 1 - dj;
    const double magic = 4503599627370496.0;   // 2^52: the mantissa of (w + 2^52) is w rounded to nearest-even
    const unsigned long long mm = 0xFFFFFFFFFFFFFull;
    dp.w00 = (unsigned long long)__double_as_longlong(fma(a, b, magic)) & mm;
    dp.w10 = (unsigned long long)__double_as_longlong(fma(d, b, magic)) & mm;
    dp.w01 = (unsigned long long)__double_as_longlong(fma(a, dj, magic)) & mm;
    dp.w11 = (unsigned long long)__double_as_longlong(fma(d, dj, magic)) & mm;
    if ((unsigned)(i - tc.ilo) <= tc.ispan && (unsigned)(j - tc.jlo) <= tc.jspan) {
        dp.k = (i - tc.wx0) * APITCH + (j - tc.wy0);
        dp.mode = 1;
        return;
    }
    if (i > c.nix - 2 || j > c.niy - 2) {
        // A position within an ulp below xl (yl) rounds to the cell index ncx (ncy).  The reference then deposits the
        // whole weight on the last node row with di == 0 (src/main.cpp:657-667), which the fold adds to row 0; that is
        // the corner pair (ncx-1, di = 1) of the last real cell.
        if (i > c.nix - 2) { i = c.nix - 2; di = 1.0; }
        if (j > c.niy - 2) { j = c.niy - 2; dj = 1.0; }
        a = (1 - di) * scale; d = di * scale; b = 1 - dj;
        dp.w00 = (unsigned long long)__double_as_longlong(fma(a, b, magic)) & mm;
        dp.w10 = (unsigned long long)__double_as_longlong(fma(d, b, magic)) & mm;
        dp.w01 = (unsigned long long)__double_as_longlong(fma(a, dj, magic)) & mm;
        dp.w11 = (unsigned long long)__double_as_longlong(fma(d, dj, magic)) & mm;
    }
    dp.i = i; dp.j = j; dp.mode = 2;
}

// Exact sum of a 52-bit quantity over the lanes of `m` out of two 32-bit warp reductions (REDUX.SUM): the 26-bit
// halves each sum to < 32 * 2^26 = 2^31.
__device__ __forceinline__ unsigned long long warp_sum52(unsigned m, unsigned long long w) {
    const unsigned lo = (unsigned)w & 0x3FFFFFFu, hi = (unsigned)(w >> 26);
    return (unsigned long long)__reduce_add_sync(m, lo) + ((unsigned long long)__reduce_add_sync(m, hi) << 26);
}

#ifndef PICSP_AGG_MIN
#define PICSP_AGG_MIN 4        // lanes of a warp in one cell from which their deposits are combined before touching shared memory
#endif
#ifndef PICSP_AGG_ROUNDS
#define PICSP_AGG_ROUNDS 2     // cells per warp that get combined (0: never); lanes of further cells deposit individually
#endif

// Warp-aggregated commit.  Lanes of a warp whose particles sit in the SAME cell would hit the same eight
// shared-memory words and serialise (a cell-ordered store, or the reference's own diagonal two-stream load with
// ~5e5 particles per occupied cell).  Up to two groups per warp are combined: the lanes that share the
// cell of lane 0 (or, when lane 0 is a stray, of the first lane outside its group), then the largest-looking group
// among the rest (the two interleaved beams of the diagonal load); the fixed-point weights of a group are summed
// exactly with REDUX and its first lane alone touches the accumulators; every other lane deposits for itself.
// One SHFL + one VOTE per group instead of a MATCH (round-2 ncu: MATCH + POPC + VOTE and its loop made the
// ordered-store kernel issue-bound).  Integer sums: the result is bit-identical to the lane-by-lane deposit.
// Returns true when the deposit stayed inside the window.
__device__ __forceinline__ bool deposit_commit(const Deposit &dp, unsigned act, bool aggregate, const PushConst &c, unsigned *sLo,
                                               unsigned *sHi, long long *__restrict__ acc) {
    if (!(PICSP_AGG_ROUNDS > 0 && aggregate && act == 0xffffffffu)) {
        // the plain commit (every lane for itself), kept as a block of its own so that the common case carries none of
        // the aggregation's register traffic.  Full warps only aggregate; the tail slice of a chunk comes here too.
        if (dp.mode == 1) {
            int k = dp.k;
            if (REPL > 1) {
                const unsigned peers = __match_any_sync(__activemask(), k & 7);
                const int rank = __popc(peers & ((1u << (threadIdx.x & 31u)) - 1u));
                const int r = (((k & 7) + 8 * (rank & 3) - k) >> 3) & 3;
                k += r * ACC_COPY;
            }
            add64_limbs(sLo, sHi, k, dp.w00);
            add64_limbs(sLo, sHi, k + APITCH, dp.w10);
            add64_limbs(sLo, sHi, k + 1, dp.w01);
            add64_limbs(sLo, sHi, k + APITCH + 1, dp.w11);
            return true;
        }
        if (dp.mode == 2) {
            unsigned long long *g = reinterpret_cast<unsigned long long *>(acc) + ((long long)dp.i * c.niy + dp.j);
            atomicAdd(g, dp.w00); atomicAdd(g + c.niy, dp.w10); atomicAdd(g + 1, dp.w01); atomicAdd(g + c.niy + 1, dp.w11);
        }
        return false;
    }
    unsigned long long v00 = dp.w00, v10 = dp.w10, v01 = dp.w01, v11 = dp.w11;
    bool own = dp.mode == 1;
    {
        const unsigned lane = threadIdx.x & 31u;
        const int key = dp.mode == 1 ? dp.k : -1 - (int)lane;                 // lanes without a window deposit match nobody
        unsigned done = 0u;
#pragma unroll
        for (int r = 0; r < PICSP_AGG_ROUNDS; r++) {
            int ld = __ffs(~done) - 1;                                         // first lane not yet in a group (r == 0: lane 0)
            unsigned m = __ballot_sync(0xffffffffu, key == __shfl_sync(0xffffffffu, key, ld)) & ~done;
            if (r == 0 && __popc(m) < 8 && m != 0xffffffffu) {                // lane 0 is a stray of an ordered warp: try its first non-member
                const int l1 = __ffs(~m) - 1;
                const unsigned m1 = __ballot_sync(0xffffffffu, key == __shfl_sync(0xffffffffu, key, l1));
                if (__popc(m1) > __popc(m)) { m = m1; ld = l1; }
            }
            if (__popc(m) >= PICSP_AGG_MIN) {                                  // warp-uniform
                if ((m >> lane) & 1u) {
                    v00 = warp_sum52(m, dp.w00); v10 = warp_sum52(m, dp.w10);
                    v01 = warp_sum52(m, dp.w01); v11 = warp_sum52(m, dp.w11);
                    own = (int)lane == ld;
                }
            }
            done |= m;
            if (done == 0xffffffffu) break;
        }
    }
    if (own) {
        int k = dp.k;
        if (REPL > 1) {
            // lanes whose node index is congruent to mine mod 8 share my banks; the r-th of them takes copy r
            const unsigned peers = __match_any_sync(__activemask(), k & 7);
            const int rank = __popc(peers & ((1u << (threadIdx.x & 31u)) - 1u));
            const int r = (((k & 7) + 8 * (rank & 3) - k) >> 3) & 3;            // bank (k + 8r) mod 32 == class + 8*rank
            k += r * ACC_COPY;
        }
        add64_limbs(sLo, sHi, k, v00);
        add64_limbs(sLo, sHi, k + APITCH, v10);
        add64_limbs(sLo, sHi, k + 1, v01);
        add64_limbs(sLo, sHi, k + APITCH + 1, v11);
    } else if (dp.mode == 2) {
        unsigned long long *g = reinterpret_cast<unsigned long long *>(acc) + ((long long)dp.i * c.niy + dp.j);
        atomicAdd(g, dp.w00); atomicAdd(g + c.niy, dp.w10); atomicAdd(g + 1, dp.w01); atomicAdd(g + c.niy + 1, dp.w11);
    }
    return dp.mode == 1;
}

// ---------------------------------------------------------------------------
// the fused chunk kernel.  MODE 0: push + deposit(next step); 1: deposit only
// (standalone scatterSpecies); 2: push only (PICSP_FLAG_NO_FUSE);
// MODE 3 / 4: MODE 0 that also RE-BINS: instead of writing the pushed particle back in place it writes it into
// the new binned layout (destination = bin of the position the push STARTED from, whose histogram is known
// before the launch), so a periodic re-sort costs no extra pass over the particles.  The store is then
// "binned as of one step ago", which the window halo absorbs like any other drift.
//   MODE 4 (the normal case): the previous launch left, per chunk, how many of its particles ended in each of the
//   9 neighbouring bins (chunk_cnt) — exactly the destination populations of this launch — and k_rebin_bases has
//   turned them into one reserved range per (chunk, bin).  Slots are handed out from shared-memory counters: no
//   global atomic, no extra barrier in the loop.
//   MODE 3 (no counts available, e.g. re-sort on consecutive steps): one reservation per slice and bin; every
//   slice then waits for a global atomic round trip (measured: the launch takes twice as long as MODE 0).
// counters[0] = extra pushes, counters[1] = particles that deposited outside their window
// ---------------------------------------------------------------------------
struct RebinArgs {
    const uint32_t *id;            // current slot -> upload index (a binned store always has one)
    const long long *tile_off;     // offsets of the NEW layout
    unsigned int *cursor;          // per-bin fill cursors of the NEW layout
    double *x2, *y2, *vx2, *vy2;   // destination arrays
    uint32_t *id2;
    const unsigned int *chunk_base;// MODE 4: [chunk][9] first slot (relative to the bin) reserved for this chunk
    unsigned int *chunk_cnt;       // MODE 0: [chunk][9] particles of this chunk that ended in each neighbouring bin (or nullptr)
};

// one thread per (chunk, neighbour class): reserve the chunk's range in the destination bin
__global__ void k_rebin_bases(const Chunk *__restrict__ chunks, const int *__restrict__ nchunks, int ntx, int nty,
                              const unsigned int *__restrict__ chunk_cnt, const long long *__restrict__ tile_off,
                              unsigned int *__restrict__ cursor, unsigned int *__restrict__ chunk_base, int *__restrict__ err) {
    const long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (q >= 9ll * *nchunks) return;
    const int ch = (int)(q / 9), cls = (int)(q - 9ll * ch);
    const unsigned cnt = chunk_cnt[q];
    if (!cnt) { chunk_base[q] = 0u; return; }
    const int t = chunks[ch].tile;
    const int tx = t / nty, ty = t - tx * nty;
    const int ux = (tx + cls / 3 - 1 + ntx) % ntx, uy = (ty + cls % 3 - 1 + nty) % nty;
    const int ut = ux * nty + uy;
    const unsigned base = atomicAdd(&cursor[ut], cnt);
    // bin sizes and chunk counts come from the same launch and the same bin function: this never fires
    if ((long long)base + cnt > tile_off[ut + 1] - tile_off[ut]) atomicOr(err, ERR_BIT_REBIN);
    chunk_base[q] = base;
}

template <int MODE>
__global__ void __launch_bounds__(MOVER_THREADS, MOVER_MIN_CTAS)
k_tile_mover(const __grid_constant__ CUtensorMap tmapE, double *__restrict__ x, double *__restrict__ y,
             double *__restrict__ vx, double *__restrict__ vy, const Chunk *__restrict__ chunks,
             const int *__restrict__ nchunks, PushConst c, const double2 *__restrict__ E,
             long long *__restrict__ acc, const int *__restrict__ frac, unsigned int *__restrict__ hist_next,
             unsigned long long *__restrict__ counters, int *__restrict__ err, RebinArgs rb) {
    constexpr bool REBIN = (MODE == 3 || MODE == 4);     // built on the bulk-copy pipeline; never launched without it (op_push)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2 *sE = reinterpret_cast<double2 *>(smem_raw);                               // REPL copies of E_COPY nodes
    unsigned *sLo = reinterpret_cast<unsigned *>(smem_raw + sizeof(double2) * (size_t)REPL * E_COPY);
    unsigned *sHi = sLo + (size_t)REPL * ACC_COPY;
    __shared__ unsigned sCnt[9];
    __shared__ unsigned sRbCnt[9], sRbBase[9];     // MODE 3: per-slice population / reserved base of each neighbour bin; MODE 4: running count / chunk base
    __shared__ long long sDstOff[9];               // re-binning: first destination slot per neighbour class
    __shared__ __align__(8) unsigned long long sBar;
    __shared__ __align__(8) unsigned long long sFull[STAGES];

    if ((int)blockIdx.x >= *nchunks) return;
    const Chunk ck = chunks[blockIdx.x];
    const int tid = threadIdx.x;
    TileCtx tc;
    tc.tx = ck.tile / c.nty; tc.ty = ck.tile - tc.tx * c.nty;
    tc.wx0 = tc.tx * TILE - HALO; tc.wy0 = tc.ty * TILE - HALO;
    tc.ilo = max(tc.wx0, 0); tc.jlo = max(tc.wy0, 0);
    tc.ispan = (unsigned)(min(tc.wx0 + WIN - 2, c.nix - 2) - tc.ilo);
    tc.jspan = (unsigned)(min(tc.wy0 + WIN - 2, c.niy - 2) - tc.jlo);
    tc.ntx1 = c.ntx - 1; tc.nty1 = c.nty - 1;
    tc.xl_bits = (unsigned long long)__double_as_longlong(c.xl); tc.yl_bits = (unsigned long long)__double_as_longlong(c.yl);

    if (tid == 0) {
        mbar_init(&sBar, 1);
        for (int st = 0; st < STAGES; st++) mbar_init(&sFull[st], 1);
        fence_barrier_init();
    }
    if (MODE != 1 && tid == 0) {
        mbar_expect_tx(&sBar, (uint32_t)(REPL * WIN * WPITCH * sizeof(double2)));
        // tensor = E viewed as [nix][2*niy] doubles; box = [WIN][2*WPITCH]; out-of-grid elements arrive as 0.
        // copy r is loaded 2r nodes to the left, so node (li, lj) sits at column lj + 2r: two bank groups further
#pragma unroll
        for (int r = 0; r < REPL; r++) tma_load_2d(sE + (size_t)r * E_COPY, &tmapE, 2 * (tc.wy0 - 2 * r), tc.wx0, &sBar);
    }
    if (MODE != 2)
        for (int k = tid; k < REPL * ACC_COPY; k += MOVER_THREADS) { sLo[k] = 0u; sHi[k] = 0u; }
    if (tid < 9) { sCnt[tid] = 0u; sRbCnt[tid] = 0u; }
    if (REBIN && tid < 9) {        // first slot of each neighbouring bin in the new layout (MODE 4: of this chunk's range in it)
        const int ux = (tc.tx + tid / 3 - 1 + c.ntx) % c.ntx, uy = (tc.ty + tid % 3 - 1 + c.nty) % c.nty;
        long long off = rb.tile_off[ux * c.nty + uy];
        if (MODE == 4) off += rb.chunk_base[9ll * blockIdx.x + tid];
        sDstOff[tid] = off;
    }

    // chunk-local pointers: 32-bit indexing inside the loop
    double *__restrict__ cx = x + ck.start;
    double *__restrict__ cy = y + ck.start;
    double *__restrict__ cvx = vx + ck.start;
    double *__restrict__ cvy = vy + ck.start;
    const int count = ck.count;
    const bool nbr_ok = (c.ntx >= 3 && c.nty >= 3);
    const double inv_dx = 1.0 / c.dx;
    const double scale = (MODE != 2) ? exp2((double)frac[0]) : 0.0;
    const bool aggregate = (MODE != 2) && frac[1] != 0;      // CTA-uniform: decided by k_frac_from_hist
    unsigned extra = 0, outside = 0, same = 0;

    // everything that happens to particle k of the chunk
    // act: the lanes of this warp that process a particle in this iteration (warp-cooperative deposit), 0 = unknown
    auto process = [&](int k, double &px, double &py, double &pvx, double &pvy, int &oi, int &oj, unsigned act) {
        oi = oj = -1;
        if (MODE != 1) {
            extra += push_one(px, py, pvx, pvy, c, inv_dx, tc, sE, E, err, oi, oj);
            if (!REBIN) { cx[k] = px; cy[k] = py; cvx[k] = pvx; cvy[k] = pvy; }
        }
        int ci = -1, cj = -1;
        if (MODE != 2) {
            Deposit dp;
            deposit_prepare(px, py, c, inv_dx, tc, scale, dp, ci, cj);
            if (!deposit_commit(dp, act, aggregate, c, sLo, sHi, acc)) outside++;
        } else if (in_box(px, py, c)) {
            double fi, fj;
            ci = floor_nonneg(to_logical_fast(px, c.dx, inv_dx), fi);
            cj = floor_nonneg(to_logical_fast(py, c.dx, inv_dx), fj);
        }
        if (MODE != 1) {
            // histogram of the positions just written (bins of the next sort, bound of the next
            // scale); same bin function as tile_of(): out-of-box positions count in bin 0
            // common case in ONE test: the push started and ended in this chunk's own bin (cell - first cell of the
            // bin < TILE for all four coordinates; -1 = outside the box fails it).  ci < ncx, so ci / TILE never
            // exceeds ntx - 1 and the clamp of tile_of_cell() is the identity for it.
            const int cx0 = tc.tx * TILE, cy0 = tc.ty * TILE;
            if (((unsigned)(ci - cx0) | (unsigned)(cj - cy0) | (unsigned)(oi - cx0) | (unsigned)(oj - cy0)) < (unsigned)TILE) {
                same++;
            } else {
                const int ux = ci >= 0 ? min((int)((unsigned)ci / TILE), tc.ntx1) : 0;
                const int uy = ci >= 0 ? min((int)((unsigned)cj / TILE), tc.nty1) : 0;
                if (ux == tc.tx && uy == tc.ty) {
                    same++;
                } else {
                    int ddx = ux - tc.tx, ddy = uy - tc.ty;
                    if (ddx > 1) ddx -= c.ntx; else if (ddx < -1) ddx += c.ntx;      // periodic neighbours
                    if (ddy > 1) ddy -= c.nty; else if (ddy < -1) ddy += c.nty;
                    if (nbr_ok && ddx >= -1 && ddx <= 1 && ddy >= -1 && ddy <= 1)
                        atomicAdd(&sCnt[(ddx + 1) * 3 + (ddy + 1)], 1u);
                    else
                        atomicAdd(&hist_next[ux * c.nty + uy], 1u);
                }
                // the fixed-point scale bounds the particles that move at most one tile in ONE step; the others are counted
                const int ox = oi >= 0 ? min((int)((unsigned)oi / TILE), tc.ntx1) : 0;
                const int oy = oi >= 0 ? min((int)((unsigned)oj / TILE), tc.nty1) : 0;
                if ((ox != ux || oy != uy) && !(c.walls && ci < 0)) {        // an absorbed particle is filed under bin 0: not a displacement
                    int ax = abs(ux - ox), ay = abs(uy - oy);
                    ax = min(ax, c.ntx - ax); ay = min(ay, c.nty - ay);
                    if (MODE != 2 && (ax > 1 || ay > 1)) far_mover(counters, frac, c.far_shift, err);   // (MODE 2 deposits later, from the histogram of the positions it finds)
                }
            }
        }
    };

    if (BULK_PIPE) {
        double *sP = reinterpret_cast<double *>(smem_raw + MOVER_WINDOW_BYTES);      // [STAGES][4][STAGE_W]
        uint32_t *sI = reinterpret_cast<uint32_t *>(smem_raw + MOVER_WINDOW_BYTES + MOVER_PIPE_BYTES);   // [STAGES][STAGE_IDW]
        constexpr int NARR = (MODE == 1) ? 2 : 4;
        const int niter = (count + MOVER_THREADS - 1) / MOVER_THREADS;
        // slice `it` -> stage it % STAGES.  Bulk copies need 16-byte alignment: the slice is widened to even
        // particle indices (the extra neighbours are never read).
        auto issue = [&](int it) {
            const int st = it % STAGES;
            const long long first = ck.start + (long long)it * MOVER_THREADS;
            const int cnt = min(MOVER_THREADS, count - it * MOVER_THREADS);
            const long long a0 = first & ~1ll, a1 = (first + cnt + 1) & ~1ll;
            const uint32_t bytes = (uint32_t)((a1 - a0) * sizeof(double));
            double *dst = sP + (size_t)st * 4 * STAGE_W;
            // a re-binning launch also moves the slot -> upload-index map: it rides in the same pipeline stage
            const long long i0 = first & ~3ll, i1 = (first + cnt + 3) & ~3ll;
            const uint32_t ibytes = REBIN ? (uint32_t)((i1 - i0) * sizeof(uint32_t)) : 0u;
            mbar_expect_tx(&sFull[st], NARR * bytes + ibytes);
            if (REBIN) bulk_load(sI + (size_t)st * STAGE_IDW, rb.id + i0, ibytes, &sFull[st]);
            bulk_load(dst, x + a0, bytes, &sFull[st]);
            bulk_load(dst + STAGE_W, y + a0, bytes, &sFull[st]);
            if (MODE != 1) {
                bulk_load(dst + 2 * STAGE_W, vx + a0, bytes, &sFull[st]);
                bulk_load(dst + 3 * STAGE_W, vy + a0, bytes, &sFull[st]);
            }
        };
        if (tid == 0)
            for (int it = 0; it < STAGES - 1 && it < niter; it++) issue(it);
        __syncthreads();                       // accumulators zeroed, barriers initialised
        if (MODE != 1) mbar_wait(&sBar, 0);    // E window landed
        for (int it = 0; it < niter; it++) {
            const int st = it % STAGES;
            // the stage refilled here was drained in iteration it-1 (barrier at the end of that iteration)
            if (tid == 0 && it + STAGES - 1 < niter) issue(it + STAGES - 1);
            mbar_wait(&sFull[st], (uint32_t)((it / STAGES) & 1));
            const int k = it * MOVER_THREADS + tid;
            const bool live = k < count;
            const unsigned act = __ballot_sync(0xffffffffu, live);
            double px = 0, py = 0, pvx = 0, pvy = 0;
            int oi = -1, oj = -1;
            if (live) {
                const double *src = sP + (size_t)st * 4 * STAGE_W + (int)((ck.start + (long long)it * MOVER_THREADS) & 1ll) + tid;
                px = src[0]; py = src[STAGE_W];
                if (MODE != 1) { pvx = src[2 * STAGE_W]; pvy = src[3 * STAGE_W]; }
                process(k, px, py, pvx, pvy, oi, oj, act);
            }
            if (REBIN) {
                // destination bin = bin of the position the push started from (same bin function as the histogram)
                // (no runtime integer division in the per-particle code: TILE is a power of two)
                int cls = 15, tpx = 0, tpy = 0;
                if (live) {
                    if (oi >= 0) { tpx = min((int)((unsigned)oi / TILE), tc.ntx1); tpy = min((int)((unsigned)oj / TILE), tc.nty1); }
                    int ddx = tpx - tc.tx, ddy = tpy - tc.ty;
                    // mirrors the counting of the positions written by the previous launch (below): own bin, neighbour, far
                    if ((ddx | ddy) == 0) cls = 4;
                    else {
                        if (ddx > 1) ddx -= c.ntx; else if (ddx < -1) ddx += c.ntx;
                        if (ddy > 1) ddy -= c.nty; else if (ddy < -1) ddy += c.nty;
                        cls = (nbr_ok && ddx >= -1 && ddx <= 1 && ddy >= -1 && ddy <= 1) ? (ddx + 1) * 3 + (ddy + 1) : 9;
                    }
                }
                // rank inside (chunk or slice, destination bin): warp-aggregated shared-memory counter
                const unsigned lane = tid & 31u;
                const unsigned peers = __match_any_sync(0xffffffffu, cls);
                const int leader = __ffs(peers) - 1;
                unsigned first = 0;
                if (cls < 9 && (int)lane == leader) first = atomicAdd(&sRbCnt[cls], (unsigned)__popc(peers));
                first = __shfl_sync(0xffffffffu, first, leader);
                const unsigned rank = first + (unsigned)__popc(peers & ((1u << lane) - 1u));
                if (MODE == 3) {
                    __syncthreads();
                    if (tid < 9 && sRbCnt[tid]) {      // one contiguous reservation per destination bin and slice
                        const int ux = (tc.tx + tid / 3 - 1 + c.ntx) % c.ntx, uy = (tc.ty + tid % 3 - 1 + c.nty) % c.nty;
                        const int ut = ux * c.nty + uy;
                        const unsigned base = atomicAdd(&rb.cursor[ut], sRbCnt[tid]);
                        // the bin sizes come from the histogram taken by the previous launch: same bin function, so this never fires
                        if ((long long)base + sRbCnt[tid] > rb.tile_off[ut + 1] - rb.tile_off[ut]) atomicOr(err, ERR_BIT_REBIN);
                        sRbBase[tid] = base;
                        sRbCnt[tid] = 0u;
                    }
                    __syncthreads();
                }
                if (live) {
                    long long dst;
                    if (cls < 9) {
                        dst = sDstOff[cls] + rank;
                        if (MODE == 3) dst += sRbBase[cls];
                    } else {                           // far bin (straggler of stragglers): individual slot
                        const int tpre = tpx * c.nty + tpy;
                        dst = rb.tile_off[tpre] + atomicAdd(&rb.cursor[tpre], 1u);
                    }
                    rb.x2[dst] = px; rb.y2[dst] = py; rb.vx2[dst] = pvx; rb.vy2[dst] = pvy;
                    rb.id2[dst] = sI[(size_t)st * STAGE_IDW + (int)((ck.start + (long long)it * MOVER_THREADS) & 3ll) + tid];
                }
            }
            __syncthreads();
        }
    } else {
        // per-thread register prefetch: the next particle's loads are in flight while the current one is processed
        if (REBIN) __trap();               // the host only selects a re-binning launch when BULK_PIPE is on
        const int last = count - 1;
        int k = tid;
        double px, py, pvx = 0, pvy = 0;
        {
            const int k0 = min(k, last);
            px = cx[k0]; py = cy[k0];
            if (MODE != 1) { pvx = cvx[k0]; pvy = cvy[k0]; }
        }
        __syncthreads();
        if (MODE != 1) mbar_wait(&sBar, 0);
        while (k < count) {
            const int kn = k + MOVER_THREADS;
            double nx = px, ny = py, nvx = pvx, nvy = pvy;
            if (kn < count) {
                nx = cx[kn]; ny = cy[kn];
                if (MODE != 1) { nvx = cvx[kn]; nvy = cvy[kn]; }
            }
            int oi, oj;
            process(k, px, py, pvx, pvy, oi, oj, 0u);       // per-thread trip counts: no warp-cooperative deposit here
            k = kn; px = nx; py = ny; pvx = nvx; pvy = nvy;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        extra += __shfl_xor_sync(0xffffffffu, extra, o);
        outside += __shfl_xor_sync(0xffffffffu, outside, o);
        same += __shfl_xor_sync(0xffffffffu, same, o);
    }
    if ((tid & 31) == 0) {
        if (same) atomicAdd(&sCnt[4], same);
        if (extra) atomicAdd(&counters[0], (unsigned long long)extra);
        if (outside) atomicAdd(&counters[1], (unsigned long long)outside);
    }
    __syncthreads();

    // flush the window: limbs -> one native 64-bit integer RED per touched node
    if (MODE != 2) {
        for (int q = tid; q < WIN * APITCH; q += MOVER_THREADS) {      // (the padding column holds zeros)
            unsigned long long v = 0;
#pragma unroll
            for (int r = 0; r < REPL; r++)
                v += ((unsigned long long)sHi[q + r * ACC_COPY] << 32) | (unsigned long long)sLo[q + r * ACC_COPY];
            if (v) {
                int li = q / APITCH, lj = q - li * APITCH;
                long long gi = tc.wx0 + li, gj = tc.wy0 + lj;
                atomicAdd(reinterpret_cast<unsigned long long *>(acc) + (gi * c.niy + gj), v);
            }
        }
    }
    // where this chunk's particles are now, by neighbouring bin: the reservation sizes of a re-binning launch next step
    if (MODE == 0 && rb.chunk_cnt && tid < 9) rb.chunk_cnt[9ll * blockIdx.x + tid] = sCnt[tid];
    if (MODE != 1 && tid < 9 && sCnt[tid]) {
        if (tid == 4) {
            atomicAdd(&hist_next[ck.tile], sCnt[4]);
        } else {
            int ux = (tc.tx + tid / 3 - 1 + c.ntx) % c.ntx, uy = (tc.ty + tid % 3 - 1 + c.nty) % c.nty;
            atomicAdd(&hist_next[ux * c.nty + uy], sCnt[tid]);
        }
    }
}

}  // namespace picsp
