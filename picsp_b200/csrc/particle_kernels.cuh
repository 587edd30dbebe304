// picsp_b200/csrc/particle_kernels.cuh — particle-side kernels.
//
// Particle state is structure-of-arrays (x[], y[], vx[], vy[], f64) in HBM.
// Deposition accumulates 64-bit FIXED-POINT weights with integer atomics: integer
// addition is associative, so the density is bit-for-bit independent of particle
// order, launch order and atomic arrival order (deterministic within a GPU), which a
// floating-point atomic cannot give.  The number of fraction bits is chosen on the
// device from a per-tile particle histogram so that no node can overflow 2^62 while
// keeping the quantum far below double rounding of the node value.
#pragma once
#include "ctx.cuh"

namespace picsp {

struct PushConst {
    double dx;        // == dy
    double dt;
    double dtqm;      // timeStep*qm            (src/main.cpp:793: evaluated as (timeStep*qm)*E)
    double hdtqm;     // 0.5*timeStep*qm        (src/main.cpp:863: ((0.5*timeStep)*qm)*E)
    double xl, yl;    // domain.xl == xmax, domain.yl == ymax (x0 = y0 = 0)
    int nix, niy;
    int ntx, nty;
    long long nn, guard;
    int walls;        // PICSP_FLAG_WALLS (extension, no reference semantics): a particle that leaves the box is absorbed
    int far_shift;    // log2 (rounded up) of the stores that deposit into this species' accumulator grid (parts): see far_mover()
};

// XtoL / YtoL, src/main.cpp:643-652: true division, x0 = y0 = 0.
__device__ __forceinline__ double to_logical(double pos, double dx) { return (pos - 0.0) / dx; }

// ---------------------------------------------------------------------------
// gather, src/main.cpp:671-681.  (int) casts truncate toward zero; indexing is
// flat, i*niy+j, exactly as the reference, so a j == niy-1 cell reads the next row
// (Q5).  Reads outside the array — undefined behaviour in the reference — return 0:
// E carries a zeroed guard band and anything beyond it is clamped to 0 as well.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double2 load_E(const double2 *__restrict__ E, long long idx, long long nn, long long guard) {
    if (idx < -guard || idx >= nn + guard) return make_double2(0.0, 0.0);
    return E[idx];
}

__device__ __forceinline__ double2 gather_E(const double2 *__restrict__ E, const PushConst &c, double lx, double ly) {
    int i = __double2int_rz(lx);
    int j = __double2int_rz(ly);
    double di = lx - i, dj = ly - j;
    long long b = (long long)i * c.niy + j;
    double2 f00, f10, f01, f11;
    if (i >= 0 && i <= c.nix - 2 && j >= 0 && j <= c.niy - 2) {   // all four corners are real nodes
        f00 = E[b]; f10 = E[b + c.niy]; f01 = E[b + 1]; f11 = E[b + c.niy + 1];
    } else {
        f00 = load_E(E, b, c.nn, c.guard); f10 = load_E(E, b + c.niy, c.nn, c.guard);
        f01 = load_E(E, b + 1, c.nn, c.guard); f11 = load_E(E, b + c.niy + 1, c.nn, c.guard);
    }
    double2 e;
    e.x = f00.x * (1 - di) * (1 - dj) + f10.x * di * (1 - dj) + f01.x * (1 - di) * dj + f11.x * di * dj;
    e.y = f00.y * (1 - di) * (1 - dj) + f10.y * di * (1 - dj) + f01.y * (1 - di) * dj + f11.y * di * dj;
    return e;
}

// x/dx without the division sequence: one multiply by the rounded reciprocal and one
// Newton correction (error <= 1 ulp of the quotient; CIC weights are continuous across a
// cell boundary, so an index that differs by one at an exact boundary gives the same sums).
__device__ __forceinline__ double to_logical_fast(double pos, double dx, double inv_dx) {
    double q = pos * inv_dx;
    double r = fma(-q, dx, pos);
    return fma(r, inv_dx, q);
}

// floor of a non-negative double < 2^31 without a conversion instruction:
// l + 2^52 rounded toward zero leaves floor(l) in the low mantissa bits.
__device__ __forceinline__ int floor_nonneg(double l, double &fl) {
    const double magic = 4503599627370496.0;   // 2^52
    double t = __dadd_rz(l, magic);
    fl = t - magic;
    return __double2loint(t);
}

__device__ __forceinline__ bool in_box(double px, double py, const PushConst &c) {
    return px >= 0.0 && px < c.xl && py >= 0.0 && py < c.yl;
}

// Bin of a position.  ONE definition shared by the histogram, the sort and the mover, so
// that bin populations and bin contents always agree bit for bit.  Positions outside the
// box (only possible for caller-supplied garbage; the mover keeps particles inside) go to bin 0.
__device__ __forceinline__ int tile_of_cell(int ci, int cj, const PushConst &c) {
    return min(ci / TILE, c.ntx - 1) * c.nty + min(cj / TILE, c.nty - 1);
}
__device__ __forceinline__ int tile_of(double x, double y, const PushConst &c) {
    if (!in_box(x, y, c)) return 0;
    const double inv_dx = 1.0 / c.dx;
    double fi, fj;
    int ci = floor_nonneg(to_logical_fast(x, c.dx, inv_dx), fi);
    int cj = floor_nonneg(to_logical_fast(y, c.dx, inv_dx), fj);
    return tile_of_cell(ci, cj, c);
}

// ---------------------------------------------------------------------------
// fixed-point CIC scatter of one particle (weights of src/main.cpp:664-667 without
// the common factor spwt/(dx*dy), which k_deposit_finalize applies once per node)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void scatter_fixed(long long *__restrict__ acc, const PushConst &c, double x, double y,
                                              double scale) {
    double lx = to_logical(x, c.dx), ly = to_logical(y, c.dx);
    int i = __double2int_rz(lx), j = __double2int_rz(ly);
    double di = lx - i, dj = ly - j;
    if (i < 0 || j < 0 || i > c.nix - 1 || j > c.niy - 1) return;   // cannot happen for 0 <= pos < xl (main.cpp:807-824 guarantees it)
    // pos within an ulp below xl: the quotient rounds to ncx and the reference puts the whole weight on the last
    // node row (di == 0), i.e. on the (ncx-1, di = 1) corner pair of the last real cell
    if (i > c.nix - 2) { i = c.nix - 2; di = 1.0; }
    if (j > c.niy - 2) { j = c.niy - 2; dj = 1.0; }
    long long b = (long long)i * c.niy + j;
    unsigned long long *a = reinterpret_cast<unsigned long long *>(acc);
    atomicAdd(&a[b],             (unsigned long long)__double2ll_rn((1 - di) * (1 - dj) * scale));
    atomicAdd(&a[b + c.niy],     (unsigned long long)__double2ll_rn(di * (1 - dj) * scale));
    atomicAdd(&a[b + 1],         (unsigned long long)__double2ll_rn((1 - di) * dj * scale));
    atomicAdd(&a[b + c.niy + 1], (unsigned long long)__double2ll_rn(di * dj * scale));
}

// ---------------------------------------------------------------------------
// per-tile particle histogram and the fixed-point scale derived from it
// ---------------------------------------------------------------------------
// Random positions hit a few thousand bins: global atomics on so few addresses serialise in L2 (12.6 ms for 5e8
// particles, the traffic alone is 1.5 ms), so each CTA counts in shared memory and flushes its non-zero bins.
// `nsmem` = bins that fit the CTA's dynamic shared memory (0: count in global memory directly).
__global__ void k_tile_hist(const double *__restrict__ x, const double *__restrict__ y, long long n,
                            PushConst c, unsigned int *__restrict__ hist, int nsmem) {
    extern __shared__ unsigned int s_hist[];
    const int nt = c.ntx * c.nty;
    const bool local = nsmem >= nt;
    if (local) {
        for (int t = threadIdx.x; t < nt; t += blockDim.x) s_hist[t] = 0u;
        __syncthreads();
    }
    // contiguous slice per CTA (see k_sort_scatter)
    const long long per_cta = (n + gridDim.x - 1) / gridDim.x;
    const long long lo = blockIdx.x * per_cta, hi = min(n, lo + per_cta);
    for (long long p = lo + threadIdx.x; p < hi; p += blockDim.x)
        atomicAdd(local ? &s_hist[tile_of(x[p], y[p], c)] : &hist[tile_of(x[p], y[p], c)], 1u);
    if (local) {
        __syncthreads();
        for (int t = threadIdx.x; t < nt; t += blockDim.x)
            if (s_hist[t]) atomicAdd(&hist[t], s_hist[t]);
    }
}

// frac = 62 - bits(max over tiles of the periodic 5x5 neighbourhood population).
// A node receives weight only from particles in the <= 2x2 tiles around it; the
// fused mover deposits positions one step after the histogram was taken, and a
// particle may move by at most one tile per step (checked by the mover), so the
// 5x5 neighbourhood bounds every node sum by pop * 2^frac < 2^62.
// Multi-CTA: every thread sums one tile's neighbourhood (25 independent L2 loads), the CTA maximum goes
// to scratch[0] with atomicMax, and the last CTA to finish (ticket in scratch[1]) converts the maximum into
// the fraction-bit count and resets the scratch for the next call.
// sum of the parts' histograms (a species split into several stores deposits into ONE accumulator grid)
__global__ void k_hist_add(unsigned int *__restrict__ sum, const unsigned int *__restrict__ h, int nt) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nt) sum[t] += h[t];
}

__global__ void __launch_bounds__(256)
k_frac_from_hist(const unsigned int *__restrict__ hist, int ntx, int nty, int *__restrict__ frac, int cap,
                 unsigned long long *__restrict__ scratch, long long n, int force_agg) {
    __shared__ unsigned long long s_max[8], s_tmax[8];
    unsigned long long m = 0, tm = 0;          // largest 5x5 neighbourhood, largest single tile
    const int wx = ntx < 5 ? ntx : 5, wy = nty < 5 ? nty : 5;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < ntx * nty) {
        const int tx = t / nty, ty = t % nty;
        tm = hist[t];
        for (int a = 0; a < wx; a++)
            for (int b = 0; b < wy; b++) {
                int ux = (tx - wx / 2 + a + 2 * ntx) % ntx, uy = (ty - wy / 2 + b + 2 * nty) % nty;
                m += hist[ux * nty + uy];
            }
    }
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long v = __shfl_xor_sync(0xffffffffu, m, o); m = v > m ? v : m;
        v = __shfl_xor_sync(0xffffffffu, tm, o); tm = v > tm ? v : tm;
    }
    if ((threadIdx.x & 31) == 0) { s_max[threadIdx.x >> 5] = m; s_tmax[threadIdx.x >> 5] = tm; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 0; w < (blockDim.x + 31) / 32; w++) { m = s_max[w] > m ? s_max[w] : m; tm = s_tmax[w] > tm ? s_tmax[w] : tm; }
        atomicMax(&scratch[0], m);
        atomicMax(&scratch[2], tm);
        __threadfence();
        if (atomicAdd(&scratch[1], 1ull) == (unsigned long long)gridDim.x - 1) {   // last CTA
            __threadfence();
            const unsigned long long mx = atomicExch(&scratch[0], 0ull);
            const unsigned long long tmx = atomicExch(&scratch[2], 0ull);
            scratch[1] = 0ull;
            const int bits = 64 - __clzll((long long)(mx | 1ull));
            const int f = 62 - bits;
            frac[0] = f > cap ? cap : (f < 0 ? 0 : f);
            // frac[1]: the mover combines the deposits of the lanes of a warp that share a cell before they touch shared
            // memory (deposit_commit).  Worth its instructions only when whole warps share cells: a cell-ordered store
            // (force_agg), or a load concentrated on few bins — the busiest bin holds more than 8x its fair share,
            // e.g. the reference's diagonal two-stream load (src/main.cpp:597-615): 64 of 4096 bins occupied.
            const long long fair = n / ((long long)ntx * nty) + 1;
            frac[1] = (force_agg > 0 || (force_agg == 0 && (long long)tmx > 8 * fair + 4096)) ? 1 : 0;      // force_agg: 1 on, 0 automatic, -1 off
        }
    }
}

// ---------------------------------------------------------------------------
// standalone deposit (scatter loop of scatterSpecies, src/main.cpp:695-700)
// ---------------------------------------------------------------------------
__global__ void k_deposit(const double *__restrict__ x, const double *__restrict__ y, long long n, PushConst c,
                          long long *__restrict__ acc, const int *__restrict__ frac) {
    const double scale = exp2((double)*frac);
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x)
        scatter_fixed(acc, c, x[p], y[p], scale);
}

// ---------------------------------------------------------------------------
// mover: pushSpecies + gather (src/main.cpp:772-847, :671-681), optionally fused
// with the NEXT step's scatter (the positions written here are exactly the ones
// scatterSpecies reads next, src/main.cpp:482-483), which makes one particle-step
// cost 32 B read + 32 B written.
//
// The reference's wrap is an if/else-if chain whose trailing `else it++` is the
// only place the list iterator advances (:807-845): a particle that wrapped is
// pushed AGAIN from its wrapped position, until a push ends inside the box (Q4).
// ---------------------------------------------------------------------------
constexpr int ERR_BIT_DISPLACEMENT = 1;

// A particle that moves MORE THAN ONE particle tile in a step ("far mover") is legal in the reference and is handled
// here too (global gather, global deposit, individual slot at a re-binning); but the fixed-point scale of the fused
// deposit (2^frac with 2^frac * P < 2^62, P = the largest 5 x 5-tile population before the push) only bounds what the
// particles of a node's OWN neighbourhood can add to it.  Far movers may land anywhere, so they are counted
// (counters[2]): while fewer than 2^(62 - frac) of them exist in one launch — 2048 for the sparsest load, millions
// for a dense one — even all of them on one node keep its sum below 2^63.  Only beyond that does the launch raise
// PICSP_ERR_DISPLACEMENT (an absurd population: a plasma that far out of its CFL range has long stopped meaning anything).
__device__ __forceinline__ void far_mover(unsigned long long *__restrict__ counters, const int *__restrict__ frac, int far_shift,
                                          int *__restrict__ err) {
    const unsigned long long seen = atomicAdd(&counters[2], 1ull) + 1ull;
    const int room = 62 - frac[0] - far_shift;
    if (room < 0 || seen >= (1ull << room)) atomicOr(err, ERR_BIT_DISPLACEMENT);
}
constexpr int ERR_BIT_RUNAWAY = 2;
constexpr int ERR_BIT_REBIN = 4;

template <bool FUSE_DEPOSIT>
__global__ void __launch_bounds__(256)
k_push(double *__restrict__ x, double *__restrict__ y, double *__restrict__ vx, double *__restrict__ vy,
       long long n, PushConst c, const double2 *__restrict__ E, long long *__restrict__ acc,
       const int *__restrict__ frac, unsigned int *__restrict__ hist_next,
       unsigned long long *__restrict__ counters, int *__restrict__ err) {
    const double scale = FUSE_DEPOSIT ? exp2((double)*frac) : 0.0;
    unsigned int extra = 0;
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
        double px = x[p], py = y[p], pvx = vx[p], pvy = vy[p];
        const int t0 = tile_of(px, py, c);
        int wrapped, guard_iter = 0;
        do {
            double lx = to_logical(px, c.dx), ly = to_logical(py, c.dx);
            double2 e = gather_E(E, c, lx, ly);
            pvx += c.dtqm * e.x;
            pvy += c.dtqm * e.y;
            px += c.dt * pvx;
            py += c.dt * pvy;
            wrapped = 1;
            if (px < 0.0) px += c.xl;
            else if (px >= c.xl) px -= c.xl;
            else if (py < 0.0) py += c.yl;
            else if (py >= c.yl) py -= c.yl;
            else wrapped = 0;
            extra += wrapped;
            if (++guard_iter > 64) { atomicOr(err, ERR_BIT_RUNAWAY); break; }   // +-inf / absurd speeds: the reference would spin forever
        } while (wrapped);
        x[p] = px; y[p] = py; vx[p] = pvx; vy[p] = pvy;
        if (FUSE_DEPOSIT) {
            const int t1 = tile_of(px, py, c);
            int dtx = abs(t1 / c.nty - t0 / c.nty), dty = abs(t1 % c.nty - t0 % c.nty);
            dtx = min(dtx, c.ntx - dtx); dty = min(dty, c.nty - dty);
            if (dtx > 1 || dty > 1) far_mover(counters, frac, c.far_shift, err);
            atomicAdd(&hist_next[t1], 1u);
            scatter_fixed(acc, c, px, py, scale);
        }
    }
    for (int o = 16; o > 0; o >>= 1) extra += __shfl_xor_sync(0xffffffffu, extra, o);
    if ((threadIdx.x & 31) == 0 && extra) atomicAdd(&counters[0], (unsigned long long)extra);
}

// rewindSpecies, src/main.cpp:850-866 (no move, no wrap)
__global__ void k_rewind(const double *__restrict__ x, const double *__restrict__ y, double *__restrict__ vx,
                         double *__restrict__ vy, long long n, PushConst c, const double2 *__restrict__ E) {
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
        double lx = to_logical(x[p], c.dx), ly = to_logical(y[p], c.dx);
        double2 e = gather_E(E, c, lx, ly);
        vx[p] -= c.hdtqm * e.x;
        vy[p] -= c.hdtqm * e.y;
    }
}

// ---------------------------------------------------------------------------
// bench-only synthetic loader (NOT reference behaviour): counter-based RNG
// ---------------------------------------------------------------------------
__device__ __forceinline__ double u01(uint64_t seed, uint64_t idx, uint32_t draw) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (idx * 8ull + draw + 1ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

__global__ void k_fill_synthetic(double *__restrict__ x, double *__restrict__ y, double *__restrict__ vx,
                                 double *__restrict__ vy, long long n, long long first, uint64_t seed,
                                 double xl, double yl, double vth, double xdrift) {
    const double s2 = 1.4142135623730951;
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
        uint64_t g = (uint64_t)(first + p);
        x[p] = u01(seed, g, 0) * xl;
        y[p] = u01(seed, g, 4) * yl;
        double sgn = (g & 1ull) ? -1.0 : 1.0;
        vx[p] = vth * s2 * (u01(seed, g, 1) + u01(seed, g, 2) + u01(seed, g, 3) - 1.5) + sgn * xdrift;
        vy[p] = vth * s2 * (u01(seed, g, 5) + u01(seed, g, 6) + u01(seed, g, 7) - 1.5);
    }
}

// un-permute for downloads: out[id[slot]] = in[slot]
__global__ void k_unpermute(const double *__restrict__ in, const uint32_t *__restrict__ id, double *__restrict__ out,
                            long long n) {
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x)
        out[id[p]] = in[p];
}

// row-layout dump: rows[id[slot]] = {x, y, vx, vy}[slot] (id == nullptr: identity); 32 contiguous bytes per particle
__global__ void k_rows_unpermute(const double *__restrict__ x, const double *__restrict__ y, const double *__restrict__ vx,
                                 const double *__restrict__ vy, const uint32_t *__restrict__ id, long long n,
                                 double2 *__restrict__ rows) {
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
        const long long r = id ? (long long)id[p] : p;
        rows[2 * r] = make_double2(x[p], y[p]);
        rows[2 * r + 1] = make_double2(vx[p], vy[p]);
    }
}

}  // namespace picsp
