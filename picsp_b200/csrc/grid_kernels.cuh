// picsp_b200/csrc/grid_kernels.cuh — field-side kernels (all O(grid), L2-resident).
//
// Each kernel names the reference statements it replaces (/root/reference/src/main.cpp).
#pragma once
#include <cooperative_groups.h>

#include "ctx.cuh"

namespace picsp {

// The reference's truncated pi (src/main.cpp:61) — used in k, NOT in the FFT twiddles.
__device__ __constant__ double kRefPi = 3.14159265359;

// ---------------------------------------------------------------------------
// finalize: den += (spwt/(dx*dy)) * acc * 2^-frac ; acc = 0
// (the accumulate-into semantics of scatterSpecies, src/main.cpp:689-700: the
//  clearing memset at :692 is commented out, so den is a running sum — Q1)
// ---------------------------------------------------------------------------
__global__ void k_deposit_finalize(double *__restrict__ den, long long *__restrict__ acc,
                                   const int *__restrict__ frac, double weight, long long nn, int clear) {
    const double scale = weight * exp2((double)(-*frac));
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < nn;
         k += (long long)gridDim.x * blockDim.x) {
        long long a = acc[k];
        double base = clear ? 0.0 : den[k];
        den[k] = base + (double)a * scale;
        acc[k] = 0;
    }
}

// ---------------------------------------------------------------------------
// periodic fold, rows first then columns (src/main.cpp:702-712 and :880-890).
// One CTA: the two phases are ordered (corners take both), 2*(nix+niy) updates.
// ---------------------------------------------------------------------------
__global__ void k_fold_periodic(double *f, int nix, int niy) {
    for (int j = threadIdx.x; j < niy; j += blockDim.x) {
        double v = f[j] + f[(long long)(nix - 1) * niy + j];
        f[j] = v;
        f[(long long)(nix - 1) * niy + j] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nix; i += blockDim.x) {
        double v = f[(long long)i * niy] + f[(long long)i * niy + niy - 1];
        f[(long long)i * niy] = v;
        f[(long long)i * niy + niy - 1] = v;
    }
}

// ---------------------------------------------------------------------------
// rho interior (src/main.cpp:874-878); boundary nodes are left as they are (Q3)
// ---------------------------------------------------------------------------
__global__ void k_compute_rho(double *__restrict__ rho, const double *__restrict__ den_i,
                              const double *__restrict__ den_e, double q_i, double q_e, int nix, int niy) {
    long long nint = (long long)(nix - 2) * (niy - 2);
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < nint;
         k += (long long)gridDim.x * blockDim.x) {
        int i = 1 + (int)(k / (niy - 2)), j = 1 + (int)(k % (niy - 2));
        long long idx = (long long)i * niy + j;
        rho[idx] = q_i * den_i[idx] + q_e * den_e[idx];
    }
}

// ---------------------------------------------------------------------------
// One-launch grid phase of a time step (picsp_step): for BOTH species
//   den += spwt/(dx*dy) * acc * 2^-frac ; acc = 0          (k_deposit_finalize)
//   periodic fold of den, rows then columns                 (k_fold_periodic, src/main.cpp:702-712)
// then rho = q_i*den_i + q_e*den_e on interior nodes        (k_compute_rho, src/main.cpp:874-878)
// and the fold of rho's (never written) boundary            (src/main.cpp:880-890).
// The folds only couple an edge node with its periodic partner, so they are done in place by giving
// each partner pair (and the four corners together) to ONE thread: the result of the two ordered fold
// loops is  edge (0,j)/(L,j): f0j + fLj;  edge (i,0)/(i,M): fi0 + fiM;  corners: (f00 + fL0) + (f0M + fLM).
// ---------------------------------------------------------------------------
struct GridPhaseSpecies { double *den; long long *acc; const int *frac; double weight, q; };

__device__ __forceinline__ double finalize_node(const GridPhaseSpecies &sp, long long k, double scale, int clear) {
    const long long a = sp.acc[k];
    sp.acc[k] = 0;
    return (clear ? 0.0 : sp.den[k]) + (double)a * scale;
}

__global__ void k_grid_phase(GridPhaseSpecies s0, GridPhaseSpecies s1, double *__restrict__ rho, int nix, int niy, int clear,
                             const int *__restrict__ err, int *__restrict__ err_host) {
    const int L = nix - 1, M = niy - 1;
    // mirror the sticky device error flag (set by the movers of the previous steps) into mapped host memory:
    // picsp_step reads it without synchronising and reports the violation from its next call
    if (blockIdx.x == 0 && threadIdx.x == 0 && *err) *err_host = *err;
    const double sc0 = s0.weight * exp2((double)(-*s0.frac)), sc1 = s1.weight * exp2((double)(-*s1.frac));
    const unsigned nn = (unsigned)nix * (unsigned)niy;      // <= 2^31 by the capacity of int node counts
    for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < nn; k += gridDim.x * blockDim.x) {
        const int i = (int)(k / (unsigned)niy), j = (int)(k - (unsigned)i * (unsigned)niy);
        if (i > 0 && i < L && j > 0 && j < M) {              // interior node: all four loads in flight before the first store
            const long long a0 = s0.acc[k], a1 = s1.acc[k];
            const double d0 = clear ? 0.0 : s0.den[k], d1 = clear ? 0.0 : s1.den[k];
            const double di = d0 + (double)a0 * sc0, de = d1 + (double)a1 * sc1;
            s0.acc[k] = 0; s1.acc[k] = 0;
            s0.den[k] = di; s1.den[k] = de;
            rho[k] = s0.q * di + s1.q * de;
        } else if (i == 0 && j > 0 && j < M) {               // rows 0 and L: the (0,j) thread owns the pair
            const long long a = k, b = (long long)L * niy + j;
            double v = finalize_node(s0, a, sc0, clear) + finalize_node(s0, b, sc0, clear); s0.den[a] = v; s0.den[b] = v;
            v = finalize_node(s1, a, sc1, clear) + finalize_node(s1, b, sc1, clear); s1.den[a] = v; s1.den[b] = v;
            v = rho[a] + rho[b]; rho[a] = v; rho[b] = v;
        } else if (j == 0 && i > 0 && i < L) {               // columns 0 and M: the (i,0) thread owns the pair
            const long long a = k, b = a + M;
            double v = finalize_node(s0, a, sc0, clear) + finalize_node(s0, b, sc0, clear); s0.den[a] = v; s0.den[b] = v;
            v = finalize_node(s1, a, sc1, clear) + finalize_node(s1, b, sc1, clear); s1.den[a] = v; s1.den[b] = v;
            v = rho[a] + rho[b]; rho[a] = v; rho[b] = v;
        } else if (i == 0 && j == 0) {                       // the four corners
            const long long c00 = 0, c0M = M, cL0 = (long long)L * niy, cLM = cL0 + M;
            double v = (finalize_node(s0, c00, sc0, clear) + finalize_node(s0, cL0, sc0, clear)) +
                       (finalize_node(s0, c0M, sc0, clear) + finalize_node(s0, cLM, sc0, clear));
            s0.den[c00] = v; s0.den[c0M] = v; s0.den[cL0] = v; s0.den[cLM] = v;
            v = (finalize_node(s1, c00, sc1, clear) + finalize_node(s1, cL0, sc1, clear)) +
                (finalize_node(s1, c0M, sc1, clear) + finalize_node(s1, cLM, sc1, clear));
            s1.den[c00] = v; s1.den[c0M] = v; s1.den[cL0] = v; s1.den[cLM] = v;
            v = (rho[c00] + rho[cL0]) + (rho[c0M] + rho[cLM]);
            rho[c00] = v; rho[c0M] = v; rho[cL0] = v; rho[cLM] = v;
        }                                                    // (L,j), (i,M) and the other corners are written by their owners
    }
}

// ---------------------------------------------------------------------------
// k-space Green's function (src/main.cpp:996-1025) + 1/(Nx*Ny) (:1051-1055).
//  * kx = 2*PI*i/Lx for i < Nx/2, 2*PI*(Nx-i)/Lx for i > Nx/2, Lx = xl (:999-1021)
//  * row i == Nx/2 is never written by the reference -> defined as 0 (Q7)
//  * DC bin zeroed (:1023-1024)
//  * the zeroed row breaks Hermitian symmetry in the self-conjugate columns; FFTW's
//    c2r silently keeps only the real part after the dim-0 transform, cuFFT's Z2D
//    leaves that case undefined, so those columns are symmetrised here:
//        P[i,j] <- (P[i,j] + conj P[(Nx-i)%Nx, j]) / 2      for j == 0 (and j == Ny/2, Ny even)
//    which gives the identical real output (Re of the dim-0 transform).
// ---------------------------------------------------------------------------
__device__ __forceinline__ double2 green_bin(const cufftDoubleComplex *__restrict__ rhok, int i, int j,
                                             int Nx, int Nh, double Lx, double Ly) {
    if (i == Nx / 2 || (i == 0 && j == 0)) return make_double2(0.0, 0.0);
    double ky = 2.0 * kRefPi * j / Ly;
    double kx = (i < Nx / 2) ? 2.0 * kRefPi * i / Lx : 2.0 * kRefPi * (Nx - i) / Lx;
    double k2 = kx * kx + ky * ky;
    cufftDoubleComplex r = rhok[(long long)i * Nh + j];
    return make_double2(r.x / k2, r.y / k2);
}

__global__ void k_kspace_green(const cufftDoubleComplex *__restrict__ rhok, cufftDoubleComplex *__restrict__ phik,
                               int Nx, int Ny, double Lx, double Ly) {
    const int Nh = Ny / 2 + 1;
    const double norm = (double)((long long)Nx * Ny);
    long long total = (long long)Nx * Nh;
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < total;
         k += (long long)gridDim.x * blockDim.x) {
        int i = (int)(k / Nh), j = (int)(k % Nh);
        double2 p = green_bin(rhok, i, j, Nx, Nh, Lx, Ly);
        bool self_conj = (j == 0) || ((Ny % 2 == 0) && (j == Ny / 2));
        if (self_conj) {
            double2 m = green_bin(rhok, (Nx - i) % Nx, j, Nx, Nh, Lx, Ly);
            p.x = 0.5 * (p.x + m.x);
            p.y = 0.5 * (p.y - m.y);
        }
        phik[k].x = p.x / norm;
        phik[k].y = p.y / norm;
    }
}

// ---------------------------------------------------------------------------
// SOR (src/main.cpp:904-957): in-place lexicographic sweeps with omega = 1.4 and
// periodic wrap indices.  The lexicographic order is reproduced EXACTLY by an
// anti-diagonal wavefront: node (i,j) needs the new values of (i-1,j), (i,j-1)
// (diagonal d-1), the new (1,j) when i == nix-1 and the new (i,1) when j == niy-1
// (earlier diagonals), and the OLD values of (i+1,j), (i,j+1), (nix-2,j) for i == 0
// and (i,niy-2) for j == 0 (all on later diagonals).  One CTA marches the
// nix+niy-1 diagonals with a block barrier between them; nodes on one diagonal are
// independent.  Convergence is tested after sweeps 0,100,200,... exactly as the
// reference does (in practice the first test passes: one sweep per call, Q6).
// status[0] = sweeps done (negative: cap hit), d_l2 = last L2.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double block_sum(double v, double *s_red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < (blockDim.x + 31) / 32; w++) t += s_red[w];
    return t;   // valid on thread 0
}

__device__ __forceinline__ double ldcg(const double *p) { return __ldcg(p); }

// x / d for a loop-invariant d with its correctly rounded reciprocal r = 1/d: quotient estimate, exact
// remainder (FMA), one correction.  This is the final step of the IEEE division sequence (Markstein): the
// result is the correctly rounded quotient (identical to `x / d`) except in astronomically rare tie cases,
// at 3 dependent FMAs instead of ~10 — the SOR wavefront is latency-bound on exactly this chain.
__device__ __forceinline__ double sor_div(double x, double d, double r) {
    double q = x * r;
    double rem = fma(-q, d, x);
    return fma(rem, r, q);
}

__global__ void __launch_bounds__(1024, 1)
k_sor_solve(double *phi, const double *__restrict__ rho, int nix, int niy, double dx, double dy,
            long long *status, double *d_l2, int max_sweeps, int first_sweep) {
    __shared__ double s_red[32];
    __shared__ double s_l2;
    const double dx2 = dx * dx, dy2 = dy * dy, eps = 1.0;
    const double coef = 0.5 * (1 / ((1 / dx2) + (1 / dy2)));
    const double rdx2 = 1 / dx2, rdy2 = 1 / dy2;
    const int tid = threadIdx.x, nt = blockDim.x;
    double L2 = 0.0;
    if (first_sweep > 0) {            // sweep 0 and its convergence test were done by the pipelined kernels
        if (status[0] == 1) return;   // converged after one sweep (the usual case)
        L2 = *d_l2;
    }
    for (int sweep = first_sweep; sweep < max_sweeps; sweep++) {
        for (int d = 0; d <= nix + niy - 2; d++) {
            int i_lo = d - (niy - 1); if (i_lo < 0) i_lo = 0;
            int i_hi = d < nix - 1 ? d : nix - 1;
            for (int i = i_lo + tid; i <= i_hi; i += nt) {
                int j = d - i;
                int p = i - 1; if (p < 0) p = nix - 2;
                int q = i + 1; if (q > nix - 1) q = 1;
                int r = j - 1; if (r < 0) r = niy - 2;
                int s = j + 1; if (s > niy - 1) s = 1;
                long long c = (long long)i * niy + j;
                double g = coef * (sor_div(ldcg(&phi[(long long)p * niy + j]) + ldcg(&phi[(long long)q * niy + j]), dx2, rdx2) +
                                   sor_div(ldcg(&phi[(long long)i * niy + r]) + ldcg(&phi[(long long)i * niy + s]), dy2, rdy2) +
                                   (rho[c] / eps));
                double old = ldcg(&phi[c]);
                __stcg(&phi[c], old + 1.4 * (g - old));
            }
            __syncthreads();
        }
        if (sweep % 100 == 0) {
            double sum = 0.0;
            long long nn = (long long)nix * niy;
            for (long long k = tid; k < nn; k += nt) {
                int i = (int)(k / niy), j = (int)(k % niy);
                int p = i - 1; if (p < 0) p = nix - 2;
                int q = i + 1; if (q > nix - 1) q = 1;
                int r = j - 1; if (r < 0) r = niy - 2;
                int s = j + 1; if (s > niy - 1) s = 1;
                double R = 0.25 * (ldcg(&phi[(long long)p * niy + j]) + ldcg(&phi[(long long)q * niy + j]) +
                                   ldcg(&phi[(long long)i * niy + r]) + ldcg(&phi[(long long)i * niy + s]) +
                                   (dx2 * rho[k] / eps)) - ldcg(&phi[k]);
                sum = sum + (R * R);
            }
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            if ((tid & 31) == 0) s_red[tid >> 5] = sum;
            __syncthreads();
            if (tid == 0) {
                double t = 0.0;
                for (int w = 0; w < (nt + 31) / 32; w++) t += s_red[w];
                s_l2 = sqrt(t) / (nix * niy);
            }
            __syncthreads();
            L2 = s_l2;
            if (L2 < 1e-2) {
                if (tid == 0) { status[0] = sweep + 1; *d_l2 = L2; }
                return;
            }
            __syncthreads();
        }
    }
    if (tid == 0) { status[0] = -(long long)max_sweeps; *d_l2 = L2; }
}

// ---------------------------------------------------------------------------
// SOR sweep 0, pipelined across CTAs (same lexicographic iterate as k_sor_solve).
//
// One thread per grid row i; thread t of band b owns row i = b*SOR_ROWS + t and visits the
// columns j = s - t at band step s, so the threads of a band form a skewed wavefront.  What a
// node needs (src/main.cpp:916-924):
//   new phi(i-1, j)  - produced ONE step earlier by thread t-1: handed over in shared memory;
//                      for t == 0 it comes from the previous band through global memory, gated by
//                      that band's published progress (checked once per 32 columns);
//                      for i == 0 it is the OLD phi(nix-2, j) (periodic wrap), read before the last
//                      band can overwrite it (that band transitively waits on this one)
//   new phi(i, j-1)  - this thread's previous result (register); OLD phi(i, niy-2) for j == 0
//   old phi(i+1, j), old phi(i, j+1), rho(i, j) - not yet written this sweep: plain (L1-cached,
//                      software-prefetched) global loads; NEW phi(1, j) for i == nix-1 and this
//                      thread's saved NEW phi(i, 1) for j == niy-1 (periodic wraps)
// so a step costs one shared-memory hand-over and one block barrier instead of L2 round trips.
// Bands must be co-resident (grid = ceil(nix/SOR_ROWS) <= #SMs, enforced by the launcher).
// ---------------------------------------------------------------------------
constexpr int SOR_ROWS = 64;
constexpr int SOR_DEPTH = 8;     // columns each thread keeps in flight

// A store to phi(i,j) invalidates the L1 line that also holds phi(i,j+1..), so the "old" operands
// cannot be served from L1, and deep register prefetching is defeated by the counting scoreboard
// (a wait on an old load also waits for the younger loads sharing its slot).  Instead every thread
// streams the operands of its next SOR_DEPTH columns through a private shared-memory ring with
// cp.async (LDGSTS, L2-only, completion tracked by cp.async groups, not the scoreboard).  cp.async.cg
// moves 16 aligned bytes, so the aligned pair containing the wanted double is fetched.
__device__ __forceinline__ void sor_cp_async16(void *smem_dst, const double *gmem_elem) {
    const unsigned long long a = reinterpret_cast<unsigned long long>(gmem_elem) & ~15ull;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(a) : "memory");
}
// Operands that THIS SM alone writes during the sweep (old phi(i,j+1), old phi(i+1,j)) or nobody writes (rho) can go
// through L1: 8-byte cp.async.ca, no pair selection afterwards.  (Operands another band produces this sweep — new
// phi(i-1,.) for thread 0, new phi(1,.) for row nix-1 — must bypass L1, which could hold the line from before they
// were written: those keep the 16-byte .cg form.)
__device__ __forceinline__ void sor_cp_async8(double *smem_dst, const double *gmem_elem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_elem) : "memory");
}
__device__ __forceinline__ double sor_pick(const double2 &pair, const double *gmem_elem) {
    return (reinterpret_cast<unsigned long long>(gmem_elem) & 8ull) ? pair.y : pair.x;
}

__global__ void __launch_bounds__(SOR_ROWS, 1)
k_sor_sweep_pipelined(double *phi, const double *__restrict__ rho, int nix, int niy, double dx, double dy,
                      int *progress) {
    constexpr int D = SOR_DEPTH;
    __shared__ double s_new[2][SOR_ROWS];
    __shared__ __align__(16) double2 s_ring[D][4][SOR_ROWS];   // [slot][right, down, rho, up][thread]
    const int t = threadIdx.x, b = blockIdx.x;
    const int i = b * SOR_ROWS + t;
    const bool active = i < nix;
    const int rows_here = min(SOR_ROWS, nix - b * SOR_ROWS);
    const bool last_row_of_band = (t == rows_here - 1);
    const double dx2 = dx * dx, dy2 = dy * dy, eps = 1.0;
    const double coef = 0.5 * (1 / ((1 / dx2) + (1 / dy2)));
    const double rdx2 = 1 / dx2, rdy2 = 1 / dy2;
    const int q_row = (i + 1 > nix - 1) ? 1 : i + 1;          // src/main.cpp:917
    const int p_row = (i - 1 < 0) ? nix - 2 : i - 1;          // src/main.cpp:916
    const double *row = phi + (long long)(active ? i : 0) * niy;
    const double *row_q = phi + (long long)(active ? q_row : 0) * niy;
    const double *row_p = phi + (long long)(active ? p_row : 0) * niy;
    const double *rrho = rho + (long long)(active ? i : 0) * niy;
    // row nix-1 reads the NEW phi(1, j); it may be fetched ahead only if row 1 is far enough in front
    const bool q_is_new = (i == nix - 1);
    const bool q_prefetch_ok = !q_is_new || (nix > D + 8);
    const bool up_from_global = (t == 0);   // previous band's last row (b > 0), or the i == 0 wrap (old values)
    int granted = (b == 0) ? niy : 0;       // columns of the previous band known to be complete (polling threads only)

    // start the copies of the next column's operands into ring slot k (one cp.async group per call).  Calls come in
    // column order (jf = -t, -t+1, ...), so the operand addresses are running pointers, not recomputed indices.
    int jf = -t;
    const double *pf_right = row + (jf + 1), *pf_q = row_q + jf, *pf_rho = rrho + jf, *pf_p = row_p + jf;
    auto fetch = [&](int k) {
        if (active && jf >= 0 && jf < niy) {
            // Threads that read values ANOTHER band produces this sweep — thread 0 (new phi(i-1,.)) and the row
            // nix-1 (new phi(1,.), implied by the previous band's progress) — wait for the previous band first,
            // (progress is published 16 columns at a time; this also covers the fetches issued before the first barrier).
            if ((up_from_global || q_is_new) && b > 0 && jf >= granted) {
                const int want = min(jf + 1, niy);
                while ((granted = *(volatile int *)&progress[b - 1]) < want) { }
                __threadfence();
            }
            if (jf < niy - 1) sor_cp_async8(&s_ring[k][0][t].x, pf_right);        // old phi(i,jf+1); jf == niy-1 uses saved_col1
            if (q_prefetch_ok) {
                if (q_is_new) sor_cp_async16(&s_ring[k][1][t], pf_q);              // new phi(1,jf): written by another SM, L2 only
                else sor_cp_async8(&s_ring[k][1][t].x, pf_q);                      // old phi(i+1,jf)
            }
            sor_cp_async8(&s_ring[k][2][t].x, pf_rho);
            if (up_from_global) sor_cp_async16(&s_ring[k][3][t], pf_p);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        jf++; pf_right++; pf_q++; pf_rho++; pf_p++;
    };

    double left = 0.0, center = 0.0, saved_col1 = 0.0;
    if (active) { left = __ldcg(&row[niy - 2]); center = __ldcg(&row[0]); }   // r wrap for j == 0 (old), phi_old(i,0)
#pragma unroll
    for (int k = 0; k < D; k++) fetch(k);

    const int nsteps = niy + rows_here - 1;
    int j = -t;                                                   // column of this step
    double *pout = phi + ((long long)(active ? i : 0) * niy + j);    // &phi[i][j], advanced with j
    for (int s0 = 0; s0 < nsteps; s0 += D) {
#pragma unroll
        for (int k = 0; k < D; k++) {
            const int s = s0 + k;
            if (s < nsteps) {               // uniform across the CTA
                const bool work = active && j >= 0 && j < niy;
                asm volatile("cp.async.wait_group %0;" ::"n"(D - 1) : "memory");   // the group that filled slot k has landed
                if (work) {
                    const double right = (j == niy - 1) ? saved_col1 : s_ring[k][0][t].x;
                    // (the fetch pointers are D columns ahead of j)
                    const double down = q_prefetch_ok ? (q_is_new ? sor_pick(s_ring[k][1][t], pf_q - D) : s_ring[k][1][t].x)
                                                      : __ldcg(&row_q[j]);
                    const double rh = s_ring[k][2][t].x;
                    const double up = up_from_global ? sor_pick(s_ring[k][3][t], pf_p - D)
                                                     : s_new[(s + 1) & 1][t - 1];   // written at step s-1 by thread t-1
                    const double g = coef * (sor_div(up + down, dx2, rdx2) + sor_div(left + right, dy2, rdy2) + (rh / eps));
                    const double v = center + 1.4 * (g - center);
                    __stcg(pout, v);
                    s_new[s & 1][t] = v;
                    if (j == 1) saved_col1 = v;
                    left = v;
                    center = right;         // phi_old(i, j+1) is the next centre
                    if (last_row_of_band && ((j & 15) == 15 || j == niy - 1)) {   // one gpu-scope fence (~770 cycles) per 16 columns
                        __threadfence();
                        *(volatile int *)&progress[b] = j + 1;
                    }
                }
                fetch(k);                   // refill this slot for step s + D
                j++; pout++;
                __syncthreads();
            }
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ---------------------------------------------------------------------------
// SOR sweep 0 for SMALL grids (the shipped input.ini: 65 x 65 nodes): phi and rho live in shared memory for the
// whole sweep, one thread per grid row, one block barrier per anti-diagonal.  Same lexicographic iterate and the
// same arithmetic as the other two kernels (bit-identical, tested); what changes is the cost of a step: four LDS and
// a barrier among 3 warps instead of an L2 round trip (k_sor_solve) or the ring / progress machinery of the
// pipelined sweep — ~150 instead of ~560 ns per anti-diagonal.  The row pitch is even, so the threads of a
// diagonal (stride pitch - 1 doubles) fall on distinct banks.
// ---------------------------------------------------------------------------
__host__ __device__ inline int sor_smem_pitch(int niy) { return niy + (niy & 1); }
inline size_t sor_smem_bytes(int nix, int niy) { return 2 * sizeof(double) * (size_t)nix * sor_smem_pitch(niy); }

__global__ void __launch_bounds__(1024, 1)
k_sor_sweep_smem(double *phi, const double *__restrict__ rho, int nix, int niy, double dx, double dy) {
    extern __shared__ __align__(16) unsigned char sor_smem[];
    const int pitch = sor_smem_pitch(niy);
    double *sphi = reinterpret_cast<double *>(sor_smem), *srho = sphi + (size_t)nix * pitch;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int k = tid; k < nix * niy; k += nt) {
        const int i = k / niy, j = k - i * niy;
        sphi[i * pitch + j] = ldcg(&phi[k]);
        srho[i * pitch + j] = rho[k];
    }
    __syncthreads();
    const double dx2 = dx * dx, dy2 = dy * dy, eps = 1.0;
    const double coef = 0.5 * (1 / ((1 / dx2) + (1 / dy2)));
    const double rdx2 = 1 / dx2, rdy2 = 1 / dy2;
    const int i = tid;                                           // row of this thread (blockDim.x >= nix)
    const int p = (i - 1 < 0) ? nix - 2 : i - 1;                 // src/main.cpp:916-919
    const int q = (i + 1 > nix - 1) ? 1 : i + 1;
    double *rowp = sphi + i * pitch;
    const double *up_row = sphi + p * pitch, *down_row = sphi + q * pitch, *rho_row = srho + i * pitch;
    for (int d = 0; d <= nix + niy - 2; d++) {
        const int j = d - i;
        if (i < nix && j >= 0 && j < niy) {
            const int r = (j - 1 < 0) ? niy - 2 : j - 1;
            const int sidx = (j + 1 > niy - 1) ? 1 : j + 1;
            const double g = coef * (sor_div(up_row[j] + down_row[j], dx2, rdx2) + sor_div(rowp[r] + rowp[sidx], dy2, rdy2) +
                                     (rho_row[j] / eps));
            const double old = rowp[j];
            rowp[j] = old + 1.4 * (g - old);
        }
        __syncthreads();
    }
    for (int k = tid; k < nix * niy; k += nt) {
        const int ii = k / niy, j = k - ii * niy;
        __stcg(&phi[k], sphi[ii * pitch + j]);
    }
}

// residual of the reference's convergence test (src/main.cpp:930-950), multi-CTA, fixed tree
__global__ void k_sor_residual_partial(const double *__restrict__ phi, const double *__restrict__ rho, int nix, int niy,
                                       double dx, double *__restrict__ partial) {
    __shared__ double s_red[32];
    const double dx2 = dx * dx, eps = 1.0;
    const long long nn = (long long)nix * niy;
    double sum = 0.0;
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < nn; k += (long long)gridDim.x * blockDim.x) {
        int i = (int)(k / niy), j = (int)(k % niy);
        int p = i - 1; if (p < 0) p = nix - 2;
        int q = i + 1; if (q > nix - 1) q = 1;
        int r = j - 1; if (r < 0) r = niy - 2;
        int s = j + 1; if (s > niy - 1) s = 1;
        double R = 0.25 * (phi[(long long)p * niy + j] + phi[(long long)q * niy + j] + phi[(long long)i * niy + r] +
                           phi[(long long)i * niy + s] + (dx2 * rho[k] / eps)) - phi[k];
        sum = sum + (R * R);
    }
    double tsum = block_sum(sum, s_red);
    if (threadIdx.x == 0) partial[blockIdx.x] = tsum;
}
// status[0] = 1 (converged after 1 sweep) or 0 (keep sweeping); d_l2 = L2
__global__ void k_sor_residual_final(const double *__restrict__ partial, int n, int nix, int niy, long long *status,
                                     double *d_l2) {
    __shared__ double s_red[32];
    double s = 0.0;
    for (int k = threadIdx.x; k < n; k += blockDim.x) s += partial[k];
    double t = block_sum(s, s_red);
    if (threadIdx.x == 0) {
        double L2 = sqrt(t) / (nix * niy);
        *d_l2 = L2;
        status[0] = (L2 < 1e-2) ? 1 : 0;
    }
}

// ===========================================================================
// PICSP_FLAG_WALLS — EXTENSION WITHOUT REFERENCE SEMANTICS (BASELINE.json config 3, "bounded domain with wall
// boundaries").  The reference is periodic-only; these kernels are checked against the repo's own CPU restatement
// (test infrastructure outside the product tree), not against the reference.
// ===========================================================================
// grid phase without periodic folds: den += w * acc * 2^-frac on every node; rho = q_i*den_i + q_e*den_e on interior
// nodes, 0 on the walls (Dirichlet nodes carry no equation)
__global__ void k_grid_phase_walls(GridPhaseSpecies s0, GridPhaseSpecies s1, double *__restrict__ rho, int nix, int niy, int clear,
                                   const int *__restrict__ err, int *__restrict__ err_host) {
    if (blockIdx.x == 0 && threadIdx.x == 0 && *err) *err_host = *err;
    const double sc0 = s0.weight * exp2((double)(-*s0.frac)), sc1 = s1.weight * exp2((double)(-*s1.frac));
    const unsigned nn = (unsigned)nix * (unsigned)niy;
    for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < nn; k += gridDim.x * blockDim.x) {
        const int i = (int)(k / (unsigned)niy), j = (int)(k - (unsigned)i * (unsigned)niy);
        const long long a0 = s0.acc[k], a1 = s1.acc[k];
        const double di = (clear ? 0.0 : s0.den[k]) + (double)a0 * sc0, de = (clear ? 0.0 : s1.den[k]) + (double)a1 * sc1;
        s0.acc[k] = 0; s1.acc[k] = 0;
        s0.den[k] = di; s1.den[k] = de;
        rho[k] = (i > 0 && i < nix - 1 && j > 0 && j < niy - 1) ? __dadd_rn(__dmul_rn(s0.q, di), __dmul_rn(s1.q, de)) : 0.0;
    }
}
__global__ void k_zero_walls(double *__restrict__ f, int nix, int niy) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < 2 * (nix + niy); k += gridDim.x * blockDim.x) {
        if (k < niy) f[k] = 0.0;
        else if (k < 2 * niy) f[(long long)(nix - 1) * niy + (k - niy)] = 0.0;
        else if (k < 2 * niy + nix) f[(long long)(k - 2 * niy) * niy] = 0.0;
        else f[(long long)(k - 2 * niy - nix) * niy + niy - 1] = 0.0;
    }
}

// Red-black Gauss-Seidel / SOR for  lap(phi) = -rho  with phi = 0 on the four walls (5-point stencil, dx == dy):
//     g = 0.25 * ((phi[i-1][j] + phi[i+1][j]) + (phi[i][j-1] + phi[i][j+1]) + dx^2 * rho[i][j]);   phi += omega * (g - phi)
// first on the nodes with (i + j) even, then on the odd ones; warm-started from the previous phi.  After every `batch`
// sweeps the residual  L2 = sqrt(sum_interior R^2) / (nix * niy),  R = g - phi,  is evaluated (fixed reduction tree) and
// the iteration stops when L2 < tol (or after max_sweeps).  ONE cooperative launch: colours and the residual test
// are separated by grid-wide barriers, no host round trip.  Every operation is explicitly rounded (no contraction), so
// the iterate after a given number of sweeps is bit-identical to the CPU restatement used by the tests.
// status[0] = sweeps done (negative: the cap was hit), *d_l2 = last residual.
__global__ void __launch_bounds__(256)
k_rb_sor(double *phi, const double *__restrict__ rho, int nix, int niy, double dx, double omega, double tol, int max_sweeps,
         int batch, long long *status, double *d_l2, double *partial) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ double s_red[32];
    __shared__ double s_l2;
    const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    const double dx2 = __dmul_rn(dx, dx);
    const int ni = nix - 2, nj = niy - 2;
    const int half = (nj + 1) / 2;
    for (long long k = tid; k < 2ll * (nix + niy); k += nth) {      // the walls
        if (k < niy) phi[k] = 0.0;
        else if (k < 2 * niy) phi[(long long)(nix - 1) * niy + (k - niy)] = 0.0;
        else if (k < 2 * niy + nix) phi[(k - 2 * niy) * niy] = 0.0;
        else phi[(k - 2 * niy - nix) * niy + niy - 1] = 0.0;
    }
    grid.sync();
    int sweep = 0;
    double L2 = 1e300;
    while (sweep < max_sweeps) {
        for (int colour = 0; colour < 2; colour++) {
            for (long long k = tid; k < (long long)ni * half; k += nth) {
                const int i = 1 + (int)(k / half), jj = (int)(k % half);
                const int j = 1 + 2 * jj + ((i + 1 + colour) & 1);          // (i + j) & 1 == colour
                if (j <= niy - 2) {
                    const long long c = (long long)i * niy + j;
                    const double a = __dadd_rn(__ldcg(&phi[c - niy]), __ldcg(&phi[c + niy]));
                    const double b = __dadd_rn(__ldcg(&phi[c - 1]), __ldcg(&phi[c + 1]));
                    const double g = __dmul_rn(0.25, __dadd_rn(__dadd_rn(a, b), __dmul_rn(dx2, rho[c])));
                    const double old = __ldcg(&phi[c]);
                    __stcg(&phi[c], __dadd_rn(old, __dmul_rn(omega, __dadd_rn(g, -old))));
                }
            }
            grid.sync();
        }
        sweep++;
        if (sweep % batch == 0 || sweep == max_sweeps) {
            double sum = 0.0;
            for (long long k = tid; k < (long long)ni * nj; k += nth) {
                const int i = 1 + (int)(k / nj), j = 1 + (int)(k % nj);
                const long long c = (long long)i * niy + j;
                const double a = __dadd_rn(__ldcg(&phi[c - niy]), __ldcg(&phi[c + niy]));
                const double b = __dadd_rn(__ldcg(&phi[c - 1]), __ldcg(&phi[c + 1]));
                const double R = __dadd_rn(__dmul_rn(0.25, __dadd_rn(__dadd_rn(a, b), __dmul_rn(dx2, rho[c]))), -__ldcg(&phi[c]));
                sum = __dadd_rn(sum, __dmul_rn(R, R));
            }
            const double t = block_sum(sum, s_red);
            if (threadIdx.x == 0) __stcg(&partial[blockIdx.x], t);
            grid.sync();
            if (threadIdx.x == 0) {               // every CTA adds the same partials in the same order: one verdict for the grid
                double tot = 0.0;
                for (unsigned b = 0; b < gridDim.x; b++) tot += __ldcg(&partial[b]);
                s_l2 = sqrt(tot) / ((double)nix * (double)niy);
            }
            __syncthreads();
            L2 = s_l2;
            if (L2 < tol) break;
            grid.sync();                           // nobody rewrites `partial` or phi before everyone has read the verdict's inputs
        }
    }
    if (tid == 0) { status[0] = L2 < tol ? sweep : -(long long)sweep; *d_l2 = L2; }
}

// E = -grad(phi): central differences inside, full one-sided differences on the walls
__global__ void k_compute_ef_walls(const double *__restrict__ phi, double2 *__restrict__ E, int nix, int niy, double dx) {
    const long long nn = (long long)nix * niy;
    const double two_dx = __dmul_rn(2.0, dx);
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < nn; k += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(k / niy), j = (int)(k % niy);
        double2 e;
        if (i == 0) e.x = __ddiv_rn(__dadd_rn(phi[k], -phi[k + niy]), dx);
        else if (i == nix - 1) e.x = __ddiv_rn(__dadd_rn(phi[k - niy], -phi[k]), dx);
        else e.x = __ddiv_rn(__dadd_rn(phi[k - niy], -phi[k + niy]), two_dx);
        if (j == 0) e.y = __ddiv_rn(__dadd_rn(phi[k], -phi[k + 1]), dx);
        else if (j == niy - 1) e.y = __ddiv_rn(__dadd_rn(phi[k - 1], -phi[k]), dx);
        else e.y = __ddiv_rn(__dadd_rn(phi[k - 1], -phi[k + 1]), two_dx);
        E[k] = e;
    }
}

// ---------------------------------------------------------------------------
// E field (src/main.cpp:1111-1139).  Interior: central differences.  Rows i = 0 and
// i = nix-1 get the one-sided efx (divided by 2*dx as the reference does), columns
// j = 0 and j = niy-1 the one-sided efy.  efx[i][0], efx[i][niy-1] for interior i
// and efy[0][j], efy[nix-1][j] for interior j are never written (Q8).
// E is stored interleaved {efx, efy} so that a gather corner is one 16-byte load.
// ---------------------------------------------------------------------------
__global__ void k_compute_ef(const double *__restrict__ phi, double2 *__restrict__ E, int nix, int niy,
                             double dx, double dy) {
    long long nn = (long long)nix * niy;
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < nn;
         k += (long long)gridDim.x * blockDim.x) {
        int i = (int)(k / niy), j = (int)(k % niy);
        bool ii = (i >= 1 && i <= nix - 2), jj = (j >= 1 && j <= niy - 2);
        if (ii && jj) {
            double2 e;
            e.x = (phi[k - niy] - phi[k + niy]) / (2 * dx);
            e.y = (phi[k - 1] - phi[k + 1]) / (2 * dy);
            E[k] = e;
            continue;
        }
        if (i == 0)            E[k].x = -(phi[k + niy] - phi[k]) / (2 * dx);
        else if (i == nix - 1) E[k].x = -(phi[k] - phi[k - niy]) / (2 * dx);
        if (j == 0)            E[k].y = -(phi[k + 1] - phi[k]) / (2 * dy);
        else if (j == niy - 1) E[k].y = -(phi[k] - phi[k - 1]) / (2 * dy);
    }
}

// split/merge for the ABI's separate efx / efy views
__global__ void k_ef_get_component(const double2 *__restrict__ E, double *__restrict__ out, long long nn, int comp) {
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < nn; k += (long long)gridDim.x * blockDim.x)
        out[k] = comp ? E[k].y : E[k].x;
}
__global__ void k_ef_set_component(double2 *__restrict__ E, const double *__restrict__ in, long long nn, int comp) {
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < nn; k += (long long)gridDim.x * blockDim.x) {
        if (comp) E[k].y = in[k]; else E[k].x = in[k];
    }
}

// ---------------------------------------------------------------------------
// deterministic reductions (fixed grid, fixed tree): sum(vx^2+vy^2) and max(phi)
// ---------------------------------------------------------------------------
constexpr int RED_BLOCKS = 592;   // 4 per SM on 148 SMs
constexpr int RED_THREADS = 256;

__global__ void k_ke_partial(const double *__restrict__ vx, const double *__restrict__ vy, long long n,
                             double *__restrict__ partial) {
    __shared__ double s_red[32];
    double s = 0.0;
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x)
        s += vx[p] * vx[p] + vy[p] * vy[p];
    double t = block_sum(s, s_red);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
}
// KE terms written in UPLOAD order (stage[id[slot]]), so that the reduction tree — and the result, bit for
// bit — does not depend on how the tile sort happened to arrange the particles in memory.
__global__ void k_ke_terms(const double *__restrict__ vx, const double *__restrict__ vy, const unsigned int *__restrict__ id,
                           long long n, double *__restrict__ terms) {
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x)
        terms[id ? (long long)id[p] : p] = vx[p] * vx[p] + vy[p] * vy[p];
}
// KE terms of a row-layout snapshot ([n][4] rows {x, y, vx, vy} in upload order): fixed grid, fixed tree
__global__ void k_ke_rows_partial(const double *__restrict__ rows, long long n, double *__restrict__ partial) {
    __shared__ double s_red[32];
    double s = 0.0;
    const double2 *r2 = reinterpret_cast<const double2 *>(rows);
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
        const double2 v = r2[2 * p + 1];
        s += v.x * v.x + v.y * v.y;
    }
    double t = block_sum(s, s_red);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
}
__global__ void k_sum_partial(const double *__restrict__ a, long long n, double *__restrict__ partial) {
    __shared__ double s_red[32];
    double s = 0.0;
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) s += a[p];
    double t = block_sum(s, s_red);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
}
__global__ void k_sum_final(const double *__restrict__ partial, int n, double *__restrict__ out) {
    __shared__ double s_red[32];
    double s = 0.0;
    for (int k = threadIdx.x; k < n; k += blockDim.x) s += partial[k];
    double t = block_sum(s, s_red);
    if (threadIdx.x == 0) *out = t;
}
__global__ void k_max_partial(const double *__restrict__ f, long long n, double *__restrict__ partial) {
    __shared__ double s_red[32];
    double m = f[0];
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x)
        m = f[k] > m ? f[k] : m;
    for (int o = 16; o > 0; o >>= 1) { double t = __shfl_xor_sync(0xffffffffu, m, o); m = t > m ? t : m; }
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 0; w < (blockDim.x + 31) / 32; w++) m = s_red[w] > m ? s_red[w] : m;
        partial[blockIdx.x] = m;
    }
}
__global__ void k_max_final(const double *__restrict__ partial, int n, const double *__restrict__ f,
                            double *__restrict__ out_max, double *__restrict__ out_f0) {
    __shared__ double s_red[32];
    double m = f[0];
    for (int k = threadIdx.x; k < n; k += blockDim.x) m = partial[k] > m ? partial[k] : m;
    for (int o = 16; o > 0; o >>= 1) { double t = __shfl_xor_sync(0xffffffffu, m, o); m = t > m ? t : m; }
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 0; w < (blockDim.x + 31) / 32; w++) m = s_red[w] > m ? s_red[w] : m;
        *out_max = m; *out_f0 = f[0];
    }
}

}  // namespace picsp
