// picsp_b200/csrc/fft_kernels.cuh — the 2-D real DFT of spectralPotentialSolver (src/main.cpp:960-1058) for node
// counts with a LARGE PRIME FACTOR, each 1-D transform resident in shared memory.
//
// The reference transforms the (N+1) x (N+1) NODE array (Q7), so a power-of-two cell count gives awkward lengths:
// 2049 = 3 * 683, 257 (prime), 4097 = 17 * 241.  cuFFT falls back to Bluestein's algorithm through global memory there
// (2049^2: 1.5-1.7 ms per solve, 19 % of BASELINE config 5's step on 8 GPUs).  Here a length M = P * Q, gcd(P, Q) = 1, is
// split by the prime-factor map (Good-Thomas, no twiddle factors) into P transforms of length Q and Q of length P:
//   * length Q (any Q, prime or not) by Bluestein's chirp-z identity
//         X[k] = c[k] * sum_n (x[n] c[n]) conj(c)[k - n],   c[n] = exp(-i pi n^2 / Q),
//     i.e. one circular convolution of power-of-two length L >= 2Q - 1: decimation in frequency forward (natural ->
//     bit-reversed order), product with the pre-transformed chirp filter (stored in that same order, 1/L folded in),
//     decimation in time inverse (bit-reversed -> natural): no reordering pass.  Three radix-2 stages per pass on 8
//     elements held in registers (radix 8: 4 passes over shared memory for L = 2048 instead of 11), one twiddle
//     look-up per thread and pass;
//   * length P (<= 32) as a direct DFT from a P x P table.
// All P rows of a transform sit in shared memory at once (P * L complex doubles: 96 KB for 2049), one CTA per
// transform; the four passes of the 2-D transform are
//   rows forward   two REAL rows packed into one complex transform, separated into two half spectra,
//   columns forward / inverse   in place on the [Nx][Ny/2+1] half-spectrum array (same layout as cuFFT's D2Z),
//   rows inverse   two Hermitian half rows packed into one complex transform whose real / imaginary parts are the rows.
// The inverse transform is conj(DFT(conj(.))).  Tables are computed on the host in long double.
// Semantics are exactly cuFFT's D2Z / Z2D (unnormalised), so k_kspace_green sits between the passes unchanged and the
// cuFFT path stays as a cross-check (PICSP_FLAG_CUFFT_ONLY).
#pragma once
#include "ctx.cuh"

namespace picsp {

constexpr int FFT_THREADS = 512;       // launch bound; the launch uses fft_threads() <= this

struct BluePlanDev {
    int kind;                // 0: prime-factor split + Bluestein (below); 1: two DIRECT coprime factors P x Q, both <= 64 (pfa2_transform)
    int M, P, Q, L, logL;
    const double2 *chirp;    // [Q]      c[n]
    const double2 *bhat;     // [L]      DIF(b) / L, b[m] = conj(c[|m|]) wrapped, in DIF output order
    const double2 *tw;       // [L/2]    exp(-2 pi i k / L)
    const int *in_pos;       // [M]      n -> n1 * L + n2   (n = (n1 Q + n2 P) mod M)
    const int *out_idx;      // [P * Q]  k1 * Q + k2 -> k   (k = k1 mod P, k = k2 mod Q)
    const double2 *wp;       // [P * P]  exp(-2 pi i k1 n1 / P)
    const double2 *rootP;    // kind 1: [P] exp(-2 pi i m / P)
    const double2 *rootQ;    // kind 1: [Q] exp(-2 pi i m / Q)
};

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ double2 cmulc(double2 a, double2 b) {     // a * conj(b)
    return make_double2(fma(a.x, b.x, a.y * b.y), fma(a.y, b.x, -a.x * b.y));
}

// Shared-memory index with one complex of padding per 8: a thread of the last passes owns 8 CONSECUTIVE elements, so
// without it the lanes of a quarter-warp would all hit the same 16-byte bank group.
__device__ __forceinline__ int padx(int i) { return i + (i >> 3); }
__host__ __device__ constexpr size_t fft_padded(size_t n) { return n + (n >> 3) + 1; }

// multiply by exp(-2 pi i j / 2^(T+1)), j < 2^T: the rotations inside a radix-2^R butterfly
template <int T> __device__ __forceinline__ double2 rot_fwd(double2 d, int j) {
    constexpr double H = 0.70710678118654752440;
    if (T == 0 || j == 0) return d;
    if (T == 1) return make_double2(d.y, -d.x);                                   // -i
    if (j == 1) return make_double2((d.x + d.y) * H, (d.y - d.x) * H);             // exp(-i pi/4)
    if (j == 2) return make_double2(d.y, -d.x);
    return make_double2((d.y - d.x) * H, -(d.x + d.y) * H);                        // exp(-3 i pi/4)
}
template <int T> __device__ __forceinline__ double2 rot_inv(double2 d, int j) {   // the conjugate rotations
    constexpr double H = 0.70710678118654752440;
    if (T == 0 || j == 0) return d;
    if (T == 1) return make_double2(-d.y, d.x);                                   // +i
    if (j == 1) return make_double2((d.x - d.y) * H, (d.x + d.y) * H);
    if (j == 2) return make_double2(-d.y, d.x);
    return make_double2(-(d.x + d.y) * H, (d.x - d.y) * H);
}

// R radix-2 stages on the 2^R elements a thread holds in registers (halves s * 2^(R-1) .. s).  The twiddle of the pair
// (m, m + 2^t) in stage t is b^(2^(R-1-t)) * exp(-2 pi i (m mod 2^t) / 2^(t+1)), b = W_L^(lowpart * L / (2 hmax)):
// ONE table look-up per thread and pass, the rest are squarings and fixed rotations.
template <int R, int T> struct DifStage {
    static __device__ __forceinline__ void run(double2 (&v)[1 << R], double2 bt) {
#pragma unroll
        for (int m = 0; m < (1 << R); m++) {
            if (m & (1 << T)) continue;
            const double2 a = v[m], c = v[m + (1 << T)];
            v[m] = make_double2(a.x + c.x, a.y + c.y);
            v[m + (1 << T)] = rot_fwd<T>(cmul(make_double2(a.x - c.x, a.y - c.y), bt), m & ((1 << T) - 1));
        }
        DifStage<R, T - 1>::run(v, cmul(bt, bt));
    }
};
template <int R> struct DifStage<R, -1> { static __device__ __forceinline__ void run(double2 (&)[1 << R], double2) {} };

template <int R, int T> struct DitStage {       // stages t = 0 .. R-1; bpow[t] = b^(2^(R-1-t))
    static __device__ __forceinline__ void run(double2 (&v)[1 << R], const double2 (&bpow)[R]) {
#pragma unroll
        for (int m = 0; m < (1 << R); m++) {
            if (m & (1 << T)) continue;
            const double2 a = v[m];
            const double2 c = cmulc(rot_inv<T>(v[m + (1 << T)], m & ((1 << T) - 1)), bpow[T]);
            v[m] = make_double2(a.x + c.x, a.y + c.y);
            v[m + (1 << T)] = make_double2(a.x - c.x, a.y - c.y);
        }
        DitStage<R, T + 1>::run(v, bpow);
    }
};
template <int R> struct DitStage<R, R> { static __device__ __forceinline__ void run(double2 (&)[1 << R], const double2 (&)[R]) {} };

// one pass of R stages over `rows` rows of length L; ltop = log2 of the largest half of the pass.
// FWD: decimation in frequency (halves descending); otherwise the mirrored decimation in time with conjugated twiddles.
// MULB (inverse passes only): the elements are multiplied by bhat as they are loaded (the convolution product).
template <int R, bool FWD, bool MULB>
__device__ __forceinline__ void fft_pass(double2 *w, int rows, int L, int logL, int ltop, const double2 *__restrict__ tw,
                                         const double2 *__restrict__ bhat) {
    const int ls = ltop - R + 1, s = 1 << ls;
    const int gpr = L >> R, ng = rows * gpr;
    for (int g = threadIdx.x; g < ng; g += blockDim.x) {
        const int row = g >> (logL - R), gg = g & (gpr - 1);
        const int low = gg & (s - 1);
        const int off = ((gg >> ls) << (ls + R)) + low, base = row * L + off;
        double2 v[1 << R];
#pragma unroll
        for (int m = 0; m < (1 << R); m++) {
            v[m] = w[padx(base + m * s)];
            if (MULB) v[m] = cmul(v[m], bhat[off + m * s]);
        }
        const double2 b = tw[low << (logL - 1 - ltop)];
        if (FWD) {
            DifStage<R, R - 1>::run(v, b);
        } else {
            double2 bpow[R];
            bpow[R - 1] = b;
#pragma unroll
            for (int t = R - 2; t >= 0; t--) bpow[t] = cmul(bpow[t + 1], bpow[t + 1]);
            DitStage<R, 0>::run(v, bpow);
        }
#pragma unroll
        for (int m = 0; m < (1 << R); m++) w[padx(base + m * s)] = v[m];
    }
    __syncthreads();
}

// forward transform of every row (natural order in, bit-reversed out): passes of 3 bits, the remainder last
__device__ __forceinline__ void fft_dif(double2 *w, int rows, int L, int logL, const double2 *__restrict__ tw) {
    int ltop = logL - 1;
    for (; ltop >= 2; ltop -= 3) fft_pass<3, true, false>(w, rows, L, logL, ltop, tw, nullptr);
    if (ltop == 1) fft_pass<2, true, false>(w, rows, L, logL, ltop, tw, nullptr);
    else if (ltop == 0) fft_pass<1, true, false>(w, rows, L, logL, ltop, tw, nullptr);
}
// its mirror image (bit-reversed in, natural out; un-normalised inverse), the product with bhat folded into the first pass
__device__ __forceinline__ void fft_dit_inv_mulb(double2 *w, int rows, int L, int logL, const double2 *__restrict__ tw,
                                                 const double2 *__restrict__ bhat) {
    const int rem = logL % 3;                  // the forward transform ended with a pass of `rem` bits: start with it
    int ltop;
    if (rem == 2) { fft_pass<2, false, true>(w, rows, L, logL, 1, tw, bhat); ltop = 4; }
    else if (rem == 1) { fft_pass<1, false, true>(w, rows, L, logL, 0, tw, bhat); ltop = 3; }
    else { fft_pass<3, false, true>(w, rows, L, logL, 2, tw, bhat); ltop = 5; }
    for (; ltop <= logL - 1; ltop += 3) fft_pass<3, false, false>(w, rows, L, logL, ltop, tw, nullptr);
}

// One length-M DFT in shared memory.  load(n) yields input element n, store(k, v) receives output element k; with
// INVERSE the un-normalised inverse transform is computed (conjugate in, conjugate out).  `work` holds P * L complex.
template <bool INVERSE, class Load, class Store>
__device__ __forceinline__ void blue_transform(const BluePlanDev &pl, double2 *work, Load load, Store store) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int PL = pl.P * pl.L;
    for (int i = tid; i < PL; i += nt)
        if ((i & (pl.L - 1)) >= pl.Q) work[padx(i)] = make_double2(0.0, 0.0);       // zero padding of every row
    for (int n = tid; n < pl.M; n += nt) {
        double2 v = load(n);
        if (INVERSE) v.y = -v.y;
        const int pos = pl.in_pos[n];
        work[padx(pos)] = cmul(v, pl.chirp[pos & (pl.L - 1)]);
    }
    __syncthreads();
    fft_dif(work, pl.P, pl.L, pl.logL, pl.tw);
    fft_dit_inv_mulb(work, pl.P, pl.L, pl.logL, pl.tw, pl.bhat);
    // length-P transforms across the rows, one output element per (k1, k2)
    const int PQ = pl.P * pl.Q;
    for (int o = tid; o < PQ; o += nt) {
        const int k1 = o / pl.Q, k2 = o - k1 * pl.Q;
        double2 acc = make_double2(0.0, 0.0);
        for (int n1 = 0; n1 < pl.P; n1++) {
            const double2 y = work[padx(n1 * pl.L + k2)], wv = pl.wp[k1 * pl.P + n1];
            acc.x += fma(y.x, wv.x, -y.y * wv.y);
            acc.y += fma(y.x, wv.y, y.y * wv.x);
        }
        double2 v = cmul(acc, pl.chirp[k2]);
        if (INVERSE) v.y = -v.y;
        store(pl.out_idx[o], v);
    }
}

// ---------------------------------------------------------------------------
// Two direct factors (kind 1).  M = P * Q, coprime, both small (1025 = 25 * 41, 513 = 27 * 19, 65 = 5 * 13 ...): the
// prime-factor map turns the transform into Q-point DFTs along the rows of a P x Q array and P-point DFTs along its
// columns, both done DIRECTLY as small matrix products — no Bluestein padding, two passes over shared memory and three
// barriers in all.  A thread computes a block of output pairs (k, N - k) of one 1-D transform: it walks the inputs once
// and keeps, per pair, four real accumulators and the index of the current root of unity (advanced by k modulo the
// length: a table of the roots sits in shared memory).  Lanes of a warp work on neighbouring transforms of the same output block, so the
// root is a broadcast and the inputs are conflict-free.  Shared memory: 2 * M complex + the two root tables.
// ---------------------------------------------------------------------------
#ifndef PICSP_PFA2_KP
#define PICSP_PFA2_KP 3      // measured at 1025^2 (192 threads): 2 / 3 / 4 / 5 pairs -> 108 / 106 / 112 / 115 us
#endif
constexpr int PFA2_KP = PICSP_PFA2_KP;        // output PAIRS (k, N - k) per thread

// out[line][k] = sum_n W_N^(k n) in[line][n] for `lines` lines of length N; element (line, n) sits at line * ls + n * ns.
// Outputs k and N - k use conjugate roots, so with v = a + i b and W^(k n) = c - i s the four sums
//     S1 = sum a c,  S2 = sum b s,  S3 = sum b c,  S4 = sum a s
// give BOTH:  X[k] = (S1 + S2) + i (S3 - S4),  X[N-k] = (S1 - S2) + i (S3 + S4)  — four FMAs and one root per input for
// two outputs (k = 0, and k = N/2 for even N, pair with themselves and are emitted once).
template <class Emit>
__device__ __forceinline__ void pfa2_stage(const double2 *__restrict__ in, int lines, int N, int ls, int ns,
                                           const double2 *__restrict__ root, Emit emit) {
    const int npair = N / 2 + 1;
    const int nblk = (npair + PFA2_KP - 1) / PFA2_KP;
    for (int item = threadIdx.x; item < lines * nblk; item += blockDim.x) {
        const int blk = item / lines, line = item - blk * lines;       // consecutive lanes: consecutive lines, same output block
        const int k0 = blk * PFA2_KP;
        double s1[PFA2_KP], s2[PFA2_KP], s3[PFA2_KP], s4[PFA2_KP];
        int m[PFA2_KP], kk[PFA2_KP];
#pragma unroll
        for (int o = 0; o < PFA2_KP; o++) { s1[o] = s2[o] = s3[o] = s4[o] = 0.0; m[o] = 0; kk[o] = (k0 + o) % N; }
        const double2 *src = in + line * ls;
        for (int n = 0; n < N; n++) {
            const double2 v = src[n * ns];
#pragma unroll
            for (int o = 0; o < PFA2_KP; o++) {
                const double2 w = root[m[o]];                 // (c, -s)
                s1[o] = fma(v.x, w.x, s1[o]); s2[o] = fma(-v.y, w.y, s2[o]);
                s3[o] = fma(v.y, w.x, s3[o]); s4[o] = fma(-v.x, w.y, s4[o]);
                m[o] += kk[o];                                // (k n) mod N
                if (m[o] >= N) m[o] -= N;
            }
        }
#pragma unroll
        for (int o = 0; o < PFA2_KP; o++) {
            const int k = k0 + o;
            if (k >= npair) continue;
            emit(line, k, make_double2(s1[o] + s2[o], s3[o] - s4[o]));
            const int kc = N - k;
            if (k != 0 && kc != k) emit(line, kc, make_double2(s1[o] - s2[o], s3[o] + s4[o]));
        }
    }
}

template <bool INVERSE, class Load, class Store>
__device__ __forceinline__ void pfa2_transform(const BluePlanDev &pl, double2 *smem, Load load, Store store) {
    const int P = pl.P, Q = pl.Q, M = pl.M;
    double2 *buf0 = smem, *buf1 = smem + M, *rP = smem + 2 * M, *rQ = rP + P;
    for (int i = threadIdx.x; i < P; i += blockDim.x) rP[i] = pl.rootP[i];
    for (int i = threadIdx.x; i < Q; i += blockDim.x) rQ[i] = pl.rootQ[i];
    for (int n = threadIdx.x; n < M; n += blockDim.x) {
        double2 v = load(n);
        if (INVERSE) v.y = -v.y;
        buf0[pl.in_pos[n]] = v;                        // (n1, n2) at n1 * Q + n2
    }
    __syncthreads();
    // Q-point transforms along n2 for every n1: lines = n1 (stride Q), elements stride 1
    pfa2_stage(buf0, P, Q, Q, 1, rQ, [&](int n1, int k2, double2 v) { buf1[n1 * Q + k2] = v; });
    __syncthreads();
    // P-point transforms along n1 for every k2: lines = k2 (stride 1), elements stride Q
    pfa2_stage(buf1, Q, P, 1, Q, rP, [&](int k2, int k1, double2 v) {
        if (INVERSE) v.y = -v.y;
        store(pl.out_idx[k1 * Q + k2], v);
    });
}

// one 1-D transform of either kind (a template parameter: each kind gets kernels of its own, with its own register budget)
template <int KIND, bool INVERSE, class Load, class Store>
__device__ __forceinline__ void dft_transform(const BluePlanDev &pl, double2 *smem, Load load, Store store) {
    if (KIND == 1) pfa2_transform<INVERSE>(pl, smem, load, store);
    else blue_transform<INVERSE>(pl, smem, load, store);
}
#ifndef PICSP_PFA2_THREADS
#define PICSP_PFA2_THREADS 224     // upper bound; the launch uses the work items of the larger stage (1025 = 25 x 41: 175 and 205 -> 224: 97 us)
#endif
#ifndef PICSP_PFA2_CTAS
#define PICSP_PFA2_CTAS 4
#endif
constexpr int PFA2_THREADS = PICSP_PFA2_THREADS;     // 1025 = 25 x 41: 150 and 164 work items per stage
constexpr int BLUE_THREADS = 384;     // kind 0: at most this many threads, two CTAs per SM (<= 85 registers)
#define PICSP_FFT_BOUNDS(KIND) __launch_bounds__((KIND) == 1 ? PFA2_THREADS : BLUE_THREADS, (KIND) == 1 ? PICSP_PFA2_CTAS : 2)

// bhat = DIF(b) / L, computed with the very routine that will consume it (so the order matches by construction)
__global__ void __launch_bounds__(FFT_THREADS)
k_blue_bhat(const double2 *__restrict__ b, double2 *__restrict__ bhat, int L, int logL, const double2 *__restrict__ tw) {
    extern __shared__ __align__(16) unsigned char fft_smem[];
    double2 *w = reinterpret_cast<double2 *>(fft_smem);
    for (int i = threadIdx.x; i < L; i += blockDim.x) w[padx(i)] = b[i];
    __syncthreads();
    fft_dif(w, 1, L, logL, tw);
    const double s = 1.0 / (double)L;
    for (int i = threadIdx.x; i < L; i += blockDim.x) bhat[i] = make_double2(w[padx(i)].x * s, w[padx(i)].y * s);
}

// rows forward: real rows 2a, 2a+1 of x[nrows][M] -> half spectra T[2a][0..Nh), T[2a+1][0..Nh)
template <int KIND>
__global__ void PICSP_FFT_BOUNDS(KIND)
k_fft_rows_fwd(BluePlanDev pl, const double *__restrict__ x, double2 *__restrict__ T, int nrows) {
    extern __shared__ __align__(16) unsigned char fft_smem[];
    double2 *work = reinterpret_cast<double2 *>(fft_smem);
    // the packed spectrum Z[k] goes to the columns [Q, 2Q) of the work rows (free once the convolution is done; L >= 2Q):
    // element k sits in row k / Q, column Q + k % Q
    // (kind 1: the first buffer is free once the first stage has read it)
    const int M = pl.M, Nh = M / 2 + 1, Q = pl.Q, L = pl.L;
    constexpr bool direct = KIND == 1;
    auto zpos = [Q, L, direct](int k) { if (direct) return k; const int r = k / Q; return padx(r * L + Q + (k - r * Q)); };
    const int ra = 2 * blockIdx.x, rb = ra + 1;
    const double *xa = x + (size_t)ra * M, *xb = x + (size_t)rb * M;
    const bool has_b = rb < nrows;
    dft_transform<KIND, false>(pl, work,
                          [&](int n) { return make_double2(xa[n], has_b ? xb[n] : 0.0); },
                          [&](int k, double2 v) { work[zpos(k)] = v; });
    __syncthreads();
    for (int k = threadIdx.x; k < Nh; k += blockDim.x) {
        const double2 zk = work[zpos(k)], zm = work[zpos(k == 0 ? 0 : M - k)];
        // A = (Z[k] + conj Z[M-k]) / 2,  B = (Z[k] - conj Z[M-k]) / (2i)
        T[(size_t)ra * Nh + k] = make_double2(0.5 * (zk.x + zm.x), 0.5 * (zk.y - zm.y));
        if (has_b) T[(size_t)rb * Nh + k] = make_double2(0.5 * (zk.y + zm.y), -0.5 * (zk.x - zm.x));
    }
}

// columns, in place on T[M][Nh]: column blockIdx.x
template <int KIND, bool INVERSE>
__global__ void PICSP_FFT_BOUNDS(KIND)
k_fft_cols(BluePlanDev pl, double2 *__restrict__ T, int Nh) {
    extern __shared__ __align__(16) unsigned char fft_smem[];
    double2 *work = reinterpret_cast<double2 *>(fft_smem);
    double2 *col = T + blockIdx.x;
    dft_transform<KIND, INVERSE>(pl, work,
                            [&](int n) { return col[(size_t)n * Nh]; },
                            [&](int k, double2 v) { col[(size_t)k * Nh] = v; });      // every load is done before the first store
}

// rows inverse: Hermitian half rows U[2a], U[2a+1] -> real rows 2a, 2a+1 of out[nrows][M] (un-normalised, like Z2D)
template <int KIND>
__global__ void PICSP_FFT_BOUNDS(KIND)
k_fft_rows_inv(BluePlanDev pl, const double2 *__restrict__ U, double *__restrict__ out, int nrows) {
    extern __shared__ __align__(16) unsigned char fft_smem[];
    double2 *work = reinterpret_cast<double2 *>(fft_smem);
    const int M = pl.M, Nh = M / 2 + 1;
    const int ra = 2 * blockIdx.x, rb = ra + 1;
    const bool has_b = rb < nrows;
    const double2 *ua = U + (size_t)ra * Nh, *ub = U + (size_t)rb * Nh;
    double *oa = out + (size_t)ra * M, *ob = out + (size_t)rb * M;
    dft_transform<KIND, true>(pl, work,
                         [&](int k) {
                             // Z[k] = A[k] + i B[k]; beyond the stored half the rows are Hermitian: conj of the mirror
                             const bool lo = k < Nh;
                             const int q = lo ? k : M - k;
                             double2 a = ua[q], b = has_b ? ub[q] : make_double2(0.0, 0.0);
                             if (!lo) { a.y = -a.y; b.y = -b.y; }
                             if (q == 0 || 2 * q == M) { a.y = 0.0; b.y = 0.0; }     // self-conjugate bins are real (a c2r transform ignores their imaginary part)
                             return make_double2(a.x - b.y, a.y + b.x);
                         },
                         [&](int n, double2 v) { oa[n] = v.x; if (has_b) ob[n] = v.y; });
}

}  // namespace picsp
