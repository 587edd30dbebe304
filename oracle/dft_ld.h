/* oracle/dft_ld.h — TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Extended-precision (x87 long double) discrete Fourier transforms used as the
 * stand-in for FFTW3 at the two call sites the reference has
 * (/root/reference/src/main.cpp:990-993 r2c_2d, :1045-1048 c2r_2d).  FFTW3 is an
 * external, un-vendored, version-unpinned dependency (reference makefile:22
 * `-lfftw3`) that is absent from this image, so its *published contract* is
 * restated here: unnormalised DFT, forward sign -1, backward sign +1, the 2-D
 * r2c keeps n1/2+1 bins of the last dimension, the 2-D c2r does complex
 * backward DFTs over the leading dimension followed by a Hermitian-completing
 * real transform over the last one (imaginary parts of the self-conjugate bins
 * are ignored).
 *
 * Two 1-D engines: a direct O(n^2) sum with exactly reduced twiddle indices
 * (n <= ORACLE_DFT_DIRECT_MAX) and Bluestein's chirp-z on a radix-2 FFT for
 * larger n.  Both work for any n (the reference transforms numxCells+1 points,
 * which is never a power of two).  Everything is long double, so the result is
 * a *stricter* oracle than double-precision FFTW.
 *
 * Plain C99 + also valid C++ (included from oracle/shims/fftw3.h).
 */
#ifndef ORACLE_DFT_LD_H
#define ORACLE_DFT_LD_H

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef ORACLE_DFT_DIRECT_MAX
#define ORACLE_DFT_DIRECT_MAX 300
#endif

typedef struct { long double re, im; } oracle_cld;

static const long double ORACLE_PI_L = 3.14159265358979323846264338327950288L;

/* 0 = automatic (direct for small n, Bluestein above), 1 = force direct,
 * 2 = force Bluestein, 3 = fast double-precision Bluestein with cached plans (timing).  Settable by the harness for timing vs accuracy runs. */
static int oracle_dft_mode = 0;

static inline oracle_cld oracle_cmul(oracle_cld a, oracle_cld b) {
    oracle_cld r;
    r.re = a.re * b.re - a.im * b.im;
    r.im = a.re * b.im + a.im * b.re;
    return r;
}

/* in-place radix-2 FFT, n power of two, sign = -1 forward / +1 backward */
static void oracle_fft_pow2(oracle_cld *a, int n, int sign) {
    int i, j, len;
    for (i = 1, j = 0; i < n; i++) {
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { oracle_cld t = a[i]; a[i] = a[j]; a[j] = t; }
    }
    for (len = 2; len <= n; len <<= 1) {
        int half = len >> 1, k;
        /* twiddles computed directly (no recurrence) for accuracy */
        oracle_cld *w = (oracle_cld *)malloc(sizeof(oracle_cld) * (size_t)half);
        for (k = 0; k < half; k++) {
            long double ang = sign * 2.0L * ORACLE_PI_L * (long double)k / (long double)len;
            w[k].re = cosl(ang); w[k].im = sinl(ang);
        }
        for (i = 0; i < n; i += len) {
            for (k = 0; k < half; k++) {
                oracle_cld u = a[i + k];
                oracle_cld v = oracle_cmul(a[i + k + half], w[k]);
                a[i + k].re = u.re + v.re; a[i + k].im = u.im + v.im;
                a[i + k + half].re = u.re - v.re; a[i + k + half].im = u.im - v.im;
            }
        }
        free(w);
    }
}

/* out[k] = sum_j in[j] * exp(sign*2*pi*i*j*k/n); strides in elements */
static void oracle_dft_1d(const oracle_cld *in, int istride, oracle_cld *out, int ostride,
                          int n, int sign) {
    int use_direct = (oracle_dft_mode == 1) || (oracle_dft_mode == 0 && n <= ORACLE_DFT_DIRECT_MAX);
    int j, k;
    if (n == 1) { out[0] = in[0]; return; }
    if (use_direct) {
        oracle_cld *tw = (oracle_cld *)malloc(sizeof(oracle_cld) * (size_t)n);
        oracle_cld *tmp = (oracle_cld *)malloc(sizeof(oracle_cld) * (size_t)n);
        for (k = 0; k < n; k++) {
            long double ang = sign * 2.0L * ORACLE_PI_L * (long double)k / (long double)n;
            tw[k].re = cosl(ang); tw[k].im = sinl(ang);
        }
        for (k = 0; k < n; k++) {
            long double sr = 0.0L, si = 0.0L;
            long long idx = 0;
            for (j = 0; j < n; j++) {
                oracle_cld x = in[(size_t)j * istride];
                oracle_cld w = tw[idx];
                sr += x.re * w.re - x.im * w.im;
                si += x.re * w.im + x.im * w.re;
                idx += k; if (idx >= n) idx -= n;
            }
            tmp[k].re = sr; tmp[k].im = si;
        }
        for (k = 0; k < n; k++) out[(size_t)k * ostride] = tmp[k];
        free(tw); free(tmp);
    } else {
        /* Bluestein: jk = (j^2 + k^2 - (k-j)^2)/2 */
        int m = 1;
        oracle_cld *c, *a, *b;
        while (m < 2 * n - 1) m <<= 1;
        c = (oracle_cld *)malloc(sizeof(oracle_cld) * (size_t)n);
        a = (oracle_cld *)calloc((size_t)m, sizeof(oracle_cld));
        b = (oracle_cld *)calloc((size_t)m, sizeof(oracle_cld));
        for (j = 0; j < n; j++) {
            long long j2 = ((long long)j * (long long)j) % (2LL * n);   /* exact reduction */
            long double ang = sign * ORACLE_PI_L * (long double)j2 / (long double)n;
            c[j].re = cosl(ang); c[j].im = sinl(ang);
        }
        for (j = 0; j < n; j++) {
            a[j] = oracle_cmul(in[(size_t)j * istride], c[j]);
            b[j].re = c[j].re; b[j].im = -c[j].im;
            if (j) b[m - j] = b[j];
        }
        oracle_fft_pow2(a, m, -1);
        oracle_fft_pow2(b, m, -1);
        for (j = 0; j < m; j++) a[j] = oracle_cmul(a[j], b[j]);
        oracle_fft_pow2(a, m, +1);
        for (k = 0; k < n; k++) {
            oracle_cld v;
            v.re = a[k].re / (long double)m; v.im = a[k].im / (long double)m;
            out[(size_t)k * ostride] = oracle_cmul(v, c[k]);
        }
        free(c); free(a); free(b);
    }
}

/* ---------------------------------------------------------------------------
 * Fast engine for TIMING runs (oracle_dft_mode == 3): double precision Bluestein
 * with cached plans (chirp, transformed kernel, twiddles, bit reversal).  Used by
 * bench.py's reference arm so that the shim FFT does not dominate the reference's
 * timed step; accuracy ~1e-15*log(n), still far inside the 1e-12 budget.
 * --------------------------------------------------------------------------- */
typedef struct { double re, im; } oracle_cd;
typedef struct oracle_fast_plan {
    int n, m, sign;
    oracle_cd *chirp;      /* n   */
    oracle_cd *kernel_f;   /* m: FFT of the conjugate chirp, pre-divided by m */
    oracle_cd *tw;         /* m/2 forward twiddles exp(-2 pi i k/m) */
    int *rev;              /* m   */
    oracle_cd *work;       /* m   */
    struct oracle_fast_plan *next;
} oracle_fast_plan;
static oracle_fast_plan *oracle_fast_plans = NULL;

static void oracle_fast_fft(const oracle_fast_plan *p, oracle_cd *a, int inverse) {
    int m = p->m, i, len;
    for (i = 0; i < m; i++) { int j = p->rev[i]; if (i < j) { oracle_cd t = a[i]; a[i] = a[j]; a[j] = t; } }
    for (len = 2; len <= m; len <<= 1) {
        int half = len >> 1, step = m / len, k;
        for (i = 0; i < m; i += len)
            for (k = 0; k < half; k++) {
                oracle_cd w = p->tw[k * step], u = a[i + k], x = a[i + k + half], v;
                if (inverse) w.im = -w.im;
                v.re = x.re * w.re - x.im * w.im; v.im = x.re * w.im + x.im * w.re;
                a[i + k].re = u.re + v.re; a[i + k].im = u.im + v.im;
                a[i + k + half].re = u.re - v.re; a[i + k + half].im = u.im - v.im;
            }
    }
}

static oracle_fast_plan *oracle_fast_get_plan(int n, int sign) {
    oracle_fast_plan *p;
    int j, bits = 0;
    for (p = oracle_fast_plans; p; p = p->next) if (p->n == n && p->sign == sign) return p;
    p = (oracle_fast_plan *)calloc(1, sizeof(*p));
    p->n = n; p->sign = sign; p->m = 1;
    while (p->m < 2 * n - 1) { p->m <<= 1; }
    while ((1 << bits) < p->m) bits++;
    p->chirp = (oracle_cd *)malloc(sizeof(oracle_cd) * (size_t)n);
    p->kernel_f = (oracle_cd *)calloc((size_t)p->m, sizeof(oracle_cd));
    p->tw = (oracle_cd *)malloc(sizeof(oracle_cd) * (size_t)(p->m / 2 + 1));
    p->rev = (int *)malloc(sizeof(int) * (size_t)p->m);
    p->work = (oracle_cd *)malloc(sizeof(oracle_cd) * (size_t)p->m);
    for (j = 0; j < p->m; j++) {
        int r = 0, b;
        for (b = 0; b < bits; b++) if (j & (1 << b)) r |= 1 << (bits - 1 - b);
        p->rev[j] = r;
    }
    for (j = 0; j < p->m / 2; j++) {
        long double ang = -2.0L * ORACLE_PI_L * (long double)j / (long double)p->m;
        p->tw[j].re = (double)cosl(ang); p->tw[j].im = (double)sinl(ang);
    }
    for (j = 0; j < n; j++) {
        long long j2 = ((long long)j * (long long)j) % (2LL * n);
        long double ang = sign * ORACLE_PI_L * (long double)j2 / (long double)n;
        p->chirp[j].re = (double)cosl(ang); p->chirp[j].im = (double)sinl(ang);
    }
    for (j = 0; j < n; j++) {
        p->kernel_f[j].re = p->chirp[j].re; p->kernel_f[j].im = -p->chirp[j].im;
        if (j) p->kernel_f[p->m - j] = p->kernel_f[j];
    }
    oracle_fast_fft(p, p->kernel_f, 0);
    for (j = 0; j < p->m; j++) { p->kernel_f[j].re /= p->m; p->kernel_f[j].im /= p->m; }
    p->next = oracle_fast_plans; oracle_fast_plans = p;
    return p;
}

static void oracle_fast_dft_1d(const oracle_cd *in, int istride, oracle_cd *out, int ostride, int n, int sign) {
    oracle_fast_plan *p = oracle_fast_get_plan(n, sign);
    oracle_cd *a = p->work;
    int j, m = p->m;
    for (j = 0; j < n; j++) {
        oracle_cd x = in[(size_t)j * istride], c = p->chirp[j];
        a[j].re = x.re * c.re - x.im * c.im; a[j].im = x.re * c.im + x.im * c.re;
    }
    for (j = n; j < m; j++) { a[j].re = 0.0; a[j].im = 0.0; }
    oracle_fast_fft(p, a, 0);
    for (j = 0; j < m; j++) {
        oracle_cd x = a[j], k = p->kernel_f[j];
        a[j].re = x.re * k.re - x.im * k.im; a[j].im = x.re * k.im + x.im * k.re;
    }
    oracle_fast_fft(p, a, 1);
    for (j = 0; j < n; j++) {
        oracle_cd x = a[j], c = p->chirp[j];
        out[(size_t)j * ostride].re = x.re * c.re - x.im * c.im;
        out[(size_t)j * ostride].im = x.re * c.im + x.im * c.re;
    }
}

static void oracle_fast_r2c_2d(int n0, int n1, const double *in, double *out) {
    int nh = n1 / 2 + 1, i, j;
    oracle_cd *row = (oracle_cd *)malloc(sizeof(oracle_cd) * (size_t)n1);
    oracle_cd *rowo = (oracle_cd *)malloc(sizeof(oracle_cd) * (size_t)n1);
    oracle_cd *o = (oracle_cd *)out;
    for (i = 0; i < n0; i++) {
        for (j = 0; j < n1; j++) { row[j].re = in[(size_t)i * n1 + j]; row[j].im = 0.0; }
        oracle_fast_dft_1d(row, 1, rowo, 1, n1, -1);
        for (j = 0; j < nh; j++) o[(size_t)i * nh + j] = rowo[j];
    }
    for (j = 0; j < nh; j++) oracle_fast_dft_1d(o + j, nh, o + j, nh, n0, -1);
    free(row); free(rowo);
}

static void oracle_fast_c2r_2d(int n0, int n1, const double *in, double *out) {
    int nh = n1 / 2 + 1, i, j;
    oracle_cd *work = (oracle_cd *)malloc(sizeof(oracle_cd) * (size_t)n0 * (size_t)nh);
    oracle_cd *row = (oracle_cd *)malloc(sizeof(oracle_cd) * (size_t)n1);
    oracle_cd *rowo = (oracle_cd *)malloc(sizeof(oracle_cd) * (size_t)n1);
    memcpy(work, in, sizeof(oracle_cd) * (size_t)n0 * (size_t)nh);
    for (j = 0; j < nh; j++) oracle_fast_dft_1d(work + j, nh, work + j, nh, n0, +1);
    for (i = 0; i < n0; i++) {
        row[0].re = work[(size_t)i * nh].re; row[0].im = 0.0;
        for (j = 1; j < nh; j++) {
            oracle_cd v = work[(size_t)i * nh + j];
            if (2 * j == n1) { row[j].re = v.re; row[j].im = 0.0; }
            else { row[j] = v; row[n1 - j].re = v.re; row[n1 - j].im = -v.im; }
        }
        oracle_fast_dft_1d(row, 1, rowo, 1, n1, +1);
        for (j = 0; j < n1; j++) out[(size_t)i * n1 + j] = rowo[j].re;
    }
    free(work); free(row); free(rowo);
}

/* FFTW contract, fftw_plan_dft_r2c_2d(n0, n1, in, out): in n0*n1 doubles,
 * out n0*(n1/2+1) interleaved complex doubles. */
static void oracle_dft_r2c_2d(int n0, int n1, const double *in, double *out /* [n0*nh][2] */) {
    int nh = n1 / 2 + 1, i, j;
    if (oracle_dft_mode == 3) { oracle_fast_r2c_2d(n0, n1, in, out); return; }
    oracle_cld *row = (oracle_cld *)malloc(sizeof(oracle_cld) * (size_t)n1);
    oracle_cld *rowo = (oracle_cld *)malloc(sizeof(oracle_cld) * (size_t)n1);
    oracle_cld *work = (oracle_cld *)malloc(sizeof(oracle_cld) * (size_t)n0 * (size_t)nh);
    for (i = 0; i < n0; i++) {
        for (j = 0; j < n1; j++) { row[j].re = (long double)in[(size_t)i * n1 + j]; row[j].im = 0.0L; }
        oracle_dft_1d(row, 1, rowo, 1, n1, -1);
        for (j = 0; j < nh; j++) work[(size_t)i * nh + j] = rowo[j];
    }
    for (j = 0; j < nh; j++) oracle_dft_1d(work + j, nh, work + j, nh, n0, -1);
    for (i = 0; i < n0 * nh; i++) { out[2 * (size_t)i] = (double)work[i].re; out[2 * (size_t)i + 1] = (double)work[i].im; }
    free(row); free(rowo); free(work);
}

/* FFTW contract, fftw_plan_dft_c2r_2d(n0, n1, in, out). */
static void oracle_dft_c2r_2d(int n0, int n1, const double *in /* [n0*nh][2] */, double *out) {
    int nh = n1 / 2 + 1, i, j;
    if (oracle_dft_mode == 3) { oracle_fast_c2r_2d(n0, n1, in, out); return; }
    oracle_cld *work = (oracle_cld *)malloc(sizeof(oracle_cld) * (size_t)n0 * (size_t)nh);
    oracle_cld *row = (oracle_cld *)malloc(sizeof(oracle_cld) * (size_t)n1);
    oracle_cld *rowo = (oracle_cld *)malloc(sizeof(oracle_cld) * (size_t)n1);
    for (i = 0; i < n0 * nh; i++) { work[i].re = (long double)in[2 * (size_t)i]; work[i].im = (long double)in[2 * (size_t)i + 1]; }
    for (j = 0; j < nh; j++) oracle_dft_1d(work + j, nh, work + j, nh, n0, +1);
    for (i = 0; i < n0; i++) {
        /* Hermitian completion; the self-conjugate bins contribute their real part only */
        row[0].re = work[(size_t)i * nh].re; row[0].im = 0.0L;
        for (j = 1; j < nh; j++) {
            oracle_cld v = work[(size_t)i * nh + j];
            if (2 * j == n1) { row[j].re = v.re; row[j].im = 0.0L; }
            else { row[j] = v; row[n1 - j].re = v.re; row[n1 - j].im = -v.im; }
        }
        oracle_dft_1d(row, 1, rowo, 1, n1, +1);
        for (j = 0; j < n1; j++) out[(size_t)i * n1 + j] = (double)rowo[j].re;
    }
    free(work); free(row); free(rowo);
}

#endif /* ORACLE_DFT_LD_H */
