/* oracle/shims/fftw3.h — TEST INFRASTRUCTURE ONLY.
 *
 * Stand-in for the eight FFTW3 symbols the reference uses at
 * /root/reference/src/main.cpp:964-1048 so that the UNMODIFIED reference
 * translation unit compiles in an image without libfftw3.  The arithmetic is
 * oracle/dft_ld.h (long-double DFT obeying FFTW's documented contract).
 *
 * One deliberate choice: fftw_malloc returns ZEROED memory.  The reference never
 * writes row Nx/2 of `phik` (main.cpp:999,1015 skip it) and reads it in the c2r;
 * with the real FFTW that row is whatever the allocator returns (in practice a
 * fresh zero page).  calloc makes that de-facto behaviour defined.
 */
#ifndef PICSP_ORACLE_FFTW3_SHIM_H
#define PICSP_ORACLE_FFTW3_SHIM_H

#include <cstdlib>
#include <cstddef>
#include "../dft_ld.h"

typedef double fftw_complex[2];

struct picsp_shim_fftw_plan {
    int kind;            /* 0 = r2c_2d, 1 = c2r_2d */
    int n0, n1;
    double *real;
    fftw_complex *cplx;
};
typedef picsp_shim_fftw_plan *fftw_plan;

#define FFTW_ESTIMATE (1U << 6)

static inline void *fftw_malloc(size_t n) { return calloc(1, n); }
static inline void fftw_free(void *p) { free(p); }

static inline fftw_plan fftw_plan_dft_r2c_2d(int n0, int n1, double *in, fftw_complex *out, unsigned) {
    fftw_plan p = new picsp_shim_fftw_plan;
    p->kind = 0; p->n0 = n0; p->n1 = n1; p->real = in; p->cplx = out;
    return p;
}
static inline fftw_plan fftw_plan_dft_c2r_2d(int n0, int n1, fftw_complex *in, double *out, unsigned) {
    fftw_plan p = new picsp_shim_fftw_plan;
    p->kind = 1; p->n0 = n0; p->n1 = n1; p->real = out; p->cplx = in;
    return p;
}
static inline void fftw_execute(const fftw_plan p) {
    if (p->kind == 0) oracle_dft_r2c_2d(p->n0, p->n1, p->real, &p->cplx[0][0]);
    else              oracle_dft_c2r_2d(p->n0, p->n1, &p->cplx[0][0], p->real);
}
static inline void fftw_destroy_plan(fftw_plan p) { delete p; }
static inline void fftw_cleanup(void) {}

#endif
