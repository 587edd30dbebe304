/* oracle/walls_check.c — TEST INFRASTRUCTURE ONLY.  *** NOT THE REFERENCE. ***
 *
 * CPU restatement of the PICSP_FLAG_WALLS extension of picsp_b200 (BASELINE.json config 3, "bounded domain with
 * wall boundaries and Gauss-Seidel solver").  The reference (sayanadhikari/picsp) has NO bounded-domain semantics:
 * its code is periodic-only; absorbing walls and a Dirichlet solver exist only as commented-out sketches
 * (/root/reference/src/main.cpp:826-843 and :1064-1108).  The semantics below are therefore this repository's own
 * definition, and this file is the repository's own checker for its CUDA implementation — results that agree with
 * it carry the label "no reference oracle".
 *
 *   deposit   CIC weights exactly as the reference's scatter (main.cpp:655-668), accumulated into den (or into a
 *             cleared den), NO periodic fold; absorbed particles (position NaN) are skipped
 *   rho       q_i*den_i + q_e*den_e on interior nodes, 0 on the walls
 *   solve     red-black Gauss-Seidel/SOR, phi = 0 on the walls, warm start:
 *                 g = 0.25*((phi[i-1][j] + phi[i+1][j]) + (phi[i][j-1] + phi[i][j+1]) + dx^2*rho[i][j])
 *                 phi[i][j] += omega*(g - phi[i][j]),   (i+j) even first, then odd;
 *             every `batch` sweeps L2 = sqrt(sum_interior (g - phi)^2)/(nix*niy); stop when L2 < tol
 *   E         central differences inside, full one-sided differences on the walls
 *   push      the reference's kick + drift with its gather (main.cpp:671-681, 779-799); a particle that leaves
 *             [0, xl) x [0, yl) is absorbed: position NaN, velocity 0
 * Compile with -ffp-contract=off: the CUDA kernels round every operation explicitly in the same order, so a given
 * number of sweeps gives bit-identical phi.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

void walls_deposit(int nix, int niy, double dx, double *den, const double *x, const double *y, long n, double spwt, int clear) {
    double dxdy = dx * dx;
    long p;
    if (clear) memset(den, 0, sizeof(double) * (size_t)nix * niy);
    for (p = 0; p < n; p++) {
        double lx, ly, di, dj;
        int i, j;
        if (x[p] != x[p]) continue;                 /* absorbed */
        lx = (x[p] - 0.0) / dx; ly = (y[p] - 0.0) / dx;
        i = (int)lx; j = (int)ly;
        di = lx - i; dj = ly - j;
        den[i * niy + j]           += spwt * (1 - di) * (1 - dj) / dxdy;
        den[(i + 1) * niy + j]     += spwt * (di) * (1 - dj) / dxdy;
        den[i * niy + j + 1]       += spwt * (1 - di) * (dj) / dxdy;
        den[(i + 1) * niy + j + 1] += spwt * (di) * (dj) / dxdy;
    }
}

void walls_rho(int nix, int niy, double *rho, const double *den_i, const double *den_e, double q_i, double q_e) {
    int i, j;
    for (i = 0; i < nix; i++)
        for (j = 0; j < niy; j++)
            rho[i * niy + j] = (i > 0 && i < nix - 1 && j > 0 && j < niy - 1) ? q_i * den_i[i * niy + j] + q_e * den_e[i * niy + j] : 0.0;
}

double walls_omega(int ncx, int ncy) { return 2.0 / (1.0 + sin(3.14159265358979323846 / (double)(ncx > ncy ? ncx : ncy))); }

static double walls_residual(int nix, int niy, double dx2, const double *phi, const double *rho) {
    double sum = 0.0;
    int i, j;
    for (i = 1; i < nix - 1; i++)
        for (j = 1; j < niy - 1; j++) {
            long c = (long)i * niy + j;
            double a = phi[c - niy] + phi[c + niy], b = phi[c - 1] + phi[c + 1];
            double R = 0.25 * ((a + b) + dx2 * rho[c]) - phi[c];
            sum = sum + R * R;
        }
    return sqrt(sum) / ((double)nix * (double)niy);
}

/* fixed_sweeps > 0: run exactly that many sweeps (to compare a CUDA iterate bit for bit); otherwise iterate to tol.
 * Returns the sweeps done (negative: cap hit); *l2_out = last residual evaluated (or the final one). */
long walls_rb_sor(int nix, int niy, double dx, double *phi, const double *rho, double omega, double tol, long max_sweeps,
                  int batch, long fixed_sweeps, double *l2_out) {
    double dx2 = dx * dx, L2 = 1e300;
    long sweep = 0;
    int i, j, colour;
    for (j = 0; j < niy; j++) { phi[j] = 0.0; phi[(long)(nix - 1) * niy + j] = 0.0; }
    for (i = 0; i < nix; i++) { phi[(long)i * niy] = 0.0; phi[(long)i * niy + niy - 1] = 0.0; }
    if (fixed_sweeps > 0) max_sweeps = fixed_sweeps;
    while (sweep < max_sweeps) {
        for (colour = 0; colour < 2; colour++)
            for (i = 1; i < nix - 1; i++)
                for (j = 1 + ((i + 1 + colour) & 1); j < niy - 1; j += 2) {
                    long c = (long)i * niy + j;
                    double a = phi[c - niy] + phi[c + niy], b = phi[c - 1] + phi[c + 1];
                    double g = 0.25 * ((a + b) + dx2 * rho[c]);
                    double old = phi[c];
                    phi[c] = old + omega * (g - old);
                }
        sweep++;
        if (fixed_sweeps <= 0 && (sweep % batch == 0 || sweep == max_sweeps)) {
            L2 = walls_residual(nix, niy, dx2, phi, rho);
            if (L2 < tol) break;
        }
    }
    if (fixed_sweeps > 0) L2 = walls_residual(nix, niy, dx2, phi, rho);
    if (l2_out) *l2_out = L2;
    return (fixed_sweeps > 0 || L2 < tol) ? sweep : -sweep;
}

void walls_ef(int nix, int niy, double dx, const double *phi, double *efx, double *efy) {
    int i, j;
    double two_dx = 2.0 * dx;
    for (i = 0; i < nix; i++)
        for (j = 0; j < niy; j++) {
            long k = (long)i * niy + j;
            if (i == 0) efx[k] = (phi[k] - phi[k + niy]) / dx;
            else if (i == nix - 1) efx[k] = (phi[k - niy] - phi[k]) / dx;
            else efx[k] = (phi[k - niy] - phi[k + niy]) / two_dx;
            if (j == 0) efy[k] = (phi[k] - phi[k + 1]) / dx;
            else if (j == niy - 1) efy[k] = (phi[k - 1] - phi[k]) / dx;
            else efy[k] = (phi[k - 1] - phi[k + 1]) / two_dx;
        }
}

static double walls_gather(int niy, const double *f, double lx, double ly) {
    int i = (int)lx, j = (int)ly;
    double di = lx - i, dj = ly - j;
    return f[i * niy + j] * (1 - di) * (1 - dj) + f[(i + 1) * niy + j] * di * (1 - dj) + f[i * niy + j + 1] * (1 - di) * dj +
           f[(i + 1) * niy + j + 1] * di * dj;
}

/* half = 1: the half-step velocity rewind (main.cpp:850-866), no move */
long walls_push(int nix, int niy, double dx, double dt, const double *efx, const double *efy, double *x, double *y,
                double *vx, double *vy, long n, double q, double m, int half) {
    double qm = q / m, xl = (nix - 1) * dx, yl = (niy - 1) * dx;
    long p, absorbed = 0;
    for (p = 0; p < n; p++) {
        double lx, ly, ex, ey;
        if (x[p] != x[p]) continue;
        lx = (x[p] - 0.0) / dx; ly = (y[p] - 0.0) / dx;
        ex = walls_gather(niy, efx, lx, ly); ey = walls_gather(niy, efy, lx, ly);
        if (half) { vx[p] -= 0.5 * dt * qm * ex; vy[p] -= 0.5 * dt * qm * ey; continue; }
        vx[p] += dt * qm * ex; vy[p] += dt * qm * ey;
        x[p] += dt * vx[p]; y[p] += dt * vy[p];
        if (!(x[p] >= 0.0 && x[p] < xl && y[p] >= 0.0 && y[p] < yl)) { x[p] = y[p] = NAN; vx[p] = vy[p] = 0.0; absorbed++; }
    }
    return absorbed;
}
