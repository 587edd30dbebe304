// Particle loader, src/main.cpp:567-640, with the reference's RNG classes (main.cpp:49-54).
#include <cmath>

#include "host.hpp"

namespace picsp_host {

void Loader::fill(const picsp_run_config &cfg, int species, double *x, double *y, double *vx, double *vy) {
    const int nix = cfg.numxCells + 1, niy = cfg.numyCells + 1;
    const double dx = cfg.stepSize, dy = cfg.stepSize;
    const double xl = (nix - 1) * dx, yl = (niy - 1) * dy;     // main.cpp:366,374
    const int num = species == 0 ? cfg.nParticlesI : cfg.nParticlesE;
    const double vth = species == 0 ? cfg.vthI : cfg.vthE;     // Species::Temp (main.cpp:407-408)
    const double xdrift = species == 0 ? cfg.driftI : cfg.driftE, ydrift = 0.0;   // init(&ions,driftI,0) (main.cpp:437-438)
    const double PI = 3.14159265359;                           // main.cpp:61
    const double delta_x = xl / num, delta_y = yl / num, theta = 2 * PI / xl;
    for (int p = 0; p < num; p++) {
        double px, py, u, v;
        if (cfg.loadType == 1) {
            // draw order x, u(3 draws), y, v(3 draws); the three rnd() of sampleVel are taken left to right
            px = 0.0 + rnd() * (nix - 1) * dx;
            double r1 = rnd(), r2 = rnd(), r3 = rnd();
            u = vth * std::sqrt(2) * (r1 + r2 + r3 - 1.5);
            py = 0.0 + rnd() * (niy - 1) * dy;
            r1 = rnd(); r2 = rnd(); r3 = rnd();
            v = vth * std::sqrt(2) * (r1 + r2 + r3 - 1.5);
        } else {
            // `double x = ... + 0.1*sin(theta*x)` (main.cpp:599) reads x in its own initialiser; with the
            // reference's build (g++ -O0, makefile:29) the slot holds the previous particle's final x
            // (0 for the first ion; carried from the last ion into the first electron).
            px = 0.0 + (p + 0.5) * delta_x + 0.1 * std::sin(theta * x_carry);
            u = xdrift * std::pow(-1, p);
            py = (p + 0.5) * delta_y;
            v = ydrift;
        }
        if (px < 0) px = px + xl;
        if (px > xl) px = px - xl;
        if (py < 0) py = py + yl;
        if (py > yl) py = py - yl;
        if (cfg.loadType != 1) x_carry = px;
        x[p] = px; y[p] = py; vx[p] = u; vy[p] = v;
    }
}

}  // namespace picsp_host
