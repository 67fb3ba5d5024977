// CPU check of opencafemol_b200/csrc/replica_math.hpp (the FP32 bonded terms of the whole-step replica kernel):
//   1. the double instantiation against central finite differences of its own energy (F = -grad E), every kind;
//   2. the float instantiation against the double one on the same geometry (differences taken in double, as the kernel does).
// Prints "max_fd_err <x> max_f32_err <y>" (errors relative to the largest force of the term) and exits 0; the pytest
// wrapper (tests/test_replica_math.py) applies the bars.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../opencafemol_b200/csrc/replica_math.hpp"

using namespace ocm::repmath;

static std::mt19937_64 rng(12345);
static double uni(double a, double b) { return a + (b - a) * std::uniform_real_distribution<double>(0.0, 1.0)(rng); }

static int nbody(int kind) { return kind <= 3 ? 2 : (kind <= 5 ? 3 : 4); }

// energy and forces of one term on 2-4 points (x: 3 * nbody doubles); T = arithmetic type after the differences
template <typename T>
static double term(int kind, const double* pd, const double* tabd, const double* x, double* f) {
    T p[8], tab[24];
    for (int k = 0; k < 8; k++) p[k] = (T)pd[k];
    for (int k = 0; k < 24; k++) tab[k] = (T)tabd[k];
    auto diff = [&](int i, int j) { return V3<T>{(T)(x[3 * i] - x[3 * j]), (T)(x[3 * i + 1] - x[3 * j + 1]), (T)(x[3 * i + 2] - x[3 * j + 2])}; };
    const int nb = nbody(kind);
    for (int k = 0; k < 3 * nb; k++) f[k] = 0;
    if (nb == 2) {
        const double dx = x[0] - x[3], dy = x[1] - x[4], dz = x[2] - x[5];
        const double r = std::sqrt(dx * dx + dy * dy + dz * dz);
        const double r0 = kind == 0 ? pd[0] : (kind == 1 ? pd[1] : (kind == 3 ? pd[2] : 0.0));
        V3<T> fi;
        const T E = bond<T>(kind, p, tab, diff(0, 1), (T)r, (T)(r - r0), fi);
        f[0] = fi.x; f[1] = fi.y; f[2] = fi.z; f[3] = -fi.x; f[4] = -fi.y; f[5] = -fi.z;
        return E;
    }
    if (nb == 3) {
        V3<T> fi, fk;
        const T E = angle<T>(kind, p, tab, diff(0, 1), diff(2, 1), fi, fk);
        f[0] = fi.x; f[1] = fi.y; f[2] = fi.z; f[6] = fk.x; f[7] = fk.y; f[8] = fk.z;
        f[3] = -fi.x - fk.x; f[4] = -fi.y - fk.y; f[5] = -fi.z - fk.z;
        return E;
    }
    V3<T> f1, f2, f3, f4;
    const T E = dihedral<T>(kind, p, tab, diff(0, 1), diff(1, 2), diff(3, 2), f1, f2, f3, f4);
    const V3<T> ff[4] = {f1, f2, f3, f4};
    for (int b = 0; b < 4; b++) { f[3 * b] = ff[b].x; f[3 * b + 1] = ff[b].y; f[3 * b + 2] = ff[b].z; }
    return E;
}

int main() {
    double max_fd = 0, max_f32 = 0;
    double kind_abs[9] = {0}, kind_rel[9] = {0};
    for (int kind = 0; kind <= 8; kind++) {
        for (int trial = 0; trial < 400; trial++) {
            double p[8] = {0}, tab[24] = {0};
            const int nb = nbody(kind);
            double x[12];
            // a chain with bond lengths around 0.38 nm, placed 0-8 nm from the origin (what a bead of a CG protein sees)
            const double off[3] = {uni(-8, 8), uni(-8, 8), uni(-8, 8)};
            for (int b = 0; b < nb; b++)
                for (int q = 0; q < 3; q++) x[3 * b + q] = off[q] + (b ? x[3 * (b - 1) + q] - off[q] : 0.0) + uni(-0.35, 0.35);
            switch (kind) {
                case 0: p[0] = 0.38 + uni(-0.02, 0.02); p[1] = uni(1e4, 6e4); break;
                case 1: p[0] = -uni(1, 8); p[1] = 0.55 + uni(-0.1, 0.1); p[2] = uni(0.01, 0.03); break;
                case 2: p[0] = uni(0.5, 5); p[1] = uni(0.3, 0.7); p[2] = -uni(10, 100); p[3] = uni(10, 100); p[4] = (trial & 1) ? uni(0.6, 0.9) : 0.0; break;
                case 3: p[0] = uni(100, 300); p[1] = uni(1e4, 3e4); p[2] = 0.4 + uni(-0.05, 0.05); break;
                case 4: p[0] = uni(1.2, 2.6); p[1] = uni(50, 500); break;
                case 5: {
                    p[0] = uni(1, 5);
                    for (int k = 0; k < 10; k++) tab[k] = uni(0, 8);
                    tab[20] = 1.3; tab[21] = 2.9; tab[22] = uni(0, 1);
                    // natural spline second derivatives (the engine computes them with the same recurrence)
                    const double h = (tab[21] - tab[20]) / 9.0;
                    double cp[8], dp[8];
                    for (int i = 0; i < 8; i++) {
                        const double rhs = 6.0 * (tab[i + 2] - 2.0 * tab[i + 1] + tab[i]) / h;
                        const double a = i > 0 ? h : 0.0, b = 4.0 * h, c = i < 7 ? h : 0.0;
                        const double den = b - a * (i > 0 ? cp[i - 1] : 0.0);
                        cp[i] = c / den; dp[i] = (rhs - a * (i > 0 ? dp[i - 1] : 0.0)) / den;
                    }
                    for (int i = 7; i >= 0; i--) tab[10 + i + 1] = dp[i] - (i < 7 ? cp[i] * tab[10 + i + 2] : 0.0);
                } break;
                case 6: p[0] = -uni(1, 6); p[1] = uni(-3.1, 3.1); p[2] = uni(0.15, 0.4); break;
                case 7: p[0] = uni(1, 6); p[1] = uni(-3.1, 3.1); p[2] = (double)(1 + trial % 3); break;
                default: p[0] = uni(1, 5); for (int k = 0; k < 8; k++) tab[k] = uni(-3, 3); break;
            }
            double fd[12], ff[12], fnum[12];
            term<double>(kind, p, tab, x, fd);
            term<float>(kind, p, tab, x, ff);
            double fmax = 0;
            for (int k = 0; k < 3 * nb; k++) fmax = std::fmax(fmax, std::fabs(fd[k]));
            if (!(fmax > 1e-3)) continue;
            // the steep part of a 12-10 contact or a gaussian tail is not where a finite difference resolves 1e-6: central
            // differences with h = 1e-6 nm are good to ~1e-7 relative for smooth terms
            const double h = 1e-6;
            for (int k = 0; k < 3 * nb; k++) {
                double xp[12], xm[12], tmp[12];
                for (int q = 0; q < 3 * nb; q++) { xp[q] = x[q]; xm[q] = x[q]; }
                xp[k] += h; xm[k] -= h;
                fnum[k] = -(term<double>(kind, p, tab, xp, tmp) - term<double>(kind, p, tab, xm, tmp)) / (2 * h);
            }
            for (int k = 0; k < 3 * nb; k++) {
                const double e_fd = std::fabs(fnum[k] - fd[k]) / fmax, e_32 = std::fabs(ff[k] - fd[k]) / fmax;
                if (e_fd > max_fd) max_fd = e_fd;
                if (e_32 > max_f32) max_f32 = e_32;
                kind_abs[kind] = std::fmax(kind_abs[kind], std::fabs(ff[k] - fd[k]));
                kind_rel[kind] = std::fmax(kind_rel[kind], e_32);
                if (e_fd > 1e-4)
                    std::printf("kind %d trial %d comp %d: fd %.9g analytic %.9g f32 %.9g (fmax %.3g)\n", kind, trial, k, fnum[k], fd[k], ff[k], fmax);
            }
        }
    }
    for (int kind = 0; kind <= 8; kind++) std::printf("kind %d: f32 - f64 max abs %.3e kJ/mol/nm, max rel %.3e\n", kind, kind_abs[kind], kind_rel[kind]);
    std::printf("max_fd_err %.3e max_f32_err %.3e\n", max_fd, max_f32);
    return 0;
}
