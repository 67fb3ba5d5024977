// lbfgs.hpp — limited-memory BFGS with a backtracking Armijo line search (host side, double precision).
//
// Used by ocm_minimize, the engine's counterpart of OpenMM::LocalEnergyMinimizer::minimize, which the reference
// calls before a run when `energy_minimization = true` (src/Simulator.cpp:43-46).  The energy/gradient callback
// evaluates on the device; only the two-loop recursion and the line-search bookkeeping run here.
#ifndef OCM_LBFGS_HPP
#define OCM_LBFGS_HPP

#include <algorithm>
#include <cmath>
#include <deque>
#include <vector>

namespace ocmhost {

struct LbfgsReport {
    int iterations = 0;
    int evaluations = 0;
    double energy = 0.0;
    double rms_gradient = 0.0;
    bool converged = false;
};

// minimises f over x (in place).  eval(x, g) returns f(x) and fills the gradient g (same length as x).
// Converged when sqrt(|g|^2 / n_particles) < tolerance, n_particles = x.size() / 3 (OpenMM's criterion is an RMS
// force per particle).  max_iterations <= 0: until convergence.
template <class Eval>
LbfgsReport lbfgs_minimize(std::vector<double>& x, Eval&& eval, double tolerance, int max_iterations, int history = 10,
                           double first_step = 0.01) {
    const std::size_t N = x.size();
    const double npart = std::max<std::size_t>(1, N / 3);
    LbfgsReport rep;
    std::vector<double> g(N), xn(N), gn(N), dir(N);
    double E = eval(x, g);
    rep.evaluations++;
    std::deque<std::vector<double>> S, Y;
    std::deque<double> rho;
    const int itmax = max_iterations > 0 ? max_iterations : 1000000;
    int restarts = 0;
    for (int it = 0; it < itmax; it++) {
        double g2 = 0;
        for (double v : g) g2 += v * v;
        rep.rms_gradient = std::sqrt(g2 / npart);
        if (!(rep.rms_gradient >= tolerance)) { rep.converged = std::isfinite(rep.rms_gradient); break; }
        dir = g;
        std::vector<double> alpha(S.size());
        for (int k = (int)S.size() - 1; k >= 0; k--) {
            double sd = 0;
            for (std::size_t i = 0; i < N; i++) sd += S[k][i] * dir[i];
            alpha[k] = rho[k] * sd;
            for (std::size_t i = 0; i < N; i++) dir[i] -= alpha[k] * Y[k][i];
        }
        if (!S.empty()) {
            double sy = 0, yy = 0;
            for (std::size_t i = 0; i < N; i++) { sy += S.back()[i] * Y.back()[i]; yy += Y.back()[i] * Y.back()[i]; }
            const double gam = sy / yy;
            for (double& v : dir) v *= gam;
        } else {
            // first step (and after a restart): move the system by first_step nm in the steepest-descent direction
            const double sc = first_step / std::sqrt(std::max(g2, 1e-300));
            for (double& v : dir) v *= sc;
        }
        for (std::size_t k = 0; k < S.size(); k++) {
            double yd = 0;
            for (std::size_t i = 0; i < N; i++) yd += Y[k][i] * dir[i];
            const double beta = rho[k] * yd;
            for (std::size_t i = 0; i < N; i++) dir[i] += S[k][i] * (alpha[k] - beta);
        }
        double gd = 0;
        for (std::size_t i = 0; i < N; i++) gd -= g[i] * dir[i];
        if (!(gd < 0)) {
            if (S.empty() || ++restarts > 50) break;
            S.clear(); Y.clear(); rho.clear();
            continue;
        }
        double t = 1.0, En = E;
        bool ok = false;
        for (int ls = 0; ls < 50; ls++) {
            for (std::size_t i = 0; i < N; i++) xn[i] = x[i] - t * dir[i];
            En = eval(xn, gn);
            rep.evaluations++;
            if (std::isfinite(En) && En <= E + 1e-4 * t * gd) { ok = true; break; }
            t *= 0.5;
        }
        if (!ok) {
            if (S.empty() || ++restarts > 50) break;
            S.clear(); Y.clear(); rho.clear();
            continue;
        }
        std::vector<double> s(N), y(N);
        double sy = 0;
        for (std::size_t i = 0; i < N; i++) { s[i] = xn[i] - x[i]; y[i] = gn[i] - g[i]; sy += s[i] * y[i]; }
        if (sy > 1e-12) {
            S.push_back(std::move(s)); Y.push_back(std::move(y)); rho.push_back(1.0 / sy);
            if ((int)S.size() > history) { S.pop_front(); Y.pop_front(); rho.pop_front(); }
        }
        x.swap(xn);
        g.swap(gn);
        E = En;
        rep.iterations = it + 1;
    }
    rep.energy = E;
    return rep;
}

}  // namespace ocmhost

#endif  // OCM_LBFGS_HPP
