// host_system.hpp — host-side term tables behind `ocm_system` (pure C++, no CUDA).
//
// What OpenMM::System + its Force objects are in the reference (SystemGenerator.cpp:34-111):
// masses, box, and one record per force-field generator.  Instead of expression strings the
// records are typed parameter tables; ocm_context_create compiles them into device tables.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "../../include/ocm_b200.h"

namespace ocm {

enum BondedKind : int {
    B_HARMONIC_BOND = 0, B_GAUSSIAN_BOND, B_GO_CONTACT, B_3SPN2_BOND,
    B_HARMONIC_ANGLE, B_FLEX_ANGLE, B_GAUSSIAN_DIHEDRAL, B_COSINE_DIHEDRAL, B_FLEX_DIHEDRAL, B_3SPN2_STACK,
    B_EXTERNAL,   // one-bead external terms; p[0] selects the form (ExternalForm), kExternalNpar parameters per term
    B_NUM_KINDS
};

// forms of B_EXTERNAL and their parameter rows (p[0] = form):
//   X_POSITION_RESTRAINT  x0 y0 z0 k v0          X_PULLING  fx fy fz
//   X_RECTANGULAR_BOX     eps lo[3] hi[3] radius X_CYLINDRICAL_RESTRAINT  axis[3] shift[3] k v0
enum ExternalForm : int { X_POSITION_RESTRAINT = 0, X_PULLING, X_RECTANGULAR_BOX, X_CYLINDRICAL_RESTRAINT };
constexpr int kExternalNpar = 10;

inline int bonded_nbody(int kind) {
    if (kind == B_EXTERNAL) return 1;
    if (kind <= B_3SPN2_BOND) return 2;
    if (kind <= B_FLEX_ANGLE || kind == B_3SPN2_STACK) return 3;
    return 4;
}

struct BondedTerm {
    int kind, group, periodic, nbody, npar;
    std::vector<int32_t> idx;  // nbody * n
    std::vector<double> p;     // npar * n
    double tab[24];            // FLA: y[10], y2[10], tmin, tmax, emin;  FLD: coef7, emin
    int64_t n() const { return (int64_t)idx.size() / nbody; }
};

enum PairKind : int { P_EXV = 0, P_DH, P_LJREP, P_LJATT, P_3SPN2EXV, P_WCA, P_ISOLF, P_TRIG };

struct PairTerm {
    int kind, group, periodic;
    double rc, rsw;
    double g0, g1, g2;
    std::vector<double> p0, p1, p2;         // per bead, NaN = non participant (p2: attractive wells only)
    std::vector<int32_t> part;              // ascending participants
    std::vector<std::pair<int32_t, int32_t>> excl;  // sorted unique (i<j)
    std::vector<int32_t> cls;               // interaction class per bead (-1 none)
    std::vector<uint32_t> allow;            // per class: bit mask of allowed partner classes
    int ncls;
};

// donor x acceptor ("CustomHbondForce"-like) term
enum DAKind : int { DA_BASE_PAIR = 0, DA_CROSS_STACK = 1, DA_PDNS = 2, DA_COMB_GO = 3 };

struct DATerm {
    int kind, group, periodic, npar;
    double cutoff;
    double g0, g1, g2;                  // BP: alpha, K_BP; CS: K_BP, K_CS, alpha; PDNS: sigma, delta
    std::vector<int32_t> donors, acceptors;  // 3 per entry, -1 = unused
    std::vector<double> p;              // npar per donor (BP, PDNS) or per acceptor (CS); none for DA_COMB_GO (g0 = k, g1 = r0)
    std::vector<uint64_t> excl;         // sorted keys donor * n_acceptors + acceptor
    int64_t nd() const { return (int64_t)donors.size() / 3; }
    int64_t na() const { return (int64_t)acceptors.size() / 3; }
};

// harmonic spring between the mass-weighted centres of two bead groups ("CustomCentroidBondForce")
struct ComTerm {
    int group, periodic;
    double k, v0;
    std::vector<int32_t> g1, g2;
};

}  // namespace ocm

struct ocm_system {
    int n = 0;
    std::vector<double> mass;
    bool has_box = false;
    double box[3] = {0, 0, 0};
    std::vector<ocm::BondedTerm> bonded;
    std::vector<ocm::PairTerm> pairs;
    std::vector<ocm::DATerm> das;
    std::vector<ocm::ComTerm> coms;
};

namespace ocm {

extern thread_local std::string g_err;
inline int fail(const std::string& m) { g_err = m; return 1; }

// natural cubic spline second derivatives on equally spaced knots (OpenMM Continuous1DFunction
// builds the same natural spline).  Thomas algorithm on the interior knots.
inline void natural_spline_second_derivatives(const double* y, int n, double h, double* y2) {
    for (int i = 0; i < n; i++) y2[i] = 0.0;
    int m = n - 2;
    if (m < 1) return;
    std::vector<double> cp(m), dp(m);
    for (int i = 0; i < m; i++) {
        double rhs = 6.0 * (y[i + 2] - 2.0 * y[i + 1] + y[i]) / h;
        double a = (i > 0) ? h : 0.0, b = 4.0 * h, c = (i < m - 1) ? h : 0.0;
        double den = b - a * (i > 0 ? cp[i - 1] : 0.0);
        cp[i] = c / den;
        dp[i] = (rhs - a * (i > 0 ? dp[i - 1] : 0.0)) / den;
    }
    for (int i = m - 1; i >= 0; i--) y2[i + 1] = dp[i] - (i < m - 1 ? cp[i] * y2[i + 2] : 0.0);
}

inline int check_indices(const ocm_system* s, const int32_t* idx, int64_t cnt, const char* what) {
    for (int64_t k = 0; k < cnt; k++)
        if (idx[k] < 0 || idx[k] >= s->n) return fail(std::string(what) + ": particle index out of range");
    return 0;
}

inline int add_bonded(ocm_system* s, int kind, int group, int periodic, int64_t n, const int32_t* idx,
                      std::vector<const double*> cols, const char* what) {
    if (!s) return fail("null system");
    if (n < 0 || (n > 0 && !idx)) return fail(std::string(what) + ": bad arguments");
    if (group < 0 || group > 31) return fail("force group must be in 0..31");
    if (periodic && !s->has_box) return fail(std::string(what) + ": periodic term without a box");
    int nbody = bonded_nbody(kind);
    if (check_indices(s, idx, n * nbody, what)) return 1;
    BondedTerm t;
    t.kind = kind; t.group = group; t.periodic = periodic ? 1 : 0; t.nbody = nbody;
    t.npar = (int)cols.size();
    t.idx.assign(idx, idx + n * nbody);
    t.p.resize((size_t)n * t.npar);
    for (int q = 0; q < t.npar; q++) {
        if (n > 0 && !cols[q]) return fail(std::string(what) + ": null parameter array");
        for (int64_t m = 0; m < n; m++) t.p[(size_t)m * t.npar + q] = cols[q][m];
    }
    std::fill(t.tab, t.tab + 24, 0.0);
    s->bonded.push_back(std::move(t));
    return 0;
}

// Exclusions + interaction groups -> per-bead class labels and an allowed-class matrix.
// Beads with identical membership over all sets share a class; a class pair interacts if some
// group has one class in set1 and the other in set2 (CustomNonbondedForce::addInteractionGroup).
inline int add_pair(ocm_system* s, int kind, int group, double rc, double rsw, double g0, double g1, double g2,
                    const double* p0, const double* p1, const ocm_pair_topology* topo, const char* what, const double* p2 = nullptr) {
    if (!s) return fail("null system");
    if (group < 0 || group > 31) return fail("force group must be in 0..31");
    if (!topo || !p0) return fail(std::string(what) + ": null argument");
    if (!(rc > 0)) return fail(std::string(what) + ": cutoff must be positive");
    if (topo->periodic && !s->has_box) return fail(std::string(what) + ": periodic term without a box");
    if (!topo->periodic && s->has_box)
        return fail(std::string(what) + ": non-periodic pair term in a periodic system is not supported");
    if (topo->periodic)
        for (int k = 0; k < 3; k++)
            if (rc > 0.5 * s->box[k]) return fail(std::string(what) + ": cutoff exceeds half the box");
    PairTerm t;
    t.kind = kind; t.group = group; t.periodic = topo->periodic ? 1 : 0; t.rc = rc; t.rsw = rsw;
    t.g0 = g0; t.g1 = g1; t.g2 = g2;
    const int n = s->n;
    t.p0.assign(p0, p0 + n);
    if (p1) t.p1.assign(p1, p1 + n); else t.p1 = t.p0;
    if (p2) t.p2.assign(p2, p2 + n); else t.p2 = t.p0;
    for (int i = 0; i < n; i++)
        if (std::isnan(t.p0[i]) != std::isnan(t.p1[i]) || std::isnan(t.p0[i]) != std::isnan(t.p2[i]))
            return fail(std::string(what) + ": incomplete parameter set for a particle");
    if (topo->n_excl < 0 || (topo->n_excl > 0 && !topo->excl_ij)) return fail(std::string(what) + ": bad exclusion list");
    if (check_indices(s, topo->excl_ij, topo->n_excl * 2, what)) return 1;
    t.excl.reserve(topo->n_excl);
    for (int64_t e = 0; e < topo->n_excl; e++) {
        int a = topo->excl_ij[2 * e], b = topo->excl_ij[2 * e + 1];
        if (a == b) continue;
        t.excl.emplace_back(std::min(a, b), std::max(a, b));
    }
    std::sort(t.excl.begin(), t.excl.end());
    t.excl.erase(std::unique(t.excl.begin(), t.excl.end()), t.excl.end());
    t.ncls = 0;
    t.cls.assign(n, -1);
    if (topo->n_igroups > 0) {
        const int ng = topo->n_igroups;
        if (ng > 64) return fail(std::string(what) + ": more than 64 interaction groups");
        // membership signature: 2 bits per group (in set1 / in set2)
        std::vector<std::pair<uint64_t, uint64_t>> sig(n, {0, 0});
        for (int g = 0; g < ng; g++) {
            for (int64_t k = topo->set1_off[g]; k < topo->set1_off[g + 1]; k++) {
                int b = topo->set1[k];
                if (b < 0 || b >= n) return fail(std::string(what) + ": interaction-group index out of range");
                sig[b].first |= (uint64_t)1 << g;
            }
            for (int64_t k = topo->set2_off[g]; k < topo->set2_off[g + 1]; k++) {
                int b = topo->set2[k];
                if (b < 0 || b >= n) return fail(std::string(what) + ": interaction-group index out of range");
                sig[b].second |= (uint64_t)1 << g;
            }
        }
        std::map<std::pair<uint64_t, uint64_t>, int> ids;
        std::vector<std::pair<uint64_t, uint64_t>> reps;
        for (int i = 0; i < n; i++) {
            if (!sig[i].first && !sig[i].second) continue;
            if (std::isnan(t.p0[i]))
                return fail(std::string(what) + ": interaction group member without parameters");
            auto it = ids.find(sig[i]);
            if (it == ids.end()) { it = ids.emplace(sig[i], (int)reps.size()).first; reps.push_back(sig[i]); }
            t.cls[i] = it->second;
        }
        t.ncls = (int)reps.size();
        if (t.ncls > 32) return fail(std::string(what) + ": more than 32 distinct interaction classes");
        t.allow.assign(t.ncls, 0u);
        for (int a = 0; a < t.ncls; a++)
            for (int b = 0; b < t.ncls; b++) {
                uint64_t hit = (reps[a].first & reps[b].second) | (reps[a].second & reps[b].first);
                int cnt = __builtin_popcountll(hit);
                if (cnt > 1)
                    return fail(std::string(what) + ": overlapping interaction groups (a pair would be counted twice) are not supported");
                if (cnt) t.allow[a] |= 1u << b;
            }
        for (int i = 0; i < n; i++) if (t.cls[i] >= 0) t.part.push_back(i);
    } else {
        for (int i = 0; i < n; i++) {
            if (std::isnan(t.p0[i]))
                return fail(std::string(what) + ": a particle has no parameter but no interaction group was given");
            t.part.push_back(i);
        }
    }
    s->pairs.push_back(std::move(t));
    return 0;
}

inline int add_da(ocm_system* s, int kind, int group, const ocm_da_topology* topo, bool per_donor,
                  std::vector<const double*> cols, double g0, double g1, double g2, const char* what) {
    if (!s || !topo) return fail(std::string(what) + ": null argument");
    if (group < 0 || group > 31) return fail("force group must be in 0..31");
    if (topo->periodic && !s->has_box) return fail(std::string(what) + ": periodic term without a box");
    if (!topo->periodic && s->has_box) return fail(std::string(what) + ": non-periodic donor-acceptor term in a periodic system is not supported");
    if (!(topo->cutoff > 0)) return fail(std::string(what) + ": cutoff must be positive");
    if (topo->periodic)
        for (int k = 0; k < 3; k++)
            if (topo->cutoff > 0.5 * s->box[k]) return fail(std::string(what) + ": cutoff exceeds half the box");
    if (topo->n_donors < 0 || topo->n_acceptors < 0 || topo->n_excl < 0) return fail(std::string(what) + ": negative count");
    if ((topo->n_donors > 0 && !topo->donors) || (topo->n_acceptors > 0 && !topo->acceptors) || (topo->n_excl > 0 && !topo->excl_da))
        return fail(std::string(what) + ": null index array");
    if (topo->n_donors >= (1 << 24) || topo->n_acceptors >= (1 << 24)) return fail(std::string(what) + ": too many donors or acceptors");
    DATerm t;
    t.kind = kind; t.group = group; t.periodic = topo->periodic ? 1 : 0; t.cutoff = topo->cutoff;
    t.npar = (int)cols.size(); t.g0 = g0; t.g1 = g1; t.g2 = g2;
    t.donors.assign(topo->donors, topo->donors + 3 * topo->n_donors);
    t.acceptors.assign(topo->acceptors, topo->acceptors + 3 * topo->n_acceptors);
    // which slots the expression of each kind reads (d1,d2,d3 / a1,a2,a3); d1 and a1 always (cutoff)
    static const bool need_d[4][3] = {{true, true, false}, {true, true, false}, {true, true, true}, {true, false, false}};
    static const bool need_a[4][3] = {{true, true, false}, {true, true, true}, {true, true, false}, {true, false, false}};
    for (int64_t d = 0; d < topo->n_donors; d++)
        for (int q = 0; q < 3; q++) {
            int32_t v = t.donors[3 * d + q];
            if (v >= s->n || v < -1 || (v < 0 && need_d[kind][q])) return fail(std::string(what) + ": donor particle index out of range");
        }
    for (int64_t a = 0; a < topo->n_acceptors; a++)
        for (int q = 0; q < 3; q++) {
            int32_t v = t.acceptors[3 * a + q];
            if (v >= s->n || v < -1 || (v < 0 && need_a[kind][q])) return fail(std::string(what) + ": acceptor particle index out of range");
        }
    for (int64_t e = 0; e < topo->n_excl; e++) {
        int64_t d = topo->excl_da[2 * e], a = topo->excl_da[2 * e + 1];
        if (d < 0 || d >= topo->n_donors || a < 0 || a >= topo->n_acceptors) return fail(std::string(what) + ": exclusion index out of range");
        t.excl.push_back((uint64_t)d * (uint64_t)topo->n_acceptors + (uint64_t)a);
    }
    std::sort(t.excl.begin(), t.excl.end());
    t.excl.erase(std::unique(t.excl.begin(), t.excl.end()), t.excl.end());
    const int64_t cnt = per_donor ? topo->n_donors : topo->n_acceptors;
    t.p.resize((size_t)cnt * t.npar);
    for (int q = 0; q < t.npar; q++) {
        if (cnt > 0 && !cols[q]) return fail(std::string(what) + ": null parameter array");
        for (int64_t m = 0; m < cnt; m++) t.p[(size_t)m * t.npar + q] = cols[q][m];
    }
    s->das.push_back(std::move(t));
    return 0;
}

// effective interaction range of one pair term: beyond it the term is exactly zero.
// (LJ-repulsive and WCA vanish beyond 2^(1/6) sigma_ij, the 3SPN.2 excluded volume beyond sigma_ij.)
inline double effective_rc(const PairTerm& t) {
    double rc = t.rc;
    if (t.kind == P_LJREP || t.kind == P_WCA || t.kind == P_3SPN2EXV) {
        double m1 = 0, m2 = 0;
        for (int i : t.part) {
            double s = t.p1[i];
            if (s >= m1) { m2 = m1; m1 = s; } else if (s > m2) m2 = s;
        }
        if (t.part.size() < 2) m2 = m1;
        double smax = 0.5 * (m1 + m2);
        double zone = (t.kind != P_3SPN2EXV) ? smax * std::pow(2.0, 1.0 / 6.0) : smax;
        zone *= (1.0 + 1e-6);
        rc = std::min(rc, zone);
    }
    return rc;
}

}  // namespace ocm
