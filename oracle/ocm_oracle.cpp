// ocm_oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY, NOT PRODUCT CODE.
//
// A double-precision CPU restatement of the per-step hot path of yutakasi634/OpenCafeMol, i.e.
// of the OpenMM 8.0.0 semantics that the reference's force-field generators rely on (SURVEY.md
// Appendix A/B).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load this library; the product (libocm_b200.so) never does.
//
// PARITY STATUS: pinned by the reference compiled here.  The reference's own tests
// (test/test_read_toml_ff_generator.cpp) hold no numeric assertion and OpenMM (where the arithmetic lives,
// pinned 8.0.0 by .github/workflows/main.yml:25) is absent from /root/reference and cannot be built
// offline, so the pin is oracle/_ref: every reference source compiled unmodified (oracle/Makefile, target
// `ref`) against the stand-in OpenMM of oracle/ref_shim, whose Custom*Force classes evaluate the reference's
// own energy-expression strings with forward-mode differentiation.  tests/golden/make_ref_*.py (committed)
// generate tests/golden/ref_*.json.gz from it; tests/test_reference_pin.py holds this file to those vectors
// (energies 1e-10, forces 1e-9, 20-step trajectories 1e-8).  Beside that: (i) the closed-form known answers
// of the reference's sample inputs (tests/test_oracle_kat.py), (ii) finite differences
// F = -dE/dx of every term kind.  Unpinned against OpenMM itself: the Monte Carlo barostat's acceptance rule
// and OpenMM's integrator internals are restated from its published source.
//
// It exports the same C ABI as include/ocm_b200.h with the prefix `oracle_` instead of `ocm_`.
// Each potential cites the reference file:line whose expression string it restates.
//
// Build: make -C oracle   (g++ -O2 -fopenmp -ffp-contract=off)

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <set>
#include <memory>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/ocm_b200.h"

namespace {

thread_local std::string g_err;
int fail(const std::string& m) { g_err = m; return 1; }

const double kBoltz = 8.31446261815324e-3;  // kJ/mol/K  (OpenMM BOLTZ)

struct Vec3 {
    double x, y, z;
};
inline Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 operator+(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator*(Vec3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline double dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec3 cross(Vec3 a, Vec3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

enum BondedKind {
    B_HARMONIC_BOND, B_GAUSSIAN_BOND, B_GO_CONTACT, B_3SPN2_BOND,
    B_HARMONIC_ANGLE, B_FLEX_ANGLE, B_GAUSSIAN_DIHEDRAL, B_COSINE_DIHEDRAL, B_FLEX_DIHEDRAL, B_3SPN2_STACK,
    B_EXTERNAL  // one-bead "CustomExternalForce" terms, p[0] = form (ExternalForm)
};
enum ExternalForm { X_POSITION_RESTRAINT, X_PULLING, X_RECTANGULAR_BOX, X_CYLINDRICAL_RESTRAINT };

struct BondedTerm {
    BondedKind kind;
    int group, periodic;
    int nbody;                 // 2, 3 or 4
    std::vector<int32_t> idx;  // nbody * n
    std::vector<double> p;     // npar * n
    int npar;
    // FLA: natural-spline knots + second derivatives, range, e_min.  FLD: coef7 + e_min
    double tab[24];
    int64_t n() const { return (int64_t)idx.size() / nbody; }
};

enum PairKind { P_EXV, P_DH, P_LJREP, P_LJATT, P_3SPN2EXV, P_WCA, P_ISOLF, P_TRIG };

struct PairTerm {
    PairKind kind;
    int group, periodic;
    double rc, rsw;
    double g0, g1, g2;            // global parameters (meaning depends on kind)
    std::vector<double> p0, p1, p2;   // per-bead parameters (NaN = not a participant); p2: attractive wells only
    std::vector<int32_t> part;    // participants, ascending
    std::vector<std::vector<int32_t>> excl;  // per bead, sorted partner list
    std::vector<int32_t> cls;     // interaction class per bead (-1 none)
    std::vector<uint8_t> allow;   // ncls x ncls
    int ncls;
    int nper = 0;                 // replica batch: beads of different replicas (i / nper) never interact
};

// donor x acceptor ("CustomHbondForce"-like) term, SURVEY.md B.5
enum DAKind { DA_BASE_PAIR, DA_CROSS_STACK, DA_PDNS, DA_COMB_GO };
struct DATerm {
    DAKind kind;
    int group, periodic;
    double cutoff;
    std::vector<int32_t> donors, acceptors;           // 3 per entry, -1 = unused
    std::vector<std::pair<int32_t, int32_t>> excl;    // (donor index, acceptor index), sorted
    std::vector<double> p;                            // 4 or 5 per donor (BP, PDNS) / per acceptor (CS)
    int npar;
    double g0, g1, g2;                                // BP: alpha, K; CS: K_BP, K_CS, alpha; PDNS: sigma, delta
    int64_t nd() const { return (int64_t)donors.size() / 3; }
    int64_t na() const { return (int64_t)acceptors.size() / 3; }
};

// spring between the mass-weighted centres of two bead groups ("CustomCentroidBondForce", default weights)
struct ComTerm {
    int group, periodic;
    double k, v0;
    std::vector<int32_t> g1, g2;
};

}  // namespace

struct ocm_system {
    int n;
    std::vector<double> mass;
    bool has_box;
    double box[3];
    std::vector<BondedTerm> bonded;
    std::vector<PairTerm> pairs;
    std::vector<DATerm> das;
    std::vector<ComTerm> coms;
};

struct ocm_context {
    ocm_system sys;
    int nrep = 1, nper = 0;          // replica batch
    std::vector<uint64_t> rep_seeds;
    std::shared_ptr<ocm_system> base; // the single-replica system of a batch
    double T, gamma, dt;
    uint64_t seed;
    int64_t step;
    std::vector<Vec3> x, v;
    bool force_brute;
    // slab decomposition (restated as a plan only: which rank owns / sees which bead; the arithmetic is rank-free)
    int dd_rank = 0, dd_world = 1, dd_axis = -1;
    double dd_halo = 0;
    double vbox[3] = {0, 0, 0};  // the engine's virtual box of an unlimited-boundary system (canonical predicate only)
    // the CPU arm of bench.py (oracle_step on large systems): Verlet half lists with a skin, shared by the pair terms of equal
    // topology, rebuilt when a bead has moved more than half the skin; per-thread force slabs kept between calls
    struct VerletGroup {
        std::vector<int> terms;            // pair terms (same participants, exclusions, classes) evaluated over this list
        double rlist = 0;
        std::vector<int> off, nbr;         // CSR half list: partners j of participant bead i (global indices)
        std::vector<int> beads;            // row -> bead
    };
    std::vector<VerletGroup> vgroups;
    std::vector<Vec3> vl_xref;
    std::vector<std::vector<Vec3>> vl_F;   // [thread][bead]
    bool vl_valid = false, vl_enabled = true;
    double vl_skin = 0.3;
    long long vl_builds = 0;
    // Monte Carlo barostat (oracle_context_set_barostat)
    int baro_kind = 0, baro_mode[3] = {0, 0, 0}, baro_freq = 0, baro_attempted[3] = {0, 0, 0}, baro_accepted[3] = {0, 0, 0};
    double baro_pressure[3] = {0, 0, 0}, baro_tension = 0, baro_T = 0, baro_vscale[3] = {0, 0, 0};
    uint64_t baro_rng = 0;
    int64_t baro_calls = 0;               // updateContextState calls since the last trial
    double baro_trials = 0, baro_accepts = 0, baro_last[6] = {-1, 0, 0, 0, 0, 0};
    std::vector<std::vector<int>> molecules;
};

struct ocm_comm { int world; };

namespace {

// ------------------------------------------------------------------------------------------------
// geometry helpers
inline Vec3 min_image(Vec3 d, const double* box) {
    d.x -= box[0] * std::nearbyint(d.x / box[0]);
    d.y -= box[1] * std::nearbyint(d.y / box[1]);
    d.z -= box[2] * std::nearbyint(d.z / box[2]);
    return d;
}

// natural cubic spline second derivatives on equally spaced knots (OpenMM SplineFitter::
// createNaturalSpline as used by Continuous1DFunction)
void natural_spline_y2(const double* y, int n, double h, double* y2) {
    std::vector<double> a(n), b(n), c(n), r(n);
    y2[0] = y2[n - 1] = 0.0;
    if (n < 3) return;
    // tridiagonal system for interior points
    int m = n - 2;
    std::vector<double> diag(m), rhs(m), sup(m), sub(m);
    for (int i = 0; i < m; i++) {
        sub[i] = h; diag[i] = 4.0 * h; sup[i] = h;
        rhs[i] = 6.0 * ((y[i + 2] - y[i + 1]) / h - (y[i + 1] - y[i]) / h);
    }
    for (int i = 1; i < m; i++) {
        double w = sub[i] / diag[i - 1];
        diag[i] -= w * sup[i - 1];
        rhs[i] -= w * rhs[i - 1];
    }
    std::vector<double> sol(m);
    sol[m - 1] = rhs[m - 1] / diag[m - 1];
    for (int i = m - 2; i >= 0; i--) sol[i] = (rhs[i] - sup[i] * sol[i + 1]) / diag[i];
    for (int i = 0; i < m; i++) y2[i + 1] = sol[i];
}

// value and derivative of the spline at t in [tmin,tmax]
inline void spline_eval(const double* y, const double* y2, int n, double tmin, double tmax,
                        double t, double* val, double* der) {
    double h = (tmax - tmin) / (n - 1);
    int k = (int)std::floor((t - tmin) / h);
    if (k < 0) k = 0;
    if (k > n - 2) k = n - 2;
    double a = (tmin + (k + 1) * h - t) / h;
    double b = 1.0 - a;
    *val = a * y[k] + b * y[k + 1] + ((a * a * a - a) * y2[k] + (b * b * b - b) * y2[k + 1]) * h * h / 6.0;
    *der = (y[k + 1] - y[k]) / h + ((1.0 - 3.0 * a * a) * y2[k] + (3.0 * b * b - 1.0) * y2[k + 1]) * h / 6.0;
}

// ------------------------------------------------------------------------------------------------
// bonded energies: E and dE/d(coordinate) for each kind
inline void bond_like(const BondedTerm& t, int64_t m, double r, double* E, double* dEdr) {
    const double* p = &t.p[m * t.npar];
    switch (t.kind) {
        case B_HARMONIC_BOND: {  // p = r0, k     HarmonicBondForceFieldGenerator.cpp:3-14
            double d = r - p[0];
            *E = 0.5 * p[1] * d * d; *dEdr = p[1] * d;
        } break;
        case B_GAUSSIAN_BOND: {  // p = k, r0, sigma   GaussianBondForceFieldGenerator.cpp:5-9
            double d = r - p[1];
            double e = p[0] * std::exp(-d * d / (2.0 * p[2] * p[2]));
            *E = e; *dEdr = -e * d / (p[2] * p[2]);
        } break;
        case B_GO_CONTACT: {  // p = k, r0, slope, intercept, cap     GoContactForceFieldGenerator.cpp:5-7;
                              // CappedGoContactForceFieldGenerator.cpp:5-9: cap_val = step(cap - r/r0) selects the linear cap
            if (p[4] - r / p[1] >= 0.0) {
                *E = p[0] * p[2] * (r / p[1]) + p[0] * p[3];
                *dEdr = p[0] * p[2] / p[1];
            } else {
                double s = p[1] / r, s2 = s * s, s10 = s2 * s2 * s2 * s2 * s2, s12 = s10 * s2;
                *E = p[0] * (5.0 * s12 - 6.0 * s10);
                *dEdr = p[0] * 60.0 * (s10 - s12) / r;
            }
        } break;
        case B_3SPN2_BOND: {  // p = k2, k4, r0  ThreeSPN2BondForceFieldGenerator.cpp:5-10
            double d = r - p[2], d2 = d * d;
            *E = p[0] * d2 + p[1] * d2 * d2; *dEdr = 2.0 * p[0] * d + 4.0 * p[1] * d2 * d;
        } break;
        default: *E = 0; *dEdr = 0;
    }
}

inline void angle_like(const BondedTerm& t, int64_t m, double th, double* E, double* dEdth) {
    const double* p = &t.p[m * t.npar];
    if (t.kind == B_HARMONIC_ANGLE) {  // p = theta0, k   HarmonicAngleForceFieldGenerator.cpp:5-12
        double d = th - p[0];
        *E = 0.5 * p[1] * d * d; *dEdth = p[1] * d;
    } else {  // B_FLEX_ANGLE, p = k.   FlexibleLocalAngleForceFieldGenerator.cpp:21-25
        const double* y = t.tab; const double* y2 = t.tab + 10;
        double tmin = t.tab[20], tmax = t.tab[21], emin = t.tab[22];
        double val = 0, der = 0;
        if (th >= tmin && th <= tmax) spline_eval(y, y2, 10, tmin, tmax, th, &val, &der);
        double e = val - emin, de = der;
        if (tmin - th >= 0.0) { e += 30.0 * (tmin - th) + y[0]; de += -30.0; }  // step(x)=1 for x>=0
        if (th - tmax >= 0.0) { e += 30.0 * (th - tmax) + y[9]; de += 30.0; }
        *E = p[0] * e; *dEdth = p[0] * de;
    }
}

inline void dihedral_like(const BondedTerm& t, int64_t m, double phi, double* E, double* dEdphi) {
    const double* p = &t.p[m * t.npar];
    const double PI = 3.14159265358979323846;
    switch (t.kind) {
        case B_GAUSSIAN_DIHEDRAL: {  // p = k, phi0, sigma  GaussianDihedralForceFieldGenerator.cpp:5-12
            double dt = phi - p[1];
            double d = dt - std::floor((dt + PI) / (2.0 * PI)) * 2.0 * PI;
            double e = p[0] * std::exp(-d * d / (2.0 * p[2] * p[2]));
            *E = e; *dEdphi = -e * d / (p[2] * p[2]);
        } break;
        case B_COSINE_DIHEDRAL: {  // p = k, t0, n     CosineDihedralForceFieldGenerator.cpp:7-12
            double a = p[2] * (phi - p[1]);
            *E = p[0] * (1.0 - std::cos(a)); *dEdphi = p[0] * p[2] * std::sin(a);
        } break;
        case B_FLEX_DIHEDRAL: {  // p = k   FlexibleLocalDihedralForceFieldGenerator.cpp:28-37
            const double* c = t.tab; double emin = t.tab[7];
            double e = c[0] - emin, de = 0;
            for (int h = 1; h <= 3; h++) {
                double cs = std::cos(h * phi), sn = std::sin(h * phi);
                e += c[2 * h - 1] * cs + c[2 * h] * sn;
                de += h * (-c[2 * h - 1] * sn + c[2 * h] * cs);
            }
            *E = p[0] * e; *dEdphi = p[0] * de;
        } break;
        default: *E = 0; *dEdphi = 0;
    }
}

void eval_base_stacking(const ocm_context& c, const BondedTerm& t, std::vector<Vec3>* F, double* Egroup);
void eval_external(const ocm_context& c, const BondedTerm& t, std::vector<Vec3>* F, double* Egroup);

void eval_bonded(const ocm_context& c, const BondedTerm& t, std::vector<Vec3>* F, double* Egroup) {
    if (t.kind == B_3SPN2_STACK) { eval_base_stacking(c, t, F, Egroup); return; }
    if (t.kind == B_EXTERNAL) { eval_external(c, t, F, Egroup); return; }
    const double* box = c.sys.box;
    double Esum = 0;
    for (int64_t m = 0; m < t.n(); m++) {
        const int32_t* id = &t.idx[m * t.nbody];
        if (t.nbody == 2) {
            Vec3 d = c.x[id[0]] - c.x[id[1]];
            if (t.periodic) d = min_image(d, box);
            double r = std::sqrt(dot(d, d)), E, dE;
            bond_like(t, m, r, &E, &dE);
            Esum += E;
            if (F) {
                Vec3 f = d * (-dE / r);
                (*F)[id[0]] = (*F)[id[0]] + f;
                (*F)[id[1]] = (*F)[id[1]] - f;
            }
        } else if (t.nbody == 3) {
            // angle at the middle particle (OpenMM angle(p1,p2,p3))
            Vec3 a = c.x[id[0]] - c.x[id[1]], b = c.x[id[2]] - c.x[id[1]];
            if (t.periodic) { a = min_image(a, box); b = min_image(b, box); }
            double la = std::sqrt(dot(a, a)), lb = std::sqrt(dot(b, b));
            double cs = dot(a, b) / (la * lb);
            cs = std::max(-1.0, std::min(1.0, cs));
            double th = std::acos(cs), E, dE;
            angle_like(t, m, th, &E, &dE);
            Esum += E;
            if (F) {
                double sn = std::sqrt(std::max(1.0 - cs * cs, 1e-24));
                Vec3 ah = a * (1.0 / la), bh = b * (1.0 / lb);
                Vec3 fi = (bh - ah * cs) * (dE / (sn * la));
                Vec3 fk = (ah - bh * cs) * (dE / (sn * lb));
                (*F)[id[0]] = (*F)[id[0]] + fi;
                (*F)[id[2]] = (*F)[id[2]] + fk;
                (*F)[id[1]] = (*F)[id[1]] - fi - fk;
            }
        } else {
            // dihedral(p1,p2,p3,p4) with OpenMM's sign rule
            Vec3 Fv = c.x[id[0]] - c.x[id[1]], G = c.x[id[1]] - c.x[id[2]], H = c.x[id[3]] - c.x[id[2]];
            if (t.periodic) { Fv = min_image(Fv, box); G = min_image(G, box); H = min_image(H, box); }
            Vec3 A = cross(Fv, G), B = cross(H, G);
            double A2 = dot(A, A), B2 = dot(B, B), lG = std::sqrt(dot(G, G));
            double cs = dot(A, B) / std::sqrt(A2 * B2);
            cs = std::max(-1.0, std::min(1.0, cs));
            double phi = std::acos(cs);
            // sign: d0 = p1-p2 = Fv, c1 = d1 x d2 = (p3-p2) x (p3-p4) = (-G) x (-H) = G x H = -B
            if (dot(Fv, B) > 0.0) phi = -phi;   // d0.c1 < 0  <=>  Fv.B > 0
            double E, dE;
            dihedral_like(t, m, phi, &E, &dE);
            Esum += E;
            if (F) {
                double FG = dot(Fv, G), HG = dot(H, G);
                Vec3 dp1 = A * (-lG / A2);
                Vec3 dp4 = B * (lG / B2);
                Vec3 dp2 = A * (lG / A2) + A * (FG / (A2 * lG)) - B * (HG / (B2 * lG));
                Vec3 dp3 = B * (-lG / B2) - A * (FG / (A2 * lG)) + B * (HG / (B2 * lG));
                (*F)[id[0]] = (*F)[id[0]] - dp1 * dE;
                (*F)[id[1]] = (*F)[id[1]] - dp2 * dE;
                (*F)[id[2]] = (*F)[id[2]] - dp3 * dE;
                (*F)[id[3]] = (*F)[id[3]] - dp4 * dE;
            }
        }
    }
    if (Egroup) Egroup[t.group] += Esum;
}

// ------------------------------------------------------------------------------------------------
// pair potentials: E(r) and dE/dr for one pair of term t (r < rc, not excluded, groups allowed)
inline double pow6(double x) { double x2 = x * x; return x2 * x2 * x2; }
static const double kTwoSixth = std::pow(2.0, 1.0 / 6.0);

inline void pair_energy(const PairTerm& t, int i, int j, double r, double* E, double* dEdr) {
    switch (t.kind) {
        case P_EXV: {  // ExcludedVolumeForceFieldGenerator.cpp:25-27; g0 = eps, g1 = rc^-12
            double s = t.p0[i] + t.p0[j];
            double s12 = std::pow(s, 12), ir = 1.0 / r, ir12 = std::pow(ir, 12);
            *E = t.g0 * s12 * (ir12 - t.g1);
            *dEdr = -12.0 * t.g0 * s12 * ir12 * ir;
        } break;
        case P_DH: {  // DebyeHuckelForceFieldGenerator.cpp:54-57; g0 = K, g1 = debye, g2 = cc
            double qq = t.p0[i] * t.p0[j] * t.g0;
            double ex = std::exp(-r / t.g1);
            *E = qq * (ex / r - t.g2);
            *dEdr = -qq * ex * (1.0 / (r * r) + 1.0 / (t.g1 * r));
        } break;
        case P_LJREP: {  // LennardJonesRepulsiveForceFieldGenerator.cpp:35-43; p0 = eps, p1 = sigma
            double sg = 0.5 * (t.p1[i] + t.p1[j]), ep = 0.5 * (t.p0[i] + t.p0[j]);
            double th = sg * kTwoSixth;
            if (th - r >= 0.0) {
                double sr = sg / r, s6 = pow6(sr), s12 = s6 * s6;
                *E = ep * (4.0 * (s12 - s6) + 1.0);
                *dEdr = ep * 4.0 * (-12.0 * s12 + 6.0 * s6) / r;
            } else { *E = 0; *dEdr = 0; }
        } break;
        case P_WCA: {  // WeeksChandlerAndersenForceFieldGenerator.cpp:34-41; p0 = eps (geometric mean), p1 = sigma
            double sg = 0.5 * (t.p1[i] + t.p1[j]), ep = std::sqrt(t.p0[i] * t.p0[j]);
            double th = sg * kTwoSixth;
            if (th - r >= 0.0) {
                double sr = sg / r, s6 = pow6(sr), s12 = s6 * s6;
                *E = 4.0 * ep * (s12 - s6) + ep;
                *dEdr = ep * 4.0 * (-12.0 * s12 + 6.0 * s6) / r;
            } else { *E = 0; *dEdr = 0; }
        } break;
        case P_LJATT: {  // LennardJonesAttractiveForceFieldGenerator.cpp:40-50 (+ switch :113-115)
            double sg = 0.5 * (t.p1[i] + t.p1[j]), ep = 0.5 * (t.p0[i] + t.p0[j]);
            double th = sg * kTwoSixth;
            double e = 0, de = 0;
            if (r - th >= 0.0) {
                double sr = sg / r, s6 = pow6(sr), s12 = s6 * s6;
                e += ep * 4.0 * (s12 - s6);
                de += ep * 4.0 * (-12.0 * s12 + 6.0 * s6) / r;
            }
            if (th - r >= 0.0) e -= ep;
            // OpenMM switching function
            if (r > t.rsw) {
                double w = t.rc - t.rsw, x = (r - t.rsw) / w;
                double S = 1.0 + x * x * x * (-10.0 + x * (15.0 - 6.0 * x));
                double dS = x * x * (-30.0 + x * (60.0 - 30.0 * x)) / w;
                de = S * de + e * dS;
                e = e * S;
            }
            *E = e; *dEdr = de;
        } break;
        case P_3SPN2EXV: {  // ThreeSPN2ExcludedVolumeForceFieldGenerator.cpp:21-26
            double sg = 0.5 * (t.p1[i] + t.p1[j]), ep = std::sqrt(t.p0[i] * t.p0[j]);
            if (sg - r >= 0.0) {
                double sr = sg / r, s6 = pow6(sr), s12 = s6 * s6;
                *E = ep * (s12 - 2.0 * s6) + ep;
                *dEdr = ep * (-12.0 * s12 + 12.0 * s6) / r;
            } else { *E = 0; *dEdr = 0; }
        } break;
        case P_ISOLF: {  // iSoLFAttractiveForceFieldGenerator.cpp:41-50; p0 = eps (geometric), p1 = sigma, p2 = omega (arithmetic)
            const double PI = 3.1415926535897932385;
            double ep = std::sqrt(t.p0[i] * t.p0[j]), base = 0.5 * (t.p1[i] + t.p1[j]) * kTwoSixth, om = 0.5 * (t.p2[i] + t.p2[j]);
            // -eps (step(base - r) + step(r - base) step(base + omega - r) cos(0.5 pi (r - base) / omega)^2), step(0) = 1
            double e = 0, de = 0;
            if (base - r >= 0.0) e -= ep;
            if (r - base >= 0.0 && base + om - r >= 0.0) {
                double a = 0.5 * PI * (r - base) / om, cs = std::cos(a);
                e -= ep * cs * cs;
                de += ep * 2.0 * cs * std::sin(a) * 0.5 * PI / om;
            }
            *E = e; *dEdr = de;
        } break;
        case P_TRIG: {  // TrigonometricForceFieldGenerator.cpp:43-53; p0 = epsilon (geometric), p1 = sigma, p2 = omega (arithmetic)
            double ep = std::sqrt(t.p0[i] * t.p0[j]), th = 0.5 * (t.p1[i] + t.p1[j]) * kTwoSixth, om = 0.5 * (t.p2[i] + t.p2[j]);
            // -epsilon (step(th + omega - r) + step(r - th) step(th + omega - r) (2 u^3 - 3 u^2)), u = (r - th) / omega
            double e = 0, de = 0;
            if (th + om - r >= 0.0) {
                e -= ep;
                if (r - th >= 0.0) {
                    double u = (r - th) / om;
                    e -= ep * (2.0 * u * u * u - 3.0 * u * u);
                    de -= ep * (6.0 * u * u - 6.0 * u) / om;
                }
            }
            *E = e; *dEdr = de;
        } break;
    }
}

inline bool pair_allowed(const PairTerm& t, int i, int j) {
    if (t.nper > 0 && i / t.nper != j / t.nper) return false;
    if (t.ncls > 0) {
        int ci = t.cls[i], cj = t.cls[j];
        if (ci < 0 || cj < 0 || !t.allow[ci * t.ncls + cj]) return false;
    }
    const auto& e = t.excl[i];
    return !std::binary_search(e.begin(), e.end(), j);
}

// --- canonical FP32 / fixed-point predicate shared (bit-exactly) with the CUDA engine ---------
// Periodic boxes: positions are quantised to 32-bit fixed-point box fractions; the wrapped
// integer difference IS the minimum image.  Unlimited boundary: FP32 positions.
struct Canon {
    bool periodic;
    std::vector<uint32_t> fx, fy, fz;  // periodic
    std::vector<float> px, py, pz;     // non-periodic
    float sx, sy, sz;                  // box / 2^32
};

Canon canonicalise(const ocm_context& c) {
    Canon q;
    int n = c.sys.n;
    // (unlimited boundary with the engine's virtual box: the same fixed-point quantisation in that box; option
    // "virtual_box_x/y/z", as the engine reports it)
    q.periodic = c.sys.has_box || c.vbox[0] > 0;
    const double* box = c.sys.has_box ? c.sys.box : c.vbox;
    if (q.periodic) {
        q.fx.resize(n); q.fy.resize(n); q.fz.resize(n);
        double inv[3] = {1.0 / box[0], 1.0 / box[1], 1.0 / box[2]};
        for (int i = 0; i < n; i++) {
            double u[3] = {c.x[i].x * inv[0], c.x[i].y * inv[1], c.x[i].z * inv[2]};
            uint32_t f[3];
            for (int k = 0; k < 3; k++) {
                double w = u[k] - std::floor(u[k]);
                f[k] = (uint32_t)(uint64_t)(w * 4294967296.0);
            }
            q.fx[i] = f[0]; q.fy[i] = f[1]; q.fz[i] = f[2];
        }
        q.sx = (float)(box[0] / 4294967296.0);
        q.sy = (float)(box[1] / 4294967296.0);
        q.sz = (float)(box[2] / 4294967296.0);
    } else {
        q.px.resize(n); q.py.resize(n); q.pz.resize(n);
        for (int i = 0; i < n; i++) {
            q.px[i] = (float)c.x[i].x; q.py[i] = (float)c.x[i].y; q.pz[i] = (float)c.x[i].z;
        }
    }
    return q;
}

inline float canon_r2(const Canon& q, int i, int j) {
    float dx, dy, dz;
    if (q.periodic) {
        dx = (float)(int32_t)(q.fx[i] - q.fx[j]) * q.sx;
        dy = (float)(int32_t)(q.fy[i] - q.fy[j]) * q.sy;
        dz = (float)(int32_t)(q.fz[i] - q.fz[j]) * q.sz;
    } else {
        dx = q.px[i] - q.px[j]; dy = q.py[i] - q.py[j]; dz = q.pz[i] - q.pz[j];
    }
    return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}

// effective interaction cutoff of a term (pairs beyond contribute exactly zero)
double effective_rc(const PairTerm& t) {
    double rc = t.rc;
    if (t.kind == P_LJREP || t.kind == P_WCA || t.kind == P_3SPN2EXV) {
        // zero beyond max over pairs of the zone boundary
        double m1 = 0, m2 = 0;  // two largest sigma
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

// cell-list enumeration of candidate pairs (double precision positions), calling fn(i, j, d, r)
// for every unordered participant pair with r < rc.  Brute force for small systems.
template <class Fn>
void for_each_pair(const ocm_context& c, const PairTerm& t, double rc, bool parallel, Fn fn) {
    const int np = (int)t.part.size();
    const double* box = c.sys.box;
    const bool per = t.periodic != 0;
    const double rc2 = rc * rc;
    bool brute = c.force_brute || np < 1500;
    int nc[3] = {1, 1, 1};
    double lo[3] = {0, 0, 0}, w[3] = {1, 1, 1};
    if (!brute) {
        if (per) {
            for (int k = 0; k < 3; k++) {
                nc[k] = (int)std::floor(box[k] / rc);
                if (nc[k] < 3) { brute = true; }
                w[k] = box[k] / std::max(nc[k], 1);
            }
        } else {
            double hi[3] = {-1e300, -1e300, -1e300};
            lo[0] = lo[1] = lo[2] = 1e300;
            for (int i : t.part) {
                const double* xi = &c.x[i].x;
                for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], xi[k]); hi[k] = std::max(hi[k], xi[k]); }
            }
            for (int k = 0; k < 3; k++) {
                nc[k] = std::max(1, std::min(256, (int)std::floor((hi[k] - lo[k]) / rc) + 1));
                w[k] = std::max(rc, (hi[k] - lo[k]) / nc[k] * (1 + 1e-12));
            }
        }
    }
    if (brute) {
#pragma omp parallel for schedule(static, 8) if (parallel)   // (static: which thread sums what does not depend on timing)
        for (int a = 0; a < np; a++) {
            int i = t.part[a];
            for (int b = a + 1; b < np; b++) {
                int j = t.part[b];
                Vec3 d = c.x[i] - c.x[j];
                if (per) d = min_image(d, box);
                double r2 = dot(d, d);
                if (r2 < rc2) fn(i, j, d, std::sqrt(r2));
            }
        }
        return;
    }
    // bin; beads are copied in cell order (wrapped into the box when periodic) so that a cell's candidates are
    // contiguous, and the image of a neighbour cell is a constant shift per cell pair (cell width >= rc and
    // >= 3 cells per periodic axis: the pair found through adjacent cells is the minimum image)
    int64_t ncell = (int64_t)nc[0] * nc[1] * nc[2];
    std::vector<int> cell_of(np), count(ncell + 1, 0), ids(np);
    std::vector<Vec3> xw(np), xs(np);
    for (int a = 0; a < np; a++) {
        const double* xi = &c.x[t.part[a]].x;
        int ci[3];
        double u[3];
        for (int k = 0; k < 3; k++) {
            u[k] = per ? (xi[k] - box[k] * std::floor(xi[k] / box[k])) : xi[k];
            ci[k] = std::max(0, std::min(nc[k] - 1, (int)std::floor((per ? u[k] : u[k] - lo[k]) / w[k])));
        }
        xw[a] = Vec3{u[0], u[1], u[2]};
        cell_of[a] = (ci[2] * nc[1] + ci[1]) * nc[0] + ci[0];
        count[cell_of[a] + 1]++;
    }
    for (int64_t k = 0; k < ncell; k++) count[k + 1] += count[k];
    std::vector<int> cursor(count.begin(), count.end() - 1);
    for (int a = 0; a < np; a++) {
        int slot = cursor[cell_of[a]]++;
        ids[slot] = t.part[a];
        xs[slot] = xw[a];
    }
#pragma omp parallel for schedule(static, 4) if (parallel)
    for (int64_t cidx = 0; cidx < ncell; cidx++) {
        int cx = cidx % nc[0], cy = (cidx / nc[0]) % nc[1], cz = cidx / ((int64_t)nc[0] * nc[1]);
        for (int dz = -1; dz <= 1; dz++)
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++) {
                    int ox = cx + dx, oy = cy + dy, oz = cz + dz;
                    Vec3 shift{0, 0, 0};  // position of the neighbour cell's image relative to its wrapped copy
                    if (per) {
                        if (ox < 0) { ox += nc[0]; shift.x = -box[0]; } else if (ox >= nc[0]) { ox -= nc[0]; shift.x = box[0]; }
                        if (oy < 0) { oy += nc[1]; shift.y = -box[1]; } else if (oy >= nc[1]) { oy -= nc[1]; shift.y = box[1]; }
                        if (oz < 0) { oz += nc[2]; shift.z = -box[2]; } else if (oz >= nc[2]) { oz -= nc[2]; shift.z = box[2]; }
                    } else if (ox < 0 || oy < 0 || oz < 0 || ox >= nc[0] || oy >= nc[1] || oz >= nc[2]) continue;
                    int64_t oidx = ((int64_t)oz * nc[1] + oy) * nc[0] + ox;
                    if (oidx < cidx) continue;  // each unordered cell pair once (nc>=3 => no aliasing)
                    for (int ua = count[cidx]; ua < count[cidx + 1]; ua++) {
                        const int i = ids[ua];
                        const Vec3 xi = xs[ua] - shift;
                        int vb0 = (oidx == cidx) ? ua + 1 : count[oidx];
                        for (int vb = vb0; vb < count[oidx + 1]; vb++) {
                            Vec3 d = xi - xs[vb];
                            double r2 = dot(d, d);
                            if (r2 < rc2) {
                                const int j = ids[vb];
                                if (i < j) fn(i, j, d, std::sqrt(r2));
                                else fn(j, i, d * -1.0, std::sqrt(r2));
                            }
                        }
                    }
                }
    }
}

void eval_pair(const ocm_context& c, const PairTerm& t, std::vector<Vec3>* F, double* Egroup) {
    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
    const int n = c.sys.n;
    std::vector<std::vector<Vec3>> Fl;
    if (F) Fl.assign(nthreads, std::vector<Vec3>(n, Vec3{0, 0, 0}));
    std::vector<double> El(nthreads * 8, 0.0);
    for_each_pair(c, t, t.rc, true, [&](int i, int j, Vec3 d, double r) {
        if (!pair_allowed(t, i, j)) return;
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        double E, dE;
        pair_energy(t, i, j, r, &E, &dE);
        El[tid * 8] += E;
        if (F) {
            Vec3 f = d * (-dE / r);
            Fl[tid][i] = Fl[tid][i] + f;
            Fl[tid][j] = Fl[tid][j] - f;
        }
    });
    double E = 0;
    for (int k = 0; k < nthreads; k++) E += El[k * 8];
    if (Egroup) Egroup[t.group] += E;
    if (F)
        for (int k = 0; k < nthreads; k++)
            for (int i = 0; i < n; i++) (*F)[i] = (*F)[i] + Fl[k][i];
}

void eval_da(const ocm_context& c, const DATerm& t, std::vector<Vec3>* F, double* Egroup);

// HarmonicCoMPullingForceFieldGenerator.cpp:5-18: "0.5 * k * (distance(g1, g2) - v0)^2" between the centres of two groups,
// weights = masses (CustomCentroidBondForce::addGroup without weights); every bead takes its mass share of the force
void eval_com(const ocm_context& c, const ComTerm& t, std::vector<Vec3>* F, double* Egroup) {
    Vec3 cen[2];
    double M[2];
    const std::vector<int32_t>* grp[2] = {&t.g1, &t.g2};
    for (int q = 0; q < 2; q++) {
        M[q] = 0; cen[q] = Vec3{0, 0, 0};
        for (int32_t i : *grp[q]) { M[q] += c.sys.mass[i]; cen[q] = cen[q] + c.x[i] * c.sys.mass[i]; }
        cen[q] = cen[q] * (1.0 / M[q]);
    }
    Vec3 d = cen[0] - cen[1];
    if (t.periodic) d = min_image(d, c.sys.box);
    const double r = std::sqrt(dot(d, d)), dl = r - t.v0;
    if (Egroup) Egroup[t.group] += 0.5 * t.k * dl * dl;
    if (F && r > 0.0) {
        const Vec3 f = d * (-t.k * dl / r);   // force on centre 1
        for (int32_t i : t.g1) (*F)[i] = (*F)[i] + f * (c.sys.mass[i] / M[0]);
        for (int32_t i : t.g2) (*F)[i] = (*F)[i] - f * (c.sys.mass[i] / M[1]);
    }
}

void compute(const ocm_context& c, std::vector<Vec3>* F, double* Egroup, uint32_t mask) {
    if (F) F->assign(c.sys.n, Vec3{0, 0, 0});
    if (Egroup) std::fill(Egroup, Egroup + 32, 0.0);
    for (const auto& t : c.sys.bonded)
        if (mask >> t.group & 1) eval_bonded(c, t, F, Egroup);
    for (const auto& t : c.sys.pairs)
        if (mask >> t.group & 1) eval_pair(c, t, F, Egroup);
    for (const auto& t : c.sys.das)
        if (mask >> t.group & 1) eval_da(c, t, F, Egroup);
    for (const auto& t : c.sys.coms)
        if (mask >> t.group & 1) eval_com(c, t, F, Egroup);
}

// ------------------------------------------------------------------------------------------------
// 3SPN.2 base stacking / base pairing / cross stacking and PDNS.
// These restate the reference's Lepton expression strings LITERALLY and differentiate them with
// forward-mode automatic differentiation (dual numbers carrying d/d(x,y,z) of up to six particles), with
// Lepton's conventions: step(x) = (x >= 0), d step = 0, max/min differentiate the selected branch
// (ties pick the first argument).  No hand-derived gradient is involved, so the CUDA engine's analytic
// forces are checked against an independent derivation.
constexpr int kND = 18;
struct Dual {
    double v;
    double d[kND];
};
inline Dual dconst(double v) { Dual r; r.v = v; for (int k = 0; k < kND; k++) r.d[k] = 0; return r; }
inline Dual operator+(const Dual& a, const Dual& b) { Dual r; r.v = a.v + b.v; for (int k = 0; k < kND; k++) r.d[k] = a.d[k] + b.d[k]; return r; }
inline Dual operator-(const Dual& a, const Dual& b) { Dual r; r.v = a.v - b.v; for (int k = 0; k < kND; k++) r.d[k] = a.d[k] - b.d[k]; return r; }
inline Dual operator-(const Dual& a) { Dual r; r.v = -a.v; for (int k = 0; k < kND; k++) r.d[k] = -a.d[k]; return r; }
inline Dual operator*(const Dual& a, const Dual& b) { Dual r; r.v = a.v * b.v; for (int k = 0; k < kND; k++) r.d[k] = a.d[k] * b.v + a.v * b.d[k]; return r; }
inline Dual operator/(const Dual& a, const Dual& b) { Dual r; r.v = a.v / b.v; for (int k = 0; k < kND; k++) r.d[k] = (a.d[k] - r.v * b.d[k]) / b.v; return r; }
inline Dual operator+(const Dual& a, double b) { Dual r = a; r.v += b; return r; }
inline Dual operator-(const Dual& a, double b) { Dual r = a; r.v -= b; return r; }
inline Dual operator-(double a, const Dual& b) { Dual r = -b; r.v += a; return r; }
inline Dual operator+(double a, const Dual& b) { return b + a; }
inline Dual operator*(const Dual& a, double b) { Dual r; r.v = a.v * b; for (int k = 0; k < kND; k++) r.d[k] = a.d[k] * b; return r; }
inline Dual operator*(double a, const Dual& b) { return b * a; }
inline Dual chain(const Dual& a, double v, double dv) { Dual r; r.v = v; for (int k = 0; k < kND; k++) r.d[k] = dv * a.d[k]; return r; }
inline Dual dexp(const Dual& a) { double e = std::exp(a.v); return chain(a, e, e); }
inline Dual dsin(const Dual& a) { return chain(a, std::sin(a.v), std::cos(a.v)); }
inline Dual dcos(const Dual& a) { return chain(a, std::cos(a.v), -std::sin(a.v)); }
inline Dual dsqrt(const Dual& a) { double r = std::sqrt(a.v); return chain(a, r, 0.5 / r); }
inline Dual dacos(const Dual& a) { return chain(a, std::acos(a.v), -1.0 / std::sqrt(1.0 - a.v * a.v)); }
inline Dual datan2(const Dual& y, const Dual& x) {
    Dual r; r.v = std::atan2(y.v, x.v);
    const double den = x.v * x.v + y.v * y.v;
    for (int k = 0; k < kND; k++) r.d[k] = (x.v * y.d[k] - y.v * x.d[k]) / den;
    return r;
}
inline Dual dstep(const Dual& a) { return dconst(a.v >= 0.0 ? 1.0 : 0.0); }
inline Dual dmax(const Dual& a, const Dual& b) { return (a.v >= b.v) ? a : b; }
inline Dual dmin(const Dual& a, const Dual& b) { return (a.v <= b.v) ? a : b; }
inline Dual dsq(const Dual& a) { return a * a; }

struct DVec { Dual x, y, z; };
inline DVec operator-(const DVec& a, const DVec& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Dual ddot(const DVec& a, const DVec& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline DVec dcross(const DVec& a, const DVec& b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

// up to six particles; slot s carries the seeds d/d(x_s, y_s, z_s)
struct DAGeom {
    const ocm_context* c;
    bool periodic;
    int bead[6];
    DVec pos(int s) const {
        DVec p{dconst(c->x[bead[s]].x), dconst(c->x[bead[s]].y), dconst(c->x[bead[s]].z)};
        p.x.d[3 * s] = 1.0; p.y.d[3 * s + 1] = 1.0; p.z.d[3 * s + 2] = 1.0;
        return p;
    }
    // separation vector p_a - p_b with the minimum image (the image shift is a constant)
    DVec sep(int a, int b) const {
        DVec d = pos(a) - pos(b);
        if (periodic) {
            Vec3 raw{d.x.v, d.y.v, d.z.v}, mi = min_image(raw, c->sys.box);
            d.x.v = mi.x; d.y.v = mi.y; d.z.v = mi.z;
        }
        return d;
    }
    Dual distance(int a, int b) const { DVec d = sep(a, b); return dsqrt(ddot(d, d)); }
    // OpenMM angle(p1,p2,p3): angle at p2, acos of the normalised dot product
    Dual angle(int a, int b, int c3) const {
        DVec u = sep(a, b), w = sep(c3, b);
        Dual cs = ddot(u, w) / dsqrt(ddot(u, u) * ddot(w, w));
        if (cs.v > 1.0) cs.v = 1.0;
        if (cs.v < -1.0) cs.v = -1.0;
        return dacos(cs);
    }
    // OpenMM dihedral(p1,p2,p3,p4) with its sign rule, written with atan2:
    // F = p1-p2, G = p2-p3, H = p4-p3, A = F x G, B = H x G; phi = atan2([F,G,H] |G|, A.B)
    Dual dihedral(int a, int b, int c3, int d4) const {
        DVec F = sep(a, b), G = sep(b, c3), H = sep(d4, c3);
        DVec A = dcross(F, G), B = dcross(H, G);
        Dual triple = ddot(F, dcross(G, H));
        return datan2(triple * dsqrt(ddot(G, G)), ddot(A, B));
    }
};

const double kPi = 3.1415926535897932385;
// "max(f*rect0, rect1)" with f = 1 - cos(dt)^2, rect0 = step(pi+dt) step(pi-dt), rect1 = step(pi/2+dt) step(pi/2-dt)
inline Dual modulation(const Dual& dt) {
    Dual f = 1.0 - dsq(dcos(dt));
    Dual rect0 = dstep(kPi + dt) * dstep(kPi - dt);
    Dual rect1 = dstep(kPi / 2 + dt) * dstep(kPi / 2 - dt);
    return dmax(f * rect0, rect1);
}
inline Dual morse_rep(double eps, double alpha, const Dual& dr) {  // eps (1 - exp(-alpha dr))^2 step(-dr)
    return eps * dsq(1.0 - dexp(-(alpha * dr))) * dstep(-dr);
}
inline Dual morse_attr(double eps, double alpha, const Dual& dr) {  // eps (1 - exp(-alpha dr))^2 step(dr) - eps
    return eps * dsq(1.0 - dexp(-(alpha * dr))) * dstep(dr) - eps;
}

inline void scatter_dual(const Dual& E, const int* bead, int nslots, std::vector<Vec3>* F) {
    for (int s = 0; s < nslots; s++) {
        if (bead[s] < 0) continue;
        Vec3& f = (*F)[bead[s]];
        f.x -= E.d[3 * s]; f.y -= E.d[3 * s + 1]; f.z -= E.d[3 * s + 2];
    }
}

// ThreeSPN2BaseStackingForceFieldGenerator.cpp:6-21 (CustomCompoundBondForce, 3 particles)
void eval_base_stacking(const ocm_context& c, const BondedTerm& t, std::vector<Vec3>* F, double* Egroup) {
    double Esum = 0;
    for (int64_t m = 0; m < t.n(); m++) {
        DAGeom g{&c, t.periodic != 0, {t.idx[3 * m], t.idx[3 * m + 1], t.idx[3 * m + 2], -1, -1, -1}};
        const double* p = &t.p[m * t.npar];  // eps, r0, t0, alpha, K_BS
        Dual dr = g.distance(1, 2) - p[1];
        Dual dt = p[4] * (g.angle(0, 1, 2) - p[2]);
        Dual f = 1.0 - dsq(dcos(dt));
        Dual rect1 = dstep(dt + kPi / 2) * dstep(kPi / 2 - dt);
        Dual rect2 = dstep(dt + kPi) * dstep(kPi - dt);
        Dual f2 = dmax(f * rect2, rect1);
        Dual E = morse_rep(p[0], p[3], dr) + f2 * morse_attr(p[0], p[3], dr);
        Esum += E.v;
        if (F) scatter_dual(E, g.bead, 3, F);
    }
    if (Egroup) Egroup[t.group] += Esum;
}

// One-bead external terms: the reference's expression strings restated literally on dual numbers seeded with the bead's own
// coordinates (absolute, unwrapped, as CustomExternalForce sees them on the Reference platform).
inline Dual dpow12(const Dual& a) { Dual a2 = a * a, a4 = a2 * a2; return a4 * a4 * a4; }
void eval_external(const ocm_context& c, const BondedTerm& t, std::vector<Vec3>* F, double* Egroup) {
    double Esum = 0;
    for (int64_t m = 0; m < t.n(); m++) {
        const double* p = &t.p[m * t.npar];
        DAGeom g{&c, false, {t.idx[m], -1, -1, -1, -1, -1}};
        const DVec x = g.pos(0);
        Dual E = dconst(0.0);
        switch ((int)p[0]) {
            case X_POSITION_RESTRAINT:  // "k * (((x-x0)^2 + (y-y0)^2 + (z-z0)^2)^(1/2) - v0)^2", p = x0 y0 z0 k v0
                                        // PositionRestraintForceFieldGenerator.cpp:8-10 (the periodic string of :12-17 is dead code)
                E = p[4] * dsq(dsqrt(dsq(x.x - p[1]) + dsq(x.y - p[2]) + dsq(x.z - p[3])) - p[5]);
                break;
            case X_PULLING:  // "-kx * x - ky * y - kz * z"     PullingForceFieldGenerator.cpp:5-7
                E = -(p[1] * x.x) - p[2] * x.y - p[3] * x.z;
                break;
            case X_RECTANGULAR_BOX: {  // "epsilon * ((radius/(x-x_lower))^12 + (radius/(x_upper-x))^12 + ...)", p = eps lo[3] hi[3] radius
                                       // EXVRectangularBoxForceFieldGenerator.cpp:33-38
                const Dual q[3] = {x.x, x.y, x.z};
                for (int a = 0; a < 3; a++)
                    E = E + dpow12(dconst(p[8]) / (q[a] - p[2 + a])) + dpow12(dconst(p[8]) / (p[5 + a] - q[a]));
                E = p[1] * E;
            } break;
            default: {  // "k * (dr - v0)^2; dr = |x - h| or periodicdistance(x, h); h = ra_r * axis + shift; ra_r = axis . x",
                        // p = axis[3] shift[3] k v0     CylindricalRestraintForceFieldGenerator.cpp:10-26
                const Dual ra = p[1] * x.x + p[2] * x.y + p[3] * x.z;
                DVec d{x.x - (ra * p[1] + p[4]), x.y - (ra * p[2] + p[5]), x.z - (ra * p[3] + p[6])};
                if (t.periodic) {
                    const Vec3 mi = min_image(Vec3{d.x.v, d.y.v, d.z.v}, c.sys.box);
                    d.x.v = mi.x; d.y.v = mi.y; d.z.v = mi.z;
                }
                E = p[7] * dsq(dsqrt(ddot(d, d)) - p[8]);
            }
        }
        Esum += E.v;
        if (F) scatter_dual(E, g.bead, 1, F);
    }
    if (Egroup) Egroup[t.group] += Esum;
}

// slots: 0,1,2 = d1,d2,d3 ; 3,4,5 = a1,a2,a3
Dual da_energy(const ocm_context& c, const DATerm& t, int64_t d, int64_t a) {
    DAGeom g{&c, t.periodic != 0, {t.donors[3 * d], t.donors[3 * d + 1], t.donors[3 * d + 2],
                                  t.acceptors[3 * a], t.acceptors[3 * a + 1], t.acceptors[3 * a + 2]}};
    enum { d1 = 0, d2 = 1, d3 = 2, a1 = 3, a2 = 4, a3 = 5 };
    if (t.kind == DA_COMB_GO) {
        // "k * (5 * (r0 / distance(a1, d1))^12 - 6 * (r0 / distance(a1, d1))^10)", k = g0, r0 = g1
        // CombinatorialGoContactForceFieldGenerator.cpp:5-10
        const Dual s = dconst(t.g1) / g.distance(a1, d1), s2 = s * s, s4 = s2 * s2, s10 = s4 * s4 * s2;
        return t.g0 * (5.0 * (s10 * s2) - 6.0 * s10);
    }
    if (t.kind == DA_BASE_PAIR) {
        // ThreeSPN2BasePairForceFieldGenerator.cpp:8-26; per-donor eps, r0, t01, t02, phi0; alpha = g0, K = g1
        const double* p = &t.p[d * t.npar];
        Dual dphi = g.dihedral(d2, d1, a1, a2) - p[4];
        Dual dr = g.distance(d1, a1) - p[1];
        Dual dt1 = t.g1 * (g.angle(d2, d1, a1) - p[2]);
        Dual dt2 = t.g1 * (g.angle(a2, a1, d1) - p[3]);
        Dual rep = p[0] * dsq(1.0 - dexp(-(t.g0 * dr))) * (1.0 - dstep(dr));
        Dual attr = morse_attr(p[0], t.g0, dr);
        return rep + 0.5 * (1.0 + dcos(dphi)) * modulation(dt1) * modulation(dt2) * attr;
    }
    if (t.kind == DA_CROSS_STACK) {
        // ThreeSPN2CrossStackingForceFieldGenerator.cpp:8-31; per-acceptor eps, r0, t03, t0CS; g0 = K_BP, g1 = K_CS, g2 = alpha
        const double* p = &t.p[a * t.npar];
        Dual t1 = g.angle(d2, d1, a1), t2 = g.angle(d1, a1, a2), phi = g.dihedral(d2, d1, a1, a2);
        Dual cost3 = dsin(t1) * dsin(t2) * dcos(phi) - dcos(t1) * dcos(t2);
        Dual cost3lim = dmin(dmax(cost3, dconst(-0.99)), dconst(0.99));
        Dual t3 = dacos(cost3lim);
        Dual tCS = g.angle(d2, d1, a3);
        Dual dt3 = t.g0 * (t3 - p[2]), dtCS = t.g1 * (tCS - p[3]);
        Dual dr = g.distance(d1, a3) - p[1];
        return modulation(dt3) * modulation(dtCS) * morse_attr(p[0], t.g2, dr) * 0.5;
    }
    // ProteinDNANonSpecificForceFieldGenerator.cpp:60-83; per-donor k, r0, theta0, phi0; g0 = sigma, g1 = delta
    const double* p = &t.p[d * t.npar];
    const double K = kPi / (2.0 * t.g1);
    Dual t1 = g.angle(d1, d3, a1), t2 = g.angle(d3, a1, d2), phi1 = g.dihedral(d1, d3, a1, d2);
    Dual cost1 = dsin(t1) * dsin(t2) * dcos(phi1) - dcos(t1) * dcos(t2);
    Dual cost1lim = dmin(dmax(cost1, dconst(-0.99999)), dconst(0.99999));
    Dual theta = dacos(cost1lim);
    Dual phi = g.angle(d2, a1, a2);
    Dual dtheta = K * (theta - p[2]), dphi = K * (phi - p[3]);
    Dual dr = g.distance(d2, a1) - p[1];
    Dual f_r = dexp(-(dsq(dr) * (1.0 / (2.0 * t.g0 * t.g0))));
    return -(p[0] * f_r * modulation(dtheta) * modulation(dphi));
}

// all donor x acceptor combinations, not excluded, |d1 - a1| < cutoff
template <class Fn>
void for_each_da(const ocm_context& c, const DATerm& t, Fn fn) {
    for (int64_t d = 0; d < t.nd(); d++)
        for (int64_t a = 0; a < t.na(); a++) {
            Vec3 dv = c.x[t.donors[3 * d]] - c.x[t.acceptors[3 * a]];
            if (t.periodic) dv = min_image(dv, c.sys.box);
            if (!(dot(dv, dv) < t.cutoff * t.cutoff)) continue;
            if (std::binary_search(t.excl.begin(), t.excl.end(), std::make_pair((int32_t)d, (int32_t)a))) continue;
            fn(d, a);
        }
}

void eval_da(const ocm_context& c, const DATerm& t, std::vector<Vec3>* F, double* Egroup) {
    double Esum = 0;
    for_each_da(c, t, [&](int64_t d, int64_t a) {
        Dual E = da_energy(c, t, d, a);
        Esum += E.v;
        if (F) {
            int bead[6] = {t.donors[3 * d], t.donors[3 * d + 1], t.donors[3 * d + 2],
                           t.acceptors[3 * a], t.acceptors[3 * a + 1], t.acceptors[3 * a + 2]};
            // a bead may sit in two slots (e.g. d2 == a-side sugar never, but keep general): contributions add
            scatter_dual(E, bead, 6, F);
        }
    });
    if (Egroup) Egroup[t.group] += Esum;
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), the documented noise source shared with the CUDA engine:
// key = (seed_lo, seed_hi), counter = (bead, step_lo, step_hi, stream).
inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                          uint32_t out[4]) {
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

inline void normals3(uint64_t seed, uint32_t bead, uint64_t step, uint32_t stream, double g[3]) {
    uint32_t o[4];
    philox4x32_10(bead, (uint32_t)step, (uint32_t)(step >> 32), stream, (uint32_t)seed, (uint32_t)(seed >> 32), o);
    const double TWO_PI = 6.283185307179586476925;
    double u0 = (o[0] + 0.5) * (1.0 / 4294967296.0), u1 = (o[1] + 0.5) * (1.0 / 4294967296.0);
    double u2 = (o[2] + 0.5) * (1.0 / 4294967296.0), u3 = (o[3] + 0.5) * (1.0 / 4294967296.0);
    double r0 = std::sqrt(-2.0 * std::log(u0)), r1 = std::sqrt(-2.0 * std::log(u2));
    g[0] = r0 * std::cos(TWO_PI * u1);
    g[1] = r0 * std::sin(TWO_PI * u1);
    g[2] = r1 * std::cos(TWO_PI * u3);
}

int check_idx(const ocm_system* s, const int32_t* idx, int64_t cnt, const char* what) {
    for (int64_t k = 0; k < cnt; k++)
        if (idx[k] < 0 || idx[k] >= s->n) return fail(std::string(what) + ": particle index out of range");
    return 0;
}

int add_bonded(ocm_system* s, BondedKind kind, int group, int periodic, int nbody, int64_t n,
               const int32_t* idx, int npar, std::vector<const double*> cols, const char* what) {
    if (!s) return fail("null system");
    if (group < 0 || group > 31) return fail("force group must be in 0..31");
    if (periodic && !s->has_box) return fail(std::string(what) + ": periodic term without a box");
    if (check_idx(s, idx, n * nbody, what)) return 1;
    BondedTerm t;
    t.kind = kind; t.group = group; t.periodic = periodic; t.nbody = nbody; t.npar = npar;
    t.idx.assign(idx, idx + n * nbody);
    t.p.resize(n * npar);
    for (int64_t m = 0; m < n; m++)
        for (int q = 0; q < npar; q++) t.p[m * npar + q] = cols[q][m];
    std::fill(t.tab, t.tab + 24, 0.0);
    s->bonded.push_back(std::move(t));
    return 0;
}

int add_pair(ocm_system* s, PairKind kind, int group, double rc, double rsw, double g0, double g1, double g2,
             const double* p0, const double* p1, const ocm_pair_topology* topo, const char* what, const double* p2 = nullptr) {
    if (!s) return fail("null system");
    if (group < 0 || group > 31) return fail("force group must be in 0..31");
    if (!topo) return fail("null topology");
    if (topo->periodic && !s->has_box) return fail(std::string(what) + ": periodic term without a box");
    if (topo->periodic)
        for (int k = 0; k < 3; k++)
            if (rc > 0.5 * s->box[k]) return fail(std::string(what) + ": cutoff exceeds half the box");
    PairTerm t;
    t.kind = kind; t.group = group; t.periodic = topo->periodic; t.rc = rc; t.rsw = rsw;
    t.g0 = g0; t.g1 = g1; t.g2 = g2;
    int n = s->n;
    t.p0.assign(p0, p0 + n);
    if (p1) t.p1.assign(p1, p1 + n); else t.p1 = t.p0;  // single-parameter kinds
    if (p2) t.p2.assign(p2, p2 + n); else t.p2 = t.p0;
    for (int i = 0; i < n; i++) {
        bool has0 = !std::isnan(t.p0[i]), has1 = !std::isnan(t.p1[i]);
        if (has0 != has1) return fail(std::string(what) + ": incomplete parameter set for a particle");
    }
    t.excl.assign(n, {});
    if (check_idx(s, topo->excl_ij, topo->n_excl * 2, what)) return 1;
    for (int64_t e = 0; e < topo->n_excl; e++) {
        int a = topo->excl_ij[2 * e], b = topo->excl_ij[2 * e + 1];
        if (a == b) continue;
        t.excl[a].push_back(b); t.excl[b].push_back(a);
    }
    for (auto& v : t.excl) { std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end()); }
    t.ncls = 0;
    t.cls.assign(n, -1);
    if (topo->n_igroups > 0) {
        // class = membership signature over all sets
        int ng = topo->n_igroups;
        std::vector<std::vector<uint8_t>> sig(n, std::vector<uint8_t>(2 * ng, 0));
        for (int g = 0; g < ng; g++) {
            for (int64_t k = topo->set1_off[g]; k < topo->set1_off[g + 1]; k++) {
                int b = topo->set1[k];
                if (b < 0 || b >= n) return fail(std::string(what) + ": interaction-group index out of range");
                sig[b][2 * g] = 1;
            }
            for (int64_t k = topo->set2_off[g]; k < topo->set2_off[g + 1]; k++) {
                int b = topo->set2[k];
                if (b < 0 || b >= n) return fail(std::string(what) + ": interaction-group index out of range");
                sig[b][2 * g + 1] = 1;
            }
        }
        std::map<std::vector<uint8_t>, int> ids;
        std::vector<std::vector<uint8_t>> reps;
        for (int i = 0; i < n; i++) {
            bool any = false;
            for (auto v : sig[i]) any |= v;
            if (!any) continue;
            if (std::isnan(t.p0[i])) return fail(std::string(what) + ": interaction group member without parameters");
            auto it = ids.find(sig[i]);
            if (it == ids.end()) { it = ids.emplace(sig[i], (int)reps.size()).first; reps.push_back(sig[i]); }
            t.cls[i] = it->second;
        }
        t.ncls = (int)reps.size();
        t.allow.assign(t.ncls * t.ncls, 0);
        for (int a = 0; a < t.ncls; a++)
            for (int b = 0; b < t.ncls; b++) {
                int cnt = 0;
                for (int g = 0; g < ng; g++)
                    if ((reps[a][2 * g] && reps[b][2 * g + 1]) || (reps[a][2 * g + 1] && reps[b][2 * g])) cnt++;
                if (cnt > 1) return fail(std::string(what) + ": overlapping interaction groups (a pair would be counted twice) are not supported");
                t.allow[a * t.ncls + b] = (uint8_t)cnt;
            }
        for (int i = 0; i < n; i++) if (t.cls[i] >= 0) t.part.push_back(i);
    } else {
        for (int i = 0; i < n; i++) {
            if (std::isnan(t.p0[i])) return fail(std::string(what) + ": a particle has no parameter but no interaction group was given");
            t.part.push_back(i);
        }
    }
    s->pairs.push_back(std::move(t));
    return 0;
}

}  // namespace

// ================================================================================================
extern "C" {

const char* oracle_last_error(void) { return g_err.c_str(); }
const char* oracle_version(void) { return "ocm-oracle 0.1 (CPU double precision; test infrastructure)"; }
int oracle_device_count(void) { return 0; }

int oracle_system_create(int32_t n, const double* mass, const double* box, ocm_system** out) {
    if (n <= 0 || !mass || !out) return fail("system_create: bad arguments");
    auto* s = new ocm_system();
    s->n = n; s->mass.assign(mass, mass + n);
    s->has_box = box != nullptr;
    for (int k = 0; k < 3; k++) s->box[k] = box ? box[k] : 0.0;
    if (box) for (int k = 0; k < 3; k++) if (!(box[k] > 0)) { delete s; return fail("system_create: box edge must be positive"); }
    *out = s;
    return 0;
}
void oracle_system_destroy(ocm_system* s) { delete s; }
int oracle_system_num_particles(const ocm_system* s) { return s ? s->n : -1; }

int oracle_system_add_harmonic_bond(ocm_system* s, int g, int per, int64_t n, const int32_t* ij, const double* r0, const double* k) {
    return add_bonded(s, B_HARMONIC_BOND, g, per, 2, n, ij, 2, {r0, k}, "harmonic_bond");
}
int oracle_system_add_gaussian_bond(ocm_system* s, int g, int per, int64_t n, const int32_t* ij, const double* k, const double* r0, const double* sigma) {
    return add_bonded(s, B_GAUSSIAN_BOND, g, per, 2, n, ij, 3, {k, r0, sigma}, "gaussian_bond");
}
int oracle_system_add_go_contact(ocm_system* s, int g, int per, int64_t n, const int32_t* ij, const double* k, const double* r0) {
    const std::vector<double> zero((size_t)std::max<int64_t>(n, 1), 0.0);
    return add_bonded(s, B_GO_CONTACT, g, per, 2, n, ij, 5, {k, r0, zero.data(), zero.data(), zero.data()}, "go_contact");
}
int oracle_system_add_capped_go_contact(ocm_system* s, int g, int per, int64_t n, const int32_t* ij, const double* k, const double* r0,
                                        const double* slope, const double* intercept, const double* cap) {
    return add_bonded(s, B_GO_CONTACT, g, per, 2, n, ij, 5, {k, r0, slope, intercept, cap}, "capped_go_contact");
}
int oracle_system_add_3spn2_bond(ocm_system* s, int g, int per, int64_t n, const int32_t* ij, const double* k2, const double* k4, const double* r0) {
    return add_bonded(s, B_3SPN2_BOND, g, per, 2, n, ij, 3, {k2, k4, r0}, "3spn2_bond");
}
int oracle_system_add_harmonic_angle(ocm_system* s, int g, int per, int64_t n, const int32_t* ijk, const double* th0, const double* k) {
    return add_bonded(s, B_HARMONIC_ANGLE, g, per, 3, n, ijk, 2, {th0, k}, "harmonic_angle");
}
int oracle_system_add_flexible_local_angle(ocm_system* s, int g, int per, int64_t n, const int32_t* ijk, const double* k,
                                           const double* y10, double tmin, double tmax, double emin) {
    if (add_bonded(s, B_FLEX_ANGLE, g, per, 3, n, ijk, 1, {k}, "flexible_local_angle")) return 1;
    BondedTerm& t = s->bonded.back();
    for (int q = 0; q < 10; q++) t.tab[q] = y10[q];
    natural_spline_y2(y10, 10, (tmax - tmin) / 9.0, t.tab + 10);
    t.tab[20] = tmin; t.tab[21] = tmax; t.tab[22] = emin;
    return 0;
}
int oracle_system_add_gaussian_dihedral(ocm_system* s, int g, int per, int64_t n, const int32_t* ijkl, const double* k, const double* phi0, const double* sigma) {
    return add_bonded(s, B_GAUSSIAN_DIHEDRAL, g, per, 4, n, ijkl, 3, {k, phi0, sigma}, "gaussian_dihedral");
}
int oracle_system_add_cosine_dihedral(ocm_system* s, int g, int per, int64_t n, const int32_t* ijkl, const double* k, const double* t0, const int32_t* nper) {
    std::vector<double> nn(n);
    for (int64_t m = 0; m < n; m++) nn[m] = nper[m];
    return add_bonded(s, B_COSINE_DIHEDRAL, g, per, 4, n, ijkl, 3, {k, t0, nn.data()}, "cosine_dihedral");
}
int oracle_system_add_flexible_local_dihedral(ocm_system* s, int g, int per, int64_t n, const int32_t* ijkl, const double* k, const double* coef7, double emin) {
    if (add_bonded(s, B_FLEX_DIHEDRAL, g, per, 4, n, ijkl, 1, {k}, "flexible_local_dihedral")) return 1;
    BondedTerm& t = s->bonded.back();
    for (int q = 0; q < 7; q++) t.tab[q] = coef7[q];
    t.tab[7] = emin;
    return 0;
}

int oracle_system_add_3spn2_base_stacking(ocm_system* s, int g, int per, int64_t n, const int32_t* ijk, const double* eps,
                                          const double* r0, const double* theta0, double alpha, double K_BS) {
    std::vector<double> al((size_t)std::max<int64_t>(n, 0), alpha), kb((size_t)std::max<int64_t>(n, 0), K_BS);
    return add_bonded(s, B_3SPN2_STACK, g, per, 3, n, ijk, 5, {eps, r0, theta0, al.data(), kb.data()}, "3spn2_base_stacking");
}

static int add_da(ocm_system* s, DAKind kind, int group, const ocm_da_topology* topo, int npar, bool per_donor,
                  std::vector<const double*> cols, double g0, double g1, double g2, const char* what) {
    if (!s || !topo) return fail(std::string(what) + ": null argument");
    if (group < 0 || group > 31) return fail("force group must be in 0..31");
    if (topo->periodic && !s->has_box) return fail(std::string(what) + ": periodic term without a box");
    if (!(topo->cutoff > 0)) return fail(std::string(what) + ": cutoff must be positive");
    DATerm t;
    t.kind = kind; t.group = group; t.periodic = topo->periodic ? 1 : 0; t.cutoff = topo->cutoff;
    t.npar = npar; t.g0 = g0; t.g1 = g1; t.g2 = g2;
    t.donors.assign(topo->donors, topo->donors + 3 * topo->n_donors);
    t.acceptors.assign(topo->acceptors, topo->acceptors + 3 * topo->n_acceptors);
    for (int32_t v : t.donors) if (v < -1 || v >= s->n) return fail(std::string(what) + ": donor index out of range");
    for (int32_t v : t.acceptors) if (v < -1 || v >= s->n) return fail(std::string(what) + ": acceptor index out of range");
    for (int64_t e = 0; e < topo->n_excl; e++) {
        int32_t d = topo->excl_da[2 * e], a = topo->excl_da[2 * e + 1];
        if (d < 0 || d >= topo->n_donors || a < 0 || a >= topo->n_acceptors) return fail(std::string(what) + ": exclusion index out of range");
        t.excl.emplace_back(d, a);
    }
    std::sort(t.excl.begin(), t.excl.end());
    t.excl.erase(std::unique(t.excl.begin(), t.excl.end()), t.excl.end());
    const int64_t cnt = per_donor ? topo->n_donors : topo->n_acceptors;
    t.p.resize((size_t)cnt * npar);
    for (int64_t m = 0; m < cnt; m++)
        for (int q = 0; q < npar; q++) t.p[m * npar + q] = cols[q][m];
    s->das.push_back(std::move(t));
    return 0;
}
static int add_external(ocm_system* s, int form, int g, int per, int64_t n, const int32_t* index, const std::vector<std::pair<const double*, int>>& cols,
                        const std::vector<double>& globals, const char* what) {
    // rows of 10 doubles: form, the globals, then the per-term columns (pointer, width); same layout as the engine's
    const int npar = 10;
    if (n < 0 || (n > 0 && !index)) return fail(std::string(what) + ": bad arguments");
    std::vector<std::vector<double>> col(npar, std::vector<double>((size_t)std::max<int64_t>(n, 1), 0.0));
    std::fill(col[0].begin(), col[0].end(), (double)form);
    int q = 1;
    for (double v : globals) { std::fill(col[q].begin(), col[q].end(), v); q++; }
    for (auto& cw : cols) {
        if (n > 0 && !cw.first) return fail(std::string(what) + ": null parameter array");
        for (int w = 0; w < cw.second; w++, q++)
            for (int64_t m = 0; m < n; m++) col[q][m] = cw.first[(size_t)m * cw.second + w];
    }
    std::vector<const double*> ptr;
    for (auto& v : col) ptr.push_back(v.data());
    return add_bonded(s, B_EXTERNAL, g, per, 1, n, index, npar, ptr, what);
}
int oracle_system_add_position_restraint(ocm_system* s, int g, int64_t n, const int32_t* index, const double* pos3, const double* k, const double* v0) {
    return add_external(s, X_POSITION_RESTRAINT, g, 0, n, index, {{pos3, 3}, {k, 1}, {v0, 1}}, {}, "position_restraint");
}
int oracle_system_add_pulling(ocm_system* s, int g, int64_t n, const int32_t* index, const double* force3) {
    return add_external(s, X_PULLING, g, 0, n, index, {{force3, 3}}, {}, "pulling");
}
int oracle_system_add_exv_rectangular_box(ocm_system* s, int g, double eps, const double* lower3, const double* upper3, int64_t n, const int32_t* index,
                                          const double* radius) {
    if (!s || !lower3 || !upper3) return fail("exv_rectangular_box: null argument");
    if (s->has_box) return fail("exv_rectangular_box: RectangularBox interaction does not work under periodic boundary condition");
    for (int a = 0; a < 3; a++) if (!(upper3[a] > lower3[a])) return fail("exv_rectangular_box: the upper edge must be larger than the lower edge");
    return add_external(s, X_RECTANGULAR_BOX, g, 0, n, index, {{radius, 1}},
                        {eps, lower3[0], lower3[1], lower3[2], upper3[0], upper3[1], upper3[2]}, "exv_rectangular_box");
}
int oracle_system_add_cylindrical_restraint(ocm_system* s, int g, int per, int64_t n, const int32_t* index, const double* axis3, const double* shift3,
                                            const double* k, const double* v0) {
    return add_external(s, X_CYLINDRICAL_RESTRAINT, g, per, n, index, {{axis3, 3}, {shift3, 3}, {k, 1}, {v0, 1}}, {}, "cylindrical_restraint");
}
int oracle_system_add_harmonic_com_pulling(ocm_system* s, int g, int per, double k, double v0, int64_t n1, const int32_t* g1, int64_t n2, const int32_t* g2) {
    if (!s) return fail("null system");
    if (g < 0 || g > 31) return fail("force group must be in 0..31");
    if (n1 <= 0 || n2 <= 0 || !g1 || !g2) return fail("harmonic_com_pulling: both bead groups must be non-empty");
    if (per && !s->has_box) return fail("harmonic_com_pulling: periodic term without a box");
    if (check_idx(s, g1, n1, "harmonic_com_pulling") || check_idx(s, g2, n2, "harmonic_com_pulling")) return 1;
    ComTerm t;
    t.group = g; t.periodic = per ? 1 : 0; t.k = k; t.v0 = v0;
    t.g1.assign(g1, g1 + n1); t.g2.assign(g2, g2 + n2);
    s->coms.push_back(std::move(t));
    return 0;
}
int oracle_system_add_combinatorial_go_contact(ocm_system* s, int g, const ocm_da_topology* topo, double k, double r0) {
    if (!(r0 > 0)) return fail("combinatorial_go_contact: r0 must be positive");
    return add_da(s, DA_COMB_GO, g, topo, 0, true, {}, k, r0, 0, "combinatorial_go_contact");
}
int oracle_system_add_3spn2_base_pair(ocm_system* s, int g, const ocm_da_topology* topo, const double* eps, const double* r0,
                                      const double* t01, const double* t02, const double* phi0, double alpha, double K_BP) {
    return add_da(s, DA_BASE_PAIR, g, topo, 5, true, {eps, r0, t01, t02, phi0}, alpha, K_BP, 0, "3spn2_base_pair");
}
int oracle_system_add_3spn2_cross_stacking(ocm_system* s, int g, const ocm_da_topology* topo, const double* eps, const double* r0,
                                           const double* t03, const double* t0CS, double K_BP, double K_CS, double alpha) {
    return add_da(s, DA_CROSS_STACK, g, topo, 4, false, {eps, r0, t03, t0CS}, K_BP, K_CS, alpha, "3spn2_cross_stacking");
}
int oracle_system_add_pdns(ocm_system* s, int g, const ocm_da_topology* topo, const double* k, const double* r0,
                           const double* theta0, const double* phi0, double sigma, double delta) {
    return add_da(s, DA_PDNS, g, topo, 4, true, {k, r0, theta0, phi0}, sigma, delta, 0, "pdns");
}
int oracle_num_da_terms(const ocm_context* c) { return c ? (int)c->sys.das.size() : -1; }
// (donor index, acceptor index) pairs of donor-acceptor term `term` evaluated at the current positions, sorted
int oracle_get_da_pairs(ocm_context* c, int term, int32_t* da, int64_t cap, int64_t* npairs) {
    if (!c || !npairs) return fail("null argument");
    if (term < 0 || term >= (int)c->sys.das.size()) return fail("donor-acceptor term index out of range");
    std::vector<std::pair<int32_t, int32_t>> out;
    for_each_da(*c, c->sys.das[term], [&](int64_t d, int64_t a) { out.emplace_back((int32_t)d, (int32_t)a); });
    *npairs = (int64_t)out.size();
    if (da) {
        if ((int64_t)out.size() > cap) return fail("get_da_pairs: capacity too small");
        for (size_t k = 0; k < out.size(); k++) { da[2 * k] = out[k].first; da[2 * k + 1] = out[k].second; }
    }
    return 0;
}

int oracle_system_add_excluded_volume(ocm_system* s, int g, double eps, double rc, const double* radius, const ocm_pair_topology* topo) {
    return add_pair(s, P_EXV, g, rc, rc, eps, std::pow(1.0 / rc, 12), 0, radius, nullptr, topo, "excluded_volume");
}
int oracle_system_add_debye_huckel(ocm_system* s, int g, double K, double debye, double rc, double cc, const double* charge, const ocm_pair_topology* topo) {
    return add_pair(s, P_DH, g, rc, rc, K, debye, cc, charge, nullptr, topo, "debye_huckel");
}
int oracle_system_add_lj_repulsive(ocm_system* s, int g, double rc, const double* eps, const double* sigma, const ocm_pair_topology* topo) {
    return add_pair(s, P_LJREP, g, rc, rc, 0, 0, 0, eps, sigma, topo, "lj_repulsive");
}
int oracle_system_add_isolf_attractive(ocm_system* s, int g, double rc, const double* eps, const double* sigma, const double* omega,
                                       const ocm_pair_topology* topo) {
    if (!eps || !sigma || !omega) return fail("isolf_attractive: null parameter array");
    return add_pair(s, P_ISOLF, g, rc, rc, 0, 0, 0, eps, sigma, topo, "isolf_attractive", omega);
}
int oracle_system_add_trigonometric(ocm_system* s, int g, double rc, const double* eps, const double* sigma, const double* omega,
                                    const ocm_pair_topology* topo) {
    if (!eps || !sigma || !omega) return fail("trigonometric: null parameter array");
    return add_pair(s, P_TRIG, g, rc, rc, 0, 0, 0, eps, sigma, topo, "trigonometric", omega);
}
int oracle_system_add_wca(ocm_system* s, int g, double rc, const double* eps, const double* sigma, const ocm_pair_topology* topo) {
    return add_pair(s, P_WCA, g, rc, rc, 0, 0, 0, eps, sigma, topo, "wca");
}
int oracle_system_add_lj_attractive(ocm_system* s, int g, double rc, double rsw, const double* eps, const double* sigma, const ocm_pair_topology* topo) {
    if (!(rsw < rc)) return fail("lj_attractive: r_switch must be below rc");
    return add_pair(s, P_LJATT, g, rc, rsw, 0, 0, 0, eps, sigma, topo, "lj_attractive");
}
int oracle_system_add_3spn2_excluded_volume(ocm_system* s, int g, double rc, const double* eps, const double* sigma, const ocm_pair_topology* topo) {
    return add_pair(s, P_3SPN2EXV, g, rc, rc, 0, 0, 0, eps, sigma, topo, "3spn2_excluded_volume");
}

int oracle_context_create(const ocm_system* sys, double T, double gamma, double dt, uint64_t seed, int device, ocm_context** out) {
    (void)device;
    if (!sys || !out) return fail("context_create: bad arguments");
    auto* c = new ocm_context();
    c->sys = *sys; c->T = T; c->gamma = gamma; c->dt = dt; c->seed = seed; c->step = 0;
    c->x.assign(sys->n, Vec3{0, 0, 0}); c->v.assign(sys->n, Vec3{0, 0, 0});
    c->force_brute = false;
    *out = c;
    return 0;
}
// Replica ensemble restated independently of the engine's cell keying: one concatenated system in which beads of
// different replicas never pair (pair_allowed), each replica with its own noise key and local bead counter.
int oracle_context_create_batch(const ocm_system* sys, int32_t nrep, double T, double gamma, double dt, const uint64_t* seeds,
                                int device, ocm_context** out) {
    (void)device;
    if (!sys || !out || !seeds || nrep < 1) return fail("context_create_batch: bad arguments");
    if (sys->has_box) return fail("context_create_batch: replica batches need an unlimited boundary (no box)");
    if (!sys->das.empty()) return fail("context_create_batch: donor-acceptor terms are not supported in replica batches yet");
    if (!sys->coms.empty()) return fail("context_create_batch: centre-of-mass pulling terms are not supported in replica batches");
    auto* c = new ocm_context();
    const int np = sys->n;
    ocm_system& S = c->sys;
    S.n = np * nrep; S.has_box = false; S.box[0] = S.box[1] = S.box[2] = 0;
    for (int r = 0; r < nrep; r++) S.mass.insert(S.mass.end(), sys->mass.begin(), sys->mass.end());
    for (const auto& t : sys->bonded) {
        BondedTerm u = t;
        u.idx.clear(); u.p.clear();
        for (int r = 0; r < nrep; r++) {
            for (int32_t v : t.idx) u.idx.push_back(v + r * np);
            u.p.insert(u.p.end(), t.p.begin(), t.p.end());
        }
        S.bonded.push_back(std::move(u));
    }
    for (const auto& t : sys->pairs) {
        PairTerm u = t;
        u.p0.clear(); u.p1.clear(); u.p2.clear(); u.part.clear(); u.excl.clear(); u.cls.clear();
        u.nper = np;
        for (int r = 0; r < nrep; r++) {
            u.p0.insert(u.p0.end(), t.p0.begin(), t.p0.end());
            u.p1.insert(u.p1.end(), t.p1.begin(), t.p1.end());
            u.p2.insert(u.p2.end(), t.p2.begin(), t.p2.end());
            u.cls.insert(u.cls.end(), t.cls.begin(), t.cls.end());
            for (int32_t v : t.part) u.part.push_back(v + r * np);
            for (const auto& e : t.excl) {
                std::vector<int32_t> sh;
                for (int32_t v : e) sh.push_back(v + r * np);
                u.excl.push_back(std::move(sh));
            }
        }
        S.pairs.push_back(std::move(u));
    }
    c->nrep = nrep; c->nper = np; c->rep_seeds.assign(seeds, seeds + nrep);
    c->base = std::make_shared<ocm_system>(*sys);
    c->T = T; c->gamma = gamma; c->dt = dt; c->seed = seeds[0]; c->step = 0;
    c->x.assign(S.n, Vec3{0, 0, 0}); c->v.assign(S.n, Vec3{0, 0, 0});
    c->force_brute = false;
    *out = c;
    return 0;
}
int oracle_num_replicas(const ocm_context* c) { return c ? c->nrep : -1; }

// ---- slab decomposition: the oracle restates the ownership rule of include/ocm_b200.h (rank r owns the beads whose
// fixed-point coordinate q on the axis has (q * world) >> 32 == r; halo = outside the slab within halo_nm of a face).
// Forces, energies and trajectories do not depend on the decomposition, so a slab context is a plain context.
int oracle_comm_unique_id(char*) { return fail("the oracle has no NCCL transport"); }
int oracle_comm_create_nccl(const char*, int, int, int, ocm_comm**) { return fail("the oracle has no NCCL transport"); }
int oracle_comm_create_local(int world, ocm_comm** out) {
    if (!out || world < 1) return fail("comm_create_local: bad arguments");
    *out = new ocm_comm{world};
    return 0;
}
void oracle_comm_destroy(ocm_comm* c) { delete c; }
int oracle_context_create_slab(const ocm_system* sys, double T, double gamma, double dt, uint64_t seed, int device, ocm_comm* comm,
                               int rank, int axis, ocm_context** out) {
    if (!comm) return fail("context_create_slab: null communicator");
    if (!sys || !sys->has_box) return fail("context_create_slab: a slab decomposition needs a periodic box");
    if (axis < 0 || axis > 2) return fail("context_create_slab: axis must be 0, 1 or 2");
    if (rank < 0 || rank >= comm->world) return fail("context_create_slab: rank out of range");
    if (oracle_context_create(sys, T, gamma, dt, seed, device, out)) return 1;
    (*out)->dd_rank = rank; (*out)->dd_world = comm->world; (*out)->dd_axis = axis;
    return 0;
}
int oracle_slab_info(const ocm_context* c, int* rank, int* world, int* axis, double* halo_nm) {
    if (!c) return fail("null context");
    if (rank) *rank = c->dd_rank;
    if (world) *world = c->dd_world;
    if (axis) *axis = c->dd_axis;
    if (halo_nm) *halo_nm = c->dd_halo;
    return 0;
}
int oracle_get_roles(ocm_context* c, unsigned char* role) {
    if (!c || !role) return fail("null argument");
    const int n = c->sys.n;
    if (c->dd_world <= 1) { for (int i = 0; i < n; i++) role[i] = 1; return 0; }
    Canon q = canonicalise(*c);
    const std::vector<uint32_t>& f = c->dd_axis == 0 ? q.fx : (c->dd_axis == 1 ? q.fy : q.fz);
    const double two32 = 4294967296.0, len = c->sys.box[c->dd_axis];
    const int W = c->dd_world, r = c->dd_rank;
    const uint32_t lo = (uint32_t)(uint64_t)std::ceil((double)r * two32 / W);
    const uint32_t hi = r + 1 == W ? 0u : (uint32_t)(uint64_t)std::ceil((double)(r + 1) * two32 / W);
    const int32_t hf = (int32_t)((uint32_t)std::ceil(c->dd_halo / len * two32) + 2u);
    for (int i = 0; i < n; i++) {
        const int owner = (int)(((uint64_t)f[i] * (uint64_t)W) >> 32);
        const int32_t below = (int32_t)(lo - 1u - f[i]), above = (int32_t)(f[i] - hi);
        const bool halo = (below >= 0 && below < hf) || (above >= 0 && above < hf);
        role[i] = owner == r ? 1 : (halo ? 2 : 0);
    }
    return 0;
}
void oracle_context_destroy(ocm_context* c) { delete c; }

int oracle_context_set_option(ocm_context* c, const char* key, double v) {
    if (!c) return fail("null context");
    if (!strcmp(key, "brute_force")) { c->force_brute = v != 0; return 0; }
    if (!strcmp(key, "cpu_verlet")) { c->vl_enabled = v != 0; c->vl_valid = false; return 0; }       // Verlet lists in oracle_step (default on)
    if (!strcmp(key, "cpu_skin_nm")) { if (v > 0) c->vl_skin = v; c->vl_valid = false; return 0; }
    if (!strcmp(key, "slab_halo_nm")) { c->dd_halo = v; return 0; }
    if (!strcmp(key, "virtual_box_x")) { c->vbox[0] = v; return 0; }
    if (!strcmp(key, "virtual_box_y")) { c->vbox[1] = v; return 0; }
    if (!strcmp(key, "virtual_box_z")) { c->vbox[2] = v; return 0; }
    if (!strcmp(key, "threads")) {
#ifdef _OPENMP
        omp_set_num_threads((int)v);
#endif
        return 0;
    }
    return 0;  // engine tunables are meaningless for the oracle
}
int oracle_context_get_option(ocm_context* c, const char* key, double* v) {
    if (!c || !v) return fail("null argument");
    if (!strcmp(key, "cpu_verlet_builds")) { *v = (double)c->vl_builds; return 0; }
    if (!strcmp(key, "threads")) {
#ifdef _OPENMP
        *v = omp_get_max_threads();
#else
        *v = 1;
#endif
        return 0;
    }
    *v = 0; return 0;
}

int oracle_set_positions(ocm_context* c, const double* xyz) {
    if (!c || !xyz) return fail("null argument");
    for (int i = 0; i < c->sys.n; i++) c->x[i] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    return 0;
}
int oracle_set_velocities(ocm_context* c, const double* xyz) {
    if (!c || !xyz) return fail("null argument");
    for (int i = 0; i < c->sys.n; i++) c->v[i] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    return 0;
}
// v = sqrt(kT/m) N(0,1), no COM removal.  Noise stream 1 of the shared Philox.
int oracle_set_velocities_to_temperature(ocm_context* c, double T, uint64_t seed) {
    if (!c) return fail("null context");
    for (int i = 0; i < c->sys.n; i++) {
        double g[3];
        const int rep = c->nrep > 1 ? i / c->nper : 0;
        normals3(seed + (uint64_t)rep, (uint32_t)(i - rep * c->nper), 0, 1, g);
        double m = c->sys.mass[i];
        double s = m > 0 ? std::sqrt(kBoltz * T / m) : 0.0;
        c->v[i] = {s * g[0], s * g[1], s * g[2]};
    }
    return 0;
}

// SURVEY.md B.1 (OpenMM LangevinIntegrator, leap-frog)
// ---- Verlet-list force evaluation for oracle_step on large systems (the CPU arm of bench.py) -------------------
// Same pair set as the cell-list path (every pair inside the cutoff is in the list and is tested against the cutoff),
// so the same forces up to the order of double-precision additions.
static bool verlet_applicable(const ocm_context& c) {
    if (!c.vl_enabled || c.force_brute || c.sys.pairs.empty()) return false;
    size_t big = 0;
    for (const auto& t : c.sys.pairs) big = std::max(big, t.part.size());
    return big >= 1500;
}

static void verlet_build(ocm_context& c) {
    c.vgroups.clear();
    // group the pair terms by topology
    for (int ti = 0; ti < (int)c.sys.pairs.size(); ti++) {
        const PairTerm& t = c.sys.pairs[ti];
        bool placed = false;
        for (auto& g : c.vgroups) {
            const PairTerm& o = c.sys.pairs[g.terms[0]];
            if (o.periodic == t.periodic && o.part == t.part && o.excl == t.excl && o.cls == t.cls && o.allow == t.allow && o.nper == t.nper) {
                g.terms.push_back(ti); placed = true; break;
            }
        }
        if (!placed) { ocm_context::VerletGroup g; g.terms.push_back(ti); c.vgroups.push_back(std::move(g)); }
    }
    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
    for (auto& g : c.vgroups) {
        const PairTerm& t0 = c.sys.pairs[g.terms[0]];
        double rc = 0;
        for (int ti : g.terms) rc = std::max(rc, effective_rc(c.sys.pairs[ti]));
        g.rlist = rc + c.vl_skin;
        if (t0.periodic) for (int k = 0; k < 3; k++) g.rlist = std::min(g.rlist, 0.4999 * c.sys.box[k]);
        std::vector<std::vector<std::pair<int, int>>> found(nthreads);
        for_each_pair(c, t0, g.rlist, true, [&](int i, int j, Vec3, double) {
            if (!pair_allowed(t0, i, j)) return;
            int tid = 0;
#ifdef _OPENMP
            tid = omp_get_thread_num();
#endif
            found[tid].emplace_back(i, j);
        });
        std::vector<int> row_of(c.sys.n, -1);
        g.beads.assign(t0.part.begin(), t0.part.end());
        for (size_t a = 0; a < g.beads.size(); a++) row_of[g.beads[a]] = (int)a;
        g.off.assign(g.beads.size() + 1, 0);
        for (auto& v : found) for (auto& pr : v) g.off[row_of[pr.first] + 1]++;
        for (size_t a = 0; a < g.beads.size(); a++) g.off[a + 1] += g.off[a];
        g.nbr.resize(g.off.back());
        std::vector<int> cur(g.off.begin(), g.off.end() - 1);
        for (auto& v : found) for (auto& pr : v) g.nbr[cur[row_of[pr.first]]++] = pr.second;
    }
    c.vl_xref = c.x;
    c.vl_valid = true;
    c.vl_builds++;
    if ((int)c.vl_F.size() != nthreads || (c.vl_F.size() && (int)c.vl_F[0].size() != c.sys.n)) c.vl_F.assign(nthreads, std::vector<Vec3>(c.sys.n, Vec3{0, 0, 0}));
}

static void compute_forces_verlet(ocm_context& c, std::vector<Vec3>* F) {
    const int n = c.sys.n;
    // rebuild when a bead has moved more than half the skin since the lists were made
    bool stale = !c.vl_valid;
    if (!stale) {
        const double lim = 0.25 * c.vl_skin * c.vl_skin;
        int moved = 0;
#pragma omp parallel for reduction(| : moved)
        for (int i = 0; i < n; i++) { const Vec3 d = c.x[i] - c.vl_xref[i]; if (dot(d, d) > lim) moved |= 1; }
        stale = moved != 0;
    }
    if (stale) verlet_build(c);
    F->assign(n, Vec3{0, 0, 0});
    for (const auto& t : c.sys.bonded) eval_bonded(c, t, F, nullptr);
    for (const auto& t : c.sys.das) eval_da(c, t, F, nullptr);
    for (const auto& t : c.sys.coms) eval_com(c, t, F, nullptr);
    const int nthreads = (int)c.vl_F.size();
#pragma omp parallel
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        std::vector<Vec3>& Fl = c.vl_F[tid];
        for (auto& g : c.vgroups) {
            const PairTerm& t0 = c.sys.pairs[g.terms[0]];
            const bool per = t0.periodic != 0;
            const double* box = c.sys.box;
            const double ibx = per ? 1.0 / box[0] : 0, iby = per ? 1.0 / box[1] : 0, ibz = per ? 1.0 / box[2] : 0;
            double rc2max = 0;
            for (int ti : g.terms) rc2max = std::max(rc2max, c.sys.pairs[ti].rc * c.sys.pairs[ti].rc);
#pragma omp for schedule(dynamic, 32) nowait   // (the timed CPU arm: dynamic chunks are 20 % faster than static ones; partial sums may differ in the last bit from run to run)
            for (int a = 0; a < (int)g.beads.size(); a++) {
                const int i = g.beads[a];
                const Vec3 xi = c.x[i];
                Vec3 fi{0, 0, 0};
                for (int q = g.off[a]; q < g.off[a + 1]; q++) {
                    const int j = g.nbr[q];
                    Vec3 d = xi - c.x[j];
                    if (per) { d.x -= box[0] * std::nearbyint(d.x * ibx); d.y -= box[1] * std::nearbyint(d.y * iby); d.z -= box[2] * std::nearbyint(d.z * ibz); }
                    const double r2 = dot(d, d);
                    if (!(r2 < rc2max)) continue;
                    const double r = std::sqrt(r2);
                    double dEsum = 0;
                    for (int ti : g.terms) {
                        const PairTerm& t = c.sys.pairs[ti];
                        if (!(r < t.rc)) continue;
                        double E, dE;
                        pair_energy(t, i, j, r, &E, &dE);
                        dEsum += dE;
                    }
                    const Vec3 f = d * (-dEsum / r);
                    fi = fi + f;
                    Fl[j] = Fl[j] - f;
                }
                Fl[i] = Fl[i] + fi;
            }
        }
#pragma omp barrier
        // reduce the per-thread slabs in parallel over the beads, and clear them for the next call
#pragma omp for schedule(static)
        for (int i = 0; i < n; i++) {
            Vec3 s = (*F)[i];
            for (int k = 0; k < nthreads; k++) { s = s + c.vl_F[k][i]; c.vl_F[k][i] = Vec3{0, 0, 0}; }
            (*F)[i] = s;
        }
    }
}

// ---- Monte Carlo barostats -------------------------------------------------------------------------------------
// OpenMM 8.0.0 is not under /root/reference (external dependency, CI pins 8.0.0); this restates the published
// algorithm of MonteCarloAnisotropicBarostatImpl / MonteCarloMembraneBarostatImpl::updateContextState and of
// ReferenceMonteCarloBarostat::applyBarostat as the reference's generators configure them
// (MonteCarloAnisotropicBarostatGenerator.cpp:9-17, MonteCarloMembraneBarostatGenerator.cpp:6-14, SystemGenerator.cpp:89-108).
// PARITY UNPINNED for this part: nothing in /root/reference holds a barostat trajectory, and OpenMM's SFMT stream is replaced
// by SplitMix64 (the engine draws the same stream, so the two take the same decisions).
namespace {
double baro_uniform(ocm_context& c) {
    c.baro_rng += 0x9E3779B97F4A7C15ull;
    uint64_t z = c.baro_rng;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (double)(z >> 11) / 9007199254740992.0;
}
double total_potential(ocm_context& c) {
    double Eg[32];
    compute(c, nullptr, Eg, 0xFFFFFFFFu);
    double e = 0;
    for (int g = 0; g < 32; g++) e += Eg[g];
    return e;
}
// molecules = sets of beads connected through bonded terms (ContextImpl::getMolecules over getBondedParticles)
void find_molecules(ocm_context& c) {
    const int n = c.sys.n;
    std::vector<std::vector<int>> adj(n);
    for (const auto& t : c.sys.bonded)
        for (int64_t m = 0; m < t.n(); m++)
            for (int b = 0; b + 1 < t.nbody; b++) {
                const int i = t.idx[m * t.nbody + b], j = t.idx[m * t.nbody + b + 1];
                adj[i].push_back(j); adj[j].push_back(i);
            }
    std::vector<int> mol(n, -1);
    c.molecules.clear();
    for (int i = 0; i < n; i++) {
        if (mol[i] >= 0) continue;
        const int id = (int)c.molecules.size();
        std::vector<int> stack(1, i), members;
        mol[i] = id;
        while (!stack.empty()) {
            const int a = stack.back(); stack.pop_back();
            members.push_back(a);
            for (int b : adj[a]) if (mol[b] < 0) { mol[b] = id; stack.push_back(b); }
        }
        std::sort(members.begin(), members.end());
        c.molecules.push_back(members);
    }
}
void barostat_trial(ocm_context& c) {
    const double bar = 6.02214076e23 * 1e-25;  // kJ/mol/nm^3 per bar
    const double E_before = total_potential(c);
    const bool membrane = c.baro_kind == 2;
    int axis = -1;
    while (axis < 0) {
        const double r = 3.0 * baro_uniform(c);
        const int pick = r < 1.0 ? 0 : (r < 2.0 ? 1 : 2);
        bool allowed;
        if (!membrane) allowed = c.baro_mode[pick] != 0;
        else allowed = pick == 0 || (pick == 1 && c.baro_mode[0] == 1) || (pick == 2 && c.baro_mode[1] == 0);
        if (allowed) axis = pick;
    }
    const double old_box[3] = {c.sys.box[0], c.sys.box[1], c.sys.box[2]};
    const double V = old_box[0] * old_box[1] * old_box[2];
    double dV = c.baro_vscale[axis] * 2 * (baro_uniform(c) - 0.5);
    double Vn = V + dV;
    double ls[3] = {1, 1, 1};
    if (membrane && c.baro_mode[0] == 0 && axis == 0) ls[0] = ls[1] = std::sqrt(Vn / V);
    else ls[axis] = Vn / V;
    if (membrane && c.baro_mode[1] == 2) { ls[2] = 1.0 / (ls[0] * ls[1]); Vn = V; dV = 0; }
    const std::vector<Vec3> saved = c.x;
    for (const auto& mol : c.molecules) {
        Vec3 ctr = {0, 0, 0};
        for (int i : mol) ctr = ctr + c.x[i];
        const double cnt = (double)mol.size();
        ctr = {ctr.x / cnt, ctr.y / cnt, ctr.z / cnt};
        const Vec3 wr = {ctr.x - std::floor(ctr.x / old_box[0]) * old_box[0], ctr.y - std::floor(ctr.y / old_box[1]) * old_box[1],
                         ctr.z - std::floor(ctr.z / old_box[2]) * old_box[2]};
        const Vec3 shift = {wr.x * ls[0] - ctr.x, wr.y * ls[1] - ctr.y, wr.z * ls[2] - ctr.z};
        for (int i : mol) c.x[i] = c.x[i] + shift;
    }
    for (int k = 0; k < 3; k++) c.sys.box[k] = old_box[k] * ls[k];
    const double dA = c.sys.box[0] * c.sys.box[1] - old_box[0] * old_box[1];
    const double E_after = total_potential(c);
    const double kT = kBoltz * c.baro_T;
    const double P = (membrane ? c.baro_pressure[0] : c.baro_pressure[axis]) * bar;
    const double gam = membrane ? c.baro_tension * bar : 0.0;
    const double w = E_after - E_before + P * dV - gam * dA - (double)c.molecules.size() * kT * std::log(Vn / V);
    bool ok = true;
    if (w > 0 && baro_uniform(c) > std::exp(-w / kT)) {
        ok = false;
        c.x = saved;
        for (int k = 0; k < 3; k++) c.sys.box[k] = old_box[k];
    }
    c.baro_attempted[axis]++;
    if (ok) c.baro_accepted[axis]++;
    if (c.baro_attempted[axis] >= 10) {
        if (c.baro_accepted[axis] < 0.25 * c.baro_attempted[axis]) {
            c.baro_vscale[axis] /= 1.1;
            c.baro_attempted[axis] = c.baro_accepted[axis] = 0;
        } else if (c.baro_accepted[axis] > 0.75 * c.baro_attempted[axis]) {
            c.baro_vscale[axis] = std::min(c.baro_vscale[axis] * 1.1, V * 0.3);
            c.baro_attempted[axis] = c.baro_accepted[axis] = 0;
        }
    }
    c.baro_trials += 1; c.baro_accepts += ok ? 1 : 0;
    c.baro_last[0] = axis; c.baro_last[1] = dV; c.baro_last[2] = E_before; c.baro_last[3] = E_after; c.baro_last[4] = w; c.baro_last[5] = ok ? 1 : 0;
    c.vl_valid = false;
}
}  // namespace

int oracle_context_set_barostat(ocm_context* c, int kind, const double* pressure3, double tension, double T, const int32_t* mode3, int32_t frequency,
                                uint64_t seed) {
    if (!c) return fail("null context");
    if (kind == 0) { c->baro_kind = 0; return 0; }
    if (kind != 1 && kind != 2) return fail("set_barostat: unknown kind");
    if (!pressure3 || !mode3) return fail("set_barostat: null argument");
    if (!c->sys.has_box) return fail("barostat should be used with periodic boundary condition.");
    if (c->nrep > 1 || c->dd_world > 1) return fail("set_barostat: single periodic systems only");
    if (frequency < 0) return fail("set_barostat: negative frequency");
    if (!(T > 0)) return fail("set_barostat: temperature must be positive");
    if (kind == 1 && !mode3[0] && !mode3[1] && !mode3[2]) frequency = 0;
    if (kind == 2 && (mode3[0] < 0 || mode3[0] > 1 || mode3[1] < 0 || mode3[1] > 2)) return fail("set_barostat: unknown xymode / zmode");
    c->baro_kind = kind; c->baro_freq = frequency; c->baro_T = T; c->baro_tension = tension; c->baro_rng = seed; c->baro_calls = 0;
    const double V = c->sys.box[0] * c->sys.box[1] * c->sys.box[2];
    for (int k = 0; k < 3; k++) {
        c->baro_pressure[k] = pressure3[kind == 2 ? 0 : k]; c->baro_mode[k] = mode3[k];
        c->baro_vscale[k] = 0.01 * V; c->baro_attempted[k] = c->baro_accepted[k] = 0;
    }
    c->baro_trials = c->baro_accepts = 0; c->baro_last[0] = -1;
    find_molecules(*c);
    return 0;
}
int oracle_get_box(ocm_context* c, double* box3) {
    if (!c || !box3) return fail("null argument");
    for (int k = 0; k < 3; k++) box3[k] = c->sys.has_box ? c->sys.box[k] : 0.0;
    return 0;
}
int oracle_set_box(ocm_context* c, const double* box3) {
    if (!c || !box3) return fail("null argument");
    if (!c->sys.has_box) return fail("set_box: the context has no periodic box");
    for (int k = 0; k < 3; k++) if (!(box3[k] > 0)) return fail("set_box: box edge must be positive");
    for (int k = 0; k < 3; k++) c->sys.box[k] = box3[k];
    c->vl_valid = false;
    return 0;
}
int oracle_get_barostat_info(ocm_context* c, double* out) {
    if (!c || !out) return fail("null argument");
    out[0] = c->baro_trials; out[1] = c->baro_accepts;
    for (int k = 0; k < 6; k++) out[2 + k] = c->baro_last[k];
    for (int k = 0; k < 3; k++) out[8 + k] = c->baro_vscale[k];
    out[11] = (double)c->molecules.size();
    return 0;
}

int oracle_step(ocm_context* c, int64_t nsteps) {
    if (!c) return fail("null context");
    const double a = std::exp(-c->gamma * c->dt);
    const double b = c->gamma == 0 ? c->dt : (1.0 - a) / c->gamma;
    const double cc = std::sqrt(kBoltz * c->T * (1.0 - a * a));
    std::vector<Vec3> F;
    const bool verlet = verlet_applicable(*c);
    c->vl_valid = false;  // (positions may have been set since the last call)
    for (int64_t s = 0; s < nsteps; s++) {
        // ContextImpl::updateContextState before every step: the barostat counts the calls and makes a trial at every frequency-th
        if (c->baro_kind && c->baro_freq > 0 && ++c->baro_calls >= c->baro_freq) { c->baro_calls = 0; barostat_trial(*c); }
        if (verlet) compute_forces_verlet(*c, &F);
        else compute(*c, &F, nullptr, 0xFFFFFFFFu);
        for (int i = 0; i < c->sys.n; i++) {
            double m = c->sys.mass[i];
            if (!(m > 0)) continue;
            double g[3];
            const int rep = c->nrep > 1 ? i / c->nper : 0;
            normals3(c->nrep > 1 ? c->rep_seeds[rep] : c->seed, (uint32_t)(i - rep * c->nper), (uint64_t)c->step, 0, g);
            double im = 1.0 / m, sq = std::sqrt(im);
            Vec3& v = c->v[i];
            v.x = a * v.x + b * F[i].x * im + cc * sq * g[0];
            v.y = a * v.y + b * F[i].y * im + cc * sq * g[1];
            v.z = a * v.z + b * F[i].z * im + cc * sq * g[2];
            c->x[i] = c->x[i] + v * c->dt;
        }
        c->step++;
    }
    return 0;
}
int64_t oracle_current_step(const ocm_context* c) { return c ? c->step : -1; }

// LocalEnergyMinimizer::minimize (src/Simulator.cpp:43-46): L-BFGS (history 8) with step halving, until the RMS
// force per particle is below `tolerance` (kJ/mol/nm).  Written independently of the engine's host/lbfgs.hpp; the
// tests compare outcomes (a stationary point below the tolerance), not iterates.
int oracle_minimize(ocm_context* c, double tolerance, int max_iterations, int* iterations, double* final_energy) {
    if (!c) return fail("null context");
    if (!(tolerance > 0)) return fail("minimize: tolerance must be positive");
    const int n = c->sys.n;
    const size_t N = (size_t)3 * n;
    auto energy_grad = [&](const std::vector<double>& at, std::vector<double>& grad) {
        for (int i = 0; i < n; i++) c->x[i] = {at[3 * i], at[3 * i + 1], at[3 * i + 2]};
        std::vector<Vec3> F; double Eg[32];
        compute(*c, &F, Eg, 0xFFFFFFFFu);
        double e = 0;
        for (int g = 0; g < 32; g++) e += Eg[g];
        for (int i = 0; i < n; i++) {
            const bool mobile = c->sys.mass[i] > 0;
            grad[3 * i] = mobile ? -F[i].x : 0; grad[3 * i + 1] = mobile ? -F[i].y : 0; grad[3 * i + 2] = mobile ? -F[i].z : 0;
        }
        return e;
    };
    std::vector<double> x(N), g(N), xt(N), gt(N), q(N);
    for (int i = 0; i < n; i++) { x[3 * i] = c->x[i].x; x[3 * i + 1] = c->x[i].y; x[3 * i + 2] = c->x[i].z; }
    double E = energy_grad(x, g);
    std::vector<std::vector<double>> dx, dg;
    const int hist = 8, itmax = max_iterations > 0 ? max_iterations : 1000000;
    int it = 0, stalls = 0;
    for (; it < itmax; it++) {
        double gg = 0;
        for (double v : g) gg += v * v;
        if (std::sqrt(gg / std::max(1, n)) < tolerance) break;
        q = g;
        std::vector<double> a(dx.size());
        for (int k = (int)dx.size() - 1; k >= 0; k--) {
            double sy = 0, sq = 0;
            for (size_t i = 0; i < N; i++) { sy += dx[k][i] * dg[k][i]; sq += dx[k][i] * q[i]; }
            a[k] = sq / sy;
            for (size_t i = 0; i < N; i++) q[i] -= a[k] * dg[k][i];
        }
        double scale = 0.01 / std::sqrt(std::max(gg, 1e-300));
        if (!dx.empty()) {
            double sy = 0, yy = 0;
            for (size_t i = 0; i < N; i++) { sy += dx.back()[i] * dg.back()[i]; yy += dg.back()[i] * dg.back()[i]; }
            scale = sy / yy;
        }
        for (double& v : q) v *= scale;
        for (size_t k = 0; k < dx.size(); k++) {
            double sy = 0, yq = 0;
            for (size_t i = 0; i < N; i++) { sy += dx[k][i] * dg[k][i]; yq += dg[k][i] * q[i]; }
            const double b = yq / sy;
            for (size_t i = 0; i < N; i++) q[i] += dx[k][i] * (a[k] - b);
        }
        double slope = 0;
        for (size_t i = 0; i < N; i++) slope += g[i] * q[i];
        bool moved = false;
        if (slope > 0) {
            double t = 1.0;
            for (int h = 0; h < 50 && !moved; h++, t *= 0.5) {
                for (size_t i = 0; i < N; i++) xt[i] = x[i] - t * q[i];
                const double Et = energy_grad(xt, gt);
                if (std::isfinite(Et) && Et <= E - 1e-4 * t * slope) {
                    std::vector<double> s(N), y(N);
                    double sy = 0;
                    for (size_t i = 0; i < N; i++) { s[i] = xt[i] - x[i]; y[i] = gt[i] - g[i]; sy += s[i] * y[i]; }
                    if (sy > 1e-12) { dx.push_back(s); dg.push_back(y); if ((int)dx.size() > hist) { dx.erase(dx.begin()); dg.erase(dg.begin()); } }
                    x = xt; g = gt; E = Et; moved = true;
                }
            }
        }
        if (!moved) {
            if (dx.empty() || ++stalls > 50) break;
            dx.clear(); dg.clear();
        }
    }
    for (int i = 0; i < n; i++) c->x[i] = {x[3 * i], x[3 * i + 1], x[3 * i + 2]};
    if (iterations) *iterations = it;
    if (final_energy) *final_energy = E;
    return 0;
}

int oracle_get_state(ocm_context* c, uint32_t mask, double* pos, double* vel, double* force, double* epot, double* ekin, double* ebg) {
    if (!c) return fail("null context");
    int n = c->sys.n;
    if (pos) for (int i = 0; i < n; i++) { pos[3 * i] = c->x[i].x; pos[3 * i + 1] = c->x[i].y; pos[3 * i + 2] = c->x[i].z; }
    if (vel) for (int i = 0; i < n; i++) { vel[3 * i] = c->v[i].x; vel[3 * i + 1] = c->v[i].y; vel[3 * i + 2] = c->v[i].z; }
    if (force || epot || ebg) {
        std::vector<Vec3> F; double Eg[32];
        compute(*c, force ? &F : nullptr, Eg, mask);
        if (force) for (int i = 0; i < n; i++) { force[3 * i] = F[i].x; force[3 * i + 1] = F[i].y; force[3 * i + 2] = F[i].z; }
        if (epot) { double e = 0; for (int g = 0; g < 32; g++) e += Eg[g]; *epot = e; }
        if (ebg) std::copy(Eg, Eg + 32, ebg);
    }
    if (ekin) {
        // leap-frog KE shifted by dt/2 with the current (all-group) forces, SURVEY.md B.1
        std::vector<Vec3> F;
        compute(*c, &F, nullptr, 0xFFFFFFFFu);
        double ke = 0;
        for (int i = 0; i < n; i++) {
            double m = c->sys.mass[i];
            if (!(m > 0)) continue;
            Vec3 vs = c->v[i] + F[i] * (0.5 * c->dt / m);
            ke += 0.5 * m * dot(vs, vs);
        }
        *ekin = ke;
    }
    return 0;
}

// per-replica energies: every replica is evaluated as the stand-alone system it is
int oracle_get_replica_energies(ocm_context* c, uint32_t mask, double* epot, double* ekin, double* ebg) {
    if (!c) return fail("null context");
    for (int r = 0; r < c->nrep; r++) {
        ocm_context one;
        one.sys = c->nrep > 1 ? *c->base : c->sys;
        one.T = c->T; one.gamma = c->gamma; one.dt = c->dt; one.seed = c->seed; one.step = c->step; one.force_brute = c->force_brute;
        const int np = one.sys.n;
        one.x.assign(c->x.begin() + (size_t)r * np, c->x.begin() + (size_t)(r + 1) * np);
        one.v.assign(c->v.begin() + (size_t)r * np, c->v.begin() + (size_t)(r + 1) * np);
        double ep = 0, ek = 0, eg[32];
        if (oracle_get_state(&one, mask, nullptr, nullptr, nullptr, &ep, ekin ? &ek : nullptr, eg)) return 1;
        if (epot) epot[r] = ep;
        if (ekin) ekin[r] = ek;
        if (ebg) std::copy(eg, eg + 32, ebg + (size_t)32 * r);
    }
    return 0;
}

int oracle_num_pair_terms(const ocm_context* c) { return c ? (int)c->sys.pairs.size() : -1; }

// pair set with the canonical FP32 / fixed-point predicate (bit-exact contract with the engine)
int oracle_get_neighbor_pairs(ocm_context* c, int term, int32_t* ij, int64_t cap, int64_t* npairs) {
    if (!c || !npairs) return fail("null argument");
    if (term < 0 || term >= (int)c->sys.pairs.size()) return fail("pair term index out of range");
    const PairTerm& t = c->sys.pairs[term];
    Canon q = canonicalise(*c);
    double rce = effective_rc(t);
    float rc2 = (float)(rce * rce);
    std::vector<std::pair<int, int>> out;
    // candidates from a double-precision search with a margin, decided by the canonical predicate
    std::vector<std::vector<std::pair<int, int>>> loc(1);
#ifdef _OPENMP
    loc.resize(omp_get_max_threads());
#endif
    for_each_pair(*c, t, rce * 1.001 + 1e-4, true, [&](int i, int j, Vec3, double) {
        if (!pair_allowed(t, i, j)) return;
        if (canon_r2(q, i, j) < rc2) {
            int tid = 0;
#ifdef _OPENMP
            tid = omp_get_thread_num();
#endif
            loc[tid].push_back({i, j});
        }
    });
    for (auto& v : loc) out.insert(out.end(), v.begin(), v.end());
    std::sort(out.begin(), out.end());
    *npairs = (int64_t)out.size();
    if (ij) {
        if (cap < (int64_t)out.size()) return fail("get_neighbor_pairs: capacity too small");
        for (size_t k = 0; k < out.size(); k++) { ij[2 * k] = out[k].first; ij[2 * k + 1] = out[k].second; }
    }
    return 0;
}

// Canonical cell assignment, shared bit-exactly with the engine: the grid is derived from the
// term's list cutoff (effective rc + skin); see DESIGN.md "cell grid".
// rlist = list cutoff (interaction cutoff of the list serving the term + skin), reported by the engine
// through ocm_get_counter("list_cutoff_nm.<term>").
static int cell_ids_grid(ocm_context* c, int term, float wmin, long long cap, int32_t* cell_of_bead, int32_t dims[3]);

int oracle_get_cell_ids_rlist(ocm_context* c, int term, double rlist, int32_t* cell_of_bead, int32_t dims[3]) {
    if (!c || !cell_of_bead || !dims) return fail("null argument");
    if (term < 0 || term >= (int)c->sys.pairs.size()) return fail("pair term index out of range");
    // Verlet-row lists: cells of half the list radius; only the (mostly empty) virtual box of an unlimited-boundary system
    // widens them, until their number is of the order of the participants
    const long long cap = c->sys.has_box ? -1 : std::max<long long>(4096, 8 * (long long)c->sys.pairs[term].part.size() + 64);
    return cell_ids_grid(c, term, 0.5f * (float)rlist, cap, cell_of_bead, dims);
}

// The grid rule with its two inputs given explicitly (the engine reports them: counters cell_width_nm.<term> and
// cell_cap.<term>): cells of width >= wmin, widened by factors of 1.25 until there are at most `cap` of them (cap < 0: never).
// Tile lists use cells of a sixth of the list radius capped at one cell per participant.
int oracle_get_cell_ids_grid(ocm_context* c, int term, double wmin, long long cap, int32_t* cell_of_bead, int32_t dims[3]) {
    if (!c || !cell_of_bead || !dims) return fail("null argument");
    if (term < 0 || term >= (int)c->sys.pairs.size()) return fail("pair term index out of range");
    return cell_ids_grid(c, term, (float)wmin, cap, cell_of_bead, dims);
}

static int cell_ids_grid(ocm_context* c, int term, float wmin, long long cap, int32_t* cell_of_bead, int32_t dims[3]) {
    const PairTerm& t = c->sys.pairs[term];
    Canon q = canonicalise(*c);
    int n = c->sys.n;
    for (int i = 0; i < n; i++) cell_of_bead[i] = -1;
    if (q.periodic) {
        int nc[3];
        float w = wmin;
        for (;;) {
            long long tot = 1;
            for (int k = 0; k < 3; k++) {
                nc[k] = (int)floorf((float)(c->sys.has_box ? c->sys.box[k] : c->vbox[k]) / w);
                nc[k] = std::max(1, std::min(nc[k], 1024));
                dims[k] = nc[k];
                tot *= nc[k];
            }
            if (cap < 0 || tot <= cap) break;
            w *= 1.25f;
        }
        for (int i : t.part) {
            uint32_t cx = (uint32_t)(((uint64_t)q.fx[i] * (uint64_t)nc[0]) >> 32);
            uint32_t cy = (uint32_t)(((uint64_t)q.fy[i] * (uint64_t)nc[1]) >> 32);
            uint32_t cz = (uint32_t)(((uint64_t)q.fz[i] * (uint64_t)nc[2]) >> 32);
            cell_of_bead[i] = (int32_t)((cz * nc[1] + cy) * nc[0] + cx);
        }
    } else {
        float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int i : t.part) {
            float p[3] = {q.px[i], q.py[i], q.pz[i]};
            for (int k = 0; k < 3; k++) { lo[k] = fminf(lo[k], p[k]); hi[k] = fmaxf(hi[k], p[k]); }
        }
        int nc[3]; float inv[3];
        for (int k = 0; k < 3; k++) {
            float ext = hi[k] - lo[k];
            nc[k] = (int)floorf(ext / wmin) + 1;
            nc[k] = std::max(1, std::min(nc[k], 1024));
            dims[k] = nc[k];
            inv[k] = 1.0f / wmin;
        }
        for (int i : t.part) {
            float p[3] = {q.px[i], q.py[i], q.pz[i]};
            int ci[3];
            for (int k = 0; k < 3; k++) {
                int v = (int)floorf((p[k] - lo[k]) * inv[k]);
                ci[k] = std::max(0, std::min(nc[k] - 1, v));
            }
            cell_of_bead[i] = (ci[2] * nc[1] + ci[1]) * nc[0] + ci[0];
        }
    }
    return 0;
}

int oracle_get_cell_ids(ocm_context* c, int term, int32_t* cell_of_bead, int32_t dims[3]) {
    if (!c) return fail("null argument");
    if (term < 0 || term >= (int)c->sys.pairs.size()) return fail("pair term index out of range");
    return oracle_get_cell_ids_rlist(c, term, effective_rc(c->sys.pairs[term]) + 0.3, cell_of_bead, dims);
}

// the stall-free observer calls of the engine (ocm_get_state_begin / _end, ocm_set_state_async) restated: the oracle has no
// device to overlap with, so begin snapshots synchronously and end hands the snapshot out
struct ocm_state_request {
    std::vector<double> pos, vel, force, ebg;
    double epot = 0, ekin = 0;
    bool has_energy = false;
};
int oracle_get_state_begin(ocm_context* c, uint32_t mask, int want_pos, int want_vel, int want_force, int want_energy, ocm_state_request** out) {
    if (!c || !out) return fail("null argument");
    auto* rq = new ocm_state_request();
    const size_t n3 = (size_t)3 * c->sys.n;
    if (want_pos) rq->pos.resize(n3);
    if (want_vel) rq->vel.resize(n3);
    if (want_force) rq->force.resize(n3);
    rq->ebg.resize(32);
    rq->has_energy = want_energy != 0;
    int rc = oracle_get_state(c, mask, want_pos ? rq->pos.data() : nullptr, want_vel ? rq->vel.data() : nullptr, want_force ? rq->force.data() : nullptr,
                              want_energy ? &rq->epot : nullptr, want_energy ? &rq->ekin : nullptr, want_energy ? rq->ebg.data() : nullptr);
    if (rc) { delete rq; return rc; }
    *out = rq;
    return 0;
}
int oracle_get_state_end(ocm_state_request* rq, double* pos, double* vel, double* force, double* epot, double* ekin, double* ebg) {
    if (!rq) return fail("null request");
    if (pos && !rq->pos.empty()) std::copy(rq->pos.begin(), rq->pos.end(), pos);
    if (vel && !rq->vel.empty()) std::copy(rq->vel.begin(), rq->vel.end(), vel);
    if (force && !rq->force.empty()) std::copy(rq->force.begin(), rq->force.end(), force);
    if (rq->has_energy) {
        if (epot) *epot = rq->epot;
        if (ekin) *ekin = rq->ekin;
        if (ebg) std::copy(rq->ebg.begin(), rq->ebg.end(), ebg);
    }
    delete rq;
    return 0;
}
int oracle_set_state_async(ocm_context* c, const double* xyz, const double* vel) {
    if (int rc = oracle_set_positions(c, xyz)) return rc;
    return vel ? oracle_set_velocities(c, vel) : 0;
}

int oracle_get_counter(ocm_context*, const char*, double* v) { if (v) *v = 0; return 0; }

int oracle_profile_kernel(ocm_context*, const char*, int, double* v) { if (v) *v = 0; return fail("the oracle has no kernels"); }
int oracle_measure_fp32_peak(int, double* v) { if (v) *v = 0; return fail("the oracle has no device"); }

}  // extern "C"
