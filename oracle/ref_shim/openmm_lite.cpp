// openmm_lite.cpp — implementation of oracle/ref_shim/OpenMM.h (TEST INFRASTRUCTURE, see that header).
//
// Semantics restated from OpenMM 8.0's Reference platform as summarised in SURVEY.md Appendix B:
//   B.1 LangevinIntegrator (leap-frog), kinetic energy shifted by dt/2;  B.2 setVelocitiesToTemperature;
//   B.3 CustomNonbondedForce (cutoff, exclusions, switching function, interaction groups);
//   B.4 bonded Custom*Force geometry (angle via acos of the clamped cosine, OpenMM's dihedral sign rule,
//       Continuous1DFunction = natural cubic spline, 0 outside its interval);
//   B.5 CustomHbondForce (donor x acceptor, cutoff on |d1 - a1|, donor/acceptor-indexed exclusions);
//   Lepton: step(x) = (x >= 0), d step = 0; max/min differentiate the selected branch; unary minus binds
//   weaker than '^' and tighter than '*'; '^' is right-associative; "a; name = expr" intermediate values.
// Derivatives are forward-mode automatic differentiation through the expression tree and through the
// distance / angle / dihedral geometry, so no hand-derived gradient formula is shared with ocm_oracle.cpp.
#include "OpenMM.h"

#include <algorithm>
#include <cctype>
#include <cstring>
#include <functional>
#include <iomanip>
#include <random>
#include <sstream>
#include <unordered_map>
#include <unordered_set>

#include "openmm_lite_internal.h"

namespace OpenMM {
namespace lite {

// ================================================================================================
// dual numbers
constexpr int MAXD = 18;
struct Dual {
    double v;
    double d[MAXD];
};

static inline void d_const(Dual& o, double v, int nd) { o.v = v; for (int k = 0; k < nd; k++) o.d[k] = 0; }
static inline void d_chain(Dual& o, const Dual& a, double f, double df, int nd) { for (int k = 0; k < nd; k++) o.d[k] = df * a.d[k]; o.v = f; }
static inline Dual d_sub(const Dual& a, const Dual& b, int nd) { Dual o; o.v = a.v - b.v; for (int k = 0; k < nd; k++) o.d[k] = a.d[k] - b.d[k]; return o; }
static inline Dual d_add(const Dual& a, const Dual& b, int nd) { Dual o; o.v = a.v + b.v; for (int k = 0; k < nd; k++) o.d[k] = a.d[k] + b.d[k]; return o; }
static inline Dual d_mul(const Dual& a, const Dual& b, int nd) { Dual o; o.v = a.v * b.v; for (int k = 0; k < nd; k++) o.d[k] = a.d[k] * b.v + a.v * b.d[k]; return o; }
static inline Dual d_div(const Dual& a, const Dual& b, int nd) {
    Dual o; const double ib = 1.0 / b.v; o.v = a.v * ib;
    for (int k = 0; k < nd; k++) o.d[k] = (a.d[k] - o.v * b.d[k]) * ib;
    return o;
}
static inline Dual d_sqrt(const Dual& a, int nd) { Dual o; const double s = std::sqrt(a.v); d_chain(o, a, s, s > 0 ? 0.5 / s : 0.0, nd); return o; }
static inline Dual d_acos(const Dual& a, int nd) { Dual o; d_chain(o, a, std::acos(a.v), -1.0 / std::sqrt(std::max(1e-300, 1.0 - a.v * a.v)), nd); return o; }
static inline Dual d_atan2(const Dual& y, const Dual& x, int nd) {
    Dual o; o.v = std::atan2(y.v, x.v);
    const double q = x.v * x.v + y.v * y.v, iq = q > 0 ? 1.0 / q : 0.0;
    for (int k = 0; k < nd; k++) o.d[k] = (x.v * y.d[k] - y.v * x.d[k]) * iq;
    return o;
}
struct DVec { Dual c[3]; };
static inline DVec dv_sub(const DVec& a, const DVec& b, int nd) { DVec o; for (int c = 0; c < 3; c++) o.c[c] = d_sub(a.c[c], b.c[c], nd); return o; }
static inline Dual dv_dot(const DVec& a, const DVec& b, int nd) { return d_add(d_add(d_mul(a.c[0], b.c[0], nd), d_mul(a.c[1], b.c[1], nd), nd), d_mul(a.c[2], b.c[2], nd), nd); }
static inline DVec dv_cross(const DVec& a, const DVec& b, int nd) {
    DVec o;
    o.c[0] = d_sub(d_mul(a.c[1], b.c[2], nd), d_mul(a.c[2], b.c[1], nd), nd);
    o.c[1] = d_sub(d_mul(a.c[2], b.c[0], nd), d_mul(a.c[0], b.c[2], nd), nd);
    o.c[2] = d_sub(d_mul(a.c[0], b.c[1], nd), d_mul(a.c[1], b.c[0], nd), nd);
    return o;
}

// ================================================================================================
// state handed to Force::evaluate
struct EvalContext {
    const std::vector<Vec3>* pos;
    double box[3];
    bool has_box;
    const System* sys = nullptr;   // (particle masses: the default weights of CustomCentroidBondForce groups)
};

static inline Vec3 min_image(Vec3 d, const EvalContext& ec, bool periodic) {
    if (periodic) for (int c = 0; c < 3; c++) d[c] -= ec.box[c] * std::floor(d[c] / ec.box[c] + 0.5);
    return d;
}
// b - a with the minimum image applied (derivatives are those of the plain difference)
static inline DVec dv_delta(const DVec& a, const DVec& b, const EvalContext& ec, bool periodic, int nd) {
    DVec o = dv_sub(b, a, nd);
    if (periodic) for (int c = 0; c < 3; c++) o.c[c].v -= ec.box[c] * std::floor(o.c[c].v / ec.box[c] + 0.5);
    return o;
}

// ================================================================================================
// expression programs
enum Op {
    CONST, PVAR, AVAR, ADD, SUB, MUL, DIV, POW, NEG, F_EXP, F_LOG, F_SIN, F_COS, F_TAN, F_ASIN, F_ACOS, F_ATAN, F_SQRT,
    F_ABS, F_FLOOR, F_CEIL, F_STEP, F_DELTA, F_MIN, F_MAX, F_SELECT, F_TAB, F_DISTANCE, F_ANGLE, F_DIHEDRAL, F_PERIODICDISTANCE
};
struct Node {
    int op = CONST;
    double c = 0;
    int a = -1, b = -1, e = -1;  // children (node indices)
    int slot = -1;               // variable slot
    const TabulatedFunction* fn = nullptr;
    int ps[6] = {-1, -1, -1, -1, -1, -1};  // particle slots (geometry) or argument nodes (periodicdistance)
};
struct Program {
    std::vector<Node> nodes;
    int root = -1;
    bool periodic = false;  // geometry functions use the minimum image
};

struct VarRef {
    enum Kind { NONE, CONSTANT, PASSIVE, ACTIVE } kind = NONE;
    double value = 0;
    int slot = -1;
};
using VarLookup = std::function<VarRef(const std::string&)>;
using ParticleLookup = std::function<int(const std::string&)>;
using FunctionLookup = std::function<const TabulatedFunction*(const std::string&)>;

class Compiler {
  public:
    Compiler(const std::string& expr, VarLookup vars, ParticleLookup parts, FunctionLookup fns)
        : vars_(std::move(vars)), parts_(std::move(parts)), fns_(std::move(fns)) {
        // split "main; name = expr; ..." (Lepton: later definitions may be used by earlier ones)
        std::vector<std::string> pieces;
        std::string cur;
        for (char ch : expr) { if (ch == ';') { pieces.push_back(cur); cur.clear(); } else cur += ch; }
        pieces.push_back(cur);
        main_ = pieces[0];
        for (std::size_t i = 1; i < pieces.size(); i++) {
            const std::string& p = pieces[i];
            const auto eq = p.find('=');
            if (eq == std::string::npos) {
                if (trim(p).empty()) continue;
                throw OpenMMException("Lepton: subexpression without '=': " + p);
            }
            defs_[trim(p.substr(0, eq))] = p.substr(eq + 1);
        }
    }
    std::shared_ptr<Program> compile() {
        prog_ = std::make_shared<Program>();
        prog_->root = parse_text(main_);
        return prog_;
    }

  private:
    VarLookup vars_;
    ParticleLookup parts_;
    FunctionLookup fns_;
    std::string main_;
    std::map<std::string, std::string> defs_;
    std::map<std::string, int> def_nodes_;
    std::set<std::string> in_progress_;
    std::shared_ptr<Program> prog_;

    struct Tok { enum T { NUM, ID, OP, LP, RP, COMMA, END } t; std::string s; double v = 0; };
    std::vector<Tok> toks_;
    std::size_t tp_ = 0;

    static std::string trim(const std::string& s) {
        std::size_t b = 0, e = s.size();
        while (b < e && std::isspace((unsigned char)s[b])) ++b;
        while (e > b && std::isspace((unsigned char)s[e - 1])) --e;
        return s.substr(b, e - b);
    }
    static std::vector<Tok> tokenize(const std::string& s) {
        std::vector<Tok> out;
        std::size_t i = 0;
        while (i < s.size()) {
            const char c = s[i];
            if (std::isspace((unsigned char)c)) { ++i; continue; }
            if (std::isdigit((unsigned char)c) || (c == '.' && i + 1 < s.size() && std::isdigit((unsigned char)s[i + 1]))) {
                std::size_t j = i;
                while (j < s.size() && (std::isdigit((unsigned char)s[j]) || s[j] == '.')) ++j;
                if (j < s.size() && (s[j] == 'e' || s[j] == 'E')) {
                    std::size_t k = j + 1;
                    if (k < s.size() && (s[k] == '+' || s[k] == '-')) ++k;
                    if (k < s.size() && std::isdigit((unsigned char)s[k])) { while (k < s.size() && std::isdigit((unsigned char)s[k])) ++k; j = k; }
                }
                Tok t{Tok::NUM, s.substr(i, j - i)};
                t.v = std::strtod(t.s.c_str(), nullptr);
                out.push_back(t);
                i = j;
                continue;
            }
            if (std::isalpha((unsigned char)c) || c == '_') {
                std::size_t j = i;
                while (j < s.size() && (std::isalnum((unsigned char)s[j]) || s[j] == '_')) ++j;
                out.push_back({Tok::ID, s.substr(i, j - i)});
                i = j;
                continue;
            }
            if (c == '(') out.push_back({Tok::LP, "("});
            else if (c == ')') out.push_back({Tok::RP, ")"});
            else if (c == ',') out.push_back({Tok::COMMA, ","});
            else if (std::strchr("+-*/^", c)) out.push_back({Tok::OP, std::string(1, c)});
            else throw OpenMMException(std::string("Lepton: unexpected character '") + c + "' in: " + s);
            ++i;
        }
        out.push_back({Tok::END, ""});
        return out;
    }

    int push(const Node& n) { prog_->nodes.push_back(n); return (int)prog_->nodes.size() - 1; }
    int constant(double v) { Node n; n.op = CONST; n.c = v; return push(n); }
    bool is_const(int i) const { return prog_->nodes[i].op == CONST; }
    double cval(int i) const { return prog_->nodes[i].c; }

    int parse_text(const std::string& text) {
        // nested parses (definitions referenced from inside an expression) keep their own token stream
        std::vector<Tok> saved = std::move(toks_);
        const std::size_t saved_tp = tp_;
        toks_ = tokenize(text);
        tp_ = 0;
        const int r = parse_prec(0);
        if (toks_[tp_].t != Tok::END) throw OpenMMException("Lepton: trailing tokens in: " + text);
        toks_ = std::move(saved);
        tp_ = saved_tp;
        return r;
    }

    int binary(int op, int a, int b) {
        if (is_const(a) && is_const(b)) {
            const double x = cval(a), y = cval(b);
            switch (op) {
                case ADD: return constant(x + y);
                case SUB: return constant(x - y);
                case MUL: return constant(x * y);
                case DIV: return constant(x / y);
                case POW: return constant(std::pow(x, y));
            }
        }
        Node n; n.op = op; n.a = a; n.b = b;
        return push(n);
    }

    int parse_prec(int prec) {
        int lhs;
        const Tok t = toks_[tp_];
        if (t.t == Tok::OP && t.s == "-") {
            ++tp_;
            const int x = parse_prec(2);
            if (is_const(x)) lhs = constant(-cval(x));
            else { Node n; n.op = NEG; n.a = x; lhs = push(n); }
        } else if (t.t == Tok::OP && t.s == "+") {
            ++tp_;
            lhs = parse_prec(2);
        } else {
            lhs = parse_primary();
        }
        for (;;) {
            const Tok& o = toks_[tp_];
            if (o.t != Tok::OP) return lhs;
            int p, op;
            bool left = true;
            switch (o.s[0]) {
                case '+': p = 0; op = ADD; break;
                case '-': p = 0; op = SUB; break;
                case '*': p = 1; op = MUL; break;
                case '/': p = 1; op = DIV; break;
                default: p = 3; op = POW; left = false; break;
            }
            if (p < prec) return lhs;
            ++tp_;
            const int rhs = parse_prec(left ? p + 1 : p);
            lhs = binary(op, lhs, rhs);
        }
    }

    int parse_primary() {
        const Tok t = toks_[tp_++];
        if (t.t == Tok::NUM) return constant(t.v);
        if (t.t == Tok::LP) {
            const int r = parse_prec(0);
            if (toks_[tp_].t != Tok::RP) throw OpenMMException("Lepton: expected ')'");
            ++tp_;
            return r;
        }
        if (t.t != Tok::ID) throw OpenMMException("Lepton: unexpected token '" + t.s + "'");
        if (toks_[tp_].t == Tok::LP) return parse_call(t.s);
        return variable(t.s);
    }

    int variable(const std::string& name) {
        auto dn = def_nodes_.find(name);
        if (dn != def_nodes_.end()) return dn->second;
        auto df = defs_.find(name);
        if (df != defs_.end()) {
            if (in_progress_.count(name)) throw OpenMMException("Lepton: recursive definition of " + name);
            in_progress_.insert(name);
            const int r = parse_text(df->second);
            in_progress_.erase(name);
            def_nodes_[name] = r;
            return r;
        }
        const VarRef v = vars_(name);
        if (v.kind == VarRef::CONSTANT) return constant(v.value);
        if (v.kind == VarRef::PASSIVE) { Node n; n.op = PVAR; n.slot = v.slot; return push(n); }
        if (v.kind == VarRef::ACTIVE) { Node n; n.op = AVAR; n.slot = v.slot; return push(n); }
        throw OpenMMException("Lepton: unknown variable \"" + name + "\"");
    }

    int parse_call(const std::string& fname) {
        ++tp_;  // '('
        static const std::map<std::string, std::pair<int, int>> geo = {{"distance", {F_DISTANCE, 2}}, {"angle", {F_ANGLE, 3}}, {"dihedral", {F_DIHEDRAL, 4}}};
        auto g = geo.find(fname);
        if (g != geo.end()) {
            Node n; n.op = g->second.first;
            for (int k = 0; k < g->second.second; k++) {
                if (toks_[tp_].t != Tok::ID) throw OpenMMException("Lepton: " + fname + "() takes particle names");
                n.ps[k] = parts_ ? parts_(toks_[tp_].s) : -1;
                if (n.ps[k] < 0) throw OpenMMException("Lepton: unknown particle \"" + toks_[tp_].s + "\" in " + fname + "()");
                ++tp_;
                if (k + 1 < g->second.second) { if (toks_[tp_].t != Tok::COMMA) throw OpenMMException("Lepton: expected ','"); ++tp_; }
            }
            if (toks_[tp_].t != Tok::RP) throw OpenMMException("Lepton: expected ')' after " + fname);
            ++tp_;
            return push(n);
        }
        std::vector<int> args;
        if (toks_[tp_].t != Tok::RP) {
            for (;;) {
                args.push_back(parse_prec(0));
                if (toks_[tp_].t == Tok::COMMA) { ++tp_; continue; }
                break;
            }
        }
        if (toks_[tp_].t != Tok::RP) throw OpenMMException("Lepton: expected ')' after arguments of " + fname);
        ++tp_;
        static const std::map<std::string, std::pair<int, int>> fn = {
            {"exp", {F_EXP, 1}}, {"log", {F_LOG, 1}}, {"sin", {F_SIN, 1}}, {"cos", {F_COS, 1}}, {"tan", {F_TAN, 1}},
            {"asin", {F_ASIN, 1}}, {"acos", {F_ACOS, 1}}, {"atan", {F_ATAN, 1}}, {"sqrt", {F_SQRT, 1}}, {"abs", {F_ABS, 1}},
            {"floor", {F_FLOOR, 1}}, {"ceil", {F_CEIL, 1}}, {"step", {F_STEP, 1}}, {"delta", {F_DELTA, 1}},
            {"min", {F_MIN, 2}}, {"max", {F_MAX, 2}}, {"select", {F_SELECT, 3}}, {"periodicdistance", {F_PERIODICDISTANCE, 6}}};
        auto f = fn.find(fname);
        Node n;
        if (f != fn.end()) {
            if ((int)args.size() != f->second.second) throw OpenMMException("Lepton: wrong number of arguments to " + fname);
            n.op = f->second.first;
            if (n.op == F_PERIODICDISTANCE) { for (int k = 0; k < 6; k++) n.ps[k] = args[k]; return push(n); }
            n.a = args[0];
            if (args.size() > 1) n.b = args[1];
            if (args.size() > 2) n.e = args[2];
            return push(n);
        }
        const TabulatedFunction* tf = fns_ ? fns_(fname) : nullptr;
        if (!tf) throw OpenMMException("Lepton: unknown function \"" + fname + "\"");
        if (args.size() != 1) throw OpenMMException("Lepton: tabulated function " + fname + " takes one argument");
        n.op = F_TAB; n.fn = tf; n.a = args[0];
        return push(n);
    }
};

// evaluate; pv = passive variable values, av = active variables, gp = Dual positions of the bond's particles
static Dual run(const Program& P, const double* pv, const Dual* av, const DVec* gp, const EvalContext& ec, int nd) {
    static thread_local std::vector<Dual> R;
    if (R.size() < P.nodes.size()) R.resize(P.nodes.size());
    const int nn = (int)P.nodes.size();
    for (int i = 0; i < nn; i++) {
        const Node& n = P.nodes[i];
        Dual& o = R[i];
        switch (n.op) {
            case CONST: d_const(o, n.c, nd); break;
            case PVAR: d_const(o, pv[n.slot], nd); break;
            case AVAR: o = av[n.slot]; break;
            case ADD: o = d_add(R[n.a], R[n.b], nd); break;
            case SUB: o = d_sub(R[n.a], R[n.b], nd); break;
            case MUL: o = d_mul(R[n.a], R[n.b], nd); break;
            case DIV: o = d_div(R[n.a], R[n.b], nd); break;
            case NEG: d_chain(o, R[n.a], -R[n.a].v, -1.0, nd); break;
            case POW: {
                const Dual &a = R[n.a], &b = R[n.b];
                const Node& bn = P.nodes[n.b];
                if (bn.op == CONST) {
                    const double p = bn.c;
                    double f, df;
                    if (p == std::floor(p) && std::fabs(p) <= 64) {
                        // integer power: repeated product, as Lepton's PowerConstant does
                        const int ip = (int)std::fabs(p);
                        double acc = 1.0, accm1 = 1.0;
                        for (int k = 0; k < ip; k++) { accm1 = acc; acc *= a.v; }
                        if (p >= 0) { f = acc; df = ip == 0 ? 0.0 : p * accm1; }
                        else { f = 1.0 / acc; df = p * f / a.v; }
                    } else {
                        f = std::pow(a.v, p);
                        df = p * std::pow(a.v, p - 1.0);
                    }
                    d_chain(o, a, f, df, nd);
                } else {
                    const double f = std::pow(a.v, b.v);
                    o.v = f;
                    const double da = b.v * std::pow(a.v, b.v - 1.0), db = a.v > 0 ? f * std::log(a.v) : 0.0;
                    for (int k = 0; k < nd; k++) o.d[k] = da * a.d[k] + db * b.d[k];
                }
                break;
            }
            case F_EXP: { const double f = std::exp(R[n.a].v); d_chain(o, R[n.a], f, f, nd); break; }
            case F_LOG: d_chain(o, R[n.a], std::log(R[n.a].v), 1.0 / R[n.a].v, nd); break;
            case F_SIN: d_chain(o, R[n.a], std::sin(R[n.a].v), std::cos(R[n.a].v), nd); break;
            case F_COS: d_chain(o, R[n.a], std::cos(R[n.a].v), -std::sin(R[n.a].v), nd); break;
            case F_TAN: { const double t = std::tan(R[n.a].v); d_chain(o, R[n.a], t, 1.0 + t * t, nd); break; }
            case F_ASIN: d_chain(o, R[n.a], std::asin(R[n.a].v), 1.0 / std::sqrt(1.0 - R[n.a].v * R[n.a].v), nd); break;
            case F_ACOS: d_chain(o, R[n.a], std::acos(R[n.a].v), -1.0 / std::sqrt(1.0 - R[n.a].v * R[n.a].v), nd); break;
            case F_ATAN: d_chain(o, R[n.a], std::atan(R[n.a].v), 1.0 / (1.0 + R[n.a].v * R[n.a].v), nd); break;
            case F_SQRT: o = d_sqrt(R[n.a], nd); break;
            case F_ABS: d_chain(o, R[n.a], std::fabs(R[n.a].v), R[n.a].v >= 0 ? 1.0 : -1.0, nd); break;
            case F_FLOOR: d_const(o, std::floor(R[n.a].v), nd); break;
            case F_CEIL: d_const(o, std::ceil(R[n.a].v), nd); break;
            case F_STEP: d_const(o, R[n.a].v >= 0 ? 1.0 : 0.0, nd); break;
            case F_DELTA: d_const(o, R[n.a].v == 0 ? 1.0 : 0.0, nd); break;
            case F_MIN: o = (R[n.a].v - R[n.b].v >= 0) ? R[n.b] : R[n.a]; break;
            case F_MAX: o = (R[n.a].v - R[n.b].v >= 0) ? R[n.a] : R[n.b]; break;
            case F_SELECT: o = (R[n.a].v != 0) ? R[n.b] : R[n.e]; break;
            case F_TAB: d_chain(o, R[n.a], n.fn->value(R[n.a].v), n.fn->derivative(R[n.a].v), nd); break;
            case F_DISTANCE: {
                const DVec d = dv_delta(gp[n.ps[0]], gp[n.ps[1]], ec, P.periodic, nd);
                o = d_sqrt(dv_dot(d, d, nd), nd);
                break;
            }
            case F_ANGLE: {
                // angle at the middle particle: acos of the clamped cosine (0 / pi at the clamps)
                const DVec u = dv_delta(gp[n.ps[1]], gp[n.ps[0]], ec, P.periodic, nd);
                const DVec w = dv_delta(gp[n.ps[1]], gp[n.ps[2]], ec, P.periodic, nd);
                const Dual c = d_div(dv_dot(u, w, nd), d_sqrt(d_mul(dv_dot(u, u, nd), dv_dot(w, w, nd), nd), nd), nd);
                if (c.v >= 1.0) d_const(o, 0.0, nd);
                else if (c.v <= -1.0) d_const(o, 3.14159265358979323846, nd);
                else o = d_acos(c, nd);
                break;
            }
            case F_DIHEDRAL: {
                // d0 = p1-p2, d1 = p3-p2, d2 = p3-p4; c0 = d0 x d1, c1 = d1 x d2; phi = angle(c0, c1), negative when
                // d0 . c1 < 0.  |c0 x c1| = |d1| |d0 . c1|, hence phi = atan2(|d1| (d0 . c1), c0 . c1).
                const DVec d0 = dv_delta(gp[n.ps[1]], gp[n.ps[0]], ec, P.periodic, nd);
                const DVec d1 = dv_delta(gp[n.ps[1]], gp[n.ps[2]], ec, P.periodic, nd);
                const DVec d2 = dv_delta(gp[n.ps[3]], gp[n.ps[2]], ec, P.periodic, nd);
                const DVec c0 = dv_cross(d0, d1, nd), c1 = dv_cross(d1, d2, nd);
                const Dual y = d_mul(d_sqrt(dv_dot(d1, d1, nd), nd), dv_dot(d0, c1, nd), nd);
                o = d_atan2(y, dv_dot(c0, c1, nd), nd);
                break;
            }
            case F_PERIODICDISTANCE: {
                Dual s; d_const(s, 0.0, nd);
                for (int c = 0; c < 3; c++) {
                    Dual d = d_sub(R[n.ps[c]], R[n.ps[c + 3]], nd);
                    if (ec.has_box) d.v -= ec.box[c] * std::floor(d.v / ec.box[c] + 0.5);
                    s = d_add(s, d_mul(d, d, nd), nd);
                }
                o = d_sqrt(s, nd);
                break;
            }
        }
    }
    return R[P.root];
}

// ================================================================================================
// JSON sink
struct Describer {
    std::ostringstream os;
    bool first = true;
    Describer() { os << std::setprecision(17); }
    void sep() { if (!first) os << ","; first = false; }
    void key(const std::string& k) { sep(); str(k); os << ":"; }
    void str(const std::string& s) {
        os << '"';
        for (char c : s) { if (c == '"' || c == '\\') os << '\\' << c; else if (c == '\n') os << "\\n"; else os << c; }
        os << '"';
    }
    void num(double v) {
        if (std::isnan(v)) os << "null";
        else if (std::isinf(v)) os << (v > 0 ? "1e999" : "-1e999");
        else os << v;
    }
    void kv(const std::string& k, double v) { key(k); num(v); }
    void kv(const std::string& k, int v) { key(k); os << v; }
    void kv(const std::string& k, const std::string& v) { key(k); str(v); }
    void kb(const std::string& k, bool v) { key(k); os << (v ? "true" : "false"); }
    void names(const std::string& k, const std::vector<std::string>& v) {
        key(k); os << "[";
        for (std::size_t i = 0; i < v.size(); i++) { if (i) os << ","; str(v[i]); }
        os << "]";
    }
    void nums(const std::vector<double>& v) { os << "["; for (std::size_t i = 0; i < v.size(); i++) { if (i) os << ","; num(v[i]); } os << "]"; }
    void ints(const std::vector<int>& v) { os << "["; for (std::size_t i = 0; i < v.size(); i++) { if (i) os << ","; os << v[i]; } os << "]"; }
    void open(const std::string& k, char br) { key(k); os << br; first = true; }
    void close(char br) { os << br; first = false; }
};

}  // namespace lite

using lite::Describer;
using lite::Dual;
using lite::DVec;
using lite::EvalContext;
using lite::VarRef;

// ================================================================================================
// Continuous1DFunction
Continuous1DFunction::Continuous1DFunction(const std::vector<double>& values, double min, double max, bool)
    : y_(values), min_(min), max_(max) {
    const int n = (int)y_.size();
    if (n < 2) throw OpenMMException("Continuous1DFunction: a tabulated function must have at least two points");
    if (!(max > min)) throw OpenMMException("Continuous1DFunction: max <= min");
    // natural cubic spline: second derivatives from the tridiagonal system, h = (max-min)/(n-1)
    y2_.assign(n, 0.0);
    if (n > 2) {
        const double h = (max - min) / (n - 1);
        std::vector<double> a(n, 0), b(n, 0), c(n, 0), r(n, 0);
        b[0] = b[n - 1] = 1.0;
        for (int i = 1; i < n - 1; i++) {
            a[i] = h / 6.0; b[i] = 2.0 * h / 3.0; c[i] = h / 6.0;
            r[i] = (y_[i + 1] - 2.0 * y_[i] + y_[i - 1]) / h;
        }
        for (int i = 1; i < n; i++) { const double m = a[i] / b[i - 1]; b[i] -= m * c[i - 1]; r[i] -= m * r[i - 1]; }
        y2_[n - 1] = r[n - 1] / b[n - 1];
        for (int i = n - 2; i >= 0; i--) y2_[i] = (r[i] - c[i] * y2_[i + 1]) / b[i];
    }
}
void Continuous1DFunction::getFunctionParameters(std::vector<double>& values, double& min, double& max) const { values = y_; min = min_; max = max_; }
double Continuous1DFunction::value(double x) const {
    if (!(x >= min_ && x <= max_)) return 0.0;
    const int n = (int)y_.size();
    const double h = (max_ - min_) / (n - 1);
    int i = std::min(n - 2, std::max(0, (int)std::floor((x - min_) / h)));
    const double A = (min_ + (i + 1) * h - x) / h, B = 1.0 - A;
    return A * y_[i] + B * y_[i + 1] + ((A * A * A - A) * y2_[i] + (B * B * B - B) * y2_[i + 1]) * h * h / 6.0;
}
double Continuous1DFunction::derivative(double x) const {
    if (!(x >= min_ && x <= max_)) return 0.0;
    const int n = (int)y_.size();
    const double h = (max_ - min_) / (n - 1);
    int i = std::min(n - 2, std::max(0, (int)std::floor((x - min_) / h)));
    const double A = (min_ + (i + 1) * h - x) / h, B = 1.0 - A;
    return (y_[i + 1] - y_[i]) / h + (-(3.0 * A * A - 1.0) * y2_[i] + (3.0 * B * B - 1.0) * y2_[i + 1]) * h / 6.0;
}

// ================================================================================================
// Force plumbing
void Force::setForceGroup(int g) {
    if (g < 0 || g > 31) throw OpenMMException("Force group must be between 0 and 31");
    group_ = g;
}
CustomForceBase::~CustomForceBase() { for (auto& f : functions_) delete f.second; }
int CustomForceBase::addGlobalParameter(const std::string& name, double value) { globals_.emplace_back(name, value); prog_.reset(); return (int)globals_.size() - 1; }
int CustomForceBase::addTabulatedFunction(const std::string& name, TabulatedFunction* fn) { functions_.emplace_back(name, fn); prog_.reset(); return (int)functions_.size() - 1; }
void CustomForceBase::describe_common(Describer& d, const char* cls) const {
    d.kv("class", std::string(cls));
    d.kv("group", getForceGroup());
    d.kv("expression", expr_);
    d.open("globals", '{');
    for (auto& g : globals_) d.kv(g.first, g.second);
    d.close('}');
    d.open("tabulated", '{');
    for (auto& f : functions_) {
        d.key(f.first);
        if (auto* c = dynamic_cast<const Continuous1DFunction*>(f.second)) {
            std::vector<double> v; double lo, hi;
            c->getFunctionParameters(v, lo, hi);
            d.os << "{\"min\":"; d.num(lo); d.os << ",\"max\":"; d.num(hi); d.os << ",\"values\":"; d.nums(v); d.os << "}";
        } else d.os << "null";
    }
    d.close('}');
}

namespace {

VarRef find_in(const std::vector<std::pair<std::string, double>>& globals, const std::string& name) {
    for (auto& g : globals) if (g.first == name) { VarRef r; r.kind = VarRef::CONSTANT; r.value = g.second; return r; }
    return {};
}
int index_of(const std::vector<std::string>& v, const std::string& name) {
    for (std::size_t i = 0; i < v.size(); i++) if (v[i] == name) return (int)i;
    return -1;
}
VarRef passive(int slot) { VarRef r; r.kind = VarRef::PASSIVE; r.slot = slot; return r; }
VarRef active(int slot) { VarRef r; r.kind = VarRef::ACTIVE; r.slot = slot; return r; }

lite::FunctionLookup fn_lookup(const std::vector<std::pair<std::string, TabulatedFunction*>>& fns) {
    return [&fns](const std::string& n) -> const TabulatedFunction* {
        for (auto& f : fns) if (f.first == n) return f.second;
        return nullptr;
    };
}

DVec seed_position(const Vec3& p, int slot, int nd) {
    DVec o;
    for (int c = 0; c < 3; c++) { lite::d_const(o.c[c], p[c], nd); o.c[c].d[3 * slot + c] = 1.0; }
    return o;
}

}  // namespace

// ---- HarmonicBondForce / HarmonicAngleForce --------------------------------------------------------------
int HarmonicBondForce::addBond(int p1, int p2, double length, double k) { bonds_.push_back({p1, p2, length, k}); return (int)bonds_.size() - 1; }
double HarmonicBondForce::evaluate(const EvalContext& ec, std::vector<Vec3>* F) const {
    double E = 0;
    const auto& x = *ec.pos;
    for (auto& b : bonds_) {
        const Vec3 d = lite::min_image(x[b.p2] - x[b.p1], ec, periodic_);
        const double r = std::sqrt(d.dot(d)), dr = r - b.length;
        E += 0.5 * b.k * dr * dr;
        if (F && r > 0) { const Vec3 f = d * (b.k * dr / r); (*F)[b.p1] += f; (*F)[b.p2] -= f; }
    }
    return E;
}
void HarmonicBondForce::describe(Describer& d) const {
    d.kv("class", std::string("HarmonicBondForce")); d.kv("group", getForceGroup()); d.kb("periodic", periodic_);
    d.open("bonds", '[');
    for (auto& b : bonds_) { d.sep(); d.os << "[" << b.p1 << "," << b.p2 << ","; d.num(b.length); d.os << ","; d.num(b.k); d.os << "]"; }
    d.close(']');
}
int HarmonicAngleForce::addAngle(int p1, int p2, int p3, double angle, double k) { angles_.push_back({p1, p2, p3, angle, k}); return (int)angles_.size() - 1; }
double HarmonicAngleForce::evaluate(const EvalContext& ec, std::vector<Vec3>* F) const {
    double E = 0;
    const auto& x = *ec.pos;
    const int nd = 9;
    for (auto& a : angles_) {
        DVec gp[3] = {seed_position(x[a.p1], 0, nd), seed_position(x[a.p2], 1, nd), seed_position(x[a.p3], 2, nd)};
        const DVec u = lite::dv_delta(gp[1], gp[0], ec, periodic_, nd), w = lite::dv_delta(gp[1], gp[2], ec, periodic_, nd);
        const Dual c = lite::d_div(lite::dv_dot(u, w, nd), lite::d_sqrt(lite::d_mul(lite::dv_dot(u, u, nd), lite::dv_dot(w, w, nd), nd), nd), nd);
        Dual th;
        if (c.v >= 1.0) lite::d_const(th, 0.0, nd);
        else if (c.v <= -1.0) lite::d_const(th, 3.14159265358979323846, nd);
        else th = lite::d_acos(c, nd);
        const double dt = th.v - a.angle;
        E += 0.5 * a.k * dt * dt;
        if (F) {
            const int p[3] = {a.p1, a.p2, a.p3};
            for (int s = 0; s < 3; s++) for (int cc = 0; cc < 3; cc++) (*F)[p[s]][cc] -= a.k * dt * th.d[3 * s + cc];
        }
    }
    return E;
}
void HarmonicAngleForce::describe(Describer& d) const {
    d.kv("class", std::string("HarmonicAngleForce")); d.kv("group", getForceGroup()); d.kb("periodic", periodic_);
    d.open("angles", '[');
    for (auto& a : angles_) { d.sep(); d.os << "[" << a.p1 << "," << a.p2 << "," << a.p3 << ","; d.num(a.angle); d.os << ","; d.num(a.k); d.os << "]"; }
    d.close(']');
}

// ---- CustomBondForce -----------------------------------------------------------------------------------
int CustomBondForce::addPerBondParameter(const std::string& n) { params_.push_back(n); prog_.reset(); return (int)params_.size() - 1; }
int CustomBondForce::addBond(int p1, int p2, const std::vector<double>& params) {
    if (params.size() != params_.size()) throw OpenMMException("CustomBondForce: wrong number of parameters");
    bonds_.push_back({p1, p2, params});
    return (int)bonds_.size() - 1;
}
double CustomBondForce::evaluate(const EvalContext& ec, std::vector<Vec3>* F) const {
    if (!prog_) {
        lite::Compiler c(expr_, [this](const std::string& n) -> VarRef {
            if (n == "r") return active(0);
            const int k = index_of(params_, n);
            if (k >= 0) return passive(k);
            return find_in(globals_, n);
        }, nullptr, fn_lookup(functions_));
        prog_ = c.compile();
    }
    double E = 0;
    const auto& x = *ec.pos;
    for (auto& b : bonds_) {
        const Vec3 d = lite::min_image(x[b.p2] - x[b.p1], ec, periodic_);
        Dual r; r.v = std::sqrt(d.dot(d)); r.d[0] = 1.0;
        const Dual e = lite::run(*prog_, b.params.data(), &r, nullptr, ec, 1);
        E += e.v;
        if (F && r.v > 0) { const Vec3 f = d * (e.d[0] / r.v); (*F)[b.p1] += f; (*F)[b.p2] -= f; }
    }
    return E;
}
void CustomBondForce::describe(Describer& d) const {
    describe_common(d, "CustomBondForce");
    d.kb("periodic", periodic_);
    d.names("per_bond_parameters", params_);
    d.open("bonds", '[');
    for (auto& b : bonds_) { d.sep(); d.os << "[" << b.p1 << "," << b.p2 << ","; d.nums(b.params); d.os << "]"; }
    d.close(']');
}

// ---- CustomTorsionForce --------------------------------------------------------------------------------
int CustomTorsionForce::addPerTorsionParameter(const std::string& n) { params_.push_back(n); prog_.reset(); return (int)params_.size() - 1; }
int CustomTorsionForce::addTorsion(int p1, int p2, int p3, int p4, const std::vector<double>& params) {
    if (params.size() != params_.size()) throw OpenMMException("CustomTorsionForce: wrong number of parameters");
    torsions_.push_back({{p1, p2, p3, p4}, params});
    return (int)torsions_.size() - 1;
}
double CustomTorsionForce::evaluate(const EvalContext& ec, std::vector<Vec3>* F) const {
    if (!prog_) {
        // "theta" is the dihedral of the four particles: compile it as dihedral(p1,p2,p3,p4)
        lite::Compiler c(expr_ + "; theta = dihedral(p1, p2, p3, p4)", [this](const std::string& n) -> VarRef {
            const int k = index_of(params_, n);
            if (k >= 0) return passive(k);
            return find_in(globals_, n);
        }, [](const std::string& n) { return (n.size() == 2 && n[0] == 'p' && n[1] >= '1' && n[1] <= '4') ? n[1] - '1' : -1; }, fn_lookup(functions_));
        prog_ = c.compile();
        prog_->periodic = periodic_;
    }
    double E = 0;
    const auto& x = *ec.pos;
    const int nd = 12;
    for (auto& t : torsions_) {
        DVec gp[4];
        for (int s = 0; s < 4; s++) gp[s] = seed_position(x[t.p[s]], s, nd);
        const Dual e = lite::run(*prog_, t.params.data(), nullptr, gp, ec, nd);
        E += e.v;
        if (F) for (int s = 0; s < 4; s++) for (int c = 0; c < 3; c++) (*F)[t.p[s]][c] -= e.d[3 * s + c];
    }
    return E;
}
void CustomTorsionForce::describe(Describer& d) const {
    describe_common(d, "CustomTorsionForce");
    d.kb("periodic", periodic_);
    d.names("per_torsion_parameters", params_);
    d.open("torsions", '[');
    for (auto& t : torsions_) { d.sep(); d.os << "[" << t.p[0] << "," << t.p[1] << "," << t.p[2] << "," << t.p[3] << ","; d.nums(t.params); d.os << "]"; }
    d.close(']');
}

// ---- CustomCompoundBondForce ---------------------------------------------------------------------------
int CustomCompoundBondForce::addPerBondParameter(const std::string& n) { params_.push_back(n); prog_.reset(); return (int)params_.size() - 1; }
int CustomCompoundBondForce::addBond(const std::vector<int>& particles, const std::vector<double>& params) {
    if ((int)particles.size() != np_) throw OpenMMException("CustomCompoundBondForce: wrong number of particles");
    if (params.size() != params_.size()) throw OpenMMException("CustomCompoundBondForce: wrong number of parameters");
    bonds_.push_back({particles, params});
    return (int)bonds_.size() - 1;
}
double CustomCompoundBondForce::evaluate(const EvalContext& ec, std::vector<Vec3>* F) const {
    if (np_ > 6) throw OpenMMException("openmm-lite: CustomCompoundBondForce with more than 6 particles");
    if (!prog_) {
        const int np = np_;
        lite::Compiler c(expr_, [this](const std::string& n) -> VarRef {
            const int k = index_of(params_, n);
            if (k >= 0) return passive(k);
            return find_in(globals_, n);
        }, [np](const std::string& n) { return (n.size() == 2 && n[0] == 'p' && n[1] >= '1' && n[1] < '1' + np) ? n[1] - '1' : -1; }, fn_lookup(functions_));
        prog_ = c.compile();
        prog_->periodic = periodic_;
    }
    double E = 0;
    const auto& x = *ec.pos;
    const int nd = 3 * np_;
    for (auto& b : bonds_) {
        DVec gp[6];
        for (int s = 0; s < np_; s++) gp[s] = seed_position(x[b.p[s]], s, nd);
        const Dual e = lite::run(*prog_, b.params.data(), nullptr, gp, ec, nd);
        E += e.v;
        if (F) for (int s = 0; s < np_; s++) for (int c = 0; c < 3; c++) (*F)[b.p[s]][c] -= e.d[3 * s + c];
    }
    return E;
}
void CustomCompoundBondForce::describe(Describer& d) const {
    describe_common(d, "CustomCompoundBondForce");
    d.kb("periodic", periodic_);
    d.kv("particles_per_bond", np_);
    d.names("per_bond_parameters", params_);
    d.open("bonds", '[');
    for (auto& b : bonds_) { d.sep(); d.os << "["; d.ints(b.p); d.os << ","; d.nums(b.params); d.os << "]"; }
    d.close(']');
}

// ---- CustomCentroidBondForce (recorded; only evaluated for unit weights = plain centroid by mass is not needed here) ----
int CustomCentroidBondForce::addPerBondParameter(const std::string& n) { params_.push_back(n); return (int)params_.size() - 1; }
int CustomCentroidBondForce::addGroup(const std::vector<int>& particles, const std::vector<double>& weights) { groups_.push_back({particles, weights}); return (int)groups_.size() - 1; }
int CustomCentroidBondForce::addBond(const std::vector<int>& groups, const std::vector<double>& params) { bonds_.push_back({groups, params}); return (int)bonds_.size() - 1; }
// OpenMM: the position of a group is the weighted mean of its particles, weights = the given ones or, when none are given, the
// particle masses; the expression sees the groups as g1, g2, ... and the force on a group centre is shared out by weight.
double CustomCentroidBondForce::evaluate(const EvalContext& ec, std::vector<Vec3>* F) const {
    if (bonds_.empty()) return 0.0;
    if (ng_ > 6) throw OpenMMException("openmm-lite: CustomCentroidBondForce with more than 6 groups per bond");
    if (!prog_) {
        const int ng = ng_;
        lite::Compiler c(expr_, [this](const std::string& n) -> VarRef {
            const int k = index_of(params_, n);
            if (k >= 0) return passive(k);
            return find_in(globals_, n);
        }, [ng](const std::string& n) { return (n.size() == 2 && n[0] == 'g' && n[1] >= '1' && n[1] < '1' + ng) ? n[1] - '1' : -1; }, fn_lookup(functions_));
        prog_ = c.compile();
        prog_->periodic = periodic_;
    }
    const auto& x = *ec.pos;
    auto weights = [&](const Group& g) {
        std::vector<double> w = g.w;
        if (w.empty()) for (int i : g.p) w.push_back(ec.sys->getParticleMass(i));
        if (w.size() != g.p.size()) throw OpenMMException("CustomCentroidBondForce: wrong number of weights");
        double tot = 0;
        for (double v : w) tot += v;
        for (double& v : w) v /= tot;
        return w;
    };
    double E = 0;
    const int nd = 3 * ng_;
    for (auto& b : bonds_) {
        DVec gp[6];
        std::vector<std::vector<double>> w;
        for (int s = 0; s < ng_; s++) {
            const Group& g = groups_.at(b.g[s]);
            w.push_back(weights(g));
            Vec3 cen;
            for (std::size_t k = 0; k < g.p.size(); k++) cen += x[g.p[k]] * w[s][k];
            gp[s] = seed_position(cen, s, nd);
        }
        const Dual e = lite::run(*prog_, b.params.data(), nullptr, gp, ec, nd);
        E += e.v;
        if (F)
            for (int s = 0; s < ng_; s++) {
                const Group& g = groups_.at(b.g[s]);
                for (std::size_t k = 0; k < g.p.size(); k++)
                    for (int c = 0; c < 3; c++) (*F)[g.p[k]][c] -= e.d[3 * s + c] * w[s][k];
            }
    }
    return E;
}
void CustomCentroidBondForce::describe(Describer& d) const {
    describe_common(d, "CustomCentroidBondForce");
    d.kv("groups_per_bond", ng_);
    d.names("per_bond_parameters", params_);
    d.open("groups", '[');
    for (auto& g : groups_) { d.sep(); d.os << "["; d.ints(g.p); d.os << ","; d.nums(g.w); d.os << "]"; }
    d.close(']');
    d.open("bonds", '[');
    for (auto& b : bonds_) { d.sep(); d.os << "["; d.ints(b.g); d.os << ","; d.nums(b.params); d.os << "]"; }
    d.close(']');
}

// ---- CustomExternalForce -------------------------------------------------------------------------------
int CustomExternalForce::addPerParticleParameter(const std::string& n) { params_.push_back(n); prog_.reset(); return (int)params_.size() - 1; }
int CustomExternalForce::addParticle(int particle, const std::vector<double>& params) {
    if (params.size() != params_.size()) throw OpenMMException("CustomExternalForce: wrong number of parameters");
    items_.push_back({particle, params});
    return (int)items_.size() - 1;
}
double CustomExternalForce::evaluate(const EvalContext& ec, std::vector<Vec3>* F) const {
    if (!prog_) {
        lite::Compiler c(expr_, [this](const std::string& n) -> VarRef {
            if (n == "x") return active(0);
            if (n == "y") return active(1);
            if (n == "z") return active(2);
            const int k = index_of(params_, n);
            if (k >= 0) return passive(k);
            return find_in(globals_, n);
        }, nullptr, fn_lookup(functions_));
        prog_ = c.compile();
    }
    double E = 0;
    const auto& x = *ec.pos;
    for (auto& it : items_) {
        Dual av[3];
        for (int c = 0; c < 3; c++) { lite::d_const(av[c], x[it.p][c], 3); av[c].d[c] = 1.0; }
        const Dual e = lite::run(*prog_, it.params.data(), av, nullptr, ec, 3);
        E += e.v;
        if (F) for (int c = 0; c < 3; c++) (*F)[it.p][c] -= e.d[c];
    }
    return E;
}
void CustomExternalForce::describe(Describer& d) const {
    describe_common(d, "CustomExternalForce");
    d.names("per_particle_parameters", params_);
    d.open("particles", '[');
    for (auto& it : items_) { d.sep(); d.os << "[" << it.p << ","; d.nums(it.params); d.os << "]"; }
    d.close(']');
}

// ---- neighbour search shared by the cutoff forces -------------------------------------------------------
namespace {

// calls f(i, j) once for every unordered pair of `members` closer than `cutoff` (all pairs when cutoff <= 0)
template <class Fn>
void for_each_close_pair(const std::vector<Vec3>& x, const std::vector<int>& members, double cutoff, bool periodic,
                         const EvalContext& ec, Fn&& f) {
    const int m = (int)members.size();
    const double rc2 = cutoff * cutoff;
    if (cutoff <= 0 || m < 400) {
        for (int a = 0; a < m; a++)
            for (int b = a + 1; b < m; b++) {
                if (cutoff > 0) {
                    const Vec3 d = lite::min_image(x[members[b]] - x[members[a]], ec, periodic);
                    if (!(d.dot(d) < rc2)) continue;
                }
                f(members[a], members[b]);
            }
        return;
    }
    double lo[3] = {0, 0, 0}, len[3];
    int nc[3];
    if (periodic) {
        for (int c = 0; c < 3; c++) len[c] = ec.box[c];
    } else {
        double hi[3];
        for (int c = 0; c < 3; c++) { lo[c] = 1e300; hi[c] = -1e300; }
        for (int i : members) for (int c = 0; c < 3; c++) { lo[c] = std::min(lo[c], x[i][c]); hi[c] = std::max(hi[c], x[i][c]); }
        for (int c = 0; c < 3; c++) len[c] = std::max(hi[c] - lo[c], 1e-9) * (1.0 + 1e-9) + 1e-9;
    }
    for (int c = 0; c < 3; c++) nc[c] = std::max(1, std::min(64, (int)std::floor(len[c] / cutoff)));
    const int ncell = nc[0] * nc[1] * nc[2];
    std::vector<std::vector<int>> cells(ncell);
    auto cell_of = [&](const Vec3& p, int c3[3]) {
        for (int c = 0; c < 3; c++) {
            double fr = (p[c] - lo[c]) / len[c];
            if (periodic) fr -= std::floor(fr);
            c3[c] = std::min(nc[c] - 1, std::max(0, (int)(fr * nc[c])));
        }
    };
    for (int i : members) { int c3[3]; cell_of(x[i], c3); cells[(c3[2] * nc[1] + c3[1]) * nc[0] + c3[0]].push_back(i); }
    for (int cz = 0; cz < nc[2]; cz++) for (int cy = 0; cy < nc[1]; cy++) for (int cx = 0; cx < nc[0]; cx++) {
        const int c0 = (cz * nc[1] + cy) * nc[0] + cx;
        std::vector<int> nb;
        for (int dz = -1; dz <= 1; dz++) for (int dy = -1; dy <= 1; dy++) for (int dx = -1; dx <= 1; dx++) {
            int q[3] = {cx + dx, cy + dy, cz + dz};
            bool ok = true;
            for (int c = 0; c < 3; c++) {
                if (periodic) q[c] = (q[c] + nc[c]) % nc[c];
                else if (q[c] < 0 || q[c] >= nc[c]) ok = false;
            }
            if (!ok) continue;
            const int c1 = (q[2] * nc[1] + q[1]) * nc[0] + q[0];
            if (c1 >= c0) nb.push_back(c1);
        }
        std::sort(nb.begin(), nb.end());
        nb.erase(std::unique(nb.begin(), nb.end()), nb.end());
        for (int c1 : nb) {
            const auto &A = cells[c0], &B = cells[c1];
            for (std::size_t a = 0; a < A.size(); a++)
                for (std::size_t b = (c1 == c0 ? a + 1 : 0); b < B.size(); b++) {
                    const Vec3 d = lite::min_image(x[B[b]] - x[A[a]], ec, periodic);
                    if (d.dot(d) < rc2) f(std::min(A[a], B[b]), std::max(A[a], B[b]));
                }
        }
    }
}

inline std::uint64_t pair_key(int a, int b) { return ((std::uint64_t)(unsigned)std::min(a, b) << 32) | (unsigned)std::max(a, b); }

}  // namespace

// ---- CustomNonbondedForce ------------------------------------------------------------------------------
int CustomNonbondedForce::addPerParticleParameter(const std::string& n) { params_.push_back(n); prog_.reset(); return (int)params_.size() - 1; }
int CustomNonbondedForce::addParticle(const std::vector<double>& params) {
    if (params.size() != params_.size()) throw OpenMMException("CustomNonbondedForce: wrong number of parameters");
    particles_.push_back(params);
    return (int)particles_.size() - 1;
}
int CustomNonbondedForce::addExclusion(int p1, int p2) { exclusions_.emplace_back(p1, p2); return (int)exclusions_.size() - 1; }
int CustomNonbondedForce::addInteractionGroup(const std::set<int>& s1, const std::set<int>& s2) { groups_.emplace_back(s1, s2); return (int)groups_.size() - 1; }
double CustomNonbondedForce::evaluate(const EvalContext& ec, std::vector<Vec3>* F) const {
    const int n = (int)particles_.size();
    const auto& x = *ec.pos;
    if (n != (int)x.size()) throw OpenMMException("CustomNonbondedForce must have exactly as many particles as the System it belongs to.");
    const int np = (int)params_.size();
    if (!prog_) {
        lite::Compiler c(expr_, [this, np](const std::string& nm) -> VarRef {
            if (nm == "r") return active(0);
            if (!nm.empty() && (nm.back() == '1' || nm.back() == '2')) {
                const int k = index_of(params_, nm.substr(0, nm.size() - 1));
                if (k >= 0) return passive(nm.back() == '1' ? k : np + k);
            }
            return find_in(globals_, nm);
        }, nullptr, fn_lookup(functions_));
        prog_ = c.compile();
    }
    const bool periodic = method_ == CutoffPeriodic;
    if (periodic && !ec.has_box) throw OpenMMException("CustomNonbondedForce: CutoffPeriodic without a periodic box");
    if (periodic) for (int c = 0; c < 3; c++) if (cutoff_ > 0.5 * ec.box[c]) throw OpenMMException("CustomNonbondedForce: The cutoff distance cannot be greater than half the periodic box size.");
    const double cutoff = method_ == NoCutoff ? 0.0 : cutoff_;
    if (use_switch_ && (switch_ < 0 || switch_ >= cutoff_)) throw OpenMMException("CustomNonbondedForce: Switching distance must satisfy 0 <= r_switch < r_cutoff");
    std::unordered_set<std::uint64_t> excl;
    excl.reserve(exclusions_.size() * 2 + 1);
    for (auto& e : exclusions_) excl.insert(pair_key(e.first, e.second));

    double E = 0;
    std::vector<double> pv(2 * np);
    auto do_pair = [&](int i, int j) {
        if (excl.count(pair_key(i, j))) return;
        const Vec3 d = lite::min_image(x[j] - x[i], ec, periodic);
        Dual r; r.v = std::sqrt(d.dot(d)); r.d[0] = 1.0;
        for (int k = 0; k < np; k++) { pv[k] = particles_[i][k]; pv[np + k] = particles_[j][k]; }
        Dual e = lite::run(*prog_, pv.data(), &r, nullptr, ec, 1);
        if (use_switch_ && cutoff > 0 && r.v > switch_) {
            const double w = cutoff_ - switch_, t = (r.v - switch_) / w;
            const double S = 1.0 + t * t * t * (-10.0 + t * (15.0 - t * 6.0));
            const double dS = t * t * (-30.0 + t * (60.0 - t * 30.0)) / w;
            e.d[0] = S * e.d[0] + e.v * dS;
            e.v *= S;
        }
        E += e.v;
        if (F && r.v > 0) { const Vec3 f = d * (e.d[0] / r.v); (*F)[i] += f; (*F)[j] -= f; }
    };
    if (groups_.empty()) {
        std::vector<int> all(n);
        for (int i = 0; i < n; i++) all[i] = i;
        for_each_close_pair(x, all, cutoff, periodic, ec, do_pair);
    } else {
        // a pair is evaluated once for every group in which one bead is in set1 and the other in set2
        for (auto& g : groups_) {
            std::vector<char> in1(n, 0), in2(n, 0);
            std::vector<int> members;
            for (int i : g.first) { if (i < 0 || i >= n) throw OpenMMException("CustomNonbondedForce: interaction group index out of range"); in1[i] = 1; }
            for (int i : g.second) { if (i < 0 || i >= n) throw OpenMMException("CustomNonbondedForce: interaction group index out of range"); in2[i] = 1; }
            for (int i = 0; i < n; i++) if (in1[i] || in2[i]) members.push_back(i);
            for_each_close_pair(x, members, cutoff, periodic, ec, [&](int i, int j) {
                if ((in1[i] && in2[j]) || (in1[j] && in2[i])) do_pair(i, j);
            });
        }
    }
    return E;
}
void CustomNonbondedForce::describe(Describer& d) const {
    describe_common(d, "CustomNonbondedForce");
    d.kv("method", (int)method_);
    d.kv("cutoff", cutoff_);
    d.kb("use_switch", use_switch_);
    d.kv("switch_distance", switch_);
    d.names("per_particle_parameters", params_);
    d.open("particles", '[');
    for (auto& p : particles_) { d.sep(); d.nums(p); }
    d.close(']');
    d.open("exclusions", '[');
    for (auto& e : exclusions_) { d.sep(); d.os << "[" << e.first << "," << e.second << "]"; }
    d.close(']');
    d.open("interaction_groups", '[');
    for (auto& g : groups_) {
        d.sep();
        d.os << "[";
        d.ints(std::vector<int>(g.first.begin(), g.first.end()));
        d.os << ",";
        d.ints(std::vector<int>(g.second.begin(), g.second.end()));
        d.os << "]";
    }
    d.close(']');
}

// ---- CustomHbondForce ----------------------------------------------------------------------------------
int CustomHbondForce::addPerDonorParameter(const std::string& n) { dparams_.push_back(n); prog_.reset(); return (int)dparams_.size() - 1; }
int CustomHbondForce::addPerAcceptorParameter(const std::string& n) { aparams_.push_back(n); prog_.reset(); return (int)aparams_.size() - 1; }
int CustomHbondForce::addDonor(int d1, int d2, int d3, const std::vector<double>& params) {
    if (params.size() != dparams_.size()) throw OpenMMException("CustomHbondForce: wrong number of donor parameters");
    donors_.push_back({{d1, d2, d3}, params});
    return (int)donors_.size() - 1;
}
int CustomHbondForce::addAcceptor(int a1, int a2, int a3, const std::vector<double>& params) {
    if (params.size() != aparams_.size()) throw OpenMMException("CustomHbondForce: wrong number of acceptor parameters");
    acceptors_.push_back({{a1, a2, a3}, params});
    return (int)acceptors_.size() - 1;
}
int CustomHbondForce::addExclusion(int donor, int acceptor) {
    if (donor < 0 || donor >= (int)donors_.size() || acceptor < 0 || acceptor >= (int)acceptors_.size())
        throw OpenMMException("CustomHbondForce: exclusion index out of range");
    exclusions_.emplace_back(donor, acceptor);
    return (int)exclusions_.size() - 1;
}
double CustomHbondForce::evaluate(const EvalContext& ec, std::vector<Vec3>* F) const {
    const int nd_ = (int)dparams_.size(), na_ = (int)aparams_.size();
    if (!prog_) {
        lite::Compiler c(expr_, [this, nd_](const std::string& n) -> VarRef {
            int k = index_of(dparams_, n);
            if (k >= 0) return passive(k);
            k = index_of(aparams_, n);
            if (k >= 0) return passive(nd_ + k);
            return find_in(globals_, n);
        }, [](const std::string& n) {
            if (n.size() == 2 && n[1] >= '1' && n[1] <= '3') { if (n[0] == 'd') return n[1] - '1'; if (n[0] == 'a') return 3 + (n[1] - '1'); }
            return -1;
        }, fn_lookup(functions_));
        prog_ = c.compile();
        prog_->periodic = method_ == CutoffPeriodic;
    }
    const bool periodic = method_ == CutoffPeriodic;
    const auto& x = *ec.pos;
    std::unordered_set<std::uint64_t> excl;
    for (auto& e : exclusions_) excl.insert(((std::uint64_t)(unsigned)e.first << 32) | (unsigned)e.second);
    const double rc2 = cutoff_ * cutoff_;
    const int nd = 18;
    double E = 0;
    std::vector<double> pv(nd_ + na_);
    for (int di = 0; di < (int)donors_.size(); di++) {
        const auto& D = donors_[di];
        for (int ai = 0; ai < (int)acceptors_.size(); ai++) {
            const auto& A = acceptors_[ai];
            if (excl.count(((std::uint64_t)(unsigned)di << 32) | (unsigned)ai)) continue;
            if (method_ != NoCutoff) {
                const Vec3 d = lite::min_image(x[A.p[0]] - x[D.p[0]], ec, periodic);
                if (!(d.dot(d) < rc2)) continue;
            }
            DVec gp[6];
            for (int s = 0; s < 3; s++) {
                if (D.p[s] >= 0) gp[s] = seed_position(x[D.p[s]], s, nd);
                if (A.p[s] >= 0) gp[3 + s] = seed_position(x[A.p[s]], 3 + s, nd);
            }
            for (int k = 0; k < nd_; k++) pv[k] = D.params[k];
            for (int k = 0; k < na_; k++) pv[nd_ + k] = A.params[k];
            const Dual e = lite::run(*prog_, pv.data(), nullptr, gp, ec, nd);
            E += e.v;
            if (F) for (int s = 0; s < 3; s++) for (int c = 0; c < 3; c++) {
                if (D.p[s] >= 0) (*F)[D.p[s]][c] -= e.d[3 * s + c];
                if (A.p[s] >= 0) (*F)[A.p[s]][c] -= e.d[3 * (3 + s) + c];
            }
        }
    }
    return E;
}
void CustomHbondForce::describe(Describer& d) const {
    describe_common(d, "CustomHbondForce");
    d.kv("method", (int)method_);
    d.kv("cutoff", cutoff_);
    d.names("per_donor_parameters", dparams_);
    d.names("per_acceptor_parameters", aparams_);
    d.open("donors", '[');
    for (auto& s : donors_) { d.sep(); d.os << "[" << s.p[0] << "," << s.p[1] << "," << s.p[2] << ","; d.nums(s.params); d.os << "]"; }
    d.close(']');
    d.open("acceptors", '[');
    for (auto& s : acceptors_) { d.sep(); d.os << "[" << s.p[0] << "," << s.p[1] << "," << s.p[2] << ","; d.nums(s.params); d.os << "]"; }
    d.close(']');
    d.open("exclusions", '[');
    for (auto& e : exclusions_) { d.sep(); d.os << "[" << e.first << "," << e.second << "]"; }
    d.close(']');
}

// ---- barostats -----------------------------------------------------------------------------------------
void MonteCarloAnisotropicBarostat::describe(Describer& d) const {
    d.kv("class", std::string("MonteCarloAnisotropicBarostat")); d.kv("group", getForceGroup());
    d.key("pressure"); d.nums({p_[0], p_[1], p_[2]});
    d.kv("temperature", t_); d.kb("scale_x", sx_); d.kb("scale_y", sy_); d.kb("scale_z", sz_); d.kv("frequency", freq_);
}
void MonteCarloMembraneBarostat::describe(Describer& d) const {
    d.kv("class", std::string("MonteCarloMembraneBarostat")); d.kv("group", getForceGroup());
    d.kv("pressure", p_); d.kv("surface_tension", g_); d.kv("temperature", t_); d.kv("xymode", (int)xy_); d.kv("zmode", (int)z_);
    d.kv("frequency", freq_);
}

// ================================================================================================
// System / Platform
System::~System() { for (Force* f : forces_) delete f; }
bool System::usesPeriodicBoundaryConditions() const {
    for (const Force* f : forces_) if (f->usesPeriodicBoundaryConditions()) return true;
    return false;
}

struct PlatformRegistry {
    static std::vector<Platform*>& all() {
        // the stand-in answers to every platform name the reference's inputs use; all of them are this
        // double-precision CPU implementation
        static std::vector<Platform*> v = {new Platform("Reference"), new Platform("CPU"), new Platform("CUDA"), new Platform("OpenCL")};
        return v;
    }
};
const std::string& Platform::getOpenMMVersion() { static const std::string v = "8.0 (openmm-lite stand-in, oracle/ref_shim)"; return v; }
std::vector<std::string> Platform::loadPluginsFromDirectory(const std::string&) { return {}; }
std::vector<std::string> Platform::getPluginLoadFailures() { return {}; }
int Platform::getNumPlatforms() { return (int)PlatformRegistry::all().size(); }
Platform& Platform::getPlatform(int i) { return *PlatformRegistry::all().at(i); }
Platform& Platform::getPlatformByName(const std::string& name) {
    for (Platform* p : PlatformRegistry::all()) if (p->getName() == name) return *p;
    throw OpenMMException("There is no registered Platform called \"" + name + "\"");
}

// ================================================================================================
// State / Context
static const double kBoltz = 8.31446261815324e-3;  // kJ/mol/K

const std::vector<Vec3>& State::getPositions() const { if (!(types_ & Positions)) throw OpenMMException("Invoked getPositions() on a State which does not contain positions."); return pos_; }
const std::vector<Vec3>& State::getVelocities() const { if (!(types_ & Velocities)) throw OpenMMException("Invoked getVelocities() on a State which does not contain velocities."); return vel_; }
const std::vector<Vec3>& State::getForces() const { if (!(types_ & Forces)) throw OpenMMException("Invoked getForces() on a State which does not contain forces."); return frc_; }
double State::getKineticEnergy() const { if (!(types_ & Energy)) throw OpenMMException("Invoked getKineticEnergy() on a State which does not contain energies."); return ke_; }
double State::getPotentialEnergy() const { if (!(types_ & Energy)) throw OpenMMException("Invoked getPotentialEnergy() on a State which does not contain energies."); return pe_; }

static Context* g_most_recent = nullptr;
Context* Context::mostRecent() { return g_most_recent; }

Context::Context(const System& s, Integrator& i) : sys_(s), integ_(i), platform_(&Platform::getPlatformByName("Reference")) { init(); }
Context::Context(const System& s, Integrator& i, Platform& p) : sys_(s), integ_(i), platform_(&p) { init(); }
Context::Context(const System& s, Integrator& i, Platform& p, const std::map<std::string, std::string>&) : sys_(s), integ_(i), platform_(&p) { init(); }
Context::Context(Context&& o)
    : sys_(o.sys_), integ_(o.integ_), platform_(o.platform_), pos_(std::move(o.pos_)), vel_(std::move(o.vel_)), time_(o.time_),
      steps_(o.steps_), baro_(std::move(o.baro_)) {
    for (int c = 0; c < 3; c++) box_[c] = o.box_[c];
    o.moved_from_ = true;
    integ_.ctx_ = this;
    if (g_most_recent == &o) g_most_recent = this;
}
Context::~Context() {
    if (moved_from_) return;
    if (g_most_recent == this) g_most_recent = nullptr;
    integ_.ctx_ = nullptr;
}
void Context::init() {
    if (integ_.ctx_) throw OpenMMException("This Integrator is already bound to a context");
    integ_.ctx_ = this;
    const int n = sys_.getNumParticles();
    pos_.assign(n, Vec3());
    vel_.assign(n, Vec3());
    sys_.getDefaultPeriodicBoxVectors(box_[0], box_[1], box_[2]);
    for (int k = 0; k < sys_.getNumForces(); k++) {
        const Force* f = &sys_.getForce(k);
        if (auto* b = dynamic_cast<const MonteCarloAnisotropicBarostat*>(f)) {
            baro_.kind = 1; baro_.freq = b->freq_; baro_.T = b->t_;
            for (int c = 0; c < 3; c++) baro_.pressure[c] = b->p_[c];
            baro_.mode[0] = b->sx_; baro_.mode[1] = b->sy_; baro_.mode[2] = b->sz_;
            if (!b->sx_ && !b->sy_ && !b->sz_) baro_.freq = 0;
        } else if (auto* m = dynamic_cast<const MonteCarloMembraneBarostat*>(f)) {
            baro_.kind = 2; baro_.freq = m->freq_; baro_.T = m->t_; baro_.pressure[0] = m->p_; baro_.tension = m->g_;
            baro_.mode[0] = (int)m->xy_; baro_.mode[1] = (int)m->z_;
        }
    }
    if (baro_.kind) {
        if (!sys_.hasBox()) throw OpenMMException("barostat should be used with periodic boundary condition.");
        const double V = box_[0][0] * box_[1][1] * box_[2][2];
        for (int c = 0; c < 3; c++) baro_.vscale[c] = 0.01 * V;
        // molecules: connected components over every Force's bonded particles, in the order of their first particle
        std::vector<std::pair<int, int>> edges;
        for (int k = 0; k < sys_.getNumForces(); k++) sys_.getForce(k).bondedParticles(edges);
        std::vector<int> label(n);
        for (int i = 0; i < n; i++) label[i] = i;
        for (bool changed = true; changed;) {
            changed = false;
            for (auto& e : edges) {
                const int lo = std::min(label[e.first], label[e.second]);
                if (label[e.first] != lo || label[e.second] != lo) { label[e.first] = label[e.second] = lo; changed = true; }
            }
        }
        std::map<int, int> index;
        for (int i = 0; i < n; i++) {
            auto it = index.find(label[i]);
            if (it == index.end()) { it = index.emplace(label[i], (int)baro_.molecules.size()).first; baro_.molecules.emplace_back(); }
            baro_.molecules[it->second].push_back(i);
        }
    }
    g_most_recent = this;
}

// ContextImpl::updateContextState as the barostats implement it (OpenMM 8.0 MonteCarloAnisotropicBarostatImpl /
// MonteCarloMembraneBarostatImpl): a trial at every frequency-th call, before that step is integrated
static double lite_uniform(std::uint64_t& state) {
    state += 0x9E3779B97F4A7C15ull;
    std::uint64_t z = state;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}
void Context::updateContextState(std::uint64_t integrator_seed) {
    if (!baro_.kind || baro_.freq <= 0) return;
    if (!baro_.seeded) { baro_.rng = integrator_seed; baro_.seeded = true; }
    if (++baro_.calls < baro_.freq) return;
    baro_.calls = 0;
    barostatTrial();
}
void Context::barostatTrial() {
    Baro& B = baro_;
    const double bar = 6.02214076e23 * 1e-25;
    const double E0 = computeForcesAndEnergy(0xFFFFFFFF, nullptr);
    int axis;
    for (;;) {
        const double rnd = lite_uniform(B.rng) * 3.0;
        if (rnd < 1.0) { if (B.kind == 2 || B.mode[0]) { axis = 0; break; } }
        else if (rnd < 2.0) { if (B.kind == 2 ? B.mode[0] == 1 : B.mode[1] != 0) { axis = 1; break; } }
        else if (B.kind == 2 ? B.mode[1] == 0 : B.mode[2] != 0) { axis = 2; break; }
    }
    const Vec3 old_box[3] = {box_[0], box_[1], box_[2]};
    const double volume = box_[0][0] * box_[1][1] * box_[2][2];
    double deltaVolume = B.vscale[axis] * 2 * (lite_uniform(B.rng) - 0.5);
    double newVolume = volume + deltaVolume;
    double scale[3] = {1.0, 1.0, 1.0};
    if (B.kind == 2 && B.mode[0] == 0 && axis == 0) scale[0] = scale[1] = std::sqrt(newVolume / volume);
    else scale[axis] = newVolume / volume;
    if (B.kind == 2 && B.mode[1] == 2) { scale[2] = 1.0 / (scale[0] * scale[1]); newVolume = volume; deltaVolume = 0; }
    const std::vector<Vec3> saved = pos_;
    for (const auto& mol : B.molecules) {
        Vec3 centre;
        for (int i : mol) centre += pos_[i];
        centre = centre / (double)mol.size();
        Vec3 wrapped = centre;
        for (int c = 0; c < 3; c++) wrapped[c] -= std::floor(centre[c] / box_[c][c]) * box_[c][c];
        Vec3 shift;
        for (int c = 0; c < 3; c++) shift[c] = wrapped[c] * scale[c] - centre[c];
        for (int i : mol) pos_[i] += shift;
    }
    for (int c = 0; c < 3; c++) box_[c] = old_box[c] * scale[c];
    const double deltaArea = box_[0][0] * box_[1][1] - old_box[0][0] * old_box[1][1];
    const double E1 = computeForcesAndEnergy(0xFFFFFFFF, nullptr);
    const double kT = kBoltz * B.T;
    const double pressure = (B.kind == 2 ? B.pressure[0] : B.pressure[axis]) * bar;
    const double tension = (B.kind == 2 ? B.tension : 0.0) * bar;
    const double w = E1 - E0 + pressure * deltaVolume - tension * deltaArea - (double)B.molecules.size() * kT * std::log(newVolume / volume);
    if (w > 0 && lite_uniform(B.rng) > std::exp(-w / kT)) {
        pos_ = saved;
        for (int c = 0; c < 3; c++) box_[c] = old_box[c];
    } else B.accepted[axis]++;
    B.attempted[axis]++;
    if (B.attempted[axis] >= 10) {
        if (B.accepted[axis] < 0.25 * B.attempted[axis]) { B.vscale[axis] /= 1.1; B.attempted[axis] = B.accepted[axis] = 0; }
        else if (B.accepted[axis] > 0.75 * B.attempted[axis]) {
            B.vscale[axis] = std::min(B.vscale[axis] * 1.1, volume * 0.3);
            B.attempted[axis] = B.accepted[axis] = 0;
        }
    }
}
void Context::setPositions(const std::vector<Vec3>& p) {
    if ((int)p.size() != sys_.getNumParticles()) throw OpenMMException("Called setPositions() on a Context with the wrong number of positions");
    pos_ = p;
}
void Context::setVelocities(const std::vector<Vec3>& v) {
    if ((int)v.size() != sys_.getNumParticles()) throw OpenMMException("Called setVelocities() on a Context with the wrong number of velocities");
    vel_ = v;
}
void Context::setVelocitiesToTemperature(double T) {
    std::random_device rd;
    setVelocitiesToTemperature(T, (int)(rd() & 0x7fffffff));
}
// v = sqrt(kT/m) N(0,1) per component, no centre-of-mass removal; stream 1 of the Philox source
void Context::setVelocitiesToTemperature(double T, int seed) {
    for (int i = 0; i < sys_.getNumParticles(); i++) {
        double g[3];
        lite::philox_normals3((std::uint64_t)(std::int64_t)seed, (std::uint32_t)i, 0, 1, g);
        const double m = sys_.getParticleMass(i);
        const double s = m > 0 ? std::sqrt(kBoltz * T / m) : 0.0;
        vel_[i] = Vec3(s * g[0], s * g[1], s * g[2]);
    }
}
double Context::computeForcesAndEnergy(int groups, std::vector<Vec3>* forces) const {
    EvalContext ec;
    ec.pos = &pos_;
    ec.has_box = sys_.hasBox();
    ec.sys = &sys_;
    for (int c = 0; c < 3; c++) ec.box[c] = box_[c][c];
    if (forces) forces->assign(pos_.size(), Vec3());
    double E = 0;
    for (int k = 0; k < sys_.getNumForces(); k++) {
        const Force& f = sys_.getForce(k);
        if (!((unsigned)groups >> f.getForceGroup() & 1u)) continue;
        E += f.evaluate(ec, forces);
    }
    return E;
}
State Context::getState(int types, bool enforcePeriodicBox, int groups) const {
    State s;
    s.types_ = types;
    s.time_ = time_;
    for (int c = 0; c < 3; c++) s.box_[c] = box_[c];
    if (types & State::Positions) {
        s.pos_ = pos_;
        if (enforcePeriodicBox && sys_.hasBox())
            for (auto& p : s.pos_) for (int c = 0; c < 3; c++) p[c] -= box_[c][c] * std::floor(p[c] / box_[c][c]);
    }
    if (types & State::Velocities) s.vel_ = vel_;
    if (types & (State::Forces | State::Energy)) {
        std::vector<Vec3> F;
        s.pe_ = computeForcesAndEnergy(groups, (types & State::Forces) ? &F : nullptr);
        if (types & State::Forces) s.frc_ = std::move(F);
    }
    if (types & State::Energy) {
        // leap-frog kinetic energy: velocities shifted forward by dt/2 with the current forces (all groups)
        std::vector<Vec3> F;
        computeForcesAndEnergy(0xFFFFFFFF, &F);
        const double dt = integ_.getStepSize();
        double ke = 0;
        for (int i = 0; i < sys_.getNumParticles(); i++) {
            const double m = sys_.getParticleMass(i);
            if (!(m > 0)) continue;
            const Vec3 v = vel_[i] + F[i] * (0.5 * dt / m);
            ke += 0.5 * m * v.dot(v);
        }
        s.ke_ = ke;
    }
    return s;
}

// ================================================================================================
// LangevinIntegrator
LangevinIntegrator::LangevinIntegrator(double T, double friction, double dt) : T_(T), gamma_(friction) { dt_ = dt; }
void LangevinIntegrator::step(int steps) {
    if (!ctx_) throw OpenMMException("This Integrator is not bound to a context!");
    Context& c = *ctx_;
    const double a = std::exp(-gamma_ * dt_);
    const double b = gamma_ == 0 ? dt_ : (1.0 - a) / gamma_;
    const double cc = std::sqrt(kBoltz * T_ * (1.0 - a * a));
    std::vector<Vec3> F;
    const int n = c.sys_.getNumParticles();
    for (int s = 0; s < steps; s++) {
        c.updateContextState((std::uint64_t)(std::int64_t)seed_);
        c.computeForcesAndEnergy(0xFFFFFFFF, &F);
        for (int i = 0; i < n; i++) {
            const double m = c.sys_.getParticleMass(i);
            if (!(m > 0)) continue;
            double g[3];
            lite::philox_normals3((std::uint64_t)(std::int64_t)seed_, (std::uint32_t)i, (std::uint64_t)c.steps_, 0, g);
            const double im = 1.0 / m, sq = std::sqrt(im);
            Vec3& v = c.vel_[i];
            for (int k = 0; k < 3; k++) v[k] = a * v[k] + b * F[i][k] * im + cc * sq * g[k];
            c.pos_[i] += v * dt_;
        }
        c.steps_++;
        c.time_ += dt_;
    }
}

// ================================================================================================
// LocalEnergyMinimizer: L-BFGS (m = 10) with a backtracking Armijo line search
void LocalEnergyMinimizer::minimize(Context& ctx, double tolerance, int maxIterations) {
    const int n = ctx.sys_.getNumParticles(), N = 3 * n, M = 10;
    std::vector<double> x(N), g(N), xn(N), gn(N), dir(N);
    auto load = [&](const std::vector<double>& src) { for (int i = 0; i < n; i++) ctx.pos_[i] = Vec3(src[3 * i], src[3 * i + 1], src[3 * i + 2]); };
    auto eval = [&](const std::vector<double>& at, std::vector<double>& grad) {
        load(at);
        std::vector<Vec3> F;
        const double E = ctx.computeForcesAndEnergy(0xFFFFFFFF, &F);
        for (int i = 0; i < n; i++) for (int c = 0; c < 3; c++) grad[3 * i + c] = ctx.sys_.getParticleMass(i) > 0 ? -F[i][c] : 0.0;
        return E;
    };
    for (int i = 0; i < n; i++) for (int c = 0; c < 3; c++) x[3 * i + c] = ctx.pos_[i][c];
    double E = eval(x, g);
    std::vector<std::vector<double>> S, Y;
    std::vector<double> rho;
    const int itmax = maxIterations > 0 ? maxIterations : 100000;
    for (int it = 0; it < itmax; it++) {
        double g2 = 0;
        for (double v : g) g2 += v * v;
        if (std::sqrt(g2 / std::max(1, n)) < tolerance) break;
        dir = g;
        std::vector<double> alpha(S.size());
        for (int k = (int)S.size() - 1; k >= 0; k--) {
            double sd = 0;
            for (int i = 0; i < N; i++) sd += S[k][i] * dir[i];
            alpha[k] = rho[k] * sd;
            for (int i = 0; i < N; i++) dir[i] -= alpha[k] * Y[k][i];
        }
        if (!S.empty()) {
            double sy = 0, yy = 0;
            for (int i = 0; i < N; i++) { sy += S.back()[i] * Y.back()[i]; yy += Y.back()[i] * Y.back()[i]; }
            const double gam = sy / yy;
            for (double& v : dir) v *= gam;
        } else {
            const double sc = 0.01 / std::sqrt(std::max(g2, 1e-300));  // first step: 0.01 nm
            for (double& v : dir) v *= sc;
        }
        for (std::size_t k = 0; k < S.size(); k++) {
            double yd = 0;
            for (int i = 0; i < N; i++) yd += Y[k][i] * dir[i];
            const double beta = rho[k] * yd;
            for (int i = 0; i < N; i++) dir[i] += S[k][i] * (alpha[k] - beta);
        }
        double gd = 0;
        for (int i = 0; i < N; i++) gd -= g[i] * dir[i];
        if (!(gd < 0)) { S.clear(); Y.clear(); rho.clear(); continue; }
        double t = 1.0, En = E;
        bool ok = false;
        for (int ls = 0; ls < 40; ls++) {
            for (int i = 0; i < N; i++) xn[i] = x[i] - t * dir[i];
            En = eval(xn, gn);
            if (std::isfinite(En) && En <= E + 1e-4 * t * gd) { ok = true; break; }
            t *= 0.5;
        }
        if (!ok) {
            if (S.empty()) break;
            S.clear(); Y.clear(); rho.clear();
            continue;
        }
        std::vector<double> s(N), y(N);
        double sy = 0;
        for (int i = 0; i < N; i++) { s[i] = xn[i] - x[i]; y[i] = gn[i] - g[i]; sy += s[i] * y[i]; }
        if (sy > 1e-12) {
            S.push_back(s); Y.push_back(y); rho.push_back(1.0 / sy);
            if ((int)S.size() > M) { S.erase(S.begin()); Y.erase(Y.begin()); rho.erase(rho.begin()); }
        }
        x = xn; g = gn; E = En;
    }
    load(x);
}

// ================================================================================================
// description of a whole System (used by oracle/ref_shim/ref_api.cpp)
namespace lite {

std::string describe_system(const System& sys) {
    Describer d;
    d.os << "{";
    d.first = true;
    d.kv("num_particles", sys.getNumParticles());
    d.key("masses");
    std::vector<double> m(sys.getNumParticles());
    for (int i = 0; i < sys.getNumParticles(); i++) m[i] = sys.getParticleMass(i);
    d.nums(m);
    d.kb("has_box", sys.hasBox());
    Vec3 a, b, c;
    sys.getDefaultPeriodicBoxVectors(a, b, c);
    d.key("box");
    d.nums({a[0], b[1], c[2]});
    d.open("forces", '[');
    for (int k = 0; k < sys.getNumForces(); k++) {
        d.sep();
        d.os << "{";
        d.first = true;
        sys.getForce(k).describe(d);
        d.os << "}";
        d.first = false;
    }
    d.close(']');
    d.os << "}";
    return d.os.str();
}

}  // namespace lite
}  // namespace OpenMM
