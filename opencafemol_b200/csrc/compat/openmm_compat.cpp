// openmm_compat.cpp — implementation of compat/OpenMM.h on the C ABI of the engine (include/ocm_b200.h).
//
// SystemCompiler turns each recorded OpenMM::Force into one ocm_system_add_* call.  The catalogue below lists, for
// every generator of the reference that is in scope, the energy expression it hands to OpenMM
// (white space and the "<id>_" prefix of src/forcefield/ForceFieldIDGenerator.hpp removed) and the names of its
// parameters; the reference file that emits it is cited next to each entry.
#include "OpenMM.h"

#include <algorithm>
#include <cstring>
#include <functional>
#include <limits>
#include <random>

#include "../../../include/ocm_b200.h"

namespace OpenMM {

static void check(int rc, const char* what) {
    if (rc != 0) throw OpenMMException(std::string(what) + ": " + ocm_last_error());
}

// ---- recording ----------------------------------------------------------------------------------------------
void Force::setForceGroup(int g) {
    if (g < 0 || g > 31) throw OpenMMException("Force group must be between 0 and 31");
    group_ = g;
}
Continuous1DFunction::Continuous1DFunction(const std::vector<double>& values, double min, double max, bool)
    : y_(values), min_(min), max_(max) {
    if (values.size() < 2) throw OpenMMException("Continuous1DFunction: a tabulated function must have at least two points");
    if (!(max > min)) throw OpenMMException("Continuous1DFunction: max <= min");
}
int HarmonicBondForce::addBond(int p1, int p2, double length, double k) {
    p_.push_back(p1); p_.push_back(p2); length_.push_back(length); k_.push_back(k);
    return getNumBonds() - 1;
}
int HarmonicAngleForce::addAngle(int p1, int p2, int p3, double angle, double k) {
    p_.push_back(p1); p_.push_back(p2); p_.push_back(p3); angle_.push_back(angle); k_.push_back(k);
    return getNumAngles() - 1;
}
CustomForceBase::~CustomForceBase() { for (auto& f : functions_) delete f.second; }
int CustomForceBase::addGlobalParameter(const std::string& name, double value) { globals_.emplace_back(name, value); return (int)globals_.size() - 1; }
int CustomForceBase::addTabulatedFunction(const std::string& name, TabulatedFunction* fn) { functions_.emplace_back(name, fn); return (int)functions_.size() - 1; }
int CustomForceBase::add_row(std::vector<int> p, const std::vector<double>& v, const char* who) {
    if (v.size() != params_.size()) throw OpenMMException(std::string(who) + ": wrong number of parameters");
    rows_.push_back({std::move(p), v});
    return (int)rows_.size() - 1;
}
int CustomCompoundBondForce::addBond(const std::vector<int>& particles, const std::vector<double>& params) {
    if ((int)particles.size() != np_) throw OpenMMException("CustomCompoundBondForce: wrong number of particles");
    return add_row(particles, params, "CustomCompoundBondForce");
}
int CustomHbondForce::addAcceptor(int a1, int a2, int a3, const std::vector<double>& params) {
    if (params.size() != aparams_.size()) throw OpenMMException("CustomHbondForce acceptor: wrong number of parameters");
    arows_.push_back({{a1, a2, a3}, params});
    return (int)arows_.size() - 1;
}
int CustomHbondForce::addExclusion(int donor, int acceptor) {
    if (donor < 0 || donor >= (int)rows_.size() || acceptor < 0 || acceptor >= (int)arows_.size())
        throw OpenMMException("CustomHbondForce: exclusion index out of range");
    excl_.push_back(donor); excl_.push_back(acceptor);
    return (int)excl_.size() / 2 - 1;
}
System::~System() { for (Force* f : forces_) delete f; }
void System::setDefaultPeriodicBoxVectors(const Vec3& a, const Vec3& b, const Vec3& c) {
    if (a[1] != 0 || a[2] != 0 || b[0] != 0 || b[2] != 0 || c[0] != 0 || c[1] != 0)
        throw OpenMMException("B200 engine: only orthorhombic periodic boxes are supported (the reference only makes cuboids, src/SystemGenerator.cpp:52-55)");
    box_[0] = a; box_[1] = b; box_[2] = c;
    has_box_ = true;
}
bool System::usesPeriodicBoundaryConditions() const {
    for (const Force* f : forces_) if (f->usesPeriodicBoundaryConditions()) return true;
    return false;
}

// ---- platform -----------------------------------------------------------------------------------------------
struct PlatformRegistry {
    static std::vector<Platform*>& all() {
        // one engine; it also answers to the stock OpenMM platform names so that existing inputs
        // ([platform] name = "CUDA" | "CPU" | "Reference" | "OpenCL") run unchanged — getPlatformByName says so once
        static std::vector<Platform*> v = {new Platform("B200"), new Platform("CUDA"), new Platform("CPU"), new Platform("Reference"), new Platform("OpenCL")};
        return v;
    }
};
const std::string& Platform::getOpenMMVersion() {
    static const std::string v = std::string("8.0-compatible host layer of ") + ocm_version();
    return v;
}
std::vector<std::string> Platform::loadPluginsFromDirectory(const std::string&) { return {}; }
std::vector<std::string> Platform::getPluginLoadFailures() { return {}; }
int Platform::getNumPlatforms() { return (int)PlatformRegistry::all().size(); }
Platform& Platform::getPlatform(int i) { return *PlatformRegistry::all().at(i); }
Platform& Platform::getPlatformByName(const std::string& name) {
    for (Platform* p : PlatformRegistry::all()) if (p->getName() == name) {
        static bool told = false;
        if (!told && name != "B200" && name != "CUDA") {
            told = true;
            std::cerr << "[ocm-b200] platform \"" << name << "\" requested: this build has a single platform, the B200 CUDA engine" << std::endl;
        }
        return *p;
    }
    throw OpenMMException("There is no registered Platform called \"" + name + "\" (this build drives the B200 engine)");
}

// ---- Force -> C ABI -----------------------------------------------------------------------------------------
namespace b200 {

namespace {

std::string strip_ws(const std::string& s) {
    std::string o;
    for (char c : s) if (!std::isspace((unsigned char)c)) o += c;
    while (!o.empty() && o.back() == ';') o.pop_back();
    return o;
}
void replace_all(std::string& s, const std::string& from, const std::string& to) {
    if (from.empty()) return;
    for (std::size_t p = s.find(from); p != std::string::npos; p = s.find(from, p + to.size())) s.replace(p, from.size(), to);
}
// Canonical form under which an energy expression is looked up: white space removed, trailing ';' dropped, and the auxiliary
// definitions that follow the main expression ("energy; a = ...; b = ...") sorted, since Lepton does not care about their order.
// An upstream edit that only re-indents or re-orders the definitions of a generator's string still finds its catalogue entry.
std::string canonical(const std::string& expr) {
    const std::string s = strip_ws(expr);
    std::vector<std::string> parts;
    std::size_t b = 0;
    for (std::size_t p = s.find(';'); ; p = s.find(';', b)) {
        parts.push_back(s.substr(b, p == std::string::npos ? std::string::npos : p - b));
        if (p == std::string::npos) break;
        b = p + 1;
    }
    parts.erase(std::remove_if(parts.begin() + 1, parts.end(), [](const std::string& t) { return t.empty(); }), parts.end());
    std::sort(parts.begin() + 1, parts.end());
    std::string o = parts[0];
    for (std::size_t k = 1; k < parts.size(); k++) o += ";" + parts[k];
    return o;
}

// expressions emitted by the reference's generators, normalised
const std::string kGoContact = canonical("k*(5*(r0/r)^12-6*(r0/r)^10)");                       // GoContactForceFieldGenerator.cpp:5-18
const std::string kCappedGoContact = canonical(  // CappedGoContactForceFieldGenerator.cpp:5-9
    "cap_val*cap_formula+(1-cap_val)*lj_formula;cap_formula=k*slope/roverdist+k*intercept;lj_formula=k*(5*roverdist^12-6*roverdist^10);"
    "cap_val=step(cap-1/roverdist);roverdist=r0/r");
const std::string kGaussianBond = canonical("k*exp(-(r-v0)^2/(2*sigma^2))");                   // GaussianBondForceFieldGenerator.cpp:5-21
const std::string k3SPN2Bond = canonical("k2*(r-v0)^2+100.0*k4*(r-v0)^4");                     // ThreeSPN2BondForceFieldGenerator.cpp:5-20
const std::string kFLA = canonical(  // FlexibleLocalAngleForceFieldGenerator.cpp:21-27
    "k*(spline(theta)-min_energy+(step(min_theta-theta)*(30*(min_theta-theta)+min_theta_y))+(step(theta-max_theta)*(30*(theta-max_theta)+max_theta_y)));"
    "theta=angle(p1,p2,p3)");
const std::string kBaseStacking = canonical(  // ThreeSPN2BaseStackingForceFieldGenerator.cpp:6-22
    "rep+f2*attr;rep=epsilon*(1-exp(-alpha*(dr)))^2*step(-dr);attr=epsilon*(1-exp(-alpha*(dr)))^2*step(dr)-epsilon;dr=distance(p2,p3)-r0;"
    "f2=max(f*rect2,rect1);rect1=step(dt+pi/2)*step(pi/2-dt);rect2=step(dt+pi)*step(pi-dt);f=1-cos(dt)^2;dt=K_BS*(angle(p1,p2,p3)-t0);"
    "pi=3.1415926535897932385");
const std::string kGaussianDihedral = canonical(  // GaussianDihedralForceFieldGenerator.cpp:5-13
    "k*exp(-(dt_periodic)^2/(2*sigma^2));dt_periodic=dt-floor((dt+pi)/(2*pi))*(2*pi);dt=theta-theta0;pi=3.1415926535897932385");
const std::string kCosineDihedral = canonical("k_periodic*(1-cos(n0*(theta-t0)))");            // CosineDihedralForceFieldGenerator.cpp:7
const std::string kFLD = canonical(  // FlexibleLocalDihedralForceFieldGenerator.cpp:28-37
    "k*(c-min_energy+kcos1*cos(theta)+ksin1*sin(theta)+kcos2*cos(2*theta)+ksin2*sin(2*theta)+kcos3*cos(3*theta)+ksin3*sin(3*theta))");
const std::string kExcludedVolume = canonical(  // ExcludedVolumeForceFieldGenerator.cpp:25-27
    "epsilon*(sigma_sum)^12*((1/r)^12-cutoff_correction);sigma_sum=sigma1+sigma2");
const std::string kDebyeHuckel = canonical("q1*q2*inv_4_pi_eps0_epsk*(exp(-r/debye_length)/r-cutoff_correction)");  // DebyeHuckelForceFieldGenerator.cpp:54-57
const std::string kLJRepulsive = canonical(  // LennardJonesRepulsiveForceFieldGenerator.cpp:35-43
    "step(threthold-r)*epsilon*(4*(sigma_r_12-sigma_r_6)+1);sigma_r_12=sigma_r_6^2;sigma_r_6=sigma_r^6;sigma_r=sigma/r;threthold=sigma*2^(1/6);"
    "sigma=(sigma1+sigma2)*0.5;epsilon=(epsilon1+epsilon2)*0.5");
const std::string kLJAttractive = canonical(  // LennardJonesAttractiveForceFieldGenerator.cpp:40-50
    "epsilon*(step(r-threthold)*4*(sigma_r_12-sigma_r_6)-step(threthold-r));sigma_r_12=sigma_r_6^2;sigma_r_6=sigma_r^6;sigma_r=sigma/r;"
    "threthold=sigma*2^(1/6);sigma=(sigma1+sigma2)*0.5;epsilon=(epsilon1+epsilon2)*0.5");
const std::string kWCA = canonical(  // WeeksChandlerAndersenForceFieldGenerator.cpp:34-41
    "step(sigma*2^(1/6)-r)*(4*eps*(sigma_r_12-sigma_r_6)+eps);sigma_r_12=sigma_r_6^2;sigma_r_6=sigma_r^6;sigma_r=sigma/r;"
    "eps=sqrt(eps1*eps2);sigma=(sigma1+sigma2)*0.5");
const std::string kISoLFAttractive = canonical(  // iSoLFAttractiveForceFieldGenerator.cpp:41-50
    "-eps*(step(base-r)+step(r-base)*step(base+omega-r)*cos(0.5*pi*(r-base)/omega)^2);base=sigma*2^(1/6);sigma=(sigma1+sigma2)*0.5;"
    "eps=sqrt(eps1*eps2);omega=(omega1+omega2)*0.5;pi=3.1415926535897932385");
const std::string kTrigonometric = canonical(  // TrigonometricForceFieldGenerator.cpp:43-53
    "-epsilon*(step(threthold+omega-r)+step(r-threthold)*step(threthold+omega-r)*(2*r_th_omeg^3-3*r_th_omeg^2));r_th_omeg=(r-threthold)/omega;"
    "threthold=sigma*2^(1/6);epsilon=(epsilon1*epsilon2)^(1/2);sigma=(sigma1+sigma2)*0.5;omega=(omega1+omega2)*0.5");
const std::string kUniformWCA = canonical(  // UniformWeeksChandlerAndersenForceFieldGenerator.cpp:138-143
    "4*eps*(sigma_r_12-sigma_r_6+0.25);sigma_r_12=sigma_r_6^2;sigma_r_6=sigma_r^6;sigma_r=sigma/r");
const std::string kUniformLJAttractive = canonical(  // UniformLennardJonesAttractiveForceFieldGenerator.cpp:145-152
    "epsilon*(step(r-threthold)*4*(sigma_r_12-sigma_r_6)-step(threthold-r));sigma_r_12=sigma_r_6^2;sigma_r_6=sigma_r^6;sigma_r=sigma/r;"
    "threthold=sigma*2^(1/6)");
const std::string k3SPN2ExcludedVolume = canonical(  // ThreeSPN2ExcludedVolumeForceFieldGenerator.cpp:21-26
    "step(sigma-r)*(epsilon*((sigma/r)^12-2*(sigma/r)^6)+epsilon);epsilon=sqrt(epsilon1*epsilon2);sigma=0.5*(sigma1+sigma2)");
const std::string kBasePair = canonical(  // ThreeSPN2BasePairForceFieldGenerator.cpp:8-26
    "rep+1/2*(1+cos(dphi))*fdt1*fdt2*attr;rep=epsilon*(1-exp(-alpha_BP*dr))^2*(1-step(dr));attr=epsilon*(1-exp(-alpha_BP*dr))^2*step(dr)-epsilon;"
    "fdt1=max(f1*rect0t1,rect1t1);fdt2=max(f2*rect0t2,rect1t2);rect1t1=step(pi/2+dt1)*step(pi/2-dt1);rect1t2=step(pi/2+dt2)*step(pi/2-dt2);"
    "rect0t1=step(pi+dt1)*step(pi-dt1);rect0t2=step(pi+dt2)*step(pi-dt2);f1=1-cos(dt1)^2;f2=1-cos(dt2)^2;dphi=dihedral(d2,d1,a1,a2)-phi0;"
    "dr=distance(d1,a1)-r0;dt1=K_BP*(angle(d2,d1,a1)-t01);dt2=K_BP*(angle(a2,a1,d1)-t02);pi=3.1415926535897932385");
const std::string kCrossStacking = canonical(  // ThreeSPN2CrossStackingForceFieldGenerator.cpp:8-31
    "fdt3*fdtCS*attr/2;attr=epsilon*(1-exp(-alpha_CS*dr))^2*step(dr)-epsilon;fdt3=max(f1*rect0t3,rect1t3);fdtCS=max(f2*rect0tCS,rect1tCS);"
    "rect0t3=step(pi+dt3)*step(pi-dt3);rect0tCS=step(pi+dtCS)*step(pi-dtCS);rect1t3=step(pi/2+dt3)*step(pi/2-dt3);"
    "rect1tCS=step(pi/2+dtCS)*step(pi/2-dtCS);f1=1-cos(dt3)^2;f2=1-cos(dtCS)^2;dr=distance(d1,a3)-r0;dt3=K_BP*(t3-t03);dtCS=K_CS*(tCS-t0CS);"
    "tCS=angle(d2,d1,a3);t3=acos(cost3lim);cost3lim=min(max(cost3,-0.99),0.99);cost3=sin(t1)*sin(t2)*cos(phi)-cos(t1)*cos(t2);"
    "t1=angle(d2,d1,a1);t2=angle(d1,a1,a2);phi=dihedral(d2,d1,a1,a2);pi=3.1415926535897932385");
// the *Local variants: the same energies over an explicit list of (donor, acceptor) nucleotide pairs, as compound bonds
const std::string kBasePairLocal = canonical(  // ThreeSPN2BasePairLocalForceFieldGenerator.cpp:6-25
    "rep+1/2*(1+cos(dphi))*fdt1*fdt2*attr;rep=epsilon*(1-exp(-alpha_BP*dr))^2*(1-step(dr));attr=epsilon*(1-exp(-alpha_BP*dr))^2*step(dr)-epsilon;"
    "fdt1=max(f1*rect0t1,rect1t1);fdt2=max(f2*rect0t2,rect1t2);rect1t1=step(pi/2+dt1)*step(pi/2-dt1);rect1t2=step(pi/2+dt2)*step(pi/2-dt2);"
    "rect0t1=step(pi+dt1)*step(pi-dt1);rect0t2=step(pi+dt2)*step(pi-dt2);f1=1-cos(dt1)^2;f2=1-cos(dt2)^2;dphi=dihedral(p1,p2,p3,p4)-phi0;"
    "dr=distance(p2,p3)-r0;dt1=K_BP*(angle(p1,p2,p3)-t01);dt2=K_BP*(angle(p4,p3,p2)-t02);pi=3.1415926535897932385");
const std::string kCrossStackingLocal = canonical(  // ThreeSPN2CrossStackingLocalForceFieldGenerator.cpp:6-30
    "fdt3*fdtCS*attr/2;attr=epsilon*(1-exp(-alpha_CS*dr))^2*step(dr)-epsilon;fdt3=max(f1*rect0t3,rect1t3);fdtCS=max(f2*rect0tCS,rect1tCS);"
    "rect0t3=step(pi+dt3)*step(pi-dt3);rect0tCS=step(pi+dtCS)*step(pi-dtCS);rect1t3=step(pi/2+dt3)*step(pi/2-dt3);"
    "rect1tCS=step(pi/2+dtCS)*step(pi/2-dtCS);f1=1-cos(dt3)^2;f2=1-cos(dtCS)^2;dr=distance(p2,p6)-r0;dt3=K_BP*(t3-t03);dtCS=K_CS*(tCS-t0CS);"
    "tCS=angle(p1,p2,p6);t3=acos(cost3lim);cost3lim=min(max(cost3,-0.99),0.99);cost3=sin(t1)*sin(t2)*cos(phi)-cos(t1)*cos(t2);"
    "t1=angle(p1,p2,p3);t2=angle(p2,p3,p4);phi=dihedral(p1,p2,p3,p4);pi=3.1415926535897932385");
// Range inside which the *Local pairs are evaluated.  The compound bonds have no cutoff, but their energy beyond
// r0 + 1.9 nm is below eps * 2 exp(-alpha * 1.9 nm) < 1e-15 kJ/mol (alpha >= 20 / nm): nothing is lost at double precision.
constexpr double kLocalPairRange = 2.5;

const std::string kPDNS = canonical(  // ProteinDNANonSpecificForceFieldGenerator.cpp:60-83
    "-k*f_r*g_theta*g_phi;f_r=exp(-dr^2/(2*sigma*sigma));g_theta=max(g1*rect0t,rect1t);g_phi=max(g2*rect0p,rect1p);"
    "rect0t=step(pi+dtheta)*step(pi-dtheta);rect1t=step(pi/2+dtheta)*step(pi/2-dtheta);rect0p=step(pi+dphi)*step(pi-dphi);"
    "rect1p=step(pi/2+dphi)*step(pi/2-dphi);g1=1-cos(dtheta)^2;g2=1-cos(dphi)^2;dr=distance(d2,a1)-r0;dtheta=K*(theta-theta0);dphi=K*(phi-phi0);"
    "K=pi/(2*delta);phi=angle(d2,a1,a2);theta=acos(cost1lim);cost1lim=min(max(cost1,-0.99999),0.99999);"
    "cost1=sin(t1)*sin(t2)*cos(phi1)-cos(t1)*cos(t2);t1=angle(d1,d3,a1);t2=angle(d3,a1,d2);phi1=dihedral(d1,d3,a1,d2);pi=3.1415926535897932385");

// one-bead external terms, the centroid spring and the combinatorial contact
const std::string kPositionRestraint = canonical("k*(((x-x0)^2+(y-y0)^2+(z-z0)^2)^(1/2)-v0)^2");  // PositionRestraintForceFieldGenerator.cpp:8-10
const std::string kPulling = canonical("-kx*x-ky*y-kz*z");                                        // PullingForceFieldGenerator.cpp:5-7
const std::string kRectangularBox = canonical(  // EXVRectangularBoxForceFieldGenerator.cpp:33-38
    "epsilon*((radius/(x-x_lower))^12+(radius/(x_upper-x))^12+(radius/(y-y_lower))^12+(radius/(y_upper-y))^12+(radius/(z-z_lower))^12+"
    "(radius/(z_upper-z))^12)");
const std::string kCylindricalRestraint = canonical(  // CylindricalRestraintForceFieldGenerator.cpp:10-17
    "k*(dr-v0)^2;dr=((x-xh)^2+(y-yh)^2+(z-zh)^2)^(1/2);xh=ra_r*xa+x0;yh=ra_r*ya+y0;zh=ra_r*za+z0;ra_r=xa*x+ya*y+za*z");
const std::string kCylindricalRestraintPeriodic = canonical(  // CylindricalRestraintForceFieldGenerator.cpp:19-26
    "k*(dr-v0)^2;dr=periodicdistance(x,y,z,xh,yh,zh);xh=ra_r*xa+x0;yh=ra_r*ya+y0;zh=ra_r*za+z0;ra_r=xa*x+ya*y+za*z");
const std::string kHarmonicCoMPulling = canonical("0.5*k*(distance(g1,g2)-v0)^2");                // HarmonicCoMPullingForceFieldGenerator.cpp:5-7
const std::string kCombinatorialGoContact = canonical("k*(5*(r0/distance(a1,d1))^12-6*(r0/distance(a1,d1))^10)");  // CombinatorialGoContact...cpp:5-10

}  // namespace

class SystemCompiler {
  public:
    explicit SystemCompiler(const System& sys) : sys_(sys) {}

    ocm_system* compile() {
        const int n = sys_.getNumParticles();
        double box[3] = {sys_.box_[0][0], sys_.box_[1][1], sys_.box_[2][2]};
        check(ocm_system_create(n, sys_.mass_.data(), sys_.has_box_ ? box : nullptr, &out_), "ocm_system_create");
        try {
            for (const Force* f : sys_.forces_) add(*f);
        } catch (...) {
            ocm_system_destroy(out_);
            throw;
        }
        return out_;
    }

  private:
    const System& sys_;
    ocm_system* out_ = nullptr;

    // expression and parameter names with the generator id prefix removed
    struct Norm {
        std::string expr;
        std::vector<std::string> params, aparams, globals;
    };
    static Norm normalise(const CustomForceBase& f, const std::vector<std::string>* aparams = nullptr) {
        Norm nz;
        // the generator id ("EXV0", "TSPN2_CS7", ...) is the longest common prefix of all parameter names, cut back
        // to its last underscore
        std::vector<std::string> names = f.params_;
        if (aparams) names.insert(names.end(), aparams->begin(), aparams->end());
        for (auto& g : f.globals_) names.push_back(g.first);
        std::string prefix = names.empty() ? std::string() : names[0];
        for (auto& nm : names) {
            std::size_t k = 0;
            while (k < prefix.size() && k < nm.size() && prefix[k] == nm[k]) ++k;
            prefix.resize(k);
        }
        const std::size_t us = prefix.rfind('_');
        prefix = us == std::string::npos ? std::string() : prefix.substr(0, us + 1);
        auto strip = [&](const std::string& s) { return (!prefix.empty() && s.compare(0, prefix.size(), prefix) == 0) ? s.substr(prefix.size()) : s; };
        nz.expr = f.expr_;
        replace_all(nz.expr, prefix, "");
        nz.expr = canonical(nz.expr);
        for (auto& p : f.params_) nz.params.push_back(strip(p));
        if (aparams) for (auto& p : *aparams) nz.aparams.push_back(strip(p));
        for (auto& g : f.globals_) nz.globals.push_back(strip(g.first));
        return nz;
    }
    [[noreturn]] static void unsupported(const char* cls, const CustomForceBase& f) {
        throw OpenMMException(std::string("B200 engine: this ") + cls + " energy expression is not one of the potentials of the hot-path scope "
                              "(there is no expression interpreter and no CPU fallback): " + f.expr_);
    }
    static void expect_names(const std::vector<std::string>& got, std::initializer_list<const char*> want, const char* what) {
        bool ok = got.size() == want.size();
        std::size_t i = 0;
        if (ok) for (const char* w : want) ok = ok && got[i++] == w;
        if (!ok) {
            std::string s = std::string("B200 engine: unexpected parameter list for ") + what + ":";
            for (auto& g : got) s += " " + g;
            throw OpenMMException(s);
        }
    }
    static double global_value(const CustomForceBase& f, const Norm& nz, const char* name) {
        for (std::size_t i = 0; i < nz.globals.size(); i++) if (nz.globals[i] == name) return f.globals_[i].second;
        throw OpenMMException(std::string("B200 engine: global parameter \"") + name + "\" is missing");
    }
    static std::vector<std::int32_t> indices(const std::vector<CustomForceBase::Row>& rows, int per) {
        std::vector<std::int32_t> out;
        out.reserve(rows.size() * per);
        for (auto& r : rows) for (int k = 0; k < per; k++) out.push_back(r.p[k]);
        return out;
    }
    static std::vector<double> column(const std::vector<CustomForceBase::Row>& rows, int col) {
        std::vector<double> out(rows.size());
        for (std::size_t i = 0; i < rows.size(); i++) out[i] = rows[i].v[col];
        return out;
    }
    static double uniform(const std::vector<CustomForceBase::Row>& rows, int col, const char* what) {
        if (rows.empty()) return 0.0;
        for (auto& r : rows) if (r.v[col] != rows[0].v[col])
            throw OpenMMException(std::string("B200 engine: ") + what + " must be the same for every entry of one Force");
        return rows[0].v[col];
    }

    void add(const Force& f) {
        const int g = f.getForceGroup();
        if (auto* p = dynamic_cast<const HarmonicBondForce*>(&f)) {
            check(ocm_system_add_harmonic_bond(out_, g, p->periodic_, p->getNumBonds(), p->p_.data(), p->length_.data(), p->k_.data()), "HarmonicBondForce");
        } else if (auto* p = dynamic_cast<const HarmonicAngleForce*>(&f)) {
            check(ocm_system_add_harmonic_angle(out_, g, p->periodic_, p->getNumAngles(), p->p_.data(), p->angle_.data(), p->k_.data()), "HarmonicAngleForce");
        } else if (auto* p = dynamic_cast<const CustomBondForce*>(&f)) {
            add_bond(*p, g);
        } else if (auto* p = dynamic_cast<const CustomTorsionForce*>(&f)) {
            add_torsion(*p, g);
        } else if (auto* p = dynamic_cast<const CustomCompoundBondForce*>(&f)) {
            add_compound(*p, g);
        } else if (auto* p = dynamic_cast<const CustomNonbondedForce*>(&f)) {
            add_nonbonded(*p, g);
        } else if (auto* p = dynamic_cast<const CustomHbondForce*>(&f)) {
            add_hbond(*p, g);
        } else if (dynamic_cast<const MonteCarloAnisotropicBarostat*>(&f) || dynamic_cast<const MonteCarloMembraneBarostat*>(&f)) {
            // not a term of the system tables: Context::create hands it to ocm_context_set_barostat
            if (!sys_.has_box_) throw OpenMMException("barostat should be used with periodic boundary condition.");
        } else if (auto* p = dynamic_cast<const CustomExternalForce*>(&f)) {
            add_external(*p, g);
        } else if (auto* p = dynamic_cast<const CustomCentroidBondForce*>(&f)) {
            add_centroid(*p, g);
        } else if (auto* p = dynamic_cast<const CustomForceBase*>(&f)) {
            unsupported("Custom force", *p);
        } else {
            throw OpenMMException("B200 engine: unknown Force class");
        }
    }

    void add_bond(const CustomBondForce& f, int g) {
        const Norm nz = normalise(f);
        const auto ij = indices(f.rows_, 2);
        const std::int64_t n = (std::int64_t)f.rows_.size();
        if (nz.expr == kGoContact) {
            expect_names(nz.params, {"k", "r0"}, "GoContact");
            check(ocm_system_add_go_contact(out_, g, f.periodic_, n, ij.data(), column(f.rows_, 0).data(), column(f.rows_, 1).data()), "GoContact");
        } else if (nz.expr == kCappedGoContact) {
            expect_names(nz.params, {"k", "r0", "slope", "intercept", "cap"}, "CappedGoContact");
            check(ocm_system_add_capped_go_contact(out_, g, f.periodic_, n, ij.data(), column(f.rows_, 0).data(), column(f.rows_, 1).data(), column(f.rows_, 2).data(),
                                                   column(f.rows_, 3).data(), column(f.rows_, 4).data()), "CappedGoContact");
        } else if (nz.expr == kGaussianBond) {
            expect_names(nz.params, {"k", "v0", "sigma"}, "GaussianBond");
            check(ocm_system_add_gaussian_bond(out_, g, f.periodic_, n, ij.data(), column(f.rows_, 0).data(), column(f.rows_, 1).data(), column(f.rows_, 2).data()), "GaussianBond");
        } else if (nz.expr == k3SPN2Bond) {
            expect_names(nz.params, {"k2", "k4", "v0"}, "3SPN2Bond");
            std::vector<double> k4 = column(f.rows_, 1);
            for (double& v : k4) v *= 100.0;  // the expression's literal factor; the C ABI takes the final coefficient
            check(ocm_system_add_3spn2_bond(out_, g, f.periodic_, n, ij.data(), column(f.rows_, 0).data(), k4.data(), column(f.rows_, 2).data()), "3SPN2Bond");
        } else {
            unsupported("CustomBondForce", f);
        }
    }

    void add_torsion(const CustomTorsionForce& f, int g) {
        const Norm nz = normalise(f);
        const auto ijkl = indices(f.rows_, 4);
        const std::int64_t n = (std::int64_t)f.rows_.size();
        if (nz.expr == kGaussianDihedral) {
            expect_names(nz.params, {"k", "theta0", "sigma"}, "GaussianDihedral");
            check(ocm_system_add_gaussian_dihedral(out_, g, f.periodic_, n, ijkl.data(), column(f.rows_, 0).data(), column(f.rows_, 1).data(), column(f.rows_, 2).data()), "GaussianDihedral");
        } else if (nz.expr == kCosineDihedral) {
            expect_names(nz.params, {"k_periodic", "t0", "n0"}, "CosineDihedral");
            std::vector<std::int32_t> nper(f.rows_.size());
            for (std::size_t i = 0; i < f.rows_.size(); i++) {
                const double v = f.rows_[i].v[2];
                if (v != std::floor(v)) throw OpenMMException("B200 engine: CosineDihedral periodicity must be an integer");
                nper[i] = (std::int32_t)v;
            }
            check(ocm_system_add_cosine_dihedral(out_, g, f.periodic_, n, ijkl.data(), column(f.rows_, 0).data(), column(f.rows_, 1).data(), nper.data()), "CosineDihedral");
        } else if (nz.expr == kFLD) {
            expect_names(nz.params, {"k", "c", "kcos1", "ksin1", "kcos2", "ksin2", "kcos3", "ksin3", "min_energy"}, "FlexibleLocalDihedral");
            // one C-ABI call per distinct (coefficients, min_energy): the reference makes one Force per residue-pair class
            std::vector<char> done(f.rows_.size(), 0);
            for (std::size_t a = 0; a < f.rows_.size(); a++) {
                if (done[a]) continue;
                std::vector<std::int32_t> idx;
                std::vector<double> k;
                for (std::size_t b = a; b < f.rows_.size(); b++) {
                    if (done[b] || !std::equal(f.rows_[a].v.begin() + 1, f.rows_[a].v.end(), f.rows_[b].v.begin() + 1)) continue;
                    done[b] = 1;
                    idx.insert(idx.end(), f.rows_[b].p.begin(), f.rows_[b].p.end());
                    k.push_back(f.rows_[b].v[0]);
                }
                check(ocm_system_add_flexible_local_dihedral(out_, g, f.periodic_, (std::int64_t)k.size(), idx.data(), k.data(), f.rows_[a].v.data() + 1, f.rows_[a].v[8]), "FlexibleLocalDihedral");
            }
        } else {
            unsupported("CustomTorsionForce", f);
        }
    }

    void add_compound(const CustomCompoundBondForce& f, int g) {
        const Norm nz = normalise(f);
        const std::int64_t n = (std::int64_t)f.rows_.size();
        if (nz.expr == kFLA && f.np_ == 3) {
            expect_names(nz.params, {"k"}, "FlexibleLocalAngle");
            const Continuous1DFunction* spline = nullptr;
            for (auto& t : f.functions_) if (t.first == "spline") spline = dynamic_cast<const Continuous1DFunction*>(t.second);
            if (!spline) throw OpenMMException("B200 engine: FlexibleLocalAngle without its tabulated function \"spline\"");
            std::vector<double> y;
            double lo, hi;
            spline->getFunctionParameters(y, lo, hi);
            if (y.size() != 10) throw OpenMMException("B200 engine: the FlexibleLocalAngle table must have 10 knots (src/util/Constants.hpp:26-49)");
            if (global_value(f, nz, "min_theta") != lo || global_value(f, nz, "max_theta") != hi || global_value(f, nz, "min_theta_y") != y.front() ||
                global_value(f, nz, "max_theta_y") != y.back())
                throw OpenMMException("B200 engine: FlexibleLocalAngle walls do not start at the ends of the spline table");
            const auto ijk = indices(f.rows_, 3);
            check(ocm_system_add_flexible_local_angle(out_, g, f.periodic_, n, ijk.data(), column(f.rows_, 0).data(), y.data(), lo, hi, global_value(f, nz, "min_energy")), "FlexibleLocalAngle");
        } else if (nz.expr == kBaseStacking && f.np_ == 3) {
            expect_names(nz.params, {"epsilon", "r0", "t0", "alpha", "K_BS"}, "3SPN2BaseStacking");
            const auto ijk = indices(f.rows_, 3);
            check(ocm_system_add_3spn2_base_stacking(out_, g, f.periodic_, n, ijk.data(), column(f.rows_, 0).data(), column(f.rows_, 1).data(), column(f.rows_, 2).data(),
                                                     uniform(f.rows_, 3, "3SPN2BaseStacking alpha"), uniform(f.rows_, 4, "3SPN2BaseStacking K_BS")), "3SPN2BaseStacking");
        } else if ((nz.expr == kBasePairLocal && f.np_ == 4) || (nz.expr == kCrossStackingLocal && f.np_ == 6)) {
            // explicit nucleotide pairs = a donor-acceptor term in which bond i contributes donor i and acceptor i and every
            // other combination is excluded
            const bool bp = f.np_ == 4;
            if (bp) expect_names(nz.params, {"epsilon", "r0", "t01", "t02", "phi0", "alpha_BP", "K_BP"}, "3SPN2BasePairLocal");
            else expect_names(nz.params, {"epsilon", "r0", "t03", "t0CS", "K_BP", "K_CS", "alpha_CS"}, "3SPN2CrossStackingLocal");
            std::vector<std::int32_t> donors, acceptors, excl;
            for (auto& r : f.rows_) {
                // base pair: (p1, p2, p3, p4) = (sugar_d, base_d, base_a, sugar_a); cross stacking adds p6 = the cross base
                donors.insert(donors.end(), {r.p[1], r.p[0], -1});
                acceptors.insert(acceptors.end(), {r.p[2], r.p[3], bp ? -1 : r.p[5]});
            }
            for (std::int32_t d = 0; d < (std::int32_t)n; d++)
                for (std::int32_t a = 0; a < (std::int32_t)n; a++)
                    if (a != d) { excl.push_back(d); excl.push_back(a); }
            ocm_da_topology t{};
            t.n_donors = n; t.donors = donors.data(); t.n_acceptors = n; t.acceptors = acceptors.data();
            t.n_excl = (std::int64_t)excl.size() / 2; t.excl_da = excl.data();
            t.cutoff = kLocalPairRange;
            t.periodic = f.periodic_;
            if (bp)
                check(ocm_system_add_3spn2_base_pair(out_, g, &t, column(f.rows_, 0).data(), column(f.rows_, 1).data(), column(f.rows_, 2).data(), column(f.rows_, 3).data(),
                                                     column(f.rows_, 4).data(), uniform(f.rows_, 5, "3SPN2BasePairLocal alpha_BP"), uniform(f.rows_, 6, "3SPN2BasePairLocal K_BP")), "3SPN2BasePairLocal");
            else
                check(ocm_system_add_3spn2_cross_stacking(out_, g, &t, column(f.rows_, 0).data(), column(f.rows_, 1).data(), column(f.rows_, 2).data(), column(f.rows_, 3).data(),
                                                          uniform(f.rows_, 4, "3SPN2CrossStackingLocal K_BP"), uniform(f.rows_, 5, "3SPN2CrossStackingLocal K_CS"),
                                                          uniform(f.rows_, 6, "3SPN2CrossStackingLocal alpha_CS")), "3SPN2CrossStackingLocal");
        } else {
            unsupported("CustomCompoundBondForce", f);
        }
    }

    struct PairTopo {
        ocm_pair_topology t{};
        std::vector<std::int64_t> off1, off2;
        std::vector<std::int32_t> s1, s2;
    };
    void fill_topology(const CustomNonbondedForce& f, PairTopo& pt) {
        if (f.method_ == CustomNonbondedForce::NoCutoff) throw OpenMMException("B200 engine: CustomNonbondedForce needs a cutoff method (the reference always sets one)");
        pt.t.n_excl = (std::int64_t)f.excl_.size() / 2;
        pt.t.excl_ij = f.excl_.data();
        pt.t.n_igroups = (std::int32_t)f.groups_.size();
        pt.off1.push_back(0); pt.off2.push_back(0);
        for (auto& gpair : f.groups_) {
            pt.s1.insert(pt.s1.end(), gpair.first.begin(), gpair.first.end());
            pt.s2.insert(pt.s2.end(), gpair.second.begin(), gpair.second.end());
            pt.off1.push_back((std::int64_t)pt.s1.size());
            pt.off2.push_back((std::int64_t)pt.s2.size());
        }
        pt.t.set1_off = pt.off1.data(); pt.t.set1 = pt.s1.data();
        pt.t.set2_off = pt.off2.data(); pt.t.set2 = pt.s2.data();
        pt.t.periodic = f.method_ == CustomNonbondedForce::CutoffPeriodic;
    }

    void add_nonbonded(const CustomNonbondedForce& f, int g) {
        if ((int)f.rows_.size() != sys_.getNumParticles())
            throw OpenMMException("CustomNonbondedForce must have exactly as many particles as the System it belongs to.");
        const Norm nz = normalise(f);
        PairTopo pt;
        fill_topology(f, pt);
        if (f.use_switch_ && nz.expr != kLJAttractive && nz.expr != kUniformLJAttractive)
            throw OpenMMException("B200 engine: a switching function is only supported for the LennardJonesAttractive potentials");
        if (nz.expr == kExcludedVolume) {
            expect_names(nz.params, {"sigma"}, "ExcludedVolume");
            const double rc = f.cutoff_, cc = global_value(f, nz, "cutoff_correction");
            if (std::fabs(cc - std::pow(1.0 / rc, 12)) > 1e-9 * cc) throw OpenMMException("B200 engine: ExcludedVolume cutoff_correction is not (1/rc)^12");
            check(ocm_system_add_excluded_volume(out_, g, global_value(f, nz, "epsilon"), rc, column(f.rows_, 0).data(), &pt.t), "ExcludedVolume");
        } else if (nz.expr == kDebyeHuckel) {
            expect_names(nz.params, {"q"}, "DebyeHuckel");
            check(ocm_system_add_debye_huckel(out_, g, global_value(f, nz, "inv_4_pi_eps0_epsk"), global_value(f, nz, "debye_length"), f.cutoff_,
                                              global_value(f, nz, "cutoff_correction"), column(f.rows_, 0).data(), &pt.t), "DebyeHuckel");
        } else if (nz.expr == kLJRepulsive) {
            expect_names(nz.params, {"epsilon", "sigma"}, "LennardJonesRepulsive");
            check(ocm_system_add_lj_repulsive(out_, g, f.cutoff_, column(f.rows_, 0).data(), column(f.rows_, 1).data(), &pt.t), "LennardJonesRepulsive");
        } else if (nz.expr == kLJAttractive) {
            expect_names(nz.params, {"epsilon", "sigma"}, "LennardJonesAttractive");
            // without OpenMM's switching function the potential is simply cut at rc: r_switch = rc
            const double rsw = f.use_switch_ ? f.switch_ : f.cutoff_;
            check(ocm_system_add_lj_attractive(out_, g, f.cutoff_, rsw, column(f.rows_, 0).data(), column(f.rows_, 1).data(), &pt.t), "LennardJonesAttractive");
        } else if (nz.expr == kWCA) {
            expect_names(nz.params, {"sigma", "eps"}, "WCA");
            check(ocm_system_add_wca(out_, g, f.cutoff_, column(f.rows_, 1).data(), column(f.rows_, 0).data(), &pt.t), "WCA");
        } else if (nz.expr == kISoLFAttractive) {
            expect_names(nz.params, {"sigma", "eps", "omega"}, "iSoLFAttractive");
            check(ocm_system_add_isolf_attractive(out_, g, f.cutoff_, column(f.rows_, 1).data(), column(f.rows_, 0).data(), column(f.rows_, 2).data(), &pt.t),
                  "iSoLFAttractive");
        } else if (nz.expr == kTrigonometric) {
            expect_names(nz.params, {"epsilon", "sigma", "omega"}, "Trigonometric");
            check(ocm_system_add_trigonometric(out_, g, f.cutoff_, column(f.rows_, 0).data(), column(f.rows_, 1).data(), column(f.rows_, 2).data(), &pt.t),
                  "Trigonometric");
        } else if (nz.expr == kUniformWCA || nz.expr == kUniformLJAttractive) {
            // one sigma / epsilon for every member of the interaction groups (one Force per residue-name pair): hand the
            // engine per-bead arrays that carry the parameter on the members and NaN elsewhere
            if (!nz.params.empty()) throw OpenMMException("B200 engine: a uniform pair potential has no per-particle parameters");
            if (f.groups_.empty()) throw OpenMMException("B200 engine: a uniform pair potential needs its interaction groups");
            const bool wca = nz.expr == kUniformWCA;
            const double sigma = global_value(f, nz, "sigma"), eps = global_value(f, nz, wca ? "eps" : "epsilon");
            std::vector<double> sg(f.rows_.size(), std::numeric_limits<double>::quiet_NaN()), ep(sg);
            for (auto& gp : f.groups_)
                for (const std::set<int>* st : {&gp.first, &gp.second})
                    for (int i : *st) { sg.at(i) = sigma; ep.at(i) = eps; }
            if (wca) {
                if (std::fabs(f.cutoff_ - sigma * std::pow(2.0, 1.0 / 6.0)) > 1e-9 * f.cutoff_) throw OpenMMException("B200 engine: UniformWCA cutoff is not 2^(1/6) sigma");
                check(ocm_system_add_wca(out_, g, f.cutoff_, ep.data(), sg.data(), &pt.t), "UniformWeeksChandlerAndersen");
            } else {
                if (!f.use_switch_) throw OpenMMException("B200 engine: UniformLennardJonesAttractive without its switching function");
                check(ocm_system_add_lj_attractive(out_, g, f.cutoff_, f.switch_, ep.data(), sg.data(), &pt.t), "UniformLennardJonesAttractive");
            }
        } else if (nz.expr == k3SPN2ExcludedVolume) {
            expect_names(nz.params, {"epsilon", "sigma"}, "3SPN2ExcludedVolume");
            check(ocm_system_add_3spn2_excluded_volume(out_, g, f.cutoff_, column(f.rows_, 0).data(), column(f.rows_, 1).data(), &pt.t), "3SPN2ExcludedVolume");
        } else {
            unsupported("CustomNonbondedForce", f);
        }
    }

    // rows of `width` doubles starting at parameter column `col`
    static std::vector<double> columns(const std::vector<CustomForceBase::Row>& rows, int col, int width) {
        std::vector<double> out;
        out.reserve(rows.size() * width);
        for (auto& r : rows) for (int k = 0; k < width; k++) out.push_back(r.v[col + k]);
        return out;
    }

    void add_external(const CustomExternalForce& f, int g) {
        const Norm nz = normalise(f);
        const auto idx = indices(f.rows_, 1);
        const std::int64_t n = (std::int64_t)f.rows_.size();
        if (nz.expr == kPositionRestraint) {
            expect_names(nz.params, {"x0", "y0", "z0", "k", "v0"}, "PositionRestraint");
            check(ocm_system_add_position_restraint(out_, g, n, idx.data(), columns(f.rows_, 0, 3).data(), column(f.rows_, 3).data(), column(f.rows_, 4).data()),
                  "PositionRestraint");
        } else if (nz.expr == kPulling) {
            expect_names(nz.params, {"kx", "ky", "kz"}, "Pulling");
            check(ocm_system_add_pulling(out_, g, n, idx.data(), columns(f.rows_, 0, 3).data()), "Pulling");
        } else if (nz.expr == kRectangularBox) {
            expect_names(nz.params, {"radius"}, "EXVRectangularBox");
            const double lo[3] = {global_value(f, nz, "x_lower"), global_value(f, nz, "y_lower"), global_value(f, nz, "z_lower")};
            const double hi[3] = {global_value(f, nz, "x_upper"), global_value(f, nz, "y_upper"), global_value(f, nz, "z_upper")};
            check(ocm_system_add_exv_rectangular_box(out_, g, global_value(f, nz, "epsilon"), lo, hi, n, idx.data(), column(f.rows_, 0).data()), "EXVRectangularBox");
        } else if (nz.expr == kCylindricalRestraint || nz.expr == kCylindricalRestraintPeriodic) {
            expect_names(nz.params, {"xa", "ya", "za", "x0", "y0", "z0", "k", "v0"}, "CylindricalRestraint");
            check(ocm_system_add_cylindrical_restraint(out_, g, nz.expr == kCylindricalRestraintPeriodic, n, idx.data(), columns(f.rows_, 0, 3).data(),
                                                       columns(f.rows_, 3, 3).data(), column(f.rows_, 6).data(), column(f.rows_, 7).data()), "CylindricalRestraint");
        } else {
            unsupported("CustomExternalForce", f);
        }
    }

    void add_centroid(const CustomCentroidBondForce& f, int g) {
        const Norm nz = normalise(f);
        if (nz.expr != kHarmonicCoMPulling || f.ng_ != 2) unsupported("CustomCentroidBondForce", f);
        const double k = global_value(f, nz, "k"), v0 = global_value(f, nz, "v0");
        for (auto& grp : f.groups_)
            if (!grp.second.empty()) throw OpenMMException("B200 engine: CustomCentroidBondForce groups with explicit weights are not supported (the reference uses the masses)");
        for (auto& r : f.rows_) {
            if (r.p.size() != 2 || r.p[0] < 0 || r.p[1] < 0 || r.p[0] >= (int)f.groups_.size() || r.p[1] >= (int)f.groups_.size())
                throw OpenMMException("CustomCentroidBondForce: group index out of range");
            const std::vector<int>& a = f.groups_[r.p[0]].first;
            const std::vector<int>& b = f.groups_[r.p[1]].first;
            std::vector<std::int32_t> g1(a.begin(), a.end()), g2(b.begin(), b.end());
            check(ocm_system_add_harmonic_com_pulling(out_, g, f.periodic_, k, v0, (std::int64_t)g1.size(), g1.data(), (std::int64_t)g2.size(), g2.data()),
                  "HarmonicCoMPulling");
        }
    }

    void add_hbond(const CustomHbondForce& f, int g) {
        const Norm nz = normalise(f, &f.aparams_);
        if (f.method_ == CustomHbondForce::NoCutoff) throw OpenMMException("B200 engine: CustomHbondForce needs a cutoff method (the reference always sets one)");
        const auto donors = indices(f.rows_, 3), acceptors = indices(f.arows_, 3);
        ocm_da_topology t{};
        t.n_donors = (std::int64_t)f.rows_.size(); t.donors = donors.data();
        t.n_acceptors = (std::int64_t)f.arows_.size(); t.acceptors = acceptors.data();
        t.n_excl = (std::int64_t)f.excl_.size() / 2; t.excl_da = f.excl_.data();
        t.cutoff = f.cutoff_;
        t.periodic = f.method_ == CustomHbondForce::CutoffPeriodic;
        if (nz.expr == kBasePair) {
            expect_names(nz.params, {"epsilon", "r0", "t01", "t02", "phi0", "alpha_BP", "K_BP"}, "3SPN2BasePair");
            check(ocm_system_add_3spn2_base_pair(out_, g, &t, column(f.rows_, 0).data(), column(f.rows_, 1).data(), column(f.rows_, 2).data(), column(f.rows_, 3).data(),
                                                 column(f.rows_, 4).data(), uniform(f.rows_, 5, "3SPN2BasePair alpha_BP"), uniform(f.rows_, 6, "3SPN2BasePair K_BP")), "3SPN2BasePair");
        } else if (nz.expr == kCrossStacking) {
            expect_names(nz.aparams, {"epsilon", "r0", "t03", "t0CS", "K_BP", "K_CS", "alpha_CS"}, "3SPN2CrossStacking");
            check(ocm_system_add_3spn2_cross_stacking(out_, g, &t, column(f.arows_, 0).data(), column(f.arows_, 1).data(), column(f.arows_, 2).data(), column(f.arows_, 3).data(),
                                                      uniform(f.arows_, 4, "3SPN2CrossStacking K_BP"), uniform(f.arows_, 5, "3SPN2CrossStacking K_CS"),
                                                      uniform(f.arows_, 6, "3SPN2CrossStacking alpha_CS")), "3SPN2CrossStacking");
        } else if (nz.expr == kPDNS) {
            expect_names(nz.params, {"sigma", "delta", "k", "r0", "theta0", "phi0"}, "PDNS");
            check(ocm_system_add_pdns(out_, g, &t, column(f.rows_, 2).data(), column(f.rows_, 3).data(), column(f.rows_, 4).data(), column(f.rows_, 5).data(),
                                      uniform(f.rows_, 0, "PDNS sigma"), uniform(f.rows_, 1, "PDNS delta")), "PDNS");
        } else if (nz.expr == kCombinatorialGoContact) {
            check(ocm_system_add_combinatorial_go_contact(out_, g, &t, global_value(f, nz, "k"), global_value(f, nz, "r0")), "CombinatorialGoContact");
        } else {
            unsupported("CustomHbondForce", f);
        }
    }
};

}  // namespace b200

// ---- State ------------------------------------------------------------------------------------------------------
const std::vector<Vec3>& State::getPositions() const { if (!(types_ & Positions)) throw OpenMMException("Invoked getPositions() on a State which does not contain positions."); return pos_; }
const std::vector<Vec3>& State::getVelocities() const { if (!(types_ & Velocities)) throw OpenMMException("Invoked getVelocities() on a State which does not contain velocities."); return vel_; }
const std::vector<Vec3>& State::getForces() const { if (!(types_ & Forces)) throw OpenMMException("Invoked getForces() on a State which does not contain forces."); return frc_; }
double State::getKineticEnergy() const { if (!(types_ & Energy)) throw OpenMMException("Invoked getKineticEnergy() on a State which does not contain energies."); return ke_; }
double State::getPotentialEnergy() const { if (!(types_ & Energy)) throw OpenMMException("Invoked getPotentialEnergy() on a State which does not contain energies."); return pe_; }

// ---- Context ------------------------------------------------------------------------------------------------------
Context::Context(const System& s, Integrator& i) : sys_(&s), integ_(&i), platform_(&Platform::getPlatformByName("B200")) { create({}); }
Context::Context(const System& s, Integrator& i, Platform& p) : sys_(&s), integ_(&i), platform_(&p) { create({}); }
Context::Context(const System& s, Integrator& i, Platform& p, const std::map<std::string, std::string>& props) : sys_(&s), integ_(&i), platform_(&p) { create(props); }
Context::Context(Context&& o) noexcept : sys_(o.sys_), integ_(o.integ_), platform_(o.platform_), h_(o.h_) {
    o.h_ = nullptr;
    integ_->ctx_ = this;
}
Context::~Context() {
    if (h_) {
        ocm_context_destroy(h_);
        if (integ_->ctx_ == this) integ_->ctx_ = nullptr;
    }
}
void Context::create(const std::map<std::string, std::string>& props) {
    auto* li = dynamic_cast<LangevinIntegrator*>(integ_);
    if (!li) throw OpenMMException("B200 engine: only LangevinIntegrator is supported (src/IntegratorGenerator.cpp:5-14)");
    if (integ_->ctx_) throw OpenMMException("This Integrator is already bound to a context");
    int device = 0;
    for (const char* key : {"DeviceIndex", "CudaDeviceIndex"}) {
        auto it = props.find(key);
        if (it != props.end()) device = std::atoi(it->second.c_str());
    }
    std::uint64_t seed = (std::uint64_t)(std::int64_t)li->getRandomNumberSeed();
    if (seed == 0) { std::random_device rd; seed = ((std::uint64_t)rd() << 16) ^ rd(); }  // OpenMM: 0 means "pick one"
    ocm_system* sys = b200::SystemCompiler(*sys_).compile();
    const int rc = ocm_context_create(sys, li->getTemperature(), li->getFriction(), li->getStepSize(), seed, device, &h_);
    ocm_system_destroy(sys);
    check(rc, "ocm_context_create");
    // barostats (SystemGenerator.cpp:89-108).  OpenMM draws their random stream from a seed of its own (0 = from the clock); a
    // zero seed here follows the integrator's, so that a run with a fixed integrator seed is reproducible.
    for (int k = 0; k < sys_->getNumForces(); k++) {
        const Force& f = sys_->getForce(k);
        if (auto* b = dynamic_cast<const MonteCarloAnisotropicBarostat*>(&f)) {
            const double P[3] = {b->getDefaultPressure()[0], b->getDefaultPressure()[1], b->getDefaultPressure()[2]};
            const std::int32_t mode[3] = {b->getScaleX() ? 1 : 0, b->getScaleY() ? 1 : 0, b->getScaleZ() ? 1 : 0};
            const std::uint64_t bs = b->getRandomNumberSeed() ? (std::uint64_t)(std::int64_t)b->getRandomNumberSeed() : seed;
            check(ocm_context_set_barostat(h_, OCM_BAROSTAT_ANISOTROPIC, P, 0.0, b->getDefaultTemperature(), mode, b->getFrequency(), bs),
                  "MonteCarloAnisotropicBarostat");
        } else if (auto* b = dynamic_cast<const MonteCarloMembraneBarostat*>(&f)) {
            const double P[3] = {b->getDefaultPressure(), 0.0, 0.0};
            const std::int32_t mode[3] = {(std::int32_t)b->getXYMode(), (std::int32_t)b->getZMode(), 0};
            const std::uint64_t bs = b->getRandomNumberSeed() ? (std::uint64_t)(std::int64_t)b->getRandomNumberSeed() : seed;
            check(ocm_context_set_barostat(h_, OCM_BAROSTAT_MEMBRANE, P, b->getDefaultSurfaceTension(), b->getDefaultTemperature(), mode,
                                           b->getFrequency(), bs), "MonteCarloMembraneBarostat");
        }
    }
    integ_->ctx_ = this;
}
void Context::setPositions(const std::vector<Vec3>& p) {
    if ((int)p.size() != sys_->getNumParticles()) throw OpenMMException("Called setPositions() on a Context with the wrong number of positions");
    static_assert(sizeof(Vec3) == 3 * sizeof(double), "Vec3 must be three packed doubles");
    check(ocm_set_positions(h_, reinterpret_cast<const double*>(p.data())), "setPositions");
    invalidate();
}
void Context::setVelocities(const std::vector<Vec3>& v) {
    if ((int)v.size() != sys_->getNumParticles()) throw OpenMMException("Called setVelocities() on a Context with the wrong number of velocities");
    check(ocm_set_velocities(h_, reinterpret_cast<const double*>(v.data())), "setVelocities");
    invalidate();
}
void Context::setVelocitiesToTemperature(double T) {
    std::random_device rd;
    setVelocitiesToTemperature(T, (int)(rd() & 0x7fffffff));
}
void Context::setVelocitiesToTemperature(double T, int seed) {
    check(ocm_set_velocities_to_temperature(h_, T, (std::uint64_t)(std::int64_t)seed), "setVelocitiesToTemperature");
    invalidate();
}
double Context::getTime() const { return (double)ocm_current_step(h_) * integ_->getStepSize(); }
std::int64_t Context::getStepCount() const { return ocm_current_step(h_); }

State Context::getState(int types, bool enforcePeriodicBox, int groups) const {
    State s;
    s.types_ = types;
    s.time_ = getTime();
    sys_->getDefaultPeriodicBoxVectors(s.box_[0], s.box_[1], s.box_[2]);
    if (sys_->usesPeriodicBoundaryConditions()) {
        // (a Monte Carlo barostat changes the box of the Context, not the System's default)
        double b[3] = {0, 0, 0};
        check(ocm_get_box(h_, b), "getState");
        if (b[0] > 0) { s.box_[0] = Vec3(b[0], 0, 0); s.box_[1] = Vec3(0, b[1], 0); s.box_[2] = Vec3(0, 0, b[2]); }
    }
    const int n = sys_->getNumParticles();
    if (types & State::Forces) s.frc_.resize(n);
    if ((types & State::Forces) && (unsigned)groups != 0xFFFFFFFFu) throw OpenMMException("B200 engine: forces restricted to a force-group mask are not supported");
    if (types & State::Forces) {
        // forces are not part of the frame copy: blocking call
        double epot = 0;
        const bool want_energy = (types & State::Energy) && !energy_valid_;
        check(ocm_get_state(h_, 0xFFFFFFFFu, nullptr, nullptr, reinterpret_cast<double*>(s.frc_.data()), want_energy ? &epot : nullptr,
                            want_energy ? &ekin_ : nullptr, want_energy ? ebg_ : nullptr), "getState");
        if (want_energy) energy_valid_ = true;
    }
    if ((types & (State::Positions | State::Velocities | State::Energy)) && !frame_valid_) {
        frame_pos_.resize(n); frame_vel_.resize(n);
        double epot = 0;
        ocm_state_request* rq = nullptr;
        check(ocm_get_state_begin(h_, 0xFFFFFFFFu, 1, 1, 0, 1, &rq), "getState");
        check(ocm_get_state_end(rq, reinterpret_cast<double*>(frame_pos_.data()), reinterpret_cast<double*>(frame_vel_.data()), nullptr, &epot, &ekin_, ebg_),
              "getState");
        frame_valid_ = true; energy_valid_ = true;
    }
    if (types & State::Positions) s.pos_ = frame_pos_;
    if (types & State::Velocities) s.vel_ = frame_vel_;
    if (types & State::Energy) {
        double e = 0;
        for (int g = 0; g < 32; g++) if ((unsigned)groups >> g & 1u) e += ebg_[g];
        s.pe_ = e;
        s.ke_ = ekin_;
    }
    if ((types & State::Positions) && enforcePeriodicBox && sys_->usesPeriodicBoundaryConditions())
        for (auto& p : s.pos_) for (int c = 0; c < 3; c++) p[c] -= s.box_[c][c] * std::floor(p[c] / s.box_[c][c]);
    return s;
}

void LangevinIntegrator::step(int steps) {
    if (!ctx_) throw OpenMMException("This Integrator is not bound to a context!");
    check(ocm_step(ctx_->h_, steps), "LangevinIntegrator::step");
    ctx_->invalidate();
}

void LocalEnergyMinimizer::minimize(Context& context, double tolerance, int maxIterations) {
    check(ocm_minimize(context.h_, tolerance, maxIterations, nullptr, nullptr), "LocalEnergyMinimizer::minimize");
    context.invalidate();
}

}  // namespace OpenMM
