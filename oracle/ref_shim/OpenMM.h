// OpenMM.h — "openmm-lite": a small double-precision stand-in for the slice of the OpenMM 8.0 C++ API that the
// reference (yutakasi634/OpenCafeMol) calls.
//
// TEST INFRASTRUCTURE (oracle/_ref build only; never linked into the product).  OpenMM is an external
// library of the reference (`.github/workflows/main.yml:25` pins 8.0.0) that is neither vendored under
// /root/reference nor installable offline.  To run the reference's OWN reader, Topology, force-field
// generators, SystemGenerator, Simulator and Observers unmodified, this header declares the classes they use
// (SURVEY.md 8b lists the surface) and oracle/ref_shim/openmm_lite.cpp implements them the way OpenMM's
// Reference platform behaves: the Custom*Force classes keep the reference's energy
// expression STRING and evaluate it with a small Lepton-like interpreter with forward-mode automatic
// differentiation, so energies and forces come from the reference's own formulas and parameters, not from
// the formulas restated in oracle/ocm_oracle.cpp.  Nothing of OpenMM is copied; semantics are restated.
#ifndef OCM_REF_SHIM_OPENMM_H
#define OCM_REF_SHIM_OPENMM_H

#include <cassert>
#include <cmath>
#include <cstdint>
#include <exception>
#include <iosfwd>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <utility>
#include <vector>

namespace OpenMM {

static const double NmPerAngstrom = 0.1;
static const double AngstromsPerNm = 10.0;
static const double KJPerKcal = 4.184;
static const double KcalPerKJ = 1.0 / 4.184;
static const double RadiansPerDegree = 3.14159265358979323846 / 180.0;
static const double DegreesPerRadian = 180.0 / 3.14159265358979323846;
static const double SigmaPerVdwRadius = 1.7817974362806786095;

class OpenMMException : public std::exception {
  public:
    explicit OpenMMException(std::string m) : msg_(std::move(m)) {}
    const char* what() const noexcept override { return msg_.c_str(); }

  private:
    std::string msg_;
};

class Vec3 {
  public:
    Vec3() : d_{0, 0, 0} {}
    Vec3(double x, double y, double z) : d_{x, y, z} {}
    double operator[](int i) const { return d_[i]; }
    double& operator[](int i) { return d_[i]; }
    Vec3 operator+(const Vec3& o) const { return {d_[0] + o[0], d_[1] + o[1], d_[2] + o[2]}; }
    Vec3 operator-(const Vec3& o) const { return {d_[0] - o[0], d_[1] - o[1], d_[2] - o[2]}; }
    Vec3 operator-() const { return {-d_[0], -d_[1], -d_[2]}; }
    Vec3 operator*(double s) const { return {d_[0] * s, d_[1] * s, d_[2] * s}; }
    Vec3 operator/(double s) const { return {d_[0] / s, d_[1] / s, d_[2] / s}; }
    Vec3& operator+=(const Vec3& o) { d_[0] += o[0]; d_[1] += o[1]; d_[2] += o[2]; return *this; }
    Vec3& operator-=(const Vec3& o) { d_[0] -= o[0]; d_[1] -= o[1]; d_[2] -= o[2]; return *this; }
    Vec3& operator*=(double s) { d_[0] *= s; d_[1] *= s; d_[2] *= s; return *this; }
    bool operator==(const Vec3& o) const { return d_[0] == o[0] && d_[1] == o[1] && d_[2] == o[2]; }
    bool operator!=(const Vec3& o) const { return !(*this == o); }
    double dot(const Vec3& o) const { return d_[0] * o[0] + d_[1] * o[1] + d_[2] * o[2]; }
    Vec3 cross(const Vec3& o) const {
        return {d_[1] * o[2] - d_[2] * o[1], d_[2] * o[0] - d_[0] * o[2], d_[0] * o[1] - d_[1] * o[0]};
    }

  private:
    double d_[3];
};
inline Vec3 operator*(double s, const Vec3& v) { return v * s; }

namespace lite {
struct Program;      // compiled energy expression
struct EvalContext;  // positions + box handed to Force::evaluate
struct Describer;    // JSON sink for Force::describe
}  // namespace lite

// ---- tabulated functions ---------------------------------------------------------------------------------
class TabulatedFunction {
  public:
    virtual ~TabulatedFunction() = default;
    virtual double value(double x) const = 0;
    virtual double derivative(double x) const = 0;
};

// natural cubic spline through equally spaced values on [min, max]; 0 outside the interval
class Continuous1DFunction : public TabulatedFunction {
  public:
    Continuous1DFunction(const std::vector<double>& values, double min, double max, bool periodic = false);
    void getFunctionParameters(std::vector<double>& values, double& min, double& max) const;
    double value(double x) const override;
    double derivative(double x) const override;

  private:
    std::vector<double> y_, y2_;
    double min_, max_;
};

// ---- forces ------------------------------------------------------------------------------------------------
class Force {
  public:
    Force() = default;
    Force(const Force&) = delete;
    Force& operator=(const Force&) = delete;
    virtual ~Force() = default;
    int getForceGroup() const { return group_; }
    void setForceGroup(int g);
    virtual bool usesPeriodicBoundaryConditions() const { return false; }
    // openmm-lite: potential energy at the given state; forces are ADDED to *forces when it is not null
    virtual double evaluate(const lite::EvalContext& ec, std::vector<Vec3>* forces) const = 0;
    // openmm-lite: everything the reference handed to this Force, as one JSON object
    virtual void describe(lite::Describer& out) const = 0;
    // openmm-lite: ForceImpl::getBondedParticles (the edges ContextImpl::getMolecules joins into molecules)
    virtual void bondedParticles(std::vector<std::pair<int, int>>&) const {}

  private:
    int group_ = 0;
};

class HarmonicBondForce : public Force {
  public:
    int addBond(int p1, int p2, double length, double k);
    int getNumBonds() const { return (int)bonds_.size(); }
    void setUsesPeriodicBoundaryConditions(bool p) { periodic_ = p; }
    bool usesPeriodicBoundaryConditions() const override { return periodic_; }
    double evaluate(const lite::EvalContext&, std::vector<Vec3>*) const override;
    void describe(lite::Describer&) const override;
    void bondedParticles(std::vector<std::pair<int, int>>& out) const override { for (auto& b : bonds_) out.emplace_back(b.p1, b.p2); }

  private:
    struct Bond { int p1, p2; double length, k; };
    std::vector<Bond> bonds_;
    bool periodic_ = false;
};

class HarmonicAngleForce : public Force {
  public:
    int addAngle(int p1, int p2, int p3, double angle, double k);
    int getNumAngles() const { return (int)angles_.size(); }
    void setUsesPeriodicBoundaryConditions(bool p) { periodic_ = p; }
    bool usesPeriodicBoundaryConditions() const override { return periodic_; }
    double evaluate(const lite::EvalContext&, std::vector<Vec3>*) const override;
    void describe(lite::Describer&) const override;
    void bondedParticles(std::vector<std::pair<int, int>>& out) const override {
        for (auto& a : angles_) { out.emplace_back(a.p1, a.p2); out.emplace_back(a.p2, a.p3); }
    }

  private:
    struct Angle { int p1, p2, p3; double angle, k; };
    std::vector<Angle> angles_;
    bool periodic_ = false;
};

// shared bookkeeping of the Custom*Force classes
class CustomForceBase : public Force {
  public:
    explicit CustomForceBase(std::string expr) : expr_(std::move(expr)) {}
    ~CustomForceBase() override;
    const std::string& getEnergyFunction() const { return expr_; }
    int addGlobalParameter(const std::string& name, double value);
    int addTabulatedFunction(const std::string& name, TabulatedFunction* fn);  // takes ownership

  protected:
    std::string expr_;
    std::vector<std::pair<std::string, double>> globals_;
    std::vector<std::pair<std::string, TabulatedFunction*>> functions_;
    mutable std::shared_ptr<lite::Program> prog_;
    void describe_common(lite::Describer&, const char* cls) const;
};

class CustomBondForce : public CustomForceBase {
  public:
    explicit CustomBondForce(const std::string& energy) : CustomForceBase(energy) {}
    int addPerBondParameter(const std::string& name);
    int addBond(int p1, int p2, const std::vector<double>& params = {});
    int getNumBonds() const { return (int)bonds_.size(); }
    void setUsesPeriodicBoundaryConditions(bool p) { periodic_ = p; }
    bool usesPeriodicBoundaryConditions() const override { return periodic_; }
    double evaluate(const lite::EvalContext&, std::vector<Vec3>*) const override;
    void describe(lite::Describer&) const override;
    void bondedParticles(std::vector<std::pair<int, int>>& out) const override { for (auto& b : bonds_) out.emplace_back(b.p1, b.p2); }

  private:
    struct Bond { int p1, p2; std::vector<double> params; };
    std::vector<std::string> params_;
    std::vector<Bond> bonds_;
    bool periodic_ = false;
};

class CustomTorsionForce : public CustomForceBase {
  public:
    explicit CustomTorsionForce(const std::string& energy) : CustomForceBase(energy) {}
    int addPerTorsionParameter(const std::string& name);
    int addTorsion(int p1, int p2, int p3, int p4, const std::vector<double>& params = {});
    int getNumTorsions() const { return (int)torsions_.size(); }
    void setUsesPeriodicBoundaryConditions(bool p) { periodic_ = p; }
    bool usesPeriodicBoundaryConditions() const override { return periodic_; }
    double evaluate(const lite::EvalContext&, std::vector<Vec3>*) const override;
    void describe(lite::Describer&) const override;
    void bondedParticles(std::vector<std::pair<int, int>>& out) const override {
        for (auto& t : torsions_) for (int k = 1; k < 4; k++) out.emplace_back(t.p[k - 1], t.p[k]);
    }

  private:
    struct Torsion { int p[4]; std::vector<double> params; };
    std::vector<std::string> params_;
    std::vector<Torsion> torsions_;
    bool periodic_ = false;
};

class CustomCompoundBondForce : public CustomForceBase {
  public:
    CustomCompoundBondForce(int numParticles, const std::string& energy) : CustomForceBase(energy), np_(numParticles) {}
    int addPerBondParameter(const std::string& name);
    int addBond(const std::vector<int>& particles, const std::vector<double>& params = {});
    int getNumBonds() const { return (int)bonds_.size(); }
    void setUsesPeriodicBoundaryConditions(bool p) { periodic_ = p; }
    bool usesPeriodicBoundaryConditions() const override { return periodic_; }
    double evaluate(const lite::EvalContext&, std::vector<Vec3>*) const override;
    void describe(lite::Describer&) const override;
    void bondedParticles(std::vector<std::pair<int, int>>& out) const override {
        for (auto& b : bonds_) for (std::size_t k = 1; k < b.p.size(); k++) out.emplace_back(b.p[k - 1], b.p[k]);
    }

  private:
    struct Bond { std::vector<int> p; std::vector<double> params; };
    int np_;
    std::vector<std::string> params_;
    std::vector<Bond> bonds_;
    bool periodic_ = false;
};

class CustomCentroidBondForce : public CustomForceBase {
  public:
    CustomCentroidBondForce(int numGroups, const std::string& energy) : CustomForceBase(energy), ng_(numGroups) {}
    int addPerBondParameter(const std::string& name);
    int addGroup(const std::vector<int>& particles, const std::vector<double>& weights = {});
    int addBond(const std::vector<int>& groups, const std::vector<double>& params = {});
    void setUsesPeriodicBoundaryConditions(bool p) { periodic_ = p; }
    bool usesPeriodicBoundaryConditions() const override { return periodic_; }
    double evaluate(const lite::EvalContext&, std::vector<Vec3>*) const override;
    void describe(lite::Describer&) const override;

  private:
    struct Group { std::vector<int> p; std::vector<double> w; };
    struct Bond { std::vector<int> g; std::vector<double> params; };
    int ng_;
    std::vector<std::string> params_;
    std::vector<Group> groups_;
    std::vector<Bond> bonds_;
    bool periodic_ = false;
};

class CustomExternalForce : public CustomForceBase {
  public:
    explicit CustomExternalForce(const std::string& energy) : CustomForceBase(energy) {}
    int addPerParticleParameter(const std::string& name);
    int addParticle(int particle, const std::vector<double>& params = {});
    double evaluate(const lite::EvalContext&, std::vector<Vec3>*) const override;
    void describe(lite::Describer&) const override;

  private:
    struct Item { int p; std::vector<double> params; };
    std::vector<std::string> params_;
    std::vector<Item> items_;
};

class CustomNonbondedForce : public CustomForceBase {
  public:
    enum NonbondedMethod { NoCutoff = 0, CutoffNonPeriodic = 1, CutoffPeriodic = 2 };
    explicit CustomNonbondedForce(const std::string& energy) : CustomForceBase(energy) {}
    int addPerParticleParameter(const std::string& name);
    int addParticle(const std::vector<double>& params = {});
    int getNumParticles() const { return (int)particles_.size(); }
    int addExclusion(int p1, int p2);
    int getNumExclusions() const { return (int)exclusions_.size(); }
    int addInteractionGroup(const std::set<int>& set1, const std::set<int>& set2);
    int getNumInteractionGroups() const { return (int)groups_.size(); }
    NonbondedMethod getNonbondedMethod() const { return method_; }
    void setNonbondedMethod(NonbondedMethod m) { method_ = m; }
    double getCutoffDistance() const { return cutoff_; }
    void setCutoffDistance(double d) { cutoff_ = d; }
    bool getUseSwitchingFunction() const { return use_switch_; }
    void setUseSwitchingFunction(bool u) { use_switch_ = u; }
    double getSwitchingDistance() const { return switch_; }
    void setSwitchingDistance(double d) { switch_ = d; }
    bool usesPeriodicBoundaryConditions() const override { return method_ == CutoffPeriodic; }
    double evaluate(const lite::EvalContext&, std::vector<Vec3>*) const override;
    void describe(lite::Describer&) const override;

  private:
    std::vector<std::string> params_;
    std::vector<std::vector<double>> particles_;
    std::vector<std::pair<int, int>> exclusions_;
    std::vector<std::pair<std::set<int>, std::set<int>>> groups_;
    NonbondedMethod method_ = NoCutoff;
    double cutoff_ = 1.0, switch_ = -1.0;
    bool use_switch_ = false;
};

class CustomHbondForce : public CustomForceBase {
  public:
    enum NonbondedMethod { NoCutoff = 0, CutoffNonPeriodic = 1, CutoffPeriodic = 2 };
    explicit CustomHbondForce(const std::string& energy) : CustomForceBase(energy) {}
    int addPerDonorParameter(const std::string& name);
    int addPerAcceptorParameter(const std::string& name);
    int addDonor(int d1, int d2, int d3, const std::vector<double>& params = {});
    int addAcceptor(int a1, int a2, int a3, const std::vector<double>& params = {});
    int addExclusion(int donor, int acceptor);
    int getNumDonors() const { return (int)donors_.size(); }
    int getNumAcceptors() const { return (int)acceptors_.size(); }
    NonbondedMethod getNonbondedMethod() const { return method_; }
    void setNonbondedMethod(NonbondedMethod m) { method_ = m; }
    double getCutoffDistance() const { return cutoff_; }
    void setCutoffDistance(double d) { cutoff_ = d; }
    bool usesPeriodicBoundaryConditions() const override { return method_ == CutoffPeriodic; }
    double evaluate(const lite::EvalContext&, std::vector<Vec3>*) const override;
    void describe(lite::Describer&) const override;

  private:
    struct Site { int p[3]; std::vector<double> params; };
    std::vector<std::string> dparams_, aparams_;
    std::vector<Site> donors_, acceptors_;
    std::vector<std::pair<int, int>> exclusions_;
    NonbondedMethod method_ = NoCutoff;
    double cutoff_ = 1.0;
};

// barostats: the trial moves run inside LangevinIntegrator::step (Context::updateContextState; OpenMM 8.0 algorithm restated,
// random stream = SplitMix64 of the integrator seed as in the engine and the oracle)
class MonteCarloAnisotropicBarostat : public Force {
  public:
    MonteCarloAnisotropicBarostat(const Vec3& pressure, double temperature, bool scaleX = true, bool scaleY = true,
                                  bool scaleZ = true, int frequency = 25)
        : p_(pressure), t_(temperature), sx_(scaleX), sy_(scaleY), sz_(scaleZ), freq_(frequency) {}
    double evaluate(const lite::EvalContext&, std::vector<Vec3>*) const override { return 0.0; }
    void describe(lite::Describer&) const override;

  private:
    friend class Context;
    Vec3 p_;
    double t_;
    bool sx_, sy_, sz_;
    int freq_;
};

class MonteCarloMembraneBarostat : public Force {
  public:
    enum XYMode { XYIsotropic = 0, XYAnisotropic = 1 };
    enum ZMode { ZFree = 0, ZFixed = 1, ConstantVolume = 2 };
    MonteCarloMembraneBarostat(double pressure, double surfaceTension, double temperature, XYMode xymode, ZMode zmode,
                               int frequency = 25)
        : p_(pressure), g_(surfaceTension), t_(temperature), xy_(xymode), z_(zmode), freq_(frequency) {}
    double evaluate(const lite::EvalContext&, std::vector<Vec3>*) const override { return 0.0; }
    void describe(lite::Describer&) const override;

  private:
    friend class Context;
    double p_, g_, t_;
    XYMode xy_;
    ZMode z_;
    int freq_;
};

// ---- system ------------------------------------------------------------------------------------------------
class System {
  public:
    System() = default;
    System(const System&) = delete;
    ~System();
    int addParticle(double mass) { mass_.push_back(mass); return (int)mass_.size() - 1; }
    int getNumParticles() const { return (int)mass_.size(); }
    double getParticleMass(int i) const { return mass_.at(i); }
    void setDefaultPeriodicBoxVectors(const Vec3& a, const Vec3& b, const Vec3& c) { box_[0] = a; box_[1] = b; box_[2] = c; has_box_ = true; }
    void getDefaultPeriodicBoxVectors(Vec3& a, Vec3& b, Vec3& c) const { a = box_[0]; b = box_[1]; c = box_[2]; }
    int addForce(Force* f) { forces_.push_back(f); return (int)forces_.size() - 1; }  // takes ownership
    int getNumForces() const { return (int)forces_.size(); }
    const Force& getForce(int i) const { return *forces_.at(i); }
    Force& getForce(int i) { return *forces_.at(i); }
    bool usesPeriodicBoundaryConditions() const;
    bool hasBox() const { return has_box_; }  // openmm-lite

  private:
    std::vector<double> mass_;
    std::vector<Force*> forces_;
    Vec3 box_[3] = {Vec3(2, 0, 0), Vec3(0, 2, 0), Vec3(0, 0, 2)};
    bool has_box_ = false;
};

// ---- platform ----------------------------------------------------------------------------------------------
class Platform {
  public:
    const std::string& getName() const { return name_; }
    static const std::string& getOpenMMVersion();
    static std::vector<std::string> loadPluginsFromDirectory(const std::string& dir);
    static std::vector<std::string> getPluginLoadFailures();
    static int getNumPlatforms();
    static Platform& getPlatform(int i);
    static Platform& getPlatformByName(const std::string& name);

  private:
    explicit Platform(std::string n) : name_(std::move(n)) {}
    std::string name_;
    friend struct PlatformRegistry;
};

// ---- integrators -------------------------------------------------------------------------------------------
class Context;
class Integrator {
  public:
    virtual ~Integrator() = default;
    virtual void step(int steps) = 0;
    double getStepSize() const { return dt_; }
    void setStepSize(double dt) { dt_ = dt; }

  protected:
    friend class Context;
    Context* ctx_ = nullptr;
    double dt_ = 0.0;
};

// leap-frog Langevin.  Noise: the repository's documented Philox4x32-10 stream
// (key = seed, counter = (bead, step, stream 0)), so runs are comparable with oracle/ocm_oracle.cpp step by step.
class LangevinIntegrator : public Integrator {
  public:
    LangevinIntegrator(double temperature, double frictionCoeff, double stepSize);
    double getTemperature() const { return T_; }
    double getFriction() const { return gamma_; }
    int getRandomNumberSeed() const { return seed_; }
    void setRandomNumberSeed(int seed) { seed_ = seed; }
    void step(int steps) override;

  private:
    double T_, gamma_;
    int seed_ = 0;
};

// ---- state / context ---------------------------------------------------------------------------------------
class State {
  public:
    enum DataType { Positions = 1, Velocities = 2, Forces = 4, Energy = 8, Parameters = 16, ParameterDerivatives = 32, IntegratorParameters = 64 };
    double getTime() const { return time_; }
    const std::vector<Vec3>& getPositions() const;
    const std::vector<Vec3>& getVelocities() const;
    const std::vector<Vec3>& getForces() const;
    double getKineticEnergy() const;
    double getPotentialEnergy() const;
    void getPeriodicBoxVectors(Vec3& a, Vec3& b, Vec3& c) const { a = box_[0]; b = box_[1]; c = box_[2]; }
    double getPeriodicBoxVolume() const { return box_[0][0] * box_[1][1] * box_[2][2]; }

  private:
    friend class Context;
    int types_ = 0;
    double time_ = 0, ke_ = 0, pe_ = 0;
    std::vector<Vec3> pos_, vel_, frc_;
    Vec3 box_[3];
};

class Context {
  public:
    Context(const System& system, Integrator& integrator);
    Context(const System& system, Integrator& integrator, Platform& platform);
    Context(const System& system, Integrator& integrator, Platform& platform, const std::map<std::string, std::string>& properties);
    Context(const Context&) = delete;
    Context(Context&& other);  // the reference returns its Simulator (which holds a Context by value) from functions
    ~Context();
    const System& getSystem() const { return sys_; }
    Integrator& getIntegrator() { return integ_; }
    Platform& getPlatform() { return *platform_; }
    State getState(int types, bool enforcePeriodicBox = false, int groups = 0xFFFFFFFF) const;
    void setPositions(const std::vector<Vec3>& positions);
    void setVelocities(const std::vector<Vec3>& velocities);
    void setVelocitiesToTemperature(double temperature);  // random seed
    void setVelocitiesToTemperature(double temperature, int randomSeed);
    void setPeriodicBoxVectors(const Vec3& a, const Vec3& b, const Vec3& c) { box_[0] = a; box_[1] = b; box_[2] = c; }
    double getTime() const { return time_; }
    std::int64_t getStepCount() const { return steps_; }

    // openmm-lite: energy (sum over the forces whose group bit is set) and optionally forces
    double computeForcesAndEnergy(int groups, std::vector<Vec3>* forces) const;
    // openmm-lite: the Context most recently constructed and still alive in this process (the reference's
    // Simulator keeps its Context private; oracle/ref_shim/ref_api.cpp reaches it through this)
    static Context* mostRecent();

  private:
    friend class LangevinIntegrator;
    friend class LocalEnergyMinimizer;
    const System& sys_;
    Integrator& integ_;
    Platform* platform_;
    std::vector<Vec3> pos_, vel_;
    Vec3 box_[3];
    double time_ = 0;
    std::int64_t steps_ = 0;
    bool moved_from_ = false;
    void init();
    // Monte Carlo barostat state (one barostat per System)
    struct Baro {
        int kind = 0, mode[3] = {0, 0, 0}, freq = 0, calls = 0, attempted[3] = {0, 0, 0}, accepted[3] = {0, 0, 0};
        double pressure[3] = {0, 0, 0}, tension = 0, T = 0, vscale[3] = {0, 0, 0};
        std::uint64_t rng = 0;
        bool seeded = false;
        std::vector<std::vector<int>> molecules;
    } baro_;
    void updateContextState(std::uint64_t integrator_seed);
    void barostatTrial();
};

// L-BFGS on the potential energy until the RMS force drops below `tolerance` (kJ/mol/nm)
class LocalEnergyMinimizer {
  public:
    static void minimize(Context& context, double tolerance = 10.0, int maxIterations = 0);
};

}  // namespace OpenMM

#endif  // OCM_REF_SHIM_OPENMM_H
