// OpenMM.h — the C++ host layer of ocm-b200: the OpenMM C++ API surface the reference is written against,
// implemented on the B200 engine's C ABI (include/ocm_b200.h).
//
// The reference (yutakasi634/OpenCafeMol) reaches its per-step hot path only through OpenMM classes: every
// src/forcefield/*Generator::generate() returns an OpenMM::Force, SystemGenerator::generate() an OpenMM::System
// (src/SystemGenerator.cpp:34-111), Simulator owns an OpenMM::Context and calls integrator->step(save_step)
// (src/Simulator.cpp:14-17,60), the Observers call context.getState (src/Observer.cpp:34,107,113,340,344).
// Compiling the reference with this directory on its include path instead of OpenMM's, and linking
// libocm_openmm_compat.so instead of libOpenMM, runs its unmodified reader / generators / Simulator / Observers on
// the CUDA engine: same TOML and GENESIS inputs, same .ene / .dcd / .pdb outputs.  INTEGRATION.md has the recipe.
//
// Scope: exactly the classes and methods the reference uses.  Force objects only record what they
// are given; when a Context is created every Force is translated into one ocm_system_add_* call.  A Custom*Force is
// recognised by its energy expression (the catalogue of expressions the reference's generators emit, compared after
// removing white space and the per-generator id prefix); an expression outside the catalogue, a barostat or an
// external force makes the Context constructor throw OpenMMException — there is no interpreter and no CPU fallback.
#ifndef OCM_B200_OPENMM_COMPAT_H
#define OCM_B200_OPENMM_COMPAT_H

#include <cassert>
#include <cmath>
#include <cstdint>
#include <exception>
#include <iosfwd>
#include <iostream>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

struct ocm_context;
struct ocm_system;

namespace OpenMM {

static const double NmPerAngstrom = 0.1;
static const double AngstromsPerNm = 10.0;
static const double KJPerKcal = 4.184;
static const double KcalPerKJ = 1.0 / 4.184;
static const double RadiansPerDegree = 3.14159265358979323846 / 180.0;
static const double DegreesPerRadian = 180.0 / 3.14159265358979323846;

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
    Vec3 operator+(const Vec3& o) const { return Vec3(d_[0] + o[0], d_[1] + o[1], d_[2] + o[2]); }
    Vec3 operator-(const Vec3& o) const { return Vec3(d_[0] - o[0], d_[1] - o[1], d_[2] - o[2]); }
    Vec3 operator-() const { return Vec3(-d_[0], -d_[1], -d_[2]); }
    Vec3 operator*(double s) const { return Vec3(d_[0] * s, d_[1] * s, d_[2] * s); }
    Vec3 operator/(double s) const { return Vec3(d_[0] / s, d_[1] / s, d_[2] / s); }
    Vec3& operator+=(const Vec3& o) { for (int k = 0; k < 3; k++) d_[k] += o[k]; return *this; }
    Vec3& operator-=(const Vec3& o) { for (int k = 0; k < 3; k++) d_[k] -= o[k]; return *this; }
    Vec3& operator*=(double s) { for (int k = 0; k < 3; k++) d_[k] *= s; return *this; }
    bool operator==(const Vec3& o) const { return d_[0] == o[0] && d_[1] == o[1] && d_[2] == o[2]; }
    bool operator!=(const Vec3& o) const { return !(*this == o); }
    double dot(const Vec3& o) const { return d_[0] * o[0] + d_[1] * o[1] + d_[2] * o[2]; }
    Vec3 cross(const Vec3& o) const {
        return Vec3(d_[1] * o[2] - d_[2] * o[1], d_[2] * o[0] - d_[0] * o[2], d_[0] * o[1] - d_[1] * o[0]);
    }

  private:
    double d_[3];
};
inline Vec3 operator*(double s, const Vec3& v) { return v * s; }

namespace b200 {
class SystemCompiler;  // Force -> ocm_system_add_* translation (openmm_compat.cpp)
}

// ---- tabulated functions -----------------------------------------------------------------------------------
class TabulatedFunction {
  public:
    virtual ~TabulatedFunction() = default;
};
// the engine evaluates the natural cubic spline itself (ocm_system_add_flexible_local_angle); this only carries the knots
class Continuous1DFunction : public TabulatedFunction {
  public:
    Continuous1DFunction(const std::vector<double>& values, double min, double max, bool periodic = false);
    void getFunctionParameters(std::vector<double>& values, double& min, double& max) const { values = y_; min = min_; max = max_; }

  private:
    std::vector<double> y_;
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

  private:
    int group_ = 0;
};

class HarmonicBondForce : public Force {
  public:
    int addBond(int p1, int p2, double length, double k);
    int getNumBonds() const { return (int)p_.size() / 2; }
    void setUsesPeriodicBoundaryConditions(bool p) { periodic_ = p; }
    bool usesPeriodicBoundaryConditions() const override { return periodic_; }

  private:
    friend class b200::SystemCompiler;
    std::vector<std::int32_t> p_;
    std::vector<double> length_, k_;
    bool periodic_ = false;
};

class HarmonicAngleForce : public Force {
  public:
    int addAngle(int p1, int p2, int p3, double angle, double k);
    int getNumAngles() const { return (int)p_.size() / 3; }
    void setUsesPeriodicBoundaryConditions(bool p) { periodic_ = p; }
    bool usesPeriodicBoundaryConditions() const override { return periodic_; }

  private:
    friend class b200::SystemCompiler;
    std::vector<std::int32_t> p_;
    std::vector<double> angle_, k_;
    bool periodic_ = false;
};

// what all Custom*Force classes share: the expression, global parameters, tabulated functions, and rows of
// (particle indices, per-item parameters)
class CustomForceBase : public Force {
  public:
    ~CustomForceBase() override;
    const std::string& getEnergyFunction() const { return expr_; }
    int addGlobalParameter(const std::string& name, double value);
    int addTabulatedFunction(const std::string& name, TabulatedFunction* fn);  // takes ownership

  protected:
    explicit CustomForceBase(std::string expr) : expr_(std::move(expr)) {}
    friend class b200::SystemCompiler;
    struct Row {
        std::vector<int> p;
        std::vector<double> v;
    };
    std::string expr_;
    std::vector<std::pair<std::string, double>> globals_;
    std::vector<std::pair<std::string, TabulatedFunction*>> functions_;
    std::vector<std::string> params_;  // per-bond / per-torsion / per-particle / per-donor names
    std::vector<Row> rows_;
    bool periodic_ = false;
    int add_param(const std::string& n) { params_.push_back(n); return (int)params_.size() - 1; }
    int add_row(std::vector<int> p, const std::vector<double>& v, const char* who);
};

class CustomBondForce : public CustomForceBase {
  public:
    explicit CustomBondForce(const std::string& energy) : CustomForceBase(energy) {}
    int addPerBondParameter(const std::string& name) { return add_param(name); }
    int addBond(int p1, int p2, const std::vector<double>& params = {}) { return add_row({p1, p2}, params, "CustomBondForce"); }
    int getNumBonds() const { return (int)rows_.size(); }
    void setUsesPeriodicBoundaryConditions(bool p) { periodic_ = p; }
    bool usesPeriodicBoundaryConditions() const override { return periodic_; }
};

class CustomTorsionForce : public CustomForceBase {
  public:
    explicit CustomTorsionForce(const std::string& energy) : CustomForceBase(energy) {}
    int addPerTorsionParameter(const std::string& name) { return add_param(name); }
    int addTorsion(int p1, int p2, int p3, int p4, const std::vector<double>& params = {}) { return add_row({p1, p2, p3, p4}, params, "CustomTorsionForce"); }
    int getNumTorsions() const { return (int)rows_.size(); }
    void setUsesPeriodicBoundaryConditions(bool p) { periodic_ = p; }
    bool usesPeriodicBoundaryConditions() const override { return periodic_; }
};

class CustomCompoundBondForce : public CustomForceBase {
  public:
    CustomCompoundBondForce(int numParticles, const std::string& energy) : CustomForceBase(energy), np_(numParticles) {}
    int addPerBondParameter(const std::string& name) { return add_param(name); }
    int addBond(const std::vector<int>& particles, const std::vector<double>& params = {});
    int getNumBonds() const { return (int)rows_.size(); }
    int getNumParticlesPerBond() const { return np_; }
    void setUsesPeriodicBoundaryConditions(bool p) { periodic_ = p; }
    bool usesPeriodicBoundaryConditions() const override { return periodic_; }

  private:
    friend class b200::SystemCompiler;
    int np_;
};

class CustomNonbondedForce : public CustomForceBase {
  public:
    enum NonbondedMethod { NoCutoff = 0, CutoffNonPeriodic = 1, CutoffPeriodic = 2 };
    explicit CustomNonbondedForce(const std::string& energy) : CustomForceBase(energy) {}
    int addPerParticleParameter(const std::string& name) { return add_param(name); }
    int addParticle(const std::vector<double>& params = {}) { return add_row({}, params, "CustomNonbondedForce"); }
    int getNumParticles() const { return (int)rows_.size(); }
    int addExclusion(int p1, int p2) { excl_.push_back(p1); excl_.push_back(p2); return (int)excl_.size() / 2 - 1; }
    int getNumExclusions() const { return (int)excl_.size() / 2; }
    int addInteractionGroup(const std::set<int>& set1, const std::set<int>& set2) { groups_.emplace_back(set1, set2); return (int)groups_.size() - 1; }
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

  private:
    friend class b200::SystemCompiler;
    std::vector<std::int32_t> excl_;
    std::vector<std::pair<std::set<int>, std::set<int>>> groups_;
    NonbondedMethod method_ = NoCutoff;
    double cutoff_ = 1.0, switch_ = -1.0;
    bool use_switch_ = false;
};

class CustomHbondForce : public CustomForceBase {
  public:
    enum NonbondedMethod { NoCutoff = 0, CutoffNonPeriodic = 1, CutoffPeriodic = 2 };
    explicit CustomHbondForce(const std::string& energy) : CustomForceBase(energy) {}
    int addPerDonorParameter(const std::string& name) { return add_param(name); }
    int addPerAcceptorParameter(const std::string& name) { aparams_.push_back(name); return (int)aparams_.size() - 1; }
    int addDonor(int d1, int d2, int d3, const std::vector<double>& params = {}) { return add_row({d1, d2, d3}, params, "CustomHbondForce donor"); }
    int addAcceptor(int a1, int a2, int a3, const std::vector<double>& params = {});
    int addExclusion(int donor, int acceptor);
    int getNumDonors() const { return (int)rows_.size(); }
    int getNumAcceptors() const { return (int)arows_.size(); }
    NonbondedMethod getNonbondedMethod() const { return method_; }
    void setNonbondedMethod(NonbondedMethod m) { method_ = m; }
    double getCutoffDistance() const { return cutoff_; }
    void setCutoffDistance(double d) { cutoff_ = d; }
    bool usesPeriodicBoundaryConditions() const override { return method_ == CutoffPeriodic; }

  private:
    friend class b200::SystemCompiler;
    std::vector<std::string> aparams_;
    std::vector<Row> arows_;
    std::vector<std::int32_t> excl_;
    NonbondedMethod method_ = NoCutoff;
    double cutoff_ = 1.0;
};

// One-bead external terms and the centroid spring (PositionRestraint, Pulling, EXVRectangularBox, CylindricalRestraint,
// HarmonicCoMPulling generators): translated like every other Custom*Force, by their expression.
class CustomExternalForce : public CustomForceBase {
  public:
    explicit CustomExternalForce(const std::string& energy) : CustomForceBase(energy) {}
    int addPerParticleParameter(const std::string& name) { return add_param(name); }
    int addParticle(int particle, const std::vector<double>& params = {}) { return add_row({particle}, params, "CustomExternalForce"); }
};
class CustomCentroidBondForce : public CustomForceBase {
  public:
    CustomCentroidBondForce(int numGroups, const std::string& energy) : CustomForceBase(energy), ng_(numGroups) {}
    int addPerBondParameter(const std::string& name) { return add_param(name); }
    int addGroup(const std::vector<int>& particles, const std::vector<double>& weights = {}) { groups_.emplace_back(particles, weights); return (int)groups_.size() - 1; }
    int addBond(const std::vector<int>& groups, const std::vector<double>& params = {}) { return add_row(groups, params, "CustomCentroidBondForce"); }
    void setUsesPeriodicBoundaryConditions(bool p) { periodic_ = p; }
    bool usesPeriodicBoundaryConditions() const override { return periodic_; }

  private:
    friend class b200::SystemCompiler;
    int ng_;
    std::vector<std::pair<std::vector<int>, std::vector<double>>> groups_;
};
// Monte Carlo barostats (MonteCarloAnisotropicBarostatGenerator.cpp:9-17, MonteCarloMembraneBarostatGenerator.cpp:6-14): the
// parameters are handed to ocm_context_set_barostat when the Context is created; the trials run inside ocm_step.
class MonteCarloAnisotropicBarostat : public Force {
  public:
    MonteCarloAnisotropicBarostat(const Vec3& defaultPressure, double defaultTemperature, bool scaleX = true, bool scaleY = true, bool scaleZ = true,
                                  int frequency = 25)
        : pressure_(defaultPressure), temperature_(defaultTemperature), frequency_(frequency) { scale_[0] = scaleX; scale_[1] = scaleY; scale_[2] = scaleZ; }
    const Vec3& getDefaultPressure() const { return pressure_; }
    double getDefaultTemperature() const { return temperature_; }
    bool getScaleX() const { return scale_[0]; }
    bool getScaleY() const { return scale_[1]; }
    bool getScaleZ() const { return scale_[2]; }
    int getFrequency() const { return frequency_; }
    void setFrequency(int f) { frequency_ = f; }
    int getRandomNumberSeed() const { return seed_; }
    void setRandomNumberSeed(int seed) { seed_ = seed; }
    bool usesPeriodicBoundaryConditions() const override { return true; }

  private:
    Vec3 pressure_;
    double temperature_;
    bool scale_[3];
    int frequency_, seed_ = 0;
};
class MonteCarloMembraneBarostat : public Force {
  public:
    enum XYMode { XYIsotropic = 0, XYAnisotropic = 1 };
    enum ZMode { ZFree = 0, ZFixed = 1, ConstantVolume = 2 };
    MonteCarloMembraneBarostat(double defaultPressure, double defaultSurfaceTension, double defaultTemperature, XYMode xymode, ZMode zmode,
                               int frequency = 25)
        : pressure_(defaultPressure), tension_(defaultSurfaceTension), temperature_(defaultTemperature), xymode_(xymode), zmode_(zmode),
          frequency_(frequency) {}
    double getDefaultPressure() const { return pressure_; }
    double getDefaultSurfaceTension() const { return tension_; }
    double getDefaultTemperature() const { return temperature_; }
    XYMode getXYMode() const { return xymode_; }
    ZMode getZMode() const { return zmode_; }
    int getFrequency() const { return frequency_; }
    void setFrequency(int f) { frequency_ = f; }
    int getRandomNumberSeed() const { return seed_; }
    void setRandomNumberSeed(int seed) { seed_ = seed; }
    bool usesPeriodicBoundaryConditions() const override { return true; }

  private:
    double pressure_, tension_, temperature_;
    XYMode xymode_;
    ZMode zmode_;
    int frequency_, seed_ = 0;
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
    void setDefaultPeriodicBoxVectors(const Vec3& a, const Vec3& b, const Vec3& c);
    void getDefaultPeriodicBoxVectors(Vec3& a, Vec3& b, Vec3& c) const { a = box_[0]; b = box_[1]; c = box_[2]; }
    int addForce(Force* f) { forces_.push_back(f); return (int)forces_.size() - 1; }  // takes ownership
    int getNumForces() const { return (int)forces_.size(); }
    const Force& getForce(int i) const { return *forces_.at(i); }
    Force& getForce(int i) { return *forces_.at(i); }
    bool usesPeriodicBoundaryConditions() const;

  private:
    friend class b200::SystemCompiler;
    std::vector<double> mass_;
    std::vector<Force*> forces_;
    Vec3 box_[3] = {Vec3(2, 0, 0), Vec3(0, 2, 0), Vec3(0, 0, 2)};
    bool has_box_ = false;
};

// ---- platform: one platform, the B200 engine.  It also answers to "CUDA", the name the reference's inputs and
// its default (src/input/ReadTOMLInput.cpp:1057) ask for.  Property "DeviceIndex" (OpenMM's CUDA property name)
// selects the GPU.
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

class LangevinIntegrator : public Integrator {
  public:
    LangevinIntegrator(double temperature, double frictionCoeff, double stepSize) : T_(temperature), gamma_(frictionCoeff) { dt_ = stepSize; }
    double getTemperature() const { return T_; }
    double getFriction() const { return gamma_; }
    int getRandomNumberSeed() const { return seed_; }
    void setRandomNumberSeed(int seed) { seed_ = seed; }  // 0: a random seed is drawn when the Context is created
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
    Context(Context&& other) noexcept;
    ~Context();
    const System& getSystem() const { return *sys_; }
    Integrator& getIntegrator() { return *integ_; }
    Platform& getPlatform() { return *platform_; }
    State getState(int types, bool enforcePeriodicBox = false, int groups = 0xFFFFFFFF) const;
    void setPositions(const std::vector<Vec3>& positions);
    void setVelocities(const std::vector<Vec3>& velocities);
    void setVelocitiesToTemperature(double temperature);
    void setVelocitiesToTemperature(double temperature, int randomSeed);
    double getTime() const;
    std::int64_t getStepCount() const;
    ocm_context* engineHandle() const { return h_; }  // the C-ABI context underneath

  private:
    friend class LangevinIntegrator;
    friend class LocalEnergyMinimizer;
    const System* sys_;
    Integrator* integ_;
    Platform* platform_;
    ocm_context* h_ = nullptr;
    // per-group energies are evaluated once per state: the reference's EnergyObserver asks for them group by group
    // one frame of observer output: the first getState() after a step fetches positions, velocities and all 32 group
    // energies with ONE asynchronous request (ocm_get_state_begin / _end: side-stream download from pinned staging); the
    // other getState() calls of the frame (Observer.cpp:103-118: positions, then velocities; Observer.cpp:324-346: one
    // call per force group) are served from this copy
    mutable bool energy_valid_ = false, frame_valid_ = false;
    mutable double ebg_[32];
    mutable double ekin_ = 0;
    mutable std::vector<Vec3> frame_pos_, frame_vel_;
    void create(const std::map<std::string, std::string>& properties);
    void invalidate() const { energy_valid_ = false; frame_valid_ = false; }
};

class LocalEnergyMinimizer {
  public:
    static void minimize(Context& context, double tolerance = 10.0, int maxIterations = 0);
};

}  // namespace OpenMM

#endif  // OCM_B200_OPENMM_COMPAT_H
