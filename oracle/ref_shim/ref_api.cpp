// ref_api.cpp — C entry points of oracle/_ref/libocm_ref.so (TEST INFRASTRUCTURE).
//
// libocm_ref.so = the reference's OWN sources (src/input/*.cpp, src/forcefield/*.cpp, src/Topology.cpp,
// src/SystemGenerator.cpp, src/IntegratorGenerator.cpp, src/Simulator.cpp, src/Observer.cpp, src/util/*.cpp),
// compiled unmodified from /root/reference by oracle/Makefile against the stand-in headers of this directory
// (OpenMM.h = openmm-lite, toml.hpp = toml11 interface over the repository's TOML reader).  These functions
// let the tests drive the reference's reader -> generators -> System pipeline and read back energies, forces
// and the complete content of every OpenMM::Force it built.
#include <OpenMM.h>
#include <toml.hpp>

#include <cstring>
#include <fstream>
#include <memory>
#include <sstream>
#include <string>

#include "openmm_lite_internal.h"
#include "toml_lite.hpp"  // this repository's TOML reader (cross-checked below against the toml11 the reference uses)
#include "src/IntegratorGenerator.hpp"
#include "src/Simulator.hpp"
#include "src/SystemGenerator.hpp"
#include "src/input/ReadGenesisInput.hpp"
#include "src/input/ReadTOMLInput.hpp"
#include "src/input/Utility.hpp"

namespace {

thread_local std::string g_err;

struct RefHandle {
    // low-level pipeline: reference reader -> SystemGenerator -> OpenMM::System (+ LangevinIntegrator, Context)
    std::unique_ptr<OpenMM::System> system;
    std::unique_ptr<IntegratorGeneratorBase> integ_gen;
    std::unique_ptr<OpenMM::Integrator> integrator;
    std::unique_ptr<OpenMM::Context> context;
    std::map<std::string, std::size_t> groups;
    std::string description;
    // or: the reference's Simulator as read_toml_input / read_genesis_input builds it
    std::unique_ptr<Simulator> simulator;
    OpenMM::Context* ctx = nullptr;
};

template <class F>
int guarded(F&& f) {
    try {
        f();
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    } catch (...) {
        g_err = "unknown exception";
        return 1;
    }
}

}  // namespace

extern "C" {

const char* ref_last_error(void) { return g_err.c_str(); }

// reference pipeline on a TOML input: ReadTOMLInput.cpp read_toml_system / read_toml_integrator_gen /
// read_toml_initial_conf / read_toml_initial_vel, SystemGenerator::generate, IntegratorGenerator::generate.
// The working directory must be the input's directory when it uses relative `include` files.
int ref_open_toml(const char* path, void** out) {
    return guarded([&] {
        auto h = std::make_unique<RefHandle>();
        auto data = toml::parse(std::string(path));
        expand_include(data);
        SystemGenerator sysgen = read_toml_system(data);
        h->groups = sysgen.ffname_groupid_map();
        h->system = sysgen.generate();
        h->integ_gen = read_toml_integrator_gen(data);
        h->integrator = h->integ_gen->generate();
        h->context = std::make_unique<OpenMM::Context>(*h->system, *h->integrator, OpenMM::Platform::getPlatformByName("Reference"));
        h->ctx = h->context.get();
        h->ctx->setPositions(read_toml_initial_conf(data));
        const auto& particles = toml::find<toml::array>(toml::find(data, "systems")[0], "particles");
        if (particles.at(0).contains("vel") || particles.at(0).contains("velocity")) h->ctx->setVelocities(read_toml_initial_vel(data));
        *out = h.release();
    });
}

// the reference's complete front end: read_toml_input / read_genesis_input -> Simulator (observers included;
// energy minimisation runs here when the input asks for it).  ref_run() then executes Simulator::run().
int ref_open_simulator(const char* path, void** out) {
    return guarded([&] {
        auto h = std::make_unique<RefHandle>();
        const std::string p(path);
        const std::string suffix = p.size() >= 5 ? p.substr(p.rfind('.')) : std::string();
        if (suffix == ".toml") h->simulator.reset(new Simulator(read_toml_input(p)));
        else if (suffix == ".inp") h->simulator.reset(new Simulator(read_genesis_input(p)));
        else throw std::runtime_error("ref_open_simulator: expected a .toml or .inp input");
        // the Simulator's Context member registers itself with openmm-lite when it is constructed
        h->ctx = OpenMM::Context::mostRecent();
        if (!h->ctx) throw std::runtime_error("ref_open_simulator: no live OpenMM::Context");
        *out = h.release();
    });
}

int ref_run(void* hv) {
    return guarded([&] {
        auto* h = static_cast<RefHandle*>(hv);
        if (!h->simulator) throw std::runtime_error("ref_run: handle was not opened with ref_open_simulator");
        h->simulator->run();
    });
}

void ref_close(void* hv) { delete static_cast<RefHandle*>(hv); }

int ref_num_particles(void* hv) { return static_cast<RefHandle*>(hv)->ctx->getSystem().getNumParticles(); }

// "name=groupid;name=groupid;..." in the order of SystemGenerator::ffname_groupid_map (the .ene columns)
int ref_group_map(void* hv, char* buf, long cap) {
    return guarded([&] {
        auto* h = static_cast<RefHandle*>(hv);
        std::string s;
        for (auto& kv : h->groups) s += kv.first + "=" + std::to_string(kv.second) + ";";
        if ((long)s.size() + 1 > cap) throw std::runtime_error("ref_group_map: buffer too small");
        std::memcpy(buf, s.c_str(), s.size() + 1);
    });
}

// JSON description of the OpenMM::System the reference built; call with buf == NULL to get the size
long ref_describe(void* hv, char* buf, long cap) {
    auto* h = static_cast<RefHandle*>(hv);
    try {
        if (h->description.empty()) h->description = OpenMM::lite::describe_system(h->ctx->getSystem());
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
    const long need = (long)h->description.size() + 1;
    if (buf && cap >= need) std::memcpy(buf, h->description.c_str(), need);
    return need;
}

int ref_set_positions(void* hv, const double* xyz) {
    return guarded([&] {
        auto* h = static_cast<RefHandle*>(hv);
        const int n = h->ctx->getSystem().getNumParticles();
        std::vector<OpenMM::Vec3> p(n);
        for (int i = 0; i < n; i++) p[i] = OpenMM::Vec3(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
        h->ctx->setPositions(p);
    });
}
int ref_set_velocities(void* hv, const double* xyz) {
    return guarded([&] {
        auto* h = static_cast<RefHandle*>(hv);
        const int n = h->ctx->getSystem().getNumParticles();
        std::vector<OpenMM::Vec3> p(n);
        for (int i = 0; i < n; i++) p[i] = OpenMM::Vec3(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
        h->ctx->setVelocities(p);
    });
}
int ref_set_velocities_to_temperature(void* hv, double T, int seed) {
    return guarded([&] { static_cast<RefHandle*>(hv)->ctx->setVelocitiesToTemperature(T, seed); });
}

// Context::getState(types, false, groups): any output may be NULL
int ref_get_state(void* hv, unsigned groups, double* pos, double* vel, double* frc, double* epot, double* ekin) {
    return guarded([&] {
        auto* h = static_cast<RefHandle*>(hv);
        int types = 0;
        if (pos) types |= OpenMM::State::Positions;
        if (vel) types |= OpenMM::State::Velocities;
        if (frc) types |= OpenMM::State::Forces;
        if (epot || ekin) types |= OpenMM::State::Energy;
        const OpenMM::State s = h->ctx->getState(types, false, (int)groups);
        const int n = h->ctx->getSystem().getNumParticles();
        auto copy = [n](const std::vector<OpenMM::Vec3>& v, double* o) { for (int i = 0; i < n; i++) for (int c = 0; c < 3; c++) o[3 * i + c] = v[i][c]; };
        if (pos) copy(s.getPositions(), pos);
        if (vel) copy(s.getVelocities(), vel);
        if (frc) copy(s.getForces(), frc);
        if (epot) *epot = s.getPotentialEnergy();
        if (ekin) *ekin = s.getKineticEnergy();
    });
}

// potential energy only (no kinetic-energy force pass): cheaper for large systems
int ref_potential_energy(void* hv, unsigned groups, double* epot, double* frc) {
    return guarded([&] {
        auto* h = static_cast<RefHandle*>(hv);
        std::vector<OpenMM::Vec3> F;
        *epot = h->ctx->computeForcesAndEnergy((int)groups, frc ? &F : nullptr);
        if (frc) for (std::size_t i = 0; i < F.size(); i++) for (int c = 0; c < 3; c++) frc[3 * i + c] = F[i][c];
    });
}

// State::getPeriodicBoxVectors (cuboid: the three diagonal entries)
int ref_get_box(void* hv, double* box3) {
    return guarded([&] {
        OpenMM::Vec3 a, b, c;
        static_cast<RefHandle*>(hv)->ctx->getState(0).getPeriodicBoxVectors(a, b, c);
        box3[0] = a[0]; box3[1] = b[1]; box3[2] = c[2];
    });
}

int ref_step(void* hv, int nsteps) {
    return guarded([&] { static_cast<RefHandle*>(hv)->ctx->getIntegrator().step(nsteps); });
}

int ref_minimize(void* hv, double tolerance, int max_iterations) {
    return guarded([&] { OpenMM::LocalEnergyMinimizer::minimize(*static_cast<RefHandle*>(hv)->ctx, tolerance, max_iterations); });
}

// integrator parameters as the reference's IntegratorGenerator produced them
int ref_integrator(void* hv, double* temperature, double* friction, double* dt, int* seed) {
    return guarded([&] {
        auto* li = dynamic_cast<OpenMM::LangevinIntegrator*>(&static_cast<RefHandle*>(hv)->ctx->getIntegrator());
        if (!li) throw std::runtime_error("not a LangevinIntegrator");
        *temperature = li->getTemperature();
        *friction = li->getFriction();
        *dt = li->getStepSize();
        *seed = li->getRandomNumberSeed();
    });
}

// the repository's TOML reader alone: file -> JSON (cross-checked against Python's tomllib and against toml11)
long ref_toml_to_json(const char* path, char* buf, long cap) {
    try {
        std::ostringstream os;
        tomllite::to_json(tomllite::parse_file(path), os);
        const std::string s = os.str();
        const long need = (long)s.size() + 1;
        if (buf && cap >= need) std::memcpy(buf, s.c_str(), need);
        return need;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

// the same file through the toml library this build of the reference was compiled with (its own toml11 submodule when
// that is checked out): toml::parse -> JSON
static void toml11_json(const toml::value& v, std::ostream& os) {
    if (v.is_boolean()) os << (v.as_boolean() ? "true" : "false");
    else if (v.is_integer()) os << v.as_integer();
    else if (v.is_floating()) {
        const double d = v.as_floating();
        if (std::isnan(d)) os << "\"nan\"";
        else if (std::isinf(d)) os << (d > 0 ? "\"inf\"" : "\"-inf\"");
        else { std::ostringstream t; t.precision(17); t << d; std::string q = t.str(); if (q.find_first_of(".en") == std::string::npos) q += ".0"; os << q; }
    } else if (v.is_string()) {
        os << '"';
        for (unsigned char c : std::string(v.as_string())) {
            if (c == '"' || c == '\\') os << '\\' << c;
            else if (c < 0x20) { char b[8]; std::snprintf(b, sizeof b, "\\u%04x", c); os << b; }
            else os << c;
        }
        os << '"';
    } else if (v.is_array()) {
        os << '[';
        bool first = true;
        for (const auto& e : v.as_array()) { if (!first) os << ','; first = false; toml11_json(e, os); }
        os << ']';
    } else if (v.is_table()) {
        // sorted keys, as the other dumper writes them
        std::map<std::string, const toml::value*> sorted;
        for (const auto& kv : v.as_table()) sorted.emplace(kv.first, &kv.second);
        os << '{';
        bool first = true;
        for (const auto& kv : sorted) { if (!first) os << ','; first = false; toml11_json(toml::value(kv.first), os); os << ':'; toml11_json(*kv.second, os); }
        os << '}';
    } else os << "null";
}
long ref_toml11_to_json(const char* path, char* buf, long cap) {
    try {
        std::ostringstream os;
        toml11_json(toml::parse(std::string(path)), os);
        const std::string s = os.str();
        const long need = (long)s.size() + 1;
        if (buf && cap >= need) std::memcpy(buf, s.c_str(), need);
        return need;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

}  // extern "C"
