// Host layer (opencafemol_b200/csrc/compat): a Custom*Force is recognised by its energy expression in canonical form (white space
// removed, auxiliary definitions sorted, generator id prefix removed).  Built and run by tests/test_frontend_dropin.py; no GPU needed:
// the translation of the System into the C ABI is host code, and it runs before ocm_context_create looks for a device.
//   exit 0: every case behaved as expected, 1 otherwise (the offending case on stderr)
#include <OpenMM.h>

#include <cstdio>
#include <string>

// "translated" = the Context constructor got as far as the device (no GPU here) or succeeded (on a GPU box)
static int run(const char* what, OpenMM::Force* force, bool expect_translated) {
    OpenMM::System sys;
    for (int i = 0; i < 4; i++) sys.addParticle(12.0);
    sys.addForce(force);
    OpenMM::LangevinIntegrator integ(300.0, 0.2, 0.002);
    integ.setRandomNumberSeed(7);
    std::string msg;
    bool translated = true;
    try {
        OpenMM::Context ctx(sys, integ);
    } catch (const std::exception& e) {
        msg = e.what();
        translated = msg.find("no CUDA device") != std::string::npos;
    }
    if (translated != expect_translated) {
        std::fprintf(stderr, "%s: expected %s, got: %s\n", what, expect_translated ? "a translated force" : "a refusal", msg.c_str());
        return 1;
    }
    return 0;
}

int main() {
    int bad = 0;
    {
        // GaussianDihedralForceFieldGenerator.cpp:5-13 as the generator writes it
        auto* f = new OpenMM::CustomTorsionForce(
            "GD3_k * exp(-(dt_periodic)^2 / (2*GD3_sigma^2));"
            "dt_periodic = dt - floor((dt+pi)/(2*pi))*(2*pi);"
            "dt = theta - GD3_theta0;"
            "pi = 3.1415926535897932385;");
        f->addPerTorsionParameter("GD3_k"); f->addPerTorsionParameter("GD3_theta0"); f->addPerTorsionParameter("GD3_sigma");
        f->addTorsion(0, 1, 2, 3, {-1.0, 0.5, 0.3});
        bad += run("gaussian dihedral, as written", f, true);
    }
    {
        // the same energy with its definitions in another order and other spacing
        auto* f = new OpenMM::CustomTorsionForce(
            "GD9_k*exp( -(dt_periodic)^2/(2*GD9_sigma^2) ) ;\n"
            "  pi = 3.1415926535897932385 ;\n"
            "  dt = theta - GD9_theta0 ;\n"
            "  dt_periodic = dt - floor((dt+pi)/(2*pi))*(2*pi)");
        f->addPerTorsionParameter("GD9_k"); f->addPerTorsionParameter("GD9_theta0"); f->addPerTorsionParameter("GD9_sigma");
        f->addTorsion(0, 1, 2, 3, {-1.0, 0.5, 0.3});
        bad += run("gaussian dihedral, definitions re-ordered", f, true);
    }
    {
        // a different energy must be refused, not silently mapped onto a neighbour
        auto* f = new OpenMM::CustomTorsionForce(
            "GD4_k * exp(-(dt_periodic)^4 / (2*GD4_sigma^2));"
            "dt_periodic = dt - floor((dt+pi)/(2*pi))*(2*pi);"
            "dt = theta - GD4_theta0;"
            "pi = 3.1415926535897932385;");
        f->addPerTorsionParameter("GD4_k"); f->addPerTorsionParameter("GD4_theta0"); f->addPerTorsionParameter("GD4_sigma");
        f->addTorsion(0, 1, 2, 3, {-1.0, 0.5, 0.3});
        bad += run("modified exponent", f, false);
    }
    {
        // PullingForceFieldGenerator.cpp:5-7 and an external force that is not one of the reference's
        auto* f = new OpenMM::CustomExternalForce("-PL0_kx * x - PL0_ky * y - PL0_kz * z");
        f->addPerParticleParameter("PL0_kx"); f->addPerParticleParameter("PL0_ky"); f->addPerParticleParameter("PL0_kz");
        f->addParticle(2, {1.0, 0.0, -1.0});
        bad += run("pulling", f, true);
        auto* g = new OpenMM::CustomExternalForce("Q_k * x^2");
        g->addPerParticleParameter("Q_k");
        g->addParticle(1, {3.0});
        bad += run("unknown external force", g, false);
    }
    return bad ? 1 : 0;
}
