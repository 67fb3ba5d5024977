// internal helpers of openmm-lite shared between openmm_lite.cpp and ref_api.cpp (TEST INFRASTRUCTURE)
#ifndef OCM_REF_SHIM_OPENMM_LITE_INTERNAL_H
#define OCM_REF_SHIM_OPENMM_LITE_INTERNAL_H

#include <cmath>
#include <cstdint>
#include <string>

namespace OpenMM {
class System;
namespace lite {

// Philox4x32-10 (Salmon et al. 2011): the repository's documented noise source, key = (seed_lo, seed_hi),
// counter = (bead, step_lo, step_hi, stream); three standard normals per call by Box-Muller.
inline void philox4x32_10(std::uint32_t c0, std::uint32_t c1, std::uint32_t c2, std::uint32_t c3, std::uint32_t k0,
                          std::uint32_t k1, std::uint32_t out[4]) {
    for (int r = 0; r < 10; r++) {
        const std::uint64_t p0 = (std::uint64_t)0xD2511F53u * c0, p1 = (std::uint64_t)0xCD9E8D57u * c2;
        const std::uint32_t n0 = (std::uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (std::uint32_t)p1;
        const std::uint32_t n2 = (std::uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (std::uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
inline void philox_normals3(std::uint64_t seed, std::uint32_t bead, std::uint64_t step, std::uint32_t stream, double g[3]) {
    std::uint32_t o[4];
    philox4x32_10(bead, (std::uint32_t)step, (std::uint32_t)(step >> 32), stream, (std::uint32_t)seed, (std::uint32_t)(seed >> 32), o);
    const double two_pi = 6.283185307179586476925, s = 1.0 / 4294967296.0;
    const double u0 = (o[0] + 0.5) * s, u1 = (o[1] + 0.5) * s, u2 = (o[2] + 0.5) * s, u3 = (o[3] + 0.5) * s;
    const double r0 = std::sqrt(-2.0 * std::log(u0)), r1 = std::sqrt(-2.0 * std::log(u2));
    g[0] = r0 * std::cos(two_pi * u1);
    g[1] = r0 * std::sin(two_pi * u1);
    g[2] = r1 * std::cos(two_pi * u3);
}

std::string describe_system(const System& sys);

}  // namespace lite
}  // namespace OpenMM
#endif
