// replica_math.hpp — single-precision bonded terms for the whole-step replica kernel (replica_kernels.cuh).
//
// Same terms as bonded_eval / bonded_term of kernels.cuh (which cite the reference generators), restated so that the
// arithmetic after the bead-to-bead differences runs in FP32: the general path evaluates them in double, and a 56-bead
// replica then spends its whole step in the FP64 pipe (measured, see replica_kernels.cuh).  What keeps FP32 inside the
// 1e-5 force bar:
//   * the differences p_i - p_j are taken in double on the master positions and only then rounded to float;
//   * bond lengths of the stiff 2-body kinds (harmonic, gaussian, 3SPN.2 bond) are taken in double too (r - r0 cancels);
//   * angles and dihedrals come from atan2 of (|a x b|, a.b), which has no ill-conditioned range, instead of acos;
//   * the angle gradient uses the cross-product form (OpenMM's angle kernel), finite for every non-degenerate geometry.
// Everything is templated on the real type so that the double instantiation can be checked against the float one and
// against finite differences on the CPU (tests/c/replica_math_check.cpp); host and device compile the same source.
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define OCM_HD __host__ __device__ __forceinline__
#else
#define OCM_HD inline
#endif

namespace ocm {
namespace repmath {

template <typename T> struct V3 { T x, y, z; };
template <typename T> OCM_HD T dot(const V3<T>& a, const V3<T>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <typename T> OCM_HD V3<T> cross(const V3<T>& a, const V3<T>& b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
template <typename T> OCM_HD V3<T> scale(T s, const V3<T>& a) { return {s * a.x, s * a.y, s * a.z}; }

// Device FP32: hardware approximations (MUFU, <= 2 ulp) instead of the IEEE sequences of sqrtf and '/' — each saves ~8
// instructions and a slow-path branch, and a relative 2e-7 on a force is far inside the 1e-5 bar.  Host and double: exact.
#if defined(__CUDA_ARCH__)
OCM_HD float r_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
OCM_HD float r_rsqrt(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
OCM_HD float r_sqrt(float x) { return x * r_rsqrt(x > 1e-37f ? x : 1e-37f); }
#else
OCM_HD float r_rcp(float x) { return 1.0f / x; }
OCM_HD float r_rsqrt(float x) { return 1.0f / sqrtf(x); }
OCM_HD float r_sqrt(float x) { return sqrtf(x); }
#endif
OCM_HD double r_rcp(double x) { return 1.0 / x; }
OCM_HD double r_rsqrt(double x) { return 1.0 / sqrt(x); }
OCM_HD double r_sqrt(double x) { return sqrt(x); }
OCM_HD float r_exp(float x) { return expf(x); }
OCM_HD double r_exp(double x) { return exp(x); }
OCM_HD float r_atan2(float y, float x) { return atan2f(y, x); }
OCM_HD double r_atan2(double y, double x) { return atan2(y, x); }
OCM_HD float r_floor(float x) { return floorf(x); }
OCM_HD double r_floor(double x) { return floor(x); }
#if defined(__CUDA_ARCH__)
// argument folded into [-pi, pi] first (there __sincosf is good to 2^-21 absolute), then MUFU.SIN / MUFU.COS
OCM_HD void r_sincos(float x, float& s, float& c) {
    const float k = rintf(x * 0.15915494309189535f);
    x = fmaf(k, -6.2831854820251465f, x); x = fmaf(k, 1.7484555e-7f, x);   // 2 pi = 6.2831854820251465 - 1.7484555e-7
    __sincosf(x, &s, &c);
}
#else
OCM_HD void r_sincos(float x, float& s, float& c) { s = sinf(x); c = cosf(x); }
#endif
OCM_HD void r_sincos(double x, double& s, double& c) { s = sin(x); c = cos(x); }

// E(q) and dE/dq of one bonded term, q = r, theta or phi.  Parameter layouts as in kernels.cuh::bonded_eval
// (HarmonicBond r0,k | GaussianBond k,r0,sigma | GoContact k,r0,slope,intercept,cap | 3SPN2Bond k2,k4,r0 |
//  HarmonicAngle theta0,k | FlexibleLocalAngle k + table | GaussianDihedral k,phi0,sigma | CosineDihedral k,t0,n |
//  FlexibleLocalDihedral k + table).  For kinds 0, 1, 3 the caller passes dq = r - r0 computed in double (see above).
template <typename T>
OCM_HD void eval(int kind, const T* p, const T* tab, T q, T dq, T& E, T& dE) {
    switch (kind) {
        case 0: { E = T(0.5) * p[1] * dq * dq; dE = p[1] * dq; } break;
        case 1: { const T is2 = r_rcp(p[2] * p[2]); E = p[0] * r_exp(T(-0.5) * dq * dq * is2); dE = -E * dq * is2; } break;
        case 2: {
            const T ir0 = r_rcp(p[1]), iq = r_rcp(q);
            const T u = q * ir0;
            if (p[4] - u >= T(0)) { E = p[0] * (p[2] * u + p[3]); dE = p[0] * p[2] * ir0; }
            else {
                const T s = p[1] * iq, s2 = s * s, s4 = s2 * s2, s10 = s4 * s4 * s2, s12 = s10 * s2;
                E = p[0] * (T(5) * s12 - T(6) * s10); dE = p[0] * T(60) * (s10 - s12) * iq;
            }
        } break;
        case 3: { const T d2 = dq * dq; E = p[0] * d2 + p[1] * d2 * d2; dE = T(2) * p[0] * dq + T(4) * p[1] * d2 * dq; } break;
        case 4: { const T d = q - p[0]; E = T(0.5) * p[1] * d * d; dE = p[1] * d; } break;
        case 5: {
            const T tmin = tab[20], tmax = tab[21];
            T val = 0, der = 0;
            if (q >= tmin && q <= tmax) {
                const T h = (tmax - tmin) * T(1.0 / 9.0), ih = r_rcp(h);
                int k = (int)r_floor((q - tmin) * ih);
                k = k < 0 ? 0 : (k > 8 ? 8 : k);
                const T a = (tmin + T(k + 1) * h - q) * ih, b = T(1) - a;
                const T* y = tab; const T* y2 = tab + 10;
                val = a * y[k] + b * y[k + 1] + ((a * a * a - a) * y2[k] + (b * b * b - b) * y2[k + 1]) * h * h * T(1.0 / 6.0);
                der = (y[k + 1] - y[k]) * ih + ((T(1) - T(3) * a * a) * y2[k] + (T(3) * b * b - T(1)) * y2[k + 1]) * h * T(1.0 / 6.0);
            }
            T e = val - tab[22], de = der;
            if (tmin - q >= T(0)) { e += T(30) * (tmin - q) + tab[0]; de -= T(30); }
            if (q - tmax >= T(0)) { e += T(30) * (q - tmax) + tab[9]; de += T(30); }
            E = p[0] * e; dE = p[0] * de;
        } break;
        case 6: {
            const T PI = T(3.14159265358979323846);
            const T dt = q - p[1];
            const T d = dt - r_floor((dt + PI) * T(0.15915494309189533577)) * T(2) * PI;
            const T is2 = r_rcp(p[2] * p[2]);
            E = p[0] * r_exp(T(-0.5) * d * d * is2); dE = -E * d * is2;
        } break;
        case 7: {
            T sn, cs;
            r_sincos(p[2] * (q - p[1]), sn, cs);
            E = p[0] * (T(1) - cs); dE = p[0] * p[2] * sn;
        } break;
        default: {
            T s1, c1;
            r_sincos(q, s1, c1);
            const T s2 = T(2) * s1 * c1, c2 = c1 * c1 - s1 * s1;
            const T s3 = s1 * c2 + c1 * s2, c3 = c1 * c2 - s1 * s2;
            const T e = tab[0] - tab[7] + tab[1] * c1 + tab[2] * s1 + tab[3] * c2 + tab[4] * s2 + tab[5] * c3 + tab[6] * s3;
            const T de = -tab[1] * s1 + tab[2] * c1 + T(2) * (-tab[3] * s2 + tab[4] * c2) + T(3) * (-tab[5] * s3 + tab[6] * c3);
            E = p[0] * e; dE = p[0] * de;
        }
    }
}

// 2-body term on d = p_i - p_j: force on i (the force on j is its negative).  r_minus_r0: see eval().
template <typename T>
OCM_HD T bond(int kind, const T* p, const T* tab, const V3<T>& d, T r, T r_minus_r0, V3<T>& fi) {
    T E, dE;
    eval(kind, p, tab, r, r_minus_r0, E, dE);
    fi = scale(-dE * r_rcp(r), d);
    return E;
}

// 3-body term, a = p_i - p_j, b = p_k - p_j (vertex j): theta = atan2(|a x b|, a.b);
// F_i = dE (c x a) / (|a|^2 |c|), F_k = dE (b x c) / (|b|^2 |c|), c = a x b   (OpenMM angleForce: force1, force3)
template <typename T>
OCM_HD T angle(int kind, const T* p, const T* tab, const V3<T>& a, const V3<T>& b, V3<T>& fi, V3<T>& fk) {
    const V3<T> c = cross(a, b);
    const T lc2 = dot(c, c);
    const T lc = r_sqrt(lc2 > T(1e-30) ? lc2 : T(1e-30));
    const T th = r_atan2(lc, dot(a, b));
    T E, dE;
    eval(kind, p, tab, th, T(0), E, dE);
    fi = scale(dE * r_rcp(dot(a, a) * lc), cross(c, a));
    fk = scale(dE * r_rcp(dot(b, b) * lc), cross(b, c));
    return E;
}

// 4-body term, F = p1 - p2, G = p2 - p3, H = p4 - p3, OpenMM's sign rule (phi < 0 when F.(H x G) > 0);
// sin(phi) = -|G| F.B / (|A||B|), cos(phi) = A.B / (|A||B|), A = F x G, B = H x G
template <typename T>
OCM_HD T dihedral(int kind, const T* p, const T* tab, const V3<T>& Fv, const V3<T>& G, const V3<T>& H,
                  V3<T>& f1, V3<T>& f2, V3<T>& f3, V3<T>& f4) {
    const V3<T> A = cross(Fv, G), B = cross(H, G);
    const T iA2 = r_rcp(dot(A, A)), iB2 = r_rcp(dot(B, B)), G2 = dot(G, G), ilG = r_rsqrt(G2), lG = G2 * ilG;
    const T phi = r_atan2(-lG * dot(Fv, B), dot(A, B));
    T E, dE;
    eval(kind, p, tab, phi, T(0), E, dE);
    const T FG = dot(Fv, G), HG = dot(H, G);
    const T ca = lG * iA2, cb = lG * iB2, ea = FG * iA2 * ilG, eb = HG * iB2 * ilG;
    // dphi/dp1 = -ca A ; dphi/dp4 = cb B ; dphi/dp2 = (ca+ea) A - eb B ; dphi/dp3 = -(cb-eb) B - ea A ; force = -dE dphi/dp
    f1 = scale(dE * ca, A);
    f4 = scale(-dE * cb, B);
    f2 = {-dE * ((ca + ea) * A.x - eb * B.x), -dE * ((ca + ea) * A.y - eb * B.y), -dE * ((ca + ea) * A.z - eb * B.z)};
    f3 = {dE * ((cb - eb) * B.x + ea * A.x), dE * ((cb - eb) * B.y + ea * A.y), dE * ((cb - eb) * B.z + ea * A.z)};
    return E;
}

}  // namespace repmath
}  // namespace ocm
