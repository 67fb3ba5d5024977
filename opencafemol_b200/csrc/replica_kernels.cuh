// replica_kernels.cuh — whole-step persistent kernel for small systems (configs C1 and C4: an SH3 domain of 56 beads,
// alone or as an ensemble of independent replicas).
//
// One thread block owns one replica for a whole launch of `nsteps` MD steps: positions, velocities and the fixed-point force
// accumulator live in shared memory, every non-excluded bead pair of the replica is visited each step from a static pair
// table (no neighbour list: 1 540 pairs for 56 beads), the bonded terms are the same device functions as in force_kernel
// (double precision), the Langevin update draws the same Philox4x32-10 numbers as integrate_kernel.  Only __syncthreads
// separates the phases of a step: no grid synchronisation, no global counters, no per-step kernel launches.
//
// MEASURED (B200, round 2) and therefore OFF by default (option replica_kernel = 1 turns it on): one SH3 replica steps at
// 75 k steps/s here against 118 k steps/s through the general path (2 launches per step); 512 replicas 17.2 M against 19.3 M
// replica-steps/s.  A step of one replica costs ~11 k double-precision warp instructions in its ~480 bonded terms, and one
// SM issues 0.5 of those per clock: 11 us per step on one SM, where the general path spreads the same terms over 4 SMs.
// So these small systems are bound by FP64 throughput of the bonded terms, not by launches; a version spreading one replica
// over a cluster of 8 CTAs (DSMEM) measured 59 k steps/s and was removed.  The lever that remains is FP32 bonded arithmetic.
//
// What it stands for on the reference side: Simulator::run's integrator->step(save_step) (src/Simulator.cpp:55-60) for the
// small AICG2+ samples (sample/toml_input/sh3_AICG2+.toml); OpenMM launches 10-20 kernels per step for such a system.
#pragma once
#include "kernels.cuh"

namespace ocm {

constexpr int kRepMaxLists = 6;
constexpr int kRepBlock = 256;

// static pair table of one pair list of ONE replica: every participant pair that is neither excluded nor forbidden by the
// interaction classes, with its coefficients (two-zone 12-6: the entry of PairListDev::tabF; screened Coulomb: K q_i q_j)
struct RepList {
    int npairs, form, has_switch, pad;
    float rc2, k_in, rsw, inv_sw, inv_debye;
    const ushort2* ij;     // local bead indices, i < j
    const float4* coef;
};

struct RepParams {
    int nper, nrep, nlists, periodic_unused;
    RepList lists[kRepMaxLists];
    BondedDev bonded[10];  // local bead indices (B_NUM_KINDS entries)
    double dt;
    float a, b, c;         // Langevin coefficients (StepParams)
    unsigned int seed_lo, seed_hi;
    const uint2* rep_seed; // per-replica noise keys or null
};

// two-zone 12-6 force over r (tz_pair of kernels.cuh, one pair, no votes)
__device__ __forceinline__ float rep_pair_tz(const RepList& L, const float4 c, float r2) {
    const float ir2 = fast_rcp(r2);
    const float ir6 = ir2 * ir2 * ir2;
    const float zone = (r2 < c.z) ? L.k_in : c.w;
    const float a12 = c.x * ir6;
    float fr = zone * (a12 - c.y) * ir6 * ir2;
    if (L.has_switch && r2 > L.rsw * L.rsw) {
        const float ir = fast_rsqrt(r2);
        const float tt = __saturatef((r2 * ir - L.rsw) * L.inv_sw), u = 1.f - tt;
        const float S = u * u * u * fmaf(tt, fmaf(tt, 6.f, 3.f), 1.f);
        const float tu = tt * u;
        const float dS = -30.f * L.inv_sw * tu * tu;
        const float E = zone * (a12 * (1.f / 12.f) - c.y * (1.f / 6.f)) * ir6;
        fr = S * fr - E * dS * ir;
    }
    return fr;
}

// Forces are accumulated with fire-and-forget 64-bit integer atomics on the engine's global accumulator (L2): 64-bit adds
// on SHARED memory compile to a compare-and-swap spin loop (ATOMS.CAST.SPIN), which under ~40 contributions per address
// and step made the step 4 x slower than the arithmetic.
__global__ void __launch_bounds__(kRepBlock) replica_step_kernel(const __grid_constant__ RepParams R, double* __restrict__ x_g, float4* __restrict__ v_g,
                                                                 unsigned long long* __restrict__ fbuf, int n_total, unsigned long long step0, int nsteps,
                                                                 int* __restrict__ flags) {
    extern __shared__ __align__(16) unsigned char rep_smem[];
    const int n = R.nper, rep = blockIdx.x, tid = threadIdx.x;
    double* xs = reinterpret_cast<double*>(rep_smem);                              // [3n] master positions
    float4* vs = reinterpret_cast<float4*>(xs + 3 * n);                             // [n] velocities, w = 1/m
    float4* xf = vs + n;                                                            // [n] FP32 offsets from the replica's origin
    unsigned long long* fb = fbuf + (size_t)rep * n;                                // this replica's slice: fb[q * n_total + i]
    __shared__ double origin[3];
    __shared__ int bad;
    for (int k = tid; k < 3 * n; k += kRepBlock) xs[k] = x_g[(size_t)3 * n * rep + k];
    for (int i = tid; i < n; i += kRepBlock) vs[i] = v_g[(size_t)n * rep + i];
    if (tid == 0) bad = 0;
    const uint2 key = R.rep_seed ? R.rep_seed[rep] : make_uint2(R.seed_lo, R.seed_hi);
    // the bonded device functions read a StepParams: n is the stride of the accumulator, the box is unused (unlimited boundary)
    StepParams P{};
    P.n = n_total; P.periodic = 0;
    for (int k = 0; k < 3; k++) { P.box[k] = 1.0; P.inv_box[k] = 1.0; }
    P.role = nullptr;
    __syncthreads();
    for (int s = 0; s < nsteps; s++) {
        // ---- FP32 pair-kernel positions relative to the replica's centroid (refreshed every 256 steps) ----
        if ((s & 255) == 0) {
            if (tid < 3) {
                double c = 0;
                for (int i = 0; i < n; i++) c += xs[3 * i + tid];
                origin[tid] = c / n;
            }
            __syncthreads();
        }
        for (int i = tid; i < n; i += kRepBlock)
            xf[i] = make_float4((float)(xs[3 * i] - origin[0]), (float)(xs[3 * i + 1] - origin[1]), (float)(xs[3 * i + 2] - origin[2]), 0.f);
        __syncthreads();
        // ---- pair terms: the static table of every list ----
        for (int l = 0; l < R.nlists; l++) {
            const RepList& L = R.lists[l];
            for (int p = tid; p < L.npairs; p += kRepBlock) {
                const ushort2 ij = __ldg(&L.ij[p]);
                const float4 pi = xf[ij.x], pj = xf[ij.y];
                const float dx = __fsub_rn(pi.x, pj.x), dy = __fsub_rn(pi.y, pj.y), dz = __fsub_rn(pi.z, pj.z);
                const float r2 = norm2(dx, dy, dz);
                if (!(r2 < L.rc2)) continue;
                const float4 c = __ldg(&L.coef[p]);
                float fr;
                if (L.form == SEG_PAIR_DH) {
                    const float ir = fast_rsqrt(r2), r = r2 * ir;
                    fr = c.x * __expf(-r * L.inv_debye) * (ir + L.inv_debye) * ir * ir;
                } else fr = rep_pair_tz(L, c, r2);
                const float fx = fr * dx, fy = fr * dy, fz = fr * dz;
                const long long qx = (long long)(fx * 4294967296.0f), qy = (long long)(fy * 4294967296.0f), qz = (long long)(fz * 4294967296.0f);
                atomicAdd(&fb[ij.x], (unsigned long long)qx); atomicAdd(&fb[n_total + ij.x], (unsigned long long)qy); atomicAdd(&fb[2 * (size_t)n_total + ij.x], (unsigned long long)qz);
                atomicAdd(&fb[ij.y], (unsigned long long)(-qx)); atomicAdd(&fb[n_total + ij.y], (unsigned long long)(-qy)); atomicAdd(&fb[2 * (size_t)n_total + ij.y], (unsigned long long)(-qz));
            }
        }
        // ---- bonded terms (double precision on the master positions) ----
#pragma unroll 1
        for (int kind = 0; kind < 10; kind++) {
            const BondedDev& B = R.bonded[kind];
            for (int m = tid; m < B.n; m += kRepBlock) { int g; bonded_term(kind, B, m, P, xs, fb, &g); }
        }
        __threadfence();   // the atomics of this block are performed at L2 before anybody reads the sums
        __syncthreads();
        // ---- leap-frog Langevin update (integrate_kernel, same noise stream) ----
        const unsigned long long step = step0 + (unsigned long long)s;
        for (int i = tid; i < n; i += kRepBlock) {
            const float inv32 = 2.3283064365386963e-10f;
            const float fx = (float)(long long)__ldcg(&fb[i]) * inv32, fy = (float)(long long)__ldcg(&fb[n_total + i]) * inv32,
                        fz = (float)(long long)__ldcg(&fb[2 * (size_t)n_total + i]) * inv32;
            __stcg(&fb[i], 0ull); __stcg(&fb[n_total + i], 0ull); __stcg(&fb[2 * (size_t)n_total + i], 0ull);
            float4 vi = vs[i];
            const float im = vi.w;
            if (im > 0.f) {
                float g0, g1, g2;
                normals3(philox4x32_10(make_uint4((unsigned int)i, (unsigned int)step, (unsigned int)(step >> 32), 0u), key), g0, g1, g2);
                const float cs = R.c * sqrtf(im), bm = R.b * im;
                vi.x = R.a * vi.x + bm * fx + cs * g0;
                vi.y = R.a * vi.y + bm * fy + cs * g1;
                vi.z = R.a * vi.z + bm * fz + cs * g2;
                vs[i] = vi;
                const double px = xs[3 * i] + (double)vi.x * R.dt, py = xs[3 * i + 1] + (double)vi.y * R.dt, pz = xs[3 * i + 2] + (double)vi.z * R.dt;
                xs[3 * i] = px; xs[3 * i + 1] = py; xs[3 * i + 2] = pz;
                if (!(px == px) || !(py == py) || !(pz == pz) || !(fx == fx)) bad = 1;
            }
        }
        __threadfence();   // the cleared accumulator is visible before the next step's atomics
        __syncthreads();
        if (bad) break;  // (block-uniform: read after the barrier)
    }
    for (int k = tid; k < 3 * n; k += kRepBlock) x_g[(size_t)3 * n * rep + k] = xs[k];
    for (int i = tid; i < n; i += kRepBlock) v_g[(size_t)n * rep + i] = vs[i];
    if (tid == 0 && bad) flags[FLAG_NAN] = 1;
}

}  // namespace ocm
