// replica_kernels.cuh — whole-step persistent kernel for small systems (configs C1 and C4: an SH3 domain of 56 beads,
// alone or as an ensemble of independent replicas).
//
// One thread block owns one replica for a whole launch of `nsteps` MD steps.  Positions, velocities, the noise of the
// step and every force contribution live in shared memory; two __syncthreads per step are the only synchronisation: no
// grid synchronisation, no global counters, no per-step kernel launches, and NO ATOMICS:
//   phase 1  warp w runs the warp tasks w, w + nwarps, ... of a static list sorted longest first (drawing tasks from a shared
//            atomic counter instead measured 6 % slower for 512 replicas: with 4 blocks per SM the other blocks fill the gaps) —
//              pair task    8 beads x 4 lanes: each bead gathers its force from its partners, every pair seen from both
//                           sides instead of being added atomically.  Partners come from a per-bead Verlet list in shared
//                           memory (4 interleaved sub-lists per bead, one per lane, one byte per entry), rebuilt by the
//                           same task from a scan of all beads when any bead has moved half the skin since the last
//                           build; coefficients from a full n x n table (zero for excluded / non-interacting pairs);
//              bonded task  32 terms of one kind in FP32 (replica_math.hpp: differences and stiff bond lengths in double,
//                           everything after that in single precision); each term writes the force on each of its beads
//                           into its own scratch slot (slot = position in the bead's list * n + bead: beads that are
//                           neighbours along the chain are neighbours in memory, writes and reads without bank conflicts);
//              noise task   Philox4x32-10 + Box-Muller for 32 beads (the same streams as integrate_kernel);
//   phase 2  one thread per (bead, component): sum of the bead's scratch slots in a fixed order, pair force,
//            Langevin leap-frog update, master position in double, FP32 pair-kernel position for the next step.
// Sums run in a fixed order, so a run is bit-reproducible.
//
// Why this shape (measured on B200, round 2).  The first version of this kernel kept the double-precision bonded terms
// of force_kernel and accumulated with 64-bit fixed-point atomics: 75 k steps/s for one SH3 replica against 118 k through
// the general path (2 launches per step), 17.2 M against 19.3 M replica-steps/s for 512 replicas.  A step of one replica
// cost ~11 k FP64 warp instructions in its ~480 bonded terms, and shared-memory atomics cost 2 cycles per LANE
// (64-bit ones are a compare-and-swap loop).  Hence FP32 bonded arithmetic and gather instead of scatter.
//
// What it stands for on the reference side: Simulator::run's integrator->step(save_step) (src/Simulator.cpp:55-60) for the
// small AICG2+ samples (sample/toml_input/sh3_AICG2+.toml); OpenMM launches 10-20 kernels per step for such a system.
#pragma once
#include "kernels.cuh"
#include "replica_math.hpp"

namespace ocm {

constexpr int kRepMaxLists = 6;
constexpr int kRepBlockMax = 1024;   // threads per block: 256 for ensembles that fill the GPU, up to 1024 for a few replicas
constexpr int kRepMaxWarps = kRepBlockMax / 32;
constexpr int kRepNbMax = 16;        // entries per (bead, lane) sub-list of the in-kernel Verlet lists; overflow = scan every step
constexpr float kRepSkin = 0.2f;     // nm
constexpr int kRepTaskPair = 100;    // RepTask::kind: 0..8 = bonded kind, 100 + l = pair list l, 200 = noise
constexpr int kRepTaskNoise = 200;

// one pair list of ONE replica: full coefficient table (two-zone 12-6: the entry of PairListDev::tabF; screened
// Coulomb: K q_i q_j); zero rows / entries for beads that do not take part, excluded pairs and the diagonal
struct RepList {
    int form, has_switch;
    float rc2, rl2, k_in, rsw, inv_sw, inv_debye;   // rl2 = (cutoff + kRepSkin)^2
    const float4* coef;    // [nper * nper]
};

// one bonded term: 48 bytes, fetched with three 16-byte loads
struct RepTerm {
    unsigned short idx[4];          // local bead indices (nbody used)
    unsigned short slot[4];         // scratch slot of the force on each of them
    float p[5];
    unsigned short tab, pad;        // table index (kinds 5 and 8)
    double r0;                      // kinds 0, 1, 3: equilibrium length (r - r0 is taken in double)
};
static_assert(sizeof(RepTerm) == 48, "RepTerm layout");
struct RepBonded {
    int n, npar;                    // terms, parameters per term
    const uint4* rec;               // [3 n] RepTerm records
};

struct RepTask { int kind, m0; };

struct RepParams {
    int nper, nrep, nlists, ntasks, nslots;   // nslots = nper * (largest number of bonded contributions to one bead, rounded up to 8)
    RepList lists[kRepMaxLists];
    RepBonded bonded[9];
    const float* tabs;              // 24 floats per table
    const RepTask* tasks;           // longest first; warp w runs tasks w, w + nwarps, ...
    const int* nslot_of_bead;       // [nper] bonded contributions to each bead rounded up to 8: its slots are p * nper + bead, p < that
    double dt;
    float a, b, c;                  // Langevin coefficients (StepParams)
    unsigned int seed_lo, seed_hi;
    const uint2* rep_seed;          // per-replica noise keys or null
};

// dynamic shared memory of one block: xf, vs (float4 n), scratch (float 3 nslots), xs (double 3n), origin (double 3), fpair,
// noise, xref (float 3n each), Verlet sub-list counts and entries (bytes)
__host__ __device__ inline size_t rep_smem_bytes(int nper, int nlists, int nslots) {
    return (size_t)nper * (16 + 16 + 24 + 12 * (size_t)nlists + 12 + 12) + 24 + 12 * (size_t)nslots + (size_t)nlists * nper * 4 * (1 + kRepNbMax);
}

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

__device__ __forceinline__ float rep_pair_force(const RepList& L, const float4 c, float r2) {
    if (L.form == SEG_PAIR_DH) {
        const float ir = fast_rsqrt(r2), r = r2 * ir;
        return c.x * __expf(-r * L.inv_debye) * (ir + L.inv_debye) * ir * ir;
    }
    return rep_pair_tz(L, c, r2);
}

__device__ __forceinline__ repmath::V3<float> rep_diff(const double* __restrict__ xs, int i, int j) {
    return {(float)(xs[3 * i] - xs[3 * j]), (float)(xs[3 * i + 1] - xs[3 * j + 1]), (float)(xs[3 * i + 2] - xs[3 * j + 2])};
}
__device__ __forceinline__ void rep_put(float* __restrict__ scratch, int nslots, int slot, const repmath::V3<float>& f) {
    scratch[slot] = f.x; scratch[nslots + slot] = f.y; scratch[2 * nslots + slot] = f.z;
}

// 32 terms of bonded kind KIND (compile-time: the switch in repmath::eval folds away)
template <int KIND>
__device__ __forceinline__ void rep_bonded_task(const RepParams& R, int m, const double* __restrict__ xs, float* __restrict__ scratch) {
    using namespace repmath;
    const RepBonded& B = R.bonded[KIND];
    if (m >= B.n) return;
    union { uint4 q[3]; RepTerm t; } u;
    u.q[0] = __ldg(&B.rec[3 * m]); u.q[1] = __ldg(&B.rec[3 * m + 1]); u.q[2] = __ldg(&B.rec[3 * m + 2]);
    const float* p = u.t.p;
    const float* tab = R.tabs + 24 * (int)u.t.tab;
    const int ns = R.nslots;
    if (KIND <= 3) {
        const int i = u.t.idx[0], j = u.t.idx[1];
        const double dx = xs[3 * i] - xs[3 * j], dy = xs[3 * i + 1] - xs[3 * j + 1], dz = xs[3 * i + 2] - xs[3 * j + 2];
        const V3<float> d = {(float)dx, (float)dy, (float)dz};
        float r, dq = 0.f;
        if (KIND == 0 || KIND == 1 || KIND == 3) {
            const double rd = sqrt(dx * dx + dy * dy + dz * dz);
            r = (float)rd; dq = (float)(rd - u.t.r0);
        } else r = sqrtf(dot(d, d));
        V3<float> fi;
        bond<float>(KIND, p, tab, d, r, dq, fi);
        rep_put(scratch, ns, u.t.slot[0], fi);
        rep_put(scratch, ns, u.t.slot[1], {-fi.x, -fi.y, -fi.z});
    } else if (KIND <= 5) {
        const int i = u.t.idx[0], j = u.t.idx[1], k = u.t.idx[2];
        V3<float> fi, fk;
        angle<float>(KIND, p, tab, rep_diff(xs, i, j), rep_diff(xs, k, j), fi, fk);
        rep_put(scratch, ns, u.t.slot[0], fi);
        rep_put(scratch, ns, u.t.slot[1], {-fi.x - fk.x, -fi.y - fk.y, -fi.z - fk.z});
        rep_put(scratch, ns, u.t.slot[2], fk);
    } else {
        const int i = u.t.idx[0], j = u.t.idx[1], k = u.t.idx[2], l = u.t.idx[3];
        V3<float> f1, f2, f3, f4;
        dihedral<float>(KIND, p, tab, rep_diff(xs, i, j), rep_diff(xs, j, k), rep_diff(xs, l, k), f1, f2, f3, f4);
        rep_put(scratch, ns, u.t.slot[0], f1);
        rep_put(scratch, ns, u.t.slot[1], f2);
        rep_put(scratch, ns, u.t.slot[2], f3);
        rep_put(scratch, ns, u.t.slot[3], f4);
    }
}

template <int NT>
__global__ void __launch_bounds__(NT, NT == 256 ? 5 : (NT == 512 ? 2 : 1)) replica_step_kernel(const __grid_constant__ RepParams R, double* __restrict__ x_g, float4* __restrict__ v_g,
                                                                 unsigned long long step0, int nsteps, int* __restrict__ flags) {
    extern __shared__ __align__(16) unsigned char rep_smem[];
    const int n = R.nper, rep = blockIdx.x, tid = threadIdx.x, lane = tid & 31, nthr = NT;
    float4* xf = reinterpret_cast<float4*>(rep_smem);              // [n] FP32 offsets from the replica's origin
    float4* vs = xf + n;                                           // [n] velocities, w = 1/m
    float* scratch = reinterpret_cast<float*>(vs + n);             // [3][nslots] bonded force contributions, 16-byte aligned
    double* xs = reinterpret_cast<double*>(scratch + 3 * R.nslots);   // [3n] master positions
    double* origin = xs + 3 * n;                                   // [3]
    float* fpair = reinterpret_cast<float*>(origin + 3);           // [nlists][3][n] pair forces
    float* noise = fpair + 3 * n * R.nlists;                       // [3][n] unit normals of the step
    float* xref = noise + 3 * n;                                   // [3][n] xf when the Verlet lists were built
    unsigned char* nbcnt = reinterpret_cast<unsigned char*>(xref + 3 * n);   // [nlists][n][4] sub-list lengths
    unsigned char* nb = nbcnt + R.nlists * n * 4;                  // [nlists][n][4][kRepNbMax] partners
    __shared__ int rebuild[2];   // [step parity]: scan all pairs and rebuild the Verlet lists in this step
    __shared__ int nb_overflow;  // a sub-list overflowed: scan in every step from now on
    __shared__ int bad;
    for (int k = tid; k < 3 * n; k += nthr) xs[k] = x_g[(size_t)3 * n * rep + k];
    for (int i = tid; i < n; i += nthr) vs[i] = v_g[(size_t)n * rep + i];
    for (int k = tid; k < 3 * R.nslots; k += nthr) scratch[k] = 0.f;   // (the padding slots of every bead stay zero)
    if (tid == 0) { bad = 0; rebuild[0] = 1; rebuild[1] = 0; nb_overflow = 0; }
    const float disp_max = 0.5f * kRepSkin * 0.57735f;   // per component: |d| <= sqrt(3) max_q |d_q|
    const uint2 key = R.rep_seed ? R.rep_seed[rep] : make_uint2(R.seed_lo, R.seed_hi);
    __syncthreads();
    // constants of this thread's first (component, bead) item of phase 2
    const int q0 = (tid >= n) + (tid >= 2 * n), i0 = min(tid - q0 * n, n - 1);
    const int cnt0 = __ldg(&R.nslot_of_bead[i0]);
    const float im0 = vs[i0].w, bm0 = R.b * im0, cs0 = R.c * sqrtf(fmaxf(im0, 0.f));
    for (int s = 0; s < nsteps; s++) {
        // ---- FP32 pair-kernel positions relative to the replica's centroid (moved every 256 steps) ----
        if ((s & 255) == 0) {
            if (tid < 3) {
                double c = 0;
                for (int i = 0; i < n; i++) c += xs[3 * i + tid];
                origin[tid] = c / n;
            }
            if (tid == 0) rebuild[s & 1] = 1;   // (the reference positions of the lists are in the old frame)
            __syncthreads();
            for (int i = tid; i < n; i += nthr)
                xf[i] = make_float4((float)(xs[3 * i] - origin[0]), (float)(xs[3 * i + 1] - origin[1]), (float)(xs[3 * i + 2] - origin[2]), 0.f);
            __syncthreads();
        }
        const unsigned long long step = step0 + (unsigned long long)s;
        const bool scan = rebuild[s & 1] != 0;   // (written before the last barrier; cleared after the next one)
        // ---- phase 1: warp tasks ----
        for (int t = tid >> 5; t < R.ntasks; t += NT / 32) {
            const RepTask task = R.tasks[t];
            if (task.kind >= kRepTaskNoise) {
                const int i = task.m0 + lane;
                if (i < n) {
                    float g0, g1, g2;
                    normals3(philox4x32_10(make_uint4((unsigned int)i, (unsigned int)step, (unsigned int)(step >> 32), 0u), key), g0, g1, g2);
                    noise[i] = g0; noise[n + i] = g1; noise[2 * n + i] = g2;
                }
            } else if (task.kind >= kRepTaskPair) {
                const int l = task.kind - kRepTaskPair;
                const RepList& L = R.lists[l];
                const int i = task.m0 + (lane >> 2), sub = lane & 3;
                float fx = 0.f, fy = 0.f, fz = 0.f;
                if (i < n) {
                    const float4 pi = xf[i];
                    const float4* __restrict__ crow = L.coef + (size_t)i * n;
                    unsigned char* my_nb = nb + ((size_t)(l * n + i) * 4 + sub) * kRepNbMax;
                    if (scan) {
                        // all beads: forces inside the cutoff, partners inside cutoff + skin into this lane's sub-list
                        int cnt = 0;
                        for (int j = sub; j < n; j += 4) {
                            const float4 pj = xf[j];
                            const float dx = __fsub_rn(pi.x, pj.x), dy = __fsub_rn(pi.y, pj.y), dz = __fsub_rn(pi.z, pj.z);
                            const float r2 = norm2(dx, dy, dz);
                            if (r2 < L.rl2 && j != i) {
                                const float4 c = __ldg(&crow[j]);
                                if (c.x == 0.f && c.y == 0.f) continue;   // excluded, not interacting
                                if (cnt < kRepNbMax) my_nb[cnt] = (unsigned char)j; else nb_overflow = 1;
                                cnt++;
                                if (r2 < L.rc2) {
                                    const float fr = rep_pair_force(L, c, r2);
                                    fx = fmaf(fr, dx, fx); fy = fmaf(fr, dy, fy); fz = fmaf(fr, dz, fz);
                                }
                            }
                        }
                        nbcnt[(l * n + i) * 4 + sub] = (unsigned char)min(cnt, kRepNbMax);
                        if (sub == 0 && l == 0) { xref[i] = pi.x; xref[n + i] = pi.y; xref[2 * n + i] = pi.z; }
                    } else {
                        const int cnt = nbcnt[(l * n + i) * 4 + sub];
                        for (int k = 0; k < cnt; k++) {
                            const int j = my_nb[k];
                            const float4 pj = xf[j];
                            const float4 c = __ldg(&crow[j]);
                            const float dx = __fsub_rn(pi.x, pj.x), dy = __fsub_rn(pi.y, pj.y), dz = __fsub_rn(pi.z, pj.z);
                            const float r2 = norm2(dx, dy, dz);
                            if (r2 < L.rc2) {
                                const float fr = rep_pair_force(L, c, r2);
                                fx = fmaf(fr, dx, fx); fy = fmaf(fr, dy, fy); fz = fmaf(fr, dz, fz);
                            }
                        }
                    }
                }
                fx += __shfl_xor_sync(0xffffffffu, fx, 1); fy += __shfl_xor_sync(0xffffffffu, fy, 1); fz += __shfl_xor_sync(0xffffffffu, fz, 1);
                fx += __shfl_xor_sync(0xffffffffu, fx, 2); fy += __shfl_xor_sync(0xffffffffu, fy, 2); fz += __shfl_xor_sync(0xffffffffu, fz, 2);
                if (i < n && sub == 0) { float* fp = fpair + 3 * n * l; fp[i] = fx; fp[n + i] = fy; fp[2 * n + i] = fz; }
            } else {
                const int m = task.m0 + lane;
                switch (task.kind) {
                    case 0: rep_bonded_task<0>(R, m, xs, scratch); break;
                    case 1: rep_bonded_task<1>(R, m, xs, scratch); break;
                    case 2: rep_bonded_task<2>(R, m, xs, scratch); break;
                    case 3: rep_bonded_task<3>(R, m, xs, scratch); break;
                    case 4: rep_bonded_task<4>(R, m, xs, scratch); break;
                    case 5: rep_bonded_task<5>(R, m, xs, scratch); break;
                    case 6: rep_bonded_task<6>(R, m, xs, scratch); break;
                    case 7: rep_bonded_task<7>(R, m, xs, scratch); break;
                    default: rep_bonded_task<8>(R, m, xs, scratch);
                }
            }
        }
        __syncthreads();
        if (tid == 0) rebuild[s & 1] = 0;
        if (tid == 1 && nb_overflow) rebuild[(s + 1) & 1] = 1;
        // ---- phase 2: one thread per (component, bead): fixed-order force sum, leap-frog Langevin update ----
        auto item = [&](const int t, const int q, const int i, const int cnt8, const float im, const float bm, const float cs) {
            float f = 0.f, f1 = 0.f, f2 = 0.f, f3 = 0.f;
#pragma unroll 1
            for (int l = 0; l < R.nlists; l++) f += fpair[3 * n * l + t];
            // the bead's scratch slots p * n + i, p < cnt8: its contributions, padded with zero slots to a multiple of 8
            const float* __restrict__ sc = scratch + q * R.nslots + i;
#pragma unroll 1
            for (int e = 0; e < cnt8; e += 8, sc += 8 * n) {
                f += sc[0]; f1 += sc[n]; f2 += sc[2 * n]; f3 += sc[3 * n];
                f += sc[4 * n]; f1 += sc[5 * n]; f2 += sc[6 * n]; f3 += sc[7 * n];
            }
            f = (f + f1) + (f2 + f3);
            if (im > 0.f) {
                float* vp = reinterpret_cast<float*>(vs + i) + q;
                const float vq = R.a * *vp + bm * f + cs * noise[t];
                *vp = vq;
                const double px = xs[3 * i + q] + (double)vq * R.dt;
                xs[3 * i + q] = px;
                const float xq = (float)(px - origin[q]);
                reinterpret_cast<float*>(xf + i)[q] = xq;
                if (fabsf(xq - xref[t]) > disp_max) rebuild[(s + 1) & 1] = 1;
                if (!(px == px) || !(f == f)) bad = 1;
            }
        };
        if (tid < 3 * n) item(tid, q0, i0, cnt0, im0, bm0, cs0);
        for (int t = tid + nthr; t < 3 * n; t += nthr) {   // (more than nthr / 3 beads)
            const int q = t / n, i = t - q * n;
            const float im = vs[i].w;
            item(t, q, i, __ldg(&R.nslot_of_bead[i]), im, R.b * im, R.c * sqrtf(fmaxf(im, 0.f)));
        }
        __syncthreads();
        if (bad) break;  // (block-uniform: read after the barrier)
    }
    for (int k = tid; k < 3 * n; k += nthr) x_g[(size_t)3 * n * rep + k] = xs[k];
    for (int i = tid; i < n; i += nthr) v_g[(size_t)n * rep + i] = vs[i];
    if (tid == 0 && bad) flags[FLAG_NAN] = 1;
}

}  // namespace ocm
