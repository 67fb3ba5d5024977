// kernels.cuh — sm_100a device code of the OpenCafeMol hot path.
//
// Kernel inventory (DESIGN.md has the roofline for each):
//   canon_kernel          master double positions -> 16-byte pair-kernel positions
//                         (32-bit fixed-point box fractions when periodic, FP32 otherwise)
//   bbox/grid/bin/scan/scatter/sortcell/build_list   Verlet-list construction over a cell grid
//   force_kernel          ONE fused launch per step: every pair list + every bonded kind
//   integrate_kernel      leap-frog Langevin (OpenMM LangevinIntegrator semantics), Philox4x32-10
//   energy_kernel, kinetic_kernel, extract_pairs_kernel     observer / introspection side
//
// Replaces, on the OpenMM side, the per-Force JIT kernels that the reference's generators trigger
// (one CustomNonbondedForce kernel per src/forcefield/*ForceFieldGenerator.cpp::generate()).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ocm {

constexpr int kBlock = 128;            // threads per block of the force kernel
constexpr double kFixScale = 4294967296.0;  // force accumulator: 64-bit fixed point, 2^32 per kJ/mol/nm
constexpr int kMaxMembers = 4;

enum SegKind : int { SEG_PAIR_TZ = 0, SEG_PAIR_DH = 1, SEG_BONDED = 2 };
enum FlagSlot : int { FLAG_VIOLATION = 0, FLAG_OVERFLOW = 1, FLAG_TICKET = 2, FLAG_NAN = 3, FLAG_MAXDISP = 4, FLAG_COUNT = 8 };

struct GridDev {
    int nc[3];
    int ncell;
    float lo[3];
    float inv_w;
    int bmin[3], bmax[3];  // ordered-int encoded bounding box accumulators (non-periodic)
};

struct PairListDev {
    int np, stride, ntypes, ncls, form, nmem, has_switch, smem_tab;
    const int* part;            // [np] bead index of each participant slot, ascending
    const unsigned char* type;  // [np]
    const signed char* cls;     // [np] interaction class or null
    unsigned int allow[32];     // allowed partner classes per class
    const int* excl_off;        // [np+1]
    const int* excl_idx;        // excluded partner bead ids, ascending per slot
    GridDev* grid;
    int max_cells;
    float wmin;                 // minimum cell width = (rc_list + skin)/2
    int* cell_of_slot;          // [np]
    int* cell_count;            // [max_cells+1]
    int* cell_start;            // [max_cells+1]
    int* cell_cursor;           // [max_cells]
    int* sorted_slots;          // [np]
    int* nbr_count;             // [np]
    unsigned int* nbr;          // [np*stride]  bead | type<<24
    float rlist2, rc2, rsw, inv_sw;
    // two-zone 12-6 family: E = c12/r^12 - c6/r^6 + c0 inside a2, outer coefficients beyond
    const float4* tabF;         // [nt*nt] {c12_in, c6_in, c12_out, c6_out} summed over members
    const float* tabA2;         // [nt*nt] zone boundary squared
    // screened Coulomb: E = qq (exp(-r/debye)/r - cc)
    const float* tabQ;          // [nt*nt]
    float inv_debye, dh_cc;
    // per-member energy tables (observer side)
    int mem_group[kMaxMembers];
    int mem_switch[kMaxMembers];
    float mem_rc2[kMaxMembers];
    const float4* memE[kMaxMembers];  // [2*nt*nt] {c12_in,c6_in,c0_in,a2},{c12_out,c6_out,c0_out,0}
};

struct BondedDev {
    int n, npar;
    const int* idx;       // nbody * n
    const double* par;    // npar * n
    const int* meta;      // group | periodic<<8 | table<<16
    const double* tabs;   // 24 doubles per table
};

struct WorkSeg {
    int block_begin, nblocks, kind, id, tpb;  // tpb: threads per bead (pair segments)
};

struct StepParams {
    int n;
    int periodic;
    double box[3], inv_box[3];
    float fscale[3];   // box / 2^32
    double dt;
    float a, b, c;     // Langevin coefficients
    unsigned int seed_lo, seed_hi;
    float half_skin2;
};

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int to_fix(double x, double inv_l) {
    double u = x * inv_l;
    u -= floor(u);
    return (unsigned int)(unsigned long long)(u * 4294967296.0);
}

template <bool PER>
__device__ __forceinline__ void delta(const uint4& pi, const uint4& pj, const float* fs, float& dx, float& dy, float& dz) {
    if (PER) {
        dx = __fmul_rn(__int2float_rn((int)(pi.x - pj.x)), fs[0]);
        dy = __fmul_rn(__int2float_rn((int)(pi.y - pj.y)), fs[1]);
        dz = __fmul_rn(__int2float_rn((int)(pi.z - pj.z)), fs[2]);
    } else {
        dx = __fsub_rn(__uint_as_float(pi.x), __uint_as_float(pj.x));
        dy = __fsub_rn(__uint_as_float(pi.y), __uint_as_float(pj.y));
        dz = __fsub_rn(__uint_as_float(pi.z), __uint_as_float(pj.z));
    }
}
// canonical squared distance: same operation order as the oracle's canon_r2 (bit-exact pair sets)
__device__ __forceinline__ float norm2(float dx, float dy, float dz) {
    return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}

__device__ __forceinline__ int enc_float(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7FFFFFFF; }
__device__ __forceinline__ float dec_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

__device__ __forceinline__ void add_force_fixed(unsigned long long* fb, int n, int i, float fx, float fy, float fz) {
    atomicAdd(&fb[i], (unsigned long long)(long long)(fx * 4294967296.0f));
    atomicAdd(&fb[n + i], (unsigned long long)(long long)(fy * 4294967296.0f));
    atomicAdd(&fb[2 * n + i], (unsigned long long)(long long)(fz * 4294967296.0f));
}
__device__ __forceinline__ void add_force_fixed_d(unsigned long long* fb, int n, int i, double fx, double fy, double fz) {
    atomicAdd(&fb[i], (unsigned long long)(long long)(fx * kFixScale));
    atomicAdd(&fb[n + i], (unsigned long long)(long long)(fy * kFixScale));
    atomicAdd(&fb[2 * n + i], (unsigned long long)(long long)(fz * kFixScale));
}

// ------------------------------------------------------------------------------------------------
// positions
__global__ void canon_kernel(StepParams P, const double* __restrict__ x, uint4* __restrict__ posq) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    double px = x[3 * i], py = x[3 * i + 1], pz = x[3 * i + 2];
    uint4 q;
    if (P.periodic) {
        q.x = to_fix(px, P.inv_box[0]); q.y = to_fix(py, P.inv_box[1]); q.z = to_fix(pz, P.inv_box[2]);
    } else {
        q.x = __float_as_uint((float)px); q.y = __float_as_uint((float)py); q.z = __float_as_uint((float)pz);
    }
    q.w = 0;
    posq[i] = q;
}

// ------------------------------------------------------------------------------------------------
// neighbour-list construction
__global__ void bbox_kernel(const PairListDev* lists, int li, const uint4* __restrict__ posq) {
    const PairListDev& L = lists[li];
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    if (a < L.np) {
        uint4 p = posq[L.part[a]];
        lo[0] = hi[0] = __uint_as_float(p.x); lo[1] = hi[1] = __uint_as_float(p.y); lo[2] = hi[2] = __uint_as_float(p.z);
    }
    for (int k = 0; k < 3; k++) {
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
        if ((threadIdx.x & 31) == 0 && lo[k] <= hi[k]) {
            atomicMin(&L.grid->bmin[k], enc_float(lo[k]));
            atomicMax(&L.grid->bmax[k], enc_float(hi[k]));
        }
    }
}

// one thread: derive the non-periodic grid from the bounding box, reset the accumulators
__global__ void grid_kernel(const PairListDev* lists, int li) {
    const PairListDev& L = lists[li];
    GridDev* g = L.grid;
    float w = L.wmin;
    float lo[3], ext[3];
    for (int k = 0; k < 3; k++) {
        lo[k] = dec_float(g->bmin[k]);
        float hi = dec_float(g->bmax[k]);
        ext[k] = hi - lo[k];
        if (!(ext[k] >= 0.f)) { ext[k] = 0.f; lo[k] = 0.f; }
    }
    int nc[3];
    for (int it = 0; it < 64; it++) {
        long long tot = 1;
        for (int k = 0; k < 3; k++) {
            int v = (int)floorf(ext[k] / w) + 1;
            nc[k] = max(1, min(v, 1024));
            tot *= nc[k];
        }
        if (tot <= (long long)L.max_cells) break;
        w *= 1.26f;
    }
    for (int k = 0; k < 3; k++) {
        g->nc[k] = nc[k]; g->lo[k] = lo[k];
        g->bmin[k] = 0x7FFFFFFF; g->bmax[k] = (int)0x80000000;
    }
    g->ncell = nc[0] * nc[1] * nc[2];
    g->inv_w = 1.0f / w;
}

template <bool PER>
__device__ __forceinline__ void cell_coords(const GridDev& g, const uint4& p, int& cx, int& cy, int& cz) {
    if (PER) {
        cx = (int)(((unsigned long long)p.x * (unsigned long long)g.nc[0]) >> 32);
        cy = (int)(((unsigned long long)p.y * (unsigned long long)g.nc[1]) >> 32);
        cz = (int)(((unsigned long long)p.z * (unsigned long long)g.nc[2]) >> 32);
    } else {
        cx = max(0, min(g.nc[0] - 1, (int)floorf(__fmul_rn(__fsub_rn(__uint_as_float(p.x), g.lo[0]), g.inv_w))));
        cy = max(0, min(g.nc[1] - 1, (int)floorf(__fmul_rn(__fsub_rn(__uint_as_float(p.y), g.lo[1]), g.inv_w))));
        cz = max(0, min(g.nc[2] - 1, (int)floorf(__fmul_rn(__fsub_rn(__uint_as_float(p.z), g.lo[2]), g.inv_w))));
    }
}

template <bool PER>
__global__ void bin_kernel(const PairListDev* lists, int li, const uint4* __restrict__ posq) {
    const PairListDev& L = lists[li];
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= L.np) return;
    const GridDev& g = *L.grid;
    int cx, cy, cz;
    cell_coords<PER>(g, posq[L.part[a]], cx, cy, cz);
    int c = (cz * g.nc[1] + cy) * g.nc[0] + cx;
    L.cell_of_slot[a] = c;
    atomicAdd(&L.cell_count[c], 1);
}

// single block: exclusive scan of cell_count[0..ncell) -> cell_start, cursor; zero the counts
__global__ void scan_kernel(const PairListDev* lists, int li) {
    const PairListDev& L = lists[li];
    const int ncell = L.grid->ncell;
    __shared__ int warp_sums[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int base = 0; base < ncell; base += blockDim.x) {
        int i = base + threadIdx.x;
        int v = (i < ncell) ? L.cell_count[i] : 0;
        int s = v;
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        if (lane == 31) warp_sums[wid] = s;
        __syncthreads();
        if (wid == 0) {
            int ws = (lane < nw) ? warp_sums[lane] : 0;
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, ws, o);
                if (lane >= o) ws += t;
            }
            warp_sums[lane] = ws;  // inclusive
        }
        __syncthreads();
        int excl = carry + (wid > 0 ? warp_sums[wid - 1] : 0) + s - v;
        if (i < ncell) { L.cell_start[i] = excl; L.cell_cursor[i] = excl; L.cell_count[i] = 0; }
        __syncthreads();
        if (threadIdx.x == 0) carry += warp_sums[nw - 1];
        __syncthreads();
    }
    if (threadIdx.x == 0) L.cell_start[ncell] = carry;
}

__global__ void scatter_kernel(const PairListDev* lists, int li) {
    const PairListDev& L = lists[li];
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= L.np) return;
    int pos = atomicAdd(&L.cell_cursor[L.cell_of_slot[a]], 1);
    L.sorted_slots[pos] = a;
}

// order the members of every cell by slot (= ascending bead index) so that list order, and with
// it the FP32 summation order of the force kernel, is reproducible run to run
__global__ void sortcell_kernel(const PairListDev* lists, int li) {
    const PairListDev& L = lists[li];
    const int ncell = L.grid->ncell;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ncell; c += gridDim.x * blockDim.x) {
        int s = L.cell_start[c], e = L.cell_start[c + 1];
        for (int i = s + 1; i < e; i++) {
            int v = L.sorted_slots[i], j = i - 1;
            while (j >= s && L.sorted_slots[j] > v) { L.sorted_slots[j + 1] = L.sorted_slots[j]; j--; }
            L.sorted_slots[j + 1] = v;
        }
    }
}

// one warp per participant: walk the 5x5 rows of neighbouring cells (x-runs are contiguous in the
// cell-sorted order), test the canonical distance, exclusions and interaction classes, compact with
// ballots.  count_only: just record the neighbour count (used to size the list).
template <bool PER>
__global__ void build_list_kernel(const PairListDev* lists, int li, StepParams P, const uint4* __restrict__ posq,
                                  int* __restrict__ flags, int count_only) {
    const PairListDev& L = lists[li];
    const int lane = threadIdx.x & 31;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= L.np) return;
    const GridDev& g = *L.grid;
    const int a = L.sorted_slots[w];
    const int i = L.part[a];
    const uint4 pi = posq[i];
    const int ci = L.cell_of_slot[a];
    const int ncx = g.nc[0], ncy = g.nc[1], ncz = g.nc[2];
    const int cx = ci % ncx, cy = (ci / ncx) % ncy, cz = ci / (ncx * ncy);
    const int e0 = L.excl_off[a], e1 = L.excl_off[a + 1];
    const unsigned int my_allow = L.cls ? L.allow[(int)L.cls[a]] : 0xffffffffu;
    const float rl2 = L.rlist2;
    unsigned int* row = L.nbr + (size_t)a * L.stride;
    int count = 0;

    int nz, ny, z0, y0;
    if (PER) { nz = min(5, ncz); z0 = (ncz >= 5) ? cz - 2 : 0; ny = min(5, ncy); y0 = (ncy >= 5) ? cy - 2 : 0; }
    else { z0 = max(cz - 2, 0); nz = min(cz + 2, ncz - 1) - z0 + 1; y0 = max(cy - 2, 0); ny = min(cy + 2, ncy - 1) - y0 + 1; }
    // x segments: at most two contiguous runs
    int xs[2], xe[2], nseg = 1;
    if (PER) {
        if (ncx < 5) { xs[0] = 0; xe[0] = ncx - 1; }
        else {
            int lo = cx - 2, hi = cx + 2;
            if (lo < 0) { xs[0] = 0; xe[0] = hi; xs[1] = lo + ncx; xe[1] = ncx - 1; nseg = 2; }
            else if (hi >= ncx) { xs[0] = 0; xe[0] = hi - ncx; xs[1] = lo; xe[1] = ncx - 1; nseg = 2; }
            else { xs[0] = lo; xe[0] = hi; }
        }
    } else { xs[0] = max(cx - 2, 0); xe[0] = min(cx + 2, ncx - 1); }

    for (int kz = 0; kz < nz; kz++) {
        int oz = z0 + kz;
        if (PER) { if (oz < 0) oz += ncz; else if (oz >= ncz) oz -= ncz; }
        for (int ky = 0; ky < ny; ky++) {
            int oy = y0 + ky;
            if (PER) { if (oy < 0) oy += ncy; else if (oy >= ncy) oy -= ncy; }
            const int rowbase = (oz * ncy + oy) * ncx;
            for (int sgm = 0; sgm < nseg; sgm++) {
                const int kb = L.cell_start[rowbase + xs[sgm]], ke = L.cell_start[rowbase + xe[sgm] + 1];
                for (int k0 = kb; k0 < ke; k0 += 32) {
                    const int k = k0 + lane;
                    bool ok = false;
                    unsigned int entry = 0;
                    if (k < ke) {
                        const int b = L.sorted_slots[k];
                        const int j = L.part[b];
                        if (j != i) {
                            float dx, dy, dz;
                            delta<PER>(pi, posq[j], P.fscale, dx, dy, dz);
                            if (norm2(dx, dy, dz) < rl2) {
                                ok = true;
                                if (L.cls) ok = (my_allow >> (int)L.cls[b]) & 1u;
                                for (int e = e0; ok && e < e1; e++) ok = L.excl_idx[e] != j;
                                entry = (unsigned int)j | ((unsigned int)L.type[b] << 24);
                            }
                        }
                    }
                    const unsigned int m = __ballot_sync(0xffffffffu, ok);
                    if (ok && !count_only) {
                        const int pos = count + __popc(m & ((1u << lane) - 1u));
                        if (pos < L.stride) row[pos] = entry;
                    }
                    count += __popc(m);
                }
            }
        }
    }
    if (lane == 0) {
        if (count_only) { L.nbr_count[a] = count; atomicMax(&flags[FLAG_MAXDISP + 1], count); }
        else {
            if (count > L.stride) { flags[FLAG_OVERFLOW] = 1; count = L.stride; }
            L.nbr_count[a] = count;
        }
    }
}

__global__ void save_ref_kernel(int n, const double* __restrict__ x, double* __restrict__ xref) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 3 * n) xref[i] = x[i];
}

// ------------------------------------------------------------------------------------------------
// pair forces.  T threads walk one participant's Verlet row (coalesced 4-byte entries), gather the
// 16-byte partner position, and reduce the FP32 force with shuffles; one fixed-point atomic triple
// per bead merges the result with the other lists and the bonded terms.
template <bool PER, int T, bool SMEM>
__device__ __forceinline__ void pair_tz_block(const PairListDev& L, int blk, const StepParams& P,
                                              const uint4* __restrict__ posq, unsigned long long* __restrict__ fbuf,
                                              float* smem) {
    const int nt = L.ntypes, nt2 = nt * nt;
    const float4* tabF;
    const float* tabA2;
    if (SMEM) {
        float4* sF = reinterpret_cast<float4*>(smem);
        float* sA = smem + 4 * nt2;
        for (int k = threadIdx.x; k < nt2; k += kBlock) { sF[k] = L.tabF[k]; sA[k] = L.tabA2[k]; }
        __syncthreads();
        tabF = sF; tabA2 = sA;
    } else { tabF = L.tabF; tabA2 = L.tabA2; }
    constexpr int BPB = kBlock / T;
    const int a = blk * BPB + threadIdx.x / T, s = threadIdx.x % T;
    float fx = 0.f, fy = 0.f, fz = 0.f;
    int i = 0;
    const bool active = a < L.np;
    if (active) {
        i = L.part[a];
        const uint4 pi = posq[i];
        const int trow = (int)L.type[a] * nt;
        const int n = L.nbr_count[a];
        const unsigned int* __restrict__ row = L.nbr + (size_t)a * L.stride;
        const float rc2 = L.rc2, rsw = L.rsw, inv_sw = L.inv_sw;
        const float rsw2 = rsw * rsw;
        const bool sw = L.has_switch != 0;
#pragma unroll 4
        for (int k = s; k < n; k += T) {
            const unsigned int e = __ldg(&row[k]);
            const uint4 pj = __ldg(&posq[e & 0xFFFFFFu]);
            float dx, dy, dz;
            delta<PER>(pi, pj, P.fscale, dx, dy, dz);
            const float r2 = norm2(dx, dy, dz);
            if (r2 < rc2) {
                const int ti = trow + (int)(e >> 24);
                const float4 c = tabF[ti];
                const bool inner = r2 < tabA2[ti];
                const float c12 = inner ? c.x : c.z, c6 = inner ? c.y : c.w;
                const float ir2 = __frcp_rn(r2);
                const float ir6 = ir2 * ir2 * ir2;
                float fr = (12.f * c12 * ir6 - 6.f * c6) * ir6 * ir2;  // -dE/dr / r
                if (sw && r2 > rsw2) {
                    const float r = sqrtf(r2);
                    const float t = (r - rsw) * inv_sw;
                    const float S = 1.f + t * t * t * (-10.f + t * (15.f - 6.f * t));
                    const float dS = t * t * (-30.f + t * (60.f - 30.f * t)) * inv_sw;
                    const float E = (c12 * ir6 - c6) * ir6;
                    fr = S * fr - E * dS / r;
                }
                fx = fmaf(fr, dx, fx); fy = fmaf(fr, dy, fy); fz = fmaf(fr, dz, fz);
            }
        }
    }
#pragma unroll
    for (int o = T / 2; o > 0; o >>= 1) {
        fx += __shfl_xor_sync(0xffffffffu, fx, o);
        fy += __shfl_xor_sync(0xffffffffu, fy, o);
        fz += __shfl_xor_sync(0xffffffffu, fz, o);
    }
    if (active && s == 0) add_force_fixed(fbuf, P.n, i, fx, fy, fz);
}

template <bool PER, int T>
__device__ __forceinline__ void pair_dh_block(const PairListDev& L, int blk, const StepParams& P,
                                              const uint4* __restrict__ posq, unsigned long long* __restrict__ fbuf) {
    constexpr int BPB = kBlock / T;
    const int a = blk * BPB + threadIdx.x / T, s = threadIdx.x % T;
    float fx = 0.f, fy = 0.f, fz = 0.f;
    int i = 0;
    const bool active = a < L.np;
    if (active) {
        i = L.part[a];
        const uint4 pi = posq[i];
        const int nt = L.ntypes;
        const int trow = (int)L.type[a] * nt;
        const int n = L.nbr_count[a];
        const unsigned int* __restrict__ row = L.nbr + (size_t)a * L.stride;
        const float rc2 = L.rc2, ik = L.inv_debye;
        for (int k = s; k < n; k += T) {
            const unsigned int e = __ldg(&row[k]);
            const uint4 pj = __ldg(&posq[e & 0xFFFFFFu]);
            float dx, dy, dz;
            delta<PER>(pi, pj, P.fscale, dx, dy, dz);
            const float r2 = norm2(dx, dy, dz);
            if (r2 < rc2) {
                const float qq = __ldg(&L.tabQ[trow + (int)(e >> 24)]);
                const float ir = rsqrtf(r2);
                const float r = r2 * ir;
                const float ex = __expf(-r * ik);
                const float fr = qq * ex * (ir + ik) * ir * ir;
                fx = fmaf(fr, dx, fx); fy = fmaf(fr, dy, fy); fz = fmaf(fr, dz, fz);
            }
        }
    }
#pragma unroll
    for (int o = T / 2; o > 0; o >>= 1) {
        fx += __shfl_xor_sync(0xffffffffu, fx, o);
        fy += __shfl_xor_sync(0xffffffffu, fy, o);
        fz += __shfl_xor_sync(0xffffffffu, fz, o);
    }
    if (active && s == 0) add_force_fixed(fbuf, P.n, i, fx, fy, fz);
}

// ------------------------------------------------------------------------------------------------
// bonded terms (double precision on the master positions; one thread per term)
struct D3 { double x, y, z; };
__device__ __forceinline__ D3 d3sub(const double* __restrict__ x, int i, int j, bool per, const StepParams& P) {
    D3 d = {x[3 * i] - x[3 * j], x[3 * i + 1] - x[3 * j + 1], x[3 * i + 2] - x[3 * j + 2]};
    if (per) {
        d.x -= P.box[0] * rint(d.x * P.inv_box[0]);
        d.y -= P.box[1] * rint(d.y * P.inv_box[1]);
        d.z -= P.box[2] * rint(d.z * P.inv_box[2]);
    }
    return d;
}
__device__ __forceinline__ double d3dot(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ D3 d3cross(D3 a, D3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

__device__ __forceinline__ void spline_eval_dev(const double* y, const double* y2, double tmin, double tmax, double t,
                                                double& val, double& der) {
    const double h = (tmax - tmin) / 9.0;
    int k = (int)floor((t - tmin) / h);
    k = max(0, min(8, k));
    const double a = (tmin + (k + 1) * h - t) / h, b = 1.0 - a;
    val = a * y[k] + b * y[k + 1] + ((a * a * a - a) * y2[k] + (b * b * b - b) * y2[k + 1]) * h * h / 6.0;
    der = (y[k + 1] - y[k]) / h + ((1.0 - 3.0 * a * a) * y2[k] + (3.0 * b * b - 1.0) * y2[k + 1]) * h / 6.0;
}

// E and dE/dq (q = r, theta or phi) of one bonded term; formulas cite the reference generators
__device__ __forceinline__ void bonded_eval(int kind, const double* __restrict__ p, const double* __restrict__ tab, double q,
                                            double& E, double& dE) {
    switch (kind) {
        case 0: {  // HarmonicBond  p = r0, k          HarmonicBondForceFieldGenerator.cpp:3-14
            const double d = q - p[0];
            E = 0.5 * p[1] * d * d; dE = p[1] * d;
        } break;
        case 1: {  // GaussianBond  p = k, r0, sigma   GaussianBondForceFieldGenerator.cpp:5-9
            const double d = q - p[1], is2 = 1.0 / (p[2] * p[2]);
            E = p[0] * exp(-0.5 * d * d * is2); dE = -E * d * is2;
        } break;
        case 2: {  // GoContact     p = k, r0          GoContactForceFieldGenerator.cpp:5-7
            const double s = p[1] / q, s2 = s * s, s4 = s2 * s2, s10 = s4 * s4 * s2, s12 = s10 * s2;
            E = p[0] * (5.0 * s12 - 6.0 * s10); dE = p[0] * 60.0 * (s10 - s12) / q;
        } break;
        case 3: {  // 3SPN2Bond     p = k2, k4, r0     ThreeSPN2BondForceFieldGenerator.cpp:5-10
            const double d = q - p[2], d2 = d * d;
            E = p[0] * d2 + p[1] * d2 * d2; dE = 2.0 * p[0] * d + 4.0 * p[1] * d2 * d;
        } break;
        case 4: {  // HarmonicAngle p = theta0, k      HarmonicAngleForceFieldGenerator.cpp:5-12
            const double d = q - p[0];
            E = 0.5 * p[1] * d * d; dE = p[1] * d;
        } break;
        case 5: {  // FlexibleLocalAngle p = k         FlexibleLocalAngleForceFieldGenerator.cpp:21-25
            const double tmin = tab[20], tmax = tab[21];
            double val = 0.0, der = 0.0;
            if (q >= tmin && q <= tmax) spline_eval_dev(tab, tab + 10, tmin, tmax, q, val, der);
            double e = val - tab[22], de = der;
            if (tmin - q >= 0.0) { e += 30.0 * (tmin - q) + tab[0]; de -= 30.0; }
            if (q - tmax >= 0.0) { e += 30.0 * (q - tmax) + tab[9]; de += 30.0; }
            E = p[0] * e; dE = p[0] * de;
        } break;
        case 6: {  // GaussianDihedral p = k, phi0, sigma   GaussianDihedralForceFieldGenerator.cpp:5-12
            const double PI = 3.14159265358979323846;
            const double dt = q - p[1];
            const double d = dt - floor((dt + PI) / (2.0 * PI)) * 2.0 * PI;
            const double is2 = 1.0 / (p[2] * p[2]);
            E = p[0] * exp(-0.5 * d * d * is2); dE = -E * d * is2;
        } break;
        case 7: {  // CosineDihedral p = k, t0, n      CosineDihedralForceFieldGenerator.cpp:7-12
            double sn, cs;
            sincos(p[2] * (q - p[1]), &sn, &cs);
            E = p[0] * (1.0 - cs); dE = p[0] * p[2] * sn;
        } break;
        default: {  // FlexibleLocalDihedral p = k, tab = c,a1,b1,a2,b2,a3,b3,emin   …cpp:28-37
            double s1, c1;
            sincos(q, &s1, &c1);
            const double s2 = 2.0 * s1 * c1, c2 = c1 * c1 - s1 * s1;
            const double s3 = s1 * c2 + c1 * s2, c3 = c1 * c2 - s1 * s2;
            const double e = tab[0] - tab[7] + tab[1] * c1 + tab[2] * s1 + tab[3] * c2 + tab[4] * s2 + tab[5] * c3 + tab[6] * s3;
            const double de = -tab[1] * s1 + tab[2] * c1 + 2.0 * (-tab[3] * s2 + tab[4] * c2) + 3.0 * (-tab[5] * s3 + tab[6] * c3);
            E = p[0] * e; dE = p[0] * de;
        }
    }
}

// computes one bonded term; accumulates forces (fbuf != null) and returns the energy
__device__ __forceinline__ double bonded_term(int kind, const BondedDev& B, int m, const StepParams& P,
                                              const double* __restrict__ x, unsigned long long* __restrict__ fbuf, int* group) {
    const int meta = B.meta[m];
    *group = meta & 0xff;
    const bool per = (meta >> 8) & 1;
    const double* p = B.par + (size_t)m * B.npar;
    const double* tab = B.tabs + (size_t)(meta >> 16) * 24;
    double E, dE;
    if (kind <= 3) {
        const int i = B.idx[2 * m], j = B.idx[2 * m + 1];
        const D3 d = d3sub(x, i, j, per, P);
        const double r = sqrt(d3dot(d, d));
        bonded_eval(kind, p, tab, r, E, dE);
        if (fbuf) {
            const double s = -dE / r;
            add_force_fixed_d(fbuf, P.n, i, s * d.x, s * d.y, s * d.z);
            add_force_fixed_d(fbuf, P.n, j, -s * d.x, -s * d.y, -s * d.z);
        }
    } else if (kind <= 5) {
        const int i = B.idx[3 * m], j = B.idx[3 * m + 1], k = B.idx[3 * m + 2];
        const D3 a = d3sub(x, i, j, per, P), b = d3sub(x, k, j, per, P);
        const double la = sqrt(d3dot(a, a)), lb = sqrt(d3dot(b, b));
        double cs = d3dot(a, b) / (la * lb);
        cs = fmax(-1.0, fmin(1.0, cs));
        bonded_eval(kind, p, tab, acos(cs), E, dE);
        if (fbuf) {
            const double sn = sqrt(fmax(1.0 - cs * cs, 1e-24));
            const double ci = dE / (sn * la), ck = dE / (sn * lb);
            const double ila = 1.0 / la, ilb = 1.0 / lb;
            const double fix = ci * (b.x * ilb - cs * a.x * ila), fiy = ci * (b.y * ilb - cs * a.y * ila), fiz = ci * (b.z * ilb - cs * a.z * ila);
            const double fkx = ck * (a.x * ila - cs * b.x * ilb), fky = ck * (a.y * ila - cs * b.y * ilb), fkz = ck * (a.z * ila - cs * b.z * ilb);
            add_force_fixed_d(fbuf, P.n, i, fix, fiy, fiz);
            add_force_fixed_d(fbuf, P.n, k, fkx, fky, fkz);
            add_force_fixed_d(fbuf, P.n, j, -fix - fkx, -fiy - fky, -fiz - fkz);
        }
    } else {
        const int i = B.idx[4 * m], j = B.idx[4 * m + 1], k = B.idx[4 * m + 2], l = B.idx[4 * m + 3];
        // F = p1-p2, G = p2-p3, H = p4-p3; OpenMM sign rule
        const D3 Fv = d3sub(x, i, j, per, P), G = d3sub(x, j, k, per, P), H = d3sub(x, l, k, per, P);
        const D3 A = d3cross(Fv, G), Bv = d3cross(H, G);
        const double A2 = d3dot(A, A), B2 = d3dot(Bv, Bv), lG = sqrt(d3dot(G, G));
        double cs = d3dot(A, Bv) / sqrt(A2 * B2);
        cs = fmax(-1.0, fmin(1.0, cs));
        double phi = acos(cs);
        if (d3dot(Fv, Bv) > 0.0) phi = -phi;
        bonded_eval(kind, p, tab, phi, E, dE);
        if (fbuf) {
            const double FG = d3dot(Fv, G), HG = d3dot(H, G);
            const double ca = lG / A2, cb = lG / B2, ea = FG / (A2 * lG), eb = HG / (B2 * lG);
            // dphi/dp1 = -ca A ; dphi/dp4 = cb B ; dphi/dp2 = (ca+ea) A - eb B ; dphi/dp3 = -(cb-eb) B - ea A
            const double g1x = -ca * A.x, g1y = -ca * A.y, g1z = -ca * A.z;
            const double g4x = cb * Bv.x, g4y = cb * Bv.y, g4z = cb * Bv.z;
            const double g2x = (ca + ea) * A.x - eb * Bv.x, g2y = (ca + ea) * A.y - eb * Bv.y, g2z = (ca + ea) * A.z - eb * Bv.z;
            const double g3x = -(cb - eb) * Bv.x - ea * A.x, g3y = -(cb - eb) * Bv.y - ea * A.y, g3z = -(cb - eb) * Bv.z - ea * A.z;
            add_force_fixed_d(fbuf, P.n, i, -dE * g1x, -dE * g1y, -dE * g1z);
            add_force_fixed_d(fbuf, P.n, j, -dE * g2x, -dE * g2y, -dE * g2z);
            add_force_fixed_d(fbuf, P.n, k, -dE * g3x, -dE * g3y, -dE * g3z);
            add_force_fixed_d(fbuf, P.n, l, -dE * g4x, -dE * g4y, -dE * g4z);
        }
    }
    return E;
}

// ------------------------------------------------------------------------------------------------
// the fused per-step force launch
template <bool PER>
__global__ void __launch_bounds__(kBlock) force_kernel(const WorkSeg* __restrict__ segs, int nsegs,
                                                       const PairListDev* __restrict__ lists,
                                                       const BondedDev* __restrict__ bonded, StepParams P,
                                                       const uint4* __restrict__ posq, const double* __restrict__ x,
                                                       unsigned long long* __restrict__ fbuf) {
    extern __shared__ __align__(16) float smem[];
    int s = 0;
    while (s + 1 < nsegs && (int)blockIdx.x >= segs[s + 1].block_begin) s++;
    const WorkSeg sg = segs[s];
    const int blk = blockIdx.x - sg.block_begin;
    if (sg.kind == SEG_PAIR_TZ) {
        const PairListDev& L = lists[sg.id];
        if (L.smem_tab) {
            switch (sg.tpb) {
                case 4: pair_tz_block<PER, 4, true>(L, blk, P, posq, fbuf, smem); break;
                case 8: pair_tz_block<PER, 8, true>(L, blk, P, posq, fbuf, smem); break;
                case 16: pair_tz_block<PER, 16, true>(L, blk, P, posq, fbuf, smem); break;
                default: pair_tz_block<PER, 32, true>(L, blk, P, posq, fbuf, smem);
            }
        } else {
            switch (sg.tpb) {
                case 4: pair_tz_block<PER, 4, false>(L, blk, P, posq, fbuf, smem); break;
                case 8: pair_tz_block<PER, 8, false>(L, blk, P, posq, fbuf, smem); break;
                case 16: pair_tz_block<PER, 16, false>(L, blk, P, posq, fbuf, smem); break;
                default: pair_tz_block<PER, 32, false>(L, blk, P, posq, fbuf, smem);
            }
        }
    } else if (sg.kind == SEG_PAIR_DH) {
        const PairListDev& L = lists[sg.id];
        switch (sg.tpb) {
            case 4: pair_dh_block<PER, 4>(L, blk, P, posq, fbuf); break;
            case 8: pair_dh_block<PER, 8>(L, blk, P, posq, fbuf); break;
            case 16: pair_dh_block<PER, 16>(L, blk, P, posq, fbuf); break;
            default: pair_dh_block<PER, 32>(L, blk, P, posq, fbuf);
        }
    } else {
        const BondedDev& B = bonded[sg.id];
        const int m = blk * kBlock + threadIdx.x;
        if (m < B.n) { int g; bonded_term(sg.id, B, m, P, x, fbuf, &g); }
    }
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based noise: key = seed, counter = (bead, step_lo, step_hi, stream)
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const unsigned int hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const unsigned int hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
    }
    return c;
}
__device__ __forceinline__ void normals3(uint4 o, float& g0, float& g1, float& g2) {
    const float s = 2.3283064365386963e-10f;  // 2^-32
    const float u0 = ((float)o.x + 0.5f) * s, u1 = ((float)o.y + 0.5f) * s;
    const float u2 = ((float)o.z + 0.5f) * s, u3 = ((float)o.w + 0.5f) * s;
    const float r0 = sqrtf(-2.f * logf(u0)), r1 = sqrtf(-2.f * logf(u2));
    float sn, cs;
    sincospif(2.f * u1, &sn, &cs);
    g0 = r0 * cs; g1 = r0 * sn;
    g2 = r1 * cospif(2.f * u3);
}

// leap-frog Langevin step, OpenMM LangevinIntegrator semantics (reference
// src/IntegratorGenerator.cpp:5-14): v <- a v + b F/m + c sqrt(1/m) xi ; x <- x + v dt.
// Also: convert + clear the fixed-point force accumulator, refresh the pair-kernel positions, and
// flag a Verlet-buffer violation (displacement from the list's reference positions > skin/2)
// unless this is the last step before a rebuild.
__global__ void integrate_kernel(StepParams P, double* __restrict__ x, float4* __restrict__ v,
                                 unsigned long long* __restrict__ fbuf, uint4* __restrict__ posq,
                                 const double* __restrict__ xref, unsigned long long* __restrict__ step_base,
                                 int step_offset, int window, int last_in_window, int* __restrict__ flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long step = *step_base + (unsigned long long)step_offset;
    float d2 = 0.f;
    if (i < P.n) {
        const int n = P.n;
        const float inv32 = 2.3283064365386963e-10f;
        const float fx = (float)(long long)fbuf[i] * inv32, fy = (float)(long long)fbuf[n + i] * inv32,
                    fz = (float)(long long)fbuf[2 * n + i] * inv32;
        fbuf[i] = 0ull; fbuf[n + i] = 0ull; fbuf[2 * n + i] = 0ull;
        float4 vi = v[i];
        const float im = vi.w;
        if (im > 0.f) {
            float g0, g1, g2;
            normals3(philox4x32_10(make_uint4((unsigned int)i, (unsigned int)step, (unsigned int)(step >> 32), 0u),
                                   make_uint2(P.seed_lo, P.seed_hi)), g0, g1, g2);
            const float cs = P.c * sqrtf(im), bm = P.b * im;
            vi.x = P.a * vi.x + bm * fx + cs * g0;
            vi.y = P.a * vi.y + bm * fy + cs * g1;
            vi.z = P.a * vi.z + bm * fz + cs * g2;
            v[i] = vi;
            const double px = x[3 * i] + (double)vi.x * P.dt, py = x[3 * i + 1] + (double)vi.y * P.dt,
                         pz = x[3 * i + 2] + (double)vi.z * P.dt;
            x[3 * i] = px; x[3 * i + 1] = py; x[3 * i + 2] = pz;
            uint4 q;
            if (P.periodic) { q.x = to_fix(px, P.inv_box[0]); q.y = to_fix(py, P.inv_box[1]); q.z = to_fix(pz, P.inv_box[2]); }
            else { q.x = __float_as_uint((float)px); q.y = __float_as_uint((float)py); q.z = __float_as_uint((float)pz); }
            q.w = 0;
            posq[i] = q;
            const float ex = (float)(px - xref[3 * i]), ey = (float)(py - xref[3 * i + 1]), ez = (float)(pz - xref[3 * i + 2]);
            d2 = ex * ex + ey * ey + ez * ez;
            if (!(d2 == d2) || !(fx == fx)) flags[FLAG_NAN] = 1;
        }
    }
    if (!last_in_window) {
        if (d2 > P.half_skin2) flags[FLAG_VIOLATION] = 1;
    } else {
        // record the largest displacement of the window (drives the adaptive rebuild interval)
        for (int o = 16; o > 0; o >>= 1) d2 = fmaxf(d2, __shfl_xor_sync(0xffffffffu, d2, o));
        if ((threadIdx.x & 31) == 0) atomicMax(&flags[FLAG_MAXDISP], __float_as_int(d2));
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            const int t = atomicAdd(&flags[FLAG_TICKET], 1);
            if (t == (int)gridDim.x - 1) { *step_base += (unsigned long long)window; flags[FLAG_TICKET] = 0; }
        }
    }
}

// v = sqrt(kT/m) N(0,1), noise stream 1 (Context::setVelocitiesToTemperature)
__global__ void init_velocity_kernel(int n, float kT, unsigned int seed_lo, unsigned int seed_hi, float4* __restrict__ v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 vi = v[i];
    float g0, g1, g2;
    normals3(philox4x32_10(make_uint4((unsigned int)i, 0u, 0u, 1u), make_uint2(seed_lo, seed_hi)), g0, g1, g2);
    const float s = vi.w > 0.f ? sqrtf(kT * vi.w) : 0.f;
    vi.x = s * g0; vi.y = s * g1; vi.z = s * g2;
    v[i] = vi;
}

// ------------------------------------------------------------------------------------------------
// observer side: per-group energies (double accumulation), kinetic energy, forces read-out
__device__ __forceinline__ void block_add(double val, double* target) {
    for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
    __shared__ double part[kBlock / 32];
    __syncthreads();
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = val;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0;
        for (int k = 0; k < kBlock / 32; k++) s += part[k];
        if (s != 0.0) atomicAdd(target, s);
    }
}

template <bool PER>
__global__ void __launch_bounds__(kBlock) energy_kernel(const WorkSeg* __restrict__ segs, int nsegs,
                                                        const PairListDev* __restrict__ lists,
                                                        const BondedDev* __restrict__ bonded, StepParams P,
                                                        const uint4* __restrict__ posq, const double* __restrict__ x,
                                                        double* __restrict__ egroup) {
    int s = 0;
    while (s + 1 < nsegs && (int)blockIdx.x >= segs[s + 1].block_begin) s++;
    const WorkSeg sg = segs[s];
    const int blk = blockIdx.x - sg.block_begin;
    if (sg.kind == SEG_BONDED) {
        const BondedDev& B = bonded[sg.id];
        const int m = blk * kBlock + threadIdx.x;
        double E = 0.0;
        int g = 0;
        if (m < B.n) E = bonded_term(sg.id, B, m, P, x, nullptr, &g);
        // terms of one kind may belong to different groups: accumulate per distinct group
        for (int gg = 0; gg < 32; gg++) {
            const unsigned int any = __syncthreads_or(m < B.n && g == gg);
            if (any) block_add((m < B.n && g == gg) ? E : 0.0, &egroup[gg]);
        }
        return;
    }
    const PairListDev& L = lists[sg.id];
    const int T = sg.tpb, BPB = kBlock / T;
    const int a = blk * BPB + threadIdx.x / T, sl = threadIdx.x % T;
    double Em[kMaxMembers] = {0, 0, 0, 0};
    if (a < L.np) {
        const int i = L.part[a];
        const uint4 pi = posq[i];
        const int nt = L.ntypes, nt2 = nt * nt;
        const int trow = (int)L.type[a] * nt;
        const int n = L.nbr_count[a];
        const unsigned int* row = L.nbr + (size_t)a * L.stride;
        for (int k = sl; k < n; k += T) {
            const unsigned int e = row[k];
            float dx, dy, dz;
            delta<PER>(pi, posq[e & 0xFFFFFFu], P.fscale, dx, dy, dz);
            const float r2 = norm2(dx, dy, dz);
            const int ti = trow + (int)(e >> 24);
            if (sg.kind == SEG_PAIR_DH) {
                if (r2 < L.rc2) {
                    const float ir = rsqrtf(r2), r = r2 * ir;
                    Em[0] += (double)(L.tabQ[ti] * (__expf(-r * L.inv_debye) * ir - L.dh_cc));
                }
            } else {
                for (int m = 0; m < L.nmem; m++) {
                    if (!(r2 < L.mem_rc2[m])) continue;
                    const float4 ci = L.memE[m][2 * ti], co = L.memE[m][2 * ti + 1];
                    const bool inner = r2 < ci.w;
                    const float c12 = inner ? ci.x : co.x, c6 = inner ? ci.y : co.y, c0 = inner ? ci.z : co.z;
                    const float ir2 = 1.0f / r2, ir6 = ir2 * ir2 * ir2;
                    float E = (c12 * ir6 - c6) * ir6 + c0;
                    if (L.mem_switch[m] && r2 > L.rsw * L.rsw) {
                        const float t = (sqrtf(r2) - L.rsw) * L.inv_sw;
                        E *= 1.f + t * t * t * (-10.f + t * (15.f - 6.f * t));
                    }
                    Em[m] += (double)E;
                }
            }
        }
    }
    const int nm = (sg.kind == SEG_PAIR_DH) ? 1 : L.nmem;
    for (int m = 0; m < nm; m++) block_add(0.5 * Em[m], &egroup[L.mem_group[m]]);  // full list: each pair twice
}

__global__ void kinetic_kernel(StepParams P, const float4* __restrict__ v, const unsigned long long* __restrict__ fbuf,
                               double* __restrict__ ke) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double e = 0.0;
    if (i < P.n) {
        const float4 vi = v[i];
        if (vi.w > 0.f) {
            const double im = vi.w, h = 0.5 * P.dt * im, inv = 1.0 / kFixScale;
            const double vx = vi.x + h * (double)(long long)fbuf[i] * inv, vy = vi.y + h * (double)(long long)fbuf[P.n + i] * inv,
                         vz = vi.z + h * (double)(long long)fbuf[2 * P.n + i] * inv;
            e = 0.5 / im * (vx * vx + vy * vy + vz * vz);
        }
    }
    block_add(e, ke);
}

// pairs (i<j) of list `li` inside rc2_term, appended unordered (host sorts)
template <bool PER>
__global__ void extract_pairs_kernel(const PairListDev* lists, int li, StepParams P, const uint4* __restrict__ posq,
                                     float rc2_term, int* __restrict__ out, long long capacity,
                                     unsigned long long* __restrict__ counter) {
    const PairListDev& L = lists[li];
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= L.np) return;
    const int i = L.part[a];
    const uint4 pi = posq[i];
    const int n = L.nbr_count[a];
    const unsigned int* row = L.nbr + (size_t)a * L.stride;
    for (int k = 0; k < n; k++) {
        const int j = (int)(row[k] & 0xFFFFFFu);
        if (j <= i) continue;
        float dx, dy, dz;
        delta<PER>(pi, posq[j], P.fscale, dx, dy, dz);
        if (norm2(dx, dy, dz) < rc2_term) {
            const unsigned long long pos = atomicAdd(counter, 1ull);
            if (out && (long long)pos < capacity) { out[2 * pos] = i; out[2 * pos + 1] = j; }
        }
    }
}

// register-resident FMA chain: 8 independent accumulators per thread, 2 flop per FMA
__global__ void fma_peak_kernel(float* out, int iters, float a, float b) {
    float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

}  // namespace ocm
