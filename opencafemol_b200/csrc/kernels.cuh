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
#ifndef OCM_FORCE_MINBLOCKS
#define OCM_FORCE_MINBLOCKS 6            // resident blocks per SM the force kernel is compiled for (register cap 80: no spills;
#endif                                   // measured on B200, C2 bench state: 8 -> 41.4 us, 6 -> 38.6 us, 5 -> 38.9 us, 4 -> 41.8 us per launch)
constexpr double kFixScale = 4294967296.0;  // force accumulator: 64-bit fixed point, 2^32 per kJ/mol/nm
constexpr int kMaxMembers = 4;

enum SegKind : int { SEG_PAIR_TZ = 0, SEG_PAIR_DH = 1, SEG_BONDED = 2, SEG_DA = 3, SEG_PAIR_WELL = 4 };
enum FlagSlot : int { FLAG_VIOLATION = 0, FLAG_OVERFLOW = 1, FLAG_TICKET = 2, FLAG_NAN = 3, FLAG_MAXDISP = 4, FLAG_DA_OVERFLOW = 6,
                       // slab decomposition: halo message overflow, a bonded term reaching beyond the halo, largest message (entries)
                       FLAG_DD_OVERFLOW = 7, FLAG_DD_REACH = 8, FLAG_DD_MAXSEND = 9,
                       // unlimited boundary: a bead left the range the virtual box was laid out for
                       FLAG_VBOX = 10, FLAG_COUNT = 12 };

struct GridDev {
    int nc[3];
    int ncell;
    float lo[3];
    float inv_w;
    int bmin[3], bmax[3];  // ordered-int encoded bounding box accumulators (non-periodic)
};

struct PairListDev {
    int np, stride, ntypes, ncls, form, nmem, has_switch, smem_tab;
    int nrep, nper;             // replica batch: cells are keyed (replica, local cell); pairs never cross replicas
    const unsigned char* role;  // slab decomposition: per bead 0 = not on this rank, 1 = owned, 2 = halo; null = single domain
    const int* part;            // [np] bead index of each participant slot, ascending
    const unsigned char* type;  // [np]
    const signed char* cls;     // [np] interaction class or null
    unsigned int allow[32];     // allowed partner classes per class
    const int* excl_off;        // [np+1]
    const int* excl_idx;        // excluded partner bead ids, ascending per slot
    GridDev* grid;
    int max_cells;
    int row_capacity;           // rows covered by the force launch (slab decomposition: below np)
    float wmin;                 // minimum cell width = (rc_list + skin)/2
    int* cell_of_slot;          // [np]
    int* cell_count;            // [max_cells+1]
    int* cell_start;            // [max_cells+1]
    int* cell_cursor;           // [max_cells]
    int* scan_tmp;              // [max_cells / 8192 + 2] block totals of the three-launch scan of large grids
    int* tmp_slots;             // [np] cell-grouped slots, arbitrary order inside a cell (scatter output)
    int* sorted_slots;          // [np] rank k -> slot; cells in grid order, ascending slot inside a cell
    int* rank_of_bead;          // [N]  bead -> rank k in this list's cell-sorted order, -1 = non participant
    uint4* spos;                // [np] cell-sorted pair-kernel positions; w = bead | type<<24 (two-zone lists) or the
                                //      bits of the scaled charge q sqrt(K) (screened-Coulomb lists)
    int* bead_of_rank;          // [np] rank k -> bead
    const float* qscaled;       // [np] per slot q sqrt(K) (screened-Coulomb lists)
    int* nbr_count;             // [np] row length of rank k
    int* nbr_inner;             // [np] leading entries of the row that are guaranteed inside the inner zone (see zone_a2)
    float zone_a2, zone_b2, zone_c2;  // build-time distance zones: rows are ordered [A | A' | B | C]
    unsigned int* nbr;          // [np*stride]  row of rank k: partner rank | partner type<<24
    float rlist2, rc2, rsw, inv_sw;
    // two-zone 12-6 family: E = c12/r^12 - c6/r^6 + c0 inside a2, outer coefficients beyond
    const float4* tabF;         // [nt*nt] {12 c12, 6 c6, a2, w}: -dE/dr / r = zone * (12 c12/r^12 - 6 c6/r^6)/r^2
    float k_in;                 // zone factor = r^2 < a2 ? k_in : w
    // screened Coulomb: E = q_i q_j K (exp(-r/debye)/r - cc); the charges ride in spos.w
    float inv_debye, dh_cc;
    // per-member energy tables (observer side)
    int mem_group[kMaxMembers];
    int mem_switch[kMaxMembers];
    float mem_rc2[kMaxMembers];
    const float4* memE[kMaxMembers];  // [2*nt*nt] {c12_in,c6_in,c0_in,a2},{c12_out,c6_out,c0_out,0}
    // ---- cluster-pair ("tile") form of the list (tile_kernels.cuh); Verlet rows are not built when tiled != 0 ----
    int tiled, halo_split;
    const int* cell_rank;       // [grid cells] Hilbert rank of each cell: tile lists sort their beads along the curve
    int* sbead;                 // [rows padded to 64] bead | type << 24 of rank k, -1 = padding
    float2* sparam;             // [rows padded] per-bead mixing parameters (inner, outer strength); spos.w carries a
    uint4* saux;                // [rows padded] {inner, outer strength, bead | type << 24, 0}: what the force kernel loads per j-bead
    const float* slot_a;        // [np] per slot: half sigma / radius / scaled charge
    const float2* slot_b;       // [np]
    uint4* sc_center;           // [superclusters] bounding-box centre (fixed point)
    float4* sc_half;            // [superclusters] half extents (nm)
    int tl_nsc, tl_stride;      // superclusters allocated; entries per supercluster row
    int* tl_count;              // [tl_nsc] entries in use
    int* tl_count_in;           // [tl_nsc] leading entries that touch owned beads only (slab decomposition: they need no halo data)
    int* tl_start_in;           // [tl_nsc + 1] prefix sums of the interior weights
    int* tl_weight;             // [tl_nsc] estimated work of the row (tile evaluations + overheads)
    int* tl_weight_in;          // [tl_nsc] the same for the interior entries
    int* tl_start;              // [tl_nsc + 1] exclusive prefix sum of tl_weight (tile_scan_kernel)
    uint2* tl_entries;          // [tl_nsc * tl_stride] {j-group, i-cluster mask | mask block << 8}
    unsigned int* tl_masks;     // [tl_mask_cap * 8] lane masks of the entries that need them (block index is 1-based)
    int* tl_mask_count;
    int tl_mask_cap;
    const int2* excl_pairs;     // excluded bead pairs (tile_excl_kernel)
    int n_excl_pairs;
    float t_alpha, t_beta, t_qz, t_kappa;  // pair form, see TileForm
    int t_geo;
};

struct BondedDev {
    int n, npar;
    const int* idx;       // nbody * n
    const double* par;    // npar * n
    const int* meta;      // group | periodic<<8 | table<<16
    const double* tabs;   // 24 doubles per table
    // slab decomposition: the terms with an owned bead on this rank, compacted at every list rebuild (dd_check_bonded_kernel);
    // null otherwise
    const int* dd_terms;
    const int* dd_count;
};

// donor x acceptor terms (3SPN.2 base pairing / cross stacking, PDNS): per-term tables + one Verlet list of
// (term, donor, acceptor) combinations inside cutoff + skin, rebuilt together with the pair lists
struct DATermDev {
    int kind, group, periodic, npar;
    int nd, na, nexcl, pad;
    long long cand_begin;              // offset of the term in the flattened donor x acceptor candidate space
    double rc2, rlist2;                // cutoff^2, (cutoff + skin)^2
    double g0, g1, g2;
    const int* donors;                 // 3 * nd
    const int* acceptors;              // 3 * na
    const double* par;                 // npar per donor (base pair, PDNS) or per acceptor (cross stacking)
    const unsigned long long* excl;    // sorted keys donor * na + acceptor
};
struct DAListDev {
    int nterms, capacity;
    long long ncand;
    const DATermDev* terms;
    unsigned int* count;               // entries in use
    uint2* entries;                    // x = term << 24 | donor, y = acceptor
};

// spring between the mass-weighted centres of two bead groups: idx / w hold group 1 then group 2, w = mass / group mass
struct ComDev {
    int group, periodic, n1, n2, begin, pad;
    double k, v0;
    const int* idx;
    const double* w;
};
constexpr int kComBlock = 256;

struct WorkSeg {
    int block_begin, nblocks, kind, id, tpb, chunk;  // pair segments: tpb threads per bead, chunk beads per block
    int staged;                                       // unused (the bulk-copy row staging experiment of round 1 was removed)
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
    // replica batch (unlimited boundary only): beads [r*nper, (r+1)*nper) form replica r, which has its own noise
    // key and its own local frame (pair-kernel positions are FP32 offsets from the replica's origin)
    int nrep, nper;
    const double* origin;       // [3*nrep] or null (single system, origin 0)
    const uint2* rep_seed;      // [nrep] or null
    const unsigned char* role;  // slab decomposition: 0 = not on this rank, 1 = owned, 2 = halo; null = single domain
    const int* present;         // slab decomposition: the beads present on this rank (owned + halo), compacted at every rebuild
    const int* n_present;
    int pdl_early;              // programmatic dependent launch: release the successor's launch at block start
};

// rows (ranks of the cell-sorted order) in use: under a slab decomposition only the beads present on this rank are binned
// number of sort keys: one per cell (per replica); slab-decomposed tile lists keep two copies of the grid, owned beads in the
// first and halo beads in the second, so that the halo beads come last in the sorted order (halo_split)
__device__ __forceinline__ int list_nkeys(const PairListDev& L) { return L.grid->ncell * L.nrep * (1 + L.halo_split); }
__device__ __forceinline__ int list_rows(const PairListDev& L) { return L.role ? L.cell_start[list_nkeys(L)] : L.np; }
// rows of owned beads (they come first when halo_split is set)
__device__ __forceinline__ int list_rows_owned(const PairListDev& L) { return L.halo_split ? L.cell_start[L.grid->ncell * L.nrep] : list_rows(L); }
__device__ __forceinline__ int bonded_nbody_dev(int kind) { return kind == 10 ? 1 : (kind <= 3 ? 2 : ((kind <= 5 || kind == 9) ? 3 : 4)); }

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

__device__ __forceinline__ float fast_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));  // <= 1 ulp
    return y;
}

__device__ __forceinline__ float fast_rsqrt(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// programmatic dependent launch (sm_90+): a kernel launched with the programmatic-stream-serialization attribute may
// begin while its predecessor drains; it must not touch the predecessor's output before pdl_wait().  pdl_trigger() lets
// the successor's launch start as soon as every block of this grid has begun.  Both are no-ops in a plain launch.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

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
        if (P.origin) { const double* o = P.origin + 3 * (i / P.nper); px -= o[0]; py -= o[1]; pz -= o[2]; }
        q.x = __float_as_uint((float)px); q.y = __float_as_uint((float)py); q.z = __float_as_uint((float)pz);
    }
    q.w = 0;
    posq[i] = q;
}

// Monte Carlo barostat trial (OpenMM ApplyMonteCarloBarostatKernel::scaleCoordinates as MonteCarloAnisotropicBarostat /
// MonteCarloMembraneBarostat call it): every molecule moves rigidly so that its centre (plain mean of its beads), wrapped
// into the old box, is multiplied by the length scale of each axis.  One thread per molecule, beads in ascending order
// (the oracle adds in the same order).
__global__ void scale_molecules_kernel(int nmol, const int* __restrict__ mol_start, const int* __restrict__ mol_bead, double* __restrict__ x,
                                       double bx, double by, double bz, double sx, double sy, double sz) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nmol) return;
    const int lo = mol_start[m], hi = mol_start[m + 1];
    double cx = 0, cy = 0, cz = 0;
    for (int k = lo; k < hi; k++) {
        const int i = mol_bead[k];
        cx += x[3 * i]; cy += x[3 * i + 1]; cz += x[3 * i + 2];
    }
    const double cnt = (double)(hi - lo);
    cx /= cnt; cy /= cnt; cz /= cnt;
    const double wx = cx - floor(cx / bx) * bx, wy = cy - floor(cy / by) * by, wz = cz - floor(cz / bz) * bz;
    const double dx = wx * sx - cx, dy = wy * sy - cy, dz = wz * sz - cz;
    for (int k = lo; k < hi; k++) {
        const int i = mol_bead[k];
        x[3 * i] += dx; x[3 * i + 1] += dy; x[3 * i + 2] += dz;
    }
}

// replica batch: move every replica's local frame to its current centroid (one warp per replica).  Forces are
// translation invariant, so this only keeps the FP32 offsets small and the shared cell grid compact.
__global__ void recentre_kernel(StepParams P, const double* __restrict__ x, double* __restrict__ origin) {
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (r >= P.nrep) return;
    double sx = 0, sy = 0, sz = 0;
    for (int k = lane; k < P.nper; k += 32) {
        const int i = r * P.nper + k;
        sx += x[3 * i]; sy += x[3 * i + 1]; sz += x[3 * i + 2];
    }
    for (int o = 16; o > 0; o >>= 1) {
        sx += __shfl_xor_sync(0xffffffffu, sx, o); sy += __shfl_xor_sync(0xffffffffu, sy, o); sz += __shfl_xor_sync(0xffffffffu, sz, o);
    }
    if (lane == 0) { origin[3 * r] = sx / P.nper; origin[3 * r + 1] = sy / P.nper; origin[3 * r + 2] = sz / P.nper; }
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
        if (tot <= (long long)(L.max_cells / L.nrep)) break;
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
    const int bead = L.part[a];
    if (L.role && L.role[bead] == 0) { L.cell_of_slot[a] = -1; L.rank_of_bead[bead] = -1; return; }
    cell_coords<PER>(g, posq[bead], cx, cy, cz);
    int c = (cz * g.nc[1] + cy) * g.nc[0] + cx;
    if (L.cell_rank) c = L.cell_rank[c];  // tile lists: cells are keyed by their rank along the Hilbert curve
    if (L.nrep > 1) c += (bead / L.nper) * g.ncell;
    if (L.halo_split && L.role[bead] == 2) c += g.ncell;
    L.cell_of_slot[a] = c;
    atomicAdd(&L.cell_count[c], 1);
}

// single block: exclusive scan of cell_count[0..ncell*nrep) -> cell_start, cursor; zero the counts.
// Every thread owns kScanItems consecutive cells per pass, so large (replica-keyed) grids stay cheap.
constexpr int kScanItems = 8;
__global__ void scan_kernel(const PairListDev* lists, int li) {
    const PairListDev& L = lists[li];
    const int ncell = list_nkeys(L);
    __shared__ int warp_sums[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int base = 0; base < ncell; base += blockDim.x * kScanItems) {
        const int i0 = base + threadIdx.x * kScanItems;
        int vals[kScanItems];
        int v = 0;
#pragma unroll
        for (int q = 0; q < kScanItems; q++) { vals[q] = (i0 + q < ncell) ? L.cell_count[i0 + q] : 0; v += vals[q]; }
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
        int run = carry + (wid > 0 ? warp_sums[wid - 1] : 0) + s - v;
#pragma unroll
        for (int q = 0; q < kScanItems; q++) {
            if (i0 + q < ncell) { L.cell_start[i0 + q] = run; L.cell_cursor[i0 + q] = run; L.cell_count[i0 + q] = 0; }
            run += vals[q];
        }
        __syncthreads();
        if (threadIdx.x == 0) carry += warp_sums[nw - 1];
        __syncthreads();
    }
    if (threadIdx.x == 0) L.cell_start[ncell] = carry;
}

// The same scan for large grids (a million fine cells at 1 M beads took the single block 0.65 ms): three launches.
//   scan_blocks_kernel  every block scans its 8192 cells locally and leaves its total in scan_tmp[block]
//   scan_tops_kernel    one block turns the block totals into exclusive offsets (up to 8192 blocks = 2^26 cells)
//   scan_finish_kernel  adds the offsets, sets the cursors, clears the counts, writes cell_start[ncell]
constexpr int kScanChunk = 1024 * kScanItems;
__global__ void __launch_bounds__(1024) scan_blocks_kernel(const PairListDev* lists, int li) {
    const PairListDev& L = lists[li];
    const int ncell = list_nkeys(L);
    __shared__ int warp_sums[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int i0 = blockIdx.x * kScanChunk + threadIdx.x * kScanItems;
    int vals[kScanItems];
    int v = 0;
#pragma unroll
    for (int q = 0; q < kScanItems; q++) { vals[q] = (i0 + q < ncell) ? L.cell_count[i0 + q] : 0; v += vals[q]; }
    int s = v;
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += t; }
    if (lane == 31) warp_sums[wid] = s;
    __syncthreads();
    if (wid == 0) {
        int ws = warp_sums[lane];
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, ws, o); if (lane >= o) ws += t; }
        warp_sums[lane] = ws;
    }
    __syncthreads();
    int run = (wid > 0 ? warp_sums[wid - 1] : 0) + s - v;
#pragma unroll
    for (int q = 0; q < kScanItems; q++) { if (i0 + q < ncell) L.cell_start[i0 + q] = run; run += vals[q]; }
    if (threadIdx.x == 1023) L.scan_tmp[blockIdx.x] = warp_sums[31];
}
__global__ void __launch_bounds__(1024) scan_tops_kernel(const PairListDev* lists, int li, int nblocks) {
    const PairListDev& L = lists[li];
    __shared__ int warp_sums[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int i0 = threadIdx.x * kScanItems;
    int vals[kScanItems];
    int v = 0;
#pragma unroll
    for (int q = 0; q < kScanItems; q++) { vals[q] = (i0 + q < nblocks) ? L.scan_tmp[i0 + q] : 0; v += vals[q]; }
    int s = v;
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += t; }
    if (lane == 31) warp_sums[wid] = s;
    __syncthreads();
    if (wid == 0) {
        int ws = warp_sums[lane];
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, ws, o); if (lane >= o) ws += t; }
        warp_sums[lane] = ws;
    }
    __syncthreads();
    int run = (wid > 0 ? warp_sums[wid - 1] : 0) + s - v;
#pragma unroll
    for (int q = 0; q < kScanItems; q++) { if (i0 + q < nblocks) L.scan_tmp[i0 + q] = run; run += vals[q]; }
    if (threadIdx.x == 1023) L.scan_tmp[nblocks] = warp_sums[31];
}
__global__ void __launch_bounds__(1024) scan_finish_kernel(const PairListDev* lists, int li) {
    const PairListDev& L = lists[li];
    const int ncell = list_nkeys(L);
    const int off = L.scan_tmp[blockIdx.x];
    const int i0 = blockIdx.x * kScanChunk + threadIdx.x * kScanItems;
#pragma unroll
    for (int q = 0; q < kScanItems; q++)
        if (i0 + q < ncell) { const int v = L.cell_start[i0 + q] + off; L.cell_start[i0 + q] = v; L.cell_cursor[i0 + q] = v; L.cell_count[i0 + q] = 0; }
    if (blockIdx.x == 0 && threadIdx.x == 0) L.cell_start[ncell] = L.scan_tmp[gridDim.x];
}

__global__ void scatter_kernel(const PairListDev* lists, int li) {
    const PairListDev& L = lists[li];
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= L.np) return;
    const int c = L.cell_of_slot[a];
    if (c < 0) return;
    int pos = atomicAdd(&L.cell_cursor[c], 1);
    L.tmp_slots[pos] = a;
}

// Deterministic cell-sorted order: inside a cell, members are ranked by slot (= ascending bead index), so the
// list order, and with it the FP32 summation order of the force kernel, is reproducible run to run.  Also
// fills the list's cell-sorted position array (w = bead | type<<24) and the bead -> rank map.
__global__ void rank_fill_kernel(const PairListDev* lists, int li, const uint4* __restrict__ posq) {
    const PairListDev& L = lists[li];
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= list_rows(L)) return;
    const int a = L.tmp_slots[p];
    const int c = L.cell_of_slot[a];
    const int s = L.cell_start[c], e = L.cell_start[c + 1];
    int r = 0;
    for (int q = s; q < e; q++) r += (L.tmp_slots[q] < a) ? 1 : 0;
    const int k = s + r;
    const int bead = L.part[a];
    L.sorted_slots[k] = a;
    L.rank_of_bead[bead] = k;
    L.bead_of_rank[k] = bead;
    uint4 q4 = posq[bead];
    q4.w = (L.form == SEG_PAIR_DH) ? __float_as_uint(L.qscaled[a]) : ((unsigned int)bead | ((unsigned int)L.type[a] << 24));
    L.spos[k] = q4;
}

// one warp per participant (in cell-sorted order): walk the 5x5 rows of neighbouring cells -- x-runs are
// contiguous in the cell-sorted position array, so candidates stream in with coalesced 16-byte loads -- test
// the canonical distance, exclusions and interaction classes, compact with ballots.
// The row is assembled in shared memory and written out ordered by build-time distance zone [A | A' | B | C]
// (scan order inside a zone, so it stays deterministic):
//   A : r^2 < zone_a2 = (min(r_switch, rc) - skin)^2  -> inside the cutoff and below the switch until the next
//       rebuild (beads move at most skin/2 each): the force kernel evaluates these with no tests (nbr_inner)
//   A': r^2 < zone_b2   probably still below the switching distance  (warp-uniform vote skips the switch)
//   B : r^2 < zone_c2   switching region / cutoff shell
//   C : the rest of the skin                                      (warp-uniform vote skips everything)
// count_only: just record the neighbour count (used to size the list).  zoned == 0: no shared memory, one zone.
// DH / CLS: compile-time copies of L.form == SEG_PAIR_DH and L.cls != null.  The candidate loop is the cost of a rebuild (C2: 63
// groups of 32 candidates per bead): everything it needs from the list descriptor is hoisted into registers before it, and the
// per-zone counts are kept per lane and summed once (ncu, round 2: 195 -> ~70 warp instructions per candidate group).
template <bool PER, bool DH, bool CLS>
__global__ void __launch_bounds__(256) build_list_kernel(const PairListDev* lists, int li, StepParams P, int* __restrict__ flags, int count_only, int zoned) {
    extern __shared__ __align__(16) unsigned int bl_smem[];
    const PairListDev& L = lists[li];
    const int lane = threadIdx.x & 31;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w == 0 && lane == 0 && !count_only && list_rows(L) > L.row_capacity) flags[FLAG_OVERFLOW] = 1;
    if (w >= list_rows(L)) return;
    const GridDev& g = *L.grid;
    const uint4* __restrict__ spos = L.spos;
    const int* __restrict__ excl_idx = L.excl_idx;
    const signed char* __restrict__ cls = L.cls;
    const int* __restrict__ sslots = L.sorted_slots;
    const int* __restrict__ bor = L.bead_of_rank;
    const int a = L.sorted_slots[w];
    if (L.role && L.role[L.part[a]] != 1) {
        // halo bead: present as a partner only, its own row stays empty (its owner computes its forces)
        if (lane == 0) { L.nbr_count[w] = 0; L.nbr_inner[w] = 0; }
        return;
    }
    const uint4 pi = spos[w];
    const int cig = L.cell_of_slot[a];
    const int ncx = g.nc[0], ncy = g.nc[1], ncz = g.nc[2];
    const int cell_base = (L.nrep > 1) ? (cig / g.ncell) * g.ncell : 0;  // first cell of this bead's replica
    const int ci = cig - cell_base;
    const int cx = ci % ncx, cy = (ci / ncx) % ncy, cz = ci / (ncx * ncy);
    const int e0 = L.excl_off[a], e1 = L.excl_off[a + 1];
    const unsigned int my_allow = CLS ? L.allow[(int)L.cls[a]] : 0xffffffffu;
    const float rl2 = L.rlist2, za2 = L.zone_a2, zb2 = L.zone_b2, zc2 = L.zone_c2;
    const int stride = L.stride;
    unsigned int* row = L.nbr + (size_t)w * stride;
    const bool stage = zoned && !count_only;
    const int wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    unsigned int* s_row = bl_smem + (size_t)wib * stride;
    unsigned char* s_zone = reinterpret_cast<unsigned char*>(bl_smem + (size_t)wpb * stride) + (size_t)wib * stride;
    int count = 0, cz0 = 0, cz1 = 0, cz2 = 0;
    int n0 = 0, n1 = 0, n2 = 0;   // this lane's accepted entries per zone
    const unsigned int lt = (1u << lane) - 1u;

    int nz, ny, z0, y0;
    if (PER) { nz = min(5, ncz); z0 = (ncz >= 5) ? cz - 2 : 0; ny = min(5, ncy); y0 = (ncy >= 5) ? cy - 2 : 0; }
    else { z0 = max(cz - 2, 0); nz = min(cz + 2, ncz - 1) - z0 + 1; y0 = max(cy - 2, 0); ny = min(cy + 2, ncy - 1) - y0 + 1; }
    // Geometric pruning of the 5x5x5 block: a (y,z) row of cells whose nearest edge is farther than the list radius
    // is skipped, and the x-run of the others is cut to the chord of the sphere.  Conservative (1e-4 margins): every
    // surviving candidate still passes through the canonical distance test, so the pair set is unchanged.
    float fx_c, fy_c, fz_c, wx, wy, wz;  // position inside the own cell (cell units), cell widths (nm)
    bool prune;
    if (PER) {
        const float s32 = 2.3283064365386963e-10f;
        fx_c = (float)pi.x * s32 * (float)ncx - (float)cx; fy_c = (float)pi.y * s32 * (float)ncy - (float)cy; fz_c = (float)pi.z * s32 * (float)ncz - (float)cz;
        wx = P.fscale[0] * 4294967296.f / (float)ncx; wy = P.fscale[1] * 4294967296.f / (float)ncy; wz = P.fscale[2] * 4294967296.f / (float)ncz;
        prune = ncx >= 5 && ncy >= 5 && ncz >= 5;
    } else {
        fx_c = (__uint_as_float(pi.x) - g.lo[0]) * g.inv_w - (float)cx; fy_c = (__uint_as_float(pi.y) - g.lo[1]) * g.inv_w - (float)cy;
        fz_c = (__uint_as_float(pi.z) - g.lo[2]) * g.inv_w - (float)cz;
        wx = wy = wz = 1.0f / g.inv_w;
        prune = true;
    }
    fx_c = fminf(fmaxf(fx_c, 0.f), 1.f); fy_c = fminf(fmaxf(fy_c, 0.f), 1.f); fz_c = fminf(fmaxf(fz_c, 0.f), 1.f);
    const float rl2m = rl2 * 1.0002f;

    // lane r < nz*ny sets up (y,z) row r: up to two contiguous runs [kb0,ke0) [kb1,ke1) of the cell-sorted order.
    // All rows are resolved in parallel (their cell_start look-ups overlap) and broadcast by shuffles below.
    int kb0 = 0, ke0 = 0, kb1 = 0, ke1 = 0;
    const int nrows = nz * ny;
    if (lane < nrows) {
        const int kz = lane / ny, ky = lane - kz * ny;
        int oz = z0 + kz, oy = y0 + ky;
        const int dzc = oz - cz, dyc = oy - cy;
        if (PER) {
            if (oz < 0) oz += ncz; else if (oz >= ncz) oz -= ncz;
            if (oy < 0) oy += ncy; else if (oy >= ncy) oy -= ncy;
        }
        const float gz = prune ? wz * fmaxf(0.f, dzc > 0 ? (float)dzc - fz_c - 1e-4f : (dzc < 0 ? fz_c - (float)dzc - 1.f - 1e-4f : 0.f)) : 0.f;
        const float gy = prune ? wy * fmaxf(0.f, dyc > 0 ? (float)dyc - fy_c - 1e-4f : (dyc < 0 ? fy_c - (float)dyc - 1.f - 1e-4f : 0.f)) : 0.f;
        const float rem2 = rl2m - gz * gz - gy * gy;
        if (rem2 >= 0.f) {
            // x cells reached by the chord: [cx + dlo, cx + dhi] within +-2
            int dlo = -2, dhi = 2;
            if (prune) {
                const float rx = sqrtf(rem2) / wx + 1e-4f;
                dlo = max(-2, (int)floorf(fx_c - rx)); dhi = min(2, (int)floorf(fx_c + rx));
            }
            int xs0, xe0, xs1 = 0, xe1 = -1;
            if (PER) {
                if (ncx < 5) { xs0 = 0; xe0 = ncx - 1; }
                else {
                    const int lo = cx + dlo, hi = cx + dhi;
                    if (lo < 0) { xs0 = 0; xe0 = hi; xs1 = lo + ncx; xe1 = ncx - 1; }
                    else if (hi >= ncx) { xs0 = 0; xe0 = hi - ncx; xs1 = lo; xe1 = ncx - 1; }
                    else { xs0 = lo; xe0 = hi; }
                }
            } else { xs0 = max(cx + dlo, 0); xe0 = min(cx + dhi, ncx - 1); }
            const int rowbase = cell_base + (oz * ncy + oy) * ncx;
            if (xs0 <= xe0) { kb0 = L.cell_start[rowbase + xs0]; ke0 = L.cell_start[rowbase + xe0 + 1]; }
            if (xs1 <= xe1) { kb1 = L.cell_start[rowbase + xs1]; ke1 = L.cell_start[rowbase + xe1 + 1]; }
        }
    }
    // the first four exclusions of this bead live in registers
    const int ex0 = (e0 < e1) ? excl_idx[e0] : -1, ex1 = (e0 + 1 < e1) ? excl_idx[e0 + 1] : -1;
    const int ex2 = (e0 + 2 < e1) ? excl_idx[e0 + 2] : -1, ex3 = (e0 + 3 < e1) ? excl_idx[e0 + 3] : -1;
    const bool more_excl = e0 + 4 < e1;

    // one group of 32 candidates: lane's candidate k (rank in the cell-sorted order) with its position pj; ke = end of the run
    auto group = [&](const int k, const uint4 pj, const int ke) {
        bool ok = false;
        unsigned int entry = 0;
        int zone = 3;
        if (k < ke && k != w) {
            float dx, dy, dz;
            delta<PER>(pi, pj, P.fscale, dx, dy, dz);
            const float r2 = norm2(dx, dy, dz);
            if (r2 < rl2) {
                ok = true;
                if (CLS) ok = (my_allow >> (int)cls[sslots[k]]) & 1u;
                const int j = DH ? bor[k] : (int)(pj.w & 0xFFFFFFu);
                ok = ok && j != ex0 && j != ex1 && j != ex2 && j != ex3;
                if (more_excl) for (int e = e0 + 4; ok && e < e1; e++) ok = excl_idx[e] != j;
                entry = DH ? (unsigned int)k : ((unsigned int)k | (pj.w & 0xFF000000u));
                zone = (r2 < za2) ? 0 : (r2 < zb2) ? 1 : (r2 < zc2) ? 2 : 3;
            }
        }
        const unsigned int m = __ballot_sync(0xffffffffu, ok);
        if (m == 0u) return;
        if (!count_only) {
            const int pos = count + __popc(m & lt);
            if (ok && pos < stride) {
                if (stage) { s_row[pos] = entry; s_zone[pos] = (unsigned char)zone; }
                else row[pos] = entry;
            }
            if (ok) { n0 += (zone == 0); n1 += (zone == 1); n2 += (zone == 2); }
        }
        count += __popc(m);
    };

    for (int r = 0; r < 2 * nrows; r++) {
        const int kb = __shfl_sync(0xffffffffu, (r & 1) ? kb1 : kb0, r >> 1);
        const int ke = __shfl_sync(0xffffffffu, (r & 1) ? ke1 : ke0, r >> 1);
        for (int k0 = kb; k0 < ke; k0 += 64) {
            // two 32-candidate groups per iteration: both position loads are in flight together
            const int ka = k0 + lane, kbb = k0 + 32 + lane;
            uint4 pja = make_uint4(0, 0, 0, 0), pjb = make_uint4(0, 0, 0, 0);
            if (ka < ke) pja = __ldg(&spos[ka]);
            if (kbb < ke) pjb = __ldg(&spos[kbb]);
            group(ka, pja, ke);
            if (k0 + 32 < ke) group(kbb, pjb, ke);   // warp-uniform
        }
    }
    if (stage) {
        for (int o = 16; o > 0; o >>= 1) {
            n0 += __shfl_xor_sync(0xffffffffu, n0, o); n1 += __shfl_xor_sync(0xffffffffu, n1, o); n2 += __shfl_xor_sync(0xffffffffu, n2, o);
        }
        cz0 = n0; cz1 = n1; cz2 = n2;
    }
    if (count_only) {
        if (lane == 0) { L.nbr_count[w] = count; atomicMax(&flags[FLAG_MAXDISP + 1], count); }
        return;
    }
    int n_inner = 0;
    if (count > stride) {
        if (lane == 0) flags[FLAG_OVERFLOW] = 1;  // the engine re-sizes the rows and rebuilds
        count = stride;
        if (stage) for (int t = lane; t < count; t += 32) row[t] = s_row[t];
    } else if (stage) {
        __syncwarp();
        int off0 = 0, off1 = cz0, off2 = cz0 + cz1, off3 = cz0 + cz1 + cz2;
        n_inner = cz0;
        for (int t0 = 0; t0 < count; t0 += 32) {
            const int t = t0 + lane;
            const bool live = t < count;
            const unsigned int e = live ? s_row[t] : 0u;
            const int z = live ? (int)s_zone[t] : -1;
            const unsigned int m0 = __ballot_sync(0xffffffffu, z == 0), m1 = __ballot_sync(0xffffffffu, z == 1);
            const unsigned int m2 = __ballot_sync(0xffffffffu, z == 2), m3 = __ballot_sync(0xffffffffu, z == 3);
            if (z == 0) row[off0 + __popc(m0 & lt)] = e;
            else if (z == 1) row[off1 + __popc(m1 & lt)] = e;
            else if (z == 2) row[off2 + __popc(m2 & lt)] = e;
            else if (z == 3) row[off3 + __popc(m3 & lt)] = e;
            off0 += __popc(m0); off1 += __popc(m1); off2 += __popc(m2); off3 += __popc(m3);
        }
    }
    // pad the row to a multiple of 128 with the self entry (distance 0, never interacts): the T = 32 force path reads
    // four whole 32-entry groups per iteration (the stride is a multiple of 128)
    for (int pad = count + lane; pad < ((count + 127) & ~127) && pad < stride; pad += 32)
        row[pad] = DH ? (unsigned int)w : ((unsigned int)w | (pi.w & 0xFF000000u));
    if (lane == 0) { L.nbr_count[w] = count; L.nbr_inner[w] = n_inner; }
}

// unlimited boundary (virtual box, see ocm_engine.cu): every bead must stay inside [lo, hi], the range for which the
// virtual box keeps periodic images out of every list radius
__global__ void vbox_check_kernel(int n, const double* __restrict__ x, double lo0, double lo1, double lo2, double hi0, double hi1, double hi2,
                                  int* __restrict__ flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double px = x[3 * i], py = x[3 * i + 1], pz = x[3 * i + 2];
    if (!(px >= lo0 && px <= hi0 && py >= lo1 && py <= hi1 && pz >= lo2 && pz <= hi2)) flags[FLAG_VBOX] = 1;
}

__global__ void save_ref_kernel(int n, const double* __restrict__ x, double* __restrict__ xref) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 3 * n) xref[i] = x[i];
}

// ------------------------------------------------------------------------------------------------
// pair forces.  T lanes walk the Verlet row of one participant (coalesced 4-byte entries), gather 16-byte
// partner positions from the list's cell-sorted array (neighbouring entries are neighbouring addresses, so a
// warp-wide gather touches few cache lines), and reduce the FP32 force with shuffles; one fixed-point atomic
// triple per bead merges the result with the other lists and the bonded terms.  A block owns the contiguous
// rank range [blk*chunk, (blk+1)*chunk): beads that are close in space, so rows overlap in L1.
//
// Two-zone 12-6 family: per type pair one float4 {12 c12, 6 c6, a2, w};
//   -dE/dr / r = zone * (12 c12 / r^12 - 6 c6 / r^6) / r^2,  zone = r^2 < a2 ? k_in : k_out + k_w * w
template <bool PER, int T, bool SW>
__device__ __forceinline__ void pair_tz_rows(const PairListDev& L, int k_lo, int k_hi, const StepParams& P,
                                             unsigned long long* __restrict__ fbuf, const float4* __restrict__ tab) {
    constexpr int G = kBlock / T;
    const int g = threadIdx.x / T, s = threadIdx.x % T;
    const uint4* __restrict__ spos = L.spos;
    const int nt = L.ntypes;
    const float rc2 = L.rc2, rsw = L.rsw, inv_sw = L.inv_sw, rsw2 = rsw * rsw;
    const float kin = L.k_in;
    for (int base = k_lo; base < k_hi; base += G) {
        const int k = base + g;
        const bool active = k < k_hi;
        float fx = 0.f, fy = 0.f, fz = 0.f;
        unsigned int wi = 0;
        if (active) {
            const uint4 pi = spos[k];
            wi = pi.w;
            const int trow = (int)(wi >> 24) * nt;
            const int n = L.nbr_count[k];
            const unsigned int* __restrict__ row = L.nbr + (size_t)k * L.stride;
#pragma unroll 4
            for (int t = s; t < n; t += T) {
                const unsigned int e = __ldg(&row[t]);
                const uint4 pj = __ldg(&spos[(int)(e & 0xFFFFFFu)]);
                float dx, dy, dz;
                delta<PER>(pi, pj, P.fscale, dx, dy, dz);
                const float r2 = norm2(dx, dy, dz);
                const float4 c = tab[trow + (int)(e >> 24)];
                const float ir2 = fast_rcp(r2);
                const float ir6 = ir2 * ir2 * ir2;
                const float zone = (r2 < c.z) ? kin : c.w;
                const float a12 = c.x * ir6;
                float fr = zone * (a12 - c.y) * ir6 * ir2;
                if (SW) {
                    if (r2 > rsw2) {
                        // OpenMM switching function on [rsw, rc]: S = (1-t)^3 (1 + 3t + 6t^2), S' = -30 t^2 (1-t)^2 / (rc-rsw)
                        const float ir = fast_rsqrt(r2);
                        const float tt = __saturatef((r2 * ir - rsw) * inv_sw), u = 1.f - tt;
                        const float S = u * u * u * fmaf(tt, fmaf(tt, 6.f, 3.f), 1.f);
                        const float tu = tt * u;
                        const float dS = -30.f * inv_sw * tu * tu;
                        const float E = zone * (a12 * (1.f / 12.f) - c.y * (1.f / 6.f)) * ir6;
                        fr = S * fr - E * dS * ir;
                    }
                }
                fr = (r2 < rc2) ? fr : 0.f;
                fx = fmaf(fr, dx, fx); fy = fmaf(fr, dy, fy); fz = fmaf(fr, dz, fz);
            }
        }
#pragma unroll
        for (int o = T / 2; o > 0; o >>= 1) {
            fx += __shfl_xor_sync(0xffffffffu, fx, o);
            fy += __shfl_xor_sync(0xffffffffu, fy, o);
            fz += __shfl_xor_sync(0xffffffffu, fz, o);
        }
        if (active && s == 0) add_force_fixed(fbuf, P.n, (int)(wi & 0xFFFFFFu), fx, fy, fz);
    }
}

// one pair evaluation of the two-zone 12-6 family; `in` = inside the cutoff (false for padding lanes).
// Branches are warp-uniform votes: rows are ordered near -> far at build time (zones A' | B | C), so most warp
// iterations of A' skip the switching function and most of C (the outer skin) skip everything.
template <bool SW>
__device__ __forceinline__ void tz_pair(const float4 c, float r2, bool in, float dx, float dy, float dz, float kin,
                                        float rsw, float rsw2, float inv_sw, float& fx, float& fy, float& fz) {
    if (!__any_sync(0xffffffffu, in)) return;
    const float ir2 = fast_rcp(r2);
    const float ir6 = ir2 * ir2 * ir2;
    const float zone = (r2 < c.z) ? kin : c.w;
    const float a12 = c.x * ir6;
    float fr = zone * (a12 - c.y) * ir6 * ir2;
    if (SW) {
        if (__any_sync(0xffffffffu, in && r2 > rsw2)) {
            // OpenMM switching function on [rsw, rc]: S = (1-t)^3 (1 + 3t + 6t^2), S' = -30 t^2 (1-t)^2 / (rc-rsw)
            const float ir = fast_rsqrt(r2);
            const float tt = __saturatef((r2 * ir - rsw) * inv_sw), u = 1.f - tt;
            const float S = u * u * u * fmaf(tt, fmaf(tt, 6.f, 3.f), 1.f);
            const float tu = tt * u;
            const float dS = -30.f * inv_sw * tu * tu;
            const float E = zone * (a12 * (1.f / 12.f) - c.y * (1.f / 6.f)) * ir6;
            fr = S * fr - E * dS * ir;
        }
    }
    fr = in ? fr : 0.f;
    fx = fmaf(fr, dx, fx); fy = fmaf(fr, dy, fy); fz = fmaf(fr, dz, fz);
}
// zone A of a row: the pair was closer than (r_switch or rc) - skin when the list was built, so it is inside
// the cutoff and below the switching distance for the whole life of the list: no tests at all
__device__ __forceinline__ void tz_pair_inner(const float4 c, float r2, float dx, float dy, float dz, float kin,
                                              float& fx, float& fy, float& fz) {
    const float ir2 = fast_rcp(r2);
    const float ir6 = ir2 * ir2 * ir2;
    const float zone = (r2 < c.z) ? kin : c.w;
    const float fr = zone * (c.x * ir6 - c.y) * ir6 * ir2;
    fx = fmaf(fr, dx, fx); fy = fmaf(fr, dy, fy); fz = fmaf(fr, dz, fz);
}

// T = 32 path, rows streamed from global memory with register double-buffering: the next four row entries
// (and the first four of the next bead's row) are requested before the current four pairs are evaluated, so the
// row stream's L2 latency overlaps arithmetic; four gathers + four table look-ups are in flight per lane.
// OCM_ROW_UNROLL: 32-entry groups per lane and iteration = gathers in flight per lane.  Measured on B200 at the C2 bench state
// (tools/lib_ab.py): 8 gathers in flight need 4 - 5 blocks per SM and lose (force launch 42.7 - 43.0 us against 37.2 us); 7 blocks
// per SM with 4 gathers spill and gain nothing (37.0 us).  The rows are padded to 4 x 32 entries, so 4 it is.
#ifndef OCM_ROW_UNROLL
#define OCM_ROW_UNROLL 4
#endif
template <bool PER, bool SW>
__device__ __forceinline__ void pair_tz_rows_warp(const PairListDev& L, int k_lo, int k_hi, const StepParams& P,
                                                  unsigned long long* __restrict__ fbuf, const float4* tab) {
    constexpr int G = kBlock / 32, U = OCM_ROW_UNROLL, W = 32 * U;
    static_assert(W == 128, "build_list_kernel pads the rows to whole groups of 128 entries");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint4* __restrict__ spos = L.spos;
    const unsigned int* __restrict__ nbr = L.nbr;
    const int nt = L.ntypes, stride = L.stride;
    const float rc2 = L.rc2, rsw = L.rsw, inv_sw = L.inv_sw, rsw2 = rsw * rsw;
    const float kin = L.k_in;
    int k = k_lo + warp;
    if (k >= k_hi) return;
    int n = L.nbr_count[k];
    int n_in = L.nbr_inner[k];
    const unsigned int* row = nbr + (size_t)k * stride + lane;
    unsigned int en[U];
#pragma unroll
    for (int u = 0; u < U; u++) en[u] = (n > 0) ? __ldg(row + u * 32) : 0u;   // (rows are padded to whole groups of W entries)
    for (; k < k_hi; k += G) {
        const uint4 pi = spos[k];
        const int trow = (int)(pi.w >> 24) * nt;
        const int kn = k + G;
        const int n_next = (kn < k_hi) ? L.nbr_count[kn] : 0;
        const int n_in_next = (kn < k_hi) ? L.nbr_inner[kn] : 0;
        const unsigned int* row_next = nbr + (size_t)kn * stride + lane;
        float fx = 0.f, fy = 0.f, fz = 0.f;
        for (int t0 = 0; t0 < n; t0 += W) {
            unsigned int e[U];
            uint4 pj[U];
            float4 c[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                e[u] = en[u];   // beyond the row's end: the self entries the builder padded with (r2 = 0, never inside)
                pj[u] = __ldg(&spos[(int)(e[u] & 0xFFFFFFu)]);
                c[u] = tab[trow + (int)(e[u] >> 24)];
            }
            // request the following group: rest of this row, or the head of the next bead's row
            if (t0 + W < n) {
#pragma unroll
                for (int u = 0; u < U; u++) en[u] = __ldg(row + t0 + W + u * 32);
            } else if (n_next > 0) {
#pragma unroll
                for (int u = 0; u < U; u++) en[u] = __ldg(row_next + u * 32);
            }
            // evaluated four groups at a time (the four evaluations interleave), whatever the number of gathers in flight
#pragma unroll
            for (int h = 0; h < U; h += 4) {
                if (t0 + (h + 4) * 32 <= n_in) {
                    // all four groups lie in zone A
#pragma unroll
                    for (int u = h; u < h + 4; u++) {
                        float dx, dy, dz;
                        delta<PER>(pi, pj[u], P.fscale, dx, dy, dz);
                        tz_pair_inner(c[u], norm2(dx, dy, dz), dx, dy, dz, kin, fx, fy, fz);
                    }
                } else {
#pragma unroll
                    for (int u = h; u < h + 4; u++) {
                        if (t0 + u * 32 >= n) break;  // warp-uniform
                        float dx, dy, dz;
                        delta<PER>(pi, pj[u], P.fscale, dx, dy, dz);
                        const float r2 = norm2(dx, dy, dz);
                        const bool in = (r2 < rc2) && (r2 > 0.f);
                        tz_pair<SW>(c[u], r2, in, dx, dy, dz, kin, rsw, rsw2, inv_sw, fx, fy, fz);
                    }
                }
            }
        }
        if (n == 0) {
            // empty row (a halo bead of a slab decomposition, or a bead without partners): nothing to reduce or add
#pragma unroll
            for (int u = 0; u < U; u++) en[u] = (n_next > 0) ? __ldg(row_next + u * 32) : 0u;
        } else {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                fx += __shfl_xor_sync(0xffffffffu, fx, o);
                fy += __shfl_xor_sync(0xffffffffu, fy, o);
                fz += __shfl_xor_sync(0xffffffffu, fz, o);
            }
            if (lane == 0) add_force_fixed(fbuf, P.n, (int)(pi.w & 0xFFFFFFu), fx, fy, fz);
        }
        n = n_next;
        n_in = n_in_next;
        row = row_next;
    }
}

template <bool PER, int T>
__device__ __forceinline__ void pair_tz_block(const PairListDev& L, int blk, int chunk, const StepParams& P,
                                              unsigned long long* __restrict__ fbuf, float* smem) {
    const int nt2 = L.ntypes * L.ntypes;
    const float4* tab = L.tabF;
    if (L.smem_tab) {
        float4* sF = reinterpret_cast<float4*>(smem);
        for (int k = threadIdx.x; k < nt2; k += kBlock) sF[k] = L.tabF[k];
        __syncthreads();
        tab = sF;
    }
    const int k_lo = blk * chunk, k_hi = min(list_rows(L), k_lo + chunk);
    if (T == 32 && L.smem_tab) {
        const float4* sF = reinterpret_cast<const float4*>(smem);
        if (L.has_switch) pair_tz_rows_warp<PER, true>(L, k_lo, k_hi, P, fbuf, sF);
        else pair_tz_rows_warp<PER, false>(L, k_lo, k_hi, P, fbuf, sF);
        return;
    }
    if (L.has_switch) pair_tz_rows<PER, T, true>(L, k_lo, k_hi, P, fbuf, tab);
    else pair_tz_rows<PER, T, false>(L, k_lo, k_hi, P, fbuf, tab);
}

// screened Coulomb: E = qq (exp(-r/debye)/r - cc)
template <bool PER, int T>
__device__ __forceinline__ void pair_dh_block(const PairListDev& L, int blk, int chunk, const StepParams& P,
                                              unsigned long long* __restrict__ fbuf) {
    constexpr int G = kBlock / T;
    const int g = threadIdx.x / T, s = threadIdx.x % T;
    const uint4* __restrict__ spos = L.spos;
    const float rc2 = L.rc2, ik = L.inv_debye;
    const int k_lo = blk * chunk, k_hi = min(list_rows(L), k_lo + chunk);
    for (int base = k_lo; base < k_hi; base += G) {
        const int k = base + g;
        const bool active = k < k_hi;
        float fx = 0.f, fy = 0.f, fz = 0.f;
        unsigned int wi = 0;
        if (active) {
            const uint4 pi = spos[k];
            wi = (unsigned int)L.bead_of_rank[k];
            const float qi = __uint_as_float(pi.w);
            const int n = L.nbr_count[k];
            const unsigned int* __restrict__ row = L.nbr + (size_t)k * L.stride;
#pragma unroll 2
            for (int t = s; t < n; t += T) {
                const unsigned int e = __ldg(&row[t]);
                const uint4 pj = __ldg(&spos[(int)(e & 0xFFFFFFu)]);
                float dx, dy, dz;
                delta<PER>(pi, pj, P.fscale, dx, dy, dz);
                const float r2 = norm2(dx, dy, dz);
                const float qq = qi * __uint_as_float(pj.w);
                const float ir = fast_rsqrt(r2);
                const float r = r2 * ir;
                const float ex = __expf(-r * ik);
                float fr = qq * ex * (ir + ik) * ir * ir;
                fr = (r2 < rc2) ? fr : 0.f;
                fx = fmaf(fr, dx, fx); fy = fmaf(fr, dy, fy); fz = fmaf(fr, dz, fz);
            }
        }
#pragma unroll
        for (int o = T / 2; o > 0; o >>= 1) {
            fx += __shfl_xor_sync(0xffffffffu, fx, o);
            fy += __shfl_xor_sync(0xffffffffu, fy, o);
            fz += __shfl_xor_sync(0xffffffffu, fz, o);
        }
        if (active && s == 0) add_force_fixed(fbuf, P.n, (int)(wi & 0xFFFFFFu), fx, fy, fz);
    }
}

// attractive wells (iSoLFAttractiveForceFieldGenerator.cpp:41-50, TrigonometricForceFieldGenerator.cpp:43-53):
//   E = -eps for r < base, -eps f(u) for 0 <= u = (r - base) / omega < 1, 0 beyond;  f = cos^2(pi u / 2) or 1 + 2 u^3 - 3 u^2.
// The force acts in the shell only.  Per type pair {base, 1 / omega, amp, shape}:
//   dE/dr = amp sin(pi u), amp = eps pi / (2 omega)   (shape 0)     dE/dr = amp u (1 - u), amp = 6 eps / omega   (shape 1)
__device__ __forceinline__ float well_dedr(const float4 c, float r) {
    const float u = (r - c.x) * c.y;
    if (!(u > 0.f && u < 1.f)) return 0.f;
    return c.w == 0.f ? c.z * sinpif(u) : c.z * u * (1.f - u);
}
template <bool PER, int T>
__device__ __forceinline__ void pair_well_block(const PairListDev& L, int blk, int chunk, const StepParams& P,
                                                unsigned long long* __restrict__ fbuf) {
    constexpr int G = kBlock / T;
    const int g = threadIdx.x / T, s = threadIdx.x % T;
    const uint4* __restrict__ spos = L.spos;
    const float4* __restrict__ tab = L.tabF;
    const float rc2 = L.rc2;
    const int nt = L.ntypes;
    const int k_lo = blk * chunk, k_hi = min(list_rows(L), k_lo + chunk);
    for (int base = k_lo; base < k_hi; base += G) {
        const int k = base + g;
        const bool active = k < k_hi;
        float fx = 0.f, fy = 0.f, fz = 0.f;
        unsigned int wi = 0;
        if (active) {
            const uint4 pi = spos[k];
            wi = pi.w;
            const int trow = (int)(wi >> 24) * nt;
            const int n = L.nbr_count[k];
            const unsigned int* __restrict__ row = L.nbr + (size_t)k * L.stride;
#pragma unroll 2
            for (int t = s; t < n; t += T) {
                const unsigned int e = __ldg(&row[t]);
                const uint4 pj = __ldg(&spos[(int)(e & 0xFFFFFFu)]);
                float dx, dy, dz;
                delta<PER>(pi, pj, P.fscale, dx, dy, dz);
                const float r2 = norm2(dx, dy, dz);
                if (r2 < rc2 && r2 > 0.f) {
                    const float ir = fast_rsqrt(r2);
                    const float fr = -well_dedr(__ldg(&tab[trow + (int)(e >> 24)]), r2 * ir) * ir;
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
        if (active && s == 0) add_force_fixed(fbuf, P.n, (int)(wi & 0xFFFFFFu), fx, fy, fz);
    }
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

// ---- geometry with gradients (double precision), shared by base stacking and the donor x acceptor terms ----
// angle at the vertex between a = p_i - p_j and b = p_k - p_j; gi = d(theta)/d(p_i), gk = d(theta)/d(p_k)
__device__ __forceinline__ double angle_grad(D3 a, D3 b, D3& gi, D3& gk) {
    const double la = sqrt(d3dot(a, a)), lb = sqrt(d3dot(b, b));
    double cs = d3dot(a, b) / (la * lb);
    cs = fmax(-1.0, fmin(1.0, cs));
    const double sn = sqrt(fmax(1.0 - cs * cs, 1e-24));
    const double ila = 1.0 / la, ilb = 1.0 / lb, ci = -1.0 / (sn * la), ck = -1.0 / (sn * lb);
    gi = {ci * (b.x * ilb - cs * a.x * ila), ci * (b.y * ilb - cs * a.y * ila), ci * (b.z * ilb - cs * a.z * ila)};
    gk = {ck * (a.x * ila - cs * b.x * ilb), ck * (a.y * ila - cs * b.y * ilb), ck * (a.z * ila - cs * b.z * ilb)};
    return acos(cs);
}
// angle between two free vectors U and W with the cosine clamped to +-lim (the reference's
// "acos(min(max(sin t1 sin t2 cos phi - cos t1 cos t2, -lim), lim))" is exactly the angle between the two
// bond vectors); gU = d(theta)/dU, gW = d(theta)/dW, zero where the clamp is active
__device__ __forceinline__ double vector_angle_grad(D3 U, D3 W, double lim, D3& gU, D3& gW) {
    const double lu = sqrt(d3dot(U, U)), lw = sqrt(d3dot(W, W));
    const double cs = d3dot(U, W) / (lu * lw);
    if (cs >= lim || cs <= -lim) {
        gU = {0, 0, 0}; gW = {0, 0, 0};
        return acos(cs >= lim ? lim : -lim);
    }
    const double sn = sqrt(1.0 - cs * cs);
    const double ilu = 1.0 / lu, ilw = 1.0 / lw, cu = -1.0 / (sn * lu), cw = -1.0 / (sn * lw);
    gU = {cu * (W.x * ilw - cs * U.x * ilu), cu * (W.y * ilw - cs * U.y * ilu), cu * (W.z * ilw - cs * U.z * ilu)};
    gW = {cw * (U.x * ilu - cs * W.x * ilw), cw * (U.y * ilu - cs * W.y * ilw), cw * (U.z * ilu - cs * W.z * ilw)};
    return acos(cs);
}
// dihedral(p1,p2,p3,p4) with OpenMM's sign rule, F = p1-p2, G = p2-p3, H = p4-p3; g1..g4 = d(phi)/d(p1..p4)
__device__ __forceinline__ double dihedral_grad(D3 Fv, D3 G, D3 H, D3& g1, D3& g2, D3& g3, D3& g4) {
    const D3 A = d3cross(Fv, G), Bv = d3cross(H, G);
    const double A2 = d3dot(A, A), B2 = d3dot(Bv, Bv), lG = sqrt(d3dot(G, G));
    double cs = d3dot(A, Bv) / sqrt(A2 * B2);
    cs = fmax(-1.0, fmin(1.0, cs));
    double phi = acos(cs);
    if (d3dot(Fv, Bv) > 0.0) phi = -phi;
    const double FG = d3dot(Fv, G), HG = d3dot(H, G);
    const double ca = lG / A2, cb = lG / B2, ea = FG / (A2 * lG), eb = HG / (B2 * lG);
    g1 = {-ca * A.x, -ca * A.y, -ca * A.z};
    g4 = {cb * Bv.x, cb * Bv.y, cb * Bv.z};
    g2 = {(ca + ea) * A.x - eb * Bv.x, (ca + ea) * A.y - eb * Bv.y, (ca + ea) * A.z - eb * Bv.z};
    g3 = {-(cb - eb) * Bv.x - ea * A.x, -(cb - eb) * Bv.y - ea * A.y, -(cb - eb) * Bv.z - ea * A.z};
    return phi;
}
// f(x) = 1 for |x| <= pi/2, sin^2 x for |x| <= pi, 0 beyond ("max(f*rect0, rect1)" of the 3SPN.2 / PDNS strings)
__device__ __forceinline__ void modulation(double xx, double& f, double& df) {
    const double PI = 3.1415926535897932385, ax = fabs(xx);
    if (ax <= 0.5 * PI) { f = 1.0; df = 0.0; }
    else if (ax <= PI) { double sn, cs; sincos(xx, &sn, &cs); f = sn * sn; df = 2.0 * sn * cs; }
    else { f = 0.0; df = 0.0; }
}
// rep = eps (1-e^{-a dr})^2 [dr < 0];  attr = eps (1-e^{-a dr})^2 [dr >= 0] - eps;  and d/d(dr)
__device__ __forceinline__ void morse_terms(double eps, double alpha, double dr, double& rep, double& drep, double& attr, double& dattr) {
    const double ex = exp(-alpha * dr), om = 1.0 - ex;
    const double mval = eps * om * om, dm = 2.0 * eps * alpha * ex * om;
    if (dr < 0.0) { rep = mval; drep = dm; attr = -eps; dattr = 0.0; }
    else { rep = 0.0; drep = 0.0; attr = mval - eps; dattr = dm; }
}
__device__ __forceinline__ void d3axpy(D3& y, double a, const D3& v) { y.x += a * v.x; y.y += a * v.y; y.z += a * v.z; }

// one donor x acceptor combination: energy, and forces into fbuf when it is non-null
__device__ __forceinline__ double da_eval(const DATermDev& T, int d, int a, const StepParams& P, const double* __restrict__ x,
                                          unsigned long long* __restrict__ fbuf) {
    const bool per = T.periodic != 0;
    const int d1 = T.donors[3 * d], d2 = T.donors[3 * d + 1], d3 = T.donors[3 * d + 2];
    const int a1 = T.acceptors[3 * a], a2 = T.acceptors[3 * a + 1], a3 = T.acceptors[3 * a + 2];
    D3 fd1 = {0, 0, 0}, fd2 = {0, 0, 0}, fd3 = {0, 0, 0}, fa1 = {0, 0, 0}, fa2 = {0, 0, 0}, fa3 = {0, 0, 0};
    double E;
    if (T.kind == 3) {
        // combinatorial Go contact, one bead per donor / acceptor; g0 = k, g1 = r0     CombinatorialGoContact...cpp:5-10
        const D3 rv = d3sub(x, a1, d1, per, P);                 // a1 - d1
        const double r2 = d3dot(rv, rv), s2 = T.g1 * T.g1 / r2, s4 = s2 * s2, s10 = s4 * s4 * s2, s12 = s10 * s2;
        E = T.g0 * (5.0 * s12 - 6.0 * s10);
        if (fbuf) {
            const double sr = T.g0 * 60.0 * (s10 - s12) / r2;   // (dE/dr) / r
            add_force_fixed_d(fbuf, P.n, a1, -sr * rv.x, -sr * rv.y, -sr * rv.z);
            add_force_fixed_d(fbuf, P.n, d1, sr * rv.x, sr * rv.y, sr * rv.z);
        }
        return E;
    }
    if (T.kind == 0) {
        // 3SPN.2 base pair, per-donor eps, r0, theta0_1, theta0_2, phi0; g0 = alpha, g1 = K_BP
        // ThreeSPN2BasePairForceFieldGenerator.cpp:8-26
        const double* p = T.par + (size_t)d * T.npar;
        const D3 rv = d3sub(x, d1, a1, per, P);                 // d1 - a1
        const D3 sd = d3sub(x, d2, d1, per, P);                 // d2 - d1
        const D3 sa = d3sub(x, a2, a1, per, P);                 // a2 - a1
        const D3 nrv = {-rv.x, -rv.y, -rv.z};                   // a1 - d1
        const double r = sqrt(d3dot(rv, rv));
        D3 g1i, g1k, g2i, g2k, h1, h2, h3, h4;
        const double th1 = angle_grad(sd, nrv, g1i, g1k);       // angle(d2, d1, a1): vertex d1
        const double th2 = angle_grad(sa, rv, g2i, g2k);        // angle(a2, a1, d1): vertex a1
        const double phi = dihedral_grad(sd, rv, sa, h1, h2, h3, h4);  // dihedral(d2, d1, a1, a2)
        double rep, drep, attr, dattr, f1, df1, f2, df2;
        morse_terms(p[0], T.g0, r - p[1], rep, drep, attr, dattr);
        modulation(T.g1 * (th1 - p[2]), f1, df1);
        modulation(T.g1 * (th2 - p[3]), f2, df2);
        double sphi, cphi;
        sincos(phi - p[4], &sphi, &cphi);
        const double hc = 0.5 * (1.0 + cphi);
        E = rep + hc * f1 * f2 * attr;
        if (fbuf) {
            const double dEdr = drep + hc * f1 * f2 * dattr;
            const double dE1 = hc * df1 * T.g1 * f2 * attr, dE2 = hc * f1 * df2 * T.g1 * attr, dEp = -0.5 * sphi * f1 * f2 * attr;
            const double sr = dEdr / r;
            d3axpy(fd1, -sr, rv); d3axpy(fa1, sr, rv);
            // theta1: i = d2, vertex d1, k = a1
            d3axpy(fd2, -dE1, g1i); d3axpy(fa1, -dE1, g1k); d3axpy(fd1, dE1, g1i); d3axpy(fd1, dE1, g1k);
            // theta2: i = a2, vertex a1, k = d1
            d3axpy(fa2, -dE2, g2i); d3axpy(fd1, -dE2, g2k); d3axpy(fa1, dE2, g2i); d3axpy(fa1, dE2, g2k);
            d3axpy(fd2, -dEp, h1); d3axpy(fd1, -dEp, h2); d3axpy(fa1, -dEp, h3); d3axpy(fa2, -dEp, h4);
        }
    } else if (T.kind == 1) {
        // 3SPN.2 cross stacking, per-acceptor eps, r0, theta0_3, theta0_CS; g0 = K_BP, g1 = K_CS, g2 = alpha
        // ThreeSPN2CrossStackingForceFieldGenerator.cpp:8-31
        const double* p = T.par + (size_t)a * T.npar;
        const D3 rv = d3sub(x, d1, a3, per, P);                 // d1 - a3
        const D3 sd = d3sub(x, d2, d1, per, P);                 // d2 - d1
        const D3 sa = d3sub(x, a2, a1, per, P);                 // a2 - a1
        const D3 nrv = {-rv.x, -rv.y, -rv.z};                   // a3 - d1
        const double r = sqrt(d3dot(rv, rv));
        D3 gci, gck, gU, gW;
        const double tcs = angle_grad(sd, nrv, gci, gck);       // angle(d2, d1, a3): vertex d1
        const double t3 = vector_angle_grad(sd, sa, 0.99, gU, gW);
        double rep, drep, attr, dattr, f3, df3, fc, dfc;
        morse_terms(p[0], T.g2, r - p[1], rep, drep, attr, dattr);
        modulation(T.g0 * (t3 - p[2]), f3, df3);
        modulation(T.g1 * (tcs - p[3]), fc, dfc);
        E = 0.5 * f3 * fc * attr;
        if (fbuf) {
            const double dEdr = 0.5 * f3 * fc * dattr, dE3 = 0.5 * df3 * T.g0 * fc * attr, dEc = 0.5 * f3 * dfc * T.g1 * attr;
            const double sr = dEdr / r;
            d3axpy(fd1, -sr, rv); d3axpy(fa3, sr, rv);
            d3axpy(fd2, -dEc, gci); d3axpy(fa3, -dEc, gck); d3axpy(fd1, dEc, gci); d3axpy(fd1, dEc, gck);
            d3axpy(fd2, -dE3, gU); d3axpy(fd1, dE3, gU); d3axpy(fa2, -dE3, gW); d3axpy(fa1, dE3, gW);
        }
    } else {
        // PDNS, per-donor k, r0, theta0, phi0; g0 = sigma, g1 = delta   ProteinDNANonSpecificForceFieldGenerator.cpp:60-83
        const double* p = T.par + (size_t)d * T.npar;
        const double K = 3.1415926535897932385 / (2.0 * T.g1);
        const D3 rv = d3sub(x, d2, a1, per, P);                 // d2 - a1  (CA - P)
        const D3 uu = d3sub(x, d1, d3, per, P);                 // d1 - d3  (CA_N - CA_C)
        const D3 sa = d3sub(x, a2, a1, per, P);                 // a2 - a1  (S3' - P)
        const double r = sqrt(d3dot(rv, rv));
        D3 gU, gW, gpi, gpk;
        const double th = vector_angle_grad(uu, rv, 0.99999, gU, gW);
        const double ph = angle_grad(rv, sa, gpi, gpk);         // angle(d2, a1, a2): vertex a1
        const double dr = r - p[1], is2 = 1.0 / (T.g0 * T.g0);
        const double fr = exp(-0.5 * dr * dr * is2);
        double gt, dgt, gp, dgp;
        modulation(K * (th - p[2]), gt, dgt);
        modulation(K * (ph - p[3]), gp, dgp);
        E = -p[0] * fr * gt * gp;
        if (fbuf) {
            const double dEdr = p[0] * dr * is2 * fr * gt * gp, dEt = -p[0] * fr * dgt * K * gp, dEp = -p[0] * fr * gt * dgp * K;
            const double sr = dEdr / r;
            d3axpy(fd2, -sr, rv); d3axpy(fa1, sr, rv);
            d3axpy(fd1, -dEt, gU); d3axpy(fd3, dEt, gU); d3axpy(fd2, -dEt, gW); d3axpy(fa1, dEt, gW);
            d3axpy(fd2, -dEp, gpi); d3axpy(fa2, -dEp, gpk); d3axpy(fa1, dEp, gpi); d3axpy(fa1, dEp, gpk);
        }
    }
    if (fbuf) {
        add_force_fixed_d(fbuf, P.n, d1, fd1.x, fd1.y, fd1.z);
        add_force_fixed_d(fbuf, P.n, d2, fd2.x, fd2.y, fd2.z);
        add_force_fixed_d(fbuf, P.n, a1, fa1.x, fa1.y, fa1.z);
        add_force_fixed_d(fbuf, P.n, a2, fa2.x, fa2.y, fa2.z);
        if (T.kind == 2) add_force_fixed_d(fbuf, P.n, d3, fd3.x, fd3.y, fd3.z);
        if (T.kind == 1) add_force_fixed_d(fbuf, P.n, a3, fa3.x, fa3.y, fa3.z);
    }
    return E;
}

// canonical donor-acceptor distance test, identical (operation for operation) to the oracle's for_each_da
__device__ __forceinline__ double da_r2(const DATermDev& T, int d, int a, const StepParams& P, const double* __restrict__ x) {
    const int i = T.donors[3 * d], j = T.acceptors[3 * a];
    double dx = __dsub_rn(x[3 * i], x[3 * j]), dy = __dsub_rn(x[3 * i + 1], x[3 * j + 1]), dz = __dsub_rn(x[3 * i + 2], x[3 * j + 2]);
    if (T.periodic) {
        dx = __dsub_rn(dx, __dmul_rn(P.box[0], nearbyint(__ddiv_rn(dx, P.box[0]))));
        dy = __dsub_rn(dy, __dmul_rn(P.box[1], nearbyint(__ddiv_rn(dy, P.box[1]))));
        dz = __dsub_rn(dz, __dmul_rn(P.box[2], nearbyint(__ddiv_rn(dz, P.box[2]))));
    }
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// Verlet list of donor x acceptor combinations inside cutoff + skin (one thread per candidate)
__global__ void da_build_kernel(DAListDev D, StepParams P, const double* __restrict__ x, int* __restrict__ flags) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= D.ncand) return;
    int t = 0;
    while (t + 1 < D.nterms && c >= D.terms[t + 1].cand_begin) t++;
    const DATermDev& T = D.terms[t];
    const long long local = c - T.cand_begin;
    const int d = (int)(local / T.na), a = (int)(local % T.na);
    if (!(da_r2(T, d, a, P, x) < T.rlist2)) return;
    const unsigned long long key = (unsigned long long)d * (unsigned long long)T.na + (unsigned long long)a;
    int lo = 0, hi = T.nexcl;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (T.excl[mid] < key) lo = mid + 1; else hi = mid; }
    if (lo < T.nexcl && T.excl[lo] == key) return;
    const unsigned int pos = atomicAdd(D.count, 1u);
    if (pos < (unsigned int)D.capacity) D.entries[pos] = make_uint2(((unsigned int)t << 24) | (unsigned int)d, (unsigned int)a);
    else flags[FLAG_DA_OVERFLOW] = 1;
}

// (donor, acceptor) combinations of term `term` inside the interaction cutoff at the current positions
__global__ void da_extract_kernel(DAListDev D, int term, StepParams P, const double* __restrict__ x, int* __restrict__ out,
                                  long long capacity, unsigned long long* __restrict__ counter) {
    const unsigned int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= min(*D.count, (unsigned int)D.capacity)) return;
    const uint2 e = D.entries[m];
    if ((int)(e.x >> 24) != term) return;
    const DATermDev& T = D.terms[term];
    const int d = (int)(e.x & 0xFFFFFFu), a = (int)e.y;
    if (da_r2(T, d, a, P, x) < T.rc2) {
        const unsigned long long pos = atomicAdd(counter, 1ull);
        if (out && (long long)pos < capacity) { out[2 * pos] = d; out[2 * pos + 1] = a; }
    }
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
        case 2: {  // GoContact     p = k, r0, slope, intercept, cap   GoContactForceFieldGenerator.cpp:5-7,
                   // CappedGoContactForceFieldGenerator.cpp:5-9 (cap = 0: plain contact; step(cap - r/r0) selects the line)
            const double u = q / p[1];
            if (p[4] - u >= 0.0) {
                E = p[0] * (p[2] * u + p[3]); dE = p[0] * p[2] / p[1];
            } else {
                const double s = p[1] / q, s2 = s * s, s4 = s2 * s2, s10 = s4 * s4 * s2, s12 = s10 * s2;
                E = p[0] * (5.0 * s12 - 6.0 * s10); dE = p[0] * 60.0 * (s10 - s12) / q;
            }
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

// one-bead external terms on the absolute (unwrapped) master position; p[0] = form, rows as in host_system.hpp
__device__ __forceinline__ double external_term(const double* __restrict__ p, int i, bool per, const StepParams& P,
                                                const double* __restrict__ x, unsigned long long* __restrict__ fbuf) {
    const double px = x[3 * i], py = x[3 * i + 1], pz = x[3 * i + 2];
    const int form = (int)p[0];
    double E, gx, gy, gz;  // energy and its gradient
    if (form == 0) {
        // PositionRestraint: k (|x - x0| - v0)^2     PositionRestraintForceFieldGenerator.cpp:8-10 (the periodic string is
        // never used, :12-17)
        const double dx = px - p[1], dy = py - p[2], dz = pz - p[3];
        const double r = sqrt(dx * dx + dy * dy + dz * dz), d = r - p[5];
        E = p[4] * d * d;
        const double s = (r > 0.0) ? 2.0 * p[4] * d / r : 0.0;
        gx = s * dx; gy = s * dy; gz = s * dz;
    } else if (form == 1) {
        // Pulling: -kx x - ky y - kz z               PullingForceFieldGenerator.cpp:5-7
        E = -(p[1] * px + p[2] * py + p[3] * pz);
        gx = -p[1]; gy = -p[2]; gz = -p[3];
    } else if (form == 2) {
        // EXVRectangularBox: eps sum (radius / (x - lo))^12 + (radius / (hi - x))^12     EXVRectangularBox...cpp:33-38
        const double eps = p[1], rad = p[8];   // row: eps lo[3] hi[3] radius
        const double q[3] = {px, py, pz};
        double g[3];
        E = 0.0;
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const double il = 1.0 / (q[a] - p[2 + a]), iu = 1.0 / (p[5 + a] - q[a]);
            const double sl = rad * il, su = rad * iu;
            const double sl2 = sl * sl, su2 = su * su, sl6 = sl2 * sl2 * sl2, su6 = su2 * su2 * su2;
            const double el = sl6 * sl6, eu = su6 * su6;
            E += eps * (el + eu);
            g[a] = 12.0 * eps * (eu * iu - el * il);
        }
        gx = g[0]; gy = g[1]; gz = g[2];
    } else {
        // CylindricalRestraint: k (dr - v0)^2, dr = |x - h|, h = (axis . x) axis + shift     Cylindrical...cpp:10-26
        // x - h = (1 - a a^T) x - shift; its derivative with respect to x is the (symmetric) matrix 1 - a a^T
        const double ax = p[1], ay = p[2], az = p[3];
        const double ra = ax * px + ay * py + az * pz;
        double dx = px - (ra * ax + p[4]), dy = py - (ra * ay + p[5]), dz = pz - (ra * az + p[6]);
        if (per) {
            dx -= P.box[0] * rint(dx * P.inv_box[0]);
            dy -= P.box[1] * rint(dy * P.inv_box[1]);
            dz -= P.box[2] * rint(dz * P.inv_box[2]);
        }
        const double r = sqrt(dx * dx + dy * dy + dz * dz), d = r - p[8];
        E = p[7] * d * d;
        const double s = (r > 0.0) ? 2.0 * p[7] * d / r : 0.0;
        const double da = ax * dx + ay * dy + az * dz;
        gx = s * (dx - da * ax); gy = s * (dy - da * ay); gz = s * (dz - da * az);
    }
    if (fbuf) add_force_fixed_d(fbuf, P.n, i, -gx, -gy, -gz);
    return E;
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
    if (kind == 9) {
        // 3SPN.2 base stacking, p = eps, r0, theta0, alpha, K_BS   ThreeSPN2BaseStackingForceFieldGenerator.cpp:6-21
        const int i = B.idx[3 * m], j = B.idx[3 * m + 1], k = B.idx[3 * m + 2];
        const D3 a = d3sub(x, i, j, per, P), b = d3sub(x, k, j, per, P);
        D3 gi, gk;
        const double lb = sqrt(d3dot(b, b));
        const double th = angle_grad(a, b, gi, gk);
        double rep, drep, attr, dattr, f, df;
        morse_terms(p[0], p[3], lb - p[1], rep, drep, attr, dattr);
        modulation(p[4] * (th - p[2]), f, df);
        E = rep + f * attr;
        if (fbuf) {
            const double dEdr = drep + f * dattr, dEdth = df * p[4] * attr;
            const double sr = dEdr / lb;  // d|b|/dp_k = b/|b|, d|b|/dp_j = -b/|b|
            const D3 fi = {-dEdth * gi.x, -dEdth * gi.y, -dEdth * gi.z};
            const D3 fk = {-dEdth * gk.x - sr * b.x, -dEdth * gk.y - sr * b.y, -dEdth * gk.z - sr * b.z};
            add_force_fixed_d(fbuf, P.n, i, fi.x, fi.y, fi.z);
            add_force_fixed_d(fbuf, P.n, k, fk.x, fk.y, fk.z);
            add_force_fixed_d(fbuf, P.n, j, -fi.x - fk.x, -fi.y - fk.y, -fi.z - fk.z);
        }
        return E;
    }
    if (kind == 10) return external_term(p, B.idx[m], per, P, x, fbuf);
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
// the fused per-step force launch: every pair list and every bonded kind in one grid (block ranges map to
// segments).  Donor x acceptor terms, when a system has them, run in local_kernel on a forked graph branch.
template <bool PER>
__global__ void __launch_bounds__(kBlock, OCM_FORCE_MINBLOCKS) force_kernel(const WorkSeg* __restrict__ segs, int nsegs,
                                                       const PairListDev* __restrict__ lists,
                                                       const BondedDev* __restrict__ bonded, StepParams P,
                                                       const double* __restrict__ x, unsigned long long* __restrict__ fbuf) {
    extern __shared__ __align__(16) float smem[];
    if (P.pdl_early) pdl_trigger();
    pdl_wait();
    int s = 0;
    while (s + 1 < nsegs && (int)blockIdx.x >= segs[s + 1].block_begin) s++;
    const WorkSeg sg = segs[s];
    const int blk = blockIdx.x - sg.block_begin;
    if (sg.kind == SEG_PAIR_TZ) {
        const PairListDev& L = lists[sg.id];
        // (under a slab decomposition the grid covers L.row_capacity rows, the rows in use at the last sizing pass plus
        // head-room; build_list_kernel raises the overflow flag, i.e. a re-sizing pass, if more rows ever exist)
        switch (sg.tpb) {
            case 4: pair_tz_block<PER, 4>(L, blk, sg.chunk, P, fbuf, smem); break;
            case 8: pair_tz_block<PER, 8>(L, blk, sg.chunk, P, fbuf, smem); break;
            case 16: pair_tz_block<PER, 16>(L, blk, sg.chunk, P, fbuf, smem); break;
            default: pair_tz_block<PER, 32>(L, blk, sg.chunk, P, fbuf, smem);
        }
    } else if (sg.kind == SEG_PAIR_DH) {
        const PairListDev& L = lists[sg.id];
        switch (sg.tpb) {
            case 4: pair_dh_block<PER, 4>(L, blk, sg.chunk, P, fbuf); break;
            case 8: pair_dh_block<PER, 8>(L, blk, sg.chunk, P, fbuf); break;
            case 16: pair_dh_block<PER, 16>(L, blk, sg.chunk, P, fbuf); break;
            default: pair_dh_block<PER, 32>(L, blk, sg.chunk, P, fbuf);
        }
    } else if (sg.kind == SEG_PAIR_WELL) {
        const PairListDev& L = lists[sg.id];
        switch (sg.tpb) {
            case 4: pair_well_block<PER, 4>(L, blk, sg.chunk, P, fbuf); break;
            case 8: pair_well_block<PER, 8>(L, blk, sg.chunk, P, fbuf); break;
            case 16: pair_well_block<PER, 16>(L, blk, sg.chunk, P, fbuf); break;
            default: pair_well_block<PER, 32>(L, blk, sg.chunk, P, fbuf);
        }
    } else {
        const BondedDev& B = bonded[sg.id];
        int m = blk * kBlock + threadIdx.x;
        // slab decomposition: a term is evaluated by every rank that owns one of its beads (the forces on the other beads
        // are discarded with the accumulator); those terms were compacted at the last rebuild
        if (B.dd_terms) m = (m < *B.dd_count) ? B.dd_terms[m] : B.n;
        if (m < B.n) {
            int g;
            bonded_term(sg.id, B, m, P, x, fbuf, &g);
        }
    }
}

// HarmonicCoMPulling (HarmonicCoMPullingForceFieldGenerator.cpp:5-18): E = 1/2 k (|c1 - c2| - v0)^2.  One block per term: both
// centres are reduced in a fixed order (lane strides, xor tree, warp partials summed 0..7), so the result does not depend on
// the schedule; every bead then takes its mass share of the spring force.  egroup != null: the energy only.
__global__ void __launch_bounds__(kComBlock) com_kernel(const ComDev* __restrict__ terms, StepParams P, const double* __restrict__ x,
                                                         unsigned long long* __restrict__ fbuf, double* __restrict__ egroup) {
    const ComDev T = terms[blockIdx.x];
    __shared__ double part[6][kComBlock / 32];
    double c1x = 0, c1y = 0, c1z = 0, c2x = 0, c2y = 0, c2z = 0;
    const int nt = T.n1 + T.n2;
    for (int m = threadIdx.x; m < nt; m += kComBlock) {
        const int i = T.idx[m];
        const double w = T.w[m], wx = w * x[3 * i], wy = w * x[3 * i + 1], wz = w * x[3 * i + 2];
        if (m < T.n1) { c1x += wx; c1y += wy; c1z += wz; } else { c2x += wx; c2y += wy; c2z += wz; }
    }
    double v[6] = {c1x, c1y, c1z, c2x, c2y, c2z};
#pragma unroll
    for (int q = 0; q < 6; q++) {
        for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
        if ((threadIdx.x & 31) == 0) part[q][threadIdx.x >> 5] = v[q];
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 6; q++) {
        double s = 0;
        for (int wv = 0; wv < kComBlock / 32; wv++) s += part[q][wv];
        v[q] = s;
    }
    double dx = v[0] - v[3], dy = v[1] - v[4], dz = v[2] - v[5];
    if (T.periodic) {
        dx -= P.box[0] * rint(dx * P.inv_box[0]);
        dy -= P.box[1] * rint(dy * P.inv_box[1]);
        dz -= P.box[2] * rint(dz * P.inv_box[2]);
    }
    const double r = sqrt(dx * dx + dy * dy + dz * dz), d = r - T.v0;
    if (egroup) {
        if (threadIdx.x == 0) atomicAdd(&egroup[T.group], 0.5 * T.k * d * d);
        return;
    }
    const double s = (r > 0.0) ? T.k * d / r : 0.0;   // (dE/dr) / r
    for (int m = threadIdx.x; m < nt; m += kComBlock) {
        const double f = (m < T.n1 ? -s : s) * T.w[m];
        add_force_fixed_d(fbuf, P.n, T.idx[m], f * dx, f * dy, f * dz);
    }
}

// donor x acceptor terms (double precision, one thread per listed combination).  Launched beside force_kernel on
// a forked graph branch, so the pair kernel keeps its own tight register budget and the two overlap.
__global__ void __launch_bounds__(kBlock) local_kernel(DAListDev da, StepParams P, const double* __restrict__ x,
                                                       unsigned long long* __restrict__ fbuf) {
    const unsigned int m = blockIdx.x * kBlock + threadIdx.x;
    if (m < min(*da.count, (unsigned int)da.capacity)) {
        const uint2 e = da.entries[m];
        const DATermDev& T = da.terms[e.x >> 24];
        const int d = (int)(e.x & 0xFFFFFFu), a = (int)e.y;
        if (da_r2(T, d, a, P, x) < T.rc2) da_eval(T, d, a, P, x, fbuf);
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
// EARLY: every independent load of the bead is issued before the first use — one memory round trip instead of three.  Small
// systems are latency-bound here (C2, 16 300 beads: 6.2 -> 4.6 us per launch); at a million beads the extra registers cost
// occupancy and the launch gets slower (0.099 -> 0.110 ms), so the engine picks by size.
template <bool EARLY>
__global__ void integrate_kernel(StepParams P, double* __restrict__ x, float4* __restrict__ v,
                                 unsigned long long* __restrict__ fbuf, uint4* __restrict__ posq,
                                 const PairListDev* __restrict__ lists, int nlists,
                                 const double* __restrict__ xref, unsigned long long* __restrict__ step_base,
                                 int step_offset, int window, int last_in_window, int* __restrict__ flags) {
    if (P.pdl_early) pdl_trigger();
    pdl_wait();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    // slab decomposition: one thread per bead PRESENT on this rank (compacted at the last rebuild), not per bead of the system
    if (P.present) {
        const int np = *P.n_present;
        if (i == 0 && np > (int)(gridDim.x * blockDim.x)) flags[FLAG_OVERFLOW] = 1;   // more beads than the launch covers: re-size
        i = (i < np) ? P.present[i] : P.n;
    }
    const unsigned long long step = *step_base + (unsigned long long)step_offset;
    float d2 = 0.f;
    // slab decomposition: beads that are not on this rank cost one byte; halo beads only have their accumulator cleared
    // (bonded terms evaluated here add to it); owned beads move
    const int role_i = (i < P.n && P.role) ? (int)P.role[i] : 1;
    if (i < P.n && role_i == 2) { fbuf[i] = 0ull; fbuf[P.n + i] = 0ull; fbuf[2 * P.n + i] = 0ull; }
    if (i < P.n && role_i == 1) {
        const int n = P.n;
        const float inv32 = 2.3283064365386963e-10f;
        const unsigned long long q0 = fbuf[i], q1 = fbuf[n + i], q2 = fbuf[2 * n + i];
        float4 vi;
        double x0, x1, x2, r0, r1, r2;
        int kr[4] = {-1, -1, -1, -1};
        if (EARLY) {
            vi = v[i];
            x0 = x[3 * i]; x1 = x[3 * i + 1]; x2 = x[3 * i + 2];
            r0 = xref[3 * i]; r1 = xref[3 * i + 1]; r2 = xref[3 * i + 2];
#pragma unroll
            for (int l = 0; l < 4; l++) if (l < nlists) kr[l] = lists[l].rank_of_bead[i];
        }
        const float fx = (float)(long long)q0 * inv32, fy = (float)(long long)q1 * inv32, fz = (float)(long long)q2 * inv32;
        fbuf[i] = 0ull; fbuf[n + i] = 0ull; fbuf[2 * n + i] = 0ull;
        if (!EARLY) vi = v[i];
        const float im = vi.w;
        if (im > 0.f) {
            float g0, g1, g2;
            const int rep = (P.nrep > 1) ? i / P.nper : 0;
            const uint2 key = P.rep_seed ? P.rep_seed[rep] : make_uint2(P.seed_lo, P.seed_hi);
            normals3(philox4x32_10(make_uint4((unsigned int)(i - rep * P.nper), (unsigned int)step, (unsigned int)(step >> 32), 0u), key),
                     g0, g1, g2);
            const float cs = P.c * sqrtf(im), bm = P.b * im;
            vi.x = P.a * vi.x + bm * fx + cs * g0;
            vi.y = P.a * vi.y + bm * fy + cs * g1;
            vi.z = P.a * vi.z + bm * fz + cs * g2;
            v[i] = vi;
            if (!EARLY) { x0 = x[3 * i]; x1 = x[3 * i + 1]; x2 = x[3 * i + 2]; }
            const double px = x0 + (double)vi.x * P.dt, py = x1 + (double)vi.y * P.dt, pz = x2 + (double)vi.z * P.dt;
            x[3 * i] = px; x[3 * i + 1] = py; x[3 * i + 2] = pz;
            uint4 q;
            if (P.periodic) { q.x = to_fix(px, P.inv_box[0]); q.y = to_fix(py, P.inv_box[1]); q.z = to_fix(pz, P.inv_box[2]); }
            else {
                double lx = px, ly = py, lz = pz;
                if (P.origin) { const double* o = P.origin + 3 * rep; lx -= o[0]; ly -= o[1]; lz -= o[2]; }
                q.x = __float_as_uint((float)lx); q.y = __float_as_uint((float)ly); q.z = __float_as_uint((float)lz);
            }
            q.w = 0;
            posq[i] = q;
            // refresh the cell-sorted copies the pair kernels read; w (bead | type, or the mixing parameter of a tile list) is not
            // touched: an 8-byte and a 4-byte store instead of a load - modify - store of 16 bytes (one dependent L2 round trip less)
            if (EARLY) {
#pragma unroll
                for (int l = 0; l < 4; l++) {
                    if (l < nlists && kr[l] >= 0) {
                        uint4* sp = lists[l].spos + kr[l];
                        *reinterpret_cast<uint2*>(sp) = make_uint2(q.x, q.y);
                        sp->z = q.z;
                    }
                }
            }
            for (int l = EARLY ? 4 : 0; l < nlists; l++) {
                const int k = lists[l].rank_of_bead[i];
                if (k >= 0) {
                    uint4* sp = lists[l].spos + k;
                    *reinterpret_cast<uint2*>(sp) = make_uint2(q.x, q.y);
                    sp->z = q.z;
                }
            }
            if (!EARLY) { r0 = xref[3 * i]; r1 = xref[3 * i + 1]; r2 = xref[3 * i + 2]; }
            const float ex = (float)(px - r0), ey = (float)(py - r1), ez = (float)(pz - r2);
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
__global__ void init_velocity_kernel(int n, int nper, float kT, unsigned long long seed, float4* __restrict__ v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 vi = v[i];
    float g0, g1, g2;
    // replica r of a batch draws from seed + r with its local bead index: the same numbers as a standalone run
    const int rep = i / nper;
    const unsigned long long sd = seed + (unsigned long long)rep;
    normals3(philox4x32_10(make_uint4((unsigned int)(i - rep * nper), 0u, 0u, 1u), make_uint2((unsigned int)sd, (unsigned int)(sd >> 32))), g0, g1, g2);
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
                                                        const BondedDev* __restrict__ bonded, DAListDev da, StepParams P,
                                                        const double* __restrict__ x, double* __restrict__ egroup) {
    int s = 0;
    while (s + 1 < nsegs && (int)blockIdx.x >= segs[s + 1].block_begin) s++;
    const WorkSeg sg = segs[s];
    const int blk = blockIdx.x - sg.block_begin;
    if (sg.kind == SEG_BONDED || sg.kind == SEG_DA) {
        double E = 0.0;
        int g = -1;
        if (sg.kind == SEG_BONDED) {
            const BondedDev& B = bonded[sg.id];
            int m = blk * kBlock + threadIdx.x;
            if (B.dd_terms) m = (m < *B.dd_count) ? B.dd_terms[m] : B.n;
            // slab decomposition: the rank owning the first bead of a term reports its energy
            if (m < B.n && (!P.role || P.role[B.idx[bonded_nbody_dev(sg.id) * m]] == 1)) E = bonded_term(sg.id, B, m, P, x, nullptr, &g);
        } else {
            const unsigned int m = blk * kBlock + threadIdx.x;
            if (m < min(*da.count, (unsigned int)da.capacity)) {
                const uint2 e = da.entries[m];
                const DATermDev& T = da.terms[e.x >> 24];
                const int d = (int)(e.x & 0xFFFFFFu), a = (int)e.y;
                if (da_r2(T, d, a, P, x) < T.rc2) { E = da_eval(T, d, a, P, x, nullptr); g = T.group; }
            }
        }
        if (P.nrep > 1) {
            // replica batch: energies are kept per replica (40 slots each); bonded terms never span replicas
            if (g >= 0 && sg.kind == SEG_BONDED) {
                const BondedDev& B = bonded[sg.id];
                const int nbody = bonded_nbody_dev(sg.id);
                const int rep = B.idx[nbody * (blk * kBlock + threadIdx.x)] / P.nper;
                atomicAdd(&egroup[rep * 40 + g], E);
            }
            return;
        }
        // terms handled by one block may belong to different groups: accumulate per distinct group
        for (int gg = 0; gg < 32; gg++) {
            const unsigned int any = __syncthreads_or(g == gg);
            if (any) block_add((g == gg) ? E : 0.0, &egroup[gg]);
        }
        return;
    }
    const PairListDev& L = lists[sg.id];
    const int T = sg.tpb, G = kBlock / T;
    const int g = threadIdx.x / T, sl = threadIdx.x % T;
    double Em[kMaxMembers] = {0, 0, 0, 0};
    const int k_lo = blk * sg.chunk, k_hi = min(list_rows(L), k_lo + sg.chunk);
    for (int k = k_lo + g; k < k_hi; k += G) {
        const uint4 pi = L.spos[k];
        const int nt = L.ntypes;
        const int trow = (int)(pi.w >> 24) * nt;
        const int n = L.nbr_count[k];
        const unsigned int* row = L.nbr + (size_t)k * L.stride;
        double Ekeep[kMaxMembers];
        if (P.nrep > 1) for (int m = 0; m < kMaxMembers; m++) { Ekeep[m] = Em[m]; Em[m] = 0.0; }
        for (int t = sl; t < n; t += T) {
            const unsigned int e = row[t];
            const uint4 pj = L.spos[e & 0xFFFFFFu];
            float dx, dy, dz;
            delta<PER>(pi, pj, P.fscale, dx, dy, dz);
            const float r2 = norm2(dx, dy, dz);
            const int ti = trow + (int)(e >> 24);
            if (sg.kind == SEG_PAIR_DH) {
                if (r2 < L.rc2) {
                    const float ir = rsqrtf(r2), r = r2 * ir;
                    Em[0] += (double)(__uint_as_float(pi.w) * __uint_as_float(pj.w) * (__expf(-r * L.inv_debye) * ir - L.dh_cc));
                }
            } else if (sg.kind == SEG_PAIR_WELL) {
                if (r2 < L.rc2 && r2 > 0.f) {
                    // {base, 1 / omega, amp, shape}: eps = amp (2 / pi) omega (shape 0), amp omega / 6 (shape 1)
                    const float4 c = L.tabF[ti];
                    const double u = ((double)sqrtf(r2) - (double)c.x) * (double)c.y;
                    const double eps = (c.w == 0.f) ? (double)c.z * 0.63661977236758134 / (double)c.y : (double)c.z / (6.0 * (double)c.y);
                    if (u < 0.0) Em[0] -= eps;
                    else if (u < 1.0) {
                        const double cs = cospi(0.5 * u);
                        Em[0] -= eps * ((c.w == 0.f) ? cs * cs : 1.0 + u * u * (2.0 * u - 3.0));
                    }
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
        if (P.nrep > 1) {
            // this bead's share goes to its replica's slots (full list: each pair is seen twice)
            const int rep = L.bead_of_rank[k] / P.nper;
            const int nmm = (sg.kind == SEG_PAIR_TZ) ? L.nmem : 1;
            for (int m = 0; m < nmm; m++) {
                if (Em[m] != 0.0) atomicAdd(&egroup[rep * 40 + L.mem_group[m]], 0.5 * Em[m]);
                Em[m] = Ekeep[m];
            }
        }
    }
    if (P.nrep > 1) return;
    const int nm = (sg.kind == SEG_PAIR_TZ) ? L.nmem : 1;
    for (int m = 0; m < nm; m++) block_add(0.5 * Em[m], &egroup[L.mem_group[m]]);  // full list: each pair twice
}

__global__ void kinetic_kernel(StepParams P, const float4* __restrict__ v, const unsigned long long* __restrict__ fbuf,
                               double* __restrict__ ke) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double e = 0.0;
    if (i < P.n) {
        const float4 vi = v[i];
        if (vi.w > 0.f && (!P.role || P.role[i] == 1)) {
            const double im = vi.w, h = 0.5 * P.dt * im, inv = 1.0 / kFixScale;
            const double vx = vi.x + h * (double)(long long)fbuf[i] * inv, vy = vi.y + h * (double)(long long)fbuf[P.n + i] * inv,
                         vz = vi.z + h * (double)(long long)fbuf[2 * P.n + i] * inv;
            e = 0.5 / im * (vx * vx + vy * vy + vz * vz);
        }
    }
    if (P.nrep > 1) { if (i < P.n) atomicAdd(&ke[(i / P.nper) * 40], e); }
    else block_add(e, ke);
}

// observer staging: what the host asked for, converted on the device into one block of doubles
// [positions 3n | velocities 3n | forces 3n | 32 group energies, kinetic energy, 7 spare], so that one copy into pinned
// host memory serves the frame (ocm_get_state_begin)
__global__ void pack_state_kernel(int n, const double* __restrict__ x, const float4* __restrict__ v, const unsigned long long* __restrict__ fbuf,
                                  double* __restrict__ out_x, double* __restrict__ out_v, double* __restrict__ out_f) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (out_x) { out_x[3 * i] = x[3 * i]; out_x[3 * i + 1] = x[3 * i + 1]; out_x[3 * i + 2] = x[3 * i + 2]; }
    if (out_v) { const float4 vi = v[i]; out_v[3 * i] = vi.x; out_v[3 * i + 1] = vi.y; out_v[3 * i + 2] = vi.z; }
    if (out_f) {
        const double inv = 1.0 / kFixScale;
        out_f[3 * i] = (double)(long long)fbuf[i] * inv; out_f[3 * i + 1] = (double)(long long)fbuf[n + i] * inv;
        out_f[3 * i + 2] = (double)(long long)fbuf[2 * (size_t)n + i] * inv;
    }
}
// ocm_set_state_async: staged doubles -> master positions, FP32 velocities (w = 1/m is kept)
__global__ void unpack_state_kernel(int n, const double* __restrict__ in_x, const double* __restrict__ in_v, double* __restrict__ x, float4* __restrict__ v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (in_x) { x[3 * i] = in_x[3 * i]; x[3 * i + 1] = in_x[3 * i + 1]; x[3 * i + 2] = in_x[3 * i + 2]; }
    if (in_v) { float4 vi = v[i]; vi.x = (float)in_v[3 * i]; vi.y = (float)in_v[3 * i + 1]; vi.z = (float)in_v[3 * i + 2]; v[i] = vi; }
}

// pairs (i<j) of list `li` inside rc2_term, appended unordered (host sorts)
template <bool PER>
__global__ void extract_pairs_kernel(const PairListDev* lists, int li, StepParams P, float rc2_term,
                                     int* __restrict__ out, long long capacity, unsigned long long* __restrict__ counter) {
    const PairListDev& L = lists[li];
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= list_rows(L)) return;
    const uint4 pi = L.spos[k];
    const int i = L.bead_of_rank[k];
    const int n = L.nbr_count[k];
    const unsigned int* row = L.nbr + (size_t)k * L.stride;
    for (int t = 0; t < n; t++) {
        const int kj = (int)(row[t] & 0xFFFFFFu);
        const uint4 pj = L.spos[kj];
        const int j = L.bead_of_rank[kj];
        if (j <= i) continue;
        float dx, dy, dz;
        delta<PER>(pi, pj, P.fscale, dx, dy, dz);
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
