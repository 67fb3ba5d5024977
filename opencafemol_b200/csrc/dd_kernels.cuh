// dd_kernels.cuh — slab domain decomposition of one periodic system over several GPUs (config C5).
//
// The box is cut into `world` slabs along one axis; rank r owns the beads whose fixed-point coordinate q on that axis
// satisfies (q * world) >> 32 == r, an exact integer rule every rank evaluates identically.  Every rank keeps arrays
// of the full system size (global bead indices everywhere: no renumbering when a bead migrates) but only the beads
// present on it are live: role 1 = owned (integrated here), role 2 = halo (within `halo` of a slab face, read as
// pair/bonded partners), role 0 = not on this rank (stale, never read).
//
// At every list rebuild:
//   A. dd_select_kernel   : owned beads within `halo` of a face -- or already across it -- go into the message for that
//                           neighbour with their full state (index, x, v).  After the exchange both sides hold
//                           bit-identical state for those beads, so ownership (dd_assign_kernel) needs no negotiation:
//                           a bead that crossed a face simply evaluates to role 1 on the far side and 2 on the near side.
//   B. dd_steplist_kernel : the (newly) owned beads within `halo` of a face form the per-step halo list; the index
//                           lists are exchanged once, and until the next rebuild every MD step ships only positions
//                           (dd_pack_kernel -> exchange -> dd_unpack_kernel) in that fixed order.
// What the reference would need OpenMM to do across devices (it cannot: one Context = one device,
// src/Simulator.cpp:14-17).
#pragma once
#include "kernels.cuh"

namespace ocm {

struct DDGeom {
    int world, rank, axis;
    unsigned int lo, hi;      // owned interval [lo, hi) of the fixed-point coordinate (hi wraps to 0 for the last rank)
    unsigned int halo_fix;    // halo width in fixed-point units
};

__device__ __forceinline__ unsigned int dd_coord(const uint4& q, int axis) { return axis == 0 ? q.x : (axis == 1 ? q.y : q.z); }
__device__ __forceinline__ int dd_owner(unsigned int q, int world) { return (int)(((unsigned long long)q * (unsigned long long)world) >> 32); }
// inside, or within the halo width of, the slab on the low / high side (wrapped signed distances)
__device__ __forceinline__ bool dd_near_lo(const DDGeom& G, unsigned int q) { return (int)(q - G.lo) < (int)G.halo_fix; }
__device__ __forceinline__ bool dd_near_hi(const DDGeom& G, unsigned int q) { return (int)(G.hi - 1u - q) < (int)G.halo_fix; }
// outside the slab but within the halo width of one of its faces
__device__ __forceinline__ bool dd_in_halo(const DDGeom& G, unsigned int q) {
    const int below = (int)(G.lo - 1u - q), above = (int)(q - G.hi);
    return (below >= 0 && below < (int)G.halo_fix) || (above >= 0 && above < (int)G.halo_fix);
}

struct DDMsg {
    int* count;               // [1] entries appended (may exceed capacity: overflow)
    int* idx;                 // [cap] bead indices
    double* x;                // [3*cap]
    float4* v;                // [cap]
};

// role by position alone (after ocm_set_positions, when every rank holds the whole system)
__global__ void dd_role_from_position_kernel(int n, DDGeom G, const uint4* __restrict__ posq, unsigned char* __restrict__ role) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned int q = dd_coord(posq[i], G.axis);
    role[i] = dd_owner(q, G.world) == G.rank ? 1 : (dd_in_halo(G, q) ? 2 : 0);
}

// phase A: full state of the owned beads near (or across) a face, one message per neighbour
__global__ void dd_select_kernel(int n, DDGeom G, int cap, const unsigned char* __restrict__ role, const uint4* __restrict__ posq,
                                 const double* __restrict__ x, const float4* __restrict__ v, DDMsg lo, DDMsg hi,
                                 int* __restrict__ flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || role[i] != 1) return;
    const unsigned int q = dd_coord(posq[i], G.axis);
    for (int side = 0; side < 2; side++) {
        if (!(side ? dd_near_hi(G, q) : dd_near_lo(G, q))) continue;
        const DDMsg& M = side ? hi : lo;
        const int p = atomicAdd(M.count, 1);
        atomicMax(&flags[FLAG_DD_MAXSEND], p + 1);
        if (p >= cap) { flags[FLAG_DD_OVERFLOW] = 1; continue; }
        M.idx[p] = i;
        M.x[3 * p] = x[3 * i]; M.x[3 * p + 1] = x[3 * i + 1]; M.x[3 * p + 2] = x[3 * i + 2];
        M.v[p] = v[i];
    }
}

// received phase-A message: install the state, mark the bead as known on this rank
__global__ void dd_install_kernel(DDGeom G, StepParams P, int cap, DDMsg M, double* __restrict__ x, float4* __restrict__ v,
                                  uint4* __restrict__ posq, unsigned char* __restrict__ known) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= min(*M.count, cap)) return;
    const int i = M.idx[p];
    const double px = M.x[3 * p], py = M.x[3 * p + 1], pz = M.x[3 * p + 2];
    x[3 * i] = px; x[3 * i + 1] = py; x[3 * i + 2] = pz;
    v[i] = M.v[p];
    posq[i] = make_uint4(to_fix(px, P.inv_box[0]), to_fix(py, P.inv_box[1]), to_fix(pz, P.inv_box[2]), 0u);
    known[i] = 1;
}

// new roles of every bead this rank has current state for (owned so far, or just received)
// ... and the compact list of the beads present here (owned + halo): integrate_kernel runs over it instead of over the system
__global__ void dd_assign_kernel(int n, DDGeom G, const uint4* __restrict__ posq, unsigned char* __restrict__ role,
                                 unsigned char* __restrict__ known, int* __restrict__ present, int* __restrict__ n_present) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned char r = 0;
    if (role[i] == 1 || known[i]) {
        const unsigned int q = dd_coord(posq[i], G.axis);
        r = dd_owner(q, G.world) == G.rank ? 1 : (dd_in_halo(G, q) ? 2 : 0);
    }
    role[i] = r;
    known[i] = 0;
    if (r) present[atomicAdd(n_present, 1)] = i;
}

// phase B: per-step halo lists of the owned beads (indices only)
__global__ void dd_steplist_kernel(int n, DDGeom G, int cap, const unsigned char* __restrict__ role, const uint4* __restrict__ posq,
                                   int* __restrict__ cnt_lo, int* __restrict__ idx_lo, int* __restrict__ cnt_hi, int* __restrict__ idx_hi,
                                   int* __restrict__ flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || role[i] != 1) return;
    const unsigned int q = dd_coord(posq[i], G.axis);
    if (dd_near_lo(G, q)) {
        const int p = atomicAdd(cnt_lo, 1);
        if (p < cap) idx_lo[p] = i; else flags[FLAG_DD_OVERFLOW] = 1;
        atomicMax(&flags[FLAG_DD_MAXSEND], p + 1);
    }
    if (dd_near_hi(G, q)) {
        const int p = atomicAdd(cnt_hi, 1);
        if (p < cap) idx_hi[p] = i; else flags[FLAG_DD_OVERFLOW] = 1;
        atomicMax(&flags[FLAG_DD_MAXSEND], p + 1);
    }
}

// every bonded term with an owned bead must have all its beads on this rank
// ... and the terms with an owned bead are compacted into terms[] (the bonded blocks of force_kernel run over that list)
__global__ void dd_check_bonded_kernel(const BondedDev* __restrict__ bonded, int kind, const unsigned char* __restrict__ role,
                                       int* __restrict__ flags, int* __restrict__ terms, int* __restrict__ count, int capacity) {
    const BondedDev& B = bonded[kind];
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= B.n) return;
    const int nb = bonded_nbody_dev(kind);
    bool mine = false, all = true;
    for (int q = 0; q < nb; q++) { const unsigned char r = role[B.idx[nb * m + q]]; mine = mine || r == 1; all = all && r != 0; }
    if (mine && !all) flags[FLAG_DD_REACH] = 1;
    if (mine) {
        const int p = atomicAdd(count, 1);
        if (p < capacity) terms[p] = m; else flags[FLAG_OVERFLOW] = 1;   // more terms than the force launch covers: re-size
    }
}

// per MD step: positions of the halo-list beads, both directions in one launch
__global__ void dd_pack_kernel(const int* __restrict__ cnt_lo, const int* __restrict__ idx_lo, double* __restrict__ out_lo,
                               const int* __restrict__ cnt_hi, const int* __restrict__ idx_hi, double* __restrict__ out_hi,
                               int cap, const double* __restrict__ x) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int nlo = min(*cnt_lo, cap), nhi = min(*cnt_hi, cap);
    if (p < nlo) { const int i = idx_lo[p]; out_lo[3 * p] = x[3 * i]; out_lo[3 * p + 1] = x[3 * i + 1]; out_lo[3 * p + 2] = x[3 * i + 2]; }
    if (p < nhi) { const int i = idx_hi[p]; out_hi[3 * p] = x[3 * i]; out_hi[3 * p + 1] = x[3 * i + 1]; out_hi[3 * p + 2] = x[3 * i + 2]; }
}

// per MD step over peer memory: the positions go straight into the neighbours' receive buffers (NVLink stores); the
// last block to finish publishes the step's sequence number in both neighbours' flags (release at system scope), on
// which their streams wait (cuStreamWaitValue32) before they unpack.  No staging copy, no collective call.
__global__ void dd_push_kernel(const int* __restrict__ cnt_lo, const int* __restrict__ idx_lo, double* __restrict__ out_lo,
                               const int* __restrict__ cnt_hi, const int* __restrict__ idx_hi, double* __restrict__ out_hi,
                               int cap, const double* __restrict__ x, unsigned int* flag_lo, unsigned int* flag_hi,
                               unsigned int seq, unsigned int* __restrict__ ticket) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int nlo = min(*cnt_lo, cap), nhi = min(*cnt_hi, cap);
    if (p < nlo) { const int i = idx_lo[p]; out_lo[3 * p] = x[3 * i]; out_lo[3 * p + 1] = x[3 * i + 1]; out_lo[3 * p + 2] = x[3 * i + 2]; }
    if (p < nhi) { const int i = idx_hi[p]; out_hi[3 * p] = x[3 * i]; out_hi[3 * p + 1] = x[3 * i + 1]; out_hi[3 * p + 2] = x[3 * i + 2]; }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(ticket, 1u);
        if (t == gridDim.x - 1) {
            *ticket = 0u;
            __threadfence_system();
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag_lo), "r"(seq) : "memory");
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag_hi), "r"(seq) : "memory");
        }
    }
}

__global__ void dd_unpack_kernel(const int* __restrict__ cnt_lo, const int* __restrict__ idx_lo, const double* __restrict__ in_lo,
                                 const int* __restrict__ cnt_hi, const int* __restrict__ idx_hi, const double* __restrict__ in_hi,
                                 int cap, StepParams P, double* __restrict__ x, uint4* __restrict__ posq,
                                 const PairListDev* __restrict__ lists, int nlists) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int nlo = min(*cnt_lo, cap), nhi = min(*cnt_hi, cap);
    for (int side = 0; side < 2; side++) {
        if (p >= (side ? nhi : nlo)) continue;
        const int i = side ? idx_hi[p] : idx_lo[p];
        const double* in = side ? in_hi : in_lo;
        const double px = in[3 * p], py = in[3 * p + 1], pz = in[3 * p + 2];
        x[3 * i] = px; x[3 * i + 1] = py; x[3 * i + 2] = pz;
        uint4 q = make_uint4(to_fix(px, P.inv_box[0]), to_fix(py, P.inv_box[1]), to_fix(pz, P.inv_box[2]), 0u);
        posq[i] = q;
        for (int l = 0; l < nlists; l++) {
            const int k = lists[l].rank_of_bead[i];
            if (k >= 0) { uint4* sp = lists[l].spos + k; q.w = sp->w; *sp = q; }
        }
    }
}

// state gather (ocm_get_state): zero what this rank does not own, the sum over ranks is then the whole system
__global__ void dd_mask_state_kernel(int n, const unsigned char* __restrict__ role, const double* __restrict__ x, const float4* __restrict__ v,
                                     const unsigned long long* __restrict__ fbuf, double* __restrict__ gx, float* __restrict__ gv,
                                     long long* __restrict__ gf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool mine = role[i] == 1;
    if (gx) for (int k = 0; k < 3; k++) gx[3 * i + k] = mine ? x[3 * i + k] : 0.0;
    if (gv) { const float4 vi = v[i]; gv[3 * i] = mine ? vi.x : 0.f; gv[3 * i + 1] = mine ? vi.y : 0.f; gv[3 * i + 2] = mine ? vi.z : 0.f; }
    if (gf) for (int k = 0; k < 3; k++) gf[(size_t)k * n + i] = mine ? (long long)fbuf[(size_t)k * n + i] : 0ll;
}

}  // namespace ocm
