// tile_kernels.cuh — cluster-pair ("tile") pair lists and the warp-tiled pair kernel (sm_100a).
//
// What it replaces: the per-bead Verlet rows of kernels.cuh (one 4-byte entry and one 16-byte gather per bead pair,
// every pair evaluated from both sides).  Here the beads of a pair list are sorted along a Hilbert curve over a fine
// cell grid; 8 consecutive beads form an i-cluster, 8 i-clusters a supercluster, 4 consecutive beads a j-group.
// The list of a supercluster S holds one 8-byte entry per j-group that has at least one bead pair inside the list
// radius, with an 8-bit mask of the i-clusters concerned.  A warp evaluates one (i-cluster, j-group) tile per pass:
// lane = tj * 8 + ti works on the pair (i = ci * 8 + ti, j = tj), so
//   * the j-bead is loaded once per entry (coalesced 16-byte loads, 4 addresses per warp) and reused by 8 i-clusters,
//   * the 64 i-beads sit in shared memory (LDS.128 with 8 distinct addresses per warp),
//   * Newton's third law: the i-forces stay in registers over the whole list of S (24 accumulators), the j-force
//     is summed over the 8 i-clusters in registers and reduced over the 8 ti lanes with 9 shuffles per entry,
//   * per-pair parameters are mixed from per-bead values in registers (no type decode, no coefficient-table look-up).
// Exclusions, the i < j rule inside a supercluster, interaction classes and padding are 32-bit lane masks, stored only
// for the entries that need them.  Forces are accumulated in the same 64-bit fixed-point buffer as everything else.
//
// OpenMM side this stands for: the CustomNonbondedForce kernels the reference's generators trigger
// (src/forcefield/*ForceFieldGenerator.cpp::generate), as kernels.cuh.
#pragma once
#include "kernels.cuh"

namespace ocm {

constexpr int kTileBlock = 128;       // threads per block of the tile force kernel (4 warps share one supercluster)
constexpr int kBuildBlock = 256;      // threads per block of the tile list build (8 warps per supercluster)
constexpr int kBuildBatch = 128;      // candidate superclusters per batch of the build

// supercluster-pair ownership: which of two superclusters lists the other (each unordered pair exactly once, balanced)
__device__ __forceinline__ bool tile_lists(int s, int o) { return (o > s) == (((s + o) & 1) == 0); }

__device__ __forceinline__ int tile_rows_pad(const PairListDev& L) { return (list_rows(L) + 63) & ~63; }

// ------------------------------------------------------------------------------------------------
// sorted per-bead arrays of a tile list (after bin / scan / scatter, which key the cells by their Hilbert rank):
// rank k -> position + mixing parameter a (spos), bead | type << 24 (sbead), (b_in, b_out) (sparam); ranks past the
// beads present are padding up to a multiple of 64 (sbead = -1)
__global__ void tile_fill_kernel(const PairListDev* lists, int li, const uint4* __restrict__ posq) {
    const PairListDev& L = lists[li];
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int rows = list_rows(L);
    if (p >= rows) {
        if (p < ((rows + 63) & ~63)) {
            L.spos[p] = make_uint4(0u, 0u, 0u, 0u); L.sbead[p] = -1; L.sparam[p] = make_float2(0.f, 0.f);
            L.saux[p] = make_uint4(0u, 0u, 0xFFFFFFFFu, 0u);
        }
        return;
    }
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
    q4.w = __float_as_uint(L.slot_a[a]);
    L.spos[k] = q4;
    const int sb = bead | ((L.form == SEG_PAIR_DH) ? 0 : ((int)L.type[a] << 24));
    const float2 b = L.slot_b[a];
    L.sbead[k] = sb;
    L.sparam[k] = b;
    L.saux[k] = make_uint4(__float_as_uint(b.x), __float_as_uint(b.y), (unsigned int)sb, 0u);
}

// bounding boxes of the superclusters (64 consecutive ranks): centre (fixed point) and half extents (nm).
// One warp per supercluster.  The extents are taken relative to the first bead with wrapped differences, so the box
// is a valid periodic interval whatever the spread of the beads.
__global__ void tile_bbox_kernel(const PairListDev* lists, int li, StepParams P) {
    const PairListDev& L = lists[li];
    const int S = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int rows = list_rows(L);
    if (S * 64 >= rows) return;
    const uint4 p0 = L.spos[S * 64];
    int lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    for (int h = 0; h < 2; h++) {
        const int k = S * 64 + h * 32 + lane;
        if (k < rows) {
            const uint4 p = L.spos[k];
            const int d[3] = {(int)(p.x - p0.x), (int)(p.y - p0.y), (int)(p.z - p0.z)};
            for (int q = 0; q < 3; q++) { lo[q] = min(lo[q], d[q]); hi[q] = max(hi[q], d[q]); }
        }
    }
    for (int q = 0; q < 3; q++)
        for (int o = 16; o > 0; o >>= 1) {
            lo[q] = min(lo[q], __shfl_xor_sync(0xffffffffu, lo[q], o));
            hi[q] = max(hi[q], __shfl_xor_sync(0xffffffffu, hi[q], o));
        }
    if (lane == 0) {
        const unsigned int base[3] = {p0.x, p0.y, p0.z};
        unsigned int c[3];
        float hf[3];
        for (int q = 0; q < 3; q++) {
            const long long mid = ((long long)lo[q] + (long long)hi[q]) >> 1;
            c[q] = base[q] + (unsigned int)(int)mid;
            // half extent, rounded up generously (one fixed-point unit + float rounding)
            hf[q] = ((float)(((long long)hi[q] - (long long)lo[q] + 3) >> 1)) * P.fscale[q] * 1.000001f;
        }
        L.sc_center[S] = make_uint4(c[0], c[1], c[2], 0u);
        L.sc_half[S] = make_float4(hf[0], hf[1], hf[2], 0.f);
    }
}

// ------------------------------------------------------------------------------------------------
// list build.  One block (8 warps) per supercluster S:
//   1. candidate superclusters O (those S lists by the ownership rule, and S itself) whose bounding box comes within
//      the list radius of S's box, found in batches of kBuildBatch by an ordered block-wide compaction;
//   2. every j-group of a candidate is tested against the 64 beads of S with the canonical distance: warp-wide, lane
//      (ti, tj), one i-cluster per pass.  An entry is emitted when some allowed pair lies inside the list radius; its
//      lane masks are stored only when they are not all ones.
// Entries come out ordered by j-group (deterministic).  count_only: only count entries (to size the rows).
template <bool PER>
__global__ void __launch_bounds__(kBuildBlock) tile_build_kernel(const PairListDev* lists, int li, StepParams P, int* __restrict__ flags,
                                                                 int count_only) {
    const PairListDev& L = lists[li];
    const int S = blockIdx.x;
    const int rows = list_rows(L);
    if (S == 0 && threadIdx.x == 0 && !count_only && rows > L.row_capacity) flags[FLAG_OVERFLOW] = 1;
    if (S * 64 >= rows) {
        if (threadIdx.x == 0 && S < L.tl_nsc) { L.tl_count[S] = 0; L.tl_weight[S] = 0; L.tl_count_in[S] = 0; L.tl_weight_in[S] = 0; }
        return;
    }
    const int nsc = (rows + 63) >> 6;
    __shared__ uint4 s_pos[64];
    __shared__ int s_info[64];            // bead | cls << 24, -1 = padding; bit 31 of a real bead never set
    __shared__ unsigned char s_own[64];
    __shared__ int s_cand[kBuildBatch];
    __shared__ int s_warp_cnt[kBuildBlock / 32];
    __shared__ uint2 s_stage[kBuildBatch * 16];
    __shared__ int s_total;   // lane pairs' tiles of the row (sum of the i-cluster mask bits)
    __shared__ int s_total_in, s_count_in;   // the same and the entry count for the interior entries (owned beads only)
    __shared__ uint4 s_cc[8];   // bounding boxes of the 8 i-clusters: centre (fixed point) ...
    __shared__ float4 s_ch[8];  // ... and half extents (nm); .w < 0: no real bead
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 64) {
        const int k = S * 64 + threadIdx.x;
        s_pos[threadIdx.x] = L.spos[k];
        const int sb = L.sbead[k];
        int info = -1;
        unsigned char own = 0;
        if (sb >= 0) {
            const int bead = sb & 0xFFFFFF;
            const int cl = L.cls ? (int)L.cls[L.sorted_slots[k]] : 0;
            info = bead | (cl << 24);
            own = L.role ? (L.role[bead] == 1) : 1;
        }
        s_info[threadIdx.x] = info;
        s_own[threadIdx.x] = own;
    }
    if (threadIdx.x == 0) { s_total = 0; s_total_in = 0; s_count_in = 0; }
    // slab decomposition with halo beads sorted last: a row of owned beads only, and its entries on j-groups of owned beads only,
    // can be evaluated before the step's halo positions have arrived (entries are ordered by j-group: they form a prefix)
    const int rows_owned = list_rows_owned(L);
    const bool interior_S = (S + 1) * 64 <= rows_owned;
    const int j4_owned = rows_owned >> 2;
    __syncthreads();
    if (threadIdx.x < 8) {
        // box of i-cluster threadIdx.x, relative to its first real bead (wrapped differences: a valid periodic interval)
        int lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0}, first = -1;
        uint4 p0 = make_uint4(0u, 0u, 0u, 0u);
        for (int t = 0; t < 8; t++) {
            const int il = threadIdx.x * 8 + t;
            if (s_info[il] < 0) continue;
            const uint4 p = s_pos[il];
            if (first < 0) { first = il; p0 = p; continue; }
            const int d[3] = {(int)(p.x - p0.x), (int)(p.y - p0.y), (int)(p.z - p0.z)};
            for (int q = 0; q < 3; q++) { lo[q] = min(lo[q], d[q]); hi[q] = max(hi[q], d[q]); }
        }
        const unsigned int base[3] = {p0.x, p0.y, p0.z};
        unsigned int c[3];
        float hf[3];
        for (int q = 0; q < 3; q++) {
            c[q] = base[q] + (unsigned int)(int)(((long long)lo[q] + (long long)hi[q]) >> 1);
            hf[q] = ((float)(((long long)hi[q] - (long long)lo[q] + 3) >> 1)) * P.fscale[q] * 1.000001f;
        }
        s_cc[threadIdx.x] = make_uint4(c[0], c[1], c[2], 0u);
        s_ch[threadIdx.x] = make_float4(hf[0], hf[1], hf[2], first < 0 ? -1.f : 0.f);
    }
    __syncthreads();
    const uint4 cS = L.sc_center[S];
    const float4 hS = L.sc_half[S];
    const float rl2 = L.rlist2;
    const int stride = L.tl_stride;
    uint2* out = L.tl_entries + (size_t)S * stride;
    const int ti = lane & 7, tj = lane >> 3;
    int n_out = 0;  // entries written so far (uniform across the block)

    for (int base = 0; base < nsc; base += kBuildBlock) {
        // ---- 1. candidates among superclusters [base, base + kBuildBlock) ----
        const int O = base + threadIdx.x;
        bool cand = false;
        if (O < nsc && (O == S || tile_lists(S, O))) {
            const uint4 cO = L.sc_center[O];
            const float4 hO = L.sc_half[O];
            float dx, dy, dz;
            delta<PER>(cS, cO, P.fscale, dx, dy, dz);
            dx = fmaxf(0.f, fabsf(dx) - (hS.x + hO.x));
            dy = fmaxf(0.f, fabsf(dy) - (hS.y + hO.y));
            dz = fmaxf(0.f, fabsf(dz) - (hS.z + hO.z));
            cand = (dx * dx + dy * dy + dz * dz) < rl2 * 1.0001f;
        }
        const unsigned int bal = __ballot_sync(0xffffffffu, cand);
        if (lane == 0) s_warp_cnt[warp] = __popc(bal);
        __syncthreads();
        int before = 0, ncand = 0;
        for (int w = 0; w < kBuildBlock / 32; w++) { if (w < warp) before += s_warp_cnt[w]; ncand += s_warp_cnt[w]; }
        // (kBuildBlock may exceed kBuildBatch: process the candidates of this sweep in batches)
        for (int b0 = 0; b0 < ncand; b0 += kBuildBatch) {
            __syncthreads();
            if (cand) {
                const int pos = before + __popc(bal & ((1u << lane) - 1u)) - b0;
                if (pos >= 0 && pos < kBuildBatch) s_cand[pos] = O;
            }
            __syncthreads();
            const int nb = min(kBuildBatch, ncand - b0);
            // ---- 2. the j-groups of the batch: candidate q = (candidate index) * 16 + (j-group within it) ----
            for (int q = warp; q < nb * 16; q += kBuildBlock / 32) {
                const int Oq = s_cand[q >> 4];
                const int j4 = Oq * 16 + (q & 15);
                const int kj = j4 * 4 + tj;
                uint2 entry = make_uint2(0xFFFFFFFFu, 0u);
                if (j4 * 4 < rows) {
                    const uint4 pj = L.spos[kj];
                    const int sbj = L.sbead[kj];
                    const bool real_j = sbj >= 0;
                    const int bead_j = sbj & 0xFFFFFF;
                    const int cls_j = (real_j && L.cls) ? (int)L.cls[L.sorted_slots[kj]] : 0;
                    const bool own_j = real_j && (L.role ? (L.role[bead_j] == 1) : true);
                    // quick rejection: lane (ti, tj) measures the j-bead tj against the box of i-CLUSTER ti; a cluster none of the
                    // four j-beads comes near is skipped by the exact pass below
                    float qx, qy, qz;
                    const float4 hC = s_ch[ti];
                    delta<PER>(s_cc[ti], pj, P.fscale, qx, qy, qz);
                    qx = fmaxf(0.f, fabsf(qx) - hC.x); qy = fmaxf(0.f, fabsf(qy) - hC.y); qz = fmaxf(0.f, fabsf(qz) - hC.z);
                    const bool near = real_j && hC.w >= 0.f && (qx * qx + qy * qy + qz * qz) < rl2 * 1.0001f;
                    const unsigned int nm = __ballot_sync(0xffffffffu, near);
                    const unsigned int cnear = (nm | (nm >> 8) | (nm >> 16) | (nm >> 24)) & 0xffu;
                    if (cnear) {
                        unsigned int act = 0;
                        bool partial = false;
                        unsigned int masks[8];
#pragma unroll
                        for (int ci = 0; ci < 8; ci++) {
                            masks[ci] = 0xffffffffu;
                            if (!((cnear >> ci) & 1u)) continue;  // warp-uniform
                            const int il = ci * 8 + ti;
                            const uint4 pi = s_pos[il];
                            const int info = s_info[il];
                            bool allowed = real_j && info >= 0;
                            if (Oq == S) allowed = allowed && (S * 64 + il) < kj;
                            if (L.cls) allowed = allowed && ((L.allow[(info >> 24) & 31] >> cls_j) & 1u);
                            if (L.role) allowed = allowed && (own_j || s_own[il]);
                            float dx, dy, dz;
                            delta<PER>(pi, pj, P.fscale, dx, dy, dz);
                            const bool in = allowed && norm2(dx, dy, dz) < rl2;
                            const unsigned int m = __ballot_sync(0xffffffffu, allowed);
                            masks[ci] = m;
                            if (__any_sync(0xffffffffu, in)) { act |= 1u << ci; partial = partial || (m != 0xffffffffu); }
                        }
                        if (act) {
                            unsigned int midx = 0;
                            if (partial && !count_only) {
                                if (lane == 0) midx = (unsigned int)atomicAdd(L.tl_mask_count, 1) + 1u;
                                midx = __shfl_sync(0xffffffffu, midx, 0);
                                if (midx > (unsigned int)L.tl_mask_cap) { if (lane == 0) flags[FLAG_OVERFLOW] = 1; midx = 0; }
                                else if (lane < 8) {
                                    unsigned int mv = 0;
#pragma unroll
                                    for (int ci = 0; ci < 8; ci++) mv = (lane == ci) ? masks[ci] : mv;
                                    L.tl_masks[(size_t)(midx - 1u) * 8 + lane] = mv;
                                }
                            } else if (partial && lane == 0) atomicAdd(L.tl_mask_count, 1);
                            entry = make_uint2((unsigned int)j4, act | (midx << 8));
                        }
                    }
                }
                if (lane == 0) s_stage[q] = entry;
            }
            __syncthreads();
            // ---- ordered compaction of the batch's entries into the row of S ----
            for (int q0 = 0; q0 < nb * 16; q0 += kBuildBlock) {
                const int q = q0 + threadIdx.x;
                const uint2 e = (q < nb * 16) ? s_stage[q] : make_uint2(0xFFFFFFFFu, 0u);
                const bool hit = e.x != 0xFFFFFFFFu;
                const unsigned int hb = __ballot_sync(0xffffffffu, hit);
                __syncthreads();
                if (lane == 0) s_warp_cnt[warp] = __popc(hb);
                __syncthreads();
                int off = 0, tot = 0;
                for (int w = 0; w < kBuildBlock / 32; w++) { if (w < warp) off += s_warp_cnt[w]; tot += s_warp_cnt[w]; }
                if (hit && !count_only) {
                    const int pos = n_out + off + __popc(hb & ((1u << lane) - 1u));
                    if (pos < stride) out[pos] = e;
                    atomicAdd(&s_total, __popc(e.y & 0xffu));
                    if (interior_S && (int)e.x < j4_owned) { atomicAdd(&s_total_in, __popc(e.y & 0xffu)); atomicAdd(&s_count_in, 1); }
                }
                n_out += tot;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (count_only) atomicMax(&flags[FLAG_MAXDISP + 1], n_out);
        else {
            if (n_out > stride) { flags[FLAG_OVERFLOW] = 1; n_out = stride; }
            L.tl_count[S] = n_out;
            // work estimate of the row in units of one tile evaluation: tiles + entry overheads + the per-row prologue / epilogue;
            // tl_weight is what the launch over entries [count_in, count) does, tl_weight_in the launch over [0, count_in)
            const int n_in = L.halo_split ? min(s_count_in, n_out) : 0, t_in = L.halo_split ? s_total_in : 0;
            L.tl_count_in[S] = n_in;
            L.tl_weight_in[S] = n_in ? t_in + n_in + 8 : 0;
            L.tl_weight[S] = (n_out - n_in) ? (s_total - t_in) + (n_out - n_in) + 8 : 0;
        }
    }
}

// exclusions: clear the lane bit of every excluded bead pair that the list holds (one thread per exclusion).
// An entry that had no masks yet gets a block of all ones first (compare-and-swap on its mask index).
__global__ void tile_excl_kernel(const PairListDev* lists, int li, int* __restrict__ flags, int count_only) {
    const PairListDev& L = lists[li];
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= L.n_excl_pairs) return;
    const int2 ab = L.excl_pairs[x];
    const int ka = L.rank_of_bead[ab.x], kb = L.rank_of_bead[ab.y];
    if (ka < 0 || kb < 0) return;
    const int Sa = ka >> 6, Sb = kb >> 6;
    int ki, kj;
    if (Sa == Sb) { ki = min(ka, kb); kj = max(ka, kb); }
    else if (tile_lists(Sa, Sb)) { ki = ka; kj = kb; }
    else { ki = kb; kj = ka; }
    const int S = ki >> 6, ci = (ki >> 3) & 7, lane = ((kj & 3) << 3) | (ki & 7);
    const unsigned int j4 = (unsigned int)(kj >> 2);
    uint2* row = L.tl_entries + (size_t)S * L.tl_stride;
    int lo = 0, hi = count_only ? 0 : L.tl_count[S];
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (row[mid].x < j4) lo = mid + 1; else hi = mid; }
    if (count_only) { atomicAdd(L.tl_mask_count, 1); return; }  // worst case: every exclusion opens a mask block
    if (lo >= L.tl_count[S] || row[lo].x != j4) return;
    unsigned int* py = &row[lo].y;
    unsigned int y = *(volatile unsigned int*)py;
    if ((y >> 8) == 0u) {
        const unsigned int mine = (unsigned int)atomicAdd(L.tl_mask_count, 1) + 1u;
        if (mine > (unsigned int)L.tl_mask_cap) { flags[FLAG_OVERFLOW] = 1; return; }
        for (int w = 0; w < 8; w++) L.tl_masks[(size_t)(mine - 1u) * 8 + w] = 0xffffffffu;
        __threadfence();
        const unsigned int old = atomicCAS(py, y, y | (mine << 8));
        y = (old == y) ? (y | (mine << 8)) : old;  // lost the race: use the winner's block (ours stays unused)
    }
    atomicAnd(&L.tl_masks[(size_t)((y >> 8) - 1u) * 8 + ci], ~(1u << lane));
}

// exclusive prefix sum of the row weights (tl_weight: estimated work of a row): the force launch cuts the concatenated
// entry stream into shares of equal weight, one per block.  tl_start[nsc] = total.  Single block.
__global__ void tile_scan_kernel(const PairListDev* lists, int li, int nsc, int interior) {
    const PairListDev& L = lists[li];
    const int* __restrict__ weight = interior ? L.tl_weight_in : L.tl_weight;
    int* __restrict__ start = interior ? L.tl_start_in : L.tl_start;
    __shared__ int warp_sums[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int base = 0; base < nsc; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const int v = (i < nsc) ? weight[i] : 0;
        int s = v;
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += t; }
        if (lane == 31) warp_sums[wid] = s;
        __syncthreads();
        if (wid == 0) {
            int ws = (lane < nw) ? warp_sums[lane] : 0;
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, ws, o); if (lane >= o) ws += t; }
            warp_sums[lane] = ws;
        }
        __syncthreads();
        if (i < nsc) start[i] = carry + (wid > 0 ? warp_sums[wid - 1] : 0) + s - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += warp_sums[nw - 1];
        __syncthreads();
    }
    if (threadIdx.x == 0) start[nsc] = carry;
}

// ------------------------------------------------------------------------------------------------
// pair evaluation
struct TileForm {
    float rc2;
    float alpha, beta;     // -dE/dr / r = e q3 (alpha q3 - beta) / r^2, q = s^2 / r^2, s = a_i + a_j
    float qz;              // inner zone: q > qz
    float rsw, inv_sw, sw_c0, sw_c1;  // switch: t = sat(r inv_sw - sw_c0), -kappa S'(r) = sw_c1 (t - t^2)^2
    float kappa;           // switched energy: E = e kappa q3 (q3 - 1)
    float inv_debye;       // screened Coulomb
};

// two-zone 12-6 family with per-bead mixing.  a = s_i + s_j partial sums, (b, c) = inner / outer strength.
template <bool SW, bool GEO>
__device__ __forceinline__ float tile_pair_tz(const TileForm& F, float r2, float ai, float aj, float2 bi, float2 bj) {
    const float s = ai + aj;
    const float s2 = s * s;
    const float ir = fast_rsqrt(r2);
    const float ir2 = ir * ir;
    const float q = s2 * ir2;
    const float q3 = q * q * q;
    const float e_in = GEO ? bi.x * bj.x : bi.x + bj.x;
    const float e_out = bi.y + bj.y;
    const float e = (q > F.qz) ? e_in : e_out;
    const float g = fmaf(F.alpha, q3, -F.beta);
    if (!SW) return (e * q3) * (g * ir2);
    // OpenMM switching function on [rsw, rc]: S = 1 - 10 t^3 + 15 t^4 - 6 t^5, S' = -30 t^2 (1-t)^2 / (rc - rsw);
    //   -d(S E)/dr / r = S (e q3 g / r^2) - E S'/r,  E = e kappa q3 (q3 - 1)
    //                  = (e q3 / r) [ (g / r) S - kappa S'(r) (q3 - 1) ]
    const float r = r2 * ir;
    const float t = __saturatef(fmaf(r, F.inv_sw, -F.sw_c0));
    const float t2 = t * t;
    const float Sw = fmaf(t2 * t, fmaf(t, fmaf(t, -6.f, 15.f), -10.f), 1.f);
    const float w = t - t2;
    const float u3 = (q3 - 1.f) * (w * w);
    const float inner = fmaf(u3, F.sw_c1, (g * ir) * Sw);   // sw_c1 = 30 kappa / (rc - rsw)
    return ((e * q3) * ir) * inner;
}

// screened Coulomb: E = q_i q_j K (exp(-r / debye) / r - cc); a = q sqrt(K)
__device__ __forceinline__ float tile_pair_dh(const TileForm& F, float r2, float ai, float aj) {
    const float ir = fast_rsqrt(r2);
    const float r = r2 * ir;
    const float ex = __expf(-r * F.inv_debye);
    return ai * aj * ex * (ir + F.inv_debye) * ir * ir;
}

// Launch description of the tile force kernel (one launch per tile list), passed BY VALUE: every field is a constant-bank
// operand, no dependent global loads before the first useful instruction and no registers spent on per-list constants.
struct TileSegP {
    int nblocks, nsc, stride, pad;
    const uint2* entries; const int* count; const int* count_lo; const int* start; const uint4* spos; const uint4* saux; const unsigned int* masks; const uint4* sc_center;
    float rc2, alpha, beta, qz, inv_sw, sw_c0, sw_c1, inv_debye;
};
struct TileLaunch {
    int n, pdl_early;
    int debug;   // timing experiments only (option tile_debug): 1 = no j-force atomics, 2 = no i-force atomics, 8 = skip every tile evaluation
    float fscale[3];
    float half_box[3];
    TileSegP seg;
};

// one (i, j) lane pair in two phases (so that two tiles can share a basic block and a vote).
// pi = (x, y, z fixed point, a) of the i-bead, pj likewise.  The difference of the fixed-point integers is the minimum
// image whatever the box (FP32 offsets from the supercluster centre would save 6 instructions per pair but are only
// the minimum image when supercluster extent + cutoff < half the box, which a 15 nm slab box does not guarantee).
struct TilePair { float dx, dy, dz, r2; bool in; };
__device__ __forceinline__ TilePair tile_geom(const TileLaunch& T, const uint4 pi, const uint4 pj, bool ok) {
    TilePair t;
    t.dx = __int2float_rn((int)(pi.x - pj.x)) * T.fscale[0];
    t.dy = __int2float_rn((int)(pi.y - pj.y)) * T.fscale[1];
    t.dz = __int2float_rn((int)(pi.z - pj.z)) * T.fscale[2];
    t.r2 = norm2(t.dx, t.dy, t.dz);
    t.in = ok && t.r2 < T.seg.rc2;
    return t;
}
// ---- two tiles at once with the packed FP32 instructions of sm_100 (FFMA2 / FMUL2 / FADD2: two independent FP32 results
// per issue slot).  The kernel is bound by instruction issue, not by the FMA pipe, so halving the issue slots of the
// floating-point part of a pair of evaluations is a direct gain.  Component .x is tile A, .y is tile B; the operations are
// those of the scalar path (up to the compiler's FMA contraction there).
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 f2(float a) { return make_float2(a, a); }
struct TilePair2 { float2 dx, dy, dz, r2; bool in_a, in_b; };
__device__ __forceinline__ TilePair2 tile_geom2(const TileLaunch& T, const uint4 pa, const uint4 pb, const uint4 pj, bool ok_a, bool ok_b) {
    TilePair2 t;
    t.dx = __fmul2_rn(f2(__int2float_rn((int)(pa.x - pj.x)), __int2float_rn((int)(pb.x - pj.x))), f2(T.fscale[0]));
    t.dy = __fmul2_rn(f2(__int2float_rn((int)(pa.y - pj.y)), __int2float_rn((int)(pb.y - pj.y))), f2(T.fscale[1]));
    t.dz = __fmul2_rn(f2(__int2float_rn((int)(pa.z - pj.z)), __int2float_rn((int)(pb.z - pj.z))), f2(T.fscale[2]));
    t.r2 = __ffma2_rn(t.dz, t.dz, __ffma2_rn(t.dy, t.dy, __fmul2_rn(t.dx, t.dx)));   // norm2() on both
    t.in_a = ok_a && t.r2.x < T.seg.rc2;
    t.in_b = ok_b && t.r2.y < T.seg.rc2;
    return t;
}
template <bool SW, bool GEO>
__device__ __forceinline__ float2 tile_pair_tz2(const TileSegP& G, float2 r2, float2 ai, float aj, float2 bi_in, float2 bi_out, float2 bj) {
    const float2 s = __fadd2_rn(ai, f2(aj));
    const float2 s2 = __fmul2_rn(s, s);
    const float2 ir = f2(fast_rsqrt(r2.x), fast_rsqrt(r2.y));
    const float2 ir2 = __fmul2_rn(ir, ir);
    const float2 q = __fmul2_rn(s2, ir2);
    const float2 q3 = __fmul2_rn(__fmul2_rn(q, q), q);
    const float2 e_in = GEO ? __fmul2_rn(bi_in, f2(bj.x)) : __fadd2_rn(bi_in, f2(bj.x));
    const float2 e_out = __fadd2_rn(bi_out, f2(bj.y));
    const float2 e = f2((q.x > G.qz) ? e_in.x : e_out.x, (q.y > G.qz) ? e_in.y : e_out.y);
    const float2 g = __ffma2_rn(f2(G.alpha), q3, f2(-G.beta));
    if (!SW) return __fmul2_rn(__fmul2_rn(e, q3), __fmul2_rn(g, ir2));
    const float2 r = __fmul2_rn(r2, ir);
    float2 t = __ffma2_rn(r, f2(G.inv_sw), f2(-G.sw_c0));
    t = f2(__saturatef(t.x), __saturatef(t.y));
    const float2 t2 = __fmul2_rn(t, t);
    const float2 Sw = __ffma2_rn(__fmul2_rn(t2, t), __ffma2_rn(t, __ffma2_rn(t, f2(-6.f), f2(15.f)), f2(-10.f)), f2(1.f));
    const float2 w = __fadd2_rn(t, f2(-t2.x, -t2.y));
    const float2 u3 = __fmul2_rn(__fadd2_rn(q3, f2(-1.f)), __fmul2_rn(w, w));
    const float2 inner = __ffma2_rn(u3, f2(G.sw_c1), __fmul2_rn(__fmul2_rn(g, ir), Sw));
    return __fmul2_rn(__fmul2_rn(__fmul2_rn(e, q3), ir), inner);
}
__device__ __forceinline__ float2 tile_pair_dh2(const TileSegP& G, float2 r2, float2 ai, float aj) {
    const float2 ir = f2(fast_rsqrt(r2.x), fast_rsqrt(r2.y));
    const float2 r = __fmul2_rn(r2, ir);
    const float2 ex = f2(__expf(-r.x * G.inv_debye), __expf(-r.y * G.inv_debye));
    return __fmul2_rn(__fmul2_rn(__fmul2_rn(__fmul2_rn(ai, f2(aj)), ex), __fadd2_rn(ir, f2(G.inv_debye))), __fmul2_rn(ir, ir));
}
// evaluate tiles A and B; fi* are the (A, B) i-force accumulators, fj* the j-force accumulators of the two tiles (summed later)
template <int FORM, bool SW, bool GEO>
__device__ __forceinline__ void tile_eval2(const TileLaunch& T, const TilePair2& t, float2 ai, float2 bi_in, float2 bi_out, float aj, float2 bj,
                                           float2& fix, float2& fiy, float2& fiz, float2& fjx, float2& fjy, float2& fjz) {
    float2 fr;
    if (FORM == SEG_PAIR_TZ) fr = tile_pair_tz2<SW, GEO>(T.seg, t.r2, ai, aj, bi_in, bi_out, bj);
    else fr = tile_pair_dh2(T.seg, t.r2, ai, aj);
    fr = f2(t.in_a ? fr.x : 0.f, t.in_b ? fr.y : 0.f);
    fix = __ffma2_rn(fr, t.dx, fix); fiy = __ffma2_rn(fr, t.dy, fiy); fiz = __ffma2_rn(fr, t.dz, fiz);
    const float2 nfr = f2(-fr.x, -fr.y);
    fjx = __ffma2_rn(nfr, t.dx, fjx); fjy = __ffma2_rn(nfr, t.dy, fjy); fjz = __ffma2_rn(nfr, t.dz, fjz);
}

template <int FORM, bool SW, bool GEO>
__device__ __forceinline__ void tile_eval(const TileLaunch& T, const TilePair& t, float ai, const float2 bi, float aj, const float2 bj,
                                          float& fix, float& fiy, float& fiz, float& fjx, float& fjy, float& fjz) {
    TileForm F;
    const TileSegP& G = T.seg;
    F.rc2 = G.rc2; F.alpha = G.alpha; F.beta = G.beta; F.qz = G.qz; F.inv_sw = G.inv_sw; F.sw_c0 = G.sw_c0; F.sw_c1 = G.sw_c1; F.inv_debye = G.inv_debye;
    float fr;
    if (FORM == SEG_PAIR_TZ) fr = tile_pair_tz<SW, GEO>(F, t.r2, ai, aj, bi, bj);
    else fr = tile_pair_dh(F, t.r2, ai, aj);
    fr = t.in ? fr : 0.f;
    fix = fmaf(fr, t.dx, fix); fiy = fmaf(fr, t.dy, fiy); fiz = fmaf(fr, t.dz, fiz);
    fjx = fmaf(-fr, t.dx, fjx); fjy = fmaf(-fr, t.dy, fjy); fjz = fmaf(-fr, t.dz, fjz);
}

// Shared memory of one block of the force kernel.  Everything an entry needs is staged here for 32 entries at a time by
// all 128 threads at once (one memory latency per batch instead of one per entry), and the j-forces leave through it too.
constexpr int kTileBatch = 32;   // entries per batch (8 per warp)
struct TileShared {
    uint4 pos[64];                       // the 64 i-beads of the supercluster: fixed-point position + mixing parameter a
    float2 par[64];                      // (b_in, b_out)
    float4 pp4[32];                      // the same per PAIR of i-clusters (2p, 2p+1) and lane ti: (a_A, a_B, b_in_A, b_in_B) ...
    float2 pp2[32];                      // ... and (b_out_A, b_out_B): aligned operands of the packed two-tile evaluation
    int bead[64];
    float acc[(kTileBlock / 32) * 192];  // per-warp i-force partial sums
    uint2 ent[kTileBatch];
    uint4 jpos[kTileBatch * 4], jaux[kTileBatch * 4];
    unsigned int mask[kTileBatch * 8];
    float fj[kTileBatch * 4 * 3];
};

// One piece [e_begin, e_end) of the entry row of supercluster S, by the 4 warps of a block.
template <int FORM, bool SW, bool GEO>
__device__ __forceinline__ void tile_chunk(const TileLaunch& T, int S, int e_begin, int e_end, unsigned long long* __restrict__ fbuf, TileShared& sh) {
    const TileSegP& G = T.seg;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ti = lane & 7, tj = lane >> 3;
    __syncthreads();  // (a block works on several rows in turn: the previous one is done with shared memory)
    if (tid < 64) {
        const uint4 ax = __ldg(&G.saux[S * 64 + tid]);
        const uint4 p = __ldg(&G.spos[S * 64 + tid]);
        sh.pos[tid] = p;
        sh.par[tid] = make_float2(__uint_as_float(ax.x), __uint_as_float(ax.y));
        sh.bead[tid] = (int)ax.z;
        // pair layout: bead tid = cluster c, lane t; pair p = c / 2, half h = c & 1
        const int c = tid >> 3, t = tid & 7, pidx = (c >> 1) * 8 + t, h = c & 1;
        float* q4 = reinterpret_cast<float*>(&sh.pp4[pidx]);
        float* q2 = reinterpret_cast<float*>(&sh.pp2[pidx]);
        q4[h] = __uint_as_float(p.w); q4[2 + h] = __uint_as_float(ax.x); q2[h] = __uint_as_float(ax.y);
    }
    const uint2* __restrict__ row = G.entries + (size_t)S * G.stride;
    const uint4* my_pos = sh.pos + ti;
    const float2* my_par = sh.par + ti;
    // i-force accumulators: component .x = cluster 2p, .y = cluster 2p + 1
    float2 fi[4][3];
#pragma unroll
    for (int p2 = 0; p2 < 4; p2++) { fi[p2][0] = f2(0.f); fi[p2][1] = f2(0.f); fi[p2][2] = f2(0.f); }

    for (int eb = e_begin; eb < e_end; eb += kTileBatch) {
        const int n_e = min(kTileBatch, e_end - eb);
        // ---- stage the batch: entries, their j-beads, their lane masks ----
        if (tid < n_e) sh.ent[tid] = __ldg(&row[eb + tid]);
        __syncthreads();
        if ((tid >> 2) < n_e) {
            const int kj = (int)sh.ent[tid >> 2].x * 4 + (tid & 3);
            sh.jpos[tid] = __ldg(&G.spos[kj]);
            sh.jaux[tid] = __ldg(&G.saux[kj]);
        }
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int w = tid + h * kTileBlock;   // word (entry w / 8, i-cluster w % 8)
            if ((w >> 3) < n_e) {
                const unsigned int midx = sh.ent[w >> 3].y >> 8;
                if (midx) sh.mask[w] = __ldg(&G.masks[(size_t)(midx - 1u) * 8 + (w & 7)]);
            }
        }
        __syncthreads();
        // ---- evaluate: warp w takes entries w, w + 4, ... of the batch ----
        for (int e = warp; e < n_e; e += kTileBlock / 32) {
            const unsigned int ey = sh.ent[e].y;
            const unsigned int act = (T.debug & 8) ? 0u : (ey & 0xffu);
            const uint4 pj = sh.jpos[e * 4 + tj];
            const uint4 xj = sh.jaux[e * 4 + tj];
            // lane masks (only the entries with exclusions, the i < j rule or padding have any)
            unsigned int lane_ok = 0xffu;
            if (ey >> 8) {
                lane_ok = 0u;
#pragma unroll
                for (int ci = 0; ci < 8; ci++) lane_ok |= ((sh.mask[e * 8 + ci] >> lane) & 1u) << ci;
            }
            const float aj = __uint_as_float(pj.w);
            const float2 bj = make_float2(__uint_as_float(xj.x), __uint_as_float(xj.y));
            float2 fjx = f2(0.f), fjy = f2(0.f), fjz = f2(0.f);
            // i-clusters in pairs: when both are active the two tiles are evaluated together with packed FP32 instructions
#pragma unroll
            for (int p2 = 0; p2 < 4; p2++) {
                const int c2 = 2 * p2;
                const unsigned int a2 = (act >> c2) & 3u;
                if (a2 == 0u) continue;  // warp-uniform
                if (a2 == 3u) {
                    const uint4 pa = my_pos[c2 * 8], pb = my_pos[c2 * 8 + 8];
                    const TilePair2 t = tile_geom2(T, pa, pb, pj, (lane_ok >> c2) & 1u, (lane_ok >> (c2 + 1)) & 1u);
                    if (!__any_sync(0xffffffffu, t.in_a || t.in_b)) continue;  // both tiles sit in the skin at this step
                    const float4 q4 = sh.pp4[p2 * 8 + ti];
                    const float2 q2 = sh.pp2[p2 * 8 + ti];
                    tile_eval2<FORM, SW, GEO>(T, t, f2(q4.x, q4.y), f2(q4.z, q4.w), q2, aj, bj, fi[p2][0], fi[p2][1], fi[p2][2], fjx, fjy, fjz);
                } else if (a2 == 1u) {
                    const uint4 pa = my_pos[c2 * 8];
                    const TilePair ta = tile_geom(T, pa, pj, (lane_ok >> c2) & 1u);
                    if (!__any_sync(0xffffffffu, ta.in)) continue;
                    tile_eval<FORM, SW, GEO>(T, ta, __uint_as_float(pa.w), my_par[c2 * 8], aj, bj, fi[p2][0].x, fi[p2][1].x, fi[p2][2].x, fjx.x, fjy.x, fjz.x);
                } else {
                    const uint4 pb = my_pos[c2 * 8 + 8];
                    const TilePair tb = tile_geom(T, pb, pj, (lane_ok >> (c2 + 1)) & 1u);
                    if (!__any_sync(0xffffffffu, tb.in)) continue;
                    tile_eval<FORM, SW, GEO>(T, tb, __uint_as_float(pb.w), my_par[c2 * 8 + 8], aj, bj, fi[p2][0].y, fi[p2][1].y, fi[p2][2].y, fjx.x, fjy.x, fjz.x);
                }
            }
            // j-force: the two tile accumulators, then over the 8 ti lanes of each tj group with 6 shuffles; x, y, z end up in
            // lanes ti = 0, 1, 2
            float gx = fjx.x + fjx.y, gy = fjy.x + fjy.y, gz = fjz.x + fjz.y;
            gx += __shfl_down_sync(0xffffffffu, gx, 1);
            gy += __shfl_up_sync(0xffffffffu, gy, 1);
            gz += __shfl_down_sync(0xffffffffu, gz, 1);
            if (ti & 1) gx = gy;
            gx += __shfl_down_sync(0xffffffffu, gx, 2);
            gz += __shfl_up_sync(0xffffffffu, gz, 2);
            if (ti & 2) gx = gz;
            gx += __shfl_down_sync(0xffffffffu, gx, 4);
            if (ti < 3) sh.fj[(e * 4 + tj) * 3 + ti] = gx;
        }
        __syncthreads();
        // ---- the j-forces of the batch leave together: one fixed-point atomic per (j-bead, component) ----
#pragma unroll
        for (int h = 0; h < 3; h++) {
            const int v = tid + h * kTileBlock;
            const int j = v / 3, q = v - j * 3;
            if (j < n_e * 4) {
                const int sbj = (int)sh.jaux[j].z;
                const float f = sh.fj[v];
                if (sbj >= 0 && f != 0.f && !(T.debug & 1)) atomicAdd(&fbuf[(size_t)q * T.n + (sbj & 0xFFFFFF)], (unsigned long long)(long long)(f * 4294967296.0f));
            }
        }
        // (the next batch's first barrier comes after its entry loads: jaux / fj are not overwritten before every thread is past here)
    }
    // i-forces: over the 4 tj groups by a transposing shuffle reduction (24 values -> 12 -> 6 per lane), over the 4 warps
    // in shared memory in a fixed order (run-to-run reproducible), then one atomic triple per i-bead
    float r1[12];
#pragma unroll
    for (int m = 0; m < 12; m++) {
        // value index m = cluster (m / 3), component (m % 3); its partner is cluster + 4
        const float2 a = fi[(m / 3) >> 1][m % 3], bb = fi[((m / 3) + 4) >> 1][m % 3];
        const float lo = ((m / 3) & 1) ? a.y : a.x, hi = ((m / 3) & 1) ? bb.y : bb.x;
        const float send = (tj & 1) ? lo : hi, keep = (tj & 1) ? hi : lo;
        r1[m] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int m = 0; m < 6; m++) {
        const float send = (tj & 2) ? r1[m] : r1[m + 6], keep = (tj & 2) ? r1[m + 6] : r1[m];
        const float v = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        // this lane now holds component q of i-cluster ci, both functions of (m, tj)
        const int M = m + ((tj & 2) ? 6 : 0) + ((tj & 1) ? 12 : 0);
        const int ci = M / 3, q = M - ci * 3;
        sh.acc[warp * 192 + q * 64 + ci * 8 + ti] = v;
    }
    __syncthreads();
    for (int t = tid; t < 192; t += kTileBlock) {
        const float v = (sh.acc[t] + sh.acc[192 + t]) + (sh.acc[384 + t] + sh.acc[576 + t]);
        const int il = t & 63, q = t >> 6;
        const int sb = sh.bead[il];
        if (sb >= 0 && v != 0.f && !(T.debug & 2)) atomicAdd(&fbuf[(size_t)q * T.n + (sb & 0xFFFFFF)], (unsigned long long)(long long)(v * 4294967296.0f));
    }
}

#ifndef OCM_TILE_MINBLOCKS
#define OCM_TILE_MINBLOCKS 5
#endif
// Persistent launch, one wave: block b works on the share [W b / B, W (b+1) / B) of the list's work W (row weights, prefix
// sums in start[]); inside a row the weight is taken as spread evenly over its entries.
template <int FORM, bool SW, bool GEO>
__global__ void __launch_bounds__(kTileBlock, OCM_TILE_MINBLOCKS) tile_force_kernel(const __grid_constant__ TileLaunch T, unsigned long long* __restrict__ fbuf) {
    __shared__ TileShared sh;
    __shared__ int s_first;
    if (T.pdl_early) pdl_trigger();
    pdl_wait();
    const TileSegP& G = T.seg;
    const int* __restrict__ start = G.start;
    const long long W = start[G.nsc];
    const int t0 = (int)(W * blockIdx.x / G.nblocks), t1 = (int)(W * (blockIdx.x + 1) / G.nblocks);
    if (t0 >= t1) return;
    // first row: the last S with start[S] <= t0.  One warp looks at 32 evenly spaced rows per round (two rounds cover 1024 rows)
    if (threadIdx.x < 32) {
        int lo = 0, hi = G.nsc;  // start[lo] <= t0 < start[hi]  (start[nsc] = W > t0)
        while (hi - lo > 1) {
            const int step = (hi - lo + 31) / 32;
            const int probe = min(hi, lo + (int)threadIdx.x * step);
            const bool le = probe < hi && start[probe] <= t0;
            const unsigned int m = __ballot_sync(0xffffffffu, le);
            const int top = 31 - __clz((int)m);   // lane 0 probes lo itself, which qualifies
            const int nlo = lo + top * step;
            hi = min(hi, nlo + step); lo = nlo;
        }
        if (threadIdx.x == 0) s_first = lo;
    }
    __syncthreads();
    for (int S = s_first; S < G.nsc; S++) {
        const int w0 = start[S], w1 = start[S + 1];
        if (w0 >= t1) break;
        if (w1 == w0) continue;
        // the launch covers the entries [lo, hi) of the row: all of them, or (slab decomposition) the interior prefix in one
        // launch and the rest in another
        const int lo_e = G.count_lo ? G.count_lo[S] : 0, hi_e = G.count[S];
        const int cnt = hi_e - lo_e;
        // entry range of this block inside the row (multiples of 4: one entry per warp and pass); both neighbours of a cut use
        // the same formula, so the shares tile the row exactly
        const int e_begin = (t0 <= w0) ? 0 : (int)(((long long)(t0 - w0) * cnt / (w1 - w0)) & ~3ll);
        const int e_end = (t1 >= w1) ? cnt : (int)(((long long)(t1 - w0) * cnt / (w1 - w0)) & ~3ll);
        if (e_begin < e_end) tile_chunk<FORM, SW, GEO>(T, S, lo_e + e_begin, lo_e + e_end, fbuf, sh);
    }
}

// ------------------------------------------------------------------------------------------------
// observer side on tile lists: per-member energies (double accumulation), one warp per (supercluster, 8 entries)
template <bool PER>
__global__ void __launch_bounds__(kBlock) tile_energy_kernel(const PairListDev* __restrict__ lists, int li, StepParams P, double* __restrict__ egroup) {
    const PairListDev& L = lists[li];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ti = lane & 7, tj = lane >> 3;
    const int S = blockIdx.x;
    double Em[kMaxMembers] = {0, 0, 0, 0};
    if (S * 64 < list_rows(L)) {
        const int cnt = L.tl_count[S];
        const uint2* row = L.tl_entries + (size_t)S * L.tl_stride;
        const int nt = L.ntypes;
        for (int e = warp; e < cnt; e += kBlock / 32) {
            const uint2 ent = row[e];
            const int kj = (int)ent.x * 4 + tj;
            const uint4 pj = L.spos[kj];
            const int sbj = L.sbead[kj];
            const unsigned int act = ent.y & 0xffu, midx = ent.y >> 8;
            for (int ci = 0; ci < 8; ci++) {
                if (!((act >> ci) & 1u)) continue;
                const int ki = S * 64 + ci * 8 + ti;
                const uint4 pi = L.spos[ki];
                const int sbi = L.sbead[ki];
                bool ok = sbi >= 0 && sbj >= 0;
                if (midx) ok = ok && ((L.tl_masks[(size_t)(midx - 1u) * 8 + ci] >> lane) & 1u);
                if (!ok) continue;
                float dx, dy, dz;
                delta<PER>(pi, pj, P.fscale, dx, dy, dz);
                const float r2 = norm2(dx, dy, dz);
                // slab decomposition: a pair of beads owned by two ranks is listed on both; each reports half
                double wgt = 1.0;
                if (L.role) wgt = 0.5 * ((L.role[sbi & 0xFFFFFF] == 1 ? 1.0 : 0.0) + (L.role[sbj & 0xFFFFFF] == 1 ? 1.0 : 0.0));
                if (L.form == SEG_PAIR_DH) {
                    if (r2 < L.rc2) {
                        const float ir = rsqrtf(r2), r = r2 * ir;
                        Em[0] += wgt * (double)(__uint_as_float(pi.w) * __uint_as_float(pj.w) * (__expf(-r * L.inv_debye) * ir - L.dh_cc));
                    }
                } else {
                    const int tix = (int)((unsigned int)sbi >> 24) * nt + (int)((unsigned int)sbj >> 24);
                    for (int m = 0; m < L.nmem; m++) {
                        if (!(r2 < L.mem_rc2[m])) continue;
                        const float4 cin = L.memE[m][2 * tix], co = L.memE[m][2 * tix + 1];
                        const bool inner = r2 < cin.w;
                        const float c12 = inner ? cin.x : co.x, c6 = inner ? cin.y : co.y, c0 = inner ? cin.z : co.z;
                        const float ir2 = 1.0f / r2, ir6 = ir2 * ir2 * ir2;
                        float E = (c12 * ir6 - c6) * ir6 + c0;
                        if (L.mem_switch[m] && r2 > L.rsw * L.rsw) {
                            const float t = (sqrtf(r2) - L.rsw) * L.inv_sw;
                            E *= 1.f + t * t * t * (-10.f + t * (15.f - 6.f * t));
                        }
                        Em[m] += wgt * (double)E;
                    }
                }
            }
        }
    }
    const int nm = (L.form == SEG_PAIR_DH) ? 1 : L.nmem;
    for (int m = 0; m < nm; m++) block_add(Em[m], &egroup[L.mem_group[m]]);
}

// pairs (i < j by bead index) of a tile list inside rc2_term, appended unordered (host sorts); canonical distance
template <bool PER>
__global__ void tile_extract_pairs_kernel(const PairListDev* lists, int li, StepParams P, float rc2_term, int* __restrict__ out, long long capacity,
                                          unsigned long long* __restrict__ counter) {
    const PairListDev& L = lists[li];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ti = lane & 7, tj = lane >> 3;
    const int S = blockIdx.x;
    if (S * 64 >= list_rows(L)) return;
    const int cnt = L.tl_count[S];
    const uint2* row = L.tl_entries + (size_t)S * L.tl_stride;
    for (int e = warp; e < cnt; e += blockDim.x / 32) {
        const uint2 ent = row[e];
        const int kj = (int)ent.x * 4 + tj;
        const uint4 pj = L.spos[kj];
        const int sbj = L.sbead[kj];
        const unsigned int act = ent.y & 0xffu, midx = ent.y >> 8;
        for (int ci = 0; ci < 8; ci++) {
            if (!((act >> ci) & 1u)) continue;
            const int ki = S * 64 + ci * 8 + ti;
            const int sbi = L.sbead[ki];
            bool ok = sbi >= 0 && sbj >= 0;
            if (midx) ok = ok && ((L.tl_masks[(size_t)(midx - 1u) * 8 + ci] >> lane) & 1u);
            if (!ok) continue;
            float dx, dy, dz;
            delta<PER>(L.spos[ki], pj, P.fscale, dx, dy, dz);
            if (norm2(dx, dy, dz) < rc2_term) {
                const int bi = sbi & 0xFFFFFF, bj = sbj & 0xFFFFFF;
                const unsigned long long pos = atomicAdd(counter, 1ull);
                if (out && (long long)pos < capacity) { out[2 * pos] = min(bi, bj); out[2 * pos + 1] = max(bi, bj); }
            }
        }
    }
}

// number of (i-cluster, j-group) tiles the force launch evaluates (32 lane pairs each)
__global__ void tile_count_kernel(const PairListDev* lists, int li, unsigned long long* __restrict__ counter) {
    const PairListDev& L = lists[li];
    const int S = blockIdx.x;
    if (S * 64 >= list_rows(L)) return;
    const int cnt = L.tl_count[S];
    const uint2* row = L.tl_entries + (size_t)S * L.tl_stride;
    unsigned long long n = 0;
    for (int e = threadIdx.x; e < cnt; e += blockDim.x) n += __popc(row[e].y & 0xffu);
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if ((threadIdx.x & 31) == 0 && n) atomicAdd(counter, n);
}

}  // namespace ocm
