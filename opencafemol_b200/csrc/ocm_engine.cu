// ocm_engine.cu — host orchestration + C ABI (include/ocm_b200.h) of the B200 engine.
//
// Host side of what OpenMM's Context/Integrator/Platform layer is to the reference
// (src/Simulator.cpp:14-17,41,60; src/Observer.cpp:34,107,113,340,344): it compiles the term tables of
// an ocm_system into device-resident tables, plans the pair lists (fusing compatible pair terms),
// builds Verlet lists, and drives the per-step kernels through CUDA graphs.
//
// Stepping model: a "window" = [rebuild every pair list] + nst x [force_kernel, integrate_kernel],
// captured once per window length as a CUDA graph.  The integrate kernel flags any bead that leaves
// the Verlet buffer (displacement > skin/2) before the window's last step; on a flag the engine rolls
// back to the backup taken at the start of the batch and replays with a shorter window.  Noise is
// counter-based (Philox keyed by seed, counter = bead/step), so a replay reproduces the trajectory.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "host_system.hpp"
#include "host/lbfgs.hpp"
#include "kernels.cuh"
#include "tile_kernels.cuh"
#include "replica_kernels.cuh"
#include "dd_kernels.cuh"
#include "dd_comm.hpp"

namespace ocm {
thread_local std::string g_err;

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " + \
                                     __FILE__ + ":" + std::to_string(__LINE__) + " (" #expr ")");   \
        }                                                                                           \
    } while (0)

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0, cap = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; cap = 0; }
    // (contents are undefined afterwards.  An allocation that still fits is kept: a Monte Carlo barostat lays the cell grids out
    // again at every trial.  cudaFree waits for the device; the kept buffer does the same, so that a host copy into it cannot
    // overtake a kernel that still reads the old contents)
    void alloc(size_t count) {
        if (p && count && count <= cap && 2 * count >= cap) { CUDA_TRY(cudaDeviceSynchronize()); n = count; return; }
        release();
        n = count; cap = count;
        if (count) CUDA_TRY(cudaMalloc(&p, count * sizeof(T)));
    }
    // synchronous (the engine's kernels run on a non-blocking stream, which does not order with the legacy stream: an
    // asynchronous memset there could race with them); zero_async takes the stream that will use the buffer
    void zero() { if (n) CUDA_TRY(cudaMemset(p, 0, n * sizeof(T))); }
    void zero_async(cudaStream_t s) { if (n) CUDA_TRY(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
    void upload(const std::vector<T>& h) {
        alloc(h.size());
        if (n) CUDA_TRY(cudaMemcpy(p, h.data(), n * sizeof(T), cudaMemcpyHostToDevice));
    }
};

const double kBoltz = 8.31446261815324e-3;  // kJ/mol/K

struct PairListHost {
    std::vector<int> terms;  // member pair terms (indices into sys.pairs)
    int form = 0;            // SEG_PAIR_TZ / SEG_PAIR_DH / SEG_PAIR_WELL
    int np = 0, ntypes = 0, stride = 0, tpb = 8;
    int row_len32 = 0;                      // longest row (with head-room) rounded to 32: what pick_tpb looks at; stride is rounded to 128
    double rc_list = 0;      // largest effective cutoff of the members
    std::vector<double> rc_eff;  // per member
    // device storage
    DevBuf<int> part, excl_off, excl_idx, cell_of_slot, cell_count, cell_start, cell_cursor, scan_tmp, tmp_slots, sorted_slots,
        rank_of_bead, bead_of_rank, nbr_count, nbr_inner;
    DevBuf<uint4> spos;
    DevBuf<float> qscaled;
    DevBuf<unsigned char> type;
    DevBuf<signed char> cls;
    DevBuf<unsigned int> nbr;
    DevBuf<GridDev> grid;
    DevBuf<float4> tabF;
    DevBuf<float4> memE[kMaxMembers];
    PairListDev dev{};       // host mirror of the device descriptor
    int max_cells = 0;
    // cluster-pair ("tile") form (tile_kernels.cuh)
    bool tile_ok = false;    // the static tile tables exist (periodic or virtual box, single system)
    bool tile_auto = false;  // large enough for tiles to be the default (option pair_kernel = -1)
    bool tiled = false;      // the list is built and evaluated in tile form
    bool t_switch = false;
    int np_pad = 0, tl_stride = 0, tl_nsc = 0, tl_mask_cap = 0;
    int grid_dims[3] = {0, 0, 0};
    float cell_width = 0.f;  // width the grid was derived from (before any widening), reported for the oracle
    long long cell_cap = -1; // cap on the number of cells that widened them (-1: none)
    std::vector<int> cell_inv;  // Hilbert rank -> linear cell id
    DevBuf<int> cell_rank, sbead, tl_count, tl_weight, tl_start, tl_count_in, tl_weight_in, tl_start_in, tl_mask_count;
    TileLaunch launch_in{};  // slab decomposition: the launch over the interior entries (no halo data needed)
    int tile_nblocks_in = 0;
    TileLaunch launch{};     // by-value description of the list's force launch
    int tile_nblocks = 0;
    DevBuf<float2> sparam, slot_b;
    DevBuf<float> slot_a;
    DevBuf<uint4> sc_center, saux;
    DevBuf<float4> sc_half;
    DevBuf<uint2> tl_entries;
    DevBuf<unsigned int> tl_masks;
    DevBuf<int2> excl_pairs;
};

}  // namespace ocm

using namespace ocm;

// slab decomposition state of one rank (dd_kernels.cuh)
struct DDState {
    DDComm* comm = nullptr;
    int rank = 0, world = 1, axis = 1;
    DDGeom G{};
    double halo = 0;
    int cap = 0;                                  // entries per halo message
    DevBuf<unsigned char> role, role_bak, role_tmp, known;
    DevBuf<char> a_send[2], a_recv[2];            // phase A: [count | idx | x | v], lo / hi
    DevBuf<char> b_send[2], b_recv[2];            // phase B: [count | idx]
    DevBuf<double> s_send[2], s_recv[2];          // per-step positions
    DevBuf<double> gx; DevBuf<float> gv; DevBuf<long long> gf;  // state gather
    std::vector<int> rows_est;                    // per pair list: rows in use at the last sizing pass
    // compact work lists of this rank, rebuilt with the pair lists: beads present (owned + halo), bonded terms with an owned bead
    DevBuf<int> present, counts;                  // counts: [0] beads present, [1 + kind] bonded terms
    DevBuf<int> bterms[B_NUM_KINDS];
    int present_est = 0, bterm_est[B_NUM_KINDS] = {0};
    int integrate_cover(int n) const { return present_est > 0 ? std::min(n, (int)(1.3 * present_est) + 1024) : n; }
    int bonded_cover(int kind, int nterms) const { return bterm_est[kind] > 0 ? std::min(nterms, (int)(1.3 * bterm_est[kind]) + 1024) : nterms; }
    double n_exchanges = 0;
    // per-step halo over peer memory (dd_push_kernel): [4 flags x 128 B][side 0/1][parity 0/1][cap x 3 doubles]
    DevBuf<char> push;
    DevBuf<unsigned int> ticket;
    char* peer_lo = nullptr; char* peer_hi = nullptr;  // the neighbours' push buffers, mapped here
    bool push_ok = false, push_wanted = true;
    unsigned int seq = 0;
    size_t push_flag_off(int side, int parity) const { return (size_t)(side * 2 + parity) * 128; }
    size_t push_data_off(int side, int parity) const { return 512 + (size_t)(side * 2 + parity) * bytes_s(); }
    size_t bytes_push() const { return 512 + 4 * bytes_s(); }
    size_t bytes_a() const { return 16 + (size_t)cap * (4 + 24 + 16); }
    size_t bytes_b() const { return 16 + (size_t)cap * 4; }
    size_t bytes_s() const { return (size_t)cap * 24; }
    static DDMsg msg(char* base, int cap) {
        DDMsg m;
        m.count = reinterpret_cast<int*>(base);
        m.idx = reinterpret_cast<int*>(base + 16);
        m.x = reinterpret_cast<double*>(base + 16 + (size_t)cap * 4);
        m.v = reinterpret_cast<float4*>(base + 16 + (size_t)cap * 28);
        return m;
    }
};

struct ocm_context {
    ocm_system sys;  // host copy of the term tables
    int n = 0;
    bool periodic = false;
    double T = 0, gamma = 0, dt = 0;
    uint64_t seed = 0;
    int64_t step = 0;
    int counters_minimize_evals = 0;
    int device = 0;
    int nrep = 1, nper = 0;                 // replica batch: nrep independent copies of nper beads
    std::vector<uint64_t> rep_seeds;
    std::unique_ptr<DDState> dd;            // slab decomposition (null: the whole system lives on this device)
    DevBuf<double> origin;                  // [3*nrep] local-frame origins (replica batch)
    DevBuf<uint2> d_rep_seed;
    cudaStream_t stream = nullptr;
    StepParams P{};

    DevBuf<double> x, xref, xbak;
    DevBuf<float4> v, vbak;
    DevBuf<uint4> posq;
    DevBuf<unsigned long long> fbuf;
    DevBuf<int> flags;
    DevBuf<unsigned long long> step_base;
    DevBuf<double> egroup;  // 32 group energies + [32] kinetic
    DevBuf<unsigned long long> pair_counter;

    std::vector<std::unique_ptr<PairListHost>> lists;
    std::vector<int> term_list, term_member;  // pair term -> (list, member)
    DevBuf<PairListDev> d_lists;
    int n_coms = 0;                         // centre-of-mass pulling terms (com_kernel: one block each)
    DevBuf<ComDev> d_coms;
    DevBuf<int> com_idx;
    DevBuf<double> com_w;
    BondedDev bonded_h[B_NUM_KINDS]{};
    DevBuf<int> b_idx[B_NUM_KINDS], b_meta[B_NUM_KINDS];
    DevBuf<double> b_par[B_NUM_KINDS], b_tabs;
    DevBuf<BondedDev> d_bonded;
    // donor x acceptor terms
    std::vector<DATermDev> da_terms_h;
    std::vector<std::unique_ptr<DevBuf<int>>> da_idx;
    std::vector<std::unique_ptr<DevBuf<double>>> da_par;
    std::vector<std::unique_ptr<DevBuf<unsigned long long>>> da_excl;
    DevBuf<DATermDev> d_da_terms;
    DevBuf<unsigned int> da_count;
    DevBuf<uint2> da_entries;
    DAListDev da{};
    double da_rcmax = 0;
    std::vector<WorkSeg> segs;
    DevBuf<WorkSeg> d_segs;
    int tile_blocks = 0;                    // blocks of all tile force launches (tile lists have one launch each)
    // observer staging (ocm_get_state_begin / _end, ocm_set_state_async): device block + pinned host mirrors
    DevBuf<double> obs_dev[2], in_dev;      // two observer blocks: a download may still run while the next frame is staged
    double* obs_pin[2] = {nullptr, nullptr}; double* in_pin = nullptr;
    size_t obs_pin_n[2] = {0, 0}, in_pin_n = 0;
    cudaEvent_t ev_obs = nullptr, ev_obs_done[2] = {nullptr, nullptr}, ev_in_free = nullptr;
    bool obs_busy[2] = {false, false}, obs_used[2] = {false, false}, in_pending = false;
    int obs_next = 0;
    double n_async_states = 0, n_obs_rebuilds_skipped = 0;
    // whole-step persistent kernel for small unlimited-boundary systems and replica ensembles (replica_kernels.cuh)
    ocm_system sys1;                        // the system of ONE replica (replica batches keep it beside the replicated tables)
    bool rep_ok = false;
    int rep_block = 256;                    // threads per block of the replica kernel (compile_replica_tables)
    int rep_opt = 1;                        // option "replica_kernel": 1 (default) = use it where possible, 0 = general path
    int rep_block_opt = 0;                  // option "replica_block": threads per block (256 / 512 / 1024), 0 = by ensemble size
    int rep_carveout_opt = -1;              // option "replica_carveout": shared-memory carve-out in per cent, -1 = what the resident blocks need
    RepParams rep{};
    size_t rep_smem = 0;
    std::vector<std::unique_ptr<DevBuf<unsigned char>>> rep_blobs;   // its static tables (compile_replica_tables)
    double n_rep_launches = 0;
    int tile_opt = -1;                      // option "pair_kernel": 1 = cluster-pair tiles where possible, 0 = Verlet rows, -1 = by size
    int tile_debug = 0;
    int tile_chunk_opt = 0;                 // blocks of the tile force launch per list, 0 = auto (one wave)
    int total_blocks = 0, pair_blocks = 0, local_blocks = 0, n_pair_segs = 0, n_local_segs = 0;
    size_t smem_bytes = 0;
    cudaStream_t side = nullptr;            // forked branch: bonded + donor-acceptor kernel beside the pair kernel
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_halo_go = nullptr, ev_halo_done = nullptr, ev_rb_fork = nullptr, ev_rb_join = nullptr;
    bool halo_overlap = true;               // option "halo_overlap": interior tile entries run while the halo positions travel

    // options
    double skin = 0.3;
    int nstlist_fixed = 0;   // 0 = adaptive
    // Unlimited boundary, single system: the pair kernels run in a VIRTUAL periodic box laid out around the beads, large
    // enough that no periodic image ever comes within a list radius: positions are then 32-bit fixed-point fractions of
    // that box (uniform resolution L / 2^32 ~ 1e-8 nm) instead of FP32 coordinates, whose resolution at 60 nm from the
    // origin (4e-6 nm) costs 1e-4 in the r^-12 forces.  Bonded and donor-acceptor terms keep their plain differences.
    bool vbox = false;
    double vb_lo[3] = {0, 0, 0}, vb_hi[3] = {0, 0, 0};  // beads must stay inside (vbox_check_kernel at every rebuild)
    int n_vbox_resize = 0;
    // optional phase timing (option "phase_timing"): events between the phases of directly launched windows
    bool phase_timing = false;
    bool pdl = false;        // programmatic dependent launch of the per-step kernels (option "pdl")
    std::vector<cudaEvent_t> ph_ev;
    std::vector<int> ph_tag;
    size_t ph_used = 0;
    double ph_ms[4] = {0, 0, 0, 0};  // rebuild, force, integrate, halo exchange
    int nst = 10;
    int tpb_opt = 0;         // 0 = auto
    int chunk_opt = 0;       // beads per block of the pair kernels, 0 = auto
    bool lists_ready = false, positions_set = false, lists_current = false;
    bool lists_fresh = false;    // lists_current because the lists were built at exactly the present positions (ensure_lists_current)
    bool last_disp2_ok = false;  // every window of the last batch ended with all beads inside their Verlet buffer
    std::map<int, cudaGraphExec_t> graphs;
    // counters
    double n_rebuilds = 0, n_rollbacks = 0, n_launches = 0, n_windows = 0, last_step_ms = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // Monte Carlo barostat (ocm_context_set_barostat)
    struct Barostat {
        int kind = 0;                        // OCM_BAROSTAT_*
        double pressure[3] = {0, 0, 0};      // bar
        double tension = 0, T = 0;           // bar nm, K
        int mode[3] = {0, 0, 0};             // anisotropic: scale X, Y, Z; membrane: xymode, zmode
        int frequency = 25;
        uint64_t rng = 0;                    // SplitMix64 state
        double volume_scale[3] = {0, 0, 0};  // trial step of each axis (nm^3)
        int attempted[3] = {0, 0, 0}, accepted[3] = {0, 0, 0};
        int64_t until_trial = 0;             // MD steps that may still run before the next trial
        double n_trials = 0, n_accepted = 0;
        double last[6] = {-1, 0, 0, 0, 0, 0};  // axis, dV, E before, E after, w, accepted
        int nmol = 0;
        DevBuf<int> mol_start, mol_bead;
        DevBuf<double> xsave;
    } baro;

    ~ocm_context() {
        for (auto& kv : graphs) cudaGraphExecDestroy(kv.second);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (ev_obs) cudaEventDestroy(ev_obs);
        for (int b = 0; b < 2; b++) { if (ev_obs_done[b]) cudaEventDestroy(ev_obs_done[b]); if (obs_pin[b]) cudaFreeHost(obs_pin[b]); }
        if (ev_in_free) cudaEventDestroy(ev_in_free);
        if (in_pin) cudaFreeHost(in_pin);
        if (ev_halo_go) cudaEventDestroy(ev_halo_go);
        if (ev_halo_done) cudaEventDestroy(ev_halo_done);
        if (ev_rb_fork) cudaEventDestroy(ev_rb_fork);
        if (ev_rb_join) cudaEventDestroy(ev_rb_join);
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_join) cudaEventDestroy(ev_join);
        if (side) cudaStreamDestroy(side);
        if (stream) cudaStreamDestroy(stream);
    }
};

namespace ocm {

static const int kTileAutoBeads = 131072;
static const int kIntegrateEarlyBeads = 262144;   // below: integrate_kernel<true> (loads hoisted: latency-bound sizes)
static const int kNstChoices[] = {1, 2, 3, 4, 5, 6, 8, 10, 12, 14, 16, 20, 24, 28, 32, 40, 48, 56, 64, 80, 100};

static void dd_configure(ocm_context* c);

static void drop_graphs(ocm_context* c) {
    for (auto& kv : c->graphs) cudaGraphExecDestroy(kv.second);
    c->graphs.clear();
}

// ---- two-zone 12-6 coefficient sets ------------------------------------------------------------
struct TZ { double c12i, c6i, c0i, a2, c12o, c6o, c0o; };

static TZ tz_coeffs(const PairTerm& t, int i, int j) {
    TZ z{0, 0, 0, INFINITY, 0, 0, 0};
    switch (t.kind) {
        case P_EXV: {  // eps (s_i+s_j)^12 (r^-12 - rc^-12)
            double A = t.g0 * std::pow(t.p0[i] + t.p0[j], 12);
            z.c12i = A; z.c0i = -A * t.g1;
        } break;
        case P_LJREP: {
            double sg = 0.5 * (t.p1[i] + t.p1[j]), ep = 0.5 * (t.p0[i] + t.p0[j]);
            double th = sg * std::pow(2.0, 1.0 / 6.0);
            z.c12i = 4 * ep * std::pow(sg, 12); z.c6i = 4 * ep * std::pow(sg, 6); z.c0i = ep; z.a2 = th * th;
        } break;
        case P_WCA: {  // as LJ-repulsive, with the geometric mean of epsilon
            double sg = 0.5 * (t.p1[i] + t.p1[j]), ep = std::sqrt(t.p0[i] * t.p0[j]);
            double th = sg * std::pow(2.0, 1.0 / 6.0);
            z.c12i = 4 * ep * std::pow(sg, 12); z.c6i = 4 * ep * std::pow(sg, 6); z.c0i = ep; z.a2 = th * th;
        } break;
        case P_LJATT: {
            double sg = 0.5 * (t.p1[i] + t.p1[j]), ep = 0.5 * (t.p0[i] + t.p0[j]);
            double th = sg * std::pow(2.0, 1.0 / 6.0);
            z.c0i = -ep; z.a2 = th * th;
            z.c12o = 4 * ep * std::pow(sg, 12); z.c6o = 4 * ep * std::pow(sg, 6);
        } break;
        case P_3SPN2EXV: {
            double sg = 0.5 * (t.p1[i] + t.p1[j]), ep = std::sqrt(t.p0[i] * t.p0[j]);
            z.c12i = ep * std::pow(sg, 12); z.c6i = 2 * ep * std::pow(sg, 6); z.c0i = ep; z.a2 = sg * sg;
        } break;
        default: break;
    }
    return z;
}

static bool same_topology(const PairTerm& a, const PairTerm& b) {
    return a.periodic == b.periodic && a.part == b.part && a.excl == b.excl && a.cls == b.cls && a.allow == b.allow;
}

// ---- list planning: which pair terms share one Verlet list ------------------------------------
static void plan_lists(ocm_context* c) {
    const auto& pairs = c->sys.pairs;
    c->term_list.assign(pairs.size(), -1);
    c->term_member.assign(pairs.size(), -1);
    for (size_t t = 0; t < pairs.size(); t++) {
        const PairTerm& pt = pairs[t];
        int target = -1;
        if (pt.kind == P_LJREP || pt.kind == P_LJATT) {
            // Ashbaugh-Hatch: LJ-repulsive + LJ-attractive with identical sigma (same zone boundary)
            for (size_t l = 0; l < c->lists.size() && target < 0; l++) {
                auto& L = *c->lists[l];
                if (L.form != SEG_PAIR_TZ || L.terms.size() != 1) continue;
                const PairTerm& o = pairs[L.terms[0]];
                bool partner = (o.kind == P_LJREP && pt.kind == P_LJATT) || (o.kind == P_LJATT && pt.kind == P_LJREP);
                if (!partner || !same_topology(o, pt)) continue;
                bool same_sigma = true;
                const PairTerm& rep = (o.kind == P_LJREP) ? o : pt;
                const PairTerm& att = (o.kind == P_LJREP) ? pt : o;
                // the fused table stores the repulsive coefficients and the ratio eps_att/eps_rep
                for (int i : pt.part) if (o.p1[i] != pt.p1[i] || !(rep.p0[i] > 0)) { same_sigma = false; break; }
                // the repulsive term must be zero beyond its zone inside both cutoffs, and the switch
                // must start outside the zone
                double zone = effective_rc(rep);
                if (same_sigma && zone <= rep.rc && zone <= att.rsw) target = (int)l;
            }
        }
        if (target < 0) {
            c->lists.emplace_back(new PairListHost());
            target = (int)c->lists.size() - 1;
            c->lists[target]->form = (pt.kind == P_DH) ? SEG_PAIR_DH : ((pt.kind == P_ISOLF || pt.kind == P_TRIG) ? SEG_PAIR_WELL : SEG_PAIR_TZ);
        }
        c->term_list[t] = target;
        c->term_member[t] = (int)c->lists[target]->terms.size();
        c->lists[target]->terms.push_back((int)t);
    }
}

// ---- cluster-pair ("tile") form: static tables ------------------------------------------------------
// Hilbert index of a grid cell (Skilling's transpose algorithm), `bits` bits per axis
static uint64_t hilbert_index(uint32_t x, uint32_t y, uint32_t z, int bits) {
    uint32_t X[3] = {x, y, z};
    const uint32_t M = 1u << (bits - 1);
    for (uint32_t Q = M; Q > 1; Q >>= 1) {
        const uint32_t Pm = Q - 1;
        for (int i = 0; i < 3; i++) {
            if (X[i] & Q) X[0] ^= Pm;
            else { const uint32_t t = (X[0] ^ X[i]) & Pm; X[0] ^= t; X[i] ^= t; }
        }
    }
    for (int i = 1; i < 3; i++) X[i] ^= X[i - 1];
    uint32_t t = 0;
    for (uint32_t Q = M; Q > 1; Q >>= 1) if (X[2] & Q) t ^= Q - 1;
    for (int i = 0; i < 3; i++) X[i] ^= t;
    uint64_t h = 0;
    for (int b = bits - 1; b >= 0; b--)
        for (int i = 0; i < 3; i++) h = (h << 1) | ((X[i] >> b) & 1u);
    return h;
}

// per-bead mixing parameters of the two-zone 12-6 family (tile_pair_tz): with s = a_i + a_j, q = s^2 / r^2,
//   -dE/dr / r = e q^3 (alpha q^3 - beta) / r^2,  e = inner zone (q > qz) ? mix(b_i, b_j) : c_i + c_j
// which restates tz_coeffs() term by term:
//   ExcludedVolume      eps (s_i+s_j)^12 / r^12                 a = radius,   b = 6 eps (sum),        alpha 1, beta 0, qz 0
//   LennardJonesRepulsive / WCA   4 eps (q^6 - q^3) + eps inside 2^(1/6) sigma   a = sigma/2, b = 12 eps_i (sum) or sqrt(24 eps_i) (product)
//   LennardJonesAttractive        4 eps (q^6 - q^3) beyond 2^(1/6) sigma, switched   c = 12 eps_i
//   3SPN2ExcludedVolume eps (q^6 - 2 q^3) + eps inside sigma    a = sigma/2,  b = sqrt(12 eps_i) (product), alpha 1, beta 1, qz 1
static void compile_tile_tables(ocm_context* c, PairListHost& L) {
    // (small lists -- a few hundred charged beads -- stay Verlet rows inside the fused launch: a tile launch of their own costs more)
    L.tile_ok = c->periodic && c->nrep == 1 && L.np >= 1024 && L.form != SEG_PAIR_WELL;   // (wells: Verlet rows only)
    L.tile_auto = L.np >= kTileAutoBeads;
    if (!L.tile_ok) return;
    const auto& pairs = c->sys.pairs;
    const PairTerm& t0 = pairs[L.terms[0]];
    PairListDev& D = L.dev;
    std::vector<float> sa(L.np, 0.f);
    std::vector<float2> sb(L.np, make_float2(0.f, 0.f));
    D.t_alpha = 2.f; D.t_beta = 1.f; D.t_qz = (float)std::pow(2.0, -1.0 / 3.0); D.t_kappa = (float)(1.0 / 6.0); D.t_geo = 0;
    L.t_switch = false;
    if (L.form == SEG_PAIR_DH) {
        const double sk = std::sqrt(t0.g0);
        for (int a = 0; a < L.np; a++) sa[a] = (float)(t0.p0[t0.part[a]] * sk);
    } else {
        for (int m : L.terms) {
            const PairTerm& t = pairs[m];
            for (int a = 0; a < L.np; a++) {
                const int i = t0.part[a];
                switch (t.kind) {
                    case P_EXV: sa[a] = (float)t.p0[i]; sb[a].x = (float)(6.0 * t.g0); break;
                    case P_LJREP: sa[a] = (float)(0.5 * t.p1[i]); sb[a].x = (float)(12.0 * t.p0[i]); break;
                    case P_WCA: sa[a] = (float)(0.5 * t.p1[i]); sb[a].x = (float)std::sqrt(24.0 * t.p0[i]); break;
                    case P_LJATT: sa[a] = (float)(0.5 * t.p1[i]); sb[a].y = (float)(12.0 * t.p0[i]); break;
                    case P_3SPN2EXV: sa[a] = (float)(0.5 * t.p1[i]); sb[a].x = (float)std::sqrt(12.0 * t.p0[i]); break;
                    default: throw std::runtime_error("internal: pair kind without a tile form");
                }
            }
            if (t.kind == P_EXV) { D.t_alpha = 1.f; D.t_beta = 0.f; D.t_qz = 0.f; }
            if (t.kind == P_3SPN2EXV) { D.t_alpha = 1.f; D.t_beta = 1.f; D.t_qz = 1.f; D.t_geo = 1; }
            if (t.kind == P_WCA) D.t_geo = 1;
            if (t.kind == P_LJATT) L.t_switch = true;
        }
    }
    L.slot_a.upload(sa); L.slot_b.upload(sb);
    D.slot_a = L.slot_a.p; D.slot_b = L.slot_b.p;
    std::vector<int2> ex;
    {
        std::vector<char> in_list(c->n, 0);
        for (int i : t0.part) in_list[i] = 1;
        for (auto& e : t0.excl) if (in_list[e.first] && in_list[e.second]) ex.push_back(make_int2(e.first, e.second));
    }
    D.n_excl_pairs = (int)ex.size();
    if (ex.empty()) ex.push_back(make_int2(0, 0));
    L.excl_pairs.upload(ex);
    D.excl_pairs = L.excl_pairs.p;
    // sorted arrays padded to whole superclusters
    L.np_pad = (L.np + 63) / 64 * 64;
    L.tl_nsc = L.np_pad / 64;
    L.spos.alloc(L.np_pad); L.spos.zero();
    L.sbead.alloc(L.np_pad); CUDA_TRY(cudaMemset(L.sbead.p, 0xFF, (size_t)L.np_pad * sizeof(int)));
    L.sparam.alloc(L.np_pad); L.sparam.zero();
    L.saux.alloc(L.np_pad); L.saux.zero();
    L.sc_center.alloc(L.tl_nsc); L.sc_half.alloc(L.tl_nsc); L.tl_count.alloc(L.tl_nsc); L.tl_count.zero(); L.tl_start.alloc(L.tl_nsc + 1); L.tl_start.zero(); L.tl_weight.alloc(L.tl_nsc); L.tl_weight.zero();
    L.tl_count_in.alloc(L.tl_nsc); L.tl_count_in.zero(); L.tl_weight_in.alloc(L.tl_nsc); L.tl_weight_in.zero(); L.tl_start_in.alloc(L.tl_nsc + 1); L.tl_start_in.zero();
    L.tl_mask_count.alloc(1); L.tl_mask_count.zero();
    D.spos = L.spos.p; D.sbead = L.sbead.p; D.sparam = L.sparam.p; D.saux = L.saux.p; D.sc_center = L.sc_center.p; D.sc_half = L.sc_half.p;
    D.tl_nsc = L.tl_nsc; D.tl_count = L.tl_count.p; D.tl_start = L.tl_start.p; D.tl_weight = L.tl_weight.p; D.tl_count_in = L.tl_count_in.p; D.tl_weight_in = L.tl_weight_in.p; D.tl_start_in = L.tl_start_in.p; D.tl_mask_count = L.tl_mask_count.p;
}

static void compile_list(ocm_context* c, PairListHost& L) {
    const auto& pairs = c->sys.pairs;
    const PairTerm& t0 = pairs[L.terms[0]];
    L.np = (int)t0.part.size();
    const int nm = (int)L.terms.size();
    // types: unique tuples of the members' per-bead parameters
    std::map<std::vector<double>, int> ids;
    std::vector<int> rep_bead;
    std::vector<unsigned char> type(L.np);
    for (int a = 0; a < L.np && L.form != SEG_PAIR_DH; a++) {
        int i = t0.part[a];
        std::vector<double> key;
        for (int m = 0; m < nm; m++) { key.push_back(pairs[L.terms[m]].p0[i]); key.push_back(pairs[L.terms[m]].p1[i]); key.push_back(pairs[L.terms[m]].p2[i]); }
        auto it = ids.find(key);
        if (it == ids.end()) {
            if (ids.size() >= 255) throw std::runtime_error("pair term with more than 255 distinct per-bead parameter sets is not supported");
            it = ids.emplace(key, (int)ids.size()).first;
            rep_bead.push_back(i);
        }
        type[a] = (unsigned char)it->second;
    }
    L.ntypes = std::max<int>(1, (int)ids.size());
    const int nt = L.ntypes, nt2 = nt * nt;
    L.type.upload(type);
    L.part.upload(std::vector<int>(t0.part.begin(), t0.part.end()));
    // exclusions CSR per slot
    {
        std::vector<int> slot_of(c->n, -1);
        for (int a = 0; a < L.np; a++) slot_of[t0.part[a]] = a;
        std::vector<std::vector<int>> ex(L.np);
        for (auto& e : t0.excl) {
            int a = slot_of[e.first], b = slot_of[e.second];
            if (a < 0 || b < 0) continue;
            ex[a].push_back(e.second); ex[b].push_back(e.first);
        }
        std::vector<int> off(L.np + 1, 0), idx;
        for (int a = 0; a < L.np; a++) {
            std::sort(ex[a].begin(), ex[a].end());
            off[a + 1] = off[a] + (int)ex[a].size();
            idx.insert(idx.end(), ex[a].begin(), ex[a].end());
        }
        if (idx.empty()) idx.push_back(0);
        L.excl_off.upload(off); L.excl_idx.upload(idx);
    }
    if (t0.ncls > 0) {
        std::vector<signed char> cl(L.np);
        for (int a = 0; a < L.np; a++) cl[a] = (signed char)t0.cls[t0.part[a]];
        L.cls.upload(cl);
    }
    // cutoffs
    L.rc_eff.clear();
    L.rc_list = 0;
    for (int m = 0; m < nm; m++) { L.rc_eff.push_back(effective_rc(pairs[L.terms[m]])); L.rc_list = std::max(L.rc_list, L.rc_eff.back()); }
    PairListDev& D = L.dev;
    std::memset(&D, 0, sizeof(D));
    D.np = L.np; D.ntypes = nt; D.ncls = t0.ncls; D.form = L.form; D.nmem = nm;
    D.nrep = c->nrep; D.nper = c->nper;
    D.role = c->dd ? c->dd->role.p : nullptr;
    for (int k = 0; k < t0.ncls && k < 32; k++) D.allow[k] = t0.allow[k];
    D.rc2 = (float)(L.rc_list * L.rc_list);
    D.rsw = (float)L.rc_list; D.inv_sw = 0.f; D.has_switch = 0;
    if (L.form == SEG_PAIR_DH) {
        // the charges ride in the w slot of the cell-sorted positions, pre-scaled so that w_i w_j = K q_i q_j
        if (!(t0.g0 >= 0)) throw std::runtime_error("debye_huckel: the Coulomb constant must not be negative");
        std::vector<float> qs(std::max(L.np, 1), 0.f);
        const double sk = std::sqrt(t0.g0);
        for (int a = 0; a < L.np; a++) qs[a] = (float)(t0.p0[t0.part[a]] * sk);
        L.qscaled.upload(qs);
        D.qscaled = L.qscaled.p;
        D.inv_debye = (float)(1.0 / t0.g1); D.dh_cc = (float)t0.g2;
        D.mem_group[0] = t0.group; D.mem_rc2[0] = D.rc2;
    } else if (L.form == SEG_PAIR_WELL) {
        // attractive wells: per type pair {base = 2^(1/6) sigma_ij, 1 / omega_ij, amp, shape} (well_dedr in kernels.cuh);
        // p0 = eps (geometric mean), p1 = sigma, p2 = omega (arithmetic means)
        if (nm != 1) throw std::runtime_error("internal: unsupported pair-list fusion");
        std::vector<float4> tf(nt2, make_float4(0, 0, 0, 0));
        const double pi = 3.1415926535897932385;
        for (int a = 0; a < nt && !rep_bead.empty(); a++)
            for (int b = 0; b < nt; b++) {
                const int i = rep_bead[a], j = rep_bead[b];
                const double ep = std::sqrt(t0.p0[i] * t0.p0[j]), sg = 0.5 * (t0.p1[i] + t0.p1[j]), om = 0.5 * (t0.p2[i] + t0.p2[j]);
                if (!(om > 0)) throw std::runtime_error("attractive well: omega must be positive");
                const double amp = (t0.kind == P_ISOLF) ? ep * pi / (2.0 * om) : 6.0 * ep / om;
                tf[a * nt + b] = make_float4((float)(sg * std::pow(2.0, 1.0 / 6.0)), (float)(1.0 / om), (float)amp, t0.kind == P_ISOLF ? 0.f : 1.f);
            }
        L.tabF.upload(tf);
        D.k_in = 1.f;
        D.mem_group[0] = t0.group; D.mem_rc2[0] = D.rc2;
    } else {
        // force table: {12 c12, 6 c6, a2, w}; zone factor = r^2 < a2 ? k_in : k_out + k_w w
        std::vector<float4> tf(nt2, make_float4(0, 0, INFINITY, 0));
        const PairTerm* rep = nullptr; const PairTerm* att = nullptr;
        if (nm == 2) {
            rep = &pairs[L.terms[0]]; att = &pairs[L.terms[1]];
            if (rep->kind != P_LJREP) std::swap(rep, att);
            D.k_in = 1.f;
        } else if (nm == 1) {
            D.k_in = (t0.kind == P_LJATT) ? 0.f : 1.f;
        } else throw std::runtime_error("internal: unsupported pair-list fusion");
        for (int a = 0; a < nt && !rep_bead.empty(); a++)
            for (int b = 0; b < nt; b++) {
                const int k = a * nt + b;
                if (nm == 2) {
                    TZ zr = tz_coeffs(*rep, rep_bead[a], rep_bead[b]), za = tz_coeffs(*att, rep_bead[a], rep_bead[b]);
                    tf[k] = make_float4((float)(12.0 * zr.c12i), (float)(6.0 * zr.c6i), (float)zr.a2, (float)(za.c12o / zr.c12i));
                } else {
                    TZ z = tz_coeffs(t0, rep_bead[a], rep_bead[b]);
                    // w = factor of the outer zone: LJ-attractive alone acts only there, the others vanish there
                    if (t0.kind == P_LJATT) tf[k] = make_float4((float)(12.0 * z.c12o), (float)(6.0 * z.c6o), (float)z.a2, 1.f);
                    else tf[k] = make_float4((float)(12.0 * z.c12i), (float)(6.0 * z.c6i), std::isfinite(z.a2) ? (float)z.a2 : INFINITY, 0.f);
                }
            }
        for (int m = 0; m < nm; m++) {
            const PairTerm& t = pairs[L.terms[m]];
            std::vector<float4> te(2 * nt2, make_float4(0, 0, 0, 0));
            for (int a = 0; a < nt && !rep_bead.empty(); a++)
                for (int b = 0; b < nt; b++) {
                    TZ z = tz_coeffs(t, rep_bead[a], rep_bead[b]);
                    int k = a * nt + b;
                    te[2 * k] = make_float4((float)z.c12i, (float)z.c6i, (float)z.c0i, (float)z.a2);
                    te[2 * k + 1] = make_float4((float)z.c12o, (float)z.c6o, (float)z.c0o, 0.f);
                }
            L.memE[m].upload(te);
            D.memE[m] = L.memE[m].p;
            D.mem_group[m] = t.group;
            D.mem_rc2[m] = (float)(t.rc * t.rc);
            D.mem_switch[m] = (t.kind == P_LJATT) ? 1 : 0;
            if (t.kind == P_LJATT) {
                D.has_switch = 1; D.rsw = (float)t.rsw; D.inv_sw = (float)(1.0 / (t.rc - t.rsw));
            }
        }
        L.tabF.upload(tf);
    }
    D.part = L.part.p; D.type = L.type.p; D.cls = L.cls.p;
    D.excl_off = L.excl_off.p; D.excl_idx = L.excl_idx.p;
    D.tabF = L.tabF.p;
    D.smem_tab = (L.form == SEG_PAIR_TZ && nt <= 64) ? 1 : 0;
    const size_t np1 = (size_t)std::max(L.np, 1);
    L.cell_of_slot.alloc(np1); L.tmp_slots.alloc(np1); L.sorted_slots.alloc(np1); L.nbr_count.alloc(np1); L.spos.alloc(np1);
    L.bead_of_rank.alloc(np1); L.nbr_inner.alloc(np1);
    L.rank_of_bead.alloc(c->n);
    L.nbr_count.zero(); L.nbr_inner.zero(); L.spos.zero(); L.bead_of_rank.zero();
    D.bead_of_rank = L.bead_of_rank.p; D.nbr_inner = L.nbr_inner.p;
    CUDA_TRY(cudaMemset(L.rank_of_bead.p, 0xFF, (size_t)c->n * sizeof(int)));
    D.cell_of_slot = L.cell_of_slot.p; D.tmp_slots = L.tmp_slots.p; D.sorted_slots = L.sorted_slots.p; D.nbr_count = L.nbr_count.p;
    D.spos = L.spos.p; D.rank_of_bead = L.rank_of_bead.p;
    compile_tile_tables(c, L);
}

// (re)derive everything that depends on the skin: list radius, cell grid
static void configure_grid(ocm_context* c, PairListHost& L) {
    PairListDev& D = L.dev;
    const double rlist = L.rc_list + c->skin;
    D.rlist2 = (float)(rlist * rlist);
    D.wmin = 0.5f * (float)rlist;
    {
        // build-time distance zones of a row (see build_list_kernel).  A is a guarantee (pairs cannot leave it before
        // the next rebuild), the others only order the row so that warp-uniform votes can skip work.
        const double r_hi = L.rc_list, r_lo = D.has_switch ? (double)D.rsw : L.rc_list, d = 0.25 * c->skin;
        const double za = std::max(0.0, r_lo - c->skin), zb = std::max(za, r_lo - d), zc = r_hi + d;
        D.zone_a2 = (L.form == SEG_PAIR_TZ) ? (float)(za * za) * (1.f - 1e-6f) : 0.f;
        D.zone_b2 = (float)(zb * zb); D.zone_c2 = (float)(zc * zc);
    }
    GridDev g{};
    for (int k = 0; k < 3; k++) { g.bmin[k] = 0x7FFFFFFF; g.bmax[k] = (int)0x80000000; g.lo[k] = 0.f; }
    if (c->periodic) {
        long long tot = 1;
        // a virtual box (unlimited boundary) is mostly empty: widen the cells until their number is of the order of the
        // participants (same rule in the oracle's restatement of the grid)
        long long cap = c->vbox ? std::max(4096, 8 * L.np + 64) : -1;
        float w = D.wmin;
        if (L.tiled) {
            // tile lists only sort by cell (the list build works on cluster bounding boxes): fine cells, a sixth of the
            // list radius, about one bead per cell at most, ordered along a Hilbert curve
            w = (float)rlist / 6.0f;
            cap = std::max(4096, L.np);
        }
        L.cell_width = w; L.cell_cap = cap;
        for (;;) {
            tot = 1;
            for (int k = 0; k < 3; k++) {
                int nc = (int)floorf((float)c->sys.box[k] / w);
                g.nc[k] = std::max(1, std::min(nc, 1024));
                tot *= g.nc[k];
            }
            if (cap < 0 || tot <= cap) break;
            w *= 1.25f;
        }
        if (tot > (1ll << 26)) throw std::runtime_error("cell grid too large");
        g.ncell = (int)tot;
        g.inv_w = 1.0f / D.wmin;
        // (slab-decomposed tile lists sort halo beads behind the owned ones: two copies of the key range)
        D.halo_split = (L.tiled && c->dd) ? 1 : 0;
        L.max_cells = g.ncell * (1 + D.halo_split);
        if (L.tiled && (L.grid_dims[0] != g.nc[0] || L.grid_dims[1] != g.nc[1] || L.grid_dims[2] != g.nc[2] || !L.cell_rank.p)) {
            // rank of every cell along the Hilbert curve through the enclosing power-of-two cube
            int bits = 1;
            while ((1 << bits) < std::max(g.nc[0], std::max(g.nc[1], g.nc[2]))) bits++;
            std::vector<std::pair<uint64_t, int>> key((size_t)g.ncell);
            for (int z = 0; z < g.nc[2]; z++)
                for (int y = 0; y < g.nc[1]; y++)
                    for (int x = 0; x < g.nc[0]; x++) {
                        const int cidx = (z * g.nc[1] + y) * g.nc[0] + x;
                        key[cidx] = {hilbert_index((uint32_t)x, (uint32_t)y, (uint32_t)z, bits), cidx};
                    }
            std::sort(key.begin(), key.end());
            std::vector<int> rank((size_t)g.ncell);
            L.cell_inv.resize((size_t)g.ncell);
            for (int r = 0; r < g.ncell; r++) { rank[key[r].second] = r; L.cell_inv[r] = key[r].second; }
            L.cell_rank.upload(rank);
            for (int k = 0; k < 3; k++) L.grid_dims[k] = g.nc[k];
        }
        D.cell_rank = L.tiled ? L.cell_rank.p : nullptr;
    } else {
        L.cell_width = D.wmin; L.cell_cap = -1;
        g.nc[0] = g.nc[1] = g.nc[2] = 1; g.ncell = 1; g.inv_w = 1.0f / D.wmin;
        L.max_cells = std::max(std::max(4096, 64 * c->nrep), std::min(1 << 21, 8 * L.np + 64));
    }
    L.grid.upload(std::vector<GridDev>(1, g));
    L.cell_count.alloc(L.max_cells + 1); L.cell_count.zero();
    L.cell_start.alloc(L.max_cells + 1); L.cell_start.zero();
    L.cell_cursor.alloc(L.max_cells);
    L.scan_tmp.alloc(L.max_cells / kScanChunk + 3);
    D.grid = L.grid.p; D.max_cells = L.max_cells;
    D.cell_count = L.cell_count.p; D.cell_start = L.cell_start.p; D.cell_cursor = L.cell_cursor.p; D.scan_tmp = L.scan_tmp.p;
}

static void upload_lists(ocm_context* c) {
    std::vector<PairListDev> h;
    for (auto& L : c->lists) h.push_back(L->dev);
    if (h.empty()) h.emplace_back();
    c->d_lists.upload(h);
}

static int pick_tpb(const ocm_context* c, const PairListHost& L) {
    if (c->tpb_opt == 4 || c->tpb_opt == 8 || c->tpb_opt == 16 || c->tpb_opt == 32) return c->tpb_opt;
    // one warp per bead unless the rows are too short to keep 32 lanes busy
    int t = 32;
    if (L.row_len32 > 0) while (t > 4 && L.row_len32 < 4 * t) t >>= 1;
    return t;
}

static void build_segments(ocm_context* c) {
    c->segs.clear();
    int blk = 0, tblk = 0;
    size_t smem = 0;
    for (size_t l = 0; l < c->lists.size(); l++) {
        auto& L = *c->lists[l];
        if (L.np == 0) continue;
        if (L.tiled) {
            // tile lists: block = (supercluster, chunk of its entry row).  About 32 entries per block (8 per warp) keeps
            // the per-block prologue / i-force epilogue below 5 % and gives some 2 000 blocks at 16 k beads.
            const int rows = (c->dd && l < c->dd->rows_est.size() && c->dd->rows_est[l] > 0) ? std::min(L.np_pad, (int)(1.25 * c->dd->rows_est[l]) + 64) : L.np_pad;
            L.dev.row_capacity = std::min(L.np_pad, (rows + 63) / 64 * 64);
            const int nsc = L.dev.row_capacity / 64;
            // persistent launch, one wave of OCM_TILE_MINBLOCKS blocks per SM (fewer when the list is small)
            int sms = 148;
            { cudaDeviceProp prop; if (cudaGetDeviceProperties(&prop, c->device) == cudaSuccess) sms = prop.multiProcessorCount; }
            const int want = c->tile_chunk_opt > 0 ? c->tile_chunk_opt : sms * OCM_TILE_MINBLOCKS;
            TileLaunch& T = L.launch;
            T = TileLaunch{};
            T.n = c->n; T.pdl_early = c->P.pdl_early; T.debug = c->tile_debug;
            for (int k = 0; k < 3; k++) { T.fscale[k] = c->P.fscale[k]; T.half_box[k] = (float)(0.5 * c->sys.box[k]); }
            TileSegP& ts = T.seg;
            ts.nblocks = std::max(1, std::min(want, nsc * 8)); ts.nsc = nsc; ts.stride = L.tl_stride; ts.start = L.tl_start.p;
            L.tile_nblocks = ts.nblocks;
            ts.entries = L.tl_entries.p; ts.count = L.tl_count.p; ts.spos = L.spos.p; ts.saux = L.saux.p; ts.masks = L.tl_masks.p; ts.sc_center = L.sc_center.p;
            ts.rc2 = L.dev.rc2; ts.alpha = L.dev.t_alpha; ts.beta = L.dev.t_beta; ts.qz = L.dev.t_qz; ts.inv_sw = L.dev.inv_sw;
            ts.sw_c0 = L.dev.rsw * L.dev.inv_sw; ts.sw_c1 = 30.f * L.dev.inv_sw * L.dev.t_kappa; ts.inv_debye = L.dev.inv_debye;
            ts.count_lo = nullptr;
            L.tile_nblocks_in = 0;
            if (L.dev.halo_split) {
                // two launches per step: the interior entries [0, count_in) while the halo positions travel, then the rest
                L.launch_in = T;
                L.launch_in.seg.count = L.tl_count_in.p; L.launch_in.seg.count_lo = nullptr; L.launch_in.seg.start = L.tl_start_in.p;
                L.tile_nblocks_in = ts.nblocks;
                ts.count_lo = L.tl_count_in.p;
            }
            tblk += L.tile_nblocks;
            continue;
        }
        L.tpb = pick_tpb(c, L);
        // a block owns `chunk` consecutive ranks of the cell-sorted order (spatially compact: rows overlap in L1,
        // one coefficient-table load per block); about 16 blocks per SM for load balance
        const int gpb = kBlock / L.tpb;
        // under a slab decomposition only the rows of the beads present on this rank are in use
        const int rows = (c->dd && l < c->dd->rows_est.size() && c->dd->rows_est[l] > 0) ? std::min(L.np, c->dd->rows_est[l]) : L.np;
        // Rows per block ~ sqrt(rows) / 12 (measured optimum on B200: 8-16 at 16 k rows, 16-32 at 130 k, 64-128 at 1 M).
        // Neighbour counts vary across a slab (dense centre, thin surface) and, under a slab decomposition, owned rows
        // alternate with empty halo rows: with the former "16 blocks per SM" rule the last blocks left SMs idle
        // (130 k rows per GPU: force launch 0.202 -> 0.169 ms).
        // (small systems keep the finer "16 blocks per SM" split: 8 rows per block at 16 k rows)
        int chunk = c->chunk_opt > 0 ? c->chunk_opt : std::min((rows + 148 * 16 - 1) / (148 * 16), (int)(std::sqrt((double)std::max(rows, 1)) / 12.0));
        chunk = std::max(gpb, (chunk + gpb - 1) / gpb * gpb);
        // (slab decomposition: 25 % head-room over the rows in use at the sizing pass; more rows than that raise the
        // overflow flag in build_list_kernel and the lists are sized again)
        int nb = ((c->dd ? std::min(L.np, (int)(1.25 * rows) + chunk) : L.np) + chunk - 1) / chunk;
        L.dev.row_capacity = c->dd ? std::min(L.np, nb * chunk) : L.np;
        // (measured, round 2: one wave of blocks owning equal numbers of 32-entry groups instead of equal numbers of rows did not
        // shorten the C2 force launch, 38.9 against 38.4 us: the idle SM cycles at its end are not a load imbalance between blocks)
        const size_t tab_bytes = (size_t)L.ntypes * L.ntypes * 16;
        c->segs.push_back({blk, nb, L.form, (int)l, L.tpb, chunk, 0});
        blk += nb;
        if (L.dev.smem_tab) smem = std::max(smem, tab_bytes);
    }
    for (int k = 0; k < B_NUM_KINDS; k++) {
        if (c->bonded_h[k].n == 0) continue;
        int nb = ((c->dd ? c->dd->bonded_cover(k, c->bonded_h[k].n) : c->bonded_h[k].n) + kBlock - 1) / kBlock;
        c->segs.push_back({blk, nb, SEG_BONDED, k, 0, 0, 0});
        blk += nb;
    }
    c->pair_blocks = blk;  // blocks of force_kernel: pair lists + bonded kinds
    c->n_pair_segs = (int)c->segs.size();
    if (c->da.nterms > 0 && c->da.capacity > 0) {
        int nb = (c->da.capacity + kBlock - 1) / kBlock;
        c->segs.push_back({blk, nb, SEG_DA, 0, 0, 0, 0});
        blk += nb;
    }
    c->total_blocks = blk;
    c->local_blocks = blk - c->pair_blocks;
    c->n_local_segs = (int)c->segs.size() - c->n_pair_segs;
    c->smem_bytes = smem;
    if (c->segs.empty()) c->segs.push_back({0, 0, SEG_BONDED, 0, 0, 0, 0});
    c->d_segs.upload(c->segs);
    c->tile_blocks = tblk;
    if (smem > 48 * 1024) {
        CUDA_TRY(cudaFuncSetAttribute(force_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(cudaFuncSetAttribute(force_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }

    drop_graphs(c);
}

static void compile_bonded(ocm_context* c) {
    std::vector<double> tabs(24, 0.0);  // table 0 = dummy
    for (int k = 0; k < B_NUM_KINDS; k++) {
        std::vector<int> idx, meta;
        std::vector<double> par;
        int npar = 0, nbody = bonded_nbody(k);
        for (const auto& t : c->sys.bonded) {
            if (t.kind != k) continue;
            npar = t.npar;
            int tab_id = 0;
            if (k == B_FLEX_ANGLE || k == B_FLEX_DIHEDRAL) {
                tab_id = (int)(tabs.size() / 24);
                if (tab_id >= 32768) throw std::runtime_error("too many flexible-local tables");
                tabs.insert(tabs.end(), t.tab, t.tab + 24);
            }
            idx.insert(idx.end(), t.idx.begin(), t.idx.end());
            par.insert(par.end(), t.p.begin(), t.p.end());
            for (int64_t m = 0; m < t.n(); m++) meta.push_back(t.group | (t.periodic << 8) | (tab_id << 16));
        }
        BondedDev& B = c->bonded_h[k];
        B.n = (int)(idx.size() / nbody); B.npar = npar;
        B.dd_terms = nullptr; B.dd_count = nullptr;
        if (B.n) {
            c->b_idx[k].upload(idx); c->b_par[k].upload(par); c->b_meta[k].upload(meta);
            B.idx = c->b_idx[k].p; B.par = c->b_par[k].p; B.meta = c->b_meta[k].p;
            if (c->dd) {
                c->dd->bterms[k].alloc(B.n);
                B.dd_terms = c->dd->bterms[k].p; B.dd_count = c->dd->counts.p + 1 + k;
            }
        }
    }
    c->b_tabs.upload(tabs);
    for (int k = 0; k < B_NUM_KINDS; k++) c->bonded_h[k].tabs = c->b_tabs.p;
    c->d_bonded.upload(std::vector<BondedDev>(c->bonded_h, c->bonded_h + B_NUM_KINDS));
}

static void compile_coms(ocm_context* c) {
    c->n_coms = (int)c->sys.coms.size();
    if (!c->n_coms) return;
    std::vector<int> idx;
    std::vector<double> w;
    std::vector<ComDev> terms;
    for (const auto& t : c->sys.coms) {
        ComDev D{};
        D.group = t.group; D.periodic = t.periodic; D.n1 = (int)t.g1.size(); D.n2 = (int)t.g2.size(); D.begin = (int)idx.size();
        D.k = t.k; D.v0 = t.v0;
        for (const auto* grp : {&t.g1, &t.g2}) {
            double M = 0;
            for (int32_t i : *grp) M += c->sys.mass[i];
            for (int32_t i : *grp) { idx.push_back(i); w.push_back(c->sys.mass[i] / M); }
        }
        terms.push_back(D);
    }
    c->com_idx.upload(idx); c->com_w.upload(w);
    for (auto& D : terms) { D.idx = c->com_idx.p + D.begin; D.w = c->com_w.p + D.begin; }
    c->d_coms.upload(terms);
}

// ---- whole-step persistent kernel: static tables of one replica ------------------------------------------
template <typename T>
static const T* rep_upload(ocm_context* c, const std::vector<T>& v) {
    if (v.empty()) return nullptr;
    c->rep_blobs.emplace_back(new DevBuf<unsigned char>());
    c->rep_blobs.back()->alloc(v.size() * sizeof(T));
    CUDA_TRY(cudaMemcpy(c->rep_blobs.back()->p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return reinterpret_cast<const T*>(c->rep_blobs.back()->p);
}

static void compile_replica_tables(ocm_context* c) {
    c->rep_ok = false;
    c->rep_blobs.clear();
    const ocm_system& S1 = c->nrep > 1 ? c->sys1 : c->sys;
    const int nper = c->nper;
    if (c->dd || (c->sys.has_box && !c->vbox) || !S1.das.empty() || nper > 256 || nper < 2) return;
    if ((int)c->lists.size() > kRepMaxLists) return;
    for (const auto& t : S1.bonded) if (t.kind >= B_3SPN2_STACK && t.n() > 0) return;   // (base stacking, external terms)
    if (!S1.coms.empty()) return;
    RepParams& R = c->rep;
    R = RepParams{};
    R.nper = nper; R.nrep = c->nrep; R.nlists = 0;
    std::vector<RepTask> tasks;
    for (auto& Lp : c->lists) if (Lp->form == SEG_PAIR_WELL) return;
    for (auto& Lp : c->lists) {
        PairListHost& L = *Lp;
        if (L.np == 0) continue;
        const int nm = (int)L.terms.size();
        const PairTerm& t0 = S1.pairs[L.terms[0]];
        const PairTerm* rep = nullptr; const PairTerm* att = nullptr;
        if (nm == 2) { rep = &S1.pairs[L.terms[0]]; att = &S1.pairs[L.terms[1]]; if (rep->kind != P_LJREP) std::swap(rep, att); }
        // full table: coefficients of every interacting pair from both sides, zero elsewhere
        std::vector<float4> coef((size_t)nper * nper, make_float4(0.f, 0.f, 0.f, 0.f));
        size_t e = 0;  // exclusions are sorted (i < j): walk them alongside the pair loop
        for (size_t a = 0; a < t0.part.size(); a++)
            for (size_t b = a + 1; b < t0.part.size(); b++) {
                const int i = t0.part[a], j = t0.part[b];
                while (e < t0.excl.size() && (t0.excl[e].first < i || (t0.excl[e].first == i && t0.excl[e].second < j))) e++;
                if (e < t0.excl.size() && t0.excl[e].first == i && t0.excl[e].second == j) continue;
                if (t0.ncls > 0 && !((t0.allow[t0.cls[i]] >> t0.cls[j]) & 1u)) continue;
                float4 cf;
                if (L.form == SEG_PAIR_DH) cf = make_float4((float)(t0.g0 * t0.p0[i] * t0.p0[j]), 0.f, 0.f, 0.f);
                else if (nm == 2) {
                    const TZ zr = tz_coeffs(*rep, i, j), za = tz_coeffs(*att, i, j);
                    cf = make_float4((float)(12.0 * zr.c12i), (float)(6.0 * zr.c6i), (float)zr.a2, (float)(za.c12o / zr.c12i));
                } else {
                    const TZ z = tz_coeffs(t0, i, j);
                    if (t0.kind == P_LJATT) cf = make_float4((float)(12.0 * z.c12o), (float)(6.0 * z.c6o), (float)z.a2, 1.f);
                    else cf = make_float4((float)(12.0 * z.c12i), (float)(6.0 * z.c6i), std::isfinite(z.a2) ? (float)z.a2 : INFINITY, 0.f);
                }
                coef[(size_t)i * nper + j] = cf; coef[(size_t)j * nper + i] = cf;
            }
        const int l = R.nlists++;
        RepList& Q = R.lists[l];
        Q.form = L.form; Q.has_switch = L.dev.has_switch;
        Q.rc2 = L.dev.rc2; Q.rl2 = (float)std::pow(std::sqrt((double)L.dev.rc2) + (double)kRepSkin, 2); Q.k_in = L.dev.k_in; Q.rsw = L.dev.rsw; Q.inv_sw = L.dev.inv_sw; Q.inv_debye = L.dev.inv_debye;
        Q.coef = rep_upload(c, coef);
        for (int i0 = 0; i0 < nper; i0 += 8) tasks.push_back({kRepTaskPair + l, i0});
    }
    // bonded terms of one replica, kind by kind: local bead indices, FP32 parameters, own table numbering
    std::vector<float> tabs(24, 0.f);
    std::vector<RepTerm> rec[9];
    std::vector<int> nslot_of_bead(nper, 0);
    for (int k = 0; k < 9; k++) {
        int npar = 0;
        const int nbody = bonded_nbody(k);
        for (const auto& t : S1.bonded) {
            if (t.kind != k || t.n() == 0) continue;
            if (t.periodic) return;
            if ((npar && npar != t.npar) || t.npar > 5) return;
            npar = t.npar;
            int tab_id = 0;
            if (k == B_FLEX_ANGLE || k == B_FLEX_DIHEDRAL) {
                tab_id = (int)(tabs.size() / 24);
                if (tab_id >= 65536) return;
                for (int q = 0; q < 24; q++) tabs.push_back((float)t.tab[q]);
            }
            for (int64_t m = 0; m < t.n(); m++) {
                RepTerm r{};
                for (int b = 0; b < nbody; b++) { r.idx[b] = (unsigned short)t.idx[(size_t)m * nbody + b]; nslot_of_bead[r.idx[b]]++; }
                r.tab = (unsigned short)tab_id;
                for (int q = 0; q < t.npar; q++) r.p[q] = (float)t.p[(size_t)m * t.npar + q];
                if (k == B_HARMONIC_BOND) r.r0 = t.p[(size_t)m * t.npar + 0];
                else if (k == B_GAUSSIAN_BOND) r.r0 = t.p[(size_t)m * t.npar + 1];
                else if (k == B_3SPN2_BOND) r.r0 = t.p[(size_t)m * t.npar + 2];
                rec[k].push_back(r);
            }
        }
        R.bonded[k] = RepBonded{};
        R.bonded[k].n = (int)rec[k].size(); R.bonded[k].npar = npar;
        for (int m0 = 0; m0 < (int)rec[k].size(); m0 += 32) tasks.push_back({k, m0});
    }
    // scratch slots: contribution number p to bead i (in (kind, term, body) order) lives in slot p * nper + i
    const int nslots = nper * ((std::max(1, *std::max_element(nslot_of_bead.begin(), nslot_of_bead.end())) + 7) & ~7);
    if (nslots >= 65536) return;
    {
        std::vector<int> fill(nper, 0);
        for (int k = 0; k < 9; k++) {
            const int nbody = bonded_nbody(k);
            for (auto& r : rec[k])
                for (int b = 0; b < nbody; b++) r.slot[b] = (unsigned short)(fill[r.idx[b]]++ * nper + r.idx[b]);
            if (!rec[k].empty()) R.bonded[k].rec = reinterpret_cast<const uint4*>(rep_upload(c, rec[k]));
        }
    }
    for (int i0 = 0; i0 < nper; i0 += 32) tasks.push_back({kRepTaskNoise, i0});
    // Threads per block: a block is one replica.  An ensemble that fills the GPU runs 256-thread blocks (several per SM);
    // a few replicas get up to 1024 threads each, so that every warp has one task and a step is as short as its longest task.
    c->rep_block = c->nrep > 296 ? 256 : (c->nrep > 148 ? 512 : 1024);
    while (c->rep_block > 256 && c->rep_block / 32 >= 2 * (int)tasks.size()) c->rep_block /= 2;
    if (c->rep_block_opt == 256 || c->rep_block_opt == 512 || c->rep_block_opt == 1024) c->rep_block = c->rep_block_opt;
    // warp w runs tasks w, w + nwarps, ... of this list: longest first (estimated warp instructions per task, from the SASS of the
    // kinds: FLD 380, gaussian / cosine dihedral 340, FLA 260, a walk over a bead group's Verlet lists 250, harmonic angle 200,
    // noise 170, bonds 140-160, 12-10 contact 110)
    auto cost = [](int kind) {
        static const int bonded_cost[9] = {140, 160, 110, 150, 200, 260, 340, 340, 380};
        return kind >= kRepTaskNoise ? 170 : (kind >= kRepTaskPair ? 250 : bonded_cost[kind]);
    };
    std::stable_sort(tasks.begin(), tasks.end(), [&](const RepTask& a, const RepTask& b) { return cost(a.kind) > cost(b.kind); });
    R.nslots = nslots; R.ntasks = (int)tasks.size();
    R.tabs = rep_upload(c, tabs); R.tasks = rep_upload(c, tasks);
    // (phase 2 sums a bead's slots eight at a time: counts rounded up, the padding slots hold zeros)
    for (int& v : nslot_of_bead) v = (v + 7) & ~7;
    R.nslot_of_bead = rep_upload(c, nslot_of_bead);
    R.rep_seed = c->nrep > 1 ? c->d_rep_seed.p : nullptr;
    c->rep_smem = rep_smem_bytes(nper, R.nlists, nslots);
    if (c->rep_smem > 96 * 1024) return;   // (scratch slots of a few thousand terms: beyond that the general path)
    auto* kern = c->rep_block == 256 ? replica_step_kernel<256> : (c->rep_block == 512 ? replica_step_kernel<512> : replica_step_kernel<1024>);
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(c->rep_smem, 1024)));
    // Shared memory of the blocks that can be resident on one SM, and no more: what is left of the 256 KB stays L1, where the
    // read-only tables of the kernel live (coefficient table n x n x 16 B, term records 48 B each).  With the largest carve-out the
    // L1 is 28 KB and those loads come from L2.
    {
        const int per_sm_limit = c->rep_block == 256 ? 5 : (c->rep_block == 512 ? 2 : 1);
        const int resident = std::max(1, std::min(per_sm_limit, (c->nrep + 147) / 148));
        cudaFuncAttributes fa{};
        CUDA_TRY(cudaFuncGetAttributes(&fa, kern));
        // (per block: dynamic + static shared memory + 1 KB the system reserves)
        int pct = (int)std::ceil(100.0 * (double)resident * (double)(c->rep_smem + fa.sharedSizeBytes + 1024) / (228.0 * 1024.0)) + 1;
        pct = std::max(0, std::min(100, pct));
        if (c->rep_carveout_opt >= 0) pct = std::min(100, c->rep_carveout_opt);
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    }
    c->rep_ok = true;
}

static void compile_da(ocm_context* c) {
    long long cand = 0;
    for (const auto& t : c->sys.das) {
        DATermDev D{};
        D.kind = t.kind; D.group = t.group; D.periodic = t.periodic; D.npar = t.npar;
        D.nd = (int)t.nd(); D.na = (int)t.na(); D.nexcl = (int)t.excl.size();
        D.cand_begin = cand;
        cand += (long long)D.nd * D.na;
        D.rc2 = t.cutoff * t.cutoff;
        D.g0 = t.g0; D.g1 = t.g1; D.g2 = t.g2;
        auto bi = std::make_unique<DevBuf<int>>();
        std::vector<int> idx(t.donors.begin(), t.donors.end());
        idx.insert(idx.end(), t.acceptors.begin(), t.acceptors.end());
        if (idx.empty()) idx.push_back(0);
        bi->upload(idx);
        D.donors = bi->p; D.acceptors = bi->p + t.donors.size();
        auto bp = std::make_unique<DevBuf<double>>();
        bp->upload(t.p.empty() ? std::vector<double>(1, 0.0) : t.p);
        D.par = bp->p;
        auto be = std::make_unique<DevBuf<unsigned long long>>();
        std::vector<unsigned long long> ex(t.excl.begin(), t.excl.end());
        if (ex.empty()) ex.push_back(~0ull);
        be->upload(ex);
        D.excl = be->p;
        c->da_idx.push_back(std::move(bi)); c->da_par.push_back(std::move(bp)); c->da_excl.push_back(std::move(be));
        c->da_terms_h.push_back(D);
        c->da_rcmax = std::max(c->da_rcmax, t.cutoff);
    }
    if (c->da_terms_h.size() > 255) throw std::runtime_error("more than 255 donor-acceptor terms");
    c->da_count.alloc(1); c->da_count.zero();
    c->da.nterms = (int)c->da_terms_h.size();
    c->da.ncand = cand;
    c->da.capacity = 0;
    c->da.count = c->da_count.p;
    c->da.entries = nullptr;
    c->da.terms = nullptr;
}

// (re)derive what depends on the skin and upload the term table
static void upload_da(ocm_context* c) {
    if (c->da_terms_h.empty()) return;
    for (size_t k = 0; k < c->da_terms_h.size(); k++) {
        const double rl = c->sys.das[k].cutoff + c->skin;
        c->da_terms_h[k].rlist2 = rl * rl;
    }
    c->d_da_terms.upload(c->da_terms_h);
    c->da.terms = c->d_da_terms.p;
}

static void enqueue_da_build(ocm_context* c, cudaStream_t s) {
    if (c->da.nterms == 0 || c->da.ncand == 0) return;
    CUDA_TRY(cudaMemsetAsync(c->da_count.p, 0, sizeof(unsigned int), s));
    const long long nb = (c->da.ncand + 255) / 256;
    da_build_kernel<<<(unsigned int)nb, 256, 0, s>>>(c->da, c->P, c->x.p, c->flags.p);
    c->n_launches += 1;
}

static void update_params(ocm_context* c);
// lay the virtual box of an unlimited-boundary system out around the given positions: returns true if it changed
static bool vbox_layout(ocm_context* c, const double* xyz, bool force) {
    const int n = c->n;
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int i = 0; i < n; i++)
        for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], xyz[3 * i + k]); hi[k] = std::max(hi[k], xyz[3 * i + k]); }
    for (int k = 0; k < 3; k++) if (!(lo[k] <= hi[k]) || !std::isfinite(lo[k]) || !std::isfinite(hi[k])) throw std::runtime_error("positions are not finite");
    bool inside = !force;
    for (int k = 0; k < 3 && inside; k++) inside = lo[k] >= c->vb_lo[k] && hi[k] <= c->vb_hi[k];
    if (inside) return false;
    double R = c->da_rcmax;
    for (auto& L : c->lists) R = std::max(R, L->rc_list);
    R += c->skin;
    R = std::max(R, 0.5);
    for (int k = 0; k < 3; k++) {
        const double ext = hi[k] - lo[k];
        const double grow = std::max(R, 0.15 * ext) * (1.0 + 0.5 * std::min(c->n_vbox_resize, 8));  // room to move before a re-layout
        c->vb_lo[k] = lo[k] - grow; c->vb_hi[k] = hi[k] + grow;
        // images stay out of reach: L > (extent the beads may reach) + R; and at least 4 R for the cell grid
        c->sys.box[k] = std::max(4.0 * R, (ext + 2.0 * grow + R) * 1.01);
    }
    c->n_vbox_resize++;
    update_params(c);
    c->lists_ready = false; c->lists_current = false;
    for (auto& L : c->lists) { L->stride = 0; L->row_len32 = 0; }
    drop_graphs(c);
    return true;
}

static void update_params(ocm_context* c) {
    StepParams& P = c->P;
    P.n = c->n; P.periodic = c->periodic ? 1 : 0;
    for (int k = 0; k < 3; k++) {
        P.box[k] = c->periodic ? c->sys.box[k] : 1.0;
        P.inv_box[k] = 1.0 / P.box[k];
        P.fscale[k] = (float)(P.box[k] / 4294967296.0);
    }
    P.dt = c->dt;
    const double a = std::exp(-c->gamma * c->dt);
    P.a = (float)a;
    P.b = (float)(c->gamma == 0 ? c->dt : (1.0 - a) / c->gamma);
    P.c = (float)std::sqrt(kBoltz * c->T * (1.0 - a * a));
    P.seed_lo = (unsigned int)c->seed; P.seed_hi = (unsigned int)(c->seed >> 32);
    P.half_skin2 = (float)(0.25 * c->skin * c->skin);
    P.nrep = c->nrep; P.nper = c->nper;
    P.origin = c->nrep > 1 ? c->origin.p : nullptr;
    P.rep_seed = c->nrep > 1 ? c->d_rep_seed.p : nullptr;
    P.role = c->dd ? c->dd->role.p : nullptr;
    P.present = c->dd ? c->dd->present.p : nullptr;
    P.n_present = c->dd ? c->dd->counts.p : nullptr;
    if (c->dd) dd_configure(c);
}

// ---- enqueue helpers ----------------------------------------------------------------------------
static inline int nblk(int n, int b) { return std::max(1, (n + b - 1) / b); }

// ---- slab decomposition: host side ---------------------------------------------------------------
// (re)derive what depends on the skin: the halo width = largest list radius
static void dd_configure(ocm_context* c) {
    DDState& d = *c->dd;
    double h = 0;
    for (auto& L : c->lists) if (L->np) h = std::max(h, L->rc_list + c->skin);
    h = std::max(h, 4.0 * c->skin);
    const double len = c->sys.box[d.axis];
    if (len / d.world < h + c->skin)
        throw std::runtime_error("slab decomposition: a slab (" + std::to_string(len / d.world) + " nm) must be wider than the halo (" +
                                 std::to_string(h) + " nm) plus the skin; use fewer ranks or another axis");
    d.halo = h;
    const double two32 = 4294967296.0;
    auto face = [&](int r) { return (unsigned int)(unsigned long long)std::ceil((double)r * two32 / d.world); };
    d.G.world = d.world; d.G.rank = d.rank; d.G.axis = d.axis;
    d.G.lo = face(d.rank);
    d.G.hi = d.rank + 1 == d.world ? 0u : face(d.rank + 1);
    d.G.halo_fix = (unsigned int)std::ceil(h / len * two32) + 2u;
}

// cuStreamWaitValue32 through the runtime's driver entry point (the engine links only the CUDA runtime)
typedef int (*StreamWaitValue32Fn)(cudaStream_t, unsigned long long, unsigned int, unsigned int);
static StreamWaitValue32Fn stream_wait_value32() {
    static StreamWaitValue32Fn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { cudaGetLastError(); f = nullptr; }
        return (StreamWaitValue32Fn)f;
    }();
    return fn;
}

static void dd_alloc(ocm_context* c, int cap) {
    DDState& d = *c->dd;
    cap = (cap + 255) / 256 * 256;
    // neighbours still map the old push buffer: unmap everywhere before anybody frees (collective)
    if (d.push_ok) {
        d.comm->peer_unmap(d.rank);
        d.comm->allreduce(d.rank, c->flags.p + FLAG_COUNT, 1, DD_I32, true, c->stream);
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        d.push_ok = false;
    }
    d.cap = cap;
    for (int k = 0; k < 2; k++) {
        d.a_send[k].alloc(d.bytes_a()); d.a_recv[k].alloc(d.bytes_a());
        d.b_send[k].alloc(d.bytes_b()); d.b_recv[k].alloc(d.bytes_b());
        d.s_send[k].alloc((size_t)3 * cap); d.s_recv[k].alloc((size_t)3 * cap);
        d.a_send[k].zero(); d.a_recv[k].zero(); d.b_send[k].zero(); d.b_recv[k].zero(); d.s_send[k].zero(); d.s_recv[k].zero();
    }
    d.push.alloc(d.bytes_push());
    d.push.zero();
    if (!d.ticket.p) { d.ticket.alloc(1); d.ticket.zero(); }
    CUDA_TRY(cudaDeviceSynchronize());
    if (d.push_wanted && stream_wait_value32()) {
        void *lo = nullptr, *hi = nullptr;
        d.push_ok = d.comm->peer_map(d.rank, d.push.p, &lo, &hi, c->stream);
        d.peer_lo = (char*)lo; d.peer_hi = (char*)hi;
    }
}

// rebuild-time exchange: migration + halo state (phase A), per-step halo lists (phase B)
static void dd_exchange_rebuild(ocm_context* c, cudaStream_t s) {
    DDState& d = *c->dd;
    const int n = c->n, cap = d.cap;
    const DDMsg slo = DDState::msg(d.a_send[0].p, cap), shi = DDState::msg(d.a_send[1].p, cap);
    const DDMsg rlo = DDState::msg(d.a_recv[0].p, cap), rhi = DDState::msg(d.a_recv[1].p, cap);
    CUDA_TRY(cudaMemsetAsync(d.a_send[0].p, 0, 16, s));
    CUDA_TRY(cudaMemsetAsync(d.a_send[1].p, 0, 16, s));
    dd_select_kernel<<<nblk(n, 256), 256, 0, s>>>(n, d.G, cap, d.role.p, c->posq.p, c->x.p, c->v.p, slo, shi, c->flags.p);
    d.comm->ring_exchange(d.rank, d.a_send[0].p, d.a_send[1].p, d.a_recv[0].p, d.a_recv[1].p, d.bytes_a(), s);
    dd_install_kernel<<<nblk(cap, 256), 256, 0, s>>>(d.G, c->P, cap, rlo, c->x.p, c->v.p, c->posq.p, d.known.p);
    dd_install_kernel<<<nblk(cap, 256), 256, 0, s>>>(d.G, c->P, cap, rhi, c->x.p, c->v.p, c->posq.p, d.known.p);
    CUDA_TRY(cudaMemsetAsync(d.counts.p, 0, (1 + B_NUM_KINDS) * sizeof(int), s));
    dd_assign_kernel<<<nblk(n, 256), 256, 0, s>>>(n, d.G, c->posq.p, d.role.p, d.known.p, d.present.p, d.counts.p);
    CUDA_TRY(cudaMemsetAsync(d.b_send[0].p, 0, 16, s));
    CUDA_TRY(cudaMemsetAsync(d.b_send[1].p, 0, 16, s));
    dd_steplist_kernel<<<nblk(n, 256), 256, 0, s>>>(n, d.G, cap, d.role.p, c->posq.p, reinterpret_cast<int*>(d.b_send[0].p),
                                                    reinterpret_cast<int*>(d.b_send[0].p + 16), reinterpret_cast<int*>(d.b_send[1].p),
                                                    reinterpret_cast<int*>(d.b_send[1].p + 16), c->flags.p);
    d.comm->ring_exchange(d.rank, d.b_send[0].p, d.b_send[1].p, d.b_recv[0].p, d.b_recv[1].p, d.bytes_b(), s);
    for (int k = 0; k < B_NUM_KINDS; k++)
        if (c->bonded_h[k].n > 0)
            dd_check_bonded_kernel<<<nblk(c->bonded_h[k].n, 256), 256, 0, s>>>(c->d_bonded.p, k, d.role.p, c->flags.p, d.bterms[k].p, d.counts.p + 1 + k,
                                                                                 nblk(d.bonded_cover(k, c->bonded_h[k].n), kBlock) * kBlock);
    c->n_launches += 5; d.n_exchanges += 2;
}

// per MD step: ship the positions of the halo-list beads to both neighbours
static void dd_halo_step(ocm_context* c, cudaStream_t s) {
    DDState& d = *c->dd;
    const int cap = d.cap;
    auto cnt = [](DevBuf<char>& b) { return reinterpret_cast<int*>(b.p); };
    auto idx = [](DevBuf<char>& b) { return reinterpret_cast<int*>(b.p + 16); };
    if (d.push_ok) {
        // peer-memory path: my low-side list lands in the previous rank's "from the high side" buffer and vice versa;
        // two parities so that a neighbour one step ahead never overwrites what is still being unpacked here
        const unsigned int seq = ++d.seq;
        const int par = (int)(seq & 1u);
        dd_push_kernel<<<nblk(cap, 256), 256, 0, s>>>(cnt(d.b_send[0]), idx(d.b_send[0]), reinterpret_cast<double*>(d.peer_lo + d.push_data_off(1, par)),
                                                      cnt(d.b_send[1]), idx(d.b_send[1]), reinterpret_cast<double*>(d.peer_hi + d.push_data_off(0, par)), cap, c->x.p,
                                                      reinterpret_cast<unsigned int*>(d.peer_lo + d.push_flag_off(1, par)),
                                                      reinterpret_cast<unsigned int*>(d.peer_hi + d.push_flag_off(0, par)), seq, d.ticket.p);
        for (int side = 0; side < 2; side++)
            if (stream_wait_value32()(s, (unsigned long long)(uintptr_t)(d.push.p + d.push_flag_off(side, par)), seq, 0u /* CU_STREAM_WAIT_VALUE_GEQ */) != 0)
                throw std::runtime_error("slab decomposition: cuStreamWaitValue32 failed");
        dd_unpack_kernel<<<nblk(cap, 256), 256, 0, s>>>(cnt(d.b_recv[0]), idx(d.b_recv[0]), reinterpret_cast<const double*>(d.push.p + d.push_data_off(0, par)),
                                                        cnt(d.b_recv[1]), idx(d.b_recv[1]), reinterpret_cast<const double*>(d.push.p + d.push_data_off(1, par)), cap, c->P,
                                                        c->x.p, c->posq.p, c->d_lists.p, (int)c->lists.size());
        c->n_launches += 2; d.n_exchanges += 1;
        return;
    }
    dd_pack_kernel<<<nblk(cap, 256), 256, 0, s>>>(cnt(d.b_send[0]), idx(d.b_send[0]), d.s_send[0].p, cnt(d.b_send[1]), idx(d.b_send[1]),
                                                  d.s_send[1].p, cap, c->x.p);
    d.comm->ring_exchange(d.rank, d.s_send[0].p, d.s_send[1].p, d.s_recv[0].p, d.s_recv[1].p, d.bytes_s(), s);
    dd_unpack_kernel<<<nblk(cap, 256), 256, 0, s>>>(cnt(d.b_recv[0]), idx(d.b_recv[0]), d.s_recv[0].p, cnt(d.b_recv[1]), idx(d.b_recv[1]),
                                                    d.s_recv[1].p, cap, c->P, c->x.p, c->posq.p, c->d_lists.p, (int)c->lists.size());
    c->n_launches += 2; d.n_exchanges += 1;
}

// flags as every rank must see them (max over ranks under a slab decomposition), on the host
static void read_flags(ocm_context* c, int* fl) {
    if (c->dd) c->dd->comm->allreduce(c->dd->rank, c->flags.p, FLAG_COUNT, DD_I32, true, c->stream);
    CUDA_TRY(cudaMemcpyAsync(fl, c->flags.p, FLAG_COUNT * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    // (an overflowed halo message leaves beads missing: the reach check is only meaningful without overflow)
    if (fl[FLAG_DD_REACH] && !fl[FLAG_DD_OVERFLOW]) throw std::runtime_error("slab decomposition: a bonded term reaches beyond the halo");
}

// the list-build launch: rows are staged in shared memory (8 warps x stride x 5 bytes) to be written out in zone order
typedef void (*BuildListKernel)(const PairListDev*, int, StepParams, int*, int, int);
static BuildListKernel build_list_variant(bool per, bool dh, bool cls) {
    static const BuildListKernel tab[8] = {
        build_list_kernel<false, false, false>, build_list_kernel<false, false, true>, build_list_kernel<false, true, false>,
        build_list_kernel<false, true, true>,   build_list_kernel<true, false, false>, build_list_kernel<true, false, true>,
        build_list_kernel<true, true, false>,   build_list_kernel<true, true, true>};
    return tab[(per ? 4 : 0) + (dh ? 2 : 0) + (cls ? 1 : 0)];
}

static void launch_build(ocm_context* c, PairListHost& L, int li, cudaStream_t s, bool count_only) {
    const size_t bytes = (size_t)8 * L.stride * 5 + 16;
    const int zoned = (!count_only && L.stride > 0 && bytes <= 160 * 1024) ? 1 : 0;
    const size_t smem = zoned ? bytes : 0;
    // one warp per row; under a slab decomposition only the rows the force launch covers (more rows than that: overflow flag)
    const int rows = (c->dd && !count_only && L.dev.row_capacity > 0) ? L.dev.row_capacity : L.np;
    auto* kern = build_list_variant(c->periodic, L.form == SEG_PAIR_DH, L.dev.cls != nullptr);
    kern<<<nblk(rows * 32, 256), 256, smem, s>>>(c->d_lists.p, li, c->P, c->flags.p, count_only ? 1 : 0, zoned);
}

// sort the beads of one list into cells and build its pair list (Verlet rows or cluster-pair tiles)
static void enqueue_list_rebuild(ocm_context* c, PairListHost& L, int li, cudaStream_t s, bool count_only) {
    if (c->periodic) {
        bin_kernel<true><<<nblk(L.np, 256), 256, 0, s>>>(c->d_lists.p, li, c->posq.p);
    } else {
        bbox_kernel<<<nblk(L.np, 256), 256, 0, s>>>(c->d_lists.p, li, c->posq.p);
        grid_kernel<<<1, 1, 0, s>>>(c->d_lists.p, li);
        bin_kernel<false><<<nblk(L.np, 256), 256, 0, s>>>(c->d_lists.p, li, c->posq.p);
    }
    if (c->periodic && c->nrep == 1 && L.max_cells > 4 * kScanChunk) {
        // large grid (its size is known on the host for a periodic or virtual box): block-parallel scan
        const int nb = (L.max_cells + kScanChunk - 1) / kScanChunk;
        scan_blocks_kernel<<<nb, 1024, 0, s>>>(c->d_lists.p, li);
        scan_tops_kernel<<<1, 1024, 0, s>>>(c->d_lists.p, li, nb);
        scan_finish_kernel<<<nb, 1024, 0, s>>>(c->d_lists.p, li);
        c->n_launches += 2;
    } else scan_kernel<<<1, 1024, 0, s>>>(c->d_lists.p, li);
    scatter_kernel<<<nblk(L.np, 256), 256, 0, s>>>(c->d_lists.p, li);
    if (L.tiled) {
        // (the grids cover the rows the force launch covers: all of them, or the estimate of a slab decomposition)
        const int rows = (c->dd && !count_only && L.dev.row_capacity > 0) ? L.dev.row_capacity : L.np_pad;
        const int nsc = (rows + 63) / 64;
        tile_fill_kernel<<<nblk(L.np_pad, 128), 128, 0, s>>>(c->d_lists.p, li, c->posq.p);
        tile_bbox_kernel<<<nblk(nsc * 32, 128), 128, 0, s>>>(c->d_lists.p, li, c->P);
        CUDA_TRY(cudaMemsetAsync(L.tl_mask_count.p, 0, sizeof(int), s));
        tile_build_kernel<true><<<nsc, kBuildBlock, 0, s>>>(c->d_lists.p, li, c->P, c->flags.p, count_only ? 1 : 0);
        if (L.dev.n_excl_pairs > 0)
            tile_excl_kernel<<<nblk(L.dev.n_excl_pairs, 256), 256, 0, s>>>(c->d_lists.p, li, c->flags.p, count_only ? 1 : 0);
        if (!count_only) {
            tile_scan_kernel<<<1, 1024, 0, s>>>(c->d_lists.p, li, nsc, 0);
            if (L.dev.halo_split) tile_scan_kernel<<<1, 1024, 0, s>>>(c->d_lists.p, li, nsc, 1);
        }
        c->n_launches += 9;
        return;
    }
    rank_fill_kernel<<<nblk(L.np, 128), 128, 0, s>>>(c->d_lists.p, li, c->posq.p);
    launch_build(c, L, li, s, count_only);
    c->n_launches += c->periodic ? 5 : 7;
}

static void enqueue_rebuild(ocm_context* c, cudaStream_t s, bool count_only) {
    const int n = c->n;
    if (c->nrep > 1) {
        // replica batch: re-centre every replica's local frame, then refresh the FP32 offsets
        recentre_kernel<<<nblk(c->nrep * 32, 128), 128, 0, s>>>(c->P, c->x.p, c->origin.p);
        canon_kernel<<<nblk(n, 256), 256, 0, s>>>(c->P, c->x.p, c->posq.p);
        c->n_launches += 2;
    }
    if (c->dd && !count_only) dd_exchange_rebuild(c, s);
    if (!count_only) save_ref_kernel<<<nblk(3 * n, 256), 256, 0, s>>>(n, c->x.p, c->xref.p);
    if (c->vbox && !count_only)
        vbox_check_kernel<<<nblk(n, 256), 256, 0, s>>>(n, c->x.p, c->vb_lo[0], c->vb_lo[1], c->vb_lo[2], c->vb_hi[0], c->vb_hi[1], c->vb_hi[2], c->flags.p);
    // A periodic single-domain system with several pair lists: the lists after the first are rebuilt on the side stream beside
    // it (each list has its own cell arrays and rows; they only share the read-only positions and the flag words).  C2: the
    // Coulomb list's five short launches (37 us) disappear behind the Ashbaugh-Hatch row build.
    const bool fork = !count_only && c->periodic && !c->dd && c->nrep == 1 && c->lists.size() > 1 && c->side != nullptr;
    if (fork) { CUDA_TRY(cudaEventRecord(c->ev_rb_fork, s)); CUDA_TRY(cudaStreamWaitEvent(c->side, c->ev_rb_fork, 0)); }
    for (size_t l = 0; l < c->lists.size(); l++) {
        auto& L = *c->lists[l];
        if (L.np == 0) continue;
        enqueue_list_rebuild(c, L, (int)l, (fork && l > 0) ? c->side : s, count_only);
    }
    if (fork) { CUDA_TRY(cudaEventRecord(c->ev_rb_join, c->side)); CUDA_TRY(cudaStreamWaitEvent(s, c->ev_rb_join, 0)); }
    if (!count_only) {
        enqueue_da_build(c, s);
        c->n_launches += 1; c->n_rebuilds += 1;
    }
}

// launch with the programmatic-stream-serialization attribute (option "pdl"): the grid may be scheduled while the
// previous kernel of the stream drains; force_kernel and integrate_kernel wait (griddepcontrol.wait) before they read
template <class... KArgs, class... Args>
static void launch_chained(ocm_context* c, void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = c->pdl ? 1 : 0;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...));
}

static void launch_tile(ocm_context* c, PairListHost& L, const TileLaunch& T, int nblocks, cudaStream_t s) {
    // one launch per tile list; the pair form is a template parameter, the list's constants ride in the kernel parameters
    if (L.form == SEG_PAIR_DH) launch_chained(c, tile_force_kernel<SEG_PAIR_DH, false, false>, nblocks, kTileBlock, 0, s, T, c->fbuf.p);
    else if (L.t_switch && !(c->tile_debug & 4)) launch_chained(c, tile_force_kernel<SEG_PAIR_TZ, true, false>, nblocks, kTileBlock, 0, s, T, c->fbuf.p);
    else if (L.dev.t_geo) launch_chained(c, tile_force_kernel<SEG_PAIR_TZ, false, true>, nblocks, kTileBlock, 0, s, T, c->fbuf.p);
    else launch_chained(c, tile_force_kernel<SEG_PAIR_TZ, false, false>, nblocks, kTileBlock, 0, s, T, c->fbuf.p);
    c->n_launches += 1;
}

// phase 0: all forces.  Slab decomposition with split tile lists: phase 1 = the interior tile entries only (owned beads on
// both sides: they can run while the step's halo positions are still travelling), phase 2 = everything else.
static void enqueue_forces(ocm_context* c, cudaStream_t s, int phase = 0) {
    if (c->total_blocks == 0 && c->tile_blocks == 0 && c->n_coms == 0) return;
    if (phase == 1) {
        for (auto& Lp : c->lists)
            if (Lp->tiled && Lp->np && Lp->tile_nblocks_in > 0) launch_tile(c, *Lp, Lp->launch_in, Lp->tile_nblocks_in, s);
        return;
    }
    const bool fork = c->local_blocks > 0 && c->pair_blocks > 0;
    if (c->local_blocks > 0) {
        cudaStream_t sl = s;
        if (fork) {
            CUDA_TRY(cudaEventRecord(c->ev_fork, s));
            CUDA_TRY(cudaStreamWaitEvent(c->side, c->ev_fork, 0));
            sl = c->side;
        }
        local_kernel<<<c->local_blocks, kBlock, 0, sl>>>(c->da, c->P, c->x.p, c->fbuf.p);
        if (fork) CUDA_TRY(cudaEventRecord(c->ev_join, c->side));
        c->n_launches += 1;
    }
    for (auto& Lp : c->lists) {
        PairListHost& L = *Lp;
        if (!L.tiled || L.np == 0) continue;
        if (phase == 0 && L.tile_nblocks_in > 0) launch_tile(c, L, L.launch_in, L.tile_nblocks_in, s);
        launch_tile(c, L, L.launch, L.tile_nblocks, s);
    }
    if (c->pair_blocks > 0) {
        // (not chained when the donor-acceptor kernel runs on the forked side stream: the join is an event wait)
        if (c->periodic) launch_chained(c, force_kernel<true>, c->pair_blocks, kBlock, c->smem_bytes, s, c->d_segs.p, c->n_pair_segs, c->d_lists.p, c->d_bonded.p, c->P, c->x.p, c->fbuf.p);
        else launch_chained(c, force_kernel<false>, c->pair_blocks, kBlock, c->smem_bytes, s, c->d_segs.p, c->n_pair_segs, c->d_lists.p, c->d_bonded.p, c->P, c->x.p, c->fbuf.p);
        c->n_launches += 1;
    }
    if (fork) CUDA_TRY(cudaStreamWaitEvent(s, c->ev_join, 0));
    if (c->n_coms > 0) {
        com_kernel<<<c->n_coms, kComBlock, 0, s>>>(c->d_coms.p, c->P, c->x.p, c->fbuf.p, nullptr);
        c->n_launches += 1;
    }
}

// size (or re-size) the Verlet rows from a counting pass at the current positions
static void size_lists(ocm_context* c, double growth) {
    cudaStream_t s = c->stream;
    for (auto& L : c->lists) {
        L->tiled = L->tile_ok && (c->tile_opt > 0 || (c->tile_opt < 0 && L->tile_auto));
        L->dev.tiled = L->tiled ? 1 : 0;
        configure_grid(c, *L);
    }
    upload_lists(c);
    if (c->nrep > 1) {
        recentre_kernel<<<nblk(c->nrep * 32, 128), 128, 0, s>>>(c->P, c->x.p, c->origin.p);
        canon_kernel<<<nblk(c->n, 256), 256, 0, s>>>(c->P, c->x.p, c->posq.p);
    }
    if (c->dd) {
        // bring roles and halo up to date first (collective: every rank sizes its lists at the same points); the halo
        // messages grow until nothing overflows
        DDState& d = *c->dd;
        if (d.cap == 0) dd_alloc(c, (int)(1.5 * c->n * d.halo / c->sys.box[d.axis]) + 1024);
        CUDA_TRY(cudaMemcpyAsync(d.role_tmp.p, d.role.p, (size_t)c->n, cudaMemcpyDeviceToDevice, s));
        for (int attempt = 0;; attempt++) {
            CUDA_TRY(cudaMemsetAsync(c->flags.p, 0, FLAG_COUNT * sizeof(int), s));
            dd_exchange_rebuild(c, s);
            int fl[FLAG_COUNT];
            read_flags(c, fl);
            if (!fl[FLAG_DD_OVERFLOW]) break;
            if (attempt >= 8) throw std::runtime_error("slab decomposition: halo messages keep overflowing");
            CUDA_TRY(cudaMemcpyAsync(d.role.p, d.role_tmp.p, (size_t)c->n, cudaMemcpyDeviceToDevice, s));
            dd_alloc(c, std::max(2 * d.cap, (int)(1.5 * fl[FLAG_DD_MAXSEND])));
        }
        d.rows_est.assign(c->lists.size(), 0);
        // how many beads / bonded terms this rank works on: the integrate and bonded launches cover 1.3 x that
        // (more than that at a later rebuild raises the overflow flag: sized again)
        int h[1 + B_NUM_KINDS];
        CUDA_TRY(cudaMemcpy(h, d.counts.p, sizeof(h), cudaMemcpyDeviceToHost));
        d.present_est = (int)(h[0] * std::max(1.0, growth / 1.25));
        for (int k = 0; k < B_NUM_KINDS; k++) d.bterm_est[k] = (int)(h[1 + k] * std::max(1.0, growth / 1.25));
    }
    std::vector<int> fl(FLAG_COUNT, 0);
    // one counting pass per list: FLAG_MAXDISP+1 receives the largest neighbour count of that list
    for (size_t l = 0; l < c->lists.size(); l++) {
        auto& L = *c->lists[l];
        if (L.np == 0) continue;
        CUDA_TRY(cudaMemsetAsync(c->flags.p, 0, FLAG_COUNT * sizeof(int), s));
        const int li = (int)l;
        const double launches = c->n_launches;
        enqueue_list_rebuild(c, L, li, s, true);
        c->n_launches = launches;
        CUDA_TRY(cudaMemcpyAsync(fl.data(), c->flags.p, FLAG_COUNT * sizeof(int), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        if (c->dd) CUDA_TRY(cudaMemcpy(&c->dd->rows_est[l], L.cell_start.p + L.max_cells, sizeof(int), cudaMemcpyDeviceToHost));
        int maxn = fl[FLAG_MAXDISP + 1];
        if (L.tiled) {
            // entries per supercluster row, and the lane-mask blocks (counted: partial tiles + one per exclusion at most)
            int want = ((int)(maxn * growth) + 16 + 7) / 8 * 8;
            if (want > L.tl_stride) {
                L.tl_stride = want;
                L.tl_entries.alloc((size_t)L.tl_nsc * L.tl_stride);
            }
            int nmask = 0;
            CUDA_TRY(cudaMemcpy(&nmask, L.tl_mask_count.p, sizeof(int), cudaMemcpyDeviceToHost));
            const int cap = (int)(nmask * growth) + 64;
            if (cap > L.tl_mask_cap) { L.tl_mask_cap = cap; L.tl_masks.alloc((size_t)cap * 8); }
            L.dev.tl_stride = L.tl_stride; L.dev.tl_entries = L.tl_entries.p; L.dev.tl_masks = L.tl_masks.p; L.dev.tl_mask_cap = L.tl_mask_cap;
            L.stride = std::max(L.stride, 32);  // (rows are not built; a non-zero stride marks the list as sized)
            L.row_len32 = std::max(L.row_len32, 32);
            continue;
        }
        int want = (int)(maxn * growth) + 16;
        want = (want + 31) / 32 * 32;
        L.row_len32 = std::max(L.row_len32, want);
        // rows are padded with self entries to whole 128-entry groups (the T = 32 force path reads four 32-entry groups per
        // iteration without testing where the row ends)
        want = (want + 127) / 128 * 128;
        if (want > L.stride) {
            L.stride = want;
            L.nbr.alloc((size_t)std::max(L.np, 1) * L.stride);
        }
        L.dev.stride = L.stride; L.dev.nbr = L.nbr.p;
    }
    if (c->da.nterms > 0 && c->da.ncand > 0) {
        // counting pass with capacity 0: every hit only bumps the counter
        upload_da(c);
        DAListDev probe = c->da;
        probe.capacity = 0; probe.entries = nullptr;
        CUDA_TRY(cudaMemsetAsync(c->da_count.p, 0, sizeof(unsigned int), s));
        da_build_kernel<<<(unsigned int)((c->da.ncand + 255) / 256), 256, 0, s>>>(probe, c->P, c->x.p, c->flags.p);
        unsigned int cnt = 0;
        CUDA_TRY(cudaMemcpyAsync(&cnt, c->da_count.p, sizeof(cnt), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        int want = (int)(cnt * growth) + 256;
        want = (want + 127) / 128 * 128;
        if (want > c->da.capacity) {
            c->da_entries.alloc((size_t)want);
            c->da.capacity = want;
            c->da.entries = c->da_entries.p;
        }
    }
    CUDA_TRY(cudaMemsetAsync(c->flags.p, 0, FLAG_COUNT * sizeof(int), s));
    build_segments(c);
    upload_lists(c);  // after build_segments: it sets row_capacity
    c->lists_ready = true;
    c->lists_current = false;
}

static void ensure_lists_current(ocm_context* c) {
    if (!c->positions_set) throw std::runtime_error("positions have not been set");
    if (!c->lists_ready) size_lists(c, 1.25);
    if (c->lists_current) return;
    for (int attempt = 0; attempt < 6; attempt++) {
        if (c->dd) CUDA_TRY(cudaMemcpyAsync(c->dd->role_tmp.p, c->dd->role.p, (size_t)c->n, cudaMemcpyDeviceToDevice, c->stream));
        enqueue_rebuild(c, c->stream, false);
        int fl[FLAG_COUNT];
        read_flags(c, fl);
        if (!fl[FLAG_OVERFLOW] && !fl[FLAG_DA_OVERFLOW] && !fl[FLAG_DD_OVERFLOW] && !fl[FLAG_VBOX]) { c->lists_current = true; c->lists_fresh = true; return; }
        if (fl[FLAG_VBOX]) {
            // unlimited boundary: the beads left the range of the virtual box; lay it out again around them
            std::vector<double> hx((size_t)3 * c->n);
            CUDA_TRY(cudaMemcpy(hx.data(), c->x.p, hx.size() * sizeof(double), cudaMemcpyDeviceToHost));
            vbox_layout(c, hx.data(), true);
            canon_kernel<<<nblk(c->n, 256), 256, 0, c->stream>>>(c->P, c->x.p, c->posq.p);
            CUDA_TRY(cudaMemsetAsync(c->flags.p, 0, FLAG_COUNT * sizeof(int), c->stream));
            size_lists(c, 1.25);
            continue;
        }
        if (c->dd) {
            // an overflowed halo message may have dropped a migrating bead: go back to the roles before the attempt
            CUDA_TRY(cudaMemcpyAsync(c->dd->role.p, c->dd->role_tmp.p, (size_t)c->n, cudaMemcpyDeviceToDevice, c->stream));
            if (fl[FLAG_DD_OVERFLOW]) dd_alloc(c, std::max(2 * c->dd->cap, (int)(1.5 * fl[FLAG_DD_MAXSEND])));
        }
        CUDA_TRY(cudaMemsetAsync(c->flags.p, 0, FLAG_COUNT * sizeof(int), c->stream));
        size_lists(c, 1.5);
    }
    throw std::runtime_error("neighbour list keeps overflowing");
}

static void phase_mark(ocm_context* c, cudaStream_t s, int tag) {
    if (!c->phase_timing) return;
    if (c->ph_used == c->ph_ev.size()) {
        cudaEvent_t e;
        CUDA_TRY(cudaEventCreate(&e));
        c->ph_ev.push_back(e);
        c->ph_tag.push_back(0);
    }
    c->ph_tag[c->ph_used] = tag;
    CUDA_TRY(cudaEventRecord(c->ph_ev[c->ph_used++], s));
}
static void phase_collect(ocm_context* c) {
    for (size_t i = 1; i < c->ph_used; i++) {
        if (c->ph_tag[i] < 0) continue;  // start of a window
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, c->ph_ev[i - 1], c->ph_ev[i]));
        c->ph_ms[c->ph_tag[i]] += ms;
    }
    c->ph_used = 0;
}

// one window: rebuild, then nst x [forces, integrate (+ halo positions to the neighbouring slabs)]
static void enqueue_window(ocm_context* c, cudaStream_t s, int nst, bool lists_are_fresh = false) {
    phase_mark(c, s, -1);
    // (a Monte Carlo barostat trial has just built the lists at these very positions for its energy evaluation)
    if (!lists_are_fresh) enqueue_rebuild(c, s, false);
    phase_mark(c, s, 0);
    // slab decomposition with split tile lists and the peer-memory halo: the halo positions of a step travel on the side stream
    // while the interior tile entries (owned beads only) are already evaluated on the compute stream
    bool overlap = false;
    if (c->dd && c->dd->push_ok && c->halo_overlap) for (auto& L : c->lists) overlap = overlap || (L->tiled && L->np && L->tile_nblocks_in > 0);
    bool interior_done = false;
    for (int k = 0; k < nst; k++) {
        if (interior_done) {
            CUDA_TRY(cudaStreamWaitEvent(s, c->ev_halo_done, 0));
            phase_mark(c, s, 3);
            enqueue_forces(c, s, 2);
        } else enqueue_forces(c, s, 0);
        interior_done = false;
        phase_mark(c, s, 1);
        launch_chained(c, c->n < kIntegrateEarlyBeads ? integrate_kernel<true> : integrate_kernel<false>, nblk(c->dd ? c->dd->integrate_cover(c->n) : c->n, 128), 128, 0, s, c->P, c->x.p, c->v.p, c->fbuf.p, c->posq.p, c->d_lists.p, (int)c->lists.size(), c->xref.p,
                       c->step_base.p, k, nst, k == nst - 1 ? 1 : 0, c->flags.p);
        phase_mark(c, s, 2);
        // the next window starts with a full halo refresh, so the last step of a window ships nothing
        if (c->dd && k < nst - 1) {
            if (overlap) {
                CUDA_TRY(cudaEventRecord(c->ev_halo_go, s));
                CUDA_TRY(cudaStreamWaitEvent(c->side, c->ev_halo_go, 0));
                dd_halo_step(c, c->side);
                CUDA_TRY(cudaEventRecord(c->ev_halo_done, c->side));
                enqueue_forces(c, s, 1);
                phase_mark(c, s, 1);
                interior_done = true;
            } else { dd_halo_step(c, s); phase_mark(c, s, 3); }
        }
    }
}

static cudaGraphExec_t window_graph(ocm_context* c, int nst) {
    auto it = c->graphs.find(nst);
    if (it != c->graphs.end()) return it->second;
    cudaStream_t s = c->stream;
    cudaGraph_t g;
    double launches_before = c->n_launches, rebuilds_before = c->n_rebuilds;
    CUDA_TRY(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    enqueue_window(c, s, nst);
    CUDA_TRY(cudaStreamEndCapture(s, &g));
    c->n_launches = launches_before; c->n_rebuilds = rebuilds_before;  // counted per launch of the graph instead
    cudaGraphExec_t ge;
    CUDA_TRY(cudaGraphInstantiate(&ge, g, 0));
    CUDA_TRY(cudaGraphDestroy(g));
    c->graphs[nst] = ge;
    return ge;
}

static double launches_per_window(const ocm_context* c, int nst) {
    double per_rebuild = 1;
    for (auto& L : c->lists) if (L->np) per_rebuild += L->tiled ? 9 : (c->periodic ? 5 : 7);
    if (c->nrep > 1) per_rebuild += 2;
    if (c->da.nterms > 0 && c->da.ncand > 0) per_rebuild += 1;
    if (c->dd) per_rebuild += 5;
    double n_tiled = 0;
    for (auto& L : c->lists) if (L->np && L->tiled) n_tiled += 1;
    const double per_step = 1.0 + n_tiled + (c->pair_blocks > 0 ? 1.0 : 0.0) + (c->local_blocks > 0 ? 1.0 : 0.0);
    if (c->dd) return per_rebuild + 2.0 * (nst - 1) + nst * per_step;
    return per_rebuild + nst * per_step;
}

static int quantise_nst(double want) {
    int best = 1;
    for (int v : kNstChoices) if (v <= want) best = v;
    return best;
}

static int initial_nst(const ocm_context* c) {
    // ballistic bound: fastest bead ~ 5 sigma_v, lightest mass
    double mmin = 1e300;
    for (double m : c->sys.mass) if (m > 0) mmin = std::min(mmin, m);
    if (!(mmin < 1e300)) return 1;
    double vmax = 5.0 * std::sqrt(kBoltz * std::max(c->T, 1.0) / mmin);
    double n = 0.5 * c->skin / (vmax * c->dt);
    return quantise_nst(std::max(1.0, std::min(100.0, n)));
}

// small unlimited-boundary systems and replica ensembles: the whole MD step in one persistent kernel, one block per replica
static void do_steps_replica_kernel(ocm_context* c, int64_t nsteps) {
    cudaStream_t s = c->stream;
    RepParams& R = c->rep;
    R.dt = c->P.dt; R.a = c->P.a; R.b = c->P.b; R.c = c->P.c; R.seed_lo = c->P.seed_lo; R.seed_hi = c->P.seed_hi;
    CUDA_TRY(cudaMemsetAsync(c->flags.p, 0, FLAG_COUNT * sizeof(int), s));
    CUDA_TRY(cudaEventRecord(c->ev0, s));
    // (launches of at most 65 536 steps: a block holds its replica for the whole launch)
    for (int64_t done = 0; done < nsteps;) {
        const int chunk = (int)std::min<int64_t>(65536, nsteps - done);
        auto* kern = c->rep_block == 256 ? replica_step_kernel<256> : (c->rep_block == 512 ? replica_step_kernel<512> : replica_step_kernel<1024>);
        kern<<<c->nrep, c->rep_block, c->rep_smem, s>>>(R, c->x.p, c->v.p, (unsigned long long)(c->step + done), chunk, c->flags.p);
        c->n_launches += 1; c->n_rep_launches += 1;
        done += chunk;
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(c->ev1, s));
    // the general path (observer kernels, list builds) reads posq: bring it up to date
    canon_kernel<<<nblk(c->n, 256), 256, 0, s>>>(c->P, c->x.p, c->posq.p);
    int fl[FLAG_COUNT];
    CUDA_TRY(cudaMemcpyAsync(fl, c->flags.p, FLAG_COUNT * sizeof(int), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    if (fl[FLAG_NAN]) throw std::runtime_error("particle coordinate or force is NaN");
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->last_step_ms = ms;
    c->step += nsteps;
    c->lists_current = false;
}

static void do_steps(ocm_context* c, int64_t nsteps) {
    if (nsteps <= 0) return;
    if (!c->positions_set) throw std::runtime_error("positions have not been set");
    if (c->rep_ok && c->rep_opt) { do_steps_replica_kernel(c, nsteps); return; }
    if (!c->lists_ready) size_lists(c, 1.25);
    cudaStream_t s = c->stream;
    const size_t xb = (size_t)3 * c->n * sizeof(double), vb = (size_t)c->n * sizeof(float4);
    int64_t done = 0;
    int replays = 0;   // consecutive rollbacks of the same batch
    if (c->nstlist_fixed > 0) c->nst = c->nstlist_fixed;
    CUDA_TRY(cudaEventRecord(c->ev0, s));
    while (done < nsteps) {
        // one batch = up to 16 windows between host synchronisations
        CUDA_TRY(cudaMemcpyAsync(c->xbak.p, c->x.p, xb, cudaMemcpyDeviceToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(c->vbak.p, c->v.p, vb, cudaMemcpyDeviceToDevice, s));
        if (c->dd) CUDA_TRY(cudaMemcpyAsync(c->dd->role_bak.p, c->dd->role.p, (size_t)c->n, cudaMemcpyDeviceToDevice, s));
        const int64_t step0 = c->step + done;
        unsigned long long sb = (unsigned long long)step0;
        CUDA_TRY(cudaMemcpyAsync(c->step_base.p, &sb, sizeof(sb), cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemsetAsync(c->flags.p, 0, FLAG_COUNT * sizeof(int), s));
        int64_t batch = 0;
        int windows = 0;
        double launches = 0;
        const int nst = c->nst;
        while (done + batch < nsteps && windows < 16) {
            int w = (int)std::min<int64_t>(nst, nsteps - done - batch);
            if (c->dd || c->phase_timing || c->baro.kind) {
                // slab decomposition: direct launches (the exchanges are host-ordered across ranks); also when phases are timed,
                // and under a Monte Carlo barostat (the box, a kernel argument, changes every few steps)
                const double l0 = c->n_launches, r0 = c->n_rebuilds;
                // the first window after an accepted barostat trial starts on the lists of the trial's energy evaluation
                const bool fresh = c->baro.kind && !c->dd && windows == 0 && replays == 0 && c->lists_current && c->lists_fresh;
                enqueue_window(c, s, w, fresh);
                c->lists_current = false; c->lists_fresh = false;
                c->n_launches = l0; c->n_rebuilds = r0;
            } else CUDA_TRY(cudaGraphLaunch(window_graph(c, w), s));
            launches += launches_per_window(c, w);
            batch += w; windows++;
        }
        int fl[FLAG_COUNT];
        read_flags(c, fl);
        if (c->phase_timing) phase_collect(c);
        if (fl[FLAG_NAN]) throw std::runtime_error("particle coordinate or force is NaN");
        if (fl[FLAG_VIOLATION] || fl[FLAG_OVERFLOW] || fl[FLAG_DA_OVERFLOW] || fl[FLAG_DD_OVERFLOW] || fl[FLAG_VBOX]) {
            // roll back and replay the batch with a safer configuration
            CUDA_TRY(cudaMemcpyAsync(c->x.p, c->xbak.p, xb, cudaMemcpyDeviceToDevice, s));
            CUDA_TRY(cudaMemcpyAsync(c->v.p, c->vbak.p, vb, cudaMemcpyDeviceToDevice, s));
            CUDA_TRY(cudaMemsetAsync(c->fbuf.p, 0, (size_t)3 * c->n * sizeof(unsigned long long), s));
            canon_kernel<<<nblk(c->n, 256), 256, 0, s>>>(c->P, c->x.p, c->posq.p);
            c->n_rollbacks += 1;
            // (a batch that keeps failing must not spin: the retries of ensure_lists_current are bounded the same way)
            if (++replays > 8) throw std::runtime_error("a batch of steps keeps being rolled back (list overflow or Verlet-buffer violation)");
            if (c->dd) {
                CUDA_TRY(cudaMemcpyAsync(c->dd->role.p, c->dd->role_bak.p, (size_t)c->n, cudaMemcpyDeviceToDevice, s));
                if (fl[FLAG_DD_OVERFLOW]) dd_alloc(c, std::max(2 * c->dd->cap, (int)(1.5 * fl[FLAG_DD_MAXSEND])));
            }
            if (fl[FLAG_VBOX]) {
                // the system outgrew the range its virtual box was laid out for: lay it out again around the restored state
                std::vector<double> hx((size_t)3 * c->n);
                CUDA_TRY(cudaMemcpyAsync(hx.data(), c->x.p, xb, cudaMemcpyDeviceToHost, s));
                CUDA_TRY(cudaStreamSynchronize(s));
                vbox_layout(c, hx.data(), true);
                canon_kernel<<<nblk(c->n, 256), 256, 0, s>>>(c->P, c->x.p, c->posq.p);
                size_lists(c, 1.25);
            }
            // the counting pass of size_lists sees the restored positions, not the step that overflowed: from the second replay on
            // grow beyond what it asks for
            if (fl[FLAG_OVERFLOW] || fl[FLAG_DA_OVERFLOW] || fl[FLAG_DD_OVERFLOW]) size_lists(c, replays > 1 ? 1.5 + 0.5 * replays : 1.5);
            if (fl[FLAG_VIOLATION]) {
                if (c->nst <= 1) throw std::runtime_error("a particle crosses the Verlet buffer within one step; increase skin_nm or reduce the time step");
                c->nst = quantise_nst(std::max(1.0, c->nst * 0.6));
            }
            continue;
        }
        c->n_launches += launches; c->n_windows += windows; c->n_rebuilds += windows;
        done += batch;
        replays = 0;
        {
            float d2max;
            std::memcpy(&d2max, &fl[FLAG_MAXDISP], sizeof(float));
            c->last_disp2_ok = d2max <= (float)(0.25 * c->skin * c->skin);
        }
        // adapt the rebuild interval from the largest end-of-window displacement (only full windows count)
        if (c->nstlist_fixed <= 0 && batch >= nst) {
            float d2;
            std::memcpy(&d2, &fl[FLAG_MAXDISP], sizeof(float));
            double disp = std::sqrt((double)d2), half = 0.5 * c->skin;
            if (disp > 0) {
                double ratio = 0.70 * half / disp;  // aim the last step at 70 % of the buffer
                double want = c->nst * std::min(1.5, std::max(0.6, ratio));
                int q = quantise_nst(std::max(1.0, want));
                if (q < c->nst || ratio > 1.25) c->nst = q;
            }
        }
    }
    CUDA_TRY(cudaEventRecord(c->ev1, s));
    CUDA_TRY(cudaEventSynchronize(c->ev1));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->last_step_ms = ms;
    c->step += nsteps;
    // The pair lists of the last window still hold every pair inside the cutoffs if no bead has moved more than half the skin
    // from the list's reference positions (the cell-sorted positions are refreshed by every step): the observer can evaluate
    // energies and forces on them without a list build.  (Slab decomposition: halos are only complete right after a rebuild.)
    c->lists_current = !c->dd && c->nrep == 1 && c->last_disp2_ok;
    c->lists_fresh = false;
}

// per-group energies of every term at the current positions (lists must be current), added into egroup
static void enqueue_energies(ocm_context* c, cudaStream_t s) {
    if (c->total_blocks > 0) {
        if (c->periodic) energy_kernel<true><<<c->total_blocks, kBlock, 0, s>>>(c->d_segs.p, (int)c->segs.size(), c->d_lists.p, c->d_bonded.p, c->da, c->P, c->x.p, c->egroup.p);
        else energy_kernel<false><<<c->total_blocks, kBlock, 0, s>>>(c->d_segs.p, (int)c->segs.size(), c->d_lists.p, c->d_bonded.p, c->da, c->P, c->x.p, c->egroup.p);
    }
    for (size_t l = 0; l < c->lists.size(); l++) {
        auto& L = *c->lists[l];
        if (L.np == 0 || !L.tiled) continue;
        tile_energy_kernel<true><<<L.dev.row_capacity / 64, kBlock, 0, s>>>(c->d_lists.p, (int)l, c->P, c->egroup.p);
    }
    if (c->n_coms > 0) com_kernel<<<c->n_coms, kComBlock, 0, s>>>(c->d_coms.p, c->P, c->x.p, nullptr, c->egroup.p);
}

static void compute_forces_into_fbuf(ocm_context* c) {
    ensure_lists_current(c);
    CUDA_TRY(cudaMemsetAsync(c->fbuf.p, 0, (size_t)3 * c->n * sizeof(unsigned long long), c->stream));
    enqueue_forces(c, c->stream);
}

// ---- Monte Carlo barostat -----------------------------------------------------------------------------------
// Context::setPeriodicBoxVectors: everything laid out in the box (fixed-point positions, cell grids, tile tables, list
// capacities) is derived again at the next list build; positions are kept.
static void apply_box(ocm_context* c, const double* nb) {
    if (!c->periodic || c->vbox) throw std::runtime_error("set_box: the context has no periodic box");
    for (int k = 0; k < 3; k++) if (!(nb[k] > 0) || !std::isfinite(nb[k])) throw std::runtime_error("set_box: box edge must be positive");
    double rmax = c->da_rcmax;
    for (auto& L : c->lists) if (L->np) rmax = std::max(rmax, L->rc_list);
    double skin = c->skin;
    for (int k = 0; k < 3; k++) {
        if (2.0 * rmax > nb[k]) throw std::runtime_error("The periodic box size has decreased to less than twice the nonbonded cutoff.");
        if (rmax + skin > 0.5 * nb[k]) skin = std::max(0.0, 0.5 * nb[k] - rmax) * 0.999;
    }
    c->skin = skin;
    for (int k = 0; k < 3; k++) c->sys.box[k] = nb[k];
    update_params(c);
    if (c->positions_set) canon_kernel<<<nblk(c->n, 256), 256, 0, c->stream>>>(c->P, c->x.p, c->posq.p);
    if (c->lists_ready) {
        // the capacities found by the last sizing pass stay (a barostat trial changes an edge by a per cent or so; a list that
        // outgrows its rows raises the overflow flag and is sized again): only the cell grids and what is derived from the box
        bool tiled = false;
        for (auto& L : c->lists) { configure_grid(c, *L); tiled = tiled || L->tiled; }
        if (tiled) build_segments(c);   // (the tile launch records hold the box)
        upload_lists(c);
    }
    c->lists_current = false;
    drop_graphs(c);
}

static double potential_energy(ocm_context* c) {
    cudaStream_t s = c->stream;
    ensure_lists_current(c);
    CUDA_TRY(cudaMemsetAsync(c->egroup.p, 0, 40 * sizeof(double), s));
    enqueue_energies(c, s);
    double he[40];
    CUDA_TRY(cudaMemcpyAsync(he, c->egroup.p, sizeof(he), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    double e = 0;
    for (int g = 0; g < 32; g++) e += he[g];
    return e;
}

// OpenMM's getMolecules(): connected components of the graph whose edges are the bonded terms' consecutive bodies
static void barostat_molecules(ocm_context* c) {
    const int n = c->n;
    std::vector<int> parent(n);
    for (int i = 0; i < n; i++) parent[i] = i;
    auto find = [&](int i) { while (parent[i] != i) { parent[i] = parent[parent[i]]; i = parent[i]; } return i; };
    for (const auto& t : c->sys.bonded)
        for (int64_t m = 0; m < t.n(); m++)
            for (int b = 1; b < t.nbody; b++) {
                int ra = find(t.idx[(size_t)m * t.nbody + b - 1]), rb = find(t.idx[(size_t)m * t.nbody + b]);
                if (ra != rb) parent[std::max(ra, rb)] = std::min(ra, rb);
            }
    // molecules in the order of their first bead, beads ascending
    std::vector<int> id(n, -1), count;
    for (int i = 0; i < n; i++) { int r = find(i); if (id[r] < 0) { id[r] = (int)count.size(); count.push_back(0); } count[id[r]]++; }
    std::vector<int> start(count.size() + 1, 0), bead(n);
    for (size_t m = 0; m < count.size(); m++) start[m + 1] = start[m] + count[m];
    std::vector<int> cur(start.begin(), start.end() - 1);
    for (int i = 0; i < n; i++) bead[cur[id[find(i)]]++] = i;
    c->baro.nmol = (int)count.size();
    c->baro.mol_start.upload(start);
    c->baro.mol_bead.upload(bead);
    c->baro.xsave.alloc((size_t)3 * n);
}

static inline double baro_uniform(ocm_context::Barostat& B) {
    uint64_t z = (B.rng += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

// one trial (MonteCarloAnisotropicBarostatImpl / MonteCarloMembraneBarostatImpl::updateContextState of OpenMM 8.0)
static void barostat_trial(ocm_context* c) {
    auto& B = c->baro;
    cudaStream_t s = c->stream;
    const double kBar = 6.02214076e23 * 1e-25;  // bar -> kJ/mol/nm^3
    const double e0 = potential_energy(c);
    int axis = 0;
    for (;;) {
        const double rnd = baro_uniform(B) * 3.0;
        if (B.kind == OCM_BAROSTAT_ANISOTROPIC) {
            if (rnd < 1.0) { if (B.mode[0]) { axis = 0; break; } }
            else if (rnd < 2.0) { if (B.mode[1]) { axis = 1; break; } }
            else if (B.mode[2]) { axis = 2; break; }
        } else {
            if (rnd < 1.0) { axis = 0; break; }
            else if (rnd < 2.0) { if (B.mode[0] == 1) { axis = 1; break; } }
            else if (B.mode[1] == 0) { axis = 2; break; }
        }
    }
    double box[3] = {c->sys.box[0], c->sys.box[1], c->sys.box[2]};
    const double volume = box[0] * box[1] * box[2];
    double dV = B.volume_scale[axis] * 2.0 * (baro_uniform(B) - 0.5);
    double new_volume = volume + dV;
    double scale[3] = {1.0, 1.0, 1.0};
    if (B.kind == OCM_BAROSTAT_MEMBRANE && B.mode[0] == 0 && axis == 0) scale[0] = scale[1] = std::sqrt(new_volume / volume);
    else scale[axis] = new_volume / volume;
    if (B.kind == OCM_BAROSTAT_MEMBRANE && B.mode[1] == 2 && axis != 2) {
        scale[2] = 1.0 / (scale[0] * scale[1]);
        new_volume = volume; dV = 0.0;
    }
    const double nb[3] = {box[0] * scale[0], box[1] * scale[1], box[2] * scale[2]};
    const double dA = nb[0] * nb[1] - box[0] * box[1];
    const double skin0 = c->skin;
    CUDA_TRY(cudaMemcpyAsync(B.xsave.p, c->x.p, (size_t)3 * c->n * sizeof(double), cudaMemcpyDeviceToDevice, s));
    scale_molecules_kernel<<<nblk(B.nmol, 128), 128, 0, s>>>(B.nmol, B.mol_start.p, B.mol_bead.p, c->x.p, box[0], box[1], box[2], scale[0], scale[1], scale[2]);
    apply_box(c, nb);
    const double e1 = potential_energy(c);
    const double kT = kBoltz * B.T;
    const double pressure = (B.kind == OCM_BAROSTAT_ANISOTROPIC ? B.pressure[axis] : B.pressure[0]) * kBar;
    const double tension = (B.kind == OCM_BAROSTAT_MEMBRANE ? B.tension : 0.0) * kBar;
    const double w = e1 - e0 + pressure * dV - tension * dA - B.nmol * kT * std::log(new_volume / volume);
    bool accept = true;
    if (w > 0 && baro_uniform(B) > std::exp(-w / kT)) {
        accept = false;
        CUDA_TRY(cudaMemcpyAsync(c->x.p, B.xsave.p, (size_t)3 * c->n * sizeof(double), cudaMemcpyDeviceToDevice, s));
        c->skin = skin0;
        apply_box(c, box);
    } else B.accepted[axis]++;
    B.attempted[axis]++;
    if (B.attempted[axis] >= 10) {
        if (B.accepted[axis] < 0.25 * B.attempted[axis]) { B.volume_scale[axis] /= 1.1; B.attempted[axis] = 0; B.accepted[axis] = 0; }
        else if (B.accepted[axis] > 0.75 * B.attempted[axis]) {
            B.volume_scale[axis] = std::min(B.volume_scale[axis] * 1.1, volume * 0.3);
            B.attempted[axis] = 0; B.accepted[axis] = 0;
        }
    }
    B.n_trials += 1; B.n_accepted += accept ? 1 : 0;
    B.last[0] = axis; B.last[1] = dV; B.last[2] = e0; B.last[3] = e1; B.last[4] = w; B.last[5] = accept ? 1 : 0;
}

// ocm_step under a barostat: OpenMM calls the barostat before every MD step; at the call where its counter reaches `frequency`
// it makes a trial, i.e. before MD steps frequency, 2 frequency, ... (counted from 1) of the context
static void do_steps_barostat(ocm_context* c, int64_t nsteps) {
    auto& B = c->baro;
    double ms = 0;
    while (nsteps > 0) {
        if (B.until_trial == 0) {
            const auto t0 = std::chrono::steady_clock::now();
            barostat_trial(c);
            ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            B.until_trial = B.frequency;
        }
        const int64_t k = std::min(nsteps, B.until_trial);
        do_steps(c, k);
        ms += c->last_step_ms;
        B.until_trial -= k; nsteps -= k;
    }
    c->last_step_ms = ms;
}

// Context::getState under a slab decomposition: every rank contributes what it owns, the sums over ranks (exact for
// positions, velocities and fixed-point forces: all other contributions are zero) give every rank the whole state
static void get_state_dd(ocm_context* c, uint32_t mask, double* pos, double* vel, double* force, double* epot, double* ekin, double* ebg) {
    DDState& d = *c->dd;
    cudaStream_t s = c->stream;
    const int n = c->n;
    if (force && mask != 0xFFFFFFFFu) throw std::runtime_error("get_state: forces restricted to a force-group mask are not supported; request all groups");
    ensure_lists_current(c);
    std::vector<double> he(40, 0.0);
    CUDA_TRY(cudaMemsetAsync(c->egroup.p, 0, he.size() * sizeof(double), s));
    if (epot || ebg) enqueue_energies(c, s);
    if (force || ekin) {
        compute_forces_into_fbuf(c);
        if (ekin) kinetic_kernel<<<nblk(n, kBlock), kBlock, 0, s>>>(c->P, c->v.p, c->fbuf.p, c->egroup.p + 32);
    }
    if (pos && d.gx.n < (size_t)3 * n) d.gx.alloc((size_t)3 * n);
    if (vel && d.gv.n < (size_t)3 * n) d.gv.alloc((size_t)3 * n);
    if (force && d.gf.n < (size_t)3 * n) d.gf.alloc((size_t)3 * n);
    if (pos || vel || force)
        dd_mask_state_kernel<<<nblk(n, 256), 256, 0, s>>>(n, d.role.p, c->x.p, c->v.p, c->fbuf.p, pos ? d.gx.p : nullptr, vel ? d.gv.p : nullptr,
                                                          force ? d.gf.p : nullptr);
    if (force || ekin) CUDA_TRY(cudaMemsetAsync(c->fbuf.p, 0, (size_t)3 * n * sizeof(unsigned long long), s));
    if (pos) d.comm->allreduce(d.rank, d.gx.p, (size_t)3 * n, DD_F64, false, s);
    if (vel) d.comm->allreduce(d.rank, d.gv.p, (size_t)3 * n, DD_F32, false, s);
    if (force) d.comm->allreduce(d.rank, d.gf.p, (size_t)3 * n, DD_I64, false, s);
    if (epot || ekin || ebg) d.comm->allreduce(d.rank, c->egroup.p, 40, DD_F64, false, s);
    std::vector<float> hv;
    std::vector<long long> hf;
    if (pos) CUDA_TRY(cudaMemcpyAsync(pos, d.gx.p, (size_t)3 * n * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (vel) { hv.resize((size_t)3 * n); CUDA_TRY(cudaMemcpyAsync(hv.data(), d.gv.p, hv.size() * sizeof(float), cudaMemcpyDeviceToHost, s)); }
    if (force) { hf.resize((size_t)3 * n); CUDA_TRY(cudaMemcpyAsync(hf.data(), d.gf.p, hf.size() * sizeof(long long), cudaMemcpyDeviceToHost, s)); }
    CUDA_TRY(cudaMemcpyAsync(he.data(), c->egroup.p, he.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    if (vel) for (size_t k = 0; k < hv.size(); k++) vel[k] = hv[k];
    if (force) for (int i = 0; i < n; i++)
        for (int k = 0; k < 3; k++) force[3 * i + k] = (double)hf[(size_t)k * n + i] / kFixScale;
    if (epot) { double e = 0; for (int g = 0; g < 32; g++) if (mask >> g & 1) e += he[g]; *epot = e; }
    if (ebg) std::copy(he.begin(), he.begin() + 32, ebg);
    if (ekin) *ekin = he[32];
}

}  // namespace ocm

// ================================================================================================
#define OCM_GUARD_BEGIN try {
#define OCM_GUARD_END                                             \
    }                                                             \
    catch (const std::exception& e) { return ocm::fail(e.what()); } \
    catch (...) { return ocm::fail("unknown error"); }

extern "C" {

const char* ocm_last_error(void) { return g_err.c_str(); }
const char* ocm_version(void) { return "ocm-b200 0.1 (sm_100a)"; }
int ocm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int ocm_system_create(int32_t n, const double* mass, const double* box, ocm_system** out) {
    if (n <= 0 || !mass || !out) return fail("system_create: bad arguments");
    if (n >= (1 << 24)) return fail("system_create: at most 16,777,215 particles are supported");
    if (box) for (int k = 0; k < 3; k++) if (!(box[k] > 0)) return fail("system_create: box edge must be positive");
    auto* s = new ocm_system();
    s->n = n; s->mass.assign(mass, mass + n); s->has_box = box != nullptr;
    for (int k = 0; k < 3; k++) s->box[k] = box ? box[k] : 0.0;
    *out = s;
    return 0;
}
void ocm_system_destroy(ocm_system* s) { delete s; }
int ocm_system_num_particles(const ocm_system* s) { return s ? s->n : -1; }

int ocm_system_add_harmonic_bond(ocm_system* s, int g, int per, int64_t n, const int32_t* ij, const double* r0, const double* k) {
    return add_bonded(s, B_HARMONIC_BOND, g, per, n, ij, {r0, k}, "harmonic_bond");
}
int ocm_system_add_gaussian_bond(ocm_system* s, int g, int per, int64_t n, const int32_t* ij, const double* k, const double* r0, const double* sigma) {
    return add_bonded(s, B_GAUSSIAN_BOND, g, per, n, ij, {k, r0, sigma}, "gaussian_bond");
}
// (Go contacts carry five parameters k, r0, slope, intercept, cap; cap = 0 is the plain 12-10 contact)
int ocm_system_add_go_contact(ocm_system* s, int g, int per, int64_t n, const int32_t* ij, const double* k, const double* r0) {
    if (n < 0) return fail("go_contact: negative count");
    const std::vector<double> zero((size_t)std::max<int64_t>(n, 1), 0.0);
    return add_bonded(s, B_GO_CONTACT, g, per, n, ij, {k, r0, zero.data(), zero.data(), zero.data()}, "go_contact");
}
int ocm_system_add_capped_go_contact(ocm_system* s, int g, int per, int64_t n, const int32_t* ij, const double* k, const double* r0,
                                     const double* slope, const double* intercept, const double* cap) {
    if (!slope || !intercept || !cap) return fail("capped_go_contact: null argument");
    return add_bonded(s, B_GO_CONTACT, g, per, n, ij, {k, r0, slope, intercept, cap}, "capped_go_contact");
}
int ocm_system_add_3spn2_bond(ocm_system* s, int g, int per, int64_t n, const int32_t* ij, const double* k2, const double* k4, const double* r0) {
    return add_bonded(s, B_3SPN2_BOND, g, per, n, ij, {k2, k4, r0}, "3spn2_bond");
}
int ocm_system_add_harmonic_angle(ocm_system* s, int g, int per, int64_t n, const int32_t* ijk, const double* th0, const double* k) {
    return add_bonded(s, B_HARMONIC_ANGLE, g, per, n, ijk, {th0, k}, "harmonic_angle");
}
int ocm_system_add_flexible_local_angle(ocm_system* s, int g, int per, int64_t n, const int32_t* ijk, const double* k,
                                        const double* y10, double tmin, double tmax, double emin) {
    if (!y10 || !(tmax > tmin)) return fail("flexible_local_angle: bad table");
    if (add_bonded(s, B_FLEX_ANGLE, g, per, n, ijk, {k}, "flexible_local_angle")) return 1;
    BondedTerm& t = s->bonded.back();
    for (int q = 0; q < 10; q++) t.tab[q] = y10[q];
    natural_spline_second_derivatives(y10, 10, (tmax - tmin) / 9.0, t.tab + 10);
    t.tab[20] = tmin; t.tab[21] = tmax; t.tab[22] = emin;
    return 0;
}
int ocm_system_add_gaussian_dihedral(ocm_system* s, int g, int per, int64_t n, const int32_t* ijkl, const double* k, const double* phi0, const double* sigma) {
    return add_bonded(s, B_GAUSSIAN_DIHEDRAL, g, per, n, ijkl, {k, phi0, sigma}, "gaussian_dihedral");
}
int ocm_system_add_cosine_dihedral(ocm_system* s, int g, int per, int64_t n, const int32_t* ijkl, const double* k, const double* t0, const int32_t* nper) {
    if (n > 0 && !nper) return fail("cosine_dihedral: null periodicity array");
    std::vector<double> nn((size_t)std::max<int64_t>(n, 0));
    for (int64_t m = 0; m < n; m++) nn[m] = nper[m];
    return add_bonded(s, B_COSINE_DIHEDRAL, g, per, n, ijkl, {k, t0, nn.data()}, "cosine_dihedral");
}
int ocm_system_add_flexible_local_dihedral(ocm_system* s, int g, int per, int64_t n, const int32_t* ijkl, const double* k, const double* coef7, double emin) {
    if (!coef7) return fail("flexible_local_dihedral: null coefficient table");
    if (add_bonded(s, B_FLEX_DIHEDRAL, g, per, n, ijkl, {k}, "flexible_local_dihedral")) return 1;
    BondedTerm& t = s->bonded.back();
    for (int q = 0; q < 7; q++) t.tab[q] = coef7[q];
    t.tab[7] = emin;
    return 0;
}

int ocm_system_add_3spn2_base_stacking(ocm_system* s, int g, int per, int64_t n, const int32_t* ijk, const double* eps, const double* r0,
                                       const double* theta0, double alpha, double K_BS) {
    std::vector<double> al((size_t)std::max<int64_t>(n, 0), alpha), kb((size_t)std::max<int64_t>(n, 0), K_BS);
    return add_bonded(s, B_3SPN2_STACK, g, per, n, ijk, {eps, r0, theta0, al.data(), kb.data()}, "3spn2_base_stacking");
}
int ocm_system_add_3spn2_base_pair(ocm_system* s, int g, const ocm_da_topology* topo, const double* eps, const double* r0,
                                   const double* t01, const double* t02, const double* phi0, double alpha, double K_BP) {
    return add_da(s, DA_BASE_PAIR, g, topo, true, {eps, r0, t01, t02, phi0}, alpha, K_BP, 0, "3spn2_base_pair");
}
int ocm_system_add_3spn2_cross_stacking(ocm_system* s, int g, const ocm_da_topology* topo, const double* eps, const double* r0,
                                        const double* t03, const double* t0CS, double K_BP, double K_CS, double alpha) {
    return add_da(s, DA_CROSS_STACK, g, topo, false, {eps, r0, t03, t0CS}, K_BP, K_CS, alpha, "3spn2_cross_stacking");
}
static int add_external(ocm_system* s, int form, int g, int per, int64_t n, const int32_t* index, const std::vector<std::pair<const double*, int>>& cols,
                        const std::vector<double>& globals, const char* what) {
    // one row of kExternalNpar doubles per term: form, the globals, then the per-term columns (pointer, width)
    if (n < 0 || (n > 0 && !index)) return fail(std::string(what) + ": bad arguments");
    for (auto& cw : cols) if (n > 0 && !cw.first) return fail(std::string(what) + ": null parameter array");
    std::vector<std::vector<double>> col(kExternalNpar, std::vector<double>((size_t)n, 0.0));
    std::fill(col[0].begin(), col[0].end(), (double)form);
    int q = 1;
    for (double v : globals) { std::fill(col[q].begin(), col[q].end(), v); q++; }
    for (auto& cw : cols)
        for (int w = 0; w < cw.second; w++, q++)
            for (int64_t m = 0; m < n; m++) col[q][m] = cw.first[(size_t)m * cw.second + w];
    std::vector<const double*> ptr;
    for (auto& v : col) ptr.push_back(n > 0 ? v.data() : reinterpret_cast<const double*>(&q));
    return add_bonded(s, B_EXTERNAL, g, per, n, index, ptr, what);
}
int ocm_system_add_position_restraint(ocm_system* s, int g, int64_t n, const int32_t* index, const double* pos3, const double* k, const double* v0) {
    return add_external(s, X_POSITION_RESTRAINT, g, 0, n, index, {{pos3, 3}, {k, 1}, {v0, 1}}, {}, "position_restraint");
}
int ocm_system_add_pulling(ocm_system* s, int g, int64_t n, const int32_t* index, const double* force3) {
    return add_external(s, X_PULLING, g, 0, n, index, {{force3, 3}}, {}, "pulling");
}
int ocm_system_add_exv_rectangular_box(ocm_system* s, int g, double eps, const double* lower3, const double* upper3, int64_t n, const int32_t* index,
                                       const double* radius) {
    if (!s || !lower3 || !upper3) return fail("exv_rectangular_box: null argument");
    if (s->has_box) return fail("exv_rectangular_box: RectangularBox interaction does not work under periodic boundary condition");
    for (int a = 0; a < 3; a++) if (!(upper3[a] > lower3[a])) return fail("exv_rectangular_box: the upper edge must be larger than the lower edge");
    return add_external(s, X_RECTANGULAR_BOX, g, 0, n, index, {{radius, 1}},
                        {eps, lower3[0], lower3[1], lower3[2], upper3[0], upper3[1], upper3[2]}, "exv_rectangular_box");
}
int ocm_system_add_cylindrical_restraint(ocm_system* s, int g, int per, int64_t n, const int32_t* index, const double* axis3, const double* shift3,
                                         const double* k, const double* v0) {
    return add_external(s, X_CYLINDRICAL_RESTRAINT, g, per, n, index, {{axis3, 3}, {shift3, 3}, {k, 1}, {v0, 1}}, {}, "cylindrical_restraint");
}
int ocm_system_add_harmonic_com_pulling(ocm_system* s, int g, int per, double k, double v0, int64_t n1, const int32_t* g1, int64_t n2, const int32_t* g2) {
    if (!s) return fail("null system");
    if (g < 0 || g > 31) return fail("force group must be in 0..31");
    if (n1 <= 0 || n2 <= 0 || !g1 || !g2) return fail("harmonic_com_pulling: both bead groups must be non-empty");
    if (per && !s->has_box) return fail("harmonic_com_pulling: periodic term without a box");
    if (check_indices(s, g1, n1, "harmonic_com_pulling") || check_indices(s, g2, n2, "harmonic_com_pulling")) return 1;
    ComTerm t;
    t.group = g; t.periodic = per ? 1 : 0; t.k = k; t.v0 = v0;
    t.g1.assign(g1, g1 + n1); t.g2.assign(g2, g2 + n2);
    for (const auto* grp : {&t.g1, &t.g2}) {
        double M = 0;
        for (int32_t i : *grp) M += s->mass[i];
        if (!(M > 0)) return fail("harmonic_com_pulling: a bead group has no mass");
    }
    s->coms.push_back(std::move(t));
    return 0;
}
int ocm_system_add_combinatorial_go_contact(ocm_system* s, int g, const ocm_da_topology* topo, double k, double r0) {
    if (!(r0 > 0)) return fail("combinatorial_go_contact: r0 must be positive");
    return add_da(s, DA_COMB_GO, g, topo, true, {}, k, r0, 0, "combinatorial_go_contact");
}
int ocm_system_add_pdns(ocm_system* s, int g, const ocm_da_topology* topo, const double* k, const double* r0, const double* theta0,
                        const double* phi0, double sigma, double delta) {
    if (!(sigma > 0) || !(delta > 0)) return fail("pdns: sigma and delta must be positive");
    return add_da(s, DA_PDNS, g, topo, true, {k, r0, theta0, phi0}, sigma, delta, 0, "pdns");
}

int ocm_system_add_excluded_volume(ocm_system* s, int g, double eps, double rc, const double* radius, const ocm_pair_topology* topo) {
    return add_pair(s, P_EXV, g, rc, rc, eps, std::pow(1.0 / rc, 12), 0, radius, nullptr, topo, "excluded_volume");
}
int ocm_system_add_debye_huckel(ocm_system* s, int g, double K, double debye, double rc, double cc, const double* charge, const ocm_pair_topology* topo) {
    if (!(debye > 0)) return fail("debye_huckel: debye length must be positive");
    return add_pair(s, P_DH, g, rc, rc, K, debye, cc, charge, nullptr, topo, "debye_huckel");
}
int ocm_system_add_lj_repulsive(ocm_system* s, int g, double rc, const double* eps, const double* sigma, const ocm_pair_topology* topo) {
    if (!sigma) return fail("lj_repulsive: null sigma");
    return add_pair(s, P_LJREP, g, rc, rc, 0, 0, 0, eps, sigma, topo, "lj_repulsive");
}
static int add_well(ocm_system* s, int kind, int g, double rc, const double* eps, const double* sigma, const double* omega,
                    const ocm_pair_topology* topo, const char* what) {
    if (!eps || !sigma || !omega) return fail(std::string(what) + ": null parameter array");
    return add_pair(s, kind, g, rc, rc, 0, 0, 0, eps, sigma, topo, what, omega);
}
int ocm_system_add_isolf_attractive(ocm_system* s, int g, double rc, const double* eps, const double* sigma, const double* omega,
                                    const ocm_pair_topology* topo) {
    return add_well(s, P_ISOLF, g, rc, eps, sigma, omega, topo, "isolf_attractive");
}
int ocm_system_add_trigonometric(ocm_system* s, int g, double rc, const double* eps, const double* sigma, const double* omega,
                                 const ocm_pair_topology* topo) {
    return add_well(s, P_TRIG, g, rc, eps, sigma, omega, topo, "trigonometric");
}
int ocm_system_add_wca(ocm_system* s, int g, double rc, const double* eps, const double* sigma, const ocm_pair_topology* topo) {
    if (!sigma) return fail("wca: null sigma");
    return add_pair(s, P_WCA, g, rc, rc, 0, 0, 0, eps, sigma, topo, "wca");
}
int ocm_system_add_lj_attractive(ocm_system* s, int g, double rc, double rsw, const double* eps, const double* sigma, const ocm_pair_topology* topo) {
    if (!sigma) return fail("lj_attractive: null sigma");
    if (!(rsw < rc)) return fail("lj_attractive: r_switch must be below rc");
    return add_pair(s, P_LJATT, g, rc, rsw, 0, 0, 0, eps, sigma, topo, "lj_attractive");
}
int ocm_system_add_3spn2_excluded_volume(ocm_system* s, int g, double rc, const double* eps, const double* sigma, const ocm_pair_topology* topo) {
    if (!sigma) return fail("3spn2_excluded_volume: null sigma");
    return add_pair(s, P_3SPN2EXV, g, rc, rc, 0, 0, 0, eps, sigma, topo, "3spn2_excluded_volume");
}

static int create_context(const ocm_system* sys, int nrep, double T, double gamma, double dt, const uint64_t* seeds, int device,
                          ocm_context** out, ocm_comm* comm = nullptr, int dd_rank = 0, int dd_axis = 1) {
    if (!sys || !out || !seeds) return fail("context_create: bad arguments");
    if (!(dt > 0)) return fail("context_create: time step must be positive");
    if (nrep < 1) return fail("context_create_batch: at least one replica");
    if (nrep > 1 && sys->has_box) return fail("context_create_batch: replica batches need an unlimited boundary (no box)");
    if (nrep > 1 && !sys->das.empty()) return fail("context_create_batch: donor-acceptor terms are not supported in replica batches yet");
    if (nrep > 1 && !sys->coms.empty()) return fail("context_create_batch: centre-of-mass pulling terms are not supported in replica batches");
    if ((int64_t)sys->n * nrep >= (1 << 24)) return fail("context_create: at most 16,777,215 particles are supported");
    if (comm) {
        if (!comm->impl) return fail("context_create_slab: communicator is closed");
        if (!sys->has_box) return fail("context_create_slab: a slab decomposition needs a periodic box");
        if (!sys->das.empty()) return fail("context_create_slab: donor-acceptor terms are not supported under a slab decomposition");
        if (!sys->coms.empty()) return fail("context_create_slab: centre-of-mass pulling terms are not supported under a slab decomposition");
        if (dd_axis < 0 || dd_axis > 2) return fail("context_create_slab: axis must be 0, 1 or 2");
        if (dd_rank < 0 || dd_rank >= comm->impl->world) return fail("context_create_slab: rank out of range");
        if (comm->rank >= 0 && comm->rank != dd_rank) return fail("context_create_slab: the communicator belongs to another rank");
        for (const auto& t : sys->pairs) if (!t.periodic) return fail("context_create_slab: every pair term must be periodic");
    }
    int ndev = ocm_device_count();
    if (ndev <= 0) return fail("context_create: no CUDA device available (this engine has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail("context_create: device index out of range");
    ocm_context* c = nullptr;
    OCM_GUARD_BEGIN
    CUDA_TRY(cudaSetDevice(device));
    c = new ocm_context();
    c->nrep = nrep; c->nper = sys->n;
    if (comm && comm->impl->world > 1) {
        c->dd.reset(new DDState());
        c->dd->comm = comm->impl; c->dd->rank = dd_rank; c->dd->world = comm->impl->world; c->dd->axis = dd_axis;
        c->dd->role.alloc(sys->n); c->dd->role_bak.alloc(sys->n); c->dd->role_tmp.alloc(sys->n); c->dd->known.alloc(sys->n);
        c->dd->role.zero(); c->dd->role_bak.zero(); c->dd->role_tmp.zero(); c->dd->known.zero();
        c->dd->present.alloc(sys->n); c->dd->counts.alloc(1 + B_NUM_KINDS); c->dd->counts.zero();
        if (auto* lc = dynamic_cast<LocalComm*>(comm->impl)) lc->attach(dd_rank, device);
    }
    if (nrep == 1) c->sys = *sys;
    else {
        c->sys1 = *sys;
        // the batch is one big system: every table replicated with shifted bead indices
        const int np = sys->n;
        ocm_system& S = c->sys;
        S.n = np * nrep; S.has_box = false;
        S.mass.reserve((size_t)S.n);
        for (int r = 0; r < nrep; r++) S.mass.insert(S.mass.end(), sys->mass.begin(), sys->mass.end());
        for (const auto& t : sys->bonded) {
            BondedTerm u = t;
            u.idx.clear(); u.p.clear();
            u.idx.reserve(t.idx.size() * nrep); u.p.reserve(t.p.size() * nrep);
            for (int r = 0; r < nrep; r++) {
                for (int32_t v : t.idx) u.idx.push_back(v + r * np);
                u.p.insert(u.p.end(), t.p.begin(), t.p.end());
            }
            S.bonded.push_back(std::move(u));
        }
        for (const auto& t : sys->pairs) {
            PairTerm u = t;
            u.p0.clear(); u.p1.clear(); u.p2.clear(); u.part.clear(); u.excl.clear(); u.cls.clear();
            for (int r = 0; r < nrep; r++) {
                u.p0.insert(u.p0.end(), t.p0.begin(), t.p0.end());
                u.p1.insert(u.p1.end(), t.p1.begin(), t.p1.end());
                u.p2.insert(u.p2.end(), t.p2.begin(), t.p2.end());
                u.cls.insert(u.cls.end(), t.cls.begin(), t.cls.end());
                for (int32_t v : t.part) u.part.push_back(v + r * np);
                for (auto& e : t.excl) u.excl.emplace_back(e.first + r * np, e.second + r * np);
            }
            S.pairs.push_back(std::move(u));
        }
    }
    c->n = c->sys.n; c->periodic = c->sys.has_box;
    if (!c->sys.has_box && nrep == 1) {
        c->vbox = true; c->periodic = true;
        for (int k = 0; k < 3; k++) c->sys.box[k] = 1.0e4;  // placeholder until the positions are known
    }
    c->rep_seeds.assign(seeds, seeds + nrep);
    const uint64_t seed = seeds[0];
    c->T = T; c->gamma = gamma; c->dt = dt; c->seed = seed; c->device = device;
    CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreate(&c->ev0));
    CUDA_TRY(cudaEventCreate(&c->ev1));
    for (int v = 0; v < 8; v++)
        CUDA_TRY(cudaFuncSetAttribute(build_list_variant(v & 4, v & 2, v & 1), cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024 + 16));
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_obs, cudaEventDisableTiming));
    for (int b = 0; b < 2; b++) CUDA_TRY(cudaEventCreateWithFlags(&c->ev_obs_done[b], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_in_free, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_halo_go, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_halo_done, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_rb_fork, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_rb_join, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    const int n = c->n;
    c->x.alloc(3 * (size_t)n); c->xref.alloc(3 * (size_t)n); c->xbak.alloc(3 * (size_t)n);
    c->v.alloc(n); c->vbak.alloc(n); c->posq.alloc(n); c->fbuf.alloc(3 * (size_t)n);
    c->flags.alloc(FLAG_COUNT + 4); /* + scratch words for barriers */ c->step_base.alloc(1); c->egroup.alloc((size_t)40 * nrep); c->pair_counter.alloc(1);
    c->x.zero(); c->xref.zero(); c->fbuf.zero(); c->flags.zero(); c->step_base.zero(); c->posq.zero();
    if (nrep > 1) {
        c->origin.alloc((size_t)3 * nrep); c->origin.zero();
        std::vector<uint2> ks(nrep);
        for (int r = 0; r < nrep; r++) ks[r] = make_uint2((unsigned int)seeds[r], (unsigned int)(seeds[r] >> 32));
        c->d_rep_seed.upload(ks);
    }
    std::vector<float4> v0(n);
    for (int i = 0; i < n; i++) v0[i] = make_float4(0, 0, 0, c->sys.mass[i] > 0 ? (float)(1.0 / c->sys.mass[i]) : 0.f);
    CUDA_TRY(cudaMemcpy(c->v.p, v0.data(), n * sizeof(float4), cudaMemcpyHostToDevice));
    // default skin: 8 % of the largest cutoff, within [0.1, 0.4] nm
    plan_lists(c);
    double rcmax = 0;
    for (auto& L : c->lists) { compile_list(c, *L); rcmax = std::max(rcmax, L->rc_list); }
    compile_da(c);
    rcmax = std::max(rcmax, c->da_rcmax);
    c->skin = std::min(0.4, std::max(0.1, 0.08 * rcmax));
    if (c->periodic)
        for (int k = 0; k < 3; k++)
            if (c->da_rcmax > 0 && c->da_rcmax + c->skin > 0.5 * c->sys.box[k]) c->skin = std::max(0.0, 0.5 * c->sys.box[k] - c->da_rcmax) * 0.999;
    if (c->periodic)
        for (auto& L : c->lists)
            for (int k = 0; k < 3; k++)
                if (L->rc_list + c->skin > 0.5 * c->sys.box[k]) c->skin = std::max(0.0, 0.5 * c->sys.box[k] - L->rc_list) * 0.999;
    compile_bonded(c);
    compile_coms(c);
    update_params(c);
    compile_replica_tables(c);
    c->nst = initial_nst(c);
    CUDA_TRY(cudaDeviceSynchronize());
    *out = c;
    return 0;
    }
    catch (const std::exception& e) { delete c; return ocm::fail(e.what()); }
    catch (...) { delete c; return ocm::fail("unknown error"); }
}

int ocm_context_create(const ocm_system* sys, double T, double gamma, double dt, uint64_t seed, int device, ocm_context** out) {
    return create_context(sys, 1, T, gamma, dt, &seed, device, out);
}
int ocm_context_create_batch(const ocm_system* sys, int32_t n_replicas, double T, double gamma, double dt, const uint64_t* seeds,
                             int device, ocm_context** out) {
    return create_context(sys, n_replicas, T, gamma, dt, seeds, device, out);
}
int ocm_num_replicas(const ocm_context* c) { return c ? c->nrep : -1; }

// ---- slab decomposition --------------------------------------------------------------------------
int ocm_comm_unique_id(char* id128) {
    if (!id128) return fail("null argument");
    OCM_GUARD_BEGIN
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    NcclApi& A = NcclApi::get();
    A.ok(A.GetUniqueId(&id), "ncclGetUniqueId");
    std::memcpy(id128, &id, 128);
    return 0;
    OCM_GUARD_END
}
int ocm_comm_create_nccl(const char* id128, int world, int rank, int device, ocm_comm** out) {
    if (!id128 || !out) return fail("null argument");
    if (world < 1 || rank < 0 || rank >= world) return fail("comm_create_nccl: rank/world out of range");
    OCM_GUARD_BEGIN
    int ndev = ocm_device_count();
    if (device < 0 || device >= ndev) return fail("comm_create_nccl: device index out of range");
    ncclUniqueId id;
    std::memcpy(&id, id128, 128);
    auto cm = std::make_unique<ocm_comm>();
    cm->impl = new NcclComm(id, world, rank, device);
    cm->rank = rank;
    *out = cm.release();
    return 0;
    OCM_GUARD_END
}
int ocm_comm_create_local(int world, ocm_comm** out) {
    if (!out) return fail("null argument");
    if (world < 1 || world > kMaxLocalRanks) return fail("comm_create_local: world must be in 1..16");
    OCM_GUARD_BEGIN
    auto cm = std::make_unique<ocm_comm>();
    cm->impl = new LocalComm(world);
    *out = cm.release();
    return 0;
    OCM_GUARD_END
}
void ocm_comm_destroy(ocm_comm* comm) { delete comm; }
int ocm_context_create_slab(const ocm_system* sys, double T, double gamma, double dt, uint64_t seed, int device, ocm_comm* comm,
                            int rank, int axis, ocm_context** out) {
    if (!comm) return fail("context_create_slab: null communicator");
    return create_context(sys, 1, T, gamma, dt, &seed, device, out, comm, rank, axis);
}
int ocm_slab_info(const ocm_context* c, int* rank, int* world, int* axis, double* halo_nm) {
    if (!c) return fail("null context");
    if (rank) *rank = c->dd ? c->dd->rank : 0;
    if (world) *world = c->dd ? c->dd->world : 1;
    if (axis) *axis = c->dd ? c->dd->axis : -1;
    if (halo_nm) *halo_nm = c->dd ? c->dd->halo : 0.0;
    return 0;
}
int ocm_get_roles(ocm_context* c, unsigned char* role) {
    if (!c || !role) return fail("null argument");
    OCM_GUARD_BEGIN
    CUDA_TRY(cudaSetDevice(c->device));
    if (!c->dd) { std::memset(role, 1, (size_t)c->n); return 0; }
    ensure_lists_current(c);
    CUDA_TRY(cudaMemcpy(role, c->dd->role.p, (size_t)c->n, cudaMemcpyDeviceToHost));
    return 0;
    OCM_GUARD_END
}
void ocm_context_destroy(ocm_context* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    if (c->dd && c->dd->push_ok) {
        // neighbours map this rank's push buffer: everybody unmaps before anybody frees (slab contexts are destroyed
        // collectively, like every other call on them)
        try {
            c->dd->comm->peer_unmap(c->dd->rank);
            c->dd->comm->allreduce(c->dd->rank, c->flags.p + FLAG_COUNT, 1, DD_I32, true, c->stream);
            cudaStreamSynchronize(c->stream);
        } catch (...) {
        }
    }
    delete c;
}

int ocm_context_set_option(ocm_context* c, const char* key, double value) {
    if (!c || !key) return fail("null argument");
    OCM_GUARD_BEGIN
    CUDA_TRY(cudaSetDevice(c->device));
    std::string k(key);
    if (k == "skin_nm") {
        if (!(value > 0)) return fail("skin_nm must be positive");
        if (c->periodic)
            for (auto& L : c->lists)
                for (int d = 0; d < 3; d++)
                    if (L->rc_list + value > 0.5 * c->sys.box[d]) return fail("skin_nm: cutoff + skin exceeds half the box");
        if (c->periodic && c->da_rcmax > 0)
            for (int d = 0; d < 3; d++)
                if (c->da_rcmax + value > 0.5 * c->sys.box[d]) return fail("skin_nm: cutoff + skin exceeds half the box");
        c->skin = value;
        update_params(c);
        c->lists_ready = false; c->lists_current = false;
        for (auto& L : c->lists) { L->stride = 0; L->row_len32 = 0; }
        drop_graphs(c);
        if (c->nstlist_fixed <= 0) c->nst = initial_nst(c);
    } else if (k == "pdl") {
        // 0 = plain launches, 1 = chained launches (successor released when the grid ends), 2 = released at block start
        c->pdl = value != 0;
        c->P.pdl_early = value >= 2 ? 1 : 0;
        drop_graphs(c);
    } else if (k == "phase_timing") {
        c->phase_timing = value != 0;
        for (double& v : c->ph_ms) v = 0;
    } else if (k == "halo_overlap") {
        c->halo_overlap = value != 0;
    } else if (k == "halo_push") {
        // per-step halo over peer memory (1, default where the transport can map neighbours) or NCCL send/recv (0); collective
        if (!c->dd) return fail("halo_push: not a slab context");
        c->dd->push_wanted = value != 0;
        if (c->dd->cap > 0) dd_alloc(c, c->dd->cap);
        c->lists_current = false;
    } else if (k == "halo_capacity") {
        // entries per halo message of a slab decomposition (grown automatically on overflow)
        if (!c->dd) return fail("halo_capacity: not a slab context");
        if (value < 1 || value > (1 << 24)) return fail("halo_capacity out of range");
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        dd_alloc(c, (int)value);
        c->lists_current = false;
    } else if (k == "nstlist") {
        c->nstlist_fixed = (int)value;
        if (c->nstlist_fixed > 0) c->nst = c->nstlist_fixed; else c->nst = initial_nst(c);
    } else if (k == "threads_per_bead") {
        int t = (int)value;
        if (!(t == 0 || t == 4 || t == 8 || t == 16 || t == 32)) return fail("threads_per_bead must be 0 (auto), 4, 8, 16 or 32");
        c->tpb_opt = t;
        if (c->lists_ready) { build_segments(c); upload_lists(c); c->lists_current = false; }
    } else if (k == "row_staging") {
        // (round-1 experiment: Verlet rows staged in shared memory by bulk async copies.  It never paid, and its mbarrier phase
        // bookkeeping was wrong for beads without neighbours; removed in round 2)
        if (value != 0) return fail("row_staging: the bulk-copy row staging path has been removed");
    } else if (k == "pair_kernel") {
        // 1: cluster-pair tile lists where the system allows them; 0: per-bead Verlet rows; -1 (default): tiles for the lists of
        // at least kTileAutoBeads beads (measured on B200: rows 38.6 us vs tiles 48 us per force launch at 16 300 beads, where one wave
        // of blocks is short; at 1 043 200 beads the launches tie at 2.0 ms and the tile list builds 3 x faster and is 60 x smaller)
        c->tile_opt = value > 0 ? 1 : (value < 0 ? -1 : 0);
        c->lists_ready = false; c->lists_current = false;
        drop_graphs(c);
    } else if (k == "replica_kernel") {
        // 1 (default): small unlimited-boundary systems and replica ensembles step in the whole-step persistent kernel (measured on
        // B200: 2 x the general path for one SH3 replica, 4 x for 512, see replica_kernels.cuh); 0: general path
        c->rep_opt = value != 0 ? 1 : 0;
    } else if (k == "replica_block" || k == "replica_carveout") {
        if (k == "replica_block") c->rep_block_opt = (int)value; else c->rep_carveout_opt = (int)value;
        compile_replica_tables(c);
    } else if (k == "tile_debug") {
        // timing experiments only (wrong forces): see TileLaunch::debug; 4 = evaluate the list without its switching function
        c->tile_debug = (int)value;
        if (c->lists_ready) { build_segments(c); upload_lists(c); }
    } else if (k == "tile_blocks") {
        if (value < 0 || value > 65536) return fail("tile_blocks out of range");
        c->tile_chunk_opt = (int)value;
        if (c->lists_ready) { build_segments(c); upload_lists(c); }
    } else if (k == "beads_per_block") {
        if (value < 0 || value > 65536) return fail("beads_per_block out of range");
        c->chunk_opt = (int)value;
        if (c->lists_ready) { build_segments(c); upload_lists(c); c->lists_current = false; }
    } else {
        return fail("unknown option: " + k);
    }
    return 0;
    OCM_GUARD_END
}
int ocm_context_get_option(ocm_context* c, const char* key, double* value) {
    if (!c || !key || !value) return fail("null argument");
    std::string k(key);
    if (k == "skin_nm") *value = c->skin;
    else if (k == "nstlist") *value = c->nstlist_fixed;
    else if (k == "threads_per_bead") *value = c->tpb_opt;
    else if (k == "beads_per_block") *value = c->chunk_opt;
    else if (k == "row_staging") *value = 0;
    else if (k == "halo_capacity") *value = c->dd ? c->dd->cap : 0;
    else if (k == "pair_kernel") *value = c->tile_opt;
    else if (k == "replica_kernel") *value = (c->rep_ok && c->rep_opt) ? 1 : 0;
    else if (k == "tile_blocks") *value = c->tile_chunk_opt;
    else return fail("unknown option: " + k);
    return 0;
}

int ocm_set_positions(ocm_context* c, const double* xyz) {
    if (!c || !xyz) return fail("null argument");
    OCM_GUARD_BEGIN
    CUDA_TRY(cudaSetDevice(c->device));
    if (c->vbox) vbox_layout(c, xyz, !c->positions_set);
    CUDA_TRY(cudaMemcpyAsync(c->x.p, xyz, (size_t)3 * c->n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    canon_kernel<<<nblk(c->n, 256), 256, 0, c->stream>>>(c->P, c->x.p, c->posq.p);
    // slab decomposition: every rank is handed the whole system; ownership follows from the positions
    if (c->dd) dd_role_from_position_kernel<<<nblk(c->n, 256), 256, 0, c->stream>>>(c->n, c->dd->G, c->posq.p, c->dd->role.p);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->positions_set = true; c->lists_current = false;
    return 0;
    OCM_GUARD_END
}
int ocm_set_velocities(ocm_context* c, const double* xyz) {
    if (!c || !xyz) return fail("null argument");
    OCM_GUARD_BEGIN
    CUDA_TRY(cudaSetDevice(c->device));
    std::vector<float4> v(c->n);
    for (int i = 0; i < c->n; i++) {
        double m = c->sys.mass[i];
        v[i] = make_float4((float)xyz[3 * i], (float)xyz[3 * i + 1], (float)xyz[3 * i + 2], m > 0 ? (float)(1.0 / m) : 0.f);
    }
    CUDA_TRY(cudaMemcpy(c->v.p, v.data(), c->n * sizeof(float4), cudaMemcpyHostToDevice));
    return 0;
    OCM_GUARD_END
}
int ocm_set_velocities_to_temperature(ocm_context* c, double T, uint64_t seed) {
    if (!c) return fail("null context");
    OCM_GUARD_BEGIN
    CUDA_TRY(cudaSetDevice(c->device));
    init_velocity_kernel<<<nblk(c->n, 256), 256, 0, c->stream>>>(c->n, c->nper, (float)(kBoltz * T), (unsigned long long)seed, c->v.p);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
    OCM_GUARD_END
}

int ocm_step(ocm_context* c, int64_t nsteps) {
    if (!c) return fail("null context");
    if (nsteps < 0) return fail("step: negative step count");
    OCM_GUARD_BEGIN
    CUDA_TRY(cudaSetDevice(c->device));
    if (c->baro.kind && c->baro.frequency > 0) do_steps_barostat(c, nsteps);
    else do_steps(c, nsteps);
    return 0;
    OCM_GUARD_END
}
int64_t ocm_current_step(const ocm_context* c) { return c ? c->step : -1; }

int ocm_context_set_barostat(ocm_context* c, int kind, const double* pressure3, double tension, double T, const int32_t* mode3, int32_t frequency,
                             uint64_t seed) {
    if (!c) return fail("null context");
    OCM_GUARD_BEGIN
    CUDA_TRY(cudaSetDevice(c->device));
    auto& B = c->baro;
    if (kind == OCM_BAROSTAT_NONE) { B.kind = 0; return 0; }
    if (kind != OCM_BAROSTAT_ANISOTROPIC && kind != OCM_BAROSTAT_MEMBRANE) return fail("set_barostat: unknown kind");
    if (!pressure3 || !mode3) return fail("set_barostat: null argument");
    if (!c->periodic || c->vbox) return fail("barostat should be used with periodic boundary condition.");
    if (c->dd) return fail("set_barostat: not available under slab decomposition");
    if (c->nrep > 1) return fail("set_barostat: not available for replica batches");
    if (frequency < 0) return fail("set_barostat: negative frequency");
    if (!(T > 0)) return fail("set_barostat: temperature must be positive");
    if (kind == OCM_BAROSTAT_ANISOTROPIC && !mode3[0] && !mode3[1] && !mode3[2]) frequency = 0;  // (OpenMM: nothing to scale, never a trial)
    if (kind == OCM_BAROSTAT_MEMBRANE && (mode3[0] < 0 || mode3[0] > 1 || mode3[1] < 0 || mode3[1] > 2)) return fail("set_barostat: unknown xymode / zmode");
    B.kind = kind;
    for (int k = 0; k < 3; k++) { B.pressure[k] = pressure3[kind == OCM_BAROSTAT_MEMBRANE ? 0 : k]; B.mode[k] = mode3[k]; }
    B.tension = tension; B.T = T; B.frequency = frequency; B.rng = seed;
    const double volume = c->sys.box[0] * c->sys.box[1] * c->sys.box[2];
    for (int k = 0; k < 3; k++) { B.volume_scale[k] = 0.01 * volume; B.attempted[k] = 0; B.accepted[k] = 0; }
    B.until_trial = std::max(0, frequency - 1);
    B.n_trials = 0; B.n_accepted = 0; B.last[0] = -1;
    barostat_molecules(c);
    return 0;
    OCM_GUARD_END
}

int ocm_get_box(ocm_context* c, double* box3) {
    if (!c || !box3) return fail("null argument");
    for (int k = 0; k < 3; k++) box3[k] = (c->periodic && !c->vbox) ? c->sys.box[k] : 0.0;
    return 0;
}

int ocm_set_box(ocm_context* c, const double* box3) {
    if (!c || !box3) return fail("null argument");
    OCM_GUARD_BEGIN
    CUDA_TRY(cudaSetDevice(c->device));
    if (c->dd) return fail("set_box: not available under slab decomposition");
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    apply_box(c, box3);
    return 0;
    OCM_GUARD_END
}

int ocm_get_barostat_info(ocm_context* c, double* out) {
    if (!c || !out) return fail("null argument");
    const auto& B = c->baro;
    out[0] = B.n_trials; out[1] = B.n_accepted;
    for (int k = 0; k < 6; k++) out[2 + k] = B.last[k];
    for (int k = 0; k < 3; k++) out[8 + k] = B.volume_scale[k];
    out[11] = B.nmol;
    return 0;
}

int ocm_minimize(ocm_context* c, double tolerance, int max_iterations, int* iterations, double* final_energy) {
    if (!c) return fail("null context");
    if (c->dd) return fail("minimize: not available under slab decomposition; minimise on one device first");
    if (c->nrep > 1) return fail("minimize: not available for replica batches");
    if (!(tolerance > 0)) return fail("minimize: tolerance must be positive");
    OCM_GUARD_BEGIN
    const int n = c->n;
    std::vector<double> x((size_t)3 * n), f((size_t)3 * n);
    if (ocm_get_state(c, 0xFFFFFFFFu, x.data(), nullptr, nullptr, nullptr, nullptr, nullptr)) return 1;
    bool bad = false;
    auto eval = [&](const std::vector<double>& at, std::vector<double>& grad) {
        double e = 0;
        if (ocm_set_positions(c, at.data()) || ocm_get_state(c, 0xFFFFFFFFu, nullptr, nullptr, f.data(), &e, nullptr, nullptr)) {
            bad = true;
            return std::numeric_limits<double>::infinity();
        }
        for (int i = 0; i < n; i++) {
            const bool mobile = c->sys.mass[i] > 0;
            for (int k = 0; k < 3; k++) grad[3 * i + k] = mobile ? -f[3 * i + k] : 0.0;
        }
        return e;
    };
    const ocmhost::LbfgsReport rep = ocmhost::lbfgs_minimize(x, eval, tolerance, max_iterations);
    if (bad) return 1;
    if (ocm_set_positions(c, x.data())) return 1;
    if (iterations) *iterations = rep.iterations;
    if (final_energy) *final_energy = rep.energy;
    c->counters_minimize_evals = rep.evaluations;
    return 0;
    OCM_GUARD_END
}

int ocm_get_state(ocm_context* c, uint32_t mask, double* pos, double* vel, double* force, double* epot, double* ekin, double* ebg) {
    if (!c) return fail("null context");
    OCM_GUARD_BEGIN
    CUDA_TRY(cudaSetDevice(c->device));
    if (c->dd) { get_state_dd(c, mask, pos, vel, force, epot, ekin, ebg); return 0; }
    cudaStream_t s = c->stream;
    const int n = c->n;
    if (pos) CUDA_TRY(cudaMemcpyAsync(pos, c->x.p, (size_t)3 * n * sizeof(double), cudaMemcpyDeviceToHost, s));
    std::vector<float4> hv;
    if (vel) { hv.resize(n); CUDA_TRY(cudaMemcpyAsync(hv.data(), c->v.p, n * sizeof(float4), cudaMemcpyDeviceToHost, s)); }
    std::vector<unsigned long long> hf;
    std::vector<double> he((size_t)40 * c->nrep, 0.0);
    if (force || epot || ekin || ebg) {
        ensure_lists_current(c);
        if (force && mask != 0xFFFFFFFFu) {
            // forces restricted to a group mask: only at the granularity of whole force segments
            for (auto& L : c->lists) {
                int in = 0, tot = 0;
                for (int t : L->terms) { tot++; in += (mask >> c->sys.pairs[t].group) & 1; }
                if (in != 0 && in != tot) return fail("get_state: the force-group mask splits a fused pair list");
            }
            return fail("get_state: forces restricted to a force-group mask are not supported; request all groups");
        }
        CUDA_TRY(cudaMemsetAsync(c->egroup.p, 0, he.size() * sizeof(double), s));
        if (epot || ebg) enqueue_energies(c, s);
        if (force || ekin) {
            compute_forces_into_fbuf(c);
            if (ekin) kinetic_kernel<<<nblk(n, kBlock), kBlock, 0, s>>>(c->P, c->v.p, c->fbuf.p, c->egroup.p + 32);
            if (force) { hf.resize((size_t)3 * n); CUDA_TRY(cudaMemcpyAsync(hf.data(), c->fbuf.p, hf.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s)); }
            CUDA_TRY(cudaMemsetAsync(c->fbuf.p, 0, (size_t)3 * n * sizeof(unsigned long long), s));
        }
        CUDA_TRY(cudaMemcpyAsync(he.data(), c->egroup.p, he.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
    }
    CUDA_TRY(cudaStreamSynchronize(s));
    if (vel) for (int i = 0; i < n; i++) { vel[3 * i] = hv[i].x; vel[3 * i + 1] = hv[i].y; vel[3 * i + 2] = hv[i].z; }
    if (force) for (int i = 0; i < n; i++)
        for (int k = 0; k < 3; k++) force[3 * i + k] = (double)(long long)hf[(size_t)k * n + i] / kFixScale;
    // a replica batch reports sums over its replicas here; ocm_get_replica_energies has them one by one
    for (int r = 1; r < c->nrep; r++)
        for (int q = 0; q < 40; q++) he[q] += he[(size_t)40 * r + q];
    if (epot) { double e = 0; for (int g = 0; g < 32; g++) if (mask >> g & 1) e += he[g]; *epot = e; }
    if (ebg) std::copy(he.begin(), he.begin() + 32, ebg);
    if (ekin) *ekin = he[32];
    return 0;
    OCM_GUARD_END
}

// ---- observer without stalls ------------------------------------------------------------------------------
struct ocm_state_request {
    ocm_context* c = nullptr;
    uint32_t mask = 0xFFFFFFFFu;
    bool pos = false, vel = false, force = false, energy = false;
    bool deferred = false;   // replica batches / slab contexts: served synchronously in ocm_get_state_end
    int buf = 0;             // which of the two staging blocks
    size_t off_x = 0, off_v = 0, off_f = 0, off_e = 0;
};

static void ensure_pinned(double*& p, size_t& have, size_t want) {
    if (have >= want) return;
    if (p) cudaFreeHost(p);
    p = nullptr; have = 0;
    CUDA_TRY(cudaHostAlloc((void**)&p, want * sizeof(double), cudaHostAllocDefault));
    have = want;
}

int ocm_get_state_begin(ocm_context* c, uint32_t mask, int want_pos, int want_vel, int want_force, int want_energy, ocm_state_request** out) {
    if (!c || !out) return fail("null argument");
    OCM_GUARD_BEGIN
    CUDA_TRY(cudaSetDevice(c->device));
    auto rq = std::make_unique<ocm_state_request>();
    rq->c = c; rq->mask = mask; rq->pos = want_pos != 0; rq->vel = want_vel != 0; rq->force = want_force != 0; rq->energy = want_energy != 0;
    if (c->dd || c->nrep > 1) { rq->deferred = true; *out = rq.release(); return 0; }
    const int b = c->obs_next;
    if (c->obs_busy[b]) return fail("get_state_begin: two requests are already open; end one first");
    rq->buf = b;
    if (rq->force && mask != 0xFFFFFFFFu) return fail("get_state: forces restricted to a force-group mask are not supported; request all groups");
    cudaStream_t s = c->stream;
    const int n = c->n;
    const size_t n3 = (size_t)3 * n;
    size_t off = 0;
    if (rq->pos) { rq->off_x = off; off += n3; }
    if (rq->vel) { rq->off_v = off; off += n3; }
    if (rq->force) { rq->off_f = off; off += n3; }
    rq->off_e = off; off += 40;
    // (block b was last downloaded two requests ago: make sure that copy has left it before it is overwritten)
    if (c->obs_used[b]) CUDA_TRY(cudaStreamWaitEvent(s, c->ev_obs_done[b], 0));
    if (c->obs_dev[b].n < off) { CUDA_TRY(cudaStreamSynchronize(c->side)); c->obs_dev[b].alloc(off); }
    ensure_pinned(c->obs_pin[b], c->obs_pin_n[b], off);
    double* D = c->obs_dev[b].p;
    if (rq->force || rq->energy) {
        if (c->lists_current) c->n_obs_rebuilds_skipped += 1;
        ensure_lists_current(c);
        CUDA_TRY(cudaMemsetAsync(c->egroup.p, 0, 40 * sizeof(double), s));
        if (rq->energy) enqueue_energies(c, s);
        compute_forces_into_fbuf(c);   // (also what the kinetic energy needs: velocities are half a step behind)
        if (rq->energy) kinetic_kernel<<<nblk(n, kBlock), kBlock, 0, s>>>(c->P, c->v.p, c->fbuf.p, c->egroup.p + 32);
        CUDA_TRY(cudaMemcpyAsync(D + rq->off_e, c->egroup.p, 40 * sizeof(double), cudaMemcpyDeviceToDevice, s));
    }
    if (rq->pos || rq->vel || rq->force)
        pack_state_kernel<<<nblk(n, 256), 256, 0, s>>>(n, c->x.p, c->v.p, c->fbuf.p, rq->pos ? D + rq->off_x : nullptr, rq->vel ? D + rq->off_v : nullptr,
                                                        rq->force ? D + rq->off_f : nullptr);
    if (rq->force || rq->energy) CUDA_TRY(cudaMemsetAsync(c->fbuf.p, 0, n3 * sizeof(unsigned long long), s));
    // the download runs on the side stream: the compute stream is free for the next window
    CUDA_TRY(cudaEventRecord(c->ev_obs, s));
    CUDA_TRY(cudaStreamWaitEvent(c->side, c->ev_obs, 0));
    CUDA_TRY(cudaMemcpyAsync(c->obs_pin[b], D, off * sizeof(double), cudaMemcpyDeviceToHost, c->side));
    CUDA_TRY(cudaEventRecord(c->ev_obs_done[b], c->side));
    c->obs_busy[b] = true; c->obs_used[b] = true; c->obs_next = b ^ 1;
    c->n_async_states += 1;
    *out = rq.release();
    return 0;
    OCM_GUARD_END
}

int ocm_get_state_end(ocm_state_request* rq, double* pos, double* vel, double* force, double* epot, double* ekin, double* ebg) {
    if (!rq) return fail("null request");
    std::unique_ptr<ocm_state_request> own(rq);
    ocm_context* c = rq->c;
    if (rq->deferred)
        return ocm_get_state(c, rq->mask, rq->pos ? pos : nullptr, rq->vel ? vel : nullptr, rq->force ? force : nullptr, rq->energy ? epot : nullptr,
                             rq->energy ? ekin : nullptr, rq->energy ? ebg : nullptr);
    OCM_GUARD_BEGIN
    CUDA_TRY(cudaSetDevice(c->device));
    const int b = rq->buf;
    c->obs_busy[b] = false;
    CUDA_TRY(cudaEventSynchronize(c->ev_obs_done[b]));
    const size_t nb = (size_t)3 * c->n * sizeof(double);
    const double* H = c->obs_pin[b];
    if (pos && rq->pos) std::memcpy(pos, H + rq->off_x, nb);
    if (vel && rq->vel) std::memcpy(vel, H + rq->off_v, nb);
    if (force && rq->force) std::memcpy(force, H + rq->off_f, nb);
    if (rq->energy) {
        const double* he = H + rq->off_e;
        if (epot) { double e = 0; for (int g = 0; g < 32; g++) if (rq->mask >> g & 1) e += he[g]; *epot = e; }
        if (ebg) std::copy(he, he + 32, ebg);
        if (ekin) *ekin = he[32];
    }
    return 0;
    OCM_GUARD_END
}

int ocm_set_state_async(ocm_context* c, const double* xyz, const double* vel) {
    if (!c || !xyz) return fail("null argument");
    OCM_GUARD_BEGIN
    CUDA_TRY(cudaSetDevice(c->device));
    // (a virtual box may have to be laid out again around the new positions, slab contexts assign roles: blocking forms)
    if (c->vbox || c->dd) {
        if (ocm_set_positions(c, xyz)) return 1;
        return vel ? ocm_set_velocities(c, vel) : 0;
    }
    cudaStream_t s = c->stream;
    const int n = c->n;
    const size_t n3 = (size_t)3 * n, tot = vel ? 2 * n3 : n3;
    if (c->in_pending) CUDA_TRY(cudaEventSynchronize(c->ev_in_free));   // the previous upload has left the pinned block
    ensure_pinned(c->in_pin, c->in_pin_n, tot);
    if (c->in_dev.n < tot) c->in_dev.alloc(tot);
    std::memcpy(c->in_pin, xyz, n3 * sizeof(double));
    if (vel) std::memcpy(c->in_pin + n3, vel, n3 * sizeof(double));
    CUDA_TRY(cudaMemcpyAsync(c->in_dev.p, c->in_pin, tot * sizeof(double), cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaEventRecord(c->ev_in_free, s));
    c->in_pending = true;
    unpack_state_kernel<<<nblk(n, 256), 256, 0, s>>>(n, c->in_dev.p, vel ? c->in_dev.p + n3 : nullptr, c->x.p, c->v.p);
    canon_kernel<<<nblk(n, 256), 256, 0, s>>>(c->P, c->x.p, c->posq.p);
    c->positions_set = true; c->lists_current = false;
    return 0;
    OCM_GUARD_END
}

int ocm_get_replica_energies(ocm_context* c, uint32_t mask, double* epot, double* ekin, double* ebg) {
    if (!c) return fail("null context");
    OCM_GUARD_BEGIN
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const int n = c->n;
    std::vector<double> he((size_t)40 * c->nrep, 0.0);
    ensure_lists_current(c);
    CUDA_TRY(cudaMemsetAsync(c->egroup.p, 0, he.size() * sizeof(double), s));
    if (epot || ebg) enqueue_energies(c, s);
    if (ekin) {
        compute_forces_into_fbuf(c);
        kinetic_kernel<<<nblk(n, kBlock), kBlock, 0, s>>>(c->P, c->v.p, c->fbuf.p, c->egroup.p + 32);
        CUDA_TRY(cudaMemsetAsync(c->fbuf.p, 0, (size_t)3 * n * sizeof(unsigned long long), s));
    }
    CUDA_TRY(cudaMemcpyAsync(he.data(), c->egroup.p, he.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    for (int r = 0; r < c->nrep; r++) {
        const double* h = &he[(size_t)40 * r];
        if (epot) { double e = 0; for (int g = 0; g < 32; g++) if (mask >> g & 1) e += h[g]; epot[r] = e; }
        if (ebg) std::copy(h, h + 32, ebg + (size_t)32 * r);
        if (ekin) ekin[r] = h[32];
    }
    return 0;
    OCM_GUARD_END
}

int ocm_num_pair_terms(const ocm_context* c) { return c ? (int)c->sys.pairs.size() : -1; }

int ocm_get_neighbor_pairs(ocm_context* c, int term, int32_t* ij, int64_t capacity, int64_t* n_pairs) {
    if (!c || !n_pairs) return fail("null argument");
    if (term < 0 || term >= (int)c->sys.pairs.size()) return fail("pair term index out of range");
    OCM_GUARD_BEGIN
    CUDA_TRY(cudaSetDevice(c->device));
    ensure_lists_current(c);
    cudaStream_t s = c->stream;
    const int li = c->term_list[term];
    PairListHost& L = *c->lists[li];
    const double rce = effective_rc(c->sys.pairs[term]);
    const float rc2 = (float)(rce * rce);
    DevBuf<int> out;
    if (ij && capacity > 0) out.alloc((size_t)2 * capacity);
    CUDA_TRY(cudaMemsetAsync(c->pair_counter.p, 0, sizeof(unsigned long long), s));
    if (L.np > 0) {
        if (L.tiled) tile_extract_pairs_kernel<true><<<L.dev.row_capacity / 64, 128, 0, s>>>(c->d_lists.p, li, c->P, rc2, out.p, capacity, c->pair_counter.p);
        else if (c->periodic) extract_pairs_kernel<true><<<nblk(L.np, 128), 128, 0, s>>>(c->d_lists.p, li, c->P, rc2, out.p, capacity, c->pair_counter.p);
        else extract_pairs_kernel<false><<<nblk(L.np, 128), 128, 0, s>>>(c->d_lists.p, li, c->P, rc2, out.p, capacity, c->pair_counter.p);
    }
    unsigned long long cnt = 0;
    CUDA_TRY(cudaMemcpyAsync(&cnt, c->pair_counter.p, sizeof(cnt), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    *n_pairs = (int64_t)cnt;
    if (ij) {
        if ((int64_t)cnt > capacity) return fail("get_neighbor_pairs: capacity too small");
        std::vector<std::pair<int, int>> pr(cnt);
        std::vector<int> raw((size_t)2 * cnt);
        if (cnt) CUDA_TRY(cudaMemcpy(raw.data(), out.p, raw.size() * sizeof(int), cudaMemcpyDeviceToHost));
        for (size_t k = 0; k < cnt; k++) pr[k] = {raw[2 * k], raw[2 * k + 1]};
        std::sort(pr.begin(), pr.end());
        for (size_t k = 0; k < cnt; k++) { ij[2 * k] = pr[k].first; ij[2 * k + 1] = pr[k].second; }
    }
    return 0;
    OCM_GUARD_END
}

int ocm_num_da_terms(const ocm_context* c) { return c ? (int)c->sys.das.size() : -1; }

int ocm_get_da_pairs(ocm_context* c, int term, int32_t* da, int64_t capacity, int64_t* n_pairs) {
    if (!c || !n_pairs) return fail("null argument");
    if (term < 0 || term >= (int)c->sys.das.size()) return fail("donor-acceptor term index out of range");
    OCM_GUARD_BEGIN
    CUDA_TRY(cudaSetDevice(c->device));
    ensure_lists_current(c);
    cudaStream_t s = c->stream;
    DevBuf<int> out;
    if (da && capacity > 0) out.alloc((size_t)2 * capacity);
    CUDA_TRY(cudaMemsetAsync(c->pair_counter.p, 0, sizeof(unsigned long long), s));
    if (c->da.capacity > 0)
        da_extract_kernel<<<nblk(c->da.capacity, 128), 128, 0, s>>>(c->da, term, c->P, c->x.p, out.p, capacity, c->pair_counter.p);
    unsigned long long cnt = 0;
    CUDA_TRY(cudaMemcpyAsync(&cnt, c->pair_counter.p, sizeof(cnt), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    *n_pairs = (int64_t)cnt;
    if (da) {
        if ((int64_t)cnt > capacity) return fail("get_da_pairs: capacity too small");
        std::vector<int> raw((size_t)2 * cnt);
        if (cnt) CUDA_TRY(cudaMemcpy(raw.data(), out.p, raw.size() * sizeof(int), cudaMemcpyDeviceToHost));
        std::vector<std::pair<int, int>> pr(cnt);
        for (size_t k = 0; k < cnt; k++) pr[k] = {raw[2 * k], raw[2 * k + 1]};
        std::sort(pr.begin(), pr.end());
        for (size_t k = 0; k < cnt; k++) { da[2 * k] = pr[k].first; da[2 * k + 1] = pr[k].second; }
    }
    return 0;
    OCM_GUARD_END
}

int ocm_get_cell_ids(ocm_context* c, int term, int32_t* cell_of_bead, int32_t dims[3]) {
    if (!c || !cell_of_bead || !dims) return fail("null argument");
    if (term < 0 || term >= (int)c->sys.pairs.size()) return fail("pair term index out of range");
    OCM_GUARD_BEGIN
    CUDA_TRY(cudaSetDevice(c->device));
    c->lists_current = false;  // (after a run the lists of the last window may still be in use: cell ids are those of a fresh build)
    ensure_lists_current(c);
    PairListHost& L = *c->lists[c->term_list[term]];
    std::vector<int> cs(std::max(L.np, 1));
    GridDev g;
    CUDA_TRY(cudaMemcpy(cs.data(), L.cell_of_slot.p, cs.size() * sizeof(int), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(&g, L.grid.p, sizeof(g), cudaMemcpyDeviceToHost));
    for (int i = 0; i < c->n; i++) cell_of_bead[i] = -1;
    const auto& part = c->sys.pairs[L.terms[0]].part;
    // (tile lists key their cells by Hilbert rank: report the linear cell id)
    for (int a = 0; a < L.np; a++) cell_of_bead[part[a]] = (L.tiled && cs[a] >= 0) ? L.cell_inv[cs[a]] : cs[a];
    for (int k = 0; k < 3; k++) dims[k] = g.nc[k];
    return 0;
    OCM_GUARD_END
}

int ocm_get_counter(ocm_context* c, const char* key, double* value) {
    if (!c || !key || !value) return fail("null argument");
    OCM_GUARD_BEGIN
    std::string k(key);
    if (k == "rebuilds") *value = c->n_rebuilds;
    else if (k == "minimize_evaluations") *value = c->counters_minimize_evals;
    else if (k == "phase_ms.rebuild") *value = c->ph_ms[0];
    else if (k == "phase_ms.force") *value = c->ph_ms[1];
    else if (k == "phase_ms.integrate") *value = c->ph_ms[2];
    else if (k == "phase_ms.halo") *value = c->ph_ms[3];
    else if (k == "rollbacks") *value = c->n_rollbacks;
    else if (k == "replica_kernel_launches") *value = c->n_rep_launches;
    else if (k == "async_states") *value = c->n_async_states;
    else if (k == "observer_rebuilds_skipped") *value = c->n_obs_rebuilds_skipped;
    else if (k == "kernel_launches") *value = c->n_launches;
    else if (k == "nstlist") *value = c->nst;
    else if (k == "last_step_ms") *value = c->last_step_ms;
    else if (k == "num_pair_lists") *value = (double)c->lists.size();
    else if (k == "halo_exchanges") *value = c->dd ? c->dd->n_exchanges : 0.0;
    else if (k == "halo_push") *value = (c->dd && c->dd->push_ok) ? 1.0 : 0.0;
    else if (k == "virtual_box_x") *value = c->vbox ? c->sys.box[0] : 0.0;
    else if (k == "virtual_box_y") *value = c->vbox ? c->sys.box[1] : 0.0;
    else if (k == "virtual_box_z") *value = c->vbox ? c->sys.box[2] : 0.0;
    else if (k == "virtual_box_layouts") *value = c->n_vbox_resize;
    else if (k == "halo_message_bytes") *value = c->dd ? (double)c->dd->bytes_s() : 0.0;
    else if (k == "owned_beads") {
        if (!c->dd) *value = c->n;
        else {
            CUDA_TRY(cudaSetDevice(c->device));
            std::vector<unsigned char> r(c->n);
            CUDA_TRY(cudaMemcpy(r.data(), c->dd->role.p, r.size(), cudaMemcpyDeviceToHost));
            double cnt = 0; for (unsigned char q : r) cnt += q == 1;
            *value = cnt;
        }
    }
    else if (k == "max_neighbors") { int m = 0; for (auto& L : c->lists) m = std::max(m, L->stride); *value = m; }
    else if (k == "pair_evals_per_step") {
        CUDA_TRY(cudaSetDevice(c->device));
        ensure_lists_current(c);
        double tot = 0;
        for (size_t l = 0; l < c->lists.size(); l++) {
            auto& L = c->lists[l];
            if (!L->np) continue;
            if (L->tiled) {
                // lane pairs the tile kernel evaluates: 32 per (i-cluster, j-group) tile
                unsigned long long nt = 0;
                CUDA_TRY(cudaMemsetAsync(c->pair_counter.p, 0, sizeof(nt), c->stream));
                tile_count_kernel<<<L->dev.row_capacity / 64, 128, 0, c->stream>>>(c->d_lists.p, (int)l, c->pair_counter.p);
                CUDA_TRY(cudaMemcpyAsync(&nt, c->pair_counter.p, sizeof(nt), cudaMemcpyDeviceToHost, c->stream));
                CUDA_TRY(cudaStreamSynchronize(c->stream));
                tot += 32.0 * (double)nt;
                continue;
            }
            int rows = L->np;
            if (c->dd) CUDA_TRY(cudaMemcpy(&rows, L->cell_start.p + L->max_cells, sizeof(int), cudaMemcpyDeviceToHost));
            std::vector<int> cnt(rows);
            if (rows) CUDA_TRY(cudaMemcpy(cnt.data(), L->nbr_count.p, cnt.size() * sizeof(int), cudaMemcpyDeviceToHost));
            for (int v : cnt) tot += v;
        }
        *value = tot;
    } else if (k.rfind("cell_width_nm.", 0) == 0 || k.rfind("cell_cap.", 0) == 0 || k.rfind("tiled.", 0) == 0) {
        // the grid rule of the list serving a pair term (the oracle restates it: oracle_get_cell_ids_grid)
        int term = std::atoi(k.c_str() + k.find('.') + 1);
        if (term < 0 || term >= (int)c->sys.pairs.size()) return fail("pair term index out of range");
        // (not collective: no list build here.  The grid exists once the lists have been sized, i.e. after the first force,
        // energy or step call)
        const PairListHost& L = *c->lists[c->term_list[term]];
        if (k[0] == 't') *value = (L.tile_ok && (c->tile_opt > 0 || (c->tile_opt < 0 && L.tile_auto))) ? 1.0 : 0.0;
        else {
            if (!c->lists_ready) return fail("the cell grid is laid out with the first list build: evaluate forces or energies first");
            *value = k[5] == 'w' ? (double)L.cell_width : (double)L.cell_cap;
        }
    } else if (k.rfind("list_cutoff_nm.", 0) == 0) {
        int term = std::atoi(k.c_str() + 15);
        if (term < 0 || term >= (int)c->sys.pairs.size()) return fail("pair term index out of range");
        *value = c->lists[c->term_list[term]]->rc_list + c->skin;
    } else if (k.rfind("threads_per_bead.", 0) == 0) {
        int term = std::atoi(k.c_str() + 17);
        if (term < 0 || term >= (int)c->sys.pairs.size()) return fail("pair term index out of range");
        *value = c->lists[c->term_list[term]]->tpb;
    } else return fail("unknown counter: " + k);
    return 0;
    OCM_GUARD_END
}

int ocm_profile_kernel(ocm_context* c, const char* which, int reps, double* avg_ms) {
    if (!c || !which || !avg_ms || reps <= 0) return fail("profile_kernel: bad arguments");
    OCM_GUARD_BEGIN
    CUDA_TRY(cudaSetDevice(c->device));
    ensure_lists_current(c);
    cudaStream_t s = c->stream;
    std::string w(which);
    const size_t fb = (size_t)3 * c->n * sizeof(unsigned long long);
    double launches = c->n_launches, rebuilds = c->n_rebuilds;
    auto run = [&](int count) {
        for (int r = 0; r < count; r++) {
            if (w == "force") enqueue_forces(c, s);
            else if (w == "rebuild") enqueue_rebuild(c, s, false);
            else if (w == "integrate") {
                // dt = 0 copy of the parameters would change the physics; instead integrate a scratch copy
                (c->n < kIntegrateEarlyBeads ? integrate_kernel<true> : integrate_kernel<false>)<<<nblk(c->n, 128), 128, 0, s>>>(c->P, c->xbak.p, c->vbak.p, c->fbuf.p, c->posq.p, c->d_lists.p, (int)c->lists.size(), c->xref.p,
                                                                 c->step_base.p, 0, 1, 0, c->flags.p);
            } else throw std::runtime_error("profile_kernel: unknown kernel " + w);
        }
    };
    if (w == "integrate") {
        CUDA_TRY(cudaMemcpyAsync(c->xbak.p, c->x.p, (size_t)3 * c->n * sizeof(double), cudaMemcpyDeviceToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(c->vbak.p, c->v.p, (size_t)c->n * sizeof(float4), cudaMemcpyDeviceToDevice, s));
    }
    run(3);
    CUDA_TRY(cudaEventRecord(c->ev0, s));
    run(reps);
    CUDA_TRY(cudaEventRecord(c->ev1, s));
    CUDA_TRY(cudaEventSynchronize(c->ev1));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    *avg_ms = ms / reps;
    // leave the state as it was: clear the accumulator and flags, restore the pair-kernel positions
    CUDA_TRY(cudaMemsetAsync(c->fbuf.p, 0, fb, s));
    CUDA_TRY(cudaMemsetAsync(c->flags.p, 0, FLAG_COUNT * sizeof(int), s));
    canon_kernel<<<nblk(c->n, 256), 256, 0, s>>>(c->P, c->x.p, c->posq.p);
    CUDA_TRY(cudaStreamSynchronize(s));
    c->n_launches = launches; c->n_rebuilds = rebuilds;
    c->lists_current = false;
    return 0;
    OCM_GUARD_END
}

int ocm_measure_fp32_peak(int device, double* tflops) {
    if (!tflops) return fail("null argument");
    OCM_GUARD_BEGIN
    int ndev = ocm_device_count();
    if (device < 0 || device >= ndev) return fail("measure_fp32_peak: device index out of range");
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    DevBuf<float> out;
    out.alloc((size_t)blocks * threads);
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1));
    double best = 0;
    for (int rep = 0; rep < 6; rep++) {
        CUDA_TRY(cudaEventRecord(e0));
        fma_peak_kernel<<<blocks, threads>>>(out.p, iters, 0.999f, 1e-3f);
        CUDA_TRY(cudaEventRecord(e1));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        double flop = (double)blocks * threads * iters * 16.0 * 8.0 * 2.0;
        if (rep > 0) best = std::max(best, flop / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *tflops = best;
    return 0;
    OCM_GUARD_END
}

}  // extern "C"
