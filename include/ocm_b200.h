/*
 * ocm_b200.h — C ABI of the B200-native coarse-grained MD engine.
 *
 * This is the drop-in boundary.  In the reference (yutakasi634/OpenCafeMol) the
 * per-step hot path sits behind the OpenMM C++ API: every
 * `*ForceFieldGenerator::generate()` emits an `OpenMM::Force`, `SystemGenerator::generate()`
 * an `OpenMM::System`, `Simulator` owns an `OpenMM::Context` and calls
 * `integrator->step(save_step)` / `context.getState(...)`.  Each entry point below names the
 * reference interface it replaces (file:line under the reference's `src/`).
 *
 * Conventions
 *   - units are OpenMM's: nm, ps, amu, kJ/mol, rad, K, elementary charge;
 *   - every function returns 0 on success, non-zero on failure; the message is then readable
 *     with ocm_last_error() (thread-local);
 *   - handles are opaque, owned by the library, released with the matching *_destroy;
 *   - all arrays are caller-owned host memory, copied during the call;
 *   - `group` is the force-group id (0..31) whose energy getState reports separately
 *     (SystemGenerator.cpp:4-32, Observer.cpp:324-346);
 *   - `periodic` != 0 applies the orthorhombic minimum image (setUsesPeriodicBoundaryConditions /
 *     CutoffPeriodic); it requires a box at ocm_system_create.
 *   - index arrays are int32, flattened row-major (ij: 2 per term, ijk: 3, ijkl: 4).
 */
#ifndef OCM_B200_H
#define OCM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ocm_system  ocm_system;
typedef struct ocm_context ocm_context;

/* Exclusions and interaction groups of one pair ("CustomNonbondedForce"-like) term.
 * Replaces CustomNonbondedForce::addExclusion / addInteractionGroup as used by e.g.
 * forcefield/ExcludedVolumeForceFieldGenerator.cpp:62-90 and forcefield/InteractionGroup.hpp:19-101.
 * n_igroups == 0 means "every bead with a parameter interacts with every other".
 * Group g couples set1 (members set1[set1_off[g] .. set1_off[g+1])) with set2 likewise; a pair
 * interacts if one bead is in set1 and the other in set2 of some group. */
typedef struct ocm_pair_topology {
    int64_t        n_excl;     /* number of excluded pairs                              */
    const int32_t* excl_ij;    /* 2*n_excl bead indices                                 */
    int32_t        n_igroups;  /* number of interaction groups (0 = none)               */
    const int64_t* set1_off;   /* n_igroups+1 offsets into set1                         */
    const int32_t* set1;
    const int64_t* set2_off;   /* n_igroups+1 offsets into set2                         */
    const int32_t* set2;
    int32_t        periodic;   /* CutoffPeriodic (1) or CutoffNonPeriodic (0)           */
} ocm_pair_topology;

/* ---- library ------------------------------------------------------------------------------ */
const char* ocm_last_error(void);
const char* ocm_version(void);
/* number of visible CUDA devices (0 when there is none); never fails */
int ocm_device_count(void);

/* ---- system (replaces OpenMM::System as built by SystemGenerator.cpp:34-111) -------------- */
/* mass_amu[n]; box_nm = {Lx,Ly,Lz} or NULL for an unlimited boundary */
int  ocm_system_create(int32_t n, const double* mass_amu, const double* box_nm, ocm_system** out);
void ocm_system_destroy(ocm_system* sys);
int  ocm_system_num_particles(const ocm_system* sys);

/* ---- bonded terms -------------------------------------------------------------------------- */
/* E = 1/2 k (r-r0)^2                      HarmonicBondForceFieldGenerator.cpp:3-14 */
int ocm_system_add_harmonic_bond(ocm_system*, int group, int periodic, int64_t n,
                                 const int32_t* ij, const double* r0, const double* k);
/* E = k exp(-(r-r0)^2/(2 sigma^2))        GaussianBondForceFieldGenerator.cpp:3-23 */
int ocm_system_add_gaussian_bond(ocm_system*, int group, int periodic, int64_t n,
                                 const int32_t* ij, const double* k, const double* r0,
                                 const double* sigma);
/* E = k (5 (r0/r)^12 - 6 (r0/r)^10)       GoContactForceFieldGenerator.cpp:3-20 */
int ocm_system_add_go_contact(ocm_system*, int group, int periodic, int64_t n,
                              const int32_t* ij, const double* k, const double* r0);
/* Go contact with a linear cap at short range: for r/r0 <= cap  E = k (slope r/r0 + intercept), else as above
 * (slope, intercept: the tangent of the 12-10 curve at r = cap r0, as the generator computes them)
 *                                         CappedGoContactForceFieldGenerator.cpp:3-30 */
int ocm_system_add_capped_go_contact(ocm_system*, int group, int periodic, int64_t n, const int32_t* ij,
                                     const double* k, const double* r0, const double* slope,
                                     const double* intercept, const double* cap);
/* E = k2 (r-r0)^2 + k4 (r-r0)^4           ThreeSPN2BondForceFieldGenerator.cpp:3-25
 * (the reference passes k4 = 100*k_4th; give the final coefficient here) */
int ocm_system_add_3spn2_bond(ocm_system*, int group, int periodic, int64_t n,
                              const int32_t* ij, const double* k2, const double* k4,
                              const double* r0);
/* E = 1/2 k (theta-theta0)^2              HarmonicAngleForceFieldGenerator.cpp:3-15 */
int ocm_system_add_harmonic_angle(ocm_system*, int group, int periodic, int64_t n,
                                  const int32_t* ijk, const double* theta0, const double* k);
/* E = k (spline(theta) - e_min + walls)   FlexibleLocalAngleForceFieldGenerator.cpp:3-45
 * y[10]: spline knots on [theta_min, theta_max] (natural cubic spline, 0 outside the range,
 * OpenMM Continuous1DFunction); linear walls of slope 30 outside the range. */
int ocm_system_add_flexible_local_angle(ocm_system*, int group, int periodic, int64_t n,
                                        const int32_t* ijk, const double* k, const double* y10,
                                        double theta_min, double theta_max, double e_min);
/* E = k exp(-d^2/(2 sigma^2)), d = wrap(phi-phi0)   GaussianDihedralForceFieldGenerator.cpp:3-28 */
int ocm_system_add_gaussian_dihedral(ocm_system*, int group, int periodic, int64_t n,
                                     const int32_t* ijkl, const double* k, const double* phi0,
                                     const double* sigma);
/* E = k (1 - cos(nper (phi - t0)))        CosineDihedralForceFieldGenerator.cpp:5-43 */
int ocm_system_add_cosine_dihedral(ocm_system*, int group, int periodic, int64_t n,
                                   const int32_t* ijkl, const double* k, const double* t0,
                                   const int32_t* nper);
/* E = k (c - e_min + sum_m a_m cos m phi + b_m sin m phi), coef7 = {c,a1,b1,a2,b2,a3,b3}
 *                                         FlexibleLocalDihedralForceFieldGenerator.cpp:5-63 */
int ocm_system_add_flexible_local_dihedral(ocm_system*, int group, int periodic, int64_t n,
                                           const int32_t* ijkl, const double* k,
                                           const double* coef7, double e_min);

/* 3SPN.2 base stacking on (p1 = sugar 5', p2 = base 5', p3 = base 3'):
 *   E = rep + f(K_BS (angle(p1,p2,p3) - theta0)) attr,  dr = |p2-p3| - r0,
 *   rep = eps (1-exp(-alpha dr))^2 [dr <= 0],  attr = eps (1-exp(-alpha dr))^2 [dr >= 0] - eps,
 *   f(x) = 1 (|x| <= pi/2), sin^2 x (|x| <= pi), 0 otherwise
 *                                         ThreeSPN2BaseStackingForceFieldGenerator.cpp:3-50 */
int ocm_system_add_3spn2_base_stacking(ocm_system*, int group, int periodic, int64_t n,
                                       const int32_t* ijk, const double* eps, const double* r0,
                                       const double* theta0, double alpha, double K_BS);

/* ---- donor x acceptor terms ("CustomHbondForce"-like) ------------------------
 * donors: 3 bead indices (d1,d2,d3) per donor, acceptors: 3 (a1,a2,a3) per acceptor, -1 = unused.
 * Every donor x acceptor combination is evaluated iff it is not excluded and |d1 - a1| < cutoff
 * (no switching; the energy is discontinuous at the cutoff, as in OpenMM).  Exclusions are
 * (donor index, acceptor index) pairs, as CustomHbondForce::addExclusion takes them. */
typedef struct ocm_da_topology {
    int64_t        n_donors;
    const int32_t* donors;      /* 3*n_donors    */
    int64_t        n_acceptors;
    const int32_t* acceptors;   /* 3*n_acceptors */
    int64_t        n_excl;
    const int32_t* excl_da;     /* 2*n_excl: donor index, acceptor index */
    double         cutoff;      /* nm */
    int32_t        periodic;
} ocm_da_topology;

/* 3SPN.2 base pairing, donor (d1 = base, d2 = sugar), acceptor (a1 = base, a2 = sugar):
 *   E = rep + 1/2 (1 + cos(dihedral(d2,d1,a1,a2) - phi0)) f(K (angle(d2,d1,a1) - theta0_1))
 *             f(K (angle(a2,a1,d1) - theta0_2)) attr,   dr = |d1-a1| - r0
 * per-donor parameter arrays of length n_donors    ThreeSPN2BasePairForceFieldGenerator.cpp:5-100 */
int ocm_system_add_3spn2_base_pair(ocm_system*, int group, const ocm_da_topology* topo,
                                   const double* eps, const double* r0, const double* theta0_1,
                                   const double* theta0_2, const double* phi0, double alpha,
                                   double K_BP);
/* 3SPN.2 cross stacking, donor (d1 = base, d2 = sugar, d3 unused by the expression), acceptor
 * (a1 = paired base, a2 = sugar, a3 = cross base):
 *   E = f(K_BP (t3 - theta0_3)) f(K_CS (angle(d2,d1,a3) - theta0_CS)) attr / 2,  dr = |d1-a3| - r0,
 *   t3 = acos(clamp(sin t1 sin t2 cos phi - cos t1 cos t2, +-0.99)), t1 = angle(d2,d1,a1),
 *   t2 = angle(d1,a1,a2), phi = dihedral(d2,d1,a1,a2)
 * per-acceptor parameter arrays of length n_acceptors
 *                                         ThreeSPN2CrossStackingForceFieldGenerator.cpp:5-104 */
int ocm_system_add_3spn2_cross_stacking(ocm_system*, int group, const ocm_da_topology* topo,
                                        const double* eps, const double* r0, const double* theta0_3,
                                        const double* theta0_CS, double K_BP, double K_CS,
                                        double alpha);
/* protein-DNA non-specific hydrogen bond, donor (d1 = CA of the N-side neighbour, d2 = CA,
 * d3 = CA of the C-side neighbour), acceptor (a1 = phosphate, a2 = 3' sugar):
 *   E = -k exp(-dr^2 / 2 sigma^2) f(K (theta - theta0)) f(K (phi - phi0)),  K = pi / (2 delta),
 *   dr = |d2-a1| - r0, phi = angle(d2,a1,a2), theta = acos(clamp(sin t1 sin t2 cos p - cos t1 cos t2,
 *   +-0.99999)), t1 = angle(d1,d3,a1), t2 = angle(d3,a1,d2), p = dihedral(d1,d3,a1,d2)
 * per-donor parameter arrays of length n_donors; the cutoff applies to |d1 - a1| (OpenMM's rule)
 *                                         ProteinDNANonSpecificForceFieldGenerator.cpp:38-141 */
int ocm_system_add_pdns(ocm_system*, int group, const ocm_da_topology* topo, const double* k,
                        const double* r0, const double* theta0, const double* phi0, double sigma,
                        double delta);

/* ---- pair terms (per-bead parameter arrays have length N; NaN = bead does not take part) --- */
/* E = eps (s_i+s_j)^12 (r^-12 - rc^-12), r < rc     ExcludedVolumeForceFieldGenerator.cpp:23-92 */
int ocm_system_add_excluded_volume(ocm_system*, int group, double eps, double rc,
                                   const double* radius, const ocm_pair_topology* topo);
/* E = q_i q_j K (exp(-r/debye)/r - cc), r < rc      DebyeHuckelForceFieldGenerator.cpp:52-109 */
int ocm_system_add_debye_huckel(ocm_system*, int group, double K, double debye_nm, double rc,
                                double cutoff_correction, const double* charge,
                                const ocm_pair_topology* topo);
/* E = step(2^(1/6) s - r) e (4((s/r)^12-(s/r)^6) + 1), s,e arithmetic means
 *                                         LennardJonesRepulsiveForceFieldGenerator.cpp:33-115 */
int ocm_system_add_lj_repulsive(ocm_system*, int group, double rc, const double* eps,
                                const double* sigma, const ocm_pair_topology* topo);
/* Weeks-Chandler-Andersen: E = step(2^(1/6) s - r) (4 e ((s/r)^12-(s/r)^6) + e), e GEOMETRIC, s arithmetic mean
 *                                         WeeksChandlerAndersenForceFieldGenerator.cpp:32-105
 * (also serves UniformWeeksChandlerAndersenForceFieldGenerator.cpp:136-175: one sigma, eps for all members of the
 * interaction groups) */
int ocm_system_add_wca(ocm_system*, int group, double rc, const double* eps, const double* sigma,
                       const ocm_pair_topology* topo);
/* Attractive wells of the lipid model: E = -e for r < b, -e f((r - b) / w) for b <= r < b + w, 0 beyond; b = 2^(1/6) s;
 * e GEOMETRIC mean, s and w (omega) arithmetic means; pairs beyond rc contribute nothing (the reference's cutoffs, below,
 * are handed over as they are: Trigonometric's lies inside b + w for unequal sigmas, and the energy jumps there as in OpenMM).
 *   iSoLF:          f(u) = cos^2(pi u / 2)       iSoLFAttractiveForceFieldGenerator.cpp:41-50, cutoff :129-134
 *   Trigonometric:  f(u) = 1 + 2 u^3 - 3 u^2     TrigonometricForceFieldGenerator.cpp:43-53, cutoff :119-123 */
int ocm_system_add_isolf_attractive(ocm_system*, int group, double rc, const double* eps, const double* sigma,
                                    const double* omega, const ocm_pair_topology* topo);
int ocm_system_add_trigonometric(ocm_system*, int group, double rc, const double* eps, const double* sigma,
                                 const double* omega, const ocm_pair_topology* topo);
/* E = e (step(r - 2^(1/6) s) 4((s/r)^12-(s/r)^6) - step(2^(1/6) s - r)), OpenMM switching
 * function on [r_switch, rc]             LennardJonesAttractiveForceFieldGenerator.cpp:38-123 */
int ocm_system_add_lj_attractive(ocm_system*, int group, double rc, double r_switch,
                                 const double* eps, const double* sigma,
                                 const ocm_pair_topology* topo);
/* E = step(s-r) (e ((s/r)^12 - 2 (s/r)^6) + e), e geometric, s arithmetic mean
 *                                         ThreeSPN2ExcludedVolumeForceFieldGenerator.cpp:19-91 */
int ocm_system_add_3spn2_excluded_volume(ocm_system*, int group, double rc, const double* eps,
                                         const double* sigma, const ocm_pair_topology* topo);

/* ---- external terms (one bead each, "CustomExternalForce"; absolute, unwrapped coordinates) ------------------------
 * The reference reaches these through [[forcefields.external]] (ReadTOMLInput.cpp:763-870). */
/* E = k (|x - x0| - v0)^2 (no factor 1/2).  The generator's periodic branch assigns to a shadowed local, so the plain
 * distance is what runs in either boundary; restated as it runs.
 *                                         PositionRestraintForceFieldGenerator.cpp:6-36 */
int ocm_system_add_position_restraint(ocm_system*, int group, int64_t n, const int32_t* index, const double* pos3,
                                      const double* k, const double* v0);
/* E = -f . x (a constant force f on the bead)     PullingForceFieldGenerator.cpp:3-22 */
int ocm_system_add_pulling(ocm_system*, int group, int64_t n, const int32_t* index, const double* force3);
/* E = eps sum_axes ((radius / (x - lower))^12 + (radius / (upper - x))^12); unlimited boundary only (the generator
 * throws under a periodic one)                    EXVRectangularBoxForceFieldGenerator.cpp:7-56 */
int ocm_system_add_exv_rectangular_box(ocm_system*, int group, double eps, const double* lower3, const double* upper3,
                                       int64_t n, const int32_t* index, const double* radius);
/* E = k (dr - v0)^2, dr = |x - h|, h = (axis . x) axis + shift: the distance from the line through `shift` along the unit
 * vector `axis` (shift orthogonal to axis); periodic = 1: minimum image of x - h (OpenMM periodicdistance)
 *                                         CylindricalRestraintForceFieldGenerator.cpp:4-47 */
int ocm_system_add_cylindrical_restraint(ocm_system*, int group, int periodic, int64_t n, const int32_t* index,
                                         const double* axis3, const double* shift3, const double* k, const double* v0);
/* E = 1/2 k (|c1 - c2| - v0)^2, c1, c2 the mass-weighted centres of two bead groups ("CustomCentroidBondForce" with its
 * default weights); periodic = 1: minimum image of c1 - c2     HarmonicCoMPullingForceFieldGenerator.cpp:3-20 */
int ocm_system_add_harmonic_com_pulling(ocm_system*, int group, int periodic, double k, double v0, int64_t n1,
                                        const int32_t* group1, int64_t n2, const int32_t* group2);
/* Go-type 12-10 contact between every donor bead and every acceptor bead closer than the cutoff (one bead per donor and
 * per acceptor: topo->donors / acceptors carry d1 / a1 and -1, -1):  E = k (5 (r0/r)^12 - 6 (r0/r)^10), r = |a1 - d1|,
 * discontinuous at the cutoff as in OpenMM      CombinatorialGoContactForceFieldGenerator.cpp:3-43 */
int ocm_system_add_combinatorial_go_contact(ocm_system*, int group, const ocm_da_topology* topo, double k, double r0);

/* ---- context (replaces OpenMM::Context + LangevinIntegrator) -------------------------------
 * Simulator.cpp:14-17 (Context), IntegratorGenerator.cpp:5-14 (LangevinIntegrator(T, gamma, dt),
 * setRandomNumberSeed).  The system is compiled into device-resident term tables; the system
 * handle may be destroyed afterwards.  device_index: CUDA ordinal. */
int  ocm_context_create(const ocm_system* sys, double temperature_K, double friction_inv_ps,
                        double dt_ps, uint64_t seed, int device_index, ocm_context** out);
void ocm_context_destroy(ocm_context* ctx);

/* Replica ensemble: n_replicas independent copies of `sys` (unlimited boundary only) advanced by the same kernel
 * launches; replica r integrates with seeds[r] and follows exactly the trajectory a stand-alone context with that
 * seed would.  What the reference runs as n_replicas processes, each with its own OpenMM::Context
 * (src/Simulator.cpp:14-17).  Every per-bead array of the calls below then has n_replicas * N entries, replica-major;
 * ocm_set_velocities_to_temperature gives replica r the seed `seed + r`; ocm_get_state reports energies summed over
 * the replicas. */
int ocm_context_create_batch(const ocm_system* sys, int32_t n_replicas, double temperature_K,
                             double friction_inv_ps, double dt_ps, const uint64_t* seeds,
                             int device_index, ocm_context** out);
int ocm_num_replicas(const ocm_context* ctx);
/* per-replica energies: epot[n_replicas] (groups in group_mask), ekin[n_replicas], epot_by_group[32 * n_replicas];
 * any pointer may be NULL */
int ocm_get_replica_energies(ocm_context*, uint32_t group_mask, double* epot, double* ekin,
                             double* epot_by_group);

/* ---- slab domain decomposition of one periodic system over several GPUs --------------------------
 * Replaces nothing in the reference (one OpenMM::Context drives one device, src/Simulator.cpp:14-17); this is the
 * multi-GPU extension for large condensates.  The box is cut into `world` slabs along `axis` (0,1,2 = x,y,z); each
 * rank owns the beads inside its slab, keeps the beads within one list radius of its faces as a halo, and exchanges
 * halo positions with its two neighbours every MD step (full state and migrations at every list rebuild).
 * Every rank builds the SAME ocm_system and makes the SAME sequence of calls (SPMD): ocm_set_positions /
 * ocm_set_velocities take the whole system on every rank, ocm_step / ocm_get_state are collective, and
 * ocm_get_state returns the whole system (and the total energies) on every rank.  ocm_get_neighbor_pairs returns the
 * pairs found by this rank (i owned here); the union over ranks is the pair set.
 *
 * Communicators:
 *   nccl  : one process per GPU.  Rank 0 calls ocm_comm_unique_id, the host program broadcasts the 128 bytes
 *           (e.g. torch.distributed.broadcast), every rank calls ocm_comm_create_nccl.
 *   local : all ranks live in one process, each driven by its own host thread (one or several devices). */
typedef struct ocm_comm ocm_comm;
int  ocm_comm_unique_id(char* id128);
int  ocm_comm_create_nccl(const char* id128, int world, int rank, int device, ocm_comm** out);
int  ocm_comm_create_local(int world, ocm_comm** out);
void ocm_comm_destroy(ocm_comm* comm);   /* after the contexts using it */
int  ocm_context_create_slab(const ocm_system* sys, double temperature_K, double friction_inv_ps, double dt_ps,
                             uint64_t seed, int device_index, ocm_comm* comm, int rank, int axis,
                             ocm_context** out);
int  ocm_slab_info(const ocm_context*, int* rank, int* world, int* axis, double* halo_nm);
/* role of every bead on this rank: 0 = absent, 1 = owned, 2 = halo (all 1 without a decomposition) */
int  ocm_get_roles(ocm_context*, unsigned char* role);

/* tunables (set_option / get_option):
 *   "skin_nm" (Verlet buffer), "nstlist" (0 = adaptive), "threads_per_bead" (0 = auto, 4/8/16/32),
 *   "beads_per_block" (rows per force block, 0 = auto ~ sqrt(rows)/12),
 *   "pair_kernel" (-1 = by size (default): cluster-pair tile lists for pair lists of >= 131 072 beads, Verlet rows below;
 *                  1 = tiles wherever possible (periodic or virtual box, single system, >= 1024 beads in the list); 0 = rows),
 *   "tile_blocks" (blocks of the persistent tile force launch per list, 0 = one wave),
 *   "replica_kernel" (1 (default) = whole-step persistent kernel, one thread block per replica, for unlimited-boundary systems of
 *                  <= 256 beads without donor-acceptor terms, alone or as a replica batch; 0 = general path),
 *   "pdl" (programmatic dependent launch of the per-step kernels: 0 off, 1, 2; no gain inside CUDA graphs),
 *   "phase_timing" (CUDA events between the phases of a step; counters "phase_ms.rebuild|force|integrate|halo"),
 *   slab contexts: "halo_capacity", "halo_push" (1 = per-step halo over NVLink peer memory, 0 = ncclSend/ncclRecv),
 *                  "halo_overlap" (1 default: interior tile entries run while the halo positions travel on a side stream). */
int ocm_context_set_option(ocm_context*, const char* key, double value);
int ocm_context_get_option(ocm_context*, const char* key, double* value);

/* Context::setPositions / setVelocities / setVelocitiesToTemperature
 * (Simulator.cpp:41, Simulator.hpp:49-65); xyz arrays are 3*N doubles */
int ocm_set_positions(ocm_context*, const double* xyz_nm);
int ocm_set_velocities(ocm_context*, const double* xyz_nm_per_ps);
int ocm_set_velocities_to_temperature(ocm_context*, double temperature_K, uint64_t seed);

/* LangevinIntegrator::step(n) (Simulator.cpp:60) */
int ocm_step(ocm_context*, int64_t nsteps);
int64_t ocm_current_step(const ocm_context*);

/* Observer side of a frame without stalling the run loop (Observer.cpp:103-118: two synchronous getState downloads per
 * frame; Observer.cpp:324-346: G+1 energy evaluations; caller Simulator.cpp:55-60).
 *   ocm_get_state_begin  snapshots what was asked for at the current step: the energy / force kernels are queued on the
 *                        compute stream behind the steps already issued (they reuse the pair lists of the last window
 *                        when no bead has left its Verlet buffer, so no list build), positions, velocities, forces and
 *                        energies are converted on the device into one staging block, and its download into PINNED host
 *                        memory runs on a side stream.  Returns at once; ocm_step may be called right away and the
 *                        download overlaps the next window.
 *   ocm_get_state_end    waits for that download only (not for steps issued after begin) and copies to the caller's
 *                        arrays (any may be NULL); the request is consumed.
 * ocm_get_state is the blocking form.  Replica batches and slab contexts accept the calls but serve them synchronously
 * inside ocm_get_state_end (at the state of that moment).
 * ocm_set_state_async: positions and velocities through pinned staging, converted on the device, no host synchronisation
 * (velocities may be NULL). */
typedef struct ocm_state_request ocm_state_request;
int ocm_get_state_begin(ocm_context*, uint32_t group_mask, int want_positions, int want_velocities, int want_forces,
                        int want_energies, ocm_state_request** out);
int ocm_get_state_end(ocm_state_request*, double* pos, double* vel, double* force, double* e_pot, double* e_kin,
                      double* e_by_group32);
int ocm_set_state_async(ocm_context*, const double* xyz_nm, const double* vel_nm_per_ps);

/* LocalEnergyMinimizer::minimize(context) (Simulator.cpp:43-46; OpenMM defaults: tolerance 10 kJ/mol/nm,
 * maxIterations 0 = until convergence): L-BFGS on the potential energy from the current positions until the
 * RMS force per particle drops below `tolerance`.  Energies and forces are evaluated on the device, the
 * two-loop recursion runs on the host.  Velocities are untouched.  Outputs may be NULL. */
int ocm_minimize(ocm_context*, double tolerance_kj_mol_nm, int max_iterations, int* iterations,
                 double* final_energy);

/* Monte Carlo barostats: OpenMM::MonteCarloAnisotropicBarostat (MonteCarloAnisotropicBarostatGenerator.cpp:9-17,
 * ensemble "NPT", ReadTOMLInput.cpp:88-109) and OpenMM::MonteCarloMembraneBarostat
 * (MonteCarloMembraneBarostatGenerator.cpp:6-14, ensemble "NPgammaT", ReadTOMLInput.cpp:110-169), added to the System by
 * SystemGenerator.cpp:89-108.  OpenMM 8.0.0 semantics restated: before the MD step at which the barostat's call counter
 * reaches `frequency`, one trial changes the box volume by a uniform random amount along one axis (membrane barostat: X and Y
 * together in the isotropic mode, Z compensating in the constant-volume mode), every molecule (the connected components of
 * the bonded terms) is translated rigidly so that its centre, wrapped into the box, scales with the box, and the trial is
 * accepted with the Metropolis rule on
 *   w = dE + P dV - gamma dA - N_mol kT ln(V'/V)      (P in bar, gamma in bar nm; 1 bar = N_A 1e-25 kJ/mol/nm^3);
 * the step size of an axis is tuned every 10 trials to keep the acceptance between 25 % and 75 %.  The energy evaluations,
 * the molecule scaling and the restore of a rejected trial run on the device; the decision is one host read of two energies.
 * kind OCM_BAROSTAT_ANISOTROPIC: pressure3_bar = {Px, Py, Pz}, mode3 = {scaleX, scaleY, scaleZ} (0/1), surface tension
 * ignored.  kind OCM_BAROSTAT_MEMBRANE: pressure3_bar[0] = P, mode3 = {xymode (0 isotropic, 1 anisotropic), zmode (0 free,
 * 1 fixed, 2 constant volume), 0}.  OpenMM draws from SFMT with a seed of its own; here the stream is SplitMix64 of
 * `seed`, restated in the oracle, so that engine and oracle take the same decisions.  kind OCM_BAROSTAT_NONE removes it.
 * Periodic single-system contexts only (no replica batch, no slab decomposition). */
enum { OCM_BAROSTAT_NONE = 0, OCM_BAROSTAT_ANISOTROPIC = 1, OCM_BAROSTAT_MEMBRANE = 2 };
int ocm_context_set_barostat(ocm_context*, int kind, const double* pressure3_bar, double surface_tension_bar_nm,
                             double temperature_K, const int32_t* mode3, int32_t frequency, uint64_t seed);
/* State::getPeriodicBoxVectors (Observer.cpp:193): the current box edges in nm (zeros without a box).
 * ocm_set_box = Context::setPeriodicBoxVectors: positions are kept, cell grids and lists are laid out again. */
int ocm_get_box(ocm_context*, double* box3_nm);
int ocm_set_box(ocm_context*, const double* box3_nm);
/* barostat bookkeeping: out12 = {trials, accepted, last axis, last dV, last E before, last E after, last w, last accepted (0/1),
 * volume step of X, Y, Z, number of molecules} */
int ocm_get_barostat_info(ocm_context*, double* out12);

/* Context::getState (Observer.cpp:34,107,113,340,344).  Any output pointer may be NULL.
 * pos/vel/force: 3*N doubles; epot: sum over groups in group_mask; ekin: leap-frog kinetic
 * energy shifted by dt/2 as OpenMM reports it; epot_by_group: 32 doubles (all groups). */
int ocm_get_state(ocm_context*, uint32_t group_mask, double* pos, double* vel, double* force,
                  double* epot, double* ekin, double* epot_by_group);

/* ---- introspection used by the bit-exact set tests ---------------------------------------- */
int ocm_num_pair_terms(const ocm_context*);
/* unordered pairs (i<j, sorted) of pair term `term` (index in order of the add_* calls) that
 * are inside its interaction cutoff at the current positions, not excluded, and allowed by its
 * interaction groups.  First call with ij == NULL to get the count. */
int ocm_get_neighbor_pairs(ocm_context*, int term, int32_t* ij, int64_t capacity, int64_t* n_pairs);
/* cell index of every participant of the pair list serving `term` (-1 for non-participants),
 * and the grid dimensions */
int ocm_get_cell_ids(ocm_context*, int term, int32_t* cell_of_bead, int32_t dims[3]);
/* donor-acceptor terms: number of terms, and the (donor index, acceptor index) combinations of term `term`
 * that are evaluated at the current positions (not excluded, |d1 - a1| < cutoff), sorted.  First call with
 * da == NULL to get the count. */
int ocm_num_da_terms(const ocm_context*);
int ocm_get_da_pairs(ocm_context*, int term, int32_t* da, int64_t capacity, int64_t* n_pairs);
/* counters: "pair_evals_per_step" (neighbor-list entries walked), "rebuilds", "rollbacks",
 * "kernel_launches", "nstlist", "max_neighbors", "num_pair_lists", "last_step_ms" (device time of the
 * last ocm_step, CUDA events on the engine's stream), "list_cutoff_nm.<term>", "threads_per_bead.<term>",
 * "virtual_box_x|y|z" (unlimited boundary: the periodic box the pair kernels run in, 0 otherwise; the oracle takes it
 * as an option of the same name so that pair sets and cell ids stay bit-exact), "virtual_box_layouts",
 * "minimize_evaluations", "owned_beads", "halo_exchanges", "halo_message_bytes", "halo_push" (slab contexts) */
int ocm_get_counter(ocm_context*, const char* key, double* value);

/* ---- measurement hooks (bench.py; CUDA events on the engine's own stream) ------------------- */
/* average device time (ms) of `reps` back-to-back launches of one kernel at the current state:
 * "force" (the fused per-step force launch), "integrate", "rebuild" (whole list rebuild sequence) */
int ocm_profile_kernel(ocm_context*, const char* which, int reps, double* avg_ms);
/* register-resident FP32 FMA microbenchmark: the attainable CUDA-core rate (TFLOP/s) that the pair
 * kernel's roofline is quoted against */
int ocm_measure_fp32_peak(int device_index, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* OCM_B200_H */
