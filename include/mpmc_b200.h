/* mpmc_b200.h — C-ABI of the B200 energy engine for mpmc++ (b-tudor/mpmcxx).
 *
 * This is the drop-in boundary for ONE path of the reference: `double System::energy()`
 * (src/System.h:315, body src/System.Energy.cpp:19-171) and the per-bead aggregator
 * `SimulationControl::PI_calculate_potential()` (src/SimulationControl.PathIntegral.cpp:752-805),
 * i.e. what the reference's dormant `cuda on` switch (src/SimulationControl.cpp:1329-1336,
 * src/System.Energy.cpp:67-74 `polar_cuda`) was meant to reach.  The reference has no plugin API;
 * each entry point below cites the reference routine(s) it replaces.  INTEGRATION.md shows the
 * few lines a maintainer adds to System::energy() to call it.
 *
 * Conventions
 *   - plain C types only; every function returns an int status (0 = ok).  Non-zero codes reuse the
 *     reference's exception integers (src/constants.h:108-147) so a C++ caller can `throw rc;`.
 *   - sites are passed flat, in the reference's list order (Molecule -> Atom walk, src/System.cpp:881-904).
 *     Site order is significant: it is the Gauss-Seidel sweep order (System.Energy.cpp:3570).
 *   - lengths in Angstrom, energies in Kelvin, charges already in reduced units sqrt(K*A)
 *     (e * 408.7816, src/System.cpp:624), polarizabilities in A^3.
 *   - one engine per reference `System` (or per set of path-integral bead systems); calls on one
 *     engine must be externally serialised (the reference's energy() is not re-entrant either).
 *   - there is NO CPU fallback: without a CUDA device mpmc_create fails with MPMC_ERR_CUDA.
 */
#ifndef MPMC_B200_H
#define MPMC_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define MPMC_ABI_VERSION 1

/* status codes: src/constants.h:108-147 */
enum {
	MPMC_OK                  = 0,
	MPMC_ERR_INTERNAL        = 101,   /* internal_error */
	MPMC_ERR_ALLOC           = 2000,  /* memory_request_fail */
	MPMC_ERR_ALLOC_INVALID   = 2001,  /* memory_request_invalid */
	MPMC_ERR_INVALID_INPUT   = 3000,  /* invalid_input */
	MPMC_ERR_NO_MOLECULES    = 3001,  /* no_molecules_in_system */
	MPMC_ERR_INVALID_SETTING = 4000,  /* invalid_setting */
	MPMC_ERR_INCOMPATIBLE    = 4002,  /* incompatible_settings */
	MPMC_ERR_MISSING_SETTING = 4003,  /* missing_setting */
	MPMC_ERR_UNSUPPORTED     = 4004,  /* unsupported_setting */
	MPMC_ERR_INVALID_BOX     = 6004,  /* invalid_box_dimensions */
	MPMC_ERR_BEADS           = 6005,  /* incongruent_bead_states */
	MPMC_ERR_CUDA            = 30000  /* CUDA runtime failure; mpmc_last_error() has the text */
};

enum { MPMC_DAMPING_OFF = 0, MPMC_DAMPING_LINEAR = 1, MPMC_DAMPING_EXPONENTIAL = 2 };  /* constants.h:67-71 */

/* The input keywords that parameterise energy() (src/SimulationControl.cpp:1186-1327), with the
 * reference's defaults noted.  Zero-initialise, then set. */
typedef struct mpmc_config {
	double basis[9];            /* pbc.basis[i][j] row-major: row i = lattice vector `basis<i+1>` (:1494) */
	int    n_beads;             /* 1 for classic ensembles; P_local bead systems for pi_nvt */
	int    capacity;            /* max sites per bead system (uVT growth); 0 = grow on demand */
	int    device;              /* CUDA device ordinal */
	/* repulsion-dispersion / electrostatics */
	int    rd_lrc;              /* default 1 (System.h:631) */
	int    rd_only;             /* skip Coulomb + polarization (System.Energy.cpp:46) */
	int    ewald_kmax;          /* default 7 (System.h:22) */
	double ewald_alpha;         /* <= 0: derive 3.5/cutoff (System.cpp:871) */
	/* Thole polarization (System.Energy.cpp:2534-3762) */
	int    polarization;
	int    polar_ewald;         /* static field by Ewald (recip_term + real_term); 0 = thole_field_nopbc */
	int    polar_iterative;     /* must be 1 (the reference's own `cuda on` validator demands it, SimulationControl.cpp:2612) */
	int    damp_type;           /* MPMC_DAMPING_* */
	int    polar_gs, polar_gs_ranked, polar_palmo, polar_sor, polar_esor, polar_zodid, polar_rrms;
	int    polar_max_iter;      /* fixed-iteration mode when polar_precision == 0 */
	double polar_damp;          /* Thole lambda */
	double polar_gamma;         /* default 1.0 (System.h:698) */
	double polar_precision;     /* Debye; > 0 selects convergence mode (<= 128 iterations) */
	double polar_ewald_alpha;   /* <= 0: derive 3.5/cutoff (System.cpp:873) */
	int    reserved[8];
} mpmc_config;

/* What energy() leaves in System::observables / nodestats (src/System.h:94-113), plus every sub-term
 * separately (the Ewald pieces cancel by ~5 orders of magnitude, so parity is judged per sub-term). */
typedef struct mpmc_energy_out {
	double energy;                 /* observables->energy = rd + coulombic + polarization */
	double rd_energy;              /* lj(): pair + pair LRC + self LRC (System.Energy.cpp:897-1032) */
	double coulombic_energy;       /* coulombic(): real - intra + reciprocal + self (:1396-1416) */
	double polarization_energy;    /* polar() (:2534-2635) */
	double vdw_energy;             /* always 0 (polarvdw is out of scope) */
	double rd_pair, rd_lrc_pair, rd_lrc_self;
	double es_real, es_self_intra, es_reciprocal, es_self;
	double dipole_rrms;            /* observables->dipole_rrms */
	double n_pairs_in_cutoff;      /* LJ pairs actually inside the cutoff sphere */
	double n_pair_evals;           /* pair distances evaluated by the pair sweep */
	int    polarization_iterations;/* nodestats->polarization_iterations */
	int    iterator_failed;        /* System::iterator_failed (:3483-3494) -> caller rejects the move */
	int    reserved[4];
} mpmc_energy_out;

typedef struct mpmc_engine mpmc_engine;   /* opaque */

int         mpmc_abi_version(void);
const char *mpmc_last_error(void);                          /* text of the last failure on this thread */
int         mpmc_device_count(int *count);

/* lifetime: replaces allocate_pair_lists() / thole_resize_matrices() (src/System.Pairs.cpp:21,
 * src/System.cpp:1430) — the engine owns device buffers instead of Pair lists and the A matrix. */
int mpmc_create(const mpmc_config *cfg, mpmc_engine **out);
int mpmc_destroy(mpmc_engine *e);

/* update_pbc() + PeriodicBoundary::update() (src/System.cpp:859-876, src/PeriodicBoundary.cpp:31-101):
 * volume, reciprocal basis, cutoff = half the shortest lattice vector, Ewald alphas, k-vector table. */
int mpmc_set_cell(mpmc_engine *e, const double basis[9]);
/* out[22] = basis[9], reciprocal_basis[9], volume, cutoff, ewald_alpha, polar_ewald_alpha */
int mpmc_get_cell(mpmc_engine *e, double out[22]);

/* rebuild_arrays() + pair_exclusions() inputs (src/System.cpp:881-904, 1035-1177): the whole site table.
 * pos is n_beads * n * 3 (bead-major); all other arrays are per site and shared by every bead system.
 * mol[] must be non-decreasing (sites of a molecule are contiguous, as in the reference's lists). */
int mpmc_upload_sites(mpmc_engine *e, int n, const double *pos, const double *charge, const double *alpha,
                      const double *epsilon, const double *sigma, const double *mass, const int *mol, const int *frozen);

/* a displace / rotate / bead-perturb move: new coordinates for `count` consecutive sites starting at `first`
 * in bead system `bead` (make_move(), src/System.MonteCarlo.cpp:875; restore(), :1510).  pos: count*3. */
int mpmc_update_sites(mpmc_engine *e, int bead, int first, int count, const double *pos);
/* all beads at once (PI_displace / restore_PI_systems): pos is n_beads * count * 3, bead-major */
int mpmc_update_sites_all_beads(mpmc_engine *e, int first, int count, const double *pos);

/* uVT: update_pairs_insert()/update_pairs_remove() (src/System.Pairs.cpp:53,100).  The new molecule's sites go
 * BEFORE site `before` (the reference links an inserted molecule in front of the selected one,
 * src/System.MonteCarlo.cpp:799-805); pos is n_beads*count*3. */
int mpmc_insert_sites(mpmc_engine *e, int before, int count, const double *pos, const double *charge, const double *alpha,
                      const double *epsilon, const double *sigma, const double *mass, int frozen);
int mpmc_remove_sites(mpmc_engine *e, int first, int count);
int mpmc_num_sites(mpmc_engine *e, int *n);

/* System::energy() (src/System.Energy.cpp:19-171), one full evaluation per bead system; out has n_beads entries. */
int mpmc_energy(mpmc_engine *e, mpmc_energy_out *out);
/* the same, split so that a caller can overlap host work: enqueue all kernels, then fetch the result */
int mpmc_energy_enqueue(mpmc_engine *e);
int mpmc_energy_fetch(mpmc_engine *e, mpmc_energy_out *out);

/* Atom::mu / ef_static / ef_induced / ef_induced_change of bead system `bead` after energy()
 * (read by write_dipole/write_field, src/System.Output.cpp:1132-1232); each n*3, any may be NULL. */
int mpmc_download_dipoles(mpmc_engine *e, int bead, double *mu, double *ef_static, double *ef_induced, double *ef_induced_change);
/* Atom::rank_metric (src/System.cpp:1000-1029), n doubles */
int mpmc_download_rank_metric(mpmc_engine *e, int bead, double *rank_metric);

/* PI_calculate_potential() (src/SimulationControl.PathIntegral.cpp:752-805) over this engine's bead systems:
 * per_bead[n_beads*4] = rd, coulombic, polarization, vdw of each bead (what the reference all-gathers, :763-766);
 * sums[4] = their sums over the local beads (divide by the GLOBAL P after the cross-GPU all-reduce). */
int mpmc_pi_potential(mpmc_engine *e, double *per_bead, double sums[4]);
/* molecular centres of mass of every bead system (Molecule::update_COM, src/Molecule.cpp:256-281): com is
 * n_beads*n_mol*3, mol_mass is n_mol; and the bead-spring sum of PI_chain_mass_length2_ENTIRE_SYSTEM
 * (PathIntegral.cpp:859-970) over the links BETWEEN LOCAL beads (bead b -> b+1, b+1 < n_beads), plus, when
 * `closed` != 0, the link from the last local bead back to the first (single-GPU ring).  kg*m^2. */
int mpmc_pi_chain(mpmc_engine *e, int closed, double *chain_mass_len2, double *com, double *mol_mass, int *n_mol);

/* Beads sharded over the GPUs of one box (replaces the 4 MPI_Allgather of PathIntegral.cpp:763-766 and the rank-per-bead layout of
 * SimulationControl.cpp:53-65): rank 0 makes an id, every rank joins with its engine (NCCL is loaded with dlopen on first use).
 * mpmc_pi_potential_allreduce: one sweep over the local beads, device-side assembly of the four per-bead-sum scalars, ONE ncclAllReduce
 * of 4 doubles on the engine's stream, then means[4] = rd, coulombic, polarization, vdw divided by the global Trotter number and
 * *potential = their sum (what PI_calculate_potential returns).  Works without mpmc_nccl_init too (single GPU).
 * mpmc_pi_chain_allreduce: PI_chain_mass_length2_ENTIRE_SYSTEM over the whole ring, the link between consecutive ranks included. */
int mpmc_nccl_get_unique_id(char id[128]);
int mpmc_nccl_init(mpmc_engine *e, const char id[128], int rank, int nranks);
int mpmc_pi_potential_allreduce(mpmc_engine *e, int P_global, double means[4], double *potential);
int mpmc_pi_chain_allreduce(mpmc_engine *e, double *chain_mass_len2);
/* how mpmc_pi_potential_allreduce sums over ranks: 0 = single GPU, 1 = ncclAllReduce, 2 = peer-memory mailboxes written over NVLink by
 * the assembly kernel itself (set up by mpmc_nccl_init when CUDA IPC + peer access are available; MPMC_PI_P2P=0 forces 1) */
int mpmc_pi_collective(mpmc_engine *e);

/* measurement hooks (bench.py): the CUDA stream every kernel of this engine is launched on (a cudaStream_t),
 * and the number of kernels this engine has launched so far. */
/* per-kernel-class device time, measured with CUDA events recorded on the engine's stream around each launch
 * (accumulated from mpmc_set_timing(e,1) on; read after an energy fetch). */
enum {
	MPMC_K_ENERGY_TOTAL = 0,  /* everything one energy() enqueues */
	MPMC_K_PAIR,              /* k_pair_gather + k_pair_sweep: lj() + coulombic_real() */
	MPMC_K_STRUCTURE,         /* k_structure_partial: S(k) chunk partials */
	MPMC_K_FIELD_RECIP,       /* k_field_recip: recip_term() */
	MPMC_K_FIELD_REAL,        /* k_field_parts + k_field_finish: real_term() / thole_field_nopbc() */
	MPMC_K_RANK,              /* k_rank_min_parts, k_rank_lim, k_rank_count_parts: the Gauss-Seidel rank metric */
	MPMC_K_DIPOLE_SWEEP,      /* k_contract_parts + k_contract_finish: Jacobi contract_dipoles(), and the Gauss-Seidel pipeline's initial contraction */
	MPMC_K_GS_SWEEP,          /* ONE Gauss-Seidel sweep: k_gs_pipeline (solver + helpers) with k_gs_updaters beside it; count = sweeps */
	MPMC_K_PALMO,             /* palmo_contraction(): k_contract_parts over the needed rows / k_gs_palmo */
	MPMC_K_GS_PRECOMPUTE,     /* per sweep order: k_gs_gather, k_gs_inverse, k_gs_near (on the second stream unless timing is on) */
	MPMC_NUM_KERNEL_CLASSES
};
int mpmc_set_timing(mpmc_engine *e, int on);
/* developer hook: SM-clock stamps of the Gauss-Seidel pipeline (solver and one updater CTA), see tools/gs_profile.py */
int mpmc_debug_gs_profile(mpmc_engine *e, int enable, long long *out, int max_blocks, int *nblk);
int mpmc_debug_pair_profile(mpmc_engine *e, int enable, long long *out, int max_warps, int *nwarps);   /* per-warp timeline of the pair sweep */
int mpmc_debug_mark_moved(mpmc_engine *e, int first, int count);   /* bench hook: next evaluation treats these sites as moved */
int mpmc_get_timing(mpmc_engine *e, double ms[MPMC_NUM_KERNEL_CLASSES], long long count[MPMC_NUM_KERNEL_CLASSES]);
/* test hooks for the host-side numerics (no device needed).
 * mpmc_debug_radial_table: build the r^2-indexed piecewise-polynomial table the kernels use (kind 0: erfc(a r)/r of
 * coulombic_real, src/System.Energy.cpp:1493-1497; kind 1: the two radial factors of real_term, :2921-2929; param = alpha)
 * on [u_lo, u_hi) and evaluate it exactly as the device does at u[0..n).
 * mpmc_debug_cutoff_thresholds: out[0] = largest r^2 with sqrt(r^2) - 1e-12 < cutoff (lj, :934), out[1] = largest r^2 with
 * !(sqrt(r^2) > cutoff) (coulombic_real, :1490; real_term, :2917). */
int mpmc_debug_radial_table(int kind, double param, double u_lo, double u_hi, const double *u, int n, double *out0, double *out1);
int mpmc_debug_cutoff_thresholds(double cutoff, double out[2]);
void     *mpmc_stream(mpmc_engine *e);
long long mpmc_kernel_launches(mpmc_engine *e);

/* FP64 FMA peak probe used as the roofline denominator (MEASURED_PEAKS.json has no FP64 entry): runs a
 * register-resident DFMA loop on all SMs and returns TFLOP/s (2 flops per FMA). */
int mpmc_probe_fp64_peak(int device, double *tflops, double *sm_clock_mhz_guess);

#ifdef __cplusplus
}
#endif
#endif /* MPMC_B200_H */
