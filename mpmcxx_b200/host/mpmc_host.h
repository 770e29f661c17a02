// mpmc_host.h — host-side mirror of the reference's System / Molecule / Atom / SimulationControl API for the paths that
// call the energy hot path: the classic Markov chain (System::mc, src/System.MonteCarlo.cpp:20-134) and the path-integral
// chain (SimulationControl::PI_nvt_mc, src/SimulationControl.PathIntegral.cpp:31-196).  Class, member and field names follow
// the reference so that its callers read the same; the bodies are written for this engine: energy() flattens the lists and
// calls the C-ABI of include/mpmc_b200.h, there are no Pair lists and no A matrix.
//
// Accept/reject trajectories are reproducible against the reference because the two RNG streams (System::mt_rand and the
// global Rando) are consumed in the reference's order with libstdc++'s own distributions, and the
// move arithmetic rounds like the reference's (-ffp-contract=off).
#pragma once
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <random>
#include <string>
#include <vector>

#include "../../include/mpmc_b200.h"

namespace mpmc_host {

// src/constants.h
constexpr double pi = 3.141592653589793238462643383279502884L;
constexpr double h = 6.626068e-34, hBar2 = 1.11211999e-68, kB = 1.3806503e-23;
constexpr double AMU2KG = 1.66053873e-27, METER2ANGSTROM = 1.0e10, ANGSTROM2METER = 1.0e-10, E2REDUCED = 408.7816, ATM2REDUCED = 0.0073389366;
constexpr double MAXVALUE = 1.0e40;
enum { ENSEMBLE_UVT, ENSEMBLE_NVT, ENSEMBLE_SURF, ENSEMBLE_SURF_FIT, ENSEMBLE_NVE, ENSEMBLE_TE, ENSEMBLE_NPT, ENSEMBLE_REPLAY, ENSEMBLE_PATH_INTEGRAL_NVT, ENSEMBLE_NVT_GIBBS };
enum { MOVETYPE_INSERT, MOVETYPE_REMOVE, MOVETYPE_DISPLACE, MOVETYPE_ADIABATIC, MOVETYPE_SPINFLIP, MOVETYPE_VOLUME, MOVETYPE_PERTURB_BEADS };
enum { DAMPING_OFF, DAMPING_LINEAR, DAMPING_EXPONENTIAL };
// error codes thrown as int, exactly like the reference (constants.h:108-147)
constexpr int internal_error = 101, invalid_monte_carlo_move = 102, fopen_fail_read = 1000, invalid_input = 3000, no_molecules_in_system = 3001,
              invalid_setting = 4000, invalid_ensemble = 4001, incompatible_settings = 4002, missing_setting = 4003, unsupported_setting = 4004,
              missing_required_datum = 6000, invalid_datum = 6001, molecule_wo_atoms = 6002, invalid_box_dimensions = 6004,
              invalid_MPI_size_for_PI = 12000;

// src/Rando.h: one global engine shared by a uniform and a (stateful) normal distribution
class Rando {
public:
	// reset(): a fresh reference process starts with no cached Box-Muller value; seeding here restores exactly that state
	static void seed(unsigned int s) { mt.seed(s); normal_distribution.reset(); uniform_distribution.reset(); }
	static double rand() { return uniform_distribution(mt); }
	static double rand_normal() { return normal_distribution(mt); }
private:
	static std::mt19937 mt;
	static std::normal_distribution<double> normal_distribution;
	static std::uniform_real_distribution<double> uniform_distribution;
};

struct PeriodicBoundary {            // src/PeriodicBoundary.h
	double cutoff = 0, volume = 0;
	double basis[3][3] = {{0}}, reciprocal_basis[3][3] = {{0}};
};

class alignas(64) Atom {             // src/Atom.h (the fields the energy path and the PQR format use)
public:
	// what a move touches — the list link, the mass (centre of mass) and the position — shares the object's first cache line: a
	// path-integral move walks the picked molecule's atoms in every bead system, and none of them is in cache
	Atom *next = nullptr;
	double mass = 0;
	double pos[3] = {0, 0, 0};
	int id = 0, frozen = 0, adiabatic = 0, spectre = 0, target = 0;
	double charge = 0, polarizability = 0, epsilon = 0, sigma = 0, omega = 0;
	double gwp_alpha = 0, c6 = 0, c8 = 0, c10 = 0, c9 = 0;   // carried from the PQR file to the PQR file (not used on this path)
	double wrapped_pos[3] = {0, 0, 0};
	double ef_static[3] = {0, 0, 0}, ef_induced[3] = {0, 0, 0}, ef_induced_change[3] = {0, 0, 0}, mu[3] = {0, 0, 0};
	double rank_metric = 0;
	char atomtype[64] = {0};
};

class Molecule {                     // src/Molecule.h
public:
	Molecule() = default;
	Molecule(const Molecule &orig);  // deep copy of the atom list; `next` is not copied
	~Molecule();
	Molecule &operator=(const Molecule &) = delete;
	void rotate_rand(double scale);
	void rotate(double x, double y, double z, double angle_degrees);
	void translate_rand_pbc(double scale, const PeriodicBoundary &pbc, std::mt19937 *mt_rand);
	void translate_rand_pbc(double scale, const PeriodicBoundary &pbc, double dice[6]);
	void translate(double x, double y, double z);
	void move_to_(double x, double y, double z);
	void orient(const double orientation[3], int orientation_site);   // src/Molecule.cpp:211-254
	void update_COM();
	int natoms() const;

	char moleculetype[32] = {0};
	int id = 0;
	double mass = 0;
	int frozen = 0, adiabatic = 0, spectre = 0, target = 0;
	double com[3] = {0, 0, 0}, wrapped_com[3] = {0, 0, 0};
	Atom *atoms = nullptr;
	Molecule *next = nullptr;
};

class System {
public:
	struct observables_t {           // src/System.h:94-113
		double energy = 0, coulombic_energy = 0, rd_energy = 0, polarization_energy = 0, vdw_energy = 0, three_body_energy = 0, dipole_rrms = 0,
		       kinetic_energy = 0, temperature = 0, volume = 0, N = 0, NU = 0, spin_ratio = 0, frozen_mass = 0, total_mass = 0;
		double potential() const { return coulombic_energy + rd_energy + polarization_energy + vdw_energy + three_body_energy; }
	};
	struct checkpoint_t {            // src/System.h:115-124
		int movetype = MOVETYPE_DISPLACE, biased_move = 0;
		Molecule *molecule_backup = nullptr, *molecule_altered = nullptr, *head = nullptr, *tail = nullptr;
		observables_t observables;
	};
	struct avg_observables_t {       // src/System.h:44-92 (what a constant-volume Markov chain reports)
		double energy = 0, energy_sq = 0, energy_error = 0, N = 0, N_sq = 0, N_error = 0;
		double coulombic_energy = 0, coulombic_energy_sq = 0, coulombic_energy_error = 0, rd_energy = 0, rd_energy_sq = 0, rd_energy_error = 0;
		double polarization_energy = 0, polarization_energy_sq = 0, polarization_energy_error = 0;
		double kinetic_energy = 0, kinetic_energy_sq = 0, kinetic_energy_error = 0;
		double density = 0, density_sq = 0, density_error = 0, pore_density = 0, percent_wt = 0, percent_wt_me = 0, excess_ratio = 0;
		double NU = 0, qst = 0, heat_capacity = 0, heat_capacity_error = 0, compressibility = 0, compressibility_error = 0;
	};
	struct nodestats_t { double boltzmann_factor = 0, polarization_iterations = 0; int accept = 0, reject = 0; };
	struct step_record { int movetype; double final_energy, boltzmann_factor; int accepted; double N; };

	System() = default;
	System(const System &orig);      // copies settings and deep-copies the molecule list (what `new System(sys)` does for PI beads)
	~System();
	System &operator=(const System &) = delete;

	// geometry
	void read_molecules(const char *pqr_file);      // src/System.cpp:507-770
	void read_pqr_box(const char *pqr_file);         // src/System.cpp:775-850: "REMARK BOX BASIS[k] = x y z" lines override the input's basis
	int write_molecules(FILE *fp);                   // src/System.Output.cpp:900-1091: the PQR format, field for field
	int write_molecules_wrapper(const char *filename);   // :837-895: previous file -> "<name>.last", then write
	void update_pbc();                               // src/System.cpp:859-876 + PeriodicBoundary::update
	int countNatoms() const;
	unsigned int countN();
	void update_com();
	void wrap_all();
	// the hot path
	double energy();                                 // src/System.Energy.cpp:19-171, through the engine
	void download_dipoles();                         // Atom::mu / ef_* after energy()
	// the classic Markov chain
	bool mc(std::vector<step_record> *log = nullptr);
	double mc_initial_energy();
	void do_checkpoint();
	void make_move();
	void boltzmann_factor(double initial_energy, double final_energy);
	void restore();
	void displace(Molecule *molecule, const PeriodicBoundary &pbc, double trans_scale, double rot_scale);
	double get_rand() { return dist(mt_rand); }
	// averages over the samples taken every correlation time and at the very end (src/System.Averages.cpp:8-208,
	// src/System.MonteCarlo.cpp:104-106 -> do_corrtime_bookkeeping)
	void calc_system_mass();                         // src/System.cpp:1537-1550
	void update_root_averages(observables_t *obs);

	// settings (names as in src/System.h)
	int cuda = 1, ensemble = ENSEMBLE_NVT;
	char job_name[256] = "untitled", pqr_input[512] = {0}, pqr_output[512] = {0}, pqr_restart[512] = {0};
	int long_output = 0, independent_particle = 0;
	double loop_seconds = 0;         // wall seconds of the last mc() step loop (initial energy excluded)
	bool write_files = true;         // restart / final PQR files are written like the reference's (tests and benches may switch it off)
	uint32_t numsteps = 0, corrtime = 0, step = 0;
	double move_factor = 1.0, rot_factor = 1.0, insert_probability = 0, bead_perturb_probability = 0;
	double temperature = 0, pressure = 0, free_volume = 0, scale_charge = 1.0;
	int preset_seed_on = 0; unsigned int preset_seed = 0;
	int rd_lrc = 1, rd_only = 0, wrapall = 1, parallel_restarts = 0, read_pqr_box_on = 0;
	int ewald_alpha_set = 0, polar_ewald_alpha_set = 0, ewald_kmax = 7;
	double ewald_alpha = 0.5, polar_ewald_alpha = 0.5;
	int polarization = 0, polar_iterative = 0, polar_ewald = 0, polar_zodid = 0, polar_palmo = 0, polar_rrms = 0, polar_gs = 0, polar_gs_ranked = 0,
	    polar_sor = 0, polar_esor = 0, polar_max_iter = 0, damp_type = DAMPING_OFF, iterator_failed = 0;
	double polar_gamma = 1.0, polar_damp = 0, polar_precision = 0;
	int gpu_device = 0;

	// state
	PeriodicBoundary pbc;
	Molecule *molecules = nullptr;
	int natoms = 0;
	observables_t observables_store, *observables = &observables_store;
	checkpoint_t checkpoint_store, *checkpoint = &checkpoint_store;
	nodestats_t nodestats_store, *nodestats = &nodestats_store;
	avg_observables_t avg_observables_store, *avg_observables = &avg_observables_store;
	int avg_counter = 0;             // the function-static sample counter of update_root_averages
	double fugacities[1] = {0};      // fugacities[0] of the reference: 0 unless a fugacity keyword sets it (this mirror accepts none)
	double last_volume = 0;
	std::mt19937 mt_rand;
	std::uniform_real_distribution<double> dist{0, 1};

	// engine binding
	mpmc_engine *gpu = nullptr;
	std::vector<double> gpu_pos;
	bool gpu_table_stale = true;     // set by insert / remove / restore of those: the whole site table is re-sent
	void fill_config(mpmc_config &c, int n_beads) const;
	void flatten(std::vector<double> &pos, std::vector<double> &q, std::vector<double> &al, std::vector<double> &ep, std::vector<double> &sg,
	             std::vector<double> &ms, std::vector<int> &mol, std::vector<int> &fz) const;
};

class SimulationControl {
public:
	SimulationControl(const char *inFilename, int P);
	~SimulationControl();
	void initializeSimulationObjects();
	bool runSimulation(std::vector<System::step_record> *log = nullptr);

	System sys;                      // template system / aggregate observables (src/SimulationControl.h:23)
	int nSys = 0;                    // Trotter number
	int PI_trial_chain_length = 0;
	std::vector<System *> systems;   // one per bead
	// bead sharding over GPUs (one process per GPU): this process evaluates beads [P rank / nranks, P (rank + 1) / nranks)
	int rank = 0, nranks = 1;
	char nccl_id[128] = {0};
	// what the last run cost: wall seconds of the step loop alone (set-up and the initial energy excluded) and potential sweeps in it
	double loop_seconds = 0;
	long long loop_sweeps = 0, pi_sweeps = 0;
	void set_sharding(int r, int nr, const char id[128]) { rank = r; nranks = nr; memcpy(nccl_id, id, 128); }
	// file names per system (check_io_files_options, src/SimulationControl.cpp:2196-2360)
	std::vector<std::string> pqr_input_filenames, pqr_restart_filenames, pqr_final_filenames;
	void check_io_files_options();
	static std::string make_filename(const char *basename, int fileno);   // src/Output.cpp:46-92

	// path integrals (src/SimulationControl.PathIntegral.cpp)
	bool PI_nvt_mc(std::vector<System::step_record> *log = nullptr);
	double PI_calculate_energy();
	double PI_calculate_potential();
	double PI_calculate_kinetic();
	double PI_chain_mass_length2_ENTIRE_SYSTEM();
	double PI_chain_mass_length2();
	double PI_chain_mass_length2(std::vector<Molecule *> &m);
	int PI_pick_NVT_move();
	void PI_make_move(int move);
	void PI_displace();
	void PI_perturb_beads();
	void PI_perturb_bead_COMs_ENTIRE_SYSTEM();
	void PI_perturb_bead_COMs();
	void PI_perturb_bead_COMs(int n);
	void restore_PI_systems();
	double PI_NVT_boltzmann_factor(double d_potential, double d_chain, double d_orient, int movetype);
	// orientational degree of freedom of a diatomic sorbate (sorbate_orientation_site / sorbate_bondlength / sorbate_reducedMass)
	struct molecular_metadata { int orientation_site; double bond_length, reduced_mass; };   // src/SimulationControl.h:66-74
	std::vector<molecular_metadata> sorbate_data;          // (static in the reference: one input file per process either way)
	std::map<std::string, uint32_t> sorbate_data_index;
	std::vector<double> orientations;                      // 3 per bead system, kept from one call to the next like the reference's
	void add_orientation_site_entry(const char *id, int site_idx);
	void add_bond_length_entry(const char *id, double bond_length);
	void add_reduced_mass_entry(const char *id, double reduced_mass);
	int get_orientation_site(const std::string &molecule_id);
	double get_bond_length(const std::string &molecule_id);
	double get_reduced_mass(const std::string &molecule_id);
	double PI_orientational_mu_length2();
	void PI_calc_system_mass();                                       // :833-837
	void average_current_observables_into_PI_avgObservables();        // :211-232
	void PI_perturb_beads_orientations();
	void generate_orientation_configs();
	void generate_orientation_configs(unsigned int start, unsigned int end, unsigned int p, unsigned int numBeads, double b2, double ukT);
	void apply_orientation_configs();

private:
	void read_config(const char *inFilename);
	bool process_command(const std::vector<std::string> &token);
	void check_system();
	void initialize_PI_NVT_Systems();
	int starterBead = 0;             // the function-static of PI_perturb_bead_COMs(int) (PathIntegral.cpp:1480)
	mpmc_engine *pi_gpu = nullptr;   // one batched engine for this rank's bead systems
	std::vector<int> pi_mol_first;   // list position of a molecule -> its first site (+ total at the end)
	std::vector<int> pi_dirty;       // list positions whose coordinates changed since the device last saw them
	int pi_target_pos = 0;
	std::vector<std::vector<Molecule *>> pi_mols;   // [bead system][list position]
	std::vector<int> pi_movable;     // list positions the Markov chain may pick
	std::vector<double> pi_chain_term;              // per molecule: PI_chain_mass_length2 of its chain, as last computed
	std::vector<int> pi_chain_stale;                // list positions whose term must be recomputed
	void pi_index_build();
	std::vector<std::vector<double>> pi_backup;     // per bead system: mass, COM and site coordinates of the molecule picked for the move
	std::vector<Atom *> pi_walk;                    // scratch of PI_pick_NVT_move's level-by-level prefetch
	int pi_N = -1;                                  // molecules the Markov chain may move (constant over a path-integral run)
	double pi_last_means[4] = {0, 0, 0, 0};         // the last sweep's bead means (reused when nothing has moved since)
	bool pi_have_means = false;
};

} // namespace mpmc_host
