// sim_control.cpp — SimulationControl mirror: the input keywords that parameterise the hot path and its two callers
// (src/SimulationControl.cpp:258-1616 subset), and the path-integral Markov chain (src/SimulationControl.PathIntegral.cpp).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <sstream>

#include "mpmc_host.h"

namespace mpmc_host {

static bool ieq(const std::string &a, const char *b) {
	size_t n = strlen(b);
	if (a.size() != n) return false;
	for (size_t i = 0; i < n; i++) if (tolower((unsigned char)a[i]) != tolower((unsigned char)b[i])) return false;
	return true;
}
static double num(const std::string &s) {
	size_t idx = 0; double d;
	try { d = std::stod(s, &idx); } catch (...) { throw invalid_input; }
	if (idx != s.size()) throw invalid_input;
	return d;
}
static int onoff(const std::string &s) {
	if (ieq(s, "on")) return 1;
	if (ieq(s, "off")) return 0;
	throw invalid_input;
}

SimulationControl::SimulationControl(const char *inFilename, int P) : nSys(P) {
	orientations.assign(3 * (size_t)std::max(P, 0), 0.0);      // one orientation vector per bead (src/SimulationControl.cpp:74)
	read_config(inFilename);
	check_system();
}

SimulationControl::~SimulationControl() {
	if (pi_gpu) mpmc_destroy(pi_gpu);
	for (System *s : systems) delete s;
}

void SimulationControl::read_config(const char *inFilename) {
	std::ifstream in(inFilename);
	if (!in) throw fopen_fail_read;
	std::string line;
	while (std::getline(in, line)) {
		std::istringstream ss(line);
		std::vector<std::string> t;
		for (std::string w; ss >> w;) t.push_back(w);
		if (t.empty() || t[0][0] == '!' || t[0][0] == '#') continue;
		if (!process_command(t)) throw invalid_input;
	}
}

// the keywords that reach energy() and the two MC loops; everything else the reference knows is either accepted and ignored
// here (output/bookkeeping options) or refused (physics this engine does not implement)
bool SimulationControl::process_command(const std::vector<std::string> &t) {
	const std::string &k = t[0];
	auto arg = [&](size_t i) -> const std::string & { if (t.size() <= i) throw invalid_input; return t[i]; };
	if (ieq(k, "job_name")) { strncpy(sys.job_name, arg(1).c_str(), sizeof sys.job_name - 1); return true; }
	if (ieq(k, "ensemble")) {
		if (ieq(arg(1), "nvt")) sys.ensemble = ENSEMBLE_NVT;
		else if (ieq(arg(1), "uvt")) sys.ensemble = ENSEMBLE_UVT;
		else if (ieq(arg(1), "pi_nvt")) sys.ensemble = ENSEMBLE_PATH_INTEGRAL_NVT;
		else {
			// ensembles the reference knows and this mirror does not drive; any other word is a malformed line (:280-301 -> invalid_input)
			for (const char *w : {"surf", "surf_fit", "nve", "total_energy", "npt", "replay", "nvt_gibbs"}) if (ieq(arg(1), w)) throw unsupported_setting;
			return false;
		}
		return true;
	}
	if (ieq(k, "seed")) { sys.preset_seed = (unsigned int)std::stoul(arg(1)); sys.preset_seed_on = 1; return true; }
	if (ieq(k, "numsteps")) { sys.numsteps = (uint32_t)num(arg(1)); return true; }
	if (ieq(k, "corrtime")) { sys.corrtime = (uint32_t)num(arg(1)); return true; }
	if (ieq(k, "move_factor")) { sys.move_factor = num(arg(1)); return true; }
	if (ieq(k, "rot_factor")) { sys.rot_factor = num(arg(1)); return true; }
	if (ieq(k, "insert_probability")) { sys.insert_probability = num(arg(1)); return true; }
	if (ieq(k, "bead_perturb_probability")) { sys.bead_perturb_probability = num(arg(1)); return true; }
	if (ieq(k, "PI_trial_chain_length")) { PI_trial_chain_length = (int)num(arg(1)); return true; }
	// <molecule type> <value>  (src/SimulationControl.cpp:306-339)
	if (ieq(k, "sorbate_orientation_site")) { add_orientation_site_entry(arg(1).c_str(), (int)num(arg(2))); return true; }
	if (ieq(k, "sorbate_bondlength")) { add_bond_length_entry(arg(1).c_str(), num(arg(2))); return true; }
	if (ieq(k, "sorbate_reducedMass")) { add_reduced_mass_entry(arg(1).c_str(), num(arg(2))); return true; }
	if (ieq(k, "temperature")) { sys.temperature = num(arg(1)); return true; }
	if (ieq(k, "pressure")) { sys.pressure = num(arg(1)); return true; }
	if (ieq(k, "free_volume")) { sys.free_volume = num(arg(1)); return true; }
	if (ieq(k, "scale_charge")) { sys.scale_charge = num(arg(1)); return true; }
	if (ieq(k, "basis1") || ieq(k, "basis2") || ieq(k, "basis3")) {
		const int r = k[5] - '1';
		for (int j = 0; j < 3; j++) sys.pbc.basis[r][j] = num(arg(1 + j));
		return true;
	}
	if (ieq(k, "pqr_input")) { strncpy(sys.pqr_input, arg(1).c_str(), sizeof sys.pqr_input - 1); return true; }
	if (ieq(k, "pqr_output")) { strncpy(sys.pqr_output, arg(1).c_str(), sizeof sys.pqr_output - 1); return true; }
	if (ieq(k, "pqr_restart")) { strncpy(sys.pqr_restart, arg(1).c_str(), sizeof sys.pqr_restart - 1); return true; }
	if (ieq(k, "long_output")) { sys.long_output = onoff(arg(1)); return true; }
	if (ieq(k, "rd_only")) { sys.rd_only = onoff(arg(1)); return true; }
	if (ieq(k, "rd_lrc")) { sys.rd_lrc = onoff(arg(1)); return true; }
	if (ieq(k, "wrapall")) { sys.wrapall = onoff(arg(1)); return true; }
	if (ieq(k, "parallel_restarts")) { sys.parallel_restarts = onoff(arg(1)); return true; }
	if (ieq(k, "read_pqr_box")) { sys.read_pqr_box_on = onoff(arg(1)); return true; }
	// (the reference rejects these two by name, :806-813)
	if (ieq(k, "move_probability") || ieq(k, "rot_probability")) return false;
	// insertions from a list of candidate molecules: not driven by this mirror
	if (ieq(k, "insert_input")) throw unsupported_setting;
	if (ieq(k, "cuda")) { sys.cuda = onoff(arg(1)); return true; }
	if (ieq(k, "ewald_alpha")) { sys.ewald_alpha = num(arg(1)); sys.ewald_alpha_set = 1; return true; }
	if (ieq(k, "ewald_kmax")) { sys.ewald_kmax = (int)num(arg(1)); return true; }
	if (ieq(k, "polarization")) { sys.polarization = onoff(arg(1)); return true; }
	if (ieq(k, "polar_ewald")) { sys.polar_ewald = onoff(arg(1)); return true; }
	if (ieq(k, "polar_ewald_alpha")) { sys.polar_ewald_alpha = num(arg(1)); sys.polar_ewald_alpha_set = 1; return true; }
	if (ieq(k, "polar_iterative")) { sys.polar_iterative = onoff(arg(1)); return true; }
	if (ieq(k, "polar_zodid")) { sys.polar_zodid = onoff(arg(1)); return true; }
	if (ieq(k, "polar_palmo")) { sys.polar_palmo = onoff(arg(1)); return true; }
	if (ieq(k, "polar_gs")) { sys.polar_gs = onoff(arg(1)); return true; }
	if (ieq(k, "polar_gs_ranked")) { sys.polar_gs_ranked = onoff(arg(1)); return true; }
	if (ieq(k, "polar_sor")) { sys.polar_sor = onoff(arg(1)); return true; }
	if (ieq(k, "polar_esor")) { sys.polar_esor = onoff(arg(1)); return true; }
	if (ieq(k, "polar_rrms")) { sys.polar_rrms = onoff(arg(1)); return true; }
	if (ieq(k, "polar_gamma")) { sys.polar_gamma = num(arg(1)); return true; }
	if (ieq(k, "polar_damp")) { sys.polar_damp = num(arg(1)); return true; }
	if (ieq(k, "polar_precision")) { sys.polar_precision = num(arg(1)); return true; }
	if (ieq(k, "polar_max_iter")) { sys.polar_max_iter = (int)num(arg(1)); return true; }
	if (ieq(k, "polar_damp_type")) {
		if (ieq(arg(1), "none") || ieq(arg(1), "off")) sys.damp_type = DAMPING_OFF;
		else if (ieq(arg(1), "linear")) sys.damp_type = DAMPING_LINEAR;
		else if (ieq(arg(1), "exponential")) sys.damp_type = DAMPING_EXPONENTIAL;
		else return false;
		return true;
	}
	// physics switches this engine refuses unless they are off
	for (const char *w : {"feynman_hibbs", "h2_fugacity", "co2_fugacity", "ch4_fugacity", "n2_fugacity", "wolf", "sg", "dreiding", "spectre", "gwp", "cavity_bias",
	                      "rd_anharmonic", "polar_ewald_full", "polar_wolf", "polar_wolf_full", "quantum_rotation", "simulated_annealing", "waldmanhagler",
	                      "halgren_mixing", "c6_mixing", "axilrod_teller", "disp_expansion", "lj_buffered_14_7", "cdvdw", "rd_crystal"})
		if (ieq(k, w)) { if (onoff(arg(1))) throw unsupported_setting; return true; }
	// bookkeeping / output options: accepted, not used on this path
	for (const char *w : {"pop_histogram", "traj_output", "energy_output", "energy_output_csv", "dipole_output", "field_output",
	                      "frozen_output", "pop_histogram_output", "pop_hist_resolution", "traj_input",
	                      "max_bondlength", "calc_pressure"})
		if (ieq(k, w)) return true;
	return false;
}

// A job that fails the reference's validation (check_system / check_mc_options, src/SimulationControl.cpp:1617-1850: steps,
// correlation time, temperature, ...) makes its constructor throw invalid_input (:67-72); the path-integral checks throw their own codes.
void SimulationControl::check_system() {
	if (sys.numsteps < 1 || sys.corrtime < 1 || sys.temperature <= 0) throw invalid_input;
	check_io_files_options();
	if (sys.ensemble == ENSEMBLE_PATH_INTEGRAL_NVT) {      // check_PI_options, PathIntegral.cpp:552-606
		int bits = 0;
		for (unsigned v = (unsigned)nSys; v; v >>= 1) bits += v & 1;
		if (nSys < 4 || bits != 1) throw invalid_MPI_size_for_PI;
		if (!PI_trial_chain_length || PI_trial_chain_length >= nSys) throw invalid_setting;
	}
	if (sys.ensemble == ENSEMBLE_UVT && sys.pressure <= 0) throw invalid_input;
}

// Output::make_filename (src/Output.cpp:46-92): "<base>-%04d<.ext>" when the name ends in a three-character extension, else
// "<base>-%04d"; /dev/null stays /dev/null
std::string SimulationControl::make_filename(const char *basename, int fileno) {
	if (!strncmp("/dev/null", basename, 9)) return "/dev/null";
	const size_t len = strlen(basename);
	char num[32];
	snprintf(num, sizeof num, "-%04d", fileno);
	if (len > 4 && basename[len - 4] == '.') return std::string(basename, len - 4) + num + std::string(basename + len - 4);
	return std::string(basename) + num;
}

// The restart / final / input file names of every system (src/SimulationControl.cpp:2196-2360, single-process branches).  With
// `parallel_restarts on` system j restarts from "<restart>-000j<.ext>" when that file exists, else from the ".last" copy — the reference
// builds that name from its rank, which is 0 in a single process, so it is system 0's for every j — else from pqr_input.
void SimulationControl::check_io_files_options() {
	const int file_count = std::max(nSys, 1);
	if (ieq(sys.pqr_restart, "off")) strcpy(sys.pqr_restart, "/dev/null");
	else if (!sys.pqr_restart[0]) { strncpy(sys.pqr_restart, sys.job_name, sizeof sys.pqr_restart - 16); strcat(sys.pqr_restart, ".restart.pqr"); }
	if (ieq(sys.pqr_output, "off")) strcpy(sys.pqr_output, "/dev/null");
	else if (!sys.pqr_output[0]) { strncpy(sys.pqr_output, sys.job_name, sizeof sys.pqr_output - 16); strcat(sys.pqr_output, ".final.pqr"); }
	pqr_restart_filenames.clear(); pqr_final_filenames.clear(); pqr_input_filenames.clear();
	for (int j = 0; j < file_count; j++) {
		pqr_restart_filenames.push_back(file_count > 1 ? make_filename(sys.pqr_restart, j) : std::string(sys.pqr_restart));
		pqr_final_filenames.push_back(file_count > 1 ? make_filename(sys.pqr_output, j) : std::string(sys.pqr_output));
	}
	auto exists = [](const std::string &f) { FILE *t = fopen(f.c_str(), "r"); if (t) fclose(t); return t != nullptr; };
	if (sys.parallel_restarts) {
		if (file_count > 1)
			for (int j = 0; j < file_count; j++) {
				std::string filename = make_filename(sys.pqr_restart, j);
				if (!exists(filename)) {
					filename = make_filename(sys.pqr_restart, 0) + ".last";
					if (!exists(filename) && !sys.pqr_input[0]) filename = std::string(sys.job_name) + ".initial.pqr";
				}
				pqr_input_filenames.push_back(filename);
			}
	} else if (!sys.pqr_input[0]) { strncpy(sys.pqr_input, sys.job_name, sizeof sys.pqr_input - 16); strcat(sys.pqr_input, ".initial.pqr"); }
}

void SimulationControl::initializeSimulationObjects() {      // src/SimulationControl.cpp:80-199
	if (!sys.preset_seed_on) throw missing_setting;          // trajectories are only comparable with a preset seed
	Rando::seed(sys.preset_seed);
	if (sys.ensemble == ENSEMBLE_PATH_INTEGRAL_NVT) { initialize_PI_NVT_Systems(); return; }
	sys.read_molecules(sys.pqr_input);
	if (sys.read_pqr_box_on) sys.read_pqr_box(sys.pqr_input);
	sys.update_pbc();
	sys.mt_rand.seed(sys.preset_seed);
}

bool SimulationControl::runSimulation(std::vector<System::step_record> *log) {
	if (sys.ensemble == ENSEMBLE_PATH_INTEGRAL_NVT) return PI_nvt_mc(log);
	return sys.mc(log);
}

// ------------------------------------------------------------------------------------------------------------
// path integrals
// ------------------------------------------------------------------------------------------------------------
void SimulationControl::initialize_PI_NVT_Systems() {        // PathIntegral.cpp:611-694
	for (int i = 0; i < nSys; i++) {
		System *s = new System(sys);
		if (sys.parallel_restarts && i < (int)pqr_input_filenames.size()) strncpy(s->pqr_input, pqr_input_filenames[i].c_str(), sizeof s->pqr_input - 1);
		strncpy(s->pqr_output, pqr_final_filenames[i].c_str(), sizeof s->pqr_output - 1);
		strncpy(s->pqr_restart, pqr_restart_filenames[i].c_str(), sizeof s->pqr_restart - 1);
		s->read_molecules(s->pqr_input);
		if (s->read_pqr_box_on) s->read_pqr_box(s->pqr_input);
		s->update_pbc();
		for (Molecule *m = s->molecules; m; m = m->next) m->update_COM();
		systems.push_back(s);
	}
	sys.pbc = systems[0]->pbc;
}

// energy() of every bead system in one batched engine call; means over beads (PathIntegral.cpp:752-805).  With nranks > 1 the bead
// systems are sharded over the ranks' GPUs (every rank keeps all P systems on the host and replays the same random stream, as the
// reference's MPI ranks do); the engine sums the per-bead energies across ranks (replaces MPI_Allgather x 4, :763-766).
// Only what moved since the last call is sent: the Markov chain alters one molecule (the same list position in every bead system)
// per step, so the driver records the list positions it touched (pi_dirty) instead of flattening and comparing P x N sites.
double SimulationControl::PI_calculate_potential() {
	const int P = nSys;
	const int b_lo = (int)((long long)P * rank / nranks), b_hi = (int)((long long)P * (rank + 1) / nranks), PL = b_hi - b_lo;
	if (PL < 1) throw invalid_MPI_size_for_PI;
	int rc;
	bool nothing_moved = false;
	if (!pi_gpu) {
		const int n = systems[0]->countNatoms();
		std::vector<double> pos, q, al, ep, sg, ms;
		std::vector<int> mol, fz;
		pos.reserve((size_t)PL * n * 3);
		for (int s = 0; s < P; s++) {
			std::vector<double> p1, q1, a1, e1, s1, m1;
			std::vector<int> mo1, f1;
			systems[s]->flatten(s >= b_lo && s < b_hi ? pos : p1, q1, a1, e1, s1, m1, mo1, f1);
			if (s == 0) { q.swap(q1); al.swap(a1); ep.swap(e1); sg.swap(s1); ms.swap(m1); mol.swap(mo1); fz.swap(f1); }
			else if ((int)q1.size() != n) throw 6005;          // incongruent_bead_states
		}
		// list position -> first site, once (path-integral runs neither insert nor remove molecules)
		pi_mol_first.clear();
		int at = 0;
		for (Molecule *m = systems[0]->molecules; m; m = m->next) { pi_mol_first.push_back(at); at += m->natoms(); }
		pi_mol_first.push_back(at);
		mpmc_config c;
		systems[0]->fill_config(c, PL);
		if ((rc = mpmc_create(&c, &pi_gpu))) throw rc;
		if ((rc = mpmc_upload_sites(pi_gpu, n, pos.data(), q.data(), al.data(), ep.data(), sg.data(), ms.data(), mol.data(), fz.data()))) throw rc;
		if (nranks > 1 && (rc = mpmc_nccl_init(pi_gpu, nccl_id, rank, nranks))) throw rc;
		pi_dirty.clear();
	} else {
		nothing_moved = pi_dirty.empty();
		std::sort(pi_dirty.begin(), pi_dirty.end());
		pi_dirty.erase(std::unique(pi_dirty.begin(), pi_dirty.end()), pi_dirty.end());
		for (int lp : pi_dirty) {
			const int first = pi_mol_first[lp], cnt = pi_mol_first[lp + 1] - first;
			std::vector<double> seg((size_t)PL * cnt * 3);
			pi_index_build();
			for (int s = b_lo; s < b_hi; s++) {
				Molecule *m = pi_mols[s][lp];
				double *o = &seg[(size_t)(s - b_lo) * cnt * 3];
				int k = 0;
				for (Atom *a = m->atoms; a; a = a->next, k++) { if (k >= cnt) throw internal_error; o[3 * k] = a->pos[0]; o[3 * k + 1] = a->pos[1]; o[3 * k + 2] = a->pos[2]; }
				if (k != cnt) throw internal_error;
			}
			if ((rc = mpmc_update_sites_all_beads(pi_gpu, first, cnt, seg.data()))) throw rc;
		}
		pi_dirty.clear();
	}
	// Nothing moved since the last sweep (the reference recomputes the energy of a move it has just accepted, :148: same
	// coordinates, same answer): the device holds exactly what it was asked about last time, so its answer is reused.  Every rank
	// sees the same list of moved molecules, so every rank skips the same sweeps.
	double means[4], U;
	if (nothing_moved && pi_have_means) memcpy(means, pi_last_means, sizeof means);
	else {
		pi_sweeps++;
		if ((rc = mpmc_pi_potential_allreduce(pi_gpu, P, means, &U))) throw rc;
		memcpy(pi_last_means, means, sizeof means);
		pi_have_means = true;
	}
	System::observables_t *obs = sys.observables;
	obs->rd_energy = means[0]; obs->coulombic_energy = means[1]; obs->polarization_energy = means[2]; obs->vdw_energy = means[3];
	return obs->rd_energy + obs->coulombic_energy + obs->vdw_energy + obs->polarization_energy;
}

double SimulationControl::PI_calculate_kinetic() {            // :810-828
	// (countN() walks the molecule list: a path-integral run neither inserts nor removes, so the count is taken once)
	if (pi_N < 0) pi_N = (int)systems[0]->countN(); else systems[0]->observables->N = pi_N;
	const double d = 3.0, N = (double)pi_N, P = (double)nSys, T = sys.temperature;
	const double beta = 1.0 / (kB * T), omega2 = P / (beta * beta * hBar2);
	const double chain_mass_len2 = PI_chain_mass_length2_ENTIRE_SYSTEM();
	const double term_1 = 0.5 * d * N * kB * T * P, term_2 = 0.5 * omega2 * chain_mass_len2;
	sys.observables->kinetic_energy = (1.0 / kB) * (term_1 - term_2);
	return sys.observables->kinetic_energy;
}

double SimulationControl::PI_calculate_energy() {             // :734-748
	const double kinetic = PI_calculate_kinetic();
	const double potential = PI_calculate_potential();
	sys.observables->energy = kinetic + potential;
	return sys.observables->energy;
}

// The sum over every mobile molecule of its chain's mass-weighted squared length (:859-904), in list order.  A molecule's term
// depends on that molecule's bead coordinates only, so the terms are kept and only those of molecules altered since the last call
// are recomputed: the same doubles added in the same order as the reference's full loop, without walking P x M molecules per step.
double SimulationControl::PI_chain_mass_length2_ENTIRE_SYSTEM() {   // :859-904
	pi_index_build();
	const int M = (int)pi_mols[0].size();
	if ((int)pi_chain_term.size() != M) { pi_chain_term.assign(M, 0.0); pi_chain_stale.clear(); for (int lp = 0; lp < M; lp++) pi_chain_stale.push_back(lp); }
	std::sort(pi_chain_stale.begin(), pi_chain_stale.end());
	pi_chain_stale.erase(std::unique(pi_chain_stale.begin(), pi_chain_stale.end()), pi_chain_stale.end());
	std::vector<Molecule *> ptr(nSys);
	for (int lp : pi_chain_stale) {
		Molecule *m0 = pi_mols[0][lp];
		if (m0->frozen || m0->adiabatic || m0->target) { pi_chain_term[lp] = 0.0; continue; }
		for (int s = 0; s < nSys; s++) ptr[s] = pi_mols[s][lp];
		pi_chain_term[lp] = PI_chain_mass_length2(ptr);
	}
	pi_chain_stale.clear();
	double sum = 0;
	for (int lp = 0; lp < M; lp++) {
		Molecule *m0 = pi_mols[0][lp];
		if (!(m0->frozen || m0->adiabatic || m0->target)) sum += pi_chain_term[lp];
	}
	return sum;
}

// list position -> molecule, per bead system (path-integral runs neither insert nor remove molecules; restore() swaps one object)
void SimulationControl::pi_index_build() {
	if ((int)pi_mols.size() == nSys && !pi_mols[0].empty()) return;
	pi_mols.assign(nSys, {});
	for (int s = 0; s < nSys; s++) for (Molecule *m = systems[s]->molecules; m; m = m->next) pi_mols[s].push_back(m);
	pi_movable.clear();
	for (int lp = 0; lp < (int)pi_mols[0].size(); lp++) {
		Molecule *m = pi_mols[0][lp];
		if (!(m->frozen || m->adiabatic || m->target)) pi_movable.push_back(lp);
	}
}

double SimulationControl::PI_chain_mass_length2() {           // :905-915
	std::vector<Molecule *> mol;
	for (System *s : systems) {
		if (!s->checkpoint->molecule_altered) throw internal_error;
		mol.push_back(s->checkpoint->molecule_altered);
	}
	return PI_chain_mass_length2(mol);
}

double SimulationControl::PI_chain_mass_length2(std::vector<Molecule *> &molecule) {   // :916-970
	std::vector<double> c(3 * (size_t)nSys);
	for (int s = 0; s < nSys; s++) {
		molecule[s]->update_COM();
		for (int p = 0; p < 3; p++) c[3 * s + p] = molecule[s]->com[p];
	}
	double len2 = 0;
	for (int i = 0; i < nSys; i++) {
		const int j = (i + 1) % nSys;
		const double dx = c[3 * i] - c[3 * j], dy = c[3 * i + 1] - c[3 * j + 1], dz = c[3 * i + 2] - c[3 * j + 2];
		len2 += dx * dx + dy * dy + dz * dz;
	}
	len2 *= (molecule[0]->mass * AMU2KG) * (ANGSTROM2METER * ANGSTROM2METER);
	return len2;
}

// two uniforms from Rando: move type, then the target index (the same index in every bead system) (:1047-1116)
int SimulationControl::PI_pick_NVT_move() {
	const double dice_move = Rando::rand(), dice_target = Rando::rand();
	pi_index_build();
	if ((int)pi_backup.size() != nSys) pi_backup.assign(nSys, {});
	if (pi_movable.empty()) throw no_molecules_in_system;
	const int lp = pi_movable[(int)std::floor(pi_movable.size() * dice_target)];
	pi_target_pos = lp;
	pi_dirty.push_back(lp);
	pi_chain_stale.push_back(lp);
	// The picked molecule's images live in nSys different lists, none of them in cache (the bead systems of config 5 hold 80 MB of
	// list nodes): walked one after the other that is ~400 dependent cache misses, half of the driver's own time per move.  Fetch
	// them level by level instead — all images' molecule nodes, then all first atoms, ... — so that the misses of a level overlap.
	{
		pi_walk.resize(nSys);
		auto fetch = [](const void *p, size_t bytes) { for (size_t o = 0; o < bytes; o += 64) __builtin_prefetch((const char *)p + o); };
		for (int s = 0; s < nSys; s++) fetch(pi_mols[s][lp], sizeof(Molecule));
		for (int s = 0; s < nSys; s++) pi_walk[s] = pi_mols[s][lp]->atoms;
		for (bool any = true; any;) {
			any = false;
			for (int s = 0; s < nSys; s++) if (pi_walk[s]) fetch(pi_walk[s], 64);      // link, mass, position: the first line (mpmc_host.h)
			for (int s = 0; s < nSys; s++) if (pi_walk[s]) { pi_walk[s] = pi_walk[s]->next; any = any || pi_walk[s]; }
		}
	}
	for (int s = 0; s < nSys; s++) {
		System *S = systems[s];
		Molecule *m = pi_mols[s][lp];
		S->checkpoint->molecule_altered = m;
		S->checkpoint->movetype = (dice_move < sys.bead_perturb_probability) ? MOVETYPE_PERTURB_BEADS : MOVETYPE_DISPLACE;
		S->checkpoint->head = lp > 0 ? pi_mols[s][lp - 1] : nullptr;
		S->checkpoint->tail = m->next;
		// The reference backs the molecule up as a deep copy and, on rejection, links the copy into the list in place of the altered
		// object (System::restore).  A path-integral move alters nothing but the sites' coordinates and the molecule's mass / COM
		// fields, so the mirror keeps exactly those and writes them back into the same object: no list surgery, and none of the
		// ~400 heap objects a move of 64 five-site images would create and destroy (60 % of the driver's own time per move).
		std::vector<double> &b = pi_backup[s];
		b.clear();
		b.push_back(m->mass);
		for (int p = 0; p < 3; p++) b.push_back(m->com[p]);
		for (Atom *a = m->atoms; a; a = a->next) for (int p = 0; p < 3; p++) b.push_back(a->pos[p]);
	}
	return systems[0]->checkpoint->movetype;
}

void SimulationControl::PI_make_move(int move) {
	switch (move) {
	case MOVETYPE_DISPLACE: PI_displace(); break;
	case MOVETYPE_PERTURB_BEADS: PI_perturb_beads(); break;
	default: throw invalid_monte_carlo_move;
	}
}

// the same translation for every bead, then one rotation of the whole chain about the chain's centre (:1320-1387).
// Rando order: six uniforms, three normals, one uniform (angle = u * rot_factor DEGREES, no factor 360).
void SimulationControl::PI_displace() {
	double dice[6];
	for (int i = 0; i < 6; i++) dice[i] = Rando::rand();
	const int n = (int)systems.size();
	double pc[3] = {0, 0, 0};
	std::vector<Molecule *> alt;
	for (int s = 0; s < n; s++) {
		Molecule *m = systems[s]->checkpoint->molecule_altered;
		alt.push_back(m);
		m->update_COM();
		m->translate_rand_pbc(sys.move_factor, systems[s]->pbc, dice);
		pc[0] = pc[0] + m->com[0]; pc[1] = pc[1] + m->com[1]; pc[2] = pc[2] + m->com[2];
	}
	pc[0] /= n; pc[1] /= n; pc[2] /= n;
	const double dx = Rando::rand_normal(), dy = Rando::rand_normal(), dz = Rando::rand_normal();
	const double angle = Rando::rand() * sys.rot_factor;
	// Quaternion(x, y, z, angle, AXIS_ANGLE_DEGREE) and Quaternion::rotate(v) = (q v) q*   (src/Quaternion.cpp)
	double ang = angle / 57.2957795, X, Y, Z, W;
	{
		double mag = std::sqrt(dx * dx + dy * dy + dz * dz);
		if (mag == 0.0) { X = Y = Z = 0; W = 1; }
		else {
			const double x = dx / mag, y = dy / mag, z = dz / mag, sn = std::sin(ang / 2.0);
			X = x * sn; Y = y * sn; Z = z * sn; W = std::cos(ang / 2.0);
		}
	}
	for (Molecule *m : alt) {
		m->translate(-pc[0], -pc[1], -pc[2]);
		for (Atom *a = m->atoms; a; a = a->next) {
			const double vx = a->pos[0], vy = a->pos[1], vz = a->pos[2], vw = 0;
			// t = q * v
			const double tw = W * vw - X * vx - Y * vy - Z * vz;
			const double tx = W * vx + X * vw + Y * vz - Z * vy;
			const double ty = W * vy - X * vz + Y * vw + Z * vx;
			const double tz = W * vz + X * vy - Y * vx + Z * vw;
			// r = t * conj(q)
			const double cx = -X, cy = -Y, cz = -Z, cw = W;
			a->pos[0] = tw * cx + tx * cw + ty * cz - tz * cy;
			a->pos[1] = tw * cy - tx * cz + ty * cw + tz * cx;
			a->pos[2] = tw * cz + tx * cy - ty * cx + tz * cw;
		}
		m->translate(pc[0], pc[1], pc[2]);
		m->update_COM();
	}
}

void SimulationControl::PI_perturb_beads() {       // :1391-1396
	PI_perturb_beads_orientations();
	PI_perturb_bead_COMs();
}

// ---- orientational degree of freedom of a diatomic sorbate -------------------------------------------------------------------
// One metadata record per molecule type, created by whichever of the three keywords names the type first
// (src/SimulationControl.cpp:2976-3071).
static SimulationControl::molecular_metadata &sorbate_entry(SimulationControl &sc, const char *id) {
	auto it = sc.sorbate_data_index.find(id);
	if (it == sc.sorbate_data_index.end()) {
		sc.sorbate_data.push_back({-1, 0.0, 0.0});
		it = sc.sorbate_data_index.insert(std::make_pair(std::string(id), (uint32_t)sc.sorbate_data.size() - 1)).first;
	}
	return sc.sorbate_data[it->second];
}
void SimulationControl::add_orientation_site_entry(const char *id, int site_idx) { sorbate_entry(*this, id).orientation_site = site_idx; }
void SimulationControl::add_bond_length_entry(const char *id, double bond_length) { sorbate_entry(*this, id).bond_length = bond_length; }
void SimulationControl::add_reduced_mass_entry(const char *id, double reduced_mass) { sorbate_entry(*this, id).reduced_mass = reduced_mass; }

// NOTE the reference returns the INDEX of the type's metadata record, not the configured site (:2996-3004): with one sorbate type
// configured the "orientation site" is atom 0 whatever the input said, with two types it is atom 0 for the first and atom 1 for the
// second.  Reproduced, because the trajectories are compared with the reference's.
int SimulationControl::get_orientation_site(const std::string &molecule_id) {
	auto it = sorbate_data_index.find(molecule_id);
	return it == sorbate_data_index.end() ? -1 : (int)it->second;
}
double SimulationControl::get_bond_length(const std::string &molecule_id) {
	auto it = sorbate_data_index.find(molecule_id);
	return it == sorbate_data_index.end() ? 0 : sorbate_data[it->second].bond_length;
}
double SimulationControl::get_reduced_mass(const std::string &molecule_id) {
	auto it = sorbate_data_index.find(molecule_id);
	return it == sorbate_data_index.end() ? -1.0 : sorbate_data[it->second].reduced_mass;
}

// Sum over the bead ring of |b_s - b_{s+1}|^2, b_s = bond_length * unit(COM -> handle site) of the moved molecule's image in bead
// system s, in m^2 (:978-1039).  0 unless a site record and a positive bond length exist for the molecule's type.
double SimulationControl::PI_orientational_mu_length2() {
	const char *moleculeID = systems[0]->checkpoint->molecule_altered->moleculetype;
	const int orientation_site = get_orientation_site(moleculeID);
	const double bond_length = get_bond_length(moleculeID);
	if (orientation_site < 0 || bond_length <= 0) return 0.0;
	std::vector<double> bond(3 * (size_t)nSys);
	for (int s = 0; s < nSys; s++) {
		Molecule *m = systems[s]->checkpoint->molecule_altered;
		m->update_COM();
		Atom *a = m->atoms;
		for (int site = 0; site != orientation_site; site++) a = a->next;
		double x = a->pos[0] - m->com[0], y = a->pos[1] - m->com[1], z = a->pos[2] - m->com[2];
		const double mag = std::sqrt(x * x + y * y + z * z);
		if (mag != 0) { x = x / mag; y = y / mag; z = z / mag; } else x = y = z = 0;
		bond[3 * s] = bond_length * x; bond[3 * s + 1] = bond_length * y; bond[3 * s + 2] = bond_length * z;
	}
	double diff = 0.0;
	for (int i = 0; i < nSys; i++) {
		const int j = (i + 1) % nSys;
		const double dx = bond[3 * i] - bond[3 * j], dy = bond[3 * i + 1] - bond[3 * j + 1], dz = bond[3 * i + 2] - bond[3 * j + 2];
		diff += dx * dx + dy * dy + dz * dz;
	}
	diff *= (ANGSTROM2METER * ANGSTROM2METER);
	return diff;
}

void SimulationControl::PI_perturb_beads_orientations() {     // :1559-1572
	const char *moleculeID = systems[0]->checkpoint->molecule_altered->moleculetype;
	if (get_orientation_site(moleculeID) < 0 || get_bond_length(moleculeID) <= 0) return;
	generate_orientation_configs();
	apply_orientation_configs();
}

namespace {
struct V3 { double x, y, z; };
inline double dot(const V3 &a, const V3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline double norm(const V3 &a) { return std::sqrt(dot(a, a)); }
inline V3 unit(V3 a) { const double m = norm(a); if (m != 0) return {a.x / m, a.y / m, a.z / m}; return {0, 0, 0}; }
inline V3 cross(const V3 &a, const V3 &b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
// Quaternion(axis, angle, AXIS_ANGLE_RADIAN).rotate(v) = (q v) q*   (src/Quaternion.cpp)
inline V3 rotate_about(const V3 &axis, double angle, const V3 &v) {
	double X = 0, Y = 0, Z = 0, W = 1;
	const double mag = std::sqrt(axis.x * axis.x + axis.y * axis.y + axis.z * axis.z);
	if (mag != 0.0) {
		const double x = axis.x / mag, y = axis.y / mag, z = axis.z / mag, sn = std::sin(angle / 2.0);
		X = x * sn; Y = y * sn; Z = z * sn; W = std::cos(angle / 2.0);
	}
	const double vw = 0;
	const double tw = W * vw - X * v.x - Y * v.y - Z * v.z;
	const double tx = W * v.x + X * vw + Y * v.z - Z * v.y;
	const double ty = W * v.y - X * v.z + Y * vw + Z * v.x;
	const double tz = W * v.z + X * v.y - Y * v.x + Z * vw;
	const double cx = -X, cy = -Y, cz = -Z, cw = W;
	return {tw * cx + tx * cw + ty * cz - tz * cy, tw * cy - tx * cz + ty * cw + tz * cx, tw * cz + tx * cy - ty * cx + tz * cw};
}
} // namespace

// Rando order: three normals for bead 0's orientation, then per placed bead one uniform (C) and one uniform (beta)  (:1577-1680)
void SimulationControl::generate_orientation_configs() {
	const char *moleculeID = systems[0]->checkpoint->molecule_altered->moleculetype;
	const double sorbate_reduced_mass = get_reduced_mass(moleculeID);
	if (sorbate_reduced_mass < 0) throw missing_required_datum;
	double sorbate_bond_length = get_bond_length(moleculeID);
	if (sorbate_bond_length < 0) throw missing_required_datum;
	sorbate_bond_length /= METER2ANGSTROM;
	const double b2 = sorbate_bond_length * sorbate_bond_length;
	const double u_kB_T = sorbate_reduced_mass * kB * sys.temperature;
	V3 o0;
	o0.x = Rando::rand_normal(); o0.y = Rando::rand_normal(); o0.z = Rando::rand_normal();      // Vector3D::randomize
	o0 = unit(o0);
	orientations[0] = o0.x; orientations[1] = o0.y; orientations[2] = o0.z;
	const unsigned int n = (unsigned int)(orientations.size() / 3);
	generate_orientation_configs(0, n, 2, n, b2, u_kB_T);
}

// Recursive bisection of the bead ring (Subramanian et al., J. Chem. Phys. 146, 094105): bead J halfway between I and K gets the
// normalised mean of their orientations, tilted by alpha drawn from the spring distribution of stiffness K = 4 kh p cos(psi/2)
// about an axis picked uniformly (beta) around that mean.
void SimulationControl::generate_orientation_configs(unsigned int start, unsigned int end, unsigned int p, unsigned int numBeads, double b2, double ukT) {
	const double two_PI = 2.0 * pi;
	if (p > numBeads) return;
	const unsigned int J_idx = (start + end) / 2, K_idx = (end == numBeads) ? 0 : end;
	const V3 vec_I{orientations[3 * start], orientations[3 * start + 1], orientations[3 * start + 2]};
	const V3 vec_K{orientations[3 * K_idx], orientations[3 * K_idx + 1], orientations[3 * K_idx + 2]};
	const V3 bisector = unit({(vec_I.x + vec_K.x) / 2.0, (vec_I.y + vec_K.y) / 2.0, (vec_I.z + vec_K.z) / 2.0});
	V3 vec_IK;
	double psi_IK = 0;
	if (p > 2) {
		vec_IK = {vec_K.x - vec_I.x, vec_K.y - vec_I.y, vec_K.z - vec_I.z};
		psi_IK = std::acos(dot(vec_I, vec_K) / (norm(vec_I) * norm(vec_K)));
	} else {                                         // I == K: any vector orthogonal to the bisector
		const V3 different_vec = unit({1 + bisector.x, 2 + bisector.y, -3 + bisector.z});
		vec_IK = cross(different_vec, bisector);
	}
	const double C = Rando::rand();
	const double lambda2 = h * h / (two_PI * ukT);
	const double kh = pi * b2 / lambda2;
	const double K = 4.0 * kh * p * std::cos(psi_IK * 0.5);
	const double angle_A = std::acos(1.0 + (1.0 / K) * std::log(1.0 - C * (1.0 - std::exp(-2.0 * K))));
	const double angle_B = Rando::rand() * two_PI;
	const V3 vec_Beta = rotate_about(bisector, angle_B, vec_IK);
	const V3 vec_J = rotate_about(vec_Beta, angle_A, bisector);
	orientations[3 * J_idx] = vec_J.x; orientations[3 * J_idx + 1] = vec_J.y; orientations[3 * J_idx + 2] = vec_J.z;
	if (p < numBeads) {
		generate_orientation_configs(start, J_idx, p * 2, numBeads, b2, ukT);
		generate_orientation_configs(J_idx, end, p * 2, numBeads, b2, ukT);
	}
}

void SimulationControl::apply_orientation_configs() {         // :1684-1697
	const int orientation_site = get_orientation_site(systems[0]->checkpoint->molecule_altered->moleculetype);
	if (orientation_site < 0) return;
	for (int s = 0; s < (int)systems.size(); s++) systems[s]->checkpoint->molecule_altered->orient(&orientations[3 * (size_t)s], orientation_site);
}

void SimulationControl::PI_perturb_bead_COMs_ENTIRE_SYSTEM() {      // :1402-1449
	std::vector<Molecule *> ptr(nSys), backup(nSys);
	for (int s = 0; s < nSys; s++) { ptr[s] = systems[s]->molecules; backup[s] = systems[s]->checkpoint->molecule_altered; }
	while (ptr[0]) {
		if (!(ptr[0]->frozen || ptr[0]->adiabatic || ptr[0]->target)) {
			for (int s = 0; s < nSys; s++) { if (!ptr[s]) throw internal_error; systems[s]->checkpoint->molecule_altered = ptr[s]; }
			PI_perturb_bead_COMs(nSys);
		}
		for (int s = 0; s < nSys; s++) ptr[s] = ptr[s]->next;
	}
	for (int s = 0; s < nSys; s++) systems[s]->checkpoint->molecule_altered = backup[s];
	if (pi_gpu) for (int lp = 0; lp + 1 < (int)pi_mol_first.size(); lp++) pi_dirty.push_back(lp);
	pi_chain_term.clear();
}

void SimulationControl::PI_perturb_bead_COMs() { PI_perturb_bead_COMs(PI_trial_chain_length); }

// Coker et al. staging of n consecutive beads between two anchors, then removal of the chain-COM drift (:1453-1554).
// Rando order: three normals per moved bead.  The anchor advances by one bead per call.
void SimulationControl::PI_perturb_bead_COMs(int n) {
	const double beta = 1.0 / (kB * sys.temperature), P = (double)nSys;
	const double Mass = AMU2KG * systems[0]->checkpoint->molecule_altered->mass;
	int prev = starterBead, bead = (prev + 1) % nSys;
	const int last = (prev + n + 1) % nSys;
	starterBead = (starterBead + 1) % nSys;
	std::vector<double> b(3 * (size_t)nSys);
	double cc[3] = {0, 0, 0};
	for (int s = 0; s < nSys; s++) {
		Molecule *m = systems[s]->checkpoint->molecule_altered;
		m->update_COM();
		for (int p = 0; p < 3; p++) { b[3 * s + p] = m->com[p]; cc[p] = cc[p] + m->com[p]; }
	}
	for (int p = 0; p < 3; p++) cc[p] /= P;
	double tB = (double)n, tA = 1.0 + n;
	for (int j = 1; j <= n; j++) {
		const double init_factor = tB-- / tA--;
		const double term_factor = 1.0 - init_factor;
		const double sigma_factor = std::sqrt((hBar2 * beta * init_factor) / (P * Mass)) * METER2ANGSTROM;
		// The reference builds `Vector3D perturbation(rand_normal(), rand_normal(), rand_normal())` (:1523): the order in which a
		// compiler evaluates call arguments is unspecified, and the g++ build that serves as the oracle evaluates them right to
		// left, so the FIRST draw is z.  Reproduced here explicitly.
		const double pz = Rando::rand_normal();
		const double py = Rando::rand_normal();
		const double px = Rando::rand_normal();
		const double pert[3] = {px, py, pz};
		for (int p = 0; p < 3; p++) b[3 * bead + p] = (init_factor * b[3 * prev + p] + term_factor * b[3 * last + p]) + sigma_factor * pert[p];
		prev = (prev + 1) % nSys;
		bead = (prev + 1) % nSys;
	}
	double dc[3] = {0, 0, 0};
	for (int s = 0; s < nSys; s++) for (int p = 0; p < 3; p++) dc[p] = dc[p] + b[3 * s + p];
	for (int p = 0; p < 3; p++) dc[p] = (dc[p] / P) - cc[p];
	for (int s = 0; s < nSys; s++) for (int p = 0; p < 3; p++) b[3 * s + p] -= dc[p];
	for (int s = 0; s < nSys; s++) systems[s]->checkpoint->molecule_altered->move_to_(b[3 * s], b[3 * s + 1], b[3 * s + 2]);
}

void SimulationControl::restore_PI_systems() {
	for (int s = 0; s < nSys; s++) {
		System *S = systems[s];
		S->iterator_failed = 0;
		*S->observables = S->checkpoint->observables;     // System::restore(), first line
		Molecule *m = S->checkpoint->molecule_altered;
		const std::vector<double> &b = pi_backup[s];
		size_t k = 0;
		m->mass = b[k++];
		for (int p = 0; p < 3; p++) m->com[p] = b[k++];
		for (Atom *a = m->atoms; a; a = a->next) for (int p = 0; p < 3; p++) a->pos[p] = b[k++];
		if (k != b.size()) throw internal_error;
	}
	pi_dirty.push_back(pi_target_pos);                  // the device still holds the rejected coordinates
	// (the chain terms need nothing: they were last computed for exactly the coordinates that have just been put back)
}

// The orientational contribution uses the same factor as the COM chain's although its length carries no (reduced) mass — the
// reference reads the reduced mass and does not use it (:520-524) — so with an orientation configured the factor is exp(-/+ ~1e26):
// such a bead move is accepted exactly when the ring of bond vectors got shorter.  Reproduced as it is.
void SimulationControl::PI_calc_system_mass() {
	systems[0]->calc_system_mass();
	sys.observables->frozen_mass = systems[0]->observables->frozen_mass;
	sys.observables->total_mass = systems[0]->observables->total_mass;
}

// The aggregate observables (means over the bead systems, kinetic estimator) averaged the way a classic chain's are; what the
// aggregate does not carry comes from the first bead system, whose molecule list also serves for the sorbate's mass.  (The reference
// takes NU from that bead system's own energy(); the engine evaluates the bead systems as one batch and returns their means, so NU
// — and with it qst — is not accumulated here.)
void SimulationControl::average_current_observables_into_PI_avgObservables() {
	sys.molecules = systems[0]->molecules;
	sys.pbc = systems[0]->pbc;
	sys.observables->N = systems[0]->observables->N;
	sys.observables->volume = systems[0]->observables->volume;
	sys.observables->temperature = systems[0]->observables->temperature;
	sys.observables->spin_ratio = systems[0]->observables->spin_ratio;
	sys.update_root_averages(sys.observables);
	sys.molecules = nullptr;                             // (the list belongs to the bead system)
}

double SimulationControl::PI_NVT_boltzmann_factor(double d_potential, double d_chain, double d_orient, int movetype) {   // :490-547
	const double P = (double)nSys, T = sys.temperature;
	if (movetype == MOVETYPE_PERTURB_BEADS) {
		const double PIchain_2_K = (P * pi * pi * kB * T) / (2.0 * h * h);
		const double potential_contrib = d_potential / T, PI_COM_contrib = d_chain * PIchain_2_K;
		double PI_orientation_contrib = 0;
		if (sorbate_data_index.find(systems[0]->checkpoint->molecule_altered->moleculetype) != sorbate_data_index.end())
			PI_orientation_contrib = d_orient * PIchain_2_K;
		return exp(-potential_contrib - PI_COM_contrib - PI_orientation_contrib);
	}
	return exp(-d_potential / T);
}

bool SimulationControl::PI_nvt_mc(std::vector<System::step_record> *log) {     // :31-196
	for (System *S : systems) { S->observables->temperature = sys.temperature; S->observables->volume = S->pbc.volume; }
	if (!sys.parallel_restarts) PI_perturb_bead_COMs_ENTIRE_SYSTEM();
	PI_calculate_energy();
	PI_calc_system_mass();
	average_current_observables_into_PI_avgObservables();      // the initial state counts once (:63-66)
	int move = PI_pick_NVT_move();
	System::observables_t saved = *sys.observables;          // backup_observables_ALL_SYSTEMS (:699-709): the aggregate ...
	for (System *S : systems) S->checkpoint->observables = *S->observables;   // ... and every bead system's own (restore() puts them back)
	double pot_current = sys.observables->potential();
	if (!std::isfinite(pot_current)) sys.observables->energy = pot_current = MAXVALUE;
	const auto t_loop = std::chrono::steady_clock::now();
	const long long sweeps0 = pi_sweeps;
	for (sys.step = 1; sys.step <= sys.numsteps; sys.step++) {
		const double pot_init = pot_current;
		const double chain_init = (move == MOVETYPE_PERTURB_BEADS) ? PI_chain_mass_length2() : 0;
		const double orient_init = (move == MOVETYPE_PERTURB_BEADS) ? PI_orientational_mu_length2() : 0;
		PI_make_move(move);
		double pot_trial = PI_calculate_potential();
		const double chain_trial = (move == MOVETYPE_PERTURB_BEADS) ? PI_chain_mass_length2() : 0;
		const double orient_trial = (move == MOVETYPE_PERTURB_BEADS) ? PI_orientational_mu_length2() : 0;
		double bf;
		if (!std::isfinite(pot_trial)) { pot_trial = sys.observables->energy = MAXVALUE; bf = 0; }
		else bf = PI_NVT_boltzmann_factor(pot_trial - pot_init, chain_trial - chain_init, orient_trial - orient_init, move);
		sys.nodestats->boltzmann_factor = bf;
		int accepted;
		if ((Rando::rand() < bf) && (systems[0]->iterator_failed == 0)) {
			accepted = 1;
			pot_current = pot_trial;
			PI_calculate_energy();                           // (:148) nothing has moved since the trial sweep: its answer is reused
			saved = *sys.observables;
			for (System *S : systems) S->checkpoint->observables = *S->observables;
			sys.nodestats->accept++;
		} else {
			accepted = 0;
			restore_PI_systems();
			*sys.observables = saved;
			sys.nodestats->reject++;
		}
		if (log) log->push_back({move, pot_trial, bf, accepted, sys.observables->kinetic_energy});
		move = PI_pick_NVT_move();
		// every correlation time and at the very end (:176-178): the averages (do_PI_corrtime_bookkeeping, :248-270) ...
		if (sys.corrtime && (!(sys.step % sys.corrtime) || sys.step == sys.numsteps)) {
			for (System *S : systems) S->calc_system_mass();
			sys.observables->total_mass = systems[0]->observables->total_mass;
			sys.observables->frozen_mass = systems[0]->observables->frozen_mass;
			average_current_observables_into_PI_avgObservables();
		}
		// ... and the restart geometry of every bead system (:280-310)
		if (sys.write_files && rank == 0 && sys.corrtime && (!(sys.step % sys.corrtime) || sys.step == sys.numsteps))
			for (System *S : systems) { S->update_com(); S->wrap_all(); S->write_molecules_wrapper(S->pqr_restart); }
	}
	loop_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_loop).count();
	loop_sweeps = pi_sweeps - sweeps0;
	if (sys.write_files && rank == 0)                  // the final state (:182-194)
		for (System *S : systems) { S->update_com(); S->wrap_all(); S->write_molecules_wrapper(S->pqr_output); }
	return true;
}

} // namespace mpmc_host
