// host_capi.cpp — C entry points over the host mirror, used by the trajectory-parity tests and the command-line driver.
#include <cstring>
#include <vector>

#include "mpmc_host.h"

using namespace mpmc_host;

static double g_last_loop_seconds = 0, g_last_loop_sweeps = 0;
static double g_last_averages[27] = {0};
static void keep_averages(const System &sys) {
	const System::avg_observables_t &a = *sys.avg_observables;
	const double v[27] = {a.energy, a.energy_error, a.N, a.N_error, a.coulombic_energy, a.coulombic_energy_error, a.rd_energy, a.rd_energy_error,
	                      a.polarization_energy, a.polarization_energy_error, a.density, a.density_error, a.heat_capacity, a.heat_capacity_error,
	                      a.compressibility, a.compressibility_error, a.percent_wt, a.percent_wt_me, a.excess_ratio, a.qst, a.pore_density, a.NU,
	                      sys.observables->frozen_mass, sys.pbc.volume, (double)sys.avg_counter, a.kinetic_energy, a.kinetic_energy_error};
	memcpy(g_last_averages, v, sizeof v);
}

extern "C" {

// wall seconds of the last run's step loop (set-up and initial energy excluded) and the potential sweeps it made (path integrals)
void mpmc_host_last_stats(double out[2]) { out[0] = g_last_loop_seconds; out[1] = g_last_loop_sweeps; }

// the averages the last run accumulated every correlation time and at the end (System::update_root_averages; a path-integral run
// also counts its initial state): energy, N, coulombic, rd, polarization (value, error each), density (value, error), heat capacity
// (value, error), compressibility (value, error), percent_wt, percent_wt_me, excess_ratio, qst, pore_density, NU, then frozen mass,
// volume, samples taken, kinetic energy (value, error)
void mpmc_host_last_averages(double out[27]) { memcpy(out, g_last_averages, sizeof g_last_averages); }

// Run the simulation an input file describes (ensemble nvt | uvt | pi_nvt), with `P` beads for pi_nvt, for at most max_steps
// steps (0 = numsteps from the file).  log receives 5 doubles per step: move type, trial energy (potential for pi_nvt),
// Boltzmann factor, accepted (0/1), N (classic) or kinetic energy (pi_nvt).  summary[8] = final observables: energy, rd, coulombic,
// polarization, kinetic, N, accepted count, rejected count.  Returns 0 or the reference-style error code that was thrown.
int mpmc_host_run(const char *input_file, int P, int max_steps, double *log, int log_capacity, int *n_logged, double *summary) {
	try {
		SimulationControl sc(input_file, P);
		if (max_steps > 0) sc.sys.numsteps = (uint32_t)max_steps;
		sc.initializeSimulationObjects();
		std::vector<System::step_record> rec;
		sc.runSimulation(&rec);
		g_last_loop_seconds = sc.sys.ensemble == ENSEMBLE_PATH_INTEGRAL_NVT ? sc.loop_seconds : sc.sys.loop_seconds;
		g_last_loop_sweeps = (double)sc.loop_sweeps;
		keep_averages(sc.sys);
		const int n = (int)std::min<size_t>(rec.size(), (size_t)log_capacity);
		for (int i = 0; i < n; i++) {
			log[5 * i] = rec[i].movetype; log[5 * i + 1] = rec[i].final_energy; log[5 * i + 2] = rec[i].boltzmann_factor;
			log[5 * i + 3] = rec[i].accepted; log[5 * i + 4] = rec[i].N;
		}
		if (n_logged) *n_logged = n;
		if (summary) {
			const System::observables_t &o = *sc.sys.observables;
			summary[0] = o.energy; summary[1] = o.rd_energy; summary[2] = o.coulombic_energy; summary[3] = o.polarization_energy;
			summary[4] = o.kinetic_energy; summary[5] = o.N; summary[6] = sc.sys.nodestats->accept; summary[7] = sc.sys.nodestats->reject;
		}
	} catch (int e) {
		return e ? e : internal_error;
	}
	return 0;
}

// Molecule::orient on a bare list of sites (test hook): pos[n][3] in / out, mass[n], the handle site and the target orientation
int mpmc_host_debug_orient(int n, double *pos, const double *mass, int orientation_site, const double orientation[3]) {
	if (n < 1 || orientation_site < 0 || orientation_site >= n) return invalid_datum;
	Molecule m;
	Atom **tail = &m.atoms;
	for (int i = 0; i < n; i++) {
		Atom *a = new Atom;
		a->mass = mass[i];
		for (int p = 0; p < 3; p++) a->pos[p] = pos[3 * i + p];
		*tail = a; tail = &a->next;
	}
	m.orient(orientation, orientation_site);
	int i = 0;
	for (Atom *a = m.atoms; a; a = a->next, i++) for (int p = 0; p < 3; p++) pos[3 * i + p] = a->pos[p];
	return 0;
}

// The same for a path-integral run whose bead systems are sharded over `nranks` processes, one GPU each (launched e.g. by torchrun):
// every rank calls this with its rank, the device it owns and the 128-byte NCCL id rank 0 obtained from mpmc_nccl_get_unique_id().
// All ranks replay the same random stream and return the same log.
int mpmc_host_run_sharded(const char *input_file, int P, int max_steps, int rank, int nranks, int device, const char *nccl_id,
                          double *log, int log_capacity, int *n_logged, double *summary) {
	try {
		SimulationControl sc(input_file, P);
		if (sc.sys.ensemble != ENSEMBLE_PATH_INTEGRAL_NVT) return invalid_ensemble;
		if (max_steps > 0) sc.sys.numsteps = (uint32_t)max_steps;
		sc.sys.gpu_device = device;
		sc.set_sharding(rank, nranks, nccl_id);
		sc.initializeSimulationObjects();
		std::vector<System::step_record> rec;
		sc.runSimulation(&rec);
		g_last_loop_seconds = sc.sys.ensemble == ENSEMBLE_PATH_INTEGRAL_NVT ? sc.loop_seconds : sc.sys.loop_seconds;
		g_last_loop_sweeps = (double)sc.loop_sweeps;
		keep_averages(sc.sys);
		const int n = (int)std::min<size_t>(rec.size(), (size_t)log_capacity);
		for (int i = 0; i < n; i++) {
			log[5 * i] = rec[i].movetype; log[5 * i + 1] = rec[i].final_energy; log[5 * i + 2] = rec[i].boltzmann_factor;
			log[5 * i + 3] = rec[i].accepted; log[5 * i + 4] = rec[i].N;
		}
		if (n_logged) *n_logged = n;
		if (summary) {
			const System::observables_t &o = *sc.sys.observables;
			summary[0] = o.energy; summary[1] = o.rd_energy; summary[2] = o.coulombic_energy; summary[3] = o.polarization_energy;
			summary[4] = o.kinetic_energy; summary[5] = o.N; summary[6] = sc.sys.nodestats->accept; summary[7] = sc.sys.nodestats->reject;
		}
	} catch (int e) {
		return e ? e : internal_error;
	}
	return 0;
}

// The PQR file the mirror writes for (bead system `s` of) the job an input file describes, as read — no device work: readers +
// update_com + wrap_all + write_molecules_wrapper.  Also returns the three file names chosen for that system ('\n'-separated:
// pqr_input, pqr_restart, pqr_output) in names (capacity cap).
int mpmc_host_write_pqr(const char *input_file, int P, int s, const char *out_path, char *names, int cap) {
	try {
		SimulationControl sc(input_file, P);
		sc.sys.write_files = false;
		sc.initializeSimulationObjects();
		System &S = sc.systems.empty() ? sc.sys : *sc.systems[s];
		S.update_com();
		S.wrap_all();
		if (out_path && out_path[0]) S.write_molecules_wrapper(out_path);
		if (names) snprintf(names, cap, "%s\n%s\n%s", S.pqr_input, S.pqr_restart, S.pqr_output);
	} catch (int e) {
		return e ? e : internal_error;
	}
	return 0;
}

// One energy() of the system an input file describes, through the mirror (reader + flatten + engine): out[5] = energy, rd,
// coulombic, polarization, iterations.
int mpmc_host_energy(const char *input_file, double *out) {
	try {
		SimulationControl sc(input_file, 0);
		sc.initializeSimulationObjects();
		out[0] = sc.sys.energy();
		out[1] = sc.sys.observables->rd_energy; out[2] = sc.sys.observables->coulombic_energy;
		out[3] = sc.sys.observables->polarization_energy; out[4] = sc.sys.nodestats->polarization_iterations;
	} catch (int e) {
		return e ? e : internal_error;
	}
	return 0;
}

// What the mirror's readers make of an input file + PQR, without any device work: the flat site table in list order (the
// engine's upload layout, System::flatten) and the cell.  cell[22] = basis (9), reciprocal basis (9), volume, cutoff, ewald alpha,
// polar ewald alpha.  For pi_nvt (P > 0) the first bead system is described.  *n = number of sites; arrays hold `capacity` sites.
int mpmc_host_describe(const char *input_file, int P, int capacity, int *n, double *pos, double *q, double *alpha, double *eps, double *sigma,
                       double *mass, int *mol, int *frozen, double *cell) {
	try {
		SimulationControl sc(input_file, P);
		sc.initializeSimulationObjects();
		const System &s = sc.systems.empty() ? sc.sys : *sc.systems[0];
		std::vector<double> vp, vq, va, ve, vs, vm;
		std::vector<int> vmol, vfz;
		s.flatten(vp, vq, va, ve, vs, vm, vmol, vfz);
		*n = (int)vq.size();
		if (*n > capacity) return invalid_input;
		memcpy(pos, vp.data(), vp.size() * sizeof(double)); memcpy(q, vq.data(), vq.size() * sizeof(double));
		memcpy(alpha, va.data(), va.size() * sizeof(double)); memcpy(eps, ve.data(), ve.size() * sizeof(double));
		memcpy(sigma, vs.data(), vs.size() * sizeof(double)); memcpy(mass, vm.data(), vm.size() * sizeof(double));
		memcpy(mol, vmol.data(), vmol.size() * sizeof(int)); memcpy(frozen, vfz.data(), vfz.size() * sizeof(int));
		for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { cell[3 * i + j] = s.pbc.basis[i][j]; cell[9 + 3 * i + j] = s.pbc.reciprocal_basis[i][j]; }
		cell[18] = s.pbc.volume; cell[19] = s.pbc.cutoff; cell[20] = s.ewald_alpha; cell[21] = s.polar_ewald_alpha;
	} catch (int e) {
		return e ? e : internal_error;
	}
	return 0;
}

} // extern "C"
