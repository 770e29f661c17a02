// oracle/ref_harness.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product path).
//
// Thin C-ABI around the UNMODIFIED reference (b-tudor/mpmcxx) compiled from the sources where
// they lie under /root/reference/src (see oracle/Makefile, target `ref`).  It lets the tests,
// the golden-vector generator (tests/golden/make_golden.py) and bench.py's `--impl reference`
// arm drive the reference's own System::energy() and its public term functions
// (System.h:314-410) and read doubles at full precision instead of the 6-decimal energy.dat.
//
// What it steps around: a non-MPI build leaves `size`=0 (main.cpp:19) — we set
// size=1; SimulationControl's PI members are private — opened with the usual test-only macro.
// Nothing here re-implements reference arithmetic: every number comes from reference code.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <unistd.h>
#include <fcntl.h>
#include <vector>
#include <map>
#include <string>
#include <random>
#include <iostream>
#include <sstream>
#include <stdint.h>

// standard headers are all included above, so the macro only touches the reference's own classes
#define private public
#define protected public
#include "SimulationControl.h"
#undef private
#undef protected
#include "Atom.h"
#include "Molecule.h"
#include "Pair.h"
#include "Rando.h"

// globals the reference declares `extern` (System.cpp:13, Output.cpp, ...) and main.cpp defines
int  rank = 0;
int  size = 1;
bool mpi  = false;

static int g_last_error = 0;

namespace {

struct Quiet {   // the reference prints a banner per call; silence stdout while it runs
	int saved;
	Quiet() { fflush(stdout); saved = dup(1); int nul = open("/dev/null", O_WRONLY); dup2(nul, 1); close(nul); }
	~Quiet() { fflush(stdout); dup2(saved, 1); close(saved); }
};

System *pick(SimulationControl *sc, int s) {
	if (s < 0 || sc->systems.empty()) return &sc->sys;
	return sc->systems[(size_t)s];
}

} // namespace

extern "C" {

int mref_last_error() { return g_last_error; }

void *mref_open(const char *input_path, int P) {
	g_last_error = 0;
	char *path = strdup(input_path);
	SimulationControl *sc = nullptr;
	try {
		Quiet q;
		sc = new SimulationControl(path, P, false, nullptr);
		sc->initializeSimulationObjects();
	} catch (int e) {
		g_last_error = e;
		sc = nullptr;   // leak on failure: the reference's dtor is not exception-safe
	}
	free(path);
	return sc;
}

void mref_close(void *h) {
	// the reference's ~System double-frees in some configurations; tests are short-lived, so leak.
	(void)h;
}

int mref_nsys(void *h) { return (int)((SimulationControl *)h)->systems.size(); }

int mref_natoms(void *h, int s) { return pick((SimulationControl *)h, s)->countNatoms(); }

// Flatten the molecule list in list order (the order pairs()/atom_array use, System.cpp:881-904).
void mref_get_sites(void *h, int s, double *pos, double *charge, double *alpha, double *eps,
                    double *sigma, double *mass, int *mol, int *frozen) {
	System *sy = pick((SimulationControl *)h, s);
	int n = 0, m = 0;
	for (Molecule *mp = sy->molecules; mp; mp = mp->next, m++)
		for (Atom *ap = mp->atoms; ap; ap = ap->next, n++) {
			for (int p = 0; p < 3; p++) pos[3 * n + p] = ap->pos[p];
			charge[n] = ap->charge; alpha[n] = ap->polarizability; eps[n] = ap->epsilon;
			sigma[n] = ap->sigma; mass[n] = ap->mass; mol[n] = m; frozen[n] = ap->frozen;
		}
}

void mref_set_pos(void *h, int s, const double *pos) {
	System *sy = pick((SimulationControl *)h, s);
	int n = 0;
	for (Molecule *mp = sy->molecules; mp; mp = mp->next)
		for (Atom *ap = mp->atoms; ap; ap = ap->next, n++)
			for (int p = 0; p < 3; p++) ap->pos[p] = pos[3 * n + p];
}

// out: basis[9], recip[9], volume, cutoff, ewald_alpha, polar_ewald_alpha  (22 doubles)
void mref_get_cell(void *h, int s, double *out) {
	System *sy = pick((SimulationControl *)h, s);
	for (int i = 0; i < 3; i++)
		for (int j = 0; j < 3; j++) {
			out[3 * i + j]     = sy->pbc.basis[i][j];
			out[9 + 3 * i + j] = sy->pbc.reciprocal_basis[i][j];
		}
	out[18] = sy->pbc.volume; out[19] = sy->pbc.cutoff;
	out[20] = sy->ewald_alpha; out[21] = sy->polar_ewald_alpha;
}

// One reference energy() exactly as the MC loop calls it (cache semantics included).
// out[0..8] = energy, rd, coulombic, polarization, vdw, polarization_iterations, dipole_rrms,
//             iterator_failed, N
int mref_energy(void *h, int s, double *out) {
	System *sy = pick((SimulationControl *)h, s);
	try {
		Quiet q;
		sy->iterator_failed = 0;
		out[0] = sy->energy();
	} catch (int e) { return g_last_error = e; }
	out[1] = sy->observables->rd_energy;
	out[2] = sy->observables->coulombic_energy;
	out[3] = sy->observables->polarization_energy;
	out[4] = sy->observables->vdw_energy;
	out[5] = sy->nodestats ? sy->nodestats->polarization_iterations : 0;
	out[6] = sy->observables->dipole_rrms;
	out[7] = sy->iterator_failed;
	out[8] = sy->observables->N;
	return 0;
}

// Cold evaluation of every sub-term through the reference's public term functions, all pairs
// flagged.  out (16 doubles):
//  0 lj() total            1 Σ pair rd_energy      2 Σ pair lrc         3 Σ lj_lrc_self
//  4 coulombic_real()      5 Σ es_real_energy      6 Σ es_self_intra    7 coulombic_reciprocal()
//  8 coulombic_self()      9 polar() (0 if off)   10 iterations        11 dipole_rrms
// 12 iterator_failed      13 n pairs              14 n pairs with rimg<cutoff (not frozen)  15 spare
int mref_terms(void *h, int s, double *out) {
	System *sy = pick((SimulationControl *)h, s);
	for (int i = 0; i < 16; i++) out[i] = 0;
	try {
		Quiet q;
		sy->natoms = sy->countNatoms();
		sy->pairs();
		sy->flag_all_pairs();
		sy->iterator_failed = 0;
		if (!(sy->use_sg || sy->rd_only)) {
			out[4] = sy->coulombic_real();
			out[7] = sy->coulombic_reciprocal();
			out[8] = sy->coulombic_self();
			if (sy->polarization) {
				out[9]  = sy->polar();
				out[10] = sy->nodestats ? sy->nodestats->polarization_iterations : 0;
				out[11] = sy->observables->dipole_rrms;
				out[12] = sy->iterator_failed;
			}
		}
		out[0] = sy->lj();
		double cutoff = sy->pbc.cutoff;
		for (Molecule *mp = sy->molecules; mp; mp = mp->next)
			for (Atom *ap = mp->atoms; ap; ap = ap->next) {
				if (sy->rd_lrc) out[3] += sy->lj_lrc_self(ap, cutoff);
				for (Pair *pp = ap->pairs; pp; pp = pp->next) {
					out[1] += pp->rd_energy;
					out[2] += pp->lrc;
					out[5] += pp->es_real_energy;
					out[6] += pp->es_self_intra_energy;
					out[13] += 1;
					if (!pp->frozen && pp->rimg - SMALL_dR < cutoff) out[14] += 1;
				}
			}
	} catch (int e) { return g_last_error = e; }
	return 0;
}

// per-site polarization state after polar(): mu, ef_static, ef_induced, ef_induced_change (3n each),
// rank_metric (n)
void mref_get_dipoles(void *h, int s, double *mu, double *efs, double *efi, double *efic, double *rankm) {
	System *sy = pick((SimulationControl *)h, s);
	int n = 0;
	for (Molecule *mp = sy->molecules; mp; mp = mp->next)
		for (Atom *ap = mp->atoms; ap; ap = ap->next, n++) {
			for (int p = 0; p < 3; p++) {
				mu[3 * n + p]  = ap->mu[p];
				efs[3 * n + p] = ap->ef_static[p];
				efi[3 * n + p] = ap->ef_induced[p];
				efic[3 * n + p] = ap->ef_induced_change[p];
			}
			rankm[n] = ap->rank_metric;
		}
}

// Path-integral aggregates (SimulationControl.PathIntegral.cpp:752-828, 859-904).
// out: potential, rd, coulombic, polarization, vdw, kinetic, chain_mass_len2   (7 doubles)
int mref_pi_energy(void *h, double *out) {
	SimulationControl *sc = (SimulationControl *)h;
	try {
		Quiet q;
		out[5] = sc->PI_calculate_kinetic();
		out[0] = sc->PI_calculate_potential();
		out[6] = sc->PI_chain_mass_length2_ENTIRE_SYSTEM();
	} catch (int e) { return g_last_error = e; }
	out[1] = sc->sys.observables->rd_energy;
	out[2] = sc->sys.observables->coulombic_energy;
	out[3] = sc->sys.observables->polarization_energy;
	out[4] = sc->sys.observables->vdw_energy;
	return 0;
}

// Just the potential sweep (what BFC.potential.trial costs per move, PathIntegral.cpp:118)
double mref_pi_potential(void *h) {
	SimulationControl *sc = (SimulationControl *)h;
	Quiet q;
	return sc->PI_calculate_potential();
}

// The reference's own classic Markov chain, step by step: the body of System::mc() (System.MonteCarlo.cpp:38-94) driven from
// here so that every step's move type, trial energy, Boltzmann factor and decision can be read as doubles (the reference only
// prints 6-decimal averages at corrtime, and its bookkeeping writes through an unallocated struct in a non-MPI build).
// log: 5 doubles per step = movetype, final_energy, boltzmann_factor, accepted, observables->N.
int mref_mc_trajectory(void *h, int nsteps, double *log) {
	SimulationControl *sc = (SimulationControl *)h;
	System &s = sc->sys;
	try {
		Quiet q;
		s.observables->volume = s.pbc.volume;
		double initial_energy = s.mc_initial_energy(), final_energy = 0;
		s.do_checkpoint();
		for (int step = 1; step <= nsteps; step++) {
			s.step = step;
			initial_energy = s.observables->energy;
			s.make_move();
			final_energy = s.energy();
			if (!std::isfinite(final_energy)) { s.observables->energy = MAXVALUE; s.nodestats->boltzmann_factor = 0; }
			else s.boltzmann_factor(initial_energy, final_energy);
			double *L = log + 5 * (step - 1);
			L[0] = s.checkpoint->movetype; L[1] = final_energy; L[2] = s.nodestats->boltzmann_factor;
			if ((s.get_rand() < s.nodestats->boltzmann_factor) && !s.iterator_failed) { L[3] = 1; s.do_checkpoint(); }
			else { L[3] = 0; s.iterator_failed = 0; s.restore(); }
			L[4] = s.observables->N;
		}
	} catch (int e) { return g_last_error = e; }
	return 0;
}

// The same chain with the averaging the reference does for the initial state (setup_mpi, :186-190), every correlation time and at the
// very end (System.MonteCarlo.cpp:104-106 ->
// do_corrtime_bookkeeping :1902-1912, 1973-2022: calc_system_mass, then update_root_averages over the node's observables — the
// rest of that routine is file output and the MPI gather that a non-MPI build cannot run).  out[25] as mref_root_averages.
int mref_mc_averages(void *h, int nsteps, int corrtime, double *out) {
	SimulationControl *sc = (SimulationControl *)h;
	System &s = sc->sys;
	try {
		Quiet q;
		if (!s.avg_observables) s.avg_observables = (System::avg_observables_t *)calloc(1, sizeof(System::avg_observables_t));
		s.observables->volume = s.pbc.volume;
		double initial_energy = s.mc_initial_energy(), final_energy = 0;
		s.calc_system_mass(); s.update_root_averages(s.observables);      // setup_mpi (:186-190): the initial values count once
		s.do_checkpoint();
		for (int step = 1; step <= nsteps; step++) {
			s.step = step;
			initial_energy = s.observables->energy;
			s.make_move();
			final_energy = s.energy();
			if (!std::isfinite(final_energy)) { s.observables->energy = MAXVALUE; s.nodestats->boltzmann_factor = 0; }
			else s.boltzmann_factor(initial_energy, final_energy);
			if ((s.get_rand() < s.nodestats->boltzmann_factor) && !s.iterator_failed) s.do_checkpoint();
			else { s.iterator_failed = 0; s.restore(); }
			if (!(step % corrtime) || step == nsteps) { s.calc_system_mass(); s.update_root_averages(s.observables); }
		}
		const System::avg_observables_t &a = *s.avg_observables;
		const double v[22] = {a.energy, a.energy_error, a.N, a.N_error, a.coulombic_energy, a.coulombic_energy_error, a.rd_energy, a.rd_energy_error,
		                      a.polarization_energy, a.polarization_energy_error, a.density, a.density_error, a.heat_capacity, a.heat_capacity_error,
		                      a.compressibility, a.compressibility_error, a.percent_wt, a.percent_wt_me, a.excess_ratio, a.qst, a.pore_density, a.NU};
		memcpy(out, v, sizeof v);
		out[22] = s.observables->frozen_mass; out[23] = s.pbc.volume; out[24] = s.fugacities[0];
	} catch (int e) { return g_last_error = e; }
	return 0;
}

// The reference's path-integral chain, step by step: the body of SimulationControl::PI_nvt_mc() (PathIntegral.cpp:31-196) without
// the file output and statistics.  log: 5 doubles per step = move, trial potential, boltzmann_factor, accepted, kinetic energy.
int mref_pi_trajectory(void *h, int nsteps, double *log) {
	SimulationControl *sc = (SimulationControl *)h;
	try {
		Quiet q;
		for (System *S : sc->systems) { S->observables->temperature = sc->sys.temperature; S->observables->volume = S->pbc.volume; }
		if (!sc->sys.parallel_restarts) sc->PI_perturb_bead_COMs_ENTIRE_SYSTEM();
		sc->PI_calculate_energy();
		int move = sc->PI_pick_NVT_move();
		sc->backup_observables_ALL_SYSTEMS();
		auto &B = sc->BFC;
		B.potential.current = sc->sys.observables->potential();
		if (!std::isfinite(B.potential.current)) sc->sys.observables->energy = B.potential.current = MAXVALUE;
		B.chain_mass_len2.current = 0; B.orient_mu_len2.current = 0;
		for (int step = 1; step <= nsteps; step++) {
			sc->sys.step = step;
			B.potential.init = B.potential.current;
			B.chain_mass_len2.init = (move == MOVETYPE_PERTURB_BEADS) ? sc->PI_chain_mass_length2() : 0;
			B.orient_mu_len2.init = (move == MOVETYPE_PERTURB_BEADS) ? sc->PI_orientational_mu_length2() : 0;
			sc->PI_make_move(move);
			B.potential.trial = sc->PI_calculate_potential();
			B.chain_mass_len2.trial = (move == MOVETYPE_PERTURB_BEADS) ? sc->PI_chain_mass_length2() : 0;
			B.orient_mu_len2.trial = (move == MOVETYPE_PERTURB_BEADS) ? sc->PI_orientational_mu_length2() : 0;
			double bf;
			if (!std::isfinite(B.potential.trial)) { B.potential.trial = sc->sys.observables->energy = MAXVALUE; bf = 0; }
			else bf = sc->PI_NVT_boltzmann_factor(B);
			double *L = log + 5 * (step - 1);
			L[0] = move; L[1] = B.potential.trial; L[2] = bf;
			if ((Rando::rand() < bf) && (sc->systems[0]->iterator_failed == 0)) {
				L[3] = 1;
				B.potential.current = B.potential.trial;
				sc->PI_calculate_energy();
				sc->backup_observables_ALL_SYSTEMS();
			} else {
				L[3] = 0;
				sc->restore_PI_systems();
				*sc->sys.observables = *sc->sys.checkpoint->observables;
			}
			L[4] = sc->sys.observables->kinetic_energy;
			move = sc->PI_pick_NVT_move();
		}
	} catch (int e) { return g_last_error = e; }
	return 0;
}

// The path-integral chain with the reference's averaging: once for the initial state (PathIntegral.cpp:63-66), then every
// correlation time and at the very end (:176-178 -> do_PI_corrtime_bookkeeping :237-270: system masses, then
// average_current_observables_into_PI_avgObservables -> sys.update_root_averages; the rest of that routine is file output).
// out[20]: energy, kinetic, rd, coulombic, polarization, N, density, heat capacity, compressibility (value, error each), then frozen
// mass and volume.
int mref_pi_averages(void *h, int nsteps, int corrtime, double *out) {
	SimulationControl *sc = (SimulationControl *)h;
	try {
		Quiet q;
		if (!sc->sys.avg_observables) sc->sys.avg_observables = (System::avg_observables_t *)calloc(1, sizeof(System::avg_observables_t));
		for (System *S : sc->systems) { S->observables->temperature = sc->sys.temperature; S->observables->volume = S->pbc.volume; }
		if (!sc->sys.parallel_restarts) sc->PI_perturb_bead_COMs_ENTIRE_SYSTEM();
		sc->PI_calculate_energy();
		sc->PI_calc_system_mass();
		sc->average_current_observables_into_PI_avgObservables();
		int move = sc->PI_pick_NVT_move();
		sc->backup_observables_ALL_SYSTEMS();
		auto &B = sc->BFC;
		B.potential.current = sc->sys.observables->potential();
		if (!std::isfinite(B.potential.current)) sc->sys.observables->energy = B.potential.current = MAXVALUE;
		B.chain_mass_len2.current = 0; B.orient_mu_len2.current = 0;
		for (int step = 1; step <= nsteps; step++) {
			sc->sys.step = step;
			B.potential.init = B.potential.current;
			B.chain_mass_len2.init = (move == MOVETYPE_PERTURB_BEADS) ? sc->PI_chain_mass_length2() : 0;
			B.orient_mu_len2.init = (move == MOVETYPE_PERTURB_BEADS) ? sc->PI_orientational_mu_length2() : 0;
			sc->PI_make_move(move);
			B.potential.trial = sc->PI_calculate_potential();
			B.chain_mass_len2.trial = (move == MOVETYPE_PERTURB_BEADS) ? sc->PI_chain_mass_length2() : 0;
			B.orient_mu_len2.trial = (move == MOVETYPE_PERTURB_BEADS) ? sc->PI_orientational_mu_length2() : 0;
			double bf;
			if (!std::isfinite(B.potential.trial)) { B.potential.trial = sc->sys.observables->energy = MAXVALUE; bf = 0; }
			else bf = sc->PI_NVT_boltzmann_factor(B);
			if ((Rando::rand() < bf) && (sc->systems[0]->iterator_failed == 0)) {
				B.potential.current = B.potential.trial;
				sc->PI_calculate_energy();
				sc->backup_observables_ALL_SYSTEMS();
			} else {
				sc->restore_PI_systems();
				*sc->sys.observables = *sc->sys.checkpoint->observables;
			}
			move = sc->PI_pick_NVT_move();
			if (!(step % corrtime) || step == nsteps) {
				for (System *S : sc->systems) S->calc_system_mass();
				sc->sys.observables->total_mass = sc->systems[0]->observables->total_mass;
				sc->sys.observables->frozen_mass = sc->systems[0]->observables->frozen_mass;
				sc->average_current_observables_into_PI_avgObservables();
			}
		}
		const System::avg_observables_t &a = *sc->sys.avg_observables;
		const double v[18] = {a.energy, a.energy_error, a.kinetic_energy, a.kinetic_energy_error, a.rd_energy, a.rd_energy_error, a.coulombic_energy,
		                      a.coulombic_energy_error, a.polarization_energy, a.polarization_energy_error, a.N, a.N_error, a.density, a.density_error,
		                      a.heat_capacity, a.heat_capacity_error, a.compressibility, a.compressibility_error};
		memcpy(out, v, sizeof v);
		out[18] = sc->sys.observables->frozen_mass; out[19] = sc->sys.pbc.volume;
	} catch (int e) { return g_last_error = e; }
	return 0;
}

/* System::write_molecules_wrapper (src/System.Output.cpp:837-1091) of system s into `path`: the reference's own PQR writer, for
 * pinning the mirror's writer byte for byte.  Runs energy() first when `after_energy` so that wrapped coordinates are current. */
int mref_write_pqr(void *h, int s, const char *path, int after_energy) {
	SimulationControl *sc = (SimulationControl *)h;
	System *S = (s < 0 || sc->systems.empty()) ? &sc->sys : sc->systems[s];
	try {
		Quiet q;
		if (after_energy) S->energy();
		char *p = strdup(path);
		int rc = S->write_molecules_wrapper(p);
		free(p);
		return rc;
	} catch (int e) { return e ? e : -1; }
}

/* pqr_input / pqr_restart / pqr_output file names the reference chose for system s (check_io_files_options,
 * src/SimulationControl.cpp:2196-2360), '\n'-separated into out (capacity cap) */
// System::update_root_averages (src/System.Averages.cpp:8-208) fed with n samples of (energy, coulombic, rd, polarization, N, NU);
// out[25]: the averages and derived adsorption observables the reference reports, then frozen mass, volume, fugacities[0].  The routine keeps its sample counter in a
// function-static: one call per process.
int mref_root_averages(void *h, int s, int n, const double *x, double *out) {
	System *sys = pick((SimulationControl *)h, s);
	try {
		Quiet q;
		if (!sys->avg_observables) sys->avg_observables = (System::avg_observables_t *)calloc(1, sizeof(System::avg_observables_t));
		sys->calc_system_mass();
		for (int i = 0; i < n; i++) {
			System::observables_t o = *sys->observables;
			o.energy = x[6 * i]; o.coulombic_energy = x[6 * i + 1]; o.rd_energy = x[6 * i + 2]; o.polarization_energy = x[6 * i + 3];
			o.N = x[6 * i + 4]; o.NU = x[6 * i + 5];
			o.temperature = sys->temperature; o.volume = sys->pbc.volume;
			sys->update_root_averages(&o);
		}
		const System::avg_observables_t &a = *sys->avg_observables;
		const double v[20] = {a.energy, a.energy_error, a.N, a.N_error, a.coulombic_energy, a.coulombic_energy_error, a.rd_energy, a.rd_energy_error,
		                      a.polarization_energy, a.polarization_energy_error, a.density, a.density_error, a.heat_capacity, a.heat_capacity_error,
		                      a.compressibility, a.compressibility_error, a.percent_wt, a.percent_wt_me, a.excess_ratio, a.qst};
		memcpy(out, v, sizeof v);
		out[20] = a.pore_density; out[21] = a.NU; out[22] = sys->observables->frozen_mass; out[23] = sys->pbc.volume; out[24] = sys->fugacities[0];
	} catch (int e) { return g_last_error = e; }
	return 0;
}

int mref_io_filenames(void *h, int s, char *out, int cap) {
	SimulationControl *sc = (SimulationControl *)h;
	System *S = (s < 0 || sc->systems.empty()) ? &sc->sys : sc->systems[s];
	snprintf(out, cap, "%s\n%s\n%s", S->pqr_input, S->pqr_restart, S->pqr_output);
	return 0;
}

} // extern "C"
