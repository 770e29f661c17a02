// host.cpp — implementation of the host-side mirror (see mpmc_host.h).  Compile with -ffp-contract=off.
#include "mpmc_host.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <sstream>

namespace mpmc_host {

std::normal_distribution<double> Rando::normal_distribution(0.0, 1.0);
std::uniform_real_distribution<double> Rando::uniform_distribution(0.0, 1.0);
std::mt19937 Rando::mt(5489u);

static bool ieq(const std::string &a, const char *b) {
	size_t n = strlen(b);
	if (a.size() != n) return false;
	for (size_t i = 0; i < n; i++) if (tolower((unsigned char)a[i]) != tolower((unsigned char)b[i])) return false;
	return true;
}
static double to_double(const std::string &s) {       // SafeOps::atod: the whole token must parse (src/SafeOps.cpp:70-82)
	size_t idx = 0;
	double d;
	try { d = std::stod(s, &idx); } catch (...) { throw invalid_datum; }
	if (idx != s.size()) throw invalid_datum;
	return d;
}
static int to_int(const std::string &s) {
	size_t idx = 0;
	int v;
	try { v = std::stoi(s, &idx); } catch (...) { throw invalid_datum; }
	if (idx != s.size()) throw invalid_datum;
	return v;
}

// ------------------------------------------------------------------------------------------------------------
// Molecule
// ------------------------------------------------------------------------------------------------------------
Molecule::Molecule(const Molecule &o) {
	memcpy(moleculetype, o.moleculetype, sizeof moleculetype);
	id = o.id; mass = o.mass; frozen = o.frozen; adiabatic = o.adiabatic; spectre = o.spectre; target = o.target;
	for (int p = 0; p < 3; p++) { com[p] = o.com[p]; wrapped_com[p] = o.wrapped_com[p]; }
	Atom **tail = &atoms;
	for (const Atom *a = o.atoms; a; a = a->next) {
		Atom *c = new Atom(*a);
		c->next = nullptr;
		*tail = c;
		tail = &c->next;
	}
}

Molecule::~Molecule() {
	for (Atom *a = atoms; a;) { Atom *n = a->next; delete a; a = n; }
}

int Molecule::natoms() const { int n = 0; for (const Atom *a = atoms; a; a = a->next) n++; return n; }

// Molecule::update_COM (src/Molecule.cpp:256-281)
void Molecule::update_COM() {
	mass = 0; com[0] = com[1] = com[2] = 0;
	for (Atom *a = atoms; a; a = a->next) {
		mass += a->mass;
		com[0] += a->mass * a->pos[0]; com[1] += a->mass * a->pos[1]; com[2] += a->mass * a->pos[2];
	}
	com[0] /= mass; com[1] /= mass; com[2] /= mass;
}

void Molecule::translate(double x, double y, double z) {
	com[0] += x; com[1] += y; com[2] += z;
	for (Atom *a = atoms; a; a = a->next) { a->pos[0] += x; a->pos[1] += y; a->pos[2] += z; }
}

void Molecule::move_to_(double x, double y, double z) { translate(x - com[0], y - com[1], z - com[2]); }

// six uniforms: three magnitudes scale*u*cutoff, three sign tests u < 0.5 (src/Molecule.cpp:286-321)
void Molecule::translate_rand_pbc(double scale, const PeriodicBoundary &pbc, std::mt19937 *mt_rand) {
	double dice[6];
	std::uniform_real_distribution<double> d{0, 1};
	for (int i = 0; i < 6; i++) dice[i] = d(*mt_rand);
	translate_rand_pbc(scale, pbc, dice);
}
void Molecule::translate_rand_pbc(double scale, const PeriodicBoundary &pbc, double dice[6]) {
	double t[3];
	for (int p = 0; p < 3; p++) {
		t[p] = scale * dice[p] * pbc.cutoff;
		if (dice[3 + p] < 0.5) t[p] *= -1.0;
	}
	for (Atom *a = atoms; a; a = a->next) { a->pos[0] += t[0]; a->pos[1] += t[1]; a->pos[2] += t[2]; }
	update_COM();
}

namespace {
struct Quat {                                    // src/Quaternion.cpp
	double X, Y, Z, W;
	static Quat axis_angle(double x, double y, double z, double angle_rad) {
		double mag = std::sqrt(x * x + y * y + z * z);
		if (mag == 0.0) return {0, 0, 0, 1};
		x = x / mag; y = y / mag; z = z / mag;
		double s = std::sin(angle_rad / 2.0);
		return {x * s, y * s, z * s, std::cos(angle_rad / 2.0)};
	}
	Quat conj() const { return {-X, -Y, -Z, W}; }
	Quat operator*(const Quat &r) const {
		double w = W * r.W - X * r.X - Y * r.Y - Z * r.Z;
		double x = W * r.X + X * r.W + Y * r.Z - Z * r.Y;
		double y = W * r.Y - X * r.Z + Y * r.W + Z * r.X;
		double z = W * r.Z + X * r.Y - Y * r.X + Z * r.W;
		return {x, y, z, w};
	}
};
} // namespace

// three normals for the axis, one uniform for the angle = u * 360 * scale degrees (src/Molecule.cpp:128-136)
void Molecule::rotate_rand(double scale) {
	double x = Rando::rand_normal();
	double y = Rando::rand_normal();
	double z = Rando::rand_normal();
	double angle = Rando::rand() * 360 * scale;
	rotate(x, y, z, angle);
}

// rotation about the COM: p' = R (p R*) (src/Molecule.cpp:138-206; the association matters for the last bit)
void Molecule::rotate(double x, double y, double z, double angle_degrees) {
	const Quat R = Quat::axis_angle(x, y, z, angle_degrees / 57.2957795), Rc = R.conj();
	for (Atom *a = atoms; a; a = a->next) { a->pos[0] -= com[0]; a->pos[1] -= com[1]; a->pos[2] -= com[2]; }
	for (Atom *a = atoms; a; a = a->next) {
		const Quat p{a->pos[0], a->pos[1], a->pos[2], 0.0};
		const Quat ans = R * (p * Rc);
		a->pos[0] = ans.X; a->pos[1] = ans.Y; a->pos[2] = ans.Z;
		a->pos[0] += com[0]; a->pos[1] += com[1]; a->pos[2] += com[2];
	}
}

// Turn the molecule about its COM so that the site `orientation_site` points along `orientation` (src/Molecule.cpp:211-254):
// angle = acos(c . o / |o|) with c the normalised COM->site vector, axis = c x o, p' = (R p) R*.  A site that sits ON the COM gives
// c = 0, axis = 0 and the identity rotation, exactly like the reference.
void Molecule::orient(const double o[3], int orientation_site) {
	update_COM();
	const double rc[3] = {com[0], com[1], com[2]};
	translate(-rc[0], -rc[1], -rc[2]);
	Atom *a = atoms;
	for (int site = 0; site != orientation_site; site++) a = a->next;
	double cx = a->pos[0], cy = a->pos[1], cz = a->pos[2];
	const double mag = std::sqrt(cx * cx + cy * cy + cz * cz);
	if (mag != 0) { cx = cx / mag; cy = cy / mag; cz = cz / mag; } else cx = cy = cz = 0;
	const double angle = std::acos((cx * o[0] + cy * o[1] + cz * o[2]) / std::sqrt(o[0] * o[0] + o[1] * o[1] + o[2] * o[2]));
	const Quat R = Quat::axis_angle(cy * o[2] - cz * o[1], cz * o[0] - cx * o[2], cx * o[1] - cy * o[0], angle), Rc = R.conj();
	for (a = atoms; a; a = a->next) {
		const Quat p{a->pos[0], a->pos[1], a->pos[2], 0.0};
		const Quat ans = (R * p) * Rc;
		a->pos[0] = ans.X; a->pos[1] = ans.Y; a->pos[2] = ans.Z;
	}
	translate(rc[0], rc[1], rc[2]);
}

// ------------------------------------------------------------------------------------------------------------
// System: geometry
// ------------------------------------------------------------------------------------------------------------
System::System(const System &o) {
	// settings only; the bead systems read their own geometry (initialize_PI_NVT_Systems, PathIntegral.cpp:618-631)
	cuda = o.cuda; ensemble = o.ensemble;
	memcpy(job_name, o.job_name, sizeof job_name); memcpy(pqr_input, o.pqr_input, sizeof pqr_input);
	memcpy(pqr_output, o.pqr_output, sizeof pqr_output); memcpy(pqr_restart, o.pqr_restart, sizeof pqr_restart);
	long_output = o.long_output; independent_particle = o.independent_particle; write_files = o.write_files;
	numsteps = o.numsteps; corrtime = o.corrtime; move_factor = o.move_factor; rot_factor = o.rot_factor;
	insert_probability = o.insert_probability; bead_perturb_probability = o.bead_perturb_probability;
	temperature = o.temperature; pressure = o.pressure; free_volume = o.free_volume; scale_charge = o.scale_charge;
	preset_seed_on = o.preset_seed_on; preset_seed = o.preset_seed;
	rd_lrc = o.rd_lrc; rd_only = o.rd_only; wrapall = o.wrapall; parallel_restarts = o.parallel_restarts; read_pqr_box_on = o.read_pqr_box_on;
	ewald_alpha_set = o.ewald_alpha_set; polar_ewald_alpha_set = o.polar_ewald_alpha_set; ewald_kmax = o.ewald_kmax;
	ewald_alpha = o.ewald_alpha; polar_ewald_alpha = o.polar_ewald_alpha;
	polarization = o.polarization; polar_iterative = o.polar_iterative; polar_ewald = o.polar_ewald; polar_zodid = o.polar_zodid;
	polar_palmo = o.polar_palmo; polar_rrms = o.polar_rrms; polar_gs = o.polar_gs; polar_gs_ranked = o.polar_gs_ranked; polar_sor = o.polar_sor;
	polar_esor = o.polar_esor; polar_max_iter = o.polar_max_iter; damp_type = o.damp_type;
	polar_gamma = o.polar_gamma; polar_damp = o.polar_damp; polar_precision = o.polar_precision; gpu_device = o.gpu_device;
	pbc = o.pbc;
}

System::~System() {
	if (gpu) mpmc_destroy(gpu);
	for (Molecule *m = molecules; m;) { Molecule *n = m->next; delete m; m = n; }
	delete checkpoint->molecule_backup;
}

// The cell from the geometry file (src/System.cpp:775-850): up to the first line that starts with END, every line whose first seven
// whitespace-separated words read  REMARK BOX BASIS[k] = x y z  sets basis row k when all three numbers convert; a row that is
// not found keeps what the input file gave.
void System::read_pqr_box(const char *file) {
	std::ifstream in(file);
	if (!in) throw fopen_fail_read;
	std::string line;
	bool have[3] = {false, false, false};
	while (std::getline(in, line)) {
		if (have[0] && have[1] && have[2]) break;
		std::istringstream ss(line);
		std::string t[7];
		int n = 0;
		while (n < 7 && (ss >> t[n])) n++;
		if (n == 0) continue;
		if (!t[0].compare(0, 3, "END")) break;
		if (n < 7 || t[0] != "REMARK" || t[1] != "BOX" || t[3] != "=") continue;
		for (int k = 0; k < 3; k++) {
			if (t[2] != "BASIS[" + std::to_string(k) + "]") continue;
			// SafeOps::atod (src/SafeOps.cpp:70-82): std::stod into the target, failure when the word has trailing characters or is no
			// number at all; the first failure abandons the line, and only a line whose three words all convert marks the row as read
			bool ok = true;
			for (int p = 0; p < 3 && ok; p++) {
				try {
					size_t idx = 0;
					pbc.basis[k][p] = std::stod(t[4 + p], &idx);
					ok = idx == t[4 + p].size();
				} catch (...) { ok = false; }
			}
			if (ok) have[k] = true;
		}
	}
}

// PQR reader (src/System.cpp:515-770): whitespace tokens ATOM id atomtype moltype F|M molid x y z mass q alpha eps sigma omega ...;
// BOX pseudo-atoms are skipped (:592); charges are converted to reduced units (:624); frozen charges are scaled (:669).
void System::read_molecules(const char *file) {
	std::ifstream in(file);
	if (!in) throw fopen_fail_read;
	std::string line;
	Molecule **mtail = &molecules, *cur = nullptr;
	Atom **atail = nullptr;
	int atom_counter = 0, moveable = 0;
	while (std::getline(in, line)) {
		std::istringstream ss(line);
		std::vector<std::string> t;
		for (std::string w; ss >> w;) t.push_back(w);
		if (t.empty()) continue;
		if (t[0].size() >= 3 && ieq(t[0].substr(0, 3), "END")) break;
		if (!ieq(t[0], "ATOM") || t.size() < 4 || ieq(t[3], "BOX")) continue;
		if (t.size() < 15) throw invalid_datum;
		const int molid = to_int(t[5]);
		to_int(t[1]);
		const int fz = ieq(t[4], "F"), ad = ieq(t[4], "A"), sp = ieq(t[4], "S"), tg = ieq(t[4], "T");
		const double x = to_double(t[6]), y = to_double(t[7]), z = to_double(t[8]), mass = to_double(t[9]);
		double q = to_double(t[10]);
		const double al = to_double(t[11]), ep = to_double(t[12]), sg = to_double(t[13]), om = to_double(t[14]);
		q *= E2REDUCED;
		if (fz) q *= scale_charge;
		if (!cur || cur->id != molid) {
			cur = new Molecule();
			*mtail = cur; mtail = &cur->next;
			atail = &cur->atoms;
		}
		strncpy(cur->moleculetype, t[3].c_str(), sizeof cur->moleculetype - 1);
		cur->id = molid; cur->frozen = fz; cur->adiabatic = ad; cur->spectre = sp; cur->target = tg; cur->mass += mass;
		Atom *a = new Atom();
		a->id = ++atom_counter; a->frozen = fz; a->adiabatic = ad; a->spectre = sp; a->target = tg;
		a->pos[0] = x; a->pos[1] = y; a->pos[2] = z; a->mass = mass; a->charge = q; a->polarizability = al; a->epsilon = ep; a->sigma = sg; a->omega = om;
		if (t.size() > 15) a->gwp_alpha = to_double(t[15]);
		if (t.size() > 16) a->c6 = to_double(t[16]);
		if (t.size() > 17) a->c8 = to_double(t[17]);
		if (t.size() > 18) a->c10 = to_double(t[18]);
		if (t.size() > 19) a->c9 = to_double(t[19]);
		strncpy(a->atomtype, t[2].c_str(), sizeof a->atomtype - 1);
		*atail = a; atail = &a->next;
	}
	for (Molecule *m = molecules; m; m = m->next) if (!m->frozen) moveable++;
	if (!atom_counter) throw molecule_wo_atoms;
	if (!moveable) throw missing_required_datum;      // "no moveable molecules found" (:757-760)
	gpu_table_stale = true;
}

// The PQR file as the reference writes it (src/System.Output.cpp:900-1091): CRYST1 record (lengths %9.3f, angles %7.2f with VMD's
// alpha <-> beta convention), one ATOM record per site — PDB-style %8.3f coordinates, or %11.6f when `long_output` is on or a basis
// component reaches 100 A; wrapped coordinates when `wrapall` — the eight corners of the cell as a frozen BOX pseudo-molecule with
// CONECT records along its edges (wrapall only), the basis as REMARK records, END.
int System::write_molecules(FILE *fp) {
	auto dot = [](const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
	bool ext = long_output != 0;
	for (int i = 0; i < 3 && !ext; i++) for (int j = 0; j < 3; j++) if (pbc.basis[i][j] >= 100.0) ext = true;
	const double *b0 = pbc.basis[0], *b1 = pbc.basis[1], *b2 = pbc.basis[2];
	fprintf(fp, "CRYST1");
	fprintf(fp, "%9.3f", std::sqrt(dot(b0, b0)));
	fprintf(fp, "%9.3f", std::sqrt(dot(b1, b1)));
	fprintf(fp, "%9.3f", std::sqrt(dot(b2, b2)));
	fprintf(fp, "%7.2f", 180.0 / pi * std::acos(dot(b2, b0) / std::sqrt(dot(b0, b0) * dot(b2, b2))));
	fprintf(fp, "%7.2f", 180.0 / pi * std::acos(dot(b1, b2) / std::sqrt(dot(b1, b1) * dot(b2, b2))));
	fprintf(fp, "%7.2f", 180.0 / pi * std::acos(dot(b0, b1) / std::sqrt(dot(b1, b1) * dot(b0, b0))));
	fprintf(fp, "\n");
	int i = 1, j = 1;
	for (Molecule *m = molecules; m; m = m->next, j++)
		for (Atom *a = m->atoms; a; a = a->next, i++) {
			fprintf(fp, "ATOM  ");
			fprintf(fp, "%5d", i);
			fprintf(fp, " %-4.45s", a->atomtype);
			fprintf(fp, " %-3.3s ", m->moleculetype);
			fprintf(fp, "%-1.1s", a->adiabatic ? "A" : a->frozen ? "F" : a->spectre ? "S" : a->target ? "T" : "M");
			fprintf(fp, " %4d   ", independent_particle ? i : j);
			const double *x = wrapall ? a->wrapped_pos : a->pos;
			for (int p = 0; p < 3; p++) { if (ext) fprintf(fp, "%11.6f ", x[p]); else fprintf(fp, "%8.3f", x[p]); }
			fprintf(fp, " %8.5f", a->mass);
			fprintf(fp, " %8.5f", a->charge / E2REDUCED);
			fprintf(fp, " %8.5f", a->polarizability);
			fprintf(fp, " %8.5f", a->epsilon);
			fprintf(fp, " %8.5f", a->sigma);
			fprintf(fp, " %8.5f", a->omega);
			fprintf(fp, " %8.5f", a->gwp_alpha);
			fprintf(fp, " %8.5f", a->c6);
			fprintf(fp, " %8.5f", a->c8);
			fprintf(fp, " %8.5f", a->c10);
			fprintf(fp, " %8.5f", a->c9);
			fprintf(fp, "\n");
		}
	if (wrapall) {
		int atom_box = i, label[2][2][2];
		const int molecule_box = j;
		for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) for (int c = 0; c < 2; c++) {
			fprintf(fp, "ATOM  ");
			fprintf(fp, "%5d", atom_box);
			fprintf(fp, " %-4.45s", "X");
			fprintf(fp, " %-3.3s ", "BOX");
			fprintf(fp, "%-1.1s", "F");
			fprintf(fp, " %4d   ", molecule_box);
			const double occ[3] = {a - 0.5, b - 0.5, c - 0.5};
			for (int p = 0; p < 3; p++) {
				double v = 0;
				for (int q = 0; q < 3; q++) v += pbc.basis[q][p] * occ[q];
				if (ext) fprintf(fp, "%11.6f ", v); else fprintf(fp, "%8.3f", v);
			}
			fprintf(fp, " %8.4f", 0.0); fprintf(fp, " %8.4f", 0.0); fprintf(fp, " %8.5f", 0.0); fprintf(fp, " %8.5f", 0.0); fprintf(fp, " %8.5f", 0.0);
			fprintf(fp, "\n");
			label[a][b][c] = atom_box++;
		}
		for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) for (int c = 0; c < 2; c++)
			for (int l = 0; l < 2; l++) for (int m = 0; m < 2; m++) for (int n = 0; n < 2; n++)
				if (std::abs(a - l) + std::abs(b - m) + std::abs(c - n) == 1) fprintf(fp, "CONECT %4d %4d\n", label[a][b][c], label[l][m][n]);
	}
	for (int r = 0; r < 3; r++) fprintf(fp, "REMARK BOX BASIS[%d] = %20.14lf %20.14lf %20.14lf\n", r, pbc.basis[r][0], pbc.basis[r][1], pbc.basis[r][2]);
	fprintf(fp, "END\n");
	fflush(fp);
	return 0;
}

// :837-895 (single-process branch): the previous file becomes "<name>.last"
int System::write_molecules_wrapper(const char *filename) {
	if (!strncmp(filename, "/dev/null", 9)) return 0;    // `pqr_restart off` / `pqr_output off`: nothing to format
	if (FILE *t = fopen(filename, "r")) {
		fclose(t);
		const std::string old = std::string(filename) + ".last";
		if (rename(filename, old.c_str())) fprintf(stderr, "WARNING: Unable to rename .last file.\n");
	}
	FILE *fp = fopen(filename, "w");
	if (!fp) throw 1001;                              // fopen_fail_write
	const int rc = write_molecules(fp);
	fclose(fp);
	return rc;
}

// PeriodicBoundary::update (src/PeriodicBoundary.cpp:31-101) + update_pbc (src/System.cpp:859-876)
void System::update_pbc() {
	double (*b)[3] = pbc.basis;
	double v = b[0][0] * (b[1][1] * b[2][2] - b[1][2] * b[2][1]);
	v += b[0][1] * (b[1][2] * b[2][0] - b[1][0] * b[2][2]);
	v += b[0][2] * (b[1][0] * b[2][1] - b[1][1] * b[2][0]);
	pbc.volume = v;
	double short_mag = MAXVALUE;
	if (v > 0) {
		for (int i = -15; i <= 15; i++) for (int j = -15; j <= 15; j++) for (int k = -15; k <= 15; k++) {
			if (!i && !j && !k) continue;
			double c[3];
			for (int p = 0; p < 3; p++) c[p] = i * b[0][p] + j * b[1][p] + k * b[2][p];
			double m = std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
			if (m < short_mag) short_mag = m;
		}
		pbc.cutoff = 0.5 * short_mag;
	} else pbc.cutoff = MAXVALUE;
	if (pbc.volume <= 0.0 || pbc.cutoff <= 0.0) throw invalid_box_dimensions;
	double iv = 1.0 / v;
	double (*r)[3] = pbc.reciprocal_basis;
	r[0][0] = iv * (b[1][1] * b[2][2] - b[1][2] * b[2][1]); r[0][1] = iv * (b[0][2] * b[2][1] - b[0][1] * b[2][2]); r[0][2] = iv * (b[0][1] * b[1][2] - b[0][2] * b[1][1]);
	r[1][0] = iv * (b[1][2] * b[2][0] - b[1][0] * b[2][2]); r[1][1] = iv * (b[0][0] * b[2][2] - b[0][2] * b[2][0]); r[1][2] = iv * (b[0][2] * b[1][0] - b[0][0] * b[1][2]);
	r[2][0] = iv * (b[1][0] * b[2][1] - b[1][1] * b[2][0]); r[2][1] = iv * (b[0][1] * b[2][0] - b[0][0] * b[2][1]); r[2][2] = iv * (b[0][0] * b[1][1] - b[0][1] * b[1][0]);
	if (ewald_alpha_set != 1) ewald_alpha = 3.5 / pbc.cutoff;
	if (polar_ewald_alpha_set != 1) polar_ewald_alpha = 3.5 / pbc.cutoff;
}

int System::countNatoms() const {
	int n = 0;
	for (Molecule *m = molecules; m; m = m->next) for (Atom *a = m->atoms; a; a = a->next) n++;
	return n;
}

unsigned int System::countN() {                      // src/System.cpp:909-931
	unsigned int count = 0;
	observables->spin_ratio = 0;
	for (Molecule *m = molecules; m; m = m->next) if (!(m->frozen || m->adiabatic || m->target)) count++;
	observables->N = count;
	return count;
}

void System::update_com() {                          // src/System.cpp:1347-1374
	for (Molecule *m = molecules; m; m = m->next) {
		m->com[0] = m->com[1] = m->com[2] = 0;
		if (m->spectre || m->target) continue;
		m->mass = 0;
		for (Atom *a = m->atoms; a; a = a->next) {
			m->mass += a->mass;
			for (int i = 0; i < 3; i++) m->com[i] += a->mass * a->pos[i];
		}
		for (int i = 0; i < 3; i++) m->com[i] /= m->mass;
	}
}

void System::wrap_all() {                            // src/System.cpp:1379-1425
	for (Molecule *m = molecules; m; m = m->next) {
		double dimg[3] = {0, 0, 0};
		if (!m->frozen) {
			double d[3];
			for (int i = 0; i < 3; i++) {
				d[i] = 0;
				for (int j = 0; j < 3; j++) d[i] += pbc.reciprocal_basis[j][i] * m->com[j];
				d[i] = std::rint(d[i]);
			}
			for (int i = 0; i < 3; i++) {
				dimg[i] = 0;
				for (int j = 0; j < 3; j++) dimg[i] += pbc.basis[j][i] * d[j];
				m->wrapped_com[i] = dimg[i];
			}
		}
		for (Atom *a = m->atoms; a; a = a->next) for (int i = 0; i < 3; i++) a->wrapped_pos[i] = a->pos[i] - dimg[i];
	}
}

// ------------------------------------------------------------------------------------------------------------
// System: the hot path through the engine
// ------------------------------------------------------------------------------------------------------------
void System::fill_config(mpmc_config &c, int n_beads) const {
	memset(&c, 0, sizeof c);
	for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) c.basis[3 * i + j] = pbc.basis[i][j];
	c.n_beads = n_beads; c.device = gpu_device;
	c.rd_lrc = rd_lrc; c.rd_only = rd_only; c.ewald_kmax = ewald_kmax;
	c.ewald_alpha = ewald_alpha_set ? ewald_alpha : 0.0;
	c.polar_ewald_alpha = polar_ewald_alpha_set ? polar_ewald_alpha : 0.0;
	c.polarization = polarization; c.polar_ewald = polar_ewald; c.polar_iterative = polar_iterative; c.damp_type = damp_type;
	c.polar_gs = polar_gs; c.polar_gs_ranked = polar_gs_ranked; c.polar_palmo = polar_palmo; c.polar_sor = polar_sor; c.polar_esor = polar_esor;
	c.polar_zodid = polar_zodid; c.polar_rrms = polar_rrms; c.polar_max_iter = polar_max_iter;
	c.polar_damp = polar_damp; c.polar_gamma = polar_gamma; c.polar_precision = polar_precision;
}

void System::flatten(std::vector<double> &pos, std::vector<double> &q, std::vector<double> &al, std::vector<double> &ep, std::vector<double> &sg,
                     std::vector<double> &ms, std::vector<int> &mol, std::vector<int> &fz) const {
	int m = 0;
	for (Molecule *mp = molecules; mp; mp = mp->next, m++)
		for (Atom *ap = mp->atoms; ap; ap = ap->next) {
			for (int p = 0; p < 3; p++) pos.push_back(ap->pos[p]);
			q.push_back(ap->charge); al.push_back(ap->polarizability); ep.push_back(ap->epsilon); sg.push_back(ap->sigma); ms.push_back(ap->mass);
			mol.push_back(m); fz.push_back(ap->frozen);
		}
}

double System::energy() {
	natoms = countNatoms();
	if (!cuda) throw unsupported_setting;            // this mirror has no CPU energy path
	int rc;
	if (!gpu) {
		mpmc_config c;
		fill_config(c, 1);
		if ((rc = mpmc_create(&c, &gpu))) throw rc;
		gpu_table_stale = true;
	}
	std::vector<double> pos, q, al, ep, sg, ms;
	std::vector<int> mol, fz;
	pos.reserve(3 * natoms);
	flatten(pos, q, al, ep, sg, ms, mol, fz);
	if (gpu_table_stale || (int)gpu_pos.size() != 3 * natoms) {
		if ((rc = mpmc_upload_sites(gpu, natoms, pos.data(), q.data(), al.data(), ep.data(), sg.data(), ms.data(), mol.data(), fz.data()))) throw rc;
		gpu_table_stale = false;
	} else {
		for (int i = 0; i < natoms;) {
			if (memcmp(&pos[3 * i], &gpu_pos[3 * i], 3 * sizeof(double))) {
				int j = i;
				while (j < natoms && memcmp(&pos[3 * j], &gpu_pos[3 * j], 3 * sizeof(double))) j++;
				if ((rc = mpmc_update_sites(gpu, 0, i, j - i, &pos[3 * i]))) throw rc;
				i = j;
			} else i++;
		}
	}
	gpu_pos.swap(pos);
	mpmc_energy_out o;
	if ((rc = mpmc_energy(gpu, &o))) throw rc;
	observables->coulombic_energy = o.coulombic_energy;
	observables->polarization_energy = o.polarization_energy;
	observables->rd_energy = o.rd_energy;
	observables->vdw_energy = 0;
	observables->energy = o.energy;
	observables->dipole_rrms = o.dipole_rrms;
	nodestats->polarization_iterations = o.polarization_iterations;
	iterator_failed = o.iterator_failed;
	update_com();
	wrap_all();
	countN();
	observables->spin_ratio /= observables->N;
	observables->NU = observables->N * observables->energy;
	last_volume = pbc.volume;
	return o.energy;
}

void System::download_dipoles() {
	if (!gpu || !polarization) return;
	const int n = countNatoms();
	std::vector<double> mu(3 * n), es(3 * n), ei(3 * n), ec(3 * n);
	int rc = mpmc_download_dipoles(gpu, 0, mu.data(), es.data(), ei.data(), ec.data());
	if (rc) throw rc;
	int i = 0;
	for (Molecule *m = molecules; m; m = m->next)
		for (Atom *a = m->atoms; a; a = a->next, i++)
			for (int p = 0; p < 3; p++) { a->mu[p] = mu[3 * i + p]; a->ef_static[p] = es[3 * i + p]; a->ef_induced[p] = ei[3 * i + p]; a->ef_induced_change[p] = ec[3 * i + p]; }
}

// ------------------------------------------------------------------------------------------------------------
// System: the classic Markov chain (src/System.MonteCarlo.cpp)
// ------------------------------------------------------------------------------------------------------------
double System::mc_initial_energy() {
	double e = energy();
	if (!std::isfinite(e)) e = observables->energy = MAXVALUE;    // "be a bit forgiving of the initial state" (:165-167)
	return e;
}

// (a) back up the state, (b) choose the next move type, (c) choose its target molecule (:252-504).
// RNG order: uVT draws insert-vs-not, then insert-vs-remove; every ensemble then draws the molecule index floor(u * N).
void System::do_checkpoint() {
	checkpoint->observables = *observables;
	std::vector<Molecule *> exchange;
	for (Molecule *m = molecules; m; m = m->next) if (!(m->frozen || m->adiabatic || m->target)) exchange.push_back(m);
	switch (ensemble) {
	case ENSEMBLE_UVT:
		if (get_rand() < insert_probability) checkpoint->movetype = (get_rand() < 0.5) ? MOVETYPE_INSERT : MOVETYPE_REMOVE;
		else checkpoint->movetype = MOVETYPE_DISPLACE;
		break;
	case ENSEMBLE_NVT:
		checkpoint->movetype = MOVETYPE_DISPLACE;
		break;
	default:
		throw invalid_ensemble;
	}
	int num_exchange = (int)exchange.size() - 1;
	const int altered = (int)std::floor(get_rand() * observables->N);
	if (altered < 0 || altered >= (int)exchange.size()) throw internal_error;
	checkpoint->molecule_altered = exchange[altered];
	if (!num_exchange && checkpoint->movetype == MOVETYPE_REMOVE) checkpoint->movetype = MOVETYPE_DISPLACE;   // never empty the system
	Molecule *prev = nullptr;
	for (Molecule *m = molecules; m; m = m->next) {
		if (m == checkpoint->molecule_altered) { checkpoint->head = prev; checkpoint->tail = m->next; break; }
		prev = m;
	}
	delete checkpoint->molecule_backup;
	checkpoint->molecule_backup = new Molecule(*checkpoint->molecule_altered);
}

void System::displace(Molecule *molecule, const PeriodicBoundary &PBC, double trans_scale, double rot_scale) {
	molecule->translate_rand_pbc(trans_scale, PBC, &mt_rand);
	molecule->rotate_rand(rot_scale);
}

// apply the move chosen by do_checkpoint (:719-900)
void System::make_move() {
	switch (checkpoint->movetype) {
	case MOVETYPE_INSERT: {
		double rnd[3], com[3];
		for (int p = 0; p < 3; p++) rnd[p] = 0.5 - get_rand();
		for (int p = 0; p < 3; p++) {
			com[p] = 0;
			for (int q = 0; q < 3; q++) com[p] += pbc.basis[q][p] * rnd[q];
		}
		Molecule *ins = checkpoint->molecule_backup;        // a copy of the selected molecule becomes the new one
		for (Atom *a = ins->atoms; a; a = a->next) for (int p = 0; p < 3; p++) a->pos[p] += com[p] - ins->com[p];
		for (int p = 0; p < 3; p++) ins->com[p] = com[p];
		ins->rotate_rand(1.0);
		if (!checkpoint->head) molecules = ins; else checkpoint->head->next = ins;   // linked in FRONT of the selected molecule (:799-805)
		ins->next = checkpoint->molecule_altered;
		checkpoint->molecule_altered = ins;
		checkpoint->tail = ins->next;
		checkpoint->molecule_backup = nullptr;
		gpu_table_stale = true;
	} break;
	case MOVETYPE_REMOVE:
		if (!checkpoint->head) { checkpoint->molecule_altered = molecules; molecules = molecules->next; }
		else checkpoint->head->next = checkpoint->tail;
		delete checkpoint->molecule_altered;
		checkpoint->molecule_altered = nullptr;
		gpu_table_stale = true;
		break;
	case MOVETYPE_DISPLACE:
		displace(checkpoint->molecule_altered, pbc, move_factor, rot_factor);
		break;
	default:
		throw invalid_monte_carlo_move;
	}
}

// :1345-1470, the branches without cavity bias; fugacity = pressure when no equation of state is selected (:1358-1363)
void System::boltzmann_factor(double initial_energy, double final_energy) {
	const double delta = final_energy - initial_energy;
	switch (ensemble) {
	case ENSEMBLE_UVT: {
		const double fugacity = pressure;
		switch (checkpoint->movetype) {
		case MOVETYPE_INSERT:
			nodestats->boltzmann_factor = pbc.volume * fugacity * ATM2REDUCED / (temperature * (double)(observables->N)) * exp(-delta / temperature) * (double)(1);
			break;
		case MOVETYPE_REMOVE:
			nodestats->boltzmann_factor = temperature * ((double)(observables->N) + 1.0) / (pbc.volume * fugacity * ATM2REDUCED) * exp(-delta / temperature) / (double)(1);
			break;
		case MOVETYPE_DISPLACE:
			nodestats->boltzmann_factor = exp(-delta / temperature);
			break;
		default:
			throw invalid_monte_carlo_move;
		}
	} break;
	case ENSEMBLE_NVT:
		nodestats->boltzmann_factor = exp(-delta / temperature);
		break;
	default:
		throw invalid_ensemble;
	}
}

// undo make_move, then choose the next move (:1510-1590)
void System::restore() {
	*observables = checkpoint->observables;
	switch (checkpoint->movetype) {
	case MOVETYPE_INSERT:
		if (!checkpoint->head) molecules = molecules->next; else checkpoint->head->next = checkpoint->tail;
		delete checkpoint->molecule_altered;
		checkpoint->molecule_altered = nullptr;
		gpu_table_stale = true;
		break;
	case MOVETYPE_REMOVE:
		if (!checkpoint->head) molecules = checkpoint->molecule_backup; else checkpoint->head->next = checkpoint->molecule_backup;
		checkpoint->molecule_backup->next = checkpoint->tail;
		checkpoint->molecule_backup = nullptr;
		gpu_table_stale = true;
		break;
	default:
		if (checkpoint->head) checkpoint->head->next = checkpoint->molecule_backup; else molecules = checkpoint->molecule_backup;
		checkpoint->molecule_backup->next = checkpoint->tail;
		checkpoint->molecule_backup = nullptr;
		delete checkpoint->molecule_altered;
		checkpoint->molecule_altered = nullptr;
	}
	if (ensemble == ENSEMBLE_PATH_INTEGRAL_NVT || ensemble == ENSEMBLE_NVT_GIBBS) return;
	do_checkpoint();
}

void System::calc_system_mass() {
	observables->total_mass = 0;
	observables->frozen_mass = 0;
	for (Molecule *m = molecules; m; m = m->next) {
		observables->total_mass += m->mass;
		if (m->frozen || m->adiabatic) observables->frozen_mass += m->mass;
	}
}

// Running means (weight (m-1)/m for the old average, 1/m for the new sample), errors sqrt(<x^2> - <x>^2) / sqrt(m - 1), and what the
// reference derives from them: density, heat capacity and compressibility with the Stirling form of their errors, weight percent,
// excess adsorption, pore density, isosteric heat.  `fugacities` is an array member in the reference, so its excess adsorption always
// uses fugacities[0] (0 without a fugacity keyword), never the pressure (:189-194).
void System::update_root_averages(observables_t *obs) {
	constexpr double NA = 6.0221415e23, A32CM3 = 1.0e-24, ATM2PASCALS = 101325.0;
	avg_observables_t *a = avg_observables;
	const double frozen_mass = obs->frozen_mass;
	avg_counter++;
	const double m = (double)avg_counter;
	const double sdom = 1.0 / sqrt(m - 1.0), factor = (m - 1.0) / m;
	auto fold = [&](double &avg, double &sq, double &err, double x) {
		avg = factor * avg + x / m;
		sq = factor * sq + (x * x) / m;
		err = sdom * sqrt(sq - avg * avg);
	};
	fold(a->energy, a->energy_sq, a->energy_error, obs->energy);
	fold(a->coulombic_energy, a->coulombic_energy_sq, a->coulombic_energy_error, obs->coulombic_energy);
	fold(a->rd_energy, a->rd_energy_sq, a->rd_energy_error, obs->rd_energy);
	fold(a->polarization_energy, a->polarization_energy_sq, a->polarization_energy_error, obs->polarization_energy);
	fold(a->kinetic_energy, a->kinetic_energy_sq, a->kinetic_energy_error, obs->kinetic_energy);
	fold(a->N, a->N_sq, a->N_error, obs->N);
	a->NU = factor * a->NU + obs->NU / m;
	double particle_mass = 0;
	for (Molecule *mp = molecules; mp; mp = mp->next) if (!mp->frozen && !mp->adiabatic) { particle_mass = mp->mass; break; }
	const double curr_density = obs->N * particle_mass / (pbc.volume * NA * A32CM3);
	fold(a->density, a->density_sq, a->density_error, curr_density);
	double gammaratio = pow((m - 2.0) / (m - 1.0), 0.5 * m - 1.0) * sqrt(0.5 * (m - 2.0)) * exp(0.5);
	gammaratio = sqrt(1.0 / avg_counter * (m - 1.0 - 2.0 * gammaratio * gammaratio));
	a->heat_capacity = (kB * NA / 1000.0) * (a->energy_sq - a->energy * a->energy) / (temperature * temperature);
	a->heat_capacity_error = sdom * 2.0 * gammaratio * a->heat_capacity;
	a->compressibility = ATM2PASCALS * (pbc.volume / pow(METER2ANGSTROM, 3)) * (a->N_sq - a->N * a->N) / (kB * temperature * a->N * a->N);
	a->compressibility_error = sdom * 2.0 * gammaratio * a->compressibility;
	if (frozen_mass > 0.0) {
		a->percent_wt = 100.0 * a->N * particle_mass / (frozen_mass + a->N * particle_mass);
		a->percent_wt_me = 100.0 * a->N * particle_mass / frozen_mass;
		if (free_volume > 0.0) {
			a->excess_ratio = 1000.0 * (a->N * particle_mass - (particle_mass * free_volume * fugacities[0] * ATM2REDUCED) / temperature) / frozen_mass;
			a->pore_density = curr_density * pbc.volume / free_volume;
		}
		a->qst = -(a->NU - a->N * a->energy);
		a->qst /= (a->N_sq - a->N * a->N);
		a->qst += temperature;
		a->qst *= kB * NA / 1000.0;
	}
}

bool System::mc(std::vector<step_record> *log) {       // :20-134
	observables->volume = pbc.volume;
	double initial_energy = mc_initial_energy(), final_energy = 0;
	if (corrtime) { calc_system_mass(); update_root_averages(observables); }    // the initial values count once (setup_mpi, :186-190)
	do_checkpoint();
	const auto t_loop = std::chrono::steady_clock::now();
	for (step = 1; step <= numsteps; step++) {
		initial_energy = observables->energy;
		make_move();
		final_energy = energy();
		if (!std::isfinite(final_energy)) { observables->energy = MAXVALUE; nodestats->boltzmann_factor = 0; }
		else boltzmann_factor(initial_energy, final_energy);
		const int movetype = checkpoint->movetype;
		const double bf = nodestats->boltzmann_factor;
		int accepted;
		if ((get_rand() < nodestats->boltzmann_factor) && !iterator_failed) {
			accepted = 1;
			do_checkpoint();
			nodestats->accept++;
		} else {
			accepted = 0;
			iterator_failed = 0;
			restore();
			nodestats->reject++;
		}
		if (log) log->push_back({movetype, final_energy, bf, accepted, observables->N});
		// every correlation time: the restart geometry (do_corrtime_bookkeeping, System.MonteCarlo.cpp:1925-1934)
		if (write_files && corrtime && !(step % corrtime) && pqr_restart[0]) { update_com(); wrap_all(); write_molecules_wrapper(pqr_restart); }
		// ... and the node's averages, also at the very end (:104-106, :1912, :1973-2022)
		if (corrtime && (!(step % corrtime) || step == numsteps)) { calc_system_mass(); update_root_averages(observables); }
	}
	loop_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_loop).count();
	if (write_files && pqr_output[0]) { update_com(); wrap_all(); write_molecules_wrapper(pqr_output); }   // the final state (:112-120)
	return true;
}

} // namespace mpmc_host
