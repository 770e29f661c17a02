/* tests/shim/oracle_engine.c — TEST INFRASTRUCTURE ONLY.
 *
 * The subset of the engine's C-ABI (include/mpmc_b200.h) that the C++ host mirror calls, implemented on the CPU oracle
 * (oracle/oracle.c).  tests/test_host_trajectory_cpu.py links the host mirror against this instead of libmpmc_b200.so, so that
 * the mirror's Monte Carlo drivers — move generation, RNG order, Boltzmann factors, accept/reject, restore, uVT insert/remove,
 * path-integral bead moves — are checked against the reference's golden trajectories without a GPU.  Nothing under
 * mpmcxx_b200/ builds, links or loads this file. */
#include <stdlib.h>
#include <string.h>

#include "mpmc_b200.h"
#include "oracle.h"

struct mpmc_engine {
	mpmc_config cfg;
	int n;
	double *pos, *q, *al, *ep, *sg, *ms;      /* pos: n_beads * n * 3 */
	int *mol, *fz;
	double *mu, *es, *ei, *ec;                /* bead 0 of the last energy() */
};

static const char *g_err = "";
const char *mpmc_last_error(void) { return g_err; }
int mpmc_abi_version(void) { return MPMC_ABI_VERSION; }

int mpmc_create(const mpmc_config *cfg, mpmc_engine **out) {
	mpmc_engine *e = (mpmc_engine *)calloc(1, sizeof *e);
	e->cfg = *cfg;
	if (e->cfg.n_beads < 1) e->cfg.n_beads = 1;
	*out = e;
	return MPMC_OK;
}

static void drop(mpmc_engine *e) {
	free(e->pos); free(e->q); free(e->al); free(e->ep); free(e->sg); free(e->ms); free(e->mol); free(e->fz);
	free(e->mu); free(e->es); free(e->ei); free(e->ec);
	e->pos = e->q = e->al = e->ep = e->sg = e->ms = e->mu = e->es = e->ei = e->ec = NULL;
	e->mol = e->fz = NULL;
}

int mpmc_destroy(mpmc_engine *e) {
	if (e) { drop(e); free(e); }
	return MPMC_OK;
}

static double *dupd(const double *p, size_t n) { double *r = (double *)malloc(n * sizeof(double) + 8); memcpy(r, p, n * sizeof(double)); return r; }
static int *dupi(const int *p, size_t n) { int *r = (int *)malloc(n * sizeof(int) + 8); memcpy(r, p, n * sizeof(int)); return r; }

int mpmc_upload_sites(mpmc_engine *e, int n, const double *pos, const double *charge, const double *alpha, const double *epsilon,
                      const double *sigma, const double *mass, const int *mol, const int *frozen) {
	drop(e);
	e->n = n;
	e->pos = dupd(pos, (size_t)e->cfg.n_beads * n * 3);
	e->q = dupd(charge, n); e->al = dupd(alpha, n); e->ep = dupd(epsilon, n); e->sg = dupd(sigma, n); e->ms = dupd(mass, n);
	e->mol = dupi(mol, n); e->fz = dupi(frozen, n);
	e->mu = (double *)calloc(3 * (size_t)n + 1, sizeof(double)); e->es = (double *)calloc(3 * (size_t)n + 1, sizeof(double));
	e->ei = (double *)calloc(3 * (size_t)n + 1, sizeof(double)); e->ec = (double *)calloc(3 * (size_t)n + 1, sizeof(double));
	return MPMC_OK;
}

int mpmc_update_sites(mpmc_engine *e, int bead, int first, int count, const double *pos) {
	if (bead < 0 || bead >= e->cfg.n_beads || first < 0 || first + count > e->n) return MPMC_ERR_INVALID_INPUT;
	memcpy(e->pos + ((size_t)bead * e->n + first) * 3, pos, (size_t)count * 3 * sizeof(double));
	return MPMC_OK;
}

int mpmc_update_sites_all_beads(mpmc_engine *e, int first, int count, const double *pos) {
	for (int b = 0; b < e->cfg.n_beads; b++) {
		int rc = mpmc_update_sites(e, b, first, count, pos + (size_t)b * count * 3);
		if (rc) return rc;
	}
	return MPMC_OK;
}

static void options(const mpmc_config *c, int *iopt, double *dopt) {
	memset(iopt, 0, sizeof(int) * ORC_NIOPT);
	memset(dopt, 0, sizeof(double) * ORC_NDOPT);
	iopt[ORC_RD_LRC] = c->rd_lrc; iopt[ORC_RD_ONLY] = c->rd_only; iopt[ORC_POLARIZATION] = c->polarization; iopt[ORC_DAMP_TYPE] = c->damp_type;
	iopt[ORC_POLAR_EWALD] = c->polar_ewald; iopt[ORC_POLAR_ITERATIVE] = c->polar_iterative; iopt[ORC_POLAR_GS] = c->polar_gs;
	iopt[ORC_POLAR_GS_RANKED] = c->polar_gs_ranked; iopt[ORC_POLAR_PALMO] = c->polar_palmo; iopt[ORC_POLAR_SOR] = c->polar_sor;
	iopt[ORC_POLAR_ESOR] = c->polar_esor; iopt[ORC_POLAR_ZODID] = c->polar_zodid; iopt[ORC_POLAR_RRMS] = c->polar_rrms;
	iopt[ORC_POLAR_MAX_ITER] = c->polar_max_iter; iopt[ORC_EWALD_KMAX] = c->ewald_kmax;
	dopt[ORC_POLAR_DAMP] = c->polar_damp; dopt[ORC_POLAR_GAMMA] = c->polar_gamma; dopt[ORC_POLAR_PRECISION] = c->polar_precision;
	dopt[ORC_EWALD_ALPHA] = c->ewald_alpha; dopt[ORC_POLAR_EWALD_ALPHA] = c->polar_ewald_alpha;
}

static int one_energy(mpmc_engine *e, int bead, double *o, int want_sites) {
	int iopt[ORC_NIOPT];
	double dopt[ORC_NDOPT];
	options(&e->cfg, iopt, dopt);
	return orc_energy(e->n, e->pos + (size_t)bead * e->n * 3, e->q, e->al, e->ep, e->sg, e->mol, e->fz, e->cfg.basis, iopt, dopt, o,
	                  want_sites ? e->mu : NULL, want_sites ? e->es : NULL, want_sites ? e->ei : NULL, want_sites ? e->ec : NULL, NULL);
}

int mpmc_energy(mpmc_engine *e, mpmc_energy_out *out) {
	double o[ORC_NOUT];
	memset(out, 0, sizeof *out);
	if (e->n < 1) return MPMC_ERR_NO_MOLECULES;
	int rc = one_energy(e, 0, o, 1);
	if (rc) return MPMC_ERR_INVALID_INPUT;
	out->energy = o[ORC_O_ENERGY]; out->rd_energy = o[ORC_O_RD_TOTAL]; out->coulombic_energy = o[ORC_O_COULOMBIC];
	out->polarization_energy = o[ORC_O_POLAR]; out->rd_pair = o[ORC_O_RD_PAIR]; out->rd_lrc_pair = o[ORC_O_LRC_PAIR];
	out->rd_lrc_self = o[ORC_O_LRC_SELF]; out->es_real = o[ORC_O_ES_REAL]; out->es_self_intra = o[ORC_O_ES_SELF_INTRA];
	out->es_reciprocal = o[ORC_O_ES_RECIP]; out->es_self = o[ORC_O_ES_SELF]; out->dipole_rrms = o[ORC_O_DIPOLE_RRMS];
	out->n_pairs_in_cutoff = o[ORC_O_NPAIR_IN_CUTOFF];
	out->polarization_iterations = (int)o[ORC_O_ITERATIONS]; out->iterator_failed = (int)o[ORC_O_ITERATOR_FAILED];
	return MPMC_OK;
}

int mpmc_download_dipoles(mpmc_engine *e, int bead, double *mu, double *ef_static, double *ef_induced, double *ef_induced_change) {
	(void)bead;
	const size_t len = 3 * (size_t)e->n * sizeof(double);
	if (mu) memcpy(mu, e->mu, len);
	if (ef_static) memcpy(ef_static, e->es, len);
	if (ef_induced) memcpy(ef_induced, e->ei, len);
	if (ef_induced_change) memcpy(ef_induced_change, e->ec, len);
	return MPMC_OK;
}

/* PI_calculate_potential (PathIntegral.cpp:786-804): bead sums of rd, coulombic, polarization, vdw, divided by P */
int mpmc_pi_potential_allreduce(mpmc_engine *e, int P_global, double means[4], double *potential) {
	double s[4] = {0, 0, 0, 0}, o[ORC_NOUT];
	for (int b = 0; b < e->cfg.n_beads; b++) {
		if (one_energy(e, b, o, 0)) return MPMC_ERR_INVALID_INPUT;
		s[0] += o[ORC_O_RD_TOTAL]; s[1] += o[ORC_O_COULOMBIC]; s[2] += o[ORC_O_POLAR];
	}
	for (int q = 0; q < 4; q++) means[q] = s[q] / P_global;
	if (potential) *potential = means[0] + means[1] + means[3] + means[2];
	return MPMC_OK;
}

/* bead sharding needs GPUs and NCCL: the CPU shim only ever runs single-rank */
int mpmc_nccl_init(mpmc_engine *e, const char id[128], int rank, int nranks) {
	(void)e; (void)id; (void)rank; (void)nranks;
	g_err = "the CPU test shim has no collective";
	return MPMC_ERR_UNSUPPORTED;
}
