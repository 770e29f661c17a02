/* oracle/oracle.c — plain-C restatement of the mpmc++ energy hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file restates, on flat list-ordered arrays, what the reference computes over its linked
 * lists of Molecule -> Atom -> Pair nodes.  Each function cites the reference file:line it follows
 * (paths relative to /root/reference/src).  Nothing here is shipped or measured as the product:
 * the CUDA engine is checked against it (tests/, smoke()) and it is the `port` CPU baseline leg.
 * Parity status: PINNED against oracle/_ref and tests/golden (see oracle.h).
 *
 * Differences from the reference that are deliberate and do not change results beyond rounding:
 *  - the Thole A matrix (3N x 3N) is not materialised; T_ij is recomputed from the same pair
 *    geometry wherever the reference reads A_matrix (System.Energy.cpp:2748-2764);
 *  - O(N^2) loops are OpenMP-parallel over i with per-i partial sums added in index order, so
 *    results are deterministic for any thread count (the reference adds pair by pair);
 *  - no per-pair cache: every call is a "cold" evaluation (all pairs flagged, System.cpp:1284).
 * Compile with -ffp-contract=off: the reference (g++ -O2, x86-64 baseline) has no FMA contraction,
 * and the cutoff tests below are sensitive to the last bit of rimg on lattice configurations.
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* constants.h:12-55 */
static const double PI_ = 3.141592653589793238462643383279502884;
static const double ONE_OVER_SQRT_PI = 0.5641895835477562869480794515607725858440506293289988;
static const double MAXVALUE = 1.0e40;
static const double SMALL_dR = 1.0e-12;
static const double MAX_ITERATION_COUNT = 128;
static const double DEBYE2SKA = 85.10597636;
static const double kB = 1.3806503e-23;
static const double hBar2 = 1.11211999e-68;
static const double AMU2KG = 1.66053873e-27;
static const double ANGSTROM2METER = 1.0e-10;

typedef struct {
	double basis[3][3], recip[3][3], volume, cutoff, ewald_alpha, polar_ewald_alpha;
} cell_t;

/* PeriodicBoundary.cpp:31-101 (volume, cutoff over +-15 images, reciprocal = plain inverse);
 * System.cpp:859-876 (alphas default to 3.5/cutoff). */
static void cell_update(cell_t *c, const double basis[9], double ewald_alpha, double polar_ewald_alpha) {
	double (*b)[3] = c->basis;
	for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) b[i][j] = basis[3 * i + j];
	double v = b[0][0] * (b[1][1] * b[2][2] - b[1][2] * b[2][1]);
	v += b[0][1] * (b[1][2] * b[2][0] - b[1][0] * b[2][2]);
	v += b[0][2] * (b[1][0] * b[2][1] - b[1][1] * b[2][0]);
	c->volume = v;
	double short_mag = MAXVALUE;
	if (v > 0) {
		for (int i = -15; i <= 15; i++) for (int j = -15; j <= 15; j++) for (int k = -15; k <= 15; k++) {
			if (!i && !j && !k) continue;
			double cv[3];
			for (int p = 0; p < 3; p++) cv[p] = i * b[0][p] + j * b[1][p] + k * b[2][p];
			double m = sqrt(cv[0] * cv[0] + cv[1] * cv[1] + cv[2] * cv[2]);
			if (m < short_mag) short_mag = m;
		}
		c->cutoff = 0.5 * short_mag;
	} else c->cutoff = MAXVALUE;
	double iv = 1.0 / v;
	double (*r)[3] = c->recip;
	r[0][0] = iv * (b[1][1] * b[2][2] - b[1][2] * b[2][1]);
	r[0][1] = iv * (b[0][2] * b[2][1] - b[0][1] * b[2][2]);
	r[0][2] = iv * (b[0][1] * b[1][2] - b[0][2] * b[1][1]);
	r[1][0] = iv * (b[1][2] * b[2][0] - b[1][0] * b[2][2]);
	r[1][1] = iv * (b[0][0] * b[2][2] - b[0][2] * b[2][0]);
	r[1][2] = iv * (b[0][2] * b[1][0] - b[0][0] * b[1][2]);
	r[2][0] = iv * (b[1][0] * b[2][1] - b[1][1] * b[2][0]);
	r[2][1] = iv * (b[0][1] * b[2][0] - b[0][0] * b[2][1]);
	r[2][2] = iv * (b[0][0] * b[1][1] - b[0][1] * b[1][0]);
	c->ewald_alpha = ewald_alpha > 0 ? ewald_alpha : 3.5 / c->cutoff;
	c->polar_ewald_alpha = polar_ewald_alpha > 0 ? polar_ewald_alpha : 3.5 / c->cutoff;
}

void orc_cell(const double basis[9], double ewald_alpha, double polar_ewald_alpha, double cell[22]) {
	cell_t c;
	cell_update(&c, basis, ewald_alpha, polar_ewald_alpha);
	for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { cell[3 * i + j] = c.basis[i][j]; cell[9 + 3 * i + j] = c.recip[i][j]; }
	cell[18] = c.volume; cell[19] = c.cutoff; cell[20] = c.ewald_alpha; cell[21] = c.polar_ewald_alpha;
}

/* System.cpp:1202-1279 minimum_image(): d = r_i - r_j, img = rint(d . recip), dimg = d - img . basis. */
static inline void min_image(const cell_t *c, const double *pi, const double *pj, double *r, double *rimg, double dimg[3]) {
	double d[3], img[3], di[3];
	for (int p = 0; p < 3; p++) d[p] = pi[p] - pj[p];
	for (int p = 0; p < 3; p++) {
		img[p] = 0;
		for (int q = 0; q < 3; q++) img[p] += c->recip[q][p] * d[q];
		img[p] = rint(img[p]);
	}
	for (int p = 0; p < 3; p++) {
		di[p] = 0;
		for (int q = 0; q < 3; q++) di[p] += c->basis[q][p] * img[q];
	}
	for (int p = 0; p < 3; p++) di[p] = d[p] - di[p];
	double r2 = 0, ri2 = 0;
	for (int p = 0; p < 3; p++) { r2 += d[p] * d[p]; ri2 += di[p] * di[p]; }
	*r = sqrt(r2);
	double ri = sqrt(ri2);
	if (isnan(ri)) { *rimg = *r; for (int p = 0; p < 3; p++) dimg[p] = d[p]; }
	else { *rimg = ri; for (int p = 0; p < 3; p++) dimg[p] = di[p]; }
}

typedef struct {
	int n;
	const double *pos, *q, *alpha, *eps, *sigma;
	const int *mol, *frozen;
	const int *iopt;
	const double *dopt;
	cell_t cell;
} sys_t;

/* System.cpp:1035-1177 pair_exclusions(), Lorentz-Berthelot branch only (:1166-1177). A fresh Pair has
 * epsilon = 0 (Pair.h:20-41), and the negative-sigma branch (:1167-1169) never assigns epsilon. */
static inline void pair_mix(const sys_t *s, int i, int j, int *rd_excl, int *es_excl, int *frozen, double *eps, double *sig) {
	int same = s->mol[i] == s->mol[j];
	if (same) { *rd_excl = 1; *es_excl = 1; }
	else {
		*rd_excl = (s->eps[i] == 0.0 || s->sigma[i] == 0.0 || s->eps[j] == 0.0 || s->sigma[j] == 0.0);
		*es_excl = (s->q[i] == 0.0 || s->q[j] == 0.0);
	}
	*frozen = s->frozen[i] && s->frozen[j];
	if (s->sigma[i] < 0.0 || s->sigma[j] < 0.0) { *sig = 0.5 * (fabs(s->sigma[i]) + fabs(s->sigma[j])); *eps = 0.0; }
	else if (s->sigma[i] == 0 || s->sigma[j] == 0) { *sig = 0; *eps = sqrt(s->eps[i] * s->eps[j]); }
	else { *sig = 0.5 * (s->sigma[i] + s->sigma[j]); *eps = sqrt(s->eps[i] * s->eps[j]); }
}

/* System.Energy.cpp:1036-1069 lj_lrc_corr() / :1072-1096 lj_lrc_self(), plain-LJ branch. */
static inline double lrc_formula(double eps, double sigma, double cutoff, double volume) {
	double sig_cut = fabs(sigma) / cutoff;
	double sig3 = fabs(sigma);
	sig3 *= sig3 * sig3;
	double sig_cut3 = sig_cut * sig_cut * sig_cut;
	double sig_cut9 = sig_cut3 * sig_cut3 * sig_cut3;
	return ((16.0 / 3.0) * PI_ * eps * sig3) * ((1.0 / 3.0) * sig_cut9 - sig_cut3) / volume;
}

/* lj(): System.Energy.cpp:897-1032; coulombic_real(): :1466-1517.  One triangular sweep for both. */
static void pair_energies(const sys_t *s, double *rd, double *lrc_pair, double *es_real, double *es_intra, double *n_in, double *abs_sums) {
	const int n = s->n;
	const double cutoff = s->cell.cutoff, a = s->cell.ewald_alpha;
	const int do_es = !s->iopt[ORC_RD_ONLY], polar = s->iopt[ORC_POLARIZATION], do_lrc = s->iopt[ORC_RD_LRC];
	double *acc = (double *)calloc((size_t)n * 8, sizeof(double));
#pragma omp parallel for schedule(dynamic, 16)
	for (int i = 0; i < n - 1; i++) {
		double a_rd = 0, a_lrc = 0, a_re = 0, a_in = 0, a_cnt = 0, b_rd = 0, b_re = 0, b_in = 0;   /* b_*: sums of |term| (scale of the rounding noise; not a reference quantity) */
		for (int j = i + 1; j < n; j++) {
			int rdx, esx, fz; double e, sg;
			pair_mix(s, i, j, &rdx, &esx, &fz, &e, &sg);
			double r = 0, rimg = 0, dimg[3] = {0, 0, 0};
			if (!fz || polar) min_image(&s->cell, s->pos + 3 * i, s->pos + 3 * j, &r, &rimg, dimg);   /* System.cpp:985 */
			/* :929-931 pair LRC (includes intramolecular pairs: exclusions are not consulted, :1045-1050) */
			if (do_lrc && e != 0 && sg != 0 && !fz) a_lrc += lrc_formula(e, sg, cutoff, s->cell.volume);
			/* :934-937, :965-993 */
			if ((rimg - SMALL_dR < cutoff) && !rdx && !fz) {
				double sor = fabs(sg) / rimg;
				double s6 = sor * sor * sor; s6 *= s6;
				double s12 = s6 * s6;
				a_rd += 4.0 * e * (s12 - s6);
				b_rd += fabs(4.0 * e * s12) + fabs(4.0 * e * s6);
				a_cnt += 1;
			}
			if (do_es && !fz) {
				if (!((rimg > cutoff) || esx)) { const double t = s->q[i] * s->q[j] * erfc(a * rimg) / rimg; a_re += t; b_re += fabs(t); }   /* :1490-1497 */
				else if (esx) { const double t = s->q[i] * s->q[j] * erf(a * r) / r; a_in += t; b_in += fabs(t); }                            /* :1503-1504, un-imaged r */
			}
		}
		acc[8 * i] = a_rd; acc[8 * i + 1] = a_lrc; acc[8 * i + 2] = a_re; acc[8 * i + 3] = a_in; acc[8 * i + 4] = a_cnt;
		acc[8 * i + 5] = b_rd; acc[8 * i + 6] = b_re; acc[8 * i + 7] = b_in;
	}
	*rd = *lrc_pair = *es_real = *es_intra = *n_in = 0;
	abs_sums[0] = abs_sums[1] = abs_sums[2] = 0;
	for (int i = 0; i < n; i++) {
		*rd += acc[8 * i]; *lrc_pair += acc[8 * i + 1]; *es_real += acc[8 * i + 2]; *es_intra += acc[8 * i + 3]; *n_in += acc[8 * i + 4];
		for (int q = 0; q < 3; q++) abs_sums[q] += acc[8 * i + 5 + q];
	}
	free(acc);
}

/* hemisphere of k-vectors: System.Energy.cpp:1577-1590 (identical loop at :2849-2860) */
static int kvectors(const cell_t *c, int kmax, double **kout) {
	int cap = (2 * kmax + 1) * (2 * kmax + 1) * (kmax + 1), nk = 0;
	double *k = (double *)malloc(sizeof(double) * 3 * (size_t)cap);
	int l[3];
	for (l[0] = 0; l[0] <= kmax; l[0]++)
		for (l[1] = (!l[0] ? 0 : -kmax); l[1] <= kmax; l[1]++)
			for (l[2] = ((!l[0] && !l[1]) ? 1 : -kmax); l[2] <= kmax; l[2]++) {
				if (l[0] * l[0] + l[1] * l[1] + l[2] * l[2] > kmax * kmax) continue;
				for (int p = 0; p < 3; p++) {
					double kp = 0;
					for (int q = 0; q < 3; q++) kp += 2.0 * PI_ * c->recip[p][q] * l[q];
					k[3 * nk + p] = kp;
				}
				nk++;
			}
	*kout = k;
	return nk;
}

/* coulombic_reciprocal(): System.Energy.cpp:1561-1622.  Frozen and uncharged sites are skipped (:1599-1602). */
static double es_reciprocal(const sys_t *s, const double *kv, int nk) {
	const double a = s->cell.ewald_alpha;
	double *part = (double *)malloc(sizeof(double) * (size_t)nk);
#pragma omp parallel for schedule(static)
	for (int ik = 0; ik < nk; ik++) {
		const double *k = kv + 3 * ik;
		double k2 = k[0] * k[0] + k[1] * k[1] + k[2] * k[2], re = 0, im = 0;
		for (int j = 0; j < s->n; j++) {
			if (s->frozen[j] || s->q[j] == 0.0) continue;
			const double *p = s->pos + 3 * j;
			double dot = k[0] * p[0] + k[1] * p[1] + k[2] * p[2];
			re += s->q[j] * cos(dot);
			im += s->q[j] * sin(dot);
		}
		part[ik] = exp(-k2 / (4.0 * a * a)) / k2 * (re * re + im * im);
	}
	double pot = 0;
	for (int ik = 0; ik < nk; ik++) pot += part[ik];
	free(part);
	return pot * 4.0 * PI_ / s->cell.volume;
}

/* coulombic_self(): System.Energy.cpp:1626-1643 (uses sqrt(pi), not the truncated SqrtPi) */
static double es_self(const sys_t *s) {
	double self = 0;
	for (int i = 0; i < s->n; i++) {
		if (s->frozen[i]) continue;
		self -= s->cell.ewald_alpha * s->q[i] * s->q[i] / sqrt(PI_);
	}
	return self;
}

/* recip_term(): System.Energy.cpp:2834-2896.  All sites, frozen included (:2868-2872). */
static void field_recip(const sys_t *s, const double *kv, int nk, double *ef) {
	const int n = s->n;
	const double ea = s->cell.polar_ewald_alpha;
	double *sre = (double *)malloc(sizeof(double) * (size_t)nk), *sim = (double *)malloc(sizeof(double) * (size_t)nk);
#pragma omp parallel for schedule(static)
	for (int ik = 0; ik < nk; ik++) {
		const double *k = kv + 3 * ik;
		double f1 = 0, f2 = 0;
		for (int j = 0; j < n; j++) {
			const double *p = s->pos + 3 * j;
			double dot = k[0] * p[0] + k[1] * p[1] + k[2] * p[2];
			f1 += s->q[j] * cos(dot);
			f2 += s->q[j] * sin(dot);
		}
		sre[ik] = f1; sim[ik] = f2;
	}
#pragma omp parallel for schedule(static)
	for (int i = 0; i < n; i++) {
		const double *p = s->pos + 3 * i;
		double e[3] = {0, 0, 0};
		for (int ik = 0; ik < nk; ik++) {
			const double *k = kv + 3 * ik;
			double k2 = k[0] * k[0] + k[1] * k[1] + k[2] * k[2];
			double w = exp(-k2 / (4.0 * ea * ea));
			double dot = k[0] * p[0] + k[1] * p[1] + k[2] * p[2];
			double sn = sin(dot), cs = cos(dot);
			for (int q = 0; q < 3; q++) {
				double kw = k[q] / k2 * w;
				e[q] += kw * sn * sre[ik];
				e[q] -= kw * cs * sim[ik];
			}
		}
		for (int q = 0; q < 3; q++) ef[3 * i + q] = e[q] * (8.0 * PI_ / s->cell.volume);   /* :2886-2893 */
	}
	free(sre); free(sim);
}

/* real_term(): System.Energy.cpp:2900-2940, restated per ordered pair (dimg(j,i) = -dimg(i,j) exactly). */
static void field_real(const sys_t *s, double *ef) {
	const int n = s->n;
	const double a = s->cell.polar_ewald_alpha, cutoff = s->cell.cutoff;
#pragma omp parallel for schedule(dynamic, 16)
	for (int i = 0; i < n; i++) {
		double e[3] = {0, 0, 0};
		for (int j = 0; j < n; j++) {
			if (j == i) continue;
			if (s->frozen[i] && s->frozen[j]) continue;                      /* :2915 */
			double r, rimg, dimg[3];
			min_image(&s->cell, s->pos + 3 * i, s->pos + 3 * j, &r, &rimg, dimg);
			if ((rimg > cutoff) || (rimg == 0.0)) continue;                   /* :2917 */
			double r2 = rimg * rimg, factor;
			int esx = (s->mol[i] == s->mol[j]) || s->q[i] == 0.0 || s->q[j] == 0.0;
			if (esx) factor = (2.0 * a * ONE_OVER_SQRT_PI * exp(-a * a * r2) * rimg - erf(a * rimg)) / (rimg * r2);   /* :2921 */
			else     factor = (2.0 * a * ONE_OVER_SQRT_PI * exp(-a * a * r2) * rimg + erfc(a * rimg)) / (r2 * rimg);  /* :2929 */
			for (int p = 0; p < 3; p++) e[p] += factor * s->q[j] * dimg[p];
		}
		for (int p = 0; p < 3; p++) ef[3 * i + p] += e[p];
	}
}

/* thole_field_nopbc(): System.Energy.cpp:3300-3333 (polar_ewald off) */
static void field_nopbc(const sys_t *s, double *ef) {
	const int n = s->n;
	const double cutoff = s->cell.cutoff;
#pragma omp parallel for schedule(dynamic, 16)
	for (int i = 0; i < n; i++) {
		double e[3] = {0, 0, 0};
		for (int j = 0; j < n; j++) {
			if (j == i || (s->frozen[i] && s->frozen[j]) || s->mol[i] == s->mol[j]) continue;
			double r, rimg, dimg[3];
			min_image(&s->cell, s->pos + 3 * i, s->pos + 3 * j, &r, &rimg, dimg);
			if ((rimg - SMALL_dR < cutoff) && (rimg != 0.)) for (int p = 0; p < 3; p++) e[p] += s->q[j] * dimg[p] / (rimg * rimg * rimg);
		}
		for (int p = 0; p < 3; p++) ef[3 * i + p] += e[p];
	}
}

/* One 3x3 block of thole_amatrix() (System.Energy.cpp:2694-2767) contracted with mu_j:
 * out += T_ij mu_j, T[p][q] = delta_pq damp1/r^3 - 3 d_p d_q damp2/r^5 on rimg/dimg, no cutoff (q6). */
static inline void tensor_dot(const sys_t *s, int i, int j, const double *mu_j, double out[3]) {
	double r, rimg, d[3], ir, ir3, ir5, damp1 = 1, damp2 = 1;
	min_image(&s->cell, s->pos + 3 * i, s->pos + 3 * j, &r, &rimg, d);
	double r2 = rimg * rimg;
	if (rimg == 0.) { ir = 0; ir3 = ir5 = MAXVALUE; }
	else { ir = 1.0 / rimg; ir3 = ir * ir * ir; ir5 = ir3 * ir * ir; }
	const double l = s->dopt[ORC_POLAR_DAMP];
	switch (s->iopt[ORC_DAMP_TYPE]) {
	case 0: {
		int esx = (s->mol[i] == s->mol[j]) || s->q[i] == 0.0 || s->q[j] == 0.0;
		damp1 = damp2 = esx ? 0.0 : 1.0;
	} break;
	case 1: {
		double sc = l * pow(s->alpha[i] * s->alpha[j], 1.0 / 6.0), v = rimg / sc;
		if (rimg < sc) { damp1 = (4.0 - 3.0 * v) * v * v * v; damp2 = v * v * v * v; } else damp1 = damp2 = 1.0;
	} break;
	default: {
		double l2 = l * l, l3 = l2 * l, explr = exp(-l * rimg);
		damp1 = 1.0 - explr * (0.5 * l2 * r2 + l * rimg + 1.0);
		damp2 = damp1 - explr * (l3 * r2 * rimg / 6.0);
	}
	}
	for (int p = 0; p < 3; p++) {
		double A[3];
		for (int q = 0; q < 3; q++) {
			A[q] = -3.0 * d[p] * d[q] * damp2 * ir5;
			if (p == q) A[q] += damp1 * ir3;
		}
		out[p] += A[0] * mu_j[0] + A[1] * mu_j[1] + A[2] * mu_j[2];      /* UsefulMath::dddotprod */
	}
}

/* rank metric: System.cpp:1000-1029 (rmin from rimg, the count test uses the UN-imaged r) */
static void rank_metric(const sys_t *s, double *rank) {
	const int n = s->n;
	double rmin = MAXVALUE;
	for (int i = 0; i < n; i++) rank[i] = 0;
	for (int i = 0; i < n; i++) {
		if (s->alpha[i] == 0.0) continue;
		for (int j = i + 1; j < n; j++) {
			if (s->alpha[j] == 0.0) continue;
			double r, rimg, d[3];
			min_image(&s->cell, s->pos + 3 * i, s->pos + 3 * j, &r, &rimg, d);
			if (rimg < rmin) rmin = rimg;
		}
	}
	for (int i = 0; i < n; i++) {
		if (s->alpha[i] == 0.0) continue;
		for (int j = i + 1; j < n; j++) {
			if (s->alpha[j] == 0.0) continue;
			double r, rimg, d[3];
			min_image(&s->cell, s->pos + 3 * i, s->pos + 3 * j, &r, &rimg, d);
			if (r <= rmin * 1.5) { rank[i] += 1.0; rank[j] += 1.0; }
		}
	}
}

/* contract_dipoles(): System.Energy.cpp:3564-3598 (Jacobi is parallel over i; GS is sequential in ranked order) */
static void contract(const sys_t *s, const int *ranked, const double *efs, double *efi, double *mu, double *new_mu) {
	const int n = s->n, gs = s->iopt[ORC_POLAR_GS] || s->iopt[ORC_POLAR_GS_RANKED];
	if (!gs) {
#pragma omp parallel for schedule(dynamic, 8)
		for (int i = 0; i < n; i++) {
			if (s->alpha[i] == 0) { for (int p = 0; p < 3; p++) new_mu[3 * i + p] = 0; continue; }
			double acc[3] = {0, 0, 0};
			for (int j = 0; j < n; j++) if (j != i && s->alpha[j] != 0) tensor_dot(s, i, j, mu + 3 * j, acc);
			for (int p = 0; p < 3; p++) { efi[3 * i + p] -= acc[p]; new_mu[3 * i + p] = s->alpha[i] * (efs[3 * i + p] + efi[3 * i + p]); }
		}
		for (int i = 0; i < n; i++) if (s->alpha[i] == 0) for (int p = 0; p < 3; p++) mu[3 * i + p] = 0;
		return;
	}
	for (int ii = 0; ii < n; ii++) {
		int i = ranked[ii];
		if (s->alpha[i] == 0) { for (int p = 0; p < 3; p++) new_mu[3 * i + p] = mu[3 * i + p] = 0; continue; }
		double a0 = 0, a1 = 0, a2 = 0;
#pragma omp parallel for schedule(static) reduction(+ : a0, a1, a2)
		for (int j = 0; j < n; j++) {
			if (j == i || s->alpha[j] == 0) continue;
			double t[3] = {0, 0, 0};
			tensor_dot(s, i, j, mu + 3 * j, t);
			a0 += t[0]; a1 += t[1]; a2 += t[2];
		}
		efi[3 * i] -= a0; efi[3 * i + 1] -= a1; efi[3 * i + 2] -= a2;
		for (int p = 0; p < 3; p++) {
			new_mu[3 * i + p] = s->alpha[i] * (efs[3 * i + p] + efi[3 * i + p]);
			mu[3 * i + p] = new_mu[3 * i + p];                                  /* :3590-3592 */
		}
	}
}

/* palmo_contraction(): System.Energy.cpp:3602-3627 */
static void palmo(const sys_t *s, const double *efi, const double *mu, double *efic) {
	const int n = s->n;
#pragma omp parallel for schedule(dynamic, 8)
	for (int i = 0; i < n; i++) {
		double acc[3] = {0, 0, 0};
		for (int j = 0; j < n; j++) if (j != i && s->alpha[j] != 0) tensor_dot(s, i, j, mu + 3 * j, acc);
		for (int p = 0; p < 3; p++) efic[3 * i + p] = -efi[3 * i + p] - acc[p];
	}
}

/* update_ranking(): System.Energy.cpp:3631-3656 — bubble sort, descending, stable */
static void update_ranking(int n, const double *rank, int *ranked) {
	for (int i = 0; i < n; i++) {
		int sorted = 1;
		for (int j = 0; j < n - 1; j++)
			if (rank[ranked[j]] < rank[ranked[j + 1]]) { sorted = 0; int t = ranked[j]; ranked[j] = ranked[j + 1]; ranked[j + 1] = t; }
		if (sorted) break;
	}
}

/* polar(): System.Energy.cpp:2534-2635 with thole_field (:3271), thole_iterative (:3450-3543),
 * init_dipoles (:3547), calc_dipole_rrms (:3147), are_we_done_yet (:3215), get_dipole_rrms (:2639). */
static double polar(const sys_t *s, const double *kv, int nk, double *mu, double *efs, double *efi, double *efic,
                    double *rankm, int *iterations, double *rrms_out, int *failed) {
	const int n = s->n;
	const int *io = s->iopt;
	const double gamma = s->dopt[ORC_POLAR_GAMMA], precision = s->dopt[ORC_POLAR_PRECISION];
	double *old_mu = (double *)calloc((size_t)3 * n, sizeof(double)), *new_mu = (double *)calloc((size_t)3 * n, sizeof(double));
	double *rrms = (double *)calloc((size_t)n, sizeof(double));
	int *ranked = (int *)malloc(sizeof(int) * (size_t)n);
	memset(efs, 0, sizeof(double) * 3 * (size_t)n);
	memset(efi, 0, sizeof(double) * 3 * (size_t)n);
	memset(efic, 0, sizeof(double) * 3 * (size_t)n);
	if (io[ORC_POLAR_EWALD]) { field_recip(s, kv, nk, efs); field_real(s, efs); }     /* ewald_estatic :3400 */
	else field_nopbc(s, efs);
	if (io[ORC_POLAR_ITERATIVE] && io[ORC_POLAR_GS_RANKED]) rank_metric(s, rankm); else for (int i = 0; i < n; i++) rankm[i] = 0;
	for (int i = 0; i < n; i++) ranked[i] = i;
	for (int i = 0; i < 3 * n; i++) {                                                 /* init_dipoles */
		mu[i] = s->alpha[i / 3] * efs[i];
		if (!io[ORC_POLAR_SOR] && !io[ORC_POLAR_ESOR]) mu[i] *= gamma;
	}
	int it = 0, keep = !io[ORC_POLAR_ZODID];
	*failed = 0;
	while (keep) {
		it++;
		if (it >= MAX_ITERATION_COUNT && precision) {                                  /* :3483-3494 */
			for (int i = 0; i < 3 * n; i++) { mu[i] = s->alpha[i / 3] * efs[i]; efic[i] = 0; }
			*failed = 1;
			break;
		}
		memset(efi, 0, sizeof(double) * 3 * (size_t)n);
		if (io[ORC_POLAR_RRMS] || precision > 0 || io[ORC_POLAR_SOR] || io[ORC_POLAR_ESOR]) memcpy(old_mu, mu, sizeof(double) * 3 * (size_t)n);
		contract(s, ranked, efs, efi, mu, new_mu);
		if (io[ORC_POLAR_RRMS] || precision > 0)
			for (int i = 0; i < n; i++) {
				double c = 0;
				for (int p = 0; p < 3; p++) { double e = new_mu[3 * i + p] - old_mu[3 * i + p]; c += e * e; }
				c /= new_mu[3 * i] * new_mu[3 * i] + new_mu[3 * i + 1] * new_mu[3 * i + 1] + new_mu[3 * i + 2] * new_mu[3 * i + 2];
				c = sqrt(c);
				rrms[i] = isfinite(c) ? c : 0;
			}
		if (precision == 0.0) keep = (it != io[ORC_POLAR_MAX_ITER]);                   /* are_we_done_yet */
		else {
			double allowed = precision * precision * DEBYE2SKA * DEBYE2SKA;
			keep = 0;
			for (int i = 0; i < 3 * n && !keep; i++) { double e = new_mu[i] - old_mu[i]; if (e * e > allowed) keep = 1; }
		}
		if (io[ORC_POLAR_PALMO] && !keep) palmo(s, efi, mu, efic);
		if (io[ORC_POLAR_GS_RANKED] && keep) update_ranking(n, rankm, ranked);
		for (int i = 0; i < 3 * n; i++) {                                              /* :3526-3536 */
			if (io[ORC_POLAR_SOR]) mu[i] = gamma * new_mu[i] + (1.0 - gamma) * old_mu[i];
			else if (io[ORC_POLAR_ESOR]) mu[i] = (1.0 - exp(-gamma * it)) * new_mu[i] + exp(-gamma * it) * old_mu[i];
			else mu[i] = new_mu[i];
		}
	}
	*iterations = it;
	double rr = 0;
	for (int i = 0; i < n; i++) if (isfinite(rrms[i])) rr += rrms[i];
	*rrms_out = rr / n;
	double pot = 0;
	for (int i = 0; i < n; i++) {                                                      /* :2609-2618 */
		pot += mu[3 * i] * efs[3 * i] + mu[3 * i + 1] * efs[3 * i + 1] + mu[3 * i + 2] * efs[3 * i + 2];
		if (io[ORC_POLAR_PALMO]) pot += mu[3 * i] * efic[3 * i] + mu[3 * i + 1] * efic[3 * i + 1] + mu[3 * i + 2] * efic[3 * i + 2];
	}
	free(old_mu); free(new_mu); free(rrms); free(ranked);
	return -0.5 * pot;
}

/* energy(): System.Energy.cpp:19-171 (LJ + Ewald + Thole branch only) */
int orc_energy(int n, const double *pos, const double *charge, const double *alpha, const double *eps,
               const double *sigma, const int *mol, const int *frozen, const double basis[9],
               const int *iopt, const double *dopt, double *out,
               double *mu, double *ef_static, double *ef_induced, double *ef_induced_change, double *rankm) {
	sys_t s = {n, pos, charge, alpha, eps, sigma, mol, frozen, iopt, dopt, {{{0}}, {{0}}, 0, 0, 0, 0}};
	cell_update(&s.cell, basis, dopt[ORC_EWALD_ALPHA], dopt[ORC_POLAR_EWALD_ALPHA]);
	for (int i = 0; i < ORC_NOUT; i++) out[i] = 0;
	double *kv = NULL;
	int nk = kvectors(&s.cell, iopt[ORC_EWALD_KMAX], &kv);
	double rd, lrcp, esr, esi, nin;
	double abs_sums[3];
	pair_energies(&s, &rd, &lrcp, &esr, &esi, &nin, abs_sums);
	double lrcs = 0;
	if (iopt[ORC_RD_LRC])
		for (int i = 0; i < n; i++)
			if (sigma[i] != 0 && eps[i] != 0 && !frozen[i]) lrcs += lrc_formula(eps[i], sigma[i], s.cell.cutoff, s.cell.volume);
	double coul = 0, pol = 0;
	if (!iopt[ORC_RD_ONLY]) {
		out[ORC_O_ES_REAL] = esr; out[ORC_O_ES_SELF_INTRA] = esi;
		out[ORC_O_ES_RECIP] = es_reciprocal(&s, kv, nk);
		out[ORC_O_ES_SELF] = es_self(&s);
		coul = (esr - esi) + out[ORC_O_ES_RECIP] + out[ORC_O_ES_SELF];                 /* :1407-1412 */
		if (iopt[ORC_POLARIZATION]) {
			double *b = (double *)calloc((size_t)13 * n, sizeof(double));
			double *m = mu ? mu : b, *es = ef_static ? ef_static : b + 3 * n, *ei = ef_induced ? ef_induced : b + 6 * n;
			double *ec = ef_induced_change ? ef_induced_change : b + 9 * n, *rk = rankm ? rankm : b + 12 * n;
			int it = 0, failed = 0; double rr = 0;
			pol = polar(&s, kv, nk, m, es, ei, ec, rk, &it, &rr, &failed);
			out[ORC_O_ITERATIONS] = it; out[ORC_O_DIPOLE_RRMS] = rr; out[ORC_O_ITERATOR_FAILED] = failed;
			free(b);
		}
	}
	out[ORC_O_RD_PAIR] = rd; out[ORC_O_LRC_PAIR] = lrcp; out[ORC_O_LRC_SELF] = lrcs;
	out[ORC_O_RD_TOTAL] = rd + lrcp + lrcs;
	out[ORC_O_COULOMBIC] = coul; out[ORC_O_POLAR] = pol;
	out[ORC_O_ENERGY] = out[ORC_O_RD_TOTAL] + coul + pol;                               /* :136 */
	out[ORC_O_VOLUME] = s.cell.volume; out[ORC_O_CUTOFF] = s.cell.cutoff;
	out[ORC_O_EWALD_ALPHA] = s.cell.ewald_alpha; out[ORC_O_POLAR_EWALD_ALPHA] = s.cell.polar_ewald_alpha;
	out[ORC_O_NKVEC] = nk; out[ORC_O_NPAIR_IN_CUTOFF] = nin;
	out[ORC_O_RD_ABS] = abs_sums[0]; out[ORC_O_ES_REAL_ABS] = abs_sums[1]; out[ORC_O_ES_INTRA_ABS] = abs_sums[2];
	free(kv);
	return 0;
}

/* PI_calculate_potential (SimulationControl.PathIntegral.cpp:752-805), PI_chain_mass_length2_ENTIRE_SYSTEM
 * (:859-904) + PI_chain_mass_length2 (:916-970), PI_calculate_kinetic (:810-828). */
int orc_pi_energy(int P, int n, const double *pos, const double *charge, const double *alpha, const double *eps,
                  const double *sigma, const double *mass, const int *mol, const int *frozen, const double basis[9],
                  const int *iopt, const double *dopt, double temperature, double *out, double *per_bead) {
	double rd = 0, es = 0, pol = 0;
	for (int b = 0; b < P; b++) {
		double o[ORC_NOUT];
		orc_energy(n, pos + (size_t)3 * n * b, charge, alpha, eps, sigma, mol, frozen, basis, iopt, dopt, o, NULL, NULL, NULL, NULL, NULL);
		rd += o[ORC_O_RD_TOTAL]; es += o[ORC_O_COULOMBIC]; pol += o[ORC_O_POLAR];
		if (per_bead) { per_bead[4 * b] = o[ORC_O_RD_TOTAL]; per_bead[4 * b + 1] = o[ORC_O_COULOMBIC]; per_bead[4 * b + 2] = o[ORC_O_POLAR]; per_bead[4 * b + 3] = 0; }
	}
	rd /= P; es /= P; pol /= P;
	out[1] = rd; out[2] = es; out[3] = pol; out[4] = 0;
	out[0] = rd + es + 0 + pol;
	/* chain: molecules are runs of equal mol[]; frozen molecules are skipped (:885) */
	double sum = 0; int nmobile = 0;
	double *com = (double *)malloc(sizeof(double) * 3 * (size_t)P);
	for (int a = 0; a < n;) {
		int e = a; while (e < n && mol[e] == mol[a]) e++;
		if (!frozen[a]) {
			nmobile++;
			double m = 0;
			for (int b = 0; b < P; b++) {
				const double *pb = pos + (size_t)3 * n * b;
				double c[3] = {0, 0, 0}; m = 0;
				for (int i = a; i < e; i++) { m += mass[i]; for (int p = 0; p < 3; p++) c[p] += mass[i] * pb[3 * i + p]; }
				for (int p = 0; p < 3; p++) com[3 * b + p] = c[p] / m;
			}
			double len = 0;
			for (int b = 0; b < P; b++) {
				int b2 = (b + 1) % P;
				double dx = com[3 * b] - com[3 * b2], dy = com[3 * b + 1] - com[3 * b2 + 1], dz = com[3 * b + 2] - com[3 * b2 + 2];
				len += dx * dx + dy * dy + dz * dz;
			}
			sum += len * (m * AMU2KG) * (ANGSTROM2METER * ANGSTROM2METER);
		}
		a = e;
	}
	free(com);
	out[5] = sum;
	double beta = 1.0 / (kB * temperature), omega2 = P / (beta * beta * hBar2);
	double t1 = 0.5 * 3.0 * nmobile * kB * temperature * P, t2 = 0.5 * omega2 * sum;
	out[6] = (1.0 / kB) * (t1 - t2);
	return 0;
}
