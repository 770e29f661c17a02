// kernels_polar.cuh — Thole induced-dipole solve without ever materialising the 3N x 3N A matrix
// (reference: thole_amatrix + thole_iterative + contract_dipoles + palmo_contraction, src/System.Energy.cpp:2661-2770,
// 3450-3656).  T_ij is recomputed from the pair geometry inside every contraction: the kernels are FP64-pipe bound
// and read only O(N) bytes per sweep.
#pragma once
#include <cooperative_groups.h>
#include "kernels_pair.cuh"

namespace mpmc {
namespace cg = cooperative_groups;

// init_dipoles() (System.Energy.cpp:3547-3560): mu = alpha E_static (* gamma unless SOR/ESOR); clears the per-sweep fields.
__global__ void k_dipole_init(const double *__restrict__ alpha, const double *__restrict__ efs, int n, int nbeads, double gamma_init,
                              double *__restrict__ mu, double *__restrict__ new_mu, double *__restrict__ old_mu,
                              double *__restrict__ efi, double *__restrict__ efic, double *__restrict__ rrms) {
	const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= (size_t)n * nbeads * 3) return;
	const int i = (int)((t / 3) % n);
	const double m = alpha[i] * efs[t] * gamma_init;
	mu[t] = m; new_mu[t] = m; old_mu[t] = 0; efi[t] = 0; efic[t] = 0;
	if (t % 3 == 0) rrms[t / 3] = 0;
}

// modes of the contraction sweeps (kernels_polar2.cuh: k_contract_finish)
enum { SWEEP_JACOBI = 0, SWEEP_PALMO = 1, SWEEP_PALMO_NONPOLAR = 2, SWEEP_ACC = 3 };

// calc_dipole_rrms() (:3147-3177) and the precision branch of are_we_done_yet() (:3227-3236).  flags[bead] is set to 1 when
// some component still moves by more than the allowed error.
__global__ void k_dipole_check(const double *__restrict__ new_mu, const double *__restrict__ old_mu, int n, int nbeads,
                               int want_rrms, double allowed_sqerr, double *__restrict__ rrms, int *__restrict__ flags) {
	const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= (size_t)n * nbeads) return;
	const double *nm = new_mu + 3 * s, *om = old_mu + 3 * s;
	double carry = 0, nn = 0;
	bool broke = false;
	for (int q = 0; q < 3; q++) {
		const double e = nm[q] - om[q];
		carry += e * e; nn += nm[q] * nm[q];
		if (e * e > allowed_sqerr) broke = true;
	}
	if (want_rrms) {
		double r = sqrt(carry / nn);
		rrms[s] = isfinite(r) ? r : 0.0;
	}
	if (allowed_sqerr > 0 && broke) flags[s / n] = 1;
}

// "save the dipoles for the next pass" (:3526-3536): plain, SOR or ESOR relaxation.  esor_w = exp(-gamma * iteration).
__global__ void k_mu_update(const double *__restrict__ new_mu, const double *__restrict__ old_mu, size_t len, int sor, int esor,
                            double gamma, double esor_w, double *__restrict__ mu) {
	const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= len) return;
	if (sor) mu[t] = gamma * new_mu[t] + (1.0 - gamma) * old_mu[t];
	else if (esor) mu[t] = (1.0 - esor_w) * new_mu[t] + esor_w * old_mu[t];
	else mu[t] = new_mu[t];
}

// convergence failure after MAX_ITERATION_COUNT (:3483-3494): mu = alpha E_static, efic = 0
__global__ void k_dipole_fail(const double *__restrict__ alpha, const double *__restrict__ efs, int n, size_t off, double *__restrict__ mu,
                              double *__restrict__ efic) {
	const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= (size_t)n * 3) return;
	mu[off + t] = alpha[t / 3] * efs[off + t];
	efic[off + t] = 0;
}

// tail of polar() (:2609-2618) and get_dipole_rrms() (:2639-2657): result record slots 5..7 = sum mu.E_s, sum mu.dE_ind, sum rrms
__global__ void __launch_bounds__(1024)
k_polar_energy(const double *__restrict__ mu, const double *__restrict__ efs, const double *__restrict__ efic,
               const double *__restrict__ rrms, int n, double *__restrict__ out) {
	__shared__ double s_red[3][1024];
	const int bead = blockIdx.x;
	double a = 0, b = 0, r = 0;
	for (int i = threadIdx.x; i < n; i += blockDim.x) {
		const size_t o = ((size_t)bead * n + i) * 3;
		a += mu[o] * efs[o] + mu[o + 1] * efs[o + 1] + mu[o + 2] * efs[o + 2];
		b += mu[o] * efic[o] + mu[o + 1] * efic[o + 1] + mu[o + 2] * efic[o + 2];
		const double rr = rrms[(size_t)bead * n + i];
		if (isfinite(rr)) r += rr;
	}
	s_red[0][threadIdx.x] = a; s_red[1][threadIdx.x] = b; s_red[2][threadIdx.x] = r;
	__syncthreads();
	for (int o = blockDim.x / 2; o > 0; o >>= 1) {
		if ((int)threadIdx.x < o)
			for (int q = 0; q < 3; q++) s_red[q][threadIdx.x] += s_red[q][threadIdx.x + o];
		__syncthreads();
	}
	if (threadIdx.x == 0) { out[kResStride * bead + 5] = s_red[0][0]; out[kResStride * bead + 6] = s_red[1][0]; out[kResStride * bead + 7] = s_red[2][0]; }
}

// stable descending order of the polarizable sites by rank metric (update_ranking, :3631-3656, restricted to alpha != 0:
// the other sites only ever set mu = 0).  order[pos] = site.
__global__ void __launch_bounds__(kOrdThreads)
k_rank_order_plist(const double *__restrict__ rank, const int *__restrict__ plist, int np, int *__restrict__ order) {
	// CTA = 32 sites x 8 j-lanes (the first version walked all np partners with one thread per site: 127 us at np = 9200)
	__shared__ double s_m[kOrdTileJ];
	const int tid = threadIdx.x, jl = tid % kOrdJ, il = tid / kOrdJ;
	const int t = blockIdx.x * kOrdI + il;
	const int i = t < np ? plist[t] : 0;
	const double mi = t < np ? rank[i] : 0;
	int pos = 0;
	for (int j0 = 0; j0 < np; j0 += kOrdTileJ) {
		__syncthreads();
		if (j0 + tid < np) s_m[tid] = rank[plist[j0 + tid]];
		__syncthreads();
		const int jn = min(kOrdTileJ, np - j0);
		for (int jj = jl; jj < jn; jj += kOrdJ) {
			const double mj = s_m[jj];
			pos += (mj > mi) || (mj == mi && j0 + jj < t);
		}
	}
#pragma unroll
	for (int o = kOrdJ / 2; o > 0; o >>= 1) pos += __shfl_xor_sync(0xffffffffu, pos, o);
	if (jl == 0 && t < np) order[pos] = i;
}

// bead-chain bookkeeping ------------------------------------------------------------------------------------
// Molecule::update_COM (src/Molecule.cpp:256-281): com = sum m r / sum m over the sites of one molecule, in list order
__global__ void k_mol_com(const double4 *__restrict__ posq, int stride, const double *__restrict__ mass,
                          const int *__restrict__ mol_start, int nmol, int nbeads, double *__restrict__ com, double *__restrict__ mol_mass) {
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= nmol * nbeads) return;
	const int bead = t / nmol, m = t % nmol;
	const double4 *pq = posq + (size_t)bead * stride;
	double ms = 0, cx = 0, cy = 0, cz = 0;
	for (int i = mol_start[m]; i < mol_start[m + 1]; i++) {
		const double w = mass[i];
		ms = __dadd_rn(ms, w);
		cx = __dadd_rn(cx, __dmul_rn(w, pq[i].x)); cy = __dadd_rn(cy, __dmul_rn(w, pq[i].y)); cz = __dadd_rn(cz, __dmul_rn(w, pq[i].z));
	}
	com[3 * (size_t)t] = cx / ms; com[3 * (size_t)t + 1] = cy / ms; com[3 * (size_t)t + 2] = cz / ms;
	if (bead == 0) mol_mass[m] = ms;
}

// PI_chain_mass_length2 (src/SimulationControl.PathIntegral.cpp:916-970) for every mobile molecule over the LOCAL links
// b -> b+1 (and last -> first when closed); per-molecule results, summed by the caller in molecule order.
__global__ void k_chain_len2(const double *__restrict__ com, const double *__restrict__ mol_mass, const unsigned char *__restrict__ mol_mobile,
                             int nmol, int nbeads, int closed, double *__restrict__ per_mol) {
	const int m = blockIdx.x * blockDim.x + threadIdx.x;
	if (m >= nmol) return;
	double len = 0;
	if (mol_mobile[m]) {
		const int links = closed ? nbeads : nbeads - 1;
		for (int b = 0; b < links; b++) {
			const int b2 = (b + 1) % nbeads;
			const double *a = com + 3 * ((size_t)b * nmol + m), *z = com + 3 * ((size_t)b2 * nmol + m);
			const double dx = a[0] - z[0], dy = a[1] - z[1], dz = a[2] - z[2];
			len += dx * dx + dy * dy + dz * dz;
		}
		len *= (mol_mass[m] * 1.66053873e-27) * (1.0e-10 * 1.0e-10);   // AMU2KG, ANGSTROM2METER^2 (constants.h:35,27)
	}
	per_mol[m] = len;
}

// the link between this rank's last bead and the next rank's first bead (bead chains sharded over GPUs)
__global__ void k_chain_boundary(const double *__restrict__ com_last, const double *__restrict__ com_next_first, const double *__restrict__ mol_mass,
                                 const unsigned char *__restrict__ mol_mobile, int nmol, double *__restrict__ per_mol) {
	const int m = blockIdx.x * blockDim.x + threadIdx.x;
	if (m >= nmol || !mol_mobile[m]) return;
	const double dx = com_last[3 * m] - com_next_first[3 * m], dy = com_last[3 * m + 1] - com_next_first[3 * m + 1], dz = com_last[3 * m + 2] - com_next_first[3 * m + 2];
	per_mol[m] += (dx * dx + dy * dy + dz * dz) * (mol_mass[m] * 1.66053873e-27) * (1.0e-10 * 1.0e-10);
}

// out[0] = sum of v[0..n) in a fixed order (one CTA)
__global__ void k_sum_array(const double *__restrict__ v, int n, double *__restrict__ out) {
	__shared__ double s_red[256];
	double a = 0;
	for (int i = threadIdx.x; i < n; i += blockDim.x) a += v[i];
	s_red[threadIdx.x] = a;
	__syncthreads();
	for (int o = blockDim.x / 2; o > 0; o >>= 1) {
		if ((int)threadIdx.x < o) s_red[threadIdx.x] += s_red[threadIdx.x + o];
		__syncthreads();
	}
	if (threadIdx.x == 0) out[0] = s_red[0];
}

// FP64 FMA peak probe: 8 independent chains per thread, 4096 x 8 FMAs each
__global__ void k_fp64_probe(double *out, int iters) {
	double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
	const double m = 1.0000001, b = 1e-7;
	for (int i = 0; i < iters; i++) {
		a0 = fma(a0, m, b); a1 = fma(a1, m, b); a2 = fma(a2, m, b); a3 = fma(a3, m, b);
		a4 = fma(a4, m, b); a5 = fma(a5, m, b); a6 = fma(a6, m, b); a7 = fma(a7, m, b);
	}
	out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

} // namespace mpmc
