// kernels_polar2.cuh — the dipole-field contraction acc_i = sum_{j != i} T_ij mu_j, second generation
// (reference: thole_amatrix + contract_dipoles + palmo_contraction, src/System.Energy.cpp:2661-2770, 3564-3627; the 3N x 3N
// matrix is never formed).
//
//   * work is cut in (32-row block) x (column part) units, one CTA each, so that the hardware scheduler keeps 148 SMs busy to the
//     end (a plain one-CTA-per-row-block launch is 2.1 waves at N = 10^4: 30 % of the machine idles in the tail); the per-part
//     sums are combined in a fixed order by k_contract_finish, which also applies the caller's epilogue;
//   * rows come from a list (polarizable sites, non-polarizable sites, or everything), so no CTA idles on rows that are not wanted;
//   * with exponential damping (the `cuda on` configuration) the pair geometry uses FMA and the damping factors are evaluated only
//     where they differ from 1 (device_math.cuh: tensor_contract_exp): ~37-45 FP64 instructions per ordered pair instead of ~75.
#pragma once
#include "kernels_polar.cuh"

namespace mpmc {

constexpr int kCtTile = 256;

template <bool ORTHO, bool EXPD>
__global__ void __launch_bounds__(kOrdThreads)
k_contract_parts(const double4 *__restrict__ posq, const double *__restrict__ alpha, const int *__restrict__ meta,
                 const int *__restrict__ plist, int np, int part_len, const int *__restrict__ rowlist, int nrows, int n, int stride,
                 CellDev c, PolarDev p, const double *__restrict__ mu, double *__restrict__ part) {
	__shared__ double4 s_a[kCtTile];     // x, y, z, mu_x
	__shared__ double2 s_b[kCtTile];     // mu_y, mu_z
	__shared__ double  s_al[kCtTile];
	__shared__ int     s_meta[kCtTile];
	__shared__ int     s_idx[kCtTile];
	const int bead = blockIdx.z, q = blockIdx.y;
	const double4 *pq = posq + (size_t)bead * stride;
	const double *mub = mu + (size_t)bead * n * 3;
	const int tid = threadIdx.x, jl = tid % kOrdJ, il = tid / kOrdJ;
	const int ri = blockIdx.x * kOrdI + il;
	const int i = ri < nrows ? (rowlist ? rowlist[ri] : ri) : -1;
	double4 pi = make_double4(0, 0, 0, 0);
	double ai = 0; int mi = 0;
	if (i >= 0) { pi = pq[i]; ai = alpha[i]; mi = meta[i]; }
	double ax = 0, ay = 0, az = 0;
	const int jbeg = q * part_len, jend = min(np, jbeg + part_len);
	for (int j0 = jbeg; j0 < jend; j0 += kCtTile) {
		__syncthreads();
		if (j0 + tid < jend) {
			const int j = plist[j0 + tid];
			const double4 pj = pq[j];
			s_a[tid] = make_double4(pj.x, pj.y, pj.z, mub[3 * j]);
			s_b[tid] = make_double2(mub[3 * j + 1], mub[3 * j + 2]);
			if (!EXPD) { s_al[tid] = alpha[j]; s_meta[tid] = meta[j] | (pj.w != 0.0 ? 0x40000000 : 0); s_idx[tid] = j; }
		}
		__syncthreads();
		const int jn = min(kCtTile, jend - j0);
		if (i >= 0) {
			if (EXPD) {
#pragma unroll 2
				for (int jj = jl; jj < jn; jj += kOrdJ) {
					const double4 a = s_a[jj];
					const double2 b = s_b[jj];
					tensor_contract_exp<ORTHO>(c, p.damp, p.u_damp, pi.x, pi.y, pi.z, a.x, a.y, a.z, a.w, b.x, b.y, ax, ay, az);   // i == j: r = 0 adds nothing
				}
			} else {
				for (int jj = jl; jj < jn; jj += kOrdJ) {
					if (s_idx[jj] == i) continue;
					const double4 a = s_a[jj];
					const double2 b = s_b[jj];
					const int mj = s_meta[jj];
					const bool excl = (meta_mol(mi) == (mj & 0x3fffffff)) || pi.w == 0.0 || !(mj & 0x40000000);
					tensor_contract<ORTHO>(c, p, pi.x, pi.y, pi.z, a.x, a.y, a.z, excl, ai * s_al[jj], a.w, b.x, b.y, ax, ay, az);
				}
			}
		}
	}
	ax = jlane_sum(ax); ay = jlane_sum(ay); az = jlane_sum(az);
	if (jl == 0 && i >= 0) {
		double *o = part + (((size_t)q * gridDim.z + bead) * nrows + ri) * 3;
		o[0] = ax; o[1] = ay; o[2] = az;
	}
}

// sum the column parts in a fixed order and apply the caller's epilogue:
//   SWEEP_JACOBI : contract_dipoles() in Jacobi form (:3564-3598): efi = -acc, new_mu = alpha (E_s + efi)   [rows = polarizable sites]
//   SWEEP_PALMO / SWEEP_PALMO_NONPOLAR : palmo_contraction() (:3602-3627): efic = -efi - acc              [all rows / alpha == 0 rows]
//   SWEEP_ACC    : out_acc = acc (the running contraction the Gauss-Seidel pipeline keeps up to date)      [rows = polarizable sites]
template <int MODE>
__global__ void k_contract_finish(const double *__restrict__ part, int nparts, const int *__restrict__ rowlist, int nrows, int n, int nbeads,
                                  const double *__restrict__ alpha, const double *__restrict__ efs, double *__restrict__ efi,
                                  double *__restrict__ new_mu, double *__restrict__ efic, double *__restrict__ out_acc) {
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= nrows * nbeads) return;
	const int bead = t / nrows, ri = t - bead * nrows;
	const int i = rowlist ? rowlist[ri] : ri;
	double a[3] = {0, 0, 0};
	for (int q = 0; q < nparts; q++) {
		const double *s = part + (((size_t)q * nbeads + bead) * nrows + ri) * 3;
		a[0] += s[0]; a[1] += s[1]; a[2] += s[2];
	}
	const size_t o = ((size_t)bead * n + i) * 3;
	if (MODE == SWEEP_PALMO || MODE == SWEEP_PALMO_NONPOLAR) {
		for (int k = 0; k < 3; k++) efic[o + k] = -efi[o + k] - a[k];
	} else if (MODE == SWEEP_ACC) {
		for (int k = 0; k < 3; k++) out_acc[o + k] = a[k];
	} else {
		const double ai = alpha[i];
		for (int k = 0; k < 3; k++) { efi[o + k] = -a[k]; new_mu[o + k] = ai * (efs[o + k] - a[k]); }
	}
}

} // namespace mpmc
