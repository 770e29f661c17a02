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

// ---------------------------------------------------------------------------------------------------------
// Static field, real-space part: real_term() (src/System.Energy.cpp:2900-2940) when EWALD, thole_field_nopbc() (:3300-3333)
// otherwise, in the same (row block x column part) decomposition.  Rows and columns come from lists; the host launches
// (mobile rows x all charged columns) and (frozen rows x mobile charged columns), which is every pair except the frozen-frozen
// ones the reference skips (:2915 / :3311).  The cutoff test `r > rc` acts on the reference's own rounding of r, so it is done as
// in the pair sweep: a threshold on r^2 (PairParams::t2_es / t2_lj) with an exact recomputation inside a 1e-9 band; the image is
// always the reference's (min_image_fast).  With EWALD the radial factor (2a/sqrt(pi) e^{-a^2 r^2} r +/- erf[c](a r))/r^3 comes
// from a two-function table in r^2 (normal form, excluded form: a lane reads only the one it needs).
// ---------------------------------------------------------------------------------------------------------
struct FieldParams {
	double t2_in;       // largest exact r^2 that passes the cutoff test
	double t2_adm, t2_safe;
	double u_tab_lo;    // EWALD: the table covers [u_tab_lo, > t2_adm)
	int tab_base, tab_rows, tab_len;
};
constexpr int kFieldRow = 2 * (kTabDeg + 1) + kTabPad;

template <bool ORTHO, bool EWALD>
__global__ void __launch_bounds__(kOrdThreads)
k_field_parts(const double4 *__restrict__ posq, const int *__restrict__ meta, const int *__restrict__ collist, int ncols, int part_len,
              const int *__restrict__ rowlist, int nrows, int stride, CellDev c, FieldParams fp, const double *__restrict__ tab,
              double *__restrict__ part) {
	extern __shared__ __align__(16) double s_tabrows[];
	__shared__ double4 s_pq[kCtTile];
	__shared__ int     s_meta[kCtTile];
	const int bead = blockIdx.z, q = blockIdx.y;
	const double4 *pq = posq + (size_t)bead * stride;
	const int tid = threadIdx.x, jl = tid % kOrdJ, il = tid / kOrdJ;
	const int ri = blockIdx.x * kOrdI + il;
	const int i = ri < nrows ? rowlist[ri] : -1;
	double4 pi = make_double4(0, 0, 0, 0);
	int mi = 0;
	if (i >= 0) { pi = pq[i]; mi = meta_mol(meta[i]); }
	if (EWALD) stage_table(s_tabrows, tab, fp.tab_len);
	const double a = c.polar_alpha;
	double ex = 0, ey = 0, ez = 0;
	const int jbeg = q * part_len, jend = min(ncols, jbeg + part_len);
	for (int j0 = jbeg; j0 < jend; j0 += kCtTile) {
		__syncthreads();
		if (j0 + tid < jend) { const int j = collist[j0 + tid]; s_pq[tid] = pq[j]; s_meta[tid] = meta_mol(meta[j]); }
		if (EWALD && j0 == jbeg) stage_table_wait();
		__syncthreads();
		const int jn = min(kCtTile, jend - j0);
		if (i >= 0) {
#pragma unroll 2
			for (int jj = jl; jj < jn; jj += kOrdJ) {
				const double4 pj = s_pq[jj];
				const bool same = mi == s_meta[jj];
				double dx, dy, dz;
				min_image_fast<ORTHO>(c, __dsub_rn(pi.x, pj.x), __dsub_rn(pi.y, pj.y), __dsub_rn(pi.z, pj.z), dx, dy, dz);
				const double u = fma(dz, dz, fma(dy, dy, dx * dx));
				bool in = u <= fp.t2_adm && u > 0.0 && (EWALD || !same);               // :2917 / :3313, :3319; the i == j column has u = 0
				if (in && u > fp.t2_safe) {
					// within 1e-9 of the cutoff: decide on the reference's own rounding of r^2 (System.cpp:1228-1255)
					double qx, qy, qz;
					min_image<ORTHO>(c, __dsub_rn(pi.x, pj.x), __dsub_rn(pi.y, pj.y), __dsub_rn(pi.z, pj.z), qx, qy, qz);
					in = norm2_nofma(qx, qy, qz) <= fp.t2_in;
				}
				double f;
				if (EWALD) {
					const bool excl = same || pi.w == 0.0;                               // es_excluded form of the factor, :2921
					const int hi = __double2hiint(u);
					const double mid = __hiloint2double((hi & ~((1 << kTabShiftCoarse) - 1)) | (1 << (kTabShiftCoarse - 1)), 0);
					const double d = u - mid;
					const int row = min(max((hi >> kTabShiftCoarse) - fp.tab_base, 0), fp.tab_rows - 1);
					const double2 *r = reinterpret_cast<const double2 *>(s_tabrows + row * kFieldRow + (excl ? kTabDeg + 1 : 0));
					const double2 c67 = r[3], c45 = r[2], c23 = r[1], c01 = r[0];
					f = fma(c67.y, d, c67.x);
					f = fma(f, d, c45.y); f = fma(f, d, c45.x);
					f = fma(f, d, c23.y); f = fma(f, d, c23.x);
					f = fma(f, d, c01.y); f = fma(f, d, c01.x);
					if (in && u < fp.u_tab_lo) {                                         // closer than the table starts
						const double rr = sqrt(u), g = 2.0 * a * kOneOverSqrtPi * exp(-a * a * u) * rr;
						f = excl ? (g - erf(a * rr)) / (rr * u) : (g + erfc(a * rr)) / (u * rr);
					}
				} else {
					const double ir = rsqrt(u);
					f = ir * ir * ir;
				}
				const double fq = in ? f * pj.w : 0.0;
				ex = fma(fq, dx, ex); ey = fma(fq, dy, ey); ez = fma(fq, dz, ez);
			}
		}
	}
	ex = jlane_sum(ex); ey = jlane_sum(ey); ez = jlane_sum(ez);
	if (jl == 0 && i >= 0) {
		double *o = part + (((size_t)q * gridDim.z + bead) * nrows + ri) * 3;
		o[0] = ex; o[1] = ey; o[2] = ez;
	}
}

// ef[i] += sum of the column parts (fixed order), rows from a list
__global__ void k_field_finish(const double *__restrict__ part, int nparts, const int *__restrict__ rowlist, int nrows, int n, int nbeads,
                               double *__restrict__ ef) {
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= nrows * nbeads) return;
	const int bead = t / nrows, ri = t - bead * nrows;
	double a[3] = {0, 0, 0};
	for (int q = 0; q < nparts; q++) {
		const double *s = part + (((size_t)q * nbeads + bead) * nrows + ri) * 3;
		a[0] += s[0]; a[1] += s[1]; a[2] += s[2];
	}
	double *e = ef + ((size_t)bead * n + rowlist[ri]) * 3;
	e[0] += a[0]; e[1] += a[1]; e[2] += a[2];
}

// ---------------------------------------------------------------------------------------------------------
// Gauss-Seidel rank metric (src/System.cpp:1000-1029): rmin = smallest minimum-image separation between two polarizable sites,
// then per polarizable site the number of polarizable partners whose UN-imaged separation is <= 1.5 rmin.  Both are exact
// integer / ordering statements about the reference's own roundings, so the geometry here is the reference's (no FMA); sqrt is
// monotone and correctly rounded, so the minimum is taken over r^2 and the count compares r^2 with the largest r^2 whose sqrt
// is <= 1.5 rmin (k_rank_lim).  Rows and columns come from lists: the frozen-frozen part (most of a framework system) depends
// only on the frozen coordinates and on rmin, so the engine computes it once and reuses it while both stay the same.
// ---------------------------------------------------------------------------------------------------------
template <bool ORTHO>
__global__ void __launch_bounds__(kOrdThreads)
k_rank_min_parts(const double4 *__restrict__ posq, const int *__restrict__ collist, int ncols, int part_len, const int *__restrict__ rowlist,
                 int nrows, int stride, CellDev c, unsigned long long *__restrict__ r2min_bits) {
	__shared__ double4 s_pq[kCtTile];
	__shared__ int     s_idx[kCtTile];
	const int bead = blockIdx.z, q = blockIdx.y;
	const double4 *pq = posq + (size_t)bead * stride;
	const int tid = threadIdx.x, jl = tid % kOrdJ, il = tid / kOrdJ;
	const int ri = blockIdx.x * kOrdI + il;
	const int i = ri < nrows ? rowlist[ri] : -1;
	const double4 pi = i >= 0 ? pq[i] : make_double4(0, 0, 0, 0);
	double best = kMaxValue;
	const int jbeg = q * part_len, jend = min(ncols, jbeg + part_len);
	for (int j0 = jbeg; j0 < jend; j0 += kCtTile) {
		__syncthreads();
		if (j0 + tid < jend) { const int j = collist[j0 + tid]; s_pq[tid] = pq[j]; s_idx[tid] = j; }
		__syncthreads();
		const int jn = min(kCtTile, jend - j0);
		if (i >= 0)
#pragma unroll 2
			for (int jj = jl; jj < jn; jj += kOrdJ) {
				const double4 pj = s_pq[jj];
				double dx, dy, dz;
				min_image<ORTHO>(c, __dsub_rn(pi.x, pj.x), __dsub_rn(pi.y, pj.y), __dsub_rn(pi.z, pj.z), dx, dy, dz);
				const double r2 = norm2_nofma(dx, dy, dz);
				if (s_idx[jj] != i) best = fmin(best, r2);
			}
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) best = fmin(best, __shfl_xor_sync(0xffffffffu, best, o));
	if ((tid & 31) == 0) atomicMin(r2min_bits + bead, (unsigned long long)__double_as_longlong(best));   // positive doubles order like their bits
}

// per bead: rmin = sqrt(min r^2); t2 = the largest r^2 with sqrt(r^2) <= 1.5 rmin; recount = (t2 differs from the value the cached
// frozen-frozen counts were made with); the cache key is updated
__global__ void k_rank_lim(const unsigned long long *__restrict__ r2min_bits, const unsigned long long *__restrict__ r2min_ff_bits, int nbeads,
                           double *__restrict__ t2, double *__restrict__ t2_cached, int *__restrict__ recount) {
	const int b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= nbeads) return;
	const unsigned long long m = min(r2min_bits[b], r2min_ff_bits[b]);
	const double rmin = sqrt(__longlong_as_double((long long)m));
	const double lim = __dmul_rn(rmin, 1.5);
	// bisection on the bit pattern: pred(x) = sqrt(x) <= lim is true up to some point and false beyond
	unsigned long long lo = 0, hi = (unsigned long long)__double_as_longlong(4.0 * lim * lim + 1.0);
	if (!(lim < 1.0e150)) { lo = hi = 0x7fefffffffffffffull; }          // no polarizable pair at all: everything counts (as in the reference)
	while (hi - lo > 1) {
		const unsigned long long mid = lo + (hi - lo) / 2;
		if (sqrt(__longlong_as_double((long long)mid)) <= lim) lo = mid; else hi = mid;
	}
	const double v = __longlong_as_double((long long)lo);
	t2[b] = v;
	recount[b] = (v != t2_cached[b]);
	t2_cached[b] = v;
}

// out[row site] += number of columns j != i with |r_i - r_j|^2 (un-imaged, the reference's rounding) <= t2.  Integer counts added
// as doubles: exact, so the order of the atomic adds does not matter.  `gate`: when non-null, the whole launch is skipped for a
// bead unless gate[bead] != 0 (cached frozen-frozen counts).
__global__ void __launch_bounds__(kOrdThreads)
k_rank_count_parts(const double4 *__restrict__ posq, const int *__restrict__ collist, int ncols, int part_len, const int *__restrict__ rowlist,
                   int nrows, int n, int stride, const double *__restrict__ t2, const int *__restrict__ gate, double *__restrict__ out) {
	__shared__ double4 s_pq[kCtTile];
	__shared__ int     s_idx[kCtTile];
	const int bead = blockIdx.z, q = blockIdx.y;
	if (gate && !gate[bead]) return;
	const double4 *pq = posq + (size_t)bead * stride;
	const int tid = threadIdx.x, jl = tid % kOrdJ, il = tid / kOrdJ;
	const int ri = blockIdx.x * kOrdI + il;
	const int i = ri < nrows ? rowlist[ri] : -1;
	const double4 pi = i >= 0 ? pq[i] : make_double4(0, 0, 0, 0);
	const double lim2 = t2[bead];
	int cnt = 0;
	const int jbeg = q * part_len, jend = min(ncols, jbeg + part_len);
	for (int j0 = jbeg; j0 < jend; j0 += kCtTile) {
		__syncthreads();
		if (j0 + tid < jend) { const int j = collist[j0 + tid]; s_pq[tid] = pq[j]; s_idx[tid] = j; }
		__syncthreads();
		const int jn = min(kCtTile, jend - j0);
		if (i >= 0)
#pragma unroll 4
			for (int jj = jl; jj < jn; jj += kOrdJ) {
				const double4 pj = s_pq[jj];
				// |d| is symmetric in (i,j) bit for bit, so counting over ordered pairs equals the reference's i<j double update
				const double r2 = norm2_nofma(__dsub_rn(pi.x, pj.x), __dsub_rn(pi.y, pj.y), __dsub_rn(pi.z, pj.z));
				cnt += (r2 <= lim2 && s_idx[jj] != i) ? 1 : 0;
			}
	}
#pragma unroll
	for (int o = kOrdJ / 2; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
	if (jl == 0 && i >= 0 && cnt) atomicAdd(out + (size_t)bead * n + i, (double)cnt);
}

// rank[i] = cached frozen-frozen count for frozen polarizable sites, 0 elsewhere (the mobile parts are added on top);
// when the cache is being rebuilt (recount) the cached array is cleared instead and filled by the gated count launch
__global__ void k_rank_init(const double *__restrict__ cnt_ff, int n, int nbeads, double *__restrict__ rank) {
	const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t < (size_t)n * nbeads) rank[t] = cnt_ff[t];
}
__global__ void k_rank_clear_gated(const int *__restrict__ gate, int n, int nbeads, double *__restrict__ cnt_ff) {
	const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t < (size_t)n * nbeads && gate[t / n]) cnt_ff[t] = 0.0;
}

} // namespace mpmc
