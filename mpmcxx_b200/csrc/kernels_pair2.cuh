// kernels_pair2.cuh — the pair-energy sweep, second generation: lj() + coulombic_real() (reference src/System.Energy.cpp:897-1032,
// 1466-1517) over every pair i<j that pairs() visits (src/System.cpp:967-991).
//
// What bounds this sweep on B200 is the FP64 pipe (64 lanes/SM: one warp instruction every 2 cycles per scheduler) and, right behind
// it, the issue slots (tools/ubench/fp64_ubench.cu).  So the kernel is built to issue as few instructions per pair as the reference's
// semantics allow:
//   * geometry with FMA (16 FP64 instructions per pair in an orthorhombic cell).  The reference's cutoff tests act on the last bit of
//     rimg = sqrt(r^2) computed WITHOUT FMA; sqrt and the subtraction of 1e-12 are monotone, so each test is equivalent to
//     r2_exact <= T for a double T the host finds by bisection (pair_thresholds()).  The FMA r^2 differs from the exact one by < 1e-13
//     relative, so only pairs within 1e-9 of a threshold (never, except on lattices) are recomputed the reference's way and tested
//     exactly; every other decision is provably the reference's.
//   * the ~48 % of pairs outside the cutoff sphere cost nothing more: the lanes of a warp push the pairs that do interact into a
//     per-warp queue (ballot + popc, fixed order) and the expensive part runs on full warps, 32 queued pairs at a time.
//   * erfc(alpha r)/r comes from a table in r^2 (radial_table.h): 1 add + 7 FMA, no sqrt, no exp; LJ needs only 1/r^2.
//   * the pair sum does not depend on the order of the sites, so the sweep runs over a copy of the site table sorted by CLASS
//     (frozen?, charged?, LJ-active?).  A block of pairs between two classes needs LJ only, Coulomb only, both or nothing at all
//     — frozen x frozen and (charge-only) x (LJ-only) blocks are simply not in the work list — and the inner loop is compiled once
//     per kind, so a five-site H2 system (one LJ+charge site, two charge-only, two LJ-only) does half the arithmetic it would
//     with every pair going through both formulas.
//   * work is dealt to warps, not CTAs: the (i-group of 32 sites) x (j site) columns of every class block are flattened and cut
//     into equal ranges, so 148 SMs x 16 warps stay busy to the end whatever N is.
// Sums are accumulated per item in a fixed order and reduced by k_reduce_partials: results are bit-reproducible.
#pragma once
#include "device_math.cuh"
#include "kernels_pair.cuh"
#include "radial_table.h"

namespace mpmc {

constexpr int kPwWarps = 8, kPwThreads = kPwWarps * 32;
constexpr int kErfRow = (kTabDeg + 1) + kTabPad;        // doubles per row of the erfc table

// per-site word of the pair sweep: the molecule index (same molecule = rd_excluded / es_excluded); padding lanes and columns carry
// kPmPad, which equals no molecule and marks the site as absent
constexpr int kPmPad = -1;
enum { kPairLJ = 1, kPairES = 2 };                      // what a block of pairs needs

// i-group = sorted sites [i_begin, min(i_begin + 32, i_end)) meets sorted sites [j_begin, j_end), which need `kind`; col0 = columns before it
struct PairSeg { int i_begin, i_end, j_begin, j_end, col0, kind; };

struct PairParams {
	double t2_lj;       // largest r^2 with  sqrt(r^2) - 1e-12 < cutoff      (lj, :934)
	double t2_es;       // largest r^2 with  !(sqrt(r^2) > cutoff)            (coulombic_real, :1490)
	double t2_adm;      // admission threshold on the FMA r^2: t2_lj (1 + 1e-9)
	double t2_safe;     // below this both tests hold without looking closer: t2_es (1 - 1e-9)
	double u_tab_lo;    // the table covers [u_tab_lo, > t2_adm)
	int tab_base, tab_rows;
	int ncols;          // columns per bead system
	int items_per_bead, cols_per_item;
	int nseg;
};

// row lookup of a radial table held in shared memory: value of function 0 of a 1-function table at u
// (the row index is clamped to the table: callers discard the value when u lies outside [u_lo, u_hi))
__device__ __forceinline__ double tab_eval1(const double *rows, int base, int nrows, double u) {
	const int hi = __double2hiint(u);
	const double mid = __hiloint2double((hi & ~((1 << kTabShift) - 1)) | (1 << (kTabShift - 1)), 0);
	const double d = u - mid;
	const int row = min(max((hi >> kTabShift) - base, 0), nrows - 1);
	const double2 *r = reinterpret_cast<const double2 *>(rows + row * kErfRow);
	const double2 c67 = r[3], c45 = r[2], c23 = r[1], c01 = r[0];
	double v = fma(c67.y, d, c67.x);
	v = fma(v, d, c45.y); v = fma(v, d, c45.x);
	v = fma(v, d, c23.y); v = fma(v, d, c23.x);
	v = fma(v, d, c01.y); v = fma(v, d, c01.x);
	return v;
}

// minimum-image r^2 with FMA contraction (NOT the reference's rounding; see the header comment)
template <bool ORTHO>
__device__ __forceinline__ double r2_fast(const CellDev &c, double dx, double dy, double dz) {
	const double M = 6755399441055744.0;
	double ix, iy, iz;
	if (ORTHO) {
		const double fx = fma(c.rb[0][0], dx, M) - M, fy = fma(c.rb[1][1], dy, M) - M, fz = fma(c.rb[2][2], dz, M) - M;
		ix = fma(-c.b[0][0], fx, dx); iy = fma(-c.b[1][1], fy, dy); iz = fma(-c.b[2][2], fz, dz);
	} else {
		const double f0 = (fma(c.rb[2][0], dz, fma(c.rb[1][0], dy, c.rb[0][0] * dx)) + M) - M;
		const double f1 = (fma(c.rb[2][1], dz, fma(c.rb[1][1], dy, c.rb[0][1] * dx)) + M) - M;
		const double f2 = (fma(c.rb[2][2], dz, fma(c.rb[1][2], dy, c.rb[0][2] * dx)) + M) - M;
		ix = fma(-c.b[2][0], f2, fma(-c.b[1][0], f1, fma(-c.b[0][0], f0, dx)));
		iy = fma(-c.b[2][1], f2, fma(-c.b[1][1], f1, fma(-c.b[0][1], f0, dy)));
		iz = fma(-c.b[2][2], f2, fma(-c.b[1][2], f1, fma(-c.b[0][2], f0, dz)));
	}
	return fma(iz, iz, fma(iy, iy, ix * ix));
}

struct PairAcc { double rd, re, in; int cnt; };

#ifndef MPMC_PW_CTAS
#define MPMC_PW_CTAS 2
#endif
constexpr int kPwCtas = MPMC_PW_CTAS;         // resident CTAs per SM the kernel is compiled for
constexpr int kPwCols = 4;                      // columns per pass of the inner loop (independent FP64 chains)
constexpr int kPwJ = 32 + kPwCols;              // staged j sites per warp (+ padding columns)
constexpr int kPwWarpDoubles = kPwJ * 4 + kPwJ * 2 + kPwJ / 2;   // (x y z q), (sqrt eps, sigma/2), site word

// 1/x to double precision from the hardware seed (relative error < 2^-20): one cubically convergent step e = 1 - x y,
// y <- y (1 + e + e^2) leaves 2^-60.  3 FMAs; the seed instruction runs on the XU pipe, not the FP64 pipe.
__device__ __forceinline__ double rcp_full(double x) {
	double y;
	asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
	const double e = fma(-x, y, 1.0);
	return fma(y, fma(e, e, e), y);
}

// the class-sorted copy of the coordinates: spq[bead][k] = posq[bead][perm[k]]
__global__ void k_pair_gather(const double4 *__restrict__ posq, const int *__restrict__ perm, int n, int stride, int nbeads, double4 *__restrict__ spq) {
	const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= (size_t)n * nbeads) return;
	const int bead = (int)(t / n), k = (int)(t - (size_t)bead * n);
	spq[(size_t)bead * stride + k] = posq[(size_t)bead * stride + perm[k]];
}

// a move of `count` consecutive sites in every bead system: stage[b][k] -> posq[bead_lo + b][first + k]
__global__ void k_scatter_sites(const double4 *__restrict__ stage, double4 *__restrict__ posq, int stride, int bead_lo, int nb, int first, int count) {
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= nb * count) return;
	const int b = t / count, k = t - b * count;
	posq[(size_t)(bead_lo + b) * stride + first + k] = stage[t];
}

// one staged chunk (cnt <= 32 columns in shared memory) against the warp's 32 i sites, for a block of pairs of one kind
template <bool ORTHO, int KIND>
__device__ __forceinline__ void pair_chunk(const CellDev &c, const PairParams &pp, const double *s_tab, const double4 *s_pq, const double2 *s_lj,
                                           const int *s_pm, int cnt, int dlim, const double4 pi, const double2 li, int mi, PairAcc &a) {
	constexpr bool LJ = (KIND & kPairLJ) != 0, ES = (KIND & kPairES) != 0;
	const double t2_adm = pp.t2_adm, t2_safe = pp.t2_safe;
	for (int jj = 0; jj < cnt; jj += kPwCols) {
		double r2[kPwCols], elj[kPwCols], ees[kPwCols];
		bool in[kPwCols], excl[kPwCols], band[kPwCols];
#pragma unroll
		for (int u = 0; u < kPwCols; u++) {
			const double4 pj = s_pq[jj + u];
			const int mj = s_pm[jj + u];
			r2[u] = r2_fast<ORTHO>(c, pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
			excl[u] = mi == mj;                                              // same molecule (rd_excluded / es_excluded)
			in[u] = !excl[u] && (mi | mj) >= 0 && dlim < jj + u && r2[u] <= t2_adm;   // present, i < j, inside the (1e-9 widened) cutoff sphere
			band[u] = in[u] && r2[u] > t2_safe;
			if (LJ) {
				// Lorentz-Berthelot LJ (:965-993); 4 eps_ij is applied after the sum
				const double2 ljj = s_lj[jj + u];
				const double sig = li.y + ljj.y;
				const double s2 = sig * sig * rcp_full(r2[u]), s6 = s2 * s2 * s2;
				elj[u] = (li.x * ljj.x) * fma(s6, s6, -s6);
			}
			if (ES) {                                                            // :1493-1497
				ees[u] = 0.0;
				if (in[u]) {
					double f = tab_eval1(s_tab, pp.tab_base, pp.tab_rows, r2[u]);
					if (r2[u] < pp.u_tab_lo) { const double r = sqrt(r2[u]); f = erfc(c.ewald_alpha * r) / r; }   // closer than the table starts
					ees[u] = pi.w * pj.w * f;
				}
			}
		}
		bool in_es[kPwCols];
#pragma unroll
		for (int u = 0; u < kPwCols; u++) in_es[u] = in[u];
		if (band[0] || band[1] || band[2] || band[3]) {
			// within 1e-9 of a cutoff: decide on the reference's own rounding of r^2 (System.cpp:1228-1255)
#pragma unroll
			for (int u = 0; u < kPwCols; u++)
				if (band[u]) {
					const double4 pj = s_pq[jj + u];
					double ex, ey, ez;
					min_image<ORTHO>(c, __dsub_rn(pi.x, pj.x), __dsub_rn(pi.y, pj.y), __dsub_rn(pi.z, pj.z), ex, ey, ez);
					const double r2x = norm2_nofma(ex, ey, ez);
					in[u] = r2x <= pp.t2_lj;
					in_es[u] = r2x <= pp.t2_es;
				}
		}
#pragma unroll
		for (int u = 0; u < kPwCols; u++) {
			if (LJ && in[u]) { a.rd += elj[u]; a.cnt++; }
			if (ES && in_es[u]) a.re += ees[u];
		}
		if (ES && (excl[0] || excl[1] || excl[2] || excl[3])) {
			// es_self_intra = q_i q_j erf(alpha r)/r on the UN-imaged distance (:1503-1504)
#pragma unroll
			for (int u = 0; u < kPwCols; u++)
				if (excl[u] && mi >= 0 && dlim < jj + u) {
					const double4 pj = s_pq[jj + u];
					const double dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
					const double r = sqrt(fma(dz, dz, fma(dy, dy, dx * dx)));
					a.in += pi.w * pj.w * erf(c.ewald_alpha * r) / r;
				}
		}
	}
}

template <bool ORTHO, bool ES>
__global__ void __launch_bounds__(kPwThreads, kPwCtas)
k_pair_sweep(const double4 *__restrict__ spq, const double2 *__restrict__ lj, const int *__restrict__ pmeta, int stride, int nbeads,
             const PairSeg *__restrict__ seg, const int *__restrict__ item_seg, const PairParams pp, const CellDev c,
             const double *__restrict__ tab, PairPartial *__restrict__ partials) {
	extern __shared__ __align__(16) double s_raw[];
	double *s_tab = s_raw;
	const int tab_len = ES ? pp.tab_rows * kErfRow : 0;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	double *wbase = s_raw + ((tab_len + 1) & ~1) + warp * kPwWarpDoubles;
	double4 *s_pq = reinterpret_cast<double4 *>(wbase);
	double2 *s_lj = reinterpret_cast<double2 *>(s_pq + kPwJ);
	int *s_pm = reinterpret_cast<int *>(s_lj + kPwJ);
	if (ES) {
		for (int q = tid; q < tab_len; q += kPwThreads) s_tab[q] = tab[q];
		__syncthreads();
	}
	const int W = gridDim.x * kPwWarps, gw = blockIdx.x * kPwWarps + warp;
	const int items = nbeads * pp.items_per_bead;
	if (lane < kPwCols) { s_pq[32 + lane] = make_double4(0, 0, 0, 0); s_lj[32 + lane] = make_double2(0, 0); s_pm[32 + lane] = kPmPad; }
	__syncwarp();

	for (int it = gw; it < items; it += W) {
		const int bead = it / pp.items_per_bead, k = it - bead * pp.items_per_bead;
		const double4 *pq = spq + (size_t)bead * stride;
		int col = k * pp.cols_per_item;
		const int col_end = min(col + pp.cols_per_item, pp.ncols);
		PairAcc a = {0.0, 0.0, 0.0, 0};
		if (col < col_end) {
			int s = item_seg[k];                             // segment that holds `col`
			while (col < col_end) {
				const PairSeg sg = seg[s];
				const int j0 = sg.j_begin + (col - sg.col0);
				const int j1 = min(sg.j_end, j0 + (col_end - col));
				const int i = sg.i_begin + lane;
				double4 pi = make_double4(0, 0, 0, 0);
				double2 li = make_double2(0, 0);
				int mi = kPmPad;
				if (i < sg.i_end) { pi = pq[i]; li = lj[i]; mi = pmeta[i]; }
				// first chunk of the segment into registers; later chunks are fetched while the previous one is being swept
				double4 npq = make_double4(0, 0, 0, 0); double2 nlj = make_double2(0, 0); int npm = kPmPad;
				if (j0 + lane < j1) { npq = pq[j0 + lane]; nlj = lj[j0 + lane]; npm = pmeta[j0 + lane]; }
				for (int jc = j0; jc < j1; jc += 32) {
					const int cnt = min(32, j1 - jc);
					__syncwarp();
					s_pq[lane] = npq; s_lj[lane] = nlj; s_pm[lane] = npm;
					__syncwarp();
					{
						const int jn = jc + 32 + lane;
						npm = kPmPad; npq = make_double4(0, 0, 0, 0); nlj = make_double2(0, 0);
						if (jn < j1) { npq = pq[jn]; nlj = lj[jn]; npm = pmeta[jn]; }
					}
					const int dlim = (jc < sg.i_begin + 32) ? i - jc : -1;   // a chunk that overlaps the i-group (same class): keep i < j only
					if (ES) {
						if (sg.kind == (kPairLJ | kPairES)) pair_chunk<ORTHO, kPairLJ | kPairES>(c, pp, s_tab, s_pq, s_lj, s_pm, cnt, dlim, pi, li, mi, a);
						else if (sg.kind == kPairES) pair_chunk<ORTHO, kPairES>(c, pp, s_tab, s_pq, s_lj, s_pm, cnt, dlim, pi, li, mi, a);
						else pair_chunk<ORTHO, kPairLJ>(c, pp, s_tab, s_pq, s_lj, s_pm, cnt, dlim, pi, li, mi, a);
					} else pair_chunk<ORTHO, kPairLJ>(c, pp, s_tab, s_pq, s_lj, s_pm, cnt, dlim, pi, li, mi, a);
				}
				col += j1 - j0;
				s++;
			}
		}
		const double rd = warp_sum(a.rd), re = warp_sum(a.re), in = warp_sum(a.in), cn = warp_sum((double)a.cnt);
		if (lane == 0) { PairPartial p; p.rd = 4.0 * rd; p.es_real = re; p.es_intra = in; p.n_in = cn; partials[it] = p; }
	}
}

inline int pair_ctas_per_sm(bool) { return kPwCtas; }
inline size_t pair_sweep_smem(bool es, int tab_rows) {
	const size_t tab_len = es ? ((size_t)tab_rows * kErfRow + 1) & ~(size_t)1 : 0;
	return sizeof(double) * (tab_len + (size_t)kPwWarps * kPwWarpDoubles);
}

} // namespace mpmc
