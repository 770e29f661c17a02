// kernels_pair2.cuh — the pair-energy sweep, second generation: lj() + coulombic_real() (reference src/System.Energy.cpp:897-1032,
// 1466-1517) over every pair i<j that pairs() visits (src/System.cpp:967-991).
//
// What bounds this sweep on B200 is the FP64 pipe (64 lanes/SM: one warp instruction every 2 cycles per scheduler, during which the
// scheduler issues nothing else: cost of a pair ~ 2 x FP64 instructions + other instructions, tools/ubench/fp64_ubench.cu and the
// instruction counts of the ncu source page).  So the kernel is built to issue as few instructions per pair as the reference's
// semantics allow:
//   * geometry.  The pair energies need r^2 only.  In an orthorhombic cell the sites are folded into the cell centred on the origin
//     while they are staged (once per site and chunk, not per pair), after which |d_img| = L/2 - | |dx| - L/2 | per axis: 3 adds per
//     axis with the absolute values as operand modifiers, 12 FP64 instructions per pair with the sum of squares (the rint form needs
//     15).  A triclinic cell keeps the general FMA form.
//   * the reference's cutoff tests act on the last bit of rimg = sqrt(r^2) computed WITHOUT FMA; sqrt and the subtraction of 1e-12 are
//     monotone, so each test is equivalent to r2_exact <= T for a double T the host finds by bisection (prepare_pair_sweep()).  The
//     fast r^2 differs from the exact one by < 1e-13 relative, and the tests are done on the HIGH WORD of r^2 with integer compares
//     (2^-20 relative granularity, no FP64 instruction): below the word of T (1 - 1e-9) the pair is inside both cutoffs, above the
//     word of T (1 + 1e-9) it is outside, and the few pairs in between (never, except on lattices) are recomputed from the unfolded
//     coordinates the reference's way and tested exactly; every other decision is provably the reference's.
//   * padding lanes and columns carry NaN coordinates: their r^2 is NaN, whose high word is above every threshold — no extra test.
//   * erfc(alpha r)/r comes from a table in r^2 (radial_table.h): 1 add + 7 FMA, no sqrt, no exp; LJ needs only 1/r^2.
//   * the pair sum does not depend on the order of the sites, so the sweep runs over a copy of the site table sorted by CLASS
//     (frozen?, charged?, LJ-active?).  A block of pairs between two classes needs LJ only, Coulomb only, both or nothing at all
//     — frozen x frozen and (charge-only) x (LJ-only) blocks are simply not in the work list — and the inner loop is compiled once
//     per kind, so a five-site H2 system (one LJ+charge site, two charge-only, two LJ-only) does half the arithmetic it would
//     with every pair going through both formulas.
//   * work is dealt to warps, not CTAs: the (i-group of 32 sites) x (j site) columns of every class block are flattened and cut
//     into items by COST (columns weighted by what their kind issues); a warp takes its first item by its index and the following,
//     smaller and smaller ones from a counter, so 148 SMs x 16 warps stay busy to the end whatever N and the mix of kinds are.
// Sums are accumulated per item in a fixed order and reduced by k_reduce_partials: results are bit-reproducible whichever warp
// ends up with an item.
#pragma once
#include "device_math.cuh"
#include "kernels_pair.cuh"
#include "radial_table.h"

namespace mpmc {

// CTA shapes: 2 CTAs x 8 warps at <= 128 registers for both kernels.  Measured alternatives on config 3 (tools/pair_timeline.py,
// tools/pair3_probe.py): 3 CTAs x 6 warps at 96 registers 28.6 us instead of 26.2; one CTA of 16 warps 24.9 us against 24.7 — the
// schedulers serve the CTA that arrived first on the SM before the later one (its warps end after 11 us, the other's after 17.5 us;
// with one CTA all warps end after ~16.5 us), but two warps per scheduler already reach ~3/4 of the rate of four, so the even
// finish buys nothing, and the Coulomb kernels were 3 % slower with it.
__host__ __device__ constexpr int pair_warps(bool) { return 8; }
__host__ __device__ constexpr int pair_ctas(bool) { return 2; }
constexpr int kErfRow = (kTabDeg + 1) + kTabPad;        // doubles per row of the erfc table

// per-site word of the pair sweep: the molecule index (same molecule = rd_excluded / es_excluded); padding lanes and columns carry
// kPmPad, which equals no molecule and marks the site as absent
constexpr int kPmPad = -1;
enum { kPairLJ = 1, kPairES = 2 };                      // what a block of pairs needs

// i-group = sorted sites [i_begin, min(i_begin + 32, i_end)) meets sorted sites [j_begin, j_end), which need `kind`; col0 = columns before it
struct PairSeg { int i_begin, i_end, j_begin, j_end, col0, kind; };
// one item of the work list: columns [col, col_end) of the flattened blocks, starting in segment `seg`, whose record rides along so
// that a warp knows which sites to fetch after ONE round trip to L2 instead of three dependent ones (a single-round sweep is one
// item per warp: its prologue is not hidden by anything)
struct __align__(16) PairItem { int col, col_end, seg, pad0; PairSeg first; int pad1, pad2; };
static_assert(sizeof(PairItem) == 48, "PairItem is read as three 16-byte words");

struct PairParams {
	double t2_lj;       // largest r^2 with  sqrt(r^2) - 1e-12 < cutoff      (lj, :934)
	double t2_es;       // largest r^2 with  !(sqrt(r^2) > cutoff)            (coulombic_real, :1490)
	double t2_adm;      // admission threshold on the FMA r^2: t2_lj (1 + 1e-9)
	double t2_safe;     // below this both tests hold without looking closer: t2_es (1 - 1e-9)
	double u_tab_lo;    // the table covers [u_tab_lo, > t2_adm)
	unsigned h_adm, h_safe, h_tab_lo;   // high words of t2_adm, t2_safe, u_tab_lo (u_tab_lo sits on a word boundary)
	int tab_base, tab_rows;
	int ncols;          // columns per bead system
	int items_per_bead; // items of one bead system
	int nseg;
	int single_round;   // every warp owns exactly one item: no counter traffic
};

// relative cost of one column (32 pairs) of each kind, ~ 2 x FP64 + other instructions of its inner loop: equal-cost items
inline int pair_kind_weight(int kind) { return kind == kPairLJ ? 60 : kind == kPairES ? 90 : 100; }   // the table look-ups of the Coulomb kinds are bound by shared-memory bandwidth

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

__device__ __forceinline__ long long pair_gtime() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ long long pair_gtime_after(double a, double b) { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t) : "d"(a), "d"(b)); return t; }
__device__ __forceinline__ unsigned pair_smid() { unsigned r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }

struct PairAcc { double rd, re, in; int cnt; double rd1, re1; };

// acc += v (and cnt += 1) under a predicate, as predicated instructions: the compiler's own if-conversion turns `if (p) acc += v`
// into two selects and an unconditional add, three issue slots more per pair
__device__ __forceinline__ void add_if(double &acc, int &cnt, double v, bool p) {
	asm("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %3, 0;\n\t@q add.f64 %0, %0, %2;\n\t@q add.s32 %1, %1, 1;\n\t}" : "+d"(acc), "+r"(cnt) : "d"(v), "r"((int)p));
}
__device__ __forceinline__ void add_if(double &acc, double v, bool p) {
	asm("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %2, 0;\n\t@q add.f64 %0, %0, %1;\n\t}" : "+d"(acc) : "d"(v), "r"((int)p));
}

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

// a move of `count` consecutive sites in every bead system: stage[b][k] -> posq[bead_lo + b][first + k], and — where the pair sweep
// reads a class-sorted copy — the same site in spq (iperm = site -> sorted position), so the copy never has to be re-gathered
__global__ void k_scatter_sites(const double4 *__restrict__ stage, double4 *__restrict__ posq, int stride, int bead_lo, int nb, int first, int count,
                                const int *__restrict__ iperm, double4 *__restrict__ spq) {
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= nb * count) return;
	const int b = t / count, k = t - b * count;
	const double4 v = stage[t];
	posq[(size_t)(bead_lo + b) * stride + first + k] = v;
	if (spq) spq[(size_t)(bead_lo + b) * stride + iperm[first + k]] = v;
}

// folds a site into the cell centred on the origin along each axis of an orthorhombic cell (coordinates end up in [-L/2, L/2] up to
// rounding; a site already there is not touched, so a cluster around the origin of a huge box keeps its full precision); the fold
// moves a site by whole lattice vectors, which changes no image distance by more than the rounding of the coordinate itself
template <bool ORTHO>
__device__ __forceinline__ double4 fold_site(const CellDev &c, double4 p) {
	if (ORTHO) {
		p.x = fma(-c.b[0][0], rint(p.x * c.rb[0][0]), p.x);
		p.y = fma(-c.b[1][1], rint(p.y * c.rb[1][1]), p.y);
		p.z = fma(-c.b[2][2], rint(p.z * c.rb[2][2]), p.z);
	}
	return p;
}

// minimum-image r^2 of two FOLDED sites: per axis |d_img| = h - | |d| - h |, h = L/2 (orthorhombic); general form otherwise
template <bool ORTHO>
__device__ __forceinline__ double r2_pair(const CellDev &c, double hx, double hy, double hz, double dx, double dy, double dz) {
	if (ORTHO) {
		const double ux = hx - fabs(fabs(dx) - hx), uy = hy - fabs(fabs(dy) - hy), uz = hz - fabs(fabs(dz) - hz);
		return fma(uz, uz, fma(uy, uy, ux * ux));
	}
	return r2_fast<false>(c, dx, dy, dz);
}

__device__ __forceinline__ double4 nan_site() { const double q = __longlong_as_double(0x7ff8000000000000LL); return make_double4(q, q, q, 0.0); }

// one staged chunk (cnt <= 32 columns in shared memory, folded) against the warp's 32 i sites, for a block of pairs of one kind.
// a.rd and a.re collect sum_j sqrt(eps_j) (..) and sum_j q_j (..): the factor of site i is applied by the caller when the segment ends.
// raw = the bead's unfolded sorted table, i / jc = sorted index of the lane's row / of the chunk's first column (rare exact paths).
template <bool ORTHO, int KIND, bool DIAG>
__device__ __forceinline__ void pair_chunk(const CellDev &c, const PairParams &pp, double hx, double hy, double hz, const double *s_tab, const double4 *s_pq,
                                           const double2 *s_lj, const int *s_pm, int cnt, int dlim, const double4 pi, const double2 li, int mi,
                                           const double4 *__restrict__ raw, int i, int jc, PairAcc &a) {
	constexpr bool LJ = (KIND & kPairLJ) != 0, ES = (KIND & kPairES) != 0;
	const unsigned h_adm = pp.h_adm, h_safe = pp.h_safe;
	for (int jj = 0; jj < cnt; jj += kPwCols) {
		double r2[kPwCols], elj[kPwCols], ees[kPwCols];
		bool in[kPwCols], excl[kPwCols], band[kPwCols], nearu[kPwCols];
		const int dl = DIAG ? dlim - jj : -1;                                // diagonal chunk: column u counts for this lane when dl < u (i < j)
#pragma unroll
		for (int u = 0; u < kPwCols; u++) {
			const double4 pj = s_pq[jj + u];
			const int mj = s_pm[jj + u];
			r2[u] = r2_pair<ORTHO>(c, hx, hy, hz, pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
			const unsigned hw = (unsigned)__double2hiint(r2[u]);               // NaN (padding) compares above every threshold
			excl[u] = mi == mj;                                              // same molecule (rd_excluded / es_excluded)
			in[u] = !excl[u] && (!DIAG || dl < u) && hw <= h_adm;             // i < j, inside the (widened) cutoff sphere
			band[u] = in[u] && hw >= h_safe;
			if (LJ) {
				// Lorentz-Berthelot LJ (:965-993); 4 eps_ij is applied after the sum
				const double2 ljj = s_lj[jj + u];
				const double sig = li.y + ljj.y;
				const double s2 = sig * sig * rcp_full(r2[u]), s6 = s2 * s2 * s2;
				elj[u] = ljj.x * fma(s6, s6, -s6);                                 // x sqrt(eps_i) once per segment, by the caller
			}
			if (ES) {                                                            // :1493-1497
				ees[u] = pj.w * tab_eval1(s_tab, pp.tab_base, pp.tab_rows, r2[u]);   // x q_i once per segment, by the caller
				nearu[u] = in[u] && hw < pp.h_tab_lo;                              // closer than the table starts
			}
		}
		if (ES && (nearu[0] || nearu[1] || nearu[2] || nearu[3])) {
#pragma unroll
			for (int u = 0; u < kPwCols; u++)
				if (nearu[u]) { const double r = sqrt(r2[u]); ees[u] = s_pq[jj + u].w * erfc(c.ewald_alpha * r) / r; }
		}
		if (!(band[0] || band[1] || band[2] || band[3])) {
			// even and odd columns into two accumulators each: the serial chain of dependent adds per pass is halved
#pragma unroll
			for (int u = 0; u < kPwCols; u++) {
				if (LJ) add_if((u & 1) ? a.rd1 : a.rd, a.cnt, elj[u], in[u]);
				if (ES) add_if((u & 1) ? a.re1 : a.re, ees[u], in[u]);
			}
		} else {
			// next to a cutoff: decide on the reference's own rounding of r^2 (System.cpp:1228-1255), from the unfolded coordinates.
			// The tests are redone from r^2 here (behind an opaque copy) so that the common path keeps no flag alive across the branch.
			const double4 qi = raw[max(i, 0)];
#pragma unroll
			for (int u = 0; u < kPwCols; u++) {
				unsigned hw = (unsigned)__double2hiint(r2[u]);
				asm volatile("" : "+r"(hw));
				bool in_lj = mi != s_pm[jj + u] && (!DIAG || dl < u) && hw <= h_adm, in_es = in_lj;
				if (in_lj && hw >= h_safe) {
					const double4 qj = raw[jc + jj + u];
					double ex, ey, ez;
					min_image<ORTHO>(c, __dsub_rn(qi.x, qj.x), __dsub_rn(qi.y, qj.y), __dsub_rn(qi.z, qj.z), ex, ey, ez);
					const double r2x = norm2_nofma(ex, ey, ez);
					in_lj = r2x <= pp.t2_lj;
					in_es = r2x <= pp.t2_es;
				}
				if (LJ && in_lj) { a.rd += elj[u]; a.cnt++; }
				if (ES && in_es) a.re += ees[u];
			}
		}
		if (ES && (excl[0] || excl[1] || excl[2] || excl[3])) {
			// es_self_intra = q_i q_j erf(alpha r)/r on the UN-imaged distance (:1503-1504): unfolded coordinates
#pragma unroll
			for (int u = 0; u < kPwCols; u++)
				if (excl[u] && mi >= 0 && dl < u) {
					const double4 qi = raw[i], qj = raw[jc + jj + u];
					const double dx = qi.x - qj.x, dy = qi.y - qj.y, dz = qi.z - qj.z;
					const double r = sqrt(fma(dz, dz, fma(dy, dy, dx * dx)));
					a.in += qi.w * qj.w * erf(c.ewald_alpha * r) / r;
				}
		}
	}
}

// item_list[k], k < items_per_bead: the work list; ctr: the next item to hand out (gridDim.x * warps per CTA when the kernel starts —
// k_reduce_partials puts it back)
template <bool ORTHO, bool ES>
__global__ void __launch_bounds__(pair_warps(ES) * 32, pair_ctas(ES))
k_pair_sweep(const double4 *__restrict__ spq, const double2 *__restrict__ lj, const int *__restrict__ pmeta, int stride, int nbeads,
             const PairSeg *__restrict__ seg, const PairItem *__restrict__ item_list, const PairParams pp, const CellDev c,
             const double *__restrict__ tab, PairPartial *__restrict__ partials, int *__restrict__ ctr, long long *__restrict__ prof) {
	extern __shared__ __align__(16) double s_raw[];
	double *s_tab = s_raw;
	const int tab_len = ES ? pp.tab_rows * kErfRow : 0;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	double *wbase = s_raw + ((tab_len + 1) & ~1) + warp * kPwWarpDoubles;
	double4 *s_pq = reinterpret_cast<double4 *>(wbase);
	double2 *s_lj = reinterpret_cast<double2 *>(s_pq + kPwJ);
	int *s_pm = reinterpret_cast<int *>(s_lj + kPwJ);
	if (ES) {
		// the table (~32 KB) in 16-byte asynchronous copies, all in flight at once: a plain copy loop is 16 dependent round trips to L2,
		// ~5 us during which the CTA does nothing else
		stage_table(s_tab, tab, tab_len);
	}
	const int gw = blockIdx.x * pair_warps(ES) + warp;
	const int items = nbeads * pp.items_per_bead;
	// developer timeline (mpmc_debug_pair_profile): per warp { entry, first sites in registers, last item summed, items done | SM }
	if (prof && lane == 0) { prof[4 * gw] = pair_gtime(); prof[4 * gw + 1] = 0; }
	const double hx = 0.5 * fabs(c.b[0][0]), hy = 0.5 * fabs(c.b[1][1]), hz = 0.5 * fabs(c.b[2][2]);
	if (lane < kPwCols) { s_pq[32 + lane] = nan_site(); s_lj[32 + lane] = make_double2(0, 0); s_pm[32 + lane] = kPmPad; }
	if (ES) { stage_table_wait(); __syncthreads(); }
	__syncwarp();

	for (int it = gw; it < items;) {
		const int k = it / nbeads, bead = it - k * nbeads;      // k-major: the rounds of prepare_pair_sweep() follow each other
		const double4 *pq = spq + (size_t)bead * stride;
		const int4 w0 = reinterpret_cast<const int4 *>(item_list + k)[0], w1 = reinterpret_cast<const int4 *>(item_list + k)[1],
		           w2 = reinterpret_cast<const int4 *>(item_list + k)[2];
		int col = w0.x;
		const int col_end = w0.y;
		PairAcc a = {0.0, 0.0, 0.0, 0, 0.0, 0.0};
		if (col < col_end) {
			int s = w0.z;                                    // segment that holds `col`
			PairSeg sg = {w1.x, w1.y, w1.z, w1.w, w2.x, w2.y};
			while (col < col_end) {
				const int j0 = sg.j_begin + (col - sg.col0);
				const int j1 = min(sg.j_end, j0 + (col_end - col));
				int i = sg.i_begin + lane;
				double4 pi = nan_site();
				double2 li = make_double2(0, 0);
				int mi = kPmPad;
				if (i < sg.i_end) { pi = fold_site<ORTHO>(c, pq[i]); li = lj[i]; mi = pmeta[i]; } else i = -1;
				PairAcc sa = {0.0, 0.0, 0.0, 0, 0.0, 0.0};       // this segment's sums, without the factors of site i
				// first chunk of the segment into registers; later chunks are fetched while the previous one is being swept
				double4 npq = nan_site(); double2 nlj = make_double2(0, 0); int npm = kPmPad;
				if (j0 + lane < j1) { npq = pq[j0 + lane]; nlj = lj[j0 + lane]; npm = pmeta[j0 + lane]; }
				if (prof && lane == 0 && prof[4 * gw + 1] == 0) prof[4 * gw + 1] = pair_gtime_after(pi.x, npq.x);
				for (int jc = j0; jc < j1; jc += 32) {
					const int cnt = min(32, j1 - jc);
					__syncwarp();
					s_pq[lane] = fold_site<ORTHO>(c, npq); s_lj[lane] = nlj; s_pm[lane] = npm;
					__syncwarp();
					{
						const int jn = jc + 32 + lane;
						npm = kPmPad; npq = nan_site(); nlj = make_double2(0, 0);
						if (jn < j1) { npq = pq[jn]; nlj = lj[jn]; npm = pmeta[jn]; }
					}
					const int dlim = (jc < sg.i_begin + 32) ? sg.i_begin + lane - jc : -1;   // a chunk that overlaps the i-group (same class): keep i < j only
#define MPMC_PAIR_CHUNK(KIND, DIAG) pair_chunk<ORTHO, KIND, DIAG>(c, pp, hx, hy, hz, s_tab, s_pq, s_lj, s_pm, cnt, dlim, pi, li, mi, pq, i, jc, sa)
					const int kind = ES ? sg.kind : kPairLJ;
					if (jc < sg.i_begin + 32) {
						if (kind == (kPairLJ | kPairES)) MPMC_PAIR_CHUNK(kPairLJ | kPairES, true);
						else if (kind == kPairES) MPMC_PAIR_CHUNK(kPairES, true);
						else MPMC_PAIR_CHUNK(kPairLJ, true);
					} else {
						if (kind == (kPairLJ | kPairES)) MPMC_PAIR_CHUNK(kPairLJ | kPairES, false);
						else if (kind == kPairES) MPMC_PAIR_CHUNK(kPairES, false);
						else MPMC_PAIR_CHUNK(kPairLJ, false);
					}
#undef MPMC_PAIR_CHUNK
				}
				a.rd = fma(li.x, sa.rd + sa.rd1, a.rd); a.re = fma(pi.w, sa.re + sa.re1, a.re); a.in += sa.in; a.cnt += sa.cnt;
				col += j1 - j0;
				s++;
				if (col < col_end) sg = seg[s];
			}
		}
		// the four sums over the warp in one shrinking xor tree (6 shuffles instead of 20, fixed order): lanes 0, 8, 16, 24 end up
		// with the totals of rd, es_real, es_intra, n_in and store them
		{
			const unsigned F = 0xffffffffu;
			double v0 = a.rd, v1 = a.re, v2 = a.in, v3 = (double)a.cnt;
			const bool h16 = (lane & 16) != 0, h8 = (lane & 8) != 0;
			const double r0 = __shfl_xor_sync(F, h16 ? v0 : v2, 16), r1 = __shfl_xor_sync(F, h16 ? v1 : v3, 16);
			v0 = (h16 ? v2 : v0) + r0; v1 = (h16 ? v3 : v1) + r1;
			const double r2 = __shfl_xor_sync(F, h8 ? v0 : v1, 8);
			v0 = (h8 ? v1 : v0) + r2;
			v0 += __shfl_xor_sync(F, v0, 4); v0 += __shfl_xor_sync(F, v0, 2); v0 += __shfl_xor_sync(F, v0, 1);
			if ((lane & 7) == 0) reinterpret_cast<double *>(partials + ((size_t)bead * pp.items_per_bead + k))[lane >> 3] = lane == 0 ? 4.0 * v0 : v0;
		}
		if (prof && lane == 0) { prof[4 * gw + 2] = pair_gtime(); prof[4 * gw + 3] += 1 + ((long long)pair_smid() << 32) * (prof[4 * gw + 3] == 0); }
		if (pp.single_round) break;
		if (lane == 0) it = atomicAdd(ctr, 1);
		it = __shfl_sync(0xffffffffu, it, 0);
	}
}

inline int pair_ctas_per_sm(bool es) { return pair_ctas(es); }
inline size_t pair_sweep_smem(bool es, int tab_rows) {
	const size_t tab_len = es ? ((size_t)tab_rows * kErfRow + 1) & ~(size_t)1 : 0;
	return sizeof(double) * (tab_len + (size_t)pair_warps(es) * kPwWarpDoubles);
}

} // namespace mpmc
