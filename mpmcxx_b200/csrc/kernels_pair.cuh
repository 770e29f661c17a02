// kernels_pair.cuh — all-pairs sweeps: LJ + Ewald real-space energy (triangular), static-field real term and
// GS rank metric (ordered).  FP64 throughout; j-sites are staged through shared memory one tile at a time and
// every reduction is a fixed-shape tree, so results are bit-reproducible run to run.
#pragma once
#include "device_math.cuh"

namespace mpmc {

constexpr int kPairTile = 128;   // largest tile side of the triangular energy sweep (= threads per CTA); 64 and 32 are used for small systems

// site metadata word: molecule index in the low 31 bits, frozen flag in the sign bit
__device__ __forceinline__ int  meta_mol(int m)    { return m & 0x7fffffff; }
__device__ __forceinline__ bool meta_frozen(int m) { return m < 0; }

// per-CTA partial sums of the energy sweep
struct PairPartial { double rd, es_real, es_intra, n_in; };

// ---------------------------------------------------------------------------------------------------------
// K1/K2: lj() + coulombic_real()  (reference src/System.Energy.cpp:897-1032, 1466-1517) over the pairs i<j that
// pairs() would visit (src/System.cpp:967-991).  One CTA per (tile_i <= tile_j) entry of the host-built tile list
// (tiles whose two blocks are entirely frozen are not listed: pair->frozen pairs contribute nothing, :936, :1487).
// Thread t owns site i = tile_i*T + t in registers and walks the j tile in shared memory.  The tile side T (128, 64 or 32) is
// chosen by the host so that there are enough CTAs to fill and balance 148 SMs even for a few hundred sites per bead.
// ---------------------------------------------------------------------------------------------------------
template <bool ORTHO, bool ES, int T>
__global__ void __launch_bounds__(T)
k_pair_energy(const double4 *__restrict__ posq, const double2 *__restrict__ lj, const int *__restrict__ meta,
              int n, int stride, const int2 *__restrict__ tiles, int ntiles, CellDev c, PairPartial *__restrict__ partials) {
	__shared__ double4 s_pq[T];
	__shared__ double2 s_lj[T];
	__shared__ int     s_meta[T];
	__shared__ double  s_red[4][T / 32];

	const int bead = blockIdx.y;
	const int2 tile = tiles[blockIdx.x];
	const double4 *pq = posq + (size_t)bead * stride;
	const int tid = threadIdx.x;
	const int i = tile.x * T + tid;
	const int j0 = tile.y * T;

	double4 pi = make_double4(0, 0, 0, 0);
	double2 li = make_double2(0, 0);
	int mi = 0;
	if (i < n) { pi = pq[i]; li = lj[i]; mi = meta[i]; }
	{
		int j = j0 + tid;
		if (j < n) { s_pq[tid] = pq[j]; s_lj[tid] = lj[j]; s_meta[tid] = meta[j]; }
	}
	__syncthreads();

	double a_rd = 0, a_re = 0, a_in = 0, a_cnt = 0;
	const int jn = min(T, n - j0);
	const int jstart = (tile.x == tile.y) ? tid + 1 : 0;
	const double rc = c.cutoff, alpha = c.ewald_alpha;
	if (i < n) {
		for (int jj = jstart; jj < jn; jj++) {
			const double4 pj = s_pq[jj];
			const int mj = s_meta[jj];
			if (meta_frozen(mi) && meta_frozen(mj)) continue;                 // pair->frozen
			const bool same = meta_mol(mi) == meta_mol(mj);
			const double ddx = __dsub_rn(pi.x, pj.x), ddy = __dsub_rn(pi.y, pj.y), ddz = __dsub_rn(pi.z, pj.z);
			double dx, dy, dz;
			min_image<ORTHO>(c, ddx, ddy, ddz, dx, dy, dz);
			const double rimg = sqrt(norm2_nofma(dx, dy, dz));
			const double2 ljj = s_lj[jj];
			const double eps = li.x * ljj.x;                                    // sqrt(eps_i) sqrt(eps_j); 0 when either site is LJ-null
			const bool in_lj = (rimg - kSmallDr < rc);
			double inv_r = 0;
			if (in_lj || ES) inv_r = 1.0 / rimg;
			if (in_lj && !same && eps != 0.0) {                                  // :934-937, rd_excluded
				double s = (li.y + ljj.y) * inv_r;                               // sigma_ij / rimg, sigma_ij = (s_i + s_j)/2
				double s3 = s * s * s, s6 = s3 * s3;
				a_rd += 4.0 * eps * (s6 * s6 - s6);
				a_cnt += 1.0;
			}
			if (ES) {
				const double qq = pi.w * pj.w;
				if (qq != 0.0) {
					if (!same) { if (!(rimg > rc)) a_re += qq * erfc(alpha * rimg) * inv_r; }   // :1490-1497
					else {                                                               // :1503-1504: un-imaged distance
						const double r = sqrt(norm2_nofma(ddx, ddy, ddz));
						a_in += qq * erf(alpha * r) / r;
					}
				}
			}
		}
	}
	a_rd = warp_sum(a_rd); a_re = warp_sum(a_re); a_in = warp_sum(a_in); a_cnt = warp_sum(a_cnt);
	const int w = tid >> 5;
	if ((tid & 31) == 0) { s_red[0][w] = a_rd; s_red[1][w] = a_re; s_red[2][w] = a_in; s_red[3][w] = a_cnt; }
	__syncthreads();
	if (tid == 0) {
		PairPartial p = {0, 0, 0, 0};
		for (int k = 0; k < T / 32; k++) { p.rd += s_red[0][k]; p.es_real += s_red[1][k]; p.es_intra += s_red[2][k]; p.n_in += s_red[3][k]; }
		partials[(size_t)bead * ntiles + blockIdx.x] = p;
	}
}

// Sum the per-tile partials of one bead in a fixed order (strided per thread, then a shared-memory tree) into the per-bead
// result record: res[bead*kResStride + 0..3] = rd_pair, es_real, es_intra, n_in.
constexpr int kResStride = 8;   // + 4 es_recip, 5 sum mu.E_s, 6 sum mu.dE_ind, 7 sum rrms
__global__ void __launch_bounds__(256)
k_reduce_partials(const PairPartial *__restrict__ partials, int ntiles, double *__restrict__ res) {
	__shared__ double s_red[4][256];
	const int bead = blockIdx.x, tid = threadIdx.x;
	double a = 0, b = 0, c = 0, d = 0;
	for (int t = tid; t < ntiles; t += 256) {
		const PairPartial p = partials[(size_t)bead * ntiles + t];
		a += p.rd; b += p.es_real; c += p.es_intra; d += p.n_in;
	}
	s_red[0][tid] = a; s_red[1][tid] = b; s_red[2][tid] = c; s_red[3][tid] = d;
	__syncthreads();
	for (int o = 128; o > 0; o >>= 1) {
		if (tid < o) for (int q = 0; q < 4; q++) s_red[q][tid] += s_red[q][tid + o];
		__syncthreads();
	}
	if (tid < kResStride) res[bead * kResStride + tid] = tid < 4 ? s_red[tid][0] : 0.0;   // slots 4..7 are filled by later kernels
}

// PI_calculate_potential (SimulationControl.PathIntegral.cpp:786-796): sums over this engine's bead systems of
// rd, coulombic, polarization, vdw — assembled on the device so that a collective can follow without a host round trip.
// konst = { lrc_pair + lrc_self, es_self, palmo flag, polarization flag }
__global__ void k_pi_sums(const double *__restrict__ res, int nbeads, double rd_const, double es_self, int es_on, int polar_on, int palmo,
                          double *__restrict__ per_bead, double *__restrict__ sums) {
	if (threadIdx.x != 0 || blockIdx.x != 0) return;
	double s0 = 0, s1 = 0, s2 = 0;
	for (int b = 0; b < nbeads; b++) {
		const double *r = res + b * kResStride;
		const double rd = r[0] + rd_const;
		const double es = es_on ? (r[1] - r[2]) + r[4] + es_self : 0.0;
		double pol = 0;
		if (es_on && polar_on) pol = -0.5 * (r[5] + (palmo ? r[6] : 0.0));
		if (per_bead) { per_bead[4 * b] = rd; per_bead[4 * b + 1] = es; per_bead[4 * b + 2] = pol; per_bead[4 * b + 3] = 0; }
		s0 += rd; s1 += es; s2 += pol;
	}
	sums[0] = s0; sums[1] = s1; sums[2] = s2; sums[3] = 0;
}

// The same assembly fused with the cross-GPU sum (replaces MPI_Allgather x 4, PathIntegral.cpp:763-766, and the ncclAllReduce of the
// first version): every rank owns a mailbox of [2 parities][nranks] slots in its own HBM, mapped into every peer through CUDA IPC.
// A rank stores its four local sums straight into slot [parity][rank] of every peer's mailbox over NVLink, fences, then stamps the
// slot with the step number; it then waits for the nranks stamps in its own mailbox and adds the slots in rank order — the same
// order on every rank, so all ranks hold bit-identical sums.  One kernel, no host round trip, ~3 us instead of the ~70 us the
// 8-rank collective cost inside a 250 us step.  Two parities suffice: a rank can only be one step ahead of the slowest reader.
struct PiMailSlot { double v[4]; long long seq; long long pad[3]; };   // 64 bytes
__global__ void k_pi_sums_xchg(const double *__restrict__ res, int nbeads, double rd_const, double es_self, int es_on, int polar_on, int palmo,
                               double *__restrict__ sums, PiMailSlot *const *__restrict__ peers, int rank, int nranks, long long *step_counter) {
	__shared__ double s_loc[4];
	__shared__ double s_in[32][4];
	__shared__ long long s_step;
	const int t = threadIdx.x;
	if (t == 0) {
		double s0 = 0, s1 = 0, s2 = 0;
		for (int b = 0; b < nbeads; b++) {
			const double *r = res + b * kResStride;
			s0 += r[0] + rd_const;
			s1 += es_on ? (r[1] - r[2]) + r[4] + es_self : 0.0;
			if (es_on && polar_on) s2 += -0.5 * (r[5] + (palmo ? r[6] : 0.0));
		}
		s_loc[0] = s0; s_loc[1] = s1; s_loc[2] = s2; s_loc[3] = 0;
		s_step = ++(*step_counter);
	}
	__syncthreads();
	const long long step = s_step;
	const int par = (int)(step & 1);
	if (t < nranks) {
		volatile PiMailSlot *dst = peers[t] + par * nranks + rank;
		for (int q = 0; q < 4; q++) dst->v[q] = s_loc[q];
		__threadfence_system();
		dst->seq = step;
		volatile PiMailSlot *src = peers[rank] + par * nranks + t;
		const long long t0 = clock64();
		bool ok = true;
		while (src->seq != step) { if (clock64() - t0 > 8000000000ll) { ok = false; break; } }   // ~4 s: a peer died; fail loudly with NaN
		__threadfence_system();
		for (int q = 0; q < 4; q++) s_in[t][q] = ok ? src->v[q] : __longlong_as_double(0x7ff8000000000000ll);
	}
	__syncthreads();
	if (t < 4) {
		double a = 0;
		for (int r = 0; r < nranks; r++) a += s_in[r][t];
		sums[t] = a;
	}
}

// ---------------------------------------------------------------------------------------------------------
// Ordered sweeps: CTA = kOrdI sites i x kOrdJ j-lanes (256 threads); thread (il, jl) accumulates site i's sum over
// j = jl, jl+kOrdJ, ... of every j tile; the j-lanes of one site are the 8 neighbouring lanes of a warp and are
// combined with a fixed xor tree.  Each CTA covers ALL j for its sites, so no cross-CTA reduction is needed.
// ---------------------------------------------------------------------------------------------------------
constexpr int kOrdI = 32, kOrdJ = 8, kOrdThreads = kOrdI * kOrdJ, kOrdTileJ = 256;

__device__ __forceinline__ double jlane_sum(double v) {
#pragma unroll
	for (int o = kOrdJ / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}

// K5 (real part): real_term() (src/System.Energy.cpp:2900-2940) when EWALD, thole_field_nopbc() (:3300-3333) otherwise.
// Adds into ef (n*3 per bead), which already holds the reciprocal part (or zeros).
// blk_frozen[b] != 0 when every site of the 32-site block b is frozen; a j tile is skipped when it and the i block are.
template <bool ORTHO, bool EWALD>
__global__ void __launch_bounds__(kOrdThreads)
k_field_real(const double4 *__restrict__ posq, const int *__restrict__ meta, const unsigned char *__restrict__ blk_frozen,
             int n, int stride, CellDev c, double *__restrict__ ef) {
	__shared__ double4 s_pq[kOrdTileJ];
	__shared__ int     s_meta[kOrdTileJ];
	const int bead = blockIdx.y;
	const double4 *pq = posq + (size_t)bead * stride;
	const int tid = threadIdx.x, jl = tid % kOrdJ, il = tid / kOrdJ;
	const int i = blockIdx.x * kOrdI + il;
	double4 pi = make_double4(0, 0, 0, 0);
	int mi = 0;
	if (i < n) { pi = pq[i]; mi = meta[i]; }
	const bool iblk_frozen = blk_frozen[blockIdx.x] != 0;
	const double a = c.polar_alpha, rc = c.cutoff;
	double ex = 0, ey = 0, ez = 0;
	for (int j0 = 0; j0 < n; j0 += kOrdTileJ) {
		if (iblk_frozen) {   // uniform per CTA: skip j tiles that are entirely frozen too
			bool all = true;
			for (int b = j0 / kOrdI; b < min((j0 + kOrdTileJ + kOrdI - 1) / kOrdI, (n + kOrdI - 1) / kOrdI); b++) all = all && blk_frozen[b];
			if (all) continue;
		}
		__syncthreads();
		if (j0 + tid < n) { s_pq[tid] = pq[j0 + tid]; s_meta[tid] = meta[j0 + tid]; }
		__syncthreads();
		const int jn = min(kOrdTileJ, n - j0);
		if (i < n) {
			for (int jj = jl; jj < jn; jj += kOrdJ) {
				const int j = j0 + jj;
				const double4 pj = s_pq[jj];
				const int mj = s_meta[jj];
				if (j == i || pj.w == 0.0) continue;                           // q_j = 0 adds exactly nothing
				if (meta_frozen(mi) && meta_frozen(mj)) continue;               // :2915 / :3311
				const bool same = meta_mol(mi) == meta_mol(mj);
				if (!EWALD && same) continue;                                  // :3313
				double dx, dy, dz;
				min_image<ORTHO>(c, __dsub_rn(pi.x, pj.x), __dsub_rn(pi.y, pj.y), __dsub_rn(pi.z, pj.z), dx, dy, dz);
				const double r2 = norm2_nofma(dx, dy, dz);
				const double r = sqrt(r2);
				double factor;
				if (EWALD) {
					if ((r > rc) || (r == 0.0)) continue;                      // :2917
					const double rr2 = r * r;
					const double g = 2.0 * a * kOneOverSqrtPi * exp(-a * a * rr2) * r;
					if (same || pi.w == 0.0) factor = (g - erf(a * r)) / (r * rr2);   // es_excluded form, :2921
					else                     factor = (g + erfc(a * r)) / (rr2 * r);  // :2929
				} else {
					if (!((r - kSmallDr < rc) && (r != 0.0))) continue;        // :3319
					factor = 1.0 / (r * r * r);
				}
				const double fq = factor * pj.w;
				ex += fq * dx; ey += fq * dy; ez += fq * dz;
			}
		}
	}
	ex = jlane_sum(ex); ey = jlane_sum(ey); ez = jlane_sum(ez);
	if (jl == 0 && i < n) {
		double *e = ef + ((size_t)bead * n + i) * 3;
		e[0] += ex; e[1] += ey; e[2] += ez;
	}
}

__global__ void k_fill_u64(unsigned long long *p, int n, unsigned long long v) {
	for (int i = threadIdx.x; i < n; i += blockDim.x) p[i] = v;
}

// rank metric, pass 1: smallest minimum-image separation between two polarizable sites (src/System.cpp:1003-1010).
// Positive doubles order like their bit patterns, so an integer atomicMin is exact and order-independent.
template <bool ORTHO>
__global__ void __launch_bounds__(kOrdThreads)
k_rank_rmin(const double4 *__restrict__ posq, const double *__restrict__ alpha, int n, int stride, CellDev c,
            unsigned long long *__restrict__ rmin_bits) {
	__shared__ double4 s_pq[kOrdTileJ];
	__shared__ double  s_al[kOrdTileJ];
	const int bead = blockIdx.y;
	const double4 *pq = posq + (size_t)bead * stride;
	const int tid = threadIdx.x, jl = tid % kOrdJ, il = tid / kOrdJ;
	const int i = blockIdx.x * kOrdI + il;
	double4 pi = make_double4(0, 0, 0, 0);
	double ai = 0;
	if (i < n) { pi = pq[i]; ai = alpha[i]; }
	double best = kMaxValue;
	for (int j0 = 0; j0 < n; j0 += kOrdTileJ) {
		__syncthreads();
		if (j0 + tid < n) { s_pq[tid] = pq[j0 + tid]; s_al[tid] = alpha[j0 + tid]; }
		__syncthreads();
		const int jn = min(kOrdTileJ, n - j0);
		if (i < n && ai != 0.0)
			for (int jj = jl; jj < jn; jj += kOrdJ) {
				if (j0 + jj <= i || s_al[jj] == 0.0) continue;
				const double4 pj = s_pq[jj];
				double dx, dy, dz;
				min_image<ORTHO>(c, __dsub_rn(pi.x, pj.x), __dsub_rn(pi.y, pj.y), __dsub_rn(pi.z, pj.z), dx, dy, dz);
				best = fmin(best, sqrt(norm2_nofma(dx, dy, dz)));
			}
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) best = fmin(best, __shfl_xor_sync(0xffffffffu, best, o));
	if ((tid & 31) == 0) atomicMin(rmin_bits + bead, (unsigned long long)__double_as_longlong(best));
}

// rank metric, pass 2: number of polarizable partners whose UN-imaged separation is <= 1.5 rmin (src/System.cpp:1016-1027)
__global__ void __launch_bounds__(kOrdThreads)
k_rank_count(const double4 *__restrict__ posq, const double *__restrict__ alpha, int n, int stride,
             const unsigned long long *__restrict__ rmin_bits, double *__restrict__ rank) {
	__shared__ double4 s_pq[kOrdTileJ];
	__shared__ double  s_al[kOrdTileJ];
	const int bead = blockIdx.y;
	const double4 *pq = posq + (size_t)bead * stride;
	const int tid = threadIdx.x, jl = tid % kOrdJ, il = tid / kOrdJ;
	const int i = blockIdx.x * kOrdI + il;
	double4 pi = make_double4(0, 0, 0, 0);
	double ai = 0;
	if (i < n) { pi = pq[i]; ai = alpha[i]; }
	const double lim = __dmul_rn(__longlong_as_double((long long)rmin_bits[bead]), 1.5);
	double cnt = 0;
	for (int j0 = 0; j0 < n; j0 += kOrdTileJ) {
		__syncthreads();
		if (j0 + tid < n) { s_pq[tid] = pq[j0 + tid]; s_al[tid] = alpha[j0 + tid]; }
		__syncthreads();
		const int jn = min(kOrdTileJ, n - j0);
		if (i < n && ai != 0.0)
			for (int jj = jl; jj < jn; jj += kOrdJ) {
				if (j0 + jj == i || s_al[jj] == 0.0) continue;
				const double4 pj = s_pq[jj];
				// |d| is symmetric in (i,j) bit for bit, so counting over ordered pairs equals the reference's i<j double update
				const double r = sqrt(norm2_nofma(__dsub_rn(pi.x, pj.x), __dsub_rn(pi.y, pj.y), __dsub_rn(pi.z, pj.z)));
				if (r <= lim) cnt += 1.0;
			}
	}
	cnt = jlane_sum(cnt);
	if (jl == 0 && i < n) rank[(size_t)bead * n + i] = cnt;
}

// stable descending sort of the site indices by rank metric == the reference's bubble sort (System.Energy.cpp:3631-3656):
// position(i) = #{j : m_j > m_i} + #{j < i : m_j == m_i}
__global__ void k_rank_order(const double *__restrict__ rank, int n, int *__restrict__ order) {
	__shared__ double s_m[256];
	const int bead = blockIdx.y;
	const double *m = rank + (size_t)bead * n;
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	const double mi = i < n ? m[i] : 0;
	int pos = 0;
	for (int j0 = 0; j0 < n; j0 += 256) {
		__syncthreads();
		if (j0 + threadIdx.x < n) s_m[threadIdx.x] = m[j0 + threadIdx.x];
		__syncthreads();
		const int jn = min(256, n - j0);
		for (int jj = 0; jj < jn; jj++) {
			const double mj = s_m[jj];
			pos += (mj > mi) || (mj == mi && j0 + jj < i);
		}
	}
	if (i < n) order[(size_t)bead * n + pos] = i;
}

} // namespace mpmc
