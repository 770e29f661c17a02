// kernels_gs.cuh — Gauss-Seidel dipole sweep, mathematically sequential in the ranked order, as a software pipeline.
//
// contract_dipoles() with polar_gs / polar_gs_ranked (reference src/System.Energy.cpp:3570-3595) overwrites mu_i as soon
// as it is computed, so site i sees the NEW dipoles of every site swept before it and the OLD dipoles of the rest: a dense
// triangular solve with N sequential steps.  Here the engine keeps, for every polarizable site, the running contraction
//       acc_i = sum_{j != i} T_ij mu_j(current)
// so that a site's update is just  mu_i = alpha_i (E_s,i - acc_i),  and the change  dmu_i = mu_i(new) - mu_i(old)  is pushed
// into every other row:  acc_m += T_mi dmu_i.  Sites are processed in blocks of kGsB in sweep order:
//   * the SOLVER CTA (block 0) walks one block with a single warp (rows in registers, in-block tensors in shared memory,
//     precomputed by k_gs_tensors), then publishes dmu of the block and raises `solved`;
//   * the UPDATER CTAs push each published block ("panel") into the rows they own — chunks of kGsRows consecutive rows of
//     the sweep order, dealt round-robin — always in panel order, starting with the chunks right after the panel, and
//     count per chunk how many panels it has received;
//   * the solver may start block b when its rows have received panels 0..b-1.
// The only serial work is therefore the in-block walk; the O(N^2) tensor work streams behind it on all other SMs.  Every
// row receives its updates in a fixed order, so the result does not depend on timing.  Pushing panels into rows that were
// already swept prepares acc for the next sweep, and after the last sweep acc_i is exactly the contraction
// palmo_contraction() needs (:3602-3627), so Palmo costs no extra sweep.
// One cooperative launch runs `nsweeps` sweeps (grid barrier between sweeps); all CTAs are co-resident, which makes the
// flag waits safe.
#pragma once
#include "kernels_polar.cuh"

namespace mpmc {

constexpr int kGsB = 64;                  // sites per solver block
constexpr int kGsRows = 8;                // rows per updater chunk (one warp per row)
constexpr int kGsThreads = 256;
constexpr int kGsPairs = kGsB * (kGsB - 1) / 2;
__host__ __device__ constexpr int gs_tri(int a, int b) { return a * (2 * kGsB - a - 1) / 2 + (b - a - 1); }   // a < b
constexpr size_t kGsSmemBytes = sizeof(double) * (6 * (kGsPairs + 1) + 16 * kGsB) + sizeof(int) * 2 * kGsB;   // +1: an all-zero dummy pair

struct GsCtl { int solved; int pad[31]; };   // followed in memory by int applied[nchunks]

__device__ __forceinline__ int ld_flag(const int *p) { return *(const volatile int *)p; }
__device__ __forceinline__ void st_flag(int *p, int v) { *(volatile int *)p = v; }

// in-block tensors for every block of the sweep order: tri[blk][kGsPairs][6] = xx yy zz xy xz yz of T_ab, a < b in block
template <bool ORTHO>
__global__ void __launch_bounds__(kGsThreads)
k_gs_tensors(const double4 *__restrict__ pq, const double *__restrict__ alpha, const int *__restrict__ meta,
             const int *__restrict__ order, int np, CellDev c, PolarDev p, double *__restrict__ tri) {
	__shared__ double4 s_pq[kGsB];
	__shared__ double  s_al[kGsB];
	__shared__ int     s_met[kGsB];
	const int blk = blockIdx.x, base = blk * kGsB, cnt = min(kGsB, np - base), tid = threadIdx.x;
	if (tid < cnt) { const int s = order[base + tid]; s_pq[tid] = pq[s]; s_al[tid] = alpha[s]; s_met[tid] = meta[s]; }
	__syncthreads();
	double *out = tri + (size_t)blk * 6 * kGsPairs;
	for (int q = tid; q < cnt * cnt; q += kGsThreads) {
		const int a = q / cnt, b = q % cnt;
		if (a >= b) continue;
		double dx, dy, dz;
		min_image<ORTHO>(c, __dsub_rn(s_pq[a].x, s_pq[b].x), __dsub_rn(s_pq[a].y, s_pq[b].y), __dsub_rn(s_pq[a].z, s_pq[b].z), dx, dy, dz);
		const double r2 = norm2_nofma(dx, dy, dz), r = sqrt(r2);
		double ir3, ir5;
		if (r == 0.0) { ir3 = ir5 = kMaxValue; } else { const double ir = 1.0 / r, ir2 = ir * ir; ir3 = ir2 * ir; ir5 = ir3 * ir2; }
		const bool excl = (meta_mol(s_met[a]) == meta_mol(s_met[b])) || s_pq[a].w == 0.0 || s_pq[b].w == 0.0;
		double d1, d2;
		thole_damping(p, r, r2, excl, s_al[a] * s_al[b], d1, d2);
		const double ta = d1 * ir3, tb = 3.0 * d2 * ir5;
		const int t = gs_tri(a, b);
		// pair-major: xx yy zz xy xz yz of pair t are 48 contiguous bytes (three 128-bit shared loads in the solver's walk)
		out[6 * t + 0] = ta - tb * dx * dx; out[6 * t + 1] = ta - tb * dy * dy; out[6 * t + 2] = ta - tb * dz * dz;
		out[6 * t + 3] = -tb * dx * dy;     out[6 * t + 4] = -tb * dx * dz;     out[6 * t + 5] = -tb * dy * dz;
	}
}

constexpr int kGsMaxCh = 4;               // chunks an updater pushes one panel into at a time (rows interleaved for ILP)

template <bool ORTHO>
__global__ void __launch_bounds__(kGsThreads, 2)
k_gs_pipeline(const double4 *__restrict__ pq, const double *__restrict__ alpha, const int *__restrict__ meta,
              const int *__restrict__ order, int np, CellDev c, PolarDev p, const double *__restrict__ efs,
              double *mu, double *efi, double *new_mu, double *acc, double *dmu, const double *__restrict__ tri,
              GsCtl *ctl, int nsweeps, long long *prof) {
	cg::grid_group grid = cg::this_grid();
	extern __shared__ double s_raw[];
	int *applied = (int *)(ctl + 1);
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int G = gridDim.x, cta = blockIdx.x, U = G - 1;
	const int nblk = (np + kGsB - 1) / kGsB, nchunks = (np + kGsRows - 1) / kGsRows;

	for (int sweep = 0; sweep < nsweeps; sweep++) {
		if (cta == 0) {
			// ------------------------------------------------ solver ------------------------------------------------
			double *s_tri = s_raw;                       // [kGsPairs][6]
			double *s_site = s_tri + 6 * (kGsPairs + 1); // [16][kGsB]: 0 alpha, 1-3 mu_old, 4-6 E_static, 7-9 acc
			if (tid < 6) s_tri[6 * kGsPairs + tid] = 0.0;   // dummy pair: rows that must not move (m == k, m >= cnt) read zeros
			int *s_idx = (int *)(s_site + 16 * kGsB);
			for (int blk = 0; blk < nblk; blk++) {
				const int base = blk * kGsB, cnt = min(kGsB, np - base);
				if (prof && tid == 0 && sweep == 0) prof[blk * 8 + 0] = clock64();
				// geometry-only data first: it does not depend on the flags
				{
					const double2 *tsrc = (const double2 *)(tri + (size_t)blk * 6 * kGsPairs);
					double2 *tdst = (double2 *)s_tri;
#pragma unroll 8
					for (int q = tid; q < 3 * kGsPairs; q += kGsThreads) tdst[q] = tsrc[q];
				}
				if (tid < cnt) {
					const int s = order[base + tid];
					s_idx[tid] = s;
					s_site[0 * kGsB + tid] = alpha[s];
					for (int q = 0; q < 3; q++) {
						s_site[(1 + q) * kGsB + tid] = __ldcg(mu + 3 * s + q);
						s_site[(4 + q) * kGsB + tid] = efs[3 * s + q];
					}
				}
				// rows of this block must have received panels 0..blk-1 of this sweep
				const int c0 = base / kGsRows, c1 = (base + cnt + kGsRows - 1) / kGsRows;
				if (prof && tid == 0 && sweep == 0) prof[blk * 8 + 1] = clock64();
				if (tid < c1 - c0) while (ld_flag(applied + c0 + tid) < blk) { }
				__syncthreads();
				__threadfence();
				if (prof && tid == 0 && sweep == 0) prof[blk * 8 + 2] = clock64();
				if (tid < cnt) {
					const int s = s_idx[tid];
					for (int q = 0; q < 3; q++) s_site[(7 + q) * kGsB + tid] = __ldcg(acc + 3 * s + q);
				}
				__syncthreads();
				if (prof && tid == 0 && sweep == 0) prof[blk * 8 + 3] = clock64();
				if (warp == 0) {
					// lane owns rows lane (slot 0) and lane+32 (slot 1); everything a row needs lives in registers during the walk.
					// With c = alpha E_s - mu_old the change of a dipole is a single FMA:  dmu = c - alpha acc.
					double al[2], cx[2], cy[2], cz[2], ax[2], ay[2], az[2], ex[2], ey[2], ez[2];
					int mb[2], mrow[2];
#pragma unroll
					for (int h = 0; h < 2; h++) {
						const int m = lane + 32 * h, mm = min(m, kGsB - 1);
						al[h] = s_site[mm];
						cx[h] = al[h] * s_site[4 * kGsB + mm] - s_site[1 * kGsB + mm];
						cy[h] = al[h] * s_site[5 * kGsB + mm] - s_site[2 * kGsB + mm];
						cz[h] = al[h] * s_site[6 * kGsB + mm] - s_site[3 * kGsB + mm];
						ax[h] = s_site[7 * kGsB + mm]; ay[h] = s_site[8 * kGsB + mm]; az[h] = s_site[9 * kGsB + mm];
						ex[h] = ey[h] = ez[h] = 0;
						mb[h] = m * (2 * kGsB - m - 1) / 2 - m - 1;             // gs_tri(m, k) = mb + k for k > m
						mrow[h] = m < cnt ? m : -1;                               // rows past the end never match and never move
					}
					int rb = -1;                                                  // gs_tri(k, m) = rb + m for m > k;  rb(k) = k(2B-k-1)/2 - k - 1
#pragma unroll
					for (int half = 0; half < 2; half++) {
						const int kend = min(32, cnt - 32 * half);
						for (int kk = 0; kk < kend; kk++) {
							const int k = kk + 32 * half;
							// tensor entries of column k for my two rows; rows that must stay put read the all-zero dummy pair.
							// (independent of the dipoles: issued before the dependent chain)
							double2 t[2][3];
#pragma unroll
							for (int hh = 0; hh < 2; hh++) {
								const int m = mrow[hh];
								int ti = m > k ? rb + m : mb[hh] + k;
								ti = (m == k || m < 0) ? kGsPairs : ti;
								const double2 *tp = (const double2 *)(s_tri + 6 * ti);
								t[hh][0] = tp[0]; t[hh][1] = tp[1]; t[hh][2] = tp[2];        // (xx yy) (zz xy) (xz yz)
							}
							// every lane forms the candidate change of its own slot-`half` row; the owner's is the real one
							const double dxc = fma(-al[half], ax[half], cx[half]), dyc = fma(-al[half], ay[half], cy[half]), dzc = fma(-al[half], az[half], cz[half]);
							const double dx = __shfl_sync(0xffffffffu, dxc, kk), dy = __shfl_sync(0xffffffffu, dyc, kk), dz = __shfl_sync(0xffffffffu, dzc, kk);
							if (lane == kk) { ex[half] = ax[half]; ey[half] = ay[half]; ez[half] = az[half]; }   // acc at the moment of the update
#pragma unroll
							for (int hh = 0; hh < 2; hh++) {
								ax[hh] = fma(t[hh][0].x, dx, fma(t[hh][1].y, dy, fma(t[hh][2].x, dz, ax[hh])));
								ay[hh] = fma(t[hh][1].y, dx, fma(t[hh][0].y, dy, fma(t[hh][2].y, dz, ay[hh])));
								az[hh] = fma(t[hh][2].x, dx, fma(t[hh][2].y, dy, fma(t[hh][1].x, dz, az[hh])));
							}
							rb += kGsB - k - 2;
						}
					}
					if (prof && tid == 0 && sweep == 0) prof[blk * 8 + 4] = clock64();
#pragma unroll
					for (int hh = 0; hh < 2; hh++) {
						const int m = lane + 32 * hh;
						if (m < cnt) {
							const int s = s_idx[m];
							// contract_dipoles: ef_induced = -acc at the moment of the update, mu = alpha (E_s + ef_induced)  (:3583-3592);
							// the published change is recomputed exactly as the walk formed it
							const double fx = -ex[hh], fy = -ey[hh], fz = -ez[hh];
							const double nx = al[hh] * (s_site[4 * kGsB + m] + fx), ny = al[hh] * (s_site[5 * kGsB + m] + fy), nz = al[hh] * (s_site[6 * kGsB + m] + fz);
							__stcg(mu + 3 * s, nx); __stcg(mu + 3 * s + 1, ny); __stcg(mu + 3 * s + 2, nz);
							new_mu[3 * s] = nx; new_mu[3 * s + 1] = ny; new_mu[3 * s + 2] = nz;
							efi[3 * s] = fx; efi[3 * s + 1] = fy; efi[3 * s + 2] = fz;
							__stcg(acc + 3 * s, ax[hh]); __stcg(acc + 3 * s + 1, ay[hh]); __stcg(acc + 3 * s + 2, az[hh]);
							__stcg(dmu + 3 * (base + m), fma(-al[hh], ex[hh], cx[hh]));
							__stcg(dmu + 3 * (base + m) + 1, fma(-al[hh], ey[hh], cy[hh]));
							__stcg(dmu + 3 * (base + m) + 2, fma(-al[hh], ez[hh], cz[hh]));
						}
					}
					__threadfence();
					__syncwarp();
					if (lane == 0) st_flag(&ctl->solved, blk + 1);
					if (prof && tid == 0 && sweep == 0) prof[blk * 8 + 5] = clock64();
				}
				__syncthreads();
			}
		} else {
			// ------------------------------------------------ updaters ----------------------------------------------
			double4 *s_pq = (double4 *)s_raw;                 // [kGsB]
			double  *s_al = (double *)(s_pq + kGsB);          // [kGsB]
			double  *s_dm = s_al + kGsB;                      // [kGsB][3]
			int     *s_met = (int *)(s_dm + 3 * kGsB);        // [kGsB]
			const int u = cta - 1;
			for (int blk = 0; blk < nblk; blk++) {
				const int base = blk * kGsB, cnt = min(kGsB, np - base);
				__syncthreads();
				if (tid < cnt) { const int s = order[base + tid]; s_pq[tid] = pq[s]; s_al[tid] = alpha[s]; s_met[tid] = meta[s]; }
				if (prof && tid == 0 && sweep == 0 && cta == 1) prof[(nblk + blk) * 8 + 0] = clock64();
				if (tid == 0) while (ld_flag(&ctl->solved) <= blk) __nanosleep(20);
				__syncthreads();
				__threadfence();
				if (prof && tid == 0 && sweep == 0 && cta == 1) prof[(nblk + blk) * 8 + 1] = clock64();
				if (tid < cnt * 3) s_dm[tid] = __ldcg(dmu + 3 * base + tid);
				__syncthreads();
				// my chunks (ch = u mod U), starting right behind the panel and wrapping around; the panel's own chunks are the solver's.
				// They are pushed kGsMaxCh at a time: warp w owns row w of each chunk of the group and interleaves them.
				const int cb0 = base / kGsRows, cb1 = (base + cnt + kGsRows - 1) / kGsRows;
				const int cfirst = cb1 + ((u - cb1) % U + U) % U;           // first chunk >= cb1 congruent to u
				const int n_after = cfirst < nchunks ? (nchunks - cfirst + U - 1) / U : 0;
				const int n_before = u < cb0 ? (cb0 - u + U - 1) / U : 0;
				int nch_done = 0;
				// the chunk that lies in the NEXT solver block (if I own one) is pushed alone and flagged first: the solver waits for it
				const bool urgent = n_after > 0 && cfirst < cb1 + kGsB / kGsRows;
				for (int g0 = 0, gsz = urgent ? 1 : kGsMaxCh; g0 < n_after + n_before; g0 += gsz, gsz = kGsMaxCh) {
					int ch[kGsMaxCh], row_i[kGsMaxCh];
					double4 pi[kGsMaxCh];
					double ai[kGsMaxCh], ax[kGsMaxCh], ay[kGsMaxCh], az[kGsMaxCh];
					int mi[kGsMaxCh];
#pragma unroll
					for (int r = 0; r < kGsMaxCh; r++) {
						const int g = g0 + r;
						ch[r] = r >= gsz ? -1 : (g < n_after ? cfirst + g * U : (g < n_after + n_before ? u + (g - n_after) * U : -1));
						const int row = ch[r] * kGsRows + warp;
						row_i[r] = (ch[r] >= 0 && row < np) ? order[row] : -1;
						ax[r] = ay[r] = az[r] = 0;
						if (row_i[r] >= 0) { pi[r] = pq[row_i[r]]; ai[r] = alpha[row_i[r]]; mi[r] = meta[row_i[r]]; }
						else { pi[r] = make_double4(0, 0, 0, 0); ai[r] = 0; mi[r] = 0; }
					}
					for (int jj = lane; jj < cnt; jj += 32) {
						const double4 pj = s_pq[jj];
						const double aj = s_al[jj], mx = s_dm[3 * jj], my = s_dm[3 * jj + 1], mz = s_dm[3 * jj + 2];
						const int mj = s_met[jj];
#pragma unroll
						for (int r = 0; r < kGsMaxCh; r++)
							if (row_i[r] >= 0) {
								const bool excl = (meta_mol(mi[r]) == meta_mol(mj)) || pi[r].w == 0.0 || pj.w == 0.0;
								tensor_contract<ORTHO>(c, p, pi[r].x, pi[r].y, pi[r].z, pj.x, pj.y, pj.z, excl, ai[r] * aj, mx, my, mz, ax[r], ay[r], az[r]);
							}
					}
#pragma unroll
					for (int r = 0; r < kGsMaxCh; r++)
						if (row_i[r] >= 0) {                                  // warp-uniform
							const double sx = warp_sum(ax[r]), sy = warp_sum(ay[r]), sz = warp_sum(az[r]);
							if (lane == 0) {
								const int i = row_i[r];
								__stcg(acc + 3 * i, __ldcg(acc + 3 * i) + sx);
								__stcg(acc + 3 * i + 1, __ldcg(acc + 3 * i + 1) + sy);
								__stcg(acc + 3 * i + 2, __ldcg(acc + 3 * i + 2) + sz);
							}
						}
					__threadfence();
					__syncthreads();
					if (tid < kGsMaxCh) {
						const int g = g0 + tid;
						const int chv = tid >= gsz ? -1 : (g < n_after ? cfirst + g * U : (g < n_after + n_before ? u + (g - n_after) * U : -1));
						if (chv >= 0) st_flag(applied + chv, blk + 1);
					}
					if (prof && tid == 0 && sweep == 0 && cta == 1 && nch_done == 0) prof[(nblk + blk) * 8 + 2] = clock64();
					nch_done += gsz;
				}
				if (prof && tid == 0 && sweep == 0 && cta == 1) { prof[(nblk + blk) * 8 + 3] = clock64(); prof[(nblk + blk) * 8 + 4] = n_after + n_before; }
			}
		}
		__threadfence();
		grid.sync();
		// reset the flags for the next sweep
		for (int q = cta * kGsThreads + tid; q < nchunks; q += G * kGsThreads) applied[q] = 0;
		if (cta == 0 && tid == 0) ctl->solved = 0;
		__threadfence();
		grid.sync();
	}
}

// Palmo after Gauss-Seidel: efic_i = -efi_i - acc_i for the polarizable sites (acc is the final running contraction)
__global__ void k_gs_palmo(const int *__restrict__ plist, int np, const double *__restrict__ efi, const double *__restrict__ acc,
                           double *__restrict__ efic) {
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= np * 3) return;
	const int o = 3 * plist[t / 3] + t % 3;
	efic[o] = -efi[o] - acc[o];
}

} // namespace mpmc
