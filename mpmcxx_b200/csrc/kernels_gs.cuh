// kernels_gs.cuh — Gauss-Seidel dipole sweep, mathematically sequential in the sweep order, as a software pipeline.
//
// contract_dipoles() with polar_gs / polar_gs_ranked (reference src/System.Energy.cpp:3570-3595) overwrites mu_i as soon
// as it is computed, so site i sees the NEW dipoles of every site swept before it and the OLD dipoles of the rest: a dense
// triangular solve with N sequential steps.  The engine keeps, for every polarizable site, the running contraction
//       acc_i = sum_{j != i} T_ij mu_j(current)
// so that a site's update is just  mu_i = alpha_i (E_s,i - acc_i),  and the change  dmu_i = mu_i(new) - mu_i(old)  is pushed
// into every other row:  acc_m += T_mi dmu_i.  Sites are processed in blocks of kGsB = 64 in sweep order ("panels"):
//   * the SOLVER CTA (block 0) owns the critical path.  One warp walks a block (two rows per lane in registers, the block's
//     tensors in shared memory, tensor loads issued one step ahead of the dependent chain).  While it walks, six other warps
//     push every column, as soon as it is final, into the 64 rows of the NEXT block (so that push costs the critical path only
//     its tail), a seventh fetches the next block's site columns, and the next block's tensors stream into the second
//     shared-memory buffer with cp.async.  Nothing on the critical path crosses the chip.
//   * the UPDATER CTAs (one per remaining SM) push each published panel into all other rows — 8 rows per warp, 4 column lanes
//     per row, every warp on its own — in panel order, and flag each 8-row chunk when it has received a panel.  The rows of
//     panels p and p+1 are the solver's, so an updater has a whole block period before its work is needed: its flag latency is
//     off the critical path.
//   * the solver may start block b when its rows have received panels 0..b-2 from the updaters (panel b-1 is its own push).
// Every row receives its updates in a fixed order, so the result does not depend on timing.  Pushing panels into rows that
// were already swept prepares acc for the next sweep, and after the last sweep acc_i is exactly the contraction
// palmo_contraction() needs (:3602-3627), so Palmo costs no extra sweep.
// One cooperative launch runs `nsweeps` sweeps (grid barrier between sweeps); all CTAs are co-resident, which makes the
// flag waits safe.
#pragma once
#include <cuda_pipeline.h>
#include "kernels_polar2.cuh"

namespace mpmc {

constexpr int kGsB = 64;                  // sites per solver block
constexpr int kGsRows = 8;                // rows per updater chunk (one warp: 8 rows x 4 column lanes)
constexpr int kGsThreads = 256;
constexpr int kGsWarps = kGsThreads / 32;
constexpr int kGsPairs = kGsB * (kGsB - 1) / 2;
constexpr int kGsTriLen = 6 * (kGsPairs + 1);   // doubles per tensor buffer (+1: an all-zero dummy pair)
__host__ __device__ constexpr int gs_tri(int a, int b) { return a * (2 * kGsB - a - 1) / 2 + (b - a - 1); }   // a < b

// shared memory (doubles).  Solver: two tensor buffers, two site-column buffers, pending push (3 slices + sum), two row buffers,
// panel dmu, ints.  Updaters: per warp the panel's columns and dmu.
constexpr int kGsSiteCols = 10;           // 0 alpha, 1-3 mu_old, 4-6 E_static, 7-9 acc
constexpr size_t kGsSolverDoubles = 2 * (size_t)kGsTriLen + 2 * kGsSiteCols * kGsB + 4 * 3 * kGsB + 2 * 4 * kGsB + 4 * kGsB + 3 * kGsB + 16;
constexpr size_t kGsUpdaterDoubles = (size_t)kGsWarps * 8 * kGsB;
constexpr size_t kGsSmemBytes = sizeof(double) * (kGsSolverDoubles > kGsUpdaterDoubles ? kGsSolverDoubles : kGsUpdaterDoubles);

struct GsCtl { int solved; int pad[31]; };   // followed in memory by int applied[nchunks]

__device__ __forceinline__ int ld_flag(const int *p) { return *(const volatile int *)p; }
__device__ __forceinline__ void st_flag(int *p, int v) { *(volatile int *)p = v; }

// sweep-order copies of what the pipeline reads per site: gpq[pos] = x, y, z, alpha; gmeta[pos] = molecule | charged<<30
__global__ void k_gs_gather(const double4 *__restrict__ pq, const double *__restrict__ alpha, const int *__restrict__ meta,
                            const int *__restrict__ order, int np, double4 *__restrict__ gpq, int *__restrict__ gmeta) {
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= np) return;
	const int s = order[t];
	const double4 p = pq[s];
	gpq[t] = make_double4(p.x, p.y, p.z, alpha[s]);
	gmeta[t] = (meta[s] & 0x3fffffff) | (p.w != 0.0 ? 0x40000000 : 0);
}

// acc += T_rc dmu_c for one (row, column) pair of the sweep order, either damping model
template <bool ORTHO, bool EXPD>
__device__ __forceinline__ void gs_contract(const CellDev &c, const PolarDev &p, const double4 &pr, int mr, const double4 &pc, int mc,
                                            const double4 &dm, double &ax, double &ay, double &az) {
	if (EXPD) tensor_contract_exp<ORTHO>(c, p.damp, p.u_damp, pr.x, pr.y, pr.z, pc.x, pc.y, pc.z, dm.x, dm.y, dm.z, ax, ay, az);
	else {
		const bool excl = ((mr & 0x3fffffff) == (mc & 0x3fffffff)) || !(mr & 0x40000000) || !(mc & 0x40000000);
		tensor_contract<ORTHO>(c, p, pr.x, pr.y, pr.z, pc.x, pc.y, pc.z, excl, pr.w * pc.w, dm.x, dm.y, dm.z, ax, ay, az);
	}
}

// in-block tensors for every block of the sweep order: tri[blk][kGsPairs][6] = xx yy zz xy xz yz of T_ab, a < b in block
template <bool ORTHO>
__global__ void __launch_bounds__(kGsThreads)
k_gs_tensors(const double4 *__restrict__ gpq, const int *__restrict__ gmeta, int np, CellDev c, PolarDev p, double *__restrict__ tri) {
	__shared__ double4 s_pq[kGsB];
	__shared__ int     s_met[kGsB];
	const int blk = blockIdx.x, base = blk * kGsB, cnt = min(kGsB, np - base), tid = threadIdx.x;
	if (tid < cnt) { s_pq[tid] = gpq[base + tid]; s_met[tid] = gmeta[base + tid]; }
	__syncthreads();
	double *out = tri + (size_t)blk * 6 * kGsPairs;
	for (int q = tid; q < cnt * cnt; q += kGsThreads) {
		const int a = q / cnt, b = q % cnt;
		if (a >= b) continue;
		// the tensor itself = the contraction applied to the three unit dipoles (columns of T)
		double xx = 0, xy = 0, xz = 0, yy = 0, yz = 0, zz = 0, t0 = 0, t1 = 0;
		if (p.damp_type == 2) {
			gs_contract<ORTHO, true>(c, p, s_pq[a], s_met[a], s_pq[b], s_met[b], make_double4(1, 0, 0, 0), xx, xy, xz);
			gs_contract<ORTHO, true>(c, p, s_pq[a], s_met[a], s_pq[b], s_met[b], make_double4(0, 1, 0, 0), t0, yy, yz);
			gs_contract<ORTHO, true>(c, p, s_pq[a], s_met[a], s_pq[b], s_met[b], make_double4(0, 0, 1, 0), t0, t1, zz);
		} else {
			gs_contract<ORTHO, false>(c, p, s_pq[a], s_met[a], s_pq[b], s_met[b], make_double4(1, 0, 0, 0), xx, xy, xz);
			gs_contract<ORTHO, false>(c, p, s_pq[a], s_met[a], s_pq[b], s_met[b], make_double4(0, 1, 0, 0), t0, yy, yz);
			gs_contract<ORTHO, false>(c, p, s_pq[a], s_met[a], s_pq[b], s_met[b], make_double4(0, 0, 1, 0), t0, t1, zz);
		}
		const int k = gs_tri(a, b);
		// pair-major: xx yy zz xy xz yz of pair k are 48 contiguous bytes (three 128-bit shared loads in the solver's walk)
		out[6 * k + 0] = xx; out[6 * k + 1] = yy; out[6 * k + 2] = zz;
		out[6 * k + 3] = xy; out[6 * k + 4] = xz; out[6 * k + 5] = yz;
	}
}

// index of pair (m, k) in the block's triangular store; the all-zero dummy pair for rows that must not move
__device__ __forceinline__ int gs_pair_index(int m, int k) {
	const int lo = min(m, k), hi = max(m, k);
	return (m == k || m < 0) ? kGsPairs : lo * (2 * kGsB - lo - 1) / 2 + (hi - lo - 1);
}

template <bool ORTHO, bool EXPD>
__global__ void __launch_bounds__(kGsThreads, 1)
k_gs_pipeline(const double4 *__restrict__ gpq, const int *__restrict__ gmeta, const int *__restrict__ order, int np, CellDev c, PolarDev p,
              const double *__restrict__ efs, double *mu, double *efi, double *new_mu, double *acc, double *dmu,
              const double *__restrict__ tri, GsCtl *ctl, int nsweeps, long long *prof) {
	cg::grid_group grid = cg::this_grid();
	extern __shared__ __align__(16) double s_raw[];
	int *applied = (int *)(ctl + 1);
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int G = gridDim.x, cta = blockIdx.x, U = G - 1;
	const int nblk = (np + kGsB - 1) / kGsB, nchunks = (np + kGsRows - 1) / kGsRows;
	constexpr int kChunksPerBlk = kGsB / kGsRows;

	for (int sweep = 0; sweep < nsweeps; sweep++) {
		if (cta == 0) {
			// ------------------------------------------------ solver ------------------------------------------------
			double *s_tri = s_raw;                                         // [2][kGsTriLen]
			double *s_site = s_tri + 2 * kGsTriLen;                        // [2][kGsSiteCols][kGsB]
			double *s_pendp = s_site + 2 * kGsSiteCols * kGsB;             // [3 slices][kGsB][3] partial pushes, then [kGsB][3] their sum
			double *s_pend = s_pendp + 3 * 3 * kGsB;
			double4 *s_rows = (double4 *)(s_pend + 3 * kGsB);              // [2][kGsB] x y z alpha of this block / the next block
			double4 *s_dm = s_rows + 2 * kGsB;                             // [kGsB] dmu of the block being walked
			int *s_meta = (int *)(s_dm + kGsB);                            // [2][kGsB]
			int *s_idx = s_meta + 2 * kGsB;                                // [2][kGsB] site ids
			volatile int *s_prog = (volatile int *)(s_idx + 2 * kGsB);     // columns of the current walk that are final
			auto load_tri = [&](int blk) {                                 // cp.async: tensors of block blk -> buffer blk & 1 (+ commit)
				const double2 *tsrc = (const double2 *)(tri + (size_t)blk * 6 * kGsPairs);
				double2 *tdst = (double2 *)(s_tri + (blk & 1) * kGsTriLen);
				for (int q = tid; q < 3 * kGsPairs; q += kGsThreads) __pipeline_memcpy_async(tdst + q, tsrc + q, sizeof(double2));
				__pipeline_commit();
			};
			auto load_rows = [&](int blk, int m) {                         // position, alpha, meta and site id of row m of block blk
				const int pos = blk * kGsB + m, b = blk & 1;
				const bool on = pos < np;
				s_idx[b * kGsB + m] = on ? order[pos] : 0;
				s_rows[b * kGsB + m] = on ? gpq[pos] : make_double4(0, 0, 0, 0);
				s_meta[b * kGsB + m] = on ? gmeta[pos] : 0;
			};
			auto load_cols = [&](int blk, int m) {                         // site columns of row m (all but the running contraction); after load_rows
				const int b = blk & 1;
				const bool on = blk * kGsB + m < np;
				const int s = s_idx[b * kGsB + m];
				double *sc = s_site + b * kGsSiteCols * kGsB;
				sc[m] = s_rows[b * kGsB + m].w;
				for (int q = 0; q < 3; q++) {
					sc[(1 + q) * kGsB + m] = on ? __ldcg(mu + 3 * s + q) : 0.0;
					sc[(4 + q) * kGsB + m] = on ? efs[3 * s + q] : 0.0;
				}
			};
			if (tid < 12) { s_tri[6 * kGsPairs + (tid % 6) + (tid / 6) * kGsTriLen] = 0.0; }   // dummy pair of both buffers
			for (int q = tid; q < 3 * kGsB; q += kGsThreads) s_pend[q] = 0.0;
			load_tri(0);
			if (tid < kGsB) { load_rows(0, tid); load_cols(0, tid); }
			if (tid == 0) *s_prog = 0;
			for (int blk = 0; blk < nblk; blk++) {
				const int base = blk * kGsB, cnt = min(kGsB, np - base), cur = blk & 1;
				double *ss = s_site + cur * kGsSiteCols * kGsB;
				if (prof && tid == 0 && sweep == 0) prof[blk * 8 + 0] = clock64();
				// (A) rows of this block must have received panels 0..blk-2 from the updaters; then the running contraction is the
				//     global value plus my own push of the previous panel
				const int c0 = base / kGsRows, c1 = (base + cnt + kGsRows - 1) / kGsRows;
				if (blk >= 2 && tid >= 128 && tid - 128 < c1 - c0) while (ld_flag(applied + c0 + tid - 128) < blk - 1) { }
				if (tid >= kGsB && tid < 2 * kGsB && blk + 1 < nblk) load_rows(blk + 1, tid - kGsB);     // the pushing warps need them during the walk
				__syncthreads();
				__threadfence();
				if (tid < cnt) {
					const int s = s_idx[cur * kGsB + tid];
					for (int q = 0; q < 3; q++) ss[(7 + q) * kGsB + tid] = __ldcg(acc + 3 * s + q) + s_pend[3 * tid + q];
				}
				__pipeline_wait_prior(0);
				__syncthreads();
				if (prof && tid == 0 && sweep == 0) prof[blk * 8 + 1] = clock64();
				// (B) warp 0 walks; warps 1-3 and 5-7 push every column into the next block's rows as soon as it is final; warp 4
				//     (the walker's scheduler partner) only fetches the next block's site columns
				if (blk + 1 < nblk) load_tri(blk + 1);
				if (warp == 0) {
					const double *tb = s_tri + cur * kGsTriLen;
					// lane owns rows lane (slot 0) and lane+32 (slot 1); everything a row needs lives in registers during the walk.
					// With c = alpha E_s - mu_old the change of a dipole is a single FMA:  dmu = c - alpha acc.
					double al[2], cx[2], cy[2], cz[2], ax[2], ay[2], az[2], ex[2], ey[2], ez[2];
					int mrow[2];
#pragma unroll
					for (int h = 0; h < 2; h++) {
						const int m = lane + 32 * h;
						al[h] = ss[m];
						cx[h] = al[h] * ss[4 * kGsB + m] - ss[1 * kGsB + m];
						cy[h] = al[h] * ss[5 * kGsB + m] - ss[2 * kGsB + m];
						cz[h] = al[h] * ss[6 * kGsB + m] - ss[3 * kGsB + m];
						ax[h] = ss[7 * kGsB + m]; ay[h] = ss[8 * kGsB + m]; az[h] = ss[9 * kGsB + m];
						ex[h] = ey[h] = ez[h] = 0;
						mrow[h] = m < cnt ? m : -1;                               // rows past the end never match and never move
					}
					// tensor entries of column 0 for my two rows (independent of the dipoles: always one step ahead of the chain)
					double2 tn[2][3];
#pragma unroll
					for (int hh = 0; hh < 2; hh++) {
						const double2 *tp = (const double2 *)(tb + 6 * gs_pair_index(mrow[hh], 0));
						tn[hh][0] = tp[0]; tn[hh][1] = tp[1]; tn[hh][2] = tp[2];
					}
#pragma unroll
					for (int half = 0; half < 2; half++) {
						const int kend = min(32, cnt - 32 * half);
						for (int kk = 0; kk < kend; kk++) {
							const int k = kk + 32 * half;
							double2 tc[2][3];
#pragma unroll
							for (int hh = 0; hh < 2; hh++) {
								tc[hh][0] = tn[hh][0]; tc[hh][1] = tn[hh][1]; tc[hh][2] = tn[hh][2];     // (xx yy) (zz xy) (xz yz)
								const double2 *tp = (const double2 *)(tb + 6 * gs_pair_index(mrow[hh], min(k + 1, kGsB - 1)));
								tn[hh][0] = tp[0]; tn[hh][1] = tp[1]; tn[hh][2] = tp[2];
							}
							// every lane forms the candidate change of its own slot-`half` row; the owner's is the real one
							const double dxc = fma(-al[half], ax[half], cx[half]), dyc = fma(-al[half], ay[half], cy[half]), dzc = fma(-al[half], az[half], cz[half]);
							const double dx = __shfl_sync(0xffffffffu, dxc, kk), dy = __shfl_sync(0xffffffffu, dyc, kk), dz = __shfl_sync(0xffffffffu, dzc, kk);
							if (lane == kk) { ex[half] = ax[half]; ey[half] = ay[half]; ez[half] = az[half]; }   // acc at the moment of the update
#pragma unroll
							for (int hh = 0; hh < 2; hh++) {
								ax[hh] = fma(tc[hh][0].x, dx, fma(tc[hh][1].y, dy, fma(tc[hh][2].x, dz, ax[hh])));
								ay[hh] = fma(tc[hh][1].y, dx, fma(tc[hh][0].y, dy, fma(tc[hh][2].y, dz, ay[hh])));
								az[hh] = fma(tc[hh][2].x, dx, fma(tc[hh][2].y, dy, fma(tc[hh][1].x, dz, az[hh])));
							}
							// hand the finished column to the pushing warps (published in groups of 4 columns)
							if (lane == 0) {
								s_dm[k] = make_double4(dx, dy, dz, 0.0);
								if ((k & 3) == 3 || k == cnt - 1) { __threadfence_block(); *s_prog = k + 1; }
							}
						}
					}
					if (prof && tid == 0 && sweep == 0) prof[blk * 8 + 2] = clock64();
#pragma unroll
					for (int hh = 0; hh < 2; hh++) {
						const int m = lane + 32 * hh;
						if (m < cnt) {
							const int s = s_idx[cur * kGsB + m];
							// contract_dipoles: ef_induced = -acc at the moment of the update, mu = alpha (E_s + ef_induced)  (:3583-3592);
							// the published change is recomputed exactly as the walk formed it
							const double fx = -ex[hh], fy = -ey[hh], fz = -ez[hh];
							const double nx = al[hh] * (ss[4 * kGsB + m] + fx), ny = al[hh] * (ss[5 * kGsB + m] + fy), nz = al[hh] * (ss[6 * kGsB + m] + fz);
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
				} else if (warp == 4) {
					if (blk + 1 < nblk) { load_cols(blk + 1, lane); load_cols(blk + 1, lane + 32); }
				} else if (blk + 1 < nblk) {
					const int h = (warp < 4 ? warp - 1 : warp - 2) * 32 + lane;       // 0..191
					const int r = h & (kGsB - 1), sl = h >> 6;                         // row of the next block, column slice (warp-uniform)
					const double4 pr = s_rows[(cur ^ 1) * kGsB + r];
					const int mr = s_meta[(cur ^ 1) * kGsB + r];
					const bool on = base + kGsB + r < np;
					double ax = 0, ay = 0, az = 0;
					for (int k = sl; k < cnt; k += 3) {
						while (*s_prog <= k) { }
						const volatile double *vd = (const volatile double *)(s_dm + k);
						const double4 dm = make_double4(vd[0], vd[1], vd[2], 0.0);
						if (on) gs_contract<ORTHO, EXPD>(c, p, pr, mr, s_rows[cur * kGsB + k], s_meta[cur * kGsB + k], dm, ax, ay, az);
					}
					double *o = s_pendp + (sl * kGsB + r) * 3;
					o[0] = ax; o[1] = ay; o[2] = az;
				}
				__syncthreads();
				if (prof && tid == 0 && sweep == 0) prof[blk * 8 + 3] = clock64();
				// (C) my push of this panel into the next block's rows, slices summed in a fixed order
				if (tid < 3 * kGsB) {
					const int r = tid / 3, q = tid % 3;
					s_pend[3 * r + q] = (blk + 1 < nblk) ? (s_pendp[(0 * kGsB + r) * 3 + q] + s_pendp[(1 * kGsB + r) * 3 + q]) + s_pendp[(2 * kGsB + r) * 3 + q] : 0.0;
				}
				if (tid == 0) *s_prog = 0;
			}
		} else {
			// ------------------------------------------------ updaters ----------------------------------------------
			// warps work independently: global warp gwid owns the chunks ch = gwid, gwid + GW, ...  (8 consecutive rows of the
			// sweep order each) and keeps its own copy of the panel in shared memory
			double4 *w_col = (double4 *)s_raw + warp * 2 * kGsB;
			double4 *w_dm = w_col + kGsB;
			const int GW = U * kGsWarps, gwid = (cta - 1) * kGsWarps + warp;
			const int r = lane & 7, cl = lane >> 3;             // row of the chunk, column lane
			for (int blk = 0; blk < nblk; blk++) {
				const int base = blk * kGsB, cnt = min(kGsB, np - base);
				// the panel's own rows and the rows of the next block belong to the solver
				const int skip0 = blk * kChunksPerBlk, skip1 = (blk + 1 < nblk) ? (blk + 2) * kChunksPerBlk : (blk + 1) * kChunksPerBlk;
				bool any = false;
				for (int ch = gwid; ch < nchunks; ch += GW) any = any || !(ch >= skip0 && ch < skip1);
				if (!any) continue;
				__syncwarp();
				for (int cc = lane; cc < cnt; cc += 32) {
					const double4 g = gpq[base + cc];
					w_col[cc] = make_double4(g.x, g.y, g.z, __longlong_as_double((long long)gmeta[base + cc]));   // alpha travels in w_dm.w
				}
				if (lane == 0) while (ld_flag(&ctl->solved) <= blk) __nanosleep(32);
				__syncwarp();
				__threadfence();
				for (int cc = lane; cc < cnt; cc += 32)
					w_dm[cc] = make_double4(__ldcg(dmu + 3 * (base + cc)), __ldcg(dmu + 3 * (base + cc) + 1), __ldcg(dmu + 3 * (base + cc) + 2), EXPD ? 0.0 : gpq[base + cc].w);
				__syncwarp();
				for (int ch = gwid; ch < nchunks; ch += GW) {
					if (ch >= skip0 && ch < skip1) continue;
					const int pos = ch * kGsRows + r;
					const bool on = pos < np;
					const double4 pr = on ? gpq[pos] : make_double4(0, 0, 0, 0);
					const int mr = on ? gmeta[pos] : 0;
					double ax = 0, ay = 0, az = 0;
					if (on) {
#pragma unroll 4
						for (int cc = cl; cc < cnt; cc += 4) {
							double4 pc = w_col[cc];
							const double4 dm = w_dm[cc];
							const int mc = __double2loint(pc.w);
							if (!EXPD) pc.w = dm.w;                           // alpha of the column (linear damping)
							gs_contract<ORTHO, EXPD>(c, p, pr, mr, pc, mc, dm, ax, ay, az);
						}
					}
					ax += __shfl_xor_sync(0xffffffffu, ax, 8); ay += __shfl_xor_sync(0xffffffffu, ay, 8); az += __shfl_xor_sync(0xffffffffu, az, 8);
					ax += __shfl_xor_sync(0xffffffffu, ax, 16); ay += __shfl_xor_sync(0xffffffffu, ay, 16); az += __shfl_xor_sync(0xffffffffu, az, 16);
					if (cl == 0 && on) {
						const int i = order[pos];
						__stcg(acc + 3 * i, __ldcg(acc + 3 * i) + ax);
						__stcg(acc + 3 * i + 1, __ldcg(acc + 3 * i + 1) + ay);
						__stcg(acc + 3 * i + 2, __ldcg(acc + 3 * i + 2) + az);
					}
					__threadfence();
					__syncwarp();
					if (lane == 0) st_flag(applied + ch, blk + 1);
				}
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
