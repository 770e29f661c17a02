// kernels_gs.cuh — Gauss-Seidel dipole sweep, mathematically sequential in the sweep order, as a software pipeline.
//
// contract_dipoles() with polar_gs / polar_gs_ranked (reference src/System.Energy.cpp:3570-3595) overwrites mu_i as soon
// as it is computed, so site i sees the NEW dipoles of every site swept before it and the OLD dipoles of the rest: a dense
// triangular solve with N sequential steps.  The engine keeps, for every polarizable site, the running contraction
//       acc_i = sum_{j != i} T_ij mu_j(current)
// so that a site's update is just  mu_i = alpha_i (E_s,i - acc_i),  and the change  dmu_i = mu_i(new) - mu_i(old)  is pushed
// into every other row:  acc_m += T_mi dmu_i.  Sites are processed in blocks of kGsB = 64 in sweep order ("panels"):
//   * the SOLVER CTA (block 0) owns the critical path and keeps its SM quiet: one warp walks a block (two rows per lane in
//     registers, the block's tensor matrix in shared memory, tensor loads one column ahead of the dependent chain; ~135
//     cycles per site on B200, see tools/ubench/walk_ubench.cu), a second warp rolls the next block's tensors into the columns
//     the walk has left behind and fetches the next block's site columns, a third publishes the panel.
//   * three HELPER CTAs share a thread-block cluster with the solver.  They read every finished column straight out of the
//     solver's shared memory (distributed shared memory: no trip through L2, no flag round trip) and push it into the rows
//     of the next kGsAhead = 3 blocks, then drop their sums into the solver's shared memory.  Those are the rows the walk
//     needs next; doing them inside the cluster costs the critical path only the tail of the last few columns.
//   * the UPDATER CTAs (every other SM) push each published panel into all remaining rows — 8 rows per warp, 4 column lanes
//     per row, every warp on its own — in panel order, and flag each 8-row chunk when it has received a panel.  An updater
//     has kGsAhead block periods (~20 us) to deliver, several times what a flag - fence - load - compute - fence - flag round
//     trip across the chip takes (~10 us measured), so that latency never reaches the solver.
//   * the solver may start block b when its rows have received panels 0..b-1-kGsAhead from the updaters; the later panels
//     are the cluster's own pushes.
// Every row receives its updates in a fixed order, so the result does not depend on timing.  Pushing panels into rows that
// were already swept prepares acc for the next sweep, and after the last sweep acc_i is exactly the contraction
// palmo_contraction() needs (:3602-3627), so Palmo costs no extra sweep.
// One launch = one sweep; every CTA must be resident (the grid is sized from cudaOccupancyMaxActiveClusters), which makes the
// flag waits safe.
#pragma once
#include <cuda_pipeline.h>
#include "kernels_polar2.cuh"

namespace mpmc {

constexpr int kGsB = 64;                  // sites per solver block
constexpr int kGsRows = 8;                // rows per updater chunk (one warp: 8 rows x 4 column lanes)
constexpr int kGsThreads = 256;
constexpr int kGsWarps = kGsThreads / 32;
constexpr int kGsMat = kGsB * kGsB * 6;   // doubles of one block's tensor matrix: [column k][row m][xx yy zz xy xz yz]

// shared memory (doubles).  Solver: the block's tensor matrix, two site-column buffers, pending push (3 slices + sum), two row
// buffers, panel dmu, the walk's results, ints.  Updaters: per warp the panel's columns and dmu.
constexpr int kGsSiteCols = 10;           // 0 alpha, 1-3 mu_old, 4-6 E_static, 7-9 acc
constexpr int kGsAhead = 3;               // the cluster pushes a panel into this many following blocks itself; the updaters take the rest
constexpr int kGsHelpers = 3, kGsCluster = 1 + kGsHelpers;
constexpr int kGsSlots = kGsAhead + 1;    // ring of per-block buffers
constexpr size_t kGsSolverDoubles = (size_t)kGsMat + kGsSiteCols * kGsB + kGsHelpers * (kGsAhead * kGsB * 3) + kGsSlots * 3 * kGsB + 4 * kGsB + 2 * kGsB + 16;
constexpr size_t kGsUpdaterDoubles = (size_t)kGsWarps * 8 * kGsB;
constexpr size_t kGsSmemBytes = sizeof(double) * kGsSolverDoubles;
constexpr size_t kGsUpdaterSmemBytes = sizeof(double) * kGsUpdaterDoubles;

__device__ int g_gs_debug = 0;            // developer switch (mpmc_debug_gs_profile enable bits 1..): 2 = pushers idle, 4 = no rolling copy
struct GsCtl { int solved; int abort; int pad[30]; };   // followed in memory by int applied[nchunks]
// `abort`: set when a CTA has waited ~2 s for the other kernel of the pipeline — the two kernels were not run side by side (a
// tool that serialises launches).  Everybody then stops waiting, the sweep's result is meaningless and the host reports it.
constexpr int kGsWaitLimit = 20000000;

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

// in-block tensors for every block of the sweep order, as a full matrix so that the walk reads column k with unit stride:
// mat[blk][k][m][6] = xx yy zz xy xz yz of T_mk (zero on the diagonal and for rows/columns past the end)
template <bool ORTHO>
__global__ void __launch_bounds__(kGsThreads)
k_gs_tensors(const double4 *__restrict__ gpq, const int *__restrict__ gmeta, int np, CellDev c, PolarDev p, double *__restrict__ mat) {
	__shared__ double4 s_pq[kGsB];
	__shared__ int     s_met[kGsB];
	const int blk = blockIdx.x, base = blk * kGsB, cnt = min(kGsB, np - base), tid = threadIdx.x;
	if (tid < cnt) { s_pq[tid] = gpq[base + tid]; s_met[tid] = gmeta[base + tid]; }
	__syncthreads();
	double *out = mat + (size_t)blk * kGsMat;
	for (int q = tid; q < kGsB * kGsB; q += kGsThreads) {
		const int a = q / kGsB, b = q % kGsB;
		if (a > b) continue;
		double xx = 0, xy = 0, xz = 0, yy = 0, yz = 0, zz = 0, t0 = 0, t1 = 0;
		if (a < b && b < cnt) {
			// the tensor itself = the contraction applied to the three unit dipoles (columns of T)
			if (p.damp_type == 2) {
				gs_contract<ORTHO, true>(c, p, s_pq[a], s_met[a], s_pq[b], s_met[b], make_double4(1, 0, 0, 0), xx, xy, xz);
				gs_contract<ORTHO, true>(c, p, s_pq[a], s_met[a], s_pq[b], s_met[b], make_double4(0, 1, 0, 0), t0, yy, yz);
				gs_contract<ORTHO, true>(c, p, s_pq[a], s_met[a], s_pq[b], s_met[b], make_double4(0, 0, 1, 0), t0, t1, zz);
			} else {
				gs_contract<ORTHO, false>(c, p, s_pq[a], s_met[a], s_pq[b], s_met[b], make_double4(1, 0, 0, 0), xx, xy, xz);
				gs_contract<ORTHO, false>(c, p, s_pq[a], s_met[a], s_pq[b], s_met[b], make_double4(0, 1, 0, 0), t0, yy, yz);
				gs_contract<ORTHO, false>(c, p, s_pq[a], s_met[a], s_pq[b], s_met[b], make_double4(0, 0, 1, 0), t0, t1, zz);
			}
		}
		double *u = out + ((size_t)a * kGsB + b) * 6, *l = out + ((size_t)b * kGsB + a) * 6;
		u[0] = xx; u[1] = yy; u[2] = zz; u[3] = xy; u[4] = xz; u[5] = yz;
		l[0] = xx; l[1] = yy; l[2] = zz; l[3] = xy; l[4] = xz; l[5] = yz;
	}
}

// The updaters' work, shared by the updater kernel and by the single-launch fallback of the pipeline kernel: CTA `cta` of `U`.
template <bool ORTHO, bool EXPD>
__device__ __forceinline__ void gs_updater_body(double *s_raw, int cta, int U, const double4 *__restrict__ gpq, const int *__restrict__ gmeta,
                                                const int *__restrict__ order, int np, const CellDev &c, const PolarDev &p, double *acc,
                                                const double *dmu, GsCtl *ctl, long long *prof) {
	int *applied = (int *)(ctl + 1);
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int nblk = (np + kGsB - 1) / kGsB, nchunks = (np + kGsRows - 1) / kGsRows;
	constexpr int kChunksPerBlk = kGsB / kGsRows;
	{
		// warps work independently: global warp gwid owns the chunks ch = gwid, gwid + GW, ...  (8 consecutive rows of the
		// sweep order each) and keeps its own copy of the panel in shared memory
		double4 *w_col = (double4 *)s_raw + warp * 2 * kGsB;
		double4 *w_dm = w_col + kGsB;
		// (warp-major numbering: when there are more chunks than warps, the second chunks go one to each CTA instead of eight to a
		// few CTAs — an SM with twice the work falls behind by a few thousand cycles per panel and stalls the solver at the end)
		const int GW = U * kGsWarps, gwid = warp * U + cta;
		const int r = lane & 7, cl = lane >> 3;             // row of the chunk, column lane
		for (int blk = 0; blk < nblk; blk++) {
			const int base = blk * kGsB, cnt = min(kGsB, np - base);
			// the panel's own rows and the rows of the next kGsAhead blocks belong to the cluster
			const int skip0 = blk * kChunksPerBlk, skip1 = min(blk + 1 + kGsAhead, nblk) * kChunksPerBlk;
			const int first = gwid;
			bool any = false;
			for (int ch = first; ch < nchunks; ch += GW) any = any || !(ch >= skip0 && ch < skip1);
			if (!any) continue;
			__syncwarp();
			for (int cc = lane; cc < cnt; cc += 32) {
				const double4 g = gpq[base + cc];
				w_col[cc] = make_double4(g.x, g.y, g.z, __longlong_as_double((long long)gmeta[base + cc]));   // alpha travels in w_dm.w
			}
			const bool pw = prof && cta == 0 && warp == 0 && lane == 0;
			if (pw) prof[(nblk + blk) * 8 + 0] = clock64();
			if (lane == 0) {
				int spins = 0;
				while (ld_flag(&ctl->solved) <= blk && !ld_flag(&ctl->abort)) {
					__nanosleep(32);
					if (++spins > 3 * kGsWaitLimit) st_flag(&ctl->abort, 1);
				}
			}
			if (__shfl_sync(0xffffffffu, ld_flag(&ctl->abort), 0)) return;
			__syncwarp();
			__threadfence();
			if (pw) prof[(nblk + blk) * 8 + 1] = clock64();
			for (int cc = lane; cc < cnt; cc += 32)
				w_dm[cc] = make_double4(__ldcg(dmu + 3 * (base + cc)), __ldcg(dmu + 3 * (base + cc) + 1), __ldcg(dmu + 3 * (base + cc) + 2), EXPD ? 0.0 : gpq[base + cc].w);
			__syncwarp();
			// all my chunks of this panel first, then their running contractions (each row has exactly one writer, so the atomic adds
			// are ordered and the result deterministic), ONE fence, then the flags: a warp that owns two chunks must not pay two fences
			for (int ch0 = first; ch0 < nchunks; ch0 += 2 * GW) {
				double sx[2], sy[2], sz[2];
				int chs[2];
#pragma unroll
				for (int w = 0; w < 2; w++) {
					const int ch = ch0 + w * GW;
					chs[w] = (ch < nchunks && !(ch >= skip0 && ch < skip1)) ? ch : -1;
					sx[w] = sy[w] = sz[w] = 0.0;
					if (chs[w] < 0) continue;
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
					sx[w] = ax; sy[w] = ay; sz[w] = az;
				}
				if (pw) prof[(nblk + blk) * 8 + 2] = clock64();
#pragma unroll
				for (int w = 0; w < 2; w++) {
					const int pos = chs[w] * kGsRows + r;
					if (chs[w] >= 0 && cl == 0 && pos < np) {
						const int i = order[pos];
						atomicAdd(acc + 3 * i, sx[w]); atomicAdd(acc + 3 * i + 1, sy[w]); atomicAdd(acc + 3 * i + 2, sz[w]);
					}
				}
				__threadfence();
				__syncwarp();
				if (lane < 2 && chs[lane] >= 0) st_flag(applied + chs[lane], blk + 1);
				if (pw) prof[(nblk + blk) * 8 + 3] = clock64();
			}
		}
	}
}

// what the cluster shares: the solver's shared-memory words the helpers read (progress, dmu) and write (partial pushes, done)
struct GsShared { int prog; int loaded; int folded; int pad; int done[4]; };

template <bool ORTHO, bool EXPD>
__global__ void __cluster_dims__(kGsCluster, 1, 1) __launch_bounds__(kGsThreads, 1)
k_gs_pipeline(const double4 *__restrict__ gpq, const int *__restrict__ gmeta, const int *__restrict__ order, int np, CellDev c, PolarDev p,
              const double *__restrict__ efs, double *mu, double *efi, double *new_mu, double *acc, double *dmu,
              const double *__restrict__ tri, GsCtl *ctl, long long *prof, volatile int *started, int token) {
	cg::cluster_group cluster = cg::this_cluster();
	if (blockIdx.x == 0 && threadIdx.x == 0) { *started = token; __threadfence_system(); }   // tells the host that the cluster holds its SMs
	extern __shared__ __align__(16) double s_raw[];
	int *applied = (int *)(ctl + 1);
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int cta = blockIdx.x;
	const int nblk = (np + kGsB - 1) / kGsB, nchunks = (np + kGsRows - 1) / kGsRows;
	constexpr int kChunksPerBlk = kGsB / kGsRows;
	// the solver's layout (helpers address the shared part of it through the cluster)
	double *s_mat = s_raw;                                         // [kGsB columns][kGsB rows][6], rolled: column k of the next block replaces column k once the walk has passed it
	double *s_site = s_mat + kGsMat;                               // [kGsSiteCols][kGsB]
	double *s_pendp = s_site + kGsSiteCols * kGsB;                 // [kGsHelpers][kGsAhead * kGsB][3] the helpers' pushes of the last panel
	double *s_pend = s_pendp + kGsHelpers * (kGsAhead * kGsB * 3); // [kGsSlots][kGsB][3] pushes already made into the rows of blocks b .. b+kGsAhead (slot = block % kGsSlots)
	double4 *s_dm = (double4 *)(s_pend + kGsSlots * 3 * kGsB);     // [kGsB] dmu of the block being walked
	int *s_idx = (int *)(s_dm + kGsB);                             // [2][kGsB] site ids of this block / the next block
	GsShared *s_sh = (GsShared *)(s_idx + 2 * kGsB);

	if (cta == 0) {
		// ------------------------------------------------ solver ------------------------------------------------
		volatile int *s_prog = &s_sh->prog;                        // blk * kGsB + columns of the walk that are final
		volatile int *s_done = s_sh->done;                         // per helper: panels delivered
		volatile int *s_folded = &s_sh->folded;                    // panels whose deliveries have been folded (their buffers are free again)
		volatile int *s_loaded = &s_sh->loaded;                    // blocks whose site columns the walker has taken into registers
		auto load_cols = [&](int blk, int m) {                     // site id and site columns of row m (all but the running contraction)
			const int pos = blk * kGsB + m;
			const bool on = pos < np;
			const int s = on ? order[pos] : 0;
			s_idx[(blk & 1) * kGsB + m] = s;
			s_site[m] = on ? gpq[pos].w : 0.0;
			for (int q = 0; q < 3; q++) {
				s_site[(1 + q) * kGsB + m] = on ? __ldcg(mu + 3 * s + q) : 0.0;
				s_site[(4 + q) * kGsB + m] = on ? efs[3 * s + q] : 0.0;
			}
		};
		auto load_acc = [&](int blk, int m) {                      // running contraction of row m as the updaters left it
			const bool on = blk * kGsB + m < np;
			const int s = s_idx[(blk & 1) * kGsB + m];
			for (int q = 0; q < 3; q++) s_site[(7 + q) * kGsB + m] = on ? __ldcg(acc + 3 * s + q) : 0.0;
		};
		for (int q = tid; q < kGsSlots * 3 * kGsB; q += kGsThreads) s_pend[q] = 0.0;
		{
			const double2 *src = (const double2 *)tri;
			double2 *dst = (double2 *)s_mat;
			for (int q = tid; q < kGsMat / 2; q += kGsThreads) __pipeline_memcpy_async(dst + q, src + q, sizeof(double2));
			__pipeline_commit();
		}
		if (tid < kGsB) { load_cols(0, tid); load_acc(0, tid); }
		if (tid == 0) { *s_prog = 0; *s_loaded = 0; *s_folded = 0; for (int h = 0; h < 4; h++) s_done[h] = 0; }
		cluster.sync();                                            // the helpers may look at prog / done from here on
		for (int blk = 0; blk < nblk; blk++) {
			const int base = blk * kGsB, cnt = min(kGsB, np - base);
			if (prof && tid == 0) prof[blk * 8 + 0] = clock64();
			// (A) fold the helpers' pushes of the previous panel into the pending sums (fixed order), then the running contraction
			//     of this block's rows = what the updaters left (fetched during the previous walk) + the cluster's own pushes
			if (blk > 0) {
				if (tid < kGsHelpers) while (s_done[tid] < blk) { }         // every helper has delivered panel blk-1
				__syncthreads();
				asm volatile("fence.acq_rel.cluster;" ::: "memory");
				if (tid < kGsAhead * kGsB) {
					const int j = tid / kGsB, row = tid % kGsB;         // target block blk + j
					double *dst = s_pend + (((blk + j) % kGsSlots) * kGsB + row) * 3;
					for (int q = 0; q < 3; q++) {
						const double v = (s_pendp[(0 * kGsAhead * kGsB + tid) * 3 + q] + s_pendp[(1 * kGsAhead * kGsB + tid) * 3 + q]) + s_pendp[(2 * kGsAhead * kGsB + tid) * 3 + q];
						dst[q] = (j < kGsAhead - 1 ? dst[q] : 0.0) + v;     // the farthest target starts here; the nearer ones already hold earlier panels
					}
				}
			}
			__pipeline_wait_prior(0);
			__syncthreads();
			if (tid == 0) *s_folded = blk;                                 // the helpers may overwrite their delivery buffers (a helper with
			                                                               // no column in a short last block would otherwise run ahead)
			if (tid < kGsB) for (int q = 0; q < 3; q++) s_site[(7 + q) * kGsB + tid] += s_pend[((blk % kGsSlots) * kGsB + tid) * 3 + q];
			__syncthreads();
			if (prof && tid == 0) prof[blk * 8 + 1] = clock64();
			// (B)
			if (warp == 0) {
				// lane owns rows lane (slot 0) and lane+32 (slot 1); everything a row needs lives in registers during the walk.
				// With c = alpha E_s - mu_old the change of a dipole is a single FMA:  dmu = c - alpha acc.
				double al[2], cx[2], cy[2], cz[2], ax[2], ay[2], az[2], ex[2], ey[2], ez[2], sx[2], sy[2], sz[2];
#pragma unroll
				for (int h = 0; h < 2; h++) {
					const int m = lane + 32 * h;
					al[h] = s_site[m];
					sx[h] = s_site[4 * kGsB + m]; sy[h] = s_site[5 * kGsB + m]; sz[h] = s_site[6 * kGsB + m];
					cx[h] = al[h] * sx[h] - s_site[1 * kGsB + m];
					cy[h] = al[h] * sy[h] - s_site[2 * kGsB + m];
					cz[h] = al[h] * sz[h] - s_site[3 * kGsB + m];
					ax[h] = s_site[7 * kGsB + m]; ay[h] = s_site[8 * kGsB + m]; az[h] = s_site[9 * kGsB + m];
					ex[h] = ey[h] = ez[h] = 0;
				}
				__syncwarp();
				if (lane == 0) { *s_prog = base; *s_loaded = blk + 1; }     // the site columns are in registers: warp 1 may refill them
				// tensor entries of column k for my two rows: (xx yy) (zz xy) (xz yz); always one column ahead of the dependent chain.
				// Two register sets used alternately (the loop is unrolled by hand: a copy would cost 24 moves per step on the
				// one warp the whole sweep waits for).
				const double2 *tcol = (const double2 *)s_mat + lane * 3;
				double2 ta[2][3], tb[2][3];
#pragma unroll
				for (int hh = 0; hh < 2; hh++) { ta[hh][0] = tcol[hh * 96]; ta[hh][1] = tcol[hh * 96 + 1]; ta[hh][2] = tcol[hh * 96 + 2]; }
				const unsigned tcol_s = (unsigned)__cvta_generic_to_shared(tcol);
				auto step = [&](const int half, const int kk, double2 (&tc)[2][3], double2 (&tn)[2][3]) {
					const int k = kk + 32 * half;
					// every lane forms the candidate change of its own slot-`half` row; the owner's is the real one
					const double dxc = fma(-al[half], ax[half], cx[half]), dyc = fma(-al[half], ay[half], cy[half]), dzc = fma(-al[half], az[half], cz[half]);
					const double dx = __shfl_sync(0xffffffffu, dxc, kk), dy = __shfl_sync(0xffffffffu, dyc, kk), dz = __shfl_sync(0xffffffffu, dzc, kk);
					if (lane == kk) { ex[half] = ax[half]; ey[half] = ay[half]; ez[half] = az[half]; }   // acc at the moment of the update
					// hand the finished column to the other warps and CTAs.  No fence: both are volatile shared-memory stores of one
					// thread, which the LSU performs in program order (a MEMBAR here costs more than the whole step); every lane
					// stores the same values, so the walk has no divergent region
					volatile double *vd = (volatile double *)(s_dm + k);
					vd[0] = dx; vd[1] = dy; vd[2] = dz;
					*s_prog = base + k + 1;
					// the next column into the register set the previous step has finished with.  Volatile (ordered after the stores
					// above) so that the compiler does not hoist these loads over the previous step's FMAs, which would cost it a third
					// register set and 24 moves per step
					const unsigned tnext = tcol_s + (unsigned)(min(k + 1, kGsB - 1) * (kGsB * 3) * sizeof(double2));
#pragma unroll
					for (int hh = 0; hh < 2; hh++)
#pragma unroll
						for (int q = 0; q < 3; q++)
							asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(tn[hh][q].x), "=d"(tn[hh][q].y) : "r"(tnext + (unsigned)((hh * 96 + q) * sizeof(double2))));
#pragma unroll
					for (int hh = 0; hh < 2; hh++) {                      // the diagonal entry is zero: a row does not move itself
						ax[hh] = fma(tc[hh][0].x, dx, fma(tc[hh][1].y, dy, fma(tc[hh][2].x, dz, ax[hh])));
						ay[hh] = fma(tc[hh][1].y, dx, fma(tc[hh][0].y, dy, fma(tc[hh][2].y, dz, ay[hh])));
						az[hh] = fma(tc[hh][2].x, dx, fma(tc[hh][2].y, dy, fma(tc[hh][1].x, dz, az[hh])));
					}
				};
#pragma unroll
				for (int half = 0; half < 2; half++) {
					const int kend = min(32, cnt - 32 * half);
					int kk = 0;
					for (; kk + 1 < kend; kk += 2) { step(half, kk, ta, tb); step(half, kk + 1, tb, ta); }
					if (kk < kend) step(half, kk, ta, tb);                    // an odd count: always the last step of the block
				}
				if (prof && tid == 0) prof[blk * 8 + 2] = clock64();
				// contract_dipoles: ef_induced = -acc at the moment of the update, mu = alpha (E_s + ef_induced)  (:3583-3592)
#pragma unroll
				for (int hh = 0; hh < 2; hh++) {
					const int m = lane + 32 * hh;
					if (m < cnt) {
						const int s = s_idx[(blk & 1) * kGsB + m];
						const double nx = al[hh] * (sx[hh] - ex[hh]), ny = al[hh] * (sy[hh] - ey[hh]), nz = al[hh] * (sz[hh] - ez[hh]);
						__stcg(mu + 3 * s, nx); __stcg(mu + 3 * s + 1, ny); __stcg(mu + 3 * s + 2, nz);
						new_mu[3 * s] = nx; new_mu[3 * s + 1] = ny; new_mu[3 * s + 2] = nz;
						efi[3 * s] = -ex[hh]; efi[3 * s + 1] = -ey[hh]; efi[3 * s + 2] = -ez[hh];
						__stcg(acc + 3 * s, ax[hh]); __stcg(acc + 3 * s + 1, ay[hh]); __stcg(acc + 3 * s + 2, az[hh]);
					}
				}
				asm volatile("bar.sync 1, 96;" ::: "memory");
			} else if (warp == 1) {
				// the next block: site columns (once the walker has taken its own into registers), tensors rolled in behind the walk
				// (column k is dead once column k+1 has been fetched), running contraction once the updaters have delivered
				if (blk + 1 < nblk) {
					while (*s_loaded <= blk) __nanosleep(100);
					load_cols(blk + 1, lane); load_cols(blk + 1, lane + 32);
					// (no updater touches these rows between panel blk-kGsAhead and my own write-back: safe to fetch now)
					const int c0 = (base + kGsB) / kGsRows, c1 = min(nchunks, c0 + kChunksPerBlk);
					if (blk >= kGsAhead && lane < c1 - c0) {
						int spins = 0;
						while (ld_flag(applied + c0 + lane) < blk + 1 - kGsAhead && !ld_flag(&ctl->abort)) {
							__nanosleep(100);
							if (++spins > kGsWaitLimit) st_flag(&ctl->abort, 1);
						}
					}
					__syncwarp();
					__threadfence();
					load_acc(blk + 1, lane); load_acc(blk + 1, lane + 32);
					const double2 *src = (const double2 *)(tri + (size_t)(blk + 1) * kGsMat);
					double2 *dst = (double2 *)s_mat;
					int done = 0;
					while (done < kGsB && !(g_gs_debug & 4)) {
						int pg = max(*s_prog - base, 0);
						if (pg >= cnt) pg = kGsB;                               // the walk is over: the remaining (unused) columns too
						if (pg <= done) { __nanosleep(200); continue; }
						for (int q = done * (kGsB * 3) + lane; q < pg * (kGsB * 3); q += 32) __pipeline_memcpy_async(dst + q, src + q, sizeof(double2));
						done = pg;
					}
					__pipeline_commit();
					__pipeline_wait_prior(0);
				}
				asm volatile("bar.sync 1, 96;" ::: "memory");                   // walker + this warp + the publisher have what the next block needs
			} else if (warp == 2) {
				// publish the panel for the updaters: the change of every dipole of the block, then the flag
				while (*s_prog < base + cnt) __nanosleep(100);
				double d[2][3];
#pragma unroll
				for (int h = 0; h < 2; h++) {
					const volatile double *vd = (const volatile double *)(s_dm + min(lane + 32 * h, kGsB - 1));
					d[h][0] = vd[0]; d[h][1] = vd[1]; d[h][2] = vd[2];
				}
				asm volatile("bar.sync 1, 96;" ::: "memory");                   // s_dm may be overwritten by the next walk from here on
#pragma unroll
				for (int h = 0; h < 2; h++) {
					const int k = lane + 32 * h;
					if (k < cnt) { __stcg(dmu + 3 * (base + k), d[h][0]); __stcg(dmu + 3 * (base + k) + 1, d[h][1]); __stcg(dmu + 3 * (base + k) + 2, d[h][2]); }
				}
				__threadfence();
				__syncwarp();
				if (lane == 0) st_flag(&ctl->solved, blk + 1);
			}
			if (prof && tid == 0) prof[blk * 8 + 3] = clock64();
		}
		__syncthreads();
		cluster.sync();                                            // the helpers are done with my shared memory
	} else if (cta < kGsCluster) {
		// ------------------------------------------------ helpers -----------------------------------------------
		// helper hj takes the columns k = hj mod 3 of every panel, for all rows of the next kGsAhead blocks
		const int hj = cta - 1;
		double4 *h_rows = (double4 *)s_raw;                        // [kGsSlots][kGsB] x y z alpha, slot = block % kGsSlots
		int *h_meta = (int *)(h_rows + kGsSlots * kGsB);           // [kGsSlots][kGsB]
		double *h_part = (double *)(h_meta + kGsSlots * kGsB);     // [4 column sub-slices][kGsAhead * kGsB][3]
		const volatile int *r_prog = &cluster.map_shared_rank(s_sh, 0)->prog;
		int *r_done = cluster.map_shared_rank(s_sh, 0)->done + hj;
		const volatile int *r_folded = &cluster.map_shared_rank(s_sh, 0)->folded;
		const volatile double *r_dm = (const volatile double *)cluster.map_shared_rank(s_dm, 0);
		double *r_pendp = cluster.map_shared_rank(s_pendp, 0) + hj * (kGsAhead * kGsB * 3);
		auto load_rows = [&](int blk) {
			if (tid < kGsB) {
				const int pos = blk * kGsB + tid;
				const bool on = pos < np;
				h_rows[(blk % kGsSlots) * kGsB + tid] = on ? gpq[pos] : make_double4(0, 0, 0, 0);
				h_meta[(blk % kGsSlots) * kGsB + tid] = on ? gmeta[pos] : 0;
			}
		};
		for (int b = 0; b < kGsAhead; b++) load_rows(b);
		cluster.sync();
		const int r = tid & (kGsB - 1), s4 = tid >> 6;            // my row of each target block; my column sub-slice
		for (int blk = 0; blk < nblk; blk++) {
			const int base = blk * kGsB, cnt = min(kGsB, np - base);
			load_rows(blk + kGsAhead);
			__syncthreads();
			double4 pr[kGsAhead]; int mr[kGsAhead]; bool on[kGsAhead];
			double ax[kGsAhead], ay[kGsAhead], az[kGsAhead];
#pragma unroll
			for (int j = 0; j < kGsAhead; j++) {
				const int tb = blk + 1 + j;
				pr[j] = h_rows[(tb % kGsSlots) * kGsB + r];
				mr[j] = h_meta[(tb % kGsSlots) * kGsB + r];
				on[j] = tb < nblk && tb * kGsB + r < np;
				ax[j] = ay[j] = az[j] = 0.0;
			}
			for (int k = hj + kGsHelpers * s4; k < cnt; k += kGsHelpers * 4) {
				if (lane == 0) {
					// a column takes the walk ~180 cycles (~95 ns): sleep about as long as the columns still to come need, so that
					// 24 helper warps do not keep reading the solver's shared memory while it walks
					int pg;
					while ((pg = *r_prog) < base + k + 1) __nanosleep(min(2000, 40 + 80 * (base + k - pg)));
				}
				__syncwarp();
				const double4 dm = make_double4(r_dm[4 * k], r_dm[4 * k + 1], r_dm[4 * k + 2], 0.0);
				const double4 pc = h_rows[(blk % kGsSlots) * kGsB + k];
				const int mc = h_meta[(blk % kGsSlots) * kGsB + k];
#pragma unroll
				for (int j = 0; j < kGsAhead; j++)
					if (on[j]) gs_contract<ORTHO, EXPD>(c, p, pr[j], mr[j], pc, mc, dm, ax[j], ay[j], az[j]);
			}
#pragma unroll
			for (int j = 0; j < kGsAhead; j++) {
				double *o = h_part + ((s4 * kGsAhead + j) * kGsB + r) * 3;
				o[0] = ax[j]; o[1] = ay[j]; o[2] = az[j];
			}
			// my previous delivery must have been folded before its buffer is written again (true by construction whenever this
			// helper had a column to wait for; not in a last block shorter than the number of helpers)
			if (tid == 0) while (*r_folded < blk) __nanosleep(100);
			__syncthreads();
			// sub-slices summed in a fixed order, straight into the solver's shared memory
			if (tid < kGsAhead * kGsB) {
				for (int q = 0; q < 3; q++)
					r_pendp[tid * 3 + q] = (h_part[((0 * kGsAhead * kGsB) + tid) * 3 + q] + h_part[((1 * kGsAhead * kGsB) + tid) * 3 + q]) +
					                       (h_part[((2 * kGsAhead * kGsB) + tid) * 3 + q] + h_part[((3 * kGsAhead * kGsB) + tid) * 3 + q]);
			}
			asm volatile("fence.acq_rel.cluster;" ::: "memory");
			__syncthreads();
			if (tid == 0) *(volatile int *)r_done = blk + 1;
			// the solver folds these sums before it starts the next walk, and only then publishes columns of the next panel:
			// r_pendp is free again by the time this helper writes it
		}
		cluster.sync();
	} else {
		// single-launch fallback (MPMC_GS_FUSED=1: for tools that serialise kernel launches, e.g. ncu): the CTAs beyond the
		// cluster do the updaters' work, one CTA per SM
		gs_updater_body<ORTHO, EXPD>(s_raw, cta - kGsCluster, gridDim.x - kGsCluster, gpq, gmeta, order, np, c, p, acc, dmu, ctl, prof);
	}
}

// The updaters as their own kernel (2 CTAs per SM on every SM the solver's cluster leaves free — the solver needs a whole SM's shared
// memory, the updaters need latency hiding), launched right after the solver kernel on a second stream.
template <bool ORTHO, bool EXPD>
__global__ void __launch_bounds__(kGsThreads, 2)
k_gs_updaters(const double4 *__restrict__ gpq, const int *__restrict__ gmeta, const int *__restrict__ order, int np, CellDev c, PolarDev p,
              double *acc, const double *dmu, GsCtl *ctl, long long *prof) {
	extern __shared__ __align__(16) double s_raw[];
	gs_updater_body<ORTHO, EXPD>(s_raw, blockIdx.x, gridDim.x, gpq, gmeta, order, np, c, p, acc, dmu, ctl, prof);
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
