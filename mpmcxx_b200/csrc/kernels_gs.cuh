// kernels_gs.cuh — Gauss-Seidel dipole sweep, mathematically sequential in the sweep order, as a software pipeline.
//
// contract_dipoles() with polar_gs / polar_gs_ranked (reference src/System.Energy.cpp:3570-3595) overwrites mu_i as soon
// as it is computed, so site i sees the NEW dipoles of every site swept before it and the OLD dipoles of the rest: a dense
// triangular solve with N sequential steps.  The engine keeps, for every polarizable site, the running contraction
//       acc_i = sum_{j != i} T_ij mu_j(current)
// so that a site's update is  mu_i = alpha_i (E_s,i - acc_i),  and the change  dmu_i = mu_i(new) - mu_i(old)  is pushed
// into every other row:  acc_m += T_mi dmu_i.  Sites are processed in blocks of kGsB = 64 in sweep order ("panels").
// Everything that depends on geometry and sweep order only is computed ONCE per energy() and order, in parallel over blocks,
// and reused by all sweeps: the inverse of each block's triangular system (k_gs_inverse) and the tensors between a block and
// the kGsAhead blocks that follow it (k_gs_near).  One sweep is then:
//   * the SOLVER CTA (rank 0 of an 8-CTA cluster) owns the critical path.  Per block: fold what the cluster pushed into the
//     block's rows, form the right-hand side, one 192 x 192 triangular matrix-vector product (21 tiles of 32 x 32 over 14 warps,
//     the matrix in shared memory), write mu and the rows' running contraction back, publish the panel of dipole changes.
//     Dedicated warps fetch the next block's site data during the walk, send the panel to the helpers, fence and flag it for the
//     updaters, and issue the bulk copies of the next inverse, which is awaited piece by piece.
//   * seven HELPER CTAs share the cluster.  They push each panel into the rows of the next kGsAhead = 4 blocks — the rows
//     the solver needs before any updater could deliver: tensors prefetched into registers before the panel exists, each helper's
//     columns of the panel stored into its shared memory by the solver (st.async that signals the helper's barrier), 45 FMAs per
//     thread; a helper keeps what it has pushed into the blocks ahead in its own shared memory and returns only the sums of the
//     block solved next (one bulk copy that signals the solver's barrier).  Nothing is polled or fenced across the cluster.
//   * the UPDATER CTAs (every other SM, one CTA of 16 warps per SM) push each published panel into all remaining rows, including
//     the panel's own block — 4 rows per warp, 8 column lanes per row, tensors on the fly, a row's sum in registers from panel to
//     panel — in panel order, and flag each 4-row chunk when the solver is about to need it.  An updater has kGsAhead block
//     periods to deliver.
//   * the solver may start block b when its rows have received panels 0..b-1-kGsAhead from the updaters; the later panels
//     are the cluster's own pushes.
// Every row receives its updates in a fixed order, so the result does not depend on timing.  Pushing panels into rows that
// were already swept prepares acc for the next sweep, and after the last sweep acc_i is exactly the contraction
// palmo_contraction() needs (:3602-3627), so Palmo costs no extra sweep.  ef_induced follows from mu after the sweep (k_gs_efi).
// One launch = one sweep; every CTA must be resident (the grid is sized from cudaOccupancyMaxActiveClusters), which makes the
// flag waits safe.  What was tried on the way is in profiles/r01c_gs_pipeline.md and profiles/r02_gs_pipeline.md.
#pragma once
#include <cuda_pipeline.h>
#include "kernels_polar2.cuh"

namespace mpmc {

constexpr int kGsB = 64;                  // sites per solver block
constexpr int kGsRows = 4;                // rows per updater chunk (one warp: 4 rows x 8 column lanes)
constexpr int kGsColLanes = 32 / kGsRows;
constexpr int kGsUpdWarps = 16, kGsUpdThreads = 32 * kGsUpdWarps;   // ONE updater CTA of 16 warps per SM (four per scheduler): 140 SMs x 16 = 2240
                                          // warps for the 2300 four-row chunks of 9200 polarizable sites; the 60 surplus chunks are
                                          // split by columns over four warps of one CTA, one per scheduler (gs_updater_body)
constexpr int kGsThreads = 256;
constexpr int kGsWarps = kGsThreads / 32;
constexpr int kGsPipeThreads = 512;       // the solver/helper cluster: 192 x 2 threads walk, 512 per helper push
constexpr int kGsN = 3 * kGsB;            // components per block
constexpr int kGsInv = 9 * (kGsB * (kGsB - 1) / 2);   // 18144 doubles: one block's strictly lower inverse (k_gs_inverse)
__host__ __device__ constexpr int gs_inv_rows(int j) { return kGsN - 3 * (j + 1); }
__host__ __device__ constexpr int gs_inv_off(int j) { return 3 * (kGsN - 3) * j - 9 * (j * (j - 1) / 2); }
// The solver reads a block's inverse as 32 x 32 tiles: tile (I, J), J <= I, of the 192 x 192 strictly (by site) lower matrix X sits at
// gs_tile(I, J) * 1024 + c * 32 + r  (column-major inside the tile: a warp reads one column with consecutive lanes); the entries of the
// diagonal tiles that are not strictly lower by site are stored as zeros.
constexpr int kGsTileDim = 32, kGsTilesPerSide = kGsN / kGsTileDim, kGsTiles = kGsTilesPerSide * (kGsTilesPerSide + 1) / 2;   // 6, 21
__host__ __device__ constexpr int gs_tile(int I, int J) { return I * (I + 1) / 2 + J; }
constexpr int kGsMat = kGsTiles * kGsTileDim * kGsTileDim;   // 21504 doubles of the solver's matrix buffer (172 KB)

// shared memory (doubles).  Updaters: per warp the panel's columns and dmu.
constexpr int kGsSiteCols = 10;           // 0 alpha, 1-3 mu_old, 4-6 E_static, 7-9 acc
constexpr int kGsAhead = 4;               // the cluster pushes a panel into this many following blocks itself; the updaters take the rest
constexpr int kGsHelpers = 7, kGsCluster = 1 + kGsHelpers;
// Solver: the block's inverse, two site-column buffers, the helpers' deliveries for the block about to be solved, panel dmu, right-hand
// side, per-tile partial products, site ids, barriers.  Helpers: the panel, a ring of pending sums for the next kGsAhead blocks, the two
// column slices' partial sums, a barrier.  (Offsets in doubles from the start of dynamic shared memory; the same in every CTA, so a
// CTA can form the address of a peer's buffer from its own.)
constexpr int kGsHelpDm = 0;                                   // double4[kGsB]: the panel's dipole changes (written by the solver)
constexpr int kGsHelpAcc = kGsHelpDm + 4 * kGsB;               // [kGsAhead][kGsN]: what this helper has pushed into the next blocks' rows so far
constexpr int kGsHelpPart = kGsHelpAcc + kGsAhead * kGsN;      // [2 column slices][kGsAhead * kGsB][3]
constexpr int kGsHelpBar = kGsHelpPart + 2 * kGsAhead * kGsB * 3;   // one mbarrier (8 bytes)
constexpr int kGsSolIn = kGsMat + 2 * kGsSiteCols * kGsB;      // [kGsHelpers][kGsN]: the helpers' deliveries (written by the helpers)
constexpr int kGsSolBar = kGsSolIn + kGsHelpers * kGsN;        // two mbarriers: deliveries, inverse
constexpr size_t kGsSolverDoubles = (size_t)kGsSolBar + 2 + 4 * kGsB + kGsN + kGsTiles * kGsTileDim + 2 * kGsB + 16;
static_assert(kGsSolverDoubles >= (size_t)kGsHelpBar + 1, "helpers use the same allocation");
static_assert(kGsSolverDoubles >= (size_t)(kGsPipeThreads / 32) * 8 * kGsB, "the fused fallback runs the updaters inside the pipeline kernel");
constexpr size_t kGsSmemBytes = sizeof(double) * kGsSolverDoubles;
constexpr int kGsRing = 4;                // panels an updater CTA keeps in shared memory
constexpr size_t kGsUpdaterDoubles = (size_t)kGsRing * 8 * kGsB;
constexpr size_t kGsUpdaterSmemBytes = sizeof(double) * kGsUpdaterDoubles;

__device__ int g_gs_debug = 0;            // developer switch (mpmc_debug_gs_profile enable bits 1..): 2 = pushers idle, 4 = no rolling copy, 8 = surplus chunks not split
struct GsCtl { int solved; int abort; int pad[30]; };   // followed in memory by int applied[nchunks]
// The flags are never cleared between sweeps: every launch gets a generation number and counts from `gbase` = generation << 16
// (solved = gbase + blocks solved, applied[chunk] = gbase + panels in memory), so whatever the previous sweep left compares as "not
// yet".  The host zeroes the words when the generation wraps.  `abort` is sticky until the host has looked at it.
constexpr int kGsGenShift = 16;
// `abort`: set when a CTA has waited ~2 s for the other kernel of the pipeline — the two kernels were not run side by side (a
// tool that serialises launches).  Everybody then stops waiting, the sweep's result is meaningless and the host reports it.
constexpr int kGsWaitLimit = 20000000;

__device__ __forceinline__ long long gtime() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
// clock read that cannot run ahead of a preceding barrier (BAR.SYNC.DEFER_BLOCKING lets register-only instructions pass): it
// depends on a shared-memory load issued after the barrier
__device__ __forceinline__ long long clock_after(const volatile int *p) { const int x = *p; long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) : "r"(x)); return t; }
__device__ __forceinline__ int ld_acquire(const int *p) { int v; asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ int ld_flag(const int *p) { return *(const volatile int *)p; }
__device__ __forceinline__ void st_flag(int *p, int v) { *(volatile int *)p = v; }
// mbarrier + bulk asynchronous copy (TMA engine, no tensor map): one thread moves a block's inverse into shared memory with a few
// instructions; the waiters sleep on the barrier's phase instead of issuing 40 cp.async each
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ unsigned map_to_cta(unsigned addr, unsigned rank) { unsigned r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r; }
// bulk copy from my shared memory into a peer's, completion signalled on the PEER's barrier (one wide transfer instead of hundreds of
// 8-byte stores: the cluster port handles ~1 small store per 1.5 cycles, which made 7 x 192 doubles take 1 us)
__device__ __forceinline__ void bulk_s2peer(unsigned remote_dst, unsigned local_src, unsigned bytes, unsigned remote_bar) {
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // my generic-proxy stores to the source are visible to the copy engine
	asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(remote_dst), "r"(local_src), "r"(bytes), "r"(remote_bar) : "memory");
}
// a 16-byte store into a peer's shared memory that signals the peer's mbarrier when it lands (no flag, no fence, no polling over the
// cluster network: the reader sleeps on its own barrier).  Measured alternatives for the solver's 2 KB panel: plain peer stores +
// `mbarrier.arrive.release.cluster` (the arrival takes 1.3 us), one bulk copy per helper (640 ns to land), 8 bytes per store (twice
// the stores; the cluster port takes ~1 small store per 3-6 cycles)
__device__ __forceinline__ void st_async_v2(unsigned remote_addr, double a, double b, unsigned remote_bar) {
	asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];" ::"r"(remote_addr), "l"(__double_as_longlong(a)),
	             "l"(__double_as_longlong(b)), "r"(remote_bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
	unsigned ok;
	do {
		asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
	} while (!ok);
}

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

// The walk of a block is a forward substitution: with a = the running contraction of the block's rows when the block starts,
//       (I + diag(alpha) T_L) dmu = alpha E_s - mu_old - alpha a,        T_L = the strictly lower (earlier-site) part of the block's T,
// 64 sequential steps if done site by site (the first version of this pipeline: ~160 cycles per site on the one warp the whole sweep
// waits for).  The matrix depends on the geometry and the sweep order only — not on the dipoles — so its inverse is computed ONCE
// per energy() and sweep order, for all blocks in parallel (k_gs_inverse), and the walk becomes one 192 x 192 triangular
// matrix-vector product spread over 192 threads.  Layout of the strictly lower inverse X (unit diagonal and the identity 3x3
// diagonal blocks are implied):  inv[blk][gs_inv_off(j) + q * gs_inv_rows(j) + (r - 3 (j + 1))] = X[r][3 j + q]  for column site j,
// column component q, row component r >= 3 (j + 1): thread r of the walk reads unit-stride.
constexpr size_t kGsInverseSmemBytes = sizeof(double) * (kGsInv + 6 * kGsB) + sizeof(double4) * kGsB + sizeof(int) * kGsB;

template <bool ORTHO>
__global__ void __launch_bounds__(kGsThreads)
k_gs_inverse(const double4 *__restrict__ gpq, const int *__restrict__ gmeta, int np, CellDev c, PolarDev p, double *__restrict__ inv) {
	extern __shared__ __align__(16) double s_w[];                 // X by rows: row component r holds its 3 (r / 3) entries at row_off(r)
	double *s_L = s_w + kGsInv;                                    // alpha_i T_ij of the row being eliminated: [j][xx yy zz xy xz yz]
	double4 *s_pq = (double4 *)(s_L + 6 * kGsB);
	int *s_met = (int *)(s_pq + kGsB);
	const int blk = blockIdx.x, base = blk * kGsB, cnt = min(kGsB, np - base), tid = threadIdx.x;
	if (tid < kGsB) { s_pq[tid] = tid < cnt ? gpq[base + tid] : make_double4(0, 0, 0, 0); s_met[tid] = tid < cnt ? gmeta[base + tid] : 0; }
	__syncthreads();
	auto row_off = [](int r) { const int i = r / 3; return 9 * (i * (i - 1) / 2) + (r - 3 * i) * 3 * i; };
	// phase 1: alpha_i T_ij for every pair j < i, parked at the top of the buffer; row i lives at t_off(i) and is taken out before
	// X's rows (growing from the bottom) reach it:  9 i (i + 1) / 2  <=  t_off(i + 1)  for every i
	auto t_off = [](int i) { return kGsInv - 3 * (kGsB * (kGsB - 1) - i * (i - 1)); };
	for (int q = tid; q < kGsB * kGsB; q += kGsThreads) {
		const int i = q / kGsB, j = q % kGsB;
		if (j >= i) continue;
		double xx = 0, xy = 0, xz = 0, yy = 0, yz = 0, zz = 0, t0 = 0, t1 = 0;
		if (i < cnt) {
			// the tensor itself = the contraction applied to the three unit dipoles (columns of T)
			if (p.damp_type == 2) {
				gs_contract<ORTHO, true>(c, p, s_pq[i], s_met[i], s_pq[j], s_met[j], make_double4(1, 0, 0, 0), xx, xy, xz);
				gs_contract<ORTHO, true>(c, p, s_pq[i], s_met[i], s_pq[j], s_met[j], make_double4(0, 1, 0, 0), t0, yy, yz);
				gs_contract<ORTHO, true>(c, p, s_pq[i], s_met[i], s_pq[j], s_met[j], make_double4(0, 0, 1, 0), t0, t1, zz);
			} else {
				gs_contract<ORTHO, false>(c, p, s_pq[i], s_met[i], s_pq[j], s_met[j], make_double4(1, 0, 0, 0), xx, xy, xz);
				gs_contract<ORTHO, false>(c, p, s_pq[i], s_met[i], s_pq[j], s_met[j], make_double4(0, 1, 0, 0), t0, yy, yz);
				gs_contract<ORTHO, false>(c, p, s_pq[i], s_met[i], s_pq[j], s_met[j], make_double4(0, 0, 1, 0), t0, t1, zz);
			}
		}
		const double al = s_pq[i].w;
		double *t = s_w + t_off(i) + 6 * j;
		t[0] = al * xx; t[1] = al * yy; t[2] = al * zz; t[3] = al * xy; t[4] = al * xz; t[5] = al * yz;
	}
	__syncthreads();
	// phase 2: forward substitution by row sites.  Thread cc owns column component cc = 3 jc + q of X:
	//   X[i][cc] = - sum_{jc <= j < i} L_ij X[j][cc],   X[jc][cc] = e_q
	for (int i = 1; i < kGsB; i++) {
		for (int q = tid; q < 6 * i; q += kGsThreads) s_L[q] = s_w[t_off(i) + q];
		__syncthreads();
		if (tid < 3 * i) {
			const int jc = tid / 3, q = tid - 3 * jc;
			const double *l = s_L + 6 * jc;
			// - (column q of L_i,jc):  L = [xx xy xz; xy yy yz; xz yz zz]
			double x0 = q == 0 ? -l[0] : q == 1 ? -l[3] : -l[4];
			double x1 = q == 0 ? -l[3] : q == 1 ? -l[1] : -l[5];
			double x2 = q == 0 ? -l[4] : q == 1 ? -l[5] : -l[2];
			// (unrolled so that the shared-memory loads of four steps are in flight together: with 6 warps on the SM nothing else
			// hides their latency; the order of the operations of each sum is unchanged)
			const double *w = s_w + row_off(3 * (jc + 1)) + tid;          // rows 3j, 3j+1, 3j+2 of site j are 3j doubles apart
#pragma unroll 4
			for (int j = jc + 1; j < i; j++) {
				const double v0 = w[0], v1 = w[3 * j], v2 = w[6 * j];
				const double2 *m = reinterpret_cast<const double2 *>(s_L + 6 * j);   // (xx yy) (zz xy) (xz yz)
				const double2 m01 = m[0], m23 = m[1], m45 = m[2];
				x0 = fma(-m01.x, v0, fma(-m23.y, v1, fma(-m45.x, v2, x0)));
				x1 = fma(-m23.y, v0, fma(-m01.y, v1, fma(-m45.y, v2, x1)));
				x2 = fma(-m45.x, v0, fma(-m45.y, v1, fma(-m23.x, v2, x2)));
				w += 9 * j;
			}
			s_w[row_off(3 * i) + tid] = x0; s_w[row_off(3 * i + 1) + tid] = x1; s_w[row_off(3 * i + 2) + tid] = x2;
		}
		__syncthreads();
	}
	// the solver's layout: 32 x 32 tiles, zeros where the matrix is not strictly lower by site
	double *out = inv + (size_t)blk * kGsMat;
	for (int q = tid; q < kGsMat; q += kGsThreads) {
		const int t = q / (kGsTileDim * kGsTileDim), cl = (q / kGsTileDim) % kGsTileDim, rl = q % kGsTileDim;
		int I = 0;
		while (gs_tile(I + 1, 0) <= t) I++;
		const int J = t - gs_tile(I, 0), r = kGsTileDim * I + rl, cc = kGsTileDim * J + cl;
		out[q] = cc < 3 * (r / 3) ? s_w[row_off(r) + cc] : 0.0;
	}
}

// The cluster's own pushes (a panel into the rows of the next kGsAhead blocks) need the tensors between every site and the
// kGsAhead * 64 sites that follow its block in the sweep order.  Like the inverse they depend on geometry and order only, so they
// are computed once per energy() and order and then only READ by the helpers in every sweep (48 bytes and 9 FMAs per pair instead
// of ~70 FP64 instructions):  near[((blk * 64 + k) * kGsAhead + j) * 64 + r][6] = xx yy zz xy xz yz of T(row r of block blk+1+j, column k of block blk).
constexpr size_t kGsNearPerBlock = (size_t)kGsB * kGsAhead * kGsB * 6;
template <bool ORTHO>
__global__ void __launch_bounds__(kGsPipeThreads)
k_gs_near(const double4 *__restrict__ gpq, const int *__restrict__ gmeta, int np, CellDev c, PolarDev p, double *__restrict__ near) {
	__shared__ double4 s_col[kGsB], s_row[kGsAhead * kGsB];
	__shared__ int s_cm[kGsB], s_rm[kGsAhead * kGsB];
	const int blk = blockIdx.x, tid = threadIdx.x;
	if (tid < kGsB) { const int pos = blk * kGsB + tid; s_col[tid] = pos < np ? gpq[pos] : make_double4(0, 0, 0, 0); s_cm[tid] = pos < np ? gmeta[pos] : -1; }
	if (tid < kGsAhead * kGsB) { const int pos = (blk + 1) * kGsB + tid; s_row[tid] = pos < np ? gpq[pos] : make_double4(0, 0, 0, 0); s_rm[tid] = pos < np ? gmeta[pos] : -1; }
	__syncthreads();
	double *out = near + (size_t)blk * kGsNearPerBlock;
	for (int q = tid; q < kGsB * kGsAhead * kGsB; q += kGsPipeThreads) {
		const int k = q / (kGsAhead * kGsB), jr = q % (kGsAhead * kGsB);          // jr = j * 64 + r
		double xx = 0, xy = 0, xz = 0, yy = 0, yz = 0, zz = 0, t0 = 0, t1 = 0;
		if (blk * kGsB + k < np && (blk + 1) * kGsB + jr < np) {
			if (p.damp_type == 2) {
				gs_contract<ORTHO, true>(c, p, s_row[jr], s_rm[jr], s_col[k], s_cm[k], make_double4(1, 0, 0, 0), xx, xy, xz);
				gs_contract<ORTHO, true>(c, p, s_row[jr], s_rm[jr], s_col[k], s_cm[k], make_double4(0, 1, 0, 0), t0, yy, yz);
				gs_contract<ORTHO, true>(c, p, s_row[jr], s_rm[jr], s_col[k], s_cm[k], make_double4(0, 0, 1, 0), t0, t1, zz);
			} else {
				gs_contract<ORTHO, false>(c, p, s_row[jr], s_rm[jr], s_col[k], s_cm[k], make_double4(1, 0, 0, 0), xx, xy, xz);
				gs_contract<ORTHO, false>(c, p, s_row[jr], s_rm[jr], s_col[k], s_cm[k], make_double4(0, 1, 0, 0), t0, yy, yz);
				gs_contract<ORTHO, false>(c, p, s_row[jr], s_rm[jr], s_col[k], s_cm[k], make_double4(0, 0, 1, 0), t0, t1, zz);
			}
		}
		double2 *o = reinterpret_cast<double2 *>(out + (size_t)q * 6);
		o[0] = make_double2(xx, yy); o[1] = make_double2(zz, xy); o[2] = make_double2(xz, yz);
	}
}

// The updaters' work, shared by the updater kernel and by the single-launch fallback of the pipeline kernel: CTA `cta` of `U`.
// Warps work independently.  A warp owns up to kGsOwn chunks of kGsRows consecutive rows of the sweep order (lane = row x column
// lane) and keeps what it pushes into them IN REGISTERS from panel to panel — a row's sum only has to reach memory twice per sweep:
//   * when the solver is about to take the row over: the rows of block c belong to the cluster for panels c-kGsAhead .. c-1, so
//     after panel c-kGsAhead-1 the chunk's sums are reduced over the column lanes, added to acc, fenced and flagged
//     (applied[chunk] = number of panels that have reached memory);
//   * at the end of the sweep, with everything pushed since the solver wrote the row back (panels c .. nblk-1, which prepare acc
//     for the next sweep / are the Palmo contraction).
// (The first version added every panel to acc with atomics and paid a fence + flag per panel and chunk: 2.2 k of the 7.9 k
// cycles a warp spent per panel, and a latency the solver ran into several times per sweep.)  Each row has one writer and
// receives its panels in a fixed order, so the result does not depend on timing.  Chunks beyond kGsOwn per warp (systems with
// more than kGsOwn x kGsRows x resident-warps polarizable sites) take the per-panel path of the first version.
constexpr int kGsOwn = 2;
template <bool ORTHO, bool EXPD>
__device__ __forceinline__ void gs_updater_body(double *s_raw, int cta, int U, const double4 *__restrict__ gpq, const int *__restrict__ gmeta,
                                                const int *__restrict__ order, int np, const CellDev &c, const PolarDev &p, double *acc,
                                                const double *dmu, GsCtl *ctl, long long *prof, int gbase) {
	int *applied = (int *)(ctl + 1);
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int nblk = (np + kGsB - 1) / kGsB, nchunks = (np + kGsRows - 1) / kGsRows;
	constexpr int kChunksPerBlk = kGsB / kGsRows;
	// A panel (64 columns: position + molecule word, dipole change + alpha) is fetched ONCE per CTA into a ring of kGsRing shared
	// buffers and read from there by all its warps.  (With a private copy per warp, 2 400 warps read the same 4 KB from L2 at the same
	// moment: the few L2 lines that hold a panel became a hot spot and a panel took 1.5 us to arrive.)  Whichever warp first needs a
	// panel that is not in the ring takes a token, waits for the solver's flag (the one acquire the CTA relies on), waits until every
	// warp has finished with the buffer's previous panel, loads, and posts it; the others meanwhile work on what is already there.
	double4 *ring = (double4 *)s_raw;                               // [kGsRing][2][kGsB]
	__shared__ int s_ready, s_token, s_done[kGsRing];               // panels in the ring so far; loader token; warps finished with a buffer's panel
	__shared__ double s_split[8][2][4][kGsRows * 3];                // split chunks: [group of four warps][flush][warp of the group][row][xyz]
	__shared__ int s_split_cnt[8][2];
	// (warp-major numbering: when there are more chunks than warps, the second chunks go one to each CTA instead of eight to a
	// few CTAs — an SM with twice the work falls behind by a few thousand cycles per panel and stalls the solver at the end)
	const int GW = U * (int)(blockDim.x >> 5), gwid = warp * U + cta;
	const int r = lane & (kGsRows - 1), cl = lane / kGsRows;   // row of the chunk, column lane
	int nlive = 0;                                                  // warps of this CTA that own rows
	for (int w = 0; w < (int)(blockDim.x >> 5); w++) nlive += w * U + cta < nchunks;
	if (tid == 0) { s_ready = 0; s_token = 0; }
	if (tid < kGsRing) s_done[tid] = nlive;                         // every buffer starts free
	if (tid < 16) s_split_cnt[tid >> 1][tid & 1] = 0;
	__syncthreads();
	if (gwid >= nchunks) return;
	// More chunks than warps.  With one chunk too many on an SM one scheduler would carry five warps' worth of contraction against four
	// on the others and set the pace of the whole sweep (config 4: 2300 chunks for 2240 warps).  Up to one surplus chunk per four warps
	// of a CTA is therefore SPLIT BY COLUMNS over those four warps — consecutive warps sit on the four schedulers — each taking a
	// quarter of every panel's columns on top of its own chunk; their partial sums meet in shared memory, in a fixed order, when the
	// chunk is written out.  Beyond that (larger systems) a warp simply owns a second whole chunk.
	const int nwarp = (int)(blockDim.x >> 5), surplus = nchunks - GW;
	const bool split = !(g_gs_debug & 8) && surplus > 0 && surplus <= U * (nwarp / 4) && nwarp <= 32;   // (developer switch 8: never split)
	const int grp = warp >> 2, gq = warp & 3;
	constexpr int kIters = kGsB / kGsColLanes;                      // column iterations of a whole panel (8)
	int nflush = 0;
	// my rows (constant over the sweep) and their sums
	double4 pr[kGsOwn];
	int mr[kGsOwn], chs[kGsOwn], cblk[kGsOwn];
	double sx[kGsOwn], sy[kGsOwn], sz[kGsOwn];
#pragma unroll
	for (int w = 0; w < kGsOwn; w++) {
		int ch = gwid + w * GW;
		if (split && w > 0) ch = (w == 1 && grp < nwarp / 4 && grp * U + cta < surplus) ? GW + grp * U + cta : nchunks;
		chs[w] = ch < nchunks ? ch : -1;
		cblk[w] = ch / kChunksPerBlk;
		const int pos = ch * kGsRows + r;
		const bool on = chs[w] >= 0 && pos < np;
		pr[w] = on ? gpq[pos] : make_double4(0, 0, 0, 0);
		mr[w] = on ? gmeta[pos] : 0;
		if (!on && chs[w] >= 0) mr[w] = -1;                 // a row past the end of a last, partial chunk
		sx[w] = sy[w] = sz[w] = 0.0;
	}
	const bool more = !split && gwid + kGsOwn * GW < nchunks;
	const int it0 = split ? gq * (kIters / 4) : 0, it1 = split ? (gq + 1) * (kIters / 4) : kIters;   // slot 1's share of a panel's column iterations
	// reduce a chunk's sums over the column lanes and add them to acc (one writer per row: a plain read-modify-write)
	auto flush = [&](int w) {
		double ax = sx[w], ay = sy[w], az = sz[w];
		ax += __shfl_xor_sync(0xffffffffu, ax, 4); ay += __shfl_xor_sync(0xffffffffu, ay, 4); az += __shfl_xor_sync(0xffffffffu, az, 4);
		ax += __shfl_xor_sync(0xffffffffu, ax, 8); ay += __shfl_xor_sync(0xffffffffu, ay, 8); az += __shfl_xor_sync(0xffffffffu, az, 8);
		ax += __shfl_xor_sync(0xffffffffu, ax, 16); ay += __shfl_xor_sync(0xffffffffu, ay, 16); az += __shfl_xor_sync(0xffffffffu, az, 16);
		const int pos = chs[w] * kGsRows + r;
		if (split && w == 1) {
			// the four warps' column quarters, added in the order of the quarters by the group's first warp
			double *mine = s_split[grp][nflush & 1][gq] + 3 * r;
			if (cl == 0) { mine[0] = ax; mine[1] = ay; mine[2] = az; }
			__threadfence_block();
			__syncwarp();
			if (gq != 0) { if (lane == 0) atomicAdd(&s_split_cnt[grp][nflush & 1], 1); }
			else {
				if (lane == 0) {
					int spins = 0;
					while (*(volatile int *)&s_split_cnt[grp][nflush & 1] < 3 && !ld_flag(&ctl->abort)) {
						__nanosleep(20);
						if (++spins > 20 * kGsWaitLimit) st_flag(&ctl->abort, 1);
					}
				}
				__syncwarp();
				__threadfence_block();
				if (cl == 0 && pos < np) {
					const double *q0 = s_split[grp][nflush & 1][0] + 3 * r;
					double tx = q0[0], ty = q0[1], tz = q0[2];
#pragma unroll
					for (int q = 1; q < 4; q++) { tx += q0[q * kGsRows * 3]; ty += q0[q * kGsRows * 3 + 1]; tz += q0[q * kGsRows * 3 + 2]; }
					double *a = acc + 3 * order[pos];
					__stcg(a, __ldcg(a) + tx); __stcg(a + 1, __ldcg(a + 1) + ty); __stcg(a + 2, __ldcg(a + 2) + tz);
				}
			}
			nflush++;
		} else if (cl == 0 && pos < np) {
			double *a = acc + 3 * order[pos];
			__stcg(a, __ldcg(a) + ax); __stcg(a + 1, __ldcg(a + 1) + ay); __stcg(a + 2, __ldcg(a + 2) + az);
		}
		sx[w] = sy[w] = sz[w] = 0.0;
	};
	volatile int *v_ready = &s_ready, *v_token = &s_token, *v_done = s_done;
	for (int blk = 0; blk < nblk; blk++) {
		// the rows of the next kGsAhead blocks belong to the cluster (the panel's own rows do not: the solver writes them back
		// before it publishes the panel, without the panel's own contribution)
		const int skip0 = (blk + 1) * kChunksPerBlk, skip1 = min(blk + 1 + kGsAhead, nblk) * kChunksPerBlk;
		const bool pw = prof && cta == 0 && warp == 0 && lane == 0;
		if (pw) prof[(nblk + blk) * 8 + 0] = gtime();
		// ---- get panel blk into the ring (warp-uniform control flow: lane 0 decides, everybody follows)
		for (;;) {
			const int rd = __shfl_sync(0xffffffffu, *v_ready, 0);
			if (rd > blk) break;
			int got = 0;
			if (lane == 0) got = atomicCAS(&s_token, 0, 1) == 0;
			got = __shfl_sync(0xffffffffu, got, 0);
			if (!got) { __nanosleep(20); continue; }
			const int q = *v_ready;                                 // the next panel nobody has loaded yet (q <= blk)
			if (q > blk) { __syncwarp(); if (lane == 0) *v_token = 0; continue; }
			int bad = 0;
			if (lane == 0) {
				int spins = 0;
				while (ld_acquire(&ctl->solved) <= gbase + q && !(bad = ld_flag(&ctl->abort))) {
					__nanosleep(40);
					if (++spins > 20 * kGsWaitLimit) st_flag(&ctl->abort, 1);
				}
				while (!bad && v_done[q % kGsRing] < nlive) {        // the buffer's previous panel (q - kGsRing) is still being read
					__nanosleep(20);
					if (++spins > 20 * kGsWaitLimit) { st_flag(&ctl->abort, 1); bad = 1; }
				}
			}
			bad = __shfl_sync(0xffffffffu, bad, 0);
			if (!bad) {
				double4 *b_col = ring + (q % kGsRing) * 2 * kGsB, *b_dm = b_col + kGsB;
#pragma unroll
				for (int h = 0; h < 2; h++) {
					const int cc = lane + 32 * h, pos = q * kGsB + cc;
					const bool on = pos < np;
					const double4 g = on ? gpq[pos] : make_double4(0, 0, 0, 0);
					const int gm = on ? gmeta[pos] : 0;
					double d0 = 0, d1 = 0, d2 = 0;                         // beyond the end: zero change, contributes nothing
					if (on) { d0 = __ldcg(dmu + 3 * pos); d1 = __ldcg(dmu + 3 * pos + 1); d2 = __ldcg(dmu + 3 * pos + 2); }
					b_col[cc] = make_double4(g.x, g.y, g.z, __longlong_as_double((long long)gm));
					b_dm[cc] = make_double4(d0, d1, d2, g.w);               // .w: alpha of the column (linear damping)
				}
				__threadfence_block();
				__syncwarp();
				if (lane == 0) { v_done[q % kGsRing] = 0; __threadfence_block(); *v_ready = q + 1; }
			} else if (lane == 0) *v_ready = 0x7ffffff0;             // abort: everybody out
			__syncwarp();
			if (lane == 0) { __threadfence_block(); *v_token = 0; }
		}
		if (__shfl_sync(0xffffffffu, ld_flag(&ctl->abort), 0)) return;
		__threadfence_block();                                      // the loader's stores to the buffer, posted before `ready`
		if (pw) prof[(nblk + blk) * 8 + 1] = gtime();
		const double4 *w_col = ring + (blk % kGsRing) * 2 * kGsB, *w_dm = w_col + kGsB;
#pragma unroll
		for (int w = 0; w < kGsOwn; w++) {
			if (chs[w] < 0 || (chs[w] >= skip0 && chs[w] < skip1)) continue;       // warp-uniform
			if (mr[w] != -1) {
				double ax = 0, ay = 0, az = 0;
				// slot 0 always sweeps the whole panel (compile-time bounds); slot 1 its share [it0, it1) of the column iterations
				const int ib = w == 0 ? 0 : it0, ie = w == 0 ? kIters : it1;
				if (EXPD) {
#pragma unroll
					for (int i = 0; i < kIters; i += 2) {                // no bounds on the columns: those past the end carry a zero change
						if (w != 0 && (i < ib || i >= ie)) continue;
						const int cc = cl + kGsColLanes * i;
						const double4 p0 = w_col[cc], p1 = w_col[cc + kGsColLanes];
						const double4 m0 = w_dm[cc], m1 = w_dm[cc + kGsColLanes];
						tensor_contract_exp_x2<ORTHO>(c, p.damp, p.u_damp, pr[w].x, pr[w].y, pr[w].z, p0.x, p0.y, p0.z, m0.x, m0.y, m0.z,
						                              p1.x, p1.y, p1.z, m1.x, m1.y, m1.z, ax, ay, az);
					}
				} else {
#pragma unroll
					for (int i = 0; i < kIters; i++) {
						if (w != 0 && (i < ib || i >= ie)) continue;
						const int cc = cl + kGsColLanes * i;
						double4 pc = w_col[cc];
						const double4 dm = w_dm[cc];
						const int mc = __double2loint(pc.w);
						pc.w = dm.w;
						gs_contract<ORTHO, EXPD>(c, p, pr[w], mr[w], pc, mc, dm, ax, ay, az);
					}
				}
				sx[w] += ax; sy[w] += ay; sz[w] += az;
			}
			// the last panel before the cluster takes these rows over: to memory, then the flag
			if (blk == cblk[w] - kGsAhead - 1) {
				flush(w);
				if (!(split && w == 1) || gq == 0) {                        // (a split chunk is written out by its group's first warp)
					__threadfence();
					__syncwarp();
					if (lane == 0) st_flag(applied + chs[w], gbase + blk + 1);
					if (prof && lane == 0) prof[nblk * 16 + chs[w]] = gtime();      // when this chunk's rows were handed to the cluster
				}
			}
		}
		if (pw) prof[(nblk + blk) * 8 + 2] = gtime();
		if (more) {
			// per-panel path for the chunks this warp cannot keep in registers
			for (int ch = gwid + kGsOwn * GW; ch < nchunks; ch += GW) {
				if (ch >= skip0 && ch < skip1) continue;
				const int pos = ch * kGsRows + r;
				const bool on = pos < np;
				const double4 prr = on ? gpq[pos] : make_double4(0, 0, 0, 0);
				const int mrr = on ? gmeta[pos] : 0;
				double ax = 0, ay = 0, az = 0;
				if (on) {
#pragma unroll 4
					for (int cc = cl; cc < kGsB; cc += kGsColLanes) {
						double4 pc = w_col[cc];
						const double4 dm = w_dm[cc];
						const int mc = __double2loint(pc.w);
						if (!EXPD) pc.w = dm.w;
						gs_contract<ORTHO, EXPD>(c, p, prr, mrr, pc, mc, dm, ax, ay, az);
					}
				}
				ax += __shfl_xor_sync(0xffffffffu, ax, 4); ay += __shfl_xor_sync(0xffffffffu, ay, 4); az += __shfl_xor_sync(0xffffffffu, az, 4);
				ax += __shfl_xor_sync(0xffffffffu, ax, 8); ay += __shfl_xor_sync(0xffffffffu, ay, 8); az += __shfl_xor_sync(0xffffffffu, az, 8);
				ax += __shfl_xor_sync(0xffffffffu, ax, 16); ay += __shfl_xor_sync(0xffffffffu, ay, 16); az += __shfl_xor_sync(0xffffffffu, az, 16);
				if (cl == 0 && on) {
					const int i = order[pos];
					atomicAdd(acc + 3 * i, ax); atomicAdd(acc + 3 * i + 1, ay); atomicAdd(acc + 3 * i + 2, az);
				}
				__threadfence();
				__syncwarp();
				if (lane == 0) st_flag(applied + ch, gbase + blk + 1);
			}
		}
		// this warp is finished with the panel's buffer
		__syncwarp();
		if (lane == 0) { __threadfence_block(); atomicAdd(&s_done[blk % kGsRing], 1); }
	}
	// what was pushed since the solver wrote the rows back
#pragma unroll
	for (int w = 0; w < kGsOwn; w++) if (chs[w] >= 0) flush(w);
}

// columns of a 64-site panel a helper pushes: 2 hj + cs + 14 m  (cs = 0, 1 the thread's column slice)
__host__ __device__ constexpr int gs_helper_cols(int hj) { int n = 0; for (int k = 0; k < kGsB; k++) n += (k % (2 * kGsHelpers)) / 2 == hj; return n; }
// where helper hj's columns sit in the solver's outgoing panel (double4 units): the helpers' columns one after the other, each
// helper's in the order (m, slice) = the order its threads walk them, so that ONE bulk copy per helper delivers them
__host__ __device__ constexpr int gs_helper_first(int hj) { int n = 0; for (int h = 0; h < hj; h++) n += gs_helper_cols(h); return n; }

template <bool ORTHO, bool EXPD>
__global__ void __cluster_dims__(kGsCluster, 1, 1) __launch_bounds__(kGsPipeThreads, 1)
k_gs_pipeline(const double4 *__restrict__ gpq, const int *__restrict__ gmeta, const int *__restrict__ order, int np, CellDev c, PolarDev p,
              const double *__restrict__ efs, double *mu, double *efi, double *new_mu, double *acc, double *dmu,
              const double *__restrict__ tri, const double *__restrict__ near, GsCtl *ctl, long long *prof, int gbase) {
	cg::cluster_group cluster = cg::this_cluster();
	// Programmatic dependent launch: the updater kernel follows in the same stream and may start as soon as every CTA of this grid
	// has executed this — i.e. once the cluster holds its SMs (its CTAs want a whole SM's shared memory each, so the updaters must
	// not be placed first).  No host round trip between the two launches.
	asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
	extern __shared__ __align__(16) double s_raw[];
	int *applied = (int *)(ctl + 1);
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int cta = blockIdx.x;
	const int nblk = (np + kGsB - 1) / kGsB, nchunks = (np + kGsRows - 1) / kGsRows;
	// How the cluster talks (the first version had the helpers poll the solver's progress word, copy the whole panel out of its shared
	// memory and leave their sums for the solver to pull: 57 KB per block through the solver's cluster port, which moves ~17 bytes
	// per cycle — 3.3 k of the 8.6 k cycles a block took):
	//   solver -> helper : each dipole change goes to the ONE helper that owns its column, as an 8-byte store that signals the helper's
	//                      barrier (1.5 KB per block);
	//   helper -> solver : a helper keeps what it has pushed into the next four blocks in its own shared memory and delivers only the
	//                      sums of the block that is solved next, the same way (7 x 1.5 KB per block).
	// Readers sleep on their own barrier; nothing is polled or pulled across the cluster.
	if (cta == 0) {
		// ------------------------------------------------ solver ------------------------------------------------
		double *s_mat = s_raw;                                         // [kGsTiles][32 columns][32 rows]: the block's inverse (k_gs_inverse)
		double *s_site0 = s_mat + kGsMat;                              // [2][kGsSiteCols][kGsB], buffer = block & 1
		double *s_in = s_raw + kGsSolIn;                               // [kGsHelpers][kGsN] what the helpers pushed into the rows of the block about to be solved
		unsigned long long *s_bars = (unsigned long long *)(s_raw + kGsSolBar);
		double4 *s_dm = (double4 *)(s_raw + kGsSolBar + 2);            // [kGsB] the outgoing panel, grouped by owning helper (gs_helper_first)
		double *s_rho = (double *)(s_dm + kGsB);                       // [kGsN] right-hand side of the block being walked
		double *s_tpart = s_rho + kGsN;                                // [kGsTiles][32] per-tile partial dot products
		int *s_idx = (int *)(s_tpart + kGsTiles * kGsTileDim);         // [2][kGsB] site ids of this block / the next block
		volatile int *s_clk = s_idx;                                   // any shared word: anchors the profiling clock reads behind the barriers
		constexpr int kPublisherWarp = 12, kLoaderWarp = 13;
		constexpr int kProductWarps = kGsPipeThreads / 32 - 2;       // every warp but the publisher and the loader
		// named barriers: 1 = the product warps among themselves (right-hand side complete, product complete), 2 = top of the block
		// (all but the publisher), 3 = end of the block (the product warps + the publisher), 4 = the outgoing panel is in shared memory
		// (the row threads arrive, the publisher waits)
		auto load_cols = [&](int blk, int m) {                     // site id and site columns of row m (all but the running contraction)
			const int pos = blk * kGsB + m;
			const bool on = pos < np;
			const int s = on ? order[pos] : 0;
			double *s_site = s_site0 + (blk & 1) * kGsSiteCols * kGsB;
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
			double *s_site = s_site0 + (blk & 1) * kGsSiteCols * kGsB;
			for (int q = 0; q < 3; q++) s_site[(7 + q) * kGsB + m] = on ? __ldcg(acc + 3 * s + q) : 0.0;
		};
		// the inverse arrives in three pieces of seven tiles, each on its own barrier: the product starts on the tiles of a piece as soon
		// as that piece is there (172 KB take ~3 k cycles through this SM's port; waiting for all of it cost ~0.7 k cycles per block)
		unsigned long long *s_bars_m = (unsigned long long *)(s_raw + kGsSolverDoubles - 4);
		const unsigned bar_in = smem_u32(&s_bars[0]);
		const unsigned bar_m[3] = {smem_u32(&s_bars[1]), smem_u32(&s_bars_m[0]), smem_u32(&s_bars_m[1])};
		constexpr int kPieceTiles = kGsTiles / 3;
		static_assert(kPieceTiles * 3 == kGsTiles, "three equal pieces");
		constexpr unsigned kMatBytes = sizeof(double) * kGsMat, kMatPiece = kMatBytes / 3;
		static_assert(kMatPiece % 16 == 0, "bulk copies move multiples of 16 bytes");
		auto fetch_inverse = [&](int blk) {                        // one thread: the whole inverse of block blk -> s_mat, completion on the pieces' barriers
			const char *src = (const char *)(tri + (size_t)blk * kGsMat);
			for (int q = 0; q < 3; q++) {
				mbar_expect_tx(bar_m[q], kMatPiece);
				bulk_g2s(smem_u32(s_mat) + q * kMatPiece, src + q * kMatPiece, kMatPiece, bar_m[q]);
			}
		};
		// my component's place in the outgoing panel: column site wm belongs to helper (wm % 14) / 2, as its column 2 (wm / 14) + (wm % 2)
		const int wm = tid / 3, wq = tid - 3 * wm;                     // my site of the block and component (threads 0..191)
		int out_at = 0;
		if (tid < kGsN) {
			const int w14 = wm % (2 * kGsHelpers), owner = w14 / 2;
			out_at = 4 * (gs_helper_first(owner) + 2 * (wm / (2 * kGsHelpers)) + (w14 & 1)) + wq;
		}
		// the publisher warp's lanes 0..6 deliver: one bulk copy each, shared -> the helper's panel buffer, signalling the helper's barrier
		// (192 eight-byte st.async took the row warps ~650 cycles per block to issue)
		// Warp 15 delivers the panel: 2 KB as four 16-byte st.async per lane into the helpers' panel buffers, each signalling the owning
		// helper's barrier.  (192 eight-byte st.async from the row threads reached the helpers ~1.3 k cycles after the product and kept
		// the row warps busy for ~650 of them; a bulk copy per helper took 640 ns to land; plain peer stores followed by a
		// release.cluster arrival took 1.3 us.)
		unsigned r_unit[4] = {0, 0, 0, 0}, r_ubar[4] = {0, 0, 0, 0};
		constexpr int kSenderWarp = 15;                             // a product-only warp with no global stores of its own in flight
		if (warp == kSenderWarp) {
#pragma unroll
			for (int k = 0; k < 4; k++) {
				const int u = lane + 32 * k, d4 = u >> 1;
				int h = 0;
				while (h + 1 < kGsHelpers && gs_helper_first(h + 1) <= d4) h++;
				r_unit[k] = map_to_cta(smem_u32(s_raw + kGsHelpDm) + (unsigned)((d4 - gs_helper_first(h)) * 32 + (u & 1) * 16), 1 + h);
				r_ubar[k] = map_to_cta(smem_u32(s_raw + kGsHelpBar), 1 + h);
			}
		}
		if (tid < kGsB) s_dm[tid] = make_double4(0.0, 0.0, 0.0, 0.0);
		if (tid == 0) {
			mbar_init(bar_in, 1);
			for (int q = 0; q < 3; q++) mbar_init(bar_m[q], 1);
			asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		}
		if (tid < kGsB) { load_cols(0, tid); load_acc(0, tid); }
		cluster.sync();                                            // every barrier of the cluster is initialised
		if (tid == 0) fetch_inverse(0);
		for (int blk = 0; blk < nblk; blk++) {
			const int base = blk * kGsB, cnt = min(kGsB, np - base);
			double *s_site = s_site0 + (blk & 1) * kGsSiteCols * kGsB;
			if (prof && tid == 0) prof[blk * 8 + 0] = clock_after(s_clk);
			// (A) what the cluster pushed into this block's rows (panels blk-kGsAhead .. blk-1), delivered by the helpers after panel blk-1
			// (the publisher warp comes straight from its memory fence — several hundred cycles after the others — and owns no row:
			// it joins at the barrier before the matrix-vector product, not here)
			if (warp != kPublisherWarp) {
				if (blk > 0 && warp != kLoaderWarp) {
					if (tid == 0) mbar_expect_tx(bar_in, (unsigned)(sizeof(double) * kGsHelpers * kGsN));
					mbar_wait(bar_in, (blk - 1) & 1);
					if (prof && tid == 0) prof[blk * 8 + 3] = gtime();
				}
				asm volatile("bar.sync 2, %0;" :: "n"(kGsPipeThreads - 32) : "memory");   // the loader warp's columns of this block are in place
			}
			if (prof && tid == 0) { prof[blk * 8 + 4] = clock_after(s_clk); prof[blk * 8 + 7] = gtime(); }
			if (warp == kLoaderWarp) {
				// the NEXT block's site columns and running contraction, fetched while this block is walked (nothing of it depends
				// on this walk: the rows' dipoles are still the old ones, and no updater touches these rows between panel
				// blk-kGsAhead and the solver's own write-back).  This warp skips the walk's barriers and waits at the next block's.
				if (blk + 1 < nblk) {
					load_cols(blk + 1, lane); load_cols(blk + 1, lane + 32);
					// every lane acquires the flags of the two chunks its rows (lane, lane + 32) belong to: no fence — a MEMBAR in this
					// warp stalls the shared-memory traffic of the walk next to it
					const int c0 = (base + kGsB) / kGsRows;
					const long long tw0 = prof ? clock64() : 0;
					long long twait = 0; int late = -1;
					if (blk >= kGsAhead) {
#pragma unroll
						for (int h = 0; h < 2; h++) {
							const int ch = c0 + (lane + 32 * h) / kGsRows;
							if (ch >= nchunks) continue;
							int spins = 0;
							while (ld_acquire(applied + ch) < gbase + blk + 1 - kGsAhead && !ld_flag(&ctl->abort)) {
								__nanosleep(100);
								if (++spins > kGsWaitLimit) st_flag(&ctl->abort, 1);
							}
							if (prof && spins) { twait = clock64() - tw0; late = ch; }
						}
					}
					if (prof) {                                        // how long this block's rows kept the loader waiting, and for which chunk
						long long key = (twait << 16) | (late & 0xffff);
#pragma unroll
						for (int o = 16; o > 0; o >>= 1) { const long long k2 = __shfl_xor_sync(0xffffffffu, key, o); key = k2 > key ? k2 : key; }
						if (lane == 0) prof[blk * 8 + 5] = key;
					}
					__syncwarp();
					load_acc(blk + 1, lane); load_acc(blk + 1, lane + 32);
				}
				continue;
			}
			if (warp == kPublisherWarp) {
				// the publisher: waits for the block's write-back, sends the panel to the helpers (they may be waiting for it already), then
				// publishes it for the updaters: the change of every dipole of the block, then the flag — after the write-back of the
				// block's rows: the updaters add this panel to those rows too.  (The end-of-block barrier makes the row threads' stores part
				// of what this fence orders.)  It takes no part in the product: its fence lasts ~1 us, and while it shared the product's
				// barriers every block waited ~0.5 k cycles for it.
				asm volatile("bar.sync 3, %0;" :: "n"((kProductWarps + 1) * 32) : "memory");
				__threadfence();
				__syncwarp();
				if (lane == 0) st_flag(&ctl->solved, gbase + blk + 1);
				continue;
			}
			// (B) the walk: rhs = alpha E_s - mu_old - alpha acc, dmu = rhs + X rhs (k_gs_inverse): a 192 x 192 triangular matrix-vector
			// product cut in 21 tiles of 32 x 32, one or two tiles per warp, lane = row; every load has a compile-time offset (the
			// first version walked the packed triangle with ~100 instructions of address arithmetic per 12 FMAs: 3 000 cycles)
			double rhs_r = 0.0, mo_r = 0.0, a_r = 0.0;
			if (tid < kGsN) {
				const double al_r = s_site[wm], es_r = s_site[(4 + wq) * kGsB + wm];
				mo_r = s_site[(1 + wq) * kGsB + wm]; a_r = s_site[(7 + wq) * kGsB + wm];
				if (blk > 0) {                                             // the helpers' sums in a fixed order
					double v = s_in[tid];
#pragma unroll
					for (int h = 1; h < kGsHelpers; h++) v += s_in[h * kGsN + tid];
					a_r += v;
				}
				rhs_r = fma(-al_r, a_r, fma(al_r, es_r, -mo_r));
				s_rho[tid] = rhs_r;
			}
			asm volatile("bar.sync 1, %0;" :: "n"(kProductWarps * 32) : "memory");   // the right-hand side is complete
			if (prof && tid == 0) prof[blk * 8 + 1] = clock_after(s_clk);
			{
				const int wslot = warp < kPublisherWarp ? warp : warp - 2;   // 14 warps take part: seven of them take two tiles
				for (int t = wslot; t < kGsTiles; t += kProductWarps) {
					int I = 0;
					while (gs_tile(I + 1, 0) <= t) I++;
					const int J = t - gs_tile(I, 0);
					mbar_wait(bar_m[t / kPieceTiles], blk & 1);                // the tile's piece has landed (every reader observes the phase itself)
					const double *xt = s_mat + t * (kGsTileDim * kGsTileDim) + lane;
					const double2 *rh = reinterpret_cast<const double2 *>(s_rho + kGsTileDim * J);
					double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
#pragma unroll
					for (int cc = 0; cc < kGsTileDim; cc += 4) {
						const double2 r01 = rh[cc / 2], r23 = rh[cc / 2 + 1];
						d0 = fma(xt[cc * kGsTileDim], r01.x, d0);
						d1 = fma(xt[(cc + 1) * kGsTileDim], r01.y, d1);
						d2 = fma(xt[(cc + 2) * kGsTileDim], r23.x, d2);
						d3 = fma(xt[(cc + 3) * kGsTileDim], r23.y, d3);
					}
					s_tpart[t * kGsTileDim + lane] = (d0 + d1) + (d2 + d3);
				}
				if (prof && tid == 191) prof[blk * 8 + 6] = clock64();
			}
			asm volatile("bar.sync 1, %0;" :: "n"(kProductWarps * 32) : "memory");
			if (warp == kSenderWarp) {
				// the panel to the helpers, as soon as the row threads have put it into s_dm (this warp has signed the end-of-block barrier
				// off already: see the fetching warp below)
				asm volatile("bar.arrive 3, %0;" :: "n"((kProductWarps + 1) * 32) : "memory");
				asm volatile("bar.sync 4, %0;" :: "n"(kGsN + 32) : "memory");
				if (blk + 1 < nblk) {
#pragma unroll
					for (int k = 0; k < 4; k++) {
						const double2 v = reinterpret_cast<const double2 *>(s_dm)[lane + 32 * k];
						st_async_v2(r_unit[k], v.x, v.y, r_ubar[k]);
					}
				}
				__syncwarp();
				continue;
			}
			if (warp == kLoaderWarp + 1) {
				// s_mat is free: every warp is past the product.  This warp owns no row: it signs the end-of-block barrier off at once and
				// issues the three bulk copies behind it, so that nobody waits for their issue (the next barrier it meets is the next
				// block's `bar.sync 2`, which the others reach only after this block's barrier has completed)
				asm volatile("bar.arrive 3, %0;" :: "n"((kProductWarps + 1) * 32) : "memory");
				if (lane == 0 && blk + 1 < nblk) fetch_inverse(blk + 1);
				__syncwarp();
				continue;
			}
			if (tid < kGsN) {
				double d = rhs_r;                                              // + (X rhs)_row: the row block's tiles in a fixed order
				const int I = tid / kGsTileDim;
				for (int J = 0; J <= I; J++) d += s_tpart[gs_tile(I, J) * kGsTileDim + (tid & (kGsTileDim - 1))];
				reinterpret_cast<double *>(s_dm)[out_at] = d;              // the panel for the helpers, grouped by owner
				asm volatile("bar.arrive 4, %0;" :: "n"(kGsN + 32) : "memory");      // ... which the publisher warp sends off while the rows are written back
				if (wm < cnt) __stcg(dmu + 3 * (base + wm) + wq, d);               // for the updaters (the publisher warp fences and raises the flag)
				// contract_dipoles: mu = alpha (E_s + ef_induced), ef_induced = -acc at the moment of the update  (:3583-3592):
				// mu = mu_old + dmu; ef_induced is recovered from mu after the sweep (k_gs_efi)
				if (wm < cnt) {
					const int sidx = s_idx[(blk & 1) * kGsB + wm];
					const double nm = mo_r + d;
					__stcg(mu + 3 * sidx + wq, nm);
					new_mu[3 * sidx + wq] = nm;
					__stcg(acc + 3 * sidx + wq, a_r);                      // what the cluster knows of this row; the updaters add the panel itself
				}
			}
			asm volatile("bar.sync 3, %0;" :: "n"((kProductWarps + 1) * 32) : "memory");
			if (prof && tid == 0) { prof[blk * 8 + 2] = clock_after(s_clk); prof[(nblk + blk) * 8 + 6] = gtime(); }
		}
		__syncthreads();
		cluster.sync();                                            // nothing of the cluster is in flight towards my shared memory any more
	} else if (cta < kGsCluster) {
		// ------------------------------------------------ helpers -----------------------------------------------
		// a panel = 64 columns x the 64 rows of each of the next kGsAhead blocks.  Thread = (row r, target block j, column slice cs);
		// helper hj's two slices are 2 hj and 2 hj + 1 of 2 kGsHelpers: at most 5 columns per thread, whose tensors (k_gs_near)
		// are fetched into registers BEFORE the panel is ready — when it arrives, 45 FMAs per thread remain.
		const int hj = cta - 1;
		double4 *h_dm = (double4 *)(s_raw + kGsHelpDm);            // [kGsB] the panel's dipole changes (my columns only are written)
		double *h_acc = s_raw + kGsHelpAcc;                        // [kGsAhead][kGsN] pushed so far into the rows of the blocks ahead (slot = block % kGsAhead)
		double *h_part = s_raw + kGsHelpPart;                      // [2 column slices][kGsAhead * kGsB][3]
		const unsigned bar_p = smem_u32(s_raw + kGsHelpBar);
		const unsigned r_in = map_to_cta(smem_u32(s_raw + kGsSolIn + hj * kGsN), 0), r_bar_in = map_to_cta(smem_u32(s_raw + kGsSolBar), 0);
		const unsigned panel_bytes = (unsigned)(sizeof(double4) * gs_helper_cols(hj));   // my columns of the panel, in the order (m, slice)
		for (int q = tid; q < kGsAhead * kGsN; q += kGsPipeThreads) h_acc[q] = 0.0;
		if (tid < kGsB) h_dm[tid] = make_double4(0.0, 0.0, 0.0, 0.0);
		if (tid == 0) { mbar_init(bar_p, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
		__syncthreads();
		cluster.sync();
		const int r = tid & (kGsB - 1), j = (tid >> 6) & (kGsAhead - 1), cs = tid >> 8;
		static_assert(kGsAhead == 4 && kGsPipeThreads == 2 * kGsAhead * kGsB, "thread = (row, target block, column slice)");
		constexpr int kCols = (kGsB + 2 * kGsHelpers - 1) / (2 * kGsHelpers);   // 5
		for (int blk = 0; blk + 1 < nblk; blk++) {                 // (the last panel has no rows ahead of it)
			const int base = blk * kGsB, cnt = min(kGsB, np - base);
			const int tb = blk + 1 + j;
			const bool on = tb < nblk && tb * kGsB + r < np;
			double2 t[kCols][3];
#pragma unroll
			for (int m = 0; m < kCols; m++) {
				const int k = 2 * hj + cs + 2 * kGsHelpers * m;
				t[m][0] = t[m][1] = t[m][2] = make_double2(0.0, 0.0);
				if (on && k < cnt) {
					const double2 *src = reinterpret_cast<const double2 *>(near + (size_t)blk * kGsNearPerBlock + ((size_t)(k * kGsAhead + j) * kGsB + r) * 6);
					t[m][0] = __ldg(src); t[m][1] = __ldg(src + 1); t[m][2] = __ldg(src + 2);
				}
			}
			// the tensors of the block after next: into L2 now (they stream from HBM once per sweep), into registers one block later
			if (blk + 2 < nblk) {
				const int tb2 = blk + 2 + j;
				if (tb2 < nblk && tb2 * kGsB + r < np) {
#pragma unroll
					for (int m = 0; m < kCols; m++) {
						const int k = 2 * hj + cs + 2 * kGsHelpers * m;
						if (k < kGsB) asm volatile("prefetch.global.L2 [%0];" ::"l"(near + (size_t)(blk + 1) * kGsNearPerBlock + ((size_t)(k * kGsAhead + j) * kGsB + r) * 6));
					}
				}
			}
			if (tid == 0) mbar_expect_tx(bar_p, panel_bytes);
			mbar_wait(bar_p, blk & 1);                             // my columns of the panel have arrived
			const bool hp = prof && hj == 0 && tid == 0;
			if (hp) prof[(nblk + blk) * 8 + 4] = gtime();
			double ax = 0.0, ay = 0.0, az = 0.0;
#pragma unroll
			for (int m = 0; m < kCols; m++) {
				const double4 d = h_dm[2 * m + cs];                        // (xx yy) (zz xy) (xz yz); zero tensor (and a zero entry) where there is no column
				ax = fma(t[m][0].x, d.x, fma(t[m][1].y, d.y, fma(t[m][2].x, d.z, ax)));
				ay = fma(t[m][1].y, d.x, fma(t[m][0].y, d.y, fma(t[m][2].y, d.z, ay)));
				az = fma(t[m][2].x, d.x, fma(t[m][2].y, d.y, fma(t[m][1].x, d.z, az)));
			}
			{
				double *o = h_part + ((cs * kGsAhead + j) * kGsB + r) * 3;
				o[0] = ax; o[1] = ay; o[2] = az;
			}
			__syncthreads();
			if (hp) prof[(nblk + blk) * 8 + 5] = gtime();
			// the two slices in a fixed order, added to what earlier panels pushed into the same rows; the farthest target block
			// starts its sum here.  What belongs to the block solved next goes to the solver, the rest stays here.
			if (tid < kGsAhead * kGsB) {
				double *slot = h_acc + ((blk + 1 + j) % kGsAhead) * kGsN + 3 * r;
#pragma unroll
				for (int q = 0; q < 3; q++) {
					const double v = h_part[tid * 3 + q] + h_part[(kGsAhead * kGsB + tid) * 3 + q];
					slot[q] = (j < kGsAhead - 1 ? slot[q] : 0.0) + v;
				}
			}
			__syncthreads();                                       // h_part and the panel buffer are free for the next panel; the next block's slot is complete
			if (tid == 0) bulk_s2peer(r_in, smem_u32(h_acc + ((blk + 1) % kGsAhead) * kGsN), (unsigned)(sizeof(double) * kGsN), r_bar_in);
			if (prof && tid == 0) atomicMax((unsigned long long *)&prof[(nblk + blk) * 8 + 7], (unsigned long long)gtime());   // the LAST helper to finish
		}
		cluster.sync();
	} else {
		// single-launch fallback (MPMC_GS_FUSED=1: for tools that serialise kernel launches, e.g. ncu): the CTAs beyond the
		// cluster do the updaters' work, one CTA per SM
		gs_updater_body<ORTHO, EXPD>(s_raw, cta - kGsCluster, gridDim.x - kGsCluster, gpq, gmeta, order, np, c, p, acc, dmu, ctl, prof, gbase);
	}
}

// The updaters as their own kernel (kGsUpdCtas CTAs per SM on every SM the solver's cluster leaves free — the solver needs a whole SM's
// shared memory, the updaters need latency hiding), launched right after the solver kernel in the same stream as its programmatic
// dependent (cudaLaunchAttributeProgrammaticStreamSerialization): it starts when the cluster is resident and never calls
// griddepcontrol.wait — the two kernels synchronise through their own flags.
template <bool ORTHO, bool EXPD>
__global__ void __launch_bounds__(kGsUpdThreads, 1)
k_gs_updaters(const double4 *__restrict__ gpq, const int *__restrict__ gmeta, const int *__restrict__ order, int np, CellDev c, PolarDev p,
              double *acc, const double *dmu, GsCtl *ctl, long long *prof, int gbase) {
	extern __shared__ __align__(16) double s_raw[];
	gs_updater_body<ORTHO, EXPD>(s_raw, blockIdx.x, gridDim.x, gpq, gmeta, order, np, c, p, acc, dmu, ctl, prof, gbase);
}

// ef_induced of the polarizable sites after a Gauss-Seidel sweep: mu = alpha (E_s + ef_induced)  =>  ef_induced = mu / alpha - E_s
__global__ void k_gs_efi(const int *__restrict__ plist, int np, const double *__restrict__ mu, const double *__restrict__ alpha,
                         const double *__restrict__ efs, double *__restrict__ efi) {
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= np * 3) return;
	const int i = plist[t / 3], o = 3 * i + t % 3;
	efi[o] = mu[o] / alpha[i] - efs[o];
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
