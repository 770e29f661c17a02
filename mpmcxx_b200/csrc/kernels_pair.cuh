// kernels_pair.cuh — what the sweeps share: the per-item partial record and its fixed-order reduction, the path-integral
// assembly of the per-bead sums (alone, or fused with the cross-GPU sum over peer memory), the (row block x column lane)
// constants of the ordered sweeps.  The sweeps themselves live in kernels_pair2.cuh / kernels_polar2.cuh / kernels_gs.cuh.
#pragma once
#include "device_math.cuh"

namespace mpmc {


// site metadata word: molecule index in the low 31 bits, frozen flag in the sign bit
__device__ __forceinline__ int  meta_mol(int m)    { return m & 0x7fffffff; }
__device__ __forceinline__ bool meta_frozen(int m) { return m < 0; }

// per-CTA partial sums of the energy sweep
struct PairPartial { double rd, es_real, es_intra, n_in; };

// Sum the per-item partials of one bead in a fixed order (strided per thread, then a shared-memory tree) into the per-bead
// result record: res[bead*kResStride + 0..3] = rd_pair, es_real, es_intra, n_in.
constexpr int kResStride = 8;   // + 4 es_recip, 5 sum mu.E_s, 6 sum mu.dE_ind, 7 sum rrms
__global__ void __launch_bounds__(256)
k_reduce_partials(const PairPartial *__restrict__ partials, int nitems, double *__restrict__ res, int *__restrict__ item_ctr, int ctr_start) {
	__shared__ double s_red[4][256];
	const int bead = blockIdx.x, tid = threadIdx.x;
	if (bead == 0 && tid == 0) *item_ctr = ctr_start;      // the sweep's item counter, ready for the next launch
	double a = 0, b = 0, c = 0, d = 0;
	for (int t = tid; t < nitems; t += 256) {
		const PairPartial p = partials[(size_t)bead * nitems + t];
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
// A peer that does not show up within `timeout_cycles` (default ~60 s: rank-0 I/O, a topology rebuild or a debugger can hold a rank
// for seconds) makes the exchange FAIL ON EVERY RANK: the waiting rank raises its sticky error word (sums[5], copied back with the
// sums: the host call returns MPMC_ERR_CUDA) and poisons its slot in every peer's mailbox (seq = -1, both parities), so that a
// late peer — which may still complete the step it is in from the stamps already posted — fails at its next exchange instead of
// carrying on with a partner that has stopped.  No rank ever returns success with sums another rank does not hold.
constexpr long long kPiPoison = -1;
__device__ __forceinline__ void pi_exchange(const double *s_loc, double (*s_in)[4], double *__restrict__ sums, PiMailSlot *const *__restrict__ peers,
                                            int rank, int nranks, long long step, long long timeout_cycles) {
	const int t = threadIdx.x;
	const int par = (int)(step & 1);
	__shared__ int s_fail;
	if (t == 0) s_fail = 0;
	__syncthreads();
	if (t < nranks) {
		volatile PiMailSlot *dst = peers[t] + par * nranks + rank;
		for (int q = 0; q < 4; q++) dst->v[q] = s_loc[q];
		__threadfence_system();
		dst->seq = step;
		volatile PiMailSlot *src = peers[rank] + par * nranks + t;
		const long long t0 = clock64();
		bool ok = true;
		for (;;) {
			const long long sq = src->seq;
			if (sq == step) break;
			if (sq == kPiPoison || clock64() - t0 > timeout_cycles) { ok = false; break; }
		}
		__threadfence_system();
		for (int q = 0; q < 4; q++) s_in[t][q] = ok ? src->v[q] : __longlong_as_double(0x7ff8000000000000ll);
		if (!ok) s_fail = 1;
	}
	__syncthreads();
	if (s_fail && t < nranks) {
		for (int pp = 0; pp < 2; pp++) { volatile PiMailSlot *dst = peers[t] + pp * nranks + rank; dst->seq = kPiPoison; }
		__threadfence_system();
	}
	if (t < 4) {
		double a = 0;
		for (int r = 0; r < nranks; r++) a += s_in[r][t];
		sums[t] = a;
	}
	if (t == 0 && s_fail) sums[5] = 1.0;       // sticky: only the host clears it
}

__global__ void k_pi_sums_xchg(const double *__restrict__ res, int nbeads, double rd_const, double es_self, int es_on, int polar_on, int palmo,
                               double *__restrict__ sums, PiMailSlot *const *__restrict__ peers, int rank, int nranks, long long *step_counter,
                               long long timeout_cycles) {
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
	pi_exchange(s_loc, s_in, sums, peers, rank, nranks, s_step, timeout_cycles);
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

__global__ void k_fill_u64(unsigned long long *p, int n, unsigned long long v) {
	for (int i = threadIdx.x; i < n; i += blockDim.x) p[i] = v;
}

} // namespace mpmc
