// kernels_pi.cuh — the tail of a path-integral sweep as ONE kernel (non-polarizable systems): what k_reduce_partials,
// k_structure_reduce, k_recip_energy and k_pi_sums / k_pi_sums_xchg do in four launches.  At 8 beads per GPU a sweep is ~100 us of
// kernels, and every kernel boundary costs 3-4 us of it.
//   CTA = one bead system:  (1) the pair sweep's per-item partials in a fixed order -> rd_pair, es_real, es_intra, n_in;
//                           (2) S_mobile(k) = sum of the chunk partials in chunk order (stored for later readers), and
//                               coulombic_reciprocal()'s (4 pi / V) sum_k w(k) |S(k)|^2 (src/System.Energy.cpp:1613-1619);
//   the LAST CTA to finish:  (3) PI_calculate_potential's sums over the local beads in bead order
//                               (src/SimulationControl.PathIntegral.cpp:786-796), and, with peers, the cross-GPU exchange.
// Every sum has a fixed shape and order: the result does not depend on which CTA happens to be last.
#pragma once
#include "kernels_pair.cuh"
#include "kernels_recip.cuh"

namespace mpmc {

__global__ void __launch_bounds__(256)
k_pi_finish(const PairPartial *__restrict__ partials, int nitems, int *__restrict__ item_ctr, int ctr_start,
            const double2 *__restrict__ sk_part, int nchunks, const KVec *__restrict__ kv, int nk, double four_pi_over_v, double2 *__restrict__ S_mobile,
            double *res, int nbeads, double rd_const, double es_self, int es_on,
            double *__restrict__ sums, PiMailSlot *const *__restrict__ peers, int rank, int nranks, long long *step_counter, long long timeout_cycles,
            int *__restrict__ done, volatile double *host_out, long long *launch_counter) {
	__shared__ double s_red[4][256];
	__shared__ double s_loc[4];
	__shared__ double s_in[32][4];
	__shared__ long long s_step;
	__shared__ int s_last;
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
	const double v0 = s_red[0][0], v1 = s_red[1][0], v2 = s_red[2][0], v3 = s_red[3][0];
	__syncthreads();
	double e_recip = 0;
	if (es_on) {
		double acc = 0;
		for (int ik = tid; ik < nk; ik += 256) {
			double2 s = make_double2(0.0, 0.0);
			for (int cidx = 0; cidx < nchunks; cidx++) {
				const double2 p = sk_part[((size_t)bead * nchunks + cidx) * nk + ik];
				s.x += p.x; s.y += p.y;
			}
			S_mobile[(size_t)bead * nk + ik] = s;
			acc += kv[ik].w_energy * (s.x * s.x + s.y * s.y);
		}
		s_red[0][tid] = acc;
		__syncthreads();
		for (int o = 128; o > 0; o >>= 1) {
			if (tid < o) s_red[0][tid] += s_red[0][tid + o];
			__syncthreads();
		}
		e_recip = s_red[0][0] * four_pi_over_v;
	}
	if (tid == 0) {
		double *r = res + bead * kResStride;
		r[0] = v0; r[1] = v1; r[2] = v2; r[3] = v3; r[4] = e_recip; r[5] = r[6] = r[7] = 0.0;
		__threadfence();
		const int ticket = atomicAdd(done, 1);
		s_last = ticket == nbeads - 1;
		if (s_last) *done = 0;
	}
	__syncthreads();
	if (!s_last) return;
	__threadfence();
	if (tid == 0) {
		double s0 = 0, s1 = 0;
		for (int bb = 0; bb < nbeads; bb++) {
			const double *r = res + bb * kResStride;
			s0 += __ldcg(r) + rd_const;
			s1 += es_on ? (__ldcg(r + 1) - __ldcg(r + 2)) + __ldcg(r + 4) + es_self : 0.0;
		}
		s_loc[0] = s0; s_loc[1] = s1; s_loc[2] = 0; s_loc[3] = 0;
		if (peers) s_step = ++(*step_counter);
		else { sums[0] = s0; sums[1] = s1; sums[2] = 0; sums[3] = 0; }
	}
	__syncthreads();
	if (peers) pi_exchange(s_loc, s_in, sums, peers, rank, nranks, s_step, timeout_cycles);
	// The result straight into the caller's pinned buffer, then a launch number the host waits for: no copy node after the kernel, no
	// stream synchronisation to wake the host (together ~10-15 us of a 100 us sweep).  host_out[0..5] = sums + error word, [7] = number.
	if (host_out) {
		__syncthreads();
		if (tid == 0) {
			for (int q = 0; q < 6; q++) host_out[q] = __ldcg(sums + q);
			__threadfence_system();
			host_out[7] = (double)(++(*launch_counter));
		}
	}
}

} // namespace mpmc
