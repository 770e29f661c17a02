// kernels_recip.cuh — Ewald reciprocal space: structure factors S(k) (energy and field), reciprocal static field.
//
// The reference evaluates cos/sin(k.r) once per (site, k) (src/System.Energy.cpp:1596-1611, 2868-2880).  Here
// k.r = 2 pi sum_q l_q s_q with s = fractional coordinates (k_p = 2 pi sum_q recip[p][q] l_q, :1586-1590), so
// each site needs only three sincos; e^{i k.r} for every k is a product of three entries of a per-site table of
// powers e^{i m theta_q}, m = 0..kmax, built by complex multiplication.  That turns ~100 FP64 instructions per
// (site, k) into 8 FMAs.  Accumulated rounding after <= kmax multiplications is ~1e-15 relative.
#pragma once
#include "device_math.cuh"

namespace mpmc {

constexpr int kMaxKmax   = 15;            // table rows per axis = kmax+1 <= 16
constexpr int kSkSites   = 64;            // sites per CTA in the structure-factor kernel
constexpr int kSkThreads = 256;
constexpr int kFrSites   = 32;            // sites per CTA in the reciprocal-field kernel
constexpr int kFrLanes   = 4;             // k lanes per site (threads per CTA = kFrSites * kFrLanes)

struct KVec { double kx, ky, kz, w_energy, w_field; int l0, l1, l2, pad; };   // w_field = exp(-k^2/4a_p^2)/k^2

// phases of one site: (cos, sin)(2 pi s_q * m) for m = 0..kmax, q = 0..2, written to tab[(q*(kmax+1)+m)*ld + col]
__device__ __forceinline__ void build_phase_table(const CellDev &c, double x, double y, double z, int kmax,
                                                  double2 *tab, int ld, int col) {
#pragma unroll
	for (int q = 0; q < 3; q++) {
		double s = c.rb[0][q] * x + c.rb[1][q] * y + c.rb[2][q] * z;         // fractional coordinate q
		s -= rint(s);                                                         // exact; e^{2 pi i m s} is periodic in s for integer m
		double sn, cs;
		sincos(2.0 * kPi * s, &sn, &cs);
		double2 e = make_double2(1.0, 0.0);
		tab[(q * (kmax + 1)) * ld + col] = e;
		for (int m = 1; m <= kmax; m++) {
			e = make_double2(e.x * cs - e.y * sn, e.x * sn + e.y * cs);
			tab[(q * (kmax + 1) + m) * ld + col] = e;
		}
	}
}

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 tab_at(const double2 *tab, int kmax, int ld, int col, int q, int l) {
	double2 e = tab[(q * (kmax + 1) + abs(l)) * ld + col];
	if (l < 0) e.y = -e.y;
	return e;
}

// K3: partial structure factors of one chunk of sites for every k:  part[bead][chunk][k] = sum_j q_j e^{i k.r_j}
// over the sites listed in `list` (the mobile charged sites for the per-move sum, or the frozen charged sites for
// the cached framework sum).  Threads run over k; the chunk's phase table lives in shared memory laid out
// [site][row] so that lanes with different l hit different banks.
__global__ void __launch_bounds__(kSkThreads)
k_structure_partial(const double4 *__restrict__ posq, int stride, const int *__restrict__ list, int nlist,
                    const KVec *__restrict__ kv, int nk, int kmax, CellDev c, double2 *__restrict__ part, int nchunks, const int *__restrict__ dirty) {
	extern __shared__ double2 s_tab[];               // [kSkSites][3*(kmax+1)]
	__shared__ double s_q[kSkSites];
	// `dirty` (may be null): { count, chunk indices ... } — only the chunks that hold a site moved since the partials were last
	// computed; the others are still valid, and summing all of them in chunk order gives the bits of a full evaluation
	// In the dirty form the k vectors are also cut in gridDim.z slices: a move touches one or two chunks per bead system, and at 8 bead
	// systems per GPU that would be 8 busy CTAs of 10 us each.
	if (dirty && (int)blockIdx.x >= dirty[0]) return;
	const int bead = blockIdx.y, chunk = dirty ? dirty[1 + blockIdx.x] : blockIdx.x;
	const int kslices = gridDim.z, kper = (nk + kslices - 1) / kslices, k_lo = blockIdx.z * kper, k_hi = min(nk, k_lo + kper);
	const double4 *pq = posq + (size_t)bead * stride;
	const int rows = 3 * (kmax + 1);
	const int base = chunk * kSkSites;
	const int cnt = min(kSkSites, nlist - base);
	if ((int)threadIdx.x < cnt) {
		const double4 p = pq[list[base + threadIdx.x]];
		s_q[threadIdx.x] = p.w;
		// table stored transposed: element (row, site) at s_tab[site*rows + row]  -> ld = 1 "column stride" trick:
		// build with ld = 1 and col offset = site*rows
		build_phase_table(c, p.x, p.y, p.z, kmax, s_tab + (size_t)threadIdx.x * rows, 1, 0);
	}
	__syncthreads();
	for (int ik = k_lo + threadIdx.x; ik < k_hi; ik += kSkThreads) {
		const KVec k = kv[ik];
		double re = 0, im = 0;
		for (int s = 0; s < cnt; s++) {
			const double2 *t = s_tab + (size_t)s * rows;
			double2 e0 = t[abs(k.l0)];                 if (k.l0 < 0) e0.y = -e0.y;
			double2 e1 = t[(kmax + 1) + abs(k.l1)];    if (k.l1 < 0) e1.y = -e1.y;
			double2 e2 = t[2 * (kmax + 1) + abs(k.l2)]; if (k.l2 < 0) e2.y = -e2.y;
			const double2 e = cmul(cmul(e0, e1), e2);
			re += s_q[s] * e.x;
			im += s_q[s] * e.y;
		}
		part[((size_t)bead * nchunks + chunk) * nk + ik] = make_double2(re, im);
	}
}

// sum the chunk partials in chunk order:  S[bead][k] = (add ? S : 0) + sum_chunk part
__global__ void k_structure_reduce(const double2 *__restrict__ part, int nchunks, int nk, double2 *__restrict__ S,
                                   const double2 *__restrict__ addend) {
	const int bead = blockIdx.y;
	const int ik = blockIdx.x * blockDim.x + threadIdx.x;
	if (ik >= nk) return;
	double2 acc = addend ? addend[(size_t)bead * nk + ik] : make_double2(0.0, 0.0);
	for (int cidx = 0; cidx < nchunks; cidx++) {
		const double2 p = part[((size_t)bead * nchunks + cidx) * nk + ik];
		acc.x += p.x; acc.y += p.y;
	}
	S[(size_t)bead * nk + ik] = acc;
}

// coulombic_reciprocal() (src/System.Energy.cpp:1613-1619): (4 pi / V) sum_k exp(-k^2/4a^2)/k^2 |S_mobile(k)|^2, one CTA per bead
__global__ void k_recip_energy(const double2 *__restrict__ S, const KVec *__restrict__ kv, int nk, double four_pi_over_v,
                               double *__restrict__ out) {
	__shared__ double s_red[256];
	const int bead = blockIdx.x;
	double acc = 0;
	for (int ik = threadIdx.x; ik < nk; ik += blockDim.x) {
		const double2 s = S[(size_t)bead * nk + ik];
		acc += kv[ik].w_energy * (s.x * s.x + s.y * s.y);
	}
	s_red[threadIdx.x] = acc;
	__syncthreads();
	for (int o = blockDim.x / 2; o > 0; o >>= 1) {
		if ((int)threadIdx.x < o) s_red[threadIdx.x] += s_red[threadIdx.x + o];
		__syncthreads();
	}
	if (threadIdx.x == 0) out[bead * 8 + 4] = s_red[0] * four_pi_over_v;   // kResStride record, slot 4
}

// K5 (reciprocal part): recip_term() (src/System.Energy.cpp:2834-2896).  One thread per site; S_all(k) includes the
// frozen sites (:2868-2872).  ef[i] = (8 pi / V) sum_k k w_field(k) [sin(k.r_i) Re S - cos(k.r_i) Im S]  (overwrites ef).
__global__ void __launch_bounds__(kFrSites * kFrLanes)
k_field_recip(const double4 *__restrict__ posq, int n, int stride, const KVec *__restrict__ kv, int nk, int kmax,
              const double2 *__restrict__ S, CellDev c, double eight_pi_over_v, double *__restrict__ ef) {
	// 32 sites x 4 k-lanes per CTA: ~2.1 CTAs per SM at N = 10^4 (the first version ran 79 CTAs of 128 threads on 148 SMs)
	extern __shared__ double2 s_tab[];               // [3*(kmax+1)][kFrSites]
	const int bead = blockIdx.y;
	const int sl = threadIdx.x / kFrLanes, kl = threadIdx.x % kFrLanes;
	const int i = blockIdx.x * kFrSites + sl;
	if (kl == 0) {
		const double4 p = (i < n) ? posq[(size_t)bead * stride + i] : make_double4(0, 0, 0, 0);
		build_phase_table(c, p.x, p.y, p.z, kmax, s_tab, kFrSites, sl);
	}
	__syncthreads();
	double ex = 0, ey = 0, ez = 0;
	for (int ik = kl; ik < nk; ik += kFrLanes) {
		const KVec k = kv[ik];
		const double2 e = cmul(cmul(tab_at(s_tab, kmax, kFrSites, sl, 0, k.l0), tab_at(s_tab, kmax, kFrSites, sl, 1, k.l1)),
		                       tab_at(s_tab, kmax, kFrSites, sl, 2, k.l2));
		const double2 s = S[(size_t)bead * nk + ik];
		const double t = k.w_field * (e.y * s.x - e.x * s.y);
		ex += k.kx * t; ey += k.ky * t; ez += k.kz * t;
	}
#pragma unroll
	for (int o = kFrLanes / 2; o > 0; o >>= 1) {
		ex += __shfl_xor_sync(0xffffffffu, ex, o); ey += __shfl_xor_sync(0xffffffffu, ey, o); ez += __shfl_xor_sync(0xffffffffu, ez, o);
	}
	if (kl == 0 && i < n) {
		double *e = ef + ((size_t)bead * n + i) * 3;
		e[0] = ex * eight_pi_over_v; e[1] = ey * eight_pi_over_v; e[2] = ez * eight_pi_over_v;
	}
}

} // namespace mpmc
