// kernels_polar.cuh — Thole induced-dipole solve without ever materialising the 3N x 3N A matrix
// (reference: thole_amatrix + thole_iterative + contract_dipoles + palmo_contraction, src/System.Energy.cpp:2661-2770,
// 3450-3656).  T_ij is recomputed from the pair geometry inside every contraction: the kernels are FP64-pipe bound
// and read only O(N) bytes per sweep.
#pragma once
#include <cooperative_groups.h>
#include "kernels_pair.cuh"

namespace mpmc {
namespace cg = cooperative_groups;

// init_dipoles() (System.Energy.cpp:3547-3560): mu = alpha E_static (* gamma unless SOR/ESOR); clears the per-sweep fields.
__global__ void k_dipole_init(const double *__restrict__ alpha, const double *__restrict__ efs, int n, int nbeads, double gamma_init,
                              double *__restrict__ mu, double *__restrict__ new_mu, double *__restrict__ old_mu,
                              double *__restrict__ efi, double *__restrict__ efic, double *__restrict__ rrms) {
	const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= (size_t)n * nbeads * 3) return;
	const int i = (int)((t / 3) % n);
	const double m = alpha[i] * efs[t] * gamma_init;
	mu[t] = m; new_mu[t] = m; old_mu[t] = 0; efi[t] = 0; efic[t] = 0;
	if (t % 3 == 0) rrms[t / 3] = 0;
}

// One full contraction acc_i = sum_{j != i} T_ij mu_j over the polarizable sites j (mu_j == 0 exactly elsewhere).
//   PALMO = false: contract_dipoles() in Jacobi form (:3564-3598): efi = -acc, new_mu = alpha (E_s + efi); sites with
//                  alpha == 0 get efi = new_mu = 0.  mu is NOT touched (the caller relaxes it afterwards, :3526-3536).
//   PALMO = true : palmo_contraction() (:3602-3627): efic = -efi - acc, for every site.
// Layout as the other ordered sweeps: CTA = 32 sites x 8 j-lanes.
template <bool ORTHO, bool PALMO>
__global__ void __launch_bounds__(kOrdThreads)
k_dipole_sweep(const double4 *__restrict__ posq, const double *__restrict__ alpha, const int *__restrict__ meta,
               const int *__restrict__ plist, int np, int n, int stride, CellDev c, PolarDev p,
               const double *__restrict__ mu, const double *__restrict__ efs, double *__restrict__ efi,
               double *__restrict__ new_mu, double *__restrict__ efic) {
	__shared__ double4 s_pq[kOrdTileJ];
	__shared__ double  s_mu[kOrdTileJ][3];
	__shared__ double  s_al[kOrdTileJ];
	__shared__ int     s_meta[kOrdTileJ];
	__shared__ int     s_idx[kOrdTileJ];
	const int bead = blockIdx.y;
	const double4 *pq = posq + (size_t)bead * stride;
	const double *mub = mu + (size_t)bead * n * 3;
	const int tid = threadIdx.x, jl = tid % kOrdJ, il = tid / kOrdJ;
	const int i = blockIdx.x * kOrdI + il;
	double4 pi = make_double4(0, 0, 0, 0);
	double ai = 0; int mi = 0;
	if (i < n) { pi = pq[i]; ai = alpha[i]; mi = meta[i]; }
	const bool active = (i < n) && (PALMO || ai != 0.0);
	double ax = 0, ay = 0, az = 0;
	for (int j0 = 0; j0 < np; j0 += kOrdTileJ) {
		__syncthreads();
		if (j0 + tid < np) {
			const int j = plist[j0 + tid];
			s_idx[tid] = j; s_pq[tid] = pq[j]; s_al[tid] = alpha[j]; s_meta[tid] = meta[j];
			s_mu[tid][0] = mub[3 * j]; s_mu[tid][1] = mub[3 * j + 1]; s_mu[tid][2] = mub[3 * j + 2];
		}
		__syncthreads();
		const int jn = min(kOrdTileJ, np - j0);
		if (active)
			for (int jj = jl; jj < jn; jj += kOrdJ) {
				if (s_idx[jj] == i) continue;
				const double4 pj = s_pq[jj];
				const bool excl = (meta_mol(mi) == meta_mol(s_meta[jj])) || pi.w == 0.0 || pj.w == 0.0;
				tensor_contract<ORTHO>(c, p, pi.x, pi.y, pi.z, pj.x, pj.y, pj.z, excl, ai * s_al[jj],
				                       s_mu[jj][0], s_mu[jj][1], s_mu[jj][2], ax, ay, az);
			}
	}
	ax = jlane_sum(ax); ay = jlane_sum(ay); az = jlane_sum(az);
	if (jl == 0 && i < n) {
		const size_t o = ((size_t)bead * n + i) * 3;
		if (PALMO) {
			efic[o] = -efi[o] - ax; efic[o + 1] = -efi[o + 1] - ay; efic[o + 2] = -efi[o + 2] - az;
		} else if (ai != 0.0) {
			efi[o] = -ax; efi[o + 1] = -ay; efi[o + 2] = -az;
			new_mu[o] = ai * (efs[o] - ax); new_mu[o + 1] = ai * (efs[o + 1] - ay); new_mu[o + 2] = ai * (efs[o + 2] - az);
		} else {
			efi[o] = efi[o + 1] = efi[o + 2] = 0;
			new_mu[o] = new_mu[o + 1] = new_mu[o + 2] = 0;
		}
	}
}

// calc_dipole_rrms() (:3147-3177) and the precision branch of are_we_done_yet() (:3227-3236).  flags[bead] is set to 1 when
// some component still moves by more than the allowed error.
__global__ void k_dipole_check(const double *__restrict__ new_mu, const double *__restrict__ old_mu, int n, int nbeads,
                               int want_rrms, double allowed_sqerr, double *__restrict__ rrms, int *__restrict__ flags) {
	const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= (size_t)n * nbeads) return;
	const double *nm = new_mu + 3 * s, *om = old_mu + 3 * s;
	double carry = 0, nn = 0;
	bool broke = false;
	for (int q = 0; q < 3; q++) {
		const double e = nm[q] - om[q];
		carry += e * e; nn += nm[q] * nm[q];
		if (e * e > allowed_sqerr) broke = true;
	}
	if (want_rrms) {
		double r = sqrt(carry / nn);
		rrms[s] = isfinite(r) ? r : 0.0;
	}
	if (allowed_sqerr > 0 && broke) flags[s / n] = 1;
}

// "save the dipoles for the next pass" (:3526-3536): plain, SOR or ESOR relaxation.  esor_w = exp(-gamma * iteration).
__global__ void k_mu_update(const double *__restrict__ new_mu, const double *__restrict__ old_mu, size_t len, int sor, int esor,
                            double gamma, double esor_w, double *__restrict__ mu) {
	const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= len) return;
	if (sor) mu[t] = gamma * new_mu[t] + (1.0 - gamma) * old_mu[t];
	else if (esor) mu[t] = (1.0 - esor_w) * new_mu[t] + esor_w * old_mu[t];
	else mu[t] = new_mu[t];
}

// convergence failure after MAX_ITERATION_COUNT (:3483-3494): mu = alpha E_static, efic = 0
__global__ void k_dipole_fail(const double *__restrict__ alpha, const double *__restrict__ efs, int n, size_t off, double *__restrict__ mu,
                              double *__restrict__ efic) {
	const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= (size_t)n * 3) return;
	mu[off + t] = alpha[t / 3] * efs[off + t];
	efic[off + t] = 0;
}

// tail of polar() (:2609-2618) and get_dipole_rrms() (:2639-2657): out[bead] = { sum mu.E_s, sum mu.dE_ind, sum rrms }
__global__ void k_polar_energy(const double *__restrict__ mu, const double *__restrict__ efs, const double *__restrict__ efic,
                               const double *__restrict__ rrms, int n, double *__restrict__ out) {
	__shared__ double s_red[3][256];
	const int bead = blockIdx.x;
	double a = 0, b = 0, r = 0;
	for (int i = threadIdx.x; i < n; i += blockDim.x) {
		const size_t o = ((size_t)bead * n + i) * 3;
		a += mu[o] * efs[o] + mu[o + 1] * efs[o + 1] + mu[o + 2] * efs[o + 2];
		b += mu[o] * efic[o] + mu[o + 1] * efic[o + 1] + mu[o + 2] * efic[o + 2];
		const double rr = rrms[(size_t)bead * n + i];
		if (isfinite(rr)) r += rr;
	}
	s_red[0][threadIdx.x] = a; s_red[1][threadIdx.x] = b; s_red[2][threadIdx.x] = r;
	__syncthreads();
	for (int o = blockDim.x / 2; o > 0; o >>= 1) {
		if ((int)threadIdx.x < o)
			for (int q = 0; q < 3; q++) s_red[q][threadIdx.x] += s_red[q][threadIdx.x + o];
		__syncthreads();
	}
	if (threadIdx.x == 0) { out[3 * bead] = s_red[0][0]; out[3 * bead + 1] = s_red[1][0]; out[3 * bead + 2] = s_red[2][0]; }
}

// stable descending order of the polarizable sites by rank metric (update_ranking, :3631-3656, restricted to alpha != 0:
// the other sites only ever set mu = 0).  order[pos] = site.
__global__ void k_rank_order_plist(const double *__restrict__ rank, const int *__restrict__ plist, int np, int *__restrict__ order) {
	__shared__ double s_m[256];
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	const int i = t < np ? plist[t] : 0;
	const double mi = t < np ? rank[i] : 0;
	int pos = 0;
	for (int j0 = 0; j0 < np; j0 += 256) {
		__syncthreads();
		if (j0 + (int)threadIdx.x < np) s_m[threadIdx.x] = rank[plist[j0 + threadIdx.x]];
		__syncthreads();
		const int jn = min(256, np - j0);
		for (int jj = 0; jj < jn; jj++) {
			const double mj = s_m[jj];
			pos += (mj > mi) || (mj == mi && j0 + jj < t);
		}
	}
	if (t < np) order[pos] = i;
}

// ---------------------------------------------------------------------------------------------------------
// Gauss-Seidel sweep, mathematically sequential in `order` (contract_dipoles with polar_gs / polar_gs_ranked,
// :3570-3595: mu_i is overwritten as soon as it is computed, so site i sees the NEW dipoles of every site swept
// before it and the OLD dipoles of the rest).  Blocked: kGsB sites at a time.
//   phase A (all CTAs): ext_i = sum over every j outside the block of T_ij mu_j(current)          -> grid.sync
//   phase B (CTA 0)   : in-block tensors in shared memory, acc_i = ext_i + sum_{m in block} T_im mu_m(old);
//                       warp 0 then walks the block in order: mu_k = alpha_k (E_s,k - acc_k), and every lane adds
//                       T_mk (mu_k_new - mu_k_old) to the rows m > k it owns                          -> grid.sync
// One cooperative launch per sweep; grid = every CTA the device can hold at once.
// ---------------------------------------------------------------------------------------------------------
constexpr int kGsB = 64, kGsThreads = 256, kGsJL = kGsThreads / kGsB;
constexpr int kGsPairs = kGsB * (kGsB - 1) / 2;
__host__ __device__ constexpr int gs_tri(int a, int b) { return a * (2 * kGsB - a - 1) / 2 + (b - a - 1); }   // a < b
constexpr size_t kGsSmemBytes = sizeof(double) * (6 * kGsPairs + kGsJL * kGsB * 3 + kGsB * 16) + sizeof(int) * kGsB * 2;

template <bool ORTHO>
__global__ void __launch_bounds__(kGsThreads)
k_gs_sweep(const double4 *__restrict__ pq, const double *__restrict__ alpha, const int *__restrict__ meta,
           const int *__restrict__ order, int np, CellDev c, PolarDev p,
           double *mu, const double *__restrict__ efs, double *efi, double *new_mu, double *part) {
	cg::grid_group grid = cg::this_grid();
	extern __shared__ double s_raw[];
	double *s_tri  = s_raw;                                  // [6][kGsPairs]
	double *s_acc  = s_tri + 6 * kGsPairs;                   // [kGsJL][kGsB][3]   (phase A lane partials / phase B scratch)
	double *s_site = s_acc + kGsJL * kGsB * 3;               // [16][kGsB]: x y z q alpha mu_old(3) efs(3) acc(3) new(2 spare)
	int    *s_idx  = (int *)(s_site + 16 * kGsB);            // [kGsB] site index, [kGsB] meta
	int    *s_met  = s_idx + kGsB;
	const int tid = threadIdx.x, il = tid % kGsB, jl = tid / kGsB;
	const int G = gridDim.x, cta = blockIdx.x;
	const int nblk = (np + kGsB - 1) / kGsB;

	for (int blk = 0; blk < nblk; blk++) {
		const int base = blk * kGsB, cnt = min(kGsB, np - base);
		// ---- phase A ----
		if (tid < cnt) {
			const int s = order[base + tid];
			const double4 v = pq[s];
			s_idx[tid] = s; s_met[tid] = meta[s];
			s_site[0 * kGsB + tid] = v.x; s_site[1 * kGsB + tid] = v.y; s_site[2 * kGsB + tid] = v.z; s_site[3 * kGsB + tid] = v.w;
			s_site[4 * kGsB + tid] = alpha[s];
		}
		__syncthreads();
		double ax = 0, ay = 0, az = 0;
		if (il < cnt) {
			const double xi = s_site[il], yi = s_site[kGsB + il], zi = s_site[2 * kGsB + il], qi = s_site[3 * kGsB + il], ai = s_site[4 * kGsB + il];
			const int mi = s_met[il];
			if (ai != 0.0)
				for (int pos = cta * kGsJL + jl; pos < np; pos += G * kGsJL) {
					if (pos >= base && pos < base + cnt) continue;
					const int j = order[pos];                           // same j for the whole warp: broadcast loads
					const double4 pj = pq[j];
					const bool excl = (meta_mol(mi) == meta_mol(meta[j])) || qi == 0.0 || pj.w == 0.0;
					tensor_contract<ORTHO>(c, p, xi, yi, zi, pj.x, pj.y, pj.z, excl, ai * alpha[j], __ldcg(mu + 3 * j), __ldcg(mu + 3 * j + 1), __ldcg(mu + 3 * j + 2), ax, ay, az);
				}
		}
		s_acc[(jl * kGsB + il) * 3 + 0] = ax; s_acc[(jl * kGsB + il) * 3 + 1] = ay; s_acc[(jl * kGsB + il) * 3 + 2] = az;
		__syncthreads();
		if (tid < kGsB * 3) {
			double v = 0;
			for (int q = 0; q < kGsJL; q++) v += s_acc[q * kGsB * 3 + tid];
			part[(size_t)cta * kGsB * 3 + tid] = v;
		}
		__threadfence();
		grid.sync();
		// ---- phase B ----
		if (cta == 0) {
			if (tid < kGsB * 3) {                                   // ext = sum over CTAs in CTA order
				double v = 0;
				for (int q = 0; q < G; q++) v += __ldcg(part + (size_t)q * kGsB * 3 + tid);
				const int m = tid / 3, comp = tid % 3;
				s_site[(11 + comp) * kGsB + m] = v;                  // acc rows 11..13
				if (m < cnt) {
					const int s = s_idx[m];
					s_site[(5 + comp) * kGsB + m] = __ldcg(mu + 3 * s + comp);   // mu_old rows 5..7
					s_site[(8 + comp) * kGsB + m] = efs[3 * s + comp];    // E_static rows 8..10
				}
			}
			for (int q = tid; q < cnt * cnt; q += kGsThreads) {       // in-block tensors, a < b
				const int a = q / cnt, b = q % cnt;
				if (a >= b) continue;
				double dx, dy, dz;
				min_image<ORTHO>(c, __dsub_rn(s_site[a], s_site[b]), __dsub_rn(s_site[kGsB + a], s_site[kGsB + b]),
				                 __dsub_rn(s_site[2 * kGsB + a], s_site[2 * kGsB + b]), dx, dy, dz);
				const double r2 = norm2_nofma(dx, dy, dz), r = sqrt(r2);
				double ir3, ir5;
				if (r == 0.0) { ir3 = ir5 = kMaxValue; } else { const double ir = 1.0 / r, ir2 = ir * ir; ir3 = ir2 * ir; ir5 = ir3 * ir2; }
				const bool excl = (meta_mol(s_met[a]) == meta_mol(s_met[b])) || s_site[3 * kGsB + a] == 0.0 || s_site[3 * kGsB + b] == 0.0;
				double d1, d2;
				thole_damping(p, r, r2, excl, s_site[4 * kGsB + a] * s_site[4 * kGsB + b], d1, d2);
				const double ta = d1 * ir3, tb = 3.0 * d2 * ir5;
				const int t = gs_tri(a, b);
				s_tri[0 * kGsPairs + t] = ta - tb * dx * dx; s_tri[1 * kGsPairs + t] = ta - tb * dy * dy; s_tri[2 * kGsPairs + t] = ta - tb * dz * dz;
				s_tri[3 * kGsPairs + t] = -tb * dx * dy;     s_tri[4 * kGsPairs + t] = -tb * dx * dz;     s_tri[5 * kGsPairs + t] = -tb * dy * dz;
			}
			__syncthreads();
			if (tid < kGsB * 3) {                                   // acc_m += sum_{m' != m in block} T_mm' mu_m'(old)
				const int m = tid / 3, comp = tid % 3;
				if (m < cnt) {
					double v = s_site[(11 + comp) * kGsB + m];
					for (int o = 0; o < cnt; o++) {
						if (o == m) continue;
						const int t = m < o ? gs_tri(m, o) : gs_tri(o, m);
						// row `comp` of the symmetric 3x3: (xx xy xz / xy yy yz / xz yz zz)
						const double t0 = s_tri[(comp == 0 ? 0 : comp == 1 ? 3 : 4) * kGsPairs + t];
						const double t1 = s_tri[(comp == 0 ? 3 : comp == 1 ? 1 : 5) * kGsPairs + t];
						const double t2 = s_tri[(comp == 0 ? 4 : comp == 1 ? 5 : 2) * kGsPairs + t];
						v += t0 * s_site[5 * kGsB + o] + t1 * s_site[6 * kGsB + o] + t2 * s_site[7 * kGsB + o];
					}
					s_acc[m * 3 + comp] = v;
				}
			}
			__syncthreads();
			if (tid < 32) {                                         // sequential walk by one warp; lane owns rows lane, lane+32
				const int lane = tid;
				double acc[2][3], res_mu[2][3], res_ef[2][3];
				for (int h = 0; h < 2; h++) for (int q = 0; q < 3; q++) { acc[h][q] = s_acc[(lane + 32 * h) * 3 + q]; res_mu[h][q] = 0; res_ef[h][q] = 0; }
				for (int k = 0; k < cnt; k++) {
					const int owner = k & 31, h = k >> 5;
					double dmx = 0, dmy = 0, dmz = 0;
					if (lane == owner) {
						const double ak = s_site[4 * kGsB + k];
						const double c0 = h ? acc[1][0] : acc[0][0], c1 = h ? acc[1][1] : acc[0][1], c2 = h ? acc[1][2] : acc[0][2];
						double nx = 0, ny = 0, nz = 0, ex = 0, ey = 0, ez = 0;
						if (ak != 0.0) {
							ex = -c0; ey = -c1; ez = -c2;
							nx = ak * (s_site[8 * kGsB + k] + ex); ny = ak * (s_site[9 * kGsB + k] + ey); nz = ak * (s_site[10 * kGsB + k] + ez);
						}
						if (h == 0) { res_mu[0][0] = nx; res_mu[0][1] = ny; res_mu[0][2] = nz; res_ef[0][0] = ex; res_ef[0][1] = ey; res_ef[0][2] = ez; }
						else        { res_mu[1][0] = nx; res_mu[1][1] = ny; res_mu[1][2] = nz; res_ef[1][0] = ex; res_ef[1][1] = ey; res_ef[1][2] = ez; }
						dmx = nx - s_site[5 * kGsB + k]; dmy = ny - s_site[6 * kGsB + k]; dmz = nz - s_site[7 * kGsB + k];
					}
					dmx = __shfl_sync(0xffffffffu, dmx, owner); dmy = __shfl_sync(0xffffffffu, dmy, owner); dmz = __shfl_sync(0xffffffffu, dmz, owner);
#pragma unroll
					for (int hh = 0; hh < 2; hh++) {
						const int m = lane + 32 * hh;
						if (m > k && m < cnt) {
							const int t = gs_tri(k, m);
							const double xx = s_tri[t], yy = s_tri[kGsPairs + t], zz = s_tri[2 * kGsPairs + t];
							const double xy = s_tri[3 * kGsPairs + t], xz = s_tri[4 * kGsPairs + t], yz = s_tri[5 * kGsPairs + t];
							acc[hh][0] += xx * dmx + xy * dmy + xz * dmz;
							acc[hh][1] += xy * dmx + yy * dmy + yz * dmz;
							acc[hh][2] += xz * dmx + yz * dmy + zz * dmz;
						}
					}
				}
				for (int h = 0; h < 2; h++) {
					const int m = lane + 32 * h;
					if (m < cnt) {
						const int s = s_idx[m];
						for (int q = 0; q < 3; q++) { mu[3 * s + q] = res_mu[h][q]; new_mu[3 * s + q] = res_mu[h][q]; efi[3 * s + q] = res_ef[h][q]; }
					}
				}
			}
		}
		__threadfence();
		grid.sync();
	}
}

// bead-chain bookkeeping ------------------------------------------------------------------------------------
// Molecule::update_COM (src/Molecule.cpp:256-281): com = sum m r / sum m over the sites of one molecule, in list order
__global__ void k_mol_com(const double4 *__restrict__ posq, int stride, const double *__restrict__ mass,
                          const int *__restrict__ mol_start, int nmol, int nbeads, double *__restrict__ com, double *__restrict__ mol_mass) {
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= nmol * nbeads) return;
	const int bead = t / nmol, m = t % nmol;
	const double4 *pq = posq + (size_t)bead * stride;
	double ms = 0, cx = 0, cy = 0, cz = 0;
	for (int i = mol_start[m]; i < mol_start[m + 1]; i++) {
		const double w = mass[i];
		ms = __dadd_rn(ms, w);
		cx = __dadd_rn(cx, __dmul_rn(w, pq[i].x)); cy = __dadd_rn(cy, __dmul_rn(w, pq[i].y)); cz = __dadd_rn(cz, __dmul_rn(w, pq[i].z));
	}
	com[3 * (size_t)t] = cx / ms; com[3 * (size_t)t + 1] = cy / ms; com[3 * (size_t)t + 2] = cz / ms;
	if (bead == 0) mol_mass[m] = ms;
}

// PI_chain_mass_length2 (src/SimulationControl.PathIntegral.cpp:916-970) for every mobile molecule over the LOCAL links
// b -> b+1 (and last -> first when closed); per-molecule results, summed by the caller in molecule order.
__global__ void k_chain_len2(const double *__restrict__ com, const double *__restrict__ mol_mass, const unsigned char *__restrict__ mol_mobile,
                             int nmol, int nbeads, int closed, double *__restrict__ per_mol) {
	const int m = blockIdx.x * blockDim.x + threadIdx.x;
	if (m >= nmol) return;
	double len = 0;
	if (mol_mobile[m]) {
		const int links = closed ? nbeads : nbeads - 1;
		for (int b = 0; b < links; b++) {
			const int b2 = (b + 1) % nbeads;
			const double *a = com + 3 * ((size_t)b * nmol + m), *z = com + 3 * ((size_t)b2 * nmol + m);
			const double dx = a[0] - z[0], dy = a[1] - z[1], dz = a[2] - z[2];
			len += dx * dx + dy * dy + dz * dz;
		}
		len *= (mol_mass[m] * 1.66053873e-27) * (1.0e-10 * 1.0e-10);   // AMU2KG, ANGSTROM2METER^2 (constants.h:35,27)
	}
	per_mol[m] = len;
}

// FP64 FMA peak probe: 8 independent chains per thread, 4096 x 8 FMAs each
__global__ void k_fp64_probe(double *out, int iters) {
	double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
	const double m = 1.0000001, b = 1e-7;
	for (int i = 0; i < iters; i++) {
		a0 = fma(a0, m, b); a1 = fma(a1, m, b); a2 = fma(a2, m, b); a3 = fma(a3, m, b);
		a4 = fma(a4, m, b); a5 = fma(a5, m, b); a6 = fma(a6, m, b); a7 = fma(a7, m, b);
	}
	out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

} // namespace mpmc
