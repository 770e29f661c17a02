// device_math.cuh — FP64 building blocks shared by every kernel of the engine (sm_100a).
//
// Geometry follows System::minimum_image() (reference src/System.cpp:1202-1279) operation by operation and
// WITHOUT fused multiply-add: the reference is built for baseline x86-64 (no FMA contraction) and its cutoff
// tests (`rimg - 1e-12 < rc`, `r > rc`) act on the last bit of rimg for lattice configurations, so the pair
// distance must round exactly as the reference's does.  Everything downstream of the distance (LJ, erfc,
// exp, tensor contraction) is free to use FMA: those only move results at the 1e-16 level.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include "radial_table.h"

namespace mpmc {

constexpr double kPi           = 3.141592653589793238462643383279502884;   // constants.h:12
constexpr double kOneOverSqrtPi = 0.5641895835477562869480794515607725858440506293289988; // constants.h:48
constexpr double kMaxValue     = 1.0e40;   // constants.h:53
constexpr double kSmallDr      = 1.0e-12;  // constants.h:54
constexpr double kDebye2Ska    = 85.10597636; // constants.h:38

struct CellDev {
	double b[3][3];     // pbc.basis            (rows = lattice vectors)
	double rb[3][3];    // pbc.reciprocal_basis (plain inverse, PeriodicBoundary.cpp:83-101)
	double cutoff, volume, ewald_alpha, polar_alpha;
};

// Thole / solver parameters handed to the polarization kernels
struct PolarDev {
	double damp;        // polar_damp (lambda)
	double gamma;       // polar_gamma
	double allowed_sqerr; // (polar_precision * DEBYE2SKA)^2, System.Energy.cpp:3228
	double u_damp;      // (50 / lambda)^2: beyond this squared distance exponential damping is exactly 1 in double precision
	int    damp_type;   // 0 off, 1 linear, 2 exponential
	int    gs;          // polar_gs || polar_gs_ranked
	int    sor, esor;
};

// rint() (ties to even) for |x| < 2^51 as two FP64 adds; identical to libm rint on that range.
__device__ __forceinline__ double rint_magic(double x) {
	const double M = 6755399441055744.0;   // 1.5 * 2^52
	return __dadd_rn(__dadd_rn(x, M), -M);
}

// d = r_i - r_j  ->  minimum-image displacement (ix,iy,iz).  ORTHO is selected when every off-diagonal element
// of the basis is exactly 0; then the general formula's extra terms are exact zeros and this short form is
// bit-identical to it.
template <bool ORTHO>
__device__ __forceinline__ void min_image(const CellDev &c, double dx, double dy, double dz, double &ix, double &iy, double &iz) {
	if (ORTHO) {
		double fx = rint_magic(__dmul_rn(c.rb[0][0], dx));
		double fy = rint_magic(__dmul_rn(c.rb[1][1], dy));
		double fz = rint_magic(__dmul_rn(c.rb[2][2], dz));
		ix = __dsub_rn(dx, __dmul_rn(c.b[0][0], fx));
		iy = __dsub_rn(dy, __dmul_rn(c.b[1][1], fy));
		iz = __dsub_rn(dz, __dmul_rn(c.b[2][2], fz));
	} else {
		// img[p] = sum_q recip[q][p] * d[q]   (System.cpp:1228-1235), accumulation order q = 0,1,2
		double f0 = rint_magic(__dadd_rn(__dadd_rn(__dmul_rn(c.rb[0][0], dx), __dmul_rn(c.rb[1][0], dy)), __dmul_rn(c.rb[2][0], dz)));
		double f1 = rint_magic(__dadd_rn(__dadd_rn(__dmul_rn(c.rb[0][1], dx), __dmul_rn(c.rb[1][1], dy)), __dmul_rn(c.rb[2][1], dz)));
		double f2 = rint_magic(__dadd_rn(__dadd_rn(__dmul_rn(c.rb[0][2], dx), __dmul_rn(c.rb[1][2], dy)), __dmul_rn(c.rb[2][2], dz)));
		// di[p] = sum_q basis[q][p] * img[q]  (:1238-1242); d - di (:1245-1246)
		ix = __dsub_rn(dx, __dadd_rn(__dadd_rn(__dmul_rn(c.b[0][0], f0), __dmul_rn(c.b[1][0], f1)), __dmul_rn(c.b[2][0], f2)));
		iy = __dsub_rn(dy, __dadd_rn(__dadd_rn(__dmul_rn(c.b[0][1], f0), __dmul_rn(c.b[1][1], f1)), __dmul_rn(c.b[2][1], f2)));
		iz = __dsub_rn(dz, __dadd_rn(__dadd_rn(__dmul_rn(c.b[0][2], f0), __dmul_rn(c.b[1][2], f1)), __dmul_rn(c.b[2][2], f2)));
	}
}

// |v|^2 summed x, y, z with separate multiplies and adds (System.cpp:1250-1255)
__device__ __forceinline__ double norm2_nofma(double x, double y, double z) {
	return __dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z));
}

// Thole damping factors (System.Energy.cpp:2713-2742) at separation r.  `excluded` and `alpha_prod` are only
// read by the OFF / LINEAR variants.
__device__ __forceinline__ void thole_damping(const PolarDev &p, double r, double r2, bool excluded, double alpha_prod,
                                               double &damp1, double &damp2) {
	if (p.damp_type == 2) {
		const double l = p.damp, l2 = l * l, l3 = l2 * l;
		double explr = exp(-l * r);
		damp1 = 1.0 - explr * (0.5 * l2 * r2 + l * r + 1.0);
		damp2 = damp1 - explr * (l3 * r2 * r / 6.0);
	} else if (p.damp_type == 1) {
		double s = p.damp * pow(alpha_prod, 1.0 / 6.0);
		double v = r / s;
		if (r < s) { damp1 = (4.0 - 3.0 * v) * v * v * v; damp2 = v * v * v * v; }
		else damp1 = damp2 = 1.0;
	} else {
		damp1 = damp2 = excluded ? 0.0 : 1.0;
	}
}

// acc += T_ij mu_j with T = damp1/r^3 I - 3 damp2/r^5 d d^T on the minimum-image displacement; no cutoff and
// same-molecule pairs included (thole_amatrix, System.Energy.cpp:2694-2767).
template <bool ORTHO>
__device__ __forceinline__ void tensor_contract(const CellDev &c, const PolarDev &p, double xi, double yi, double zi,
                                                double xj, double yj, double zj, bool excluded, double alpha_prod,
                                                double mx, double my, double mz, double &ax, double &ay, double &az) {
	double dx, dy, dz;
	min_image<ORTHO>(c, __dsub_rn(xi, xj), __dsub_rn(yi, yj), __dsub_rn(zi, zj), dx, dy, dz);
	double a, b;
	if (p.damp_type == 2) {
		// exponential damping, the `cuda on` configuration: one rsqrt gives r and every inverse power.  At r = 0 the reference
		// multiplies MAXVALUE by damping factors that are exactly 0 (:2704-2705, :2733-2734): the pair contributes nothing.
		const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
		if (r2 == 0.0) return;
		const double ir = rsqrt(r2), r = r2 * ir, ir2 = ir * ir, ir3 = ir2 * ir;
		const double lr = p.damp * r;
		const double e = exp(-lr);
		const double damp1 = fma(-e, fma(0.5 * lr, lr, lr) + 1.0, 1.0);            // 1 - e (l^2 r^2 / 2 + l r + 1)
		const double damp2 = fma(-e, lr * lr * lr * (1.0 / 6.0), damp1);           // damp1 - e l^3 r^3 / 6
		a = damp1 * ir3;
		b = 3.0 * damp2 * ir3 * ir2 * (dx * mx + dy * my + dz * mz);
	} else {
		double r2 = norm2_nofma(dx, dy, dz);
		double r = sqrt(r2);
		double ir3, ir5;
		if (r == 0.0) { ir3 = ir5 = kMaxValue; }
		else { double ir = 1.0 / r; double ir2 = ir * ir; ir3 = ir2 * ir; ir5 = ir3 * ir2; }
		double damp1, damp2;
		thole_damping(p, r, r2, excluded, alpha_prod, damp1, damp2);
		a = damp1 * ir3;
		b = 3.0 * damp2 * ir5 * (dx * mx + dy * my + dz * mz);
	}
	ax += a * mx - b * dx;
	ay += a * my - b * dy;
	az += a * mz - b * dz;
}


// ---- fast paths: FMA geometry + radial tables in r^2 (radial_table.h) ------------------------------------------------------
// None of the polarization tensor work involves a cutoff decision, so nothing there depends on the last bit of the distance:
// geometry may use FMA, and the radial factors come from tables.

// minimum-image displacement for the vector/tensor kernels.  The IMAGE is chosen exactly as the reference chooses it — the
// fractional coordinate is formed with the reference's roundings before rint (System.cpp:1228-1235), because for a pair that
// sits half a cell apart (any lattice start) the two images are equally near and the choice flips the sign of that component of
// dimg, hence of the field and of the off-diagonal tensor elements — while the back-projection d - basis^T img uses FMA (that
// only moves the result by an ulp).
template <bool ORTHO>
__device__ __forceinline__ void min_image_fast(const CellDev &c, double dx, double dy, double dz, double &ix, double &iy, double &iz) {
	if (ORTHO) {
		const double fx = rint_magic(__dmul_rn(c.rb[0][0], dx)), fy = rint_magic(__dmul_rn(c.rb[1][1], dy)), fz = rint_magic(__dmul_rn(c.rb[2][2], dz));
		ix = fma(-c.b[0][0], fx, dx); iy = fma(-c.b[1][1], fy, dy); iz = fma(-c.b[2][2], fz, dz);
	} else {
		const double f0 = rint_magic(__dadd_rn(__dadd_rn(__dmul_rn(c.rb[0][0], dx), __dmul_rn(c.rb[1][0], dy)), __dmul_rn(c.rb[2][0], dz)));
		const double f1 = rint_magic(__dadd_rn(__dadd_rn(__dmul_rn(c.rb[0][1], dx), __dmul_rn(c.rb[1][1], dy)), __dmul_rn(c.rb[2][1], dz)));
		const double f2 = rint_magic(__dadd_rn(__dadd_rn(__dmul_rn(c.rb[0][2], dx), __dmul_rn(c.rb[1][2], dy)), __dmul_rn(c.rb[2][2], dz)));
		ix = fma(-c.b[2][0], f2, fma(-c.b[1][0], f1, fma(-c.b[0][0], f0, dx)));
		iy = fma(-c.b[2][1], f2, fma(-c.b[1][1], f1, fma(-c.b[0][1], f0, dy)));
		iz = fma(-c.b[2][2], f2, fma(-c.b[1][2], f1, fma(-c.b[0][2], f0, dz)));
	}
}

// acc += T_ij mu_j like tensor_contract(), for the configuration the reference's `cuda on` validator demands (exponential
// damping, thole_amatrix, System.Energy.cpp:2731-2742): FMA geometry, one rsqrt for every inverse power, and the damping
// factors only where they differ from 1 in double precision: for lambda r >= 50 the terms e^{-lr}(l^2 r^2/2 + l r + 1) and
// e^{-lr} l^3 r^3/6 are below 2^-54, so damp1 = damp2 = 1 exactly as the closed form would round.  (A radial table was tried
// here and lost: two functions x 8 coefficients per lane make the shared-memory pipe, not the FP64 pipe, the bottleneck.)
// ~37 FP64 instructions per far pair, ~65 per damped pair, against ~75 for every pair before.  u_damp = (50 / lambda)^2.
template <bool ORTHO>
__device__ __forceinline__ void tensor_contract_exp(const CellDev &c, double lambda, double u_damp, double xi, double yi, double zi,
                                                    double xj, double yj, double zj, double mx, double my, double mz,
                                                    double &ax, double &ay, double &az) {
	double dx, dy, dz;
	min_image_fast<ORTHO>(c, __dsub_rn(xi, xj), __dsub_rn(yi, yj), __dsub_rn(zi, zj), dx, dy, dz);
	const double u = fma(dz, dz, fma(dy, dy, dx * dx));
	// coincident sites (and the i == j column): the reference multiplies MAXVALUE by damping factors that are exactly 0 (:2704-2705)
	const double ir = u > 0.0 ? rsqrt(u) : 0.0;
	const double ir2 = ir * ir, ir3 = ir2 * ir;
	double a = ir3, b = 3.0 * ir3 * ir2;
	if (u < u_damp) {
		const double lr = lambda * (u * ir);
		const double e = exp(-lr);
		const double damp1 = fma(-e, fma(0.5 * lr, lr, lr) + 1.0, 1.0);            // 1 - e (l^2 r^2 / 2 + l r + 1)
		const double damp2 = fma(-e, lr * lr * lr * (1.0 / 6.0), damp1);           // damp1 - e l^3 r^3 / 6
		a *= damp1; b *= damp2;
	}
	const double bd = b * fma(dz, mz, fma(dy, my, dx * mx));
	ax = fma(-bd, dx, fma(a, mx, ax));
	ay = fma(-bd, dy, fma(a, my, ay));
	az = fma(-bd, dz, fma(a, mz, az));
}

// ---- branch-free building blocks for the contraction sweeps ---------------------------------------------------------------
// The library rsqrt() and exp() each carry a rare-case branch; with one of each per pair the compiler emits every contraction as
// its own sequence of divergence regions and cannot interleave two of them: a warp then runs its contractions back to back, one
// dependent chain at a time (~150 instructions, ~700 cycles each at 4 warps per scheduler), and the FP64 pipe idles half the
// time.  These forms have no branches, so two contractions can be issued as one interleaved stream.

// 1/sqrt(u) for normal positive u: hardware seed (2^-23) + two Newton steps; <= 1 ulp
__device__ __forceinline__ double rsqrt_nr(double u) {
	double y;
	asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(u));
	const double h = 0.5 * u;
	double e = fma(-(h * y), y, 0.5);
	y = fma(y, e, y);
	e = fma(-(h * y), y, 0.5);
	return fma(y, e, y);
}

// e^x for x in [-700, 0]: x = k ln2 + r, |r| <= ln2 / 2, degree-13 Taylor polynomial (truncation 4e-18), 2^k through the exponent field; <= 2 ulp
__device__ __forceinline__ double exp_nonpos(double x) {
	const double kMagic = 6755399441055744.0;
	double t = fma(x, 1.4426950408889634074, kMagic);
	const int k = __double2loint(t);
	t -= kMagic;
	double r = fma(-t, 6.93147180369123816490e-01, x);
	r = fma(-t, 1.90821492927058770002e-10, r);
	double p = 1.6059043836821613e-10;                 // 1/13!
	p = fma(p, r, 2.0876756987868100e-09);            // 1/12!
	p = fma(p, r, 2.5052108385441720e-08);            // 1/11!
	p = fma(p, r, 2.7557319223985888e-07);            // 1/10!
	p = fma(p, r, 2.7557319223985893e-06);            // 1/9!
	p = fma(p, r, 2.4801587301587302e-05);            // 1/8!
	p = fma(p, r, 1.9841269841269841e-04);            // 1/7!
	p = fma(p, r, 1.3888888888888889e-03);            // 1/6!
	p = fma(p, r, 8.3333333333333332e-03);            // 1/5!
	p = fma(p, r, 4.1666666666666664e-02);            // 1/4!
	p = fma(p, r, 1.6666666666666666e-01);            // 1/3!
	p = fma(p, r, 0.5);
	p = fma(p, r, 1.0);
	p = fma(p, r, 1.0);
	return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

// acc += T_i,j0 mu_j0 + T_i,j1 mu_j1 (exponential damping), the two contractions as ONE branch-free stream apart from a single test:
// the damping factors are evaluated for BOTH columns when either needs them (for lambda r >= 50 they come out as exactly 1, as
// tensor_contract_exp skips them).  The sums are taken in the order j0, j1, like two calls of tensor_contract_exp.
template <bool ORTHO>
__device__ __forceinline__ void tensor_contract_exp_x2(const CellDev &c, double lambda, double u_damp, double xi, double yi, double zi,
                                                       double x0, double y0, double z0, double m0x, double m0y, double m0z,
                                                       double x1, double y1, double z1, double m1x, double m1y, double m1z,
                                                       double &ax, double &ay, double &az) {
	double d0x, d0y, d0z, d1x, d1y, d1z;
	min_image_fast<ORTHO>(c, __dsub_rn(xi, x0), __dsub_rn(yi, y0), __dsub_rn(zi, z0), d0x, d0y, d0z);
	min_image_fast<ORTHO>(c, __dsub_rn(xi, x1), __dsub_rn(yi, y1), __dsub_rn(zi, z1), d1x, d1y, d1z);
	const double u0 = fma(d0z, d0z, fma(d0y, d0y, d0x * d0x)), u1 = fma(d1z, d1z, fma(d1y, d1y, d1x * d1x));
	// coincident sites (and the i == j column): the reference multiplies MAXVALUE by damping factors that are exactly 0 (:2704-2705)
	const double ir0 = u0 > 0.0 ? rsqrt_nr(u0 > 0.0 ? u0 : 1.0) : 0.0, ir1 = u1 > 0.0 ? rsqrt_nr(u1 > 0.0 ? u1 : 1.0) : 0.0;
	const double i20 = ir0 * ir0, i21 = ir1 * ir1, i30 = i20 * ir0, i31 = i21 * ir1;
	double a0 = i30, b0 = 3.0 * i30 * i20, a1 = i31, b1 = 3.0 * i31 * i21;
	if (u0 < u_damp || u1 < u_damp) {
		const double l0 = lambda * (u0 * ir0), l1 = lambda * (u1 * ir1);
		const double e0 = exp_nonpos(-l0), e1 = exp_nonpos(-l1);
		const double g0 = fma(-e0, fma(0.5 * l0, l0, l0) + 1.0, 1.0), g1 = fma(-e1, fma(0.5 * l1, l1, l1) + 1.0, 1.0);   // 1 - e (l^2 r^2 / 2 + l r + 1)
		const double h0 = fma(-e0, l0 * l0 * l0 * (1.0 / 6.0), g0), h1 = fma(-e1, l1 * l1 * l1 * (1.0 / 6.0), g1);       // damp1 - e l^3 r^3 / 6
		a0 *= g0; b0 *= h0; a1 *= g1; b1 *= h1;
	}
	const double q0 = b0 * fma(d0z, m0z, fma(d0y, m0y, d0x * m0x)), q1 = b1 * fma(d1z, m1z, fma(d1y, m1y, d1x * m1x));
	ax = fma(-q1, d1x, fma(a1, m1x, fma(-q0, d0x, fma(a0, m0x, ax))));
	ay = fma(-q1, d1y, fma(a1, m1y, fma(-q0, d0y, fma(a0, m0y, ay))));
	az = fma(-q1, d1z, fma(a1, m1z, fma(-q0, d0z, fma(a0, m0z, az))));
}

// copy a table into shared memory (all threads of the CTA; len even, both 16-byte aligned): asynchronous 16-byte copies, all in
// flight at once.  The caller runs stage_table_wait() and then synchronises the CTA before the first look-up.
__device__ __forceinline__ void stage_table(double *dst, const double *__restrict__ src, int len) {
	const unsigned s0 = (unsigned)__cvta_generic_to_shared(dst);
	for (int q = threadIdx.x; q < len / 2; q += blockDim.x)
		asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s0 + 16u * q), "l"(src + 2 * q) : "memory");
	asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void stage_table_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// deterministic warp reductions (xor tree: every lane ends with the same value)
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}

} // namespace mpmc
