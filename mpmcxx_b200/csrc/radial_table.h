// radial_table.h — host-side builder of piecewise-polynomial tables for the radial functions of the hot path, indexed by the
// SQUARED pair distance u = r^2 (which the kernels hold exactly after the minimum-image step):
//     coulombic_real      erfc(alpha r)/r                                         (reference src/System.Energy.cpp:1493-1497)
//     real_term           (2 a/sqrt(pi) e^{-a^2 r^2} r + erfc(a r))/r^3  and the  "- erf" form of excluded pairs (:2921-2929)
//     thole_amatrix       damp1/r^3  and  3 damp2/r^5  with exponential damping   (:2731-2742)
// These are smooth functions of u > 0, so a degree-7 polynomial per interval of a grid that is uniform in the floating-point
// representation of u (kTabPerOctave intervals per binary octave: the row index is a shift of the high word of u, no log, no
// division, no sqrt) reproduces them to ~2e-16 of their near-field values (tests/test_tables.py).  A lookup costs 1 FP64 add
// + 7 FMAs per function instead of the ~40-60 FP64 instructions of rsqrt + exp/erfc, which is what the FP64-pipe-bound sweeps
// are made of.  Outside [u_lo, u_hi) the kernels evaluate the closed form directly.
//
// Row layout: nfun * 8 coefficients (c0..c7 of function 0, then function 1, ...) followed by 2 doubles of padding, so that the
// 128-bit shared-memory loads of 8 lanes reading 8 different rows fall into different banks.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace mpmc {

constexpr int kTabLog2PerOctave = 5;                     // 32 intervals per octave: interpolation error below the rounding of the evaluation
constexpr int kTabDeg = 7;
constexpr int kTabShift = 20 - kTabLog2PerOctave;        // high word of a double: sign(1) exponent(11) mantissa(20)
constexpr int kTabLog2PerOctaveCoarse = 4;               // 16 per octave: <= 1.3e-13 relative on r^-5 (two-function tables that must stay small)
constexpr int kTabShiftCoarse = 20 - kTabLog2PerOctaveCoarse;
constexpr int kTabPad = 2;

inline int tab_row_stride(int nfun) { return nfun * (kTabDeg + 1) + kTabPad; }

struct RadialTable {
	int nfun = 0, base = 0, nrows = 0, stride = 0, shift = kTabShift;
	double u_lo = 0, u_hi = 0;                           // valid for u_lo <= u < u_hi
	std::vector<double> rows;

	static inline int hi_word(double u) { uint64_t b; std::memcpy(&b, &u, 8); return (int)(b >> 32); }
	static inline double from_hi(int hi) { uint64_t b = (uint64_t)(uint32_t)hi << 32; double u; std::memcpy(&u, &b, 8); return u; }

	// f(u, out[nfun]) in long double; lo/hi are rounded outwards to interval boundaries
	template <class F>
	void build(int nfun_, double lo, double hi, F f, int shift_ = kTabShift) {
		nfun = nfun_; stride = tab_row_stride(nfun); shift = shift_;
		const int kTabShift = shift_;
		base = hi_word(lo) >> kTabShift;
		const int last = hi_word(hi) >> kTabShift;
		nrows = last - base + 1;
		u_lo = from_hi(base << kTabShift);
		u_hi = from_hi((last + 1) << kTabShift);
		rows.assign((size_t)nrows * stride, 0.0);
		constexpr int D = kTabDeg + 1;
		const long double pi = 3.141592653589793238462643383279502884L;
		// Chebyshev -> monomial conversion matrix: T_k(s) = sum_j tm[k][j] s^j
		long double tm[D][D] = {};
		tm[0][0] = 1; tm[1][1] = 1;
		for (int k = 2; k < D; k++)
			for (int j = 0; j < D; j++) tm[k][j] = (j ? 2 * tm[k - 1][j - 1] : 0) - tm[k - 2][j];
		std::vector<long double> val(nfun);
		for (int r = 0; r < nrows; r++) {
			const int hi0 = (base + r) << kTabShift;
			const long double a = from_hi(hi0), b = from_hi(hi0 + (1 << kTabShift));
			const double mid = from_hi(hi0 | (1 << (kTabShift - 1)));      // what the device reconstructs from the bits of u
			const long double half = (b - a) / 2;                           // mid is exactly (a+b)/2
			long double fn[8][D];                                          // function values at the Chebyshev nodes
			for (int i = 0; i < D; i++) {
				const long double s = cosl(pi * (2 * i + 1) / (2 * D));
				f((long double)mid + half * s, val.data());
				for (int q = 0; q < nfun; q++) fn[q][i] = val[q];
			}
			for (int q = 0; q < nfun; q++) {
				long double cheb[D], mono[D] = {};
				for (int k = 0; k < D; k++) {
					long double acc = 0;
					for (int i = 0; i < D; i++) acc += fn[q][i] * cosl(pi * k * (2 * i + 1) / (2 * D));
					cheb[k] = acc * (k ? 2.0L : 1.0L) / D;
				}
				for (int k = 0; k < D; k++)
					for (int j = 0; j < D; j++) mono[j] += cheb[k] * tm[k][j];
				long double scale = 1;
				for (int j = 0; j < D; j++) { rows[(size_t)r * stride + q * D + j] = (double)(mono[j] / scale); scale *= half; }
			}
		}
	}

	// the device's evaluation order (Horner with FMA), for host-side checks
	double eval(int fun, double u) const {
		const int kTabShift = shift;
		const int hi = hi_word(u);
		const int idx = (hi >> kTabShift) - base;
		const double mid = from_hi((hi & ~((1 << kTabShift) - 1)) | (1 << (kTabShift - 1)));
		const double d = u - mid;
		const double *c = &rows[(size_t)idx * stride + fun * (kTabDeg + 1)];
		double v = c[kTabDeg];
		for (int j = kTabDeg - 1; j >= 0; j--) v = std::fma(v, d, c[j]);
		return v;
	}
};

} // namespace mpmc
