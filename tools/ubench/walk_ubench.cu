// Developer microbenchmark: the Gauss-Seidel in-block walk (kernels_gs.cuh) in isolation — one warp, 64 sequential steps,
// tensors in shared memory — to see what a step costs and which part of it.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench/walk_ubench tools/ubench/walk_ubench.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int B = 64;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

// VAR bits: 1 = publish dm/prog to shared memory, 2 = capture acc at update, 4 = two rows per lane (else one),
//           8 = broadcast through shared memory instead of shuffles, 16 = tree-shaped row update
template <int VAR>
__global__ void k_walk(const double *mat_g, double *out, long long *cyc, int reps) {
	extern __shared__ __align__(16) double s_mat[];
	__shared__ double4 s_dm[B];
	__shared__ volatile int s_prog;
	__shared__ volatile double s_bc[4];
	for (int q = threadIdx.x; q < B * B * 6; q += blockDim.x) s_mat[q] = mat_g[q];
	__syncthreads();
	if (threadIdx.x >= 32) return;
	const int lane = threadIdx.x;
	constexpr int NR = (VAR & 4) ? 2 : 1;
	double al[2], cx[2], cy[2], cz[2], ax[2], ay[2], az[2], ex[2] = {0, 0}, ey[2] = {0, 0}, ez[2] = {0, 0};
	for (int h = 0; h < 2; h++) { al[h] = 1.1 + lane * 1e-3; cx[h] = 0.3 + h; cy[h] = 0.2; cz[h] = 0.1; ax[h] = lane * 1e-2; ay[h] = 0.5; az[h] = 0.25; }
	const double2 *tcol = (const double2 *)s_mat + lane * 3;
	long long t0 = clock64();
	for (int rep = 0; rep < reps; rep++) {
		double2 tn[2][3];
#pragma unroll
		for (int hh = 0; hh < NR; hh++) { tn[hh][0] = tcol[hh * 96]; tn[hh][1] = tcol[hh * 96 + 1]; tn[hh][2] = tcol[hh * 96 + 2]; }
#pragma unroll
		for (int half = 0; half < NR; half++) {
#pragma unroll 2
			for (int kk = 0; kk < 32; kk++) {
				const int k = kk + 32 * half;
				double2 tc[2][3];
				const double2 *tnext = tcol + min(k + 1, B - 1) * (B * 3);
#pragma unroll
				for (int hh = 0; hh < NR; hh++) {
					tc[hh][0] = tn[hh][0]; tc[hh][1] = tn[hh][1]; tc[hh][2] = tn[hh][2];
					tn[hh][0] = tnext[hh * 96]; tn[hh][1] = tnext[hh * 96 + 1]; tn[hh][2] = tnext[hh * 96 + 2];
				}
				const double dxc = fma(-al[half], ax[half], cx[half]), dyc = fma(-al[half], ay[half], cy[half]), dzc = fma(-al[half], az[half], cz[half]);
				double dx, dy, dz;
				if (VAR & 8) {
					if (lane == kk) { s_bc[0] = dxc; s_bc[1] = dyc; s_bc[2] = dzc; }
					__syncwarp();
					dx = s_bc[0]; dy = s_bc[1]; dz = s_bc[2];
					__syncwarp();
				} else {
					dx = __shfl_sync(0xffffffffu, dxc, kk); dy = __shfl_sync(0xffffffffu, dyc, kk); dz = __shfl_sync(0xffffffffu, dzc, kk);
				}
				if (VAR & 2) { if (lane == kk) { ex[half] = ax[half]; ey[half] = ay[half]; ez[half] = az[half]; } }
#pragma unroll
				for (int hh = 0; hh < NR; hh++) {
					if (VAR & 16) {
						ax[hh] = (fma(tc[hh][0].x, dx, ax[hh])) + fma(tc[hh][1].y, dy, tc[hh][2].x * dz);
						ay[hh] = (fma(tc[hh][1].y, dx, ay[hh])) + fma(tc[hh][0].y, dy, tc[hh][2].y * dz);
						az[hh] = (fma(tc[hh][2].x, dx, az[hh])) + fma(tc[hh][2].y, dy, tc[hh][1].x * dz);
					} else {
						ax[hh] = fma(tc[hh][0].x, dx, fma(tc[hh][1].y, dy, fma(tc[hh][2].x, dz, ax[hh])));
						ay[hh] = fma(tc[hh][1].y, dx, fma(tc[hh][0].y, dy, fma(tc[hh][2].y, dz, ay[hh])));
						az[hh] = fma(tc[hh][2].x, dx, fma(tc[hh][2].y, dy, fma(tc[hh][1].x, dz, az[hh])));
					}
				}
				if (VAR & 1) {
					volatile double *vd = (volatile double *)(s_dm + k);
					vd[0] = dx; vd[1] = dy; vd[2] = dz;
					s_prog = k + 1;
				}
			}
		}
	}
	long long t1 = clock64();
	out[lane] = ax[0] + ay[0] + az[0] + ax[1] + ex[0] + ey[0] + ez[0] + ex[1];
	if (lane == 0) *cyc = t1 - t0;
}

template <int VAR>
int run(const char *name, const double *mat, double *out, long long *cyc) {
	const int reps = 20;
	CK(cudaFuncSetAttribute(k_walk<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, B * B * 6 * 8));
	k_walk<VAR><<<1, 256, B * B * 6 * 8>>>(mat, out, cyc, reps);
	CK(cudaDeviceSynchronize());
	k_walk<VAR><<<1, 256, B * B * 6 * 8>>>(mat, out, cyc, reps);
	long long h;
	CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
	const int steps = (VAR & 4) ? 64 : 32;
	printf("%-58s %7.1f cycles/step\n", name, (double)h / reps / steps);
	return 0;
}

int main() {
	double *mat, *out; long long *cyc;
	CK(cudaMalloc(&mat, B * B * 6 * 8)); CK(cudaMalloc(&out, 256 * 8)); CK(cudaMalloc(&cyc, 8));
	CK(cudaMemset(mat, 0, B * B * 6 * 8));
	run<4>("2 rows/lane, shuffles", mat, out, cyc);
	run<4 | 1>("2 rows/lane, shuffles, publish", mat, out, cyc);
	run<4 | 2>("2 rows/lane, shuffles, capture", mat, out, cyc);
	run<4 | 1 | 2>("2 rows/lane, shuffles, publish, capture (the kernel)", mat, out, cyc);
	run<4 | 1 | 2 | 16>("  + tree-shaped update", mat, out, cyc);
	run<4 | 1 | 2 | 8>("2 rows/lane, shared-memory broadcast, publish, capture", mat, out, cyc);
	run<0>("1 row/lane, shuffles", mat, out, cyc);
	run<1 | 2>("1 row/lane, shuffles, publish, capture", mat, out, cyc);
	run<1 | 2 | 16>("1 row/lane, shuffles, publish, capture, tree", mat, out, cyc);
	run<1 | 2 | 8>("1 row/lane, shared-memory broadcast, publish, capture", mat, out, cyc);
	return 0;
}
