// Developer microbenchmark (sm_100a, B200): is the FP64 tensor-core instruction (mma.sync.m8n8k4.f64, "DMMA") worth using on this path?
//   1. raw throughput and dependent latency of DMMA against DFMA (same accumulators, register-resident operands);
//   2. the one GEMM-shaped step of the Gauss-Seidel precomputation — the rectangular update X[rows] -= L[rows, J] X[J] of the block
//      inverse (k_gs_inverse), 24 x 24 x 192 per site group — done with DFMA (thread = row x column slice) and with DMMA tiles;
//   3. the batched 3 x 3 tensor times 3-vector contraction of the dipole sweeps (9 FMAs per pair) written as DMMA: a matrix-VECTOR
//      product uses one of the eight B columns of m8n8k4, so 7/8 of the instruction's work is wasted.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench/dmma_ubench tools/ubench/dmma_ubench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int ILP, bool TENSOR>
__global__ void k_tput(double *out, int iters) {
	double c[ILP][2];
#pragma unroll
	for (int q = 0; q < ILP; q++) { c[q][0] = threadIdx.x * 1e-9 + q; c[q][1] = 0.5 * q; }
	const double a = 1.0000001, b = 1e-7 * (threadIdx.x & 3);
	for (int i = 0; i < iters; i++) {
#pragma unroll
		for (int q = 0; q < ILP; q++) {
			if (TENSOR) dmma(c[q][0], c[q][1], a, b);
			else { c[q][0] = fma(c[q][0], a, b); c[q][1] = fma(c[q][1], a, b); }
		}
	}
	double s = 0;
#pragma unroll
	for (int q = 0; q < ILP; q++) s += c[q][0] + c[q][1];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <bool TENSOR>
__global__ void k_latency(double *out, long long *cyc, int iters) {
	double c0 = threadIdx.x * 1e-9, c1 = 0.25;
	const double a = 1.0000001, b = 1e-7;
	long long t0 = clock64();
	for (int i = 0; i < iters; i++) {
		if (TENSOR) dmma(c0, c1, a, b);
		else { c0 = fma(c0, a, b); }
	}
	long long t1 = clock64();
	out[threadIdx.x] = c0 + c1;
	if (threadIdx.x == 0) *cyc = t1 - t0;
}

// (2) rectangular update of a block inverse: C[24 x 192] -= A[24 x 24] B[24 x 192], all in shared memory, one CTA of 256 threads per
// problem, `reps` problems per CTA.  DFMA form: thread = (row r of 24, column slice of 18 columns... 192 = 8 warps x 24 columns).
constexpr int M = 24, K = 24, N = 192;
__global__ void __launch_bounds__(256) k_update_dfma(double *out, int reps) {
	__shared__ double sA[M * K], sB[K * N];
	double *sC = out + (size_t)blockIdx.x * M * N;       // C lives in global memory (L1/L2 resident): 37 KB per CTA
	for (int q = threadIdx.x; q < M * K; q += 256) sA[q] = 1e-3 * (q % 7);
	for (int q = threadIdx.x; q < K * N; q += 256) sB[q] = 1e-3 * (q % 5);
	for (int q = threadIdx.x; q < M * N; q += 256) sC[q] = 0.0;
	__syncthreads();
	// thread t: column n = t % 192 ... 256 threads: use 192 of them, each owns one column and all 24 rows? that is 24 accumulators:
	// rows in registers, A broadcast from shared memory, B column in registers
	const int n = threadIdx.x;
	for (int rep = 0; rep < reps; rep++) {
		if (n < N) {
			double b[K], c[M];
#pragma unroll
			for (int k = 0; k < K; k++) b[k] = sB[k * N + n];
#pragma unroll
			for (int m = 0; m < M; m++) c[m] = sC[m * N + n];
#pragma unroll
			for (int m = 0; m < M; m++)
#pragma unroll
				for (int k = 0; k < K; k++) c[m] = fma(-sA[m * K + k], b[k], c[m]);
#pragma unroll
			for (int m = 0; m < M; m++) sC[m * N + n] = c[m];
		}
		__syncthreads();
	}
}
// DMMA form: 8 warps; warp w owns columns 24 w .. 24 w + 23 (3 n-tiles of 8) x 24 rows (3 m-tiles of 8): 9 accumulator tiles, K = 24 = 6 k-steps
__global__ void __launch_bounds__(256) k_update_dmma(double *out, int reps) {
	__shared__ double sA[M * K], sB[K * N];
	double *sC = out + (size_t)blockIdx.x * M * N;       // C lives in global memory (L1/L2 resident): 37 KB per CTA
	for (int q = threadIdx.x; q < M * K; q += 256) sA[q] = 1e-3 * (q % 7);
	for (int q = threadIdx.x; q < K * N; q += 256) sB[q] = 1e-3 * (q % 5);
	for (int q = threadIdx.x; q < M * N; q += 256) sC[q] = 0.0;
	__syncthreads();
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;   // A: row g, col t; B: row t, col g; C: row g, cols 2t, 2t+1
	for (int rep = 0; rep < reps; rep++) {
		double c[3][3][2];
#pragma unroll
		for (int mt = 0; mt < 3; mt++)
#pragma unroll
			for (int nt = 0; nt < 3; nt++) {
				const int r = 8 * mt + g, cc = 24 * warp + 8 * nt + 2 * t;
				c[mt][nt][0] = sC[r * N + cc]; c[mt][nt][1] = sC[r * N + cc + 1];
			}
#pragma unroll
		for (int ks = 0; ks < 6; ks++) {
			double a[3], b[3];
#pragma unroll
			for (int mt = 0; mt < 3; mt++) a[mt] = -sA[(8 * mt + g) * K + 4 * ks + t];
#pragma unroll
			for (int nt = 0; nt < 3; nt++) b[nt] = sB[(4 * ks + t) * N + 24 * warp + 8 * nt + g];
#pragma unroll
			for (int mt = 0; mt < 3; mt++)
#pragma unroll
				for (int nt = 0; nt < 3; nt++) dmma(c[mt][nt][0], c[mt][nt][1], a[mt], b[nt]);
		}
#pragma unroll
		for (int mt = 0; mt < 3; mt++)
#pragma unroll
			for (int nt = 0; nt < 3; nt++) {
				const int r = 8 * mt + g, cc = 24 * warp + 8 * nt + 2 * t;
				sC[r * N + cc] = c[mt][nt][0]; sC[r * N + cc + 1] = c[mt][nt][1];
			}
		__syncthreads();
	}
}

// (3) batched 3x3 tensor (symmetric, 6 numbers) times 3-vector, register resident: DFMA = 9 FMAs per pair per lane.  As DMMA the 8 rows of
// A hold (parts of) three tensors' rows and B one vector in one of its 8 columns: 8 x 8 x 4 = 256 FMAs issued for 9 x (8/3) useful ones.
template <bool TENSOR>
__global__ void k_tensor_vec(double *out, int iters) {
	double t0 = 1.0 + threadIdx.x * 1e-6, t1 = 0.5, t2 = 0.25, t3 = 0.125, t4 = 0.0625, t5 = 0.03;
	double vx = 1e-3, vy = 2e-3, vz = 3e-3, ax = 0, ay = 0, az = 0, c0 = 0, c1 = 0;
	for (int i = 0; i < iters; i++) {
		if (!TENSOR) {
			ax = fma(t0, vx, fma(t3, vy, fma(t4, vz, ax)));
			ay = fma(t3, vx, fma(t1, vy, fma(t5, vz, ay)));
			az = fma(t4, vx, fma(t5, vy, fma(t2, vz, az)));
		} else {
			dmma(c0, c1, t0, vx);       // one k4 step covers a 3-component column with a padding row
		}
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = ax + ay + az + c0 + c1;
}

int main() {
	cudaDeviceProp p;
	CK(cudaGetDeviceProperties(&p, 0));
	printf("%s, %d SMs, clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
	double *d; long long *dc;
	CK(cudaMalloc(&d, sizeof(double) * 148 * 64 * 1024)); CK(cudaMalloc(&dc, 8));
	cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
	auto timeit = [&](auto launch) { launch(); cudaDeviceSynchronize(); cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); return ms; };
	const int iters = 1 << 14, blocks = p.multiProcessorCount * 8, threads = 256;
	{
		float ms = timeit([&] { k_tput<8, false><<<blocks, threads>>>(d, iters); });
		printf("DFMA  ILP 8 x2: %.3f ms  %.2f TFLOP/s\n", ms, 2.0 * 2 * 8 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12);
		ms = timeit([&] { k_tput<8, true><<<blocks, threads>>>(d, iters); });
		printf("DMMA  m8n8k4 ILP 8: %.3f ms  %.2f TFLOP/s  (256 FMA per warp instruction)\n", ms, 2.0 * 256 / 32.0 * 8 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12);
		ms = timeit([&] { k_tput<2, true><<<blocks, threads>>>(d, iters); });
		printf("DMMA  m8n8k4 ILP 2: %.3f ms  %.2f TFLOP/s\n", ms, 2.0 * 256 / 32.0 * 2 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12);
		ms = timeit([&] { k_tput<4, true><<<p.multiProcessorCount, 128>>>(d, iters); });
		printf("DMMA  m8n8k4 ILP 4, 4 warps/SM: %.3f ms  %.2f TFLOP/s\n", ms, 2.0 * 256 / 32.0 * 4 * (double)iters * p.multiProcessorCount * 128 / (ms * 1e-3) / 1e12);
	}
	{
		long long c;
		k_latency<false><<<1, 32>>>(d, dc, 4096); CK(cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost)); printf("latency DFMA %.1f cycles\n", c / 4096.0);
		k_latency<true><<<1, 32>>>(d, dc, 4096); CK(cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost)); printf("latency DMMA %.1f cycles\n", c / 4096.0);
	}
	{
		const int reps = 256, nb = p.multiProcessorCount * 4;
		const double flop = 2.0 * M * K * N * reps * nb;
		float ms = timeit([&] { k_update_dfma<<<nb, 256>>>(d, reps); });
		printf("inverse update 24x24x192 DFMA: %.3f ms  %.2f TFLOP/s\n", ms, flop / (ms * 1e-3) / 1e12);
		ms = timeit([&] { k_update_dmma<<<nb, 256>>>(d, reps); });
		printf("inverse update 24x24x192 DMMA: %.3f ms  %.2f TFLOP/s\n", ms, flop / (ms * 1e-3) / 1e12);
	}
	{
		float ms = timeit([&] { k_tensor_vec<false><<<blocks, threads>>>(d, iters); });
		printf("3x3 tensor x vector, DFMA: %.3f ms  %.1f G contractions/s\n", ms, (double)iters * blocks * threads / (ms * 1e-3) / 1e9);
		ms = timeit([&] { k_tensor_vec<true><<<blocks, threads>>>(d, iters); });
		printf("3x3 tensor x vector, one DMMA per 8/3 contractions (upper bound of that mapping): %.3f ms  %.1f G contractions/s\n", ms, (8.0 / 3.0) / 32.0 * (double)iters * blocks * threads / (ms * 1e-3) / 1e9);
	}
	return 0;
}
