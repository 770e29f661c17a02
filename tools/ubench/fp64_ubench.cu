// Developer microbenchmarks for the FP64 path of sm_100a (B200): latencies of the dependent operations that bound the
// Gauss-Seidel walk, and FP64-pipe throughput as a function of resident warps, ILP and instruction mix.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench/fp64_ubench tools/ubench/fp64_ubench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int OP>
__global__ void k_latency(double *out, long long *cyc, int iters, double seed) {
	double a = seed + threadIdx.x * 1e-9, b = 1.0000001, c = 1e-9;
	__shared__ double s[64];
	s[threadIdx.x & 63] = seed; __syncthreads();
	int idx = threadIdx.x & 31;
	long long t0 = clock64();
	for (int i = 0; i < iters; i++) {
		if (OP == 0) a = fma(a, b, c);
		else if (OP == 1) a = __dadd_rn(a, c);
		else if (OP == 2) a = __dmul_rn(a, b);
		else if (OP == 3) a = __shfl_sync(0xffffffffu, a, (threadIdx.x + 1) & 31);
		else if (OP == 4) a = rsqrt(a) + 1.0;
		else if (OP == 5) a = exp(-a) + 0.5;
		else if (OP == 6) a = 1.0 / a + 0.5;
		else if (OP == 7) a = sqrt(a) + 0.5;
		else if (OP == 8) { idx = (int)s[idx] ; a += idx; }     // LDS + cvt dependent
		else if (OP == 9) a = erfc(a) + 0.5;
		else if (OP == 10) { float f = (float)a; f = fmaf(f, 1.0001f, 1e-3f); a = (double)f; }
		else if (OP == 11) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a)); a = r + 0.5; }
		else if (OP == 12) { double r; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a)); a = r + 0.5; }
		else if (OP == 13) a = a > 1.5 ? a - 0.5 : a + 0.25;   // DSETP + select chain
	}
	long long t1 = clock64();
	out[blockIdx.x * blockDim.x + threadIdx.x] = a + idx;
	if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// throughput: ILP independent chains per thread, MIX selects what is interleaved with the DFMAs
template <int ILP, int MIX>
__global__ void k_tput(double *out, int iters) {
	double a[ILP];
	float f[ILP];
	int n[ILP];
#pragma unroll
	for (int q = 0; q < ILP; q++) { a[q] = threadIdx.x * 1e-9 + q; f[q] = threadIdx.x + q; n[q] = threadIdx.x + q; }
	const double m = 1.0000001, b = 1e-7;
	for (int i = 0; i < iters; i++) {
#pragma unroll
		for (int q = 0; q < ILP; q++) {
			if (MIX == 0) a[q] = fma(a[q], m, b);
			else if (MIX == 1) a[q] = __dadd_rn(a[q], b);
			else if (MIX == 2) a[q] = __dmul_rn(a[q], m);
			else if (MIX == 3) { a[q] = fma(a[q], m, b); f[q] = fmaf(f[q], 1.0001f, 0.5f); }                 // 1 FFMA per DFMA
			else if (MIX == 4) { a[q] = fma(a[q], m, b); n[q] = n[q] * 3 + i; }                              // 1 IMAD per DFMA
			else if (MIX == 5) { a[q] = fma(a[q], m, b); f[q] = fmaf(f[q], 1.0001f, 0.5f); f[q] = fmaf(f[q], 0.999f, 0.25f); n[q] = n[q] * 3 + i; }  // 3 others per DFMA
			else if (MIX == 6) { a[q] = a[q] > 2.0 ? a[q] - 1.0 : a[q] + 0.5; }                              // DSETP+DADD+DADD+sel
		}
	}
	double s = 0;
#pragma unroll
	for (int q = 0; q < ILP; q++) s += a[q] + f[q] + n[q];
	out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP, int MIX>
int run_tput(const char *name, int sms, int threads, int ctas_per_sm, double *d) {
	const int iters = 4096;
	cudaEvent_t e0, e1;
	CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
	k_tput<ILP, MIX><<<sms * ctas_per_sm, threads>>>(d, iters);
	CK(cudaDeviceSynchronize());
	float best = 1e30f;
	for (int r = 0; r < 3; r++) {
		CK(cudaEventRecord(e0));
		k_tput<ILP, MIX><<<sms * ctas_per_sm, threads>>>(d, iters);
		CK(cudaEventRecord(e1));
		CK(cudaEventSynchronize(e1));
		float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
		if (ms < best) best = ms;
	}
	double ops = (double)iters * ILP * threads * ctas_per_sm * sms;
	printf("tput %-28s ILP %d warps/SM %2d : %8.3f ms  %7.2f T FP64-instr-lanes/s\n", name, ILP, threads / 32 * ctas_per_sm, best, ops / (best * 1e-3) / 1e12);
	return 0;
}

int main() {
	cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
	printf("%s, %d SMs, clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
	double *d; long long *c;
	CK(cudaMalloc(&d, sizeof(double) * 148 * 64 * 1024)); CK(cudaMalloc(&c, 8));
	const char *names[] = {"dfma", "dadd", "dmul", "shfl.f64", "rsqrt()+add", "exp()+add", "1/x+add", "sqrt()+add", "lds+cvt+add", "erfc()+add", "cvt f64->f32 ffma ->f64", "rcp.approx+add", "rsqrt.approx+add", "dsetp+sel chain"};
	const int iters = 2048;
	long long h;
#define LAT(OP) k_latency<OP><<<1, 32>>>(d, c, iters, 1.25); CK(cudaDeviceSynchronize()); k_latency<OP><<<1, 32>>>(d, c, iters, 1.25); CK(cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost)); printf("latency %-26s %7.1f cycles/iter\n", names[OP], (double)h / iters);
	LAT(0) LAT(1) LAT(2) LAT(3) LAT(4) LAT(5) LAT(6) LAT(7) LAT(8) LAT(9) LAT(10) LAT(11) LAT(12) LAT(13)
	const int S = p.multiProcessorCount;
	// FP64 pipe vs occupancy and ILP
	run_tput<1, 0>("dfma", S, 128, 1, d); run_tput<1, 0>("dfma", S, 256, 1, d); run_tput<1, 0>("dfma", S, 256, 2, d); run_tput<1, 0>("dfma", S, 256, 4, d); run_tput<1, 0>("dfma", S, 256, 8, d);
	run_tput<2, 0>("dfma", S, 128, 1, d); run_tput<4, 0>("dfma", S, 128, 1, d); run_tput<8, 0>("dfma", S, 128, 1, d);
	run_tput<4, 0>("dfma", S, 256, 2, d); run_tput<8, 0>("dfma", S, 256, 2, d); run_tput<8, 0>("dfma", S, 256, 8, d);
	run_tput<8, 1>("dadd", S, 256, 8, d); run_tput<8, 2>("dmul", S, 256, 8, d);
	run_tput<8, 3>("dfma+ffma", S, 256, 8, d); run_tput<8, 4>("dfma+imad", S, 256, 8, d); run_tput<8, 5>("dfma+2ffma+imad", S, 256, 8, d);
	run_tput<8, 6>("dsetp+2dadd+sel (per 1 count)", S, 256, 8, d);
	run_tput<4, 3>("dfma+ffma", S, 256, 2, d); run_tput<4, 5>("dfma+2ffma+imad", S, 256, 2, d);
	return 0;
}
