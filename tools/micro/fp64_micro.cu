// fp64_micro.cu -- microbenchmarks: DFMA latency / throughput, rsqrt, shuffle, LDS on this GPU
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dfma_lat(double *out, int n) {
	double a = threadIdx.x * 1e-9 + 1.0, b = 1.0000001, c = 1e-7;
	long long t0 = clock64();
	for(int i = 0; i < n; ++ i) { a = a * b + c; a = a * b + c; a = a * b + c; a = a * b + c; }
	long long t1 = clock64();
	if(threadIdx.x == 0 && blockIdx.x == 0) { out[0] = double(t1 - t0) / (4.0 * n); }
	if(a == 123.456) out[1] = a;
}
__global__ void k_dfma_tput(double *out, int n) {
	double a0 = threadIdx.x * 1e-9 + 1.0, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
	double b = 1.0000001, c = 1e-7;
	long long t0 = clock64();
	for(int i = 0; i < n; ++ i) { a0 = a0 * b + c; a1 = a1 * b + c; a2 = a2 * b + c; a3 = a3 * b + c; a4 = a4 * b + c; a5 = a5 * b + c; a6 = a6 * b + c; a7 = a7 * b + c; }
	long long t1 = clock64();
	if(threadIdx.x == 0 && blockIdx.x == 0) out[0] = double(t1 - t0) / (8.0 * n);
	if(a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 123.456) out[1] = a0;
}
__global__ void k_rsqrt_lat(double *out, int n) {
	double a = threadIdx.x * 1e-3 + 2.0;
	long long t0 = clock64();
	for(int i = 0; i < n; ++ i) { a = rsqrt(a) + 1.5; }
	long long t1 = clock64();
	if(threadIdx.x == 0 && blockIdx.x == 0) out[0] = double(t1 - t0) / n;
	if(a == 123.456) out[1] = a;
}
__global__ void k_sqrt_div_lat(double *out, int n) {
	double a = threadIdx.x * 1e-3 + 2.0;
	long long t0 = clock64();
	for(int i = 0; i < n; ++ i) { a = 1.0 / sqrt(a) + 1.5; }
	long long t1 = clock64();
	if(threadIdx.x == 0 && blockIdx.x == 0) out[0] = double(t1 - t0) / n;
	if(a == 123.456) out[1] = a;
}
__global__ void k_shfl_lat(double *out, int n) {
	double a = threadIdx.x;
	long long t0 = clock64();
	for(int i = 0; i < n; ++ i) { a = __shfl_sync(0xffffffffu, a, (threadIdx.x + 1) & 31); }
	long long t1 = clock64();
	if(threadIdx.x == 0 && blockIdx.x == 0) out[0] = double(t1 - t0) / n;
	if(a == 123.456) out[1] = a;
}
__global__ void k_lds_fma(double *out, int n) {
	__shared__ double s[1024];
	for(int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = i * 1e-6;
	__syncthreads();
	double x[16];
	for(int i = 0; i < 16; ++ i) x[i] = i + threadIdx.x;
	long long t0 = clock64();
	for(int it = 0; it < n; ++ it) {
		#pragma unroll
		for(int i = 0; i < 16; ++ i) x[i] -= s[(it * 16 + i) & 1023] * 1.0000001;
	}
	long long t1 = clock64();
	double sum = 0; for(int i = 0; i < 16; ++ i) sum += x[i];
	if(threadIdx.x == 0 && blockIdx.x == 0) out[0] = double(t1 - t0) / (16.0 * n);
	if(sum == 123.456) out[1] = sum;
}
int main() {
	double *d; cudaMalloc(&d, 64); double h[2];
	const int n = 4096;
	struct { const char *name; void (*k)(double*, int); int threads; } tests[] = {
		{"DFMA dependent latency (cycles)", k_dfma_lat, 32}, {"DFMA 8 independent chains, 1 warp (cycles/FMA)", k_dfma_tput, 32},
		{"DFMA 8 chains, 4 warps/SM (cycles/FMA/warp)", k_dfma_tput, 128}, {"DFMA 8 chains, 16 warps/SM", k_dfma_tput, 512},
		{"DFMA 8 chains, 32 warps/SM", k_dfma_tput, 1024},
		{"rsqrt(double)+add dependent (cycles)", k_rsqrt_lat, 32}, {"1/sqrt(double)+add dependent (cycles)", k_sqrt_div_lat, 32},
		{"64-bit shuffle dependent (cycles)", k_shfl_lat, 32}, {"LDS broadcast + DFMA, 16 independent, 1 warp (cycles/FMA)", k_lds_fma, 32},
		{"LDS broadcast + DFMA, 8 warps", k_lds_fma, 256},
	};
	for(auto &t : tests) {
		t.k<<<1, t.threads>>>(d, n); cudaDeviceSynchronize();
		cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
		printf("%-62s %8.2f\n", t.name, h[0]);
	}
	return 0;
}
