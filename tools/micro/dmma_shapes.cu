// dmma_shapes.cu -- which FP64 mma.sync shapes does sm_100a run, with which fragment layout, and how fast?
// (m8n8k4 is the sm_80 shape; m16n8k4 / k8 / k16 are the sm_90+ shapes.) Prints the max error of each layout
// hypothesis against a host product, then the whole-GPU throughput and the dependent-chain latency of each shape.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma884(double &c0, double &c1, double a, double b)
{
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma1684(double *c, const double *a, const double *b)
{
	asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
		: "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
}
__device__ __forceinline__ void mma1688(double *c, const double *a, const double *b)
{
	asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
		: "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void mma16816(double *c, const double *a, const double *b)
{
	asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
		: "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
		: "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// layout hypothesis H (0 / 1) for A: register i of K/2 registers
//  H0: row = g + 8 * (i & 1), col = t + 4 * (i >> 1)        H1: row = g + 8 * (i / (K/4)), col = t + 4 * (i % (K/4))
template <int K>
__global__ void k_check(const double *A /*16xK row-major*/, const double *B /*Kx8 row-major*/, double *C /*16x8*/, int H)
{
	const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
	double a[K / 2], b[K / 4], c[4] = {0, 0, 0, 0};
	for(int i = 0; i < K / 2; ++ i) {
		int row = H? g + 8 * (i / (K / 4)) : g + 8 * (i & 1), col = H? t + 4 * (i % (K / 4)) : t + 4 * (i >> 1);
		a[i] = A[row * K + col];
	}
	for(int i = 0; i < K / 4; ++ i)
		b[i] = B[(t + 4 * i) * 8 + g];
	if(K == 4) mma1684(c, a, b);
	if(K == 8) mma1688(c, a, b);
	if(K == 16) mma16816(c, a, b);
	C[g * 8 + 2 * t] = c[0]; C[g * 8 + 2 * t + 1] = c[1];
	C[(g + 8) * 8 + 2 * t] = c[2]; C[(g + 8) * 8 + 2 * t + 1] = c[3];
}

template <int SHAPE> // 0: m8n8k4, 1: m16n8k4, 2: m16n8k8, 3: m16n8k16; 8 independent accumulators per warp
__global__ void k_tput(double *out, int n)
{
	double c[8][4], a[8], b[4];
	for(int i = 0; i < 8; ++ i) { a[i] = threadIdx.x * 1e-3 + i; for(int j = 0; j < 4; ++ j) c[i][j] = 0; }
	for(int i = 0; i < 4; ++ i) b[i] = 1e-3 * i + 0.5;
	for(int it = 0; it < n; ++ it) {
		#pragma unroll
		for(int i = 0; i < 8; ++ i) {
			if(SHAPE == 0) mma884(c[i][0], c[i][1], a[i], b[0]);
			if(SHAPE == 1) mma1684(c[i], a, b);
			if(SHAPE == 2) mma1688(c[i], a, b);
			if(SHAPE == 3) mma16816(c[i], a, b);
		}
	}
	double s = 0;
	for(int i = 0; i < 8; ++ i) for(int j = 0; j < 4; ++ j) s += c[i][j];
	if(s == 123.456) out[0] = s;
}

template <int SHAPE>
__global__ void k_lat(double *out, int n) // one dependent chain
{
	double c[4] = {0, 0, 0, 0}, a[8], b[4];
	for(int i = 0; i < 8; ++ i) a[i] = threadIdx.x * 1e-3 + i;
	for(int i = 0; i < 4; ++ i) b[i] = 1e-3 * i + 0.5;
	long long t0 = clock64();
	for(int it = 0; it < n; ++ it) {
		if(SHAPE == 0) mma884(c[0], c[1], a[0], b[0]);
		if(SHAPE == 1) mma1684(c, a, b);
		if(SHAPE == 2) mma1688(c, a, b);
		if(SHAPE == 3) mma16816(c, a, b);
	}
	long long t1 = clock64();
	if(threadIdx.x == 0) out[1] = double(t1 - t0) / n;
	if(c[0] + c[1] + c[2] + c[3] == 123.456) out[0] = c[0];
}

template <int K> void check()
{
	double hA[16 * K], hB[K * 8], hC[128], ref[128];
	for(int i = 0; i < 16 * K; ++ i) hA[i] = rand() / double(RAND_MAX) - 0.5;
	for(int i = 0; i < K * 8; ++ i) hB[i] = rand() / double(RAND_MAX) - 0.5;
	for(int i = 0; i < 16; ++ i) for(int j = 0; j < 8; ++ j) { double s = 0; for(int k = 0; k < K; ++ k) s += hA[i * K + k] * hB[k * 8 + j]; ref[i * 8 + j] = s; }
	double *dA, *dB, *dC;
	cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dC, sizeof(hC));
	cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
	for(int H = 0; H < 2; ++ H) {
		k_check<K><<<1, 32>>>(dA, dB, dC, H);
		cudaError_t e = cudaDeviceSynchronize();
		cudaMemcpy(hC, dC, sizeof(hC), cudaMemcpyDeviceToHost);
		double err = 0; for(int i = 0; i < 128; ++ i) err = fmax(err, fabs(hC[i] - ref[i]));
		printf("m16n8k%-2d layout H%d: max err %.3e  (%s)\n", K, H, err, cudaGetErrorString(e));
	}
}

template <int SHAPE> void tput(const char *name, double flops_per_mma)
{
	double *d; cudaMalloc(&d, 64);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	const int n = 20000;
	for(int warps : {4, 8, 16}) {
		k_tput<SHAPE><<<148 * 2, warps * 32>>>(d, 100);
		cudaEventRecord(e0);
		k_tput<SHAPE><<<148 * 2, warps * 32>>>(d, n);
		cudaEventRecord(e1); cudaEventSynchronize(e1);
		float ms; cudaEventElapsedTime(&ms, e0, e1);
		double fl = 148.0 * 2 * warps * 8.0 * n * flops_per_mma;
		printf("%-10s %2d warps/CTA x 2 CTA/SM: %.2f TFLOP/s\n", name, warps, fl / (ms * 1e-3) / 1e12);
	}
	k_lat<SHAPE><<<1, 32>>>(d, 4096); cudaDeviceSynchronize();
	double h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
	printf("%-10s dependent-chain latency: %.1f cycles\n", name, h[1]);
}

int main()
{
	check<4>(); check<8>(); check<16>();
	tput<0>("m8n8k4", 512); tput<1>("m16n8k4", 1024); tput<2>("m16n8k8", 2048); tput<3>("m16n8k16", 4096);
	return 0;
}
