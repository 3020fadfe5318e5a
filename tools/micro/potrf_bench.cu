// potrf_bench.cu -- times k_potrf128 (slam_plus_plus_b200/csrc/potrf128.cuh) in isolation on one SPD 128 x 128
// block, checks the factor and its inverse against a host Cholesky, and prints per-warp clock64 stamps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -DPOTRF_TRACE -o potrf_bench potrf_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#define CH_NB 128
namespace spp {
#include "../../slam_plus_plus_b200/csrc/potrf128.cuh"
}
using namespace spp;

int main()
{
	const int n = CH_NB;
	std::vector<double> M(n * n), A(n * n), R(n * n, 0.0);
	srand(7);
	for(auto &v : M) v = rand() / double(RAND_MAX) - 0.5;
	for(int i = 0; i < n; ++ i) for(int j = 0; j < n; ++ j) {
		double s = (i == j)? 1.0 : 0.0;
		for(int k = 0; k < n; ++ k) s += M[i * n + k] * M[j * n + k];
		A[j * n + i] = s;
	}
	// host upper Cholesky (column-major), A = R^T R
	std::vector<double> T(A);
	for(int j = 0; j < n; ++ j) {
		for(int i = 0; i <= j; ++ i) {
			double s = T[j * n + i];
			for(int k = 0; k < i; ++ k) s -= R[i * n + k] * R[j * n + k];
			R[j * n + i] = (i == j)? sqrt(s) : s / R[i * n + i];
		}
	}
	double *dA, *dA0, *dX; int *dinfo; long long *ddbg;
	cudaMalloc(&dA, n * n * 8); cudaMalloc(&dA0, n * n * 8); cudaMalloc(&dX, n * n * 8); cudaMalloc(&dinfo, 4); cudaMalloc(&ddbg, (16 + 8 * 64) * 8);
	cudaMemcpy(dA0, A.data(), n * n * 8, cudaMemcpyHostToDevice);
	cudaMemset(dX, 0, n * n * 8); cudaMemset(dinfo, 0, 4); cudaMemset(ddbg, 0, (16 + 8 * 64) * 8);
	cudaFuncSetAttribute(k_potrf128, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POTRF_SMEM);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	float best = 1e9;
	for(int it = 0; it < 10; ++ it) {
		cudaMemcpy(dA, dA0, n * n * 8, cudaMemcpyDeviceToDevice);
		cudaEventRecord(e0);
		k_potrf128<<<1, PT, POTRF_SMEM>>>(dA, n, 0, dX, dinfo, ddbg);
		cudaEventRecord(e1); cudaEventSynchronize(e1);
		float ms; cudaEventElapsedTime(&ms, e0, e1); if(ms < best) best = ms;
	}
	printf("launch: %s; best of 10: %.2f us\n", cudaGetErrorString(cudaGetLastError()), best * 1e3);
	std::vector<double> hR(n * n), hX(n * n);
	cudaMemcpy(hR.data(), dA, n * n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(hX.data(), dX, n * n * 8, cudaMemcpyDeviceToHost);
	double eR = 0, eI = 0;
	for(int j = 0; j < n; ++ j) for(int i = 0; i <= j; ++ i) eR = fmax(eR, fabs(hR[j * n + i] - R[j * n + i]));
	for(int i = 0; i < n; ++ i) for(int j = 0; j < n; ++ j) { // R * X = I
		double s = 0; for(int k = 0; k < n; ++ k) s += ((i <= k)? R[k * n + i] : 0.0) * ((k <= j)? hX[j * n + k] : 0.0);
		eI = fmax(eI, fabs(s - (i == j)));
	}
	int info; cudaMemcpy(&info, dinfo, 4, cudaMemcpyDeviceToHost);
	printf("max |R - R_host| = %.3e, max |R X - I| = %.3e, info = %d\n", eR, eI, info);
	std::vector<long long> h(16 + 8 * 64);
	cudaMemcpy(h.data(), ddbg, h.size() * 8, cudaMemcpyDeviceToHost);
	printf("thread-0 marks: load %lld | step0 %lld | step1 %lld | steps0-4 %lld | factor %lld | store %lld | inverse %lld | storeinv %lld | total %lld\n",
		h[1] - h[0], h[2] - h[1], h[3] - h[2], h[4] - h[1], h[5] - h[1], h[6] - h[5], h[7] - h[6], h[8] - h[7], h[8] - h[0]);
	const char *names1[5] = {"top", "urgent+bar", "leaf factored", "rowsolve|lazy", "sync"};
	for(int s : {2, 4}) {
		printf("step s=%d (deltas to the previous stamp of the same warp; warps 0..7)\n", s);
		for(int m = 1; m < 5; ++ m) {
			printf("  %-14s", names1[m]);
			for(int w = 0; w < 8; ++ w) {
				long long a = h[16 + w * 64 + s * 8 + m], b = 0;
				for(int q = m - 1; q >= 0 && !b; -- q) b = h[16 + w * 64 + s * 8 + q];
				printf(" %6lld", a? a - b : 0LL);
			}
			printf("\n");
		}
	}
	for(int s : {2, 4}) {
		printf("step s=%d: leaf store (w3), urgent update, barrier wait:\n", s);
		for(int w = 0; w < 4; ++ w) {
			long long *b = &h[16 + w * 64 + s * 8];
			printf("   warp %d: store %lld  urgent %lld  barrier %lld\n", w, b[5] - b[0], b[6] - b[5], b[1] - b[6]);
		}
	}
	printf("L0 store alone (warps 0..7):");
	for(int w = 0; w < 8; ++ w) printf(" %lld", h[16 + w * 64 + 47] - h[16 + w * 64 + 41]);
	printf("\n");
	const char *names2[7] = {"L0 computed", "sync", "L0 published", "level 8", "level 16", "level 32", "level 64"};
	printf("inverse (deltas; warps 0..7)\n");
	for(int m = 0; m < 7; ++ m) {
		printf("  %-12s", names2[m]);
		for(int w = 0; w < 8; ++ w) {
			long long a = h[16 + w * 64 + 40 + m], b = m? h[16 + w * 64 + 40 + m - 1] : h[6];
			printf(" %6lld", a - b);
		}
		printf("\n");
	}
	return 0;
}
