// dense_chol.cu -- stage 3 of the hot path for a dense reduced camera system: FP64 Cholesky S = R^T R of
// the upper triangle, and the two triangular solves.
//
// Reference function replaced: CLinearSolver_DenseEigen::Solve_PosDef
// (src/slam/LinearSolver_Schur.cpp:2314-2333) = Convert_to_Dense + Eigen::LLT<MatrixXd, Eigen::Upper>::compute
// (reads the upper triangle only, fails on a non-positive pivot) + LLT::solve (two triangular solves).
//
// Layout: column-major, leading dimension ld = n rounded up to a multiple of the panel width; the padding
// carries an identity diagonal so that no kernel needs bounds checks. The right-hand side rides along as an
// extra column block right of the matrix: the panel solve and the trailing update turn it into
// y = R^-T b for free (augmented-matrix trick), which leaves one backward solve R x = y.
//
// Blocked right-looking factorisation, panel width CH_NB:
//   k_potrf_diag   one CTA, diagonal block in shared memory
//   k_trsm_panel   R_kk^-T applied to the block row right of the diagonal (thread per column)
//   k_syrk_update  trailing update C -= P^T P on FP64 tensor cores (mma.sync m8n8k4 DMMA), 64x64 CTA tiles
// Backward solve: k_invert_diag (all diagonal blocks at once, off the critical path) + k_backsolve, one
// persistent CTA per block row that consumes x_j as soon as block row j publishes it.

#include "spp_ctx.h"

namespace spp {

#define LAUNCH_CHECK(ctx) do { ++ (ctx)->n_launches; SPP_CUDA(cudaGetLastError()); } while(0)

#define CH_NB 64          // panel width
#define CH_TILE 64        // CTA tile of the trailing update
#define CH_BK 16          // k-chunk staged in shared memory
#define CH_LDS (CH_BK + 4) // padded row length: conflict-free DMMA fragment loads

// ---- diagonal block ------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) k_potrf_diag(double *__restrict__ A, size_t ld, size_t k0, int *__restrict__ info)
{
	__shared__ double T[CH_NB][CH_NB + 1];
	double *Akk = A + k0 * ld + k0;
	for(int idx = threadIdx.x; idx < CH_NB * CH_NB; idx += 256) {
		int c = idx / CH_NB, r = idx % CH_NB;
		T[r][c] = (r <= c)? Akk[(size_t)c * ld + r] : 0.0;
	}
	__syncthreads();
	for(int j = 0; j < CH_NB; ++ j) {
		if(threadIdx.x == 0) {
			double d = T[j][j];
			if(!(d > 0)) { // Eigen's LLT stops at a non-positive pivot (and so does a NaN)
				if(*info == 0)
					*info = int(k0) + j + 1;
				d = 1;
			}
			T[j][j] = sqrt(d);
		}
		__syncthreads();
		const double djj = T[j][j];
		for(int c = j + 1 + threadIdx.x; c < CH_NB; c += 256)
			T[j][c] /= djj;
		__syncthreads();
		const int m = CH_NB - 1 - j;
		for(int idx = threadIdx.x; idx < m * m; idx += 256) {
			int r = j + 1 + idx / m, c = j + 1 + idx % m;
			if(r <= c)
				T[r][c] -= T[j][r] * T[j][c];
		}
		__syncthreads();
	}
	for(int idx = threadIdx.x; idx < CH_NB * CH_NB; idx += 256) {
		int c = idx / CH_NB, r = idx % CH_NB;
		if(r <= c)
			Akk[(size_t)c * ld + r] = T[r][c];
	}
}

// ---- block row: solve R_kk^T X = A(k-block, columns right of it) -------------------------------------

__global__ void __launch_bounds__(64) k_trsm_panel(double *__restrict__ A, size_t ld, size_t k0, size_t c0, size_t n_cols)
{
	__shared__ double R[CH_NB][CH_NB + 1]; // R[j][i], upper
	const double *Akk = A + k0 * ld + k0;
	for(int idx = threadIdx.x; idx < CH_NB * CH_NB; idx += 64) {
		int c = idx / CH_NB, r = idx % CH_NB;
		R[r][c] = (r <= c)? Akk[(size_t)c * ld + r] : 0.0;
	}
	__syncthreads();
	size_t col = c0 + blockIdx.x * (size_t)64 + threadIdx.x;
	if(col >= n_cols)
		return;
	double *p = A + col * ld + k0;
	double x[CH_NB];
	#pragma unroll
	for(int i = 0; i < CH_NB; i += 2) {
		double2 t = *reinterpret_cast<const double2*>(p + i);
		x[i] = t.x; x[i + 1] = t.y;
	}
	#pragma unroll
	for(int j = 0; j < CH_NB; ++ j) {
		const double yj = x[j] / R[j][j];
		x[j] = yj;
		#pragma unroll
		for(int i = j + 1; i < CH_NB; ++ i)
			x[i] -= R[j][i] * yj;
	}
	#pragma unroll
	for(int i = 0; i < CH_NB; i += 2)
		*reinterpret_cast<double2*>(p + i) = make_double2(x[i], x[i + 1]);
}

// ---- trailing update on the FP64 tensor cores ------------------------------------------------------

__device__ __forceinline__ void dmma_m8n8k4(double &d0, double &d1, double a, double b)
{
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
		: "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// C(i0.., j0..) -= P(:, i0..)^T P(:, j0..) for the upper tiles; P = rows k0 .. k0+CH_NB of A.
// grid: x = column tile (covers the matrix columns right of the panel and the rhs block), y = row tile.
__global__ void __launch_bounds__(128) k_syrk_update(double *__restrict__ A, size_t ld, size_t k0, size_t c0,
	size_t n_rows_end)
{
	const size_t i0 = c0 + blockIdx.y * (size_t)CH_TILE, j0 = c0 + blockIdx.x * (size_t)CH_TILE;
	if(i0 > j0 || i0 >= n_rows_end)
		return;
	__shared__ double As[CH_TILE][CH_LDS];
	__shared__ double Bs[CH_TILE][CH_LDS];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int wi = (warp >> 1) * 32, wj = (warp & 1) * 32; // warp tile origin inside the CTA tile
	const int g = lane >> 2, t = lane & 3;
	double acc[4][4][2];
	#pragma unroll
	for(int a = 0; a < 4; ++ a)
		#pragma unroll
		for(int b = 0; b < 4; ++ b)
			acc[a][b][0] = acc[a][b][1] = 0;
	// staging: thread -> (column, half of the 16-deep chunk)
	const int lc = threadIdx.x >> 1, lh = (threadIdx.x & 1) * 8;
	const double *pa = A + (i0 + lc) * ld + k0 + lh;
	const double *pb = A + (j0 + lc) * ld + k0 + lh;
	for(int kk = 0; kk < CH_NB; kk += CH_BK) {
		#pragma unroll
		for(int q = 0; q < 8; q += 2) {
			double2 va = *reinterpret_cast<const double2*>(pa + kk + q);
			double2 vb = *reinterpret_cast<const double2*>(pb + kk + q);
			*reinterpret_cast<double2*>(&As[lc][lh + q]) = va;
			*reinterpret_cast<double2*>(&Bs[lc][lh + q]) = vb;
		}
		__syncthreads();
		#pragma unroll
		for(int k4 = 0; k4 < CH_BK; k4 += 4) {
			double fa[4], fb[4];
			#pragma unroll
			for(int a = 0; a < 4; ++ a)
				fa[a] = As[wi + a * 8 + g][k4 + t];
			#pragma unroll
			for(int b = 0; b < 4; ++ b)
				fb[b] = Bs[wj + b * 8 + g][k4 + t];
			#pragma unroll
			for(int a = 0; a < 4; ++ a)
				#pragma unroll
				for(int b = 0; b < 4; ++ b)
					dmma_m8n8k4(acc[a][b][0], acc[a][b][1], fa[a], fb[b]);
		}
		__syncthreads();
	}
	#pragma unroll
	for(int a = 0; a < 4; ++ a) {
		#pragma unroll
		for(int b = 0; b < 4; ++ b) {
			const size_t r = i0 + wi + a * 8 + g, c = j0 + wj + b * 8 + 2 * t;
			A[c * ld + r] -= acc[a][b][0];
			A[(c + 1) * ld + r] -= acc[a][b][1];
		}
	}
}

// ---- backward solve ---------------------------------------------------------------------------------

// Rinv[blk] = inverse of the upper-triangular diagonal block blk (column-major CH_NB x CH_NB)
__global__ void __launch_bounds__(CH_NB) k_invert_diag(const double *__restrict__ A, size_t ld, double *__restrict__ Rinv)
{
	__shared__ double X[CH_NB][CH_NB + 1]; // X[i][c]; thread c owns column c
	const size_t k0 = blockIdx.x * (size_t)CH_NB;
	const double *Akk = A + k0 * ld + k0; // R(i, j) = Akk[j * ld + i]; all threads read the same element (broadcast)
	const int c = threadIdx.x;
	for(int i = CH_NB - 1; i >= 0; -- i) {
		double v = 0;
		const double rii = Akk[(size_t)i * ld + i];
		if(i == c)
			v = 1.0 / rii;
		else if(i < c) {
			double s = 0;
			for(int j = i + 1; j <= c; ++ j)
				s += Akk[(size_t)j * ld + i] * X[j][c];
			v = -s / rii;
		}
		X[i][c] = v;
	}
	__syncthreads();
	double *out = Rinv + blockIdx.x * (size_t)(CH_NB * CH_NB);
	for(int idx = threadIdx.x; idx < CH_NB * CH_NB; idx += CH_NB) {
		int cc = idx / CH_NB, r = idx % CH_NB;
		out[idx] = X[r][cc];
	}
}

// R x = y. One CTA per block row (blockIdx 0 = last block row), CH_NB threads, thread r owns row r.
__global__ void __launch_bounds__(CH_NB) k_backsolve(const double *__restrict__ A, size_t ld, size_t n_blk,
	const double *__restrict__ Rinv, double *__restrict__ y /* in: y, out: x */, volatile int *flags)
{
	__shared__ double xs[CH_NB];
	const size_t bi = n_blk - 1 - blockIdx.x;
	const int r = threadIdx.x;
	double acc = y[bi * CH_NB + r];
	for(size_t bj = n_blk - 1; bj > bi; -- bj) {
		if(threadIdx.x == 0) {
			while(flags[bj] == 0)
				;
			__threadfence();
		}
		__syncthreads();
		xs[r] = ((volatile double*)y)[bj * CH_NB + r];
		__syncthreads();
		const double *row = A + (bj * CH_NB) * ld + bi * CH_NB + r;
		double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
		#pragma unroll 4
		for(int c = 0; c < CH_NB; c += 4) {
			s0 += row[(size_t)c * ld] * xs[c];
			s1 += row[(size_t)(c + 1) * ld] * xs[c + 1];
			s2 += row[(size_t)(c + 2) * ld] * xs[c + 2];
			s3 += row[(size_t)(c + 3) * ld] * xs[c + 3];
		}
		acc -= (s0 + s1) + (s2 + s3);
		__syncthreads();
	}
	xs[r] = acc;
	__syncthreads();
	const double *Ri = Rinv + bi * (size_t)(CH_NB * CH_NB);
	double x = 0;
	for(int c = r; c < CH_NB; ++ c) // upper triangular inverse
		x += Ri[(size_t)c * CH_NB + r] * xs[c];
	y[bi * CH_NB + r] = x;
	__threadfence();
	__syncthreads();
	if(threadIdx.x == 0)
		flags[bi] = 1;
}

__global__ void k_pad_identity(double *__restrict__ A, size_t ld, size_t n, size_t n_pad)
{
	size_t i = n + blockIdx.x * (size_t)blockDim.x + threadIdx.x;
	if(i < n_pad)
		A[i * ld + i] = 1.0;
}

__global__ void k_copy_rhs(double *__restrict__ dst, const double *__restrict__ src, size_t n)
{
	size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
	if(i < n)
		dst[i] = src[i];
}

// ---------------------------------------------------------------------------------------------------

size_t dense_chol_ld(size_t n)
{
	return (n + CH_NB - 1) / CH_NB * CH_NB;
}

// number of doubles of the augmented storage: ld x (ld + CH_TILE)
size_t dense_chol_storage(size_t n)
{
	size_t ld = dense_chol_ld(n);
	return ld * (ld + CH_TILE);
}

// A: device, augmented storage (upper triangle of the n x n matrix filled, everything else zero, rhs NOT yet
// placed). d_rhs_x: device vector of n doubles, rhs in, solution out. Returns SPP_OK / SPP_NOT_POSDEF.
int dense_chol_solve_device(spp_ctx *ctx, double *A, size_t n, double *d_rhs_x)
{
	DenseChol &ch = ctx->chol;
	const size_t ld = dense_chol_ld(n), n_blk = ld / CH_NB;
	const size_t n_cols = ld + CH_TILE; // matrix + rhs block
	cudaStream_t st = ctx->stream;
	ch.info.resize(1 + n_blk);
	SPP_CUDA(cudaMemsetAsync(ch.info.p(), 0, (1 + n_blk) * sizeof(int), st));
	ch.work.resize(n_blk * CH_NB * CH_NB);
	if(ld > n) {
		k_pad_identity<<<n_blocks(ld - n, 64), 64, 0, st>>>(A, ld, n, ld);
		LAUNCH_CHECK(ctx);
	}
	double *rhs_col = A + ld * ld; // first column of the rhs block
	k_copy_rhs<<<n_blocks(n, 256), 256, 0, st>>>(rhs_col, d_rhs_x, n);
	LAUNCH_CHECK(ctx);
	for(size_t b = 0; b < n_blk; ++ b) {
		const size_t k0 = b * CH_NB, c0 = k0 + CH_NB;
		k_potrf_diag<<<1, 256, 0, st>>>(A, ld, k0, ch.info.p());
		LAUNCH_CHECK(ctx);
		// block row right of the diagonal, including the rhs column (only its first column matters)
		const size_t n_trsm_cols = ld + 1;
		k_trsm_panel<<<n_blocks(n_trsm_cols - c0, 64), 64, 0, st>>>(A, ld, k0, c0, n_trsm_cols);
		LAUNCH_CHECK(ctx);
		if(c0 < ld) {
			dim3 grid((unsigned)((n_cols - c0) / CH_TILE), (unsigned)((ld - c0) / CH_TILE));
			k_syrk_update<<<grid, 128, 0, st>>>(A, ld, k0, c0, ld);
			LAUNCH_CHECK(ctx);
		}
	}
	k_invert_diag<<<(unsigned)n_blk, CH_NB, 0, st>>>(A, ld, ch.work.p());
	LAUNCH_CHECK(ctx);
	k_backsolve<<<(unsigned)n_blk, CH_NB, 0, st>>>(A, ld, n_blk, ch.work.p(), rhs_col, ch.info.p() + 1);
	LAUNCH_CHECK(ctx);
	k_copy_rhs<<<n_blocks(n, 256), 256, 0, st>>>(d_rhs_x, rhs_col, n);
	LAUNCH_CHECK(ctx);
	ctx->h_scalars.resize(16);
	int *h_info = reinterpret_cast<int*>(ctx->h_scalars.p());
	SPP_CUDA(cudaMemcpyAsync(h_info, ch.info.p(), sizeof(int), cudaMemcpyDeviceToHost, st));
	SPP_CUDA(cudaStreamSynchronize(st));
	return (*h_info == 0)? SPP_OK : SPP_NOT_POSDEF;
}

} // namespace spp
