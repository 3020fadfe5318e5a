// schur_kernels.cu -- stage 2 and 4 of the hot path: landmark-side Schur complement and back-substitution.
//
// Reference functions replaced (all inside
// CLinearSolver_Schur::Solve_PosDef_Blocky, include/slam/LinearSolver_Schur.h:1623-1935):
//   Permute_UpperTriangular_To / SliceTo / TransposeTo  Schur.h:1687-1709   -> no-ops here: the system is
//                                                        stored as (U, V, W) in Schur order from the start
//   InverseOf_BlockDiag_FBS_Parallel + Scale(-1)        Schur.h:1720-1735, BlockMatrixFBS.inl:1749-1868
//   U.MultiplyToWith_FBS(C^-1)                          Schur.h:1743-1745   (Y = W C^-1, per observation)
//   (-U C^-1).MultiplyToWith_FBS(V, upper) + AddTo_FBS  Schur.h:1757-1767   (S = A - Y W^T, upper blocks)
//   PreMultiply_Add_FBS (rhs)                           Schur.h:1829-1830   (b = x - Y l)
//   PostMultiply_Add_FBS_Parallel + PreMultiply_Add     Schur.h:1867-1881   (dl = C^-1 (l - W^T dx))
//
// The product S_ij = A_ij - sum_p Y_ip W_jp^T is evaluated per destination block from a precomputed list of
// observation pairs sorted by landmark -- the reference's own accumulation order (BlockMatrixFBS.h:395-448)
// -- so there are no floating-point atomics and the result is bit-reproducible.

#include "spp_ctx.h"

namespace spp {

size_t dense_chol_ld(size_t n);
size_t dense_chol_storage(size_t n);

#define LAUNCH_CHECK(ctx) do { ++ (ctx)->n_launches; SPP_CUDA(cudaGetLastError()); } while(0)

// thread per landmark: Cinv = (V + alpha I)^-1 by cofactors (what Eigen's fixed 3x3 inverse() does,
// BlockMatrixBase.h:1256-1270), then Y_o = W_o Cinv along the track
__global__ void k_landmark_inverse(size_t P, double alpha, const uint32_t *__restrict__ pt_ptr,
	const double *__restrict__ V, const double *__restrict__ W, double *__restrict__ Cinv, double *__restrict__ Y)
{
	size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
	if(p >= P) return;
	const double *Vp = V + p * 9; // column-major
	double m00 = Vp[0] + alpha, m10 = Vp[1], m20 = Vp[2];
	double m01 = Vp[3], m11 = Vp[4] + alpha, m21 = Vp[5];
	double m02 = Vp[6], m12 = Vp[7], m22 = Vp[8] + alpha;
	// cofactors
	double c00 = m11 * m22 - m12 * m21, c10 = m21 * m02 - m22 * m01, c20 = m01 * m12 - m02 * m11;
	double det = c00 * m00 + c10 * m10 + c20 * m20;
	double id = 1.0 / det;
	double i00 = c00 * id, i01 = c10 * id, i02 = c20 * id;
	double i10 = (m12 * m20 - m10 * m22) * id, i11 = (m22 * m00 - m20 * m02) * id, i12 = (m02 * m10 - m00 * m12) * id;
	double i20 = (m10 * m21 - m11 * m20) * id, i21 = (m20 * m01 - m21 * m00) * id, i22 = (m00 * m11 - m01 * m10) * id;
	double *Ci = Cinv + p * 9; // column-major
	Ci[0] = i00; Ci[1] = i10; Ci[2] = i20;
	Ci[3] = i01; Ci[4] = i11; Ci[5] = i21;
	Ci[6] = i02; Ci[7] = i12; Ci[8] = i22;
	const unsigned beg = pt_ptr[p], end = pt_ptr[p + 1];
	for(unsigned o = beg; o < end; ++ o) {
		const double *Wo = W + (size_t)o * 18;
		double *Yo = Y + (size_t)o * 18;
		double w[18];
		#pragma unroll
		for(int i = 0; i < 18; i += 2) {
			double2 t = *reinterpret_cast<const double2*>(Wo + i);
			w[i] = t.x; w[i + 1] = t.y;
		}
		double y[18];
		#pragma unroll
		for(int r = 0; r < 6; ++ r) {
			y[r] = w[r] * i00 + w[6 + r] * i10 + w[12 + r] * i20;
			y[6 + r] = w[r] * i01 + w[6 + r] * i11 + w[12 + r] * i21;
			y[12 + r] = w[r] * i02 + w[6 + r] * i12 + w[12 + r] * i22;
		}
		#pragma unroll
		for(int i = 0; i < 18; i += 2)
			*reinterpret_cast<double2*>(Yo + i) = make_double2(y[i], y[i + 1]);
	}
}

// one warp per upper-triangular 6x6 block of the reduced camera system.
// S(i,j) = [i == j] (U_i + alpha I) - sum_pairs Y_a W_b^T ; for i == j also b_i = gc_i - sum_a Y_a gp_{p(a)}
#define SB_WARPS 8

__global__ void __launch_bounds__(SB_WARPS * 32) k_schur_blocks(size_t n_blocks_total, size_t ld, double alpha,
	const uint32_t *__restrict__ blk_row, const uint32_t *__restrict__ blk_col, const uint64_t *__restrict__ blk_ptr,
	const uint32_t *__restrict__ pair_a, const uint32_t *__restrict__ pair_b, const double *__restrict__ Y,
	const double *__restrict__ W, const double *__restrict__ U, const double *__restrict__ gc,
	const double *__restrict__ gp, const uint32_t *__restrict__ obs_pt, double *__restrict__ S, double *__restrict__ b)
{
	const int lane = threadIdx.x & 31;
	const size_t blk = blockIdx.x * (size_t)SB_WARPS + (threadIdx.x >> 5);
	if(blk >= n_blocks_total) return;
	const unsigned bi = blk_row[blk], bj = blk_col[blk];
	const bool diag = bi == bj;
	double acc[36], accb[6];
	#pragma unroll
	for(int i = 0; i < 36; ++ i) acc[i] = 0;
	#pragma unroll
	for(int i = 0; i < 6; ++ i) accb[i] = 0;
	const uint64_t beg = blk_ptr[blk], end = blk_ptr[blk + 1];
	for(uint64_t k = beg + lane; k < end; k += 32) {
		const unsigned oa = pair_a[k], ob = pair_b[k];
		double y[18], w[18];
		const double *Ya = Y + (size_t)oa * 18, *Wb = W + (size_t)ob * 18;
		#pragma unroll
		for(int i = 0; i < 18; i += 2) {
			double2 t = *reinterpret_cast<const double2*>(Ya + i);
			y[i] = t.x; y[i + 1] = t.y;
			double2 s = *reinterpret_cast<const double2*>(Wb + i);
			w[i] = s.x; w[i + 1] = s.y;
		}
		// (Y_a W_b^T)(r, c) = sum_k Y(r,k) W(c,k), column-major accumulators
		#pragma unroll
		for(int c = 0; c < 6; ++ c) {
			#pragma unroll
			for(int r = 0; r < 6; ++ r)
				acc[c * 6 + r] += y[r] * w[c] + y[6 + r] * w[6 + c] + y[12 + r] * w[12 + c];
		}
		if(diag) {
			const unsigned p = obs_pt[oa];
			const double g0 = gp[(size_t)p * 3], g1 = gp[(size_t)p * 3 + 1], g2 = gp[(size_t)p * 3 + 2];
			#pragma unroll
			for(int r = 0; r < 6; ++ r)
				accb[r] += y[r] * g0 + y[6 + r] * g1 + y[12 + r] * g2;
		}
	}
	#pragma unroll
	for(int i = 0; i < 36; ++ i) {
		double v = acc[i];
		#pragma unroll
		for(int o = 16; o > 0; o >>= 1)
			v += __shfl_xor_sync(0xffffffffu, v, o);
		acc[i] = v;
	}
	if(diag) {
		#pragma unroll
		for(int i = 0; i < 6; ++ i) {
			double v = accb[i];
			#pragma unroll
			for(int o = 16; o > 0; o >>= 1)
				v += __shfl_xor_sync(0xffffffffu, v, o);
			accb[i] = v;
		}
	}
	// every lane holds the full sums; lanes 0..35 -> lane l writes entries l and (l + 32 < 36)
	#pragma unroll
	for(int i = 0; i < 36; ++ i) {
		if((i & 31) == lane) {
			const int c = i / 6, r = i % 6;
			double v = -acc[i];
			if(diag) {
				v += U[(size_t)bi * 36 + i];
				if(r == c) v += alpha;
			}
			S[((size_t)bj * 6 + c) * ld + (size_t)bi * 6 + r] = v;
		}
	}
	if(diag && lane < 6) {
		double v = 0;
		#pragma unroll
		for(int i = 0; i < 6; ++ i)
			if(i == lane) v = accb[i];
		b[(size_t)bi * 6 + lane] = gc[(size_t)bi * 6 + lane] - v;
	}
}

// thread per landmark: dl = Cinv (gp - sum_o W_o^T dxc_{c(o)})
__global__ void k_backsubstitute(size_t P, const uint32_t *__restrict__ pt_ptr, const uint32_t *__restrict__ obs_cam,
	const double *__restrict__ W, const double *__restrict__ Cinv, const double *__restrict__ gp,
	const double *__restrict__ dxc, double *__restrict__ dxp)
{
	size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
	if(p >= P) return;
	double l0 = gp[p * 3], l1 = gp[p * 3 + 1], l2 = gp[p * 3 + 2];
	const unsigned beg = pt_ptr[p], end = pt_ptr[p + 1];
	for(unsigned o = beg; o < end; ++ o) {
		const unsigned c = obs_cam[o];
		const double *Wo = W + (size_t)o * 18;
		const double *d = dxc + (size_t)c * 6;
		double dd[6];
		#pragma unroll
		for(int i = 0; i < 6; i += 2) {
			double2 t = *reinterpret_cast<const double2*>(d + i);
			dd[i] = t.x; dd[i + 1] = t.y;
		}
		double w[18];
		#pragma unroll
		for(int i = 0; i < 18; i += 2) {
			double2 t = *reinterpret_cast<const double2*>(Wo + i);
			w[i] = t.x; w[i + 1] = t.y;
		}
		double s0 = 0, s1 = 0, s2 = 0;
		#pragma unroll
		for(int r = 0; r < 6; ++ r) {
			s0 += w[r] * dd[r];
			s1 += w[6 + r] * dd[r];
			s2 += w[12 + r] * dd[r];
		}
		l0 -= s0; l1 -= s1; l2 -= s2;
	}
	const double *Ci = Cinv + p * 9;
	dxp[p * 3 + 0] = Ci[0] * l0 + Ci[3] * l1 + Ci[6] * l2;
	dxp[p * 3 + 1] = Ci[1] * l0 + Ci[4] * l1 + Ci[7] * l2;
	dxp[p * 3 + 2] = Ci[2] * l0 + Ci[5] * l1 + Ci[8] * l2;
}

// ---------------------------------------------------------------------------------------------------

// forms Cinv, Y, the dense upper-triangular reduced camera system S and its right-hand side b
// alpha damps the landmark blocks of this rank; alpha_diag is what this rank adds to the camera diagonal (alpha on
// rank 0, zero elsewhere: the partial systems are summed over ranks)
void schur_form_reduced_system(spp_ctx *ctx, double alpha, double alpha_diag)
{
	SchurSystem &s = ctx->sys;
	const size_t n = s.C * 6;
	const size_t ld = dense_chol_ld(n); // S is written straight into the dense solver's padded storage
	s.Cinv.resize(s.P * 9);
	s.Y.resize(s.O * 18);
	s.S.resize(dense_chol_storage(n));
	s.b.resize(n);
	if(s.P) {
		k_landmark_inverse<<<n_blocks(s.P, 128), 128, 0, ctx->stream>>>(s.P, alpha, s.pt_ptr.p(), s.V.p(), s.W.p(),
			s.Cinv.p(), s.Y.p());
		LAUNCH_CHECK(ctx);
	}
	s.S.zero(ctx->stream);
	if(s.n_blocks) {
		k_schur_blocks<<<n_blocks(s.n_blocks, SB_WARPS), SB_WARPS * 32, 0, ctx->stream>>>(s.n_blocks, ld, alpha_diag,
			s.blk_row.p(), s.blk_col.p(), s.blk_ptr.p(), s.pair_a.p(), s.pair_b.p(), s.Y.p(), s.W.p(), s.U.p(),
			s.gc.p(), s.gp.p(), s.obs_pt.p(), s.S.p(), s.b.p());
		LAUNCH_CHECK(ctx);
	}
}

void schur_backsubstitute(spp_ctx *ctx)
{
	SchurSystem &s = ctx->sys;
	s.dxp.resize(s.P * 3);
	if(s.P) {
		k_backsubstitute<<<n_blocks(s.P, 128), 128, 0, ctx->stream>>>(s.P, s.pt_ptr.p(), s.obs_cam.p(), s.W.p(),
			s.Cinv.p(), s.gp.p(), s.dxc.p(), s.dxp.p());
		LAUNCH_CHECK(ctx);
	}
}

} // namespace spp
