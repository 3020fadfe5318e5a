// schur_slot.cu -- slot 1 of the drop-in boundary: the landmark Schur linear solver on a lambda handed over
// by the caller in the reference's block layout (upper block-triangular, vertex id order, column-major blocks).
//
// Replaces CLinearSolver_Schur::SymbolicDecomposition_Blocky (include/slam/LinearSolver_Schur.h:1566-1606, guided
// ordering src/slam/LinearSolver_Schur.cpp:771-838) and Solve_PosDef_Blocky (Schur.h:1623-1935). The permutation,
// slicing and transposition passes of the reference (Schur.h:1687-1709) become one gather kernel that reads the
// caller's value array once and writes (U, V, W, eta_c, eta_p) in Schur order.

#include "spp_ctx.h"
#include <algorithm>

namespace spp {

void build_schur_structure(spp_ctx *ctx, size_t C, size_t P, const std::vector<uint32_t> &h_cam,
	const std::vector<uint32_t> &h_pt, std::vector<uint32_t> &obs_orig, std::vector<uint32_t> &t_cam,
	std::vector<uint32_t> &t_pt);
int schur_solve_current(spp_ctx *ctx, double alpha, spp_report_t *rep);

#define LAUNCH_CHECK(ctx) do { ++ (ctx)->n_launches; SPP_CUDA(cudaGetLastError()); } while(0)

// one thread per destination scalar
__global__ void k_slot_gather(size_t C, size_t P, size_t O, const double *__restrict__ vals, const double *__restrict__ eta,
	const uint64_t *__restrict__ u_src, const uint64_t *__restrict__ v_src, const uint64_t *__restrict__ w_src,
	const uint8_t *__restrict__ w_tr, const uint64_t *__restrict__ cam_eta, const uint64_t *__restrict__ pt_eta,
	double *__restrict__ U, double *__restrict__ V, double *__restrict__ W, double *__restrict__ gc, double *__restrict__ gp)
{
	size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
	const size_t nU = C * 36, nV = P * 9, nW = O * 18, nc = C * 6, np = P * 3;
	if(i < nU) {
		U[i] = vals[u_src[i / 36] + i % 36];
		return;
	}
	i -= nU;
	if(i < nV) {
		V[i] = vals[v_src[i / 9] + i % 9];
		return;
	}
	i -= nV;
	if(i < nW) {
		size_t o = i / 18, k = i % 18;
		size_t r = k % 6, c = k / 6; // destination: 6x3 column-major
		W[i] = vals[w_src[o] + (w_tr[o]? r * 3 + c : k)]; // source 3x6 column-major holds W^T
		return;
	}
	i -= nW;
	if(i < nc) {
		gc[i] = eta[cam_eta[i / 6] + i % 6];
		return;
	}
	i -= nc;
	if(i < np)
		gp[i] = eta[pt_eta[i / 3] + i % 3];
}

__global__ void k_slot_scatter(size_t C, size_t P, const double *__restrict__ dxc, const double *__restrict__ dxp,
	const uint64_t *__restrict__ cam_eta, const uint64_t *__restrict__ pt_eta, double *__restrict__ eta)
{
	size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
	const size_t nc = C * 6, np = P * 3;
	if(i < nc)
		eta[cam_eta[i / 6] + i % 6] = dxc[i];
	else if(i - nc < np) {
		i -= nc;
		eta[pt_eta[i / 3] + i % 3] = dxp[i];
	}
}

void slot_symbolic(spp_ctx *ctx, size_t n, const uint64_t *col_dims, const uint64_t *col_ptr, const uint64_t *row_idx,
	uint64_t *p_order, uint64_t *p_cut)
{
	SchurSlot &sl = ctx->slot;
	sl.valid = false;
	sl.filled = false;
	ctx->ba.valid = false; // the Schur system buffers are shared
	std::vector<uint32_t> local(n);
	std::vector<uint64_t> cam_cols, pt_cols;
	sl.col_base.assign(n + 1, 0);
	for(size_t i = 0; i < n; ++ i) {
		if(col_dims[i] == 6) {
			local[i] = (uint32_t)cam_cols.size();
			cam_cols.push_back(i);
		} else if(col_dims[i] == 3) {
			local[i] = (uint32_t)pt_cols.size();
			pt_cols.push_back(i);
		} else
			throw invalid_error("Schur slot: block columns must be 6 (pose) or 3 (landmark) wide");
		sl.col_base[i + 1] = sl.col_base[i] + col_dims[i];
	}
	const size_t C = cam_cols.size(), P = pt_cols.size();
	// the reference abandons the guided ordering when the pose part is not the smaller half (Schur.h:1583)
	if(!C || !P || C >= n / 2)
		throw invalid_error("Schur slot: guided ordering not applicable (needs 0 < #poses < #vertices / 2)");
	sl.order.resize(n);
	for(size_t i = 0; i < C; ++ i) sl.order[i] = cam_cols[i];
	for(size_t i = 0; i < P; ++ i) sl.order[C + i] = pt_cols[i];
	sl.cut = C;
	if(p_order) std::copy(sl.order.begin(), sl.order.end(), p_order);
	if(p_cut) *p_cut = C;

	std::vector<uint64_t> u_src(C), v_src(P), w_src_e;
	std::vector<uint8_t> w_tr_e;
	std::vector<uint32_t> h_cam, h_pt;
	std::vector<char> have_diag(n, 0);
	uint64_t off = 0;
	for(size_t c = 0; c < n; ++ c) {
		for(uint64_t k = col_ptr[c]; k < col_ptr[c + 1]; ++ k) {
			const uint64_t r = row_idx[k];
			if(r > c || r >= n)
				throw invalid_error("Schur slot: lambda must be upper block-triangular");
			if(r == c) {
				if(col_dims[c] == 6) u_src[local[c]] = off; else v_src[local[c]] = off;
				have_diag[c] = 1;
			} else if(col_dims[r] != col_dims[c]) {
				const bool row_is_cam = col_dims[r] == 6;
				h_cam.push_back(local[row_is_cam? r : c]);
				h_pt.push_back(local[row_is_cam? c : r]);
				w_src_e.push_back(off);
				w_tr_e.push_back(row_is_cam? 0 : 1);
			} else if(col_dims[r] == 6)
				throw invalid_error("Schur slot: pose-pose off-diagonal blocks are not supported yet");
			else
				throw invalid_error("Schur slot: landmark-landmark blocks (non block-diagonal C) are not supported");
			off += col_dims[r] * col_dims[c];
		}
	}
	for(size_t c = 0; c < n; ++ c)
		if(!have_diag[c]) throw invalid_error("Schur slot: missing diagonal block");
	sl.n_bcols = n;
	sl.n_scalars = sl.col_base[n];
	sl.n_values = off;

	std::vector<uint32_t> obs_orig, t_cam, t_pt;
	build_schur_structure(ctx, C, P, h_cam, h_pt, obs_orig, t_cam, t_pt);
	const size_t O = h_cam.size();
	std::vector<uint64_t> w_src(O);
	std::vector<uint8_t> w_tr(O);
	for(size_t k = 0; k < O; ++ k) {
		w_src[k] = w_src_e[obs_orig[k]];
		w_tr[k] = w_tr_e[obs_orig[k]];
	}
	std::vector<uint64_t> cam_eta(C), pt_eta(P);
	for(size_t i = 0; i < C; ++ i) cam_eta[i] = sl.col_base[cam_cols[i]];
	for(size_t i = 0; i < P; ++ i) pt_eta[i] = sl.col_base[pt_cols[i]];
	cudaStream_t st = ctx->stream;
	sl.u_src.upload(u_src, st); sl.v_src.upload(v_src, st); sl.w_src.upload(w_src, st);
	sl.w_transposed.upload(w_tr, st);
	sl.cam_eta_off.upload(cam_eta, st); sl.pt_eta_off.upload(pt_eta, st);
	sl.vals.resize(sl.n_values);
	sl.eta.resize(sl.n_scalars);
	SPP_CUDA(cudaStreamSynchronize(st));
	sl.valid = true;
}

int slot_solve(spp_ctx *ctx, const double *p_values, double *p_eta_dx)
{
	SchurSlot &sl = ctx->slot;
	SchurSystem &s = ctx->sys;
	if(!sl.valid)
		throw invalid_error("spp_schur_symbolic() has not been called");
	cudaStream_t st = ctx->stream;
	SPP_CUDA(cudaMemcpyAsync(sl.vals.p(), p_values, sl.n_values * 8, cudaMemcpyHostToDevice, st));
	SPP_CUDA(cudaMemcpyAsync(sl.eta.p(), p_eta_dx, sl.n_scalars * 8, cudaMemcpyHostToDevice, st));
	const size_t total = s.C * 36 + s.P * 9 + s.O * 18 + s.C * 6 + s.P * 3;
	k_slot_gather<<<n_blocks(total, 256), 256, 0, st>>>(s.C, s.P, s.O, sl.vals.p(), sl.eta.p(), sl.u_src.p(),
		sl.v_src.p(), sl.w_src.p(), sl.w_transposed.p(), sl.cam_eta_off.p(), sl.pt_eta_off.p(), s.U.p(), s.V.p(),
		s.W.p(), s.gc.p(), s.gp.p());
	LAUNCH_CHECK(ctx);
	sl.filled = true;
	int rc = schur_solve_current(ctx, 0.0, 0);
	if(rc != SPP_OK)
		return rc;
	k_slot_scatter<<<n_blocks(s.C * 6 + s.P * 3, 256), 256, 0, st>>>(s.C, s.P, s.dxc.p(), s.dxp.p(),
		sl.cam_eta_off.p(), sl.pt_eta_off.p(), sl.eta.p());
	LAUNCH_CHECK(ctx);
	SPP_CUDA(cudaMemcpyAsync(p_eta_dx, sl.eta.p(), sl.n_scalars * 8, cudaMemcpyDeviceToHost, st));
	SPP_CUDA(cudaStreamSynchronize(st));
	return SPP_OK;
}

} // namespace spp
