/*
 * LinearSolver_Schur_B200.h -- reference-side adapter: plugs libspp_b200.so into SLAM++'s linear-solver slot.
 *
 * This header is compiled INSIDE a SLAM++ build (it includes the reference's own headers by their usual paths);
 * it is the C++ host layer above the C ABI (include/spp_b200.h) and mirrors the interface of
 *     CLinearSolver_Schur   (include/slam/LinearSolver_Schur.h:1414-1935)  and
 *     CLinearSolver_UberBlock (include/slam/LinearSolver_UberBlock.h:44-427)
 * as required by the blockwise linear solver concept (include/slam/LinearSolverTags.h): same member names,
 * argument meaning and error behaviour --
 *     typedef CBlockwiseLinearSolverTag _Tag;
 *     copy-constructible / assignable without carrying state (solvers are passed by value, LM.h:486)
 *     void Free_Memory();
 *     bool Solve_PosDef(const CUberBlockMatrix &lambda, Eigen::VectorXd &eta);
 *     void Clear_SymbolicDecomposition();
 *     bool SymbolicDecomposition_Blocky(const CUberBlockMatrix &lambda);
 *     bool Solve_PosDef_Blocky(const CUberBlockMatrix &lambda, Eigen::VectorXd &eta);   // false <=> not pos. def.
 * Errors: CUDA / communication failures throw std::runtime_error, allocation failures std::bad_alloc (the
 * reference's solvers throw the same types).
 *
 * Use (drop-in): give this type as the CLinearSolver template argument of CNonlinearSolver_Lambda_LM /
 * CNonlinearSolver_Lambda and construct the nonlinear solver with b_use_schur = false -- the solver then hands the
 * full lambda to Solve_PosDef_Blocky() (LM.h:1523-1535) and the Schur complement, the dense Cholesky and the
 * back-substitution all run on the GPU:
 *
 *     typedef CNonlinearSolver_Lambda_LM<CSystemType, CLinearSolver_Schur_B200> CNonlinearSolverType;
 *     CNonlinearSolverType solver(system, TIncrementalSolveSetting(), TMarginalsComputationPolicy(),
 *         b_verbose, CLinearSolver_Schur_B200(), false);
 */
#pragma once
#ifndef __LINEAR_SOLVER_SCHUR_B200_INCLUDED
#define __LINEAR_SOLVER_SCHUR_B200_INCLUDED

#include <stdexcept>
#include <new>
#include <vector>
#include <string>
#include <string.h>
#include <stdint.h>

#include "slam/LinearSolverTags.h" // reference
#include "slam/BlockMatrix.h"      // reference: CUberBlockMatrix
#include "slam/OrderingMagic.h"    // reference: CMatrixOrdering (AMD), for a block-sparse reduced camera system
#include "spp_b200.h"

class CLinearSolver_Schur_B200 {
public:
	typedef CBlockwiseLinearSolverTag _Tag; /**< @brief solver type tag */

protected:
	spp_ctx_t m_p_context; /**< @brief device context (created on first use) */
	int m_n_device; /**< @brief CUDA device index */
	bool m_b_have_symbolic; /**< @brief symbolic decomposition flag */
	std::vector<uint64_t> m_col_dims, m_col_ptr, m_row_idx; /**< @brief block structure of the last lambda */
	std::vector<double> m_values; /**< @brief staging buffer for the blocks of lambda */

public:
	inline CLinearSolver_Schur_B200(int n_device = 0)
		:m_p_context(0), m_n_device(n_device), m_b_have_symbolic(false)
	{}

	/** copies carry no state (cf. CLinearSolver_UberBlock, LinearSolver_UberBlock.h:74-76) */
	inline CLinearSolver_Schur_B200(const CLinearSolver_Schur_B200 &r_other)
		:m_p_context(0), m_n_device(r_other.m_n_device), m_b_have_symbolic(false)
	{}

	inline ~CLinearSolver_Schur_B200()
	{
		Free_Memory();
	}

	inline CLinearSolver_Schur_B200 &operator =(const CLinearSolver_Schur_B200 &r_other)
	{
		m_n_device = r_other.m_n_device;
		return *this;
	}

	void Free_Memory()
	{
		if(m_p_context) {
			spp_destroy(m_p_context);
			m_p_context = 0;
		}
		m_b_have_symbolic = false;
		std::vector<uint64_t>().swap(m_col_dims);
		std::vector<uint64_t>().swap(m_col_ptr);
		std::vector<uint64_t>().swap(m_row_idx);
		std::vector<double>().swap(m_values);
	}

	/** one-shot solve: symbolic + numeric (CLinearSolver_Schur::Solve_PosDef, LinearSolver_Schur.h:1512-1518) */
	bool Solve_PosDef(const CUberBlockMatrix &r_lambda, Eigen::VectorXd &r_v_eta) // throw(std::bad_alloc, std::runtime_error)
	{
		SymbolicDecomposition_Blocky(r_lambda);
		return Solve_PosDef_Blocky(r_lambda, r_v_eta);
	}

	inline void Clear_SymbolicDecomposition()
	{
		m_b_have_symbolic = false;
	}

	/** CLinearSolver_Schur::SymbolicDecomposition_Blocky (LinearSolver_Schur.h:1566-1606) */
	bool SymbolicDecomposition_Blocky(const CUberBlockMatrix &r_lambda) // throw(std::bad_alloc, std::runtime_error)
	{
		Flatten_Structure(r_lambda);
		uint64_t n_cut = 0;
		Check(spp_schur_symbolic(p_Context(), m_col_dims.size(), &m_col_dims[0], &m_col_ptr[0],
			m_row_idx.empty()? 0 : &m_row_idx[0], 0, &n_cut));
		if(n_cut * 6 > 16384)
			Set_Reference_RCS_Ordering(size_t(n_cut));
		m_b_have_symbolic = true;
		return true;
	}

	/**
	 *	A reduced camera system this large is factored block-sparse (SPP_RCS_AUTO), the path the reference takes when
	 *	its dense solver throws std::bad_alloc (LinearSolver_Schur.h:1836-1847): CLinearSolver_UberBlock on the Schur
	 *	complement, under the AMD ordering of its block structure. This gives the library that same permutation:
	 *	the block pattern of S comes back from the library, the ordering from the reference's own CMatrixOrdering.
	 */
	void Set_Reference_RCS_Ordering(size_t n_camera_num) // throw(std::bad_alloc, std::runtime_error)
	{
		std::vector<uint8_t> pattern(n_camera_num * n_camera_num);
		Check(spp_schur_get_reduced_system(p_Context(), 0, 0, 0, &pattern[0]));
		CUberBlockMatrix S_structure; // 1 x 1 blocks: only the block graph matters to p_BlockOrdering
		Eigen::Matrix<double, 1, 1> t_one;
		t_one(0, 0) = 1;
		for(size_t i = 0; i < n_camera_num; ++ i)
			S_structure.Append_Block(t_one, i, i);
		for(size_t c = 0; c < n_camera_num; ++ c) {
			for(size_t r = 0; r < c; ++ r) {
				if(pattern[r * n_camera_num + c])
					S_structure.Append_Block(t_one, r, c);
			}
		}
		CMatrixOrdering ordering;
		const size_t *p_order = ordering.p_BlockOrdering(S_structure, true);
		std::vector<uint64_t> order(p_order, p_order + n_camera_num);
		Check(spp_schur_set_rcs_ordering(p_Context(), n_camera_num, &order[0]));
	}

	/** CLinearSolver_Schur::Solve_PosDef_Blocky (LinearSolver_Schur.h:1623-1935); r_v_eta: rhs in, solution out */
	bool Solve_PosDef_Blocky(const CUberBlockMatrix &r_lambda, Eigen::VectorXd &r_v_eta) // throw(std::bad_alloc, std::runtime_error)
	{
		_ASSERTE(r_lambda.b_SymmetricLayout());
		_ASSERTE(size_t(r_v_eta.rows()) == r_lambda.n_Column_Num());
		if(!m_b_have_symbolic || r_lambda.n_BlockColumn_Num() != m_col_dims.size() ||
		   r_lambda.n_Block_Num() != m_row_idx.size())
			SymbolicDecomposition_Blocky(r_lambda); // the structure changed (the reference re-orders in that case too)
		Flatten_Values(r_lambda);
		int n_result = spp_schur_solve(p_Context(), m_values.empty()? 0 : &m_values[0], &r_v_eta(0));
		if(n_result == SPP_NOT_POSDEF)
			return false;
		Check(n_result);
		return true;
	}

protected:
	spp_ctx_t p_Context() // throw(std::bad_alloc, std::runtime_error)
	{
		if(!m_p_context)
			Check(spp_create(m_n_device, &m_p_context));
		return m_p_context;
	}

	void Check(int n_result) const // throw(std::bad_alloc, std::runtime_error)
	{
		if(n_result == SPP_OK)
			return;
		if(n_result == SPP_ERR_NOMEM)
			throw std::bad_alloc();
		throw std::runtime_error(std::string("libspp_b200: ") + spp_last_error(m_p_context));
	}

	/** block structure through the public accessors (include/slam/BlockMatrix.h:343-430) */
	void Flatten_Structure(const CUberBlockMatrix &r_lambda) // throw(std::bad_alloc)
	{
		const size_t n = r_lambda.n_BlockColumn_Num();
		m_col_dims.resize(n);
		m_col_ptr.resize(n + 1);
		m_row_idx.clear();
		m_row_idx.reserve(r_lambda.n_Block_Num());
		size_t n_value_num = 0;
		m_col_ptr[0] = 0;
		for(size_t i = 0; i < n; ++ i) {
			m_col_dims[i] = r_lambda.n_BlockColumn_Column_Num(i);
			for(size_t j = 0, m = r_lambda.n_BlockColumn_Block_Num(i); j < m; ++ j) {
				const size_t n_row = r_lambda.n_Block_Row(i, j);
				m_row_idx.push_back(n_row);
				n_value_num += r_lambda.n_BlockColumn_Column_Num(n_row) * m_col_dims[i]; // symmetric layout
			}
			m_col_ptr[i + 1] = m_row_idx.size();
		}
		m_values.resize(n_value_num);
	}

	/** block values (dense column-major blocks) through t_Block_AtColumn (include/slam/BlockMatrix.h:470-485) */
	void Flatten_Values(const CUberBlockMatrix &r_lambda)
	{
		double *p_dest = m_values.empty()? 0 : &m_values[0];
		for(size_t i = 0, n = r_lambda.n_BlockColumn_Num(); i < n; ++ i) {
			for(size_t j = 0, m = r_lambda.n_BlockColumn_Block_Num(i); j < m; ++ j) {
				CUberBlockMatrix::_TyConstMatrixXdRef t_block = r_lambda.t_Block_AtColumn(i, j);
				const size_t n_size = t_block.rows() * t_block.cols();
				memcpy(p_dest, t_block.data(), n_size * sizeof(double)); // blocks are dense column-major
				p_dest += n_size;
			}
		}
		_ASSERTE(p_dest == (m_values.empty()? 0 : &m_values[0]) + m_values.size());
	}
};

#endif // !__LINEAR_SOLVER_SCHUR_B200_INCLUDED
