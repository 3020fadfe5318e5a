/*
 * LinearSolver_UberBlock_B200.h -- reference-side adapter: the block-sparse FP64 Cholesky of libspp_b200.so in
 * SLAM++'s linear-solver slot, for systems with one block size (pose graphs: 3 x 3 or 6 x 6 blocks; a reduced
 * camera system: 6 x 6).
 *
 * Compiled INSIDE a SLAM++ build; mirrors CLinearSolver_UberBlock<CBlockMatrixTypelist>
 * (include/slam/LinearSolver_UberBlock.h:44-427) as required by the blockwise linear solver concept
 * (include/slam/LinearSolverTags.h): same member names, argument meaning and error behaviour (false <=> not positive
 * definite; std::runtime_error / std::bad_alloc otherwise).
 *
 * Elimination ordering: exactly the reference's. SymbolicDecomposition_Blocky() calls the reference's own
 * CMatrixOrdering::p_BlockOrdering (src/slam/OrderingMagic.cpp:701-1033, AMD on the block graph), as
 * CLinearSolver_UberBlock does (LinearSolver_UberBlock.h:272-296), and hands that permutation to spp_chol_symbolic, so
 * the factor has the reference's block pattern bit for bit. The numeric phase -- permutation, block Cholesky
 * (CholeskyOf_FBS, BlockMatrixFBS.inl:2341-2513) and the two triangular solves (:2136-2275) -- runs on the GPU.
 *
 *     typedef CNonlinearSolver_Lambda<CSystemType, CLinearSolver_UberBlock_B200> CNonlinearSolverType;
 *     CNonlinearSolverType solver(system, TIncrementalSolveSetting(), TMarginalsComputationPolicy(),
 *         b_verbose, CLinearSolver_UberBlock_B200(), false);
 */
#pragma once
#ifndef __LINEAR_SOLVER_UBERBLOCK_B200_INCLUDED
#define __LINEAR_SOLVER_UBERBLOCK_B200_INCLUDED

#include <stdexcept>
#include <new>
#include <vector>
#include <string>
#include <string.h>
#include <stdint.h>

#include "slam/LinearSolverTags.h" // reference
#include "slam/BlockMatrix.h"      // reference: CUberBlockMatrix
#include "slam/OrderingMagic.h"    // reference: CMatrixOrdering (AMD)
#include "spp_b200.h"

class CLinearSolver_UberBlock_B200 {
public:
	typedef CBlockwiseLinearSolverTag _Tag; /**< @brief solver type tag */

protected:
	spp_ctx_t m_p_context; /**< @brief device context (created on first use) */
	int m_n_device; /**< @brief CUDA device index */
	bool m_b_have_symbolic; /**< @brief symbolic decomposition flag */
	size_t m_n_block_size; /**< @brief the one block size of the system */
	std::vector<uint64_t> m_col_ptr, m_row_idx, m_order; /**< @brief block structure of the last lambda, its ordering */
	std::vector<double> m_values; /**< @brief staging buffer for the blocks of lambda */
	CMatrixOrdering m_ordering; /**< @brief the reference's ordering calculator */

public:
	inline CLinearSolver_UberBlock_B200(int n_device = 0)
		:m_p_context(0), m_n_device(n_device), m_b_have_symbolic(false), m_n_block_size(0)
	{}

	/** copies carry no state (cf. LinearSolver_UberBlock.h:74-76,127-130) */
	inline CLinearSolver_UberBlock_B200(const CLinearSolver_UberBlock_B200 &r_other)
		:m_p_context(0), m_n_device(r_other.m_n_device), m_b_have_symbolic(false), m_n_block_size(0)
	{}

	inline ~CLinearSolver_UberBlock_B200()
	{
		Free_Memory();
	}

	inline CLinearSolver_UberBlock_B200 &operator =(const CLinearSolver_UberBlock_B200 &r_other)
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
		std::vector<uint64_t>().swap(m_col_ptr);
		std::vector<uint64_t>().swap(m_row_idx);
		std::vector<uint64_t>().swap(m_order);
		std::vector<double>().swap(m_values);
	}

	/** one-shot solve (CLinearSolver_UberBlock::Solve_PosDef, LinearSolver_UberBlock.h:144-262) */
	bool Solve_PosDef(const CUberBlockMatrix &r_lambda, Eigen::VectorXd &r_v_eta) // throw(std::bad_alloc, std::runtime_error)
	{
		SymbolicDecomposition_Blocky(r_lambda);
		return Solve_PosDef_Blocky(r_lambda, r_v_eta);
	}

	inline void Clear_SymbolicDecomposition()
	{
		m_b_have_symbolic = false;
	}

	/** CLinearSolver_UberBlock::SymbolicDecomposition_Blocky (LinearSolver_UberBlock.h:272-296) */
	bool SymbolicDecomposition_Blocky(const CUberBlockMatrix &r_lambda) // throw(std::bad_alloc, std::runtime_error)
	{
		_ASSERTE(r_lambda.b_SymmetricLayout());
		const size_t n = r_lambda.n_BlockColumn_Num();
		if(!n)
			return true;
		m_n_block_size = r_lambda.n_BlockColumn_Column_Num(0);
		m_col_ptr.resize(n + 1);
		m_row_idx.clear();
		m_row_idx.reserve(r_lambda.n_Block_Num());
		m_col_ptr[0] = 0;
		for(size_t i = 0; i < n; ++ i) {
			if(r_lambda.n_BlockColumn_Column_Num(i) != m_n_block_size)
				throw std::runtime_error("CLinearSolver_UberBlock_B200: all block columns must have the same width");
			for(size_t j = 0, m = r_lambda.n_BlockColumn_Block_Num(i); j < m; ++ j)
				m_row_idx.push_back(r_lambda.n_Block_Row(i, j));
			m_col_ptr[i + 1] = m_row_idx.size();
		}
		m_values.resize(m_row_idx.size() * m_n_block_size * m_n_block_size);
		const size_t *p_order = m_ordering.p_BlockOrdering(r_lambda, true); // the reference's AMD, same call as UberBlock.h:281
		m_order.assign(p_order, p_order + n);
		Check(spp_chol_symbolic(p_Context(), n, m_n_block_size, &m_col_ptr[0], &m_row_idx[0], &m_order[0], 0));
		m_b_have_symbolic = true;
		return true;
	}

	/** CLinearSolver_UberBlock::Solve_PosDef_Blocky (LinearSolver_UberBlock.h:312-426); r_v_eta: rhs in, solution out */
	bool Solve_PosDef_Blocky(const CUberBlockMatrix &r_lambda, Eigen::VectorXd &r_v_eta) // throw(std::bad_alloc, std::runtime_error)
	{
		_ASSERTE(size_t(r_v_eta.rows()) == r_lambda.n_Column_Num());
		if(!m_b_have_symbolic || r_lambda.n_BlockColumn_Num() + 1 != m_col_ptr.size() ||
		   r_lambda.n_Block_Num() != m_row_idx.size())
			SymbolicDecomposition_Blocky(r_lambda); // the structure changed
		double *p_dest = m_values.empty()? 0 : &m_values[0];
		for(size_t i = 0, n = r_lambda.n_BlockColumn_Num(); i < n; ++ i) {
			for(size_t j = 0, m = r_lambda.n_BlockColumn_Block_Num(i); j < m; ++ j) {
				CUberBlockMatrix::_TyConstMatrixXdRef t_block = r_lambda.t_Block_AtColumn(i, j);
				const size_t n_size = t_block.rows() * t_block.cols();
				memcpy(p_dest, t_block.data(), n_size * sizeof(double)); // blocks are dense column-major
				p_dest += n_size;
			}
		}
		int n_result = spp_chol_solve(p_Context(), m_values.empty()? 0 : &m_values[0], &r_v_eta(0));
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
};

#endif // !__LINEAR_SOLVER_UBERBLOCK_B200_INCLUDED
