/*
 * ref_driver_order.cpp -- TEST INFRASTRUCTURE ONLY (oracle); never linked into the product.
 *
 * Runs the UNMODIFIED reference's fill-reducing block ordering on a block pattern given by the caller:
 *     CMatrixOrdering::p_BlockOrdering(A, need_inverse = true)        (src/slam/OrderingMagic.cpp:701-1033)
 * i.e. SuiteSparse amd_l2 on the block graph of A + A^T, exactly as CLinearSolver_UberBlock::
 * SymbolicDecomposition_Blocky does (include/slam/LinearSolver_UberBlock.h:272-296) when the reference factors a
 * reduced camera system or a pose graph. The permutation is what the product's sparse Cholesky receives through the
 * reference-side adapter (spp_chol_symbolic / spp_ba_set_rcs_ordering), so that the elimination order is the
 * reference's bit for bit; the tests also compare the fill of the product's own ordering against it.
 *
 * input : u64 n, u64 nnzb, u64 col_ptr[n + 1], u64 row_idx[nnzb]   (upper block CSC, rows ascending, diagonal last)
 * output: u64 order[n]    (new position -> original block column)
 *
 * usage: ref_driver_order <pattern.bin> <order.bin>
 */

#include <stdio.h>
#include <stdint.h>
#include <vector>

#include "slam/BlockMatrix.h"
#include "slam/OrderingMagic.h"
#include "slam/Timer.h"

int n_dummy_param = 0; // the reference's solvers expect this global to exist

int main(int n_arg_num, const char **p_arg_list)
{
	if(n_arg_num < 3) {
		fprintf(stderr, "usage: ref_driver_order <pattern.bin> <order.bin>\n");
		return 2;
	}
	FILE *p_fr = fopen(p_arg_list[1], "rb");
	if(!p_fr) {
		fprintf(stderr, "ref_driver_order: cannot open %s\n", p_arg_list[1]);
		return 1;
	}
	uint64_t n = 0, nnzb = 0;
	bool b_ok = fread(&n, 8, 1, p_fr) == 1 && fread(&nnzb, 8, 1, p_fr) == 1;
	std::vector<uint64_t> col_ptr(n + 1), row_idx(nnzb);
	b_ok = b_ok && fread(&col_ptr[0], 8, n + 1, p_fr) == n + 1 && (!nnzb || fread(&row_idx[0], 8, nnzb, p_fr) == nnzb);
	fclose(p_fr);
	if(!b_ok) {
		fprintf(stderr, "ref_driver_order: truncated input\n");
		return 1;
	}
	try {
		CUberBlockMatrix A;
		Eigen::Matrix<double, 1, 1> t_one;
		t_one(0, 0) = 1;
		for(size_t i = 0; i < n; ++ i)
			A.Append_Block(t_one, i, i); // the layout first
		for(size_t c = 0; c < n; ++ c) {
			for(uint64_t k = col_ptr[c]; k < col_ptr[c + 1]; ++ k) {
				if(row_idx[k] != c)
					A.Append_Block(t_one, size_t(row_idx[k]), c);
			}
		}
		CTimer t;
		double f_start = t.f_Time();
		CMatrixOrdering mord;
		const size_t *p_order = mord.p_BlockOrdering(A, true);
		double f_time = t.f_Time() - f_start;
		std::vector<uint64_t> order(p_order, p_order + n);
		FILE *p_fw = fopen(p_arg_list[2], "wb");
		if(!p_fw || fwrite(&order[0], 8, n, p_fw) != n) {
			fprintf(stderr, "ref_driver_order: cannot write %s\n", p_arg_list[2]);
			return 1;
		}
		fclose(p_fw);
		printf("ref_driver_order: n %zu, nnzb %zu, ordering %.3f s\n", size_t(n), size_t(nnzb), f_time);
	} catch(std::exception &r_exc) {
		fprintf(stderr, "ref_driver_order: %s\n", r_exc.what());
		return 1;
	}
	return 0;
}
