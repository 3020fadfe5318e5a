/*
 * ref_driver_pose.cpp -- TEST INFRASTRUCTURE ONLY (oracle); never linked into the product.
 *
 * A thin driver on top of the UNMODIFIED reference (SLAM++ headers + sources compiled from /root/reference by
 * oracle/build_ref.sh) for the pose-graph configurations: loads a binary SE(2) or
 * SE(3) pose graph (SPPGRAF1, slam_plus_plus_b200/sppio.py), builds the reference's own
 *     CFlatSystem<CVertexPose2D, (CVertexPose2D), CEdgePose2D, (CEdgePose2D)>      (src/slam_simple_example/Main.cpp)
 *     CFlatSystem<CVertexPose3D, (CVertexPose3D), CEdgePose3D, (CEdgePose3D)>
 * and runs the reference's Gauss-Newton solver CNonlinearSolver_Lambda::Optimize()
 * (include/slam/NonlinearSolver_Lambda.h:476-667) with CLinearSolver_UberBlock (the native block Cholesky,
 * include/slam/LinearSolver_UberBlock.h:312-426).
 *
 * Modes:
 *   time  : the plain configuration; prints / records the wall time of Optimize() -> CPU baseline.
 *   dump  : a pass-through linear solver in the reference's linear-solver slot (LinearSolverTags.h concept)
 *           forwards every call to CLinearSolver_UberBlock and records lambda (block structure + values), eta and
 *           dx of every solve; for the first solve also the AMD block ordering (CMatrixOrdering::p_BlockOrdering,
 *           src/slam/OrderingMagic.cpp:701-1033) and the block pattern of the Cholesky factor R
 *           (CUberBlockMatrix::CholeskyOf_FBS on the permuted matrix, as LinearSolver_UberBlock.h:328-402 does).
 *
 * usage: ref_driver_pose <time|dump> <graph.bin> <out.dump> [max_iter=5] [min_dx=0]
 */

#include <string.h>
#include <stdio.h>
#include <omp.h>
#include <vector>
#include <string>

#include "slam/LinearSolver_UberBlock.h"
#include "slam/ConfigSolvers.h"
#include "slam/SE2_Types.h"
#include "slam/SE3_Types.h"
#include "slam/NonlinearSolver_Lambda.h"
#include "slam/OrderingMagic.h"
#include "slam/Timer.h"

#include "spp_dump.h"

int n_dummy_param = 0; // the reference's solvers expect this global to exist

static FILE *g_dump = 0; // records are appended here
static size_t g_n_solve = 0;

static void Dump_Structure(const char *p_s_tag, const CUberBlockMatrix &r_m, bool b_values)
{
	char p_s_name[32];
	const size_t n = r_m.n_BlockColumn_Num();
	std::vector<uint64_t> col_dims(n), col_ptr(n + 1), row_idx;
	std::vector<double> vals;
	col_ptr[0] = 0;
	for(size_t i = 0; i < n; ++ i) {
		col_dims[i] = r_m.n_BlockColumn_Column_Num(i);
		const size_t nb = r_m.n_BlockColumn_Block_Num(i);
		for(size_t j = 0; j < nb; ++ j) {
			row_idx.push_back(r_m.n_Block_Row(i, j));
			if(b_values) {
				CUberBlockMatrix::_TyConstMatrixXdRef t_block = r_m.t_Block_AtColumn(i, j);
				for(int c = 0; c < t_block.cols(); ++ c) {
					for(int r = 0; r < t_block.rows(); ++ r)
						vals.push_back(t_block(r, c)); // column-major, as stored
				}
			}
		}
		col_ptr[i + 1] = row_idx.size();
	}
	snprintf(p_s_name, sizeof(p_s_name), "%s.col_dims", p_s_tag);
	spp_dump_u64(g_dump, p_s_name, col_dims.size(), col_dims.empty()? 0 : &col_dims[0]);
	snprintf(p_s_name, sizeof(p_s_name), "%s.col_ptr", p_s_tag);
	spp_dump_u64(g_dump, p_s_name, col_ptr.size(), &col_ptr[0]);
	snprintf(p_s_name, sizeof(p_s_name), "%s.row_idx", p_s_tag);
	spp_dump_u64(g_dump, p_s_name, row_idx.size(), row_idx.empty()? 0 : &row_idx[0]);
	if(b_values) {
		snprintf(p_s_name, sizeof(p_s_name), "%s.vals", p_s_tag);
		spp_dump_f64(g_dump, p_s_name, vals.size(), vals.empty()? 0 : &vals[0]);
	}
}

/**
 *	@brief pass-through linear solver; satisfies the reference's blockwise linear solver concept
 */
template <class CBlockSizes>
class CRecordingSolver {
public:
	typedef CBlockwiseLinearSolverTag _Tag;
	typedef CLinearSolver_UberBlock<CBlockSizes> CRefLinearSolver;

protected:
	CRefLinearSolver m_solver;

public:
	CRecordingSolver()
	{}

	CRecordingSolver(const CRecordingSolver &UNUSED(r_other))
	{}

	CRecordingSolver &operator =(const CRecordingSolver &UNUSED(r_other))
	{
		return *this;
	}

	void Free_Memory()
	{
		m_solver.Free_Memory();
	}

	void Clear_SymbolicDecomposition()
	{
		m_solver.Clear_SymbolicDecomposition();
	}

	bool SymbolicDecomposition_Blocky(const CUberBlockMatrix &r_lambda)
	{
		return m_solver.SymbolicDecomposition_Blocky(r_lambda);
	}

	bool Solve_PosDef(const CUberBlockMatrix &r_lambda, Eigen::VectorXd &r_v_eta)
	{
		m_solver.SymbolicDecomposition_Blocky(r_lambda);
		return Solve_PosDef_Blocky(r_lambda, r_v_eta);
	}

	bool Solve_PosDef_Blocky(const CUberBlockMatrix &r_lambda, Eigen::VectorXd &r_v_eta)
	{
		char p_s_name[32], p_s_tag[16];
		snprintf(p_s_tag, sizeof(p_s_tag), "L%u", unsigned(g_n_solve));
		if(g_dump) {
			Dump_Structure(p_s_tag, r_lambda, true);
			snprintf(p_s_name, sizeof(p_s_name), "L%u.eta", unsigned(g_n_solve));
			spp_dump_f64(g_dump, p_s_name, r_v_eta.rows(), &r_v_eta(0));
			if(!g_n_solve) {
				// the ordering and the factor pattern, by the same calls as LinearSolver_UberBlock.h:272-296, 328-402
				const size_t n = r_lambda.n_BlockColumn_Num();
				CMatrixOrdering mord;
				const size_t *p_order = mord.p_BlockOrdering(r_lambda, true);
				const size_t *p_inv_order = mord.p_Get_InverseOrdering();
				std::vector<uint64_t> order(p_order, p_order + n), inv_order(p_inv_order, p_inv_order + n);
				spp_dump_u64(g_dump, "amd.order", n, &order[0]);
				spp_dump_u64(g_dump, "amd.inv_order", n, &inv_order[0]);
				CUberBlockMatrix perm, R;
				r_lambda.Permute_UpperTriangular_To(perm, p_inv_order, n, true);
				std::vector<size_t> etree(n, 0), workspace(n, 0), zeroes(n, 0);
				perm.Build_EliminationTree(etree, workspace);
				std::vector<uint64_t> etree64(etree.begin(), etree.end());
				spp_dump_u64(g_dump, "amd.etree", n, &etree64[0]);
				if(R.template CholeskyOf_FBS<CBlockSizes>(perm, etree, workspace, zeroes))
					Dump_Structure("R", R, true);
			}
		}
		bool b_result = m_solver.Solve_PosDef_Blocky(r_lambda, r_v_eta);
		if(g_dump) {
			snprintf(p_s_name, sizeof(p_s_name), "L%u.dx", unsigned(g_n_solve));
			spp_dump_f64(g_dump, p_s_name, r_v_eta.rows(), &r_v_eta(0));
			uint64_t n_ok = b_result;
			snprintf(p_s_name, sizeof(p_s_name), "L%u.ok", unsigned(g_n_solve));
			spp_dump_u64(g_dump, p_s_name, 1, &n_ok);
		}
		++ g_n_solve;
		return b_result;
	}
};

template <class CSystemType>
static void Dump_States(const CSystemType &r_system, const char *p_s_name)
{
	std::vector<double> states;
	for(size_t i = 0, n = r_system.r_Vertex_Pool().n_Size(); i < n; ++ i) {
		const typename CSystemType::_TyBaseVertex &r_vertex = r_system.r_Vertex_Pool()[i];
		for(int j = 0; j < r_vertex.r_v_State().rows(); ++ j)
			states.push_back(r_vertex.r_v_State()(j));
	}
	spp_dump_f64(g_dump, p_s_name, states.size(), &states[0]);
}

template <class CVertex, class CEdge, int n_dim>
static int Run(const spp_graph_t &g, const char *p_s_mode, size_t n_arg4, double f_arg5)
{
	typedef typename MakeTypelist(CVertex) TVertexTypelist;
	typedef typename MakeTypelist(CEdge) TEdgeTypelist;
	typedef CFlatSystem<CVertex, TVertexTypelist, CEdge, TEdgeTypelist> CSystemType;
	typedef typename CSystemType::_TyHessianMatrixBlockList TBlockSizes;
	typedef CLinearSolver_UberBlock<TBlockSizes> CRefLinearSolver;
	typedef Eigen::Matrix<double, n_dim, 1> TVec;
	typedef Eigen::Matrix<double, n_dim, n_dim> TMat;

	CTimer timer;
	CSystemType system;
	for(uint64_t i = 0; i < g.n_vertices; ++ i) {
		TVec v;
		for(int j = 0; j < n_dim; ++ j)
			v(j) = g.vdata[g.voff[i] + j];
		system.template r_Get_Vertex<CVertex>(i, v);
	}
	for(uint64_t e = 0; e < g.n_edges; ++ e) {
		TVec z;
		TMat info;
		for(int j = 0; j < n_dim; ++ j) {
			z(j) = g.z[n_dim * e + j];
			for(int k = 0; k < n_dim; ++ k)
				info(j, k) = g.info[n_dim * n_dim * e + n_dim * j + k];
		}
		system.r_Add_Edge(CEdge(g.e0[e], g.e1[e], z, info, system));
	}
	double f_build_time = timer.f_Time();

	double f_opt_time, f_chi2;
	uint64_t n_threads = omp_get_max_threads();
	if(!strcmp(p_s_mode, "steps")) {
		// bench: <warmup> <steps> calls of Optimize(1, 0) on the resident system
		const size_t n_warmup = n_arg4, n_steps = size_t(f_arg5);
		typedef CNonlinearSolver_Lambda<CSystemType, CRefLinearSolver> CSolver;
		CSolver solver(system, TIncrementalSolveSetting(), TMarginalsComputationPolicy(), false, CRefLinearSolver(), false);
		std::vector<double> step_seconds;
		double f_begin = timer.f_Time();
		for(size_t i = 0; i < n_warmup + n_steps; ++ i) {
			double f_start = timer.f_Time();
			solver.Optimize(1, 0);
			step_seconds.push_back(timer.f_Time() - f_start);
		}
		f_opt_time = timer.f_Time() - f_begin;
		f_chi2 = solver.f_Chi_Squared_Error_Denorm();
		spp_dump_f64(g_dump, "step_seconds", step_seconds.size(), &step_seconds[0]);
	} else if(!strcmp(p_s_mode, "margs")) {
		// Optimize(max_iter), then the block diagonal of the covariance (NonlinearSolver_Lambda.h:669-767 ->
		// CMarginals::Calculate_DenseMarginals_Recurrent_FBS, marginals policy mpart_Diagonal)
		typedef CNonlinearSolver_Lambda<CSystemType, CRefLinearSolver> CSolver;
		CSolver solver(system, TIncrementalSolveSetting(), TMarginalsComputationPolicy(true, frequency::Never(),
			mpart_Diagonal, mpart_Diagonal), false, CRefLinearSolver(), false);
		double f_start = timer.f_Time();
		solver.Optimize(n_arg4, f_arg5);
		f_opt_time = timer.f_Time() - f_start;
		f_chi2 = solver.f_Chi_Squared_Error_Denorm();
		if(getenv("SPP_REF_VERBOSE"))
			solver.Dump(f_opt_time);
		Dump_States(system, "states");
		const CUberBlockMatrix &r_m = solver.r_MarginalCovariance().r_SparseMatrix();
		std::vector<double> cov;
		for(size_t i = 0, n = r_m.n_BlockColumn_Num(); i < n; ++ i) {
			CUberBlockMatrix::_TyConstMatrixXdRef t_b = r_m.t_GetBlock_Log(i, i);
			for(int r = 0; r < t_b.rows(); ++ r)
				for(int c = 0; c < t_b.cols(); ++ c)
					cov.push_back(t_b(r, c));
		}
		spp_dump_f64(g_dump, "cov", cov.size(), &cov[0]); // vertex id order, row-major blocks
	} else if(strcmp(p_s_mode, "dump")) {
		typedef CNonlinearSolver_Lambda<CSystemType, CRefLinearSolver> CSolver;
		CSolver solver(system, TIncrementalSolveSetting(), TMarginalsComputationPolicy(), false, CRefLinearSolver(), false);
		double f_start = timer.f_Time();
		solver.Optimize(n_arg4, f_arg5);
		f_opt_time = timer.f_Time() - f_start;
		f_chi2 = solver.f_Chi_Squared_Error_Denorm();
	} else {
		typedef CNonlinearSolver_Lambda<CSystemType, CRecordingSolver<TBlockSizes> > CSolver;
		CSolver solver(system, TIncrementalSolveSetting(), TMarginalsComputationPolicy(), false,
			CRecordingSolver<TBlockSizes>(), false);
		Dump_States(system, "states0");
		double f_chi2_0 = solver.f_Chi_Squared_Error_Denorm();
		spp_dump_f64(g_dump, "chi2_0", 1, &f_chi2_0);
		double f_start = timer.f_Time();
		solver.Optimize(n_arg4, f_arg5);
		f_opt_time = timer.f_Time() - f_start;
		f_chi2 = solver.f_Chi_Squared_Error_Denorm();
		uint64_t n_solves = g_n_solve;
		spp_dump_u64(g_dump, "n_solves", 1, &n_solves);
		Dump_States(system, "states");
	}
	spp_dump_f64(g_dump, "chi2", 1, &f_chi2);
	spp_dump_f64(g_dump, "optimize_time", 1, &f_opt_time);
	spp_dump_f64(g_dump, "build_time", 1, &f_build_time);
	spp_dump_u64(g_dump, "omp_threads", 1, &n_threads);
	printf("ref_driver_pose: %s: optimize %.6f s, chi2 %.17g, threads %u\n", p_s_mode, f_opt_time, f_chi2, unsigned(n_threads));
	return 0;
}

int main(int n_arg_num, const char **p_arg_list)
{
	if(n_arg_num < 4) {
		fprintf(stderr, "usage: %s <time|dump|steps|margs> <graph.bin> <out.dump> [max_iter=5 | warmup] [min_dx=0 | steps]\n", p_arg_list[0]);
		return -1;
	}
	const size_t n_arg4 = (n_arg_num > 4)? atol(p_arg_list[4]) : 5;
	const double f_arg5 = (n_arg_num > 5)? atof(p_arg_list[5]) : 0.0;
	spp_graph_t g;
	if(spp_graph_read(p_arg_list[2], &g) || (g.kind != SPP_GRAPH_SE2 && g.kind != SPP_GRAPH_SE3)) {
		fprintf(stderr, "error: failed to read pose graph \'%s\'\n", p_arg_list[2]);
		return -1;
	}
	if(!(g_dump = fopen(p_arg_list[3], "wb"))) {
		fprintf(stderr, "error: failed to open \'%s\'\n", p_arg_list[3]);
		return -1;
	}
	int n_result;
	if(g.kind == SPP_GRAPH_SE2)
		n_result = Run<CVertexPose2D, CEdgePose2D, 3>(g, p_arg_list[1], n_arg4, f_arg5);
	else
		n_result = Run<CVertexPose3D, CEdgePose3D, 6>(g, p_arg_list[1], n_arg4, f_arg5);
	fclose(g_dump);
	spp_graph_free(&g);
	return n_result;
}
