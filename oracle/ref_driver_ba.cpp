/*
 * ref_driver_ba.cpp -- TEST INFRASTRUCTURE ONLY (oracle); never linked into the product.
 *
 * A thin driver on top of the UNMODIFIED reference (SLAM++ headers + sources compiled from
 * /root/reference by oracle/build_ref.sh). It loads a binary BA graph (SPPGRAF1, see
 * slam_plus_plus_b200/sppio.py), builds the reference's own
 *     CFlatSystem<CBaseVertex, (CVertexCam, CVertexXYZ), CEdgeP2C3D, (CEdgeP2C3D)>
 * (the typedefs of src/ba_interface_example/BAOptimizer.cpp:110-117) and runs the reference's
 *     CNonlinearSolver_Lambda_LM::Optimize()   (include/slam/NonlinearSolver_Lambda_LM.h:796)
 *
 * Modes:
 *   time  : exactly the CBAOptimizerCore configuration (CLinearSolver_UberBlock, b_use_schur = true);
 *           prints the wall time of Optimize() -> this is the CPU baseline.
 *   dump  : b_use_schur = false with a pass-through linear solver plugged into the reference's
 *           linear-solver slot (LinearSolverTags.h concept). The pass-through forwards every call to the
 *           reference's own CLinearSolver_Schur (same type as the hard-wired m_schur_solver,
 *           NonlinearSolver_Base.h:344-346), and records lambda (block structure + values), eta and dx of
 *           every solve, plus the LM trace (alpha, chi2, accept/reject) through a logging trust-region
 *           policy derived from CLevenbergMarquardt_Baseline (NonlinearSolver_Lambda_LM.h:134-240).
 *
 * usage: ref_driver_ba <time|dump> <graph.bin> <out.dump> [max_iter=5] [min_dx=0]
 */

#include <string.h>
#include <stdio.h>
#include <omp.h>
#include <vector>
#include <string>

#include "slam/LinearSolver_UberBlock.h"
#include "slam/ConfigSolvers.h"
#include "slam/BA_Types.h"
#include "slam/NonlinearSolver_Lambda_LM.h"
#include "slam/LinearSolver_Schur.h"
#include "slam/Timer.h"

#include "spp_dump.h"

int n_dummy_param = 0; // the reference's solvers expect this global to exist

typedef MakeTypelist_Safe((CVertexCam, CVertexXYZ)) TVertexTypelist;
typedef MakeTypelist_Safe((CEdgeP2C3D)) TEdgeTypelist;
typedef CFlatSystem<CBaseVertex, TVertexTypelist, CEdgeP2C3D, TEdgeTypelist> CSystemType;
typedef CSystemType::_TyHessianMatrixBlockList TBlockSizes;
typedef CLinearSolver_UberBlock<TBlockSizes> CRefLinearSolver;
typedef CLinearSolver_Schur<CRefLinearSolver, TBlockSizes, CSystemType> CRefSchurSolver;

static FILE *g_dump = 0; // records are appended here
static size_t g_n_solve = 0;
static std::vector<double> g_trace; // per Aftermath: alpha_before, chi2_last, chi2_new, rho_den, accepted, alpha_after

static void Dump_Lambda(const char *p_s_tag, size_t n_solve, const CUberBlockMatrix &r_lambda)
{
	char p_s_name[32];
	const size_t n = r_lambda.n_BlockColumn_Num();
	std::vector<uint64_t> col_dims(n), col_ptr(n + 1), row_idx;
	std::vector<double> vals;
	col_ptr[0] = 0;
	for(size_t i = 0; i < n; ++ i) {
		col_dims[i] = r_lambda.n_BlockColumn_Column_Num(i);
		const size_t nb = r_lambda.n_BlockColumn_Block_Num(i);
		for(size_t j = 0; j < nb; ++ j) {
			row_idx.push_back(r_lambda.n_Block_Row(i, j));
			CUberBlockMatrix::_TyConstMatrixXdRef t_block = r_lambda.t_Block_AtColumn(i, j);
			for(int c = 0; c < t_block.cols(); ++ c) {
				for(int r = 0; r < t_block.rows(); ++ r)
					vals.push_back(t_block(r, c)); // column-major, as stored
			}
		}
		col_ptr[i + 1] = row_idx.size();
	}
	snprintf(p_s_name, sizeof(p_s_name), "%s%u.col_dims", p_s_tag, unsigned(n_solve));
	spp_dump_u64(g_dump, p_s_name, col_dims.size(), col_dims.empty()? 0 : &col_dims[0]);
	snprintf(p_s_name, sizeof(p_s_name), "%s%u.col_ptr", p_s_tag, unsigned(n_solve));
	spp_dump_u64(g_dump, p_s_name, col_ptr.size(), &col_ptr[0]);
	snprintf(p_s_name, sizeof(p_s_name), "%s%u.row_idx", p_s_tag, unsigned(n_solve));
	spp_dump_u64(g_dump, p_s_name, row_idx.size(), row_idx.empty()? 0 : &row_idx[0]);
	snprintf(p_s_name, sizeof(p_s_name), "%s%u.vals", p_s_tag, unsigned(n_solve));
	spp_dump_f64(g_dump, p_s_name, vals.size(), vals.empty()? 0 : &vals[0]);
}

/**
 *	@brief pass-through linear solver; satisfies the reference's blockwise linear solver concept
 */
class CRecordingSolver {
public:
	typedef CBlockwiseLinearSolverTag _Tag;

protected:
	CRefSchurSolver m_schur;

public:
	CRecordingSolver()
		:m_schur(CRefLinearSolver())
	{}

	CRecordingSolver(const CRecordingSolver &UNUSED(r_other))
		:m_schur(CRefLinearSolver())
	{}

	CRecordingSolver &operator =(const CRecordingSolver &UNUSED(r_other))
	{
		return *this;
	}

	void Free_Memory()
	{
		m_schur.Free_Memory();
	}

	void Clear_SymbolicDecomposition()
	{
		m_schur.Clear_SymbolicDecomposition();
	}

	bool SymbolicDecomposition_Blocky(const CUberBlockMatrix &r_lambda)
	{
		return m_schur.SymbolicDecomposition_Blocky(r_lambda);
	}

	bool Solve_PosDef(const CUberBlockMatrix &r_lambda, Eigen::VectorXd &r_v_eta)
	{
		m_schur.SymbolicDecomposition_Blocky(r_lambda);
		return Solve_PosDef_Blocky(r_lambda, r_v_eta);
	}

	bool Solve_PosDef_Blocky(const CUberBlockMatrix &r_lambda, Eigen::VectorXd &r_v_eta)
	{
		char p_s_name[32];
		if(g_dump) {
			Dump_Lambda("L", g_n_solve, r_lambda);
			snprintf(p_s_name, sizeof(p_s_name), "L%u.eta", unsigned(g_n_solve));
			spp_dump_f64(g_dump, p_s_name, r_v_eta.rows(), &r_v_eta(0));
		}
		bool b_result = m_schur.Solve_PosDef_Blocky(r_lambda, r_v_eta);
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

/**
 *	@brief the reference's baseline LM policy, with a trace of its decisions
 */
template <class CLambdaLM_Solver>
class CTracingLM : public CLevenbergMarquardt_Baseline<CLambdaLM_Solver> {
public:
	double f_InitialDamping(const typename CLambdaLM_Solver::_TySystem &r_system)
	{
		double f_alpha = CLevenbergMarquardt_Baseline<CLambdaLM_Solver>::f_InitialDamping(r_system);
		if(g_dump)
			spp_dump_f64(g_dump, "alpha0", 1, &f_alpha);
		return f_alpha;
	}

	bool Aftermath(double &r_f_last_error, double f_error, double &r_f_alpha, const CUberBlockMatrix &r_lambda,
		const CLambdaLM_Solver &r_solver, const Eigen::VectorXd &r_v_dx, const Eigen::VectorXd &r_v_rhs)
	{
		double f_alpha_before = r_f_alpha, f_last = r_f_last_error;
		double f_den = (r_v_dx.transpose()).dot(r_f_alpha * r_v_dx + r_v_rhs);
		bool b_good = CLevenbergMarquardt_Baseline<CLambdaLM_Solver>::Aftermath(r_f_last_error,
			f_error, r_f_alpha, r_lambda, r_solver, r_v_dx, r_v_rhs);
		g_trace.push_back(f_alpha_before);
		g_trace.push_back(f_last);
		g_trace.push_back(f_error);
		g_trace.push_back(f_den);
		g_trace.push_back(b_good? 1.0 : 0.0);
		g_trace.push_back(r_f_alpha);
		return b_good;
	}
};

template <class CSolver>
static void Dump_States(const CSystemType &r_system, const char *p_s_name)
{
	std::vector<double> states;
	for(size_t i = 0, n = r_system.r_Vertex_Pool().n_Size(); i < n; ++ i) {
		Eigen::Map<const Eigen::VectorXd> v = ((CSystemType::_TyConstVertexRef)r_system.r_Vertex_Pool()[i]).v_StateC();
		for(int j = 0; j < v.rows(); ++ j)
			states.push_back(v(j));
	}
	spp_dump_f64(g_dump, p_s_name, states.size(), &states[0]);
}

static void Build_System(CSystemType &r_system, const spp_graph_t &g)
{
	for(uint64_t i = 0; i < g.n_vertices; ++ i) {
		const double *p = g.vdata + g.voff[i];
		if(g.vtype[i] == 0) {
			Eigen::Matrix<double, 11, 1> v_cam;
			for(int j = 0; j < 11; ++ j)
				v_cam(j) = p[j];
			r_system.r_Get_Vertex<CVertexCam>(i, v_cam);
		} else {
			Eigen::Vector3d v_pt(p[0], p[1], p[2]);
			r_system.r_Get_Vertex<CVertexXYZ>(i, v_pt);
		}
	}
	for(uint64_t e = 0; e < g.n_edges; ++ e) {
		Eigen::Vector2d v_z(g.z[2 * e], g.z[2 * e + 1]);
		Eigen::Matrix2d t_info;
		t_info << g.info[4 * e], g.info[4 * e + 1], g.info[4 * e + 2], g.info[4 * e + 3];
		r_system.r_Add_Edge(CEdgeP2C3D(g.e0[e], g.e1[e], v_z, t_info, r_system));
	}
}

int main(int n_arg_num, const char **p_arg_list)
{
	if(n_arg_num < 4) {
		fprintf(stderr, "usage: %s <time|dump|steps|margs> <graph.bin> <out.dump> [max_iter=5] [min_dx=0]\n", p_arg_list[0]);
		return -1;
	}
	const bool b_dump = !strcmp(p_arg_list[1], "dump");
	const size_t n_max_iter = (n_arg_num > 4)? atol(p_arg_list[4]) : 5;
	const double f_min_dx = (n_arg_num > 5)? atof(p_arg_list[5]) : 0.0;

	spp_graph_t g;
	if(spp_graph_read(p_arg_list[2], &g) || g.kind != SPP_GRAPH_BA) {
		fprintf(stderr, "error: failed to read BA graph \'%s\'\n", p_arg_list[2]);
		return -1;
	}
	if(!(g_dump = fopen(p_arg_list[3], "wb"))) {
		fprintf(stderr, "error: failed to open \'%s\'\n", p_arg_list[3]);
		return -1;
	}

	CTimer timer;
	CSystemType system;
	Build_System(system, g);
	double f_build_time = timer.f_Time();

	double f_opt_time, f_chi2;
	uint64_t n_threads = omp_get_max_threads();
	if(!strcmp(p_arg_list[1], "steps")) {
		// bench.py --impl reference: <warmup> <steps> calls of Optimize(1, 0) on the resident system
		const size_t n_warmup = (n_arg_num > 4)? atol(p_arg_list[4]) : 0;
		const size_t n_steps = (n_arg_num > 5)? atol(p_arg_list[5]) : 1;
		typedef CNonlinearSolver_Lambda_LM<CSystemType, CRefLinearSolver> CSolver;
		CSolver solver(system, TIncrementalSolveSetting(), TMarginalsComputationPolicy(),
			getenv("SPP_REF_VERBOSE") != 0, CRefLinearSolver(), true);
		FILE *p_keep = g_dump;
		g_dump = 0;
		std::vector<double> step_seconds;
		double f_begin = timer.f_Time();
		for(size_t i = 0; i < n_warmup + n_steps; ++ i) {
			double f_start = timer.f_Time();
			solver.Optimize(1, 0);
			step_seconds.push_back(timer.f_Time() - f_start);
		}
		f_opt_time = timer.f_Time() - f_begin;
		g_dump = p_keep;
		f_chi2 = solver.f_Chi_Squared_Error_Denorm();
		spp_dump_f64(g_dump, "step_seconds", step_seconds.size(), &step_seconds[0]);
	} else if(!strcmp(p_arg_list[1], "margs")) {
		// Optimize(max_iter), then the block diagonal of the covariance the reference recovers from the
		// Schur-complemented system (NonlinearSolver_Lambda_LM.h:1118-1350 -> BAMarginals.h:579, mpart_Diagonal)
		typedef CNonlinearSolver_Lambda_LM<CSystemType, CRefLinearSolver> CSolver;
		CSolver solver(system, TIncrementalSolveSetting(), TMarginalsComputationPolicy(true, frequency::Never(),
			mpart_Diagonal, mpart_Diagonal), getenv("SPP_REF_VERBOSE") != 0, CRefLinearSolver(), true);
		FILE *p_keep = g_dump;
		g_dump = 0;
		double f_start = timer.f_Time();
		solver.Optimize(n_max_iter, f_min_dx);
		f_opt_time = timer.f_Time() - f_start;
		g_dump = p_keep;
		f_chi2 = solver.f_Chi_Squared_Error_Denorm();
		if(getenv("SPP_REF_VERBOSE"))
			solver.Dump(f_opt_time); // time spent in the marginals among the rest
		Dump_States<CSolver>(system, "states");
		const CUberBlockMatrix &r_m = solver.r_MarginalCovariance().r_SparseMatrix();
		std::vector<double> cov6, cov3;
		for(size_t i = 0, n = r_m.n_BlockColumn_Num(); i < n; ++ i) {
			CUberBlockMatrix::_TyConstMatrixXdRef t_b = r_m.t_GetBlock_Log(i, i);
			std::vector<double> &r_dst = (t_b.cols() == 6)? cov6 : cov3;
			for(int r = 0; r < t_b.rows(); ++ r)
				for(int c = 0; c < t_b.cols(); ++ c)
					r_dst.push_back(t_b(r, c));
		}
		double f_zero = 0;
		spp_dump_f64(g_dump, "cam_cov", cov6.size(), cov6.empty()? &f_zero : &cov6[0]); // vertex id order, row-major
		spp_dump_f64(g_dump, "pt_cov", cov3.size(), cov3.empty()? &f_zero : &cov3[0]);
	} else if(!b_dump) {
		typedef CNonlinearSolver_Lambda_LM<CSystemType, CRefLinearSolver> CSolver;
		CSolver solver(system, TIncrementalSolveSetting(), TMarginalsComputationPolicy(),
			getenv("SPP_REF_VERBOSE") != 0, CRefLinearSolver(), true);
		FILE *p_keep = g_dump;
		g_dump = 0; // no recording in the timed mode
		double f_start = timer.f_Time();
		solver.Optimize(n_max_iter, f_min_dx);
		f_opt_time = timer.f_Time() - f_start;
		g_dump = p_keep;
		f_chi2 = solver.f_Chi_Squared_Error_Denorm();
		solver.Dump(f_opt_time);
	} else {
		typedef CNonlinearSolver_Lambda_LM<CSystemType, CRecordingSolver, CSystemType::_TyJacobianMatrixBlockList,
			CSystemType::_TyHessianMatrixBlockList, CTracingLM> CSolver;
		CSolver solver(system, TIncrementalSolveSetting(), TMarginalsComputationPolicy(),
			getenv("SPP_REF_VERBOSE") != 0, CRecordingSolver(), false);
		Dump_States<CSolver>(system, "states0");
		double f_chi2_0 = solver.f_Chi_Squared_Error_Denorm();
		spp_dump_f64(g_dump, "chi2_0", 1, &f_chi2_0);
		double f_start = timer.f_Time();
		solver.Optimize(n_max_iter, f_min_dx);
		f_opt_time = timer.f_Time() - f_start;
		f_chi2 = solver.f_Chi_Squared_Error_Denorm();
		uint64_t n_solves = g_n_solve;
		spp_dump_u64(g_dump, "n_solves", 1, &n_solves);
		spp_dump_f64(g_dump, "lm_trace", g_trace.size(), g_trace.empty()? 0 : &g_trace[0]);
		Dump_States<CSolver>(system, "states");
	}
	spp_dump_f64(g_dump, "chi2", 1, &f_chi2);
	spp_dump_f64(g_dump, "optimize_time", 1, &f_opt_time);
	spp_dump_f64(g_dump, "build_time", 1, &f_build_time);
	spp_dump_u64(g_dump, "omp_threads", 1, &n_threads);
	fclose(g_dump);
	printf("ref_driver_ba: %s: optimize %.6f s, chi2 %.17g, threads %u\n",
		p_arg_list[1], f_opt_time, f_chi2, unsigned(n_threads));
	spp_graph_free(&g);
	return 0;
}
