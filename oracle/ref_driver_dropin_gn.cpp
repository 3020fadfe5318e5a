/*
 * ref_driver_dropin_gn.cpp -- TEST INFRASTRUCTURE ONLY; the drop-in test of slot 3 for pose graphs.
 *
 * The UNMODIFIED reference (SLAM++ headers + sources compiled from /root/reference by oracle/build_ref.sh): its own
 * CFlatSystem, CVertexPose2D/3D and CEdgePose2D/3D, with the nonlinear solver TYPE either the reference's
 * CNonlinearSolver_Lambda (include/slam/NonlinearSolver_Lambda.h) or CNonlinearSolver_Lambda_B200
 * (include/slam_b200/NonlinearSolver_Lambda_B200.h: linearisation, block Cholesky, update and chi2 on the GPU).
 *
 *   batch:        all vertices and edges first, then Optimize(max_iter, min_dx)
 *   incremental:  the way slam_app feeds a pose graph (include/slam_app/Main.h:1108-1114, CParseLoop::AppendSystem):
 *                 edges arrive sorted by their later pose, the edge constructor creates and initialises the new pose,
 *                 solver.Incremental_Step(edge) after every edge with a nonlinear solve every <period> new vertices
 *                 (TIncrementalSolveSetting(solve::nonlinear, frequency::Every(period), max_iter, min_dx)), and a final
 *                 Optimize(max_iter, min_dx).
 * SPP_DROPIN_MARGS=1 adds the marginals policy (mpart_Diagonal) and dumps the block diagonal of the covariance.
 *
 * usage: ref_driver_dropin_gn <b200|ref> <batch|incremental> <graph.bin> <out.dump> [max_iter=5] [min_dx=0] [period=10]
 */

#include <string.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>

#include "slam/LinearSolver_UberBlock.h"
#include "slam/ConfigSolvers.h"
#include "slam/SE2_Types.h"
#include "slam/SE3_Types.h"
#include "slam/NonlinearSolver_Lambda.h"
#include "slam/Timer.h"

#include "slam_b200/NonlinearSolver_Lambda_B200.h"
#include "spp_dump.h"

int n_dummy_param = 0; // the reference's solvers expect this global to exist

template <class CSolver, class CSystemType, class CVertex, class CEdge, int n_dim>
static int Run_Solver(const spp_graph_t &g, bool b_incremental, FILE *p_fw, size_t n_max_iter, double f_min_dx, size_t n_period)
{
	typedef typename CSystemType::_TyHessianMatrixBlockList TBlockSizes;
	typedef CLinearSolver_UberBlock<TBlockSizes> CLinearSolverType;
	typedef Eigen::Matrix<double, n_dim, 1> TVec;
	typedef Eigen::Matrix<double, n_dim, n_dim> TMat;

	const bool b_marginals = getenv("SPP_DROPIN_MARGS") != 0;
	TMarginalsComputationPolicy t_margs = (b_marginals)? TMarginalsComputationPolicy(true, (b_incremental)?
		frequency::Every(n_period) : frequency::Never(), mpart_Diagonal, mpart_Diagonal) : TMarginalsComputationPolicy();
	TIncrementalSolveSetting t_incremental = (b_incremental)? TIncrementalSolveSetting(solve::nonlinear,
		frequency::Every(n_period), n_max_iter, f_min_dx) : TIncrementalSolveSetting();
	CSystemType system;
	CSolver solver(system, t_incremental, t_margs, getenv("SPP_REF_VERBOSE") != 0, CLinearSolverType(), false);

	std::vector<uint64_t> edge_order(g.n_edges);
	for(uint64_t e = 0; e < g.n_edges; ++ e) edge_order[e] = e;
	if(b_incremental) { // by the later pose; the odometry edge that creates a pose comes before the loop closures that use it
		std::vector<std::pair<std::pair<uint64_t, uint64_t>, uint64_t> > keys(g.n_edges);
		for(uint64_t e = 0; e < g.n_edges; ++ e) {
			keys[e].first.first = std::max(g.e0[e], g.e1[e]);
			keys[e].first.second = (g.e1[e] == g.e0[e] + 1)? 0 : 1;
			keys[e].second = e;
		}
		std::stable_sort(keys.begin(), keys.end());
		for(uint64_t e = 0; e < g.n_edges; ++ e) edge_order[e] = keys[e].second;
	} else {
		for(uint64_t i = 0; i < g.n_vertices; ++ i) {
			TVec v;
			for(int j = 0; j < n_dim; ++ j) v(j) = g.vdata[g.voff[i] + j];
			system.template r_Get_Vertex<CVertex>(i, v);
		}
	}
	CTimer timer;
	double f_opt_time = 0;
	std::vector<double> chi2_trace;
	for(uint64_t k = 0; k < g.n_edges; ++ k) {
		const uint64_t e = edge_order[k];
		TVec z;
		TMat info;
		for(int j = 0; j < n_dim; ++ j) {
			z(j) = g.z[n_dim * e + j];
			for(int l = 0; l < n_dim; ++ l) info(j, l) = g.info[n_dim * n_dim * e + n_dim * j + l];
		}
		CEdge &r_edge = system.r_Add_Edge(CEdge(g.e0[e], g.e1[e], z, info, system));
		if(b_incremental) {
			double f_start = timer.f_Time();
			solver.Incremental_Step(r_edge);
			f_opt_time += timer.f_Time() - f_start;
			if(getenv("SPP_TRACE_STEPS") && k < (uint64_t)atol(getenv("SPP_TRACE_STEPS"))) // debugging aid: the state after every edge
				printf("step %zu: edge %zu -> %zu, %zu vertices, chi2 %.12g\n", size_t(k), size_t(g.e0[e]), size_t(g.e1[e]),
					system.r_Vertex_Pool().n_Size(), solver.f_Chi_Squared_Error_Denorm());
		}
	}
	if(!b_incremental)
		chi2_trace.push_back(solver.f_Chi_Squared_Error_Denorm());
	double f_start = timer.f_Time();
	solver.Optimize(n_max_iter, f_min_dx);
	f_opt_time += timer.f_Time() - f_start;
	chi2_trace.push_back(solver.f_Chi_Squared_Error_Denorm());
	if(getenv("SPP_REF_DUMP_TIMING"))
		solver.Dump(f_opt_time);

	std::vector<double> states;
	for(size_t i = 0, n = system.r_Vertex_Pool().n_Size(); i < n; ++ i) {
		const typename CSystemType::_TyBaseVertex &r_vertex = system.r_Vertex_Pool()[i];
		for(int j = 0; j < r_vertex.r_v_State().rows(); ++ j) states.push_back(r_vertex.r_v_State()(j));
	}
	uint64_t n_vertices = system.r_Vertex_Pool().n_Size(), n_edges = system.r_Edge_Pool().n_Size();
	spp_dump_f64(p_fw, "chi2_trace", chi2_trace.size(), &chi2_trace[0]);
	spp_dump_f64(p_fw, "states", states.size(), &states[0]);
	spp_dump_f64(p_fw, "optimize_time", 1, &f_opt_time);
	spp_dump_u64(p_fw, "n_vertices", 1, &n_vertices);
	spp_dump_u64(p_fw, "n_edges", 1, &n_edges);
	if(b_marginals) {
		const CUberBlockMatrix &r_m = solver.r_MarginalCovariance().r_SparseMatrix();
		std::vector<double> cov;
		for(size_t i = 0, n = r_m.n_BlockColumn_Num(); i < n; ++ i) {
			CUberBlockMatrix::_TyConstMatrixXdRef t_b = r_m.t_GetBlock_Log(i, i);
			for(int r = 0; r < t_b.rows(); ++ r)
				for(int c = 0; c < t_b.cols(); ++ c) cov.push_back(t_b(r, c));
		}
		spp_dump_f64(p_fw, "cov", cov.size(), &cov[0]);
	}
	printf("ref_driver_dropin_gn: %zu vertices, %zu edges, %.6f s in the solver, final chi2 %.17g\n", size_t(n_vertices),
		size_t(n_edges), f_opt_time, chi2_trace.back());
	return 0;
}

template <class CVertex, class CEdge, int n_dim>
static int Run(const spp_graph_t &g, bool b_b200, bool b_incremental, FILE *p_fw, size_t n_max_iter, double f_min_dx, size_t n_period)
{
	typedef typename MakeTypelist(CVertex) TVertexTypelist;
	typedef typename MakeTypelist(CEdge) TEdgeTypelist;
	typedef CFlatSystem<CVertex, TVertexTypelist, CEdge, TEdgeTypelist> CSystemType;
	typedef CLinearSolver_UberBlock<typename CSystemType::_TyHessianMatrixBlockList> CLinearSolverType;
	if(b_b200) {
		return Run_Solver<CNonlinearSolver_Lambda_B200<CSystemType, CLinearSolverType>, CSystemType, CVertex, CEdge, n_dim>(g,
			b_incremental, p_fw, n_max_iter, f_min_dx, n_period);
	}
	return Run_Solver<CNonlinearSolver_Lambda<CSystemType, CLinearSolverType>, CSystemType, CVertex, CEdge, n_dim>(g,
		b_incremental, p_fw, n_max_iter, f_min_dx, n_period);
}

int main(int n_arg_num, const char **p_arg_list)
{
	if(n_arg_num < 5) {
		fprintf(stderr, "usage: %s <b200|ref> <batch|incremental> <graph.bin> <out.dump> [max_iter=5] [min_dx=0] [period=10]\n", p_arg_list[0]);
		return -1;
	}
	const bool b_b200 = !strcmp(p_arg_list[1], "b200"), b_incremental = !strcmp(p_arg_list[2], "incremental");
	const size_t n_max_iter = (n_arg_num > 5)? atol(p_arg_list[5]) : 5;
	const double f_min_dx = (n_arg_num > 6)? atof(p_arg_list[6]) : 0.0;
	const size_t n_period = (n_arg_num > 7)? atol(p_arg_list[7]) : 10;
	spp_graph_t g;
	if(spp_graph_read(p_arg_list[3], &g) || (g.kind != SPP_GRAPH_SE2 && g.kind != SPP_GRAPH_SE3)) {
		fprintf(stderr, "error: failed to read pose graph \'%s\'\n", p_arg_list[3]);
		return -1;
	}
	FILE *p_fw = fopen(p_arg_list[4], "wb");
	if(!p_fw)
		return -1;
	int n_result;
	try {
		if(g.kind == SPP_GRAPH_SE2)
			n_result = Run<CVertexPose2D, CEdgePose2D, 3>(g, b_b200, b_incremental, p_fw, n_max_iter, f_min_dx, n_period);
		else
			n_result = Run<CVertexPose3D, CEdgePose3D, 6>(g, b_b200, b_incremental, p_fw, n_max_iter, f_min_dx, n_period);
	} catch(std::exception &r_exc) {
		fprintf(stderr, "error: %s\n", r_exc.what());
		n_result = -1;
	}
	fclose(p_fw);
	spp_graph_free(&g);
	return n_result;
}
