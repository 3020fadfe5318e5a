/*
 * ref_driver_dropin_pose.cpp -- TEST INFRASTRUCTURE ONLY; the drop-in test for pose graphs (SURVEY 8(b) slot 1, row a15).
 *
 * The UNMODIFIED reference (SLAM++ headers + sources compiled from /root/reference by oracle/build_ref.sh): its own
 * CFlatSystem, CVertexPose2D/3D, CEdgePose2D/3D and Gauss-Newton solver CNonlinearSolver_Lambda
 * (include/slam/NonlinearSolver_Lambda.h:476-667), with ONE template argument changed: the linear solver is
 * CLinearSolver_UberBlock_B200 (include/slam_b200/LinearSolver_UberBlock_B200.h) instead of CLinearSolver_UberBlock,
 * i.e. libspp_b200.so factors and solves on the GPU under the reference's own AMD ordering.
 *
 * usage: ref_driver_dropin_pose <graph.bin> <out.dump> [max_iter=5] [min_dx=0]
 */

#include <string.h>
#include <stdio.h>
#include <vector>

#include "slam/LinearSolver_UberBlock.h"
#include "slam/ConfigSolvers.h"
#include "slam/SE2_Types.h"
#include "slam/SE3_Types.h"
#include "slam/NonlinearSolver_Lambda.h"
#include "slam/Timer.h"

#include "slam_b200/LinearSolver_UberBlock_B200.h"
#include "spp_dump.h"

int n_dummy_param = 0; // the reference's solvers expect this global to exist

template <class CVertex, class CEdge, int n_dim>
static int Run(const spp_graph_t &g, FILE *p_dump, size_t n_max_iter, double f_min_dx)
{
	typedef typename MakeTypelist(CVertex) TVertexTypelist;
	typedef typename MakeTypelist(CEdge) TEdgeTypelist;
	typedef CFlatSystem<CVertex, TVertexTypelist, CEdge, TEdgeTypelist> CSystemType;
	typedef CNonlinearSolver_Lambda<CSystemType, CLinearSolver_UberBlock_B200> CSolver; // <- the one changed argument
	typedef Eigen::Matrix<double, n_dim, 1> TVec;
	typedef Eigen::Matrix<double, n_dim, n_dim> TMat;

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
	CSolver solver(system, TIncrementalSolveSetting(), TMarginalsComputationPolicy(), false,
		CLinearSolver_UberBlock_B200(0), false);
	double f_chi2_0 = solver.f_Chi_Squared_Error_Denorm();
	CTimer timer;
	double f_start = timer.f_Time();
	solver.Optimize(n_max_iter, f_min_dx);
	double f_time = timer.f_Time() - f_start;
	double f_chi2 = solver.f_Chi_Squared_Error_Denorm();
	std::vector<double> states;
	for(size_t i = 0, n = system.r_Vertex_Pool().n_Size(); i < n; ++ i) {
		const typename CSystemType::_TyBaseVertex &r_vertex = system.r_Vertex_Pool()[i];
		for(int j = 0; j < r_vertex.r_v_State().rows(); ++ j)
			states.push_back(r_vertex.r_v_State()(j));
	}
	spp_dump_f64(p_dump, "chi2_0", 1, &f_chi2_0);
	spp_dump_f64(p_dump, "chi2", 1, &f_chi2);
	spp_dump_f64(p_dump, "optimize_time", 1, &f_time);
	spp_dump_f64(p_dump, "states", states.size(), &states[0]);
	printf("ref_driver_dropin_pose: optimize %.6f s, chi2 %.17g -> %.17g\n", f_time, f_chi2_0, f_chi2);
	return 0;
}

int main(int n_arg_num, const char **p_arg_list)
{
	if(n_arg_num < 3) {
		fprintf(stderr, "usage: %s <graph.bin> <out.dump> [max_iter=5] [min_dx=0]\n", p_arg_list[0]);
		return -1;
	}
	const size_t n_max_iter = (n_arg_num > 3)? atol(p_arg_list[3]) : 5;
	const double f_min_dx = (n_arg_num > 4)? atof(p_arg_list[4]) : 0.0;
	spp_graph_t g;
	if(spp_graph_read(p_arg_list[1], &g) || (g.kind != SPP_GRAPH_SE2 && g.kind != SPP_GRAPH_SE3)) {
		fprintf(stderr, "error: failed to read pose graph \'%s\'\n", p_arg_list[1]);
		return -1;
	}
	FILE *p_dump = fopen(p_arg_list[2], "wb");
	if(!p_dump) {
		fprintf(stderr, "error: failed to open \'%s\'\n", p_arg_list[2]);
		return -1;
	}
	int n_result;
	try {
		if(g.kind == SPP_GRAPH_SE2)
			n_result = Run<CVertexPose2D, CEdgePose2D, 3>(g, p_dump, n_max_iter, f_min_dx);
		else
			n_result = Run<CVertexPose3D, CEdgePose3D, 6>(g, p_dump, n_max_iter, f_min_dx);
	} catch(std::exception &r_exc) {
		fprintf(stderr, "error: %s\n", r_exc.what());
		n_result = -1;
	}
	fclose(p_dump);
	spp_graph_free(&g);
	return n_result;
}
