/*
 * ref_driver_dropin.cpp -- TEST INFRASTRUCTURE ONLY: the drop-in test.
 *
 * The UNMODIFIED reference (graph model, edges, Jacobians, reduction plan, Levenberg-Marquardt loop of
 * CNonlinearSolver_Lambda_LM) with ONE change: the linear solver template argument is
 * CLinearSolver_Schur_B200 (include/slam_b200/LinearSolver_Schur_B200.h), i.e. libspp_b200.so does the Schur
 * complement, the dense Cholesky and the back-substitution on the GPU through the C ABI. Everything else is the
 * reference's code compiled from /root/reference. The LM trace, final chi2 and final states are dumped so that
 * tests/test_dropin_gpu.py can compare them with the pure-reference golden vectors.
 *
 * usage: ref_driver_dropin <graph.bin> <out.dump> [max_iter=5] [min_dx=0]
 */
#include <string.h>
#include <stdio.h>
#include <omp.h>
#include <vector>

#include "slam/LinearSolver_UberBlock.h"
#include "slam/ConfigSolvers.h"
#include "slam/BA_Types.h"
#include "slam/NonlinearSolver_Lambda_LM.h"
#include "slam/Timer.h"

#include "slam_b200/LinearSolver_Schur_B200.h"
#include "spp_dump.h"

int n_dummy_param = 0;

typedef MakeTypelist_Safe((CVertexCam, CVertexXYZ)) TVertexTypelist;
typedef MakeTypelist_Safe((CEdgeP2C3D)) TEdgeTypelist;
typedef CFlatSystem<CBaseVertex, TVertexTypelist, CEdgeP2C3D, TEdgeTypelist> CSystemType;

static std::vector<double> g_trace;

template <class CLambdaLM_Solver>
class CTracingLM : public CLevenbergMarquardt_Baseline<CLambdaLM_Solver> {
public:
	bool Aftermath(double &r_f_last_error, double f_error, double &r_f_alpha, const CUberBlockMatrix &r_lambda,
		const CLambdaLM_Solver &r_solver, const Eigen::VectorXd &r_v_dx, const Eigen::VectorXd &r_v_rhs)
	{
		double f_alpha_before = r_f_alpha, f_last = r_f_last_error;
		double f_den = (r_v_dx.transpose()).dot(r_f_alpha * r_v_dx + r_v_rhs);
		bool b_good = CLevenbergMarquardt_Baseline<CLambdaLM_Solver>::Aftermath(r_f_last_error,
			f_error, r_f_alpha, r_lambda, r_solver, r_v_dx, r_v_rhs);
		const double p_row[6] = {f_alpha_before, f_last, f_error, f_den, b_good? 1.0 : 0.0, r_f_alpha};
		g_trace.insert(g_trace.end(), p_row, p_row + 6);
		return b_good;
	}
};

int main(int n_arg_num, const char **p_arg_list)
{
	if(n_arg_num < 3) {
		fprintf(stderr, "usage: %s <graph.bin> <out.dump> [max_iter=5] [min_dx=0]\n", p_arg_list[0]);
		return -1;
	}
	const size_t n_max_iter = (n_arg_num > 3)? atol(p_arg_list[3]) : 5;
	const double f_min_dx = (n_arg_num > 4)? atof(p_arg_list[4]) : 0.0;
	spp_graph_t g;
	if(spp_graph_read(p_arg_list[1], &g) || g.kind != SPP_GRAPH_BA) {
		fprintf(stderr, "error: failed to read BA graph \'%s\'\n", p_arg_list[1]);
		return -1;
	}
	FILE *p_fw = fopen(p_arg_list[2], "wb");
	if(!p_fw)
		return -1;

	CSystemType system;
	for(uint64_t i = 0; i < g.n_vertices; ++ i) {
		const double *p = g.vdata + g.voff[i];
		if(g.vtype[i] == 0) {
			Eigen::Matrix<double, 11, 1> v_cam;
			for(int j = 0; j < 11; ++ j)
				v_cam(j) = p[j];
			system.r_Get_Vertex<CVertexCam>(i, v_cam);
		} else
			system.r_Get_Vertex<CVertexXYZ>(i, Eigen::Vector3d(p[0], p[1], p[2]));
	}
	for(uint64_t e = 0; e < g.n_edges; ++ e) {
		Eigen::Matrix2d t_info;
		t_info << g.info[4 * e], g.info[4 * e + 1], g.info[4 * e + 2], g.info[4 * e + 3];
		system.r_Add_Edge(CEdgeP2C3D(g.e0[e], g.e1[e], Eigen::Vector2d(g.z[2 * e], g.z[2 * e + 1]), t_info, system));
	}

	typedef CNonlinearSolver_Lambda_LM<CSystemType, CLinearSolver_Schur_B200, CSystemType::_TyJacobianMatrixBlockList,
		CSystemType::_TyHessianMatrixBlockList, CTracingLM> CSolver;
	CSolver solver(system, TIncrementalSolveSetting(), TMarginalsComputationPolicy(),
		getenv("SPP_REF_VERBOSE") != 0, CLinearSolver_Schur_B200(), false); // b_use_schur = false: our solver gets lambda
	double f_chi2_0 = solver.f_Chi_Squared_Error_Denorm();
	CTimer timer;
	double f_start = timer.f_Time();
	try {
		solver.Optimize(n_max_iter, f_min_dx);
	} catch(std::exception &r_exc) {
		fprintf(stderr, "error: %s\n", r_exc.what());
		return -2;
	}
	double f_time = timer.f_Time() - f_start;
	double f_chi2 = solver.f_Chi_Squared_Error_Denorm();

	std::vector<double> states;
	for(size_t i = 0, n = system.r_Vertex_Pool().n_Size(); i < n; ++ i) {
		Eigen::Map<const Eigen::VectorXd> v = ((CSystemType::_TyConstVertexRef)system.r_Vertex_Pool()[i]).v_StateC();
		for(int j = 0; j < v.rows(); ++ j)
			states.push_back(v(j));
	}
	spp_dump_f64(p_fw, "chi2_0", 1, &f_chi2_0);
	spp_dump_f64(p_fw, "chi2", 1, &f_chi2);
	spp_dump_f64(p_fw, "lm_trace", g_trace.size(), g_trace.empty()? 0 : &g_trace[0]);
	spp_dump_f64(p_fw, "states", states.size(), &states[0]);
	spp_dump_f64(p_fw, "optimize_time", 1, &f_time);
	fclose(p_fw);
	printf("ref_driver_dropin: optimize %.6f s, chi2 %.17g -> %.17g\n", f_time, f_chi2_0, f_chi2);
	spp_graph_free(&g);
	return 0;
}
