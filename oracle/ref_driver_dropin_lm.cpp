/*
 * ref_driver_dropin_lm.cpp -- TEST INFRASTRUCTURE ONLY; the drop-in test for slot 3.
 *
 * The UNMODIFIED reference's CFlatSystem with CVertexCam / CVertexXYZ / CEdgeP2C3D, with the nonlinear solver type
 * changed from CNonlinearSolver_Lambda_LM to CNonlinearSolver_Lambda_LM_B200
 * (include/slam_b200/NonlinearSolver_Lambda_LM_B200.h): the whole LM loop, linearisation included, runs on the GPU.
 * Two modes, both also run with the reference's own solver for comparison ("ref" as the 2nd argument):
 *   batch        Optimize(max_iter, min_dx) on the whole graph
 *   incremental  the marker-driven incremental BA of the application (CParseLoop_ConsistencyMarker,
 *                include/slam_app/IncBAParsePrimitives.h:154-168): cameras arrive in id order, a landmark and its
 *                observations are added once two of its cameras are present, Optimize() is called every <batch> cameras
 *                on the SAME system and solver objects (append-only, as the application does)
 *   periodic     no markers: the solver is built with TIncrementalSolveSetting(solve::nonlinear, frequency::Every(batch),
 *                max_iter, min_dx) and decides inside Incremental_Step() when to solve (NonlinearSolver_Base.h:557-622);
 *                the dump lists the edges after which the states changed ("solve_edges"), so the two solver types can be
 *                compared solve by solve
 *
 * usage: ref_driver_dropin_lm <b200|ref> <batch|incremental|periodic> <graph.bin> <out.dump> [max_iter=5] [min_dx=0] [batch=10]
 */

#include <string.h>
#include <stdio.h>
#include <vector>
#include <map>

#include "slam/LinearSolver_UberBlock.h"
#include "slam/ConfigSolvers.h"
#include "slam/BA_Types.h"
#include "slam/NonlinearSolver_Lambda_LM.h"
#include "slam/Timer.h"

#include "slam_b200/NonlinearSolver_Lambda_LM_B200.h"
#include "spp_dump.h"

int n_dummy_param = 0;

typedef MakeTypelist_Safe((CVertexCam, CVertexXYZ)) TVertexTypelist;
typedef MakeTypelist_Safe((CEdgeP2C3D)) TEdgeTypelist;
typedef CFlatSystem<CBaseVertex, TVertexTypelist, CEdgeP2C3D, TEdgeTypelist> CSystemType;
typedef CLinearSolver_UberBlock<CSystemType::_TyHessianMatrixBlockList> CLinearSolverType;

struct TStates {
	std::vector<double> v;
	template <class CVertex>
	void operator ()(const CVertex &r_vertex)
	{
		for(int j = 0; j < r_vertex.r_v_State().rows(); ++ j) v.push_back(r_vertex.r_v_State()(j));
	}
};

template <class CSolver>
static int Run(const spp_graph_t &g, bool b_incremental, bool b_periodic, FILE *p_fw, size_t n_max_iter, double f_min_dx, size_t n_batch)
{
	CSystemType system;
	const bool b_marginals = getenv("SPP_DROPIN_MARGS") != 0; // every Optimize() ends with the block diagonal of the covariance
	std::vector<uint64_t> solve_edges; // periodic mode: edges after which Incremental_Step() changed the states
	std::vector<double> last_states;
	CSolver solver(system, (b_periodic)? TIncrementalSolveSetting(solve::nonlinear, frequency::Every(n_batch), n_max_iter, f_min_dx) :
		TIncrementalSolveSetting(), (b_marginals)? TMarginalsComputationPolicy(true, frequency::Never(),
		mpart_Diagonal, mpart_Diagonal) : TMarginalsComputationPolicy(), getenv("SPP_REF_VERBOSE") != 0, CLinearSolverType(), true);
	std::vector<double> chi2_trace;
	CTimer timer;
	double f_opt_time = 0;
	if(!b_incremental) {
		for(uint64_t i = 0; i < g.n_vertices; ++ i) {
			const double *p = g.vdata + g.voff[i];
			if(g.vtype[i] == 0) {
				Eigen::Matrix<double, 11, 1> v_cam;
				for(int j = 0; j < 11; ++ j) v_cam(j) = p[j];
				system.r_Get_Vertex<CVertexCam>(i, v_cam);
			} else
				system.r_Get_Vertex<CVertexXYZ>(i, Eigen::Vector3d(p[0], p[1], p[2]));
		}
		for(uint64_t e = 0; e < g.n_edges; ++ e) {
			Eigen::Matrix2d t_info;
			t_info << g.info[4 * e], g.info[4 * e + 1], g.info[4 * e + 2], g.info[4 * e + 3];
			system.r_Add_Edge(CEdgeP2C3D(g.e0[e], g.e1[e], Eigen::Vector2d(g.z[2 * e], g.z[2 * e + 1]), t_info, system));
		}
		chi2_trace.push_back(solver.f_Chi_Squared_Error_Denorm());
		double f_start = timer.f_Time();
		solver.Optimize(n_max_iter, f_min_dx);
		f_opt_time += timer.f_Time() - f_start;
		chi2_trace.push_back(solver.f_Chi_Squared_Error_Denorm());
	} else {
		// vertices get their ids in order of arrival (the system is append-only): cameras in graph order, a landmark
		// when its second camera arrives
		std::vector<uint64_t> cams;
		for(uint64_t i = 0; i < g.n_vertices; ++ i)
			if(g.vtype[i] == 0) cams.push_back(i);
		std::vector<std::vector<uint64_t> > cam_edges(g.n_vertices);
		for(uint64_t e = 0; e < g.n_edges; ++ e) cam_edges[g.e1[e]].push_back(e); // e0 = point, e1 = camera
		std::map<uint64_t, size_t> new_id; // graph vertex -> system vertex
		std::vector<std::vector<uint64_t> > waiting(g.n_vertices); // per landmark: observations whose camera is present
		std::vector<char> added(g.n_vertices, 0);
		size_t n_next_id = 0;
		for(size_t c = 0; c < cams.size(); ++ c) {
			const uint64_t i = cams[c];
			const double *p = g.vdata + g.voff[i];
			Eigen::Matrix<double, 11, 1> v_cam;
			for(int j = 0; j < 11; ++ j) v_cam(j) = p[j];
			new_id[i] = n_next_id;
			system.r_Get_Vertex<CVertexCam>(n_next_id ++, v_cam);
			CEdgeP2C3D *p_last_edge = 0;
			for(size_t k = 0; k < cam_edges[i].size(); ++ k) {
				const uint64_t e = cam_edges[i][k], pt = g.e0[e];
				waiting[pt].push_back(e);
				if(!added[pt] && waiting[pt].size() < 2)
					continue;
				if(!added[pt]) {
					const double *q = g.vdata + g.voff[pt];
					new_id[pt] = n_next_id;
					system.r_Get_Vertex<CVertexXYZ>(n_next_id ++, Eigen::Vector3d(q[0], q[1], q[2]));
					added[pt] = 1;
				}
				for(size_t w = 0; w < waiting[pt].size(); ++ w) {
					const uint64_t ee = waiting[pt][w];
					Eigen::Matrix2d t_info;
					t_info << g.info[4 * ee], g.info[4 * ee + 1], g.info[4 * ee + 2], g.info[4 * ee + 3];
					CEdgeP2C3D &r_edge = system.r_Add_Edge(CEdgeP2C3D(new_id[pt], new_id[g.e1[ee]], // (point, camera)
						Eigen::Vector2d(g.z[2 * ee], g.z[2 * ee + 1]), t_info, system));
					if(!b_periodic)
						solver.Incremental_Step(r_edge);
					p_last_edge = &r_edge;
				}
				waiting[pt].clear();
			}
			if(b_periodic) {
				// one Incremental_Step() per camera, after all of its edges (a solve in the middle of a camera's edges would meet
				// landmarks with a single observation: the reference itself does not survive that)
				if(p_last_edge) {
					solver.Incremental_Step(*p_last_edge);
					TStates now = system.r_Vertex_Pool().For_Each(TStates());
					bool b_moved = false;
					for(size_t q = 0; q < last_states.size() && !b_moved; ++ q) // vertices are appended: compare the common prefix
						b_moved = now.v[q] != last_states[q];
					if(b_moved)
						solve_edges.push_back(system.r_Edge_Pool().n_Size() - 1);
					last_states.swap(now.v);
				}
				if(c + 1 == cams.size())
					chi2_trace.push_back(solver.f_Chi_Squared_Error_Denorm());
			} else if((c + 1) % n_batch == 0 || c + 1 == cams.size()) { // CONSISTENCY_MARKER
				double f_start = timer.f_Time();
				solver.Optimize(n_max_iter, f_min_dx);
				f_opt_time += timer.f_Time() - f_start;
				chi2_trace.push_back(solver.f_Chi_Squared_Error_Denorm());
			}
		}
	}
	if(getenv("SPP_REF_DUMP_TIMING"))
		solver.Dump(f_opt_time);
	TStates states = system.r_Vertex_Pool().For_Each(TStates()); // the functor travels by value
	uint64_t n_vertices = system.r_Vertex_Pool().n_Size(), n_edges = system.r_Edge_Pool().n_Size();
	spp_dump_f64(p_fw, "chi2_trace", chi2_trace.size(), &chi2_trace[0]);
	spp_dump_f64(p_fw, "states", states.v.size(), &states.v[0]);
	spp_dump_f64(p_fw, "optimize_time", 1, &f_opt_time);
	spp_dump_u64(p_fw, "n_vertices", 1, &n_vertices);
	spp_dump_u64(p_fw, "n_edges", 1, &n_edges);
	if(b_periodic) {
		uint64_t n_zero = 0;
		spp_dump_u64(p_fw, "solve_edges", solve_edges.size(), solve_edges.empty()? &n_zero : &solve_edges[0]);
	}
	if(b_marginals) { // diagonal blocks in vertex id order, row-major
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
		spp_dump_f64(p_fw, "cam_cov", cov6.size(), cov6.empty()? &f_zero : &cov6[0]);
		spp_dump_f64(p_fw, "pt_cov", cov3.size(), cov3.empty()? &f_zero : &cov3[0]);
	}
	printf("ref_driver_dropin_lm: %zu vertices, %zu edges, %zu optimisations, %.6f s in Optimize(), final chi2 %.17g\n",
		size_t(n_vertices), size_t(n_edges), chi2_trace.size() - (b_incremental? 0 : 1), f_opt_time, chi2_trace.back());
	return 0;
}

int main(int n_arg_num, const char **p_arg_list)
{
	if(n_arg_num < 5) {
		fprintf(stderr, "usage: %s <b200|ref> <batch|incremental> <graph.bin> <out.dump> [max_iter=5] [min_dx=0] [batch=10]\n", p_arg_list[0]);
		return -1;
	}
	const bool b_b200 = !strcmp(p_arg_list[1], "b200"), b_periodic = !strcmp(p_arg_list[2], "periodic");
	const bool b_incremental = b_periodic || !strcmp(p_arg_list[2], "incremental");
	const size_t n_max_iter = (n_arg_num > 5)? atol(p_arg_list[5]) : 5;
	const double f_min_dx = (n_arg_num > 6)? atof(p_arg_list[6]) : 0.0;
	const size_t n_batch = (n_arg_num > 7)? atol(p_arg_list[7]) : 10;
	spp_graph_t g;
	if(spp_graph_read(p_arg_list[3], &g) || g.kind != SPP_GRAPH_BA) {
		fprintf(stderr, "error: failed to read BA graph \'%s\'\n", p_arg_list[3]);
		return -1;
	}
	FILE *p_fw = fopen(p_arg_list[4], "wb");
	if(!p_fw)
		return -1;
	int n_result;
	try {
		if(b_b200)
			n_result = Run<CNonlinearSolver_Lambda_LM_B200<CSystemType, CLinearSolverType> >(g, b_incremental, b_periodic, p_fw, n_max_iter, f_min_dx, n_batch);
		else
			n_result = Run<CNonlinearSolver_Lambda_LM<CSystemType, CLinearSolverType> >(g, b_incremental, b_periodic, p_fw, n_max_iter, f_min_dx, n_batch);
	} catch(std::exception &r_exc) {
		fprintf(stderr, "error: %s\n", r_exc.what());
		n_result = -1;
	}
	fclose(p_fw);
	spp_graph_free(&g);
	return n_result;
}
