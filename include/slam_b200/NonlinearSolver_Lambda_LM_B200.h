/*
 * NonlinearSolver_Lambda_LM_B200.h -- reference-side adapter for slot 3: a nonlinear solver type with the
 * interface of CNonlinearSolver_Lambda_LM (include/slam/NonlinearSolver_Lambda_LM.h:318-1116) for bundle-adjustment
 * systems, whose Optimize() runs ENTIRELY on the GPU through libspp_b200.so -- linearisation of the CEdgeP2C3D edges,
 * Levenberg-Marquardt control, landmark Schur complement, Cholesky of the reduced camera system, back-substitution,
 * update and chi2 (spp_ba_set_graph / spp_ba_optimize / spp_ba_get_states).
 *
 * Compiled INSIDE a SLAM++ build. The user keeps the reference's CFlatSystem with CVertexCam / CVertexXYZ vertices and
 * CEdgeP2C3D edges (include/slam/BA_Types.h:54-110,355-390,403-531) and changes one type:
 *
 *     typedef CNonlinearSolver_Lambda_LM_B200<CSystemType, CLinearSolverType> CNonlinearSolverType; // was CNonlinearSolver_Lambda_LM
 *     CNonlinearSolverType solver(system, TIncrementalSolveSetting(), TMarginalsComputationPolicy(), b_verbose,
 *         CLinearSolverType(), b_use_schur);
 *     solver.Optimize(n_max_iteration_num, f_min_dx_norm);
 *
 * It is a valid CNonlinearSolverType template-template argument of the application's solver list (ctor signature and the
 * solver_Has* / solver_Exports* traits of include/slam/NonlinearSolver_Lambda_LM.h:351-365; include/slam_app/Main.h:
 * 1108-1114,1376-1377). The system is read and written back through public accessors only: vertex pool For_Each with
 * r_v_State() / v_Intrinsics(), edge pool For_Each with n_Vertex_Id(), v_Measurement(), t_Sigma_Inv()
 * (include/slam/FlatSystem.h:1355-1366, include/slam/BaseTypes_Binary.h:344-366). Optimize() re-reads the system every
 * time it is called, so vertices and edges added since the last call are picked up: the marker-driven incremental
 * bundle adjustment of the application (CParseLoop_ConsistencyMarker, include/slam_app/IncBAParsePrimitives.h:154-168)
 * works unchanged. The CLinearSolver argument is accepted for interface compatibility and not used: the reduced camera
 * system is solved by the library (dense or block-sparse, spp_schur_set_rcs_solver).
 *
 * Marginal covariances: with TMarginalsComputationPolicy(true, ..., mpart_Diagonal, mpart_Diagonal) every Optimize()
 * ends as the reference's does (NonlinearSolver_Lambda_LM.h:1118-1350): the block diagonal of lambda^-1 at zero damping,
 * recovered from the Schur-complemented system on the device (spp_ba_marginals), is what
 * r_MarginalCovariance().r_SparseMatrix() holds afterwards (vertex id order). Other matrix parts are not provided.
 *
 * Not provided (the traits say so): Jacobian / Hessian export.
 */
#pragma once
#ifndef __NONLINEAR_SOLVER_LAMBDA_LM_B200_INCLUDED
#define __NONLINEAR_SOLVER_LAMBDA_LM_B200_INCLUDED

#include <stdexcept>
#include <new>
#include <vector>
#include <algorithm>
#include <string>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "slam/FlatSystem.h"         // reference
#include "slam/BA_Types.h"           // reference: CVertexCam, CVertexXYZ, CEdgeP2C3D
#include "slam/IncrementalPolicy.h"  // reference: TIncrementalSolveSetting, TMarginalsComputationPolicy
#include "slam/Timer.h"              // reference: CTimer
#include "slam/Marginals.h"          // reference: CMarginalCovariance
#include "slam/NonlinearSolver_Base.h" // reference: nonlinear_detail::CNonlinearSolver_Base (incremental policy, loop-closure detection)
#include "spp_b200.h"

template <class CSystem, class CLinearSolver, class CAMatrixBlockSizes = typename CSystem::_TyJacobianMatrixBlockList,
	class CLambdaMatrixBlockSizes = typename CSystem::_TyHessianMatrixBlockList>
class CNonlinearSolver_Lambda_LM_B200 : public nonlinear_detail::CNonlinearSolver_Base<CSystem, CLinearSolver, CAMatrixBlockSizes, false, true> {
public:
	typedef nonlinear_detail::CNonlinearSolver_Base<CSystem, CLinearSolver, CAMatrixBlockSizes, false, true> _TyBase; /**< @brief the reference's solver base: configuration, marginals cache, t_Incremental_Step() */
	typedef CSystem _TySystem; /**< @brief system type */
	typedef CLinearSolver _TyLinearSolver; /**< @brief linear solver type (unused) */
	typedef typename CSystem::_TyBaseVertex _TyBaseVertex; /**< @brief the data type for storing vertices */
	typedef typename CSystem::_TyBaseEdge _TyBaseEdge; /**< @brief the data type for storing measurements */

	/** solver interface properties (cf. NonlinearSolver_Lambda_LM.h:351-365) */
	enum {
		solver_HasDump = true,
		solver_HasChi2 = true,
		solver_HasMarginals = true,
		solver_HasGaussNewton = false,
		solver_HasLevenberg = true,
		solver_HasGradient = false,
		solver_HasSchur = true,
		solver_HasDelayedOptimization = false,
		solver_IsPreferredBatch = true,
		solver_IsPreferredIncremental = false,
		solver_ExportsJacobian = false,
		solver_ExportsHessian = false,
		solver_ExportsFactor = false
	};

protected:
	using _TyBase::m_r_system; // the system, the incremental / marginals configuration, the verbosity flag and the marginals
	using _TyBase::m_t_incremental_config; // cache live in the reference's base class
	using _TyBase::m_t_marginals_config;
	using _TyBase::m_b_verbose;
	using _TyBase::m_marginals;
	spp_ctx_t m_p_context; /**< @brief device context */
	size_t m_n_iteration_num; /**< @brief linear solves so far */
	size_t m_n_gathered_edge_num; /**< @brief edges already flattened (edges are immutable once added: only new ones are read) */
	double m_f_device_ms; /**< @brief device time spent in Optimize() so far */
	double m_f_upload_time, m_f_optimize_time, m_f_download_time; /**< @brief wall-clock split of Optimize() */
	double m_f_marginals_time; /**< @brief wall-clock time of the marginals recovery */
	size_t m_n_optimize_num; /**< @brief Optimize() calls that ran the solver (incremental drop-in test: same count as the reference) */
	size_t m_n_append_num; /**< @brief uploads that only sent the new vertices and edges (spp_ba_append_graph) */
	double m_f_gather_time, m_f_library_upload_time; /**< @brief split of the upload time: reading the system / the library call */

	bool m_b_uploaded; /**< @brief the device holds the system described by the arrays below */
	std::vector<uint8_t> m_vertex_type, m_prev_vertex_type;
	std::vector<double> m_cams, m_points, m_z, m_info, m_cam_states, m_prev_cams, m_prev_points;
	std::vector<uint64_t> m_obs_point, m_obs_camera;

	/** gathers the vertices: cameras 11 numbers (state 6 + intrinsics 5), points 3 */
	struct CGatherVertices {
		CNonlinearSolver_Lambda_LM_B200 &m_r;
		CGatherVertices(CNonlinearSolver_Lambda_LM_B200 &r) :m_r(r) {}
		void operator ()(const CVertexCam &r_vertex)
		{
			m_r.m_vertex_type.push_back(0);
			for(int i = 0; i < 6; ++ i) m_r.m_cams.push_back(r_vertex.r_v_State()(i));
			for(int i = 0; i < 5; ++ i) m_r.m_cams.push_back(r_vertex.v_Intrinsics()(i));
		}
		void operator ()(const CVertexXYZ &r_vertex)
		{
			m_r.m_vertex_type.push_back(1);
			for(int i = 0; i < 3; ++ i) m_r.m_points.push_back(r_vertex.r_v_State()(i));
		}
	};

	/** gathers the observations in edge insertion order */
	struct CGatherEdges {
		CNonlinearSolver_Lambda_LM_B200 &m_r;
		CGatherEdges(CNonlinearSolver_Lambda_LM_B200 &r) :m_r(r) {}
		void operator ()(const CEdgeP2C3D &r_edge)
		{
			m_r.m_obs_camera.push_back(r_edge.n_Vertex_Id(0)); // vertex 0 is the camera (BA_Types.h:403)
			m_r.m_obs_point.push_back(r_edge.n_Vertex_Id(1));
			for(int i = 0; i < 2; ++ i) m_r.m_z.push_back(r_edge.v_Measurement()(i));
			for(int i = 0; i < 2; ++ i)
				for(int j = 0; j < 2; ++ j) m_r.m_info.push_back(r_edge.t_Sigma_Inv()(i, j));
		}
	};

	/** writes the optimized states back into the system */
	struct CScatterVertices {
		const double *m_p_cam, *m_p_point;
		CScatterVertices(const double *p_cam, const double *p_point) :m_p_cam(p_cam), m_p_point(p_point) {}
		void operator ()(CVertexCam &r_vertex)
		{
			for(int i = 0; i < 6; ++ i) r_vertex.r_v_State()(i) = m_p_cam[i];
			m_p_cam += 6;
		}
		void operator ()(CVertexXYZ &r_vertex)
		{
			for(int i = 0; i < 3; ++ i) r_vertex.r_v_State()(i) = m_p_point[i];
			m_p_point += 3;
		}
	};

public:
	/** same arguments as CNonlinearSolver_Lambda_LM (NonlinearSolver_Lambda_LM.h:470-492) */
	CNonlinearSolver_Lambda_LM_B200(CSystem &r_system,
		TIncrementalSolveSetting t_incremental_config = TIncrementalSolveSetting(),
		TMarginalsComputationPolicy t_marginals_config = TMarginalsComputationPolicy(),
		bool b_verbose = false, CLinearSolver linear_solver = CLinearSolver(), bool UNUSED(b_use_schur) = true,
		int n_device = 0)
		:_TyBase(r_system, t_incremental_config, t_marginals_config, b_verbose, linear_solver, false),
		m_p_context(0), m_n_iteration_num(0),
		m_n_gathered_edge_num(0), m_f_device_ms(0), m_f_upload_time(0), m_f_optimize_time(0), m_f_download_time(0),
		m_f_marginals_time(0), m_n_optimize_num(0), m_n_append_num(0), m_f_gather_time(0), m_f_library_upload_time(0), m_b_uploaded(false)
	{
		if(t_marginals_config.b_calculate && t_marginals_config.n_relinearize_policy != mpart_Diagonal)
			throw std::runtime_error("CNonlinearSolver_Lambda_LM_B200: only the block diagonal of the marginal covariances (mpart_Diagonal) is provided");
		Check(spp_create(n_device, &m_p_context));
	}

	~CNonlinearSolver_Lambda_LM_B200()
	{
		spp_destroy(m_p_context);
	}

	// t_IncrementalConfig(), t_MarginalsPolicy(), r_MarginalCovariance(): inherited (NonlinearSolver_Base.h:466-473,740-763)

	/** number of Optimize() calls that ran the LM loop so far */
	inline size_t n_Optimize_Num() const
	{
		return m_n_optimize_num;
	}

	/** number of uploads that only sent what was new (incremental use) */
	inline size_t n_Append_Num() const
	{
		return m_n_append_num;
	}

	/** the device context, e.g. for spp_schur_set_rcs_solver() */
	inline spp_ctx_t p_Context()
	{
		return m_p_context;
	}

	/** timing statistics (cf. CNonlinearSolver_Lambda_LM::Dump, LM.h:547-...) */
	void Dump(double f_total_time = -1) const
	{
		printf("solver took " PRIsize " iterations\n", m_n_iteration_num); // debug, to be able to say we didn't botch it numerically
		if(f_total_time > 0)
			printf("solver spent %f seconds in parallelizable section (updating lambda; disparity %g seconds)\n",
				m_f_device_ms * 1e-3, f_total_time - m_f_device_ms * 1e-3);
		printf("out of which:\n\tdevice (libspp_b200: lambda, rhs, schur, linsolve, update, chi2): %f\n", m_f_device_ms * 1e-3);
		printf("host side of Optimize(): flatten + upload + structure %f (" PRIsize " appends of new vertices / edges only), spp_ba_optimize %f, download + write-back %f\n",
			m_f_upload_time, m_n_append_num, m_f_optimize_time, m_f_download_time);
		printf("\tupload: reading the system %f, library (spp_ba_set_graph / spp_ba_append_graph / spp_ba_set_states) %f\n",
			m_f_gather_time, m_f_library_upload_time);
		if(m_t_marginals_config.b_calculate)
			printf("solver spent %f seconds in marginals (spp_ba_marginals + the block matrix)\n", m_f_marginals_time);
	}

	/** f_Chi_Squared_Error_Denorm (NonlinearSolver_Base.h:278-297) of the system as it is now */
	double f_Chi_Squared_Error_Denorm() // throw(std::bad_alloc, std::runtime_error)
	{
		Upload();
		double f_chi2 = 0;
		Check(spp_ba_chi2(m_p_context, &f_chi2));
		return f_chi2;
	}

	/** incremental optimization function: CNonlinearSolver_Lambda_LM::Incremental_Step (LM.h:671-760) on top of the
	 *	reference's own period counting and loop-closure detection (the inherited t_Incremental_Step,
	 *	NonlinearSolver_Base.h:557-622): a nonlinear solve when the nonlinear period elapsed after a loop closure, a
	 *	single step (Optimize(1, 0)) for the linear period, marginals only (Optimize(0, 0)) when vertices were added
	 *	without a solve and the marginals policy is on */
	void Incremental_Step(_TyBaseEdge &r_last_edge) // throw(std::bad_alloc, std::runtime_error)
	{
		std::pair<bool, int> t_optimize = this->t_Incremental_Step(r_last_edge);
		if(t_optimize.second == 2)
			Optimize(m_t_incremental_config.n_max_nonlinear_iteration_num, m_t_incremental_config.f_nonlinear_error_thresh);
		else if(t_optimize.second == 1)
			Optimize(1, 0);
		if(t_optimize.first && !t_optimize.second && m_t_marginals_config.b_calculate)
			Optimize(0, 0);
	}

	/** CNonlinearSolver_Lambda_LM::Optimize (LM.h:796-1116) on the device; the system receives the optimized states */
	void Optimize(size_t n_max_iteration_num = 5, double f_min_dx_norm = .01) // throw(std::bad_alloc, std::runtime_error)
	{
		if(m_r_system.r_Edge_Pool().b_Empty())
			return; // nothing to optimize
		CTimer timer;
		double f_t0 = timer.f_Time();
		Upload();
		double f_t1 = timer.f_Time();
		if(!n_max_iteration_num) { // Optimize(0, 0): the marginals follow the system, no solve (LM.h:752-754)
			m_f_upload_time += f_t1 - f_t0;
			if(m_t_marginals_config.b_calculate)
				Calculate_Marginals();
			return;
		}
		spp_report_t t_report;
		Check(spp_ba_optimize(m_p_context, n_max_iteration_num, f_min_dx_norm, &t_report));
		++ m_n_optimize_num;
		double f_t2 = timer.f_Time();
		m_f_upload_time += f_t1 - f_t0;
		m_f_optimize_time += f_t2 - f_t1;
		m_n_iteration_num += t_report.n_iterations;
		m_f_device_ms += t_report.ms_total;
		if(m_b_verbose) {
			for(int i = 0; i < t_report.n_iterations && i < SPP_MAX_TRACE; ++ i) {
				printf("chi2: %f%s, alpha %g, residual norm: %.4f\n", t_report.trace_chi2[i],
					(t_report.trace_accepted[i])? "" : " (rising: step rejected)", t_report.trace_alpha[i], t_report.trace_dx_norm[i]);
			}
		}
		if(t_report.status == SPP_NOT_POSDEF)
			fprintf(stderr, "warning: Cholesky failed\n"); // as the reference (the loop stops)
		m_cam_states.resize((m_cams.size() / 11) * 6);
		// (all vertices: on a context that was given a communicator -- spp_set_nccl(p_Context(), ...) on every rank, each
		// with the whole system -- the landmark slices of the ranks are summed into one array on the device)
		Check(spp_ba_gather_states(m_p_context, m_cam_states.empty()? 0 : &m_cam_states[0], m_points.empty()? 0 : &m_points[0]));
		m_r_system.r_Vertex_Pool().For_Each(CScatterVertices(m_cam_states.empty()? 0 : &m_cam_states[0],
			m_points.empty()? 0 : &m_points[0]));
		for(size_t i = 0, n = m_cams.size() / 11; i < n; ++ i) // the cached copy follows: the device and the system agree
			for(int j = 0; j < 6; ++ j) m_cams[i * 11 + j] = m_cam_states[i * 6 + j];
		m_f_download_time += timer.f_Time() - f_t2;
		if(m_t_marginals_config.b_calculate)
			Calculate_Marginals();
	}

	/** block diagonal of lambda^-1 at the current states and zero damping -> r_MarginalCovariance()
	 *	(NonlinearSolver_Lambda_LM.h:1118-1350 with mpart_Diagonal -> BAMarginals.h:579-760) */
	void Calculate_Marginals() // throw(std::bad_alloc, std::runtime_error)
	{
		CTimer timer;
		const size_t n_cam_num = m_cams.size() / 11, n_point_num = m_points.size() / 3;
		std::vector<double> cam_cov(n_cam_num * 36), point_cov(n_point_num * 9);
		Check(spp_ba_marginals(m_p_context, 0.0, cam_cov.empty()? 0 : &cam_cov[0], point_cov.empty()? 0 : &point_cov[0]));
		CUberBlockMatrix margs;
		size_t n_cam = 0, n_point = 0;
		for(size_t i = 0, n = m_vertex_type.size(); i < n; ++ i) {
			if(m_vertex_type[i] == 0) {
				Eigen::Map<const Eigen::Matrix<double, 6, 6> > t_block(&cam_cov[n_cam * 36]);
				++ n_cam;
				margs.t_GetBlock_Log(i, i, 6, 6, true, false) = t_block;
			} else {
				Eigen::Map<const Eigen::Matrix<double, 3, 3> > t_block(&point_cov[n_point * 9]);
				++ n_point;
				margs.t_GetBlock_Log(i, i, 3, 3, true, false) = t_block;
			}
		}
		m_marginals.Swap_SparseMatrix(margs);
		m_marginals.EnableUpdate();
		m_marginals.Set_Edge_Num(m_r_system.r_Edge_Pool().n_Size());
		m_f_marginals_time += timer.f_Time();
	}

protected:
	/** flattens the system and hands it to the library; what is already on the device is not sent again */
	void Upload() // throw(std::bad_alloc, std::runtime_error)
	{
		CTimer upload_timer;
		struct CSplitTimes { // whichever way Upload() is left: the rest of its time was spent in the library
			CTimer &m_r_timer; double &m_r_f_gather, &m_r_f_library; double m_f_gathered;
			CSplitTimes(CTimer &r_timer, double &r_f_gather, double &r_f_library)
				:m_r_timer(r_timer), m_r_f_gather(r_f_gather), m_r_f_library(r_f_library), m_f_gathered(0) {}
			~CSplitTimes() { m_r_f_gather += m_f_gathered; m_r_f_library += m_r_timer.f_Time() - m_f_gathered; }
		} split_times(upload_timer, m_f_gather_time, m_f_library_upload_time);
		m_prev_vertex_type.swap(m_vertex_type); m_prev_cams.swap(m_cams); m_prev_points.swap(m_points);
		m_vertex_type.clear(); m_cams.clear(); m_points.clear();
		m_vertex_type.reserve(m_prev_vertex_type.size() + 1024);
		m_cams.reserve(m_prev_cams.size() + 1024);
		m_points.reserve(m_prev_points.size() + 4096);
		m_r_system.r_Vertex_Pool().For_Each(CGatherVertices(*this)); // the states may have been changed by the caller
		split_times.m_f_gathered = upload_timer.f_Time();
		const size_t n_edge_num = m_r_system.r_Edge_Pool().n_Size();
		const bool b_same_structure = m_b_uploaded && n_edge_num == m_n_gathered_edge_num && m_vertex_type == m_prev_vertex_type;
		if(b_same_structure) {
			if(m_cams == m_prev_cams && m_points == m_prev_points)
				return; // the device holds exactly this system
			bool b_same_intrinsics = true;
			for(size_t i = 0, n = m_cams.size() / 11; i < n && b_same_intrinsics; ++ i)
				for(int j = 6; j < 11; ++ j) b_same_intrinsics = b_same_intrinsics && m_cams[i * 11 + j] == m_prev_cams[i * 11 + j];
			if(b_same_intrinsics) { // only the states moved
				m_cam_states.resize((m_cams.size() / 11) * 6);
				for(size_t i = 0, n = m_cams.size() / 11; i < n; ++ i)
					for(int j = 0; j < 6; ++ j) m_cam_states[i * 6 + j] = m_cams[i * 11 + j];
				Check(spp_ba_set_states(m_p_context, m_cam_states.empty()? 0 : &m_cam_states[0], m_points.empty()? 0 : &m_points[0]));
				return;
			}
		}
		if(n_edge_num < m_n_gathered_edge_num) { // not an append-only change: start over
			m_obs_point.clear(); m_obs_camera.clear(); m_z.clear(); m_info.clear();
			m_n_gathered_edge_num = 0;
			m_b_uploaded = false;
		}
		const size_t n_prev_edge_num = m_n_gathered_edge_num;
		if(n_edge_num > m_n_gathered_edge_num)
			m_r_system.r_Edge_Pool().For_Each(m_n_gathered_edge_num, n_edge_num, CGatherEdges(*this)); // the new edges only
		m_n_gathered_edge_num = n_edge_num;
		split_times.m_f_gathered = upload_timer.f_Time();
		// the system only grew (incremental BA: vertices and edges are appended between two Optimize() calls): the new
		// vertices and edges alone go to the device, the structure is rebuilt there (spp_ba_append_graph)
		const size_t n_prev_vertex_num = m_prev_vertex_type.size(), n_prev_cam_num = m_prev_cams.size() / 11,
			n_prev_point_num = m_prev_points.size() / 3;
		bool b_grew = m_b_uploaded && !getenv("SPP_ADAPTER_NO_APPEND") && m_vertex_type.size() >= n_prev_vertex_num &&
			std::equal(m_prev_vertex_type.begin(), m_prev_vertex_type.end(), m_vertex_type.begin());
		for(size_t i = 0; i < n_prev_cam_num && b_grew; ++ i) // (the intrinsics of the cameras already there must not have changed)
			for(int j = 6; j < 11; ++ j) b_grew = b_grew && m_cams[i * 11 + j] == m_prev_cams[i * 11 + j];
		if(b_grew) {
			const size_t n_new_vertex_num = m_vertex_type.size() - n_prev_vertex_num, n_new_edge_num = n_edge_num - n_prev_edge_num;
			int n_result = spp_ba_append_graph(m_p_context, n_new_vertex_num, (n_new_vertex_num)? &m_vertex_type[n_prev_vertex_num] : 0,
				(m_cams.size() > n_prev_cam_num * 11)? &m_cams[n_prev_cam_num * 11] : 0,
				(m_points.size() > n_prev_point_num * 3)? &m_points[n_prev_point_num * 3] : 0, n_new_edge_num,
				(n_new_edge_num)? &m_obs_point[n_prev_edge_num] : 0, (n_new_edge_num)? &m_obs_camera[n_prev_edge_num] : 0,
				(n_new_edge_num)? &m_z[n_prev_edge_num * 2] : 0, (n_new_edge_num)? &m_info[n_prev_edge_num * 4] : 0);
			if(n_result == SPP_OK) {
				++ m_n_append_num;
				bool b_same_states = std::equal(m_prev_points.begin(), m_prev_points.end(), m_points.begin());
				for(size_t i = 0; i < n_prev_cam_num && b_same_states; ++ i)
					for(int j = 0; j < 6; ++ j) b_same_states = b_same_states && m_cams[i * 11 + j] == m_prev_cams[i * 11 + j];
				if(!b_same_states) { // the caller moved vertices that are on the device already
					m_cam_states.resize((m_cams.size() / 11) * 6);
					for(size_t i = 0, n = m_cams.size() / 11; i < n; ++ i)
						for(int j = 0; j < 6; ++ j) m_cam_states[i * 6 + j] = m_cams[i * 11 + j];
					Check(spp_ba_set_states(m_p_context, m_cam_states.empty()? 0 : &m_cam_states[0], m_points.empty()? 0 : &m_points[0]));
				}
				return;
			}
			if(n_result != SPP_ERR_INVALID) // (invalid: this context cannot append -- e.g. several ranks -- the whole graph follows)
				Check(n_result);
		}
		Check(spp_ba_set_graph(m_p_context, m_vertex_type.size(), m_vertex_type.empty()? 0 : &m_vertex_type[0],
			m_cams.empty()? 0 : &m_cams[0], m_points.empty()? 0 : &m_points[0], m_obs_point.size(),
			m_obs_point.empty()? 0 : &m_obs_point[0], m_obs_camera.empty()? 0 : &m_obs_camera[0],
			m_z.empty()? 0 : &m_z[0], m_info.empty()? 0 : &m_info[0]));
		m_b_uploaded = true;
	}

	void Check(int n_result) const // throw(std::bad_alloc, std::runtime_error)
	{
		if(n_result == SPP_OK)
			return;
		if(n_result == SPP_ERR_NOMEM)
			throw std::bad_alloc();
		throw std::runtime_error(std::string("libspp_b200: ") + spp_last_error(m_p_context));
	}

private:
	CNonlinearSolver_Lambda_LM_B200(const CNonlinearSolver_Lambda_LM_B200 &r_solver); // no copy
	CNonlinearSolver_Lambda_LM_B200 &operator =(const CNonlinearSolver_Lambda_LM_B200 &r_solver); // no copy
};

#endif // !__NONLINEAR_SOLVER_LAMBDA_LM_B200_INCLUDED
