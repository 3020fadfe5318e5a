/*
 * NonlinearSolver_Lambda_B200.h -- slot 3 for pose graphs: a nonlinear solver TYPE with the interface of the reference's
 * CNonlinearSolver_Lambda (include/slam/NonlinearSolver_Lambda.h:130-667, Gauss-Newton on the information form), to be
 * used where an application names that type (include/slam_app/Main.h:1108-1114, src/slam_simple_example/Main.cpp):
 *
 *     typedef CNonlinearSolver_Lambda_B200<CSystemType, CLinearSolverType> CNonlinearSolverType;   // was CNonlinearSolver_Lambda
 *
 * The system stays the reference's CFlatSystem with its own vertex and edge types -- CVertexPose2D / CEdgePose2D
 * (SE2_Types.h) or CVertexPose3D / CEdgePose3D (SE3_Types.h): new vertices are still initialised by the edge
 * constructors on the host. Optimize() flattens the system through the public pool iterators (states, vertex ids,
 * measurements, information matrices), runs the whole Gauss-Newton loop on the GPU (linearisation, block Cholesky,
 * update, chi2: spp_pose_set_graph / spp_pose_optimize) and writes the states back.
 *
 * Incremental_Step() follows CNonlinearSolver_Base::t_Incremental_Step (NonlinearSolver_Base.h:502-622, the default
 * __SLAM_COUNT_ITERATIONS_AS_VERTICES counting): loop-closure detection on the last edge, a nonlinear solve every
 * t_nonlinear_freq new vertices or a single linear step every t_linear_freq, only when a loop was closed since the last
 * one -- so the incremental configurations of slam_app ("-nsp N" / "-lsp N") behave as with the reference's solver. Every
 * solve re-reads the system (only the new edges) and runs a batch factorisation on the device; the incremental factor
 * refresh of CNonlinearSolver_FastL is not provided.
 *
 * Marginal covariances: policy mpart_Diagonal as in NonlinearSolver_Lambda.h:669-767 (spp_pose_marginals).
 */
#pragma once
#ifndef __NONLINEAR_SOLVER_LAMBDA_B200_INCLUDED
#define __NONLINEAR_SOLVER_LAMBDA_B200_INCLUDED

#include <stdexcept>
#include <new>
#include <vector>
#include <string>
#include <algorithm>
#include <stdint.h>
#include <stdio.h>

#include "slam/FlatSystem.h"         // reference
#include "slam/ConfigSolvers.h"      // reference: __SLAM_COUNT_ITERATIONS_AS_VERTICES
#include "slam/IncrementalPolicy.h"  // reference: TIncrementalSolveSetting, TMarginalsComputationPolicy
#include "slam/Timer.h"              // reference: CTimer
#include "slam/Marginals.h"          // reference: CMarginalCovariance
#include "slam/NonlinearSolver_Base.h" // reference: nonlinear_detail::CNonlinearSolver_Base (incremental policy, loop-closure detection)
#include "spp_b200.h"

template <class CSystem, class CLinearSolver, class CAMatrixBlockSizes = typename CSystem::_TyJacobianMatrixBlockList,
	class CLambdaMatrixBlockSizes = typename CSystem::_TyHessianMatrixBlockList>
class CNonlinearSolver_Lambda_B200 : public nonlinear_detail::CNonlinearSolver_Base<CSystem, CLinearSolver, CAMatrixBlockSizes, false, true> {
public:
	typedef nonlinear_detail::CNonlinearSolver_Base<CSystem, CLinearSolver, CAMatrixBlockSizes, false, true> _TyBase; /**< @brief the reference's solver base: configuration, marginals cache, t_Incremental_Step() */
	typedef CSystem _TySystem; /**< @brief system type */
	typedef CLinearSolver _TyLinearSolver; /**< @brief linear solver type (unused) */
	typedef typename CSystem::_TyBaseVertex _TyBaseVertex; /**< @brief the data type for storing vertices */
	typedef typename CSystem::_TyBaseEdge _TyBaseEdge; /**< @brief the data type for storing measurements */

	/** solver interface properties (cf. NonlinearSolver_Lambda.h:141-157) */
	enum {
		solver_HasDump = true,
		solver_HasChi2 = true,
		solver_HasMarginals = true,
		solver_HasGaussNewton = true,
		solver_HasLevenberg = false,
		solver_HasGradient = false,
		solver_HasSchur = false,
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
	size_t m_n_optimize_num; /**< @brief calls of Optimize() that reached the device */
	size_t m_n_gathered_edge_num; /**< @brief edges already flattened (edges are immutable once added) */
	double m_f_device_ms, m_f_upload_time, m_f_optimize_time, m_f_download_time, m_f_marginals_time;

	int m_n_dim; /**< @brief 3 (SE(2)) or 6 (SE(3)); 0 until the first vertex is seen */
	bool m_b_uploaded; /**< @brief the device holds the graph described by the arrays below */
	size_t m_n_uploaded_vertex_num, m_n_uploaded_edge_num;
	std::vector<double> m_states, m_z, m_info;
	std::vector<uint64_t> m_from, m_to;

	struct CGatherVertices { /**< gathers the vertex states in id order */
		CNonlinearSolver_Lambda_B200 &m_r;
		CGatherVertices(CNonlinearSolver_Lambda_B200 &r) :m_r(r) {}
		template <class CVertex>
		void operator ()(const CVertex &r_vertex)
		{
			const int n_dim = int(r_vertex.r_v_State().rows());
			if(!m_r.m_n_dim)
				m_r.m_n_dim = n_dim;
			if(n_dim != m_r.m_n_dim || (n_dim != 3 && n_dim != 6))
				throw std::runtime_error("CNonlinearSolver_Lambda_B200: pose graphs of one vertex type, SE(2) or SE(3), only");
			for(int i = 0; i < n_dim; ++ i) m_r.m_states.push_back(r_vertex.r_v_State()(i));
		}
	};

	struct CGatherEdges { /**< gathers the measurements in edge insertion order */
		CNonlinearSolver_Lambda_B200 &m_r;
		CGatherEdges(CNonlinearSolver_Lambda_B200 &r) :m_r(r) {}
		template <class CEdge>
		void operator ()(const CEdge &r_edge)
		{
			const int n_dim = int(r_edge.v_Measurement().rows());
			if(r_edge.n_Vertex_Num() != 2 || n_dim != m_r.m_n_dim)
				throw std::runtime_error("CNonlinearSolver_Lambda_B200: binary pose-pose edges only");
			m_r.m_from.push_back(r_edge.n_Vertex_Id(0));
			m_r.m_to.push_back(r_edge.n_Vertex_Id(1));
			for(int i = 0; i < n_dim; ++ i) m_r.m_z.push_back(r_edge.v_Measurement()(i));
			for(int i = 0; i < n_dim; ++ i)
				for(int j = 0; j < n_dim; ++ j) m_r.m_info.push_back(r_edge.t_Sigma_Inv()(i, j));
		}
	};

	struct CScatterVertices { /**< writes the optimized states back into the system */
		const double *m_p_state;
		CScatterVertices(const double *p_state) :m_p_state(p_state) {}
		template <class CVertex>
		void operator ()(CVertex &r_vertex)
		{
			const int n_dim = int(r_vertex.r_v_State().rows());
			for(int i = 0; i < n_dim; ++ i) r_vertex.r_v_State()(i) = m_p_state[i];
			m_p_state += n_dim;
		}
	};

public:
	/** same arguments as CNonlinearSolver_Lambda (NonlinearSolver_Lambda.h:159-176) */
	CNonlinearSolver_Lambda_B200(CSystem &r_system,
		TIncrementalSolveSetting t_incremental_config = TIncrementalSolveSetting(),
		TMarginalsComputationPolicy t_marginals_config = TMarginalsComputationPolicy(),
		bool b_verbose = false, CLinearSolver linear_solver = CLinearSolver(), bool UNUSED(b_use_schur) = false,
		int n_device = 0)
		:_TyBase(r_system, t_incremental_config, t_marginals_config, b_verbose, linear_solver, false),
		m_p_context(0), m_n_iteration_num(0), m_n_optimize_num(0), m_n_gathered_edge_num(0), m_f_device_ms(0),
		m_f_upload_time(0), m_f_optimize_time(0), m_f_download_time(0), m_f_marginals_time(0), m_n_dim(0),
		m_b_uploaded(false), m_n_uploaded_vertex_num(0), m_n_uploaded_edge_num(0)
	{
		if(t_marginals_config.b_calculate && t_marginals_config.n_relinearize_policy != mpart_Diagonal)
			throw std::runtime_error("CNonlinearSolver_Lambda_B200: only the block diagonal of the marginal covariances (mpart_Diagonal) is provided");
		Check(spp_create(n_device, &m_p_context));
	}

	~CNonlinearSolver_Lambda_B200()
	{
		spp_destroy(m_p_context);
	}

	// t_IncrementalConfig(), t_MarginalsPolicy(), r_MarginalCovariance(): inherited (NonlinearSolver_Base.h:466-473,740-763)

	/** the device context, e.g. for spp_pose_set_ordering() */
	inline spp_ctx_t p_Context()
	{
		return m_p_context;
	}

	/** number of Optimize() calls that ran on the device (the incremental tests compare it with the reference's) */
	inline size_t n_Optimize_Num() const
	{
		return m_n_optimize_num;
	}

	/** timing statistics (cf. CNonlinearSolver_Lambda::Dump, NonlinearSolver_Lambda.h:246-...) */
	void Dump(double f_total_time = -1) const
	{
		printf("solver took " PRIsize " iterations\n", m_n_iteration_num); // debug, to be able to say we didn't botch it numerically
		if(f_total_time > 0)
			printf("solver spent %f seconds in parallelizable section (disparity %g seconds)\n",
				m_f_device_ms * 1e-3, f_total_time - m_f_device_ms * 1e-3);
		printf("out of which:\n\tdevice (libspp_b200: lambda, rhs, Cholesky, update, chi2): %f\n", m_f_device_ms * 1e-3);
		printf("host side of Optimize() (" PRIsize " calls): flatten + upload + ordering %f, spp_pose_optimize %f, download + write-back %f\n",
			m_n_optimize_num, m_f_upload_time, m_f_optimize_time, m_f_download_time);
		if(m_t_marginals_config.b_calculate)
			printf("solver spent %f seconds in marginals\n", m_f_marginals_time);
	}

	/** f_Chi_Squared_Error_Denorm (NonlinearSolver_Base.h:278-297) of the system as it is now */
	double f_Chi_Squared_Error_Denorm() // throw(std::bad_alloc, std::runtime_error)
	{
		if(m_r_system.r_Edge_Pool().b_Empty())
			return 0;
		Upload();
		double f_chi2 = 0;
		Check(spp_pose_chi2(m_p_context, &f_chi2));
		return f_chi2;
	}

	/** incremental optimization function: CNonlinearSolver_Lambda::Incremental_Step (NonlinearSolver_Lambda.h:334-441) on top
	 *	of the reference's own period counting and loop-closure detection (the inherited t_Incremental_Step,
	 *	NonlinearSolver_Base.h:557-622) */
	void Incremental_Step(_TyBaseEdge &r_last_edge) // throw(std::bad_alloc, std::runtime_error)
	{
		std::pair<bool, int> t_optimize = this->t_Incremental_Step(r_last_edge);
		if(t_optimize.second == 2)
			Optimize(m_t_incremental_config.n_max_nonlinear_iteration_num, m_t_incremental_config.f_nonlinear_error_thresh);
		else if(t_optimize.second == 1)
			Optimize(1, 0);
		if(t_optimize.first && !t_optimize.second && m_t_marginals_config.b_calculate)
			Optimize(0, 0); // the marginals follow the system (NonlinearSolver_Lambda.h:436-437)
	}

	/** CNonlinearSolver_Lambda::Optimize (NonlinearSolver_Lambda.h:476-667) on the device; the system receives the states */
	void Optimize(size_t n_max_iteration_num = 5, double f_min_dx_norm = .01) // throw(std::bad_alloc, std::runtime_error)
	{
		if(m_r_system.r_Edge_Pool().b_Empty())
			return; // nothing to optimize
		CTimer timer;
		double f_t0 = timer.f_Time();
		Upload();
		double f_t1 = timer.f_Time();
		if(n_max_iteration_num) {
			spp_report_t t_report;
			Check(spp_pose_optimize(m_p_context, n_max_iteration_num, f_min_dx_norm, &t_report));
			++ m_n_optimize_num;
			m_n_iteration_num += t_report.n_iterations;
			m_f_device_ms += t_report.ms_total;
			if(m_b_verbose) {
				for(int i = 0; i < t_report.n_iterations && i < SPP_MAX_TRACE; ++ i)
					printf("residual norm: %.4f\n", t_report.trace_dx_norm[i]);
			}
			if(t_report.status == SPP_NOT_POSDEF)
				fprintf(stderr, "warning: Cholesky failed\n"); // as the reference (the loop stops)
		}
		double f_t2 = timer.f_Time();
		if(n_max_iteration_num) {
			Check(spp_pose_get_states(m_p_context, &m_states[0]));
			m_r_system.r_Vertex_Pool().For_Each(CScatterVertices(&m_states[0]));
		}
		m_f_upload_time += f_t1 - f_t0;
		m_f_optimize_time += f_t2 - f_t1;
		m_f_download_time += timer.f_Time() - f_t2;
		if(m_t_marginals_config.b_calculate)
			Calculate_Marginals();
	}

	/** block diagonal of lambda^-1 at the current states -> r_MarginalCovariance() (NonlinearSolver_Lambda.h:669-767) */
	void Calculate_Marginals() // throw(std::bad_alloc, std::runtime_error)
	{
		CTimer timer;
		const size_t n_vertex_num = m_states.size() / m_n_dim, n_block = size_t(m_n_dim) * m_n_dim;
		std::vector<double> cov(n_vertex_num * n_block);
		Check(spp_pose_marginals(m_p_context, &cov[0]));
		CUberBlockMatrix margs;
		for(size_t i = 0; i < n_vertex_num; ++ i) {
			Eigen::Map<const Eigen::MatrixXd> t_block(&cov[i * n_block], m_n_dim, m_n_dim); // symmetric
			margs.t_GetBlock_Log(i, i, m_n_dim, m_n_dim, true, false) = t_block;
		}
		m_marginals.Swap_SparseMatrix(margs);
		m_marginals.EnableUpdate();
		m_marginals.Set_Edge_Num(m_r_system.r_Edge_Pool().n_Size());
		m_f_marginals_time += timer.f_Time();
	}

protected:
	/** flattens the system and hands it to the library; a graph the device already holds only gets its states refreshed */
	void Upload() // throw(std::bad_alloc, std::runtime_error)
	{
		const size_t n_vertex_num = m_r_system.r_Vertex_Pool().n_Size(), n_edge_num = m_r_system.r_Edge_Pool().n_Size();
		m_states.clear();
		m_r_system.r_Vertex_Pool().For_Each(CGatherVertices(*this));
		if(n_edge_num < m_n_gathered_edge_num) { // not an append-only change: start over
			m_from.clear(); m_to.clear(); m_z.clear(); m_info.clear();
			m_n_gathered_edge_num = 0;
			m_b_uploaded = false;
		}
		if(n_edge_num > m_n_gathered_edge_num)
			m_r_system.r_Edge_Pool().For_Each(m_n_gathered_edge_num, n_edge_num, CGatherEdges(*this)); // the new edges only
		m_n_gathered_edge_num = n_edge_num;
		if(m_b_uploaded && n_vertex_num == m_n_uploaded_vertex_num && n_edge_num == m_n_uploaded_edge_num) {
			Check(spp_pose_set_states(m_p_context, &m_states[0]));
			return;
		}
		Check(spp_pose_set_graph(m_p_context, m_n_dim, n_vertex_num, &m_states[0], n_edge_num, &m_from[0], &m_to[0],
			&m_z[0], &m_info[0]));
		m_b_uploaded = true;
		m_n_uploaded_vertex_num = n_vertex_num;
		m_n_uploaded_edge_num = n_edge_num;
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
	CNonlinearSolver_Lambda_B200(const CNonlinearSolver_Lambda_B200 &r_solver); // no copy
	CNonlinearSolver_Lambda_B200 &operator =(const CNonlinearSolver_Lambda_B200 &r_solver); // no copy
};

#endif // !__NONLINEAR_SOLVER_LAMBDA_B200_INCLUDED
