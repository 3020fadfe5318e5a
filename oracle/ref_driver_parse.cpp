/*
 * ref_driver_parse.cpp -- TEST INFRASTRUCTURE ONLY (oracle); never linked into the product.
 *
 * Runs the UNMODIFIED reference's graph-file parser (CParserTemplate, include/slam/Parser.h:1137-..., with the parse
 * primitives of include/slam_app/ParsePrimitives.h: CVertex2DParsePrimitive, CEdge2DParsePrimitive,
 * CVertexXYZParsePrimitive, CVertexCam3DParsePrimitive, CEdgeP2C3DParsePrimitive, CVertex3DParsePrimitive,
 * CEdge3DParsePrimitive, CEdge3DParsePrimitiveAxisAngle) on a text file and dumps what the
 * parse loop receives: vertex ids / states and edge ids / measurements / information matrices, in file order. This pins
 * the ingest of slam_plus_plus_b200/graphfile.py (parser-side camera pose inversion, edge inversion of descending 2D
 * edges, information matrix layouts) -- tests/golden/make_golden_parse.py, tests/test_graphfile_cpu.py.
 *
 * usage: ref_driver_parse <graph.txt> <out.dump>
 */

#include <stdio.h>
#include <vector>

#include "slam/Parser.h"
#include "slam/2DSolverBase.h"
#include "slam/3DSolverBase.h"
#include "slam_app/ParsePrimitives.h"
#include "spp_dump.h"

int n_dummy_param = 0;

class CRecordingParseLoop {
public:
	std::vector<double> v2, vxyz, vcam, v3, e2, ep2c, e3; // records: [id, state...] / [id0, id1, z..., info (row-major)...]

	void InitializeVertex(const CParserBase::TVertex2D &r_v)
	{
		v2.push_back(r_v.m_n_id);
		for(int i = 0; i < 3; ++ i) v2.push_back(r_v.m_v_position(i));
	}

	void InitializeVertex(const CParserBase::TVertexXYZ &r_v)
	{
		vxyz.push_back(r_v.m_n_id);
		for(int i = 0; i < 3; ++ i) vxyz.push_back(r_v.m_v_position(i));
	}

	void InitializeVertex(const CParserBase::TVertexCam3D &r_v)
	{
		vcam.push_back(r_v.m_n_id);
		for(int i = 0; i < 11; ++ i) vcam.push_back(r_v.m_v_position(i));
	}

	void InitializeVertex(const CParserBase::TVertex3D &r_v)
	{
		v3.push_back(r_v.m_n_id);
		for(int i = 0; i < 6; ++ i) v3.push_back(r_v.m_v_position(i));
	}

	void AppendSystem(const CParserBase::TEdge3D &r_e)
	{
		e3.push_back(double(r_e.m_n_node_0));
		e3.push_back(double(r_e.m_n_node_1));
		for(int i = 0; i < 6; ++ i) e3.push_back(r_e.m_v_delta(i));
		for(int i = 0; i < 6; ++ i)
			for(int j = 0; j < 6; ++ j) e3.push_back(r_e.m_t_inv_sigma(i, j));
	}

	void AppendSystem(const CParserBase::TEdge2D &r_e)
	{
		e2.push_back(double(r_e.m_n_node_0));
		e2.push_back(double(r_e.m_n_node_1));
		for(int i = 0; i < 3; ++ i) e2.push_back(r_e.m_v_delta(i));
		for(int i = 0; i < 3; ++ i)
			for(int j = 0; j < 3; ++ j) e2.push_back(r_e.m_t_inv_sigma(i, j));
	}

	void AppendSystem(const CParserBase::TEdgeP2C3D &r_e)
	{
		ep2c.push_back(double(r_e.m_n_node_0));
		ep2c.push_back(double(r_e.m_n_node_1));
		for(int i = 0; i < 2; ++ i) ep2c.push_back(r_e.m_v_delta(i));
		for(int i = 0; i < 2; ++ i)
			for(int j = 0; j < 2; ++ j) ep2c.push_back(r_e.m_t_inv_sigma(i, j));
	}
};

int main(int n_arg_num, const char **p_arg_list)
{
	if(n_arg_num < 3) {
		fprintf(stderr, "usage: ref_driver_parse <graph.txt> <out.dump>\n");
		return 2;
	}
	typedef MakeTypelist_Safe((CEdge2DParsePrimitive, CVertex2DParsePrimitive, CVertexXYZParsePrimitive,
		CVertexCam3DParsePrimitive, CEdgeP2C3DParsePrimitive, CVertex3DParsePrimitive, CEdge3DParsePrimitive,
		CEdge3DParsePrimitiveAxisAngle)) TPrimitives;
	CRecordingParseLoop loop;
	CParserTemplate<CRecordingParseLoop, TPrimitives> parser;
	if(!parser.Parse(p_arg_list[1], loop)) {
		fprintf(stderr, "ref_driver_parse: failed to parse %s\n", p_arg_list[1]);
		return 1;
	}
	FILE *p_fw = fopen(p_arg_list[2], "wb");
	if(!p_fw)
		return 1;
	double f_zero = 0;
	spp_dump_f64(p_fw, "vertex2d", loop.v2.size(), loop.v2.empty()? &f_zero : &loop.v2[0]);
	spp_dump_f64(p_fw, "vertex_xyz", loop.vxyz.size(), loop.vxyz.empty()? &f_zero : &loop.vxyz[0]);
	spp_dump_f64(p_fw, "vertex_cam", loop.vcam.size(), loop.vcam.empty()? &f_zero : &loop.vcam[0]);
	spp_dump_f64(p_fw, "edge2d", loop.e2.size(), loop.e2.empty()? &f_zero : &loop.e2[0]);
	spp_dump_f64(p_fw, "edge_p2c", loop.ep2c.size(), loop.ep2c.empty()? &f_zero : &loop.ep2c[0]);
	spp_dump_f64(p_fw, "vertex3d", loop.v3.size(), loop.v3.empty()? &f_zero : &loop.v3[0]);
	spp_dump_f64(p_fw, "edge3d", loop.e3.size(), loop.e3.empty()? &f_zero : &loop.e3[0]);
	fclose(p_fw);
	printf("ref_driver_parse: %zu 3D vertices, %zu 3D edges\n", loop.v3.size() / 7, loop.e3.size() / 44);
	printf("ref_driver_parse: %zu 2D vertices, %zu points, %zu cameras, %zu 2D edges, %zu projections\n", loop.v2.size() / 4,
		loop.vxyz.size() / 4, loop.vcam.size() / 12, loop.e2.size() / 14, loop.ep2c.size() / 8);
	return 0;
}
