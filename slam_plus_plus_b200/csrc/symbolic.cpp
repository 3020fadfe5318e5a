// symbolic.cpp -- host-side symbolic analysis of the landmark Schur system (integer work, one-time).
//
// Replaces the structural part of lambda_utils::CLambdaOps2::AddEntriesInSparseSystem
// (include/slam/NonlinearSolver_Lambda_Base.h:1853-1931), the guided Schur ordering
// (src/slam/LinearSolver_Schur.cpp:771-838; cameras first, points after, id order kept) and the symbolic
// phase of the two block products (include/slam/BlockMatrixFBS.inl:684-834): the list of upper-triangular
// 6x6 blocks of the reduced camera system and, per block, the observation pairs that feed it, in
// ascending landmark order -- the order in which the reference accumulates them.

#include "spp_ctx.h"
#include <algorithm>
#include <numeric>

namespace spp {

// h_cam / h_pt: local camera / point index per observation in ORIGINAL edge order.
// Fills the track-ordered structure of s (device) and the permutation track position -> original edge.
void build_schur_structure(spp_ctx *ctx, size_t C, size_t P, const std::vector<uint32_t> &h_cam,
	const std::vector<uint32_t> &h_pt, std::vector<uint32_t> &obs_orig, std::vector<uint32_t> &t_cam,
	std::vector<uint32_t> &t_pt)
{
	SchurSystem &s = ctx->sys;
	const size_t O = h_cam.size();
	if(O >= 0xffffffffu || C >= 0xffffffffu || P >= 0xffffffffu)
		throw invalid_error("graph too large for 32-bit indices");
	s.C = C; s.P = P; s.O = O;

	// stable counting sort of the observations by point -> tracks
	std::vector<uint32_t> pt_ptr(P + 1, 0);
	for(size_t e = 0; e < O; ++ e)
		++ pt_ptr[h_pt[e] + 1];
	for(size_t p = 0; p < P; ++ p)
		pt_ptr[p + 1] += pt_ptr[p];
	obs_orig.resize(O);
	std::vector<uint32_t> pos_of_edge(O);
	{
		std::vector<uint32_t> fill(pt_ptr.begin(), pt_ptr.end() - 1);
		for(size_t e = 0; e < O; ++ e) {
			uint32_t pos = fill[h_pt[e]] ++;
			obs_orig[pos] = (uint32_t)e;
			pos_of_edge[e] = pos;
		}
	}
	t_cam.resize(O);
	t_pt.resize(O);
	for(size_t k = 0; k < O; ++ k) {
		t_cam[k] = h_cam[obs_orig[k]];
		t_pt[k] = h_pt[obs_orig[k]];
	}

	// per camera: its observations (track positions), in edge insertion order
	std::vector<uint32_t> cam_ptr(C + 1, 0), cam_obs(O);
	for(size_t e = 0; e < O; ++ e)
		++ cam_ptr[h_cam[e] + 1];
	for(size_t c = 0; c < C; ++ c)
		cam_ptr[c + 1] += cam_ptr[c];
	{
		std::vector<uint32_t> fill(cam_ptr.begin(), cam_ptr.end() - 1);
		for(size_t e = 0; e < O; ++ e)
			cam_obs[fill[h_cam[e]] ++] = pos_of_edge[e];
	}

	// reduced camera system: the upper blocks (i <= j) and the observation pairs of each.
	// Block list layout: blocks 0 .. C-1 are the diagonal blocks (i, i) -- every camera owns one, with or without
	// observations; their pair lists are the camera's own observations (a, a). The off-diagonal blocks follow,
	// in 8 x 8 camera tiles when C^2 is affordable (the warps of a CTA then work on one block row of a tile and
	// share the Y / W lines of its cameras through L1 / L2), in row-major key order otherwise.
	std::vector<uint32_t> blk_row, blk_col;
	std::vector<uint64_t> blk_ptr;
	std::vector<uint32_t> pair_a, pair_b;
	size_t n_pairs = 0;
	for(size_t p = 0; p < P; ++ p) {
		size_t k = pt_ptr[p + 1] - pt_ptr[p];
		n_pairs += k * (k + 1) / 2;
	}
	const bool b_dense_table = C * C <= (size_t(1) << 28);
	std::vector<uint64_t> keys; // sorted unique off-diagonal keys (sparse mode)
	std::vector<uint32_t> table; // key -> block index + 1 (dense mode)
	auto key_of = [C](uint32_t i, uint32_t j) { return (uint64_t)i * C + j; };
	blk_ptr.push_back(0);
	for(size_t i = 0; i < C; ++ i) {
		blk_row.push_back((uint32_t)i);
		blk_col.push_back((uint32_t)i);
		blk_ptr.push_back(blk_ptr.back() + (cam_ptr[i + 1] - cam_ptr[i]));
	}
	const size_t SCHUR_TILE = 8;
	if(b_dense_table) {
		std::vector<uint32_t> count(C * C, 0);
		for(size_t p = 0; p < P; ++ p) {
			for(uint32_t a = pt_ptr[p]; a < pt_ptr[p + 1]; ++ a) {
				for(uint32_t b = pt_ptr[p]; b < pt_ptr[p + 1]; ++ b) {
					uint32_t ca = t_cam[a], cb = t_cam[b];
					if(ca < cb)
						++ count[key_of(ca, cb)];
					else if(ca == cb && a != b)
						throw invalid_error("a landmark is observed twice by the same camera (duplicate edge)");
				}
			}
		}
		table.assign(C * C, 0);
		for(size_t ti = 0; ti < C; ti += SCHUR_TILE) {
			for(size_t tj = ti; tj < C; tj += SCHUR_TILE) {
				for(size_t i = ti; i < std::min(C, ti + SCHUR_TILE); ++ i) {
					for(size_t j = std::max(tj, i + 1); j < std::min(C, tj + SCHUR_TILE); ++ j) {
						uint32_t n = count[i * C + j];
						if(n) {
							blk_row.push_back((uint32_t)i);
							blk_col.push_back((uint32_t)j);
							table[i * C + j] = (uint32_t)blk_row.size();
							blk_ptr.push_back(blk_ptr.back() + n);
						}
					}
				}
			}
		}
	} else {
		keys.reserve(n_pairs - O);
		for(size_t p = 0; p < P; ++ p) {
			for(uint32_t a = pt_ptr[p]; a < pt_ptr[p + 1]; ++ a) {
				for(uint32_t b = pt_ptr[p]; b < pt_ptr[p + 1]; ++ b) {
					uint32_t ca = t_cam[a], cb = t_cam[b];
					if(ca < cb)
						keys.push_back(key_of(ca, cb));
					else if(ca == cb && a != b)
						throw invalid_error("a landmark is observed twice by the same camera (duplicate edge)");
				}
			}
		}
		std::sort(keys.begin(), keys.end());
		std::vector<uint64_t> uniq;
		for(size_t i = 0; i < keys.size();) {
			size_t j = i;
			while(j < keys.size() && keys[j] == keys[i]) ++ j;
			uniq.push_back(keys[i]);
			blk_row.push_back((uint32_t)(keys[i] / C));
			blk_col.push_back((uint32_t)(keys[i] % C));
			blk_ptr.push_back(blk_ptr.back() + (j - i));
			i = j;
		}
		keys.swap(uniq);
	}
	const size_t n_blk = blk_row.size();
	if(blk_ptr.back() != n_pairs)
		throw invalid_error("internal error: pair count mismatch in the Schur symbolic phase");
	// pass 2: scatter the pairs; iterating landmarks in ascending order keeps every list landmark-sorted
	pair_a.resize(n_pairs);
	pair_b.resize(n_pairs);
	{
		std::vector<uint64_t> fill(blk_ptr.begin(), blk_ptr.end() - 1);
		for(size_t p = 0; p < P; ++ p) {
			for(uint32_t a = pt_ptr[p]; a < pt_ptr[p + 1]; ++ a) {
				for(uint32_t b = pt_ptr[p]; b < pt_ptr[p + 1]; ++ b) {
					uint32_t ca = t_cam[a], cb = t_cam[b];
					if(!(ca < cb || a == b))
						continue;
					size_t blk;
					if(a == b)
						blk = ca;
					else if(b_dense_table)
						blk = table[key_of(ca, cb)] - 1;
					else
						blk = C + (std::lower_bound(keys.begin(), keys.end(), key_of(ca, cb)) - keys.begin());
					uint64_t dst = fill[blk] ++;
					pair_a[dst] = a;
					pair_b[dst] = b;
				}
			}
		}
	}

	cudaStream_t st = ctx->stream;
	s.obs_cam.upload(t_cam, st);
	s.obs_pt.upload(t_pt, st);
	s.pt_ptr.upload(pt_ptr, st);
	s.cam_ptr.upload(cam_ptr, st);
	s.cam_obs.upload(cam_obs, st);
	s.n_blocks = n_blk;
	s.n_pairs = n_pairs;
	s.blk_row.upload(blk_row, st);
	s.blk_col.upload(blk_col, st);
	s.blk_ptr.upload(blk_ptr, st);
	s.pair_a.upload(pair_a, st);
	s.pair_b.upload(pair_b, st);
	s.h_blk_row.swap(blk_row);
	s.h_blk_col.swap(blk_col);
	s.U.resize(C * 36); s.V.resize(P * 9); s.W.resize(O * 18);
	s.gc.resize(C * 6); s.gp.resize(P * 3);
	s.dxc.resize(C * 6); s.dxp.resize(P * 3);
	SPP_CUDA(cudaStreamSynchronize(st)); // the host vectors go out of scope
}

// Several ranks + block-sparse reduced camera system: the upper block list of the WHOLE graph (every rank sees all
// observations in spp_ba_set_graph and computes the same list): diagonal blocks 0 .. C-1 first, then the off-diagonal
// blocks in row-major order. h_cam / h_pt: local camera / point index of every observation of the whole graph.
void build_global_rcs_pattern(size_t C, size_t P, const std::vector<uint32_t> &h_cam, const std::vector<uint32_t> &h_pt,
	std::vector<uint32_t> &g_row, std::vector<uint32_t> &g_col)
{
	const size_t O = h_cam.size();
	std::vector<uint32_t> pt_ptr(P + 1, 0), t_cam(O);
	for(size_t e = 0; e < O; ++ e) ++ pt_ptr[h_pt[e] + 1];
	for(size_t p = 0; p < P; ++ p) pt_ptr[p + 1] += pt_ptr[p];
	{
		std::vector<uint32_t> fill(pt_ptr.begin(), pt_ptr.end() - 1);
		for(size_t e = 0; e < O; ++ e) t_cam[fill[h_pt[e]] ++] = h_cam[e];
	}
	std::vector<uint64_t> bits((C * C + 63) / 64, 0);
	for(size_t p = 0; p < P; ++ p) {
		for(uint32_t a = pt_ptr[p]; a < pt_ptr[p + 1]; ++ a) {
			for(uint32_t b = a + 1; b < pt_ptr[p + 1]; ++ b) {
				const uint32_t ca = std::min(t_cam[a], t_cam[b]), cb = std::max(t_cam[a], t_cam[b]);
				const uint64_t k = (uint64_t)ca * C + cb;
				bits[k >> 6] |= (uint64_t)1 << (k & 63);
			}
		}
	}
	g_row.clear();
	g_col.clear();
	for(size_t i = 0; i < C; ++ i) { g_row.push_back((uint32_t)i); g_col.push_back((uint32_t)i); }
	for(size_t i = 0; i < C; ++ i) {
		for(size_t j = i + 1; j < C; ++ j) {
			const uint64_t k = (uint64_t)i * C + j;
			if(!(k & 63) && j + 64 <= C && !bits[k >> 6]) { j += 63; continue; } // skip an empty word
			if(bits[k >> 6] >> (k & 63) & 1) { g_row.push_back((uint32_t)i); g_col.push_back((uint32_t)j); }
		}
	}
}

// position of every block of this rank's list in the global list built above
void map_blocks_to_global(size_t C, const std::vector<uint32_t> &l_row, const std::vector<uint32_t> &l_col,
	const std::vector<uint32_t> &g_row, const std::vector<uint32_t> &g_col, std::vector<uint32_t> &slot)
{
	std::vector<uint64_t> row_start(C + 1, 0); // off-diagonal blocks of the global list, by row
	for(size_t b = C; b < g_row.size(); ++ b) ++ row_start[g_row[b] + 1];
	for(size_t i = 0; i < C; ++ i) row_start[i + 1] += row_start[i];
	slot.resize(l_row.size());
	for(size_t b = 0; b < l_row.size(); ++ b) {
		const uint32_t i = l_row[b], j = l_col[b];
		if(i == j) { slot[b] = i; continue; }
		const uint32_t *beg = &g_col[C + row_start[i]], *end = &g_col[0] + C + row_start[i + 1];
		const uint32_t *it = std::lower_bound(beg, end, j);
		if(it == end || *it != j)
			throw invalid_error("internal error: a block of this rank is missing from the global block list");
		slot[b] = (uint32_t)(it - &g_col[0]);
	}
}

} // namespace spp

extern "C" int spp_rcs_block_pattern(size_t n_cameras, size_t n_points, size_t n_observations, const uint32_t *p_obs_camera,
	const uint32_t *p_obs_point, uint64_t *p_n_blocks, uint32_t *p_block_row, uint32_t *p_block_col)
{
	if(!p_n_blocks || (n_observations && (!p_obs_camera || !p_obs_point)))
		return SPP_ERR_INVALID;
	try {
		for(size_t e = 0; e < n_observations; ++ e)
			if(p_obs_camera[e] >= n_cameras || p_obs_point[e] >= n_points) return SPP_ERR_INVALID;
		std::vector<uint32_t> h_cam(p_obs_camera, p_obs_camera + n_observations), h_pt(p_obs_point, p_obs_point + n_observations), r, c;
		spp::build_global_rcs_pattern(n_cameras, n_points, h_cam, h_pt, r, c);
		if(p_block_row && p_block_col) {
			if(*p_n_blocks < r.size()) return SPP_ERR_INVALID;
			std::copy(r.begin(), r.end(), p_block_row);
			std::copy(c.begin(), c.end(), p_block_col);
		}
		*p_n_blocks = r.size();
	} catch(const std::bad_alloc&) {
		return SPP_ERR_NOMEM;
	}
	return SPP_OK;
}

extern "C" int spp_partition_landmarks(size_t n_points, const uint32_t *p_track_length, int world, uint64_t *p_bounds)
{
	if(world < 1 || !p_bounds || (n_points && !p_track_length))
		return SPP_ERR_INVALID;
	long double total = 0;
	for(size_t p = 0; p < n_points; ++ p) {
		const long double k = p_track_length[p];
		total += k * (k + 1) / 2 + k;
	}
	p_bounds[0] = 0;
	long double acc = 0;
	size_t p = 0;
	for(int r = 1; r < world; ++ r) {
		const long double target = total * r / world;
		while(p < n_points) {
			const long double k = p_track_length[p];
			const long double w = k * (k + 1) / 2 + k;
			if(acc + w / 2 > target)
				break;
			acc += w;
			++ p;
		}
		p_bounds[r] = p;
	}
	p_bounds[world] = n_points;
	return SPP_OK;
}

