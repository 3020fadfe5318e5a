// symbolic_gpu.cu -- the symbolic analysis of the landmark Schur system (symbolic.cpp) done on the device, for the
// graph-upload path (spp_ba_set_graph): the end-to-end cost of one solve is otherwise dominated by this integer work
// on one host core.
//
// Same output as build_schur_structure() in symbolic.cpp, bit for bit (tests compare the two):
//   tracks         stable sort of the observations by landmark (edge insertion order kept inside a track)
//   camera lists   per camera, its observations in edge insertion order (summation order of the camera blocks,
//                  include/slam/NonlinearSolver_Lambda_Base.h:152-197, 563-607) and in landmark order (the pair
//                  list of the diagonal block of the reduced camera system)
//   block list     diagonal blocks 0 .. C-1, then the non-empty off-diagonal blocks in 8 x 8 camera tiles
//   pair lists     per block, the observation pairs (a, b) of one landmark with cam(a) < cam(b), in ascending
//                  landmark order (the accumulation order of include/slam/BlockMatrixFBS.h:395-448)
// Sorting and scanning use CUB (radix sort is stable, which is what keeps every list in landmark order); the
// enumeration kernels are below. Integer work only; nothing here runs per LM iteration.

#include <chrono>
#include <string>
#include "spp_ctx.h"
#include <algorithm>
#include <cub/cub.cuh>

namespace spp {

#define LAUNCH_CHECK(ctx) do { ++ (ctx)->n_launches; SPP_CUDA(cudaGetLastError()); } while(0)
#define SG_TILE 8

__global__ void k_sg_iota(uint32_t *p, size_t n)
{
	size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
	if(i < n) p[i] = (uint32_t)i;
}

// local camera / point index of every observation from the vertex ids; flags bad references
__global__ void k_sg_edge_local(size_t O, size_t n_vertices, const uint64_t *__restrict__ obs_pt_id,
	const uint64_t *__restrict__ obs_cam_id, const uint8_t *__restrict__ vtype, const uint32_t *__restrict__ vlocal,
	uint32_t *__restrict__ ocam, uint32_t *__restrict__ opt, int *__restrict__ err)
{
	size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
	if(e >= O) return;
	const uint64_t vp = obs_pt_id[e], vc = obs_cam_id[e];
	if(vp >= n_vertices || vc >= n_vertices || vtype[vp] != 1 || vtype[vc] != 0) {
		*err = 1;
		ocam[e] = opt[e] = 0;
		return;
	}
	ocam[e] = vlocal[vc];
	opt[e] = vlocal[vp];
}

__global__ void k_sg_count(size_t n, const uint32_t *__restrict__ key, uint32_t *__restrict__ cnt)
{
	size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
	if(i < n) atomicAdd(&cnt[key[i]], 1u);
}

// t_cam[k] = ocam[obs_orig[k]], pos_of_edge[obs_orig[k]] = k
__global__ void k_sg_track_gather(size_t O, const uint32_t *__restrict__ obs_orig, const uint32_t *__restrict__ ocam,
	uint32_t *__restrict__ t_cam, uint32_t *__restrict__ pos_of_edge)
{
	size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
	if(k >= O) return;
	const uint32_t e = obs_orig[k];
	t_cam[k] = ocam[e];
	pos_of_edge[e] = (uint32_t)k;
}

__global__ void k_sg_gather_u32(size_t n, const uint32_t *__restrict__ idx, const uint32_t *__restrict__ src, uint32_t *__restrict__ dst)
{
	size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
	if(i < n) dst[i] = src[idx[i]];
}

// number of off-diagonal pairs of every landmark: k (k - 1) / 2
__global__ void k_sg_pair_count(size_t P, const uint32_t *__restrict__ pt_ptr, uint64_t *__restrict__ n_off)
{
	size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
	if(p >= P) return;
	const uint64_t k = pt_ptr[p + 1] - pt_ptr[p];
	n_off[p] = k * (k - (k > 0)) / 2;
}

// position of the block (i, j), i < j, in the tile-major enumeration space (n_t x n_t tiles of 8 x 8)
__device__ __forceinline__ size_t sg_lin(uint32_t i, uint32_t j, size_t n_t)
{
	return (((size_t)(i / SG_TILE) * n_t + j / SG_TILE) * SG_TILE + i % SG_TILE) * SG_TILE + j % SG_TILE;
}

// thread per landmark: the pairs (a, b), cam(a) < cam(b), in the order of the host code (a outer, b inner over the
// track); pass 0 counts them per block, pass 1 emits (block id, a, b) at the landmark's offset
template <int PASS>
__global__ void k_sg_pairs(size_t P, size_t n_t, const uint32_t *__restrict__ pt_ptr, const uint32_t *__restrict__ t_cam,
	uint32_t *__restrict__ cnt_lin, const uint32_t *__restrict__ blk_of_lin, const uint64_t *__restrict__ pair_off,
	uint32_t *__restrict__ key_out, uint64_t *__restrict__ val_out, int *__restrict__ err)
{
	size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
	if(p >= P) return;
	const uint32_t beg = pt_ptr[p], end = pt_ptr[p + 1];
	uint64_t w = PASS? pair_off[p] : 0;
	for(uint32_t a = beg; a < end; ++ a) {
		const uint32_t ca = t_cam[a];
		for(uint32_t b = beg; b < end; ++ b) {
			const uint32_t cb = t_cam[b];
			if(ca < cb) {
				const size_t lin = sg_lin(ca, cb, n_t);
				if(PASS == 0)
					atomicAdd(&cnt_lin[lin], 1u);
				else {
					key_out[w] = blk_of_lin[lin];
					val_out[w] = (uint64_t)a | ((uint64_t)b << 32);
					++ w;
				}
			} else if(PASS == 0 && ca == cb && a != b)
				*err = 2; // a landmark observed twice by the same camera
		}
	}
}

// flags of the non-empty enumeration slots and their pair counts as 64-bit values for the scans
__global__ void k_sg_flags(size_t L, const uint32_t *__restrict__ cnt_lin, uint32_t *__restrict__ flag, uint64_t *__restrict__ cnt64)
{
	size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
	if(i >= L) return;
	const uint32_t c = cnt_lin[i];
	flag[i] = c != 0;
	cnt64[i] = c;
}

// block list entries of the non-empty slots: id = C + rank; blk_ptr[id] = O + pairs before it
__global__ void k_sg_blocks(size_t L, size_t n_t, size_t C, size_t O, const uint32_t *__restrict__ cnt_lin,
	const uint32_t *__restrict__ rank, const uint64_t *__restrict__ before, uint32_t *__restrict__ blk_of_lin,
	uint32_t *__restrict__ blk_row, uint32_t *__restrict__ blk_col, uint64_t *__restrict__ blk_ptr)
{
	size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
	if(i >= L || !cnt_lin[i]) return;
	const size_t id = C + rank[i];
	const size_t tile = i / (SG_TILE * SG_TILE), in = i % (SG_TILE * SG_TILE);
	blk_of_lin[i] = (uint32_t)id;
	blk_row[id] = (uint32_t)((tile / n_t) * SG_TILE + in / SG_TILE);
	blk_col[id] = (uint32_t)((tile % n_t) * SG_TILE + in % SG_TILE);
	blk_ptr[id] = O + before[i];
}

// diagonal block entries: (i, i), pairs = the camera's observations
__global__ void k_sg_diag_blocks(size_t C, const uint32_t *__restrict__ cam_ptr, uint32_t *__restrict__ blk_row,
	uint32_t *__restrict__ blk_col, uint64_t *__restrict__ blk_ptr)
{
	size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
	if(i >= C) return;
	blk_row[i] = blk_col[i] = (uint32_t)i;
	blk_ptr[i] = cam_ptr[i];
}

__global__ void k_sg_split_pairs(size_t n, const uint64_t *__restrict__ val, uint32_t *__restrict__ pa, uint32_t *__restrict__ pb)
{
	size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
	if(i >= n) return;
	pa[i] = (uint32_t)val[i];
	pb[i] = (uint32_t)(val[i] >> 32);
}

__global__ void k_sg_set_u64(uint64_t *p, uint64_t v) { *p = v; }

// dst row k = src row idx[k], NC doubles per row (measurements into track order)
template <int NC>
__global__ void k_sg_gather_rows(size_t n, const uint32_t *__restrict__ idx, const double *__restrict__ src, double *__restrict__ dst)
{
	size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
	if(i >= n * NC) return;
	const size_t k = i / NC, c = i % NC;
	dst[i] = src[(size_t)idx[k] * NC + c];
}

static int bits_for(size_t n) // number of key bits needed for values < n
{
	int b = 1;
	while(b < 32 && (size_t(1) << b) < n) ++ b;
	return b;
}

struct SgTemp { // grows-only scratch of the CUB calls
	DBuf<uint8_t> &buf;
	explicit SgTemp(DBuf<uint8_t> &b) : buf(b) {}
	void *get(size_t n) { if(buf.size() < n) buf.resize(n); return buf.p(); }
};

template <class K, class V>
static void sort_pairs(spp_ctx *ctx, SgTemp &tmp, const K *k_in, K *k_out, const V *v_in, V *v_out, size_t n, int end_bit)
{
	if(!n) return;
	size_t bytes = 0;
	SPP_CUDA(cub::DeviceRadixSort::SortPairs(0, bytes, k_in, k_out, v_in, v_out, n, 0, end_bit, ctx->stream));
	SPP_CUDA(cub::DeviceRadixSort::SortPairs(tmp.get(bytes), bytes, k_in, k_out, v_in, v_out, n, 0, end_bit, ctx->stream));
	ctx->n_launches += 3;
}

template <class T>
static void exclusive_scan(spp_ctx *ctx, SgTemp &tmp, const T *in, T *out, size_t n)
{
	if(!n) return;
	size_t bytes = 0;
	SPP_CUDA(cub::DeviceScan::ExclusiveSum(0, bytes, in, out, n, ctx->stream));
	SPP_CUDA(cub::DeviceScan::ExclusiveSum(tmp.get(bytes), bytes, in, out, n, ctx->stream));
	ctx->n_launches += 2;
}

bool schur_structure_device_supported(size_t C)
{
	const size_t n_t = (C + SG_TILE - 1) / SG_TILE;
	return n_t * n_t * SG_TILE * SG_TILE <= (size_t(1) << 30); // enumeration space (4 B + 4 B + 8 B + 8 B per slot)
}

// d_ocam / d_opt: local camera / point index per observation in edge insertion order (device).
// Fills the device structure of ctx->sys and d_obs_orig (track position -> edge index).
// SPP_SYMBOLIC_TIMING: wall-clock marks of the device-side analysis on stderr (each mark synchronises the stream)
struct SgMarks {
	bool on;
	cudaStream_t st;
	std::chrono::steady_clock::time_point t0;
	std::string line;
	explicit SgMarks(cudaStream_t s) : on(getenv("SPP_SYMBOLIC_TIMING") != 0), st(s), t0(std::chrono::steady_clock::now()) {}
	void mark(const char *what)
	{
		if(!on) return;
		cudaStreamSynchronize(st);
		const auto t1 = std::chrono::steady_clock::now();
		char buf[96];
		snprintf(buf, sizeof(buf), " %s %.2f", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
		line += buf;
		t0 = t1;
	}
	~SgMarks() { if(on && !line.empty()) fprintf(stderr, "[spp symbolic ms]%s\n", line.c_str()); }
};

void build_schur_structure_device(spp_ctx *ctx, size_t C, size_t P, size_t O, const uint32_t *d_ocam, const uint32_t *d_opt,
	DBuf<uint32_t> &d_obs_orig)
{
	SgMarks marks(ctx->stream);
	SchurSystem &s = ctx->sys;
	SymbolicScratch &w = ctx->sym;
	cudaStream_t st = ctx->stream;
	if(O >= 0xffffffffu || C >= 0xffffffffu || P >= 0xffffffffu)
		throw invalid_error("graph too large for 32-bit indices");
	s.C = C; s.P = P; s.O = O;
	s.h_blk_row.clear(); s.h_blk_col.clear();
	SgTemp tmp(w.cub_temp);
	const unsigned T = 256;
	w.err.resize(1);
	SPP_CUDA(cudaMemsetAsync(w.err.p(), 0, sizeof(int), st));

	// ---- tracks: stable sort of the edges by landmark
	w.iota.resize(O);
	s.obs_pt.resize(O); s.obs_cam.resize(O);
	d_obs_orig.resize(O);
	w.pos_of_edge.resize(O);
	if(O) {
		k_sg_iota<<<n_blocks(O, T), T, 0, st>>>(w.iota.p(), O);
		LAUNCH_CHECK(ctx);
	}
	sort_pairs(ctx, tmp, d_opt, s.obs_pt.p(), w.iota.p(), d_obs_orig.p(), O, bits_for(P));
	if(O) {
		k_sg_track_gather<<<n_blocks(O, T), T, 0, st>>>(O, d_obs_orig.p(), d_ocam, s.obs_cam.p(), w.pos_of_edge.p());
		LAUNCH_CHECK(ctx);
	}
	marks.mark("tracks");
	// pt_ptr / cam_ptr: histogram + exclusive scan
	w.cnt.resize(std::max(P, C) + 1);
	s.pt_ptr.resize(P + 1);
	s.cam_ptr.resize(C + 1);
	SPP_CUDA(cudaMemsetAsync(w.cnt.p(), 0, (P + 1) * sizeof(uint32_t), st));
	if(O) {
		k_sg_count<<<n_blocks(O, T), T, 0, st>>>(O, d_opt, w.cnt.p());
		LAUNCH_CHECK(ctx);
	}
	exclusive_scan(ctx, tmp, w.cnt.p(), s.pt_ptr.p(), P + 1);
	SPP_CUDA(cudaMemsetAsync(w.cnt.p(), 0, (C + 1) * sizeof(uint32_t), st));
	if(O) {
		k_sg_count<<<n_blocks(O, T), T, 0, st>>>(O, d_ocam, w.cnt.p());
		LAUNCH_CHECK(ctx);
	}
	exclusive_scan(ctx, tmp, w.cnt.p(), s.cam_ptr.p(), C + 1);

	// ---- camera lists: edge insertion order (values = track positions of the edges) ...
	s.cam_obs.resize(O);
	w.keys_out.resize(O);
	sort_pairs(ctx, tmp, d_ocam, w.keys_out.p(), w.pos_of_edge.p(), s.cam_obs.p(), O, bits_for(C));

	marks.mark("ptr+camlists");
	// ---- off-diagonal pairs: count per enumeration slot
	const size_t n_t = (C + SG_TILE - 1) / SG_TILE, L = n_t * n_t * SG_TILE * SG_TILE;
	w.cnt_lin.resize(L);
	w.flag.resize(L + 1);
	w.rank.resize(L + 1);
	w.cnt64.resize(L + 1);
	w.before.resize(L + 1);
	w.blk_of_lin.resize(L);
	SPP_CUDA(cudaMemsetAsync(w.cnt_lin.p(), 0, L * sizeof(uint32_t), st));
	SPP_CUDA(cudaMemsetAsync(w.flag.p() + L, 0, sizeof(uint32_t), st));
	SPP_CUDA(cudaMemsetAsync(w.cnt64.p() + L, 0, sizeof(uint64_t), st));
	if(P) {
		k_sg_pairs<0><<<n_blocks(P, 128), 128, 0, st>>>(P, n_t, s.pt_ptr.p(), s.obs_cam.p(), w.cnt_lin.p(), 0, 0, 0, 0, w.err.p());
		LAUNCH_CHECK(ctx);
	}
	if(L) {
		k_sg_flags<<<n_blocks(L, T), T, 0, st>>>(L, w.cnt_lin.p(), w.flag.p(), w.cnt64.p());
		LAUNCH_CHECK(ctx);
	}
	exclusive_scan(ctx, tmp, w.flag.p(), w.rank.p(), L + 1);     // rank[L] = number of off-diagonal blocks
	exclusive_scan(ctx, tmp, w.cnt64.p(), w.before.p(), L + 1);  // before[L] = number of off-diagonal pairs
	// the two totals decide the allocation sizes: one small read-back
	ctx->h_scalars.resize(16);
	uint32_t *h_nblk = reinterpret_cast<uint32_t*>(ctx->h_scalars.p());
	uint64_t *h_noff = reinterpret_cast<uint64_t*>(ctx->h_scalars.p() + 1);
	int *h_err = reinterpret_cast<int*>(ctx->h_scalars.p() + 2);
	SPP_CUDA(cudaMemcpyAsync(h_nblk, w.rank.p() + L, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
	SPP_CUDA(cudaMemcpyAsync(h_noff, w.before.p() + L, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
	SPP_CUDA(cudaMemcpyAsync(h_err, w.err.p(), sizeof(int), cudaMemcpyDeviceToHost, st));
	SPP_CUDA(cudaStreamSynchronize(st));
	if(*h_err == 1)
		throw invalid_error("observation references a vertex of the wrong type or out of range");
	if(*h_err == 2)
		throw invalid_error("a landmark is observed twice by the same camera (duplicate edge)");
	const size_t n_off_blk = *h_nblk, n_off = *h_noff, n_blk = C + n_off_blk, n_pairs = O + n_off;
	marks.mark("count");

	// ---- block list
	s.blk_row.resize(n_blk); s.blk_col.resize(n_blk); s.blk_ptr.resize(n_blk + 1);
	if(C) {
		k_sg_diag_blocks<<<n_blocks(C, T), T, 0, st>>>(C, s.cam_ptr.p(), s.blk_row.p(), s.blk_col.p(), s.blk_ptr.p());
		LAUNCH_CHECK(ctx);
	}
	if(L) {
		k_sg_blocks<<<n_blocks(L, T), T, 0, st>>>(L, n_t, C, O, w.cnt_lin.p(), w.rank.p(), w.before.p(), w.blk_of_lin.p(),
			s.blk_row.p(), s.blk_col.p(), s.blk_ptr.p());
		LAUNCH_CHECK(ctx);
	}
	k_sg_set_u64<<<1, 1, 0, st>>>(s.blk_ptr.p() + n_blk, n_pairs);
	LAUNCH_CHECK(ctx);

	marks.mark("blocks");
	// ---- pair lists: diagonal blocks = the camera's observations in landmark order (ascending track position) ...
	s.pair_a.resize(n_pairs); s.pair_b.resize(n_pairs);
	sort_pairs(ctx, tmp, s.obs_cam.p(), w.keys_out.p(), w.iota.p(), s.pair_a.p(), O, bits_for(C));
	if(O)
		SPP_CUDA(cudaMemcpyAsync(s.pair_b.p(), s.pair_a.p(), O * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
	// ... off-diagonal blocks: emit (block id, a, b) landmark by landmark, stable sort by block id
	if(n_off) {
		w.n_off.resize(P + 1);
		w.pair_off.resize(P + 1);
		SPP_CUDA(cudaMemsetAsync(w.n_off.p() + P, 0, sizeof(uint64_t), st));
		k_sg_pair_count<<<n_blocks(P, T), T, 0, st>>>(P, s.pt_ptr.p(), w.n_off.p());
		LAUNCH_CHECK(ctx);
		exclusive_scan(ctx, tmp, w.n_off.p(), w.pair_off.p(), P + 1);
		w.pkey.resize(n_off); w.pkey_out.resize(n_off);
		w.pval.resize(n_off); w.pval_out.resize(n_off);
		k_sg_pairs<1><<<n_blocks(P, 128), 128, 0, st>>>(P, n_t, s.pt_ptr.p(), s.obs_cam.p(), 0, w.blk_of_lin.p(), w.pair_off.p(),
			w.pkey.p(), w.pval.p(), w.err.p());
		LAUNCH_CHECK(ctx);
		sort_pairs(ctx, tmp, w.pkey.p(), w.pkey_out.p(), w.pval.p(), w.pval_out.p(), n_off, bits_for(n_blk));
		k_sg_split_pairs<<<n_blocks(n_off, T), T, 0, st>>>(n_off, w.pval_out.p(), s.pair_a.p() + O, s.pair_b.p() + O);
		LAUNCH_CHECK(ctx);
	}
	marks.mark("pairs");
	s.n_blocks = n_blk;
	s.n_pairs = n_pairs;
	s.U.resize(C * 36); s.V.resize(P * 9); s.W.resize(O * 18);
	s.gc.resize(C * 6); s.gp.resize(P * 3);
	s.dxc.resize(C * 6); s.dxp.resize(P * 3);
	marks.mark("lambda buffers");
}

// copies z / info to the staging buffers on the context's copy stream; records ctx->copy_done
static void upload_measurements_async(spp_ctx *ctx, const double *p_z, const double *p_info, size_t O)
{
	SymbolicScratch &w = ctx->sym;
	if(!ctx->copy_stream) {
		SPP_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
		SPP_CUDA(cudaEventCreateWithFlags(&ctx->copy_done, cudaEventDisableTiming));
	}
	w.z_in.resize(O * 2);     // (re)allocation happens here, on the host thread, before anything is enqueued
	w.info_in.resize(O * 4);
	// whatever still reads the staging buffers of the previous graph is ordered on the main stream
	SPP_CUDA(cudaEventRecord(ctx->copy_done, ctx->stream));
	SPP_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_done, 0));
	if(O) {
		SPP_CUDA(cudaMemcpyAsync(w.z_in.p(), p_z, O * 2 * sizeof(double), cudaMemcpyHostToDevice, ctx->copy_stream));
		SPP_CUDA(cudaMemcpyAsync(w.info_in.p(), p_info, O * 4 * sizeof(double), cudaMemcpyHostToDevice, ctx->copy_stream));
	}
	SPP_CUDA(cudaEventRecord(ctx->copy_done, ctx->copy_stream));
}

void ba_analyse_device_resident(spp_ctx *ctx, size_t C, size_t P, size_t O);

// uploads the caller's observation arrays, derives the per-edge local indices, builds the structure and gathers the
// measurements into track order -- the device-side body of spp_ba_set_graph for a single-GPU context
void ba_upload_and_analyse_device(spp_ctx *ctx, size_t C, size_t P, size_t O, const uint64_t *p_obs_point,
	const uint64_t *p_obs_camera, const double *p_z, const double *p_info)
{
	BAProblem &ba = ctx->ba;
	SymbolicScratch &w = ctx->sym;
	cudaStream_t st = ctx->stream;
	const unsigned T = 256;
	w.vtype.upload(ba.vtype.data(), ba.n_vertices, st);
	w.vlocal.upload(ba.vertex_local.data(), ba.n_vertices, st);
	w.obs_pt_id.upload(p_obs_point, O, st);
	w.obs_cam_id.upload(p_obs_camera, O, st);
	// the measurements (two thirds of the bytes) are only needed when the structure is known: their copy runs on a
	// second stream beside the analysis kernels (truly asynchronous when the caller's buffers are pinned)
	upload_measurements_async(ctx, p_z, p_info, O);
	ba_analyse_device_resident(ctx, C, P, O);
}

// the part of the analysis that only needs what is already on the device: per-edge local indices from the staged vertex
// ids, the structure, the measurements gathered into track order (behind the measurement copy of the side stream)
void ba_analyse_device_resident(spp_ctx *ctx, size_t C, size_t P, size_t O)
{
	BAProblem &ba = ctx->ba;
	SymbolicScratch &w = ctx->sym;
	cudaStream_t st = ctx->stream;
	const unsigned T = 256;
	w.ocam.resize(O); w.opt.resize(O);
	w.err.resize(1);
	SPP_CUDA(cudaMemsetAsync(w.err.p(), 0, sizeof(int), st));
	if(O) {
		k_sg_edge_local<<<n_blocks(O, T), T, 0, st>>>(O, ba.n_vertices, w.obs_pt_id.p(), w.obs_cam_id.p(), w.vtype.p(),
			w.vlocal.p(), w.ocam.p(), w.opt.p(), w.err.p());
		LAUNCH_CHECK(ctx);
		// a bad reference must be reported before the structure is built on clamped indices
		int h_err = 0;
		SPP_CUDA(cudaMemcpyAsync(&h_err, w.err.p(), sizeof(int), cudaMemcpyDeviceToHost, st));
		SPP_CUDA(cudaStreamSynchronize(st));
		if(h_err)
			throw invalid_error("observation references a vertex of the wrong type or out of range");
	}
	build_schur_structure_device(ctx, C, P, O, w.ocam.p(), w.opt.p(), ba.d_obs_orig);
	ba.z.resize(O * 2);
	ba.info.resize(O * 4);
	SPP_CUDA(cudaStreamWaitEvent(st, ctx->copy_done, 0)); // the measurements have arrived
	if(O) {
		k_sg_gather_rows<2><<<n_blocks(O * 2, T), T, 0, st>>>(O, ba.d_obs_orig.p(), w.z_in.p(), ba.z.p());
		LAUNCH_CHECK(ctx);
		k_sg_gather_rows<4><<<n_blocks(O * 4, T), T, 0, st>>>(O, ba.d_obs_orig.p(), w.info_in.p(), ba.info.p());
		LAUNCH_CHECK(ctx);
	}
	ba.host_maps_valid = false;
}

// Appends to the staged graph (spp_ba_append_graph): the vertex tables are sent again (a few bytes per vertex), the new
// observations' ids and measurements go behind the ones already staged in edge insertion order, and the analysis runs
// on the whole graph from device memory -- nothing that is already there crosses the link again
void ba_append_and_analyse_device(spp_ctx *ctx, size_t C, size_t P, size_t O_old, size_t O_new, const uint64_t *p_obs_point,
	const uint64_t *p_obs_camera, const double *p_z, const double *p_info)
{
	BAProblem &ba = ctx->ba;
	SymbolicScratch &w = ctx->sym;
	cudaStream_t st = ctx->stream;
	if(w.obs_pt_id.size() != O_old || w.obs_cam_id.size() != O_old || w.z_in.size() != O_old * 2 || w.info_in.size() != O_old * 4)
		throw invalid_error("spp_ba_append_graph: the staged graph is gone (another graph was analysed in between)");
	const size_t O = O_old + O_new;
	w.vtype.upload(ba.vtype.data(), ba.n_vertices, st);
	w.vlocal.upload(ba.vertex_local.data(), ba.n_vertices, st);
	SPP_CUDA(cudaStreamWaitEvent(st, ctx->copy_done, 0)); // nothing of the previous graph's copy is still in flight
	w.obs_pt_id.grow_keep(O, st);
	w.obs_cam_id.grow_keep(O, st);
	w.z_in.grow_keep(O * 2, st);
	w.info_in.grow_keep(O * 4, st);
	if(O_new) {
		SPP_CUDA(cudaMemcpyAsync(w.obs_pt_id.p() + O_old, p_obs_point, O_new * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
		SPP_CUDA(cudaMemcpyAsync(w.obs_cam_id.p() + O_old, p_obs_camera, O_new * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
		SPP_CUDA(cudaMemcpyAsync(w.z_in.p() + O_old * 2, p_z, O_new * 2 * sizeof(double), cudaMemcpyHostToDevice, st));
		SPP_CUDA(cudaMemcpyAsync(w.info_in.p() + O_old * 4, p_info, O_new * 4 * sizeof(double), cudaMemcpyHostToDevice, st));
	}
	SPP_CUDA(cudaEventRecord(ctx->copy_done, st)); // what ba_analyse_device_resident waits for
	ba_analyse_device_resident(ctx, C, P, O);
}

// ---- several ranks: this rank keeps a contiguous slice of the landmarks ---------------------------------------------

__global__ void k_sg_slice_flags(size_t O, const uint32_t *__restrict__ opt, uint32_t pt_begin, uint32_t pt_end, uint32_t *__restrict__ flag)
{
	size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
	if(e < O) flag[e] = (opt[e] >= pt_begin && opt[e] < pt_end)? 1u : 0u;
}

// stable compaction of the edges of the slice: kept[pos] = edge, local camera / point index (point relative to the slice)
__global__ void k_sg_slice_compact(size_t O, const uint32_t *__restrict__ flag, const uint32_t *__restrict__ pos,
	const uint32_t *__restrict__ ocam, const uint32_t *__restrict__ opt, uint32_t pt_begin, uint32_t *__restrict__ kept,
	uint32_t *__restrict__ ocam_l, uint32_t *__restrict__ opt_l)
{
	size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
	if(e >= O || !flag[e]) return;
	const uint32_t k = pos[e];
	kept[k] = (uint32_t)e;
	ocam_l[k] = ocam[e];
	opt_l[k] = opt[e] - pt_begin;
}

__global__ void k_sg_compose(size_t n, uint32_t *__restrict__ map, const uint32_t *__restrict__ kept)
{
	size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
	if(i < n) map[i] = kept[map[i]];
}

// The multi-rank body of spp_ba_set_graph, on the device: (a) the structure of the WHOLE graph gives the global block
// list of the reduced camera system (identical on every rank) and the track lengths; (b) the landmark slices are cut
// (spp_partition_landmarks); (c) the edges of this rank's slice are compacted (edge insertion order kept) and the
// structure is built again for the slice; (d) every block of the slice is located in the global list.
void ba_upload_and_analyse_device_sliced(spp_ctx *ctx, size_t C, size_t P, size_t O, const uint64_t *p_obs_point,
	const uint64_t *p_obs_camera, const double *p_z, const double *p_info)
{
	BAProblem &ba = ctx->ba;
	SchurSystem &s = ctx->sys;
	SymbolicScratch &w = ctx->sym;
	cudaStream_t st = ctx->stream;
	const unsigned T = 256;
	w.vtype.upload(ba.vtype.data(), ba.n_vertices, st);
	w.vlocal.upload(ba.vertex_local.data(), ba.n_vertices, st);
	w.obs_pt_id.upload(p_obs_point, O, st);
	w.obs_cam_id.upload(p_obs_camera, O, st);
	upload_measurements_async(ctx, p_z, p_info, O); // beside the analysis, as in the single-rank path
	w.ocam.resize(O); w.opt.resize(O);
	w.err.resize(1);
	SPP_CUDA(cudaMemsetAsync(w.err.p(), 0, sizeof(int), st));
	if(O) {
		k_sg_edge_local<<<n_blocks(O, T), T, 0, st>>>(O, ba.n_vertices, w.obs_pt_id.p(), w.obs_cam_id.p(), w.vtype.p(),
			w.vlocal.p(), w.ocam.p(), w.opt.p(), w.err.p());
		LAUNCH_CHECK(ctx);
		int h_err = 0;
		SPP_CUDA(cudaMemcpyAsync(&h_err, w.err.p(), sizeof(int), cudaMemcpyDeviceToHost, st));
		SPP_CUDA(cudaStreamSynchronize(st));
		if(h_err)
			throw invalid_error("observation references a vertex of the wrong type or out of range");
	}
	// (a) the whole graph
	build_schur_structure_device(ctx, C, P, O, w.ocam.p(), w.opt.p(), ba.d_obs_orig);
	s.n_blocks_global = s.n_blocks;
	s.h_gblk_row.resize(s.n_blocks);
	s.h_gblk_col.resize(s.n_blocks);
	s.blk_row.download(s.h_gblk_row.data(), s.n_blocks, st);
	s.blk_col.download(s.h_gblk_col.data(), s.n_blocks, st);
	std::vector<uint32_t> pt_ptr(P + 1);
	s.pt_ptr.download(pt_ptr.data(), P + 1, st);
	SPP_CUDA(cudaStreamSynchronize(st));
	s.gblk_row.upload(s.h_gblk_row, st);
	s.gblk_col.upload(s.h_gblk_col, st);
	// (b) slices balanced by the Schur-product work
	std::vector<uint32_t> track_len(P);
	for(size_t p = 0; p < P; ++ p) track_len[p] = pt_ptr[p + 1] - pt_ptr[p];
	std::vector<uint64_t> bounds(ctx->world + 1);
	spp_partition_landmarks(P, P? &track_len[0] : 0, ctx->world, &bounds[0]);
	ba.pt_begin = bounds[ctx->rank];
	ba.pt_end = bounds[ctx->rank + 1];
	const size_t P_local = ba.pt_end - ba.pt_begin;
	const size_t O_local = pt_ptr[ba.pt_end] - pt_ptr[ba.pt_begin];
	// (c) this rank's edges
	SgTemp tmp(w.cub_temp);
	DBuf<uint32_t> &flag = w.sl_flag, &pos = w.sl_pos, &kept = w.sl_kept, &ocam_l = w.sl_ocam, &opt_l = w.sl_opt;
	flag.resize(O); pos.resize(O); kept.resize(O_local); ocam_l.resize(O_local); opt_l.resize(O_local);
	if(O) {
		k_sg_slice_flags<<<n_blocks(O, T), T, 0, st>>>(O, w.opt.p(), (uint32_t)ba.pt_begin, (uint32_t)ba.pt_end, flag.p());
		LAUNCH_CHECK(ctx);
		exclusive_scan(ctx, tmp, flag.p(), pos.p(), O);
		k_sg_slice_compact<<<n_blocks(O, T), T, 0, st>>>(O, flag.p(), pos.p(), w.ocam.p(), w.opt.p(), (uint32_t)ba.pt_begin,
			kept.p(), ocam_l.p(), opt_l.p());
		LAUNCH_CHECK(ctx);
	}
	build_schur_structure_device(ctx, C, P_local, O_local, ocam_l.p(), opt_l.p(), ba.d_obs_orig);
	if(O_local) { // track position -> edge of the slice -> original edge
		k_sg_compose<<<n_blocks(O_local, T), T, 0, st>>>(O_local, ba.d_obs_orig.p(), kept.p());
		LAUNCH_CHECK(ctx);
	}
	ba.z.resize(O_local * 2);
	ba.info.resize(O_local * 4);
	SPP_CUDA(cudaStreamWaitEvent(st, ctx->copy_done, 0)); // the measurements have arrived
	if(O_local) {
		k_sg_gather_rows<2><<<n_blocks(O_local * 2, T), T, 0, st>>>(O_local, ba.d_obs_orig.p(), w.z_in.p(), ba.z.p());
		LAUNCH_CHECK(ctx);
		k_sg_gather_rows<4><<<n_blocks(O_local * 4, T), T, 0, st>>>(O_local, ba.d_obs_orig.p(), w.info_in.p(), ba.info.p());
		LAUNCH_CHECK(ctx);
	}
	// (d) position of this rank's blocks in the global list (host: one sort of the global keys, binary searches)
	s.h_blk_row.resize(s.n_blocks);
	s.h_blk_col.resize(s.n_blocks);
	s.blk_row.download(s.h_blk_row.data(), s.n_blocks, st);
	s.blk_col.download(s.h_blk_col.data(), s.n_blocks, st);
	SPP_CUDA(cudaStreamSynchronize(st));
	{
		const size_t ng = s.n_blocks_global;
		std::vector<uint32_t> slot(s.n_blocks);
		// both lists come out of the same enumeration (diagonal blocks, then the off-diagonal ones by tile and position), so
		// the slice's list is a subsequence of the global one: one merging pass locates every block ...
		size_t n_located = 0;
		for(size_t b = 0, gb = 0; b < s.n_blocks && gb < ng; ++ b) {
			while(gb < ng && (s.h_gblk_row[gb] != s.h_blk_row[b] || s.h_gblk_col[gb] != s.h_blk_col[b])) ++ gb;
			if(gb == ng) break;
			slot[b] = (uint32_t)gb ++;
			++ n_located;
		}
		// ... and should the orders ever differ, a sort of the global keys and binary searches do
		std::vector<std::pair<uint64_t, uint32_t> > keys((n_located == s.n_blocks)? 0 : ng);
		for(size_t b = 0; b < keys.size(); ++ b)
			keys[b] = std::make_pair((uint64_t)s.h_gblk_row[b] * C + s.h_gblk_col[b], (uint32_t)b);
		std::sort(keys.begin(), keys.end());
		for(size_t b = 0; b < s.n_blocks && n_located != s.n_blocks; ++ b) {
			const uint64_t k = (uint64_t)s.h_blk_row[b] * C + s.h_blk_col[b];
			std::vector<std::pair<uint64_t, uint32_t> >::const_iterator it =
				std::lower_bound(keys.begin(), keys.end(), std::make_pair(k, (uint32_t)0));
			if(it == keys.end() || it->first != k)
				throw invalid_error("internal error: a block of this rank is missing from the global block list");
			slot[b] = it->second;
		}
		s.blk_slot.upload(slot, st);
		SPP_CUDA(cudaStreamSynchronize(st));
	}
	ba.host_maps_valid = false;
}

// host copies of the track maps, fetched on demand (spp_ba_get_lambda / spp_ba_get_blocks)
void ba_fetch_host_maps(spp_ctx *ctx)
{
	BAProblem &ba = ctx->ba;
	SchurSystem &s = ctx->sys;
	if(ba.host_maps_valid)
		return;
	ba.obs_orig.resize(s.O); ba.h_obs_cam.resize(s.O); ba.h_obs_pt.resize(s.O);
	ba.d_obs_orig.download(ba.obs_orig.data(), s.O, ctx->stream);
	s.obs_cam.download(ba.h_obs_cam.data(), s.O, ctx->stream);
	s.obs_pt.download(ba.h_obs_pt.data(), s.O, ctx->stream);
	SPP_CUDA(cudaStreamSynchronize(ctx->stream));
	ba.host_maps_valid = true;
}

// host copy of the block pattern, fetched on demand (spp_schur_get_reduced_system)
void schur_fetch_host_pattern(spp_ctx *ctx)
{
	SchurSystem &s = ctx->sys;
	if(s.h_blk_row.size() == s.n_blocks)
		return;
	s.h_blk_row.resize(s.n_blocks);
	s.h_blk_col.resize(s.n_blocks);
	s.blk_row.download(s.h_blk_row.data(), s.n_blocks, ctx->stream);
	s.blk_col.download(s.h_blk_col.data(), s.n_blocks, ctx->stream);
	SPP_CUDA(cudaStreamSynchronize(ctx->stream));
}

} // namespace spp
