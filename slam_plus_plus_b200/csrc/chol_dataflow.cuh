// chol_dataflow.cuh -- the dense FP64 Cholesky as ONE persistent dataflow kernel (included by dense_chol.cu after
// potrf128.cuh, inside namespace spp).
//
// Replaces, for matrices of at least DF_MIN_PANELS panels, the stream-scheduled right-looking factorisation of
// dense_chol_factor_panel(): no launch boundaries, no read-modify-write passes over the trailing matrix, and the
// critical chain (diagonal block -> tile next to it -> next diagonal block) never waits for a launch.
//
// The upper factor R (R^T R = S, column-major, leading dimension ld, 128 x 128 tiles (i, j), i <= j; the columns right
// of the matrix -- right-hand side / row structure of a supernode -- are more tile columns) is computed LEFT-looking:
//
//     R(i, j) = R(i, i)^-T ( S(i, j) - sum_{k < i} R(k, i)^T R(k, j) )
//
// One CTA per SM, three roles handed out by a ticket counter in the order in which the CTAs start (so a CTA that is
// not resident yet is never waited for):
//   chain   (1 CTA)  for every panel i: waits for the final diagonal tile, factors it and inverts the factor
//                    (potrf128_block), publishes f2[i]
//   helpers (8 CTAs) the two tiles on the critical chain, each split into eight 16-column slices:
//                    H1(i): R(i, i+1) = R(i, i)^-T T(i, i+1) by blocked substitution with the inverses of the 16 x 16
//                    diagonal blocks -- available a quarter into the chain CTA's inversion, whose remaining levels (the
//                    full inverse the workers multiply with) thus run beside H1 / H2, off the chain;
//                    H2(i): D(i+1) = T(i+1, i+1) - R(i, i+1)^T R(i, i+1)
//                    (T = the partial sums the workers prepared in advance)
//   workers (rest)   take tile tasks (i, 64-column half) from a queue in row-major order: accumulate the whole sum over k
//                    in registers (128 x 64 accumulator, 8 warps of 32 x 32), waiting for each operand tile's flag,
//                    then multiply by Rinv(i)^T as soon as the chain publishes it, store, publish rdy[i][half].
// Operand tiles move with TMA (cp.async.bulk.tensor, 128-byte swizzle, 16 x 128 / 16 x 64 boxes of K-contiguous
// doubles) through a 6-stage mbarrier pipeline fed by a producer warp; the FP64 tensor-core instruction is
// mma.sync.m8n8k4.f64 (tcgen05 has no FP64 kind). Within a 16-long K chunk a thread feeds the k values
// 8 (t >> 1) + 2 s + (t & 1), s = 0..3: with the TMA swizzle every fragment load is bank-conflict free.
// Every tile is summed in a fixed order (k ascending): bit-reproducible whatever the schedule.
// Inter-CTA hand-off: plain stores, CTA barrier, then one thread's st.release / red.release of a flag (cumulative over
// the barrier, as in a cooperative-groups grid barrier); the reader spins with ld.acquire and issues fence.proxy.async
// before its TMA loads. A watchdog turns a wait that never ends (a bug, not
// a data condition) into an error instead of a hung device.
// Inside a CTA, a consumer warp hands a pipeline stage back to the producer through release_stage(): the mbarrier
// arrive is made control-dependent on the values of the fragments just loaded, because an arrive issued back to back
// with the last ld.shared of the stage let the refill overtake that load ("stage release").
// -DSPP_DF_FENCE_ALL=1 (debugging aid, no effect on results): every writer thread fences before the CTA barrier that
// precedes a flag publication, instead of relying on the cumulativity of the publishing thread's release.
#pragma once
#include <cuda.h>

namespace df {

constexpr int THREADS = 288;        // 8 consumer warps + 1 producer warp
constexpr int G = 8;                // helper CTAs = 16-column slices of a chain tile
constexpr int STAGES = 6;           // worker pipeline: 6 x (16 KB A + 8 KB B)
constexpr int HSTAGES = 8;          // helper pipeline: a whole K = 128 operand in flight, 8 x (16 KB A + 2 KB B)
constexpr uint32_t A_BYTES = 16384, B_BYTES = 8192, HB_BYTES = 2048;
constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES, HSTAGE_BYTES = A_BYTES + HB_BYTES;
constexpr uint32_t STAGING_OFF = STAGES * STAGE_BYTES;           // 147456: the accumulated tile as the solve's B operand
constexpr uint32_t BAR_OFF = STAGING_OFF + 65536;                // 212992
constexpr uint32_t SMEM_BYTES = BAR_OFF + 256 + 1024;            // + alignment slack
static_assert(HSTAGES * HSTAGE_BYTES <= BAR_OFF, "helper stages overlap the barriers");
constexpr uint32_t AS_OFF = STAGING_OFF, YS_OFF = STAGING_OFF + 2048; // helpers (H1): one 16 x 16 block / the solved blocks as B operands
constexpr unsigned SPIN_LIMIT = 1u << 21;
#ifndef SPP_DF_FENCE_ALL
#define SPP_DF_FENCE_ALL 0
#endif
constexpr bool FENCE_ALL = SPP_DF_FENCE_ALL != 0; // every writer fences before the CTA barrier that precedes a flag publication

struct Args {
	double *A;               // the matrix (and the tile columns right of it)
	double *Rinv;            // [NB][128 x 128] inverses of the diagonal blocks of R
	int *info;               // first non-positive pivot (1-based); -1: watchdog
	int *flags;              // see the layout in the kernel; zero on entry
	const uint32_t *tasks;   // worker tasks: (i << 16) | half-tile column
	unsigned long long ld;
	int NB, NJH, n_tasks;
	unsigned long long *dbg; // SPP_CHOL_TIMING: time stamps (ns) of the chain, wait cycles of the workers; null otherwise
};

__device__ __forceinline__ unsigned long long gtime()
{
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
	uint32_t ok;
	asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
		: "=r"(ok) : "r"(bar), "r"(parity) : "memory");
	return ok != 0;
}
__device__ __forceinline__ int ld_acquire(const int *p)
{
	int v;
	asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ int ld_relaxed(const int *p)
{
	int v;
	asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_release(int *p, int v)
{
	asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_add(int *p, int v)
{
	asm volatile("red.release.gpu.global.add.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar)
{
	asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
		:: "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void bar_consumers() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// waits until *flag >= v; gives up (and makes every other wait give up) when the watchdog fires
__device__ __forceinline__ void wait_ge(const int *flag, int v, int *abort_flag, int code)
{
	unsigned spins = 0;
	while(ld_acquire(flag) < v) {
		if(((++ spins) & 127) == 0) {
			if(ld_relaxed(abort_flag))
				return;
			if(spins > SPIN_LIMIT) {
				atomicCAS(abort_flag, 0, code);
				return;
			}
		}
		__nanosleep(32);
	}
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int *abort_flag)
{
	unsigned spins = 0;
	while(!mbar_try_wait(bar, parity)) {
		if(((++ spins) & 63) == 0 && ld_relaxed(abort_flag))
			return;
	}
}

// one 16-long K chunk: acc (MA x NT tiles of 8 x 8) += A^T B, a_row / b_row = shared-memory byte address of the
// operand row (g of the warp tile) + (t & 1) * 8, rows 8 apart are 1024 bytes apart (128-byte rows, swizzled).
// Returns true when a fragment holds a payload no operand holds -- never, but the test makes the caller's release of the
// pipeline stage depend on the value of every fragment, i.e. on every ld.shared of the stage having returned (see
// release_stage)
constexpr int NO_SUCH_HIWORD = 0x7ff12345;
template <int MA, int NT, bool NEG>
__device__ __forceinline__ bool mma_chunk(double (&acc)[MA][NT][2], uint32_t a_row, uint32_t b_row, int g, int t)
{
	const uint32_t x0 = (uint32_t)(((4 * (t >> 1)) ^ g) << 4);
	bool odd = false;
	#pragma unroll
	for(int s = 0; s < 4; ++ s) {
		const uint32_t off = x0 ^ (uint32_t)(s << 4);
		double fa[MA], fb[NT];
		#pragma unroll
		for(int a = 0; a < MA; ++ a) {
			asm volatile("ld.shared.f64 %0, [%1];" : "=d"(fa[a]) : "r"(a_row + a * 1024 + off));
			odd |= __double2hiint(fa[a]) == NO_SUCH_HIWORD;
			if(NEG) fa[a] = -fa[a];
		}
		#pragma unroll
		for(int b = 0; b < NT; ++ b) {
			asm volatile("ld.shared.f64 %0, [%1];" : "=d"(fb[b]) : "r"(b_row + b * 1024 + off));
			odd |= __double2hiint(fb[b]) == NO_SUCH_HIWORD;
		}
		#pragma unroll
		for(int a = 0; a < MA; ++ a)
			#pragma unroll
			for(int b = 0; b < NT; ++ b)
				dmma_m8n8k4(acc[a][b][0], acc[a][b][1], fa[a], fb[b]);
	}
	return odd;
}

// A consumer warp is done with a pipeline stage. The arrive must not be issued while a ld.shared of the stage is still in
// flight: the compiler is free to move the (non-memory) DMMAs behind the arrive, which then follows the last ld.shared
// back to back, and the refill of the stage was seen to overtake that load (wrong fragments in about one factorisation
// in twenty). The branch on the fragments' values orders the arrive behind the loads without waiting for the DMMAs.
__device__ __forceinline__ void release_stage(uint32_t bar_empty_st, bool odd, int lane)
{
	if(odd)
		__nanosleep(20);
	__syncwarp();
	if(lane == 0)
		mbar_arrive(bar_empty_st);
}

} // namespace df

__global__ void __launch_bounds__(df::THREADS, 1) k_chol_dataflow(const __grid_constant__ CUtensorMap mapA,
	const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapH, const __grid_constant__ CUtensorMap mapR,
	const __grid_constant__ CUtensorMap mapX, const df::Args p)
{
	using namespace df;
	extern __shared__ __align__(16) unsigned char df_raw[];
	__shared__ int s_role;
	__shared__ uint32_t s_task[2];
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int g = lane >> 2, t = lane & 3;
	const int NB = p.NB, NJH = p.NJH;
	const size_t ld = p.ld;
	// flags: [0] next task, [1] role ticket, [2] abort, f2[NB] (diagonal block i factored and inverted), h1cnt[NB]
	// (slices of R(i, i+1) done), dcnt[NB] (slices of the final diagonal tile i done), f1[NB] (diagonal block i factored, the
	// inverses of its 16 x 16 diagonal blocks stored), rdy[NB][NJH] (R(i, half) final), part[NB][NJH] (partial sums of a
	// chain tile stored)
	int *f_next = p.flags, *f_ticket = p.flags + 1, *f_abort = p.flags + 2;
	int *f2 = p.flags + 4, *h1cnt = f2 + NB, *dcnt = h1cnt + NB, *f1 = dcnt + NB, *rdy = f1 + NB, *part = rdy + (size_t)NB * NJH;

	if(tid == 0)
		s_role = atomicAdd(f_ticket, 1);
	__syncthreads();
	const int role = s_role;

	if(role == 0) {
		// ---- the chain: diagonal blocks -------------------------------------------------------------------
		if(warp == 8)
			return;
		for(int i = 0; i < NB; ++ i) {
			if(i > 0) {
				if(tid == 0)
					wait_ge(dcnt + i, G, f_abort, 1000 + i);
				POTRF_SYNC();
			}
			if(p.dbg && tid == 0)
				p.dbg[i] = gtime();
			potrf128_block(p.A + ((size_t)i * CH_NB) * ld + (size_t)i * CH_NB, ld, p.Rinv + (size_t)i * (CH_NB * CH_NB), p.info,
				i * CH_NB + 1, 0, [&]() { // R(i, i) and the inverses of its 16 x 16 diagonal blocks are stored: the helpers may start
					if(FENCE_ALL) { __threadfence(); fence_proxy_async(); }
					POTRF_SYNC();
					if(tid == 0) {
						fence_proxy_async();
						st_release(f1 + i, 1);
						if(p.dbg)
							p.dbg[NB + i] = gtime();
					}
				});
			if(FENCE_ALL) { __threadfence(); fence_proxy_async(); }
			POTRF_SYNC();
			if(tid == 0) {
				fence_proxy_async();
				st_release(f2 + i, 1);
				if(p.dbg)
					p.dbg[4 * NB + i] = gtime();
			}
		}
		if(tid == 0 && ld_relaxed(f_abort))
			*p.info = -1;
		return;
	}

	const uint32_t base = (smem_u32(df_raw) + 1023u) & ~1023u;
	const uint32_t bar_full = base + BAR_OFF, bar_empty = bar_full + 64, bar_tfull = bar_full + 128, bar_tempty = bar_full + 144;
	if(tid == 0) {
		for(int s = 0; s < 8; ++ s) {
			mbar_init(bar_full + 8 * s, 1);
			mbar_init(bar_empty + 8 * s, 8);
		}
		for(int e = 0; e < 2; ++ e) {
			mbar_init(bar_tfull + 8 * e, 1);
			mbar_init(bar_tempty + 8 * e, 8);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		fence_proxy_async();
	}
	__syncthreads();
	uint32_t cnt = 0; // K chunks that went through the pipeline (producer and consumers count alike)

	if(role <= G) {
		// ---- helpers: the two tiles next to the diagonal, slice h = 16 columns ------------------------------
		const int h = role - 1;
		if(warp == 8) {
			if(lane != 0)
				return;
			for(int i = 0; i + 1 < NB; ++ i) {
				for(int phase = 0; phase < 2; ++ phase) {
					if(phase == 0) { // H1: R(i, i), the inverses of its 16 x 16 diagonal blocks and the partial sums of tile (i, i+1)
						wait_ge(part + (size_t)i * NJH + 2 * (i + 1), 1, f_abort, 3000 + i);
						wait_ge(part + (size_t)i * NJH + 2 * (i + 1) + 1, 1, f_abort, 3000 + i);
						mbar_arrive(bar_tfull + 8 * (i & 1)); // the consumers fetch the partial sums while the diagonal block is still being factored
						wait_ge(f1 + i, 1, f_abort, 2000 + i);
					} else { // H2: all of R(i, i+1) and the partial sums of the diagonal tile i+1
						wait_ge(h1cnt + i, G, f_abort, 4000 + i);
						wait_ge(part + (size_t)(i + 1) * NJH + 2 * (i + 1), 1, f_abort, 5000 + i);
						wait_ge(part + (size_t)(i + 1) * NJH + 2 * (i + 1) + 1, 1, f_abort, 5000 + i);
					}
					fence_proxy_async();
					for(int c = 0; c < 8; ++ c, ++ cnt) {
						const uint32_t st = cnt % HSTAGES, use = cnt / HSTAGES;
						if(use)
							mbar_wait(bar_empty + 8 * st, (use - 1) & 1, f_abort);
						const uint32_t dst = base + st * HSTAGE_BYTES;
						mbar_expect_tx(bar_full + 8 * st, HSTAGE_BYTES);
						if(phase == 0) { // rows 16 c .. of R(i, i) and the inverse of its c-th 16 x 16 diagonal block
							tma_load_2d(dst, &mapA, i * CH_NB + 16 * c, i * CH_NB, bar_full + 8 * st);
							tma_load_2d(dst + A_BYTES, &mapX, 16 * c, i * CH_NB + 16 * c, bar_full + 8 * st);
						} else { // rows 16 c .. of R(i, i+1): all columns / this helper's slice
							tma_load_2d(dst, &mapA, i * CH_NB + 16 * c, (i + 1) * CH_NB, bar_full + 8 * st);
							tma_load_2d(dst + A_BYTES, &mapH, i * CH_NB + 16 * c, (i + 1) * CH_NB + 16 * h, bar_full + 8 * st);
						}
					}
				}
			}
			return;
		}
		const uint32_t a_off = (uint32_t)((16 * warp + g) * 128 + (t & 1) * 8), b_off = A_BYTES + (uint32_t)(g * 128 + (t & 1) * 8);
		for(int i = 0; i + 1 < NB; ++ i) {
			// element (m, n) of the slice: m = 16 warp + 8 a + g, n = 16 h + 8 b + 2 t + e
			double *out1 = p.A + ((size_t)(i + 1) * CH_NB + 16 * h + 2 * t) * ld + (size_t)i * CH_NB + 16 * warp + g;
			double *out2 = out1 + CH_NB;
			{ // H1: R(i, i+1)(:, slice) = R(i, i)^-T T(:, slice) by blocked forward substitution, 16 rows (one chunk) at a time:
			  // warp w owns rows 16 w ..; step c: warp c multiplies its block by the inverse of the c-th diagonal block (the
			  // solved block goes to shared memory as a B operand and to global memory), the warps below subtract
			  // R(c, w)^T Y_c from theirs
				mbar_wait(bar_tfull + 8 * (i & 1), (i >> 1) & 1, f_abort); // the producer has seen the partial sums' flags
				double acc[2][2][2];
				#pragma unroll
				for(int a = 0; a < 2; ++ a)
					#pragma unroll
					for(int b = 0; b < 2; ++ b) {
						acc[a][b][0] = __ldcg(out1 + (size_t)(8 * b) * ld + 8 * a);
						acc[a][b][1] = __ldcg(out1 + (size_t)(8 * b + 1) * ld + 8 * a);
					}
				const uint32_t row_off = (uint32_t)(g * 128 + (t & 1) * 8);
				for(int c = 0; c < 8; ++ c, ++ cnt) {
					const uint32_t st = cnt % HSTAGES, use = cnt / HSTAGES;
					mbar_wait(bar_full + 8 * st, use & 1, f_abort);
					const uint32_t stage = base + st * HSTAGE_BYTES;
					bool odd = false;
					if(warp == c) {
						#pragma unroll
						for(int a = 0; a < 2; ++ a)
							#pragma unroll
							for(int b = 0; b < 2; ++ b)
								#pragma unroll
								for(int q = 0; q < 2; ++ q) { // element (k, n) of the block -> row n, swizzled
									const int k = 8 * a + g, n = 8 * b + 2 * t + q;
									asm volatile("st.shared.f64 [%0], %1;" :: "r"(base + AS_OFF + (uint32_t)(n * 128 + (((k >> 1) ^ (n & 7)) << 4) + (k & 1) * 8)),
										"d"(acc[a][b][q]) : "memory");
								}
						__syncwarp();
						double y[2][2][2] = {};
						odd = mma_chunk<2, 2, false>(y, stage + A_BYTES + row_off, base + AS_OFF + row_off, g, t);
						#pragma unroll
						for(int a = 0; a < 2; ++ a)
							#pragma unroll
							for(int b = 0; b < 2; ++ b) {
								#pragma unroll
								for(int q = 0; q < 2; ++ q) {
									const int k = 8 * a + g, n = 8 * b + 2 * t + q;
									asm volatile("st.shared.f64 [%0], %1;" :: "r"(base + YS_OFF + (uint32_t)(c * 2048 + n * 128 + (((k >> 1) ^ (n & 7)) << 4) + (k & 1) * 8)),
										"d"(y[a][b][q]) : "memory");
								}
								out1[(size_t)(8 * b) * ld + 8 * a] = y[a][b][0];
								out1[(size_t)(8 * b + 1) * ld + 8 * a] = y[a][b][1];
							}
					}
					bar_consumers();
					if(warp > c)
						odd = mma_chunk<2, 2, true>(acc, stage + (uint32_t)(16 * warp * 128) + row_off, base + YS_OFF + (uint32_t)(c * 2048) + row_off, g, t);
					release_stage(bar_empty + 8 * st, odd, lane);
				}
				bar_consumers();
				if(tid == 0) {
					fence_proxy_async();
					red_release_add(h1cnt + i, 1);
					if(p.dbg && h == 0)
						p.dbg[2 * NB + i] = gtime();
				}
			}
			{ // H2: D(i+1)(:, slice) = T(i+1, i+1)(:, slice) - R(i, i+1)^T R(i, i+1)(:, slice), rows above the diagonal: warps <= h
				double acc[2][2][2] = {}, c_in[2][2][2] = {};
				for(int c = 0; c < 8; ++ c, ++ cnt) {
					const uint32_t st = cnt % HSTAGES, use = cnt / HSTAGES;
					mbar_wait(bar_full + 8 * st, use & 1, f_abort);
					if(c == 0 && warp <= h) { // the producer has seen the partial sums' flags: fetch them behind the DMMAs
						#pragma unroll
						for(int a = 0; a < 2; ++ a)
							#pragma unroll
							for(int b = 0; b < 2; ++ b) {
								c_in[a][b][0] = __ldcg(out2 + (size_t)(8 * b) * ld + 8 * a);
								c_in[a][b][1] = __ldcg(out2 + (size_t)(8 * b + 1) * ld + 8 * a);
							}
					}
					bool odd = false;
					if(warp <= h)
						odd = mma_chunk<2, 2, true>(acc, base + st * HSTAGE_BYTES + a_off, base + st * HSTAGE_BYTES + b_off, g, t);
					release_stage(bar_empty + 8 * st, odd, lane);
				}
				if(warp <= h) {
					#pragma unroll
					for(int a = 0; a < 2; ++ a)
						#pragma unroll
						for(int b = 0; b < 2; ++ b) {
							out2[(size_t)(8 * b) * ld + 8 * a] = c_in[a][b][0] + acc[a][b][0];
							out2[(size_t)(8 * b + 1) * ld + 8 * a] = c_in[a][b][1] + acc[a][b][1];
						}
				}
				if(FENCE_ALL) { __threadfence(); fence_proxy_async(); }
				bar_consumers();
				if(tid == 0) {
					red_release_add(dcnt + i + 1, 1);
					if(p.dbg && h == 0)
						p.dbg[3 * NB + i] = gtime();
				}
			}
		}
		return;
	}

	// ---- workers ----------------------------------------------------------------------------------------
	if(warp == 8) {
		if(lane != 0)
			return;
		long long t_flags = 0, t_trsm = 0; // cycles spent waiting for operand flags / for the diagonal block
		for(uint32_t tn = 0;; ++ tn) {
			const uint32_t e = tn & 1;
			if(tn >= 2)
				mbar_wait(bar_tempty + 8 * e, ((tn >> 1) - 1) & 1, f_abort);
			const int tk = atomicAdd(f_next, 1);
			const uint32_t code = (tk < p.n_tasks)? p.tasks[tk] : 0xffffffffu;
			s_task[e] = code;
			mbar_arrive(bar_tfull + 8 * e);
			if(code == 0xffffffffu) {
				if(p.dbg) {
					unsigned long long *d = p.dbg + 5 * NB + 8 * blockIdx.x;
					d[0] = (unsigned long long)t_flags; d[1] = (unsigned long long)t_trsm; d[2] = tn; d[3] = gtime();
				}
				return;
			}
			const int i = int(code >> 16), jh = int(code & 0xffff), j = jh >> 1;
			const bool partial = j == i || (j == i + 1 && j < NB);
			const int kmax = (j == i)? ((i > 0)? i - 1 : 0) : i;
			for(int k = 0; k < kmax; ++ k) {
				const long long c0 = clock64();
				// operands of slab k: R(k, i) (both halves) and R(k, half jh); a tile next to the diagonal comes from the helpers
				if(i == k + 1)
					wait_ge(h1cnt + k, G, f_abort, 6000 + k);
				else {
					wait_ge(rdy + (size_t)k * NJH + 2 * i, 1, f_abort, 7000 + k);
					wait_ge(rdy + (size_t)k * NJH + 2 * i + 1, 1, f_abort, 7000 + k);
				}
				if(j != i) {
					if(j == k + 1 && j < NB)
						wait_ge(h1cnt + k, G, f_abort, 8000 + k);
					else
						wait_ge(rdy + (size_t)k * NJH + jh, 1, f_abort, 9000 + k);
				}
				t_flags += clock64() - c0;
				fence_proxy_async();
				for(int c = 0; c < 8; ++ c, ++ cnt) {
					const uint32_t st = cnt % STAGES, use = cnt / STAGES;
					if(use)
						mbar_wait(bar_empty + 8 * st, (use - 1) & 1, f_abort);
					const uint32_t dst = base + st * STAGE_BYTES;
					mbar_expect_tx(bar_full + 8 * st, STAGE_BYTES);
					tma_load_2d(dst, &mapA, k * CH_NB + 16 * c, i * CH_NB, bar_full + 8 * st);
					tma_load_2d(dst + A_BYTES, &mapB, k * CH_NB + 16 * c, jh * 64, bar_full + 8 * st);
				}
			}
			if(!partial) {
				const long long c0 = clock64();
				wait_ge(f2 + i, 1, f_abort, 10000 + i);
				t_trsm += clock64() - c0;
				fence_proxy_async();
				for(int c = 0; c < 8; ++ c, ++ cnt) {
					const uint32_t st = cnt % STAGES, use = cnt / STAGES;
					if(use)
						mbar_wait(bar_empty + 8 * st, (use - 1) & 1, f_abort);
					mbar_expect_tx(bar_full + 8 * st, A_BYTES);
					tma_load_2d(base + st * STAGE_BYTES, &mapR, 16 * c, i * CH_NB, bar_full + 8 * st);
				}
			}
		}
	}

	const int wi = (warp >> 1) * 32, wj = (warp & 1) * 32; // 4 x 2 warps of 32 x 32 over the 128 x 64 tile
	const uint32_t a_off = (uint32_t)((wi + g) * 128 + (t & 1) * 8), b_off = (uint32_t)((wj + g) * 128 + (t & 1) * 8);
	long long t_pipe = 0, t_start = clock64(); // cycles this warp waited for operand chunks
	for(uint32_t tn = 0;; ++ tn) {
		const uint32_t e = tn & 1;
		mbar_wait(bar_tfull + 8 * e, (tn >> 1) & 1, f_abort);
		// lane 0 reads the task and hands it to the others; its release of the entry depends on the value read
		uint32_t code = 0;
		if(lane == 0) {
			asm volatile("ld.shared.u32 %0, [%1];" : "=r"(code) : "r"(smem_u32(&s_task[e])) : "memory");
			if(code != 0xfffffffeu) // (no task has this code: the load has returned when the entry is released)
				mbar_arrive(bar_tempty + 8 * e);
		}
		code = __shfl_sync(0xffffffffu, code, 0);
		if(code == 0xffffffffu) {
			if(p.dbg && tid == 0) {
				unsigned long long *d = p.dbg + 5 * NB + 8 * blockIdx.x;
				d[4] = (unsigned long long)t_pipe; d[5] = (unsigned long long)(clock64() - t_start);
			}
			return;
		}
		const int i = int(code >> 16), jh = int(code & 0xffff), j = jh >> 1;
		const bool partial = j == i || (j == i + 1 && j < NB);
		const int kmax = (j == i)? ((i > 0)? i - 1 : 0) : i;
		// element (m, n) of the tile: m = wi + 8 a + g, n = wj + 8 b + 2 t + e
		double *out = p.A + ((size_t)jh * 64 + wj + 2 * t) * ld + (size_t)i * CH_NB + wi + g;
		double acc[4][4][2];
		#pragma unroll
		for(int a = 0; a < 4; ++ a)
			#pragma unroll
			for(int b = 0; b < 4; ++ b) {
				acc[a][b][0] = out[(size_t)(8 * b) * ld + 8 * a];
				acc[a][b][1] = out[(size_t)(8 * b + 1) * ld + 8 * a];
			}
		for(int c = 0; c < 8 * kmax; ++ c, ++ cnt) {
			const uint32_t st = cnt % STAGES, use = cnt / STAGES;
			const long long c0 = clock64();
			mbar_wait(bar_full + 8 * st, use & 1, f_abort);
			t_pipe += clock64() - c0;
			const bool odd = mma_chunk<4, 4, true>(acc, base + st * STAGE_BYTES + a_off, base + st * STAGE_BYTES + A_BYTES + b_off, g, t);
			release_stage(bar_empty + 8 * st, odd, lane);
		}
		if(partial) {
			if(kmax > 0) {
				#pragma unroll
				for(int a = 0; a < 4; ++ a)
					#pragma unroll
					for(int b = 0; b < 4; ++ b) {
						out[(size_t)(8 * b) * ld + 8 * a] = acc[a][b][0];
						out[(size_t)(8 * b + 1) * ld + 8 * a] = acc[a][b][1];
					}
			}
			if(FENCE_ALL) { __threadfence(); fence_proxy_async(); }
			bar_consumers();
			if(tid == 0) {
				fence_proxy_async();
				st_release(part + (size_t)i * NJH + jh, 1);
			}
			continue;
		}
		// the accumulated tile becomes the B operand of the solve: element (k, n) -> chunk k / 16, row n, swizzled
		#pragma unroll
		for(int a = 0; a < 4; ++ a) {
			const int k = wi + 8 * a + g;
			#pragma unroll
			for(int b = 0; b < 4; ++ b) {
				#pragma unroll
				for(int q = 0; q < 2; ++ q) {
					const int n = wj + 8 * b + 2 * t + q;
					const uint32_t addr = base + STAGING_OFF + (uint32_t)((k >> 4) * 8192 + n * 128 + ((((k & 15) >> 1) ^ (n & 7)) << 4) + (k & 1) * 8);
					asm volatile("st.shared.f64 [%0], %1;" :: "r"(addr), "d"(acc[a][b][q]) : "memory");
				}
			}
		}
		bar_consumers();
		#pragma unroll
		for(int a = 0; a < 4; ++ a)
			#pragma unroll
			for(int b = 0; b < 4; ++ b)
				acc[a][b][0] = acc[a][b][1] = 0;
		for(int c = 0; c < 8; ++ c, ++ cnt) { // R(i, half) = Rinv(i)^T T; Rinv(k, m) = 0 for k > m
			const uint32_t st = cnt % STAGES, use = cnt / STAGES;
			const long long c0 = clock64();
			mbar_wait(bar_full + 8 * st, use & 1, f_abort);
			t_pipe += clock64() - c0;
			bool odd = false;
			if(16 * c <= wi + 31)
				odd = mma_chunk<4, 4, false>(acc, base + st * STAGE_BYTES + a_off, base + STAGING_OFF + c * 8192 + b_off, g, t);
			release_stage(bar_empty + 8 * st, odd, lane);
		}
		#pragma unroll
		for(int a = 0; a < 4; ++ a)
			#pragma unroll
			for(int b = 0; b < 4; ++ b) {
				out[(size_t)(8 * b) * ld + 8 * a] = acc[a][b][0];
				out[(size_t)(8 * b + 1) * ld + 8 * a] = acc[a][b][1];
			}
		if(FENCE_ALL) { __threadfence(); fence_proxy_async(); }
		bar_consumers();
		if(tid == 0) {
			fence_proxy_async();
			st_release(rdy + (size_t)i * NJH + jh, 1);
		}
	}
}
