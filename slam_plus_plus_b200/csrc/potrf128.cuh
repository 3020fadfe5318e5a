// potrf128.cuh -- the diagonal-block kernel of the dense FP64 Cholesky (see dense_chol.cu): factor a 128 x 128
// block in one CTA and invert the triangular factor. Kept in a header so that tools/micro/potrf_bench.cu can
// time it in isolation. Include inside namespace spp after CH_NB is defined.
#pragma once
#include <type_traits>
#include <cuda_pipeline_primitives.h>

// ---- diagonal block: factor + invert -----------------------------------------------------------------
//
// One CTA, 256 threads, the 128 x 128 block column-major in shared memory (leading dimension 132: the DMMA fragment
// loads below are bank-conflict free for LD = 4 mod 16). Three levels:
//   8 x 8 leaf    warps 0..3 (one per SM sub-partition) each factor the leaf redundantly in registers -- no
//                 shuffles, no broadcast: the serial chain per pivot is rsqrt -> multiply -> FMA -- and solve
//                 the row block right of the leaf by forward substitution, one column per thread
//   rank-8 update of everything below the leaf's row block on DMMA, with a look-ahead inside the CTA: warps 0..2
//                 update the next leaf's tile row and go on with that leaf (the critical chain) while warps
//                 4..7 update the rest
// The inverse of the factor (needed by the GEMM-shaped panel solve and by the backward solve) is built in place
// afterwards: the sixteen leaf inverses in registers, then recursive doubling X12 = -X11 (R12 X22) on
// 8 -> 16 -> 32 -> 64 -> 128 wide blocks on DMMA, skipping the zero halves of the triangular operands.

#define P3_LD 132
#define PT 256              // threads of k_potrf128
#define PW (PT / 32)
// barrier of the PT threads that factor the block: the persistent dataflow kernel (chol_dataflow.cuh) runs this body in
// a CTA that has one more (producer) warp, which does not take part
#define POTRF_SYNC() asm volatile("bar.sync 0, 256;" ::: "memory")

__device__ __forceinline__ void dmma_m8n8k4(double &d0, double &d1, double a, double b)
{
	// volatile: the DMMAs stay in program order, which is written so that independent chains interleave
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
		: "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// one doubling level of the in-place triangular inversion: blocks [a, a + H) and [a + H, a + 2H) of every pair
// become one inverted block of width 2H. A pair is served by TH = H / 8 warps. In the first pass, T = R12 X22,
// warp wq of the pair owns the tile columns {c, TH-1-c}, c = wq / 2, and the upper or lower half of the tile rows
// (h = wq & 1); X22 is upper triangular, so column c needs (c + 1) * 8 values of k: pairing c with TH-1-c gives
// every warp the same number of DMMAs. The second pass, X12 = -X11 T, mirrors that with rows and columns swapped.
// The k loops are split so that their bodies are branch free (a branch around a DMMA keeps ptxas from overlapping
// the fragment loads of one tile with the DMMAs of another).
template <int H>
__device__ __forceinline__ void inv_level(double *__restrict__ sm, int warp, int g, int t)
{
	constexpr int TH = H / 8, NH = (TH >= 2)? TH / 2 : 1;
	const int pr = warp / TH, wq = warp % TH, a = pr * 2 * H;
	const int c = (TH >= 2)? wq >> 1 : 0, h = (TH >= 2)? wq & 1 : 0;
	const int lo = c, hi = (TH >= 2)? TH - 1 - c : 0; // lo <= hi
	double c0[2][NH], c1[2][NH];
	{
		#pragma unroll
		for(int u = 0; u < 2; ++ u)
			#pragma unroll
			for(int q = 0; q < NH; ++ q) c0[u][q] = c1[u][q] = 0;
		const double *pA = sm + (a + H + t) * P3_LD + a + h * NH * 8 + g; // R12((h NH + q) * 8 + g, k + t)
		const double *pB = sm + (a + H + g) * P3_LD + a + H + t;          // X22(k + t, col * 8 + g)
		const double *pBlo = pB + lo * 8 * P3_LD, *pBhi = pB + hi * 8 * P3_LD;
		// X22 is upper triangular: column lo needs k < (lo + 1) * 8, column hi k < (hi + 1) * 8
		int k = 0;
		if(TH >= 2) {
			#pragma unroll 2
			for(; k < (lo + 1) * 8; k += 4) {
				const double blo = pBlo[k], bhi = pBhi[k];
				#pragma unroll
				for(int q = 0; q < NH; ++ q) {
					const double av = pA[k * P3_LD + q * 8];
					dmma_m8n8k4(c0[0][q], c1[0][q], av, blo);
					dmma_m8n8k4(c0[1][q], c1[1][q], av, bhi);
				}
			}
		}
		#pragma unroll 2
		for(; k < (hi + 1) * 8; k += 4) {
			const double bhi = pBhi[k];
			#pragma unroll
			for(int q = 0; q < NH; ++ q)
				dmma_m8n8k4(c0[1][q], c1[1][q], pA[k * P3_LD + q * 8], bhi);
		}
		POTRF_SYNC();
		#pragma unroll
		for(int u = (TH >= 2)? 0 : 1; u < 2; ++ u) {
			#pragma unroll
			for(int q = 0; q < NH; ++ q) {
				double *pc = sm + (a + H + (u? hi : lo) * 8 + 2 * t) * P3_LD + a + (h * NH + q) * 8 + g;
				pc[0] = c0[u][q];
				pc[P3_LD] = c1[u][q];
			}
		}
		POTRF_SYNC();
	}
	{
		#pragma unroll
		for(int u = 0; u < 2; ++ u)
			#pragma unroll
			for(int q = 0; q < NH; ++ q) c0[u][q] = c1[u][q] = 0;
		const double *pA = sm + (a + t) * P3_LD + a + g;                        // X11(row * 8 + g, k + t)
		const double *pB = sm + (a + H + h * NH * 8 + g) * P3_LD + a + t;       // T(k + t, (h NH + q) * 8 + g)
		const double *pAlo = pA + lo * 8, *pAhi = pA + hi * 8;
		// X11 is upper triangular: row lo needs k >= lo * 8, row hi k >= hi * 8
		int k = lo * 8;
		#pragma unroll 2
		for(; k < hi * 8; k += 4) {
			const double alo = pAlo[k * P3_LD];
			#pragma unroll
			for(int q = 0; q < NH; ++ q)
				dmma_m8n8k4(c0[0][q], c1[0][q], alo, pB[q * 8 * P3_LD + k]);
		}
		#pragma unroll 2
		for(; k < H; k += 4) {
			const double alo = pAlo[k * P3_LD], ahi = pAhi[k * P3_LD];
			#pragma unroll
			for(int q = 0; q < NH; ++ q) {
				const double bv = pB[q * 8 * P3_LD + k];
				if(TH >= 2) dmma_m8n8k4(c0[0][q], c1[0][q], alo, bv);
				dmma_m8n8k4(c0[1][q], c1[1][q], ahi, bv);
			}
		}
		POTRF_SYNC();
		#pragma unroll
		for(int u = (TH >= 2)? 0 : 1; u < 2; ++ u) {
			#pragma unroll
			for(int q = 0; q < NH; ++ q) {
				double *pc = sm + (a + H + (h * NH + q) * 8 + 2 * t) * P3_LD + a + (u? hi : lo) * 8 + g;
				pc[0] = -c0[u][q];
				pc[P3_LD] = -c1[u][q];
			}
		}
		POTRF_SYNC();
	}
}

// stores the upper-triangular m[8][8] into an 8 x 8 tile (with zeros below the diagonal if ZEROS); called by ONE lane
// of a group of lanes that all hold the same m -- the register file cannot be indexed by lane, and a shared-memory
// store instruction costs ~10 issue cycles whatever the number of active lanes
template <bool ZEROS>
__device__ __forceinline__ void store_leaf(double *dst /* &TC(o, o) */, const double (&m)[8][8])
{
	#pragma unroll
	for(int j = 0; j < 8; ++ j) {
		#pragma unroll
		for(int i = 0; i < 8; i += 2) {
			if(ZEROS || i <= j)
				*reinterpret_cast<double2*>(dst + j * P3_LD + i) = make_double2((i <= j)? m[i][j] : 0.0, (i + 1 <= j)? m[i + 1][j] : 0.0);
		}
	}
}

// ---- trailing updates, split for the look-ahead inside the CTA ------------------------------------------
// After the row block of leaf s is solved, the rank-8 update C(i, j) -= sum_k P(k, i) P(k, j) with its 8 rows is split:
//   urgent      tile row s + 1 (the next leaf and its row block): warps 0..2 (warp 3 meanwhile stores the factored
//               leaf), then warps 0..3 go on with the next leaf
//   lazy        every tile row below it: warps 4..7, concurrently with that leaf
// Both are written "all loads, then two rounds of independent DMMAs" with branch-free bodies (a branch around a
// DMMA keeps ptxas from overlapping the fragment loads of one tile with the DMMAs of another).

__device__ __forceinline__ void named_barrier(int id, int n_threads)
{
	asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(n_threads) : "memory");
}

// urgent: tiles (tr, tr + w + 3 m), m < NS, of tile row tr with the rows k0 .. k0 + 7; w = warp 0..2
template <int NS>
__device__ __forceinline__ void urgent_update(double *__restrict__ sm, int k0, int tr, int w, int g, int t)
{
	const double *pa = sm + (tr * 8 + g) * P3_LD + k0 + t;
	const double a0 = -pa[0], a1 = -pa[4];
	double b0[NS], b1[NS], c0[NS], c1[NS];
	double *pc[NS];
	#pragma unroll
	for(int m = 0; m < NS; ++ m) {
		const int tj = tr + w + 3 * m;
		const double *pb = sm + (tj * 8 + g) * P3_LD + k0 + t;
		pc[m] = sm + (tj * 8 + 2 * t) * P3_LD + tr * 8 + g;
		c0[m] = pc[m][0];
		c1[m] = pc[m][P3_LD];
		b0[m] = pb[0];
		b1[m] = pb[4];
	}
	#pragma unroll
	for(int m = 0; m < NS; ++ m)
		dmma_m8n8k4(c0[m], c1[m], a0, b0[m]);
	#pragma unroll
	for(int m = 0; m < NS; ++ m)
		dmma_m8n8k4(c0[m], c1[m], a1, b1[m]);
	#pragma unroll
	for(int m = 0; m < NS; ++ m) {
		pc[m][0] = c0[m];
		pc[m][P3_LD] = c1[m];
	}
}

__device__ __forceinline__ void urgent_update_dispatch(double *__restrict__ sm, int k0, int tr, int w, int g, int t)
{
	const int ns = (16 - tr - w + 2) / 3; // tiles of this warp: columns tr + w + 3 m < 16
	if(ns >= 5) urgent_update<5>(sm, k0, tr, w, g, t);
	else if(ns == 4) urgent_update<4>(sm, k0, tr, w, g, t);
	else if(ns == 3) urgent_update<3>(sm, k0, tr, w, g, t);
	else if(ns == 2) urgent_update<2>(sm, k0, tr, w, g, t);
	else if(ns == 1) urgent_update<1>(sm, k0, tr, w, g, t);
}

// lazy: all tiles (ti, tj), tr0 <= ti <= tj, with the rows k0 .. k0 + 7. Warp wq = 0..3 owns the tile columns
// wq + 4 m and walks each of them in chunks of four tile rows (the B fragment is shared); tiles outside the
// triangle are computed and dropped.
__device__ __forceinline__ void lazy_update(double *__restrict__ sm, int k0, int tr0, int wq, int g, int t)
{
	const double *pk = sm + k0 + t; // P(k = t, column .)
	#pragma unroll 1
	for(int tj = wq + ((tr0 - wq + 3) & ~3); tj < 16; tj += 4) { // first owned column >= tr0
		const double b0 = pk[(tj * 8 + g) * P3_LD], b1 = pk[(tj * 8 + g) * P3_LD + 4];
		#pragma unroll 1
		for(int ti0 = tr0; ti0 <= tj; ti0 += 4) {
			double a0[4], a1[4], c0[4], c1[4];
			double *pc = sm + (tj * 8 + 2 * t) * P3_LD + ti0 * 8 + g;
			#pragma unroll
			for(int rr = 0; rr < 4; ++ rr) {
				const int ti = (ti0 + rr <= tj)? ti0 + rr : ti0;
				a0[rr] = -pk[(ti * 8 + g) * P3_LD];
				a1[rr] = -pk[(ti * 8 + g) * P3_LD + 4];
				c0[rr] = pc[(ti - ti0) * 8];
				c1[rr] = pc[(ti - ti0) * 8 + P3_LD];
			}
			#pragma unroll
			for(int rr = 0; rr < 4; ++ rr)
				dmma_m8n8k4(c0[rr], c1[rr], a0[rr], b0);
			#pragma unroll
			for(int rr = 0; rr < 4; ++ rr)
				dmma_m8n8k4(c0[rr], c1[rr], a1[rr], b1);
			#pragma unroll
			for(int rr = 0; rr < 4; ++ rr) {
				if(ti0 + rr <= tj) {
					pc[rr * 8] = c0[rr];
					pc[rr * 8 + P3_LD] = c1[rr];
				}
			}
		}
	}
}

// The body: factors the 128 x 128 block at Akk (leading dimension ld) in place (upper triangle), writes the inverse of
// the factor to Rinv_out; a non-positive pivot is reported as *info = n_info_value (first report wins). All PT threads of
// the CTA take part; the whole dynamic shared memory of the CTA is the work tile. Ends with global stores: the caller
// fences / synchronises before anyone else may read them. Shared between the one-block kernel below and the persistent
// chain kernel of the dataflow factorisation (chol_dataflow.cu).
// after_x16: called by all PT threads once the inverses of the eight 16 x 16 diagonal blocks of the factor are in
// Rinv_out (together with the factor itself in Akk): the dataflow kernel's helpers solve the tile next to the diagonal by
// blocked substitution with these, so the rest of the inversion (three of four doubling levels, most of its time) is
// off the critical chain.
struct PotrfNoHook { __device__ __forceinline__ void operator ()() const {} };

template <class Hook>
__device__ __forceinline__ void potrf128_block(double *__restrict__ Akk, size_t ld,
	double *__restrict__ Rinv_out, int *__restrict__ info, int n_info_value, long long *__restrict__ dbg, Hook after_x16)
{
#define DBG_MARK(i) do { if(dbg && threadIdx.x == 0) dbg[i] = clock64(); } while(0)
#ifdef POTRF_TRACE // per-warp time stamps for tools/micro/potrf_bench.cu: dbg[16 + warp * 64 + i]
#define TR(i) do { if(dbg && (threadIdx.x & 31) == 0) dbg[16 + (threadIdx.x >> 5) * 64 + (i)] = clock64(); } while(0)
#else
#define TR(i) do { } while(0)
#endif
#define TC(r, c) sm[(c) * P3_LD + (r)]
	extern __shared__ __align__(16) double sm[]; // the tile, column-major, P3_LD x 128
	double *rdv = sm + P3_LD * CH_NB;        // 128 reciprocal pivots
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int g = lane >> 2, t = lane & 3;
	DBG_MARK(0);
	// the upper 8 x 8 tiles, whole tile columns with 16-byte async copies; column c is paired with column 127 - c, which
	// makes 68 copies per pair (no idle trips over the lower triangle: the load is on the critical chain of the
	// factorisation)
	for(int idx = tid; idx < (CH_NB / 2) * 68; idx += PT) {
		const int c0 = idx / 68, q = idx - c0 * 68, n1 = ((c0 >> 3) + 1) * 4;
		const int c = (q < n1)? c0 : CH_NB - 1 - c0, r2 = 2 * ((q < n1)? q : q - n1);
		__pipeline_memcpy_async(&TC(r2, c), Akk + (size_t)c * ld + r2, 16);
	}
	__pipeline_commit();
	__pipeline_wait_prior(0);
	POTRF_SYNC();
	DBG_MARK(1);

	bool bad = false;
	double r[8][8], rd[8]; // the current leaf (warps 0..3)
	#pragma unroll 1
	for(int s = 0; s < 16; ++ s) {
		const int o = s * 8;
		if(s == 2 || s == 4) TR(s * 8 + 0);
		if(warp < 4) {
			if(s) {
				// the rank-8 update that follows leaf s - 1, urgent part: tile row s. Warp 3 meanwhile puts the
				// factored leaf s - 1 back (nobody reads that block any more)
				if(warp == 3) {
					if(lane == 0) {
						store_leaf<false>(&TC(o - 8, o - 8), r);
						#pragma unroll
						for(int j = 0; j < 8; j += 2)
							*reinterpret_cast<double2*>(rdv + o - 8 + j) = make_double2(rd[j], rd[j + 1]);
					}
				} else
					urgent_update_dispatch(sm, o - 8, s, warp, g, t);
				if(s == 2 || s == 4) TR(s * 8 + 6);
				named_barrier(1, 128);
			}
			if(s == 2 || s == 4) TR(s * 8 + 1);
			// ---- leaf: upper 8 x 8 at (o, o)
			#pragma unroll
			for(int j = 0; j < 8; ++ j) {
				#pragma unroll
				for(int i = 0; i <= j; i += 2) {
					const double2 v = *reinterpret_cast<const double2*>(&TC(o + i, o + j));
					r[i][j] = v.x;
					if(i + 1 <= j) r[i + 1][j] = v.y;
				}
			}
			#pragma unroll
			for(int j = 0; j < 8; ++ j) {
				double piv = r[j][j];
				if(!(piv > 0)) { // Eigen's LLT stops at a non-positive pivot (NaN fails the test as well)
					bad = true;
					piv = 1;
				}
				rd[j] = rsqrt(piv);
				r[j][j] = piv * rd[j];
				#pragma unroll
				for(int c = j + 1; c < 8; ++ c)
					r[j][c] *= rd[j];
				#pragma unroll
				for(int i = j + 1; i < 8; ++ i) {
					#pragma unroll
					for(int c = i; c < 8; ++ c)
						r[i][c] -= r[j][i] * r[j][c];
				}
			}
			if(s == 2 || s == 4) TR(s * 8 + 2);
			// ---- row block right of the leaf: forward substitution R_leaf^T y = b, one column per thread
			const int c = o + 8 + tid;
			if(c < CH_NB) {
				double b[8];
				#pragma unroll
				for(int i = 0; i < 8; i += 2) {
					const double2 v = *reinterpret_cast<const double2*>(&TC(o + i, c));
					b[i] = v.x; b[i + 1] = v.y;
				}
				#pragma unroll
				for(int k = 0; k < 8; ++ k) {
					b[k] *= rd[k];
					#pragma unroll
					for(int i = k + 1; i < 8; ++ i)
						b[i] -= r[k][i] * b[k];
				}
				#pragma unroll
				for(int i = 0; i < 8; i += 2)
					*reinterpret_cast<double2*>(&TC(o + i, c)) = make_double2(b[i], b[i + 1]);
			}
			if(s == 2 || s == 4) TR(s * 8 + 3);
		} else if(s) {
			// lazy part of the same update: every tile row below tile row s, concurrently with leaf s
			lazy_update(sm, o - 8, s + 1, warp - 4, g, t);
			if(s == 2 || s == 4) TR(s * 8 + 3);
		}
		POTRF_SYNC();
		if(s == 2 || s == 4) TR(s * 8 + 4);
		if(s == 0) DBG_MARK(2);
		if(s == 1) DBG_MARK(3);
		if(s == 4) DBG_MARK(4);
	}
	if(warp == 3 && lane == 0) {
		store_leaf<false>(&TC(120, 120), r);
		#pragma unroll
		for(int j = 0; j < 8; j += 2)
			*reinterpret_cast<double2*>(rdv + 120 + j) = make_double2(rd[j], rd[j + 1]);
	}
	POTRF_SYNC();
	DBG_MARK(5);
	if(bad && tid == 0 && *info == 0)
		*info = n_info_value;
	// R to global: column c (c + 1 entries) paired with column 127 - c, 129 entries per pair
	#pragma unroll 8
	for(int idx = tid; idx < (CH_NB / 2) * 129; idx += PT) {
		const int c0 = idx / 129, q = idx - c0 * 129;
		const int c = (q <= c0)? c0 : CH_NB - 1 - c0, r = (q <= c0)? q : q - c0 - 1;
		Akk[(size_t)c * ld + r] = TC(r, c);
	}
	DBG_MARK(6);

	// ---- inverse of the upper-triangular factor, in place ----
	// level 0: the leaf inverses; half-warp h of warp w inverts leaf 2 w + h in registers (every lane of the half
	// redundantly), lane j of the half publishes column j (full 8 x 8 tiles, explicit zeros below the diagonal)
	{
		const int s = warp * 2 + (lane >> 4), o = s * 8;
		double r[8][8], x[8][8], rd[8];
		#pragma unroll
		for(int j = 0; j < 8; ++ j) {
			rd[j] = rdv[o + j];
			#pragma unroll
			for(int i = 0; i <= j; i += 2) {
				const double2 v = *reinterpret_cast<const double2*>(&TC(o + i, o + j));
				r[i][j] = v.x;
				if(i + 1 <= j) r[i + 1][j] = v.y;
			}
		}
		#pragma unroll
		for(int j = 0; j < 8; ++ j) {
			x[j][j] = rd[j];
			#pragma unroll
			for(int i = j - 1; i >= 0; -- i) {
				double acc = 0; // the newest term last: the chain per row is one FMA and one multiply
				#pragma unroll
				for(int k = j; k > i; -- k)
					acc += r[i][k] * x[k][j];
				x[i][j] = -rd[i] * acc;
			}
		}
		TR(40);
		POTRF_SYNC(); // every R leaf has been read and stored to global
		TR(41);
		if(!(lane & 15))
			store_leaf<true>(&TC(o, o), x);
		TR(47);
	}
	POTRF_SYNC();
	TR(42);
	inv_level<8>(sm, warp, g, t);
	TR(43);
	if constexpr(!std::is_same<Hook, PotrfNoHook>::value) {
		// the 16 x 16 diagonal blocks of the inverse (upper parts; the buffer is zero below the diagonal)
		#pragma unroll 4
		for(int idx = tid; idx < 8 * 16 * 16; idx += PT) {
			const int b = idx >> 8, c = (idx >> 4) & 15, r = idx & 15;
			if(r <= c)
				Rinv_out[(size_t)(16 * b + c) * CH_NB + 16 * b + r] = TC(16 * b + r, 16 * b + c);
		}
		after_x16();
	}
	inv_level<16>(sm, warp, g, t);
	TR(44);
	inv_level<32>(sm, warp, g, t);
	TR(45);
	inv_level<64>(sm, warp, g, t);
	TR(46);
	DBG_MARK(7);
	// upper part of the column-major 128 x 128 inverse (the buffer is zeroed once when it is allocated)
	#pragma unroll 8
	for(int idx = tid; idx < CH_NB * CH_NB; idx += PT) {
		const int c = idx >> 7, r = idx & 127;
		if(r <= c)
			Rinv_out[idx] = TC(r, c);
	}
	DBG_MARK(8);
#undef TC
#undef DBG_MARK
#undef TR
}

__global__ void __launch_bounds__(PT, 1) k_potrf128(double *__restrict__ A, size_t ld, size_t k0,
	double *__restrict__ Rinv_out, int *__restrict__ info, long long *__restrict__ dbg)
{
	potrf128_block(A + k0 * ld + k0, ld, Rinv_out, info, int(k0) + 1, dbg, PotrfNoHook());
}

static const size_t POTRF_SMEM = (size_t)(P3_LD * CH_NB + CH_NB) * sizeof(double);

