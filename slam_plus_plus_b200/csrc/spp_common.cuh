// spp_common.cuh -- shared host/device helpers of libspp_b200 (sm_100a only, FP64).
#pragma once
#include <chrono>
#include <stdlib.h>

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include <stdexcept>
#include <new>

#include "../../include/spp_b200.h"

namespace spp {

struct cuda_error : std::runtime_error {
	explicit cuda_error(const std::string &s) : std::runtime_error(s) {}
};

struct invalid_error : std::runtime_error {
	explicit invalid_error(const std::string &s) : std::runtime_error(s) {}
};

#define SPP_CUDA(call) do { \
		cudaError_t e_ = (call); \
		if(e_ != cudaSuccess) { \
			char b_[512]; \
			snprintf(b_, sizeof(b_), "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
			if(e_ == cudaErrorMemoryAllocation) throw std::bad_alloc(); \
			throw spp::cuda_error(b_); \
		} \
	} while(0)

// device buffer owned by a context; grows, never shrinks. The first allocation is exact (batch use: no slack); a
// buffer that has to grow again grows by at least half (a system that is extended step by step -- incremental bundle
// adjustment -- does not reallocate at every step)
// allocation statistics of all device buffers (SPP_ALLOC_STATS: printed by spp_destroy)
struct DBufStats {
	static DBufStats &get() { static DBufStats s; return s; }
	size_t n_malloc, n_free, bytes;
	double ms;
	DBufStats() : n_malloc(0), n_free(0), bytes(0), ms(0) {}
};
// Device buffers come from the device's default stream-ordered memory pool (release threshold raised by spp_create, so
// that freed blocks stay cached): a system that grows step by step -- incremental bundle adjustment -- reallocates each
// of ~60 buffers a dozen times, and plain cudaMalloc / cudaFree were measured at 2.5 - 5.2 s for the 847 + 740 calls of
// the 88-marker Venice run (3 - 6 ms per call, most of the host time of the run). The semantics of the plain calls are
// kept: an allocation is usable by every stream when dbuf_malloc returns, and a block is only freed when the device is
// idle (cudaFree synchronised implicitly; buffers are read by side streams the buffer does not know about).
// SPP_NO_MEMPOOL=1: plain cudaMalloc / cudaFree.
inline bool dbuf_use_pool()
{
	static const bool b_pool = getenv("SPP_NO_MEMPOOL") == 0;
	return b_pool;
}
inline cudaError_t dbuf_malloc(void **pp, size_t bytes)
{
	const auto t0 = std::chrono::steady_clock::now();
	cudaError_t e;
	if(dbuf_use_pool()) {
		e = cudaMallocAsync(pp, bytes, (cudaStream_t)0);
		if(e == cudaSuccess)
			e = cudaStreamSynchronize((cudaStream_t)0);
	} else
		e = cudaMalloc(pp, bytes);
	DBufStats &st = DBufStats::get();
	st.ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
	++ st.n_malloc; st.bytes += bytes;
	return e;
}
inline void dbuf_free(void *p)
{
	const auto t0 = std::chrono::steady_clock::now();
	if(dbuf_use_pool()) {
		cudaDeviceSynchronize();
		cudaFreeAsync(p, (cudaStream_t)0);
	} else
		cudaFree(p);
	DBufStats &st = DBufStats::get();
	st.ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
	++ st.n_free;
}

template <class T>
class DBuf {
	T *m_p;
	size_t m_n, m_cap;
public:
	DBuf() : m_p(0), m_n(0), m_cap(0) {}
	~DBuf() { if(m_p) dbuf_free(m_p); }
	DBuf(const DBuf&) = delete;
	DBuf &operator =(const DBuf&) = delete;
	void resize(size_t n)
	{
		if(n > m_cap) {
			size_t want = (m_cap && n < m_cap + m_cap / 2)? m_cap + m_cap / 2 : n;
			if(m_p) dbuf_free(m_p);
			m_p = 0; m_cap = 0;
			if(dbuf_malloc((void**)&m_p, (want ? want : 1) * sizeof(T)) != cudaSuccess) { // no room for the slack: exact size
				(void)cudaGetLastError();
				m_p = 0;
				want = n;
				SPP_CUDA(dbuf_malloc((void**)&m_p, (want ? want : 1) * sizeof(T)));
			}
			m_cap = want;
		}
		m_n = n;
	}
	// resize that keeps the first min(old size, n) elements (an append-only system grows; the copy is ordered on s, the
	// old block is freed once it is done)
	void grow_keep(size_t n, cudaStream_t s)
	{
		if(n > m_cap) {
			const size_t want = (m_cap && n < m_cap + m_cap / 2)? m_cap + m_cap / 2 : n;
			T *p_new = 0;
			SPP_CUDA(dbuf_malloc((void**)&p_new, (want ? want : 1) * sizeof(T)));
			if(m_p && m_n) {
				if(cudaMemcpyAsync(p_new, m_p, m_n * sizeof(T), cudaMemcpyDeviceToDevice, s) != cudaSuccess ||
				   cudaStreamSynchronize(s) != cudaSuccess) {
					dbuf_free(p_new);
					SPP_CUDA(cudaGetLastError());
					throw spp::cuda_error("grow_keep: copy failed");
				}
			}
			if(m_p) dbuf_free(m_p);
			m_p = p_new;
			m_cap = want;
		}
		m_n = n;
	}
	void release() { if(m_p) dbuf_free(m_p); m_p = 0; m_n = m_cap = 0; }
	T *p() { return m_p; }
	const T *p() const { return m_p; }
	size_t size() const { return m_n; }
	void upload(const T *h, size_t n, cudaStream_t s)
	{
		resize(n);
		if(n) SPP_CUDA(cudaMemcpyAsync(m_p, h, n * sizeof(T), cudaMemcpyHostToDevice, s));
	}
	void upload(const std::vector<T> &h, cudaStream_t s) { upload(h.data(), h.size(), s); }
	void download(T *h, size_t n, cudaStream_t s) const
	{
		if(n) SPP_CUDA(cudaMemcpyAsync(h, m_p, n * sizeof(T), cudaMemcpyDeviceToHost, s));
	}
	void zero(cudaStream_t s) { if(m_n) SPP_CUDA(cudaMemsetAsync(m_p, 0, m_n * sizeof(T), s)); }
};

// pinned host scalar block for small device->host results
template <class T>
class HPinned {
	T *m_p;
	size_t m_n;
public:
	HPinned() : m_p(0), m_n(0) {}
	~HPinned() { if(m_p) cudaFreeHost(m_p); }
	void resize(size_t n)
	{
		if(n > m_n) {
			if(m_p) cudaFreeHost(m_p);
			m_p = 0;
			SPP_CUDA(cudaMallocHost((void**)&m_p, n * sizeof(T)));
			m_n = n;
		}
	}
	T *p() { return m_p; }
	T &operator [](size_t i) { return m_p[i]; }
};

struct PhaseTimer { // CUDA-event stopwatch on the context stream
	cudaEvent_t a, b;
	cudaStream_t s;
	bool armed;
	PhaseTimer() : a(0), b(0), s(0), armed(false) {}
	void init(cudaStream_t stream)
	{
		s = stream;
		SPP_CUDA(cudaEventCreate(&a));
		SPP_CUDA(cudaEventCreate(&b));
	}
	void destroy() { if(a) cudaEventDestroy(a); if(b) cudaEventDestroy(b); a = b = 0; }
};

static inline unsigned n_blocks(size_t n, unsigned bs) { return (unsigned)((n + bs - 1) / bs); }

} // namespace spp
