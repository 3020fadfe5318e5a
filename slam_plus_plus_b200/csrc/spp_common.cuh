// spp_common.cuh -- shared host/device helpers of libspp_b200 (sm_100a only, FP64).
#pragma once

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
template <class T>
class DBuf {
	T *m_p;
	size_t m_n, m_cap;
public:
	DBuf() : m_p(0), m_n(0), m_cap(0) {}
	~DBuf() { if(m_p) cudaFree(m_p); }
	DBuf(const DBuf&) = delete;
	DBuf &operator =(const DBuf&) = delete;
	void resize(size_t n)
	{
		if(n > m_cap) {
			size_t want = (m_cap && n < m_cap + m_cap / 2)? m_cap + m_cap / 2 : n;
			if(m_p) cudaFree(m_p);
			m_p = 0; m_cap = 0;
			if(cudaMalloc((void**)&m_p, (want ? want : 1) * sizeof(T)) != cudaSuccess) { // no room for the slack: exact size
				(void)cudaGetLastError();
				m_p = 0;
				want = n;
				SPP_CUDA(cudaMalloc((void**)&m_p, (want ? want : 1) * sizeof(T)));
			}
			m_cap = want;
		}
		m_n = n;
	}
	void release() { if(m_p) cudaFree(m_p); m_p = 0; m_n = m_cap = 0; }
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
