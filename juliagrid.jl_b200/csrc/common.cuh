// Shared helpers for the jgb200 CUDA translation units.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

namespace jgb {

struct CudaError : std::runtime_error {
    explicit CudaError(const std::string& m) : std::runtime_error(m) {}
};

#define JGB_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (call);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            throw jgb::CudaError(std::string(#call) + " failed: " + cudaGetErrorString(_e) + " at " + \
                                 __FILE__ + ":" + std::to_string(__LINE__));                        \
    } while (0)

// Owning device buffer (cudaMalloc / cudaFree), movable.
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    void alloc(size_t count) {
        if (count <= n && p) return;
        release();
        if (count == 0) count = 1;
        JGB_CUDA(cudaMalloc((void**)&p, count * sizeof(T)));
        n = count;
    }
    void upload(const std::vector<T>& h, cudaStream_t st = 0) {
        alloc(h.size());
        if (!h.empty()) JGB_CUDA(cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, st));
    }
    void upload(const T* h, size_t count, cudaStream_t st = 0) {
        alloc(count);
        if (count) JGB_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, st));
    }
    void download(T* h, size_t count, cudaStream_t st = 0) const {
        if (count) JGB_CUDA(cudaMemcpyAsync(h, p, count * sizeof(T), cudaMemcpyDeviceToHost, st));
    }
    void zero(cudaStream_t st = 0) {
        if (p) JGB_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), st));
    }
};

// Pinned host buffer for small per-iteration read-backs.
template <typename T>
struct PinnedBuf {
    T* p = nullptr;
    size_t n = 0;
    ~PinnedBuf() { if (p) cudaFreeHost(p); }
    void alloc(size_t count) {
        if (count <= n && p) return;
        if (p) cudaFreeHost(p);
        JGB_CUDA(cudaMallocHost((void**)&p, count * sizeof(T)));
        n = count;
    }
};

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace jgb
