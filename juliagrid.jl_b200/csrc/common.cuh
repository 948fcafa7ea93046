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

// CUDA-event phase timer on the context's stream (enabled by jgb_profile): marks are recorded while the work is
// enqueued and resolved after the next stream synchronisation; accumulates milliseconds per phase name.
struct PhaseTimer {
    bool enabled = false;
    std::vector<cudaEvent_t> pool;
    size_t used = 0;
    struct Span { int phase; size_t a, b; };
    std::vector<Span> spans;
    double ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long count[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    ~PhaseTimer() { for (auto e : pool) cudaEventDestroy(e); }
    cudaEvent_t mark(cudaStream_t st) {
        if (!enabled) return nullptr;
        if (used == pool.size()) {
            cudaEvent_t e;
            JGB_CUDA(cudaEventCreate(&e));
            pool.push_back(e);
        }
        cudaEvent_t e = pool[used++];
        JGB_CUDA(cudaEventRecord(e, st));
        return e;
    }
    cudaEvent_t reserve() {     // an event some other code records
        if (!enabled) return nullptr;
        if (used == pool.size()) {
            cudaEvent_t e;
            JGB_CUDA(cudaEventCreate(&e));
            pool.push_back(e);
        }
        return pool[used++];
    }
    size_t last() const { return used - 1; }
    void span(int phase, size_t a, size_t b) { if (enabled) spans.push_back({phase, a, b}); }
    void resolve() {            // call after the stream has been synchronised
        if (!enabled) return;
        for (auto& sp : spans) {
            float t = 0;
            if (cudaEventElapsedTime(&t, pool[sp.a], pool[sp.b]) == cudaSuccess) { ms[sp.phase] += t; count[sp.phase]++; }
        }
        spans.clear();
        used = 0;
    }
    void reset() { for (int i = 0; i < 8; ++i) { ms[i] = 0; count[i] = 0; } spans.clear(); used = 0; }
};
// One captured + instantiated CUDA graph, rebuilt when its key (scalar kernel arguments baked in) changes.
struct GraphSlot {
    cudaGraphExec_t exec = nullptr;
    double key_a = 0;
    long long key_b = -1;
    ~GraphSlot() { if (exec) cudaGraphExecDestroy(exec); }
    bool valid(double a, long long b) const { return exec && key_a == a && key_b == b; }
    template <typename F>
    void capture(cudaStream_t st, double a, long long b, F&& enqueue) {
        if (exec) { cudaGraphExecDestroy(exec); exec = nullptr; }
        cudaGraph_t g = nullptr;
        JGB_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        try {
            enqueue();
        } catch (...) {
            cudaStreamEndCapture(st, &g);
            if (g) cudaGraphDestroy(g);
            throw;
        }
        JGB_CUDA(cudaStreamEndCapture(st, &g));
        cudaError_t e = cudaGraphInstantiate(&exec, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) { exec = nullptr; throw CudaError(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e)); }
        key_a = a;
        key_b = b;
    }
    void launch(cudaStream_t st) { JGB_CUDA(cudaGraphLaunch(exec, st)); }
};

enum Phase { kPhAssemble = 0, kPhFactor = 1, kPhBacksolve = 2, kPhUpdate = 3, kPhGain = 4, kPhRows = 5 };

}  // namespace jgb
