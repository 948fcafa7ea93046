// NCCL all-gather of converged states. See comm.cuh.
#include "comm.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

namespace jgb {
namespace {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

NcclApi& api() {
    static NcclApi a;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {getenv("JGB200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            if (!nm || !*nm) continue;
            a.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (a.lib) break;
            a.error = dlerror();
        }
        if (!a.lib) return;
        auto sym = [&](const char* s) {
            void* p = dlsym(a.lib, s);
            if (!p) a.error = std::string("missing NCCL symbol ") + s;
            return p;
        };
        a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
        a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
        a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
        a.AllGather = reinterpret_cast<decltype(a.AllGather)>(sym("ncclAllGather"));
        a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(sym("ncclGroupStart"));
        a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(sym("ncclGroupEnd"));
        a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
        if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllGather || !a.GroupStart || !a.GroupEnd ||
            !a.GetErrorString) {
            dlclose(a.lib);
            a.lib = nullptr;
        }
    });
    if (!a.lib) throw std::logic_error("NCCL is not available (libnccl.so.2 could not be loaded: " + a.error + ")");
    return a;
}

void nccl_check(ncclResult_t r, const char* what) {
    if (r != ncclSuccess) throw CudaError(std::string(what) + ": " + api().GetErrorString(r));
}

static_assert(sizeof(ncclUniqueId) == 128, "jgb_comm_* pass the NCCL unique id as 128 bytes");

}  // namespace

void CommContext::unique_id(unsigned char id[128]) {
    ncclUniqueId u;
    nccl_check(api().GetUniqueId(&u), "ncclGetUniqueId");
    std::memcpy(id, &u, 128);
}

void CommContext::init(int rank_, int nranks_, const unsigned char id[128]) {
    if (comm) throw std::logic_error("comm_init: the context already has a communicator");
    if (nranks_ < 1 || rank_ < 0 || rank_ >= nranks_ || !id) throw std::invalid_argument("comm_init: bad rank / nranks / id");
    ncclUniqueId u;
    std::memcpy(&u, id, 128);
    ncclComm_t c = nullptr;
    nccl_check(api().CommInitRank(&c, nranks_, u, rank_), "ncclCommInitRank");
    comm = c;
    rank = rank_;
    nranks = nranks_;
    JGB_CUDA(cudaStreamCreateWithFlags(&cstream, cudaStreamNonBlocking));
    JGB_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    JGB_CUDA(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
}

CommContext::~CommContext() {
    if (cstream) cudaStreamSynchronize(cstream);
    if (comm) api().CommDestroy(static_cast<ncclComm_t>(comm));
    if (ready) cudaEventDestroy(ready);
    if (done) cudaEventDestroy(done);
    if (cstream) cudaStreamDestroy(cstream);
}

void CommContext::allgather_states(int64_t rows, int64_t n, const double* vm, const double* va, const int32_t* iters,
                                   const int8_t* status, double* vm_all, double* va_all, int32_t* iters_all,
                                   int8_t* status_all) {
    if (!comm) throw std::logic_error("allgather_states: jgb_comm_init has not been called");
    if (rows <= 0 || n <= 0) throw std::invalid_argument("allgather_states: rows and n must be positive");
    if ((vm && !vm_all) || (va && !va_all) || (iters && !iters_all) || (status && !status_all))
        throw std::invalid_argument("allgather_states: a send buffer without its receive buffer");
    wait(false);
    // the collective starts once everything enqueued so far on the context's stream (the batch that produced the
    // states) has finished, and runs beside whatever the caller enqueues next
    JGB_CUDA(cudaEventRecord(ready, main));
    JGB_CUDA(cudaStreamWaitEvent(cstream, ready, 0));
    ncclComm_t c = static_cast<ncclComm_t>(comm);
    NcclApi& a = api();
    nccl_check(a.GroupStart(), "ncclGroupStart");        // one fused NCCL launch for the four arrays
    if (vm) nccl_check(a.AllGather(vm, vm_all, (size_t)(rows * n), ncclFloat64, c, cstream), "ncclAllGather");
    if (va) nccl_check(a.AllGather(va, va_all, (size_t)(rows * n), ncclFloat64, c, cstream), "ncclAllGather");
    if (iters) nccl_check(a.AllGather(iters, iters_all, (size_t)rows, ncclInt32, c, cstream), "ncclAllGather");
    if (status) nccl_check(a.AllGather(status, status_all, (size_t)rows, ncclInt8, c, cstream), "ncclAllGather");
    nccl_check(a.GroupEnd(), "ncclGroupEnd");
    JGB_CUDA(cudaEventRecord(done, cstream));
    pending = true;
    ++calls;
}

void CommContext::wait(bool host_blocking) {
    if (!pending) return;
    if (host_blocking) JGB_CUDA(cudaEventSynchronize(done));
    else JGB_CUDA(cudaStreamWaitEvent(main, done, 0));
    pending = false;
}

}  // namespace jgb
