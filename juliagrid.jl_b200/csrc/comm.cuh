// The one collective of the scenario-sharded runs: an all-gather of the converged states of every rank's block of
// contingencies / Monte-Carlo draws over NCCL (NVLink 5 / NVSwitch), SURVEY.md §8(b)/(e). The reference has no
// distributed code (the user loop over scenarios is serial); this stands in for "collect the results
// of the loop". NCCL is resolved at run time (dlopen of libnccl.so.2 — the copy already mapped by the host process, the
// system one, or the path in JGB200_NCCL_LIB), so single-GPU users need no NCCL at all and the library links nothing
// but the CUDA runtime.
#pragma once
#include "common.cuh"

namespace jgb {

class CommContext {
  public:
    CommContext(cudaStream_t main_stream) : main(main_stream) {}
    ~CommContext();
    static void unique_id(unsigned char id[128]);
    void init(int rank, int nranks, const unsigned char id[128]);
    // All-gather rows_local rows of each array into [nranks * rows_local] rows, rank-major. Runs on a private stream
    // after the work already enqueued on the context's stream, so the caller can go on with the next batch; the send and
    // receive buffers must stay untouched until wait() (or the next allgather_states, which waits first).
    void allgather_states(int64_t rows_local, int64_t n, const double* vm, const double* va, const int32_t* iters,
                          const int8_t* status, double* vm_all, double* va_all, int32_t* iters_all, int8_t* status_all);
    void wait(bool host_blocking);
    int rank = -1, nranks = 0;
    long long calls = 0;

  private:
    cudaStream_t main = nullptr, cstream = nullptr;
    cudaEvent_t ready = nullptr, done = nullptr;
    void* comm = nullptr;
    bool pending = false;
};

}  // namespace jgb
