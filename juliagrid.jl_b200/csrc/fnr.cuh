// Fast Newton-Raphson (BX / XB) AC power flow: constant B' and B'' factored once on the device, then every iteration
// is two mismatch sweeps over the Ybus strips and two solves with the stored factors. Stands in for
// src/powerFlow/acPowerFlow.jl:686-727 (mismatch!), :913-983 (solve!), :1389-1433 (powerFlow! loop) of the
// reference for AcPowerFlow{FastNewtonRaphson}. Because the matrices do not depend on the injections, a block of
// injection scenarios (load / generation samples on one topology) shares the two factorisations: state and
// mismatches are scenario minor, [bus][R], and the solves are blocks of R right-hand sides (lin.cuh).
#pragma once
#include "common.cuh"
#include "lin.cuh"

namespace jgb {

struct FnrDev {
    int n, slack, npq, R;
    const int* ycolptr;
    const int* yrow;
    const double2* yt;        // Y[col, row] (nodalMatrixTranspose.nzval)
    const signed char* type;
    const int* pq;            // 0-based position among PQ buses, -1 otherwise
    const int* pvpq;          // 0-based position among non-slack buses, -1 for the slack
    const double* pinj;       // [n][R] supply - demand, active
    const double* qinj;       // [n][R]
    double* vm;               // [n][R]
    double* va;
    double* mp;               // [n-1][R] active mismatch, overwritten by the angle increments
    double* mq;               // [npq][R]
    unsigned long long* stopbits;   // [2][R]
    double* stop;             // [2][R]
    unsigned char* active;    // [R]
    int* status;              // [R]
    int* iters;               // [R]
    int* remaining;
};

class FnrContext {
  public:
    explicit FnrContext(cudaStream_t st) : stream(st), active_lin(st), reactive_lin(st) {}
    // Ybus pattern + transpose values as in jgb_nr_setup; bp / bq: the reference's active / reactive Jacobians
    // (SparseMatrixCSC, 1-based) built by fastNewtonRaphsonBX / XB.
    void setup(int64_t n, const int64_t* ycp, const int64_t* yrv, const double* yt, const int8_t* type, int64_t slack,
               const int64_t* bpcp, const int64_t* bprv, const double* bpnz, const int64_t* bqcp, const int64_t* bqrv,
               const double* bqnz);
    void set_injection(const double* ps, const double* qs, const double* pd, const double* qd);
    void set_state(const double* vm, const double* va);
    void get_state(double* vm, double* va);
    void mismatch(double* sp, double* sq);
    void solve();
    int run(int64_t max_iter, double tol, int64_t* iters, double* sp, double* sq);
    // R injection scenarios: pinj / qinj [R][n] = supply - demand per bus; every scenario starts from the state given
    // by set_state; outputs [R][n], iterations and status (0 converged, 1 iteration cap) per scenario.
    int batch(int64_t R, const double* pinj, const double* qinj, int64_t max_iter, double tol, double* vm_out,
              double* va_out, int32_t* iters, int8_t* status, int64_t* total);
    int n = 0, npq = 0;
    long long launches = 0;

  private:
    void alloc(int Rp);
    void broadcast_state();
    FnrDev view();
    void sweep(bool q_only);
    void step();
    cudaStream_t stream;
    LinContext active_lin, reactive_lin;
    int slack = -1, R = 0;
    bool have_injection = false, have_state = false;
    int64_t iteration = 0;
    std::vector<double> h_vm, h_va;      // start point (single state given by set_state)
    DevBuf<int> d_ycolptr, d_yrow, d_pq, d_pvpq, d_status, d_iters, d_remaining;
    DevBuf<double2> d_yt;
    DevBuf<signed char> d_type;
    DevBuf<double> d_pinj, d_qinj, d_vm, d_va, d_mp, d_mq, d_stop, d_io, d_io2;
    DevBuf<unsigned long long> d_stopbits;
    DevBuf<unsigned char> d_active;
    PinnedBuf<double> h_stop;
    PinnedBuf<int> h_int;
};

}  // namespace jgb
