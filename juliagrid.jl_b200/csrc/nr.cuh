// Newton-Raphson AC power-flow context: device-resident Ybus strips, Jacobian pattern, state, and the
// per-iteration kernels. Stands in for src/powerFlow/acPowerFlow.jl:39-175 (setup), :645-685 (mismatch!),
// :793-911 (solve!), :1389-1433 (powerFlow! loop) of the reference.
#pragma once
#include <cstdlib>
#include "common.cuh"
#include "solver.cuh"

namespace jgb {

struct NrDev {
    int n, slack, dim, nnzj;
    int strip_cap;        // entries of the largest staged Ybus strip (batch assembly)
    const int* ycolptr;
    const int* yrow;
    const double2* y;     // Y[row, col]   (nodalMatrix.nzval)
    const double2* yt;    // Y[col, row]   (nodalMatrixTranspose.nzval)
    const signed char* type;
    const int* pq;        // 0-based, -1 when absent
    const int* pvpq;
    const int* pcount;
    const int* jcolptr;
    const double *sup_p, *sup_q, *dem_p, *dem_q;
    double* vm;           // [n][S]
    double* va;
    double* f;            // mismatch [dim][S]
    double* jval;         // [nnzj][S]
    double* inc;          // increment [dim][S]
    unsigned long long* stopbits;   // [2][S] running max as bit patterns of non-negative doubles
    double* stop;         // [2][S]
    unsigned char* active;          // [S]
    int* status;          // [S]
    int* iters;           // [S]
    int* remaining;       // [1]
    // per-scenario outage (batch only; null otherwise)
    const int* out_from;
    const int* out_to;
    const double2* dy;    // [4][S]: ff, ft, tf, tt
};

class NrContext {
  public:
    explicit NrContext(cudaStream_t st) : stream(st) {}
    void setup(int64_t n, const int64_t* ycp, const int64_t* yrv, const double* y, const double* yt,
               const int8_t* type, int64_t slack);
    void set_injection(const double* ps, const double* qs, const double* pd, const double* qd);
    void set_state(const double* vm, const double* va);
    void get_state(double* vm, double* va);
    void update_y(int64_t k, const int64_t* pos, const double* y, const double* yt);
    void mismatch(double* sp, double* sq);
    void solve();
    void get_vectors(double* f, double* inc, double* jv, int64_t* it);
    int run(int64_t max_iter, double tol, int64_t* iters, double* sp, double* sq);
    int batch(int64_t S, const int64_t* of, const int64_t* ot, const double* dy, bool dev_in, int64_t max_iter,
              double tol, double* vm_out, double* va_out, int32_t* iters, int8_t* status, bool dev_out,
              int64_t* total);
    void set_branches(int64_t nbr, const int64_t* from, const int64_t* to, const double* yff, const double* yft,
                      const double* ytf, const double* ytt, const int8_t* status);
    void power(double* out[10]);
    double stat(const std::string& key);

    // host mirrors (the reference's own index arrays, 1-based)
    int n = 0, slack = -1, dim = 0, nnzj = 0, nnzy = 0, strip_cap = 8;
    std::vector<int64_t> pq1, pvpq1, pcount1, jcolptr1, jrowval1;
    long long launches = 0;
    PhaseTimer timer;
    GraphSlot graph_head, graph_iter;   // single-case powerFlow! loop as two CUDA graphs

  private:
    void alloc_state(int S);
    void launch_assemble(int S, bool batch);
    // Pivot guard + one step of iterative refinement (the factorisation has no pivoting; UMFPACK / KLU in the reference
    // pivot by threshold): scenarios whose elimination met a multiplier above `pivot_growth` get delta += J^-1 (f - J delta)
    // with the residual formed from the assembled Jacobian values. Returns the number of scenarios refined.
    int refine_weak(MfSolver& sol, int S, double* jval, double* f, double* inc, const unsigned char* active, int* status,
                    int* weak);
    double pivot_growth = getenv("JGB_PIVOT_GROWTH") ? atof(getenv("JGB_PIVOT_GROWTH")) : 1e6;
    long long weak_events = 0, refine_calls = 0;
    DevBuf<int> d_jrow, d_weak, b_weak, d_weakcnt;
    DevBuf<double> r_res, r_inc;
    DevBuf<unsigned char> r_mask;
    NrDev view(int S, bool batch);

    cudaStream_t stream;
    MfSolver solver;         // single case (latency-oriented amalgamation)
    MfSolver solver_batch;   // scenario batches (throughput-oriented amalgamation)
    DevBuf<int> d_ycolptr, d_yrow, d_pq, d_pvpq, d_pcount, d_jcolptr;
    DevBuf<double2> d_y, d_yt;
    DevBuf<signed char> d_type, d_brstatus;
    DevBuf<int> d_brfrom, d_brto;
    DevBuf<double2> d_yff, d_yft, d_ytf, d_ytt;
    DevBuf<double> d_pw;
    int nbr = 0;
    DevBuf<double> d_sup_p, d_sup_q, d_dem_p, d_dem_q;
    // single-case state (S = 1)
    DevBuf<double> d_vm, d_va, d_f, d_jval, d_inc, d_stop;
    DevBuf<unsigned long long> d_stopbits;
    DevBuf<unsigned char> d_active;
    DevBuf<int> d_status, d_iters, d_remaining;
    // batch state
    int batch_S = 0;
    DevBuf<double> b_vm, b_va, b_f, b_jval, b_inc, b_stop, b_out;
    DevBuf<unsigned long long> b_stopbits;
    DevBuf<unsigned char> b_active;
    DevBuf<int> b_status, b_iters, b_of, b_ot;
    DevBuf<double2> b_dy;
    DevBuf<int64_t> b_of64, b_ot64;
    DevBuf<double> b_dyraw;
    PinnedBuf<double> h_stop;
    PinnedBuf<int> h_int;
    int64_t iteration = 0;
    bool graphs_disabled = getenv("JGB_NO_GRAPH") != nullptr;
    bool staged_assembly = getenv("JGB_STAGED_ASSEMBLY") != nullptr;
    bool jac_valid = false;
    bool have_injection = false, have_state = false;
};

}  // namespace jgb
