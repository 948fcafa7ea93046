// Constant-matrix linear solves: factor a sparse symmetric matrix once, then solve for blocks of right-hand sides.
// Stands in for the `factorization / factorization! / solution!` triad of the reference (src/backend/utility.jl:470-586)
// as used by its linear analyses: DC power flow (`solve!`, src/powerFlow/dcPowerFlow.jl:93-134), DC state estimation
// (`solve!`, src/stateEstimation/dcStateEstimation.jl:342-371) and PMU state estimation
// (`solve!`, src/stateEstimation/pmuStateEstimation.jl:369-399). In the two estimators the right-hand side is
// b = H' W z; the projection W H is constant, so it is kept on the device and applied to a whole block of
// measurement vectors (one per Monte-Carlo draw).
#pragma once
#include "common.cuh"
#include "solver.cuh"

namespace jgb {

class LinContext {
  public:
    explicit LinContext(cudaStream_t st) : stream(st) {}
    // A: n x n, symmetric pattern, full CSC (both triangles), 1-based Int64 indices as in SparseMatrixCSC. Symmetric
    // values take the LDL^T path; unsymmetric values (fast Newton-Raphson B' with phase shifters) are factored twice, A
    // and A', on the same elimination tree.
    // skip (1-based, 0 = none): that row and column are replaced by the identity, like the reference's slack fix
    // (`removeRowColumn` + `A[slack, slack] = 1`, dcPowerFlow.jl:113-114; `gain[slack, slack] = 1`,
    // dcStateEstimation.jl:354), so x[skip] = b[skip].
    void setup(int64_t n, const int64_t* colptr, const int64_t* rowval, const double* nzval, int64_t skip);
    void refactor(const double* nzval);            // same pattern, new values (`factorization!`)
    // P = W H as CSC (m x n, 1-based): b = P' z
    void set_projection(int64_t m, const int64_t* colptr, const int64_t* rowval, const double* nzval);
    // in: [R][n] right-hand sides (projected = false) or [R][m] measurement vectors (projected = true), each vector
    // contiguous; out: [R][n]. dev_* select host or device pointers.
    void solve(int64_t R, const double* in, bool dev_in, double* out, bool dev_out, bool projected);
    void solve_block(double* B, int Rp);      // device block [n][Rp] (right-hand-side minor), solved in place
    bool symmetric = true;
    int64_t n = 0, m = 0;
    int64_t nnz_factor() const { return solver.sym.nnz_lu; }
    int64_t nfronts() const { return solver.sym.nfronts; }

  private:
    void factor_now();
    bool values_symmetric(const double* av) const;
    cudaStream_t stream;
    int skip = -1;
    MfSolver solver, solver_t;         // solver_t: the transposed matrix (unsymmetric values only)
    std::vector<int64_t> tpos;         // input nonzero -> input position of its transpose partner
    std::vector<double> vals_t;
    std::vector<int> slot;             // input nonzero -> position in the analysed pattern, -1 = dropped
    std::vector<double> vals;          // values in the analysed pattern's order
    int skip_diag = -1;
    size_t nnz_in = 0;
    DevBuf<double> d_aval, d_aval_t, d_zero, d_x0, d_B, d_Z, d_in, d_out, d_pval;
    DevBuf<int> d_status, d_pcolptr, d_prow;
    PinnedBuf<int> h_status;
};

}  // namespace jgb
