// Gauss-Newton WLS state-estimation context: measurement-row kernel (21 codes), gain build on the fixed
// pattern of H'WH, multifrontal solve, state update. Stands in for src/stateEstimation/acStateEstimation.jl
// :261-583 (normalEquation!), :878-904 (increment!), :1035-1047 (solve!), :1286-1329 (stateEstimation!).
#pragma once
#include <map>
#include <utility>

#include "common.cuh"
#include "solver.cuh"

namespace jgb {

struct WlsDev {
    int n, m, slack, nnzh, nnzg, nv;
    // Ybus
    const int* ycolptr;
    const int* yrow;
    const double2* y;
    const double2* yt;
    const int* ydiag;          // position of the diagonal entry of every column
    // branches
    const int *br_from, *br_to;
    const double *br_g, *br_b, *br_gsi, *br_bsi, *br_tinv, *br_phi;
    // rows
    const signed char* type;
    const int* index;
    const int* slotptr;        // [m+1]
    const int* slotpos;        // CSC positions of the row's H entries in semantic order
    const double* wdiag;       // [m]
    const double* woff;        // [m] W[row,row-1] for the Im row of a correlated pair, else 0
    // H in CSC (reference order)
    const int* hcolptr;
    const int* hrow;
    // gain gather lists (lower triangle incl. diagonal)
    const int* gentry_ptr;     // [nlow+1]
    const int* gterm_a;        // H position of the (row, a) factor
    const int* gterm_b;
    const int* gterm_w;        // index into wall: [0,m) diagonal, [m,2m) off-diagonal of pair ending at row
    const int* glow_pos;       // position of entry (a,b), a >= b, in G's CSC values
    const int* gup_pos;        // position of the mirrored entry (b,a)
    int nlow;
    int gslack_pos;            // position of G[slack,slack]
    // state and per-iteration vectors, scenario minor
    double* vm;
    double* va;
    const double* z;           // means [m][S]
    double* res;               // residual [m][S]
    double* hval;              // [nnzh][S]
    double* gval;              // [nnzg][S]
    double* rhs;               // [2n][S]
    double* inc;               // [2n][S]
    double* objpart;           // [nblocks][S] partial objective sums
    double* obj;               // [S]
    unsigned long long* maxbits;   // [S]
    double* maxinc;            // [S]
    unsigned char* active;
    int* status;
    int* iters;
    int* remaining;
};

class WlsContext {
  public:
    explicit WlsContext(cudaStream_t st) : stream(st) {}
    void setup(int64_t n, int64_t m, int64_t slack, const int64_t* hcp, const int64_t* hrv, const int8_t* type,
               const int64_t* index, const int64_t* range6, const int64_t* wcp, const int64_t* wrv, const double* wnz,
               const int64_t* ycp, const int64_t* yrv, const double* y, const double* yt, int64_t nbr,
               const int64_t* from, const int64_t* to, const double* cond, const double* susc, const double* tap,
               const double* shift, const double* adm);
    void set_mean(const double* z);
    void set_state(const double* vm, const double* va);
    void get_state(double* vm, double* va);
    void increment(double* max_inc, double* objective);
    void solve();
    void get_vectors(double* res, double* inc, double* hval, double* gval, int64_t* it);
    int run(int64_t max_iter, double tol, int64_t* iters, double* max_inc, double* objective);
    int batch(int64_t S, const double* Z, bool dev_in, int64_t max_iter, double tol, double* vm_out, double* va_out,
              int32_t* iters, int8_t* status, double* objective, bool dev_out, int64_t* total);
    // largest normalised residual (residualTest!, stateEstimation/badData.jl:181-285): the numeric part; the
    // monitoring bookkeeping stays with the host. index is 1-based, 0 when every residual is zero.
    void residual_test(double threshold, double* max_rn, int64_t* index, double* c_out);
    void remove_row(int64_t row);              // 1-based; the row leaves the model (type 0)
    // update*!(analysis; ...) of single measurement rows, value-only Ybus and branch-parameter updates: the gain pattern,
    // its gather lists and the symbolic factorisation are reused (pattern changes throw -> rebuild the context)
    void update_rows(int64_t k, const int64_t* rows, const double* mean, const double* precision,
                     const double* precision_off, const int8_t* type, const int64_t* index);
    void update_y(int64_t k, const int64_t* pos, const double* y, const double* yt);
    void update_branch(int64_t branch, double cond, double susc, double tap, double shift, const double* adm);
    double stat(const std::string& key);

    int n = 0, m = 0, slack = -1, nnzh = 0, nnzg = 0, nbr = 0, nnzy = 0;
    std::vector<int64_t> gcolptr1, growval1;
    long long launches = 0;
    long long nterms = 0;
    PhaseTimer timer;

  private:
    WlsDev view(bool batch);
    void launch_rows(int S, bool batch);
    void launch_gain(int S, bool batch);
    void alloc_batch(int S);

    cudaStream_t stream;
    MfSolver solver;         // single case
    MfSolver solver_batch;   // Monte-Carlo batches
    DevBuf<int> d_ycolptr, d_yrow, d_ydiag, d_br_from, d_br_to, d_index, d_slotptr, d_slotpos, d_hcolptr, d_hrow,
        d_gentry_ptr, d_gterm_a, d_gterm_b, d_gterm_w, d_glow_pos, d_gup_pos;
    DevBuf<double2> d_y, d_yt;
    DevBuf<double> d_br_g, d_br_b, d_br_gsi, d_br_bsi, d_br_tinv, d_br_phi, d_wdiag, d_woff;
    DevBuf<signed char> d_type;
    int nlow = 0, gslack_pos = -1, nrowblocks = 0;
    std::vector<double> h_const;   // constant H entries (codes 1, 12, 13 = status) in CSC order, 0 elsewhere
    // single-case state
    DevBuf<double> d_vm, d_va, d_z, d_res, d_hval, d_gval, d_rhs, d_inc, d_objpart, d_obj, d_maxinc;
    DevBuf<unsigned long long> d_maxbits;
    DevBuf<unsigned char> d_active;
    DevBuf<int> d_status, d_iters, d_remaining;
    // batch state
    int batch_S = 0;
    DevBuf<double> b_vm, b_va, b_z, b_res, b_hval, b_gval, b_rhs, b_inc, b_objpart, b_obj, b_maxinc, b_out, b_zraw, b_hconst;
    DevBuf<unsigned long long> b_maxbits;
    DevBuf<unsigned char> b_active;
    DevBuf<int> b_status, b_iters;
    PinnedBuf<double> h_d;
    PinnedBuf<int> h_i;
    int64_t iteration = 0;
    bool have_mean = false, have_state = false;
    // bad-data lists (built on the first residual_test)
    void build_pairs();
    bool have_pairs = false;
    std::vector<int> h_slotptr, h_slotpos, h_poscol, h_ycolptr, h_yrow, h_brfrom, h_brto, h_index;
    std::vector<int8_t> h_type;
    std::vector<double> h_woff;
    std::vector<std::vector<std::pair<int, int>>> h_rowent;     // per row: (column, CSC position) of its H entries
    std::map<int, std::vector<int>> pending_slots;
    std::vector<double> h_wdiag;
    DevBuf<int> d_pair_ptr, d_pair_pa, d_pair_pb;
    DevBuf<long long> d_pair_z;
    DevBuf<double> d_proj;
};

}  // namespace jgb
