// Fixed-pattern multifrontal LU (no pivoting, right-hand side carried as an extra front column) on the GPU.
//
// Replaces the numeric half of the reference's `factorization!` + `solution!`
// (src/backend/utility.jl:478-500, 542-548, 576-586 -> UMFPACK / KLU / CHOLMOD in the reference).
// One instance owns the device copy of a `Symbolic` and the factor workspaces for up to S scenarios;
// values are laid out scenario-minor: element e of scenario s lives at [e * S + s].
#pragma once
#include <vector>
#include "common.cuh"
#include "symbolic.hpp"
#include "tasks.hpp"

namespace jgb {

struct ChildDesc;
struct DevSym {
    const int *f_k, *f_nf, *f_rowptr, *f_rows, *f_relptr, *f_rel, *f_childptr, *f_children, *f_asmptr, *asm_src,
        *asm_dst, *f_eaptr, *ea_roundptr;
    const int2* ea_pair;
    const int* ea_roundptr_s;       // symmetric (packed lower triangle) gather lists, see symbolic.hpp
    const int2* ea_pair_s;
    const long long *f_uoff, *f_updoff;
    long long upd_size;
    // Update storage: one section per tile width W = 1, 2, 4, 8, 16, 32 (index log2 W). The block a front leaves for its
    // parent lives in the section of the PARENT's scenario-tile width, so a CTA of TS scenarios reads its children and
    // writes its own block as contiguous runs: element e of scenario s of a block at offset `off` of section W sits at
    //   sec_base[lw] + ((s / W) * sec_size[lw] + off + e) * W + s % W          (sec_size in elements per scenario)
    long long sec_base[6], sec_size[6];
    // The packed U rows are sectioned the same way, by the tile width of the back-solve launch that READS them:
    // entry e of scenario s of a front whose rows start at f_uoff sits at usec_base[lw] + ((s / W) * usec_size[lw] + f_uoff + e) * W + s % W
    long long usec_base[6], usec_size[6];
    const struct ChildDesc* child_desc;
    // pivot guard (LU without pivoting): a multiplier |F[i,p] / F[p,p]| above `growth` marks the scenario in weak[] (nullable);
    // the caller then refines the solution of that scenario with one residual step
    int* weak;
    double growth;
};

// Per-front descriptor, laid out in launch order (one 64-byte read replaces the fronts[] -> f_* double indirection)
struct __align__(16) FrontDesc {
    int f, nf, k, rowptr;
    int asm0, asm1, child0, child1;
    int ea0, ea1;          // rounds of the extend-add gather (symmetric lists when the solver is symmetric)
    int flags, wout;       // flags bit 0: the parent reads only the lower triangle + rhs of this front's update block,
                           // bits 8..15: tile width of the section this front's packed U rows are written to (the
                           // back-solve launch's TS); wout: tile width of the section its update block is written to
                           // (the parent's TS)
    long long uoff, updoff;
};

// Per-child descriptor in f_children order: what a parent needs to fetch and scatter a child's update block
struct __align__(16) ChildDesc {
    int uc, relptr;
    long long updoff;
};

// Staged extend-add of the scenario-tile LU kernel (see mf_factor_kernel): chunk list (x = element offset inside the
// tile's section, y = elements), destination list in update-storage order, ring geometry
struct StagedEa {
    const int2* chunks = nullptr;
    const int* upd_dst = nullptr;
    int sec_cum = 0;        // elements of the sections below this launch's tile width (index base into upd_dst)
    int ring_elems = 0;     // elements per ring stage
    int ring_off = 0;       // doubles from the start of dynamic shared memory to the ring
};

struct FactorLaunch {
    int begin, count;      // range in level_fronts
    int ts;                // scenarios per CTA
    int threads;
    int tr;                // row lanes (power of two), column lanes = threads / ts / tr
    size_t smem;
    bool global_front;     // front kept in a global workspace instead of shared memory
    bool bulk;             // TMA-staged small-front kernel (batch only)
    bool sym;              // packed symmetric (LDL^T) kernel
    bool dense;            // one-scenario-per-CTA LDL^T kernel with the tensor-core trailing update (mf_dense.cuh)
    bool dense_lu;         // its LU twin for unsymmetric values (mf_factor_dense_lu_kernel)
    // single case: a CTA walks a chain of fronts (each the parent of the one before) without leaving the kernel;
    // launches follow dependency slots instead of tree levels (see MfSolver::plan)
    int seq_begin = 0, nseq = 0;   // range in the sequence-pointer array; 0 sequences = one CTA per front
    int maxnf;             // bulk: register bound on the front order (kernel variant)
    int smem_elems;        // bulk: front + staging capacity in elements (x 32 lanes x 8 bytes)
    long long gstride;
    bool staged;           // extend-add through the cp.async.bulk ring (LU scenario-tile kernel, batches)
    int ring_elems, ring_off, sec_cum;
    // batches: launches form a DAG (a launch waits only for the launches that hold children of its fronts); every
    // launch class has its own in-order stream, cross-stream edges are events
    int lane = 0;                  // stream index (0 = the caller's stream)
    bool record = false;           // some launch on another stream waits for this one
    std::vector<int> deps;         // launches on other streams to wait for (the latest per stream)
};

struct SolveLaunch {
    int begin, count;      // range in depth_fronts
    int max_nf, max_k;
    int ts;                // batch: scenarios per warp (32 / ts lanes cooperate on one scenario)
    size_t smem;
    int bs_rows;           // rows per block of the blocked kernel (<= 32)
    bool blocked;          // 32-row blocked kernel (single case; batch fronts too large for the tile staging)
    int seq_begin = 0, nseq = 0;   // sequences of this launch (walked from the last front to the first)
};

class MfSolver {
  public:
    Symbolic sym;

    // symmetric_matrix: the values are symmetric (WLS gain): mid-size fronts use the LDL^T variant of the kernel
    void setup(const Symbolic& s, cudaStream_t st, bool symmetric_matrix = false);
    // Factor A (values `aval`, CSC order of the analysed pattern, [nnz][S]) and solve A x = rhs for every scenario.
    // `active` (nullable, [S]) skips scenarios whose flag is 0. `status[s]` is set to -3 on a zero / non-finite pivot.
    void factor_solve(const double* aval, const double* rhs, double* x, int S, const unsigned char* active,
                      int* status, cudaStream_t st, cudaEvent_t after_factor = nullptr);
    // After a factor_solve with S == 1 of a symmetric matrix: solve A X = B for a block of R right-hand sides with the
    // stored factor (`solution!` per draw in the reference, utility.jl:576-586). B is [n][R] (entry i of column r at
    // B[i * R + r], R a multiple of 32) and is overwritten by X.
    // Unsymmetric values on a symmetric pattern: pass as `lower` a solver (same symbolic) that factored the transposed
    // matrix — A = L D U' means A' = U D L', so its packed rows are d_p L[j,p], exactly what the forward sweep needs.
    void solve_multi(double* B, int R, cudaStream_t st, const MfSolver* lower = nullptr);
    // After a factor_solve with S == 1 of a symmetric matrix: the entries of A^-1 on the pattern of L + L' (sparse
    // selected inverse, what the reference takes from `sparseinv` / Takahashi on the CHOLMOD factor,
    // stateEstimation/badData.jl:330-347, 536-640). Returns the device array; front f holds its nf x nf block (column
    // major, rows / columns in f_rows order) at zoff_host[f].
    const double* selected_inverse(cudaStream_t st);
    std::vector<long long> zoff_host;
    // Pivot guard: weak[s] (device, nullable) is set to 1 when a multiplier of scenario s exceeds `growth` during the
    // next factor_solve calls (the partial pivoting of UMFPACK / KLU would have swapped rows there).
    void set_pivot_guard(int* weak, double growth) { dev.weak = weak; dev.growth = growth; }
    int64_t factor_bytes(int S) const;     // algorithmic HBM bytes of one factor_solve (for roofline reports)
    int launches_per_solve(int S);
    int factor_launches(int S) { plan(S); return (int)(fplan.size() + tplan.size()); }
    // batch plan statistics: fronts handled by the task kernel, tasks, update-block elements that never leave the chip
    int task_fronts = 0, task_count = 0;
    long long task_upd_on_chip = 0;

  private:
    void plan(int S);
    void build_tasks(int S, cudaStream_t st);
    std::vector<TaskLaunch> tplan;
    TaskPlan task_plan;
    std::vector<char> in_task;             // front is factored by a task launch (not by fplan)
    std::vector<int> plan_levelptr, plan_fronts;   // level schedule of the fronts left to fplan
    DevBuf<int> d_task_blob, d_plan_fronts, d_plan_pair, d_plan_pair_s, d_upd_dst, d_chunks, d_seqptr;
    std::vector<int> plan_seqptr;          // sequence starts in plan_fronts (+ end), sequence mode only
    bool seq_mode = false;
    DevBuf<ChildDesc> d_plan_child;
    DevBuf<long long> d_plan_uoff;
    DevBuf<int2> d_task_desc;
    DevBuf<FrontDesc> d_plan_desc;
    int planned_S = -1;
    bool symmetric = false;
    std::vector<FactorLaunch> fplan;
    std::vector<SolveLaunch> splan;
    DevBuf<int> d_f_k, d_f_nf, d_f_rowptr, d_f_rows, d_f_relptr, d_f_rel, d_f_childptr, d_f_children, d_f_asmptr,
        d_asm_src, d_asm_dst, d_level_fronts, d_depth_fronts, d_f_eaptr, d_ea_roundptr, d_ea_pair, d_ea_roundptr_s,
        d_ea_pair_s;
    DevBuf<long long> d_f_uoff, d_f_updoff;
    DevBuf<double> d_U, d_upd, d_gwork, d_cvec;
    DevBuf<long long> d_coff, d_zoff;
    DevBuf<double> d_Zinv;
    DevBuf<int> d_parent;
    long long csum = 0;
    DevBuf<FrontDesc> d_level_desc;
    DevBuf<ChildDesc> d_child_desc;
    DevSym dev{};
    // DAG schedule of the batch factor phase
    std::vector<cudaStream_t> lanes;       // auxiliary streams (lane l > 0 -> lanes[l - 1])
    std::vector<cudaEvent_t> lane_events;  // one per factor launch
    cudaEvent_t fork_event = nullptr;
    int nlanes = 1;
  public:
    ~MfSolver();
};

}  // namespace jgb
