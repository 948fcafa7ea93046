// Symbolic analysis for the fixed-pattern multifrontal solver (host side, done once per topology).
//
// Replaces the symbolic half of what the reference gets from SuiteSparse through
// `factorization(...)` (src/backend/utility.jl:470-476, 534-540): fill-reducing ordering,
// elimination tree, supernodes (= fronts), assembly maps and the level schedules that the
// numeric kernels (solver.cu) replay at every Newton / Gauss-Newton iteration
// (`factorization!`, utility.jl:478-484 / 542-548) and for every scenario of a batch.
#pragma once
#include <cstdint>
#include <vector>

namespace jgb {

struct SymbolicOptions {
    int relax_small = 4;      // always merge a last child into its parent if the merged pivot count <= this
    int relax_mid = 16;       // merge up to this many pivots if the zero fraction stays below relax_mid_frac
    double relax_mid_frac = 0.5;
    int relax_big = 48;
    double relax_big_frac = 0.15;
    double relax_any_frac = 0.05;
    // ordering on the bus graph: minimum local fill (Tinney scheme 3) instead of minimum degree. Measured on the 10k-bus
    // grids: 9-16 % smaller update blocks and 14-18 % fewer flops for two or three more tree levels; the batch
    // Newton-Raphson factor phase drops from 48.3 to 43.6 ms per 10 016 scenarios, the WLS Monte-Carlo one from 29.5 to
    // 21.9 ms per 512 draws. The single-case presets cap the exact fill count (fill_exact_degree) to keep the tree shallow.
    bool min_fill = true;
    // exact fill counts up to this degree, an all-pairs upper bound beyond it (keeps the ordering of a 70 000-bus grid
    // at 1.5-3 s instead of 14 s; at 10k buses degrees never reach 48, so 48 means exact there)
    int fill_exact_degree = 48;
};

// Amalgamation presets: a single case is latency bound (fewer, larger fronts = fewer levels and launches), a batch
// is throughput bound (less amalgamation = fewer explicit zeros and flops).
inline SymbolicOptions latency_options() {
    SymbolicOptions o;
    // exact fill counts only up to degree 16: slightly more flops than the exact rule but a shallower tree (21 instead
    // of 25 levels on the 10k-bus Jacobian); measured single-case NR 1 285 (min degree) / 1 363 (exact) / 1 447 (24) and
    // WLS 244 (min degree) / 238 (exact) / 254 (24) / 272 (16) iterations per second
    o.fill_exact_degree = 16;
    return o;
}
inline SymbolicOptions throughput_options() {
    SymbolicOptions o;
    // re-tuned against the round-2 kernels (small fronts run at 2-3 TB/s, the scenario-tile kernel at 0.8-1.6: less
    // amalgamation pays): 4,8,0.3,24,0.1,0.02 -> factor 35.2 / back-solve 6.7 ms per 10 016 scenarios; these 33.8 / 6.4
    o.relax_small = 2; o.relax_mid = 6; o.relax_mid_frac = 0.2; o.relax_big = 16; o.relax_big_frac = 0.05;
    o.relax_any_frac = 0.01;
    return o;
}

struct Symbolic {
    int n = 0;                       // scalar dimension
    std::vector<int> perm;           // elimination position -> original variable
    std::vector<int> iperm;          // original variable -> elimination position

    int nfronts = 0;
    std::vector<int> f_k;            // pivots per front
    std::vector<int> f_nf;           // front order (pivots + update rows)
    std::vector<int> f_rowptr;       // [nfronts+1] offsets into f_rows
    std::vector<int> f_rows;         // ORIGINAL variable ids of the front rows; first k are the pivots
    std::vector<int> f_parent;       // parent front or -1
    std::vector<int> f_relptr;       // [nfronts+1] offsets into f_rel (one per update row)
    std::vector<int> f_rel;          // position of each update row in the parent's row list
    std::vector<int> f_childptr;     // [nfronts+1]
    std::vector<int> f_children;
    std::vector<int> f_asmptr;       // [nfronts+1] offsets into asm_src/asm_dst
    std::vector<int> asm_src;        // index into the matrix's CSC nzval
    std::vector<int> asm_dst;        // r + c*nf inside the front (column major, leading dimension nf)
    // extend-add as a gather: for every front, the destinations that receive child contributions and, per
    // destination, the source offsets into the update storage (children in order, so sums are deterministic)
    // Stored as rounds: round r of a front holds the r-th source of every destination that has one, so inside a round
    // all destinations are distinct (no race, one barrier between rounds) and a record is a single (dst, src) pair.
    std::vector<int> f_eaptr;        // [nfronts+1] range of rounds of a front in ea_roundptr
    std::vector<int> ea_roundptr;    // [nrounds+1] range of pairs of a round in ea_pair
    std::vector<int> ea_pair;        // 2 ints per pair: destination r + c*nf in the parent front, source element
                                     // offset into the update storage (f_updoff[child] + i + j*u_child)
    // The same lists for a symmetric matrix factored on packed lower triangles: only destinations on or below the
    // diagonal and the right-hand side, destination = packed index (column c at c*(2nf-c+1)/2, entry r-c; rhs entry r at
    // nf(nf+1)/2 + r) — half the pairs, no index arithmetic in the kernels
    std::vector<int> f_eaptr_sym, ea_roundptr_sym, ea_pair_sym;
    std::vector<int64_t> f_uoff;     // offset of the front's packed U rows (k rows, row p has nf+1-p entries)
    std::vector<int64_t> f_updoff;   // offset of the front's update block (u x (u+1), column major, last col = rhs)
    int64_t u_size = 0, upd_size = 0;

    // schedules
    int nlevels = 0;                 // factor levels: leaves first
    std::vector<int> levelptr;       // [nlevels+1]
    std::vector<int> level_fronts;   // fronts grouped by height level (sorted by decreasing nf inside a level)
    int ndepths = 0;                 // back-solve levels: roots first
    std::vector<int> depthptr;
    std::vector<int> depth_fronts;

    // statistics
    int64_t nnz_lu = 0;              // scalar nnz(L+U) incl. diagonal, with amalgamation zeros
    double flops = 0;                // factorisation flops (LU)
    int max_front = 0;
};

// Pattern: CSC of an n x n matrix (0-based). `group[v]` (may be null) ties variables that must stay adjacent
// in the ordering (theta_i / V_i of one bus). The pattern is symmetrised (A + A') internally; entries of A that
// are structurally absent from A' simply stay zero in the fronts.
// `skip[v] != 0` (may be null) marks variables whose row/column is identity (WLS slack angle): they are ordered
// last as isolated 1x1 fronts and entries touching them are ignored by the assembly map.
void analyse(int n, const int* colptr, const int* rowidx, const int* group, const unsigned char* skip,
             const SymbolicOptions& opt, Symbolic& out);

// Host reference of the numeric phase (S = 1): used by the CPU self-check tests of the symbolic data
// only — never by the operators. Returns 0, or -3 on a zero pivot.
int host_factor_solve(const Symbolic& s, const double* aval, const double* rhs, double* x);

}  // namespace jgb
