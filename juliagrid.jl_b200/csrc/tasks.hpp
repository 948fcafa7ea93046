// Task partition of the elimination tree for the batch factorisation (host side, no CUDA): which fronts one CTA of
// mf_task_kernel (mf_task.cuh) walks, where every update block sits on the CTA's shared-memory stack, and the index
// blobs the kernel copies to shared memory. Replaces nothing in the reference: this is scheduling data for the numeric
// refactorisation that stands in for `lu!/klu!` (src/backend/utility.jl:478-500).
#pragma once
#include <vector>

#include "symbolic.hpp"

namespace jgb {

constexpr int kTaskRec = 12;       // ints per front record in a task blob (layout: mf_task.cuh)

// One launch of the task kernel: `count` CTAs per scenario tile, each walking one task list
struct TaskLaunch {
    int begin, count;      // range in the task descriptor array
    int te, maxnf;         // kernel variant
    int front_cap, stack_cap, meta_cap;   // shared-memory areas: elements (x 32 lanes x 8 bytes) / ints
    size_t smem;
};

// Measured on the 10k-bus Jacobian, 10 016 scenarios (profiles/r02_task_kernel_ab.txt): the task kernel removes 29 % of
// the update-block traffic and ~60 launches per factorisation, but one CTA needs the largest front, the stack and the
// index data of its task in shared memory (up to 150 KB), which leaves one CTA of 8 warps per SM against 2-8 CTAs for the
// per-front kernels — 53.7 ms per factor phase instead of 43.6 ms. It stays in the tree as an opt-in (JGB_TASKS=1) with
// its CPU replay test; the per-front kernels are the default.
struct TaskOptions {
    bool enabled = false;
    int maxnf = 16;        // largest front order a task may contain (<= 16: register-resident elimination)
    int stack_lim = 256;   // update-block stack, elements per scenario
    int meta_lim = 4096;   // index data per CTA, ints
    int bundle = 24;       // fronts per CTA when several small tasks are packed together
    int task_max = 64;     // fronts per task
};
TaskOptions task_options_from_env();    // JGB_NO_TASKS, JGB_TASK_MAXNF / _STACK / _META / _BUNDLE / _MAXFRONTS (tuning only)

struct TaskPlan {
    std::vector<TaskLaunch> launches;   // in dependency order
    std::vector<int> blob;              // all task blobs
    std::vector<int> descs;             // 2 ints per task: blob offset, blob length
    std::vector<char> in_task;          // per front
    // places in `blob` that hold an update-storage offset (lo word; hi word follows) or a tile width, with the front they
    // refer to: MfSolver::plan patches them once the layout of the update storage is known (the blob as built here
    // addresses the plain layout of Symbolic::f_updoff at tile width 32, which is what the host replay walks)
    std::vector<int> off_pos, off_front, wout_pos, wout_front, uoff_pos, uoff_front;
    int task_fronts = 0, task_count = 0;
    long long upd_on_chip = 0;          // update-block elements per scenario that never leave shared memory
};

void partition_tasks(const Symbolic& sym, const TaskOptions& opt, TaskPlan& out);

// Host replay of the task schedule for one scenario (CPU self-check of the blobs only — never used by the operators):
// factors the task fronts exactly as the kernel walks them, the remaining fronts like host_factor_solve, then
// back-substitutes. Returns 0, or -3 on a zero pivot.
int host_task_factor_solve(const Symbolic& s, const TaskPlan& tp, const double* aval, const double* rhs, double* x);

}  // namespace jgb
