/* jgb200 — C ABI of the B200-native Newton-Raphson / Gauss-Newton hot path behind JuliaGrid's operator surface.
 *
 * The reference (mcosovic/JuliaGrid.jl v0.6.2) has no FFI of its own; its seam is Julia dispatch on the
 * factorisation tag (`newtonRaphson(system, ::Type{T})` src/powerFlow/acPowerFlow.jl:39,
 * `gaussNewton(monitoring, ::Type{T})` src/stateEstimation/acStateEstimation.jl:43 and the triad
 * `factorization / factorization! / solution!` src/backend/utility.jl:470-586).  Each entry point below names
 * the reference function it stands in for; INTEGRATION.md shows the `ccall` bindings a maintainer adds.
 *
 * Conventions
 *  - Arrays are the reference's own: Float64, Int64 **1-based** indices, Int8 flags, ComplexF64 passed as
 *    interleaved (re, im) doubles.  The library copies during the call and never keeps a host pointer.
 *  - Every function returns an int32 status: 0 ok; 1 not converged within the iteration cap (soft, like the
 *    reference — not an error); < 0 error: -1 bad argument, -2 CUDA error, -3 singular / non-finite pivot,
 *    -4 pattern mismatch, -5 no CUDA device.  `jgb_last_error(ctx)` returns the message of the last error.
 *  - A context is bound to one GPU and one stream; calls on one context must not overlap.
 *  - There is no CPU fallback: without a CUDA device `jgb_create` fails with -5.
 */
#ifndef JGB200_H
#define JGB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct jgb_ctx jgb_ctx;

/* ---- context ---------------------------------------------------------------------------------------------- */
/* `stream` is a cudaStream_t (may be NULL for a private non-blocking stream).  *rc receives the status. */
jgb_ctx* jgb_create(int32_t device, void* stream, int32_t* rc);
void jgb_destroy(jgb_ctx* ctx);
const char* jgb_last_error(const jgb_ctx* ctx);          /* ctx may be NULL: error of the last failed jgb_create */
int32_t jgb_abi_version(void);
int32_t jgb_synchronize(jgb_ctx* ctx);

/* ---- Newton-Raphson AC power flow -------------------------------------------------------------------------- */
/* newtonRaphson(system) + newtonJacobian (acPowerFlow.jl:39-175): takes Ybus and its transpose exactly as stored
 * in system.model.ac (nodalMatrix / nodalMatrixTranspose, same pattern), bus.layout.type, bus.layout.slack.
 * Builds pq / pvpq / pcount, the Jacobian CSC pattern, the symbolic factorisation, and uploads everything. */
int32_t jgb_nr_setup(jgb_ctx* ctx, int64_t n, const int64_t* y_colptr, const int64_t* y_rowval,
                     const double* y_nzval_re_im, const double* yt_nzval_re_im, const int8_t* bus_type,
                     int64_t slack);
/* sizes of the Jacobian built by jgb_nr_setup */
int32_t jgb_nr_dims(jgb_ctx* ctx, int64_t* dim_j, int64_t* nnz_j);
/* the reference's own index arrays for bit-exact parity: NewtonRaphson.pq/pvpq/pcount, jacobian.colptr/rowval */
int32_t jgb_nr_pattern(jgb_ctx* ctx, int64_t* pq, int64_t* pvpq, int64_t* pcount, int64_t* j_colptr,
                       int64_t* j_rowval);
/* bus.supply.active/reactive, bus.demand.active/reactive (read by mismatch!, acPowerFlow.jl:676-680) */
int32_t jgb_nr_set_injection(jgb_ctx* ctx, const double* p_supply, const double* q_supply, const double* p_demand,
                             const double* q_demand);
/* analysis.voltage.magnitude / angle */
int32_t jgb_nr_set_state(jgb_ctx* ctx, const double* vm, const double* va);
int32_t jgb_nr_get_state(jgb_ctx* ctx, double* vm, double* va);
/* value-only Ybus update on the fixed pattern (acNodalUpdate!, powerSystem/model.jl:81-110):
 * k stored positions (1-based into nzval) with their new Y and Y-transpose values */
int32_t jgb_nr_update_y(jgb_ctx* ctx, int64_t k, const int64_t* nz_pos, const double* y_re_im,
                        const double* yt_re_im);
/* mismatch!(analysis) (acPowerFlow.jl:645-685) -> (max|dP|, max|dQ|) */
int32_t jgb_nr_mismatch(jgb_ctx* ctx, double* stop_p, double* stop_q);
/* solve!(analysis) (acPowerFlow.jl:793-911): Jacobian fill, numeric refactor, solve, state update, iteration += 1 */
int32_t jgb_nr_solve(jgb_ctx* ctx);
/* method.mismatch, method.increment, method.jacobian.nzval, method.iteration (any pointer may be NULL) */
int32_t jgb_nr_get_vectors(jgb_ctx* ctx, double* mismatch, double* increment, double* j_nzval, int64_t* iteration);
/* powerFlow!(analysis; iteration, tolerance) loop (acPowerFlow.jl:1389-1433), convergence test on the device.
 * Returns 0 converged / 1 iteration cap reached. */
int32_t jgb_nr_run(jgb_ctx* ctx, int64_t max_iter, double tol, int64_t* iterations, double* stop_p, double* stop_q);
/* S independent power flows sharing topology and symbolic factorisation (N-1 sweep: updateBranch!(status = 0),
 * powerSystem/branch.jl:313-431, each followed by powerFlow!).  Scenario s removes branch with end buses
 * out_from[s], out_to[s] (1-based bus indices, 0 = base case) whose Y-parameters nodalFromFrom / nodalFromTo /
 * nodalToFrom / nodalToTo are given in dy_re_im[s][0..7].  Every scenario starts from the state that is on the
 * device when the call is made: the one last set with jgb_nr_set_state, or the iterate a later jgb_nr_solve /
 * jgb_nr_run left there (call jgb_nr_set_state with the start point first for setInitialPoint! semantics).  Outputs are S x n row-major (one row per scenario); status[s] is
 * 0 converged, 1 iteration cap, -3 singular (islanding outage).  *total_iterations = sum of solve! calls.
 * Batch size: scenario tiles sit in gridDim.y, so a call takes at most 65 535 tiles of the narrowest tile width the
 * matrix needs (32 scenarios per tile for fronts up to 16 rows ... 1 above 64 rows): S <= 65 535 always works, larger S
 * only for matrices without large fronts; otherwise the call returns -1 and the caller splits the batch (the same holds
 * for jgb_wls_batch). */
int32_t jgb_nr_batch(jgb_ctx* ctx, int64_t S, const int64_t* out_from, const int64_t* out_to, const double* dy_re_im,
                     int64_t max_iter, double tol, double* vm_out, double* va_out, int32_t* iterations,
                     int8_t* status, int64_t* total_iterations);
/* Same, but inputs already resident: out_from/out_to (int64), dy (double) and the outputs vm_out/va_out (double,
 * S x n), iterations (int32), status (int8) are DEVICE pointers on the context's GPU. */
int32_t jgb_nr_batch_dev(jgb_ctx* ctx, int64_t S, const int64_t* out_from_dev, const int64_t* out_to_dev,
                         const double* dy_dev, int64_t max_iter, double tol, double* vm_out_dev, double* va_out_dev,
                         int32_t* iterations_dev, int8_t* status_dev, int64_t* total_iterations);

/* ---- post-processing on the device ------------------------------------------------------ */
/* branch.layout.from/to (1-based), model.ac.nodalFromFrom / nodalFromTo / nodalToFrom / nodalToTo (ComplexF64) and
 * branch.layout.status, needed by jgb_nr_power only */
int32_t jgb_nr_set_branches(jgb_ctx* ctx, int64_t nbranch, const int64_t* from, const int64_t* to,
                            const double* y_ff_re_im, const double* y_ft_re_im, const double* y_tf_re_im,
                            const double* y_tt_re_im, const int8_t* status);
/* power!(analysis) / current!(analysis) for the quantities the measurement generators consume
 * (postprocessing/acAnalysis.jl:30-79, 672-700), evaluated at the state on the device: bus injections (n) and, per
 * branch, from/to active and reactive flows and current magnitude / angle (out-of-service branches give 0).
 * Any pointer may be NULL. */
int32_t jgb_nr_power(jgb_ctx* ctx, double* injection_active, double* injection_reactive, double* from_active,
                     double* from_reactive, double* to_active, double* to_reactive, double* from_current_magnitude,
                     double* from_current_angle, double* to_current_magnitude, double* to_current_angle);

/* ---- Gauss-Newton WLS AC state estimation ---------------------------------------------------------------------- */
/* gaussNewton(monitoring) after acWLS (acStateEstimation.jl:43-259): takes the tables acWLS builds — the CSC pattern
 * of the Jacobian H (m x 2n; theta columns 1..n, V columns n+1..2n), type (Int8 codes 0..21), index (bus or branch),
 * range (6 row offsets), the precision matrix W (CSC; diagonal or 2x2 blocks for correlated rectangular PMUs) —
 * plus Ybus / Ybus transpose, branch.layout.from/to, branch.parameter.conductance/susceptance/turnsRatio/shiftAngle
 * and model.ac.admittance (ComplexF64).  Builds the pattern of G = H'WH, its gather lists and symbolic factorisation. */
int32_t jgb_wls_setup(jgb_ctx* ctx, int64_t n, int64_t m, int64_t slack, const int64_t* h_colptr,
                      const int64_t* h_rowval, const int8_t* type, const int64_t* index, const int64_t* range6,
                      const int64_t* w_colptr, const int64_t* w_rowval, const double* w_nzval,
                      const int64_t* y_colptr, const int64_t* y_rowval, const double* y_nzval_re_im,
                      const double* yt_nzval_re_im, int64_t nbranch, const int64_t* from, const int64_t* to,
                      const double* conductance, const double* susceptance, const double* turns_ratio,
                      const double* shift_angle, const double* admittance_re_im);
int32_t jgb_wls_dims(jgb_ctx* ctx, int64_t* nnz_h, int64_t* nnz_g);
/* CSC pattern (1-based) of the gain matrix as Julia's `transpose(H) * W * H` stores it (slack row/column kept) */
int32_t jgb_wls_gain_pattern(jgb_ctx* ctx, int64_t* g_colptr, int64_t* g_rowval);
/* method.mean (length m) */
int32_t jgb_wls_set_mean(jgb_ctx* ctx, const double* z);
int32_t jgb_wls_set_state(jgb_ctx* ctx, const double* vm, const double* va);
int32_t jgb_wls_get_state(jgb_ctx* ctx, double* vm, double* va);
/* increment!(analysis) (acStateEstimation.jl:878-904): normalEquation!, G = H'WH with the slack fix, refactor,
 * solve, increment[slack] = 0 -> returns maximum(abs, increment); *objective = method.objective */
int32_t jgb_wls_increment(jgb_ctx* ctx, double* max_increment, double* objective);
/* solve!(analysis) (acStateEstimation.jl:1035-1047): theta += d[1:n], V += d[n+1:2n], iteration += 1 */
int32_t jgb_wls_solve(jgb_ctx* ctx);
/* method.residual (m), method.increment (2n), method.jacobian.nzval (CSC order of h_colptr/h_rowval), gain values
 * (CSC order of jgb_wls_gain_pattern), method.iteration; any pointer may be NULL */
int32_t jgb_wls_get_vectors(jgb_ctx* ctx, double* residual, double* increment, double* h_nzval, double* g_nzval,
                            int64_t* iteration);
/* stateEstimation!(analysis; iteration, tolerance) loop (acStateEstimation.jl:1286-1329). 0 converged / 1 cap. */
int32_t jgb_wls_run(jgb_ctx* ctx, int64_t max_iter, double tol, int64_t* iterations, double* max_increment,
                    double* objective);
/* S independent estimations on one topology / measurement layout differing only in the means (Monte-Carlo draws):
 * Z is S x m row-major; every draw starts from the state last set with jgb_wls_set_state. Outputs S x n row-major. */
int32_t jgb_wls_batch(jgb_ctx* ctx, int64_t S, const double* Z, int64_t max_iter, double tol, double* vm_out,
                      double* va_out, int32_t* iterations, int8_t* status, double* objective,
                      int64_t* total_iterations);
/* Same with DEVICE pointers for Z and all outputs. */
int32_t jgb_wls_batch_dev(jgb_ctx* ctx, int64_t S, const double* Z_dev, int64_t max_iter, double tol,
                          double* vm_out_dev, double* va_out_dev, int32_t* iterations_dev, int8_t* status_dev,
                          double* objective_dev, int64_t* total_iterations);

/* ---- in-place updates of a Gauss-Newton analysis (the Monte-Carlo / what-if loop of test/stateEstimation/reusing.jl) ---
 * == update*!(analysis; ...) for single measurement rows (e.g. _updateWattmeter!, src/measurement/powermeter.jl:640-677,
 * and its siblings in voltmeter.jl / ammeter.jl / pmu.jl): for every listed row (1-based) mean = status * mean,
 * residual = 0, the row's Jacobian entries = 0, type = status * code, index, precision[row,row] = 1 / variance
 * (precision_off = W[row,row-1] of a correlated rectangular PMU pair). NULL arrays keep the stored values. The H and
 * gain patterns, the gather lists and the symbolic factorisation are reused; a change that needs a new pattern
 * returns -4 (build a new context, like the reference rebuilds on a pattern change). */
int32_t jgb_wls_update_rows(jgb_ctx* ctx, int64_t k, const int64_t* rows, const double* mean, const double* precision,
                            const double* precision_off, const int8_t* type, const int64_t* index);
/* == acNodalUpdate! (powerSystem/model.jl:81-110) on the model a WLS analysis reads: k stored Ybus positions */
int32_t jgb_wls_update_y(jgb_ctx* ctx, int64_t k, const int64_t* nz_pos, const double* y_re_im, const double* yt_re_im);
/* == acParameterUpdate! (powerSystem/model.jl:113-140): the parameters the flow / current rows of one branch read */
int32_t jgb_wls_update_branch(jgb_ctx* ctx, int64_t branch, double conductance, double susceptance, double turns_ratio,
                              double shift_angle, const double* admittance_re_im);

/* ---- bad-data post-step -------------------------------------------------------------------------
 * == the numeric part of residualTest!(analysis; threshold) for Gauss-Newton WLS (src/stateEstimation/badData.jl:181-285):
 * c[i] = h_i G^-1 h_i' from a sparse selected inverse of the gain factor on the device (Takahashi recurrences on the
 * elimination tree; `selectedInverse` / `rowProjection` badData.jl:287-362, 536-640), then the largest normalised
 * residual |r_i| / sqrt|1/W_ii - c_i| over the rows with a non-zero residual. Call after jgb_wls_run / the last
 * jgb_wls_increment. index: 1-based row, 0 if none; c_out: nullable [m]. `threshold` is not applied here: the call
 * reports the maximum and its row, and the caller compares (bad.detect = maximum > threshold, badData.jl:220-222) and
 * then removes the row(s) with jgb_wls_remove_row. */
int32_t jgb_wls_residual_test(jgb_ctx* ctx, double threshold, double* max_normalized_residual, int64_t* index,
                              double* c_out);
/* Row `row` (1-based) leaves the model as in badData.jl:258-282: H entries, mean and residual zeroed, type 0,
 * iteration 0 (the monitoring-side status bookkeeping stays with the caller). */
int32_t jgb_wls_remove_row(jgb_ctx* ctx, int64_t row);

/* ---- constant-matrix linear solves -------------------------------------------------------------
 * The reference's linear analyses factor one sparse symmetric matrix and call `solution!` per right-hand side:
 *   DC power flow      solve!  src/powerFlow/dcPowerFlow.jl:93-134        (B theta = P, slack row/column -> identity)
 *   DC state estimation solve! src/stateEstimation/dcStateEstimation.jl:342-371   (H'WH theta = H'W z)
 *   PMU state estimation solve! src/stateEstimation/pmuStateEstimation.jl:369-399 (H'WH [Re V; Im V] = H'W z)
 * through `factorization / factorization! / solution!` (src/backend/utility.jl:470-586). Here the matrix is factored
 * once on the device (LDL^T on the fixed elimination tree, no pivoting) and a whole block of right-hand sides — one
 * per Monte-Carlo draw or injection scenario — is solved with the stored factor. */
/* == factorization(A, F, T): A is n x n symmetric, full CSC as SparseMatrixCSC{Float64,Int64} (1-based, sorted rows,
 * both triangles). skip (1-based, 0 = none): row and column `skip` are replaced by the identity (slack bus), so
 * x[skip] = b[skip]. Returns -3 on a zero / non-finite pivot. */
int32_t jgb_lin_setup(jgb_ctx* ctx, int64_t n, const int64_t* a_colptr, const int64_t* a_rowval, const double* a_nzval,
                      int64_t skip);
/* == factorization!(A, F, T): new values on the pattern given to jgb_lin_setup. */
int32_t jgb_lin_refactor(jgb_ctx* ctx, const double* a_nzval);
/* P = precision * coefficient (W H, m x n CSC, 1-based): right-hand sides are formed on the device as b = P' z
 * (`temp * se.mean`, dcStateEstimation.jl:363, pmuStateEstimation.jl:383). Column `skip` of P is ignored. */
int32_t jgb_lin_projection(jgb_ctx* ctx, int64_t m, const int64_t* p_colptr, const int64_t* p_rowval,
                           const double* p_nzval);
/* == solution!(x, F, b) for R right-hand sides: b and x are [R][n], each vector contiguous. */
int32_t jgb_lin_solve(jgb_ctx* ctx, int64_t R, const double* b, double* x);
/* x_r = A^-1 (P' z_r) for R measurement vectors z [R][m]. */
int32_t jgb_lin_solve_projected(jgb_ctx* ctx, int64_t R, const double* z, double* x);
/* Same two calls with DEVICE pointers (projected != 0 selects the second form). */
int32_t jgb_lin_solve_dev(jgb_ctx* ctx, int64_t R, const double* in_dev, double* x_dev, int32_t projected);
int32_t jgb_lin_dims(jgb_ctx* ctx, int64_t* n, int64_t* m, int64_t* nnz_factor, int64_t* fronts);

/* ---- fast Newton-Raphson BX / XB ----------------------------------------------------------------
 * == AcPowerFlow{FastNewtonRaphson}: mismatch! (src/powerFlow/acPowerFlow.jl:686-727), solve! (:913-983), powerFlow!
 * (:1389-1433). The constant Jacobians B' (active, (n-1) x (n-1)) and B'' (reactive, npq x npq) are the reference's own
 * (`fastNewtonRaphsonBX / XB`, :215-339), passed as SparseMatrixCSC; they are factored once on the device (B' may have
 * unsymmetric values when phase shifters are present). y_t = nodalMatrixTranspose.nzval on the Ybus pattern. */
int32_t jgb_fnr_setup(jgb_ctx* ctx, int64_t n, const int64_t* y_colptr, const int64_t* y_rowval,
                      const double* yT_nzval_re_im, const int8_t* bus_type, int64_t slack,
                      const int64_t* bp_colptr, const int64_t* bp_rowval, const double* bp_nzval,
                      const int64_t* bq_colptr, const int64_t* bq_rowval, const double* bq_nzval);
int32_t jgb_fnr_set_injection(jgb_ctx* ctx, const double* p_supply, const double* q_supply, const double* p_demand,
                              const double* q_demand);
int32_t jgb_fnr_set_state(jgb_ctx* ctx, const double* vm, const double* va);
int32_t jgb_fnr_get_state(jgb_ctx* ctx, double* vm, double* va);
int32_t jgb_fnr_mismatch(jgb_ctx* ctx, double* stop_p, double* stop_q);           /* == mismatch! */
int32_t jgb_fnr_solve(jgb_ctx* ctx);                                              /* == solve!    */
int32_t jgb_fnr_run(jgb_ctx* ctx, int64_t max_iter, double tol, int64_t* iterations, double* stop_p, double* stop_q);
/* R injection scenarios on one topology (the user loop updateBus!(active / reactive) + powerFlow!): p_inj, q_inj [R][n]
 * = supply - demand per bus; every scenario starts from the state of jgb_fnr_set_state; outputs [R][n];
 * status 0 converged, 1 iteration cap, -3 diverged. Returns 1 if any scenario did not converge. */
int32_t jgb_fnr_batch(jgb_ctx* ctx, int64_t R, const double* p_inj, const double* q_inj, int64_t max_iter, double tol,
                      double* vm_out, double* va_out, int32_t* iterations, int8_t* status, int64_t* total_iterations);

/* ---- multi-GPU: scenario sharding with one all-gather of the converged states --------------------
 * The reference has no distributed code: contingencies and Monte-Carlo draws are iterations of a serial user loop
 * (updateBranch! + powerFlow!, test/powerFlow/reusing.jl:40-84; updateWattmeter!… + stateEstimation!,
 * test/stateEstimation/reusing.jl). Here every rank (one process / context per GPU) solves its block of scenarios with
 * jgb_nr_batch_dev / jgb_wls_batch_dev and the blocks are collected with ONE grouped NCCL all-gather over NVLink.
 * NCCL is loaded at run time (libnccl.so.2, or the path in JGB200_NCCL_LIB); without it these calls fail with -1 and
 * everything else works. */
/* ncclGetUniqueId: call on one rank and hand the 128 bytes to the others by any host-side means */
int32_t jgb_comm_unique_id(uint8_t* id128);
/* ncclCommInitRank on the context's GPU (collective: every rank of the job calls it) */
int32_t jgb_comm_init(jgb_ctx* ctx, int32_t rank, int32_t nranks, const uint8_t* id128);
/* All-gather rows_local scenarios of every rank into rank-major [nranks * rows_local] arrays: states [rows][n] (FP64),
 * iteration counts (int32) and status bytes (int8); all pointers are DEVICE pointers, any send pointer may be NULL.
 * rows_local must be the same on every rank (pad the last block). The collective starts when the work already enqueued
 * on the context's stream has finished and runs on a private stream beside whatever is enqueued next; send and receive
 * buffers must not be touched until jgb_comm_wait (or the next jgb_allgather_states, which waits first). */
int32_t jgb_allgather_states(jgb_ctx* ctx, int64_t rows_local, int64_t n, const double* vm_dev, const double* va_dev,
                             const int32_t* iterations_dev, const int8_t* status_dev, double* vm_all_dev,
                             double* va_all_dev, int32_t* iterations_all_dev, int8_t* status_all_dev);
/* host_blocking != 0: return when the all-gather has finished; 0: make the context's stream wait for it */
int32_t jgb_comm_wait(jgb_ctx* ctx, int32_t host_blocking);
int32_t jgb_comm_size(jgb_ctx* ctx, int32_t* rank, int32_t* nranks);     /* -1 / 0 before jgb_comm_init */

/* ---- statistics for roofline reports ------------------------------------------------------------------------ */
/* key: "nr.nnz_lu", "nr.fronts", "nr.levels", "nr.flops", "nr.max_front", "nr.launches_per_iteration",
 *      "nr.assemble_bytes" (per scenario-iteration), "nr.solve_bytes", "wls.*" likewise; kernel launch counter
 *      "launches" (since context creation).  Unknown key -> -1. */
double jgb_stat(jgb_ctx* ctx, const char* key);
/* enable / disable + reset the CUDA-event phase timers of the batch loops ("nr.time.assemble_ms", "nr.time.factor_ms",
 * "nr.time.backsolve_ms", "wls.time.*"; accumulated since the last jgb_profile call) */
int32_t jgb_profile(jgb_ctx* ctx, int32_t enable);

/* ---- symbolic self-check (host only, no GPU work): runs the analysis on a CSC pattern and a host replay of the
 *      multifrontal schedule; used by the CPU test-suite to validate the maps the kernels consume. */
int32_t jgb_selfcheck_symbolic(int64_t n, const int64_t* colptr, const int64_t* rowval, const double* nzval,
                               const int64_t* group, const double* rhs, double* x, double* stats8);

/* host only: the elimination tree of the analysis (preset 0 default / 1 latency / 2 throughput): pivots, order,
 * parent (-1 = root) and number of matrix entries of every front; arrays of capacity `cap`. For tests and tuning. */
int32_t jgb_selfcheck_tree(int64_t n, const int64_t* colptr, const int64_t* rowval, const int64_t* group, int32_t preset,
                           int64_t cap, int64_t* nfronts, int32_t* f_k, int32_t* f_nf, int32_t* f_parent,
                           int32_t* f_nasm);

/* host only: the task partition of the batch factorisation (throughput preset) replayed on the host for one
 * right-hand side. stats8: launches, tasks, fronts in tasks, update-block elements kept on chip, largest shared-memory
 * need (bytes), index-blob ints, fronts, update-block elements in total. */
int32_t jgb_selfcheck_tasks(int64_t n, const int64_t* colptr, const int64_t* rowval, const double* nzval,
                            const int64_t* group, const double* rhs, double* x, double* stats8);

#ifdef __cplusplus
}
#endif
#endif /* JGB200_H */
