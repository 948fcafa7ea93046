// extern "C" boundary of libjgb200.so (see include/jgb200.h). Plain pointers and sizes only; every C++ exception is
// mapped to a negative status code and a message retrievable with jgb_last_error().
#include "../../include/jgb200.h"

#include <memory>
#include <string>

#include "comm.cuh"
#include "common.cuh"
#include "fnr.cuh"
#include "lin.cuh"
#include "nr.cuh"
#include "symbolic.hpp"
#ifdef JGB_WITH_WLS
#include "wls.cuh"
#endif

struct jgb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    std::unique_ptr<jgb::NrContext> nr;
    std::unique_ptr<jgb::LinContext> lin;
    std::unique_ptr<jgb::FnrContext> fnr;
    std::unique_ptr<jgb::CommContext> comm;
#ifdef JGB_WITH_WLS
    std::unique_ptr<jgb::WlsContext> wls;
#endif
};

namespace {
std::string g_create_error;

template <typename F>
int32_t guarded(jgb_ctx* ctx, F&& fn) {
    if (!ctx) return -1;
    try {
        cudaError_t e = cudaSetDevice(ctx->device);
        if (e != cudaSuccess) throw jgb::CudaError(std::string("cudaSetDevice: ") + cudaGetErrorString(e));
        return fn();
    } catch (const std::invalid_argument& e) {
        ctx->err = e.what();
        return -1;
    } catch (const jgb::CudaError& e) {
        ctx->err = e.what();
        cudaGetLastError();
        return -2;
    } catch (const std::domain_error& e) {
        ctx->err = e.what();
        return -3;
    } catch (const std::logic_error& e) {
        ctx->err = e.what();
        return -1;
    } catch (const std::exception& e) {
        ctx->err = e.what();
        return -4;
    }
}

// Batch kernels put the scenario tiles in gridDim.y (<= 65535 tiles of 32): reject what cannot be launched instead of
// failing with a CUDA launch error, and the iteration cap / tolerance like the single-case entry points do.
constexpr int64_t kMaxBatch = 65535LL * 32;
void check_batch_args(const char* who, int64_t S, int64_t max_iter, double tol) {
    if (S <= 0 || S > kMaxBatch)
        throw std::invalid_argument(std::string(who) + ": batch size must be in 1.." + std::to_string(kMaxBatch));
    if (max_iter < 0 || !(tol > 0)) throw std::invalid_argument(std::string(who) + ": max_iter >= 0 and tol > 0 required");
}

jgb::NrContext& nr_of(jgb_ctx* ctx) {
    if (!ctx->nr) throw std::logic_error("jgb_nr_setup has not been called on this context");
    return *ctx->nr;
}
}  // namespace

extern "C" {

int32_t jgb_abi_version(void) { return 1; }

jgb_ctx* jgb_create(int32_t device, void* stream, int32_t* rc) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_create_error = std::string("no CUDA device available (") + cudaGetErrorString(e) +
                         "); jgb200 has no CPU fallback";
        cudaGetLastError();
        if (rc) *rc = -5;
        return nullptr;
    }
    if (device < 0 || device >= count) {
        g_create_error = "device index out of range";
        if (rc) *rc = -1;
        return nullptr;
    }
    e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
        if (rc) *rc = -2;
        return nullptr;
    }
    jgb_ctx* ctx = new jgb_ctx();
    ctx->device = device;
    if (stream) {
        ctx->stream = (cudaStream_t)stream;
    } else {
        e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            g_create_error = std::string("cudaStreamCreate: ") + cudaGetErrorString(e);
            delete ctx;
            if (rc) *rc = -2;
            return nullptr;
        }
        ctx->own_stream = true;
    }
    if (rc) *rc = 0;
    return ctx;
}

void jgb_destroy(jgb_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx->nr.reset();
#ifdef JGB_WITH_WLS
    ctx->wls.reset();
#endif
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* jgb_last_error(const jgb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int32_t jgb_synchronize(jgb_ctx* ctx) {
    return guarded(ctx, [&] { JGB_CUDA(cudaStreamSynchronize(ctx->stream)); return 0; });
}

int32_t jgb_nr_setup(jgb_ctx* ctx, int64_t n, const int64_t* y_colptr, const int64_t* y_rowval,
                     const double* y_nzval, const double* yt_nzval, const int8_t* bus_type, int64_t slack) {
    return guarded(ctx, [&] {
        auto nr = std::make_unique<jgb::NrContext>(ctx->stream);
        nr->setup(n, y_colptr, y_rowval, y_nzval, yt_nzval, bus_type, slack);
        ctx->nr = std::move(nr);
        return 0;
    });
}

int32_t jgb_nr_dims(jgb_ctx* ctx, int64_t* dim_j, int64_t* nnz_j) {
    return guarded(ctx, [&] {
        auto& nr = nr_of(ctx);
        if (dim_j) *dim_j = nr.dim;
        if (nnz_j) *nnz_j = nr.nnzj;
        return 0;
    });
}

int32_t jgb_nr_pattern(jgb_ctx* ctx, int64_t* pq, int64_t* pvpq, int64_t* pcount, int64_t* j_colptr,
                       int64_t* j_rowval) {
    return guarded(ctx, [&] {
        auto& nr = nr_of(ctx);
        if (pq) std::copy(nr.pq1.begin(), nr.pq1.end(), pq);
        if (pvpq) std::copy(nr.pvpq1.begin(), nr.pvpq1.end(), pvpq);
        if (pcount) std::copy(nr.pcount1.begin(), nr.pcount1.end(), pcount);
        if (j_colptr) std::copy(nr.jcolptr1.begin(), nr.jcolptr1.end(), j_colptr);
        if (j_rowval) std::copy(nr.jrowval1.begin(), nr.jrowval1.end(), j_rowval);
        return 0;
    });
}

int32_t jgb_nr_set_injection(jgb_ctx* ctx, const double* ps, const double* qs, const double* pd, const double* qd) {
    return guarded(ctx, [&] { nr_of(ctx).set_injection(ps, qs, pd, qd); return 0; });
}

int32_t jgb_nr_set_state(jgb_ctx* ctx, const double* vm, const double* va) {
    return guarded(ctx, [&] { nr_of(ctx).set_state(vm, va); return 0; });
}

int32_t jgb_nr_get_state(jgb_ctx* ctx, double* vm, double* va) {
    return guarded(ctx, [&] {
        if (!vm || !va) throw std::invalid_argument("nr_get_state: null output");
        nr_of(ctx).get_state(vm, va);
        return 0;
    });
}

int32_t jgb_nr_update_y(jgb_ctx* ctx, int64_t k, const int64_t* nz_pos, const double* y, const double* yt) {
    return guarded(ctx, [&] {
        if (k < 0 || (k > 0 && (!nz_pos || !y || !yt))) throw std::invalid_argument("nr_update_y: null input");
        nr_of(ctx).update_y(k, nz_pos, y, yt);
        return 0;
    });
}

int32_t jgb_nr_mismatch(jgb_ctx* ctx, double* stop_p, double* stop_q) {
    return guarded(ctx, [&] { nr_of(ctx).mismatch(stop_p, stop_q); return 0; });
}

int32_t jgb_nr_solve(jgb_ctx* ctx) {
    return guarded(ctx, [&] { nr_of(ctx).solve(); return 0; });
}

int32_t jgb_nr_get_vectors(jgb_ctx* ctx, double* mismatch, double* increment, double* j_nzval, int64_t* iteration) {
    return guarded(ctx, [&] { nr_of(ctx).get_vectors(mismatch, increment, j_nzval, iteration); return 0; });
}

int32_t jgb_nr_run(jgb_ctx* ctx, int64_t max_iter, double tol, int64_t* iterations, double* stop_p, double* stop_q) {
    return guarded(ctx, [&] {
        if (max_iter < 0) throw std::invalid_argument("nr_run: negative iteration cap");
        return nr_of(ctx).run(max_iter, tol, iterations, stop_p, stop_q);
    });
}

int32_t jgb_nr_set_branches(jgb_ctx* ctx, int64_t nbranch, const int64_t* from, const int64_t* to, const double* y_ff,
                            const double* y_ft, const double* y_tf, const double* y_tt, const int8_t* status) {
    return guarded(ctx, [&] { nr_of(ctx).set_branches(nbranch, from, to, y_ff, y_ft, y_tf, y_tt, status); return 0; });
}

int32_t jgb_nr_power(jgb_ctx* ctx, double* inj_p, double* inj_q, double* from_p, double* from_q, double* to_p,
                     double* to_q, double* from_im, double* from_ia, double* to_im, double* to_ia) {
    return guarded(ctx, [&] {
        double* out[10] = {inj_p, inj_q, from_p, from_q, to_p, to_q, from_im, from_ia, to_im, to_ia};
        nr_of(ctx).power(out);
        return 0;
    });
}

int32_t jgb_nr_batch(jgb_ctx* ctx, int64_t S, const int64_t* out_from, const int64_t* out_to, const double* dy,
                     int64_t max_iter, double tol, double* vm_out, double* va_out, int32_t* iterations,
                     int8_t* status, int64_t* total_iterations) {
    return guarded(ctx, [&] {
        if (!vm_out || !va_out) throw std::invalid_argument("nr_batch: null output");
        check_batch_args("nr_batch", S, max_iter, tol);
        return nr_of(ctx).batch(S, out_from, out_to, dy, false, max_iter, tol, vm_out, va_out, iterations, status,
                                false, total_iterations);
    });
}

int32_t jgb_nr_batch_dev(jgb_ctx* ctx, int64_t S, const int64_t* out_from, const int64_t* out_to, const double* dy,
                         int64_t max_iter, double tol, double* vm_out, double* va_out, int32_t* iterations,
                         int8_t* status, int64_t* total_iterations) {
    return guarded(ctx, [&] {
        if (!vm_out || !va_out || !iterations || !status) throw std::invalid_argument("nr_batch_dev: null output");
        check_batch_args("nr_batch_dev", S, max_iter, tol);
        return nr_of(ctx).batch(S, out_from, out_to, dy, true, max_iter, tol, vm_out, va_out, iterations, status,
                                true, total_iterations);
    });
}

#ifdef JGB_WITH_WLS
namespace {
jgb::WlsContext& wls_of(jgb_ctx* ctx) {
    if (!ctx->wls) throw std::logic_error("jgb_wls_setup has not been called on this context");
    return *ctx->wls;
}
}  // namespace

int32_t jgb_wls_setup(jgb_ctx* ctx, int64_t n, int64_t m, int64_t slack, const int64_t* h_colptr,
                      const int64_t* h_rowval, const int8_t* type, const int64_t* index, const int64_t* range6,
                      const int64_t* w_colptr, const int64_t* w_rowval, const double* w_nzval,
                      const int64_t* y_colptr, const int64_t* y_rowval, const double* y_nzval, const double* yt_nzval,
                      int64_t nbranch, const int64_t* from, const int64_t* to, const double* conductance,
                      const double* susceptance, const double* turns_ratio, const double* shift_angle,
                      const double* admittance) {
    return guarded(ctx, [&] {
        auto w = std::make_unique<jgb::WlsContext>(ctx->stream);
        w->setup(n, m, slack, h_colptr, h_rowval, type, index, range6, w_colptr, w_rowval, w_nzval, y_colptr,
                 y_rowval, y_nzval, yt_nzval, nbranch, from, to, conductance, susceptance, turns_ratio, shift_angle,
                 admittance);
        ctx->wls = std::move(w);
        return 0;
    });
}

int32_t jgb_wls_dims(jgb_ctx* ctx, int64_t* nnz_h, int64_t* nnz_g) {
    return guarded(ctx, [&] {
        auto& w = wls_of(ctx);
        if (nnz_h) *nnz_h = w.nnzh;
        if (nnz_g) *nnz_g = w.nnzg;
        return 0;
    });
}

int32_t jgb_wls_gain_pattern(jgb_ctx* ctx, int64_t* g_colptr, int64_t* g_rowval) {
    return guarded(ctx, [&] {
        auto& w = wls_of(ctx);
        if (g_colptr) std::copy(w.gcolptr1.begin(), w.gcolptr1.end(), g_colptr);
        if (g_rowval) std::copy(w.growval1.begin(), w.growval1.end(), g_rowval);
        return 0;
    });
}

int32_t jgb_wls_set_mean(jgb_ctx* ctx, const double* z) {
    return guarded(ctx, [&] { wls_of(ctx).set_mean(z); return 0; });
}

int32_t jgb_wls_set_state(jgb_ctx* ctx, const double* vm, const double* va) {
    return guarded(ctx, [&] { wls_of(ctx).set_state(vm, va); return 0; });
}

int32_t jgb_wls_get_state(jgb_ctx* ctx, double* vm, double* va) {
    return guarded(ctx, [&] {
        if (!vm || !va) throw std::invalid_argument("wls_get_state: null output");
        wls_of(ctx).get_state(vm, va);
        return 0;
    });
}

int32_t jgb_wls_increment(jgb_ctx* ctx, double* max_increment, double* objective) {
    return guarded(ctx, [&] { wls_of(ctx).increment(max_increment, objective); return 0; });
}

int32_t jgb_wls_solve(jgb_ctx* ctx) {
    return guarded(ctx, [&] { wls_of(ctx).solve(); return 0; });
}

int32_t jgb_wls_get_vectors(jgb_ctx* ctx, double* residual, double* increment, double* h_nzval, double* g_nzval,
                            int64_t* iteration) {
    return guarded(ctx, [&] { wls_of(ctx).get_vectors(residual, increment, h_nzval, g_nzval, iteration); return 0; });
}

int32_t jgb_wls_run(jgb_ctx* ctx, int64_t max_iter, double tol, int64_t* iterations, double* max_increment,
                    double* objective) {
    return guarded(ctx, [&] {
        if (max_iter < 0) throw std::invalid_argument("wls_run: negative iteration cap");
        return wls_of(ctx).run(max_iter, tol, iterations, max_increment, objective);
    });
}

int32_t jgb_wls_batch(jgb_ctx* ctx, int64_t S, const double* Z, int64_t max_iter, double tol, double* vm_out,
                      double* va_out, int32_t* iterations, int8_t* status, double* objective,
                      int64_t* total_iterations) {
    return guarded(ctx, [&] {
        if (!vm_out || !va_out) throw std::invalid_argument("wls_batch: null output");
        check_batch_args("wls_batch", S, max_iter, tol);
        return wls_of(ctx).batch(S, Z, false, max_iter, tol, vm_out, va_out, iterations, status, objective, false,
                                 total_iterations);
    });
}

int32_t jgb_wls_batch_dev(jgb_ctx* ctx, int64_t S, const double* Z, int64_t max_iter, double tol, double* vm_out,
                          double* va_out, int32_t* iterations, int8_t* status, double* objective,
                          int64_t* total_iterations) {
    return guarded(ctx, [&] {
        if (!vm_out || !va_out || !iterations || !status) throw std::invalid_argument("wls_batch_dev: null output");
        check_batch_args("wls_batch_dev", S, max_iter, tol);
        return wls_of(ctx).batch(S, Z, true, max_iter, tol, vm_out, va_out, iterations, status, objective, true,
                                 total_iterations);
    });
}
#endif  // JGB_WITH_WLS

int32_t jgb_profile(jgb_ctx* ctx, int32_t enable) {
    return guarded(ctx, [&] {
        if (ctx->nr) { ctx->nr->timer.enabled = enable != 0; ctx->nr->timer.reset(); }
#ifdef JGB_WITH_WLS
        if (ctx->wls) { ctx->wls->timer.enabled = enable != 0; ctx->wls->timer.reset(); }
#endif
        return 0;
    });
}

#ifdef JGB_WITH_WLS
int32_t jgb_wls_residual_test(jgb_ctx* ctx, double threshold, double* max_normalized_residual, int64_t* index,
                              double* c_out) {
    return guarded(ctx, [&] { wls_of(ctx).residual_test(threshold, max_normalized_residual, index, c_out); return 0; });
}

int32_t jgb_wls_update_rows(jgb_ctx* ctx, int64_t k, const int64_t* rows, const double* mean, const double* precision,
                            const double* precision_off, const int8_t* type, const int64_t* index) {
    return guarded(ctx, [&] {
        wls_of(ctx).update_rows(k, rows, mean, precision, precision_off, type, index);
        return 0;
    });
}

int32_t jgb_wls_update_y(jgb_ctx* ctx, int64_t k, const int64_t* nz_pos, const double* y_re_im, const double* yt_re_im) {
    return guarded(ctx, [&] {
        wls_of(ctx).update_y(k, nz_pos, y_re_im, yt_re_im);
        return 0;
    });
}

int32_t jgb_wls_update_branch(jgb_ctx* ctx, int64_t branch, double conductance, double susceptance, double turns_ratio,
                              double shift_angle, const double* admittance_re_im) {
    return guarded(ctx, [&] {
        wls_of(ctx).update_branch(branch, conductance, susceptance, turns_ratio, shift_angle, admittance_re_im);
        return 0;
    });
}

int32_t jgb_wls_remove_row(jgb_ctx* ctx, int64_t row) {
    return guarded(ctx, [&] { wls_of(ctx).remove_row(row); return 0; });
}
#endif

int32_t jgb_lin_setup(jgb_ctx* ctx, int64_t n, const int64_t* a_colptr, const int64_t* a_rowval, const double* a_nzval,
                      int64_t skip) {
    return guarded(ctx, [&] {
        auto lin = std::make_unique<jgb::LinContext>(ctx->stream);
        lin->setup(n, a_colptr, a_rowval, a_nzval, skip);
        ctx->lin = std::move(lin);
        return 0;
    });
}

static jgb::LinContext& lin_of(jgb_ctx* ctx) {
    if (!ctx->lin) throw std::logic_error("jgb_lin_setup has not been called on this context");
    return *ctx->lin;
}

int32_t jgb_lin_refactor(jgb_ctx* ctx, const double* a_nzval) {
    return guarded(ctx, [&] { lin_of(ctx).refactor(a_nzval); return 0; });
}

int32_t jgb_lin_projection(jgb_ctx* ctx, int64_t m, const int64_t* p_colptr, const int64_t* p_rowval,
                           const double* p_nzval) {
    return guarded(ctx, [&] { lin_of(ctx).set_projection(m, p_colptr, p_rowval, p_nzval); return 0; });
}

int32_t jgb_lin_solve(jgb_ctx* ctx, int64_t R, const double* b, double* x) {
    return guarded(ctx, [&] { lin_of(ctx).solve(R, b, false, x, false, false); return 0; });
}

int32_t jgb_lin_solve_projected(jgb_ctx* ctx, int64_t R, const double* z, double* x) {
    return guarded(ctx, [&] { lin_of(ctx).solve(R, z, false, x, false, true); return 0; });
}

int32_t jgb_lin_solve_dev(jgb_ctx* ctx, int64_t R, const double* in_dev, double* x_dev, int32_t projected) {
    return guarded(ctx, [&] { lin_of(ctx).solve(R, in_dev, true, x_dev, true, projected != 0); return 0; });
}

int32_t jgb_lin_dims(jgb_ctx* ctx, int64_t* n, int64_t* m, int64_t* nnz_factor, int64_t* fronts) {
    return guarded(ctx, [&] {
        jgb::LinContext& l = lin_of(ctx);
        if (n) *n = l.n;
        if (m) *m = l.m;
        if (nnz_factor) *nnz_factor = l.nnz_factor();
        if (fronts) *fronts = l.nfronts();
        return 0;
    });
}

static jgb::FnrContext& fnr_of(jgb_ctx* ctx) {
    if (!ctx->fnr) throw std::logic_error("jgb_fnr_setup has not been called on this context");
    return *ctx->fnr;
}

int32_t jgb_fnr_setup(jgb_ctx* ctx, int64_t n, const int64_t* y_colptr, const int64_t* y_rowval,
                      const double* yT_nzval_re_im, const int8_t* bus_type, int64_t slack,
                      const int64_t* bp_colptr, const int64_t* bp_rowval, const double* bp_nzval,
                      const int64_t* bq_colptr, const int64_t* bq_rowval, const double* bq_nzval) {
    return guarded(ctx, [&] {
        auto f = std::make_unique<jgb::FnrContext>(ctx->stream);
        f->setup(n, y_colptr, y_rowval, yT_nzval_re_im, bus_type, slack, bp_colptr, bp_rowval, bp_nzval, bq_colptr,
                 bq_rowval, bq_nzval);
        ctx->fnr = std::move(f);
        return 0;
    });
}

int32_t jgb_fnr_set_injection(jgb_ctx* ctx, const double* ps, const double* qs, const double* pd, const double* qd) {
    return guarded(ctx, [&] { fnr_of(ctx).set_injection(ps, qs, pd, qd); return 0; });
}

int32_t jgb_fnr_set_state(jgb_ctx* ctx, const double* vm, const double* va) {
    return guarded(ctx, [&] { fnr_of(ctx).set_state(vm, va); return 0; });
}

int32_t jgb_fnr_get_state(jgb_ctx* ctx, double* vm, double* va) {
    return guarded(ctx, [&] {
        if (!vm || !va) throw std::invalid_argument("fnr_get_state: null output");
        fnr_of(ctx).get_state(vm, va);
        return 0;
    });
}

int32_t jgb_fnr_mismatch(jgb_ctx* ctx, double* stop_p, double* stop_q) {
    return guarded(ctx, [&] { fnr_of(ctx).mismatch(stop_p, stop_q); return 0; });
}

int32_t jgb_fnr_solve(jgb_ctx* ctx) {
    return guarded(ctx, [&] { fnr_of(ctx).solve(); return 0; });
}

int32_t jgb_fnr_run(jgb_ctx* ctx, int64_t max_iter, double tol, int64_t* iterations, double* stop_p, double* stop_q) {
    return guarded(ctx, [&] {
        if (max_iter < 0 || !(tol > 0)) throw std::invalid_argument("fnr_run: max_iter >= 0 and tol > 0 required");
        return fnr_of(ctx).run(max_iter, tol, iterations, stop_p, stop_q);
    });
}

int32_t jgb_fnr_batch(jgb_ctx* ctx, int64_t R, const double* p_inj, const double* q_inj, int64_t max_iter, double tol,
                      double* vm_out, double* va_out, int32_t* iterations, int8_t* status, int64_t* total_iterations) {
    return guarded(ctx, [&] {
        check_batch_args("fnr_batch", R, max_iter, tol);
        return fnr_of(ctx).batch(R, p_inj, q_inj, max_iter, tol, vm_out, va_out, iterations, status, total_iterations);
    });
}


/* ---- multi-GPU ------------------------------------------------------------------------------------------------ */
int32_t jgb_comm_unique_id(uint8_t* id128) {
    try {
        if (!id128) return -1;
        jgb::CommContext::unique_id(id128);
        return 0;
    } catch (const std::exception& e) {
        g_create_error = e.what();
        return -4;
    }
}

int32_t jgb_comm_init(jgb_ctx* ctx, int32_t rank, int32_t nranks, const uint8_t* id128) {
    return guarded(ctx, [&] {
        if (!ctx->comm) ctx->comm = std::make_unique<jgb::CommContext>(ctx->stream);
        ctx->comm->init(rank, nranks, id128);
        return 0;
    });
}

int32_t jgb_allgather_states(jgb_ctx* ctx, int64_t rows_local, int64_t n, const double* vm_dev, const double* va_dev,
                             const int32_t* iterations_dev, const int8_t* status_dev, double* vm_all_dev,
                             double* va_all_dev, int32_t* iterations_all_dev, int8_t* status_all_dev) {
    return guarded(ctx, [&] {
        if (!ctx->comm) throw std::logic_error("allgather_states: jgb_comm_init has not been called");
        ctx->comm->allgather_states(rows_local, n, vm_dev, va_dev, iterations_dev, status_dev, vm_all_dev, va_all_dev,
                                    iterations_all_dev, status_all_dev);
        return 0;
    });
}

int32_t jgb_comm_wait(jgb_ctx* ctx, int32_t host_blocking) {
    return guarded(ctx, [&] {
        if (ctx->comm) ctx->comm->wait(host_blocking != 0);
        return 0;
    });
}

int32_t jgb_comm_size(jgb_ctx* ctx, int32_t* rank, int32_t* nranks) {
    return guarded(ctx, [&] {
        if (rank) *rank = ctx->comm ? ctx->comm->rank : -1;
        if (nranks) *nranks = ctx->comm ? ctx->comm->nranks : 0;
        return 0;
    });
}

double jgb_stat(jgb_ctx* ctx, const char* key) {
    if (!ctx || !key) return -1.0;
    try {
        std::string k(key);
        if (k == "launches") {
            double t = 0;
            if (ctx->nr) t += (double)ctx->nr->launches;
#ifdef JGB_WITH_WLS
            if (ctx->wls) t += (double)ctx->wls->launches;
#endif
            return t;
        }
        if (k == "comm.calls") return ctx->comm ? (double)ctx->comm->calls : 0.0;
        if (k == "comm.nranks") return ctx->comm ? (double)ctx->comm->nranks : 0.0;
        if (k.rfind("nr.", 0) == 0 && ctx->nr) return ctx->nr->stat(k);
#ifdef JGB_WITH_WLS
        if (k.rfind("wls.", 0) == 0 && ctx->wls) return ctx->wls->stat(k);
#endif
    } catch (...) {
    }
    return -1.0;
}

int32_t jgb_selfcheck_symbolic(int64_t n, const int64_t* colptr, const int64_t* rowval, const double* nzval,
                               const int64_t* group, const double* rhs, double* x, double* stats8) {
    try {
        if (n <= 0 || !colptr || !rowval) return -1;
        std::vector<int> cp(n + 1), rv(colptr[n] - 1), grp;
        for (int64_t i = 0; i <= n; ++i) cp[i] = (int)(colptr[i] - 1);
        for (size_t q = 0; q < rv.size(); ++q) rv[q] = (int)(rowval[q] - 1);
        if (group) {
            grp.resize(n);
            for (int64_t i = 0; i < n; ++i) grp[i] = (int)group[i];
        }
        jgb::Symbolic s;
        jgb::analyse((int)n, cp.data(), rv.data(), group ? grp.data() : nullptr, nullptr, jgb::SymbolicOptions(), s);
        if (stats8) {
            stats8[0] = s.nfronts; stats8[1] = s.nlevels; stats8[2] = s.ndepths; stats8[3] = (double)s.nnz_lu;
            stats8[4] = s.flops; stats8[5] = s.max_front; stats8[6] = (double)s.u_size; stats8[7] = (double)s.upd_size;
        }
        if (nzval && rhs && x) return jgb::host_factor_solve(s, nzval, rhs, x);
        return 0;
    } catch (...) {
        return -4;
    }
}

int32_t jgb_selfcheck_tree(int64_t n, const int64_t* colptr, const int64_t* rowval, const int64_t* group, int32_t preset,
                           int64_t cap, int64_t* nfronts, int32_t* f_k, int32_t* f_nf, int32_t* f_parent,
                           int32_t* f_nasm) {
    try {
        if (n <= 0 || !colptr || !rowval || !nfronts) return -1;
        std::vector<int> cp(n + 1), rv(colptr[n] - 1), grp;
        for (int64_t i = 0; i <= n; ++i) cp[i] = (int)(colptr[i] - 1);
        for (size_t q = 0; q < rv.size(); ++q) rv[q] = (int)(rowval[q] - 1);
        if (group) {
            grp.resize(n);
            for (int64_t i = 0; i < n; ++i) grp[i] = (int)group[i];
        }
        jgb::Symbolic s;
        const jgb::SymbolicOptions opt = preset == 1 ? jgb::latency_options()
                                         : preset == 2 ? jgb::throughput_options() : jgb::SymbolicOptions();
        jgb::analyse((int)n, cp.data(), rv.data(), group ? grp.data() : nullptr, nullptr, opt, s);
        *nfronts = s.nfronts;
        if (s.nfronts > cap) return -1;
        for (int f = 0; f < s.nfronts; ++f) {
            if (f_k) f_k[f] = s.f_k[f];
            if (f_nf) f_nf[f] = s.f_nf[f];
            if (f_parent) f_parent[f] = s.f_parent[f];
            if (f_nasm) f_nasm[f] = s.f_asmptr[f + 1] - s.f_asmptr[f];
        }
        return 0;
    } catch (...) {
        return -4;
    }
}

int32_t jgb_selfcheck_tasks(int64_t n, const int64_t* colptr, const int64_t* rowval, const double* nzval,
                            const int64_t* group, const double* rhs, double* x, double* stats8) {
    try {
        if (n <= 0 || !colptr || !rowval) return -1;
        std::vector<int> cp(n + 1), rv(colptr[n] - 1), grp;
        for (int64_t i = 0; i <= n; ++i) cp[i] = (int)(colptr[i] - 1);
        for (size_t q = 0; q < rv.size(); ++q) rv[q] = (int)(rowval[q] - 1);
        if (group) {
            grp.resize(n);
            for (int64_t i = 0; i < n; ++i) grp[i] = (int)group[i];
        }
        jgb::Symbolic s;
        jgb::analyse((int)n, cp.data(), rv.data(), group ? grp.data() : nullptr, nullptr, jgb::throughput_options(), s);
        jgb::TaskPlan tp;
        jgb::partition_tasks(s, jgb::task_options_from_env(), tp);
        if (stats8) {
            size_t smem = 0;
            for (auto& tl : tp.launches) smem = std::max(smem, tl.smem);
            stats8[0] = (double)tp.launches.size(); stats8[1] = tp.task_count; stats8[2] = tp.task_fronts;
            stats8[3] = (double)tp.upd_on_chip; stats8[4] = (double)smem; stats8[5] = (double)tp.blob.size();
            stats8[6] = s.nfronts; stats8[7] = (double)s.upd_size;
        }
        if (nzval && rhs && x) return jgb::host_task_factor_solve(s, tp, nzval, rhs, x);
        return 0;
    } catch (...) {
        return -4;
    }
}

}  // extern "C"
