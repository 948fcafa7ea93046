// See lin.cuh.
#include "lin.cuh"

#include <algorithm>
#include <cmath>
#include <stdexcept>

namespace jgb {

namespace {

// [R][rows] (each vector contiguous) -> [rows][Rp] (vector minor, zero padded to Rp columns) and back
__global__ void lin_transpose_in_kernel(const double* __restrict__ src, double* __restrict__ dst, int rows, int Rp,
                                        int R) {
    __shared__ double tile[32][33];
    const int i0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int q = threadIdx.y; q < 32; q += blockDim.y) {
        const int r = r0 + q, i = i0 + threadIdx.x;
        tile[q][threadIdx.x] = (r < R && i < rows) ? src[(long long)r * rows + i] : 0.0;
    }
    __syncthreads();
    for (int q = threadIdx.y; q < 32; q += blockDim.y) {
        const int i = i0 + q, r = r0 + threadIdx.x;
        if (i < rows && r < Rp) dst[(long long)i * Rp + r] = tile[threadIdx.x][q];
    }
}

__global__ void lin_transpose_out_kernel(const double* __restrict__ src, double* __restrict__ dst, int rows, int Rp,
                                         int R) {
    __shared__ double tile[32][33];
    const int i0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int q = threadIdx.y; q < 32; q += blockDim.y) {
        const int i = i0 + q, r = r0 + threadIdx.x;
        if (i < rows && r < Rp) tile[q][threadIdx.x] = src[(long long)i * Rp + r];
    }
    __syncthreads();
    for (int q = threadIdx.y; q < 32; q += blockDim.y) {
        const int r = r0 + q, i = i0 + threadIdx.x;
        if (r < R && i < rows) dst[(long long)r * rows + i] = tile[threadIdx.x][q];
    }
}

// b[i][r] = sum over the entries of column i of P = W H (measurement rows in ascending order, so the sum order
// is fixed): one thread per (state i, vector r), 256-byte lines of Z per entry.
__global__ void lin_project_kernel(int n, const int* __restrict__ cp, const int* __restrict__ rv,
                                   const double* __restrict__ val, const double* __restrict__ Z,
                                   double* __restrict__ B, int Rp) {
    const int i = blockIdx.x * blockDim.y + threadIdx.y;
    const int r = blockIdx.y * 32 + threadIdx.x;
    if (i >= n) return;
    double acc = 0.0;
    for (int t = cp[i]; t < cp[i + 1]; ++t) acc += val[t] * Z[(long long)rv[t] * Rp + r];
    B[(long long)i * Rp + r] = acc;
}

}  // namespace

void LinContext::setup(int64_t n_, const int64_t* cp, const int64_t* rv, const double* av, int64_t skip_) {
    if (n_ <= 0 || !cp || !rv || !av) throw std::invalid_argument("lin_setup: null or empty input");
    if (n_ > (1 << 30)) throw std::invalid_argument("lin_setup: matrix too large");
    if (skip_ < 0 || skip_ > n_) throw std::invalid_argument("lin_setup: skip index out of range");
    if (cp[0] != 1) throw std::invalid_argument("lin_setup: colptr must be 1-based");
    n = n_;
    skip = (int)skip_ - 1;
    nnz_in = (size_t)(cp[n] - 1);
    const int nn = (int)n;
    for (int c = 0; c < nn; ++c) {
        if (cp[c + 1] < cp[c]) throw std::invalid_argument("lin_setup: colptr not monotone");
        for (int64_t t = cp[c] - 1; t < cp[c + 1] - 1; ++t) {
            if (rv[t] < 1 || rv[t] > n) throw std::invalid_argument("lin_setup: row index out of range");
            if (t > cp[c] - 1 && rv[t] <= rv[t - 1]) throw std::invalid_argument("lin_setup: rows not sorted");
        }
    }
    // the pattern must be symmetric (fixed elimination tree, no pivoting)
    auto find = [&](int r, int c) -> int64_t {
        const int64_t* b = rv + cp[c] - 1;
        const int64_t* e = rv + cp[c + 1] - 1;
        const int64_t* it = std::lower_bound(b, e, (int64_t)r + 1);
        return (it != e && *it == r + 1) ? (it - rv) : -1;
    };
    for (int c = 0; c < nn; ++c)
        for (int64_t t = cp[c] - 1; t < cp[c + 1] - 1; ++t) {
            const int r = (int)rv[t] - 1;
            if (r != c && find(c, r) < 0) throw std::invalid_argument("lin_setup: pattern is not symmetric");
        }
    // transpose partner of every stored entry (values may be unsymmetric: fast Newton-Raphson B' with phase shifters)
    tpos.assign(nnz_in, -1);
    for (int c = 0; c < nn; ++c)
        for (int64_t t = cp[c] - 1; t < cp[c + 1] - 1; ++t) tpos[t] = find(c, (int)rv[t] - 1);
    // analysed pattern: the skip row/column reduced to a unit diagonal
    std::vector<int> colptr(nn + 1, 0), rowidx;
    slot.assign(nnz_in, -1);
    vals.clear();
    skip_diag = -1;
    for (int c = 0; c < nn; ++c) {
        colptr[c] = (int)rowidx.size();
        bool diag_done = false;
        for (int64_t t = cp[c] - 1; t < cp[c + 1] - 1; ++t) {
            const int r = (int)rv[t] - 1;
            if (skip >= 0 && (r == skip || c == skip)) {
                if (c == skip && r == skip) {
                    skip_diag = (int)rowidx.size();
                    rowidx.push_back(r);
                    vals.push_back(1.0);
                    diag_done = true;
                }
                continue;
            }
            slot[t] = (int)rowidx.size();
            rowidx.push_back(r);
            vals.push_back(av[t]);
        }
        if (c == skip && !diag_done) {
            skip_diag = (int)rowidx.size();
            rowidx.push_back(c);
            vals.push_back(1.0);
        }
    }
    colptr[nn] = (int)rowidx.size();
    vals_t = vals;
    for (size_t t = 0; t < nnz_in; ++t)
        if (slot[t] >= 0) vals_t[slot[t]] = av[tpos[t]];
    Symbolic sym;
    analyse(nn, colptr.data(), rowidx.data(), nullptr, nullptr, latency_options(), sym);
    symmetric = values_symmetric(av);
    solver.setup(sym, stream, symmetric);
    if (!symmetric) solver_t.setup(sym, stream, false);
    std::vector<double> zero(nn, 0.0);
    d_zero.upload(zero, stream);
    d_x0.alloc(nn);
    d_status.alloc(1);
    h_status.alloc(1);
    JGB_CUDA(cudaStreamSynchronize(stream));
    m = 0;
    factor_now();
}

bool LinContext::values_symmetric(const double* av) const {
    for (size_t t = 0; t < nnz_in; ++t) {
        const double a = av[t], b = av[tpos[t]];
        if (std::fabs(a - b) > 1e-12 * std::max(std::fabs(a), std::fabs(b))) return false;
    }
    return true;
}

void LinContext::refactor(const double* av) {
    if (n == 0) throw std::logic_error("jgb_lin_setup has not been called on this context");
    if (!av) throw std::invalid_argument("lin_refactor: null values");
    if (symmetric && !values_symmetric(av))
        throw std::invalid_argument("lin_refactor: the matrix was set up as symmetric; the new values are not");
    for (size_t t = 0; t < nnz_in; ++t)
        if (slot[t] >= 0) {
            vals[slot[t]] = av[t];
            if (!symmetric) vals_t[slot[t]] = av[tpos[t]];
        }
    factor_now();
}

void LinContext::factor_now() {
    d_aval.upload(vals, stream);
    JGB_CUDA(cudaMemsetAsync(d_status.p, 0, sizeof(int), stream));
    // numeric LDL^T (S = 1); the zero right-hand side rides along and is discarded
    solver.factor_solve(d_aval.p, d_zero.p, d_x0.p, 1, nullptr, d_status.p, stream);
    JGB_CUDA(cudaMemcpyAsync(h_status.p, d_status.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
    JGB_CUDA(cudaStreamSynchronize(stream));
    if (h_status.p[0] != 0) throw std::domain_error("lin: zero or non-finite pivot (singular matrix)");
    if (!symmetric) {          // the transposed matrix on the same elimination tree supplies L
        d_aval_t.upload(vals_t, stream);
        solver_t.factor_solve(d_aval_t.p, d_zero.p, d_x0.p, 1, nullptr, d_status.p, stream);
        JGB_CUDA(cudaMemcpyAsync(h_status.p, d_status.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
        JGB_CUDA(cudaStreamSynchronize(stream));
        if (h_status.p[0] != 0) throw std::domain_error("lin: zero or non-finite pivot (singular matrix)");
    }
}

void LinContext::solve_block(double* B, int Rp) { solver.solve_multi(B, Rp, stream, symmetric ? nullptr : &solver_t); }

void LinContext::set_projection(int64_t m_, const int64_t* cp, const int64_t* rv, const double* pv) {
    if (n == 0) throw std::logic_error("jgb_lin_setup has not been called on this context");
    if (m_ <= 0 || !cp || !rv || !pv) throw std::invalid_argument("lin_projection: null or empty input");
    if (cp[0] != 1) throw std::invalid_argument("lin_projection: colptr must be 1-based");
    const size_t nz = (size_t)(cp[n] - 1);
    std::vector<int> c32(n + 1), r32(nz);
    for (int64_t c = 0; c <= n; ++c) c32[c] = (int)(cp[c] - 1);
    for (size_t t = 0; t < nz; ++t) {
        if (rv[t] < 1 || rv[t] > m_) throw std::invalid_argument("lin_projection: row index out of range");
        r32[t] = (int)rv[t] - 1;
    }
    if (skip >= 0 && c32[skip + 1] != c32[skip]) {
        // the reference removes the slack column of H before forming H' W (removeColumn, sparse.jl:155-163)
        std::vector<double> v(pv, pv + nz);
        for (int t = c32[skip]; t < c32[skip + 1]; ++t) v[t] = 0.0;
        d_pval.upload(v, stream);
    } else {
        d_pval.upload(pv, nz, stream);
    }
    d_pcolptr.upload(c32, stream);
    d_prow.upload(r32, stream);
    JGB_CUDA(cudaStreamSynchronize(stream));
    m = m_;
}

void LinContext::solve(int64_t R, const double* in, bool dev_in, double* out, bool dev_out, bool projected) {
    if (n == 0) throw std::logic_error("jgb_lin_setup has not been called on this context");
    if (projected && m == 0) throw std::logic_error("jgb_lin_projection has not been called on this context");
    if (R <= 0 || !in || !out) throw std::invalid_argument("lin_solve: null or empty input");
    if (R > 65535LL * 32) throw std::invalid_argument("lin_solve: too many right-hand sides");
    const int Rp = (int)((R + 31) / 32 * 32), nn = (int)n;
    const int rows_in = projected ? (int)m : nn;
    const double* din = in;
    if (!dev_in) {
        d_in.alloc((size_t)R * rows_in);
        JGB_CUDA(cudaMemcpyAsync(d_in.p, in, (size_t)R * rows_in * sizeof(double), cudaMemcpyHostToDevice, stream));
        din = d_in.p;
    }
    d_B.alloc((size_t)nn * Rp);
    const dim3 tb(32, 8);
    if (projected) {
        d_Z.alloc((size_t)m * Rp);
        lin_transpose_in_kernel<<<dim3(ceil_div((int)m, 32), Rp / 32), tb, 0, stream>>>(din, d_Z.p, (int)m, Rp, (int)R);
        lin_project_kernel<<<dim3(ceil_div(nn, 8), Rp / 32), tb, 0, stream>>>(nn, d_pcolptr.p, d_prow.p, d_pval.p,
                                                                                d_Z.p, d_B.p, Rp);
    } else {
        lin_transpose_in_kernel<<<dim3(ceil_div(nn, 32), Rp / 32), tb, 0, stream>>>(din, d_B.p, nn, Rp, (int)R);
    }
    solve_block(d_B.p, Rp);
    double* dout = out;
    if (!dev_out) {
        d_out.alloc((size_t)R * nn);
        dout = d_out.p;
    }
    lin_transpose_out_kernel<<<dim3(ceil_div(nn, 32), Rp / 32), tb, 0, stream>>>(d_B.p, dout, nn, Rp, (int)R);
    JGB_CUDA(cudaGetLastError());
    if (!dev_out) JGB_CUDA(cudaMemcpyAsync(out, dout, (size_t)R * nn * sizeof(double), cudaMemcpyDeviceToHost, stream));
    JGB_CUDA(cudaStreamSynchronize(stream));
}

}  // namespace jgb
