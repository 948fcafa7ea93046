// See fnr.cuh.
#include "fnr.cuh"

#include <algorithm>
#include <stdexcept>

namespace jgb {

namespace {

// mismatch! (acPowerFlow.jl:686-727) / the reactive sweep inside solve! (:958-970): one thread per (bus, scenario)
// walks the Ybus column strip of the bus with the transpose values, i.e. row i of Y.
__global__ void __launch_bounds__(128)
fnr_mismatch_kernel(FnrDev d, int q_only) {
    __shared__ double red[2][128];
    const int r = blockIdx.y * 32 + threadIdx.x;
    const int i = blockIdx.x * 4 + threadIdx.y;
    double ap = 0.0, aq = 0.0;
    if (i < d.n && i != d.slack && (!d.active || d.active[r])) {
        const bool is_pq = d.type[i] == 1;
        if (!q_only || is_pq) {
            const double Ti = d.va[(size_t)i * d.R + r], Vi = d.vm[(size_t)i * d.R + r];
            double cur_p = 0.0, cur_q = 0.0;
            for (int p = d.ycolptr[i]; p < d.ycolptr[i + 1]; ++p) {
                const int j = d.yrow[p];
                const double2 y = d.yt[p];
                double sn, cs;
                sincos(Ti - d.va[(size_t)j * d.R + r], &sn, &cs);
                const double Vj = d.vm[(size_t)j * d.R + r];
                cur_p += Vj * (y.x * cs + y.y * sn);
                cur_q += Vj * (y.x * sn - y.y * cs);
            }
            const double vinv = 1.0 / Vi;
            if (!q_only) {
                const double f = cur_p - d.pinj[(size_t)i * d.R + r] * vinv;
                d.mp[(size_t)d.pvpq[i] * d.R + r] = f;
                ap = (f != f) ? INFINITY : fabs(f);      // fmax() drops NaN operands: a diverged run must not read as converged
            }
            if (is_pq) {
                const double f = q_only ? cur_q - d.qinj[(size_t)i * d.R + r] / Vi : cur_q - d.qinj[(size_t)i * d.R + r] * vinv;
                d.mq[(size_t)d.pq[i] * d.R + r] = f;
                aq = (f != f) ? INFINITY : fabs(f);
            }
        }
    }
    if (q_only) return;
    const int t = threadIdx.y * 32 + threadIdx.x;
    red[0][t] = ap;
    red[1][t] = aq;
    __syncthreads();
    if (threadIdx.y == 0) {
        double mp = 0.0, mq = 0.0;
        for (int q = 0; q < 4; ++q) { mp = fmax(mp, red[0][q * 32 + threadIdx.x]); mq = fmax(mq, red[1][q * 32 + threadIdx.x]); }
        atomicMax(&d.stopbits[r], (unsigned long long)__double_as_longlong(mp));
        atomicMax(&d.stopbits[d.R + r], (unsigned long long)__double_as_longlong(mq));
    }
}

// powerFlow! bookkeeping (acPowerFlow.jl:1406-1418); tol < 0: publish the stop values only (mismatch! operator)
__global__ void fnr_check_kernel(FnrDev d, int Rreal, double tol, int max_iter) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= d.R) return;
    if (d.active && !d.active[r]) return;
    const double sp = __longlong_as_double((long long)d.stopbits[r]);
    const double sq = __longlong_as_double((long long)d.stopbits[d.R + r]);
    d.stop[r] = sp;
    d.stop[d.R + r] = sq;
    d.stopbits[r] = 0ull;
    d.stopbits[d.R + r] = 0ull;
    if (tol < 0.0) return;
    if (r >= Rreal) { d.active[r] = 0; return; }
    if (sp < tol && sq < tol) { d.active[r] = 0; d.status[r] = 0; return; }
    if (!(sp <= 1.79e308) || !(sq <= 1.79e308)) { d.active[r] = 0; d.status[r] = -3; return; }   // NaN or Inf mismatch: diverged
    if (d.iters[r] == max_iter) { d.active[r] = 0; d.status[r] = 1; return; }
    atomicAdd(d.remaining, 1);
}

// solve!: angle += active.increment (acPowerFlow.jl:952-956) / magnitude += reactive.increment (:974-978)
__global__ void fnr_update_kernel(FnrDev d, int magnitude) {
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int i = (int)(gid / d.R), r = (int)(gid % d.R);
    if (i >= d.n) return;
    if (d.active && !d.active[r]) return;
    if (magnitude) {
        if (d.type[i] == 1) d.vm[gid] += d.mq[(size_t)d.pq[i] * d.R + r];
        if (i == 0) d.iters[r] += 1;
    } else if (i != d.slack) {
        d.va[gid] += d.mp[(size_t)d.pvpq[i] * d.R + r];
    }
}

__global__ void fnr_broadcast_kernel(const double* __restrict__ src, double* __restrict__ dst, int n, int R) {
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)n * R) return;
    dst[gid] = src[gid / R];
}

__global__ void fnr_transpose_in_kernel(const double* __restrict__ src, double* __restrict__ dst, int rows, int Rp, int R) {
    __shared__ double tile[32][33];
    const int i0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int q = threadIdx.y; q < 32; q += blockDim.y) {
        const int r = r0 + q, i = i0 + threadIdx.x;
        tile[q][threadIdx.x] = (r < R && i < rows) ? src[(size_t)r * rows + i] : 0.0;
    }
    __syncthreads();
    for (int q = threadIdx.y; q < 32; q += blockDim.y) {
        const int i = i0 + q, r = r0 + threadIdx.x;
        if (i < rows && r < Rp) dst[(size_t)i * Rp + r] = tile[threadIdx.x][q];
    }
}

__global__ void fnr_transpose_out_kernel(const double* __restrict__ src, double* __restrict__ dst, int rows, int Rp, int R) {
    __shared__ double tile[32][33];
    const int i0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int q = threadIdx.y; q < 32; q += blockDim.y) {
        const int i = i0 + q, r = r0 + threadIdx.x;
        if (i < rows && r < Rp) tile[q][threadIdx.x] = src[(size_t)i * Rp + r];
    }
    __syncthreads();
    for (int q = threadIdx.y; q < 32; q += blockDim.y) {
        const int r = r0 + q, i = i0 + threadIdx.x;
        if (r < R && i < rows) dst[(size_t)r * rows + i] = tile[threadIdx.x][q];
    }
}

}  // namespace

void FnrContext::setup(int64_t n_, const int64_t* ycp, const int64_t* yrv, const double* yt, const int8_t* type,
                       int64_t slack_, const int64_t* bpcp, const int64_t* bprv, const double* bpnz,
                       const int64_t* bqcp, const int64_t* bqrv, const double* bqnz) {
    if (n_ <= 1 || !ycp || !yrv || !yt || !type || !bpcp || !bprv || !bpnz || !bqcp || !bqrv || !bqnz)
        throw std::invalid_argument("fnr_setup: null or empty input");
    if (slack_ < 1 || slack_ > n_) throw std::invalid_argument("fnr_setup: slack index out of range");
    n = (int)n_;
    slack = (int)slack_ - 1;
    if (type[slack] != 3) throw std::invalid_argument("fnr_setup: the slack bus must have type 3");
    const int nnzy = (int)(ycp[n] - 1);
    std::vector<int> ycolptr(n + 1), yrow(nnzy), pq(n, -1), pvpq(n, -1);
    for (int i = 0; i <= n; ++i) ycolptr[i] = (int)(ycp[i] - 1);
    for (int q = 0; q < nnzy; ++q) {
        if (yrv[q] < 1 || yrv[q] > n) throw std::invalid_argument("fnr_setup: Ybus row index out of range");
        yrow[q] = (int)yrv[q] - 1;
    }
    npq = 0;
    int nps = 0;
    for (int i = 0; i < n; ++i) {            // fastNewtonJacobian (acPowerFlow.jl:345-358)
        if (type[i] < 1 || type[i] > 3) throw std::invalid_argument("fnr_setup: bus type must be 1, 2 or 3");
        if (type[i] == 3 && i != slack) throw std::invalid_argument("fnr_setup: more than one slack bus");
        if (type[i] == 1) pq[i] = npq++;
        if (type[i] != 3) pvpq[i] = nps++;
    }
    if (npq == 0) throw std::invalid_argument("fnr_setup: no demand (PQ) bus");
    if (bpcp[n - 1] < 1 || bqcp[npq] < 1) throw std::invalid_argument("fnr_setup: bad Jacobian pattern");
    active_lin.setup(n - 1, bpcp, bprv, bpnz, 0);
    reactive_lin.setup(npq, bqcp, bqrv, bqnz, 0);
    d_ycolptr.upload(ycolptr, stream); d_yrow.upload(yrow, stream);
    d_yt.upload(reinterpret_cast<const double2*>(yt), nnzy, stream);
    d_type.upload(reinterpret_cast<const signed char*>(type), n, stream);
    d_pq.upload(pq, stream); d_pvpq.upload(pvpq, stream);
    d_remaining.alloc(1);
    h_stop.alloc(2);
    h_int.alloc(2);
    JGB_CUDA(cudaStreamSynchronize(stream));
    R = 0;
    have_injection = have_state = false;
    iteration = 0;
}

void FnrContext::alloc(int Rp) {
    if (Rp == R) return;
    const size_t r = Rp;
    d_pinj.alloc(n * r); d_qinj.alloc(n * r); d_vm.alloc(n * r); d_va.alloc(n * r);
    d_mp.alloc((n - 1) * r); d_mq.alloc(npq * r); d_stop.alloc(2 * r); d_stopbits.alloc(2 * r);
    d_active.alloc(r); d_status.alloc(r); d_iters.alloc(r);
    JGB_CUDA(cudaMemsetAsync(d_stopbits.p, 0, 2 * r * sizeof(unsigned long long), stream));
    R = Rp;
    have_injection = false;
    if (have_state) broadcast_state();
}

void FnrContext::broadcast_state() {
    DevBuf<double> a, b;
    a.upload(h_vm, stream);
    b.upload(h_va, stream);
    const int blocks = (int)(((size_t)n * R + 255) / 256);
    fnr_broadcast_kernel<<<blocks, 256, 0, stream>>>(a.p, d_vm.p, n, R);
    fnr_broadcast_kernel<<<blocks, 256, 0, stream>>>(b.p, d_va.p, n, R);
    launches += 2;
    JGB_CUDA(cudaStreamSynchronize(stream));
}

FnrDev FnrContext::view() {
    FnrDev d{};
    d.n = n; d.slack = slack; d.npq = npq; d.R = R;
    d.ycolptr = d_ycolptr.p; d.yrow = d_yrow.p; d.yt = d_yt.p; d.type = d_type.p; d.pq = d_pq.p; d.pvpq = d_pvpq.p;
    d.pinj = d_pinj.p; d.qinj = d_qinj.p; d.vm = d_vm.p; d.va = d_va.p; d.mp = d_mp.p; d.mq = d_mq.p;
    d.stopbits = d_stopbits.p; d.stop = d_stop.p; d.active = d_active.p; d.status = d_status.p; d.iters = d_iters.p;
    d.remaining = d_remaining.p;
    return d;
}

void FnrContext::set_injection(const double* ps, const double* qs, const double* pd, const double* qd) {
    if (!n) throw std::logic_error("jgb_fnr_setup has not been called on this context");
    if (!ps || !qs || !pd || !qd) throw std::invalid_argument("fnr_set_injection: null input");
    alloc(32);
    std::vector<double> p(n), q(n);
    for (int i = 0; i < n; ++i) { p[i] = ps[i] - pd[i]; q[i] = qs[i] - qd[i]; }
    d_io.upload(p, stream);
    d_io2.upload(q, stream);
    const int blocks = (int)(((size_t)n * R + 255) / 256);
    fnr_broadcast_kernel<<<blocks, 256, 0, stream>>>(d_io.p, d_pinj.p, n, R);
    fnr_broadcast_kernel<<<blocks, 256, 0, stream>>>(d_io2.p, d_qinj.p, n, R);
    launches += 2;
    JGB_CUDA(cudaStreamSynchronize(stream));
    have_injection = true;
}

void FnrContext::set_state(const double* vm, const double* va) {
    if (!n) throw std::logic_error("jgb_fnr_setup has not been called on this context");
    if (!vm || !va) throw std::invalid_argument("fnr_set_state: null input");
    h_vm.assign(vm, vm + n);
    h_va.assign(va, va + n);
    if (R == 0) alloc(32);
    broadcast_state();
    have_state = true;
}

void FnrContext::get_state(double* vm, double* va) {
    if (!have_state) throw std::logic_error("no state on the device");
    std::vector<double> a((size_t)n * R), b((size_t)n * R);
    d_vm.download(a.data(), a.size(), stream);
    d_va.download(b.data(), b.size(), stream);
    JGB_CUDA(cudaStreamSynchronize(stream));
    for (int i = 0; i < n; ++i) { vm[i] = a[(size_t)i * R]; va[i] = b[(size_t)i * R]; }
}

void FnrContext::sweep(bool q_only) {
    fnr_mismatch_kernel<<<dim3(ceil_div(n, 4), R / 32), dim3(32, 4), 0, stream>>>(view(), q_only ? 1 : 0);
    ++launches;
}

void FnrContext::step() {
    FnrDev d = view();
    const int blocks = (int)(((size_t)n * R + 255) / 256);
    active_lin.solve_block(d_mp.p, R);
    fnr_update_kernel<<<blocks, 256, 0, stream>>>(d, 0);
    sweep(true);
    reactive_lin.solve_block(d_mq.p, R);
    fnr_update_kernel<<<blocks, 256, 0, stream>>>(d, 1);
    launches += 2;
    JGB_CUDA(cudaGetLastError());
}

void FnrContext::mismatch(double* sp, double* sq) {
    if (!have_injection || !have_state) throw std::logic_error("set_injection / set_state must precede mismatch");
    JGB_CUDA(cudaMemsetAsync(d_active.p, 1, R, stream));
    sweep(false);
    fnr_check_kernel<<<ceil_div(R, 128), 128, 0, stream>>>(view(), 1, -1.0, 0);
    ++launches;
    d_stop.download(h_stop.p, 1, stream);
    JGB_CUDA(cudaMemcpyAsync(h_stop.p + 1, d_stop.p + R, sizeof(double), cudaMemcpyDeviceToHost, stream));
    JGB_CUDA(cudaStreamSynchronize(stream));
    if (sp) *sp = h_stop.p[0];
    if (sq) *sq = h_stop.p[1];
}

void FnrContext::solve() {
    if (!have_injection || !have_state) throw std::logic_error("set_injection / set_state must precede solve");
    JGB_CUDA(cudaMemsetAsync(d_active.p, 1, R, stream));
    step();
    JGB_CUDA(cudaStreamSynchronize(stream));
    iteration += 1;
}

int FnrContext::run(int64_t max_iter, double tol, int64_t* iters, double* sp, double* sq) {
    if (!have_injection || !have_state) throw std::logic_error("set_injection / set_state must precede run");
    FnrDev d = view();
    JGB_CUDA(cudaMemsetAsync(d_active.p, 1, R, stream));
    JGB_CUDA(cudaMemsetAsync(d_iters.p, 0, R * sizeof(int), stream));
    JGB_CUDA(cudaMemsetAsync(d_status.p, 0, R * sizeof(int), stream));
    iteration = 0;
    int rc = 1;
    for (int64_t it = 0; it <= max_iter; ++it) {
        JGB_CUDA(cudaMemsetAsync(d_remaining.p, 0, sizeof(int), stream));
        sweep(false);
        fnr_check_kernel<<<ceil_div(R, 128), 128, 0, stream>>>(d, 1, tol, (int)max_iter);
        ++launches;
        JGB_CUDA(cudaMemcpyAsync(h_int.p, d_remaining.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
        JGB_CUDA(cudaMemcpyAsync(h_int.p + 1, d_status.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
        d_stop.download(h_stop.p, 1, stream);
        JGB_CUDA(cudaMemcpyAsync(h_stop.p + 1, d_stop.p + R, sizeof(double), cudaMemcpyDeviceToHost, stream));
        JGB_CUDA(cudaStreamSynchronize(stream));
        if (h_int.p[0] == 0) { rc = h_int.p[1]; break; }
        step();
        iteration += 1;
    }
    if (iters) *iters = iteration;
    if (sp) *sp = h_stop.p[0];
    if (sq) *sq = h_stop.p[1];
    if (rc == -3) throw std::domain_error("fast Newton-Raphson diverged to non-finite values");
    return rc;
}

int FnrContext::batch(int64_t Rreal64, const double* pinj, const double* qinj, int64_t max_iter, double tol,
                      double* vm_out, double* va_out, int32_t* iters_out, int8_t* status_out, int64_t* total) {
    if (!have_state) throw std::logic_error("set_state must precede batch (start point of every scenario)");
    if (Rreal64 <= 0 || !pinj || !qinj || !vm_out || !va_out) throw std::invalid_argument("fnr_batch: null or empty input");
    if (Rreal64 > 65535LL * 32) throw std::invalid_argument("fnr_batch: too many scenarios");
    const int Rreal = (int)Rreal64, Rp = ceil_div(Rreal, 32) * 32;
    alloc(Rp);
    FnrDev d = view();
    const size_t cnt = (size_t)Rreal * n;
    const dim3 tb(32, 8), tg(ceil_div(n, 32), Rp / 32);
    d_io.alloc(cnt);
    JGB_CUDA(cudaMemcpyAsync(d_io.p, pinj, cnt * sizeof(double), cudaMemcpyHostToDevice, stream));
    fnr_transpose_in_kernel<<<tg, tb, 0, stream>>>(d_io.p, d_pinj.p, n, Rp, Rreal);
    d_io2.alloc(cnt);
    JGB_CUDA(cudaMemcpyAsync(d_io2.p, qinj, cnt * sizeof(double), cudaMemcpyHostToDevice, stream));
    fnr_transpose_in_kernel<<<tg, tb, 0, stream>>>(d_io2.p, d_qinj.p, n, Rp, Rreal);
    JGB_CUDA(cudaStreamSynchronize(stream));
    have_injection = true;
    broadcast_state();          // every scenario starts from the state of set_state
    launches += 2;
    JGB_CUDA(cudaMemsetAsync(d_active.p, 1, Rp, stream));
    JGB_CUDA(cudaMemsetAsync(d_iters.p, 0, Rp * sizeof(int), stream));
    JGB_CUDA(cudaMemsetAsync(d_status.p, 0, Rp * sizeof(int), stream));
    for (int64_t it = 0; it <= max_iter; ++it) {
        JGB_CUDA(cudaMemsetAsync(d_remaining.p, 0, sizeof(int), stream));
        sweep(false);
        fnr_check_kernel<<<ceil_div(Rp, 128), 128, 0, stream>>>(d, Rreal, tol, (int)max_iter);
        ++launches;
        JGB_CUDA(cudaMemcpyAsync(h_int.p, d_remaining.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
        JGB_CUDA(cudaStreamSynchronize(stream));
        if (h_int.p[0] == 0) break;
        step();
    }
    fnr_transpose_out_kernel<<<tg, tb, 0, stream>>>(d_vm.p, d_io.p, n, Rp, Rreal);
    fnr_transpose_out_kernel<<<tg, tb, 0, stream>>>(d_va.p, d_io2.p, n, Rp, Rreal);
    launches += 2;
    JGB_CUDA(cudaMemcpyAsync(vm_out, d_io.p, cnt * sizeof(double), cudaMemcpyDeviceToHost, stream));
    JGB_CUDA(cudaMemcpyAsync(va_out, d_io2.p, cnt * sizeof(double), cudaMemcpyDeviceToHost, stream));
    std::vector<int> it_h(Rp), st_h(Rp);
    d_iters.download(it_h.data(), Rp, stream);
    d_status.download(st_h.data(), Rp, stream);
    JGB_CUDA(cudaStreamSynchronize(stream));
    int rc = 0;
    int64_t sum = 0;
    for (int r = 0; r < Rreal; ++r) {
        if (iters_out) iters_out[r] = it_h[r];
        if (status_out) status_out[r] = (int8_t)st_h[r];
        sum += it_h[r];
        if (st_h[r] != 0) rc = 1;
    }
    if (total) *total = sum;
    have_injection = false;     // the injection block belongs to the batch
    return rc;
}

}  // namespace jgb
