// Newton-Raphson kernels and host driver. See nr.cuh.
#include "nr.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace jgb {

namespace {

constexpr int kWriteF = 1, kWriteJ = 2;
constexpr int kBatchBusBlock = 32;     // buses per CTA of the batch assembly kernel

// K1: fused mismatch + Jacobian values + max-abs reduction (mismatch! acPowerFlow.jl:645-685 and the Jacobian fill
// of solve! :813-888, formulas backend/equations.jl:63-143).  blockDim = (TS scenario lanes, TB bus lanes); one
// thread walks the Ybus column strip of its bus for its scenario, writing the two Jacobian columns of that bus
// through the same four running cursors the reference uses, so the CSC value order is identical.
//
// STAGED (batch launches): the Ybus strip of the CTA's bus block — row indices and both value arrays, one contiguous
// run each in the CSC arrays — is fetched once with cp.async.bulk (TMA) into shared memory and signalled through an
// mbarrier; the 32 scenario lanes of a warp then read every entry as a shared-memory broadcast, and the only global
// gathers left in the loop are the neighbour voltages.
template <bool STAGED>
__global__ void __launch_bounds__(128)
nr_assemble_kernel(NrDev d, int S, int bpb, int flags) {
    __shared__ double red[2][128];
    extern __shared__ __align__(16) unsigned char strip_raw[];
    __shared__ __align__(8) unsigned long long strip_bar;
    const int sl = threadIdx.x;
    const int s = blockIdx.y * blockDim.x + sl;
    const bool act = d.active ? (d.active[s] != 0) : true;
    int e_lo = 0, a_lo = 0;
    const double2* s_y = nullptr;
    const double2* s_yt = nullptr;
    const int* s_row = nullptr;
    if constexpr (STAGED) {
        if (!__syncthreads_or(act)) return;
        const int b0 = blockIdx.x * bpb, b1 = min(d.n, b0 + bpb);
        e_lo = d.ycolptr[b0];
        const int e_hi = d.ycolptr[b1];
        a_lo = e_lo & ~3;                                   // 16-byte aligned start of the int32 run
        const int nent = e_hi - e_lo, nrow = ((e_hi - a_lo) + 3) & ~3;
        double2* sy = reinterpret_cast<double2*>(strip_raw);
        double2* syt = sy + d.strip_cap;
        int* srow = reinterpret_cast<int*>(syt + d.strip_cap);
        const int tid = threadIdx.y * blockDim.x + threadIdx.x;
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&strip_bar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) {
            const unsigned bar = (unsigned)__cvta_generic_to_shared(&strip_bar);
            const unsigned bytes = (unsigned)(nent * 32 + nrow * 4);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             (unsigned)__cvta_generic_to_shared(sy)), "l"(d.y + e_lo), "r"((unsigned)(nent * 16)), "r"(bar) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             (unsigned)__cvta_generic_to_shared(syt)), "l"(d.yt + e_lo), "r"((unsigned)(nent * 16)), "r"(bar) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             (unsigned)__cvta_generic_to_shared(srow)), "l"(d.yrow + a_lo), "r"((unsigned)(nrow * 4)), "r"(bar) : "memory");
        }
        s_y = sy; s_yt = syt; s_row = srow;
    }
    double maxp = 0.0, maxq = 0.0;
    int of = -1, ot = -1;
    double2 dff = {0, 0}, dft = {0, 0}, dtf = {0, 0}, dtt = {0, 0};
    if (d.out_from && act) {
        of = d.out_from[s];
        ot = d.out_to[s];
        if (of >= 0) {
            dff = d.dy[0 * S + s]; dft = d.dy[1 * S + s]; dtf = d.dy[2 * S + s]; dtt = d.dy[3 * S + s];
        }
    }
    const bool wf = flags & kWriteF, wj = flags & kWriteJ;
    const int iend = min(d.n, (int)(blockIdx.x + 1) * bpb);
    if constexpr (STAGED) {
        unsigned ok;
        const unsigned bar = (unsigned)__cvta_generic_to_shared(&strip_bar);
        do {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ok) : "r"(bar), "r"(0u) : "memory");
        } while (!ok);
    }
    if (act) {
        for (int i = blockIdx.x * bpb + threadIdx.y; i < iend; i += blockDim.y) {
            if (i == d.slack) continue;
            const bool is_pq = d.type[i] == 1;
            const int k = d.pvpq[i], q = d.pq[i];
            const double Vi = d.vm[(long long)i * S + s], Ti = d.va[(long long)i * S + s];
            long long pa = d.jcolptr[k], qa = pa + d.pcount[i];
            long long pm = is_pq ? d.jcolptr[q] : 0, qm = pm + d.pcount[i];
            long long dpa = 0, dqa = 0, dpm = 0, dqm = 0;
            double Gii = 0.0, Bii = 0.0, sum_plus = 0.0, sum_minus = 0.0;
            const bool touched = (i == of) || (i == ot);
            for (int ptr = d.ycolptr[i]; ptr < d.ycolptr[i + 1]; ++ptr) {
                int r;
                double2 yt, yn;             // Y[i, r], Y[r, i]
                if constexpr (STAGED) { r = s_row[ptr - a_lo]; yt = s_yt[ptr - e_lo]; yn = s_y[ptr - e_lo]; }
                else { r = d.yrow[ptr]; yt = d.yt[ptr]; yn = d.y[ptr]; }
                if (touched) {               // branch (of -> ot) taken out: subtract its Y-parameters in place
                    if (i == of) {
                        if (r == of) { yn.x -= dff.x; yn.y -= dff.y; yt.x -= dff.x; yt.y -= dff.y; }
                        if (r == ot) { yn.x -= dtf.x; yn.y -= dtf.y; yt.x -= dft.x; yt.y -= dft.y; }
                    }
                    if (i == ot) {
                        if (r == ot) { yn.x -= dtt.x; yn.y -= dtt.y; yt.x -= dtt.x; yt.y -= dtt.y; }
                        if (r == of) { yn.x -= dft.x; yn.y -= dft.y; yt.x -= dtf.x; yt.y -= dtf.y; }
                    }
                }
                const double Vr = d.vm[(long long)r * S + s], Tr = d.va[(long long)r * S + s];
                double sn, cs;
                sincos(Ti - Tr, &sn, &cs);
                sum_plus += Vr * (yt.x * cs + yt.y * sn);     // PiQiSumPlus  (equations.jl:78-87)
                sum_minus += Vr * (yt.x * sn - yt.y * cs);    // PiQiSumMinus (equations.jl:89-98)
                const int tr = d.type[r];
                if (tr == 3 || !wj) continue;
                if (r != i) {
                    const double G = yn.x, B = yn.y, sj = -sn;   // sincos(theta_r - theta_i)
                    d.jval[pa * S + s] = Vr * Vi * (G * sj - B * cs);               // Pi_theta_j (:109)
                    ++pa;
                    if (tr == 1) { d.jval[qa * S + s] = -Vr * Vi * (G * cs + B * sj); ++qa; }   // Qi_theta_j (:134)
                    if (is_pq) { d.jval[pm * S + s] = Vr * (G * cs + B * sj); ++pm; }           // Pi_V_j (:117)
                    if (is_pq && tr == 1) { d.jval[qm * S + s] = Vr * (G * sj - B * cs); ++qm; }  // Qi_V_j (:142)
                } else {
                    Gii = yn.x; Bii = yn.y;
                    dpa = pa++;
                    if (is_pq) { dqa = qa++; dpm = pm++; dqm = qm++; }
                }
            }
            if (wf) {
                const double fp = Vi * sum_plus - d.sup_p[i] + d.dem_p[i];
                d.f[(long long)k * S + s] = fp;
                maxp = fmax(maxp, fabs(fp));
                if (is_pq) {
                    const double fq = Vi * sum_minus - d.sup_q[i] + d.dem_q[i];
                    d.f[(long long)q * S + s] = fq;
                    maxq = fmax(maxq, fabs(fq));
                }
                if (fp != fp) maxp = INFINITY;
            }
            if (wj) {
                d.jval[dpa * S + s] = Vi * (-sum_minus) - Bii * (Vi * Vi);        // Pi_theta_i (:105)
                if (is_pq) {
                    d.jval[dqa * S + s] = Vi * sum_plus - Gii * (Vi * Vi);        // Qi_theta_i (:130)
                    d.jval[dpm * S + s] = sum_plus + Gii * Vi;                    // Pi_V_i (:113)
                    d.jval[dqm * S + s] = sum_minus - Bii * Vi;                   // Qi_V_i (:138)
                }
            }
        }
    }
    if (!wf) return;
    const int t = threadIdx.y * blockDim.x + threadIdx.x;
    red[0][t] = maxp;
    red[1][t] = maxq;
    __syncthreads();
    for (int h = blockDim.y / 2; h >= 1; h >>= 1) {
        if ((int)threadIdx.y < h) {
            const int o = (threadIdx.y + h) * blockDim.x + threadIdx.x;
            red[0][t] = fmax(red[0][t], red[0][o]);
            red[1][t] = fmax(red[1][t], red[1][o]);
        }
        __syncthreads();
    }
    if (threadIdx.y == 0 && act) {
        atomicMax(&d.stopbits[s], (unsigned long long)__double_as_longlong(red[0][t]));
        atomicMax(&d.stopbits[S + s], (unsigned long long)__double_as_longlong(red[1][t]));
    }
}

// Convergence bookkeeping of powerFlow! (acPowerFlow.jl:1406-1418), one thread per scenario.
// tol < 0: only publish the stop values (mismatch! operator).
__global__ void nr_check_kernel(NrDev d, int S, int Sreal, double tol, int max_iter) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    if (d.active && !d.active[s]) return;
    const double sp = __longlong_as_double((long long)d.stopbits[s]);
    const double sq = __longlong_as_double((long long)d.stopbits[S + s]);
    d.stop[s] = sp;
    d.stop[S + s] = sq;
    d.stopbits[s] = 0ull;
    d.stopbits[S + s] = 0ull;
    if (tol < 0.0 || s >= Sreal) return;
    if (d.status[s] < 0) { d.active[s] = 0; return; }        // singular pivot met in the previous solve
    if (sp < tol && sq < tol) { d.active[s] = 0; d.status[s] = 0; return; }
    if (d.iters[s] == max_iter) { d.active[s] = 0; d.status[s] = 1; return; }
    atomicAdd(d.remaining, 1);
}

// State update of solve! (acPowerFlow.jl:899-908).
__global__ void nr_update_kernel(NrDev d, int S) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int i = (int)(gid / S), s = (int)(gid % S);
    if (i >= d.n) return;
    if (d.active && !d.active[s]) return;
    if (d.type[i] == 1) d.vm[gid] -= d.inc[(long long)d.pq[i] * S + s];
    if (i != d.slack) d.va[gid] -= d.inc[(long long)d.pvpq[i] * S + s];
    if (i == 0) d.iters[s] += 1;
}

__global__ void broadcast_state_kernel(const double* __restrict__ src, double* __restrict__ dst, int n, int S) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)n * S) return;
    dst[gid] = src[gid / S];
}

// [n][S] (scenario minor) -> [Sreal][n] (one row per scenario), 32x32 tiles through shared memory
__global__ void transpose_out_kernel(const double* __restrict__ src, double* __restrict__ dst, int n, int S,
                                     int Sreal) {
    __shared__ double tile[32][33];
    const int i0 = blockIdx.x * 32, s0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int i = i0 + r, s = s0 + threadIdx.x;
        if (i < n && s < S) tile[r][threadIdx.x] = src[(long long)i * S + s];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int s = s0 + r, i = i0 + threadIdx.x;
        if (s < Sreal && i < n) dst[(long long)s * n + i] = tile[threadIdx.x][r];
    }
}

__global__ void convert_outage_kernel(const int64_t* __restrict__ of, const int64_t* __restrict__ ot,
                                      const double* __restrict__ dy, int* __restrict__ of32, int* __restrict__ ot32,
                                      double2* __restrict__ dy2, unsigned char* __restrict__ active,
                                      int* __restrict__ status, int* __restrict__ iters, int S, int Sreal) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    if (s < Sreal) {
        of32[s] = (int)of[s] - 1;
        ot32[s] = (int)ot[s] - 1;
        for (int c = 0; c < 4; ++c) dy2[c * S + s] = make_double2(dy[s * 8 + 2 * c], dy[s * 8 + 2 * c + 1]);
        active[s] = 1;
        status[s] = 1;
    } else {
        of32[s] = -1;
        ot32[s] = -1;
        for (int c = 0; c < 4; ++c) dy2[c * S + s] = make_double2(0, 0);
        active[s] = 0;
        status[s] = 0;
    }
    iters[s] = 0;
}

__global__ void copy_results_kernel(const int* __restrict__ iters, const int* __restrict__ status,
                                    int32_t* __restrict__ it_out, int8_t* __restrict__ st_out, int Sreal) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= Sreal) return;
    it_out[s] = iters[s];
    st_out[s] = (int8_t)status[s];
}

// power! / current! (postprocessing/acAnalysis.jl:30-79, 672-700): thread t < n -> bus injection
// S_i = V_i conj(sum_j Y_ij V_j) over the Ybus strip; thread n + k -> branch k: I_from = Yff V_i + Yft V_j,
// I_to = Ytf V_i + Ytt V_j, S = V conj(I). Output layout: [inj_p | inj_q] (n each), then 8 branch vectors (nbr each).
__global__ void nr_power_kernel(NrDev d, int nbr, const int* __restrict__ bf, const int* __restrict__ bt,
                                const double2* __restrict__ yff, const double2* __restrict__ yft,
                                const double2* __restrict__ ytf, const double2* __restrict__ ytt,
                                const signed char* __restrict__ st, double* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = d.n;
    if (t < n) {
        double re = 0.0, im = 0.0;
        for (int p = d.ycolptr[t]; p < d.ycolptr[t + 1]; ++p) {
            const int j = d.yrow[p];
            const double2 y = d.yt[p];      // Y[t, j]
            double sn, cs;
            sincos(d.va[j], &sn, &cs);
            const double vr = d.vm[j] * cs, vi = d.vm[j] * sn;
            re += y.x * vr - y.y * vi;
            im += y.x * vi + y.y * vr;
        }
        double sn, cs;
        sincos(d.va[t], &sn, &cs);
        const double vr = d.vm[t] * cs, vi = d.vm[t] * sn;
        out[t] = vr * re + vi * im;          // Re(V conj(I))
        out[n + t] = vi * re - vr * im;      // Im(V conj(I))
    } else if (t < n + nbr) {
        const int k = t - n;
        double* o = out + 2 * n;
        if (st[k] != 1) {
            for (int q = 0; q < 8; ++q) o[(long long)q * nbr + k] = 0.0;
            return;
        }
        const int i = bf[k], j = bt[k];
        double si, ci, sj, cj;
        sincos(d.va[i], &si, &ci);
        sincos(d.va[j], &sj, &cj);
        const double vir = d.vm[i] * ci, vii = d.vm[i] * si, vjr = d.vm[j] * cj, vji = d.vm[j] * sj;
        const double2 a = yff[k], b = yft[k], c = ytf[k], e = ytt[k];
        const double ifr = a.x * vir - a.y * vii + b.x * vjr - b.y * vji;
        const double ifi = a.x * vii + a.y * vir + b.x * vji + b.y * vjr;
        const double itr = c.x * vir - c.y * vii + e.x * vjr - e.y * vji;
        const double iti = c.x * vii + c.y * vir + e.x * vji + e.y * vjr;
        o[0LL * nbr + k] = vir * ifr + vii * ifi;
        o[1LL * nbr + k] = vii * ifr - vir * ifi;
        o[2LL * nbr + k] = vjr * itr + vji * iti;
        o[3LL * nbr + k] = vji * itr - vjr * iti;
        o[4LL * nbr + k] = hypot(ifr, ifi);
        o[5LL * nbr + k] = atan2(ifi, ifr);
        o[6LL * nbr + k] = hypot(itr, iti);
        o[7LL * nbr + k] = atan2(iti, itr);
    }
}

// ---- pivot guard: residual r = f - J * inc of the flagged scenarios (rare path: atomics are fine) -------------------------
__global__ void nr_weak_count_kernel(const int* __restrict__ weak, const unsigned char* __restrict__ active, int S,
                                     unsigned char* __restrict__ mask, int* __restrict__ counter) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const bool w = weak[s] != 0 && (!active || active[s]);
    mask[s] = w ? 1 : 0;
    if (w) atomicAdd(counter, 1);
}

__global__ void nr_refine_copy_kernel(const double* __restrict__ f, double* __restrict__ r,
                                      const unsigned char* __restrict__ mask, int dim, int S) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)dim * S) return;
    if (mask[gid % S]) r[gid] = f[gid];
}

__global__ void nr_refine_residual_kernel(const int* __restrict__ jcolptr, const int* __restrict__ jrow,
                                          const double* __restrict__ jval, const double* __restrict__ inc,
                                          double* __restrict__ r, const unsigned char* __restrict__ mask, int dim, int S) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)dim * S) return;
    const int c = (int)(gid / S), s = (int)(gid % S);
    if (!mask[s]) return;
    const double x = inc[gid];
    for (int q = jcolptr[c]; q < jcolptr[c + 1]; ++q)
        atomicAdd(&r[(long long)jrow[q] * S + s], -jval[(long long)q * S + s] * x);
}

__global__ void nr_refine_add_kernel(double* __restrict__ inc, const double* __restrict__ inc2,
                                     const unsigned char* __restrict__ mask, int dim, int S) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)dim * S) return;
    if (mask[gid % S]) inc[gid] += inc2[gid];
}

}  // namespace

int NrContext::refine_weak(MfSolver& sol, int S, double* jval, double* f, double* inc, const unsigned char* active,
                           int* status, int* weak) {
    d_weakcnt.alloc(1);
    r_mask.alloc(S);
    JGB_CUDA(cudaMemsetAsync(d_weakcnt.p, 0, sizeof(int), stream));
    nr_weak_count_kernel<<<ceil_div(S, 128), 128, 0, stream>>>(weak, active, S, r_mask.p, d_weakcnt.p);
    ++launches;
    JGB_CUDA(cudaMemcpyAsync(h_int.p + 2, d_weakcnt.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
    JGB_CUDA(cudaStreamSynchronize(stream));
    const int cnt = h_int.p[2];
    if (cnt == 0) return 0;
    weak_events += cnt;
    ++refine_calls;
    const long long ds = (long long)dim * S;
    r_res.alloc((size_t)ds);
    r_inc.alloc((size_t)ds);
    const int blocks = (int)((ds + 255) / 256);
    nr_refine_copy_kernel<<<blocks, 256, 0, stream>>>(f, r_res.p, r_mask.p, dim, S);
    nr_refine_residual_kernel<<<blocks, 256, 0, stream>>>(d_jcolptr.p, d_jrow.p, jval, inc, r_res.p, r_mask.p, dim, S);
    // the same factorisation again, carrying the residual as right-hand side (L is never stored), flagged scenarios only
    sol.set_pivot_guard(nullptr, 1e300);
    sol.factor_solve(jval, r_res.p, r_inc.p, S, r_mask.p, status, stream);
    sol.set_pivot_guard(weak, pivot_growth);
    nr_refine_add_kernel<<<blocks, 256, 0, stream>>>(inc, r_inc.p, r_mask.p, dim, S);
    launches += 3 + sol.launches_per_solve(S);
    JGB_CUDA(cudaGetLastError());
    return cnt;
}

namespace {
}  // namespace

void NrContext::set_branches(int64_t nbr_, const int64_t* from, const int64_t* to, const double* yff, const double* yft,
                             const double* ytf, const double* ytt, const int8_t* status) {
    if (!n) throw std::logic_error("nr_setup has not been called");
    if (nbr_ <= 0 || !from || !to || !yff || !yft || !ytf || !ytt || !status)
        throw std::invalid_argument("nr_set_branches: null or empty input");
    nbr = (int)nbr_;
    std::vector<int> f(nbr), t(nbr);
    for (int k = 0; k < nbr; ++k) {
        if (from[k] < 1 || from[k] > n || to[k] < 1 || to[k] > n)
            throw std::invalid_argument("nr_set_branches: bus index out of range");
        f[k] = (int)from[k] - 1;
        t[k] = (int)to[k] - 1;
    }
    d_brfrom.upload(f, stream);
    d_brto.upload(t, stream);
    d_yff.upload(reinterpret_cast<const double2*>(yff), nbr, stream);
    d_yft.upload(reinterpret_cast<const double2*>(yft), nbr, stream);
    d_ytf.upload(reinterpret_cast<const double2*>(ytf), nbr, stream);
    d_ytt.upload(reinterpret_cast<const double2*>(ytt), nbr, stream);
    d_brstatus.upload(reinterpret_cast<const signed char*>(status), nbr, stream);
    d_pw.alloc(2 * (size_t)n + 8 * (size_t)nbr);
    JGB_CUDA(cudaStreamSynchronize(stream));
}

void NrContext::power(double* out[10]) {
    if (!have_state) throw std::logic_error("no state on the device");
    if (!nbr) throw std::logic_error("nr_set_branches must precede nr_power");
    NrDev d = view(1, false);
    const int work = n + nbr;
    nr_power_kernel<<<ceil_div(work, 128), 128, 0, stream>>>(d, nbr, d_brfrom.p, d_brto.p, d_yff.p, d_yft.p, d_ytf.p,
                                                             d_ytt.p, d_brstatus.p, d_pw.p);
    ++launches;
    JGB_CUDA(cudaGetLastError());
    for (int q = 0; q < 10; ++q) {
        if (!out[q]) continue;
        const size_t off = q < 2 ? (size_t)q * n : 2 * (size_t)n + (size_t)(q - 2) * nbr;
        JGB_CUDA(cudaMemcpyAsync(out[q], d_pw.p + off, (q < 2 ? n : nbr) * sizeof(double), cudaMemcpyDeviceToHost, stream));
    }
    JGB_CUDA(cudaStreamSynchronize(stream));
}

void NrContext::setup(int64_t n_, const int64_t* ycp, const int64_t* yrv, const double* y, const double* yt,
                      const int8_t* type, int64_t slack_) {
    if (n_ <= 0 || !ycp || !yrv || !y || !yt || !type) throw std::invalid_argument("nr_setup: null or empty input");
    if (slack_ < 1 || slack_ > n_) throw std::invalid_argument("nr_setup: slack index out of range");
    if (ycp[0] != 1) throw std::invalid_argument("nr_setup: colptr must be 1-based");
    n = (int)n_;
    slack = (int)slack_ - 1;
    nnzy = (int)(ycp[n] - 1);
    if (type[slack] != 3) throw std::invalid_argument("nr_setup: bus.layout.type[slack] must be 3");
    std::vector<int> cp(n + 1), rv(nnzy);
    for (int i = 0; i <= n; ++i) cp[i] = (int)(ycp[i] - 1);
    for (int q = 0; q < nnzy; ++q) {
        if (yrv[q] < 1 || yrv[q] > n) throw std::invalid_argument("nr_setup: row index out of range");
        rv[q] = (int)(yrv[q] - 1);
    }
    // ---- index maps and Jacobian pattern (newtonJacobian, acPowerFlow.jl:89-175)
    std::vector<int> pq(n, -1), pvpq(n, -1), pcount(n, 0), qcount(n, 0);
    int npq = 0, npvpq = 0;
    for (int i = 0; i < n; ++i) {
        if (type[i] == 1) pq[i] = (npq++) + n - 1;
        if (type[i] != 3) pvpq[i] = npvpq++;
    }
    if (npvpq != n - 1) throw std::invalid_argument("nr_setup: exactly one slack bus (type 3) is required");
    dim = n + npq - 1;
    std::vector<int> jcp(dim + 1, 0);
    for (int i = 0; i < n; ++i) {
        if (i == slack) continue;
        for (int p = cp[i]; p < cp[i + 1]; ++p) {
            int t = type[rv[p]];
            pcount[i] += (t != 3);
            qcount[i] += (t == 1);
        }
        jcp[pvpq[i] + 1] = pcount[i] + qcount[i];
        if (type[i] == 1) jcp[pq[i] + 1] = pcount[i] + qcount[i];
    }
    for (int c = 0; c < dim; ++c) jcp[c + 1] += jcp[c];
    nnzj = jcp[dim];
    std::vector<int> jrv(nnzj);
    for (int i = 0; i < n; ++i) {
        if (i == slack) continue;
        const bool is_pq = type[i] == 1;
        int pa = jcp[pvpq[i]], qa = pa + pcount[i];
        int pm = is_pq ? jcp[pq[i]] : 0, qm = pm + pcount[i];
        for (int p = cp[i]; p < cp[i + 1]; ++p) {
            int r = rv[p], t = type[r];
            if (t != 3) {
                jrv[pa++] = pvpq[r];
                if (is_pq) jrv[pm++] = pvpq[r];
            }
            if (t == 1) {
                jrv[qa++] = pq[r];
                if (is_pq) jrv[qm++] = pq[r];
            }
        }
    }
    pq1.resize(n); pvpq1.resize(n); pcount1.resize(n);
    for (int i = 0; i < n; ++i) { pq1[i] = pq[i] + 1; pvpq1[i] = pvpq[i] + 1; pcount1[i] = pcount[i]; }
    jcolptr1.resize(dim + 1);
    jrowval1.resize(nnzj);
    for (int c = 0; c <= dim; ++c) jcolptr1[c] = jcp[c] + 1;
    for (int q = 0; q < nnzj; ++q) jrowval1[q] = jrv[q] + 1;

    // ---- symbolic factorisation: theta_i / V_i of one bus form a supervariable
    std::vector<int> group(dim);
    for (int i = 0; i < n; ++i) {
        if (pvpq[i] >= 0) group[pvpq[i]] = i;
        if (pq[i] >= 0) group[pq[i]] = i;
    }
    {
        Symbolic sym;
        analyse(dim, jcp.data(), jrv.data(), group.data(), nullptr, latency_options(), sym);
        solver.setup(sym, stream);
        Symbolic symb;
        analyse(dim, jcp.data(), jrv.data(), group.data(), nullptr, throughput_options(), symb);
        solver_batch.setup(symb, stream);
    }

    // ---- device upload
    d_ycolptr.upload(cp, stream);
    {
        std::vector<int> rvp(rv);
        rvp.resize(rv.size() + 8, 0);                     // the staged copy rounds the run up to 16 bytes
        d_yrow.upload(rvp, stream);
        strip_cap = 8;
        for (int b0 = 0; b0 < n; b0 += kBatchBusBlock) {
            const int b1 = std::min(n, b0 + kBatchBusBlock);
            strip_cap = std::max(strip_cap, ((cp[b1] - (cp[b0] & ~3)) + 7) & ~3);
        }
        JGB_CUDA(cudaStreamSynchronize(stream));
    }
    d_y.upload(reinterpret_cast<const double2*>(y), nnzy, stream);
    d_yt.upload(reinterpret_cast<const double2*>(yt), nnzy, stream);
    d_type.upload(reinterpret_cast<const signed char*>(type), n, stream);
    d_pq.upload(pq, stream);
    d_pvpq.upload(pvpq, stream);
    d_pcount.upload(pcount, stream);
    d_jcolptr.upload(jcp, stream);
    d_jrow.upload(jrv, stream);
    d_weak.alloc(1);
    d_weak.zero(stream);
    d_sup_p.alloc(n); d_sup_q.alloc(n); d_dem_p.alloc(n); d_dem_q.alloc(n);
    d_vm.alloc(n); d_va.alloc(n); d_f.alloc(dim); d_jval.alloc(nnzj); d_inc.alloc(dim);
    d_stop.alloc(2); d_stopbits.alloc(2); d_active.alloc(1); d_status.alloc(1); d_iters.alloc(1);
    d_remaining.alloc(1);
    d_f.zero(stream); d_jval.zero(stream); d_inc.zero(stream); d_stop.zero(stream); d_stopbits.zero(stream);
    d_status.zero(stream); d_iters.zero(stream); d_remaining.zero(stream);
    JGB_CUDA(cudaMemsetAsync(d_active.p, 1, 1, stream));
    h_stop.alloc(4);
    h_int.alloc(4);
    JGB_CUDA(cudaStreamSynchronize(stream));
    iteration = 0;
    jac_valid = false;
    have_injection = have_state = false;
    batch_S = 0;
}

void NrContext::set_injection(const double* ps, const double* qs, const double* pd, const double* qd) {
    if (!n) throw std::logic_error("nr_setup has not been called");
    if (!ps || !qs || !pd || !qd) throw std::invalid_argument("nr_set_injection: null input");
    d_sup_p.upload(ps, n, stream); d_sup_q.upload(qs, n, stream);
    d_dem_p.upload(pd, n, stream); d_dem_q.upload(qd, n, stream);
    JGB_CUDA(cudaStreamSynchronize(stream));
    have_injection = true;
}

void NrContext::set_state(const double* vm, const double* va) {
    if (!n) throw std::logic_error("nr_setup has not been called");
    if (!vm || !va) throw std::invalid_argument("nr_set_state: null input");
    d_vm.upload(vm, n, stream);
    d_va.upload(va, n, stream);
    JGB_CUDA(cudaStreamSynchronize(stream));
    jac_valid = false;
    have_state = true;
}

void NrContext::get_state(double* vm, double* va) {
    if (!have_state) throw std::logic_error("no state on the device");
    d_vm.download(vm, n, stream);
    d_va.download(va, n, stream);
    JGB_CUDA(cudaStreamSynchronize(stream));
}

void NrContext::update_y(int64_t k, const int64_t* pos, const double* y, const double* yt) {
    if (!n) throw std::logic_error("nr_setup has not been called");
    for (int64_t e = 0; e < k; ++e) {
        if (pos[e] < 1 || pos[e] > nnzy) throw std::invalid_argument("nr_update_y: position out of range");
        JGB_CUDA(cudaMemcpyAsync(d_y.p + (pos[e] - 1), y + 2 * e, sizeof(double2), cudaMemcpyHostToDevice, stream));
        JGB_CUDA(cudaMemcpyAsync(d_yt.p + (pos[e] - 1), yt + 2 * e, sizeof(double2), cudaMemcpyHostToDevice, stream));
    }
    JGB_CUDA(cudaStreamSynchronize(stream));
    jac_valid = false;
}

NrDev NrContext::view(int S, bool batch) {
    NrDev d{};
    d.n = n; d.slack = slack; d.dim = dim; d.nnzj = nnzj;
    d.ycolptr = d_ycolptr.p; d.yrow = d_yrow.p; d.y = d_y.p; d.yt = d_yt.p; d.type = d_type.p;
    d.pq = d_pq.p; d.pvpq = d_pvpq.p; d.pcount = d_pcount.p; d.jcolptr = d_jcolptr.p;
    d.sup_p = d_sup_p.p; d.sup_q = d_sup_q.p; d.dem_p = d_dem_p.p; d.dem_q = d_dem_q.p;
    d.remaining = d_remaining.p;
    d.strip_cap = strip_cap;
    if (!batch) {
        d.vm = d_vm.p; d.va = d_va.p; d.f = d_f.p; d.jval = d_jval.p; d.inc = d_inc.p;
        d.stopbits = d_stopbits.p; d.stop = d_stop.p; d.active = d_active.p; d.status = d_status.p;
        d.iters = d_iters.p;
    } else {
        d.vm = b_vm.p; d.va = b_va.p; d.f = b_f.p; d.jval = b_jval.p; d.inc = b_inc.p;
        d.stopbits = b_stopbits.p; d.stop = b_stop.p; d.active = b_active.p; d.status = b_status.p;
        d.iters = b_iters.p;
        d.out_from = b_of.p; d.out_to = b_ot.p; d.dy = b_dy.p;
    }
    (void)S;
    return d;
}

void NrContext::launch_assemble(int S, bool batch) {
    NrDev d = view(S, batch);
    if (S == 1) {
        dim3 block(1, 128);
        dim3 grid(ceil_div(n, 128), 1);
        nr_assemble_kernel<false><<<grid, block, 0, stream>>>(d, 1, 128, kWriteF | kWriteJ);
    } else {
        dim3 block(32, 4);
        dim3 grid(ceil_div(n, kBatchBusBlock), S / 32);
        // The TMA-staged variant is kept selectable (JGB_STAGED_ASSEMBLY=1) but is not the default: measured 0.96 ms
        // vs 0.87 ms per 2048 scenarios on the 10k grid — the strip reads are already L1-resident warp broadcasts and
        // the kernel is bound by the Jacobian write stream, so the extra barrier only adds latency.
        if (staged_assembly) {
            const size_t smem = (size_t)strip_cap * 36;
            nr_assemble_kernel<true><<<grid, block, smem, stream>>>(d, S, kBatchBusBlock, kWriteF | kWriteJ);
        } else {
            nr_assemble_kernel<false><<<grid, block, 0, stream>>>(d, S, kBatchBusBlock, kWriteF | kWriteJ);
        }
    }
    ++launches;
    JGB_CUDA(cudaGetLastError());
}

void NrContext::mismatch(double* sp, double* sq) {
    if (!have_injection || !have_state) throw std::logic_error("set_injection / set_state must precede mismatch");
    NrDev d = view(1, false);
    JGB_CUDA(cudaMemsetAsync(d_active.p, 1, 1, stream));
    launch_assemble(1, false);
    nr_check_kernel<<<1, 32, 0, stream>>>(d, 1, 1, -1.0, 0);
    ++launches;
    d_stop.download(h_stop.p, 2, stream);
    JGB_CUDA(cudaStreamSynchronize(stream));
    if (sp) *sp = h_stop.p[0];
    if (sq) *sq = h_stop.p[1];
    jac_valid = true;
}

void NrContext::solve() {
    if (!have_injection || !have_state) throw std::logic_error("set_injection / set_state must precede solve");
    NrDev d = view(1, false);
    JGB_CUDA(cudaMemsetAsync(d_active.p, 1, 1, stream));
    if (!jac_valid) {
        dim3 block(1, 128);
        nr_assemble_kernel<false><<<dim3(ceil_div(n, 128), 1), block, 0, stream>>>(d, 1, 128, kWriteJ);
        ++launches;
    }
    JGB_CUDA(cudaMemsetAsync(d_status.p, 0, sizeof(int), stream));
    JGB_CUDA(cudaMemsetAsync(d_weak.p, 0, sizeof(int), stream));
    solver.set_pivot_guard(d_weak.p, pivot_growth);
    solver.factor_solve(d_jval.p, d_f.p, d_inc.p, 1, nullptr, d_status.p, stream);
    launches += solver.launches_per_solve(1);
    refine_weak(solver, 1, d_jval.p, d_f.p, d_inc.p, nullptr, d_status.p, d_weak.p);
    nr_update_kernel<<<ceil_div(n, 128), 128, 0, stream>>>(d, 1);
    ++launches;
    JGB_CUDA(cudaMemcpyAsync(h_int.p, d_status.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
    JGB_CUDA(cudaStreamSynchronize(stream));
    jac_valid = false;
    iteration += 1;
    if (h_int.p[0] == -3) throw std::domain_error("singular Jacobian: zero or non-finite pivot");
}

void NrContext::get_vectors(double* f, double* inc, double* jv, int64_t* it) {
    if (!n) throw std::logic_error("nr_setup has not been called");
    if (f) d_f.download(f, dim, stream);
    if (inc) d_inc.download(inc, dim, stream);
    if (jv) d_jval.download(jv, nnzj, stream);
    JGB_CUDA(cudaStreamSynchronize(stream));
    if (it) *it = iteration;
}

int NrContext::run(int64_t max_iter, double tol, int64_t* iters, double* sp, double* sq) {
    if (!have_injection || !have_state) throw std::logic_error("set_injection / set_state must precede run");
    NrDev d = view(1, false);
    JGB_CUDA(cudaMemsetAsync(d_active.p, 1, 1, stream));
    JGB_CUDA(cudaMemsetAsync(d_iters.p, 0, sizeof(int), stream));
    JGB_CUDA(cudaMemsetAsync(d_stopbits.p, 0, 2 * sizeof(unsigned long long), stream));
    int one = 1;
    JGB_CUDA(cudaMemcpyAsync(d_status.p, &one, sizeof(int), cudaMemcpyHostToDevice, stream));
    iteration = 0;
    int rc = 1;
    // the graph loop only records weak pivots (nr.weak_pivot_scenarios); the refinement step runs in solve() and in batches
    JGB_CUDA(cudaMemsetAsync(d_weak.p, 0, sizeof(int), stream));
    solver.set_pivot_guard(d_weak.p, pivot_growth);
    const int per_solve = solver.launches_per_solve(1);      // also plans / allocates before any capture
    // mismatch! + convergence bookkeeping + the 16-byte read-back: the head of every loop trip
    auto enqueue_head = [&] {
        JGB_CUDA(cudaMemsetAsync(d_remaining.p, 0, sizeof(int), stream));
        dim3 block(1, 128);
        nr_assemble_kernel<false><<<dim3(ceil_div(n, 128), 1), block, 0, stream>>>(d, 1, 128, kWriteF | kWriteJ);
        nr_check_kernel<<<1, 32, 0, stream>>>(d, 1, 1, tol, (int)max_iter);
        JGB_CUDA(cudaMemcpyAsync(h_int.p, d_remaining.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
        JGB_CUDA(cudaMemcpyAsync(h_int.p + 1, d_status.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
        JGB_CUDA(cudaMemcpyAsync(h_stop.p, d_stop.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, stream));
    };
    // solve! (refactor + solve + update) followed by the next head
    auto enqueue_iter = [&] {
        solver.factor_solve(d_jval.p, d_f.p, d_inc.p, 1, d_active.p, d_status.p, stream);
        nr_update_kernel<<<ceil_div(n, 128), 128, 0, stream>>>(d, 1);
        enqueue_head();
    };
    const bool use_graph = !timer.enabled && !graphs_disabled;
    if (use_graph) {
        if (!graph_head.valid(tol, max_iter)) graph_head.capture(stream, tol, max_iter, enqueue_head);
        if (!graph_iter.valid(tol, max_iter)) graph_iter.capture(stream, tol, max_iter, enqueue_iter);
    }
    if (use_graph) graph_head.launch(stream); else enqueue_head();
    launches += 2;
    for (int64_t it = 0; it <= max_iter; ++it) {
        JGB_CUDA(cudaStreamSynchronize(stream));
        timer.resolve();
        if (h_int.p[0] == 0) { rc = h_int.p[1]; break; }
        if (use_graph) {
            graph_iter.launch(stream);
        } else {
            timer.mark(stream);
            const size_t f0 = timer.last();
            cudaEvent_t mid = timer.reserve();
            const size_t f1 = timer.last();
            solver.factor_solve(d_jval.p, d_f.p, d_inc.p, 1, d_active.p, d_status.p, stream, mid);
            timer.mark(stream);
            const size_t f2 = timer.last();
            timer.span(kPhFactor, f0, f1);
            timer.span(kPhBacksolve, f1, f2);
            nr_update_kernel<<<ceil_div(n, 128), 128, 0, stream>>>(d, 1);
            enqueue_head();
        }
        launches += per_solve + 3;
        iteration += 1;
    }
    JGB_CUDA(cudaGetLastError());
    JGB_CUDA(cudaMemcpyAsync(h_int.p + 2, d_weak.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
    JGB_CUDA(cudaStreamSynchronize(stream));
    if (h_int.p[2]) ++weak_events;
    jac_valid = true;
    if (iters) *iters = iteration;
    if (sp) *sp = h_stop.p[0];
    if (sq) *sq = h_stop.p[1];
    if (rc == -3) throw std::domain_error("singular Jacobian: zero or non-finite pivot");
    return rc;
}

void NrContext::alloc_state(int S) {
    if (S <= batch_S) return;
    b_vm.alloc((size_t)n * S); b_va.alloc((size_t)n * S); b_f.alloc((size_t)dim * S);
    b_jval.alloc((size_t)nnzj * S); b_inc.alloc((size_t)dim * S); b_stop.alloc(2 * (size_t)S);
    b_stopbits.alloc(2 * (size_t)S); b_active.alloc(S); b_status.alloc(S); b_iters.alloc(S);
    b_of.alloc(S); b_ot.alloc(S); b_dy.alloc(4 * (size_t)S); b_weak.alloc(S);
    batch_S = S;
}

int NrContext::batch(int64_t Sreal64, const int64_t* of, const int64_t* ot, const double* dy, bool dev_in,
                     int64_t max_iter, double tol, double* vm_out, double* va_out, int32_t* iters_out,
                     int8_t* status_out, bool dev_out, int64_t* total) {
    if (!have_injection || !have_state) throw std::logic_error("set_injection / set_state must precede batch");
    if (Sreal64 <= 0 || !of || !ot || !dy) throw std::invalid_argument("nr_batch: null or empty input");
    const int Sreal = (int)Sreal64;
    const int S = ceil_div(Sreal, 32) * 32;
    if (S != batch_S) { batch_S = 0; }
    alloc_state(S);
    const int64_t *dof = of, *dot = ot;
    const double* ddy = dy;
    if (!dev_in) {
        b_of64.upload(of, Sreal, stream);
        b_ot64.upload(ot, Sreal, stream);
        b_dyraw.upload(dy, (size_t)Sreal * 8, stream);
        dof = b_of64.p; dot = b_ot64.p; ddy = b_dyraw.p;
    }
    convert_outage_kernel<<<ceil_div(S, 128), 128, 0, stream>>>(dof, dot, ddy, b_of.p, b_ot.p, b_dy.p, b_active.p,
                                                                b_status.p, b_iters.p, S, Sreal);
    const long long ns = (long long)n * S;
    broadcast_state_kernel<<<(int)((ns + 255) / 256), 256, 0, stream>>>(d_vm.p, b_vm.p, n, S);
    broadcast_state_kernel<<<(int)((ns + 255) / 256), 256, 0, stream>>>(d_va.p, b_va.p, n, S);
    JGB_CUDA(cudaMemsetAsync(b_stopbits.p, 0, 2 * (size_t)S * sizeof(unsigned long long), stream));
    launches += 3;
    NrDev d = view(S, true);
    for (int64_t it = 0; it <= max_iter; ++it) {
        JGB_CUDA(cudaMemsetAsync(d_remaining.p, 0, sizeof(int), stream));
        timer.mark(stream);
        size_t e0 = timer.last();
        launch_assemble(S, true);
        timer.mark(stream);
        timer.span(kPhAssemble, e0, timer.last());
        nr_check_kernel<<<ceil_div(S, 128), 128, 0, stream>>>(d, S, Sreal, tol, (int)max_iter);
        ++launches;
        JGB_CUDA(cudaMemcpyAsync(h_int.p, d_remaining.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
        JGB_CUDA(cudaStreamSynchronize(stream));
        timer.resolve();
        if (h_int.p[0] == 0) break;
        timer.mark(stream);
        size_t f0 = timer.last();
        cudaEvent_t mid = timer.reserve();
        size_t f1 = timer.last();
        JGB_CUDA(cudaMemsetAsync(b_weak.p, 0, (size_t)S * sizeof(int), stream));
        solver_batch.set_pivot_guard(b_weak.p, pivot_growth);
        solver_batch.factor_solve(b_jval.p, b_f.p, b_inc.p, S, b_active.p, b_status.p, stream, mid);
        launches += solver_batch.launches_per_solve(S);
        timer.mark(stream);
        size_t f2 = timer.last();
        refine_weak(solver_batch, S, b_jval.p, b_f.p, b_inc.p, b_active.p, b_status.p, b_weak.p);
        timer.span(kPhFactor, f0, f1);
        timer.span(kPhBacksolve, f1, f2);
        nr_update_kernel<<<(int)((ns + 127) / 128), 128, 0, stream>>>(d, S);
        ++launches;
    }
    // results: [n][S] -> [Sreal][n]
    double* dvm = vm_out;
    double* dva = va_out;
    if (!dev_out) {
        b_out.alloc(2 * (size_t)Sreal * n);
        dvm = b_out.p;
        dva = b_out.p + (size_t)Sreal * n;
    }
    dim3 tb(32, 8), tg(ceil_div(n, 32), S / 32);
    transpose_out_kernel<<<tg, tb, 0, stream>>>(b_vm.p, dvm, n, S, Sreal);
    transpose_out_kernel<<<tg, tb, 0, stream>>>(b_va.p, dva, n, S, Sreal);
    launches += 2;
    std::vector<int> hit(Sreal), hst(Sreal);
    if (dev_out) {
        copy_results_kernel<<<ceil_div(Sreal, 128), 128, 0, stream>>>(b_iters.p, b_status.p, iters_out, status_out,
                                                                      Sreal);
        ++launches;
    } else {
        JGB_CUDA(cudaMemcpyAsync(vm_out, dvm, (size_t)Sreal * n * sizeof(double), cudaMemcpyDeviceToHost, stream));
        JGB_CUDA(cudaMemcpyAsync(va_out, dva, (size_t)Sreal * n * sizeof(double), cudaMemcpyDeviceToHost, stream));
    }
    b_iters.download(hit.data(), Sreal, stream);
    b_status.download(hst.data(), Sreal, stream);
    JGB_CUDA(cudaStreamSynchronize(stream));
    int64_t tot = 0;
    int worst = 0;
    for (int s = 0; s < Sreal; ++s) {
        tot += hit[s];
        if (!dev_out) {
            if (iters_out) iters_out[s] = hit[s];
            if (status_out) status_out[s] = (int8_t)hst[s];
        }
        if (hst[s] != 0) worst = 1;
    }
    if (total) *total = tot;
    return worst;
}

double NrContext::stat(const std::string& key) {
    if (key.rfind("nr.batch.", 0) == 0) {
        const Symbolic& b = solver_batch.sym;
        if (key == "nr.batch.u_size") return (double)b.u_size;
        if (key == "nr.batch.upd_size") return (double)b.upd_size;
        if (key == "nr.batch.fronts") return b.nfronts;
        if (key == "nr.batch.levels") return b.nlevels;
        if (key == "nr.batch.flops") return b.flops;
        if (key == "nr.batch.nnz_lu") return (double)b.nnz_lu;
        if (key == "nr.batch.max_front") return b.max_front;
        if (key == "nr.batch.factor_launches") return solver_batch.factor_launches(32);
        if (key == "nr.batch.task_fronts") { solver_batch.factor_launches(32); return solver_batch.task_fronts; }
        if (key == "nr.batch.task_upd_on_chip") { solver_batch.factor_launches(32); return (double)solver_batch.task_upd_on_chip; }
        if (key == "nr.batch.launches_per_solve") return solver_batch.launches_per_solve(32);
        return -1.0;
    }
    const Symbolic& s = solver.sym;
    if (key == "nr.weak_pivot_scenarios") return (double)weak_events;
    if (key == "nr.refine_calls") return (double)refine_calls;
    if (key == "nr.time.assemble_ms") return timer.ms[kPhAssemble];
    if (key == "nr.time.factor_ms") return timer.ms[kPhFactor];
    if (key == "nr.time.backsolve_ms") return timer.ms[kPhBacksolve];
    if (key == "nr.time.assemble_count") return (double)timer.count[kPhAssemble];
    if (key == "nr.time.factor_count") return (double)timer.count[kPhFactor];
    if (key == "nr.factor_launches") return solver.factor_launches(32);
    if (key == "nr.u_size") return (double)s.u_size;
    if (key == "nr.upd_size") return (double)s.upd_size;
    if (key == "nr.nnz_lu") return (double)s.nnz_lu;
    if (key == "nr.fronts") return s.nfronts;
    if (key == "nr.levels") return s.nlevels;
    if (key == "nr.flops") return s.flops;
    if (key == "nr.max_front") return s.max_front;
    if (key == "nr.dim") return dim;
    if (key == "nr.nnz_j") return nnzj;
    if (key == "nr.nnz_y") return nnzy;
    if (key == "nr.launches_per_iteration") return solver.launches_per_solve(1) + 3;
    if (key == "nr.launches_per_iteration_batch") return solver.launches_per_solve(32) + 3;
    // per scenario-iteration algorithmic bytes: Y strip (2 complex + int32 index per entry, shared across a batch
    // but counted once here), per-bus state / injections / type, mismatch and Jacobian writes
    if (key == "nr.assemble_bytes") return 36.0 * nnzy + 4.0 * (n + 1) + 49.0 * n + 8.0 * dim + 8.0 * nnzj;
    if (key == "nr.assemble_bytes_batch") return 16.0 * n + 8.0 * dim + 8.0 * nnzj;
    if (key == "nr.solve_bytes") return (double)solver.factor_bytes(1);
    return -1.0;
}

}  // namespace jgb
