// Gauss-Newton WLS kernels and host driver. See wls.cuh.
#include "wls.cuh"

#include <algorithm>
#include <cmath>
#include <map>

namespace jgb {

namespace {

constexpr int kRowBlock = 128;

struct Br {
    int i, j;
    double g, b, gsi, bsi, tinv, phi;
};

__device__ __forceinline__ Br load_branch(const WlsDev& d, int k) {
    Br r;
    r.i = d.br_from[k]; r.j = d.br_to[k];
    r.g = d.br_g[k]; r.b = d.br_b[k]; r.gsi = d.br_gsi[k]; r.bsi = d.br_bsi[k];
    r.tinv = d.br_tinv[k]; r.phi = d.br_phi[k];
    return r;
}

// K7: one thread per (measurement row, scenario): residual r = z - h(x) and the row's H entries
// (normalEquation!, acStateEstimation.jl:261-583; formulas backend/equations.jl:20-573).
// Row-parallel: every row evaluates its function once (the reference sweeps H's theta columns and evaluates
// branch rows twice); slot order of the H entries: branch rows [theta_i, V_i, theta_j, V_j], bus-injection rows
// [theta_j, V_j] per Ybus strip entry, phasor rows [theta_i, V_i], constant rows (codes 1, 12, 13) are untouched.
__global__ void __launch_bounds__(kRowBlock)
wls_rows_kernel(WlsDev d, int S) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int row = (int)(gid / S), s = (int)(gid % S);
    if (row >= d.m) return;
    if (d.active && !d.active[s]) return;
    const int code = d.type[row];
    if (code == 0) return;
    const int k = d.index[row];
    const double z = d.z[(long long)row * S + s];
    double* __restrict__ H = d.hval + s;
    const int* __restrict__ slot = d.slotpos + d.slotptr[row];
    double h;
    if (code == 1 || code == 12) {
        h = d.vm[(long long)k * S + s];
    } else if (code == 13) {
        h = d.va[(long long)k * S + s];
    } else if (code == 6 || code == 9) {
        const int i = k;
        const double Vi = d.vm[(long long)i * S + s], Ti = d.va[(long long)i * S + s];
        double sum_plus = 0.0, sum_minus = 0.0;
        const int p0 = d.ycolptr[i], p1 = d.ycolptr[i + 1];
        const double2 yd = d.y[d.ydiag[i]];
        int dslot = 0;
        for (int p = p0; p < p1; ++p) {
            const int j = d.yrow[p];
            const double2 yt = d.yt[p];     // Y[i, j]
            const double Vj = d.vm[(long long)j * S + s], Tj = d.va[(long long)j * S + s];
            double sn, cs;
            sincos(Ti - Tj, &sn, &cs);
            sum_plus += Vj * (yt.x * cs + yt.y * sn);
            sum_minus += Vj * (yt.x * sn - yt.y * cs);
            if (j == i) { dslot = 2 * (p - p0); continue; }
            if (code == 6) {
                H[(long long)slot[2 * (p - p0)] * S] = Vi * Vj * (yt.x * sn - yt.y * cs);          // Pi_theta_j
                H[(long long)slot[2 * (p - p0) + 1] * S] = Vi * (yt.x * cs + yt.y * sn);           // Pi_V_j
            } else {
                H[(long long)slot[2 * (p - p0)] * S] = -Vi * Vj * (yt.x * cs + yt.y * sn);         // Qi_theta_j
                H[(long long)slot[2 * (p - p0) + 1] * S] = Vi * (yt.x * sn - yt.y * cs);           // Qi_V_j
            }
        }
        if (code == 6) {
            h = Vi * sum_plus;
            H[(long long)slot[dslot] * S] = Vi * (-sum_minus) - yd.y * (Vi * Vi);                  // Pi_theta_i
            H[(long long)slot[dslot + 1] * S] = sum_plus + yd.x * Vi;                              // Pi_V_i
        } else {
            h = Vi * sum_minus;
            H[(long long)slot[dslot] * S] = Vi * sum_plus - yd.x * (Vi * Vi);                      // Qi_theta_i
            H[(long long)slot[dslot + 1] * S] = sum_minus - yd.y * Vi;                             // Qi_V_i
        }
    } else if (code == 16 || code == 17) {
        const double Vi = d.vm[(long long)k * S + s], Ti = d.va[(long long)k * S + s];
        double sn, cs;
        sincos(Ti, &sn, &cs);
        if (code == 16) {
            h = Vi * cs;
            H[(long long)slot[0] * S] = -Vi * sn;
            H[(long long)slot[1] * S] = cs;
        } else {
            h = Vi * sn;
            H[(long long)slot[0] * S] = Vi * cs;
            H[(long long)slot[1] * S] = sn;
        }
    } else {
        const Br br = load_branch(d, k);
        const double Vi = d.vm[(long long)br.i * S + s], Vj = d.vm[(long long)br.j * S + s];
        const double Ti = d.va[(long long)br.i * S + s], Tj = d.va[(long long)br.j * S + s];
        const double g = br.g, b = br.b, gsi = br.gsi, bsi = br.bsi, tinv = br.tinv;
        double sn, cs;
        sincos(Ti - Tj - br.phi, &sn, &cs);
        double dti, dvi, dtj, dvj;
        switch (code) {
        case 7: {
            const double A = tinv * tinv * (g + gsi), B = tinv * g, C = tinv * b;
            h = A * (Vi * Vi) - (B * cs + C * sn) * Vi * Vj;
            dti = (B * sn - C * cs) * Vi * Vj;
            dvi = 2 * A * Vi - (B * cs + C * sn) * Vj;
            dtj = -dti;
            dvj = -(B * cs + C * sn) * Vi;
        } break;
        case 8: {
            const double A = g + gsi, B = tinv * g, C = tinv * b;
            h = A * (Vj * Vj) - (B * cs - C * sn) * Vi * Vj;
            dti = (B * sn + C * cs) * Vi * Vj;
            dvi = (-B * cs + C * sn) * Vj;
            dtj = -dti;
            dvj = 2 * A * Vj - (B * cs - C * sn) * Vi;
        } break;
        case 10: {
            const double A = tinv * tinv * (b + bsi), B = tinv * g, C = tinv * b;
            h = -A * (Vi * Vi) - (B * sn - C * cs) * Vi * Vj;
            dti = -(B * cs + C * sn) * Vi * Vj;
            dvi = -2 * A * Vi - (B * sn - C * cs) * Vj;
            dtj = -dti;
            dvj = -(B * sn - C * cs) * Vi;
        } break;
        case 11: {
            const double A = b + bsi, B = tinv * g, C = tinv * b;
            h = -A * (Vj * Vj) + (B * sn + C * cs) * Vi * Vj;
            dti = (B * cs - C * sn) * Vi * Vj;
            dvi = (B * sn + C * cs) * Vj;
            dtj = -dti;
            dvj = -2 * A * Vj + (B * sn + C * cs) * Vi;
        } break;
        case 2: case 4: case 14: {
            const double t2 = tinv * tinv;
            const double A = t2 * t2 * ((g + gsi) * (g + gsi) + (b + bsi) * (b + bsi));
            const double B = t2 * (g * g + b * b);
            const double C = t2 * tinv * (g * (g + gsi) + b * (b + bsi));
            const double D = t2 * tinv * (g * bsi - b * gsi);
            if (code == 2) {
                const double iinv = 1 / (sqrt(A * (Vi * Vi) + B * (Vj * Vj) - 2 * Vi * Vj * (C * cs - D * sn)));
                h = 1 / iinv;
                dti = iinv * (C * sn + D * cs) * Vi * Vj;
                dvi = iinv * (A * Vi - (C * cs - D * sn) * Vj);
                dtj = -dti;
                dvj = iinv * (B * Vj - (C * cs - D * sn) * Vi);
            } else if (code == 4) {
                h = A * (Vi * Vi) + B * (Vj * Vj) - 2 * Vi * Vj * (C * cs - D * sn);
                dti = 2 * (C * sn + D * cs) * Vi * Vj;
                dvi = 2 * (A * Vi - (C * cs - D * sn) * Vj);
                dtj = -dti;
                dvj = 2 * (B * Vj - (C * cs - D * sn) * Vi);
            } else {
                const double pA = t2 * (g + gsi), pB = t2 * (b + bsi), pC = tinv * g, pD = tinv * b;
                double si, ci, sj, cj;
                sincos(Ti, &si, &ci);
                sincos(Tj + br.phi, &sj, &cj);
                const double re = (pA * ci - pB * si) * Vi - (pC * cj - pD * sj) * Vj;
                const double im = (pA * si + pB * ci) * Vi - (pC * sj + pD * cj) * Vj;
                const double iinv2 = 1 / (re * re + im * im);
                h = atan2(im, re);
                dti = iinv2 * (A * (Vi * Vi) - (C * cs - D * sn) * Vi * Vj);
                dvi = -iinv2 * (C * sn + D * cs) * Vj;
                dtj = iinv2 * (B * (Vj * Vj) - (C * cs - D * sn) * Vi * Vj);
                dvj = iinv2 * (C * sn + D * cs) * Vi;
            }
        } break;
        case 3: case 5: case 15: {
            const double A = tinv * tinv * (g * g + b * b);
            const double B = (g + gsi) * (g + gsi) + (b + bsi) * (b + bsi);
            const double C = tinv * (g * (g + gsi) + b * (b + bsi));
            const double D = tinv * (g * bsi - gsi * b);
            if (code == 3) {
                const double iinv = 1 / sqrt(A * (Vi * Vi) + B * (Vj * Vj) - 2 * Vi * Vj * (C * cs + D * sn));
                h = 1 / iinv;
                dti = iinv * (C * sn - D * cs) * Vi * Vj;
                dvi = iinv * (A * Vi - (C * cs + D * sn) * Vj);
                dtj = -dti;
                dvj = iinv * (B * Vj - (C * cs + D * sn) * Vi);
            } else if (code == 5) {
                h = A * (Vi * Vi) + B * (Vj * Vj) - 2 * Vi * Vj * (C * cs + D * sn);
                dti = 2 * (C * sn - D * cs) * Vi * Vj;
                dvi = 2 * (A * Vi - (C * cs + D * sn) * Vj);
                dtj = -dti;
                dvj = 2 * (B * Vj - (C * cs + D * sn) * Vi);
            } else {
                const double pA = g + gsi, pB = b + bsi, pC = tinv * g, pD = tinv * b;
                double si, ci, sj, cj;
                sincos(Ti - br.phi, &si, &ci);
                sincos(Tj, &sj, &cj);
                const double re = (pA * cj - pB * sj) * Vj - (pC * ci - pD * si) * Vi;
                const double im = (pA * sj + pB * cj) * Vj - (pC * si + pD * ci) * Vi;
                const double iinv2 = 1 / (re * re + im * im);
                h = atan2(im, re);
                dti = iinv2 * (A * (Vi * Vi) - (C * cs + D * sn) * Vi * Vj);
                dvi = -iinv2 * (C * sn - D * cs) * Vj;
                dtj = iinv2 * (B * (Vj * Vj) - (C * cs + D * sn) * Vi * Vj);
                dvj = iinv2 * (C * sn - D * cs) * Vi;
            }
        } break;
        case 18: case 20: {
            const double t2 = tinv * tinv;
            const double pA = t2 * (g + gsi), pB = t2 * (b + bsi), pC = tinv * g, pD = tinv * b;
            double si, ci, sj, cj;
            sincos(Ti, &si, &ci);
            sincos(Tj + br.phi, &sj, &cj);
            if (code == 18) {
                h = (pA * ci - pB * si) * Vi - (pC * cj - pD * sj) * Vj;
                dti = -(pA * si + pB * ci) * Vi;
                dvi = pA * ci - pB * si;
                dtj = (pC * sj + pD * cj) * Vj;
                dvj = -pC * cj + pD * sj;
            } else {
                h = (pA * si + pB * ci) * Vi - (pC * sj + pD * cj) * Vj;
                dti = (pA * ci - pB * si) * Vi;
                dvi = pA * si + pB * ci;
                dtj = (-pC * cj + pD * sj) * Vj;
                dvj = -pC * sj - pD * cj;
            }
        } break;
        case 19: case 21: {
            const double pA = g + gsi, pB = b + bsi, pC = tinv * g, pD = tinv * b;
            double si, ci, sj, cj;
            sincos(Ti - br.phi, &si, &ci);
            sincos(Tj, &sj, &cj);
            if (code == 19) {
                h = (pA * cj - pB * sj) * Vj - (pC * ci - pD * si) * Vi;
                dti = (pC * si + pD * ci) * Vi;
                dvi = -pC * ci + pD * si;
                dtj = -(pA * sj + pB * cj) * Vj;
                dvj = pA * cj - pB * sj;
            } else {
                h = (pA * sj + pB * cj) * Vj - (pC * si + pD * ci) * Vi;
                dti = (-pC * ci + pD * si) * Vi;
                dvi = -pC * si - pD * ci;
                dtj = (pA * cj - pB * sj) * Vj;
                dvj = pA * sj + pB * cj;
            }
        } break;
        default:
            return;
        }
        H[(long long)slot[0] * S] = dti;
        H[(long long)slot[1] * S] = dvi;
        H[(long long)slot[2] * S] = dtj;
        H[(long long)slot[3] * S] = dvj;
    }
    d.res[(long long)row * S + s] = z - h;
}

// Objective partial sums (seobjective, equations.jl:689-698), deterministic two-stage reduction.
// blockDim = (TS scenario lanes, TB row lanes); block b covers rows [b*rpb, (b+1)*rpb).
__global__ void __launch_bounds__(128)
wls_objective_kernel(WlsDev d, int S, int rpb) {
    __shared__ double red[128];
    const int s = blockIdx.y * blockDim.x + threadIdx.x;
    const bool act = d.active ? (d.active[s] != 0) : true;
    double acc = 0.0;
    const int rend = min(d.m, (int)(blockIdx.x + 1) * rpb);
    if (act) {
        for (int row = blockIdx.x * rpb + threadIdx.y; row < rend; row += blockDim.y) {
            const int code = d.type[row];
            if (code == 0) continue;
            const double r = d.res[(long long)row * S + s];
            acc += r * r * d.wdiag[row];
            const double wo = d.woff[row];
            if (wo != 0.0) acc += 2 * r * d.res[(long long)(row - 1) * S + s] * wo;
        }
    }
    const int t = threadIdx.y * blockDim.x + threadIdx.x;
    red[t] = acc;
    __syncthreads();
    for (int h = blockDim.y / 2; h >= 1; h >>= 1) {
        if ((int)threadIdx.y < h) red[t] += red[(threadIdx.y + h) * blockDim.x + threadIdx.x];
        __syncthreads();
    }
    if (threadIdx.y == 0) d.objpart[(long long)blockIdx.x * S + s] = red[t];
}

__global__ void wls_objective_final_kernel(WlsDev d, int S, int nblocks) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    if (d.active && !d.active[s]) return;
    double acc = 0.0;
    for (int b = 0; b < nblocks; ++b) acc += d.objpart[(long long)b * S + s];
    d.obj[s] = acc;
}

// K8a: G = H'WH on its fixed pattern, one thread per (lower-triangular entry, scenario); terms are summed in
// ascending measurement-row order like Julia's SpGEMM does, so the result does not depend on scheduling.
// The slack angle row/column gets no terms and a unit diagonal (acStateEstimation.jl:885-889).
__global__ void __launch_bounds__(128)
wls_gain_kernel(WlsDev d, int S) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int e = (int)(gid / S), s = (int)(gid % S);
    if (e >= d.nlow) return;
    if (d.active && !d.active[s]) return;
    const double* __restrict__ H = d.hval + s;
    double acc = 0.0;
    const int t1 = d.gentry_ptr[e + 1];
    for (int t = d.gentry_ptr[e]; t < t1; ++t) {
        const int w = d.gterm_w[t];
        const double wv = (w < d.m) ? d.wdiag[w] : d.woff[w - d.m];
        acc += (H[(long long)d.gterm_a[t] * S] * wv) * H[(long long)d.gterm_b[t] * S];
    }
    const int lp = d.glow_pos[e];
    if (lp == d.gslack_pos) acc = 1.0;
    d.gval[(long long)lp * S + s] = acc;
    const int up = d.gup_pos[e];
    if (up != lp) d.gval[(long long)up * S + s] = acc;
}

// K8b: rhs = H'W r, one thread per (state variable, scenario), walking H's CSC column.
__global__ void __launch_bounds__(128)
wls_rhs_kernel(WlsDev d, int S) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int a = (int)(gid / S), s = (int)(gid % S);
    if (a >= d.nv) return;
    if (d.active && !d.active[s]) return;
    double acc = 0.0;
    if (a != d.slack) {
        const int p1 = d.hcolptr[a + 1];
        for (int p = d.hcolptr[a]; p < p1; ++p) {
            const int r = d.hrow[p];
            double wr = d.wdiag[r] * d.res[(long long)r * S + s];
            const double wo = d.woff[r];
            if (wo != 0.0) wr += wo * d.res[(long long)(r - 1) * S + s];
            if (r + 1 < d.m) {
                const double wn = d.woff[r + 1];
                if (wn != 0.0) wr += wn * d.res[(long long)(r + 1) * S + s];
            }
            acc += d.hval[(long long)p * S + s] * wr;
        }
    }
    d.rhs[(long long)a * S + s] = acc;
}

// increment[slack] = 0 and max |increment| (acStateEstimation.jl:899-903)
__global__ void __launch_bounds__(128)
wls_maxinc_kernel(WlsDev d, int S, int vpb) {
    __shared__ double red[128];
    const int s = blockIdx.y * blockDim.x + threadIdx.x;
    const bool act = d.active ? (d.active[s] != 0) : true;
    double mx = 0.0;
    const int aend = min(d.nv, (int)(blockIdx.x + 1) * vpb);
    if (act) {
        for (int a = blockIdx.x * vpb + threadIdx.y; a < aend; a += blockDim.y) {
            double v = d.inc[(long long)a * S + s];
            if (a == d.slack) { v = 0.0; d.inc[(long long)a * S + s] = 0.0; }
            v = fabs(v);
            if (v != v) v = INFINITY;
            mx = fmax(mx, v);
        }
    }
    const int t = threadIdx.y * blockDim.x + threadIdx.x;
    red[t] = mx;
    __syncthreads();
    for (int h = blockDim.y / 2; h >= 1; h >>= 1) {
        if ((int)threadIdx.y < h) red[t] = fmax(red[t], red[(threadIdx.y + h) * blockDim.x + threadIdx.x]);
        __syncthreads();
    }
    if (threadIdx.y == 0 && act) atomicMax(&d.maxbits[s], (unsigned long long)__double_as_longlong(red[t]));
}

// convergence bookkeeping of stateEstimation! (acStateEstimation.jl:1303-1315); tol < 0: publish only
__global__ void wls_check_kernel(WlsDev d, int S, int Sreal, double tol, int max_iter) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    if (d.active && !d.active[s]) return;
    const double mi = __longlong_as_double((long long)d.maxbits[s]);
    d.maxinc[s] = mi;
    d.maxbits[s] = 0ull;
    if (tol < 0.0 || s >= Sreal) return;
    if (d.status[s] < 0) { d.active[s] = 0; return; }
    if (mi < tol) { d.active[s] = 0; d.status[s] = 0; return; }
    if (d.iters[s] == max_iter) { d.active[s] = 0; d.status[s] = 1; return; }
    atomicAdd(d.remaining, 1);
}

// solve! (acStateEstimation.jl:1035-1047)
__global__ void wls_update_kernel(WlsDev d, int S) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int i = (int)(gid / S), s = (int)(gid % S);
    if (i >= d.n) return;
    if (d.active && !d.active[s]) return;
    d.va[gid] += d.inc[(long long)i * S + s];
    d.vm[gid] += d.inc[(long long)(i + d.n) * S + s];
    if (i == 0) d.iters[s] += 1;
}

__global__ void wls_broadcast_kernel(const double* __restrict__ src, double* __restrict__ dst, long long count, int S) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= count * S) return;
    dst[gid] = src[gid / S];
}

// [rows][S] <-> [Sreal][rows] transposes through shared memory
__global__ void wls_transpose_out_kernel(const double* __restrict__ src, double* __restrict__ dst, int n, int S,
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

__global__ void wls_transpose_in_kernel(const double* __restrict__ src, double* __restrict__ dst, int m, int S,
                                        int Sreal) {
    __shared__ double tile[32][33];
    const int r0 = blockIdx.x * 32, s0 = blockIdx.y * 32;
    for (int q = threadIdx.y; q < 32; q += blockDim.y) {
        const int s = s0 + q, r = r0 + threadIdx.x;
        tile[q][threadIdx.x] = (s < Sreal && r < m) ? src[(long long)s * m + r] : 0.0;
    }
    __syncthreads();
    for (int q = threadIdx.y; q < 32; q += blockDim.y) {
        const int r = r0 + q, s = s0 + threadIdx.x;
        if (r < m && s < S) dst[(long long)r * S + s] = tile[threadIdx.x][q];
    }
}

__global__ void wls_init_batch_kernel(unsigned char* active, int* status, int* iters, unsigned long long* maxbits,
                                      int S, int Sreal) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    active[s] = s < Sreal;
    status[s] = s < Sreal ? 1 : 0;
    iters[s] = 0;
    maxbits[s] = 0ull;
}

__global__ void wls_copy_results_kernel(const int* __restrict__ iters, const int* __restrict__ status,
                                        const double* __restrict__ obj, int32_t* __restrict__ it_out,
                                        int8_t* __restrict__ st_out, double* __restrict__ obj_out, int Sreal) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= Sreal) return;
    it_out[s] = iters[s];
    st_out[s] = (int8_t)status[s];
    if (obj_out) obj_out[s] = obj[s];
}

}  // namespace

// ---- bad-data post-step: c[i] = h_i G^-1 h_i' from the selected inverse of the gain factor ----
// rowProjection (stateEstimation/badData.jl:349-362, 477-499). One thread per measurement row; its pairs of H entries
// (a <= b in slot order, slack column skipped) and the offsets of (G^-1)[a,b] inside the per-front blocks of the
// selected inverse were listed on the host; fixed order, no atomics.
__global__ void wls_projection_kernel(int m, const int* __restrict__ pair_ptr, const int* __restrict__ pair_pa,
                                      const int* __restrict__ pair_pb, const long long* __restrict__ pair_z,
                                      const double* __restrict__ hval, const double* __restrict__ Z,
                                      double* __restrict__ c) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= m) return;
    double acc = 0.0;
    for (int t = pair_ptr[row]; t < pair_ptr[row + 1]; ++t) {
        const int pa = pair_pa[t], pb = pair_pb[t];
        const double v = hval[pa] * hval[pb] * Z[pair_z[t]];
        acc += (pa == pb) ? v : 2.0 * v;
    }
    c[row] = acc;
}

// residualTest! bookkeeping on the device (badData.jl:258-282): the row leaves the model — H entries, mean and
// residual zeroed, type 0.
__global__ void wls_remove_row_kernel(int row, const int* __restrict__ slotpos, int s0, int s1, double* hval,
                                      double* res, double* z, signed char* type) {
    for (int t = s0 + threadIdx.x; t < s1; t += blockDim.x) hval[slotpos[t]] = 0.0;
    if (threadIdx.x == 0) { res[row] = 0.0; z[row] = 0.0; type[row] = 0; }
}

void WlsContext::setup(int64_t n_, int64_t m_, int64_t slack_, const int64_t* hcp, const int64_t* hrv,
                       const int8_t* type, const int64_t* index, const int64_t* range6, const int64_t* wcp,
                       const int64_t* wrv, const double* wnz, const int64_t* ycp, const int64_t* yrv,
                       const double* y, const double* yt, int64_t nbr_, const int64_t* from, const int64_t* to,
                       const double* cond, const double* susc, const double* tap, const double* shift,
                       const double* adm) {
    if (n_ <= 0 || m_ <= 0 || !hcp || !hrv || !type || !index || !wcp || !wrv || !wnz || !ycp || !yrv || !y || !yt)
        throw std::invalid_argument("wls_setup: null or empty input");
    if (nbr_ > 0 && (!from || !to || !cond || !susc || !tap || !shift || !adm))
        throw std::invalid_argument("wls_setup: null branch input");
    if (slack_ < 1 || slack_ > n_) throw std::invalid_argument("wls_setup: slack index out of range");
    (void)range6;
    n = (int)n_; m = (int)m_; slack = (int)slack_ - 1; nbr = (int)nbr_;
    const int nv = 2 * n;
    nnzh = (int)(hcp[nv] - 1);
    nnzy = (int)(ycp[n] - 1);
    // ---- Ybus
    std::vector<int> ycolptr(n + 1), yrow(nnzy), ydiag(n, -1);
    for (int i = 0; i <= n; ++i) ycolptr[i] = (int)(ycp[i] - 1);
    for (int q = 0; q < nnzy; ++q) yrow[q] = (int)(yrv[q] - 1);
    for (int c = 0; c < n; ++c)
        for (int q = ycolptr[c]; q < ycolptr[c + 1]; ++q)
            if (yrow[q] == c) ydiag[c] = q;
    for (int c = 0; c < n; ++c)
        if (ydiag[c] < 0) throw std::invalid_argument("wls_setup: Ybus has no stored diagonal entry");
    // ---- H: CSC (reference order) and per-row position lookup
    std::vector<int> hcolptr(nv + 1), hrow(nnzh);
    for (int c = 0; c <= nv; ++c) hcolptr[c] = (int)(hcp[c] - 1);
    for (int q = 0; q < nnzh; ++q) {
        if (hrv[q] < 1 || hrv[q] > m) throw std::invalid_argument("wls_setup: H row index out of range");
        hrow[q] = (int)(hrv[q] - 1);
    }
    std::vector<std::vector<std::pair<int, int>>> rowent(m);   // (col, csc position), ascending col
    for (int c = 0; c < nv; ++c)
        for (int q = hcolptr[c]; q < hcolptr[c + 1]; ++q) rowent[hrow[q]].push_back({c, q});
    auto hpos = [&](int row, int col) -> int {
        for (auto& e : rowent[row])
            if (e.first == col) return e.second;
        throw std::runtime_error("wls_setup: H pattern does not hold the entry a measurement row needs (-4)");
    };
    std::vector<int> idx0(m), br_from(nbr), br_to(nbr);
    for (int k = 0; k < nbr; ++k) { br_from[k] = (int)from[k] - 1; br_to[k] = (int)to[k] - 1; }
    std::vector<int> slotptr(m + 1, 0), slotpos;
    slotpos.reserve(nnzh);
    h_const.assign(nnzh, 0.0);
    for (int r = 0; r < m; ++r) {
        const int code = type[r];
        const int k = (int)index[r] - 1;
        idx0[r] = k;
        if (code < 0 || code > 21) throw std::invalid_argument("wls_setup: unknown measurement code");
        const bool bus_row = (code == 0) ? false : (code == 1 || code == 6 || code == 9 || code == 12 || code == 13 ||
                                                    code == 16 || code == 17);
        if (code != 0 && (k < 0 || k >= (bus_row ? n : nbr)))
            throw std::invalid_argument("wls_setup: measurement index out of range");
        if (code == 1 || code == 12) {
            int q = hpos(r, k + n);
            slotpos.push_back(q);
            h_const[q] = 1.0;
        } else if (code == 13) {
            int q = hpos(r, k);
            slotpos.push_back(q);
            h_const[q] = 1.0;
        } else if (code == 6 || code == 9) {
            for (int q = ycolptr[k]; q < ycolptr[k + 1]; ++q) {
                slotpos.push_back(hpos(r, yrow[q]));
                slotpos.push_back(hpos(r, yrow[q] + n));
            }
        } else if (code == 16 || code == 17) {
            slotpos.push_back(hpos(r, k));
            slotpos.push_back(hpos(r, k + n));
        } else if (code != 0) {
            const int i = br_from[k], j = br_to[k];
            slotpos.push_back(hpos(r, i));
            slotpos.push_back(hpos(r, i + n));
            slotpos.push_back(hpos(r, j));
            slotpos.push_back(hpos(r, j + n));
        }
        slotptr[r + 1] = (int)slotpos.size();
    }
    // ---- W: diagonal + 2x2 blocks of correlated rectangular PMU pairs
    std::vector<double> wdiag(m, 0.0), woff(m, 0.0);
    for (int c = 0; c < m; ++c)
        for (int64_t q = wcp[c] - 1; q < wcp[c + 1] - 1; ++q) {
            int r = (int)wrv[q] - 1;
            if (r == c) wdiag[c] = wnz[q];
            else if (r == c + 1) woff[r] = wnz[q];            // W[c+1, c]
            else if (r == c - 1) { /* symmetric twin */ }
            else throw std::invalid_argument("wls_setup: precision matrix is not block diagonal (1x1 / 2x2)");
        }
    // ---- gain pattern and gather lists. Row groups: single rows, or (r-1, r) pairs when woff[r] != 0.
    struct Term { int a, b, w; };
    std::vector<std::map<int, std::vector<Term>>> low(nv);   // low[b][a], a >= b (column b, row a)
    std::vector<std::vector<int>> gcols(nv);                  // structural pattern incl. slack row/col
    auto add_terms = [&](int r1, int r2, int w) {
        // contribution H[r1,a] * W[r1,r2] * H[r2,b] to G[a,b]
        for (auto& ea : rowent[r1])
            for (auto& eb : rowent[r2]) {
                const int a = ea.first, b = eb.first;
                gcols[b].push_back(a);
                if (a < b) continue;
                if (a == slack || b == slack) continue;
                low[b][a].push_back({ea.second, eb.second, w});
            }
    };
    for (int r = 0; r < m; ++r) {
        add_terms(r, r, r);
        if (woff[r] != 0.0) {
            add_terms(r, r - 1, m + r);
            add_terms(r - 1, r, m + r);
        }
    }
    gcols[slack].push_back(slack);
    std::vector<int> gcolptr(nv + 1, 0), grow;
    for (int c = 0; c < nv; ++c) {
        auto& v = gcols[c];
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
        gcolptr[c + 1] = gcolptr[c] + (int)v.size();
        grow.insert(grow.end(), v.begin(), v.end());
    }
    nnzg = gcolptr[nv];
    auto gpos = [&](int row, int col) -> int {
        auto b = grow.begin() + gcolptr[col], e = grow.begin() + gcolptr[col + 1];
        auto it = std::lower_bound(b, e, row);
        if (it == e || *it != row) throw std::runtime_error("wls_setup: gain pattern lookup failed");
        return (int)(it - grow.begin());
    };
    gcolptr1.resize(nv + 1);
    growval1.resize(nnzg);
    for (int c = 0; c <= nv; ++c) gcolptr1[c] = gcolptr[c] + 1;
    for (int q = 0; q < nnzg; ++q) growval1[q] = grow[q] + 1;
    // every lower entry of the structural pattern gets a record (possibly with no terms: slack row/col)
    std::vector<int> gentry_ptr{0}, gterm_a, gterm_b, gterm_w, glow_pos, gup_pos;
    for (int b = 0; b < nv; ++b)
        for (int q = gcolptr[b]; q < gcolptr[b + 1]; ++q) {
            const int a = grow[q];
            if (a < b) continue;
            auto it = low[b].find(a);
            if (it != low[b].end()) {
                auto& ts = it->second;
                // ascending measurement row of the right factor, as Julia's Gustavson SpGEMM accumulates
                std::stable_sort(ts.begin(), ts.end(), [&](const Term& x, const Term& y2) { return hrow[x.b] < hrow[y2.b]; });
                for (auto& t : ts) { gterm_a.push_back(t.a); gterm_b.push_back(t.b); gterm_w.push_back(t.w); }
            }
            gentry_ptr.push_back((int)gterm_a.size());
            glow_pos.push_back(q);
            gup_pos.push_back(gpos(b, a));
        }
    nlow = (int)glow_pos.size();
    nterms = (long long)gterm_a.size();
    gslack_pos = gpos(slack, slack);
    // ---- symbolic factorisation of G: theta_i and V_i of one bus form a supervariable
    std::vector<int> group(nv);
    for (int i = 0; i < n; ++i) { group[i] = i; group[i + n] = i; }
    {
        Symbolic sym;
        analyse(nv, gcolptr.data(), grow.data(), group.data(), nullptr, latency_options(), sym);
        solver.setup(sym, stream, true);
        Symbolic symb;
        analyse(nv, gcolptr.data(), grow.data(), group.data(), nullptr, throughput_options(), symb);
        solver_batch.setup(symb, stream, true);
    }
    // ---- branch coefficients
    std::vector<double> bg(nbr), bb(nbr), bgsi(nbr), bbsi(nbr), btinv(nbr), bphi(nbr);
    for (int k = 0; k < nbr; ++k) {
        bg[k] = adm[2 * k]; bb[k] = adm[2 * k + 1];
        bgsi[k] = 0.5 * cond[k]; bbsi[k] = 0.5 * susc[k];
        btinv[k] = 1 / tap[k]; bphi[k] = shift[k];
    }
    // ---- upload
    d_ycolptr.upload(ycolptr, stream); d_yrow.upload(yrow, stream); d_ydiag.upload(ydiag, stream);
    d_y.upload(reinterpret_cast<const double2*>(y), nnzy, stream);
    d_yt.upload(reinterpret_cast<const double2*>(yt), nnzy, stream);
    d_br_from.upload(br_from, stream); d_br_to.upload(br_to, stream);
    d_br_g.upload(bg, stream); d_br_b.upload(bb, stream); d_br_gsi.upload(bgsi, stream);
    d_br_bsi.upload(bbsi, stream); d_br_tinv.upload(btinv, stream); d_br_phi.upload(bphi, stream);
    d_type.upload(reinterpret_cast<const signed char*>(type), m, stream);
    d_index.upload(idx0, stream); d_slotptr.upload(slotptr, stream); d_slotpos.upload(slotpos, stream);
    d_wdiag.upload(wdiag, stream); d_woff.upload(woff, stream);
    d_hcolptr.upload(hcolptr, stream); d_hrow.upload(hrow, stream);
    d_gentry_ptr.upload(gentry_ptr, stream); d_gterm_a.upload(gterm_a, stream); d_gterm_b.upload(gterm_b, stream);
    d_gterm_w.upload(gterm_w, stream); d_glow_pos.upload(glow_pos, stream); d_gup_pos.upload(gup_pos, stream);
    nrowblocks = ceil_div(m, 1024);
    d_vm.alloc(n); d_va.alloc(n); d_z.alloc(m); d_res.alloc(m); d_hval.alloc(nnzh); d_gval.alloc(nnzg);
    d_rhs.alloc(nv); d_inc.alloc(nv); d_objpart.alloc(nrowblocks); d_obj.alloc(1); d_maxinc.alloc(1);
    d_maxbits.alloc(1); d_active.alloc(1); d_status.alloc(1); d_iters.alloc(1); d_remaining.alloc(1);
    d_hval.upload(h_const, stream);
    d_res.zero(stream); d_gval.zero(stream); d_rhs.zero(stream); d_inc.zero(stream); d_obj.zero(stream);
    d_maxinc.zero(stream); d_maxbits.zero(stream); d_status.zero(stream); d_iters.zero(stream);
    d_remaining.zero(stream);
    JGB_CUDA(cudaMemsetAsync(d_active.p, 1, 1, stream));
    h_d.alloc(4);
    h_i.alloc(4);
    JGB_CUDA(cudaStreamSynchronize(stream));
    iteration = 0;
    have_mean = have_state = false;
    batch_S = 0;
    // host copies for row updates (update_rows) and the lazily built bad-data lists (residual_test)
    h_rowent = rowent;
    h_ycolptr = ycolptr;
    h_yrow = yrow;
    h_brfrom = br_from;
    h_brto = br_to;
    h_type.assign(type, type + m);
    h_index = idx0;
    h_woff = woff;
    h_slotptr = slotptr;
    h_slotpos = slotpos;
    h_wdiag = wdiag;
    h_poscol.assign(nnzh, 0);
    for (int c = 0; c < nv; ++c)
        for (int q = hcolptr[c]; q < hcolptr[c + 1]; ++q) h_poscol[q] = c;
    have_pairs = false;
}

void WlsContext::set_mean(const double* z) {
    if (!m) throw std::logic_error("wls_setup has not been called");
    if (!z) throw std::invalid_argument("wls_set_mean: null input");
    d_z.upload(z, m, stream);
    JGB_CUDA(cudaStreamSynchronize(stream));
    have_mean = true;
}

void WlsContext::set_state(const double* vm, const double* va) {
    if (!m) throw std::logic_error("wls_setup has not been called");
    if (!vm || !va) throw std::invalid_argument("wls_set_state: null input");
    d_vm.upload(vm, n, stream);
    d_va.upload(va, n, stream);
    JGB_CUDA(cudaStreamSynchronize(stream));
    have_state = true;
}

void WlsContext::get_state(double* vm, double* va) {
    if (!have_state) throw std::logic_error("no state on the device");
    d_vm.download(vm, n, stream);
    d_va.download(va, n, stream);
    JGB_CUDA(cudaStreamSynchronize(stream));
}

WlsDev WlsContext::view(bool batch) {
    WlsDev d{};
    d.n = n; d.m = m; d.slack = slack; d.nnzh = nnzh; d.nnzg = nnzg; d.nv = 2 * n;
    d.ycolptr = d_ycolptr.p; d.yrow = d_yrow.p; d.y = d_y.p; d.yt = d_yt.p; d.ydiag = d_ydiag.p;
    d.br_from = d_br_from.p; d.br_to = d_br_to.p; d.br_g = d_br_g.p; d.br_b = d_br_b.p; d.br_gsi = d_br_gsi.p;
    d.br_bsi = d_br_bsi.p; d.br_tinv = d_br_tinv.p; d.br_phi = d_br_phi.p;
    d.type = d_type.p; d.index = d_index.p; d.slotptr = d_slotptr.p; d.slotpos = d_slotpos.p;
    d.wdiag = d_wdiag.p; d.woff = d_woff.p; d.hcolptr = d_hcolptr.p; d.hrow = d_hrow.p;
    d.gentry_ptr = d_gentry_ptr.p; d.gterm_a = d_gterm_a.p; d.gterm_b = d_gterm_b.p; d.gterm_w = d_gterm_w.p;
    d.glow_pos = d_glow_pos.p; d.gup_pos = d_gup_pos.p; d.nlow = nlow; d.gslack_pos = gslack_pos;
    d.remaining = d_remaining.p;
    if (!batch) {
        d.vm = d_vm.p; d.va = d_va.p; d.z = d_z.p; d.res = d_res.p; d.hval = d_hval.p; d.gval = d_gval.p;
        d.rhs = d_rhs.p; d.inc = d_inc.p; d.objpart = d_objpart.p; d.obj = d_obj.p; d.maxbits = d_maxbits.p;
        d.maxinc = d_maxinc.p; d.active = d_active.p; d.status = d_status.p; d.iters = d_iters.p;
    } else {
        d.vm = b_vm.p; d.va = b_va.p; d.z = b_z.p; d.res = b_res.p; d.hval = b_hval.p; d.gval = b_gval.p;
        d.rhs = b_rhs.p; d.inc = b_inc.p; d.objpart = b_objpart.p; d.obj = b_obj.p; d.maxbits = b_maxbits.p;
        d.maxinc = b_maxinc.p; d.active = b_active.p; d.status = b_status.p; d.iters = b_iters.p;
    }
    return d;
}

// normalEquation! + objective: rows kernel, then the two-stage objective reduction
void WlsContext::launch_rows(int S, bool batch) {
    WlsDev d = view(batch);
    timer.mark(stream);
    const size_t t0 = timer.last();
    const long long work = (long long)m * S;
    wls_rows_kernel<<<(int)((work + kRowBlock - 1) / kRowBlock), kRowBlock, 0, stream>>>(d, S);
    if (S == 1) {
        wls_objective_kernel<<<dim3(nrowblocks, 1), dim3(1, 128), 0, stream>>>(d, 1, 1024);
    } else {
        wls_objective_kernel<<<dim3(nrowblocks, S / 32), dim3(32, 4), 0, stream>>>(d, S, 1024);
    }
    wls_objective_final_kernel<<<ceil_div(S, 128), 128, 0, stream>>>(d, S, nrowblocks);
    launches += 3;
    timer.mark(stream);
    timer.span(kPhRows, t0, timer.last());
    JGB_CUDA(cudaGetLastError());
}

// gain, rhs, factor + solve, max |increment|
void WlsContext::launch_gain(int S, bool batch) {
    WlsDev d = view(batch);
    const long long gw = (long long)nlow * S, rw = (long long)2 * n * S;
    timer.mark(stream);
    const size_t g0 = timer.last();
    wls_gain_kernel<<<(int)((gw + 127) / 128), 128, 0, stream>>>(d, S);
    wls_rhs_kernel<<<(int)((rw + 127) / 128), 128, 0, stream>>>(d, S);
    launches += 2;
    timer.mark(stream);
    const size_t g1 = timer.last();
    cudaEvent_t mid = timer.reserve();
    const size_t g2 = timer.last();
    MfSolver& sv = batch ? solver_batch : solver;
    sv.factor_solve(d.gval, d.rhs, d.inc, S, batch ? d.active : nullptr, d.status, stream, mid);
    launches += sv.launches_per_solve(S);
    timer.mark(stream);
    const size_t g3 = timer.last();
    timer.span(kPhGain, g0, g1);
    timer.span(kPhFactor, g1, g2);
    timer.span(kPhBacksolve, g2, g3);
    if (S == 1) wls_maxinc_kernel<<<dim3(ceil_div(2 * n, 512), 1), dim3(1, 128), 0, stream>>>(d, 1, 512);
    else wls_maxinc_kernel<<<dim3(ceil_div(2 * n, 128), S / 32), dim3(32, 4), 0, stream>>>(d, S, 128);
    ++launches;
    JGB_CUDA(cudaGetLastError());
}

void WlsContext::increment(double* max_inc, double* objective) {
    if (!have_mean || !have_state) throw std::logic_error("set_mean / set_state must precede increment");
    WlsDev d = view(false);
    JGB_CUDA(cudaMemsetAsync(d_active.p, 1, 1, stream));
    JGB_CUDA(cudaMemsetAsync(d_status.p, 0, sizeof(int), stream));
    JGB_CUDA(cudaMemsetAsync(d_maxbits.p, 0, sizeof(unsigned long long), stream));
    launch_rows(1, false);
    launch_gain(1, false);
    wls_check_kernel<<<1, 32, 0, stream>>>(d, 1, 1, -1.0, 0);
    ++launches;
    d_maxinc.download(h_d.p, 1, stream);
    d_obj.download(h_d.p + 1, 1, stream);
    JGB_CUDA(cudaMemcpyAsync(h_i.p, d_status.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
    JGB_CUDA(cudaStreamSynchronize(stream));
    timer.resolve();
    if (h_i.p[0] == -3) throw std::domain_error("singular gain matrix: zero or non-finite pivot");
    if (max_inc) *max_inc = h_d.p[0];
    if (objective) *objective = h_d.p[1];
}

void WlsContext::solve() {
    if (!have_mean || !have_state) throw std::logic_error("set_mean / set_state must precede solve");
    WlsDev d = view(false);
    JGB_CUDA(cudaMemsetAsync(d_active.p, 1, 1, stream));
    wls_update_kernel<<<ceil_div(n, 128), 128, 0, stream>>>(d, 1);
    ++launches;
    JGB_CUDA(cudaStreamSynchronize(stream));
    iteration += 1;
}

void WlsContext::get_vectors(double* res, double* inc, double* hval, double* gval, int64_t* it) {
    if (!m) throw std::logic_error("wls_setup has not been called");
    if (res) d_res.download(res, m, stream);
    if (inc) d_inc.download(inc, 2 * n, stream);
    if (hval) d_hval.download(hval, nnzh, stream);
    if (gval) d_gval.download(gval, nnzg, stream);
    JGB_CUDA(cudaStreamSynchronize(stream));
    if (it) *it = iteration;
}

int WlsContext::run(int64_t max_iter, double tol, int64_t* iters, double* max_inc, double* objective) {
    if (!have_mean || !have_state) throw std::logic_error("set_mean / set_state must precede run");
    WlsDev d = view(false);
    JGB_CUDA(cudaMemsetAsync(d_active.p, 1, 1, stream));
    JGB_CUDA(cudaMemsetAsync(d_iters.p, 0, sizeof(int), stream));
    JGB_CUDA(cudaMemsetAsync(d_maxbits.p, 0, sizeof(unsigned long long), stream));
    int one = 1;
    JGB_CUDA(cudaMemcpyAsync(d_status.p, &one, sizeof(int), cudaMemcpyHostToDevice, stream));
    iteration = 0;
    int rc = 1;
    for (int64_t it = 0; it <= max_iter; ++it) {
        JGB_CUDA(cudaMemsetAsync(d_remaining.p, 0, sizeof(int), stream));
        launch_rows(1, false);
        launch_gain(1, false);
        wls_check_kernel<<<1, 32, 0, stream>>>(d, 1, 1, tol, (int)max_iter);
        ++launches;
        JGB_CUDA(cudaMemcpyAsync(h_i.p, d_remaining.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
        JGB_CUDA(cudaMemcpyAsync(h_i.p + 1, d_status.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
        d_maxinc.download(h_d.p, 1, stream);
        d_obj.download(h_d.p + 1, 1, stream);
        JGB_CUDA(cudaStreamSynchronize(stream));
        timer.resolve();
        if (h_i.p[0] == 0) { rc = h_i.p[1]; break; }
        wls_update_kernel<<<ceil_div(n, 128), 128, 0, stream>>>(d, 1);
        ++launches;
        iteration += 1;
    }
    if (iters) *iters = iteration;
    if (max_inc) *max_inc = h_d.p[0];
    if (objective) *objective = h_d.p[1];
    if (rc == -3) throw std::domain_error("singular gain matrix: zero or non-finite pivot");
    return rc;
}

void WlsContext::alloc_batch(int S) {
    if (S <= batch_S) return;
    const size_t s = S;
    b_vm.alloc(n * s); b_va.alloc(n * s); b_z.alloc(m * s); b_res.alloc(m * s); b_hval.alloc(nnzh * s);
    b_gval.alloc(nnzg * s); b_rhs.alloc(2 * n * s); b_inc.alloc(2 * n * s); b_objpart.alloc(nrowblocks * s);
    b_obj.alloc(s); b_maxinc.alloc(s); b_maxbits.alloc(s); b_active.alloc(s); b_status.alloc(s); b_iters.alloc(s);
    batch_S = S;
}

int WlsContext::batch(int64_t Sreal64, const double* Z, bool dev_in, int64_t max_iter, double tol, double* vm_out,
                      double* va_out, int32_t* iters_out, int8_t* status_out, double* obj_out, bool dev_out,
                      int64_t* total) {
    if (!have_state) throw std::logic_error("set_state must precede batch");
    if (Sreal64 <= 0 || !Z) throw std::invalid_argument("wls_batch: null or empty input");
    const int Sreal = (int)Sreal64;
    const int S = ceil_div(Sreal, 32) * 32;
    if (S != batch_S) batch_S = 0;
    alloc_batch(S);
    const double* dZ = Z;
    if (!dev_in) {
        b_zraw.upload(Z, (size_t)Sreal * m, stream);
        dZ = b_zraw.p;
    }
    dim3 tb(32, 8);
    wls_transpose_in_kernel<<<dim3(ceil_div(m, 32), S / 32), tb, 0, stream>>>(dZ, b_z.p, m, S, Sreal);
    wls_init_batch_kernel<<<ceil_div(S, 128), 128, 0, stream>>>(b_active.p, b_status.p, b_iters.p, b_maxbits.p, S, Sreal);
    const long long ns = (long long)n * S, hs = (long long)nnzh * S;
    wls_broadcast_kernel<<<(int)((ns + 255) / 256), 256, 0, stream>>>(d_vm.p, b_vm.p, n, S);
    wls_broadcast_kernel<<<(int)((ns + 255) / 256), 256, 0, stream>>>(d_va.p, b_va.p, n, S);
    // constant H entries (codes 1, 12, 13) and zeros for out-of-service rows
    b_hconst.upload(h_const, stream);      // persistent buffer: a cudaMalloc / cudaFree pair per batch costs 50-250 ms once
                                           // tens of GB are allocated (seen as WLS 448 vs 241 ms per step beside the NR context)
    wls_broadcast_kernel<<<(int)((hs + 255) / 256), 256, 0, stream>>>(b_hconst.p, b_hval.p, nnzh, S);
    JGB_CUDA(cudaMemsetAsync(b_res.p, 0, (size_t)m * S * sizeof(double), stream));
    launches += 5;
    WlsDev d = view(true);
    for (int64_t it = 0; it <= max_iter; ++it) {
        JGB_CUDA(cudaMemsetAsync(d_remaining.p, 0, sizeof(int), stream));
        launch_rows(S, true);
        launch_gain(S, true);
        wls_check_kernel<<<ceil_div(S, 128), 128, 0, stream>>>(d, S, Sreal, tol, (int)max_iter);
        ++launches;
        JGB_CUDA(cudaMemcpyAsync(h_i.p, d_remaining.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
        JGB_CUDA(cudaStreamSynchronize(stream));
        timer.resolve();
        if (h_i.p[0] == 0) break;
        wls_update_kernel<<<(int)((ns + 127) / 128), 128, 0, stream>>>(d, S);
        ++launches;
    }
    double* dvm = vm_out;
    double* dva = va_out;
    if (!dev_out) {
        b_out.alloc(2 * (size_t)Sreal * n);
        dvm = b_out.p;
        dva = b_out.p + (size_t)Sreal * n;
    }
    wls_transpose_out_kernel<<<dim3(ceil_div(n, 32), S / 32), tb, 0, stream>>>(b_vm.p, dvm, n, S, Sreal);
    wls_transpose_out_kernel<<<dim3(ceil_div(n, 32), S / 32), tb, 0, stream>>>(b_va.p, dva, n, S, Sreal);
    launches += 2;
    std::vector<int> hit(Sreal), hst(Sreal);
    std::vector<double> hobj(Sreal);
    if (dev_out) {
        wls_copy_results_kernel<<<ceil_div(Sreal, 128), 128, 0, stream>>>(b_iters.p, b_status.p, b_obj.p, iters_out,
                                                                          status_out, obj_out, Sreal);
        ++launches;
    } else {
        JGB_CUDA(cudaMemcpyAsync(vm_out, dvm, (size_t)Sreal * n * sizeof(double), cudaMemcpyDeviceToHost, stream));
        JGB_CUDA(cudaMemcpyAsync(va_out, dva, (size_t)Sreal * n * sizeof(double), cudaMemcpyDeviceToHost, stream));
    }
    b_iters.download(hit.data(), Sreal, stream);
    b_status.download(hst.data(), Sreal, stream);
    b_obj.download(hobj.data(), Sreal, stream);
    JGB_CUDA(cudaStreamSynchronize(stream));
    int64_t tot = 0;
    int worst = 0;
    for (int s = 0; s < Sreal; ++s) {
        tot += hit[s];
        if (!dev_out) {
            if (iters_out) iters_out[s] = hit[s];
            if (status_out) status_out[s] = (int8_t)hst[s];
            if (obj_out) obj_out[s] = hobj[s];
        }
        if (hst[s] != 0) worst = 1;
    }
    if (total) *total = tot;
    return worst;
}

double WlsContext::stat(const std::string& key) {
    if (key.rfind("wls.batch.", 0) == 0) {
        const Symbolic& b = solver_batch.sym;
        if (key == "wls.batch.u_size") return (double)b.u_size;
        if (key == "wls.batch.upd_size") return (double)b.upd_size;
        if (key == "wls.batch.nnz_lu") return (double)b.nnz_lu;
        if (key == "wls.batch.fronts") return b.nfronts;
        if (key == "wls.batch.levels") return b.nlevels;
        if (key == "wls.batch.flops") return b.flops;
        if (key == "wls.batch.max_front") return b.max_front;
        if (key == "wls.batch.launches_per_solve") return solver_batch.launches_per_solve(32);
        return -1.0;
    }
    const Symbolic& s = solver.sym;
    if (key == "wls.time.rows_count") return (double)timer.count[kPhRows];
    if (key == "wls.n") return n;
    if (key == "wls.nbr") return nbr;
    if (key == "wls.time.rows_ms") return timer.ms[kPhRows];
    if (key == "wls.time.gain_ms") return timer.ms[kPhGain];
    if (key == "wls.time.factor_ms") return timer.ms[kPhFactor];
    if (key == "wls.time.backsolve_ms") return timer.ms[kPhBacksolve];
    if (key == "wls.time.factor_count") return (double)timer.count[kPhFactor];
    if (key == "wls.u_size") return (double)s.u_size;
    if (key == "wls.upd_size") return (double)s.upd_size;
    if (key == "wls.nnz_lu") return (double)s.nnz_lu;
    if (key == "wls.fronts") return s.nfronts;
    if (key == "wls.levels") return s.nlevels;
    if (key == "wls.flops") return s.flops;
    if (key == "wls.max_front") return s.max_front;
    if (key == "wls.m") return m;
    if (key == "wls.nnz_h") return nnzh;
    if (key == "wls.nnz_g") return nnzg;
    if (key == "wls.gain_terms") return (double)nterms;
    if (key == "wls.launches_per_iteration") return solver.launches_per_solve(1) + 8;
    // per scenario-iteration algorithmic bytes
    if (key == "wls.rows_bytes") return 21.0 * m + 8.0 * m + 8.0 * nnzh + 16.0 * n + 56.0 * nbr;
    if (key == "wls.gain_bytes") return 12.0 * nnzh + 8.0 * m + 8.0 * nnzg + 16.0 * n;
    if (key == "wls.solve_bytes") return (double)solver.factor_bytes(1);
    return -1.0;
}

void WlsContext::build_pairs() {
    const Symbolic& sy = solver.sym;
    const int nv = 2 * n;
    // pivot owner of every variable and, per front, its rows sorted by variable id for local-index lookups
    std::vector<int> owner(nv, -1);
    std::vector<std::vector<std::pair<int, int>>> local(sy.nfronts);
    for (int f = 0; f < sy.nfronts; ++f) {
        const int* rows = sy.f_rows.data() + sy.f_rowptr[f];
        for (int q = 0; q < sy.f_k[f]; ++q) owner[rows[q]] = f;
        local[f].reserve(sy.f_nf[f]);
        for (int q = 0; q < sy.f_nf[f]; ++q) local[f].push_back({rows[q], q});
        std::sort(local[f].begin(), local[f].end());
    }
    auto find_local = [&](int f, int var) -> int {
        auto it = std::lower_bound(local[f].begin(), local[f].end(), std::make_pair(var, -1));
        if (it == local[f].end() || it->first != var)
            throw std::runtime_error("bad-data lists: a gain entry is missing from the factor pattern (-4)");
        return it->second;
    };
    solver.selected_inverse(stream);          // allocates the blocks and fills zoff_host
    std::vector<int> pptr(m + 1, 0), ppa, ppb;
    std::vector<long long> pz;
    for (int r = 0; r < m; ++r) {
        for (int x = h_slotptr[r]; x < h_slotptr[r + 1]; ++x) {
            const int pa = h_slotpos[x], a = h_poscol[pa];
            if (a == slack) continue;
            for (int y = x; y < h_slotptr[r + 1]; ++y) {
                const int pb = h_slotpos[y], b = h_poscol[pb];
                if (b == slack) continue;
                const int first = sy.iperm[a] < sy.iperm[b] ? a : b;
                const int f = owner[first];
                const int nf = sy.f_nf[f];
                ppa.push_back(pa);
                ppb.push_back(pb);
                pz.push_back(solver.zoff_host[f] + find_local(f, a) + (long long)find_local(f, b) * nf);
            }
        }
        pptr[r + 1] = (int)ppa.size();
    }
    d_pair_ptr.upload(pptr, stream);
    d_pair_pa.upload(ppa, stream);
    d_pair_pb.upload(ppb, stream);
    d_pair_z.upload(pz, stream);
    d_proj.alloc(m);
    JGB_CUDA(cudaStreamSynchronize(stream));
    have_pairs = true;
}

void WlsContext::residual_test(double threshold, double* max_rn, int64_t* index, double* c_out) {
    if (!have_mean || !have_state) throw std::logic_error("residual_test needs a solved estimation on the device");
    (void)threshold;
    if (!have_pairs) build_pairs();
    // H, the residual and the gain factor on the device are those of the last increment! (final state)
    const double* Z = solver.selected_inverse(stream);
    wls_projection_kernel<<<ceil_div(m, 128), 128, 0, stream>>>(m, d_pair_ptr.p, d_pair_pa.p, d_pair_pb.p, d_pair_z.p,
                                                               d_hval.p, Z, d_proj.p);
    launches += 1 + solver.sym.ndepths;
    JGB_CUDA(cudaGetLastError());
    std::vector<double> c(m), res(m);
    d_proj.download(c.data(), m, stream);
    d_res.download(res.data(), m, stream);
    JGB_CUDA(cudaStreamSynchronize(stream));
    double best = 0.0;
    int64_t where = 0;
    for (int i = 0; i < m; ++i) {               // badData.jl:207-215: strict >, ascending rows
        if (res[i] != 0.0) {
            const double rn = std::fabs(res[i]) / std::sqrt(std::fabs(1.0 / h_wdiag[i] - c[i]));
            if (rn > best) { best = rn; where = i + 1; }
        }
    }
    if (max_rn) *max_rn = best;
    if (index) *index = where;
    if (c_out) std::copy(c.begin(), c.end(), c_out);
}

// Slots of one measurement row in the semantic order the rows kernel expects (the same rules as in setup)
static void row_slots(int code, int k, int row, int n, const std::vector<std::vector<std::pair<int, int>>>& rowent,
                      const std::vector<int>& ycolptr, const std::vector<int>& yrow, const std::vector<int>& brf,
                      const std::vector<int>& brt, std::vector<int>& out) {
    auto hpos = [&](int col) -> int {
        for (auto& e : rowent[row])
            if (e.first == col) return e.second;
        throw std::runtime_error("wls_update_rows: the H pattern does not hold the entry this measurement type needs");
    };
    out.clear();
    if (code == 1 || code == 12) out.push_back(hpos(k + n));
    else if (code == 13) out.push_back(hpos(k));
    else if (code == 6 || code == 9) {
        for (int q = ycolptr[k]; q < ycolptr[k + 1]; ++q) { out.push_back(hpos(yrow[q])); out.push_back(hpos(yrow[q] + n)); }
    } else if (code == 16 || code == 17) { out.push_back(hpos(k)); out.push_back(hpos(k + n)); }
    else if (code != 0) {
        const int i = brf[k], j = brt[k];
        out.push_back(hpos(i)); out.push_back(hpos(i + n)); out.push_back(hpos(j)); out.push_back(hpos(j + n));
    }
}

// update*!(analysis; ...) on a Gauss-Newton analysis (e.g. _updateWattmeter!, src/measurement/powermeter.jl:640-677 and
// its siblings for the other meters): mean = status * mean, residual = 0, the row's Jacobian entries = 0, type =
// status * code, precision[idx, idx] = 1 / variance. The gain pattern and the symbolic factorisation are reused.
void WlsContext::update_rows(int64_t k, const int64_t* rows, const double* mean, const double* precision,
                             const double* precision_off, const int8_t* type, const int64_t* index) {
    if (!m) throw std::logic_error("wls_setup has not been called");
    if (k <= 0 || !rows) throw std::invalid_argument("wls_update_rows: null or empty input");
    bool slots_changed = false;
    std::vector<int> slots;
    for (int64_t e = 0; e < k; ++e) {
        if (rows[e] < 1 || rows[e] > m) throw std::invalid_argument("wls_update_rows: row out of range");
        const int r = (int)rows[e] - 1;
        const int code = type ? type[e] : h_type[r];
        const int idx = index ? (int)index[e] - 1 : h_index[r];
        if (code < 0 || code > 21) throw std::invalid_argument("wls_update_rows: unknown measurement code");
        const bool bus_row = code == 1 || code == 6 || code == 9 || code == 12 || code == 13 || code == 16 || code == 17;
        if (code != 0 && (idx < 0 || idx >= (bus_row ? n : nbr)))
            throw std::invalid_argument("wls_update_rows: measurement index out of range");
        if (precision_off && precision_off[e] != 0.0 && h_woff[r] == 0.0)
            throw std::runtime_error("wls_update_rows: a new off-diagonal precision entry changes the gain pattern (-4)");
        if (precision && !(precision[e] == precision[e])) throw std::invalid_argument("wls_update_rows: NaN precision");
        // the row's slot list follows its type: rebuild it when the type or location changes
        row_slots(code, idx, r, n, h_rowent, h_ycolptr, h_yrow, h_brfrom, h_brto, slots);
        std::vector<int> old(h_slotpos.begin() + h_slotptr[r], h_slotpos.begin() + h_slotptr[r + 1]);
        // Jacobian entries of the row: zeroed (constant entries of codes 1 / 12 / 13 are set again below)
        for (auto& ent : h_rowent[r]) h_const[ent.second] = 0.0;
        if (slots != old) {
            slots_changed = true;
            pending_slots[r] = slots;
        }
        if (code == 1 || code == 12 || code == 13) h_const[slots[0]] = 1.0;
        h_type[r] = (int8_t)code;
        h_index[r] = idx;
        if (precision) h_wdiag[r] = precision[e];
        if (precision_off) h_woff[r] = precision_off[e];
    }
    if (slots_changed) {     // rare: a row changed its type class or came back into service — rebuild the flat lists
        std::vector<int> sp(m + 1, 0), pos;
        pos.reserve(h_slotpos.size() + 16);
        for (int r = 0; r < m; ++r) {
            auto it = pending_slots.find(r);
            if (it != pending_slots.end()) pos.insert(pos.end(), it->second.begin(), it->second.end());
            else pos.insert(pos.end(), h_slotpos.begin() + h_slotptr[r], h_slotpos.begin() + h_slotptr[r + 1]);
            sp[r + 1] = (int)pos.size();
        }
        pending_slots.clear();
        h_slotptr.swap(sp);
        h_slotpos.swap(pos);
        d_slotptr.upload(h_slotptr, stream);
        d_slotpos.upload(h_slotpos, stream);
        have_pairs = false;
    }
    // per-row scalars and the row's Jacobian entries on the device
    for (int64_t e = 0; e < k; ++e) {
        const int r = (int)rows[e] - 1;
        const signed char tc = (signed char)h_type[r];
        JGB_CUDA(cudaMemcpyAsync(d_type.p + r, &h_type[r], 1, cudaMemcpyHostToDevice, stream));
        JGB_CUDA(cudaMemcpyAsync(d_index.p + r, &h_index[r], sizeof(int), cudaMemcpyHostToDevice, stream));
        JGB_CUDA(cudaMemcpyAsync(d_wdiag.p + r, &h_wdiag[r], sizeof(double), cudaMemcpyHostToDevice, stream));
        JGB_CUDA(cudaMemcpyAsync(d_woff.p + r, &h_woff[r], sizeof(double), cudaMemcpyHostToDevice, stream));
        if (mean) JGB_CUDA(cudaMemcpyAsync(d_z.p + r, &mean[e], sizeof(double), cudaMemcpyHostToDevice, stream));
        JGB_CUDA(cudaMemsetAsync(d_res.p + r, 0, sizeof(double), stream));
        for (auto& ent : h_rowent[r])
            JGB_CUDA(cudaMemcpyAsync(d_hval.p + ent.second, &h_const[ent.second], sizeof(double), cudaMemcpyHostToDevice,
                                     stream));
        (void)tc;
    }
    JGB_CUDA(cudaStreamSynchronize(stream));
    have_pairs = have_pairs && !precision;     // the projection lists cache 1 / W_ii
}

// acNodalUpdate! on the model a WLS analysis reads (powerSystem/model.jl:81-110): value-only Ybus update
void WlsContext::update_y(int64_t k, const int64_t* pos, const double* y, const double* yt) {
    if (!m) throw std::logic_error("wls_setup has not been called");
    if (k <= 0 || !pos || !y || !yt) throw std::invalid_argument("wls_update_y: null or empty input");
    for (int64_t e = 0; e < k; ++e) {
        if (pos[e] < 1 || pos[e] > nnzy) throw std::invalid_argument("wls_update_y: position out of range");
        JGB_CUDA(cudaMemcpyAsync(d_y.p + (pos[e] - 1), y + 2 * e, sizeof(double2), cudaMemcpyHostToDevice, stream));
        JGB_CUDA(cudaMemcpyAsync(d_yt.p + (pos[e] - 1), yt + 2 * e, sizeof(double2), cudaMemcpyHostToDevice, stream));
    }
    JGB_CUDA(cudaStreamSynchronize(stream));
}

// acParameterUpdate! (powerSystem/model.jl:113-140) for the flow / current rows of one branch
void WlsContext::update_branch(int64_t branch1, double cond, double susc, double tap, double shift, const double* adm) {
    if (!m) throw std::logic_error("wls_setup has not been called");
    if (branch1 < 1 || branch1 > nbr || !adm) throw std::invalid_argument("wls_update_branch: bad branch index or null input");
    const int b = (int)branch1 - 1;
    const double v[6] = {adm[0], adm[1], 0.5 * cond, 0.5 * susc, 1.0 / tap, shift};
    double* dst[6] = {d_br_g.p + b, d_br_b.p + b, d_br_gsi.p + b, d_br_bsi.p + b, d_br_tinv.p + b, d_br_phi.p + b};
    for (int q = 0; q < 6; ++q) JGB_CUDA(cudaMemcpyAsync(dst[q], &v[q], sizeof(double), cudaMemcpyHostToDevice, stream));
    JGB_CUDA(cudaStreamSynchronize(stream));
}

void WlsContext::remove_row(int64_t row1) {
    if (!m) throw std::logic_error("wls_setup has not been called");
    if (row1 < 1 || row1 > m) throw std::invalid_argument("wls_remove_row: row out of range");
    const int r = (int)row1 - 1;
    wls_remove_row_kernel<<<1, 32, 0, stream>>>(r, d_slotpos.p, h_slotptr[r], h_slotptr[r + 1], d_hval.p, d_res.p,
                                               d_z.p, d_type.p);
    ++launches;
    for (int t = h_slotptr[r]; t < h_slotptr[r + 1]; ++t) h_const[h_slotpos[t]] = 0.0;
    h_type[r] = 0;
    JGB_CUDA(cudaGetLastError());
    JGB_CUDA(cudaStreamSynchronize(stream));
    iteration = 0;
}

}  // namespace jgb
