// Numeric multifrontal kernels. See solver.cuh.
//
// Kernel map: mf_factor_kernel = K2 (front assembly) + K3/K4 (dense partial LU + Schur
// update) + forward substitution (the rhs rides along as column nf of every front, so L is never stored);
// mf_backsolve_* = K5 (backward substitution, top-down over the elimination tree).
#include "solver.cuh"

#include <algorithm>
#include <cmath>

namespace jgb {

namespace {

__device__ __forceinline__ long long urow_off(int p, int nf) {
    return (long long)p * (nf + 1) - (long long)p * (p - 1) / 2;
}

// One CTA = one front x TS scenarios. Thread t: scenario lane sl = t % TS, entry lane e = t / TS,
// entry lanes are arranged TR (rows) x TC (columns). Front F is column major, ld = nf, nf+1 columns,
// element (r,c) of scenario lane sl at F[(r + c*nf) * TS + sl].
__global__ void __launch_bounds__(256)
mf_factor_kernel(DevSym sy, const int* __restrict__ fronts, const double* __restrict__ aval,
                 const double* __restrict__ rhs, double* __restrict__ U, double* __restrict__ upd, int S, int TS,
                 int TR, const unsigned char* __restrict__ active, int* __restrict__ status, double* gwork,
                 long long gstride) {
    extern __shared__ double Fs[];
    // fronts too large for shared memory live in a per-CTA global (L2-resident) workspace
    double* F = gwork ? gwork + ((long long)blockIdx.y * gridDim.x + blockIdx.x) * gstride : Fs;
    const int f = fronts[blockIdx.x];
    const int sl = threadIdx.x % TS;
    const int e0 = threadIdx.x / TS;
    const int TE = blockDim.x / TS;
    const int er = e0 % TR, ec = e0 / TR, TC = TE / TR;
    const int s = blockIdx.y * TS + sl;
    const bool act = active ? (active[s] != 0) : true;
    if (!__syncthreads_or(act)) return;

    const int nf = sy.f_nf[f], k = sy.f_k[f], u = nf - k;
    const int* __restrict__ rows = sy.f_rows + sy.f_rowptr[f];
    const int total = nf * (nf + 1);

    for (int pos = e0; pos < total; pos += TE) F[pos * TS + sl] = 0.0;
    __syncthreads();
    if (act) {
        const int a1 = sy.f_asmptr[f + 1];
        for (int a = sy.f_asmptr[f] + e0; a < a1; a += TE)
            F[sy.asm_dst[a] * TS + sl] = aval[(long long)sy.asm_src[a] * S + s];
        for (int p = e0; p < k; p += TE) F[(p + nf * nf) * TS + sl] = rhs[(long long)rows[p] * S + s];
    }
    __syncthreads();
    for (int ci = sy.f_childptr[f]; ci < sy.f_childptr[f + 1]; ++ci) {
        const int c = sy.f_children[ci];
        const int uc = sy.f_nf[c] - sy.f_k[c];
        const int* __restrict__ rel = sy.f_rel + sy.f_relptr[c];
        const double* __restrict__ C = upd + sy.f_updoff[c] * S + s;
        if (act) {
            for (int j = ec; j <= uc; j += TC) {
                const int dc = (j < uc) ? rel[j] : nf;
                for (int i = er; i < uc; i += TR)
                    F[(rel[i] + dc * nf) * TS + sl] += C[(long long)(i + j * uc) * S];
            }
        }
        __syncthreads();
    }
    bool bad = false;
    for (int p = 0; p < k; ++p) {
        const double piv = F[(p + p * nf) * TS + sl];
        if (piv == 0.0 || !isfinite(piv)) bad = true;
        const double inv = 1.0 / piv;
        for (int j = p + 1 + ec; j <= nf; j += TC) {
            const double m = inv * F[(p + j * nf) * TS + sl];
            for (int i = p + 1 + er; i < nf; i += TR)
                F[(i + j * nf) * TS + sl] -= F[(i + p * nf) * TS + sl] * m;
        }
        __syncthreads();
    }
    if (!act) return;
    if (bad && e0 == 0) status[s] = -3;
    double* __restrict__ Uf = U + sy.f_uoff[f] * S + s;
    for (int p = ec; p < k; p += TC) {
        double* Urow = Uf + urow_off(p, nf) * S;
        for (int j = p + er; j <= nf; j += TR) {
            const double v = F[(p + j * nf) * TS + sl];
            Urow[(long long)(j - p) * S] = (j == p) ? 1.0 / v : v;
        }
    }
    double* __restrict__ Cf = upd + sy.f_updoff[f] * S + s;
    for (int j = ec; j <= u; j += TC)
        for (int i = er; i < u; i += TR)
            Cf[(long long)(i + j * u) * S] = F[((k + i) + (k + j) * nf) * TS + sl];
}

// Backward substitution, S == 1: one CTA per front, pivots processed in blocks of 32 rows from the bottom up.
// For each block the packed U rows are staged in shared memory, the part of every row that multiplies already
// known x (later pivots and update rows) is removed by one warp per row, and warp 0 finishes the 32 x 32 triangle.
constexpr int kBsRows = 32;
__global__ void __launch_bounds__(128)
mf_backsolve_single(DevSym sy, const int* __restrict__ fronts, const double* __restrict__ U,
                    double* __restrict__ x, const unsigned char* __restrict__ active) {
    extern __shared__ double sh[];
    if (active && !active[0]) return;
    const int f = fronts[blockIdx.x];
    const int nf = sy.f_nf[f], k = sy.f_k[f];
    const int* __restrict__ rows = sy.f_rows + sy.f_rowptr[f];
    double* xs = sh;              // nf entries: x of the front rows (pivots filled in as they are solved)
    double* Us = sh + nf;         // packed rows of the current block
    const double* __restrict__ Uf = U + sy.f_uoff[f];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int j = k + threadIdx.x; j < nf; j += blockDim.x) xs[j] = x[rows[j]];
    for (int p1 = k; p1 > 0; p1 -= kBsRows) {
        const int p0 = max(0, p1 - kBsRows);
        const long long base = urow_off(p0, nf);
        const int cnt = (int)(urow_off(p1, nf) - base);
        for (int e = threadIdx.x; e < cnt; e += blockDim.x) Us[e] = Uf[base + e];
        __syncthreads();
        // rows p0..p1-1: t_p = y_p - sum_{j >= p1} U[p,j] x_j
        for (int p = p0 + warp; p < p1; p += nwarps) {
            const double* Urow = Us + (urow_off(p, nf) - base);
            double acc = 0.0;
            for (int j = p1 + lane; j < nf; j += 32) acc += Urow[j - p] * xs[j];
            for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) xs[p] = Urow[nf - p] - acc;
        }
        __syncthreads();
        if (warp == 0) {
            for (int p = p1 - 1; p >= p0; --p) {
                const int owner = (p - p0) & 31;
                double xp = 0.0;
                if (lane == owner) {
                    xp = xs[p] * Us[urow_off(p, nf) - base];
                    xs[p] = xp;
                }
                xp = __shfl_sync(0xffffffffu, xp, owner);
                const int q = p0 + lane;
                if (q < p) xs[q] -= Us[(urow_off(q, nf) - base) + (p - q)] * xp;
                __syncwarp();
            }
        }
        __syncthreads();
    }
    for (int p = threadIdx.x; p < k; p += blockDim.x) x[rows[p]] = xs[p];
}

// Backward substitution, batch: one thread per (front, scenario); lanes of a warp are consecutive scenarios, so every
// U / x access is a coalesced 256-byte line.
__global__ void __launch_bounds__(128)
mf_backsolve_batch(DevSym sy, const int* __restrict__ fronts, int nfr, const double* __restrict__ U,
                   double* __restrict__ x, int S, const unsigned char* __restrict__ active) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int fi = (int)(gid / S);
    const int s = (int)(gid % S);
    if (fi >= nfr) return;
    if (active && !active[s]) return;
    const int f = fronts[fi];
    const int nf = sy.f_nf[f], k = sy.f_k[f];
    const int* __restrict__ rows = sy.f_rows + sy.f_rowptr[f];
    const double* __restrict__ Uf = U + sy.f_uoff[f] * S + s;
    for (int p = k - 1; p >= 0; --p) {
        const double* Urow = Uf + urow_off(p, nf) * S;
        double acc = Urow[(long long)(nf - p) * S];
        for (int j = p + 1; j < nf; ++j) acc -= Urow[(long long)(j - p) * S] * x[(long long)rows[j] * S + s];
        x[(long long)rows[p] * S + s] = acc * Urow[0];
    }
}

constexpr int kMaxSmemFront = 150;    // nf*(nf+1)*8 bytes must fit the 200 KB dynamic shared-memory budget

int pow2_floor(int v) {
    int p = 1;
    while (p * 2 <= v) p *= 2;
    return p;
}

}  // namespace

void MfSolver::setup(const Symbolic& s, cudaStream_t st) {
    sym = s;
    d_f_k.upload(sym.f_k, st);
    d_f_nf.upload(sym.f_nf, st);
    d_f_rowptr.upload(sym.f_rowptr, st);
    d_f_rows.upload(sym.f_rows, st);
    d_f_relptr.upload(sym.f_relptr, st);
    d_f_rel.upload(sym.f_rel, st);
    d_f_childptr.upload(sym.f_childptr, st);
    d_f_children.upload(sym.f_children, st);
    d_f_asmptr.upload(sym.f_asmptr, st);
    d_asm_src.upload(sym.asm_src, st);
    d_asm_dst.upload(sym.asm_dst, st);
    d_level_fronts.upload(sym.level_fronts, st);
    d_depth_fronts.upload(sym.depth_fronts, st);
    std::vector<long long> uo(sym.f_uoff.begin(), sym.f_uoff.end()), po(sym.f_updoff.begin(), sym.f_updoff.end());
    d_f_uoff.upload(uo, st);
    d_f_updoff.upload(po, st);
    JGB_CUDA(cudaStreamSynchronize(st));   // the host vectors above go out of scope
    dev.f_k = d_f_k.p; dev.f_nf = d_f_nf.p; dev.f_rowptr = d_f_rowptr.p; dev.f_rows = d_f_rows.p;
    dev.f_relptr = d_f_relptr.p; dev.f_rel = d_f_rel.p; dev.f_childptr = d_f_childptr.p;
    dev.f_children = d_f_children.p; dev.f_asmptr = d_f_asmptr.p; dev.asm_src = d_asm_src.p;
    dev.asm_dst = d_asm_dst.p; dev.f_uoff = d_f_uoff.p; dev.f_updoff = d_f_updoff.p;
    planned_S = -1;
    JGB_CUDA(cudaFuncSetAttribute(mf_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    JGB_CUDA(cudaFuncSetAttribute(mf_backsolve_single, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
}

void MfSolver::plan(int S) {
    if (S == planned_S) return;
    fplan.clear();
    splan.clear();
    size_t gwork_need = 0;
    const int cls_bound[] = {6, 16, 48, kMaxSmemFront, 1 << 30};
    auto cls = [&](int nf) { int c = 0; while (nf > cls_bound[c]) ++c; return c; };
    for (int l = 0; l < sym.nlevels; ++l) {
        int b = sym.levelptr[l], e = sym.levelptr[l + 1];
        int i = b;
        while (i < e) {   // fronts are sorted by decreasing order inside a level
            int c = cls(sym.f_nf[sym.level_fronts[i]]);
            int j = i;
            while (j < e && cls(sym.f_nf[sym.level_fronts[j]]) == c) ++j;
            int nf = sym.f_nf[sym.level_fronts[i]];
            size_t per = (size_t)nf * (nf + 1) * sizeof(double);
            FactorLaunch fl{};
            fl.begin = i;
            fl.count = j - i;
            fl.global_front = nf > kMaxSmemFront;
            if (fl.global_front) {
                fl.ts = (S == 1) ? 1 : 4;
                fl.threads = 256;
            } else if (S == 1) {
                fl.ts = 1;
                fl.threads = nf <= 6 ? 32 : nf <= 16 ? 64 : nf <= 48 ? 128 : 256;
            } else {
                int cap = (int)std::max<size_t>(1, (size_t)(160 * 1024) / per);
                fl.ts = std::min(32, pow2_floor(cap));
                int te = pow2_floor(std::max(1, std::min(256 / fl.ts, nf * (nf + 1) / 16)));
                fl.threads = fl.ts * te;
                if (fl.threads < 32) { fl.threads = 32; }
            }
            int te = fl.threads / fl.ts;
            fl.tr = std::min(te, 16);
            fl.smem = fl.global_front ? 0 : per * fl.ts;
            fl.gstride = (long long)nf * (nf + 1) * fl.ts;
            if (fl.global_front)
                gwork_need = std::max<size_t>(gwork_need, (size_t)fl.gstride * fl.count * (S / fl.ts));
            if (fl.smem > 200 * 1024) throw std::runtime_error("front too large for shared memory");
            fplan.push_back(fl);
            i = j;
        }
    }
    for (int d = 0; d < sym.ndepths; ++d) {
        SolveLaunch sl{};
        sl.begin = sym.depthptr[d];
        sl.count = sym.depthptr[d + 1] - sl.begin;
        size_t smem = 0;
        for (int i = sl.begin; i < sl.begin + sl.count; ++i) {
            int f = sym.depth_fronts[i];
            int nf = sym.f_nf[f], k = sym.f_k[f];
            int kb = std::min(k, 32);      // largest staged block: the first (longest) kb rows
            size_t usz = (size_t)kb * (nf + 1) - (size_t)kb * (kb - 1) / 2;
            smem = std::max(smem, (usz + nf) * sizeof(double));
            sl.max_nf = std::max(sl.max_nf, nf);
            sl.max_k = std::max(sl.max_k, k);
        }
        sl.smem = smem;
        if (smem > 100 * 1024) throw std::runtime_error("front too large for the back-solve staging buffer");
        splan.push_back(sl);
    }
    d_U.alloc((size_t)sym.u_size * S);
    d_upd.alloc((size_t)sym.upd_size * S);
    if (gwork_need) d_gwork.alloc(gwork_need);
    planned_S = S;
}

int MfSolver::launches_per_solve(int S) {
    plan(S);
    return (int)(fplan.size() + splan.size());
}

int64_t MfSolver::factor_bytes(int S) const {
    // read A values + rhs, write U rows, write + read update blocks, back-solve reads U and writes x
    int64_t per = 8LL * ((int64_t)sym.asm_src.size() + sym.n + 2 * sym.u_size + 2 * sym.upd_size + sym.n);
    return per * S;
}

void MfSolver::factor_solve(const double* aval, const double* rhs, double* x, int S, const unsigned char* active,
                            int* status, cudaStream_t st, cudaEvent_t after_factor) {
    plan(S);
    for (const FactorLaunch& fl : fplan) {
        dim3 grid(fl.count, S / fl.ts);
        mf_factor_kernel<<<grid, fl.threads, fl.smem, st>>>(dev, d_level_fronts.p + fl.begin, aval, rhs, d_U.p,
                                                            d_upd.p, S, fl.ts, fl.tr, active, status,
                                                            fl.global_front ? d_gwork.p : nullptr, fl.gstride);
    }
    if (after_factor) JGB_CUDA(cudaEventRecord(after_factor, st));
    for (const SolveLaunch& sl : splan) {
        if (S == 1) {
            mf_backsolve_single<<<sl.count, 128, sl.smem, st>>>(dev, d_depth_fronts.p + sl.begin, d_U.p, x, active);
        } else {
            long long work = (long long)sl.count * S;
            int blocks = (int)((work + 127) / 128);
            mf_backsolve_batch<<<blocks, 128, 0, st>>>(dev, d_depth_fronts.p + sl.begin, sl.count, d_U.p, x, S,
                                                       active);
        }
    }
    JGB_CUDA(cudaGetLastError());
}

}  // namespace jgb
