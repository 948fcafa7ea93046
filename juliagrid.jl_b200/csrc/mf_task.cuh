// Task kernel of the batch factorisation: one CTA walks a whole list of small fronts (one or several complete
// subtrees of the elimination tree, in postorder) for a tile of 32 scenarios. Included by solver.cu inside
// namespace jgb { namespace { ... } }.
//
// Why: the fronts of a power-grid Jacobian are tiny (92 % have at most 16 rows) and a (front, tile) CTA moves only
// ~10 KB, so the per-front kernels are bound by launch turnover and by the round trip of the update blocks through
// HBM. Here the update block of a front stays on a stack in shared memory until its parent (a later front of the same
// task) consumes it; only the block of a task root is written to the update storage in HBM.
//
// Work split inside the CTA: lane = scenario (32), warp e0 of TE owns the front columns c with c % TE == e0 — it zeroes
// them, scatters the matrix entries that fall into them (the host sorts every front's entry list by owner warp),
// adds the children's contributions to them and finally holds them in registers for the elimination. A lane only ever
// touches its own scenario's slice of shared memory, so none of these phases needs a barrier; the CTA synchronises
// once per pivot (the pivot column is broadcast through a double-buffered strip) and once per front (update block
// visible to the parent).
//
// Task blob (int32, built by MfSolver::build_tasks, copied to shared memory at task start):
//   [0] number of fronts; then kTaskRec ints per front:
//   0 nf, 1 k, 2 offset of the entry lists, 3 offset of the child records, 4 number of children,
//   5 stack offset of the update block in elements (-1: task root, block goes to HBM), 6/7 packed-U offset lo/hi,
//   8/9 update-storage offset lo/hi, 10 tile width of the section a root's block goes to (its parent's launch width)
//   entry lists: TE + 1 sublist offsets (in pairs, relative), then (source, destination) pairs; source >= 0 indexes the
//   matrix values, source < 0 the right-hand side (-source - 1); destination = r + c * nf
//   child record: uc, stack offset (-1: block in HBM), update-storage offset lo/hi, then uc relative indices
constexpr int kTaskPre = 8;        // matrix entries per thread fetched one front ahead

template <int TE, int M>
__device__ __forceinline__ void task_eliminate(const double* Fl, double* bcl, int nf, int k, int e0, bool act,
                                               double* __restrict__ Uf, int S, double* ub, int ust, bool& bad) {
    constexpr int NC = (M + 1 + TE - 1) / TE;
    double col[NC][M];
#pragma unroll
    for (int q = 0; q < NC; ++q) {
        const int c = e0 + q * TE;
#pragma unroll
        for (int i = 0; i < M; ++i) col[q][i] = (c <= nf && i < nf) ? Fl[(i + c * nf) * 32] : 0.0;
    }
#pragma unroll
    for (int p = 0; p < M; ++p) {
        if (p >= k) break;
        double* b = bcl + (p & 1) * (M * 32);
        if (e0 == p % TE) {
#pragma unroll
            for (int i = p; i < M; ++i)
                if (i < nf) b[i * 32] = col[p / TE][i];
        }
        __syncthreads();
        const double piv = b[p * 32];
        if (piv == 0.0 || !isfinite(piv)) bad = true;
        const double inv = 1.0 / piv;
        double* Urow = Uf + urow_off(p, nf) * 32;      // packed U rows of task fronts: section of tile width 32
        double m[NC];                                   // multiplier of each owned column, 0 for columns not updated
#pragma unroll
        for (int q = 0; q < NC; ++q) {
            const int c = e0 + q * TE;
            const double upc = col[q][p];               // U[p, c]
            const bool in = c >= p && c <= nf;
            if (act && in) Urow[(c - p) * 32] = (c == p) ? inv : upc;
            m[q] = (in && c > p) ? inv * upc : 0.0;
        }
#pragma unroll
        for (int i = p + 1; i < M; ++i) {
            const double li = (i < nf) ? b[i * 32] : 0.0;
#pragma unroll
            for (int q = 0; q < NC; ++q) col[q][i] -= li * m[q];
        }
    }
    const int u = nf - k;
#pragma unroll
    for (int q = 0; q < NC; ++q) {
        const int c = e0 + q * TE;
        if (c >= k && c <= nf) {
            double* Cj = ub + (unsigned)((c - k) * u * ust);
#pragma unroll
            for (int i = 0; i < M; ++i)
                if (i >= k && i < nf) Cj[(unsigned)((i - k) * ust)] = col[q][i];
        }
    }
}

template <int TE, int MAXNF>
__global__ void __launch_bounds__(32 * TE)
mf_task_kernel(DevSym sy, const int* __restrict__ blobs, const int2* __restrict__ tasks,
               const double* __restrict__ aval, const double* __restrict__ rhs, double* __restrict__ U,
               double* __restrict__ upd, int S, int front_cap, int stack_cap, const unsigned char* __restrict__ active,
               int* __restrict__ status) {
    extern __shared__ __align__(128) double sm[];
    const int sl = threadIdx.x & 31, e0 = threadIdx.x >> 5;
    const int s = blockIdx.y * 32 + sl;
    const bool act = active ? (active[s] != 0) : true;
    if (!__syncthreads_or(act)) return;
    const int2 td = tasks[blockIdx.x];
    double* Fl = sm + sl;
    double* bcl = sm + (size_t)front_cap * 32 + sl;
    double* stk = sm + (size_t)(front_cap + 2 * MAXNF) * 32 + sl;
    int* meta = reinterpret_cast<int*>(sm + (size_t)(front_cap + 2 * MAXNF + stack_cap) * 32);
    for (int i = threadIdx.x; i < td.y; i += 32 * TE) meta[i] = blobs[td.x + i];
    __syncthreads();
    const int nfr = meta[0];
    const double* __restrict__ av = aval + s;
    const double* __restrict__ rv = rhs + s;
    // blocks read from HBM (children outside the task) live in the section of tile width 32 of the update storage
    double* __restrict__ uptile = upd + sy.sec_base[5] + (long long)blockIdx.y * sy.sec_size[5] * 32 + sl;
    double pv[kTaskPre];
    bool bad = false;

    // values of the entries this warp owns in front `fr`, issued one front ahead so the loads overlap the elimination
    auto prefetch = [&](const int* fr) {
        const int* aw = meta + fr[2];
        const int w0 = aw[e0], w1 = aw[e0 + 1];
        const int* pairs = aw + TE + 1;
#pragma unroll
        for (int q = 0; q < kTaskPre; ++q) {
            const int a = w0 + q;
            pv[q] = 0.0;
            if (a < w1) {
                const int src = pairs[2 * a];
                pv[q] = src >= 0 ? av[wide(src, S)] : rv[wide(-src - 1, S)];
            }
        }
    };
    prefetch(meta + 1);
    for (int fi = 0; fi < nfr; ++fi) {
        const int* fr = meta + 1 + fi * kTaskRec;
        const int nf = fr[0], k = fr[1];
        for (int c = e0; c <= nf; c += TE) {
            double* colc = Fl + c * nf * 32;
            for (int i = 0; i < nf; ++i) colc[i * 32] = 0.0;
        }
        {
            const int* aw = meta + fr[2];
            const int w0 = aw[e0], w1 = aw[e0 + 1];
            const int* pairs = aw + TE + 1;
#pragma unroll
            for (int q = 0; q < kTaskPre; ++q) {
                const int a = w0 + q;
                if (a < w1) Fl[pairs[2 * a + 1] * 32] = pv[q];
            }
            for (int a = w0 + kTaskPre; a < w1; ++a) {
                const int src = pairs[2 * a];
                Fl[pairs[2 * a + 1] * 32] = src >= 0 ? av[wide(src, S)] : rv[wide(-src - 1, S)];
            }
        }
        if (fi + 1 < nfr) prefetch(fr + kTaskRec);
        const int* cp = meta + fr[3];
        for (int ci = 0; ci < fr[4]; ++ci) {
            const int uc = cp[0], soff = cp[1];
            const int* rel = cp + 4;
            if (soff >= 0) {
                const double* src = stk + soff * 32;
                for (int j = 0; j <= uc; ++j) {
                    const int C = (j < uc) ? rel[j] : nf;
                    if ((C & (TE - 1)) != e0) continue;
                    double* colC = Fl + C * nf * 32;
                    const double* sj = src + j * uc * 32;
#pragma unroll 4
                    for (int i = 0; i < uc; ++i) colC[rel[i] * 32] += sj[i * 32];
                }
            } else {
                const long long uo = ((long long)cp[3] << 32) | (unsigned)cp[2];
                const double* __restrict__ src = uptile + uo * 32;
                for (int j = 0; j <= uc; ++j) {
                    const int C = (j < uc) ? rel[j] : nf;
                    if ((C & (TE - 1)) != e0) continue;
                    double* colC = Fl + C * nf * 32;
                    const double* __restrict__ sj = src + j * uc * 32;
#pragma unroll 4
                    for (int i = 0; i < uc; ++i) colC[rel[i] * 32] += sj[i * 32];
                }
            }
            cp += 4 + uc;
        }
        const long long uoff = ((long long)fr[7] << 32) | (unsigned)fr[6];
        double* __restrict__ Uf = u_base(U, sy, 32, s) + uoff * 32;
        double* ub;
        int ust = 32;
        if (fr[5] >= 0) {
            ub = stk + fr[5] * 32;
        } else {                                  // task root: the block goes to the section its parent reads
            ust = fr[10];
            ub = upd_base(upd, sy, ust, s) + (((long long)fr[9] << 32) | (unsigned)fr[8]) * ust;
        }
        if constexpr (MAXNF > 12) {
            if (nf > 12) task_eliminate<TE, 16>(Fl, bcl, nf, k, e0, act, Uf, S, ub, ust, bad);
            else if (nf > 8) task_eliminate<TE, 12>(Fl, bcl, nf, k, e0, act, Uf, S, ub, ust, bad);
            else task_eliminate<TE, 8>(Fl, bcl, nf, k, e0, act, Uf, S, ub, ust, bad);
        } else if constexpr (MAXNF > 8) {
            if (nf > 8) task_eliminate<TE, 12>(Fl, bcl, nf, k, e0, act, Uf, S, ub, ust, bad);
            else task_eliminate<TE, 8>(Fl, bcl, nf, k, e0, act, Uf, S, ub, ust, bad);
        } else {
            task_eliminate<TE, 8>(Fl, bcl, nf, k, e0, act, Uf, S, ub, ust, bad);
        }
        __syncthreads();
    }
    if (act && bad && e0 == 0) status[s] = -3;
}

#define JGB_TASK_VARIANTS(X) X(4, 8) X(4, 12) X(8, 16)

void launch_task(int te, int maxnf, dim3 grid, size_t smem, cudaStream_t st, DevSym dev, const int* blobs,
                 const int2* tasks, const double* aval, const double* rhs, double* U, double* upd, int S,
                 int front_cap, int stack_cap, const unsigned char* active, int* status) {
#define X(TE, MAXNF)                                                                                                \
    if (te == TE && maxnf == MAXNF) {                                                                               \
        mf_task_kernel<TE, MAXNF><<<grid, 32 * TE, smem, st>>>(dev, blobs, tasks, aval, rhs, U, upd, S, front_cap,  \
                                                               stack_cap, active, status);                         \
        return;                                                                                                     \
    }
    JGB_TASK_VARIANTS(X)
#undef X
    throw std::runtime_error("unsupported task kernel variant");
}

void set_task_smem_attr() {
#define X(TE, MAXNF) \
    JGB_CUDA(cudaFuncSetAttribute(mf_task_kernel<TE, MAXNF>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    JGB_TASK_VARIANTS(X)
#undef X
}
