// Big-front LDL^T kernel: one CTA = one front of one scenario, trailing Schur update on the FP64 tensor-core MMA.
// Included by solver.cu inside namespace jgb { namespace { ... } }.
//
// Replaces, for fronts above ~48 rows of a symmetric matrix (the WLS gain matrix), the panel loop of
// mf_factor_sym_kernel<1>: there every one of the 8 rank-1 steps of a panel and the scalar trailing update sit between
// CTA-wide barriers (ncu: 12 barrier-stall cycles per issued instruction on the 1024-thread CTAs of the top fronts).
// Here a panel of 8 pivots costs three barriers:
//   1. warp 0 factors the 8 x 8 diagonal block in registers (lane = row, pivot rows exchanged with shuffles) and
//      forward-substitutes the block of the right-hand side;
//   2. one thread per row below the block solves its row of the panel against that block (W = A21 L11^-T, L21 = W D^-1,
//      36 multiply-adds, no communication), writes W back into the front (it is the packed-U output), -L21 and W into two
//      padded strips, and updates its right-hand-side entry;
//   3. the trailing update C -= L21 W^T runs as 8 x 8 tiles of mma.sync.aligned.m8n8k4.f64 (two k-steps per panel): per
//      tile 2 + 2 loads / stores of C and 4 strip loads feed 512 multiply-adds, against ~300 shared-memory loads for the
//      same work in the scalar update.
// tcgen05 has no FP64 kind, and an Ozaki-style int8 split does not pay at these sizes (k <= 55 pivots per front: the
// TMEM read-back + FP64 recombination of >= 13 int32 accumulator planes per C element costs as much as the k FP64
// multiply-adds it replaces), so the FP64 tensor instruction is the one that keeps 1e-8 parity.
//
// Shared memory: packed lower triangle of the front (column j at j*(2nf-j+1)/2, nf-j entries) | rhs [nf] |
// Lneg [8][LDP] | Wp [8][LDP] | L11 [64] | Dinv [8] | Yb [8]; LDP = nf rounded up to 4 (mod 16), so that the four
// k-columns a half-warp reads fall into distinct bank groups.
__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__host__ __device__ __forceinline__ int dense_ldp(int nf) { return ((nf + 11) / 16) * 16 + 4; }   // >= nf, = 4 (mod 16)

__host__ __device__ __forceinline__ size_t dense_smem_doubles(int nf) {
    return (size_t)nf * (nf + 1) / 2 + nf + 16 * (size_t)dense_ldp(nf) + 64 + 16;
}

__global__ void __launch_bounds__(256)
mf_factor_dense_sym_kernel(DevSym sy, const FrontDesc* __restrict__ descs, const double* __restrict__ aval,
                           const double* __restrict__ rhs, double* __restrict__ U, double* __restrict__ upd, int S,
                           const unsigned char* __restrict__ active, int* __restrict__ status) {
    extern __shared__ double Fs[];
    const int s = blockIdx.y;
    if (active && !active[s]) return;
    const FrontDesc fd = descs[blockIdx.x];
    const int nf = fd.nf, k = fd.k, u = nf - k;
    const int tid = threadIdx.x, nth = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nth >> 5;
    const int* __restrict__ rows = sy.f_rows + fd.rowptr;
    const int tri = (nf * (nf + 1)) >> 1;
    const int LDP = dense_ldp(nf);
    double* F = Fs;                          // packed lower triangle
    double* R = Fs + tri;                    // rhs
    double* Lneg = R + nf;                   // -L21 of the current panel: element (i, q) at Lneg[i + q*LDP]
    double* Wp = Lneg + 8 * LDP;             // W = D L21^T rows of the current panel, same layout
    double* L11 = Wp + 8 * LDP;              // unit lower 8 x 8 block: L11[r*8 + c], c < r
    double* Dinv = L11 + 64;
    double* Yb = Dinv + 8;                   // forward-substituted rhs of the block
    for (int pos = tid; pos < tri + nf; pos += nth) Fs[pos] = 0.0;
    __syncthreads();
    // ---- assembly. Round 0 of the children's gather (the first source of every destination: for these fronts the whole
    // block of the largest child) goes straight into the zeroed front as 8-byte cp.async copies, so all of a thread's
    // scattered loads are in flight together; the matrix entries and the rhs are added once the copies have landed, then
    // the remaining rounds (symmetric lists: lower triangle + rhs, packed destinations; the rhs vector follows the
    // triangle in shared memory, so one index addresses both).
    constexpr int W = 1;                         // this kernel reads its children in the W = 1 section: contiguous per scenario
    double* __restrict__ up = upd_base(upd, sy, 1, s);
    if (fd.ea1 > fd.ea0) {
        const int t1 = sy.ea_roundptr_s[fd.ea0 + 1];
        for (int t = sy.ea_roundptr_s[fd.ea0] + tid; t < t1; t += nth) {
            const int2 pr = sy.ea_pair_s[t];
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_addr(Fs + pr.x)), "l"(up + (unsigned)(pr.y * W))
                         : "memory");
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    {
        constexpr int NPRE = 4;
        const double* __restrict__ av = aval + s;
        double pv[NPRE], rp = 0.0;
        int pd[NPRE];
#pragma unroll
        for (int q = 0; q < NPRE; ++q) {
            const int a = fd.asm0 + tid + q * nth;
            pd[q] = -1;
            pv[q] = 0.0;
            if (a < fd.asm1) {
                const int dst = sy.asm_dst[a];
                const int c = dst / nf, r = dst - c * nf;
                if (r >= c) {
                    pd[q] = sym_col(c, nf) + r - c;
                    pv[q] = av[wide(sy.asm_src[a], S)];
                }
            }
        }
        if (tid < k) rp = rhs[wide(rows[tid], S) + s];
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();
#pragma unroll
        for (int q = 0; q < NPRE; ++q)
            if (pd[q] >= 0) F[pd[q]] += pv[q];
        for (int a = fd.asm0 + tid + NPRE * nth; a < fd.asm1; a += nth) {
            const int dst = sy.asm_dst[a];
            const int c = dst / nf, r = dst - c * nf;
            if (r >= c) F[sym_col(c, nf) + r - c] += av[wide(sy.asm_src[a], S)];
        }
        if (tid < k) R[tid] += rp;
        for (int p = tid + nth; p < k; p += nth) R[p] += rhs[wide(rows[p], S) + s];
    }
    __syncthreads();
    for (int rd = fd.ea0 + 1; rd < fd.ea1; ++rd) {
        const int t1 = sy.ea_roundptr_s[rd + 1];
#pragma unroll 8
        for (int t = sy.ea_roundptr_s[rd] + tid; t < t1; t += nth) {
            const int2 pr = sy.ea_pair_s[t];
            Fs[pr.x] += up[(unsigned)(pr.y * W)];
        }
        __syncthreads();
    }
    bool bad = false;
    for (int p0 = 0; p0 < k; p0 += 8) {
        const int pb = (k - p0 < 8) ? k - p0 : 8;
        const int pe = p0 + pb;
        // ---- 1. diagonal block: lane l of warp 0 owns row p0 + l (its entries left of and on the diagonal)
        if (warp == 0) {
            double r[8];
            const int row = p0 + lane;
            const bool mine = lane < pb;
#pragma unroll
            for (int c = 0; c < 8; ++c) r[c] = (mine && c <= lane) ? F[sym_col(p0 + c, nf) + row - (p0 + c)] : 0.0;
            double y = mine ? R[row] : 0.0;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const double dq = __shfl_sync(0xffffffffu, r[q], q);       // pivot (valid while q < pb)
                const double yq = __shfl_sync(0xffffffffu, y, q);
                const bool on = q < pb;
                if (on && (dq == 0.0 || !isfinite(dq))) bad = true;
                const double inv = 1.0 / dq;
                const double lq = (on && mine && lane > q) ? r[q] * inv : 0.0;      // L11[lane][q]
                if (on && lane == q) Dinv[q] = inv;
                if (on && mine && lane > q) L11[lane * 8 + q] = lq;
#pragma unroll
                for (int c = q + 1; c < 8; ++c) {
                    const double wc = __shfl_sync(0xffffffffu, r[q], c);            // W[c][q] = unscaled entry of row c
                    if (c <= lane) r[c] -= lq * wc;
                }
                y -= lq * yq;
            }
            if (mine) {
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    if (c <= lane) F[sym_col(p0 + c, nf) + row - (p0 + c)] = r[c];
                R[row] = y;
                Yb[lane] = y;
            } else if (lane < 8) {
                Yb[lane] = 0.0;
            }
        }
        __syncthreads();
        // ---- 2. panel rows below the block: W = A21 L11^-T, L21 = W D^-1; one thread per row
        for (int i = pe + tid; i < nf; i += nth) {
            double w[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) w[q] = (q < pb) ? F[sym_col(p0 + q, nf) + i - (p0 + q)] : 0.0;
#pragma unroll
            for (int q = 1; q < 8; ++q) {
                if (q < pb) {
#pragma unroll
                    for (int r2 = 0; r2 < q; ++r2) w[q] -= w[r2] * L11[q * 8 + r2];
                }
            }
            double acc = R[i];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                double l = 0.0;
                if (q < pb) {
                    l = w[q] * Dinv[q];
                    F[sym_col(p0 + q, nf) + i - (p0 + q)] = w[q];
                    acc -= l * Yb[q];
                }
                Lneg[i + q * LDP] = -l;
                Wp[i + q * LDP] = (q < pb) ? w[q] : 0.0;
            }
            R[i] = acc;
        }
        __syncthreads();
        // ---- 3. trailing update on 8 x 8 tiles: C[i][j] += sum_q Lneg[i][q] * Wp[j][q], lower triangle only
        {
            const int nt = nf - pe, tiles = (nt + 7) >> 3;
            const int g = lane >> 2, tg = lane & 3;
            const int ntile = tiles * (tiles + 1) / 2;
            int ti = 0, tj = 0, tcur = 0;             // (ti, tj) of tile tcur, advanced incrementally
            for (int t = warp; t < ntile; t += nwarps) {
                while (tcur < t) {
                    const int step = min(t - tcur, ti - tj + 1);      // tiles left in row ti, or all the way to t
                    if (step == ti - tj + 1) { ++ti; tj = 0; } else tj += step;
                    tcur += step;
                }
                const int i = pe + 8 * ti + g, j0 = pe + 8 * tj, ja = j0 + 2 * tg, jb = ja + 1;
                const int jr = j0 + g;                                   // row of Wp this lane feeds as B[k = tg][n = g]
                const bool iv = i < nf;
                const bool va = iv && ja < nf && ja <= i, vb = iv && jb < nf && jb <= i;
                double c0 = va ? F[sym_col(ja, nf) + i - ja] : 0.0;
                double c1 = vb ? F[sym_col(jb, nf) + i - jb] : 0.0;
                const double a0 = iv ? Lneg[i + tg * LDP] : 0.0, a1 = iv ? Lneg[i + (4 + tg) * LDP] : 0.0;
                const double b0 = jr < nf ? Wp[jr + tg * LDP] : 0.0, b1 = jr < nf ? Wp[jr + (4 + tg) * LDP] : 0.0;
                dmma_m8n8k4(c0, c1, a0, b0);
                dmma_m8n8k4(c0, c1, a1, b1);
                if (va) F[sym_col(ja, nf) + i - ja] = c0;
                if (vb) F[sym_col(jb, nf) + i - jb] = c1;
            }
        }
        __syncthreads();
    }
    if (bad && tid == 0) status[s] = -3;
    // ---- packed U rows (1 / d_p, then d_p L[j,p] = W[j,p], then the forward-substituted rhs) and the update block
    const int Wu = (fd.flags >> 8) & 0xff;      // tile width of the back-solve launch that reads these rows (1: contiguous)
    double* __restrict__ Uf = u_base(U, sy, Wu, s) + fd.uoff * Wu;
    for (int p = warp; p < k; p += nwarps) {
        double* Urow = Uf + urow_off(p, nf) * Wu;
        const double* colp = F + sym_col(p, nf) - p;
        for (int j = p + lane; j <= nf; j += 32) {
            const double v = (j < nf) ? colp[j] : R[p];
            Urow[(unsigned)((j - p) * Wu)] = (j == p) ? 1.0 / v : v;
        }
    }
    const int Wo = fd.wout;                     // the parent's tile width (1 when the parent is a dense front: contiguous)
    double* __restrict__ Cf = upd_base(upd, sy, Wo, s) + fd.updoff * Wo;
    const bool lower_only = fd.flags & 1;       // the parent is an LDL^T front too: it never reads above the diagonal
    for (int j = warp; j <= u; j += nwarps) {   // one warp per column of the block: no index division
        double* Cj = Cf + (unsigned)(j * u * Wo);
        if (j == u) {
            for (int i = lane; i < u; i += 32) Cj[(unsigned)(i * Wo)] = R[k + i];
        } else {
            const double* colj = F + sym_col(k + j, nf) - j;          // entry (k + i, k + j), i >= j, at colj[i]
            for (int i = j + lane; i < u; i += 32) Cj[(unsigned)(i * Wo)] = colj[i];
            if (!lower_only)
                for (int i = lane; i < j; i += 32) Cj[(unsigned)(i * Wo)] = F[sym_col(k + i, nf) + j - i];
        }
    }
}
