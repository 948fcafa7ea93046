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

__global__ void __launch_bounds__(512)
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

// ---- LU twin for unsymmetric values (the Newton-Raphson Jacobian): same three-barrier panel, full column-major front.
// One CTA = one front of one scenario. Per panel of 8 pivots:
//   1. warp 0 factors the 8 x 8 diagonal block in registers (lane = row, pivot rows exchanged with shuffles, no pivoting);
//   2. one thread per row below the block solves its panel row against U11 (L21 = A21 U11^-1) and, at the same time, one
//      thread per column right of the block (the rhs column included) solves U12 = L11^-1 A12; -L21 and U12 go into two
//      padded strips;
//   3. the trailing update C -= L21 U12 runs as 8 x 8 tiles of mma.sync.aligned.m8n8k4.f64.
// The scenario-tile kernel (mf_factor_kernel) spends 8 + 2 barriers per panel and ~2 shared-memory instructions per
// multiply-add; for a single case, where a front is one CTA and the top of the tree is a chain of 45-70-row fronts, that
// latency is the Newton step.
// Shared memory: F [(nf + 1) columns x LD] (column nf = rhs; LD = 2 (mod 8) so the four column pairs of a C fragment fall
// into distinct bank groups) | Lneg [8][LDP] | Up [8][LDP] | LU11 [64] | Uinv [8]; LDP >= nf + 1, = 4 (mod 16).
__host__ __device__ __forceinline__ int dense_lu_ld(int nf) { return ((nf + 5) / 8) * 8 + 2; }          // >= nf, = 2 (mod 8)
__host__ __device__ __forceinline__ int dense_lu_ldp(int nf) { return ((nf + 12) / 16) * 16 + 4; }      // >= nf + 1, = 4 (mod 16)
__host__ __device__ __forceinline__ size_t dense_lu_smem_doubles(int nf) {
    return (size_t)dense_lu_ld(nf) * (nf + 1) + 16 * (size_t)dense_lu_ldp(nf) + 64 + 8;
}

__global__ void __launch_bounds__(512)
mf_factor_dense_lu_kernel(DevSym sy, const FrontDesc* __restrict__ descs, const double* __restrict__ aval,
                          const double* __restrict__ rhs, double* __restrict__ U, double* __restrict__ upd, int S,
                          const unsigned char* __restrict__ active, int* __restrict__ status,
                          const int* __restrict__ seqptr, StagedEa sg) {
    extern __shared__ __align__(128) double Fs[];
    __shared__ __align__(8) uint64_t ea_bar[2];
    const int s = blockIdx.y;
    if (active && !active[s]) return;
    const int tid = threadIdx.x, nth = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nth >> 5;
    // staged extend-add (sg.chunks): the children's blocks — contiguous runs of this scenario's W = 1 section — stream
    // through a two-stage ring behind the front with cp.async.bulk; three dependent memory round trips (descriptor, chunk
    // list, data) whatever the number of children, against three per gather round
    int ea_cnt = 0;                              // chunks consumed so far by this CTA (ring stage and mbarrier parity)
    if (sg.chunks && tid == 0) {
        mbar_init(&ea_bar[0], 1);
        mbar_init(&ea_bar[1], 1);
    }
    // sequence mode (seqptr): the CTA walks a chain of fronts, each the parent of the one before; the update block of a
    // front is in global memory (L2) before the barrier that starts its parent
    const int d0 = seqptr ? seqptr[blockIdx.x] : blockIdx.x, d1 = seqptr ? seqptr[blockIdx.x + 1] : blockIdx.x + 1;
    for (int di = d0; di < d1; ++di) {
    if (di != d0) {        // the block just written with generic stores is read by the bulk copy (async proxy) of the parent
        __threadfence_block();
        asm volatile("fence.proxy.async;" ::: "memory");
        __syncthreads();
    }
    const FrontDesc fd = descs[di];
    const int nf = fd.nf, k = fd.k, u = nf - k;
    const int* __restrict__ rows = sy.f_rows + fd.rowptr;
    const int LD = dense_lu_ld(nf), LDP = dense_lu_ldp(nf);
    double* F = Fs;                              // element (r, c) at F[r + c*LD], c = nf is the right-hand side
    double* Lneg = Fs + LD * (nf + 1);           // -L21 of the current panel: (i, q) at Lneg[i + q*LDP]
    double* Up = Lneg + 8 * LDP;                 // U12 of the current panel: (q, j) at Up[j + q*LDP]
    double* LU11 = Up + 8 * LDP;                 // factored diagonal block, row major: L below, U on and above the diagonal
    double* Uinv = LU11 + 64;
    auto pos_of = [&](int dst) { const int c = dst / nf; return dst + c * (LD - nf); };      // r + c*nf -> r + c*LD
    for (int pos = tid; pos < LD * (nf + 1); pos += nth) Fs[pos] = 0.0;
    __syncthreads();
    // ---- assembly: round 0 of the children's gather as 8-byte cp.async copies straight into the zeroed front (all of a
    // thread's scattered loads in flight together), matrix entries and rhs meanwhile into registers, then the other rounds
    constexpr int W = 1;                         // children are read in the W = 1 section: contiguous per scenario
    double* __restrict__ up = upd_base(upd, sy, 1, s);
    const int nch = sg.chunks ? fd.child1 - fd.child0 : 0;
    double* ring = Fs + sg.ring_off;
    if (nch > 0 && tid == 0) {
        for (int c = 0; c < 2 && c < nch; ++c) {
            const int2 cd = sg.chunks[fd.child0 + c];
            uint64_t* b = &ea_bar[(ea_cnt + c) & 1];
            mbar_expect_tx(b, (uint32_t)cd.y * 8u);
            bulk_g2s(ring + (size_t)((ea_cnt + c) & 1) * sg.ring_elems, up + cd.x, (uint32_t)cd.y * 8u, b);
        }
    }
    if (!sg.chunks && fd.ea1 > fd.ea0) {
        const int t1 = sy.ea_roundptr[fd.ea0 + 1];
        for (int t = sy.ea_roundptr[fd.ea0] + tid; t < t1; t += nth) {
            const int2 pr = sy.ea_pair[t];
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_addr(F + pos_of(pr.x))),
                         "l"(up + (unsigned)(pr.y * W))
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
                pd[q] = pos_of(sy.asm_dst[a]);
                pv[q] = av[wide(sy.asm_src[a], S)];
            }
        }
        if (tid < k) rp = rhs[wide(rows[tid], S) + s];
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();
#pragma unroll
        for (int q = 0; q < NPRE; ++q)
            if (pd[q] >= 0) F[pd[q]] += pv[q];
        for (int a = fd.asm0 + tid + NPRE * nth; a < fd.asm1; a += nth) F[pos_of(sy.asm_dst[a])] += av[wide(sy.asm_src[a], S)];
        if (tid < k) F[tid + nf * LD] += rp;
        for (int p = tid + nth; p < k; p += nth) F[p + nf * LD] += rhs[wide(rows[p], S) + s];
    }
    __syncthreads();
    for (int c = 0; c < nch; ++c) {              // staged: chunks of one child each, child order (sums as in the rounds)
        const int2 cd = sg.chunks[fd.child0 + c];
        const int* __restrict__ dl = sg.upd_dst + sg.sec_cum + cd.x;
        constexpr int NQ = 4;
        int dq[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) dq[q] = (tid + q * nth < cd.y) ? dl[tid + q * nth] : -1;
        const int st = (ea_cnt + c) & 1;
        mbar_wait(&ea_bar[st], ((ea_cnt + c) >> 1) & 1);
        const double* rg = ring + (size_t)st * sg.ring_elems;
#pragma unroll
        for (int q = 0; q < NQ; ++q)
            if (dq[q] >= 0) F[dq[q]] += rg[tid + q * nth];
        for (int e = tid + NQ * nth; e < cd.y; e += nth) F[dl[e]] += rg[e];
        __syncthreads();
        if (tid == 0 && c + 2 < nch) {
            const int2 nd = sg.chunks[fd.child0 + c + 2];
            mbar_expect_tx(&ea_bar[st], (uint32_t)nd.y * 8u);
            bulk_g2s(ring + (size_t)st * sg.ring_elems, up + nd.x, (uint32_t)nd.y * 8u, &ea_bar[st]);
        }
    }
    ea_cnt += nch;
    for (int rd = fd.ea0 + 1; rd < fd.ea1 && !sg.chunks; ++rd) {
        const int t1 = sy.ea_roundptr[rd + 1];
#pragma unroll 4
        for (int t = sy.ea_roundptr[rd] + tid; t < t1; t += nth) {
            const int2 pr = sy.ea_pair[t];
            F[pos_of(pr.x)] += up[(unsigned)(pr.y * W)];
        }
        __syncthreads();
    }
    bool bad = false, weakp = false;
    for (int p0 = 0; p0 < k; p0 += 8) {
        const int pb = (k - p0 < 8) ? k - p0 : 8;
        const int pe = p0 + pb;
        // ---- 1. diagonal block: lane l of warp 0 owns row p0 + l of the block
        if (warp == 0) {
            double r[8];
            const bool mine = lane < pb;
#pragma unroll
            for (int c = 0; c < 8; ++c) r[c] = (mine && c < pb) ? F[p0 + lane + (p0 + c) * LD] : 0.0;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const double dq = __shfl_sync(0xffffffffu, r[q], q);       // pivot (valid while q < pb)
                const bool on = q < pb;
                if (on && (dq == 0.0 || !isfinite(dq))) bad = true;
                const double inv = 1.0 / dq;
                const double lq = (on && mine && lane > q) ? r[q] * inv : 0.0;
                if (on && lane == q) Uinv[q] = inv;
                if (lane > q) r[q] = lq;
#pragma unroll
                for (int c = q + 1; c < 8; ++c) {
                    const double uqc = __shfl_sync(0xffffffffu, r[c], q);
                    r[c] -= lq * uqc;
                }
            }
            if (lane < 8) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    LU11[lane * 8 + c] = r[c];
                    if (mine && c < pb && c >= lane) F[p0 + lane + (p0 + c) * LD] = r[c];      // U11 is part of the output rows
                }
            }
        }
        __syncthreads();
        // ---- 2. L21 = A21 U11^-1 (one thread per row below) and U12 = L11^-1 A12 (one thread per column to the right)
        {
            const int nrow = nf - pe, ncol = nf + 1 - pe;
            for (int t = tid; t < nrow + ncol; t += nth) {
                double w[8];
                if (t < nrow) {
                    const int i = pe + t;
#pragma unroll
                    for (int q = 0; q < 8; ++q) w[q] = (q < pb) ? F[i + (p0 + q) * LD] : 0.0;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        if (q < pb) {
#pragma unroll
                            for (int r2 = 0; r2 < q; ++r2) w[q] -= w[r2] * LU11[r2 * 8 + q];
                            w[q] *= Uinv[q];
                        }
                        Lneg[i + q * LDP] = -w[q];
                    }
                } else {
                    const int j = pe + (t - nrow);
#pragma unroll
                    for (int q = 0; q < 8; ++q) w[q] = (q < pb) ? F[p0 + q + j * LD] : 0.0;
#pragma unroll
                    for (int q = 1; q < 8; ++q) {
                        if (q < pb) {
#pragma unroll
                            for (int r2 = 0; r2 < q; ++r2) w[q] -= LU11[q * 8 + r2] * w[r2];
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        if (q < pb) F[p0 + q + j * LD] = w[q];
                        Up[j + q * LDP] = w[q];
                    }
                }
            }
        }
        __syncthreads();
        // ---- 3. trailing update on 8 x 8 tiles: C[i][j] += sum_q Lneg[i][q] * Up[q][j], rhs column included
        {
            const int nt = nf - pe;
            const int tr = (nt + 7) >> 3, tc = (nt + 8) >> 3;
            const int g = lane >> 2, tg = lane & 3;
            const int ntile = tr * tc;
            for (int t = warp; t < ntile; t += nwarps) {
                const int tj = t / tr, ti = t - tj * tr;           // consecutive warps walk down a column of tiles
                const int i = pe + 8 * ti + g, j0 = pe + 8 * tj, ja = j0 + 2 * tg, jb = ja + 1;
                const int jr = j0 + g;                              // column this lane feeds as B[k = tg][n = g]
                const bool iv = i < nf;
                const bool va = iv && ja <= nf, vb = iv && jb <= nf;
                double c0 = va ? F[i + ja * LD] : 0.0;
                double c1 = vb ? F[i + jb * LD] : 0.0;
                const double a0 = iv ? Lneg[i + tg * LDP] : 0.0, a1 = iv ? Lneg[i + (4 + tg) * LDP] : 0.0;
                const double b0 = jr <= nf ? Up[jr + tg * LDP] : 0.0, b1 = jr <= nf ? Up[jr + (4 + tg) * LDP] : 0.0;
                dmma_m8n8k4(c0, c1, a0, b0);
                dmma_m8n8k4(c0, c1, a1, b1);
                if (va) F[i + ja * LD] = c0;
                if (vb) F[i + jb * LD] = c1;
            }
        }
        __syncthreads();
    }
    if (bad && tid == 0) status[s] = -3;
    // ---- packed U rows (1 / pivot first, rhs entry last) with the pivot guard, and the update block
    const int Wu = (fd.flags >> 8) & 0xff;
    double* __restrict__ Uf = u_base(U, sy, Wu, s) + fd.uoff * Wu;
    for (int p = warp; p < k; p += nwarps) {
        double* Urow = Uf + urow_off(p, nf) * Wu;
        const double lim = sy.growth * fabs(F[p + p * LD]);
        for (int j = p + lane; j <= nf; j += 32) {
            const double v = F[p + j * LD];
            if (j < nf && fabs(v) > lim) weakp = true;
            Urow[(unsigned)((j - p) * Wu)] = (j == p) ? 1.0 / v : v;
        }
    }
    if (weakp && sy.weak) sy.weak[s] = 1;
    const int Wo = fd.wout;
    double* __restrict__ Cf = upd_base(upd, sy, Wo, s) + fd.updoff * Wo;
    for (int j = warp; j <= u; j += nwarps) {
        double* Cj = Cf + (unsigned)(j * u * Wo);
        const double* colj = F + k + (k + j) * LD;
        for (int i = lane; i < u; i += 32) Cj[(unsigned)(i * Wo)] = colj[i];
    }
    }
}
