// Numeric multifrontal kernels. See solver.cuh.
//
// Kernel map: mf_factor_kernel = K2 (front assembly) + K3/K4 (dense partial LU + Schur
// update) + forward substitution (the rhs rides along as column nf of every front, so L is never stored);
// mf_backsolve_* = K5 (backward substitution, top-down over the elimination tree).
#include "solver.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>

namespace jgb {

namespace {

// Element offsets: all operands are non-negative 32-bit values, so index products are formed with one widening
// multiply (mul.wide.u32) or stay in 32 bits, instead of sign-extended 64-bit multiply sequences.
__device__ __forceinline__ size_t wide(int a, int b) { return (size_t)(unsigned)a * (unsigned)b; }

// log2 of a power-of-two tile width (1..32)
__host__ __device__ __forceinline__ int lg2(int w) { return w >= 32 ? 5 : w >= 16 ? 4 : w >= 8 ? 3 : w >= 4 ? 2 : w >= 2 ? 1 : 0; }
// base of scenario s in the update-storage section of tile width W (see DevSym): add (off + e) * W for an element
// (W is a power of two: shifts and masks instead of the integer division a runtime W would cost in every thread)
__device__ __forceinline__ double* upd_base(double* upd, const DevSym& sy, int W, int s) {
    const int lw = lg2(W);
    return upd + sy.sec_base[lw] + (((long long)(s >> lw) * sy.sec_size[lw]) << lw) + (s & (W - 1));
}
// same for the packed U rows (usec_* sections): add (f_uoff + e) * W for an entry
__device__ __forceinline__ double* u_base(double* U, const DevSym& sy, int W, int s) {
    const int lw = lg2(W);
    return U + sy.usec_base[lw] + (((long long)(s >> lw) * sy.usec_size[lw]) << lw) + (s & (W - 1));
}
__device__ __forceinline__ const double* u_base(const double* U, const DevSym& sy, int W, int s) {
    const int lw = lg2(W);
    return U + sy.usec_base[lw] + (((long long)(s >> lw) * sy.usec_size[lw]) << lw) + (s & (W - 1));
}
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ long long urow_off(int p, int nf) {
    return (long long)p * (nf + 1) - (long long)p * (p - 1) / 2;
}

// ---- mbarrier / bulk-copy (TMA) helpers ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}

// Trailing update A22 -= L21 U12 of one panel of PB pivots (see mf_factor_kernel): lane (er, ec) owns rows er, er + TR, ...
// in pairs and every TC-th column; the PB multipliers of its two rows stay in registers, the pivot-row entries are
// shared-memory broadcasts.
template <int TS, int PB, bool GLOBAL_F, bool TWO>
__device__ __forceinline__ void trailing_rows(double* rowi, double* rowi2, const double* l, const double* l2, double* Fl,
                                              const double* Ul, int nf, int p0, int pe, int ec, int TC) {
    constexpr int B = 8;
    const int colstride = nf * TS;
    int j = pe + ec;
    for (; j + TC <= nf; j += 2 * TC) {
        const double *ua, *ub;
        if constexpr (GLOBAL_F) { ua = Ul + j * B * TS; ub = Ul + (j + TC) * B * TS; }
        else { ua = Fl + (p0 + j * nf) * TS; ub = Fl + (p0 + (j + TC) * nf) * TS; }
        double a1 = rowi[j * colstride], b1 = rowi[(j + TC) * colstride], a2 = 0.0, b2 = 0.0;
        if constexpr (TWO) { a2 = rowi2[j * colstride]; b2 = rowi2[(j + TC) * colstride]; }
#pragma unroll
        for (int q = 0; q < PB; ++q) {
            const double x = ua[q * TS], y = ub[q * TS];
            a1 -= l[q] * x;
            b1 -= l[q] * y;
            if constexpr (TWO) { a2 -= l2[q] * x; b2 -= l2[q] * y; }
        }
        rowi[j * colstride] = a1;
        rowi[(j + TC) * colstride] = b1;
        if constexpr (TWO) {
            rowi2[j * colstride] = a2;
            rowi2[(j + TC) * colstride] = b2;
        }
    }
    if (j <= nf) {
        const double* ua;
        if constexpr (GLOBAL_F) ua = Ul + j * B * TS;
        else ua = Fl + (p0 + j * nf) * TS;
        double a1 = rowi[j * colstride], a2 = 0.0;
        if constexpr (TWO) a2 = rowi2[j * colstride];
#pragma unroll
        for (int q = 0; q < PB; ++q) {
            const double x = ua[q * TS];
            a1 -= l[q] * x;
            if constexpr (TWO) a2 -= l2[q] * x;
        }
        rowi[j * colstride] = a1;
        if constexpr (TWO) rowi2[j * colstride] = a2;
    }
}

template <int TS, int PB, bool GLOBAL_F>
__device__ __forceinline__ void trailing_update(double* Fl, const double* pan, const double* Ul, int nf, int p0, int pe,
                                                int er, int ec, int TR, int TC) {
    for (int i = pe + er; i < nf; i += 2 * TR) {
        const int i2 = i + TR;
        double* rowi = Fl + i * TS;
        double l[PB], l2[PB];
#pragma unroll
        for (int q = 0; q < PB; ++q) l[q] = pan[(i + q * nf) * TS];
        if (i2 < nf) {          // two rows share every pivot-row read
#pragma unroll
            for (int q = 0; q < PB; ++q) l2[q] = pan[(i2 + q * nf) * TS];
            trailing_rows<TS, PB, GLOBAL_F, true>(rowi, Fl + i2 * TS, l, l2, Fl, Ul, nf, p0, pe, ec, TC);
        } else {
            trailing_rows<TS, PB, GLOBAL_F, false>(rowi, rowi, l, l, Fl, Ul, nf, p0, pe, ec, TC);
        }
    }
}

#ifndef JGB_MID_MINBLOCKS
#define JGB_MID_MINBLOCKS 3
#endif
constexpr int kEaChunkCache = 96;      // chunk descriptors of a front kept in shared memory (ring-mode extend-add)

// One CTA = one front x TS scenarios. Thread t: scenario lane sl = t % TS, entry lane e = t / TS,
// entry lanes are arranged TR (rows) x TC (columns). Front F is column major, ld = nf, nf+1 columns,
// element (r,c) of scenario lane sl at F[(r + c*nf) * TS + sl].  TS and the address space of F are compile-time
// so that the front is addressed with LDS/STS and shifts (a runtime select would degrade to generic LD/ST).
template <int TS, bool GLOBAL_F>
__global__ void __launch_bounds__(TS == 1 ? 512 : 256, TS == 1 ? 1 : JGB_MID_MINBLOCKS)
mf_factor_kernel(DevSym sy, const int* __restrict__ fronts, const FrontDesc* __restrict__ descs,
                 const double* __restrict__ aval, const double* __restrict__ rhs, double* __restrict__ U,
                 double* __restrict__ upd, int S, int TR, const unsigned char* __restrict__ active,
                 int* __restrict__ status, double* gwork, long long gstride, int ea_async, StagedEa sg) {
    extern __shared__ __align__(128) double Fs[];
    __shared__ __align__(8) uint64_t ea_bar[2];
    __shared__ int2 ea_ch[kEaChunkCache];
    const int sl = threadIdx.x % TS;
    double* Fl;     // this thread's scenario lane of the front: element (r,c) at Fl[(r + c*nf) * TS]
    if constexpr (GLOBAL_F) Fl = gwork + ((long long)blockIdx.y * gridDim.x + blockIdx.x) * gstride + sl;
    else Fl = Fs + sl;
    const int e0 = threadIdx.x / TS;
    const int TE = blockDim.x / TS;
    const int er = e0 % TR, ec = e0 / TR, TC = TE / TR;
    const int s = blockIdx.y * TS + sl;
    const bool act = active ? (active[s] != 0) : true;
    if (!__syncthreads_or(act)) return;
    // batch launches pass per-front descriptors (one 64-byte read instead of the fronts[] -> f_* double indirection)
    FrontDesc fd;
    if (descs) {
        fd = descs[blockIdx.x];
    } else {
        const int f = fronts[blockIdx.x];
        fd.f = f; fd.nf = sy.f_nf[f]; fd.k = sy.f_k[f]; fd.rowptr = sy.f_rowptr[f];
        fd.asm0 = sy.f_asmptr[f]; fd.asm1 = sy.f_asmptr[f + 1];
        fd.ea0 = sy.f_eaptr[f]; fd.ea1 = sy.f_eaptr[f + 1];
        fd.child0 = fd.child1 = 0;
        fd.wout = S < 32 ? S : 32;
        fd.flags = fd.wout << 8;
        fd.uoff = sy.f_uoff[f]; fd.updoff = sy.f_updoff[f];
    }
    const int nf = fd.nf, k = fd.k, u = nf - k;
    const int* __restrict__ rows = sy.f_rows + fd.rowptr;
    const int total = nf * (nf + 1);

    // the children's blocks live in the update-storage section of this launch's tile width (see DevSym)
    constexpr int W = TS;
    double* __restrict__ up = upd_base(upd, sy, TS, s);
    // Staged extend-add (batches, TS >= 2): the children's blocks are contiguous runs of this tile's section, so one
    // elected thread streams them through a two-stage ring behind the front with cp.async.bulk (chunks of one child,
    // child order), completion on an mbarrier per stage; the first two chunks fly while the front is zeroed and the
    // matrix entries are assembled. Destinations come from a list in source order (upd_dst), one coalesced read.
    constexpr bool kCanStage = (TS >= 2) && !GLOBAL_F;
    int nch = 0;
    double* ring = nullptr;
    const double* tile_src = nullptr;
    if constexpr (kCanStage) {
        if (sg.chunks) {
            nch = fd.child1 - fd.child0;
            ring = Fs + sg.ring_off;
            tile_src = upd + sy.sec_base[lg2(TS)] + (long long)blockIdx.y * sy.sec_size[lg2(TS)] * TS;
            if (threadIdx.x == 0 && nch > 0) {
                mbar_init(&ea_bar[0], 1);
                mbar_init(&ea_bar[1], 1);
                for (int c = 0; c < 2 && c < nch; ++c) {
                    const int2 cd = sg.chunks[fd.child0 + c];
                    const uint32_t bytes = (uint32_t)cd.y * (TS * 8);
                    mbar_expect_tx(&ea_bar[c], bytes);
                    bulk_g2s(ring + (size_t)c * sg.ring_elems * TS, tile_src + (size_t)cd.x * TS, bytes, &ea_bar[c]);
                }
            }
        }
    }
    if constexpr (kCanStage) {
        if (sg.chunks)
            for (int c = threadIdx.x; c < nch && c < kEaChunkCache; c += blockDim.x) ea_ch[c] = sg.chunks[fd.child0 + c];
    }
    // Data-movement phases (zeroing, staged extend-add, write-out) run on scenario PAIRS: thread = (entry, two adjacent
    // scenarios), 16-byte shared-memory and global accesses — half the instructions of the (entry, scenario) mapping
    // the elimination uses (ncu: 62 % issue utilisation, these phases were ~45 % of the executed instructions)
    constexpr int H = TS >= 2 ? TS / 2 : 1;
    const int sl2 = threadIdx.x % H, ev = threadIdx.x / H, TEv = blockDim.x / H;
    if constexpr (TS >= 2 && !GLOBAL_F) {
        double2* F2 = reinterpret_cast<double2*>(Fs);
        for (int pos = threadIdx.x; pos < total * H; pos += blockDim.x) F2[pos] = make_double2(0.0, 0.0);
    } else {
        for (int pos = e0; pos < total; pos += TE) Fl[pos * TS] = 0.0;
    }
    __syncthreads();
    // Round 0 of the extend-add (the first source of every destination — for the fronts at the top of the tree that is
    // the whole block of the largest child) goes straight into the zeroed front as 8-byte cp.async copies: no register
    // staging and no scoreboard, so all of a thread's gathers are in flight together instead of four at a time. The
    // matrix entries are fetched into registers meanwhile and added once the copies have landed.
    bool async_r0 = false;
    if constexpr (!GLOBAL_F) {
        if (ea_async && fd.ea1 > fd.ea0 && !sg.chunks) {
            async_r0 = true;
            if (act) {
                const int t1 = sy.ea_roundptr[fd.ea0 + 1];
                for (int t = sy.ea_roundptr[fd.ea0] + e0; t < t1; t += TE) {
                    const int2 pr = sy.ea_pair[t];
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_addr(Fl + pr.x * TS)),
                                 "l"(up + (unsigned)(pr.y * W))
                                 : "memory");
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
    }
    {
        const double* __restrict__ av = aval + s;
        const int a1 = fd.asm1;
        constexpr int NPRE = 4;
        double pv[NPRE], rp = 0.0;
        int pd[NPRE];
        if (async_r0) {
            if (act) {
#pragma unroll
                for (int q = 0; q < NPRE; ++q) {
                    const int a = fd.asm0 + e0 + q * TE;
                    pd[q] = -1;
                    pv[q] = 0.0;
                    if (a < a1) {
                        pd[q] = sy.asm_dst[a];
                        pv[q] = av[wide(sy.asm_src[a], S)];
                    }
                }
                if (e0 < k) rp = rhs[wide(rows[e0], S) + s];
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncthreads();
            if (act) {
#pragma unroll
                for (int q = 0; q < NPRE; ++q)
                    if (pd[q] >= 0) Fl[pd[q] * TS] += pv[q];
                for (int a = fd.asm0 + e0 + NPRE * TE; a < a1; a += TE)
                    Fl[sy.asm_dst[a] * TS] += av[wide(sy.asm_src[a], S)];
                if (e0 < k) Fl[(e0 + nf * nf) * TS] += rp;
                for (int p = e0 + TE; p < k; p += TE) Fl[(p + nf * nf) * TS] += rhs[wide(rows[p], S) + s];
            }
        } else if (act) {
            for (int a = fd.asm0 + e0; a < a1; a += TE) Fl[sy.asm_dst[a] * TS] = av[wide(sy.asm_src[a], S)];
            for (int p = e0; p < k; p += TE) Fl[(p + nf * nf) * TS] = rhs[wide(rows[p], S) + s];
        }
    }
    __syncthreads();
    if (kCanStage && sg.chunks) {
        if constexpr (kCanStage) {
            // chunk c + 1's destination indices are fetched while chunk c is waited for and added (one chunk ahead), the
            // chunk descriptors sit in shared memory: no dependent global load between two chunks
            constexpr int NQ = 2;
            auto chunk_at = [&](int c) { return c < kEaChunkCache ? ea_ch[c] : sg.chunks[fd.child0 + c]; };
            const int* __restrict__ dbase = sg.upd_dst + sg.sec_cum;
            double2* F2 = reinterpret_cast<double2*>(Fs) + sl2;
            int2 cd = nch > 0 ? chunk_at(0) : make_int2(0, 0);
            int di[NQ];
#pragma unroll
            for (int q = 0; q < NQ; ++q) di[q] = (ev + q * TEv < cd.y) ? dbase[cd.x + ev + q * TEv] : -1;
            for (int c = 0; c < nch; ++c) {
                const int2 cn = c + 1 < nch ? chunk_at(c + 1) : make_int2(0, 0);
                int dn[NQ];
#pragma unroll
                for (int q = 0; q < NQ; ++q) dn[q] = (ev + q * TEv < cn.y) ? dbase[cn.x + ev + q * TEv] : -1;
                mbar_wait(&ea_bar[c & 1], (c >> 1) & 1);
                const double2* rg = reinterpret_cast<const double2*>(ring + (size_t)(c & 1) * sg.ring_elems * TS) + sl2;
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    if (di[q] >= 0) {
                        const double2 a = rg[(ev + q * TEv) * H];
                        double2 f = F2[di[q] * H];
                        f.x += a.x; f.y += a.y;
                        F2[di[q] * H] = f;
                    }
                }
                for (int e = ev + NQ * TEv; e < cd.y; e += TEv) {
                    const int d = dbase[cd.x + e];
                    const double2 a = rg[e * H];
                    double2 f = F2[d * H];
                    f.x += a.x; f.y += a.y;
                    F2[d * H] = f;
                }
                __syncthreads();
                if (threadIdx.x == 0 && c + 2 < nch) {
                    const int2 nd = chunk_at(c + 2);
                    const uint32_t bytes = (uint32_t)nd.y * (TS * 8);
                    mbar_expect_tx(&ea_bar[c & 1], bytes);
                    bulk_g2s(ring + (size_t)(c & 1) * sg.ring_elems * TS, tile_src + (size_t)nd.x * TS, bytes, &ea_bar[c & 1]);
                }
                cd = cn;
#pragma unroll
                for (int q = 0; q < NQ; ++q) di[q] = dn[q];
            }
        }
    } else {
        // extend-add of all children as a gather in rounds (child order per destination, so sums are deterministic)
        const int r1 = fd.ea1;
        for (int r = fd.ea0 + (async_r0 ? 1 : 0); r < r1; ++r) {      // rounds: distinct destinations inside a round
            const int t1 = sy.ea_roundptr[r + 1];
            if (act) {
#pragma unroll 4
                for (int t = sy.ea_roundptr[r] + e0; t < t1; t += TE) {
                    const int2 pr = sy.ea_pair[t];
                    Fl[pr.x * TS] += up[(unsigned)(pr.y * W)];
                }
            }
            __syncthreads();
        }
    }
    bool bad = false, weakp = false;
    const int colstride = nf * TS;

    // Blocked right-looking elimination, panels of B pivots:
    //  (a) panel: rank-1 steps restricted to the panel columns (one barrier each, a handful of multiply-adds per
    //      row), the multipliers l_i = F[i,p] / F[p,p] overwrite column p (L is not an output);
    //  (b) U12 = L11^-1 A12: one lane per trailing column (incl. the rhs column), no barrier inside;
    //  (c) A22 -= L21 U12: row-owner lanes keep their B multipliers in registers and sweep the columns — the pivot-row
    //      reads are shared-memory broadcasts, ~2 instructions per multiply-add.
    constexpr int B = 8;
    // Fronts kept in global memory (GLOBAL_F) stage the current panel (nf x B) and the U12 strip (B x (nf+1)) in
    // shared memory, so the barrier-separated panel steps never wait on L2 round trips.
    double* Pl = Fs + sl;                               // GLOBAL_F: panel, element (i, q) at Pl[(i + q*nf) * TS]
    double* Ul = Fs + (size_t)nf * B * TS + sl;         // GLOBAL_F: strip, element (q, j) at Ul[(q + j*B) * TS]
    for (int p0 = 0; p0 < k; p0 += B) {
        const int pe = (p0 + B < k) ? p0 + B : k;
        const int pb = pe - p0;
        double* pan;          // panel base: element (i, p) at pan[(i + (p - p0)*nf) * TS]
        if constexpr (GLOBAL_F) {
            for (int t = e0; t < nf * pb; t += TE) Pl[t * TS] = Fl[(p0 * nf + t) * TS];
            __syncthreads();
            pan = Pl;
        } else {
            pan = Fl + p0 * colstride;
        }
        for (int p = p0; p < pe; ++p) {
            const double* colp = pan + (p - p0) * colstride;
            const double piv = colp[p * TS];
            if (piv == 0.0 || !isfinite(piv)) bad = true;
            const double inv = 1.0 / piv;
            for (int i = p + 1 + e0; i < nf; i += TE) {
                double* rowi = pan + i * TS;
                const double li = rowi[(p - p0) * colstride] * inv;

                rowi[(p - p0) * colstride] = li;
                for (int j = p + 1; j < pe; ++j) rowi[(j - p0) * colstride] -= li * pan[(p + (j - p0) * nf) * TS];
            }
            __syncthreads();
        }
        if constexpr (GLOBAL_F) {      // rows p0..pe-1 of the panel hold final U entries: back to the front
            for (int t = e0; t < pb * pb; t += TE) {
                const int r = p0 + t % pb, q = t / pb;
                Fl[(r + (p0 + q) * nf) * TS] = Pl[(r + q * nf) * TS];
            }
        }
        if constexpr (GLOBAL_F) {
            for (int j = pe + e0; j <= nf; j += TE) {
                double* colj = Fl + j * colstride;
                double uq[B];
#pragma unroll
                for (int q = 0; q < B; ++q) uq[q] = (q < pb) ? colj[(p0 + q) * TS] : 0.0;
#pragma unroll
                for (int q = 0; q < B; ++q) {
                    if (q < pb) {
                        const double* lq = pan + q * colstride;
#pragma unroll
                        for (int r = q + 1; r < B; ++r)
                            if (r < pb) uq[r] -= lq[(p0 + r) * TS] * uq[q];
                    }
                }
#pragma unroll
                for (int q = 0; q < B; ++q) {
                    if (q < pb) {
                        colj[(p0 + q) * TS] = uq[q];
                        Ul[(q + j * B) * TS] = uq[q];
                    }
                }
            }
        } else {
            for (int j = pe + e0; j <= nf; j += TE) {
                double* colj = Fl + j * colstride;
                for (int q = p0; q < pe - 1; ++q) {
                    const double uq = colj[q * TS];
                    const double* lq = Fl + q * colstride;
                    for (int r = q + 1; r < pe; ++r) colj[r * TS] -= lq[r * TS] * uq;
                }
            }
        }
        __syncthreads();
        // two rows per lane (i and i + TR) share every pivot-row read, two columns per step give four independent
        // accumulation chains; the panel width is a compile-time constant of the instantiation (most fronts here have 4-12
        // pivots, so the ragged last panel is the common case and must not pay for eight predicated steps)
        switch (pb) {
            case 1: trailing_update<TS, 1, GLOBAL_F>(Fl, pan, Ul, nf, p0, pe, er, ec, TR, TC); break;
            case 2: trailing_update<TS, 2, GLOBAL_F>(Fl, pan, Ul, nf, p0, pe, er, ec, TR, TC); break;
            case 3: trailing_update<TS, 3, GLOBAL_F>(Fl, pan, Ul, nf, p0, pe, er, ec, TR, TC); break;
            case 4: trailing_update<TS, 4, GLOBAL_F>(Fl, pan, Ul, nf, p0, pe, er, ec, TR, TC); break;
            case 5: trailing_update<TS, 5, GLOBAL_F>(Fl, pan, Ul, nf, p0, pe, er, ec, TR, TC); break;
            case 6: trailing_update<TS, 6, GLOBAL_F>(Fl, pan, Ul, nf, p0, pe, er, ec, TR, TC); break;
            case 7: trailing_update<TS, 7, GLOBAL_F>(Fl, pan, Ul, nf, p0, pe, er, ec, TR, TC); break;
            default: trailing_update<TS, 8, GLOBAL_F>(Fl, pan, Ul, nf, p0, pe, er, ec, TR, TC); break;
        }
        __syncthreads();
    }
    if (act && bad && e0 == 0) status[s] = -3;
    if constexpr (TS >= 2 && !GLOBAL_F) {
        const int Wu2 = (fd.flags >> 8) & 0xff, Wo2 = fd.wout;
        if (Wu2 >= 2 && Wo2 >= 2) {
            // pair mapping: scenarios s2, s2 + 1 (s2 even) are adjacent in every section of width >= 2. Rows and blocks of
            // scenarios that are not active are written too (nobody reads them); flags are per scenario.
            const int s2 = blockIdx.y * TS + 2 * sl2;
            const bool act0 = active ? active[s2] != 0 : true, act1 = active ? active[s2 + 1] != 0 : true;
            const int TRv = 2 * TR, TCv = TEv / TRv;
            const int erv = ev % TRv, ecv = ev / TRv;
            const double2* F2 = reinterpret_cast<const double2*>(Fs) + sl2;
            double2* __restrict__ Uf2 = reinterpret_cast<double2*>(u_base(U, sy, Wu2, s2) + fd.uoff * Wu2);
            const int hu = Wu2 >> 1, ho = Wo2 >> 1;
            bool w0 = false, w1 = false;
            for (int p = ecv; p < k; p += TCv) {
                double2* Urow = Uf2 + urow_off(p, nf) * hu;
                const double2 dp = F2[(p + p * nf) * H];
                const double lim0 = sy.growth * fabs(dp.x), lim1 = sy.growth * fabs(dp.y);
                for (int j = p + erv; j <= nf; j += TRv) {
                    double2 v = F2[(p + j * nf) * H];
                    if (j < nf) { w0 = w0 || fabs(v.x) > lim0; w1 = w1 || fabs(v.y) > lim1; }
                    if (j == p) { v.x = 1.0 / v.x; v.y = 1.0 / v.y; }
                    Urow[(unsigned)((j - p) * hu)] = v;
                }
            }
            if (sy.weak) {
                if (w0 && act0) sy.weak[s2] = 1;
                if (w1 && act1) sy.weak[s2 + 1] = 1;
            }
            double2* __restrict__ Cf2 = reinterpret_cast<double2*>(upd_base(upd, sy, Wo2, s2) + fd.updoff * Wo2);
            for (int j = ecv; j <= u; j += TCv) {
                const double2* colj = F2 + ((k + j) * nf + k) * H;
                double2* Cj = Cf2 + (unsigned)(j * u * ho);
                for (int i = erv; i < u; i += TRv) Cj[(unsigned)(i * ho)] = colj[i * H];
            }
            return;
        }
    }
    if (!act) return;
    const int Wu = (fd.flags >> 8) & 0xff;      // tile width of the back-solve launch that reads these rows
    double* __restrict__ Uf = u_base(U, sy, Wu, s) + fd.uoff * Wu;
    // pivot guard, evaluated where the U rows are written out (off the barrier-separated panel steps): an entry of row p
    // more than `growth` times its pivot — the multiplier of the transposed elimination — marks the scenario
    for (int p = ec; p < k; p += TC) {
        double* Urow = Uf + urow_off(p, nf) * Wu;
        const double lim = sy.growth * fabs(Fl[(p + p * nf) * TS]);
        for (int j = p + er; j <= nf; j += TR) {
            const double v = Fl[(p + j * nf) * TS];
            if (j < nf && fabs(v) > lim) weakp = true;
            Urow[(unsigned)((j - p) * Wu)] = (j == p) ? 1.0 / v : v;
        }
    }
    if (weakp && sy.weak) sy.weak[s] = 1;
    const int Wo = fd.wout;                     // the parent's tile width
    double* __restrict__ Cf = upd_base(upd, sy, Wo, s) + fd.updoff * Wo;
    for (int j = ec; j <= u; j += TC) {
        const double* colj = Fl + ((k + j) * nf + k) * TS;
        double* Cj = Cf + (unsigned)(j * u * Wo);
        for (int i = er; i < u; i += TR) Cj[(unsigned)(i * Wo)] = colj[i * TS];
    }
}

// ---- symmetric (LDL^T) variant for the WLS gain matrix --------------------------------------------------------
// The front holds only its lower triangle, packed by columns (column j starts at j*(2nf-j+1)/2 and has nf-j entries),
// plus the rhs vector and the scaled multipliers of the current panel, so fronts up to order 208 fit in shared memory
// and the trailing update does half the work. Column p keeps the unscaled entries c_i = F[i,p]: they are row p of
// U = D L^T, so the packed-U output and the back-solve kernels are shared with the LU path. The update block is written
// in full (upper part mirrored) so any kind of parent kernel can consume it.
constexpr int kMaxSymFront = 208;

__device__ __forceinline__ int sym_col(int j, int nf) { return (j * (2 * nf - j + 1)) >> 1; }

// Trailing update of the packed lower triangle for one panel of PB pivots (see mf_factor_sym_kernel)
template <int TS, int PB>
__device__ __forceinline__ void sym_trailing_update(double* Fl, const double* Lp, double* Rl, int nf, int p0, int pe, int er,
                                                    int ec, int TR, int TC) {
    const double* cb[PB];
#pragma unroll
    for (int q = 0; q < PB; ++q) cb[q] = Fl + (sym_col(p0 + q, nf) - (p0 + q)) * TS;
    const int nt = nf - pe;                 // trailing rows pe .. nf-1
    for (int t = er; 2 * t < nt; t += TR) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int i = half == 0 ? pe + t : nf - 1 - t;
            if (half == 1 && i == pe + t) break;
            double l[PB];
#pragma unroll
            for (int q = 0; q < PB; ++q) l[q] = Lp[(i + q * nf) * TS];
            int j = pe + ec;
            for (; j + TC <= i; j += 2 * TC) {          // two columns per step: two independent accumulation chains
                double* da = Fl + (sym_col(j, nf) + i - j) * TS;
                double* db = Fl + (sym_col(j + TC, nf) + i - j - TC) * TS;
                double a = *da, b = *db;
#pragma unroll
                for (int q = 0; q < PB; ++q) {
                    a -= l[q] * cb[q][j * TS];
                    b -= l[q] * cb[q][(j + TC) * TS];
                }
                *da = a;
                *db = b;
            }
            if (j <= i) {
                double* da = Fl + (sym_col(j, nf) + i - j) * TS;
                double a = *da;
#pragma unroll
                for (int q = 0; q < PB; ++q) a -= l[q] * cb[q][j * TS];
                *da = a;
            }
            if (ec == 0) {
                double acc = Rl[i * TS];
#pragma unroll
                for (int q = 0; q < PB; ++q) acc -= l[q] * Rl[(p0 + q) * TS];
                Rl[i * TS] = acc;
            }
        }
    }
}

template <int TS>
__global__ void __launch_bounds__(TS == 1 ? 1024 : 256)
mf_factor_sym_kernel(DevSym sy, const FrontDesc* __restrict__ descs, const double* __restrict__ aval,
                     const double* __restrict__ rhs, double* __restrict__ U, double* __restrict__ upd, int S, int TR,
                     const unsigned char* __restrict__ active, int* __restrict__ status) {
    extern __shared__ double Fs[];
    const int sl = threadIdx.x % TS;
    double* Fl = Fs + sl;
    const FrontDesc fd = descs[blockIdx.x];
    const int e0 = threadIdx.x / TS;
    const int TE = blockDim.x / TS;
    const int er = e0 % TR, ec = e0 / TR, TC = TE / TR;
    const int s = blockIdx.y * TS + sl;
    const bool act = active ? (active[s] != 0) : true;
    if (!__syncthreads_or(act)) return;
    const int nf = fd.nf, k = fd.k, u = nf - k;
    const int* __restrict__ rows = sy.f_rows + fd.rowptr;
    const int tri = (nf * (nf + 1)) >> 1;
    double* Rl = Fl + tri * TS;                 // rhs: element i at Rl[i * TS]
    double* Lp = Fl + (tri + nf) * TS;          // panel multipliers: element (i, q) at Lp[(i + q*nf) * TS]
    constexpr int B = 8;
    for (int pos = e0; pos < tri + nf; pos += TE) Fl[pos * TS] = 0.0;
    __syncthreads();
    if (act) {
        const double* __restrict__ av = aval + s;
        const int a1 = fd.asm1;
        for (int a = fd.asm0 + e0; a < a1; a += TE) {
            const int dst = sy.asm_dst[a];
            const int c = dst / nf, r = dst - c * nf;
            if (r >= c) Fl[(sym_col(c, nf) + r - c) * TS] = av[wide(sy.asm_src[a], S)];
        }
        for (int p = e0; p < k; p += TE) Rl[p * TS] = rhs[wide(rows[p], S) + s];
    }
    __syncthreads();
    constexpr int W = TS;
    double* __restrict__ up = upd_base(upd, sy, TS, s);
    // gather of the children's blocks in rounds over the symmetric lists: lower triangle + rhs only, packed destinations
    // (the rhs vector follows the triangle, so one index addresses both)
    for (int rd = fd.ea0; rd < fd.ea1; ++rd) {
        const int t1 = sy.ea_roundptr_s[rd + 1];
        if (act) {
#pragma unroll 4
            for (int t = sy.ea_roundptr_s[rd] + e0; t < t1; t += TE) {
                const int2 pr = sy.ea_pair_s[t];
                Fl[pr.x * TS] += up[(unsigned)(pr.y * W)];
            }
        }
        __syncthreads();
    }
    bool bad = false;
    for (int p0 = 0; p0 < k; p0 += B) {
        const int pe = (p0 + B < k) ? p0 + B : k;
        const int pb = pe - p0;
        for (int p = p0; p < pe; ++p) {
            const double* colp = Fl + (sym_col(p, nf) - p) * TS;      // element (i, p) at colp[i * TS]
            const double piv = colp[p * TS];
            if (piv == 0.0 || !isfinite(piv)) bad = true;
            const double inv = 1.0 / piv;
            for (int i = p + 1 + e0; i < nf; i += TE) {
                const double li = colp[i * TS] * inv;
                Lp[(i + (p - p0) * nf) * TS] = li;
                const int jend = (i + 1 < pe) ? i + 1 : pe;
                for (int j = p + 1; j < jend; ++j) Fl[(sym_col(j, nf) + i - j) * TS] -= li * colp[j * TS];
            }
            __syncthreads();
        }
        if (e0 == 0) {      // forward substitution of the rhs inside the block (tiny, sequential)
            for (int q = p0; q < pe - 1; ++q) {
                const double yq = Rl[q * TS];
                for (int r = q + 1; r < pe; ++r) Rl[r * TS] -= Lp[(r + (q - p0) * nf) * TS] * yq;
            }
        }
        __syncthreads();
        // trailing update of the lower triangle, rows folded in pairs (t-th from the top with t-th from the bottom)
        // so that every lane sweeps the same number of columns; panel width as a compile-time constant (ragged last
        // panels are the common case: most fronts have fewer than 16 pivots)
        switch (pb) {
            case 1: sym_trailing_update<TS, 1>(Fl, Lp, Rl, nf, p0, pe, er, ec, TR, TC); break;
            case 2: sym_trailing_update<TS, 2>(Fl, Lp, Rl, nf, p0, pe, er, ec, TR, TC); break;
            case 3: sym_trailing_update<TS, 3>(Fl, Lp, Rl, nf, p0, pe, er, ec, TR, TC); break;
            case 4: sym_trailing_update<TS, 4>(Fl, Lp, Rl, nf, p0, pe, er, ec, TR, TC); break;
            case 5: sym_trailing_update<TS, 5>(Fl, Lp, Rl, nf, p0, pe, er, ec, TR, TC); break;
            case 6: sym_trailing_update<TS, 6>(Fl, Lp, Rl, nf, p0, pe, er, ec, TR, TC); break;
            case 7: sym_trailing_update<TS, 7>(Fl, Lp, Rl, nf, p0, pe, er, ec, TR, TC); break;
            default: sym_trailing_update<TS, 8>(Fl, Lp, Rl, nf, p0, pe, er, ec, TR, TC); break;
        }
        __syncthreads();
    }
    if (!act) return;
    if (bad && e0 == 0) status[s] = -3;
    const int Wu = (fd.flags >> 8) & 0xff;
    double* __restrict__ Uf = u_base(U, sy, Wu, s) + fd.uoff * Wu;
    for (int p = ec; p < k; p += TC) {
        double* Urow = Uf + urow_off(p, nf) * Wu;
        const double* colp = Fl + (sym_col(p, nf) - p) * TS;
        for (int j = p + er; j <= nf; j += TR) {
            const double v = (j < nf) ? colp[j * TS] : Rl[p * TS];
            Urow[(unsigned)((j - p) * Wu)] = (j == p) ? 1.0 / v : v;
        }
    }
    const int Wo = fd.wout;
    double* __restrict__ Cf = upd_base(upd, sy, Wo, s) + fd.updoff * Wo;
    const bool lower_only = fd.flags & 1;       // the parent is an LDL^T front too: it never reads above the diagonal
    for (int j = ec; j <= u; j += TC) {
        double* Cj = Cf + (unsigned)(j * u * Wo);
        for (int i = (lower_only && j < u) ? j + ((er - j % TR + TR) % TR) : er; i < u; i += TR) {
            double v;
            if (j == u) v = Rl[(k + i) * TS];
            else {
                const int r = i >= j ? i : j, c = i >= j ? j : i;
                v = Fl[(sym_col(k + c, nf) + r - c) * TS];
            }
            Cj[(unsigned)(i * Wo)] = v;
        }
    }
}

void launch_factor_sym(int ts, dim3 grid, int threads, size_t smem, cudaStream_t st, DevSym dev,
                       const FrontDesc* fronts, const double* aval, const double* rhs, double* U, double* upd, int S, int tr,
                       const unsigned char* active, int* status) {
#define JGB_CASE(T)                                                                                                \
    case T:                                                                                                        \
        mf_factor_sym_kernel<T><<<grid, threads, smem, st>>>(dev, fronts, aval, rhs, U, upd, S, tr, active, status); \
        break;
    switch (ts) {
        JGB_CASE(1) JGB_CASE(2) JGB_CASE(4) JGB_CASE(8) JGB_CASE(16) JGB_CASE(32)
        default: throw std::runtime_error("unsupported scenario tile");
    }
#undef JGB_CASE
}

// ---- TMA-staged variant for the many small fronts of a batch -------------------------------------------------
constexpr int kBulkMaxChildren = 32;   // children staged per group

// One CTA = one front x one tile of 32 scenarios (lane = scenario), TE warps split the entries. The update blocks
// of the children are contiguous runs in the tile-major update storage, so one elected thread fetches a whole group
// of them with cp.async.bulk (TMA) into a staging area behind the front while the other threads zero the front and
// assemble the matrix entries; completion is signalled through an mbarrier. All later traffic is coalesced
// 256-byte lines. Shared memory: [front nf*(nf+1)*32 doubles][staging cap*32 doubles][rel list cap ints].
template <int TE, int MAXNF>
__global__ void __launch_bounds__(32 * TE)
mf_factor_bulk_kernel(DevSym sy, const FrontDesc* __restrict__ descs, const double* __restrict__ aval,
                      const double* __restrict__ rhs, double* __restrict__ U, double* __restrict__ upd, int S,
                      int smem_elems, const unsigned char* __restrict__ active, int* __restrict__ status, StagedEa sg) {
    extern __shared__ __align__(128) double sm[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ __align__(8) uint64_t ea_bar[2];
    __shared__ int2 ea_ch[kEaChunkCache];
    __shared__ int s_off[kBulkMaxChildren], s_uc[kBulkMaxChildren], s_relo[kBulkMaxChildren], s_relp[kBulkMaxChildren], s_gend;
    __shared__ long long s_updoff[kBulkMaxChildren];
    constexpr int TR = TE >= 4 ? 4 : TE, TC = TE / TR;
    const int sl = threadIdx.x & 31, e0 = threadIdx.x >> 5;
    const int er = e0 % TR, ec = e0 / TR;
    const FrontDesc fd = descs[blockIdx.x];          // one 64-byte read: no dependent metadata loads
    const int s = blockIdx.y * 32 + sl;
    const bool act = active ? (active[s] != 0) : true;
    if (!__syncthreads_or(act)) return;
    const int nf = fd.nf, k = fd.k, u = nf - k;
    const int* __restrict__ rows = sy.f_rows + fd.rowptr;
    const int fsz = nf * (nf + 1);
    double* Fl = sm + sl;
    double* stage = sm + fsz * 32;
    const int cap = smem_elems - fsz;
    int* srel = reinterpret_cast<int*>(sm + (size_t)smem_elems * 32);
    double* __restrict__ uptile = upd + sy.sec_base[5] + (long long)blockIdx.y * sy.sec_size[5] * 32;   // section W = 32
    const int c0 = fd.child0, c1 = fd.child1;

    // Ring mode (sg.chunks): the children's blocks stream through a two-stage ring behind the front in chunks of one child
    // (child order), destinations from the list in update-storage order. The staging area no longer has to hold the
    // largest child: 13-16-row fronts go from one CTA per SM (front + 45-60 KB of staging) to two or three.
    const bool ring_mode = sg.chunks != nullptr;
    const int nch = ring_mode ? c1 - c0 : 0;      // child0 / child1 hold the chunk range in ring mode
    if (threadIdx.x == 0 && c1 > c0) {
        if (ring_mode) {
            mbar_init(&ea_bar[0], 1);
            mbar_init(&ea_bar[1], 1);
            for (int c = 0; c < 2 && c < nch; ++c) {
                const int2 cd = sg.chunks[c0 + c];
                mbar_expect_tx(&ea_bar[c], (uint32_t)cd.y * 256u);
                bulk_g2s(stage + (size_t)c * sg.ring_elems * 32, uptile + (size_t)cd.x * 32, (uint32_t)cd.y * 256u, &ea_bar[c]);
            }
        } else {
            mbar_init(&mbar, 1);
        }
    }
    // matrix entries and right-hand side: the first few per lane are fetched into registers before the front is
    // zeroed, so their index -> value load chains overlap the zeroing pass and the barrier instead of following them
    constexpr int NPRE = 6;
    const double* __restrict__ av = aval + s;
    const int a0 = fd.asm0, a1 = fd.asm1;
    int pdst[NPRE];
    double pval[NPRE];
#pragma unroll
    for (int q = 0; q < NPRE; ++q) {
        const int a = a0 + e0 + q * TE;
        pdst[q] = -1;
        pval[q] = 0.0;
        if (a < a1) {
            pdst[q] = sy.asm_dst[a];
            pval[q] = av[wide(sy.asm_src[a], S)];
        }
    }
    constexpr int NPRE_R = 2;
    double prhs[NPRE_R];
#pragma unroll
    for (int q = 0; q < NPRE_R; ++q) {
        const int p = e0 + q * TE;
        prhs[q] = (p < k) ? rhs[wide(rows[p], S) + s] : 0.0;
    }
    if (ring_mode)
        for (int c = threadIdx.x; c < nch && c < kEaChunkCache; c += blockDim.x) ea_ch[c] = sg.chunks[c0 + c];
    for (int pos = e0; pos < fsz; pos += TE) Fl[pos * 32] = 0.0;
    __syncthreads();
    {
#pragma unroll
        for (int q = 0; q < NPRE; ++q)
            if (pdst[q] >= 0) Fl[pdst[q] * 32] = pval[q];
        for (int a = a0 + e0 + NPRE * TE; a < a1; a += TE) Fl[sy.asm_dst[a] * 32] = av[wide(sy.asm_src[a], S)];
#pragma unroll
        for (int q = 0; q < NPRE_R; ++q) {
            const int p = e0 + q * TE;
            if (p < k) Fl[(p + nf * nf) * 32] = prhs[q];
        }
        for (int p = e0 + NPRE_R * TE; p < k; p += TE) Fl[(p + nf * nf) * 32] = rhs[wide(rows[p], S) + s];
    }
    if (ring_mode) {
        __syncthreads();                          // front assembled, barriers initialised
        // destination indices are warp-uniform: lane l fetches the index of element l (and l + 32) of the NEXT chunk while
        // the current one is waited for and added; the adds take them from the lanes with shuffles
        auto chunk_at = [&](int c) { return c < kEaChunkCache ? ea_ch[c] : sg.chunks[c0 + c]; };
        const int* __restrict__ dbase = sg.upd_dst + sg.sec_cum;
        int2 cd = nch > 0 ? chunk_at(0) : make_int2(0, 0);
        int d0 = sl < cd.y ? dbase[cd.x + sl] : 0, d1 = sl + 32 < cd.y ? dbase[cd.x + sl + 32] : 0;
        for (int c = 0; c < nch; ++c) {
            const int2 cn = c + 1 < nch ? chunk_at(c + 1) : make_int2(0, 0);
            const int n0 = sl < cn.y ? dbase[cn.x + sl] : 0, n1 = sl + 32 < cn.y ? dbase[cn.x + sl + 32] : 0;
            mbar_wait(&ea_bar[c & 1], (c >> 1) & 1);
            const double* rg = stage + (size_t)(c & 1) * sg.ring_elems * 32 + sl;
            for (int e = e0; e < cd.y; e += TE) {          // ring stages hold at most 64 elements
                const int d = __shfl_sync(0xffffffffu, e < 32 ? d0 : d1, e & 31);
                Fl[d * 32] += rg[e * 32];
            }
            __syncthreads();
            if (threadIdx.x == 0 && c + 2 < nch) {
                const int2 nd = chunk_at(c + 2);
                mbar_expect_tx(&ea_bar[c & 1], (uint32_t)nd.y * 256u);
                bulk_g2s(stage + (size_t)(c & 1) * sg.ring_elems * 32, uptile + (size_t)nd.x * 32, (uint32_t)nd.y * 256u,
                         &ea_bar[c & 1]);
            }
            cd = cn; d0 = n0; d1 = n1;
        }
    }
    uint32_t parity = 0;
    for (int ci = c0; ci < c1 && !ring_mode;) {
        // child descriptors of up to kBulkMaxChildren children: one parallel 16-byte read each (no dependent chain)
        const int ncand = min(c1 - ci, kBulkMaxChildren);
        if ((int)threadIdx.x < ncand) {
            const ChildDesc cd = sy.child_desc[ci + threadIdx.x];
            s_uc[threadIdx.x] = cd.uc;
            s_relp[threadIdx.x] = cd.relptr;
            s_updoff[threadIdx.x] = cd.updoff;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int off = 0, ro = 0, g = 0;
            while (g < ncand) {
                const int blk = s_uc[g] * (s_uc[g] + 1);
                if (g > 0 && off + blk > cap) break;
                s_off[g] = off; s_relo[g] = ro;
                off += blk; ro += s_uc[g]; ++g;
            }
            s_gend = ci + g;
            mbar_expect_tx(&mbar, (uint32_t)off * 256u);
            for (int q = 0; q < g; ++q)
                bulk_g2s(stage + (size_t)s_off[q] * 32, uptile + s_updoff[q] * 32, (uint32_t)(s_uc[q] * (s_uc[q] + 1)) * 256u,
                         &mbar);
        }
        __syncthreads();
        const int gend = s_gend;
        // relative indices of the group's children into shared memory while the bulk copies are in flight
        for (int q = ci + e0; q < gend; q += TE) {
            const int* __restrict__ rel = sy.f_rel + s_relp[q - ci];
            const int uc = s_uc[q - ci];
            for (int i = sl; i < uc; i += 32) srel[s_relo[q - ci] + i] = rel[i];
        }
        __syncthreads();
        mbar_wait(&mbar, parity);
        parity ^= 1;
        for (int q = 0; q < gend - ci; ++q) {
            const int uc = s_uc[q];
            const int* r = srel + s_relo[q];
            const double* st = stage + (size_t)s_off[q] * 32 + sl;
            for (int j = ec; j <= uc; j += TC) {
                const int dc = (j < uc) ? r[j] : nf;
                double* colj = Fl + dc * nf * 32;
                const double* sj = st + j * uc * 32;
                for (int i = er; i < uc; i += TR) colj[r[i] * 32] += sj[i * 32];
            }
            __syncthreads();
        }
        ci = gend;
    }
    __syncthreads();
    // ---- elimination in registers: warp e0 owns columns e0, e0 + TE, ... of the front (nf + 1 columns incl. rhs).
    // Each pivot column is broadcast through a small double-buffered shared-memory strip; all loops are unrolled to
    // the compile-time bound MAXNF so every register index is static, rows/columns beyond nf are predicated off.
    constexpr int NC = (MAXNF + 1 + TE - 1) / TE;
    double col[NC][MAXNF];
    double* bc = stage;    // staging area is free now: 2 x MAXNF x 32 doubles
#pragma unroll
    for (int q = 0; q < NC; ++q) {
        const int c = e0 + q * TE;
#pragma unroll
        for (int i = 0; i < MAXNF; ++i) col[q][i] = (c <= nf && i < nf) ? Fl[(i + c * nf) * 32] : 0.0;
    }
    bool bad = false, weakp = false;
    const int Wu = (fd.flags >> 8) & 0xff;      // 32 when the register back-solve reads these rows
    double* __restrict__ Uf = u_base(U, sy, Wu, s) + fd.uoff * Wu;
#pragma unroll
    for (int p = 0; p < MAXNF; ++p) {
        if (p >= k) break;
        double* b = bc + (p & 1) * (MAXNF * 32) + sl;
        if (e0 == p % TE) {
#pragma unroll
            for (int i = p; i < MAXNF; ++i)
                if (i < nf) b[i * 32] = col[p / TE][i];
        }
        __syncthreads();

        const double piv = b[p * 32];
        if (piv == 0.0 || !isfinite(piv)) bad = true;
        const double inv = 1.0 / piv;
        double* Urow = Uf + urow_off(p, nf) * Wu;
        double m[NC];                                   // multiplier of each owned column, 0 for columns not updated
#pragma unroll
        for (int q = 0; q < NC; ++q) {
            const int c = e0 + q * TE;
            const double upc = col[q][p];               // U[p, c]
            const bool in = c >= p && c <= nf;
            if (act && in) Urow[(unsigned)((c - p) * Wu)] = (c == p) ? inv : upc;
            m[q] = (in && c > p) ? inv * upc : 0.0;
            if (c < nf && fabs(m[q]) > sy.growth) weakp = true;      // pivot guard: row multiplier U[p,c] / U[p,p]
        }
#pragma unroll
        for (int i = p + 1; i < MAXNF; ++i) {
            const double li = (i < nf) ? b[i * 32] : 0.0;   // pivot-column entry, read once for all owned columns
#pragma unroll
            for (int q = 0; q < NC; ++q) col[q][i] -= li * m[q];
        }
    }
    if (!act) return;
    if (bad && e0 == 0) status[s] = -3;
    if (weakp && sy.weak) sy.weak[s] = 1;
    const int Wo = fd.wout;                     // the parent's tile width (32 when the parent is a small front too)
    double* __restrict__ Cf = upd_base(upd, sy, Wo, s) + fd.updoff * Wo;
#pragma unroll
    for (int q = 0; q < NC; ++q) {
        const int c = e0 + q * TE;
        if (c >= k && c <= nf) {
            double* Cj = Cf + (unsigned)((c - k) * u * Wo);
#pragma unroll
            for (int i = 0; i < MAXNF; ++i)
                if (i >= k && i < nf) Cj[(unsigned)((i - k) * Wo)] = col[q][i];
        }
    }
}

#include "mf_task.cuh"
#include "mf_dense.cuh"

template <bool GLOBAL_F>
void launch_factor(int ts, dim3 grid, int threads, size_t smem, cudaStream_t st, DevSym dev, const int* fronts,
                   const FrontDesc* descs, const double* aval, const double* rhs, double* U, double* upd, int S, int tr,
                   const unsigned char* active, int* status, double* gwork, long long gstride, StagedEa sg = StagedEa{}) {
    // cp.async round 0 of the extend-add: measured on the 10k-bus Jacobian single case 509 -> 491 us per factorisation,
    // batch of 10 016 scenarios 43.6 -> 44.8 ms (the batch gather is not bound by loads in flight): single case only.
    // JGB_ASYNC_EA=0/1 forces it (tuning only).
    static const int ea_env = getenv("JGB_ASYNC_EA") ? atoi(getenv("JGB_ASYNC_EA")) : -1;
    const int ea_async = ea_env >= 0 ? ea_env : (S == 1);
#define JGB_CASE(T)                                                                                                  \
    case T:                                                                                                          \
        mf_factor_kernel<T, GLOBAL_F><<<grid, threads, smem, st>>>(dev, fronts, descs, aval, rhs, U, upd, S, tr,     \
                                                                   active, status, gwork, gstride, ea_async, sg);    \
        break;
    switch (ts) {
        JGB_CASE(1) JGB_CASE(2) JGB_CASE(4) JGB_CASE(8) JGB_CASE(16) JGB_CASE(32)
        default: throw std::runtime_error("unsupported scenario tile");
    }
#undef JGB_CASE
}

// (lanes per scenario, register bound on the front order) variants of the bulk kernel
#define JGB_BULK_VARIANTS(X) X(4, 6) X(4, 8) X(8, 8) X(4, 10) X(8, 10) X(4, 12) X(8, 12) X(8, 16) X(16, 16)

void launch_factor_bulk(int maxnf, int te, dim3 grid, size_t smem, cudaStream_t st, DevSym dev, const FrontDesc* fronts,
                        const double* aval, const double* rhs, double* U, double* upd, int S, int smem_elems,
                        const unsigned char* active, int* status, StagedEa sg) {
#define X(TE, MAXNF)                                                                                              \
    if (maxnf == MAXNF && te == TE) {                                                                             \
        mf_factor_bulk_kernel<TE, MAXNF><<<grid, 32 * TE, smem, st>>>(dev, fronts, aval, rhs, U, upd, S, smem_elems, \
                                                                      active, status, sg);                       \
        return;                                                                                                   \
    }
    JGB_BULK_VARIANTS(X)
#undef X
    throw std::runtime_error("unsupported bulk factor variant");
}

int bulk_variant_for(int nf) {
    // fronts above 16 go to the shared-memory kernel: a 20-row register variant spilled its columns to local memory
    // and was 7 % slower on the factor phase than the blocked kernel with 16-scenario tiles
    static const int bulk_max = getenv("JGB_BULK_MAX") ? atoi(getenv("JGB_BULK_MAX")) : 16;
    if (nf > bulk_max || nf > 16) return 0;
    // JGB_BULK_FINE=0: three register variants (8 / 12 / 16) instead of five (6 / 8 / 10 / 12 / 16)
    static const bool fine = !(getenv("JGB_BULK_FINE") && atoi(getenv("JGB_BULK_FINE")) == 0);
    if (fine) return nf <= 6 ? 6 : nf <= 8 ? 8 : nf <= 10 ? 10 : nf <= 12 ? 12 : 16;
    return nf <= 8 ? 8 : nf <= 12 ? 12 : 16;
}
// warps per CTA of the small-front kernel by register variant; JGB_BULK_LANES="a,b,c" overrides (tuning only)
int bulk_lanes_for(int maxnf) {
    // measured at 10 016 scenarios (factor phase): 4,4,8 38.6 ms; 4,8,8 37.3; 8,8,8 39.8; 4,8,16 42.3
    static int v[3] = {4, 8, 8};
    static bool init = false;
    if (!init) {
        init = true;
        if (const char* e = getenv("JGB_BULK_LANES")) sscanf(e, "%d,%d,%d", &v[0], &v[1], &v[2]);
    }
    return maxnf <= 8 ? v[0] : maxnf <= 12 ? v[1] : v[2];
}

template <int TS>
void set_factor_smem_attr() {
    JGB_CUDA(cudaFuncSetAttribute(mf_factor_kernel<TS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    JGB_CUDA(cudaFuncSetAttribute(mf_factor_kernel<TS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    JGB_CUDA(cudaFuncSetAttribute(mf_factor_sym_kernel<TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
}

// Backward substitution, S == 1: one CTA per front, pivots processed in blocks of 32 rows from the bottom up.
// For each block the packed U rows are staged in shared memory, the part of every row that multiplies already
// known x (later pivots and update rows) is removed by one warp per row, and warp 0 finishes the 32 x 32 triangle.
constexpr int kBsRows = 32;
// The same kernel serves the batch fronts too large for the tile kernel's shared-memory staging (blockIdx.y = scenario,
// element stride S): strided reads, but only the few top-of-tree fronts of very large cases (e.g. 271 rows at 70k buses).
__global__ void __launch_bounds__(512)
mf_backsolve_single(DevSym sy, const int* __restrict__ fronts, const double* __restrict__ U,
                    double* __restrict__ x, const unsigned char* __restrict__ active, int S, int bs_rows,
                    const int* __restrict__ seqptr) {
    extern __shared__ double sh[];
    const int s = blockIdx.y;
    if (active && !active[s]) return;
    // sequence mode (seqptr): the CTA walks a chain of fronts from the last (closest to the root) to the first; x of a
    // front's pivots is in global memory before the barrier that starts its child
    const int d0 = seqptr ? seqptr[blockIdx.x] : blockIdx.x, d1 = seqptr ? seqptr[blockIdx.x + 1] : blockIdx.x + 1;
    for (int di = d1 - 1; di >= d0; --di) {
    if (di != d1 - 1) __syncthreads();
    const int f = fronts[di];
    const int nf = sy.f_nf[f], k = sy.f_k[f];
    const int* __restrict__ rows = sy.f_rows + sy.f_rowptr[f];
    double* xs = sh;              // nf entries: x of the front rows (pivots filled in as they are solved)
    double* Us = sh + nf;         // packed rows of the current block
    const double* __restrict__ Uf = u_base(U, sy, 1, s) + sy.f_uoff[f];      // section W = 1: contiguous per scenario
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int j = k + threadIdx.x; j < nf; j += blockDim.x) xs[j] = x[wide(rows[j], S) + s];
    for (int p1 = k; p1 > 0; p1 -= bs_rows) {       // bs_rows <= 32: one lane of warp 0 per row of the block
        const int p0 = max(0, p1 - bs_rows);
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
    for (int p = threadIdx.x; p < k; p += blockDim.x) x[wide(rows[p], S) + s] = xs[p];
    }
}

// Backward substitution, batch: one CTA per (front, tile of TS scenarios), TE = blockDim / TS lanes per scenario.
// The front's packed U rows and the already known x of its update rows are staged in shared memory with coalesced
// loads; phase A removes the update-row part of every pivot row (rows distributed over the lanes, no reduction),
// phase B solves the k x k triangle column by column with one barrier per pivot.
template <int TS>
__global__ void __launch_bounds__(128)
mf_backsolve_tile_kernel(DevSym sy, const int* __restrict__ fronts, const double* __restrict__ U,
                         double* __restrict__ x, int S, const unsigned char* __restrict__ active) {
    extern __shared__ double sh[];
    const int sl = threadIdx.x % TS, e = threadIdx.x / TS, TE = blockDim.x / TS;
    const int s = blockIdx.y * TS + sl;
    const bool act = active ? (active[s] != 0) : true;
    if (!__syncthreads_or(act)) return;
    const int f = fronts[blockIdx.x];
    const int nf = sy.f_nf[f], k = sy.f_k[f];
    const int* __restrict__ rows = sy.f_rows + sy.f_rowptr[f];
    const int usz = (int)urow_off(k, nf);
    double* xs = sh + sl;                 // xs[j * TS]
    double* Us = sh + nf * TS + sl;       // Us[e * TS], packed rows
    const double* __restrict__ Uf = u_base(U, sy, TS, s) + sy.f_uoff[f] * TS;   // section W = TS: one contiguous run
    for (int q = e; q < usz; q += TE) Us[q * TS] = Uf[(unsigned)(q * TS)];
    for (int j = k + e; j < nf; j += TE) xs[j * TS] = x[wide(rows[j], S) + s];
    __syncthreads();
    for (int p = e; p < k; p += TE) {
        const double* Urow = Us + (int)urow_off(p, nf) * TS;
        double acc = Urow[(nf - p) * TS];
        for (int j = k; j < nf; ++j) acc -= Urow[(j - p) * TS] * xs[j * TS];
        xs[p * TS] = acc;
    }
    __syncthreads();
    for (int p = k - 1; p >= 0; --p) {
        if (e == p % TE) xs[p * TS] *= Us[(int)urow_off(p, nf) * TS];
        __syncthreads();
        const double xp = xs[p * TS];
        for (int q = e; q < p; q += TE) xs[q * TS] -= Us[((int)urow_off(q, nf) + (p - q)) * TS] * xp;
    }
    __syncthreads();
    if (act)
        for (int p = e; p < k; p += TE) x[wide(rows[p], S) + s] = xs[p * TS];
}

// Backward substitution for the small and mid fronts of a batch (nf <= MAXNF <= 32): one thread = one scenario of one
// front, everything in registers — no shared memory, no barriers. A warp's lanes are 32 consecutive scenarios, so the
// packed U rows and x are read as 256-byte lines, and all loads of a front are independent (issued back to back).
// Loops are unrolled to the compile-time bound so that x stays in registers; rows beyond nf are predicated off.
template <int MAXNF>
__global__ void __launch_bounds__(128)
mf_backsolve_reg_kernel(DevSym sy, const int* __restrict__ fronts, int count, const double* __restrict__ U,
                        double* __restrict__ x, int S, const unsigned char* __restrict__ active) {
    const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (item >= count) return;
    const int s = blockIdx.y * 32 + (threadIdx.x & 31);
    if (active && !active[s]) return;
    const int f = fronts[item];
    const int nf = sy.f_nf[f], k = sy.f_k[f];
    const int* __restrict__ rows = sy.f_rows + sy.f_rowptr[f];
    const double* __restrict__ Uf = u_base(U, sy, 32, s) + sy.f_uoff[f] * 32;    // section W = 32
    double xv[MAXNF];
#pragma unroll
    for (int j = 0; j < MAXNF; ++j) xv[j] = (j >= k && j < nf) ? x[wide(rows[j], S) + s] : 0.0;
#pragma unroll
    for (int p = MAXNF - 1; p >= 0; --p) {
        if (p < k) {
            const double* __restrict__ Urow = Uf + urow_off(p, nf) * 32;
            double acc = Urow[(nf - p) * 32];
#pragma unroll
            for (int j = p + 1; j < MAXNF; ++j)
                if (j < nf) acc -= Urow[(j - p) * 32] * xv[j];
            xv[p] = acc * Urow[0];
        }
    }
#pragma unroll
    for (int p = 0; p < MAXNF; ++p)
        if (p < k) x[wide(rows[p], S) + s] = xv[p];
}

void launch_backsolve_reg(int maxnf, int count, int S, cudaStream_t st, DevSym dev, const int* fronts, const double* U,
                          double* x, const unsigned char* active) {
    const dim3 grid((count + 3) / 4, S / 32);
    if (maxnf <= 8) mf_backsolve_reg_kernel<8><<<grid, 128, 0, st>>>(dev, fronts, count, U, x, S, active);
    else if (maxnf <= 12) mf_backsolve_reg_kernel<12><<<grid, 128, 0, st>>>(dev, fronts, count, U, x, S, active);
    else if (maxnf <= 16) mf_backsolve_reg_kernel<16><<<grid, 128, 0, st>>>(dev, fronts, count, U, x, S, active);
    else if (maxnf <= 24) mf_backsolve_reg_kernel<24><<<grid, 128, 0, st>>>(dev, fronts, count, U, x, S, active);
    else mf_backsolve_reg_kernel<32><<<grid, 128, 0, st>>>(dev, fronts, count, U, x, S, active);
}

void launch_backsolve_tile(int ts, dim3 grid, size_t smem, cudaStream_t st, DevSym dev, const int* fronts,
                           const double* U, double* x, int S, const unsigned char* active) {
#define JGB_CASE(T)                                                                                   \
    case T:                                                                                           \
        mf_backsolve_tile_kernel<T><<<grid, 128, smem, st>>>(dev, fronts, U, x, S, active);           \
        break;
    switch (ts) {
        JGB_CASE(1) JGB_CASE(2) JGB_CASE(4) JGB_CASE(8) JGB_CASE(16) JGB_CASE(32)
        default: throw std::runtime_error("unsupported scenario tile");
    }
#undef JGB_CASE
}

// ---- one factorisation, many right-hand sides (symmetric matrices) -------------------------------------------
// The linear estimators and the DC power flow factor a constant matrix once and solve for a block of right-hand
// sides (`solution!` per Monte-Carlo draw in the reference). The block B is right-hand-side minor: entry i of
// column r at B[i * R + r], so a warp's 32 lanes are 32 right-hand sides and every access is a 256-byte line.
// For a symmetric matrix L = (D^-1 U)^T, so the packed U rows written by the factor kernels (S == 1 layout) are all
// that is needed: row p holds 1/d_p, then U[p, j] = d_p L[j, p] for the later front rows.
//
// Forward substitution, one launch per level (leaves first). CTA = one front x 32 right-hand sides, TE = blockDim / 32
// lanes per right-hand side split the front rows. A front first gathers its children's contributions (each child
// wrote its update-row part into its own slot of C, so the sums are ordered and deterministic), eliminates its
// pivots (one barrier per pivot), overwrites B[pivot rows] with y and leaves its own contribution in C.
__global__ void __launch_bounds__(256)
mf_fwd_multi_kernel(DevSym sy, const int* __restrict__ fronts, const double* __restrict__ U, double* __restrict__ B,
                    double* __restrict__ C, const long long* __restrict__ coff, int R) {
    extern __shared__ double sh[];
    const int sl = threadIdx.x & 31, e = threadIdx.x >> 5, TE = blockDim.x >> 5;
    const int r = blockIdx.y * 32 + sl;
    const int f = fronts[blockIdx.x];
    const int nf = sy.f_nf[f], k = sy.f_k[f], u = nf - k;
    const int* __restrict__ rows = sy.f_rows + sy.f_rowptr[f];
    double* t = sh + sl;                                    // t[i * 32]
    for (int i = e; i < nf; i += TE) t[i * 32] = (i < k) ? B[(long long)rows[i] * R + r] : 0.0;
    __syncthreads();
    for (int ci = sy.f_childptr[f]; ci < sy.f_childptr[f + 1]; ++ci) {
        const int c = sy.f_children[ci];
        const int uc = sy.f_nf[c] - sy.f_k[c];
        const int* __restrict__ rel = sy.f_rel + sy.f_relptr[c];
        const double* __restrict__ Cc = C + coff[c] * R + r;
        for (int j = e; j < uc; j += TE) t[rel[j] * 32] += Cc[(long long)j * R];
        __syncthreads();
    }
    const double* __restrict__ Uf = U + sy.f_uoff[f];
    for (int p = 0; p < k; ++p) {
        const double* __restrict__ Urow = Uf + urow_off(p, nf);
        const double y = t[p * 32] * Urow[0];              // y_p / d_p: multiplier of the packed row
        for (int j = p + 1 + e; j < nf; j += TE) t[j * 32] -= Urow[j - p] * y;
        __syncthreads();
    }
    for (int i = e; i < k; i += TE) B[(long long)rows[i] * R + r] = t[i * 32];
    double* __restrict__ Cf = C + coff[f] * R + r;
    for (int j = e; j < u; j += TE) Cf[(long long)j * R] = t[(k + j) * 32];
}

// Backward substitution, one launch per depth level (roots first): x_p = (y_p - sum_{j>p} U[p,j] x_j) / d_p.
__global__ void __launch_bounds__(256)
mf_bwd_multi_kernel(DevSym sy, const int* __restrict__ fronts, const double* __restrict__ U, double* __restrict__ B,
                    int R) {
    extern __shared__ double sh[];
    const int sl = threadIdx.x & 31, e = threadIdx.x >> 5, TE = blockDim.x >> 5;
    const int r = blockIdx.y * 32 + sl;
    const int f = fronts[blockIdx.x];
    const int nf = sy.f_nf[f], k = sy.f_k[f];
    const int* __restrict__ rows = sy.f_rows + sy.f_rowptr[f];
    double* t = sh + sl;
    for (int i = e; i < nf; i += TE) t[i * 32] = B[(long long)rows[i] * R + r];
    __syncthreads();
    const double* __restrict__ Uf = U + sy.f_uoff[f];
    for (int p = e; p < k; p += TE) {                       // part of every pivot row that multiplies known x
        const double* __restrict__ Urow = Uf + urow_off(p, nf);
        double acc = t[p * 32];
        for (int j = k; j < nf; ++j) acc -= Urow[j - p] * t[j * 32];
        t[p * 32] = acc;
    }
    __syncthreads();
    for (int p = k - 1; p >= 0; --p) {
        if (e == p % TE) t[p * 32] *= Uf[urow_off(p, nf)];
        __syncthreads();
        const double xp = t[p * 32];
        for (int q = e; q < p; q += TE) t[q * 32] -= Uf[urow_off(q, nf) + (p - q)] * xp;
    }
    __syncthreads();
    for (int p = e; p < k; p += TE) B[(long long)rows[p] * R + r] = t[p * 32];
}

// ---- sparse selected inverse on the elimination tree (symmetric matrices, S == 1) --------------------------------
// Takahashi recurrences per front, roots first: with the front's index set I = [pivots | update rows], Z = A^-1 on
// I x I is obtained from the parent's block (the update rows are a subset of the parent's index set) and, for the
// pivots p = k-1 .. 0,   Z[j,p] = -sum_{l>p} Z[j,l] L[l,p]  (j > p),   Z[p,p] = 1/d_p - sum_{l>p} Z[p,l] L[l,p],
// with L[l,p] = U[p,l] / d_p read from the packed U rows. Every front keeps its full nf x nf block (column major) in Z
// at zoff[f]; the entries of A^-1 on the pattern of L + L' are exactly the union of these blocks. One CTA per front,
// the block lives in global memory (L2) because the largest fronts exceed shared memory; this is a one-off post-step
// (largest normalised residual, badData.jl:181-285), not part of the iteration loop.
__global__ void __launch_bounds__(1024)
mf_selinv_kernel(DevSym sy, const int* __restrict__ fronts, const int* __restrict__ parent,
                 const double* __restrict__ U, double* __restrict__ Z, const long long* __restrict__ zoff) {
    extern __shared__ double sh[];          // lvec[nf] | zcol[nf]
    __shared__ double s_diag;
    const int f = fronts[blockIdx.x];
    const int nf = sy.f_nf[f], k = sy.f_k[f], u = nf - k;
    double* lvec = sh;
    double* zcol = sh + nf;
    double* __restrict__ Zf = Z + zoff[f];
    const int pf = parent[f];
    if (pf >= 0) {
        const int nfp = sy.f_nf[pf];
        const double* __restrict__ Zp = Z + zoff[pf];
        const int* __restrict__ rel = sy.f_rel + sy.f_relptr[f];
        for (int t = threadIdx.x; t < u * u; t += blockDim.x) {
            const int a = t % u, b = t / u;
            Zf[(k + a) + (long long)(k + b) * nf] = Zp[rel[a] + (long long)rel[b] * nfp];
        }
    }
    __syncthreads();
    const double* __restrict__ Uf = U + sy.f_uoff[f];
    for (int p = k - 1; p >= 0; --p) {
        const double* __restrict__ Urow = Uf + urow_off(p, nf);
        const double inv = Urow[0];
        for (int l = p + 1 + threadIdx.x; l < nf; l += blockDim.x) lvec[l] = Urow[l - p] * inv;
        __syncthreads();
        for (int j = p + 1 + threadIdx.x; j < nf; j += blockDim.x) {
            double acc = 0.0;
            for (int l = p + 1; l < nf; ++l) acc += Zf[j + (long long)l * nf] * lvec[l];
            zcol[j] = -acc;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double acc = inv;
            for (int l = p + 1; l < nf; ++l) acc -= zcol[l] * lvec[l];
            s_diag = acc;
        }
        for (int j = p + 1 + threadIdx.x; j < nf; j += blockDim.x) {
            Zf[j + (long long)p * nf] = zcol[j];
            Zf[p + (long long)j * nf] = zcol[j];
        }
        __syncthreads();
        if (threadIdx.x == 0) Zf[p + (long long)p * nf] = s_diag;
        __syncthreads();
    }
}

int selinv_threads(int max_nf) {      // one CTA per front; JGB_SELINV_THREADS overrides (tuning only)
    static const int v = getenv("JGB_SELINV_THREADS") ? atoi(getenv("JGB_SELINV_THREADS")) : 0;
    if (v > 0) return std::min(1024, std::max(32, v));
    return max_nf <= 64 ? 256 : 1024;
}

int backsolve_single_threads() {
    // one CTA per front: measured single-case back-solve of the 10k-bus Jacobian 293 / 248 / 239 us with 128 / 256 / 512
    // threads (gain matrix: 0.83 / 0.62 / 0.53 ms). JGB_BS_THREADS overrides (<= 512; tuning only).
    static const int v = getenv("JGB_BS_THREADS") ? std::min(512, std::max(32, atoi(getenv("JGB_BS_THREADS")))) : 512;
    return v;
}

int backsolve_reg_max() {
    // largest front order of the register back-solve. Measured at 10 016 scenarios: off 9.35 ms per back-solve, up to
    // 16 rows 7.30 ms, up to 32 rows 7.89 ms (the 24- and 32-row variants lose to the shared-memory tile kernel).
    // JGB_BSREG_MAX overrides (0 = off; tuning only).
    static const int v = getenv("JGB_BSREG_MAX") ? atoi(getenv("JGB_BSREG_MAX")) : 16;
    return v > 32 ? 32 : v;
}

constexpr int kMaxSmemFront = 150;    // nf*(nf+1)*8 bytes must fit the 200 KB dynamic shared-memory budget

int pow2_floor(int v) {
    int p = 1;
    while (p * 2 <= v) p *= 2;
    return p;
}

}  // namespace

void MfSolver::setup(const Symbolic& s, cudaStream_t st, bool symmetric_matrix) {
    sym = s;
    symmetric = symmetric_matrix;
    d_coff.release();
    csum = 0;
    d_zoff.release();
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
    d_f_eaptr.upload(sym.f_eaptr, st);
    d_ea_roundptr.upload(sym.ea_roundptr, st);
    d_ea_pair.upload(sym.ea_pair, st);
    if (symmetric) {
        d_ea_roundptr_s.upload(sym.ea_roundptr_sym, st);
        d_ea_pair_s.upload(sym.ea_pair_sym, st);
    }
    d_level_fronts.upload(sym.level_fronts, st);
    {
        std::vector<ChildDesc> cds(sym.f_children.size());
        for (size_t q = 0; q < cds.size(); ++q) {
            const int c = sym.f_children[q];
            cds[q].uc = sym.f_nf[c] - sym.f_k[c];
            cds[q].relptr = sym.f_relptr[c];
            cds[q].updoff = sym.f_updoff[c];
        }
        d_child_desc.upload(cds, st);
        JGB_CUDA(cudaStreamSynchronize(st));
    }
    d_depth_fronts.upload(sym.depth_fronts, st);
    std::vector<long long> uo(sym.f_uoff.begin(), sym.f_uoff.end()), po(sym.f_updoff.begin(), sym.f_updoff.end());
    d_f_uoff.upload(uo, st);
    d_f_updoff.upload(po, st);
    JGB_CUDA(cudaStreamSynchronize(st));   // the host vectors above go out of scope
    dev.f_k = d_f_k.p; dev.f_nf = d_f_nf.p; dev.f_rowptr = d_f_rowptr.p; dev.f_rows = d_f_rows.p;
    dev.f_relptr = d_f_relptr.p; dev.f_rel = d_f_rel.p; dev.f_childptr = d_f_childptr.p;
    dev.f_children = d_f_children.p; dev.f_asmptr = d_f_asmptr.p; dev.asm_src = d_asm_src.p;
    dev.asm_dst = d_asm_dst.p; dev.f_uoff = d_f_uoff.p; dev.f_updoff = d_f_updoff.p;
    dev.f_eaptr = d_f_eaptr.p; dev.ea_roundptr = d_ea_roundptr.p;
    dev.ea_pair = reinterpret_cast<const int2*>(d_ea_pair.p);
    dev.ea_roundptr_s = d_ea_roundptr_s.p;
    dev.ea_pair_s = reinterpret_cast<const int2*>(d_ea_pair_s.p);
    planned_S = -1;
    set_factor_smem_attr<1>(); set_factor_smem_attr<2>(); set_factor_smem_attr<4>();
    set_factor_smem_attr<8>(); set_factor_smem_attr<16>(); set_factor_smem_attr<32>();
    JGB_CUDA(cudaFuncSetAttribute(mf_backsolve_tile_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    JGB_CUDA(cudaFuncSetAttribute(mf_backsolve_tile_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    JGB_CUDA(cudaFuncSetAttribute(mf_backsolve_tile_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    JGB_CUDA(cudaFuncSetAttribute(mf_backsolve_tile_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    JGB_CUDA(cudaFuncSetAttribute(mf_backsolve_tile_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    JGB_CUDA(cudaFuncSetAttribute(mf_backsolve_tile_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
#define X(TE, MAXNF) \
    JGB_CUDA(cudaFuncSetAttribute(mf_factor_bulk_kernel<TE, MAXNF>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    JGB_BULK_VARIANTS(X)
#undef X
    set_task_smem_attr();
    JGB_CUDA(cudaFuncSetAttribute(mf_factor_dense_sym_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    JGB_CUDA(cudaFuncSetAttribute(mf_factor_dense_lu_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    dev.upd_size = sym.upd_size;
    dev.child_desc = d_child_desc.p;
    for (int q = 0; q < 6; ++q) {
        dev.sec_base[q] = 0; dev.sec_size[q] = sym.upd_size;
        dev.usec_base[q] = 0; dev.usec_size[q] = sym.u_size;
    }
    dev.weak = nullptr;
    dev.growth = 1e300;
    JGB_CUDA(cudaFuncSetAttribute(mf_backsolve_single, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
}

namespace {
struct PlanRule { int maxnf, ts, threads; };

// "maxnf:ts:threads,..." (ascending maxnf); environment overrides are for tuning experiments only
std::vector<PlanRule> parse_rules(const char* env, const std::vector<PlanRule>& dflt) {
    const char* v = env ? getenv(env) : nullptr;
    if (!v || !*v) return dflt;
    std::vector<PlanRule> out;
    std::string str(v);
    size_t pos = 0;
    while (pos < str.size()) {
        size_t end = str.find(',', pos);
        if (end == std::string::npos) end = str.size();
        PlanRule r{0, 1, 32};
        if (sscanf(str.substr(pos, end - pos).c_str(), "%d:%d:%d", &r.maxnf, &r.ts, &r.threads) >= 2) out.push_back(r);
        pos = end + 1;
    }
    return out.empty() ? dflt : out;
}
}  // namespace

void MfSolver::build_tasks(int S, cudaStream_t) {
    tplan.clear();
    task_plan = TaskPlan();
    in_task.assign(sym.nfronts, 0);
    task_fronts = task_count = 0;
    task_upd_on_chip = 0;
    if (S < 32 || S % 32 != 0) return;
    partition_tasks(sym, task_options_from_env(), task_plan);
    if (task_plan.launches.empty()) return;
    tplan = task_plan.launches;
    in_task = task_plan.in_task;
    task_fronts = task_plan.task_fronts;
    task_count = task_plan.task_count;
    task_upd_on_chip = task_plan.upd_on_chip;
    for (TaskLaunch& tl : tplan)
        if (tl.smem > 220 * 1024) throw std::runtime_error("task kernel: shared-memory budget exceeded");
    // the blobs are uploaded by plan() once the update-storage layout is fixed (offsets patched in)
}

void MfSolver::plan(int S) {
    if (S == planned_S) return;
    fplan.clear();
    splan.clear();
    size_t gwork_need = 0;
    // factor launch classes: fronts of a level are sorted by decreasing order and cut at these bounds
    // single case: one launch per level (launch latency dominates; measured 625 us vs 855 us per factorisation with
    // four size classes on the 10k-bus Jacobian)
    static const std::vector<PlanRule> single_rules = {{kMaxSmemFront, 1, 256}, {kMaxSymFront, 1, 256}};
    // LDL^T (WLS gain, linear analyses): a front is one CTA on one SM and the big fronts sit alone at the top of the
    // tree, so wider CTAs pay: measured on the 10k-bus gain matrix 4.21 ms per factorisation with 256 threads, 2.74 ms
    // with 512 threads up to 64 rows and 1024 above (the LU kernel of the Jacobian, fronts <= 66, is best at 256)
    static const std::vector<PlanRule> single_rules_sym = {{64, 1, 512}, {kMaxSymFront, 1, 1024}};
    // small fronts in five register variants (6 / 8 / 10 / 12 / 16 rows: loops are unrolled to the variant's bound, so a
    // 5-row front in an 8-row variant pays for three idle rows; factor phase 36.1 -> 35.2 ms against three variants)
    static const std::vector<PlanRule> batch_rules = {{6, 32, 128}, {8, 32, 128}, {10, 32, 128}, {12, 32, 128}, {16, 32, 256}, {20, 16, 256},
                                                      {24, 8, 256}, {32, 8, 256}, {48, 4, 256}, {64, 2, 256},
                                                      {96, 1, 256}, {kMaxSmemFront, 1, 256},
                                                      {kMaxSymFront, 1, 256}};
    // LDL^T batches (WLS Monte-Carlo): the single-scenario classes of the big gain fronts take wider CTAs (factor phase
    // of 512 draws 33.8 -> 29.5 ms); the LU batch of the Jacobian is best with 256 everywhere (48.3 vs 50.2 ms)
    static const std::vector<PlanRule> batch_rules_sym = {{8, 32, 128}, {12, 32, 128}, {16, 32, 256}, {20, 16, 256},
                                                          {24, 8, 256}, {32, 8, 256}, {48, 4, 256}, {64, 2, 256},
                                                          {96, 1, 256}, {kMaxSmemFront, 1, 512},
                                                          {kMaxSymFront, 1, 1024}};
    const char* nb = getenv("JGB_NO_BULK");
    const bool bulk_enabled = !(nb && *nb == '1');
    // JGB_STAGED_EA=0 falls back to the gather in rounds (A/B runs)
    static const bool staged_enabled = !(getenv("JGB_STAGED_EA") && atoi(getenv("JGB_STAGED_EA")) == 0);
    static const bool bulk_ring = !(getenv("JGB_BULK_RING") && atoi(getenv("JGB_BULK_RING")) == 0);
    const std::vector<PlanRule> rules = (S == 1) ? parse_rules("JGB_FPLAN_SINGLE", symmetric ? single_rules_sym : single_rules)
                                                 : parse_rules("JGB_FPLAN_BATCH", symmetric ? batch_rules_sym : batch_rules);
    // LDL^T fronts above dense_min rows take the one-scenario-per-CTA kernel with the FP64 tensor-core trailing update.
    // Measured on the 10k-bus gain matrix (scripts/sweep_dense.sh, factor phase per increment): single case 2.25 ms off,
    // 1.82 from 33 rows, 1.76 from 17 rows; 1000 draws 35.7 ms off, 33.7 from 65 rows, 34.9 from 49, 39.3 from 33 (below
    // ~64 rows the scenario-tile kernels win in batches: they coalesce 2-16 scenarios per 256-byte line).
    // JGB_DENSE_MIN overrides, 0 = off (tuning only)
    static const int dense_env = getenv("JGB_DENSE_MIN") ? atoi(getenv("JGB_DENSE_MIN")) : -1;
    const int dense_min = dense_env == 0 ? (1 << 30) : dense_env > 0 ? dense_env : (S == 1 ? 16 : 64);
    auto dense_ok = [&](int nf) {
        return symmetric && nf > dense_min && nf <= kMaxSymFront && dense_smem_doubles(nf) * sizeof(double) <= 220 * 1024;
    };
    // LU twin (unsymmetric values): single case from 17 rows (the top of the tree is a chain of 45-70-row fronts, one CTA
    // each: panel latency is the Newton step; measured 492 -> 368 us per factorisation with every front on it); batches keep
    // the scenario-tile kernels. JGB_DENSE_LU_MIN = smallest front order that takes it, 0 = off (tuning only)
    static const int dense_lu_env = getenv("JGB_DENSE_LU_MIN") ? atoi(getenv("JGB_DENSE_LU_MIN")) : -1;
    static const int dense_lu_threads = getenv("JGB_DENSE_LU_THREADS") ? atoi(getenv("JGB_DENSE_LU_THREADS")) : 256;
    const int dense_lu_min = dense_lu_env == 0 ? (1 << 30) : dense_lu_env > 0 ? dense_lu_env - 1 : (S == 1 ? 0 : (1 << 30));
    auto dense_lu_ok = [&](int nf) {
        return !symmetric && nf > dense_lu_min && dense_lu_smem_doubles(nf) * sizeof(double) <= 200 * 1024;
    };
    // dense LU launches of a single case gather their children through a ring of two 16 KB stages behind the front
    auto stage_dense_lu = [&](FactorLaunch& fl) {
        if (!staged_enabled || S != 1 || fl.smem + 2 * 16384 > 200 * 1024) return;
        fl.staged = true;
        fl.ring_elems = 2048;
        fl.ring_off = (int)(fl.smem / 8);
        fl.smem += 2 * 16384;
    };
    auto cls0 = [&](int nf) { size_t c = 0; while (c < rules.size() && nf > rules[c].maxnf) ++c; return (int)c; };
    // dense fronts of a level are launched in buckets of similar order, so that the shared-memory footprint (and with
    // it the number of resident CTAs) follows the fronts of the bucket, not the largest front of the level
    // (measured at 1000 draws: one launch per level 32.6 ms per factor phase, buckets of 32 rows 32.0 ms, of 16 rows with
    // 128-thread CTAs 34.3 ms; a single case pays for the extra launches — 1.76 -> 2.42 ms — and keeps one launch per level)
    static const int dense_bucket = getenv("JGB_DENSE_BUCKET") ? std::max(8, atoi(getenv("JGB_DENSE_BUCKET"))) : 32;
    static const int dense_small = getenv("JGB_DENSE_SMALL") ? atoi(getenv("JGB_DENSE_SMALL")) : 0;
    static const int dense_threads_single = getenv("JGB_DENSE_THREADS") ? atoi(getenv("JGB_DENSE_THREADS")) : 0;
    auto cls = [&](int nf) {
        if (dense_lu_ok(nf)) return 2000 + (S == 1 ? 0 : (nf - 1) / 16);
        return dense_ok(nf) ? 1000 + (S == 1 ? 0 : (nf - 1) / dense_bucket) : cls0(nf);
    };
    build_tasks(S, 0);
    // level schedule of the fronts the task launches leave over
    plan_levelptr.assign(1, 0);
    plan_fronts.clear();
    plan_seqptr.clear();
    // Single case, every front on the dense LU kernel: chains at the narrow top of the tree (a front whose largest child
    // sits one level below, on levels of at most seq_width fronts) become sequences that one CTA walks without leaving
    // the kernel, and launches follow dependency slots — slot(sequence) = 1 + the largest slot among the sequences that
    // hold children of its members — instead of tree levels. On the 10k-bus Jacobian the two chains of 11 fronts above
    // level 9 collapse into one launch: 21 factor + 21 back-solve launches become 11 + 11. JGB_SEQ=0 turns it off.
    static const int seq_env = getenv("JGB_SEQ") ? atoi(getenv("JGB_SEQ")) : 0;
    seq_mode = (S == 1) && !symmetric && seq_env > 0 && tplan.empty();
    for (int f = 0; f < sym.nfronts && seq_mode; ++f) seq_mode = dense_lu_ok(sym.f_nf[f]);
    std::vector<int> slot_seqptr;          // per slot: range of sequences (indices into plan_seqptr)
    if (seq_mode) {
        const int F = sym.nfronts;
        std::vector<int> lvl(F, 0), seq_of(F, -1);
        std::vector<std::vector<int>> seqs;
        for (int l = 0; l < sym.nlevels; ++l)
            for (int q = sym.levelptr[l]; q < sym.levelptr[l + 1]; ++q) lvl[sym.level_fronts[q]] = l;
        for (int l = 0; l < sym.nlevels; ++l) {
            const int width = sym.levelptr[l + 1] - sym.levelptr[l];
            for (int q = sym.levelptr[l]; q < sym.levelptr[l + 1]; ++q) {
                const int f = sym.level_fronts[q];
                int best = -1;
                long long bsz = -1;
                for (int ci = sym.f_childptr[f]; ci < sym.f_childptr[f + 1]; ++ci) {
                    const int c = sym.f_children[ci];
                    const long long uc = sym.f_nf[c] - sym.f_k[c];
                    if (uc * (uc + 1) > bsz) { bsz = uc * (uc + 1); best = c; }
                }
                if (best >= 0 && width <= seq_env && lvl[best] == l - 1 && seqs[seq_of[best]].back() == best) {
                    seq_of[f] = seq_of[best];
                    seqs[seq_of[f]].push_back(f);
                } else {
                    seq_of[f] = (int)seqs.size();
                    seqs.push_back(std::vector<int>(1, f));
                }
            }
        }
        std::vector<int> order(seqs.size()), slot(seqs.size(), 0);
        for (size_t q = 0; q < seqs.size(); ++q) order[q] = (int)q;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return lvl[seqs[a].back()] < lvl[seqs[b].back()]; });
        int nslots = 0;
        for (int sq : order) {
            for (int f : seqs[sq])
                for (int ci = sym.f_childptr[f]; ci < sym.f_childptr[f + 1]; ++ci) {
                    const int cs = seq_of[sym.f_children[ci]];
                    if (cs != sq) slot[sq] = std::max(slot[sq], slot[cs] + 1);
                }
            nslots = std::max(nslots, slot[sq] + 1);
        }
        slot_seqptr.assign(1, 0);
        for (int sl = 0; sl < nslots; ++sl) {
            for (int sq : order)
                if (slot[sq] == sl) {
                    plan_seqptr.push_back((int)plan_fronts.size());
                    for (int f : seqs[sq]) plan_fronts.push_back(f);
                }
            plan_levelptr.push_back((int)plan_fronts.size());
            slot_seqptr.push_back((int)plan_seqptr.size());
        }
        plan_seqptr.push_back((int)plan_fronts.size());
        for (int sl = 0; sl < nslots; ++sl) {
            FactorLaunch fl{};
            fl.begin = plan_levelptr[sl];
            fl.count = plan_levelptr[sl + 1] - plan_levelptr[sl];
            int mx = 0;
            for (int q = fl.begin; q < fl.begin + fl.count; ++q) mx = std::max(mx, sym.f_nf[plan_fronts[q]]);
            fl.dense_lu = true;
            fl.ts = 1;
            // a slot of a few chains is latency bound on its one or two CTAs: 512 threads; the wide slots at the bottom 256
            fl.threads = fl.count <= 8 ? 512 : std::min(512, std::max(64, dense_lu_threads));
            fl.smem = dense_lu_smem_doubles(mx) * sizeof(double);
            fl.tr = 1;
            fl.seq_begin = slot_seqptr[sl];
            fl.nseq = slot_seqptr[sl + 1] - slot_seqptr[sl];
            stage_dense_lu(fl);
            fplan.push_back(fl);
        }
    } else {
        for (int l = 0; l < sym.nlevels; ++l) {
            for (int q = sym.levelptr[l]; q < sym.levelptr[l + 1]; ++q)
                if (!in_task[sym.level_fronts[q]]) plan_fronts.push_back(sym.level_fronts[q]);
            plan_levelptr.push_back((int)plan_fronts.size());
        }
    }
    for (int l = 0; l < sym.nlevels && !seq_mode; ++l) {
        int b = plan_levelptr[l], e = plan_levelptr[l + 1];
        int i = b;
        while (i < e) {   // fronts are sorted by decreasing order inside a level
            int c = cls(sym.f_nf[plan_fronts[i]]);
            int j = i;
            while (j < e && cls(sym.f_nf[plan_fronts[j]]) == c) ++j;
            int nf = sym.f_nf[plan_fronts[i]];
            size_t per = (size_t)nf * (nf + 1) * sizeof(double);
            FactorLaunch fl{};
            fl.begin = i;
            fl.count = j - i;
            fl.dense = dense_ok(nf);
            fl.dense_lu = dense_lu_ok(nf);
            if (fl.dense || fl.dense_lu) c = cls0(nf);
            if (fl.dense_lu) {
                fl.ts = 1;
                fl.threads = std::min(512, std::max(64, dense_lu_threads));
                fl.smem = dense_lu_smem_doubles(nf) * sizeof(double);
                fl.tr = 1;
                fl.gstride = 0;
                stage_dense_lu(fl);
                fplan.push_back(fl);
                i = j;
                continue;
            }
            fl.sym = symmetric && nf <= kMaxSymFront;
            fl.global_front = !fl.sym && (nf > kMaxSmemFront || c >= (int)rules.size());
            fl.bulk = false;
            if (fl.dense) {
                fl.ts = 1;
                fl.threads = nf <= dense_small ? 128 : 256;      // one thread per panel row is all step 2 can use
                // a single case: the front is alone on its SM, gather / write-out / tensor tiles scale with the CTA width
                // (measured 1.77 -> 1.38 ms per gain factorisation, 358 -> 428 GN iterations/s; JGB_DENSE_THREADS overrides)
                if (S == 1) fl.threads = dense_threads_single > 0 ? std::min(512, dense_threads_single) : 512;
                fl.smem = dense_smem_doubles(nf) * sizeof(double);
                if (S > 1) {
                    // batches: when shared memory leaves room for one or two CTAs per SM only, wider CTAs keep the SM's warp
                    // slots busy (JGB_DENSE_WIDE="t1,t2": threads at one / two CTAs per SM; tuning only)
                    static int wide1 = 512, wide2 = 384;      // measured per 1000 draws: 256,256 25.43 ms; 512,256 25.07; 512,384 24.49; 512,512 25.94
                    static bool init = false;
                    if (!init) {
                        init = true;
                        if (const char* e = getenv("JGB_DENSE_WIDE")) sscanf(e, "%d,%d", &wide1, &wide2);
                    }
                    const size_t ctas = 233472 / (fl.smem + 1024);
                    if (ctas <= 1) fl.threads = wide1;
                    else if (ctas == 2) fl.threads = wide2;
                }
                fl.tr = 1;
                fl.gstride = 0;
                fplan.push_back(fl);
                i = j;
                continue;
            }
            if (!fl.global_front && S >= 32 && bulk_enabled && bulk_variant_for(nf) != 0) {
                fl.maxnf = bulk_variant_for(nf);
                // TMA-staged kernel: front + staging for the children's update blocks must fit in shared memory
                int need = 0;
                for (int q = i; q < j; ++q) {
                    const int f = plan_fronts[q];
                    const int fsz = sym.f_nf[f] * (sym.f_nf[f] + 1);
                    int largest = 0, total_c = 0;
                    for (int ci = sym.f_childptr[f]; ci < sym.f_childptr[f + 1]; ++ci) {
                        const int cc = sym.f_children[ci];
                        const int uc = sym.f_nf[cc] - sym.f_k[cc];
                        largest = std::max(largest, uc * (uc + 1));
                        total_c += uc * (uc + 1);
                    }
                    need = std::max(need, fsz + std::max(2 * fl.maxnf, std::min(total_c, std::max(largest, 192))));
                }
                size_t bytes = (size_t)need * 256 + (size_t)need * 4 + 64;
                if (bulk_ring) {
                    // ring mode: front + two stages of 64 / 32 / 16 elements (16 / 8 / 4 KB), whichever keeps the most
                    // CTAs resident (ties: the larger stage); the ring also holds the two pivot strips (2 x maxnf elements)
                    int fmax = 0;
                    for (int q = i; q < j; ++q) fmax = std::max(fmax, sym.f_nf[plan_fronts[q]] * (sym.f_nf[plan_fronts[q]] + 1));
                    int best_stage = 0;
                    size_t best_ctas = 0;
                    for (int stg : {64, 32, 16}) {
                        if (2 * stg < 2 * fl.maxnf) continue;
                        const size_t b = (size_t)(fmax + 2 * stg) * 256;
                        const size_t ctas = std::min<size_t>(2048 / (32 * bulk_lanes_for(fl.maxnf)), 233472 / (b + 1024));
                        if (ctas > best_ctas) { best_ctas = ctas; best_stage = stg; }
                    }
                    fl.bulk = true;
                    fl.ts = 32;
                    fl.staged = true;
                    fl.ring_elems = best_stage;
                    fl.smem_elems = fmax + 2 * best_stage;
                    fl.threads = 32 * bulk_lanes_for(fl.maxnf);
                    fl.smem = (size_t)fl.smem_elems * 256 + 64;
                } else if (bytes <= 200 * 1024) {
                    fl.bulk = true;
                    fl.ts = 32;
                    fl.smem_elems = need;
                    fl.threads = 32 * bulk_lanes_for(fl.maxnf);
                    fl.smem = bytes;
                }
            }
            if (fl.bulk) {
            } else if (fl.global_front) {
                fl.ts = (S == 1) ? 1 : 4;
                fl.threads = (S == 1) ? 512 : 256;
            } else {
                const PlanRule& rl = rules[std::min<size_t>(c, rules.size() - 1)];
                fl.ts = std::min(rl.ts, S);
                fl.threads = std::max(rl.threads, fl.ts);
                // launch bounds: 256 threads for scenario tiles, 512 (LU) / 1024 (LDL^T) for single-scenario CTAs
                fl.threads = std::min(fl.threads, fl.ts > 1 ? 256 : (fl.sym ? 1024 : 512));
                while (fl.ts > 1 && per * fl.ts > 200 * 1024) fl.ts /= 2;
            }
            int te = fl.threads / fl.ts;
            int trw = 1;
            while (trw < (nf + 1) / 2 && trw < te) trw *= 2;   // rows first: one lane per pair of front rows
            fl.tr = std::min(pow2_floor(te), trw);
            if (fl.bulk) fl.sym = false;
            if (fl.sym) {
                const size_t sper = ((size_t)nf * (nf + 1) / 2 + nf + (size_t)nf * 8) * sizeof(double);
                if (c >= (int)rules.size()) { fl.ts = 1; fl.threads = 256; }
                while (fl.ts > 1 && sper * fl.ts > 200 * 1024) fl.ts /= 2;
                fl.smem = sper * fl.ts;
                int te2 = fl.threads / fl.ts, trw2 = 1;
                while (trw2 < (nf + 1) / 2 && trw2 < te2) trw2 *= 2;      // folded rows: nf/2 row pairs
                fl.tr = std::min(pow2_floor(te2), trw2);
            } else if (!fl.bulk)
                fl.smem = fl.global_front ? (size_t)(nf * 8 + 8 * (nf + 1)) * fl.ts * sizeof(double) : per * fl.ts;
            fl.gstride = (long long)nf * (nf + 1) * fl.ts;
            if (staged_enabled && !fl.sym && !fl.bulk && !fl.global_front && fl.ts >= 2 && S % fl.ts == 0) {
                // ring of two stages behind the largest front of the launch: the largest stage (8 KB .. 1 KB) that keeps
                // the number of resident CTAs the front alone allows (at most 3: registers)
                auto ctas = [](size_t bytes) { return std::min<size_t>(JGB_MID_MINBLOCKS, 233472 / (bytes + 1024)); };
                size_t stage = 8192;
                while (stage > 1024 && (ctas(fl.smem + 2 * stage) < ctas(fl.smem) || fl.smem + 2 * stage > 200 * 1024)) stage /= 2;
                if (fl.smem + 2 * stage <= 200 * 1024 && ctas(fl.smem + 2 * stage) == ctas(fl.smem)) {
                    fl.staged = true;
                    fl.ring_elems = (int)(stage / (8 * fl.ts));
                    fl.ring_off = (int)(fl.smem / 8);
                    fl.smem += 2 * stage;
                }
            }
            if (fl.global_front)
                gwork_need = std::max<size_t>(gwork_need, (size_t)fl.gstride * fl.count * (S / fl.ts));
            if (fl.smem > 200 * 1024) throw std::runtime_error("front too large for shared memory");
            fplan.push_back(fl);
            i = j;
        }
    }
    // back-solve: one launch per depth level; batch launches are additionally cut by front size so that the
    // scenario tile (shared-memory footprint of the staged U rows) matches the class
    static const std::vector<PlanRule> bs_rules = {{8, 32, 128}, {12, 32, 128}, {16, 32, 128}, {24, 16, 128},
                                                   {32, 8, 128}, {48, 4, 128}, {64, 4, 128}, {96, 2, 128},
                                                   {1 << 30, 1, 128}};
    const std::vector<PlanRule> brules = parse_rules("JGB_BPLAN_BATCH", bs_rules);
    auto bcls = [&](int nf) { size_t c = 0; while (c + 1 < brules.size() && nf > brules[c].maxnf) ++c; return (int)c; };
    for (int sl_ = (int)slot_seqptr.size() - 2; seq_mode && sl_ >= 0; --sl_) {      // sequence mode: the factor slots backwards
        SolveLaunch sl{};
        sl.begin = plan_levelptr[sl_];
        sl.count = plan_levelptr[sl_ + 1] - plan_levelptr[sl_];
        sl.blocked = true;
        sl.ts = 1;
        sl.bs_rows = kBsRows;
        sl.seq_begin = slot_seqptr[sl_];
        sl.nseq = slot_seqptr[sl_ + 1] - slot_seqptr[sl_];
        auto need = [&](int rows_) {
            size_t worst = 0;
            for (int q = sl.begin; q < sl.begin + sl.count; ++q) {
                const int f = plan_fronts[q];
                const int nf = sym.f_nf[f], kb = std::min(sym.f_k[f], rows_);
                worst = std::max(worst, ((size_t)kb * (nf + 1) - (size_t)kb * (kb - 1) / 2 + nf) * sizeof(double));
                sl.max_nf = std::max(sl.max_nf, nf);
                sl.max_k = std::max(sl.max_k, sym.f_k[f]);
            }
            return worst;
        };
        while (sl.bs_rows > 1 && need(sl.bs_rows) > 200 * 1024) sl.bs_rows /= 2;
        sl.smem = need(sl.bs_rows);
        if (sl.smem > 200 * 1024) throw std::runtime_error("front too large for the back-solve staging buffer");
        splan.push_back(sl);
    }
    for (int d = 0; d < sym.ndepths && !seq_mode; ++d) {
        int b = sym.depthptr[d], e = sym.depthptr[d + 1];
        int i = b;
        while (i < e) {
            int j = e;
            int c = 0;
            if (S > 1) {
                c = bcls(sym.f_nf[sym.depth_fronts[i]]);
                j = i;
                while (j < e && bcls(sym.f_nf[sym.depth_fronts[j]]) == c) ++j;
            }
            SolveLaunch sl{};
            sl.begin = i;
            sl.count = j - i;
            size_t smem = 0, full = 0;
            for (int q = i; q < j; ++q) {
                int f = sym.depth_fronts[q];
                int nf = sym.f_nf[f], k = sym.f_k[f];
                int kb = std::min(k, 32);      // single-case kernel: largest staged block = the first kb rows
                size_t usz = (size_t)kb * (nf + 1) - (size_t)kb * (kb - 1) / 2;
                smem = std::max(smem, (usz + nf) * sizeof(double));
                size_t uall = (size_t)k * (nf + 1) - (size_t)k * (k - 1) / 2;
                full = std::max(full, (uall + nf) * sizeof(double));
                sl.max_nf = std::max(sl.max_nf, nf);
                sl.max_k = std::max(sl.max_k, k);
            }
            int tile_ts = 0;
            if (S > 1 && full <= 200 * 1024) {
                tile_ts = std::min(brules[c].ts, S);
                while (tile_ts > 1 && full * tile_ts > 100 * 1024) tile_ts /= 2;
            }
            // single-scenario tiles (the big fronts at the top of the tree) take the blocked kernel: a warp per pivot row
            // with a shuffle reduction instead of one lane per row (JGB_BS_TS1_TILE=1 keeps the tile kernel; tuning only)
            static const bool ts1_tile = getenv("JGB_BS_TS1_TILE") && atoi(getenv("JGB_BS_TS1_TILE")) == 1;
            if (tile_ts >= 2 || (tile_ts == 1 && ts1_tile)) {
                sl.ts = tile_ts;
                smem = full * sl.ts;
            } else {
                // single case, or a batch front whose packed rows do not fit shared memory: blocks of 32 rows
                sl.blocked = true;
                sl.ts = 1;
                sl.bs_rows = kBsRows;           // fewer rows per block when 32 packed rows exceed shared memory
                auto need = [&](int rows_) {
                    size_t worst = 0;
                    for (int q = i; q < j; ++q) {
                        const int f = sym.depth_fronts[q];
                        const int nf = sym.f_nf[f], kb = std::min(sym.f_k[f], rows_);
                        worst = std::max(worst, ((size_t)kb * (nf + 1) - (size_t)kb * (kb - 1) / 2 + nf) * sizeof(double));
                    }
                    return worst;
                };
                while (sl.bs_rows > 1 && need(sl.bs_rows) > 200 * 1024) sl.bs_rows /= 2;
                smem = need(sl.bs_rows);
                if (smem > 200 * 1024) throw std::runtime_error("front too large for the back-solve staging buffer");
            }
            sl.smem = smem;
            splan.push_back(sl);
            i = j;
        }
    }
    {
        // Tile width of every front's launch, and with it the section of the update storage each block goes to: the
        // block of front f is written at the width its PARENT reads with (see DevSym).
        const int F = sym.nfronts;
        const int wmax = S < 32 ? S : 32;
        std::vector<int> ts_of(F, wmax);
        std::vector<char> lower(F, 0);
        for (const FactorLaunch& fl : fplan)
            for (int q = fl.begin; q < fl.begin + fl.count; ++q) {
                const int f = plan_fronts[q];
                ts_of[f] = fl.bulk ? wmax : fl.ts;
                if ((fl.sym || fl.dense) && !fl.bulk) lower[f] = 1;
            }
        std::vector<int> wout(F, wmax);
        std::vector<long long> off(F, 0), secsz(6, 0);
        for (int f = 0; f < F; ++f) {
            const int par = sym.f_parent[f];
            wout[f] = par < 0 ? wmax : ts_of[par];
            const long long u = sym.f_nf[f] - sym.f_k[f];
            off[f] = secsz[lg2(wout[f])];
            secsz[lg2(wout[f])] += u * (u + 1);
        }
        // packed U rows: section = tile width of the back-solve launch that reads the front's rows
        std::vector<int> wu(F, wmax);
        for (const SolveLaunch& sl : splan) {
            const bool reg = !sl.blocked && sl.max_nf <= backsolve_reg_max() && S % 32 == 0;
            const int w = sl.blocked ? 1 : reg ? 32 : sl.ts;
            for (int q = sl.begin; q < sl.begin + sl.count; ++q)
                wu[sl.nseq ? plan_fronts[q] : sym.depth_fronts[q]] = std::min(w, wmax);
        }
        std::vector<long long> uoff(F, 0), usecsz(6, 0);
        for (int f = 0; f < F; ++f) {
            uoff[f] = usecsz[lg2(wu[f])];
            usecsz[lg2(wu[f])] += sym.f_uoff.size() > (size_t)f + 1 ? sym.f_uoff[f + 1] - sym.f_uoff[f] : sym.u_size - sym.f_uoff[f];
        }
        {
            long long ub = 0;
            for (int q = 0; q < 6; ++q) {
                dev.usec_base[q] = ub * S;
                dev.usec_size[q] = usecsz[q];
                ub += usecsz[q];
                if (usecsz[q] >= (1LL << 31) / 32) throw std::runtime_error("packed-U section exceeds the 32-bit element offsets");
            }
            d_plan_uoff.alloc(F);
            if (F) JGB_CUDA(cudaMemcpy(d_plan_uoff.p, uoff.data(), (size_t)F * sizeof(long long), cudaMemcpyHostToDevice));
            dev.f_uoff = d_plan_uoff.p;
        }
        long long base = 0;
        for (int q = 0; q < 6; ++q) {
            dev.sec_base[q] = base * S;
            dev.sec_size[q] = secsz[q];
            base += secsz[q];
            if (secsz[q] >= (1LL << 31) / 32) throw std::runtime_error("update-storage section exceeds the 32-bit element offsets");
        }
        if (!tplan.empty()) {          // task blobs: update-storage offsets and root tile widths of this layout
            std::vector<int> blob(task_plan.blob);
            for (size_t q = 0; q < task_plan.off_pos.size(); ++q) {
                const long long o = off[task_plan.off_front[q]];
                blob[task_plan.off_pos[q]] = (int)(o & 0xffffffffLL);
                blob[task_plan.off_pos[q] + 1] = (int)(o >> 32);
            }
            for (size_t q = 0; q < task_plan.wout_pos.size(); ++q) blob[task_plan.wout_pos[q]] = wout[task_plan.wout_front[q]];
            for (size_t q = 0; q < task_plan.uoff_pos.size(); ++q) {
                const int f = task_plan.uoff_front[q];
                if (wu[f] != 32) throw std::logic_error("task front outside the register back-solve class");
                blob[task_plan.uoff_pos[q]] = (int)(uoff[f] & 0xffffffffLL);
                blob[task_plan.uoff_pos[q] + 1] = (int)(uoff[f] >> 32);
            }
            d_task_blob.alloc(blob.size());
            d_task_desc.alloc(task_plan.descs.size() / 2);
            JGB_CUDA(cudaMemcpy(d_task_blob.p, blob.data(), blob.size() * sizeof(int), cudaMemcpyHostToDevice));
            JGB_CUDA(cudaMemcpy(d_task_desc.p, task_plan.descs.data(), task_plan.descs.size() * sizeof(int),
                                cudaMemcpyHostToDevice));
        }
        // gather lists with the source offsets of this layout: a pair's source is (child block offset + element)
        auto remap = [&](const std::vector<int>& eaptr, const std::vector<int>& roundptr, const std::vector<int>& pairs,
                         DevBuf<int>& out) {
            std::vector<int> np(pairs);
            for (int f = 0; f < F; ++f) {
                const int c0 = sym.f_childptr[f], c1 = sym.f_childptr[f + 1];
                if (c0 == c1) continue;
                for (int t = roundptr[eaptr[f]]; t < roundptr[eaptr[f + 1]]; ++t) {
                    const long long src = pairs[2 * t + 1];
                    int lo = c0, hi = c1 - 1;               // children ascend in index and in f_updoff
                    while (lo < hi) {
                        const int mid = (lo + hi + 1) / 2;
                        if (sym.f_updoff[sym.f_children[mid]] <= src) lo = mid; else hi = mid - 1;
                    }
                    const int c = sym.f_children[lo];
                    np[2 * t + 1] = (int)(off[c] + (src - sym.f_updoff[c]));
                }
            }
            out.alloc(np.size());
            if (!np.empty()) JGB_CUDA(cudaMemcpy(out.p, np.data(), np.size() * sizeof(int), cudaMemcpyHostToDevice));
        };
        remap(sym.f_eaptr, sym.ea_roundptr, sym.ea_pair, d_plan_pair);
        dev.ea_pair = reinterpret_cast<const int2*>(d_plan_pair.p);
        if (symmetric) {
            remap(sym.f_eaptr_sym, sym.ea_roundptr_sym, sym.ea_pair_sym, d_plan_pair_s);
            dev.ea_pair_s = reinterpret_cast<const int2*>(d_plan_pair_s.p);
        }
        {
            std::vector<ChildDesc> cds(sym.f_children.size());
            for (size_t q = 0; q < cds.size(); ++q) {
                const int c = sym.f_children[q];
                cds[q].uc = sym.f_nf[c] - sym.f_k[c];
                cds[q].relptr = sym.f_relptr[c];
                cds[q].updoff = off[c];
            }
            d_plan_child.alloc(cds.size());
            if (!cds.empty()) JGB_CUDA(cudaMemcpy(d_plan_child.p, cds.data(), cds.size() * sizeof(ChildDesc), cudaMemcpyHostToDevice));
            dev.child_desc = d_plan_child.p;
        }
        // staged extend-add: destination of every update-storage element in its parent's front, and the chunk lists
        std::vector<long long> cum(7, 0);
        for (int q = 0; q < 6; ++q) cum[q + 1] = cum[q] + secsz[q];
        std::vector<int> chunk_lo(F, 0), chunk_hi(F, 0);
        {
            bool any = false;
            for (const FactorLaunch& fl : fplan) any = any || fl.staged;
            if (any) {
                std::vector<char> dense_lu_front(F, 0);
                for (const FactorLaunch& fl : fplan)
                    if (fl.dense_lu)
                        for (int q = fl.begin; q < fl.begin + fl.count; ++q) dense_lu_front[plan_fronts[q]] = 1;
                std::vector<int> dst((size_t)sym.upd_size, 0);
                for (int c = 0; c < F; ++c) {
                    const int par = sym.f_parent[c];
                    if (par < 0) continue;
                    const int uc = sym.f_nf[c] - sym.f_k[c];
                    const int nfp = dense_lu_front[par] ? dense_lu_ld(sym.f_nf[par]) : sym.f_nf[par];     // column stride
                    const int* rel = sym.f_rel.data() + sym.f_relptr[c];
                    int* d = dst.data() + cum[lg2(wout[c])] + off[c];
                    for (int j = 0; j <= uc; ++j) {
                        const int dc = j < uc ? rel[j] : sym.f_nf[par];
                        for (int i = 0; i < uc; ++i) d[(size_t)j * uc + i] = rel[i] + dc * nfp;
                    }
                }
                std::vector<int> chunks;
                for (FactorLaunch& fl : fplan) {
                    if (!fl.staged) continue;
                    fl.sec_cum = (int)cum[lg2(fl.ts)];
                    for (int q = fl.begin; q < fl.begin + fl.count; ++q) {
                        const int f = plan_fronts[q];
                        chunk_lo[f] = (int)(chunks.size() / 2);
                        for (int ci = sym.f_childptr[f]; ci < sym.f_childptr[f + 1]; ++ci) {
                            const int c = sym.f_children[ci];
                            const long long uc = sym.f_nf[c] - sym.f_k[c], total = uc * (uc + 1);
                            for (long long e = 0; e < total; e += fl.ring_elems) {
                                chunks.push_back((int)(off[c] + e));
                                chunks.push_back((int)std::min<long long>(fl.ring_elems, total - e));
                            }
                        }
                        chunk_hi[f] = (int)(chunks.size() / 2);
                    }
                }
                d_upd_dst.alloc(dst.size());
                d_chunks.alloc(std::max<size_t>(chunks.size(), 2));
                if (!dst.empty()) JGB_CUDA(cudaMemcpy(d_upd_dst.p, dst.data(), dst.size() * sizeof(int), cudaMemcpyHostToDevice));
                if (!chunks.empty()) JGB_CUDA(cudaMemcpy(d_chunks.p, chunks.data(), chunks.size() * sizeof(int), cudaMemcpyHostToDevice));
            }
        }
        std::vector<char> is_staged(F, 0);
        for (const FactorLaunch& fl : fplan)
            if (fl.staged)
                for (int q = fl.begin; q < fl.begin + fl.count; ++q) is_staged[plan_fronts[q]] = 1;
        // per-front descriptors in launch order. Fronts factored on packed lower triangles (LDL^T kernels) take the
        // symmetric gather lists, and a front whose parent is such a front writes only the lower triangle of its block.
        std::vector<FrontDesc> descs(plan_fronts.size());
        for (size_t q = 0; q < descs.size(); ++q) {
            const int f = plan_fronts[q];
            FrontDesc& d = descs[q];
            d.f = f; d.nf = sym.f_nf[f]; d.k = sym.f_k[f]; d.rowptr = sym.f_rowptr[f];
            d.asm0 = sym.f_asmptr[f]; d.asm1 = sym.f_asmptr[f + 1];
            d.child0 = sym.f_childptr[f]; d.child1 = sym.f_childptr[f + 1];
            if (is_staged[f]) { d.child0 = chunk_lo[f]; d.child1 = chunk_hi[f]; }      // chunk range instead
            if (lower[f]) { d.ea0 = sym.f_eaptr_sym[f]; d.ea1 = sym.f_eaptr_sym[f + 1]; }
            else { d.ea0 = sym.f_eaptr[f]; d.ea1 = sym.f_eaptr[f + 1]; }
            const int par = sym.f_parent[f];
            d.flags = ((par >= 0 && lower[par]) ? 1 : 0) | (wu[f] << 8);
            d.wout = wout[f];
            d.uoff = uoff[f]; d.updoff = off[f];
        }
        if (seq_mode) {
            d_seqptr.alloc(plan_seqptr.size());
            JGB_CUDA(cudaMemcpy(d_seqptr.p, plan_seqptr.data(), plan_seqptr.size() * sizeof(int), cudaMemcpyHostToDevice));
        }
        d_plan_desc.alloc(descs.size());
        d_plan_fronts.alloc(plan_fronts.size());
        if (!descs.empty()) {
            JGB_CUDA(cudaMemcpy(d_plan_desc.p, descs.data(), descs.size() * sizeof(FrontDesc), cudaMemcpyHostToDevice));
            JGB_CUDA(cudaMemcpy(d_plan_fronts.p, plan_fronts.data(), plan_fronts.size() * sizeof(int),
                                cudaMemcpyHostToDevice));
        }
    }
    {
        // DAG schedule (batches): one in-order stream per launch class; a launch waits for the latest launch of every
        // other stream that holds a child of one of its fronts. JGB_LANES=1 keeps everything on the caller's stream.
        static const int lanes_env = getenv("JGB_LANES") ? atoi(getenv("JGB_LANES")) : 0;
        nlanes = 1;
        if (S > 1 && tplan.empty() && lanes_env != 1) {
            std::vector<long long> keys;       // class key -> lane
            std::vector<int> launch_of(sym.nfronts, -1);
            for (size_t i = 0; i < fplan.size(); ++i) {
                FactorLaunch& fl = fplan[i];
                const long long key = fl.global_front ? 1 : fl.dense ? 2 : fl.dense_lu ? 3 : fl.bulk ? 100 + fl.maxnf : fl.sym ? 200 + fl.ts : 300 + fl.ts;
                size_t l = 0;
                while (l < keys.size() && keys[l] != key) ++l;
                if (l == keys.size()) keys.push_back(key);
                fl.lane = (int)l;
                if (lanes_env > 1) fl.lane %= lanes_env;
                for (int q = fl.begin; q < fl.begin + fl.count; ++q) launch_of[plan_fronts[q]] = (int)i;
            }
            nlanes = lanes_env > 1 ? std::min<int>(lanes_env, (int)keys.size()) : (int)keys.size();
            for (size_t i = 0; i < fplan.size(); ++i) {
                FactorLaunch& fl = fplan[i];
                std::vector<int> latest(nlanes, -1);
                for (int q = fl.begin; q < fl.begin + fl.count; ++q) {
                    const int f = plan_fronts[q];
                    for (int ci = sym.f_childptr[f]; ci < sym.f_childptr[f + 1]; ++ci) {
                        const int li = launch_of[sym.f_children[ci]];
                        if (li >= 0 && fplan[li].lane != fl.lane) latest[fplan[li].lane] = std::max(latest[fplan[li].lane], li);
                    }
                }
                fl.deps.clear();
                for (int l = 0; l < nlanes; ++l)
                    if (latest[l] >= 0) { fl.deps.push_back(latest[l]); fplan[latest[l]].record = true; }
            }
            while ((int)lanes.size() < nlanes - 1) {
                cudaStream_t x;
                JGB_CUDA(cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking));
                lanes.push_back(x);
            }
            while (lane_events.size() < fplan.size() + (size_t)nlanes) {
                cudaEvent_t e;
                JGB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                lane_events.push_back(e);
            }
            if (!fork_event) JGB_CUDA(cudaEventCreateWithFlags(&fork_event, cudaEventDisableTiming));
        }
    }
    // scenario tiles sit in gridDim.y: a batch whose smallest tile width needs more than 65 535 tiles cannot be launched
    for (const FactorLaunch& fl : fplan)
        if (S / std::max(1, fl.ts) > 65535)
            throw std::invalid_argument("batch too large for the scenario tiles of this matrix (more than 65 535 tiles): split it");
    for (const SolveLaunch& sl : splan)
        if ((sl.blocked ? S : S / std::max(1, sl.ts)) > 65535 && S > 1)
            throw std::invalid_argument("batch too large for the back-solve tiles of this matrix (more than 65 535 tiles): split it");
    if (getenv("JGB_PLAN_DEBUG")) {
        for (const FactorLaunch& fl : fplan)
            fprintf(stderr, "[plan S=%d] fronts %4d nf<=%3d ts %2d threads %4d smem %6zu %s%s%s%s%s ring %d\n", S, fl.count,
                    sym.f_nf[plan_fronts[fl.begin]], fl.ts, fl.threads, fl.smem, fl.bulk ? "bulk " : "", fl.sym ? "sym " : "",
                    fl.dense ? "dense " : fl.dense_lu ? "denselu " : "", fl.global_front ? "global " : "", fl.staged ? "staged" : "", fl.ring_elems);
    }
    d_U.alloc((size_t)sym.u_size * S);
    d_upd.alloc((size_t)sym.upd_size * S);
    if (gwork_need) d_gwork.alloc(gwork_need);
    planned_S = S;
}

MfSolver::~MfSolver() {
    for (cudaStream_t x : lanes) cudaStreamDestroy(x);
    for (cudaEvent_t e : lane_events) cudaEventDestroy(e);
    if (fork_event) cudaEventDestroy(fork_event);
}

int MfSolver::launches_per_solve(int S) {
    plan(S);
    return (int)(tplan.size() + fplan.size() + splan.size());
}

int64_t MfSolver::factor_bytes(int S) const {
    // read A values + rhs, write U rows, write + read update blocks, back-solve reads U and writes x
    int64_t per = 8LL * ((int64_t)sym.asm_src.size() + sym.n + 2 * sym.u_size + 2 * sym.upd_size + sym.n);
    return per * S;
}

void MfSolver::factor_solve(const double* aval, const double* rhs, double* x, int S, const unsigned char* active,
                            int* status, cudaStream_t st, cudaEvent_t after_factor) {
    plan(S);
    for (const TaskLaunch& tl : tplan)
        launch_task(tl.te, tl.maxnf, dim3(tl.count, S / 32), tl.smem, st, dev, d_task_blob.p, d_task_desc.p + tl.begin,
                    aval, rhs, d_U.p, d_upd.p, S, tl.front_cap, tl.stack_cap, active, status);
    if (nlanes > 1) {
        JGB_CUDA(cudaEventRecord(fork_event, st));
        for (int l = 1; l < nlanes; ++l) JGB_CUDA(cudaStreamWaitEvent(lanes[l - 1], fork_event, 0));
    }
    cudaStream_t const st0 = st;
    auto staged_args = [&](const FactorLaunch& fl) {
        StagedEa sg{};
        if (fl.staged) {
            sg.chunks = reinterpret_cast<const int2*>(d_chunks.p);
            sg.upd_dst = d_upd_dst.p;
            sg.sec_cum = fl.sec_cum; sg.ring_elems = fl.ring_elems; sg.ring_off = fl.ring_off;
        }
        return sg;
    };
    for (size_t li = 0; li < fplan.size(); ++li) {
        const FactorLaunch& fl = fplan[li];
        dim3 grid(fl.count, S / fl.ts);
        if (nlanes > 1) {
            st = fl.lane == 0 ? st0 : lanes[fl.lane - 1];
            for (int d : fl.deps) JGB_CUDA(cudaStreamWaitEvent(st, lane_events[d], 0));
        }
        if (fl.dense_lu && fl.nseq)
            mf_factor_dense_lu_kernel<<<dim3(fl.nseq, S), fl.threads, fl.smem, st>>>(dev, d_plan_desc.p, aval, rhs, d_U.p, d_upd.p,
                                                                                     S, active, status, d_seqptr.p + fl.seq_begin,
                                                                                     staged_args(fl));
        else if (fl.dense_lu)
            mf_factor_dense_lu_kernel<<<grid, fl.threads, fl.smem, st>>>(dev, d_plan_desc.p + fl.begin, aval, rhs, d_U.p,
                                                                          d_upd.p, S, active, status, nullptr, staged_args(fl));
        else if (fl.dense)
            mf_factor_dense_sym_kernel<<<grid, fl.threads, fl.smem, st>>>(dev, d_plan_desc.p + fl.begin, aval, rhs, d_U.p,
                                                                           d_upd.p, S, active, status);
        else if (fl.bulk)
            launch_factor_bulk(fl.maxnf, fl.threads / 32, grid, fl.smem, st, dev, d_plan_desc.p + fl.begin, aval, rhs, d_U.p,
                               d_upd.p, S, fl.smem_elems, active, status, staged_args(fl));
        else if (fl.global_front)
            launch_factor<true>(fl.ts, grid, fl.threads, fl.smem, st, dev, d_plan_fronts.p + fl.begin,
                                d_plan_desc.p + fl.begin, aval, rhs, d_U.p, d_upd.p, S, fl.tr,
                                active, status, d_gwork.p, fl.gstride);
        else if (fl.sym)
            launch_factor_sym(fl.ts, grid, fl.threads, fl.smem, st, dev, d_plan_desc.p + fl.begin, aval, rhs, d_U.p,
                              d_upd.p, S, fl.tr, active, status);
        else {
            const StagedEa sg = staged_args(fl);
            launch_factor<false>(fl.ts, grid, fl.threads, fl.smem, st, dev, d_plan_fronts.p + fl.begin,
                                 d_plan_desc.p + fl.begin, aval, rhs, d_U.p, d_upd.p, S, fl.tr,
                                 active, status, nullptr, 0, sg);
        }
        if (nlanes > 1 && fl.record) JGB_CUDA(cudaEventRecord(lane_events[li], st));
    }
    if (nlanes > 1) {          // join: the caller's stream waits for the tail of every lane
        st = st0;
        for (int l = 1; l < nlanes; ++l) {
            cudaEvent_t e = lane_events[fplan.size() + l];
            JGB_CUDA(cudaEventRecord(e, lanes[l - 1]));
            JGB_CUDA(cudaStreamWaitEvent(st, e, 0));
        }
    }
    if (after_factor) JGB_CUDA(cudaEventRecord(after_factor, st));
    for (const SolveLaunch& sl : splan) {
        if (sl.blocked && sl.nseq) {
            mf_backsolve_single<<<dim3(sl.nseq, S), backsolve_single_threads(), sl.smem, st>>>(dev, d_plan_fronts.p, d_U.p, x, active, S,
                                                                                              sl.bs_rows, d_seqptr.p + sl.seq_begin);
        } else if (sl.blocked) {
            mf_backsolve_single<<<dim3(sl.count, S), backsolve_single_threads(), sl.smem, st>>>(dev, d_depth_fronts.p + sl.begin, d_U.p, x,
                                                                        active, S, sl.bs_rows, nullptr);
        } else if (sl.max_nf <= backsolve_reg_max() && S % 32 == 0) {
            launch_backsolve_reg(sl.max_nf, sl.count, S, st, dev, d_depth_fronts.p + sl.begin, d_U.p, x, active);
        } else {
            launch_backsolve_tile(sl.ts, dim3(sl.count, S / sl.ts), sl.smem, st, dev, d_depth_fronts.p + sl.begin,
                                  d_U.p, x, S, active);
        }
    }
    JGB_CUDA(cudaGetLastError());
}

void MfSolver::solve_multi(double* B, int R, cudaStream_t st, const MfSolver* lower) {
    if (!symmetric && !lower) throw std::logic_error("solve_multi: an unsymmetric matrix needs the factor of its transpose");
    if (lower && (lower->planned_S != 1 || lower->sym.u_size != sym.u_size))
        throw std::logic_error("solve_multi: the transpose factor does not match");
    if (planned_S != 1) throw std::logic_error("solve_multi: factor the matrix (S = 1) first");
    if (R <= 0 || R % 32 != 0) throw std::invalid_argument("solve_multi: the block width must be a multiple of 32");
    if (d_coff.n == 0) {
        std::vector<long long> coff(sym.nfronts + 1, 0);
        for (int f = 0; f < sym.nfronts; ++f) coff[f + 1] = coff[f] + (sym.f_nf[f] - sym.f_k[f]);
        csum = coff[sym.nfronts];
        d_coff.upload(coff, st);
        JGB_CUDA(cudaStreamSynchronize(st));
        JGB_CUDA(cudaFuncSetAttribute(mf_fwd_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        JGB_CUDA(cudaFuncSetAttribute(mf_bwd_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    d_cvec.alloc((size_t)std::max<long long>(csum, 1) * R);
    auto launch_dims = [&](const int* list, int b, int e, int& threads, size_t& smem) {
        int mx = 0;
        for (int q = b; q < e; ++q) mx = std::max(mx, sym.f_nf[list[q]]);
        threads = mx <= 16 ? 64 : mx <= 64 ? 128 : 256;
        smem = (size_t)mx * 32 * sizeof(double);
        if (smem > 200 * 1024) throw std::runtime_error("front too large for the multi right-hand-side solve");
    };
    for (int l = 0; l < sym.nlevels; ++l) {
        const int b = sym.levelptr[l], e = sym.levelptr[l + 1];
        int threads; size_t smem;
        launch_dims(sym.level_fronts.data(), b, e, threads, smem);
        mf_fwd_multi_kernel<<<dim3(e - b, R / 32), threads, smem, st>>>(dev, d_level_fronts.p + b,
                                                                          lower ? lower->d_U.p : d_U.p, B, d_cvec.p,
                                                                          d_coff.p, R);
    }
    for (int d = 0; d < sym.ndepths; ++d) {
        const int b = sym.depthptr[d], e = sym.depthptr[d + 1];
        int threads; size_t smem;
        launch_dims(sym.depth_fronts.data(), b, e, threads, smem);
        mf_bwd_multi_kernel<<<dim3(e - b, R / 32), threads, smem, st>>>(dev, d_depth_fronts.p + b, d_U.p, B, R);
    }
    JGB_CUDA(cudaGetLastError());
}

const double* MfSolver::selected_inverse(cudaStream_t st) {
    if (!symmetric) throw std::logic_error("selected_inverse needs a symmetric matrix");
    if (planned_S != 1) throw std::logic_error("selected_inverse: factor the matrix (S = 1) first");
    if (d_zoff.n == 0) {
        zoff_host.assign(sym.nfronts + 1, 0);
        for (int f = 0; f < sym.nfronts; ++f)
            zoff_host[f + 1] = zoff_host[f] + (long long)sym.f_nf[f] * sym.f_nf[f];
        d_zoff.upload(zoff_host, st);
        d_parent.upload(sym.f_parent, st);
        JGB_CUDA(cudaStreamSynchronize(st));
        d_Zinv.alloc((size_t)zoff_host[sym.nfronts]);
    }
    for (int d = 0; d < sym.ndepths; ++d) {
        const int b = sym.depthptr[d], e = sym.depthptr[d + 1];
        int mx = 0;
        for (int q = b; q < e; ++q) mx = std::max(mx, sym.f_nf[sym.depth_fronts[q]]);
        mf_selinv_kernel<<<e - b, selinv_threads(mx), 2 * (size_t)mx * sizeof(double), st>>>(dev, d_depth_fronts.p + b, d_parent.p,
                                                                              d_U.p, d_Zinv.p, d_zoff.p);
    }
    JGB_CUDA(cudaGetLastError());
    return d_Zinv.p;
}

}  // namespace jgb
