/* CPU oracle / CPU baseline only (TEST INFRASTRUCTURE): numeric refactorisation of a sparse LU on a fixed pattern.
 *
 * The reference factors the Jacobian (and the WLS gain) symbolically once and afterwards only refactors numerically:
 * `factorization!` -> `lu!(F, A)` / `klu!(F, A)` (/root/reference/src/backend/utility.jl:478-500, called from
 * src/powerFlow/acPowerFlow.jl:890-895 and src/stateEstimation/acStateEstimation.jl:887-898). UMFPACK / KLU themselves are
 * third-party (SuiteSparse, sources not under /root/reference); this file restates the published algorithm of
 * klu_refactor (Davis & Palamadai Natarajan, ACM TOMS 37(3), 2010, Alg. "refactor"): left-looking column LU that
 * reuses the column ordering, the row (pivot) order and the patterns of L and U found by the first factorisation — no
 * pivot search, no symbolic depth-first search, one sparse triangular solve per column on the stored pattern.
 * The first factorisation (ordering, pivoting, patterns) is SciPy's SuperLU in oracle/fast.py.
 *
 * Conventions: B = Pr A Pc = L U. row_new[i] = new row of A's row i; col_src[j] = column of A that is column j of B.
 * L is unit lower triangular and U upper triangular in CSC with ascending row indices inside every column (L stores its
 * unit diagonal first, U its diagonal last), exactly as scipy's `splu(...).L / .U` after sort_indices().
 * Returns 0, or j + 1 when the pivot of column j is zero or not finite. */
#include <math.h>
#include <stdint.h>

int olu_refactor(int32_t n, const int32_t* Ap, const int32_t* Ai, const double* Ax, const int32_t* row_new,
                 const int32_t* col_src, const int32_t* Lp, const int32_t* Li, double* Lx, const int32_t* Up,
                 const int32_t* Ui, double* Ux, double* work) {
    for (int32_t j = 0; j < n; ++j) {
        for (int32_t p = Up[j]; p < Up[j + 1]; ++p) work[Ui[p]] = 0.0;
        for (int32_t p = Lp[j]; p < Lp[j + 1]; ++p) work[Li[p]] = 0.0;
        const int32_t c = col_src[j];
        for (int32_t p = Ap[c]; p < Ap[c + 1]; ++p) work[row_new[Ai[p]]] = Ax[p];
        /* x = L \ B(:, j) on the stored pattern of U(:, j), rows ascending = a topological order */
        const int32_t ulast = Up[j + 1] - 1;
        for (int32_t p = Up[j]; p < ulast; ++p) {
            const int32_t k = Ui[p];
            const double xk = work[k];
            Ux[p] = xk;
            if (xk != 0.0)
                for (int32_t q = Lp[k] + 1; q < Lp[k + 1]; ++q) work[Li[q]] -= Lx[q] * xk;
        }
        const double d = work[j];
        Ux[ulast] = d;
        if (d == 0.0 || !isfinite(d)) return j + 1;
        Lx[Lp[j]] = 1.0;
        for (int32_t q = Lp[j] + 1; q < Lp[j + 1]; ++q) Lx[q] = work[Li[q]] / d;
    }
    return 0;
}

/* Solve A x = b with the factors above: z = U \ (L \ (Pr b)), x = Pc z. `x` may not alias `b`; work has n entries. */
void olu_solve(int32_t n, const int32_t* row_new, const int32_t* col_src, const int32_t* Lp, const int32_t* Li,
               const double* Lx, const int32_t* Up, const int32_t* Ui, const double* Ux, const double* b, double* x,
               double* work) {
    for (int32_t i = 0; i < n; ++i) work[row_new[i]] = b[i];
    for (int32_t j = 0; j < n; ++j) {
        const double yj = work[j];
        if (yj != 0.0)
            for (int32_t q = Lp[j] + 1; q < Lp[j + 1]; ++q) work[Li[q]] -= Lx[q] * yj;
    }
    for (int32_t j = n - 1; j >= 0; --j) {
        const int32_t ulast = Up[j + 1] - 1;
        const double zj = work[j] / Ux[ulast];
        work[j] = zj;
        for (int32_t p = Up[j]; p < ulast; ++p) work[Ui[p]] -= Ux[p] * zj;
    }
    for (int32_t j = 0; j < n; ++j) x[col_src[j]] = work[j];
}

/* Structural patterns of L and U for B = L U without pivoting (B given by its CSC pattern, ascending rows): column j of
 * L + U is the reach of B(:, j) in the graph of the columns of L already found (Gilbert & Peierls, SIAM J. Sci. Stat.
 * Comput. 9, 1988 — the symbolic step KLU runs inside its first factorisation). Rows above the diagonal are visited in
 * ascending order through a binary heap, which is a topological order and leaves U(:, j) sorted; L(:, j) is sorted at
 * the end of the column. mark (n ints) and heap (n ints) are workspaces. Returns 0, or -1 when capL / capU is too small. */
#include <stdlib.h>

static int cmp_i32(const void* a, const void* b) {
    const int32_t x = *(const int32_t*)a, y = *(const int32_t*)b;
    return (x > y) - (x < y);
}

int olu_symbolic(int32_t n, const int32_t* Bp, const int32_t* Bi, int32_t* Lp, int32_t* Li, int64_t capL, int32_t* Up,
                 int32_t* Ui, int64_t capU, int32_t* mark, int32_t* heap) {
    int64_t nl = 0, nu = 0;
    for (int32_t i = 0; i < n; ++i) mark[i] = -1;
    for (int32_t j = 0; j < n; ++j) {
        int32_t hn = 0;
        Lp[j] = (int32_t)nl;
        Up[j] = (int32_t)nu;
        if (nl + 1 > capL) return -1;
        Li[nl++] = j;                                  /* unit diagonal first */
        mark[j] = j;
#define OLU_VISIT(r)                                                         \
    do {                                                                     \
        const int32_t r_ = (r);                                              \
        if (mark[r_] != j) {                                                 \
            mark[r_] = j;                                                    \
            if (r_ < j) {                                                    \
                int32_t c = hn++;                                            \
                while (c > 0 && heap[(c - 1) / 2] > r_) { heap[c] = heap[(c - 1) / 2]; c = (c - 1) / 2; } \
                heap[c] = r_;                                                \
            } else {                                                         \
                if (nl + 1 > capL) return -1;                                \
                Li[nl++] = r_;                                               \
            }                                                                \
        }                                                                    \
    } while (0)
        for (int32_t p = Bp[j]; p < Bp[j + 1]; ++p) OLU_VISIT(Bi[p]);
        while (hn > 0) {
            const int32_t k = heap[0];
            const int32_t last = heap[--hn];           /* pop the minimum */
            int32_t c = 0;
            for (;;) {
                int32_t ch = 2 * c + 1;
                if (ch >= hn) break;
                if (ch + 1 < hn && heap[ch + 1] < heap[ch]) ++ch;
                if (heap[ch] >= last) break;
                heap[c] = heap[ch];
                c = ch;
            }
            if (hn > 0) heap[c] = last;
            if (nu + 1 > capU) return -1;
            Ui[nu++] = k;
            for (int32_t q = Lp[k] + 1; q < Lp[k + 1]; ++q) OLU_VISIT(Li[q]);
        }
#undef OLU_VISIT
        if (nu + 1 > capU) return -1;
        Ui[nu++] = j;                                  /* diagonal last */
        qsort(Li + Lp[j] + 1, (size_t)(nl - Lp[j] - 1), sizeof(int32_t), cmp_i32);
    }
    Lp[n] = (int32_t)nl;
    Up[n] = (int32_t)nu;
    return 0;
}
