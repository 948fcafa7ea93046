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
