/* CPU oracle, C restatement of the Newton-Raphson assembly loops (TEST INFRASTRUCTURE / CPU baseline only).
 *
 * Follows /root/reference/src/powerFlow/acPowerFlow.jl:645-685 (mismatch!) and :813-888 (Jacobian fill of solve!)
 * with the scalar formulas of src/backend/equations.jl:63-143, same loop structure and summation order as the
 * reference (including the second sincos pass that re-sums the row for the diagonal entry, :859-870).
 * Single-threaded like the reference. 0-based indices; complex arrays are interleaved (re, im).
 * Validated against the NumPy oracle (oracle/nr.py) by tests/test_oracle_c.py.
 */
#include <math.h>
#include <stdint.h>

void onr_mismatch(int64_t n, const int64_t* colptr, const int64_t* rowval, const double* ynz_t, const int8_t* type,
                  int64_t slack, const int64_t* pq, const int64_t* pvpq, const double* vm, const double* va,
                  const double* sup_p, const double* sup_q, const double* dem_p, const double* dem_q, double* mism,
                  double* stop) {
    double stop_p = 0.0, stop_q = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        if (i == slack) continue;
        const int64_t k = pvpq[i], q = pq[i];
        double cur_p = 0.0, cur_q = 0.0;
        const int is_pq = type[i] == 1;
        for (int64_t ptr = colptr[i]; ptr < colptr[i + 1]; ++ptr) {
            const int64_t row = rowval[ptr];
            const double G = ynz_t[2 * ptr], B = ynz_t[2 * ptr + 1];
            const double d = va[i] - va[row];
            const double s = sin(d), c = cos(d);
            cur_p += vm[row] * (G * c + B * s);
            if (is_pq) cur_q += vm[row] * (G * s - B * c);
        }
        mism[k] = vm[i] * cur_p - sup_p[i] + dem_p[i];
        if (fabs(mism[k]) > stop_p) stop_p = fabs(mism[k]);
        if (is_pq) {
            mism[q] = vm[i] * cur_q - sup_q[i] + dem_q[i];
            if (fabs(mism[q]) > stop_q) stop_q = fabs(mism[q]);
        }
    }
    stop[0] = stop_p;
    stop[1] = stop_q;
}

void onr_jacobian(int64_t n, const int64_t* colptr, const int64_t* rowval, const double* ynz, const double* ynz_t,
                  const int8_t* type, int64_t slack, const int64_t* pq, const int64_t* pvpq, const int64_t* pcount,
                  const int64_t* jcolptr, const double* vm, const double* va, double* nz) {
    for (int64_t i = 0; i < n; ++i) {
        if (i == slack) continue;
        const int is_pq = type[i] == 1;
        int64_t pa = jcolptr[pvpq[i]];
        int64_t qa = pa + pcount[i];
        int64_t pm = is_pq ? jcolptr[pq[i]] : 0;
        int64_t qm = is_pq ? pm + pcount[i] : 0;
        for (int64_t j = colptr[i]; j < colptr[i + 1]; ++j) {
            const int64_t row = rowval[j];
            const int t = type[row];
            if (t == 3) continue;
            const double G = ynz[2 * j], B = ynz[2 * j + 1];
            if (row != i) {
                const double d = va[row] - va[i];
                const double s = sin(d), c = cos(d);
                nz[pa++] = vm[row] * vm[i] * (G * s - B * c);
                if (t == 1) nz[qa++] = -vm[row] * vm[i] * (G * c + B * s);
                if (is_pq) nz[pm++] = vm[row] * (G * c + B * s);
                if (is_pq && t == 1) nz[qm++] = vm[row] * (G * s - B * c);
            } else {
                double cur_t = 0.0, cur_v = 0.0;
                for (int64_t ptr = colptr[i]; ptr < colptr[i + 1]; ++ptr) {
                    const int64_t q = rowval[ptr];
                    const double Gk = ynz_t[2 * ptr], Bk = ynz_t[2 * ptr + 1];
                    const double d = va[i] - va[q];
                    const double s = sin(d), c = cos(d);
                    cur_t += vm[q] * (Gk * s - Bk * c);
                    if (is_pq) cur_v += vm[q] * (Gk * c + Bk * s);
                }
                nz[pa++] = vm[row] * (-cur_t) - B * (vm[row] * vm[row]);
                if (is_pq) {
                    nz[qa++] = vm[row] * cur_v - G * (vm[row] * vm[row]);
                    nz[pm++] = cur_v + G * vm[row];
                    nz[qm++] = cur_t - B * vm[row];
                }
            }
        }
    }
}

/* state update of solve! (acPowerFlow.jl:899-908) */
void onr_update(int64_t n, const int8_t* type, int64_t slack, const int64_t* pq, const int64_t* pvpq,
                const double* inc, double* vm, double* va) {
    for (int64_t i = 0; i < n; ++i) {
        if (type[i] == 1) vm[i] = vm[i] - inc[pq[i]];
        if (i != slack) va[i] = va[i] - inc[pvpq[i]];
    }
}
