/* CPU oracle, C restatement of normalEquation! for the measurement codes of the benchmark configuration
 * (TEST INFRASTRUCTURE / CPU baseline only): voltmeters (1), bus wattmeters / varmeters (6, 9), branch wattmeters /
 * varmeters (7, 8, 10, 11) and rectangular bus PMUs (16, 17).  Follows
 * /root/reference/src/stateEstimation/acStateEstimation.jl:261-583 and src/backend/equations.jl:63-276, 550-573;
 * like the reference it writes the Jacobian H in CSC order through a per-entry position lookup, here precomputed
 * per row (`slot`), and accumulates the objective (equations.jl:689-698, diagonal precision).
 * Validated against oracle/wls.py by tests/test_host.py.  0-based indices, complex arrays interleaved. */
#include <math.h>
#include <stdint.h>

double owls_normal_equation(int64_t n, int64_t m, const int64_t* ycolptr, const int64_t* yrowval, const double* ynz,
                            const double* ynz_t, const int64_t* ydiag, const int64_t* br_from, const int64_t* br_to,
                            const double* br_g, const double* br_b, const double* br_gsi, const double* br_bsi,
                            const double* br_tinv, const double* br_phi, const int8_t* type, const int64_t* index,
                            const int64_t* slotptr, const int64_t* slotpos, const double* wdiag, const double* mean,
                            const double* vm, const double* va, double* residual, double* hnz) {
    double objective = 0.0;
    for (int64_t row = 0; row < m; ++row) {
        const int code = type[row];
        if (code == 0) continue;
        const int64_t k = index[row];
        const int64_t* slot = slotpos + slotptr[row];
        double h;
        if (code == 1) {
            h = vm[k];
        } else if (code == 6 || code == 9) {
            const int64_t i = k;
            double sp = 0.0, sm = 0.0;
            int64_t dslot = 0;
            for (int64_t p = ycolptr[i]; p < ycolptr[i + 1]; ++p) {
                const int64_t j = yrowval[p];
                const double G = ynz_t[2 * p], B = ynz_t[2 * p + 1];
                const double d = va[i] - va[j];
                const double s = sin(d), c = cos(d);
                sp += vm[j] * (G * c + B * s);
                sm += vm[j] * (G * s - B * c);
                const int64_t q = 2 * (p - ycolptr[i]);
                if (j == i) { dslot = q; continue; }
                if (code == 6) {
                    hnz[slot[q]] = vm[i] * vm[j] * (G * s - B * c);
                    hnz[slot[q + 1]] = vm[i] * (G * c + B * s);
                } else {
                    hnz[slot[q]] = -vm[i] * vm[j] * (G * c + B * s);
                    hnz[slot[q + 1]] = vm[i] * (G * s - B * c);
                }
            }
            const double Gii = ynz[2 * ydiag[i]], Bii = ynz[2 * ydiag[i] + 1];
            if (code == 6) {
                h = vm[i] * sp;
                hnz[slot[dslot]] = vm[i] * (-sm) - Bii * (vm[i] * vm[i]);
                hnz[slot[dslot + 1]] = sp + Gii * vm[i];
            } else {
                h = vm[i] * sm;
                hnz[slot[dslot]] = vm[i] * sp - Gii * (vm[i] * vm[i]);
                hnz[slot[dslot + 1]] = sm - Bii * vm[i];
            }
        } else if (code == 16 || code == 17) {
            const double s = sin(va[k]), c = cos(va[k]);
            if (code == 16) { h = vm[k] * c; hnz[slot[0]] = -vm[k] * s; hnz[slot[1]] = c; }
            else { h = vm[k] * s; hnz[slot[0]] = vm[k] * c; hnz[slot[1]] = s; }
        } else {
            const int64_t i = br_from[k], j = br_to[k];
            const double g = br_g[k], b = br_b[k], gsi = br_gsi[k], bsi = br_bsi[k], tinv = br_tinv[k];
            const double Vi = vm[i], Vj = vm[j];
            const double d = va[i] - va[j] - br_phi[k];
            const double s = sin(d), c = cos(d);
            double dti, dvi, dtj, dvj;
            if (code == 7) {
                const double A = tinv * tinv * (g + gsi), B = tinv * g, C = tinv * b;
                h = A * (Vi * Vi) - (B * c + C * s) * Vi * Vj;
                dti = (B * s - C * c) * Vi * Vj; dvi = 2 * A * Vi - (B * c + C * s) * Vj;
                dtj = -dti; dvj = -(B * c + C * s) * Vi;
            } else if (code == 8) {
                const double A = g + gsi, B = tinv * g, C = tinv * b;
                h = A * (Vj * Vj) - (B * c - C * s) * Vi * Vj;
                dti = (B * s + C * c) * Vi * Vj; dvi = (-B * c + C * s) * Vj;
                dtj = -dti; dvj = 2 * A * Vj - (B * c - C * s) * Vi;
            } else if (code == 10) {
                const double A = tinv * tinv * (b + bsi), B = tinv * g, C = tinv * b;
                h = -A * (Vi * Vi) - (B * s - C * c) * Vi * Vj;
                dti = -(B * c + C * s) * Vi * Vj; dvi = -2 * A * Vi - (B * s - C * c) * Vj;
                dtj = -dti; dvj = -(B * s - C * c) * Vi;
            } else if (code == 11) {
                const double A = b + bsi, B = tinv * g, C = tinv * b;
                h = -A * (Vj * Vj) + (B * s + C * c) * Vi * Vj;
                dti = (B * c - C * s) * Vi * Vj; dvi = (B * s + C * c) * Vj;
                dtj = -dti; dvj = -2 * A * Vj + (B * s + C * c) * Vi;
            } else {
                return NAN;   /* code outside the benchmark set: use the NumPy oracle */
            }
            hnz[slot[0]] = dti; hnz[slot[1]] = dvi; hnz[slot[2]] = dtj; hnz[slot[3]] = dvj;
        }
        residual[row] = mean[row] - h;
        objective += residual[row] * residual[row] * wdiag[row];
    }
    return objective;
}
