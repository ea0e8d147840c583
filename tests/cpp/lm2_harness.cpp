// Test-only host harness: drives the product's LM state machine (csrc/fm3d_lm2.h) on a
// two-parameter curve fit so that tests/test_lm2_host.py can compare its trajectory with the
// oracle's generic lmmin restatement.  Not part of libfm3d.
#include <vector>
#include <cmath>
#include "fm3d_lm2.h"

static void model(const double* x, const double* t, const double* y, int m, double* f) {
    for (int i = 0; i < m; i++) f[i] = x[0] * std::exp(x[1] * t[i]) - y[i];
}

extern "C" int lm2_run_exp2(const double* t, const double* y, int m, double* x, double epsilon,
                            int patience, int* nfev, int* info) {
    fm3d_lm2 s;
    std::vector<double> f0(m), f1(m), f2(m);
    int cmd = fm3d_lm2_init(&s, x[0], x[1], epsilon, patience);
    while (cmd != FM3D_LM_CMD_DONE) {
        if (cmd == FM3D_LM_CMD_JAC) {
            double xa[2] = {s.x[0] + s.h[0], s.x[1]}, xb[2] = {s.x[0], s.x[1] + s.h[1]};
            model(s.x, t, y, m, f0.data());
            model(xa, t, y, m, f1.data());
            model(xb, t, y, m, f2.data());
            double ff = 0, S00 = 0, S01 = 0, S11 = 0, g0 = 0, g1 = 0;
            for (int i = 0; i < m; i++) {
                double j0 = (f1[i] - f0[i]) / s.h[0], j1 = (f2[i] - f0[i]) / s.h[1];
                ff += f0[i] * f0[i]; S00 += j0 * j0; S01 += j0 * j1; S11 += j1 * j1;
                g0 += j0 * f0[i]; g1 += j1 * f0[i];
            }
            cmd = fm3d_lm2_after_jacobian(&s, ff, S00, S01, S11, g0, g1);
        } else {
            model(s.xt, t, y, m, f0.data());
            double ff = 0;
            for (int i = 0; i < m; i++) ff += f0[i] * f0[i];
            cmd = fm3d_lm2_after_trial(&s, ff);
        }
    }
    x[0] = s.x[0]; x[1] = s.x[1];
    *nfev = s.nfev; *info = s.info;
    return 0;
}
