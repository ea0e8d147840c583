// fm3d_lm2.h -- two-parameter Levenberg-Marquardt (lmfit `lmmin` semantics) as a resumable
// state machine.
//
// The reference minimises over (phi, theta) by calling lmfit's lmmin with a callback
// (Triangulator/normaloptimizer.cpp:269-287).  On the GPU the "callback" is a CTA-wide pass
// over the disc pixels, so the solver is turned inside out: the CTA runs a pass, reduces six
// sums, and one thread advances this state machine, which answers with the next pass to run.
//
// Everything lmdif needs from the m x 2 Jacobian J and the residual f is contained in
//   ff = f.f,  S00 = J0.J0,  S01 = J0.J1,  S11 = J1.J1,  g0 = J0.f,  g1 = J1.f
// (R of the pivoted QR is the Cholesky factor of J^T J, Q^T f restricted to its first two
// components is R^-T J^T f).  The trust-region logic, lmpar and qrsolv below are MINPACK's,
// with lmfit's deltas (step rule max(eps^2, eps|x|), 0.55 in the shrink rule, first-call
// delta clip, user break -> info 11).  Host+device code; no allocation; n fixed to 2.
#ifndef FM3D_LM2_H_
#define FM3D_LM2_H_

#include <math.h>

#if defined(__CUDACC__)
#define FM3D_HD __host__ __device__ __forceinline__
#else
#define FM3D_HD inline
#endif

#define FM3D_DBL_EPS 2.220446049250313e-16
#define FM3D_DBL_MIN 2.2250738585072014e-308

enum fm3d_lm_cmd {
    FM3D_LM_CMD_JAC = 1,   // evaluate f(x), f(x+h0 e0), f(x+h1 e1) and reduce the six sums
    FM3D_LM_CMD_TRIAL = 2, // evaluate f(xt) and reduce ff
    FM3D_LM_CMD_DONE = 3   // finished: info holds the lmfit status code
};

struct fm3d_lm2 {
    // control (lm_control_struct)
    double ftol, xtol, gtol, eps, stepbound;
    int maxfev;
    // state
    double x[2], xt[2], h[2];
    double diag[2], delta, par, xnorm, fnorm, gnorm;
    double r[4], qtf[2], acnorm[2], step[2], pnorm;
    int ipvt[2];
    int iter, nfev, info;
    int first; // 1 until the first Jacobian pass has been consumed
};

FM3D_HD double fm3d_enorm2(double a, double b) { return sqrt(a * a + b * b); }

// MINPACK qrsolv, n = 2.  r: row-major 2x2 (upper = R, strict lower overwritten).
FM3D_HD void fm3d_qrsolv2(double* r, const int* ipvt, const double* diag, const double* qtb,
                          double* x, double* sdiag) {
    double wa[2];
    for (int j = 0; j < 2; j++) {
        for (int i = j; i < 2; i++) r[i * 2 + j] = r[j * 2 + i];
        x[j] = r[j * 2 + j];
        wa[j] = qtb[j];
    }
    for (int j = 0; j < 2; j++) {
        int l = ipvt[j];
        if (diag[l] != 0.0) {
            for (int k = j; k < 2; k++) sdiag[k] = 0.0;
            sdiag[j] = diag[l];
            double qtbpj = 0.0;
            for (int k = j; k < 2; k++) {
                if (sdiag[k] == 0.0) continue;
                double sn, cs;
                if (fabs(r[k * 2 + k]) < fabs(sdiag[k])) {
                    double cotan = r[k * 2 + k] / sdiag[k];
                    sn = 0.5 / sqrt(0.25 + 0.25 * cotan * cotan);
                    cs = sn * cotan;
                } else {
                    double tn = sdiag[k] / r[k * 2 + k];
                    cs = 0.5 / sqrt(0.25 + 0.25 * tn * tn);
                    sn = cs * tn;
                }
                r[k * 2 + k] = cs * r[k * 2 + k] + sn * sdiag[k];
                double temp = cs * wa[k] + sn * qtbpj;
                qtbpj = -sn * wa[k] + cs * qtbpj;
                wa[k] = temp;
                for (int i = k + 1; i < 2; i++) {
                    temp = cs * r[i * 2 + k] + sn * sdiag[i];
                    sdiag[i] = -sn * r[i * 2 + k] + cs * sdiag[i];
                    r[i * 2 + k] = temp;
                }
            }
        }
        sdiag[j] = r[j * 2 + j];
        r[j * 2 + j] = x[j];
    }
    int nsing = 2;
    for (int j = 0; j < 2; j++) {
        if (sdiag[j] == 0.0 && nsing == 2) nsing = j;
        if (nsing < 2) wa[j] = 0.0;
    }
    for (int k = 1; k <= nsing; k++) {
        int j = nsing - k;
        double sum = 0.0;
        for (int i = j + 1; i < nsing; i++) sum += r[i * 2 + j] * wa[i];
        wa[j] = (wa[j] - sum) / sdiag[j];
    }
    for (int j = 0; j < 2; j++) x[ipvt[j]] = wa[j];
}

// MINPACK lmpar, n = 2.  Returns the new par; x = step.
FM3D_HD double fm3d_lmpar2(double* r, const int* ipvt, const double* diag, const double* qtb,
                           double delta, double par, double* x) {
    double wa1[2], wa2[2], sdiag[2];
    int nsing = 2;
    for (int j = 0; j < 2; j++) {
        wa1[j] = qtb[j];
        if (r[j * 2 + j] == 0.0 && nsing == 2) nsing = j;
        if (nsing < 2) wa1[j] = 0.0;
    }
    for (int k = 1; k <= nsing; k++) {
        int j = nsing - k;
        wa1[j] /= r[j * 2 + j];
        double temp = wa1[j];
        for (int i = 0; i < j; i++) wa1[i] -= r[i * 2 + j] * temp;
    }
    for (int j = 0; j < 2; j++) x[ipvt[j]] = wa1[j];
    int iter = 0;
    for (int j = 0; j < 2; j++) wa2[j] = diag[j] * x[j];
    double dxnorm = fm3d_enorm2(wa2[0], wa2[1]);
    double fp = dxnorm - delta;
    if (fp <= 0.1 * delta) return 0.0;
    double parl = 0.0;
    if (nsing >= 2) {
        for (int j = 0; j < 2; j++) { int l = ipvt[j]; wa1[j] = diag[l] * (wa2[l] / dxnorm); }
        for (int j = 0; j < 2; j++) {
            double sum = 0.0;
            for (int i = 0; i < j; i++) sum += r[i * 2 + j] * wa1[i];
            wa1[j] = (wa1[j] - sum) / r[j * 2 + j];
        }
        double temp = fm3d_enorm2(wa1[0], wa1[1]);
        parl = fp / delta / temp / temp;
    }
    for (int j = 0; j < 2; j++) {
        double sum = 0.0;
        for (int i = 0; i <= j; i++) sum += r[i * 2 + j] * qtb[i];
        wa1[j] = sum / diag[ipvt[j]];
    }
    double gnorm = fm3d_enorm2(wa1[0], wa1[1]);
    double paru = gnorm / delta;
    if (paru == 0.0) paru = FM3D_DBL_MIN / fmin(delta, 0.1);
    par = fmax(par, parl);
    par = fmin(par, paru);
    if (par == 0.0) par = gnorm / dxnorm;
    for (;;) {
        iter++;
        if (par == 0.0) par = fmax(FM3D_DBL_MIN, 0.001 * paru);
        double temp = sqrt(par);
        for (int j = 0; j < 2; j++) wa1[j] = temp * diag[j];
        fm3d_qrsolv2(r, ipvt, wa1, qtb, x, sdiag);
        for (int j = 0; j < 2; j++) wa2[j] = diag[j] * x[j];
        dxnorm = fm3d_enorm2(wa2[0], wa2[1]);
        temp = fp;
        fp = dxnorm - delta;
        if (fabs(fp) <= 0.1 * delta || (parl == 0.0 && fp <= temp && temp < 0.0) || iter == 10) break;
        for (int j = 0; j < 2; j++) { int l = ipvt[j]; wa1[j] = diag[l] * (wa2[l] / dxnorm); }
        for (int j = 0; j < 2; j++) {
            wa1[j] /= sdiag[j];
            double t2 = wa1[j];
            for (int i = j + 1; i < 2; i++) wa1[i] -= r[i * 2 + j] * t2;
        }
        temp = fm3d_enorm2(wa1[0], wa1[1]);
        double parc = fp / delta / temp / temp;
        if (fp > 0.0) parl = fmax(parl, par);
        if (fp < 0.0) paru = fmin(paru, par);
        par = fmax(parl, par + parc);
    }
    return par;
}

FM3D_HD void fm3d_lm2_fd_steps(fm3d_lm2* s) {
    for (int j = 0; j < 2; j++) s->h[j] = fmax(s->eps * s->eps, s->eps * fabs(s->x[j]));
}

// lm_control_double with the reference's epsilon override (normaloptimizer.cpp:272-274).
// Returns the first command (always a Jacobian pass; its f(x) doubles as lmdif's initial
// evaluation, so the first pass accounts for 1 + 2 evaluations).
FM3D_HD int fm3d_lm2_init(fm3d_lm2* s, double phi, double theta, double epsilon, int patience) {
    const double usertol = 30.0 * FM3D_DBL_EPS;
    s->ftol = usertol; s->xtol = usertol; s->gtol = usertol;
    s->eps = sqrt(fmax(epsilon, FM3D_DBL_EPS));
    s->stepbound = 100.0;
    s->maxfev = patience * 3;
    s->x[0] = phi; s->x[1] = theta;
    s->iter = 0; s->nfev = 0; s->info = 0; s->first = 1;
    s->par = 0.0; s->delta = 0.0; s->xnorm = 0.0; s->fnorm = 0.0; s->gnorm = 0.0;
    s->diag[0] = s->diag[1] = 0.0;
    fm3d_lm2_fd_steps(s);
    return FM3D_LM_CMD_JAC;
}

// Inner loop head: solve for the step and publish the trial point.
FM3D_HD int fm3d_lm2_propose(fm3d_lm2* s) {
    s->par = fm3d_lmpar2(s->r, s->ipvt, s->diag, s->qtf, s->delta, s->par, s->step);
    for (int j = 0; j < 2; j++) s->xt[j] = s->x[j] - s->step[j];
    s->pnorm = fm3d_enorm2(s->diag[0] * s->step[0], s->diag[1] * s->step[1]);
    if (s->nfev <= 1 + 2) s->delta = fmin(s->delta, s->pnorm);
    return FM3D_LM_CMD_TRIAL;
}

// Consume the sums of a Jacobian pass: ff = |f(x)|^2 and the Gram entries of the
// forward-difference Jacobian (already divided by the steps).
FM3D_HD int fm3d_lm2_after_jacobian(fm3d_lm2* s, double ff, double S00, double S01, double S11,
                                    double g0, double g1) {
    if (s->first) { s->nfev = 3; s->first = 0; } else { s->nfev += 2; }
    s->fnorm = sqrt(ff);
    const double S[2] = {S00, S11};
    const double g[2] = {g0, g1};
    s->acnorm[0] = sqrt(S00); s->acnorm[1] = sqrt(S11);
    // qrfac with column pivoting (first maximum wins, as MINPACK's kmax scan)
    int p = (s->acnorm[1] > s->acnorm[0]) ? 1 : 0;
    int q = 1 - p;
    s->ipvt[0] = p; s->ipvt[1] = q;
    double r00 = s->acnorm[p], r01 = 0.0, r11 = 0.0, qt0 = 0.0, qt1 = 0.0;
    if (r00 != 0.0) {
        r01 = S01 / r00;
        double t = S[q] - r01 * r01;
        r11 = t > 0.0 ? sqrt(t) : 0.0;
        qt0 = g[p] / r00;
        if (r11 != 0.0) qt1 = (g[q] - r01 * qt0) / r11;
    }
    s->r[0] = r00; s->r[1] = r01; s->r[2] = 0.0; s->r[3] = r11;
    s->qtf[0] = qt0; s->qtf[1] = qt1;
    if (s->iter == 0) {
        for (int j = 0; j < 2; j++) s->diag[j] = s->acnorm[j] != 0.0 ? s->acnorm[j] : 1.0;
        s->xnorm = fm3d_enorm2(s->diag[0] * s->x[0], s->diag[1] * s->x[1]);
        s->delta = s->stepbound * s->xnorm;
        if (s->delta == 0.0) s->delta = s->stepbound;
    } else {
        for (int j = 0; j < 2; j++) s->diag[j] = fmax(s->diag[j], s->acnorm[j]);
    }
    double gnorm = 0.0;
    if (s->fnorm != 0.0) {
        for (int j = 0; j < 2; j++) {
            double an = s->acnorm[s->ipvt[j]];
            if (an == 0.0) continue;
            double sum = 0.0;
            for (int i = 0; i <= j; i++) sum += s->r[i * 2 + j] * s->qtf[i] / s->fnorm;
            gnorm = fmax(gnorm, fabs(sum / an));
        }
    }
    s->gnorm = gnorm;
    if (gnorm <= s->gtol) { s->info = 4; return FM3D_LM_CMD_DONE; }
    return fm3d_lm2_propose(s);
}

// Consume the squared norm of the trial residual.
FM3D_HD int fm3d_lm2_after_trial(fm3d_lm2* s, double ff_trial) {
    const double p1 = 0.1, p0001 = 1.0e-4;
    s->nfev += 1;
    double fnorm = s->fnorm, fnorm1 = sqrt(ff_trial);
    double actred = (p1 * fnorm1 < fnorm) ? 1.0 - (fnorm1 / fnorm) * (fnorm1 / fnorm) : -1.0;
    double wa3[2] = {0.0, 0.0};
    for (int j = 0; j < 2; j++) {
        wa3[j] = 0.0;
        for (int i = 0; i <= j; i++) wa3[i] -= s->r[i * 2 + j] * s->step[s->ipvt[j]];
    }
    double temp1 = fm3d_enorm2(wa3[0], wa3[1]) / fnorm;
    double temp2 = sqrt(s->par) * s->pnorm / fnorm;
    double prered = temp1 * temp1 + 2.0 * temp2 * temp2;
    double dirder = -(temp1 * temp1 + temp2 * temp2);
    double ratio = prered != 0.0 ? actred / prered : 0.0;
    if (ratio <= 0.25) {
        double temp;
        if (actred >= 0.0) temp = 0.5;
        else temp = 0.5 * dirder / (dirder + 0.55 * actred);
        if (p1 * fnorm1 >= fnorm || temp < p1) temp = p1;
        s->delta = temp * fmin(s->delta, s->pnorm / p1);
        s->par /= temp;
    } else if (s->par == 0.0 || ratio >= 0.75) {
        s->delta = s->pnorm / 0.5;
        s->par *= 0.5;
    }
    int accepted = ratio >= p0001;
    if (accepted) {
        s->x[0] = s->xt[0]; s->x[1] = s->xt[1];
        s->xnorm = fm3d_enorm2(s->diag[0] * s->x[0], s->diag[1] * s->x[1]);
        s->fnorm = fnorm1;
        s->iter++;
    }
    int info = 0;
    if (fabs(actred) <= s->ftol && prered <= s->ftol && 0.5 * ratio <= 1.0) info = 1;
    if (s->delta <= s->xtol * s->xnorm) info += 2;
    if (info != 0) { s->info = info; return FM3D_LM_CMD_DONE; }
    if (s->nfev >= s->maxfev) info = 5;
    if (fabs(actred) <= FM3D_DBL_EPS && prered <= FM3D_DBL_EPS && 0.5 * ratio <= 1.0) info = 6;
    if (s->delta <= FM3D_DBL_EPS * s->xnorm) info = 7;
    if (s->gnorm <= FM3D_DBL_EPS) info = 8;
    if (info != 0) { s->info = info; return FM3D_LM_CMD_DONE; }
    if (!accepted) return fm3d_lm2_propose(s);
    fm3d_lm2_fd_steps(s);
    return FM3D_LM_CMD_JAC;
}

#endif  // FM3D_LM2_H_
