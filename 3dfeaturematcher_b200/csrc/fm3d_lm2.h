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
// components is R^-T J^T f).  The trust-region logic and lmpar below are MINPACK's, written
// for two unknowns on the normal equations, with lmfit's deltas (step rule max(eps^2, eps|x|), 0.55 in the shrink rule, first-call
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
    // control: lm_control_double (ftol = xtol = gtol = 30 eps, stepbound 100) with the reference's
    // epsilon override (normaloptimizer.cpp:272-274); eps = sqrt(max(epsilon, DBL_EPSILON))
    double eps;
    int maxfev;
    // state
    double x[2], xt[2], h[2];
    double diag[2], delta, par, xnorm, gnorm;
    double ff;                      // |f(x)|^2           (fnorm = sqrt(ff), kept in fnorm for callers)
    double fnorm;
    double S00, S01, S11, g0, g1;   // J^T J and J^T f at x
    double step[2], pnorm;
    int iter, nfev, info;
    int first;                      // 1 until the first Jacobian pass has been consumed
    // per-Jacobian cache for lmpar (everything that does not depend on delta)
    double gn0, gn1, gn_dx2;        // Gauss-Newton step and |D p_gn|^2
    double id0, id1;                // 1 / diag
    float a00, a01, a11, b0, b1;    // D^-1 J^T J D^-1, D^-1 J^T f
    float gnf, gn_w2;               // |b|, and the Newton denominator at par = 0 (0: rank deficient)
    // per-proposal cache for the trust-region update
    double iff, prered, dirder;     // 1/ff, predicted reduction, directional derivative
};

#define FM3D_LM_USERTOL (30.0 * FM3D_DBL_EPS)

// Cheap arithmetic for the serial LM update on the GPU (one thread, dependent chains: what counts
// is the number of instructions in a row).  Host builds use the plain operators.
//   fm3d_rcp, fm3d_sqrt   fp64 from the MUFU.RCP64H / MUFU.RSQ64H seeds + Newton steps (full
//               precision up to the last ulp, no special-case branches: operands are positive, normal)
//   fm3d_fdiv, fm3d_frcp, fm3d_fsqrt   fp32 MUFU forms (lmpar's search for par: 10 % tolerance)
#if defined(__CUDA_ARCH__)
FM3D_HD double fm3d_rcp(double a) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    r = fma(r, fma(-a, r, 1.0), r);
    r = fma(r, fma(-a, r, 1.0), r);
    return r;
}
FM3D_HD double fm3d_sqrt(double a) {
    if (!(a > 0.0)) return a == 0.0 ? 0.0 : sqrt(a);
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    // r ~ a^-1/2 (20 bits): two Newton steps on r, then s = a r with one correction
    r = r * fma(-0.5 * a * r, r, 1.5);
    r = r * fma(-0.5 * a * r, r, 1.5);
    const double sq = a * r;
    return fma(0.5 * r, fma(-sq, sq, a), sq);
}
FM3D_HD float fm3d_frcp(float a) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
FM3D_HD float fm3d_fdiv(float a, float b) { return a * fm3d_frcp(b); }
FM3D_HD float fm3d_fsqrt(float a) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
#else
FM3D_HD double fm3d_rcp(double a) { return 1.0 / a; }
FM3D_HD double fm3d_sqrt(double a) { return sqrt(a); }
FM3D_HD float fm3d_frcp(float a) { return 1.0f / a; }
FM3D_HD float fm3d_fdiv(float a, float b) { return a / b; }
FM3D_HD float fm3d_fsqrt(float a) { return sqrtf(a); }
#endif

FM3D_HD double fm3d_enorm2(double a, double b) { return fm3d_sqrt(a * a + b * b); }

// a*d - b*c with one rounding of the products compensated (Kahan)
FM3D_HD double fm3d_det2(double a, double b, double c, double d) {
    const double w = b * c;
    const double e = fma(-b, c, w);
    const double f = fma(a, d, -w);
    return f + e;
}
FM3D_HD float fm3d_det2f(float a, float b, float c, float d) {
    const float w = b * c;
    const float e = fmaf(-b, c, w);
    const float f = fmaf(a, d, -w);
    return f + e;
}

// MINPACK lmpar for n = 2, on the normal equations, reference version in fp64 (host tests compare
// the mixed-precision one below with it; it is also the fall-back when fp32 leaves its range).
// lmpar works on the pivoted QR factor R of J and calls qrsolv for every candidate parameter;
// with two unknowns R^T R = J^T J is already at hand (it is what the CTA reduces), every solve
// (J^T J + par D^2) p = J^T f is a 2x2 Cramer formula and the Newton correction
// ||R_par^-T D^2 p / |D p| ||^2 is the quadratic form of the same inverse.  The iteration (bounds
// parl/paru, 0.1 delta tolerance, at most 10 steps, rank handling of the Gauss-Newton step) is
// MINPACK's; the arithmetic is scalar: no arrays, no data-dependent indexing.
FM3D_HD double fm3d_lmpar2_f64(const fm3d_lm2* s, double delta, double par, double* p0, double* p1) {
    const double S00 = s->S00, S01 = s->S01, S11 = s->S11, g0 = s->g0, g1 = s->g1;
    const double d0 = s->diag[0], d1 = s->diag[1];
    const bool piv1 = S11 > S00;
    const double Spp = piv1 ? S11 : S00, Sqq = piv1 ? S00 : S11;
    const double gp = piv1 ? g1 : g0, gq = piv1 ? g0 : g1;
    double xp = 0.0, xq = 0.0, t = 0.0;
    bool full_rank = false;
    if (Spp != 0.0) {
        const double iS = 1.0 / Spp;
        t = Sqq - S01 * S01 * iS;               // r11^2
        if (t > 0.0) {
            full_rank = true;
            xq = (gq - S01 * gp * iS) / t;
        }
        xp = (gp - S01 * xq) * iS;
    }
    double x0 = piv1 ? xq : xp, x1 = piv1 ? xp : xq;
    double dx0 = d0 * x0, dx1 = d1 * x1;
    double dxnorm = fm3d_enorm2(dx0, dx1);
    double fp = dxnorm - delta;
    if (fp <= 0.1 * delta) { *p0 = x0; *p1 = x1; return 0.0; }
    double parl = 0.0;
    if (full_rank) {
        const double y0 = d0 * (dx0 / dxnorm), y1 = d1 * (dx1 / dxnorm);
        const double w2 = (S11 * y0 * y0 - 2.0 * S01 * y0 * y1 + S00 * y1 * y1) / (Spp * t);
        parl = fp / delta / w2;
    }
    const double gnorm = fm3d_enorm2(g0 / d0, g1 / d1);
    double paru = gnorm / delta;
    if (paru == 0.0) paru = FM3D_DBL_MIN / fmin(delta, 0.1);
    par = fmax(par, parl);
    par = fmin(par, paru);
    if (par == 0.0) par = gnorm / dxnorm;
    for (int iter = 1;; iter++) {
        if (par == 0.0) par = fmax(FM3D_DBL_MIN, 0.001 * paru);
        const double m00 = fma(par * d0, d0, S00), m11 = fma(par * d1, d1, S11);
        const double idet = 1.0 / fm3d_det2(m00, S01, S01, m11);
        x0 = fm3d_det2(m11, S01, g1, g0) * idet;    // m11 g0 - S01 g1
        x1 = fm3d_det2(m00, S01, g0, g1) * idet;    // m00 g1 - S01 g0
        dx0 = d0 * x0; dx1 = d1 * x1;
        dxnorm = fm3d_enorm2(dx0, dx1);
        const double temp = fp;
        fp = dxnorm - delta;
        if (fabs(fp) <= 0.1 * delta || (parl == 0.0 && fp <= temp && temp < 0.0) || iter == 10) break;
        const double y0 = d0 * (dx0 / dxnorm), y1 = d1 * (dx1 / dxnorm);
        const double w2 = (m11 * y0 * y0 - 2.0 * S01 * y0 * y1 + m00 * y1 * y1) * idet;
        const double parc = fp / delta / w2;
        if (fp > 0.0) parl = fmax(parl, par);
        if (fp < 0.0) paru = fmin(paru, par);
        par = fmax(parl, par + parc);
    }
    *p0 = x0; *p1 = x1;
    return par;
}

// Everything lmpar needs that does not depend on delta, computed once per Jacobian: the
// Gauss-Newton step (fp64; rank test as in the pivoted QR: r11^2 = Sqq - S01^2 / Spp > 0), and the
// problem in scaled variables z = D p in fp32 (A = D^-1 J^T J D^-1 has |a_ij| <= 1 because
// diag_j >= |J_j|, b = D^-1 J^T f has |b_j| <= |f|).
FM3D_HD void fm3d_lm2_cache_jacobian(fm3d_lm2* s) {
    const double S00 = s->S00, S01 = s->S01, S11 = s->S11, g0 = s->g0, g1 = s->g1;
    const double d0 = s->diag[0], d1 = s->diag[1];
    const bool piv1 = S11 > S00;
    const double Spp = piv1 ? S11 : S00, Sqq = piv1 ? S00 : S11;
    const double gp = piv1 ? g1 : g0, gq = piv1 ? g0 : g1;
    double xp = 0.0, xq = 0.0;
    bool full_rank = false;
    if (Spp != 0.0) {
        const double det = fm3d_det2(Spp, S01, S01, Sqq);       // Spp * r11^2
        if (det > 0.0) {
            full_rank = true;
            const double idet = fm3d_rcp(det);
            xq = fm3d_det2(Spp, S01, gp, gq) * idet;            // Spp gq - S01 gp
            xp = fm3d_det2(Sqq, S01, gq, gp) * idet;            // Sqq gp - S01 gq
        } else {
            xp = gp * fm3d_rcp(Spp);
        }
    }
    const double x0 = piv1 ? xq : xp, x1 = piv1 ? xp : xq;
    const double dx0 = d0 * x0, dx1 = d1 * x1;
    s->gn0 = x0; s->gn1 = x1;
    s->gn_dx2 = dx0 * dx0 + dx1 * dx1;
    const double id0 = fm3d_rcp(d0), id1 = fm3d_rcp(d1);
    s->id0 = id0; s->id1 = id1;
    const float a00 = (float)(S00 * id0 * id0), a01 = (float)(S01 * id0 * id1), a11 = (float)(S11 * id1 * id1);
    const float b0 = (float)(g0 * id0), b1 = (float)(g1 * id1);
    s->a00 = a00; s->a01 = a01; s->a11 = a11; s->b0 = b0; s->b1 = b1;
    s->gnf = fm3d_fsqrt(b0 * b0 + b1 * b1);
    float w2 = 0.0f;
    if (full_rank) {
        // Newton denominator at par = 0: y^T A^-1 y with y = z / |z|, z = D p_gn
        const float z0 = (float)dx0, z1 = (float)dx1;
        const float det0 = fm3d_det2f(a00, a01, a01, a11);
        const float q = fm3d_fdiv(a11 * z0 * z0 - 2.0f * a01 * z0 * z1 + a00 * z1 * z1, det0 * (z0 * z0 + z1 * z1));
        if (det0 > 0.0f && q > 0.0f) w2 = q;
    }
    s->gn_w2 = w2;
}

// lmpar with the search for par in fp32.  lmpar only asks for |D p| within 10 % of delta, so the
// Newton iteration on par does not need fp64; the step that is returned solves
// (J^T J + par D^2) p = J^T f in fp64 for the par found, and the Gauss-Newton step and its
// acceptance test are fp64 too.  Needs fm3d_lm2_cache_jacobian.
FM3D_HD double fm3d_lmpar2(const fm3d_lm2* s, double delta, double par, double* p0, double* p1) {
    // fp = |D p_gn| - delta <= 0.1 delta: the Gauss-Newton step is inside the trust region
    if (s->gn_dx2 <= 1.21 * delta * delta) { *p0 = s->gn0; *p1 = s->gn1; return 0.0; }
    const float a00 = s->a00, a01 = s->a01, a11 = s->a11, b0 = s->b0, b1 = s->b1;
    const float df = (float)delta, idf = fm3d_frcp(df);
    float dxn = fm3d_fsqrt((float)s->gn_dx2);
    float fp = dxn - df;
    const float gn_w2 = s->gn_w2;
    float parl = gn_w2 > 0.0f ? fm3d_fdiv(fp * idf, gn_w2) : 0.0f;
    const float gnorm = s->gnf;
    float paru = gnorm * idf;
    if (paru == 0.0f) paru = 1.1754944e-38f / fminf(df, 0.1f);
    float parf = (float)par;
    parf = fmaxf(parf, parl);
    parf = fminf(parf, paru);
    if (parf == 0.0f) parf = fm3d_fdiv(gnorm, dxn);
    for (int iter = 1;; iter++) {
        if (parf == 0.0f) parf = fmaxf(1.1754944e-38f, 0.001f * paru);
        const float m00 = a00 + parf, m11 = a11 + parf;
        const float idet = fm3d_frcp(fm3d_det2f(m00, a01, a01, m11));
        const float z0 = fm3d_det2f(m11, a01, b1, b0) * idet;
        const float z1 = fm3d_det2f(m00, a01, b0, b1) * idet;
        const float zn2 = z0 * z0 + z1 * z1;
        dxn = fm3d_fsqrt(zn2);
        const float temp = fp;
        fp = dxn - df;
        if (fabsf(fp) <= 0.1f * df || (parl == 0.0f && fp <= temp && temp < 0.0f) || iter == 10) break;
        const float w2 = fm3d_fdiv((m11 * z0 * z0 - 2.0f * a01 * z0 * z1 + m00 * z1 * z1) * idet, zn2);
        const float parc = fm3d_fdiv(fp * idf, w2);
        if (fp > 0.0f) parl = fmaxf(parl, parf);
        if (fp < 0.0f) paru = fminf(paru, parf);
        parf = fmaxf(parl, parf + parc);
    }
    if (!(parf >= 0.0f && parf <= 3.0e38f)) return fm3d_lmpar2_f64(s, delta, par, p0, p1);
    // the step for this par, fp64
    const double pd = (double)parf;
    const double d0 = s->diag[0], d1 = s->diag[1], S01 = s->S01;
    const double m00 = fma(pd * d0, d0, s->S00), m11 = fma(pd * d1, d1, s->S11);
    const double idet = fm3d_rcp(fm3d_det2(m00, S01, S01, m11));
    *p0 = fm3d_det2(m11, S01, s->g1, s->g0) * idet;
    *p1 = fm3d_det2(m00, S01, s->g0, s->g1) * idet;
    return pd;
}

FM3D_HD void fm3d_lm2_fd_steps(fm3d_lm2* s) {
    s->h[0] = fmax(s->eps * s->eps, s->eps * fabs(s->x[0]));
    s->h[1] = fmax(s->eps * s->eps, s->eps * fabs(s->x[1]));
}

// Returns the first command (always a Jacobian pass; its f(x) doubles as lmdif's initial
// evaluation, so the first pass accounts for 1 + 2 evaluations).
FM3D_HD int fm3d_lm2_init(fm3d_lm2* s, double phi, double theta, double epsilon, int patience) {
    s->eps = sqrt(fmax(epsilon, FM3D_DBL_EPS));
    s->maxfev = patience * 3;
    s->x[0] = phi; s->x[1] = theta;
    s->iter = 0; s->nfev = 0; s->info = 0; s->first = 1;
    s->par = 0.0; s->delta = 0.0; s->xnorm = 0.0; s->fnorm = 0.0; s->ff = 0.0; s->gnorm = 0.0;
    s->diag[0] = s->diag[1] = 0.0;
    fm3d_lm2_fd_steps(s);
    return FM3D_LM_CMD_JAC;
}

// Inner loop head: solve for the step, publish the trial point, and prepare what the
// trust-region update will need once the trial's residual is known.
FM3D_HD int fm3d_lm2_propose(fm3d_lm2* s) {
    double p0, p1;
    const double par = fm3d_lmpar2(s, s->delta, s->par, &p0, &p1);
    s->par = par;
    s->step[0] = p0; s->step[1] = p1;
    s->xt[0] = s->x[0] - p0;
    s->xt[1] = s->x[1] - p1;
    const double dp0 = s->diag[0] * p0, dp1 = s->diag[1] * p1;
    const double pn2 = dp0 * dp0 + dp1 * dp1;
    const double pnorm = fm3d_sqrt(pn2);
    s->pnorm = pnorm;
    if (s->nfev <= 1 + 2) s->delta = fmin(s->delta, pnorm);
    // |J p|^2 = p^T (J^T J) p;  prered = (|J p|^2 + 2 par |D p|^2) / |f|^2
    const double jp2 = fmax(0.0, s->S00 * p0 * p0 + 2.0 * s->S01 * p0 * p1 + s->S11 * p1 * p1);
    const double iff = s->iff;
    const double t1sq = jp2 * iff, t2sq = par * pn2 * iff;
    s->prered = t1sq + 2.0 * t2sq;
    s->dirder = -(t1sq + t2sq);
    return FM3D_LM_CMD_TRIAL;
}

// Consume the sums of a Jacobian pass: ff = |f(x)|^2, the Gram entries of the Jacobian and
// J^T f.
FM3D_HD int fm3d_lm2_after_jacobian(fm3d_lm2* s, double ff, double S00, double S01, double S11,
                                    double g0, double g1) {
    if (s->first) { s->nfev = 3; s->first = 0; } else { s->nfev += 2; }
    const double fnorm = fm3d_sqrt(ff);
    s->ff = ff; s->fnorm = fnorm;
    s->iff = fm3d_rcp(ff);
    s->S00 = S00; s->S01 = S01; s->S11 = S11; s->g0 = g0; s->g1 = g1;
    const double an0 = fm3d_sqrt(S00), an1 = fm3d_sqrt(S11);
    if (s->iter == 0) {
        s->diag[0] = an0 != 0.0 ? an0 : 1.0;
        s->diag[1] = an1 != 0.0 ? an1 : 1.0;
        s->xnorm = fm3d_enorm2(s->diag[0] * s->x[0], s->diag[1] * s->x[1]);
        s->delta = 100.0 * s->xnorm;            // stepbound
        if (s->delta == 0.0) s->delta = 100.0;
    } else {
        s->diag[0] = fmax(s->diag[0], an0);
        s->diag[1] = fmax(s->diag[1], an1);
    }
    // norm of the scaled gradient: max_j |(J^T f)_j| / (|J_j| |f|)
    double gnorm = 0.0;
    if (fnorm != 0.0) {
        if (an0 != 0.0) gnorm = fmax(gnorm, fabs(g0 * fm3d_rcp(fnorm * an0)));
        if (an1 != 0.0) gnorm = fmax(gnorm, fabs(g1 * fm3d_rcp(fnorm * an1)));
    }
    s->gnorm = gnorm;
    if (gnorm <= FM3D_LM_USERTOL) { s->info = 4; return FM3D_LM_CMD_DONE; }
    fm3d_lm2_cache_jacobian(s);
    return fm3d_lm2_propose(s);
}

// Consume the squared norm of the trial residual.
FM3D_HD int fm3d_lm2_after_trial(fm3d_lm2* s, double ff_trial) {
    const double p1 = 0.1, p0001 = 1.0e-4;
    const int nfev = s->nfev + 1;
    s->nfev = nfev;
    const double ff = s->ff;
    // actred = 1 - (fnorm1 / fnorm)^2 unless the trial is more than 10x worse
    const bool much_worse = p1 * p1 * ff_trial >= ff;          // p1 fnorm1 >= fnorm
    const double actred = much_worse ? -1.0 : 1.0 - ff_trial * s->iff;
    const double prered = s->prered, dirder = s->dirder;
    const double ratio = prered != 0.0 ? actred * fm3d_rcp(prered) : 0.0;
    const double pnorm = s->pnorm;
    double delta = s->delta, par = s->par;
    if (ratio <= 0.25) {
        double temp;
        if (actred >= 0.0) temp = 0.5;
        else temp = 0.5 * dirder * fm3d_rcp(dirder + 0.55 * actred);
        if (much_worse || temp < p1) temp = p1;
        delta = temp * fmin(delta, pnorm * 10.0);
        par *= fm3d_rcp(temp);
    } else if (par == 0.0 || ratio >= 0.75) {
        delta = pnorm * 2.0;
        par *= 0.5;
    }
    s->delta = delta; s->par = par;
    const int accepted = ratio >= p0001;
    double xnorm = s->xnorm;
    if (accepted) {
        s->x[0] = s->xt[0]; s->x[1] = s->xt[1];
        xnorm = fm3d_enorm2(s->diag[0] * s->xt[0], s->diag[1] * s->xt[1]);
        s->xnorm = xnorm;
        s->ff = ff_trial;
        s->fnorm = fm3d_sqrt(ff_trial);
        s->iff = fm3d_rcp(ff_trial);
        s->iter++;
    }
    int info = 0;
    if (fabs(actred) <= FM3D_LM_USERTOL && prered <= FM3D_LM_USERTOL && 0.5 * ratio <= 1.0) info = 1;
    if (delta <= FM3D_LM_USERTOL * xnorm) info += 2;
    if (info != 0) { s->info = info; return FM3D_LM_CMD_DONE; }
    if (nfev >= s->maxfev) info = 5;
    if (fabs(actred) <= FM3D_DBL_EPS && prered <= FM3D_DBL_EPS && 0.5 * ratio <= 1.0) info = 6;
    if (delta <= FM3D_DBL_EPS * xnorm) info = 7;
    if (s->gnorm <= FM3D_DBL_EPS) info = 8;
    if (info != 0) { s->info = info; return FM3D_LM_CMD_DONE; }
    if (!accepted) return fm3d_lm2_propose(s);
    fm3d_lm2_fd_steps(s);
    return FM3D_LM_CMD_JAC;
}

#endif  // FM3D_LM2_H_
