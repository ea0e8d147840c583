"""TEST INFRASTRUCTURE -- CPU restatement of lmfit's `lmmin` (Levenberg-Marquardt).

The reference calls `lmmin(n_par, par, m_dat, &data, evaluateNormal, lm_printout_std,
&control, 0, &status)` (Triangulator/normaloptimizer.cpp:280-281) from the lmfit library
(J. Wuttke), which is NOT vendored in the reference tree and NOT version-pinned
(CMakeLists.txt:25,29 link `lmfit` from /usr/local/lib; the 9-argument call shape with an
`lm_princon_struct` places it in the lmfit 3.x/4.x line, 2010-2013).  **Parity of the LM
trajectory is therefore unpinned**: this file restates the published algorithm of that
line (a C translation of MINPACK `lmdif`/`lmpar`/`qrfac`/`qrsolv` with lmfit's deltas):

  * forward-difference step  step_j = max(eps^2, eps*|x_j|), eps = sqrt(max(epsilon, DBL_EPSILON))
    (MINPACK: eps*|x_j|, or eps if x_j == 0)
  * the trust-region shrink uses 0.5*dirder/(dirder + 0.55*actred) (MINPACK: 0.5*actred)
  * the initial step bound is clipped to pnorm only while nfev <= 1+n (MINPACK: during the
    whole first outer iteration)
  * diag (scale_diag=1): first iteration = column norms (1 if 0), later max(diag, norm)
  * max evaluations = patience*(n+1)
  * a negative *info from the user function stops immediately with status.info = 11
  * lm_control_double = {ftol=xtol=gtol=30*DBL_EPSILON, epsilon=30*DBL_EPSILON,
    stepbound=100, patience=100, scale_diag=1}; the reference overrides epsilon
    (normaloptimizer.cpp:274).

What pins it instead: with `minpack_mode=True` (MINPACK's step rule) this restatement must
reproduce `scipy.optimize.leastsq` (the original Fortran `lmdif`) -- same nfev, same
solution -- which tests/test_oracle_lm.py checks.  Only tests/, bench.py's cpu_baseline leg
and __graft_entry__.smoke() may import this module.
"""
from __future__ import annotations

import math

import numpy as np

DBL_EPSILON = 2.220446049250313e-16
DBL_MIN = 2.2250738585072014e-308
LM_USERTOL = 30.0 * DBL_EPSILON


class LMControl:
    def __init__(self, ftol=LM_USERTOL, xtol=LM_USERTOL, gtol=LM_USERTOL, epsilon=LM_USERTOL,
                 stepbound=100.0, patience=100, scale_diag=1):
        self.ftol, self.xtol, self.gtol = ftol, xtol, gtol
        self.epsilon, self.stepbound = epsilon, stepbound
        self.patience, self.scale_diag = patience, scale_diag


def enorm(v) -> float:
    v = np.asarray(v, dtype=np.float64)
    return math.sqrt(float(np.dot(v, v)))


def qrfac(a: np.ndarray, pivot: bool):
    """MINPACK qrfac on a (m x n, modified in place). Returns ipvt, rdiag, acnorm."""
    m, n = a.shape
    acnorm = np.array([enorm(a[:, j]) for j in range(n)])
    rdiag = acnorm.copy()
    wa = rdiag.copy()
    ipvt = list(range(n))
    for j in range(min(m, n)):
        if pivot:
            kmax = j
            for k in range(j, n):
                if rdiag[k] > rdiag[kmax]:
                    kmax = k
            if kmax != j:
                a[:, [j, kmax]] = a[:, [kmax, j]]
                rdiag[kmax] = rdiag[j]
                wa[kmax] = wa[j]
                ipvt[j], ipvt[kmax] = ipvt[kmax], ipvt[j]
        ajnorm = enorm(a[j:, j])
        if ajnorm == 0.0:
            rdiag[j] = 0.0
            continue
        if a[j, j] < 0.0:
            ajnorm = -ajnorm
        a[j:, j] /= ajnorm
        a[j, j] += 1.0
        for k in range(j + 1, n):
            s = float(np.dot(a[j:, j], a[j:, k]))
            temp = s / a[j, j]
            a[j:, k] -= temp * a[j:, j]
            if pivot and rdiag[k] != 0.0:
                temp = a[j, k] / rdiag[k]
                temp = max(0.0, 1.0 - temp * temp)
                rdiag[k] *= math.sqrt(temp)
                temp = rdiag[k] / wa[k]
                if 0.05 * temp * temp <= DBL_EPSILON:
                    rdiag[k] = enorm(a[j + 1:, k])
                    wa[k] = rdiag[k]
        rdiag[j] = -ajnorm
    return ipvt, rdiag, acnorm


def qrsolv(n, r, ipvt, diag, qtb):
    """MINPACK qrsolv. r: n x n (upper triangle = R; lower is overwritten with S^T).
    Returns x, sdiag."""
    x = np.zeros(n)
    sdiag = np.zeros(n)
    wa = np.zeros(n)
    for j in range(n):
        for i in range(j, n):
            r[i, j] = r[j, i]
        x[j] = r[j, j]
        wa[j] = qtb[j]
    for j in range(n):
        l = ipvt[j]
        if diag[l] != 0.0:
            sdiag[j:] = 0.0
            sdiag[j] = diag[l]
            qtbpj = 0.0
            for k in range(j, n):
                if sdiag[k] == 0.0:
                    continue
                if abs(r[k, k]) < abs(sdiag[k]):
                    cotan = r[k, k] / sdiag[k]
                    sin = 0.5 / math.sqrt(0.25 + 0.25 * cotan * cotan)
                    cos = sin * cotan
                else:
                    tan = sdiag[k] / r[k, k]
                    cos = 0.5 / math.sqrt(0.25 + 0.25 * tan * tan)
                    sin = cos * tan
                r[k, k] = cos * r[k, k] + sin * sdiag[k]
                temp = cos * wa[k] + sin * qtbpj
                qtbpj = -sin * wa[k] + cos * qtbpj
                wa[k] = temp
                for i in range(k + 1, n):
                    temp = cos * r[i, k] + sin * sdiag[i]
                    sdiag[i] = -sin * r[i, k] + cos * sdiag[i]
                    r[i, k] = temp
        sdiag[j] = r[j, j]
        r[j, j] = x[j]
    nsing = n
    for j in range(n):
        if sdiag[j] == 0.0 and nsing == n:
            nsing = j
        if nsing < n:
            wa[j] = 0.0
    for k in range(1, nsing + 1):
        j = nsing - k
        s = 0.0
        for i in range(j + 1, nsing):
            s += r[i, j] * wa[i]
        wa[j] = (wa[j] - s) / sdiag[j]
    for j in range(n):
        x[ipvt[j]] = wa[j]
    return x, sdiag


def lmpar(n, r, ipvt, diag, qtb, delta, par):
    """MINPACK lmpar. Returns par, x, sdiag."""
    dwarf = DBL_MIN
    wa1 = np.zeros(n)
    x = np.zeros(n)
    sdiag = np.zeros(n)
    nsing = n
    for j in range(n):
        wa1[j] = qtb[j]
        if r[j, j] == 0.0 and nsing == n:
            nsing = j
        if nsing < n:
            wa1[j] = 0.0
    for k in range(1, nsing + 1):
        j = nsing - k
        wa1[j] /= r[j, j]
        temp = wa1[j]
        for i in range(j):
            wa1[i] -= r[i, j] * temp
    for j in range(n):
        x[ipvt[j]] = wa1[j]
    it = 0
    wa2 = diag * x
    dxnorm = enorm(wa2)
    fp = dxnorm - delta
    if fp <= 0.1 * delta:
        return 0.0, x, sdiag
    parl = 0.0
    if nsing >= n:
        for j in range(n):
            l = ipvt[j]
            wa1[j] = diag[l] * (wa2[l] / dxnorm)
        for j in range(n):
            s = 0.0
            for i in range(j):
                s += r[i, j] * wa1[i]
            wa1[j] = (wa1[j] - s) / r[j, j]
        temp = enorm(wa1)
        parl = fp / delta / temp / temp
    for j in range(n):
        s = 0.0
        for i in range(j + 1):
            s += r[i, j] * qtb[i]
        l = ipvt[j]
        wa1[j] = s / diag[l]
    gnorm = enorm(wa1)
    paru = gnorm / delta
    if paru == 0.0:
        paru = dwarf / min(delta, 0.1)
    par = max(par, parl)
    par = min(par, paru)
    if par == 0.0:
        par = gnorm / dxnorm
    while True:
        it += 1
        if par == 0.0:
            par = max(dwarf, 0.001 * paru)
        temp = math.sqrt(par)
        wa1 = temp * diag
        x, sdiag = qrsolv(n, r, ipvt, wa1, qtb)
        wa2 = diag * x
        dxnorm = enorm(wa2)
        temp = fp
        fp = dxnorm - delta
        if abs(fp) <= 0.1 * delta or (parl == 0.0 and fp <= temp and temp < 0.0) or it == 10:
            break
        for j in range(n):
            l = ipvt[j]
            wa1[j] = diag[l] * (wa2[l] / dxnorm)
        for j in range(n):
            wa1[j] /= sdiag[j]
            temp = wa1[j]
            for i in range(j + 1, n):
                wa1[i] -= r[i, j] * temp
        temp = enorm(wa1)
        parc = fp / delta / temp / temp
        if fp > 0.0:
            parl = max(parl, par)
        if fp < 0.0:
            paru = min(paru, par)
        par = max(parl, par + parc)
    return par, x, sdiag


class LMStatus:
    def __init__(self):
        self.info = 0
        self.nfev = 0
        self.fnorm = 0.0
        self.trace = []  # (nfev, x copy, fnorm) at every accepted iterate


def lmmin(n, par, m, evaluate, control: LMControl | None = None, minpack_mode=False,
          keep_trace=False):
    """Minimise sum(fvec^2). evaluate(par) -> (fvec (m,), info); info < 0 requests a break.

    Returns (par, status). status.info follows lmfit: 1-3 converged, 4 gtol, 5 maxfev,
    6-8 tolerances too small, 10 bad input, 11 break requested by the user function.
    """
    c = control or LMControl()
    st = LMStatus()
    x = np.array(par, dtype=np.float64)
    maxfev = c.patience * (n + 1)
    ftol, xtol, gtol, factor = c.ftol, c.xtol, c.gtol, c.stepbound
    eps = math.sqrt(max(c.epsilon, DBL_EPSILON))
    if n <= 0 or m < n or ftol < 0 or xtol < 0 or gtol < 0 or maxfev <= 0 or factor <= 0:
        st.info = 10
        return x, st
    p1, p0001 = 0.1, 1.0e-4
    diag = np.zeros(n)
    itr = 0
    par_lm = 0.0
    delta = 0.0
    xnorm = 0.0

    fvec, info = evaluate(x)
    st.nfev += 1
    if info < 0:
        st.info = 11
        return x, st
    fvec = np.array(fvec, dtype=np.float64)
    fnorm = enorm(fvec)

    def finish(code):
        st.info = code
        st.fnorm = fnorm
        return x, st

    while True:
        # Jacobian by forward differences
        fjac = np.zeros((m, n))
        for j in range(n):
            temp = x[j]
            if minpack_mode:
                step = eps * abs(temp)
                if step == 0.0:
                    step = eps
            else:
                step = max(eps * eps, eps * abs(temp))
            x[j] = temp + step
            wa4, info = evaluate(x)
            st.nfev += 1
            if info < 0:
                x[j] = temp
                st.info = 11
                st.fnorm = fnorm
                return x, st
            fjac[:, j] = (np.asarray(wa4, dtype=np.float64) - fvec) / step
            x[j] = temp
        ipvt, rdiag, acnorm = qrfac(fjac, True)
        wa1 = rdiag
        wa2 = acnorm
        if itr == 0:
            if c.scale_diag:
                for j in range(n):
                    diag[j] = wa2[j] if wa2[j] != 0.0 else 1.0
            else:
                diag[:] = 1.0
            xnorm = enorm(diag * x)
            delta = factor * xnorm
            if delta == 0.0:
                delta = factor
        else:
            if c.scale_diag:
                diag = np.maximum(diag, wa2)
        # (Q^T fvec)[:n]
        wa4 = fvec.copy()
        qtf = np.zeros(n)
        for j in range(n):
            temp3 = fjac[j, j]
            if temp3 != 0.0:
                s = float(np.dot(fjac[j:, j], wa4[j:]))
                temp = -s / temp3
                wa4[j:] += fjac[j:, j] * temp
            fjac[j, j] = wa1[j]
            qtf[j] = wa4[j]
        # norm of the scaled gradient
        gnorm = 0.0
        if fnorm != 0.0:
            for j in range(n):
                if wa2[ipvt[j]] == 0.0:
                    continue
                s = 0.0
                for i in range(j + 1):
                    s += fjac[i, j] * qtf[i] / fnorm
                gnorm = max(gnorm, abs(s / wa2[ipvt[j]]))
        if gnorm <= gtol:
            return finish(4)
        r = np.zeros((n, n))
        for j in range(n):
            for i in range(j + 1):
                r[i, j] = fjac[i, j]
        while True:
            par_lm, wa1s, _ = lmpar(n, r, ipvt, diag, qtf, delta, par_lm)
            # lmpar returns the step x; the new point is x - step
            wa2n = x - wa1s
            wa3 = diag * wa1s
            pnorm = enorm(wa3)
            # lmfit: "at first call"; MINPACK: "on the first iteration" (iter == 1)
            if (itr == 0) if minpack_mode else (st.nfev <= 1 + n):
                delta = min(delta, pnorm)
            wa4, info = evaluate(wa2n)
            st.nfev += 1
            if info < 0:
                st.info = 11
                st.fnorm = fnorm
                return x, st
            wa4 = np.asarray(wa4, dtype=np.float64)
            fnorm1 = enorm(wa4)
            if p1 * fnorm1 < fnorm:
                actred = 1.0 - (fnorm1 / fnorm) ** 2
            else:
                actred = -1.0
            w3 = np.zeros(n)
            for j in range(n):
                w3[j] = 0.0
                for i in range(j + 1):
                    w3[i] -= r[i, j] * wa1s[ipvt[j]]
            temp1 = enorm(w3) / fnorm
            temp2 = math.sqrt(par_lm) * pnorm / fnorm
            prered = temp1 * temp1 + 2.0 * temp2 * temp2
            dirder = -(temp1 * temp1 + temp2 * temp2)
            ratio = actred / prered if prered != 0.0 else 0.0
            if ratio <= 0.25:
                if actred >= 0.0:
                    temp = 0.5
                else:
                    # lmfit carries 0.55 here where MINPACK has 0.5 (p5)
                    temp = 0.5 * dirder / (dirder + (0.5 if minpack_mode else 0.55) * actred)
                if p1 * fnorm1 >= fnorm or temp < p1:
                    temp = p1
                delta = temp * min(delta, pnorm / p1)
                par_lm /= temp
            elif par_lm == 0.0 or ratio >= 0.75:
                delta = pnorm / 0.5
                par_lm *= 0.5
            if ratio >= p0001:
                x = wa2n.copy()
                xnorm = enorm(diag * x)
                fvec = wa4
                fnorm = fnorm1
                itr += 1
                if keep_trace:
                    st.trace.append((st.nfev, x.copy(), fnorm))
            info_c = 0
            if abs(actred) <= ftol and prered <= ftol and 0.5 * ratio <= 1.0:
                info_c = 1
            if delta <= xtol * xnorm:
                info_c += 2
            if info_c != 0:
                return finish(info_c)
            if st.nfev >= maxfev:
                info_c = 5
            if abs(actred) <= DBL_EPSILON and prered <= DBL_EPSILON and 0.5 * ratio <= 1.0:
                info_c = 6
            if delta <= DBL_EPSILON * xnorm:
                info_c = 7
            if gnorm <= DBL_EPSILON:
                info_c = 8
            if info_c != 0:
                return finish(info_c)
            if ratio >= p0001:
                break
