"""Local solvers (oracle; test-only).

eigsolve_solver   src/local_solvers/eigsolve.jl:3-29  -> UPSTREAM KrylovKit.eigsolve (App. A.6)
exponentiate_solver src/local_solvers/exponentiate.jl:3-29 -> UPSTREAM KrylovKit.exponentiate (A.7)
runge_kutta_2/4/solver  src/local_solvers/runge_kutta.jl:2-25 (verbatim arithmetic)
"""
from __future__ import annotations

import numpy as np
import scipy.linalg

from .tensor import Tensor, inner


def _axpy(a, x: Tensor, y: Tensor) -> Tensor:
    return Tensor(y.data + a * x.array(y.labels), y.labels)


def lanczos_eigsolve(operator, init: Tensor, *, krylovdim=3, maxiter=1, tol=1e-14, which="SR", eager=False):
    """KrylovKit.eigsolve(f, x0, 1, :SR; ishermitian=true, krylovdim, maxiter=1, tol, eager=false).

    Lanczos with the full basis kept and two-pass modified Gram-Schmidt against all previous
    vectors.  Exactly `krylovdim` matvecs unless the residual drops below tol first.  Only
    maxiter=1 (no restart) is restated -- it is what the reference uses.
    Returns (value, vector, info)."""
    assert maxiter == 1, "only the no-restart case is restated"
    n = init.data.size
    kmax = min(krylovdim, n)
    V = []
    alphas, betas = [], []
    v = init / init.norm()
    nmv = 0
    beta_prev = 0.0
    while True:
        V.append(v)
        w = operator(v)
        nmv += 1
        w = Tensor(w.array(v.labels), v.labels)
        alpha = inner(v, w)
        w = _axpy(-alpha, v, w)
        if len(V) > 1:
            w = _axpy(-beta_prev, V[-2], w)
        # full re-orthogonalisation, second MGS pass (corrections folded into alpha)
        for q in V:
            c = inner(q, w)
            w = _axpy(-c, q, w)
            if q is v:
                alpha = alpha + c
        alphas.append(float(np.real(alpha)))
        beta = w.norm()
        K = len(V)
        if K == kmax or beta <= tol:
            break
        if eager:
            # KrylovKit: `eager && K >= howmany` only triggers an early Ritz / convergence test; the loop is
            # left when the wanted Ritz pair's residual |beta y_K| <= tol, otherwise the basis keeps growing
            Te = np.diag(alphas)
            for i, b in enumerate(betas[: K - 1]):
                Te[i, i + 1] = Te[i + 1, i] = b
            _, ve = np.linalg.eigh(Te)
            if abs(beta * ve[-1, 0 if which == "SR" else K - 1]) <= tol:
                break
        betas.append(beta)
        beta_prev = beta
        v = w / beta
    K = len(V)
    T = np.diag(alphas)
    for i, b in enumerate(betas[: K - 1]):
        T[i, i + 1] = T[i + 1, i] = b
    vals, vecs = np.linalg.eigh(T)
    idx = 0 if which == "SR" else K - 1
    y = vecs[:, idx]
    x = Tensor(np.zeros_like(V[0].data, dtype=np.result_type(V[0].data, y)), V[0].labels)
    for yi, q in zip(y, V):
        x = _axpy(yi, q, x)
    info = {"numops": nmv, "residual": abs(beta * y[-1]), "krylov": K}
    return float(vals[idx]), x, info


def eigsolve_solver(operator, init, howmany=1, *, which_eigval="SR", ishermitian=True, tol=1e-14,
                    krylovdim=3, maxiter=1, verbosity=0, eager=False, **kws):
    val, vec, _ = lanczos_eigsolve(operator, init, krylovdim=krylovdim, maxiter=maxiter, tol=tol,
                                   which=which_eigval, eager=eager)
    return val, vec


def runge_kutta_2(H, t, psi0):
    Hpsi = H(psi0)
    H2psi = H(Hpsi)
    return psi0 + t * Hpsi + (t**2 / 2) * H2psi


def runge_kutta_4(H, t, psi0):
    k1 = H(psi0)
    k2 = k1 + (t / 2) * H(k1)
    k3 = k1 + (t / 2) * H(k2)
    k4 = k1 + t * H(k3)
    return psi0 + (t / 6) * (k1 + 2 * k2 + 2 * k3 + k4)


def runge_kutta_solver(H, time, psi, *, order=4, **kws):
    if order == 4:
        out = runge_kutta_4(H, time, psi)
    elif order == 2:
        out = runge_kutta_2(H, time, psi)
    else:
        raise ValueError("For runge_kutta_solver, must specify `order` keyword")
    return out, {}


def exponentiate_solver(operator, time, init: Tensor, *, krylovdim=30, maxiter=100, verbosity=0,
                        tol=1e-12, ishermitian=True, issymmetric=True, eager=True, **kws):
    """KrylovKit.exponentiate(f, t, x0) = expintegrator with p = 1 (App. A.7), Lanczos variant.

    exp(t A) u0 = u0 + t * phi_1(t A) A u0; the Krylov space is built from A u0; adaptive
    sub-stepping with safety factors delta = 1.2, gamma = 0.8; eta = tol / |t| per unit time."""
    labels = init.labels
    t = complex(time)
    tau = abs(t)
    if tau == 0.0:
        return init.copy(), {"numops": 0, "converged": 1}
    sgn = t / tau
    u0 = init
    w0 = Tensor(init.data.astype(np.result_type(init.data, sgn if sgn.imag != 0 else sgn.real)), labels)
    numops = 0
    eta = tol / tau
    totalerr = 0.0
    tau0 = 0.0
    dtau = tau
    gamma = 0.8
    numiter = 1
    n = init.data.size

    def apply(x):
        y = operator(x)
        return Tensor(y.array(labels), labels)

    w1 = apply(w0)
    numops += 1
    while True:
        beta = w1.norm()
        if beta < tol:
            return w0, {"numops": numops, "converged": 1, "error": totalerr}
        # Lanczos factorisation started from w1 / beta
        V = [w1 / beta]
        alphas, betas = [], []
        r = None
        resnorm = None

        def expand():
            nonlocal r, resnorm, numops
            v = V[-1]
            w = apply(v)
            numops += 1
            a = inner(v, w)
            w = _axpy(-a, v, w)
            if len(V) > 1:
                w = _axpy(-betas[-1], V[-2], w)
            for q in V:
                c = inner(q, w)
                w = _axpy(-c, q, w)
                if q is v:
                    a = a + c
            alphas.append(float(np.real(a)))
            r = w
            resnorm = w.norm()

        expand()
        kmax = min(krylovdim, n)
        while True:
            K = len(V)
            stepped = False

            def small_exp(dt):
                Hs = np.zeros((K + 2, K + 2), dtype=complex)
                Tm = np.diag(alphas[:K]).astype(complex)
                for i, b in enumerate(betas[: K - 1]):
                    Tm[i, i + 1] = Tm[i + 1, i] = b
                Hs[:K, :K] = (sgn * dt) * Tm       # only the Rayleigh block carries the step
                Hs[0, K] = 1.0
                Hs[K, K + 1] = 1.0
                E = scipy.linalg.expm(Hs)
                eps = abs(dt * beta * resnorm * E[K - 1, K + 1])
                return E, eps

            if K == kmax:
                dtau = min(dtau, tau - tau0)
                E, eps = small_exp(dtau)
                omega = eps / (dtau * eta)
                q = K / 2
                while omega > 1.0:
                    eps_prev, dtau_prev = eps, dtau
                    dtau *= (gamma / omega) ** (1.0 / (q + 1))
                    E, eps = small_exp(dtau)
                    omega = eps / (dtau * eta)
                    if eps <= 0.0:
                        break
                    q = max(0.0, np.log(eps / eps_prev) / np.log(dtau / dtau_prev) - 1)
                step = dtau
                stepped = True
            elif resnorm <= (tau - tau0) * eta or eager:
                step = tau - tau0
                E, eps = small_exp(step)
                omega = eps / (step * eta)
                if omega < 1.0:
                    stepped = True
            if stepped:
                totalerr += eps
                coeff = E[:K, K]
                wp = Tensor(np.zeros(init.data.shape, dtype=complex), labels)
                for ci, qv in zip(coeff, V):
                    wp = _axpy(ci, qv, wp)
                wp = _axpy(E[K - 1, K + 1], r, wp)
                w0 = _axpy(beta * sgn * step, wp, Tensor(w0.data.astype(complex), labels))
                tau0 += step
                if K == kmax and omega < gamma:
                    dtau *= (gamma / max(omega, 1e-300)) ** (1.0 / (q + 1))
            if tau0 >= tau * (1 - 1e-15):
                if not np.iscomplexobj(init.data) and sgn.imag == 0:
                    w0 = Tensor(np.real(w0.data), labels)
                return w0, {"numops": numops, "converged": 1, "error": totalerr, "numiter": numiter}
            if stepped:
                break  # restart Krylov space from the propagated vector
            if K < kmax and resnorm > 0:
                betas.append(resnorm)
                V.append(r / resnorm)
                expand()
            else:
                break
        if numiter == maxiter:
            return w0, {"numops": numops, "converged": 0, "error": totalerr, "numiter": numiter}
        numiter += 1
        w1 = apply(w0)
        numops += 1
