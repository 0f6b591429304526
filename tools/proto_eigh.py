"""NumPy prototype of the device Hermitian eigensolver (csrc/eigh.cu): blocked Householder tridiagonalisation
(panel recurrences + rank-2k trailing update), Cuppen divide & conquer on the tridiagonal matrix (host deflation,
secular equation with the origin shifted to the nearer pole, Gu-Eisenstat recomputation of z), and the compact-WY
back-transformation.  Development scaffolding: the CUDA code follows these formulas statement by statement.

    python tools/proto_eigh.py
"""
import numpy as np

EPS = np.finfo(float).eps


# ----------------------------------------------------------------------------------------------------------
# stage 1: A = Q T Q^H, full storage, "lower" variant; reflector j annihilates A[j+2:, j]
# ----------------------------------------------------------------------------------------------------------
def house_gen(x):
    """x (length m >= 1): returns v (v[0] = 1), tau, beta with (I - tau v v^H)^H x = beta e_0, beta real."""
    alpha = x[0]
    sigma = np.sum(np.abs(x[1:]) ** 2)
    v = np.zeros_like(x)
    v[0] = 1.0
    if sigma == 0.0 and np.imag(alpha) == 0.0:
        return v, 0.0 * alpha, np.real(alpha)
    beta = -np.copysign(np.sqrt(np.abs(alpha) ** 2 + sigma), np.real(alpha))
    tau = (beta - alpha) / beta
    if not np.iscomplexobj(x):
        tau = np.real(tau)
    v[1:] = x[1:] / (alpha - beta)
    return v, tau, beta


def sytrd_blocked(A, nb=8):
    """Returns d, e (real), Vall (n x n, column j = reflector j with explicit 1 at row j+1, zeros above), taus.
    H_j = I - tau_j v_j v_j^H;  Q = H_0 H_1 ... H_{n-3};  T = Q^H A Q  (LAPACK zhetrd 'L' convention)."""
    A = np.array(A, copy=True)
    n = A.shape[0]
    cplx = np.iscomplexobj(A)
    d = np.zeros(n)
    e = np.zeros(max(n - 1, 0))
    Vall = np.zeros_like(A)
    taus = np.zeros(n, dtype=A.dtype)
    p = 0
    nref = n - 1                      # reflector n-2 is trivial for a real matrix, a phase for a complex one
    while p < nref:
        w = min(nb, nref - p)
        V = np.zeros((n, w), dtype=A.dtype)
        W = np.zeros((n, w), dtype=A.dtype)
        for i in range(w):
            j = p + i
            # (1) bring column j up to date with the reflectors of this panel
            if i > 0:
                A[j:, j] -= V[j:, :i] @ np.conj(W[j, :i]) + W[j:, :i] @ np.conj(V[j, :i])
            d[j] = np.real(A[j, j])
            # (2) reflector from A[j+1:, j]
            v, tau, beta = house_gen(A[j + 1:, j])
            e[j] = beta
            V[j + 1:, i] = v
            taus[j] = tau
            # (3) w = tau * (A_trail v - V (W^H v) - W (V^H v));  the trailing square is still the panel-start matrix
            vv = V[j + 1:, i]
            y = np.conj(A[j + 1:, j + 1:]).T @ vv          # = A_trail v because A_trail is Hermitian (column dots)
            if i > 0:
                p1 = np.conj(W[j + 1:, :i]).T @ vv
                p2 = np.conj(V[j + 1:, :i]).T @ vv
                y = y - V[j + 1:, :i] @ p1 - W[j + 1:, :i] @ p2
            wv = tau * y
            # (4) w -= 1/2 tau (w^H v) v
            alpha = -0.5 * tau * np.vdot(wv, vv)
            wv = wv + alpha * vv
            W[j + 1:, i] = wv
        # trailing update A -= V W^H + W V^H on the full square
        q = p + w
        A[q:, q:] -= V[q:, :] @ np.conj(W[q:, :]).T + W[q:, :] @ np.conj(V[q:, :]).T
        Vall[:, p:p + w] = V
        p = q
    d[n - 1] = np.real(A[n - 1, n - 1])
    return d, e, Vall, taus


def larft(V, tau):
    """Forward columnwise T factor: H_0 ... H_{k-1} = I - V T V^H."""
    k = V.shape[1]
    S = np.conj(V).T @ V
    T = np.zeros((k, k), dtype=V.dtype)
    for i in range(k):
        T[i, i] = tau[i]
        if i > 0:
            T[:i, i] = -tau[i] * (T[:i, :i] @ S[:i, i])
    return T


def backtransform(Vall, taus, Z, nb=8):
    """U = Q Z, Q = H_0 ... H_{n-2} applied panel by panel (last panel first)."""
    n = Vall.shape[0]
    U = Z.astype(Vall.dtype, copy=True)
    nref = n - 1
    starts = list(range(0, nref, nb))
    for p in reversed(starts):
        w = min(nb, nref - p)
        V = Vall[:, p:p + w]
        T = larft(V, taus[p:p + w])
        Y = np.conj(V[p + 1:, :]).T @ U[p + 1:, :]
        U[p + 1:, :] -= V[p + 1:, :] @ (T @ Y)
    return U


# ----------------------------------------------------------------------------------------------------------
# stage 2: divide & conquer
# ----------------------------------------------------------------------------------------------------------
def plan_merge(D, z, rho):
    """Deflation (LAPACK dlaed2 logic).  D: eigenvalues of the two sub-blocks (any order), z: coupling vector
    (norm 1), rho > 0.  Returns rotations [(pj, nj, c, s)], non-deflated indices (ascending D), deflated indices,
    and the updated D, z."""
    D = D.copy()
    z = z.copy()
    N = len(D)
    tol = 8.0 * EPS * max(np.max(np.abs(D)), np.max(np.abs(z)))
    order = np.argsort(D, kind="stable")
    rots, nd, df = [], [], []
    if rho * np.max(np.abs(z)) <= tol:
        return rots, nd, list(order), D, z
    pj = -1
    for j in order:
        if rho * abs(z[j]) <= tol:
            df.append(j)
            continue
        if pj < 0:
            pj = j
            continue
        s, c = z[pj], z[j]
        tau = np.hypot(c, s)
        t = D[j] - D[pj]
        c /= tau
        s = -s / tau
        if abs(t * c * s) <= tol:
            z[j] = tau
            z[pj] = 0.0
            rots.append((pj, j, c, s))
            t = D[pj] * c * c + D[j] * s * s
            D[j] = D[pj] * s * s + D[j] * c * c
            D[pj] = t
            df.append(pj)
            pj = j
        else:
            nd.append(pj)
            pj = j
    nd.append(pj)
    return rots, nd, df, D, z


def secular_root(i, d, z2, rho):
    """Root i of 1 + rho sum_j z2_j / (d_j - lam) in (d_i, d_{i+1}) (last: (d_K-1, d_K-1 + rho sum z2)).
    Returns (origin index, tau, delta) with delta_j = (d_j - d_origin) - tau = d_j - lam to high relative accuracy."""
    K = len(d)
    if K == 1:
        tau = rho * z2[0]
        return 0, tau, np.array([-tau])
    last = i == K - 1
    if last:
        org = K - 1
        delta0 = d - d[org]
        lo, hi = 0.0, rho * np.sum(z2)
        # f at the midpoint decides nothing here; start in the middle
        tau = 0.5 * hi
    else:
        gap = d[i + 1] - d[i]
        delta_i = d - d[i]
        fmid = 1.0 + rho * np.sum(z2 / (delta_i - 0.5 * gap))
        if fmid > 0.0:
            org = i
            delta0 = delta_i
            lo, hi = 0.0, 0.5 * gap
        else:
            org = i + 1
            delta0 = d - d[i + 1]
            lo, hi = -0.5 * gap, 0.0
        tau = 0.5 * (lo + hi)
        if fmid == 0.0:
            return org, 0.5 * gap, delta0 - 0.5 * gap
    # iterate on g(tau) = 1 + rho sum z2 / (delta0 - tau), increasing in tau
    ip = i if not last else K - 2   # psi: poles 0..ip, phi: poles ip+1..K-1
    for it in range(100):
        dl = delta0 - tau
        t = z2 / dl
        psi = rho * np.sum(t[:ip + 1])
        phi = rho * np.sum(t[ip + 1:])
        dpsi = rho * np.sum(t[:ip + 1] / dl[:ip + 1])
        dphi = rho * np.sum(t[ip + 1:] / dl[ip + 1:])
        g = 1.0 + psi + phi
        err = 8.0 * EPS * (1.0 + abs(psi) + abs(phi)) * 1.0 + EPS * K * 0  # noqa
        erretm = 1.0 + abs(psi) + abs(phi)
        if abs(g) <= 4.0 * EPS * erretm:
            break
        if g > 0.0:
            hi = tau
        else:
            lo = tau
        # two-pole rational model: psi ~ s + S/(dA - tau), phi ~ r + R/(dB - tau)
        dA, dB = dl[ip], dl[ip + 1]            # distances from the current tau to the two bracketing poles
        S = dpsi * dA * dA
        s_ = psi - dpsi * dA
        R = dphi * dB * dB
        r_ = phi - dphi * dB
        c0 = 1.0 + s_ + r_
        # solve c0 + S/(dA - x) + R/(dB - x) = 0 for the step x (x measured from the current tau)
        # c0 (dA - x)(dB - x) + S (dB - x) + R (dA - x) = 0
        a = c0
        b = -(c0 * (dA + dB) + S + R)
        cc = c0 * dA * dB + S * dB + R * dA
        x = None
        if a == 0.0:
            if b != 0.0:
                x = -cc / b
        else:
            disc = b * b - 4.0 * a * cc
            if disc >= 0.0:
                sq = np.sqrt(disc)
                # the root between the poles: choose the numerically stable form
                q = -0.5 * (b + np.copysign(sq, b))
                cands = []
                if q != 0.0:
                    cands.append(cc / q)
                if a != 0.0:
                    cands.append(q / a)
                for cnd in cands:
                    tn = tau + cnd
                    if lo < tn < hi:
                        x = cnd
                        break
        if x is None or not (lo < tau + x < hi):
            tn = 0.5 * (lo + hi)
        else:
            tn = tau + x
        if tn == tau or hi - lo <= 2.0 * EPS * max(abs(lo), abs(hi)):
            tau = tn
            break
        tau = tn
    return org, tau, delta0 - tau


def merge(D, Q, beta, n1):
    """Eigen-decomposition of diag(D) + |beta| w w^T in the basis Q (block diagonal Q1 (+) Q2)."""
    N = len(D)
    z = np.concatenate([Q[n1 - 1, :n1], np.sign(beta) * Q[n1, n1:]]) / np.sqrt(2.0)
    rho = 2.0 * abs(beta)
    rots, nd, df, D2, z2v = plan_merge(D, z, rho)
    Q = Q.copy()
    for (pj, nj, c, s) in rots:
        x, y = Q[:, pj].copy(), Q[:, nj].copy()
        Q[:, pj] = c * x + s * y
        Q[:, nj] = c * y - s * x
    K = len(nd)
    Dn = np.zeros(N)
    Qn = np.zeros_like(Q)
    if K > 0:
        dl = D2[nd]
        zz = z2v[nd]
        z2 = zz * zz
        Dt = np.zeros((K, K))     # Dt[i, j] = d_j - lam_i
        lam = np.zeros(K)
        for i in range(K):
            org, tau, delta = secular_root(i, dl, z2, rho)
            Dt[i, :] = delta
            lam[i] = dl[org] + tau
        # Gu-Eisenstat: zhat_j^2 = prod_i (lam_i - d_j) / prod_{i != j} (d_i - d_j)
        zh = np.zeros(K)
        for j in range(K):
            pr = Dt[j, j]
            for i in range(K):
                if i != j:
                    pr *= Dt[i, j] / (dl[j] - dl[i])
            zh[j] = np.copysign(np.sqrt(-pr), zz[j])
        Ut = zh[None, :] / Dt      # Ut[i, j] = zhat_j / (d_j - lam_i)
        Ut /= np.linalg.norm(Ut, axis=1)[:, None]
        Qn[:, :K] = Q[:, nd] @ Ut.T
        Dn[:K] = lam
    Qn[:, K:] = Q[:, df]
    Dn[K:] = D2[df]
    return Dn, Qn, K


def leaf_bounds(n, leaf):
    nl = 1
    while n > nl * leaf:
        nl *= 2
    b = [(n * i) // nl for i in range(nl + 1)]
    return b


def dc_tridiag(d, e, leaf=16):
    n = len(d)
    d = d.copy()
    b = leaf_bounds(n, leaf)
    for x in b[1:-1]:
        d[x - 1] -= abs(e[x - 1])
        d[x] -= abs(e[x - 1])
    D = np.zeros(n)
    Z = np.zeros((n, n))
    for k in range(len(b) - 1):
        lo, hi = b[k], b[k + 1]
        T = np.diag(d[lo:hi]) + np.diag(e[lo:hi - 1], 1) + np.diag(e[lo:hi - 1], -1)
        w, v = np.linalg.eigh(T)
        D[lo:hi] = w
        Z[lo:hi, lo:hi] = v
    stats = []
    while len(b) > 2:
        nb_ = [b[0]]
        for k in range(0, len(b) - 1, 2):
            lo, mid, hi = b[k], b[k + 1], b[k + 2]
            Dn, Qn, K = merge(D[lo:hi], Z[lo:hi, lo:hi], e[mid - 1], mid - lo)
            D[lo:hi] = Dn
            Z[lo:hi, lo:hi] = Qn
            stats.append((hi - lo, K))
            nb_.append(hi)
        b = nb_
    return D, Z, stats


def eigh_proto(A, nb=8, leaf=16):
    d, e, Vall, taus = sytrd_blocked(A, nb)
    D, Z, stats = dc_tridiag(d, e, leaf)
    U = backtransform(Vall, taus, Z, nb)
    return D, U, stats


def check(name, A, **kw):
    D, U, stats = eigh_proto(A, **kw)
    n = A.shape[0]
    nrm = np.linalg.norm(A, 2)
    res = np.linalg.norm(A @ U - U * D[None, :]) / (nrm * n)
    orth = np.linalg.norm(np.conj(U).T @ U - np.eye(n)) / n
    wref = np.linalg.eigvalsh(A)
    ev = np.max(np.abs(np.sort(D) - wref)) / nrm
    print(f"{name:28s} n={n:4d} resid/(n|A|)={res:.2e} orth/n={orth:.2e} eig err/|A|={ev:.2e} "
          f"top merge K={stats[-1][1] if stats else 0}/{stats[-1][0] if stats else 0}")
    assert res < 50 * EPS and orth < 50 * EPS and ev < 200 * EPS, name


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for n in (5, 33, 64, 150, 257):
        M = rng.standard_normal((n, n))
        check("gauss sym", M + M.T, nb=8, leaf=16)
        G = rng.standard_normal((n, 2 * n))
        check("gram (wishart)", G @ G.T, nb=8, leaf=16)
        Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
        s = np.exp(-40.0 * np.arange(n) / n)
        check("graded gram 1e-17", (Q * s) @ Q.T, nb=8, leaf=16)
        check("low rank", (Q[:, :3] * np.array([1.0, 0.5, 1e-3])) @ Q[:, :3].T, nb=8, leaf=16)
        check("identity + rank1", np.eye(n) + 1e-3 * np.outer(Q[:, 0], Q[:, 0]), nb=8, leaf=16)
        cl = np.repeat(np.arange(1, n // 8 + 2), 8)[:n].astype(float)
        check("clustered", (Q * cl) @ Q.T, nb=8, leaf=16)
        Mc = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
        check("complex hermitian", Mc + np.conj(Mc).T, nb=8, leaf=16)
        Gc = rng.standard_normal((n, n + 3)) + 1j * rng.standard_normal((n, n + 3))
        check("complex gram", Gc @ np.conj(Gc).T, nb=8, leaf=16)
    # Wilkinson / glued
    n = 101
    T = np.diag(np.abs(np.arange(n) - n // 2).astype(float)) + np.diag(np.ones(n - 1), 1) + np.diag(np.ones(n - 1), -1)
    check("wilkinson", T, nb=8, leaf=16)
    T = np.diag(np.zeros(n)) + np.diag(np.ones(n - 1), 1) + np.diag(np.ones(n - 1), -1)
    check("zero-diagonal chain", T, nb=8, leaf=16)
    print("ok")
