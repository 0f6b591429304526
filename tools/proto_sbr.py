"""NumPy statement of a two-stage Hermitian tridiagonalisation (successive band reduction), the round-2 candidate that turns the
n^3 part of the eigensolver (csrc/eigh.cu: 311 ms of HBM-bound column products at n = 8192) into GEMMs:

  stage 1  full -> band (lower bandwidth b): QR of the panel below the band of every block column, two-sided compact-WY update
           A <- Q^H A Q of the trailing matrix (all GEMM: 4/3 n^3 flop on the DMMA pipe, no per-column grid barrier);
  stage 2  band -> tridiagonal by bulge chasing: sweep j annihilates column j below the sub-diagonal with a reflector of length
           <= b and chases the bulge down the band with (n - j) / b further reflectors; O(n^2 b) flop on O(n b) data (the band
           stays in L2 / shared memory).  Task (j, s) = step s of sweep j touches rows / columns [j + 1 + s b, j + 1 + (s + 2) b);
           sweep j + 1 may run step s once sweep j has finished step s + 2 (checked below by replaying the tasks in that
           wavefront order), i.e. ~ n / (3 b) tasks run concurrently;
  stage 3  eigen-decomposition of the tridiagonal matrix (the existing divide & conquer);
  stage 4  back-transformation U = Q1 (Q2 Z): Q2 = product of the chasing reflectors, applied in reverse order -- on the device as
           compact-WY blocks of g consecutive sweeps at one step s (apply_q2: groups descending, steps ascending inside a group);
           Q1 = product of the stage-1 block reflectors (the machinery of Eigh::vectors).

    python tools/proto_sbr.py
Used by tests/test_cpu_dc.py::test_two_stage_tridiagonalisation_prototype."""
import numpy as np


def house(x):
    """Returns (v, g, beta) with v[0] = 1 and G = I - g v v^H such that G x = beta e_0, beta real (g = conj(tau) of zlarfg)."""
    x = np.asarray(x)
    alpha = x[0]
    sigma = np.vdot(x[1:], x[1:]).real
    v = x.astype(complex if np.iscomplexobj(x) else float).copy()
    if sigma == 0.0 and np.imag(alpha) == 0.0:
        v[:] = 0.0; v[0] = 1.0
        return v, 0.0, alpha
    beta = -np.copysign(np.sqrt(abs(alpha) ** 2 + sigma), np.real(alpha))
    tau = (beta - alpha) / beta
    v = v / (alpha - beta)
    v[0] = 1.0
    return v, np.conj(tau), beta           # G = I - conj(tau) v v^H (= H^H of zlarfg) maps x to beta e_0


def panel_qr_wy(P):
    """Householder QR of an m x w panel in compact-WY form: P = Q [R; 0], Q = I - V T V^H (V unit lower trapezoidal m x w,
    T upper triangular w x w) -- the device panel factorisation (house_* kernels) followed by larft."""
    P = P.astype(complex if np.iscomplexobj(P) else float).copy()
    m, w = P.shape
    k = min(m, w)
    V = np.zeros((m, k), dtype=P.dtype)
    taus = np.zeros(k, dtype=P.dtype)
    for c in range(k):
        v, g, beta = house(P[c:, c])                     # G = I - g v v^H, G x = beta e_0
        P[c:, c + 1:] -= g * np.outer(v, v.conj() @ P[c:, c + 1:])
        P[c, c] = beta
        P[c + 1:, c] = 0.0
        V[c:, c] = v
        taus[c] = np.conj(g)                             # G = H^H with H = I - tau v v^H  =>  Q = H_0 H_1 .. = I - V T V^H
    T = np.zeros((k, k), dtype=P.dtype)
    for c in range(k):
        T[c, c] = taus[c]
        if c:
            T[:c, c] = -taus[c] * (T[:c, :c] @ (V[:, :c].conj().T @ V[:, c]))
    return V, T, np.triu(P[:k, :])


def panel_qr_wy_gemm(P):
    """The same compact-WY factorisation without a column-by-column Householder loop (tensor-core friendly panel, round-2
    candidate): CholeskyQR2 gives an explicit orthonormal Q (two Gram GEMMs + two w x w Cholesky + two triangular solves),
    Householder reconstruction (Ballard et al., "Reconstructing Householder vectors from TSQR", 2014) turns it into V and T:
      S = -sign(diag) chosen during the unpivoted LU of Q - [S; 0] = L U,   V = L (unit lower trapezoidal),
      T = -U S^H V_1^{-H}  (V_1 = top w x w block of V),   P = (I - V T V^H) [S R; 0].
    Unpivoted LU is stable here because Q - [S; 0] is diagonally dominant by construction (|q_ii - s_i| >= 1).
    CholeskyQR2 needs cond(P) < ~1e8; panels of a strongly graded density matrix need the shifted three-pass variant or a
    fall-back to the Householder loop (decided per panel from the Cholesky pivots)."""
    P = np.asarray(P)
    m, w = P.shape
    if m < w:
        return panel_qr_wy(P)
    dt = complex if np.iscomplexobj(P) else float
    Q = P.astype(dt)
    Rtot = np.eye(w, dtype=dt)
    for _ in range(2):                                     # CholeskyQR2: second pass restores orthogonality to eps
        G = Q.conj().T @ Q
        Rc = np.linalg.cholesky(G).conj().T                # G = Rc^H Rc
        Q = np.linalg.solve(Rc.conj().T, Q.conj().T).conj().T   # Q Rc^{-1}
        Rtot = Rc @ Rtot
    # unpivoted LU of Q - [S; 0] with S chosen on the fly
    M = Q.copy()
    S = np.zeros(w, dtype=dt)
    for c in range(w):
        d = M[c, c]
        S[c] = -(d / abs(d)) if abs(d) > 0 else -1.0
        M[c, c] -= S[c]
        M[c + 1:, c] /= M[c, c]
        M[c + 1:, c + 1:] -= np.outer(M[c + 1:, c], M[c, c + 1:])
    V = np.tril(M, -1)[:, :w] + np.eye(m, w, dtype=dt)
    U = np.triu(M[:w, :w])
    T = -U @ np.diag(S.conj()) @ np.linalg.inv(V[:w, :w].conj().T)
    R = np.diag(S) @ Rtot
    return V, T, R


def full_to_band(A, b, panel=None):
    """Stage 1.  Returns the band matrix (dense storage, entries below the b-th sub-diagonal are zero) and the block
    reflectors [(row offset, V, T)] with Q1 = prod_k (I - V_k T_k V_k^H).  The trailing update is the GEMM sequence planned
    for the device:  X = S V T,  W = X - 1/2 V T^H (V^H X),  S <- S - W V^H - V W^H  (= Q^H S Q)."""
    A = A.copy()
    n = A.shape[0]
    blocks = []
    for k in range(0, n - b - 1, b):
        r0 = k + b
        w = min(b, n - k)
        V, T, R = (panel or panel_qr_wy)(A[r0:, k:k + w])
        A[r0:, k:k + w] = 0.0
        A[r0:r0 + R.shape[0], k:k + w] = R
        A[k:k + w, r0:] = A[r0:, k:k + w].conj().T
        S = A[r0:, r0:]
        X = S @ V @ T
        W = X - 0.5 * V @ (T.conj().T @ (V.conj().T @ X))
        A[r0:, r0:] = S - W @ V.conj().T - V @ W.conj().T
        blocks.append((r0, V, T))
    return A, blocks


def band_to_tridiagonal(B, b, order="sweeps"):
    """Stage 2 on dense storage.  order = "sweeps": sweep after sweep; "wavefront": task (j, s) at time 3 j + s (all tasks of
    one time step are independent).  Returns T (tridiagonal, dense) and the reflector list [(row offset, v, tau)] in the order
    they were applied."""
    A = B.copy()
    n = A.shape[0]
    refl = []

    def task(j, s):
        # step 0 annihilates column j below the sub-diagonal; step s > 0 annihilates the first column of the bulge that
        # step s - 1 created: column c = j + 1 + (s - 1) b, rows r0 .. r0 + b - 1 with r0 = j + 1 + s b
        r0 = j + 1 + s * b
        c = j if s == 0 else j + 1 + (s - 1) * b
        r1 = min(r0 + b, n)
        if r1 - r0 < 2:
            return False
        x = A[r0:r1, c]
        if not np.any(x[1:]):
            return r1 < n
        v, tau, beta = house(x)
        G = np.eye(r1 - r0, dtype=A.dtype) - tau * np.outer(v, v.conj())
        # G x = beta e_0: similarity A <- G A G^H on rows / columns r0 .. r1 - 1
        A[r0:r1, :] = G @ A[r0:r1, :]
        A[:, r0:r1] = A[:, r0:r1] @ G.conj().T
        refl.append((r0, v, tau, j, s))
        return r1 < n

    nsteps = lambda j: max(0, -(-(n - j - 1) // b))
    if order == "sweeps":
        for j in range(n - 2):
            for s in range(nsteps(j)):
                task(j, s)
    else:
        tmax = 3 * (n - 2) + nsteps(0)
        for t in range(tmax + 1):
            for j in range(min(n - 2, t // 3 + 1)):
                s = t - 3 * j
                if 0 <= s < nsteps(j):
                    task(j, s)
    return A, refl


def apply_q2(U, refl, n, b, group=0):
    """U <- Q2 U.  A = G^H A' G for every task, so the eigenvectors pick up G^H = I - conj(g) v v^H, last task first.
    group = 0: reflector by reflector in reverse order of application (j descending, s descending).
    group = g: the device order -- sweeps in groups J of g consecutive j (descending); within a group the steps s ASCENDING
    (a task (j', s') with j' > j, s' > s acts on rows strictly below task (j, s), so they commute; descending s would not);
    the g reflectors of one (J, s) are staggered by one row each, i.e. they form the unit lower trapezoidal V of a QR panel
    and are applied as one compact-WY block  U[rows] -= V (T^H (V^H U[rows]))  -- three GEMMs."""
    U = U.copy()
    if not group:
        for r0, v, tau, j, s_ in reversed(refl):
            U[r0:r0 + len(v), :] -= np.conj(tau) * np.outer(v, v.conj() @ U[r0:r0 + len(v), :])
        return U
    by = {}
    for r0, v, tau, j, s_ in refl:
        by[(j, s_)] = (r0, v, tau)
    jmax = max(j for j, _ in by) if by else -1
    smax = max(s_ for _, s_ in by) if by else -1
    for j0 in range((jmax // group) * group, -1, -group):
        for s_ in range(0, smax + 1):
            js = [j for j in range(j0, min(j0 + group, jmax + 1)) if (j, s_) in by]
            if not js:
                continue
            rlo = min(by[(j, s_)][0] for j in js)
            rhi = max(by[(j, s_)][0] + len(by[(j, s_)][1]) for j in js)
            V = np.zeros((rhi - rlo, len(js)), dtype=U.dtype)
            taus = np.zeros(len(js), dtype=U.dtype)
            for c, j in enumerate(js):
                r0, v, tau = by[(j, s_)]
                V[r0 - rlo:r0 - rlo + len(v), c] = v
                taus[c] = np.conj(tau)                     # G^H = I - conj(g) v v^H
            # product G_{j_first}^H ... G_{j_last}^H applied to U means the LAST sweep acts first: H_0 H_1 .. H_{k-1} = I - V T V^H
            T = np.zeros((len(js), len(js)), dtype=U.dtype)
            for c in range(len(js)):                       # forward larft: T[:c, c] = -tau_c T[:c, :c] V[:, :c]^H V[:, c]
                T[c, c] = taus[c]
                if c:
                    T[:c, c] = -taus[c] * (T[:c, :c] @ (V[:, :c].conj().T @ V[:, c]))
            U[rlo:rhi, :] -= V @ (T @ (V.conj().T @ U[rlo:rhi, :]))
    return U


def two_stage_eigh(A, b, order="sweeps", group=0, panel=None):
    n = A.shape[0]
    Bm, blocks = full_to_band(A, b, panel)
    band_err = max((np.abs(np.tril(Bm, -b - 1)).max() if n > b + 1 else 0.0), 0.0)
    T, refl = band_to_tridiagonal(Bm, b, order)
    tri_err = np.abs(np.tril(T, -2)).max() if n > 2 else 0.0
    d = np.real(np.diag(T)).copy()
    e = np.diag(T, -1).copy()
    # phases of a complex sub-diagonal are absorbed into a diagonal similarity (stage 3 works on a real tridiagonal matrix)
    ph = np.ones(n, dtype=T.dtype)
    for i in range(n - 1):
        ph[i + 1] = ph[i] * (e[i] / abs(e[i]) if abs(e[i]) > 0 else 1.0)
    Tr = np.diag(d) + np.diag(np.abs(e), -1) + np.diag(np.abs(e), 1)
    w, Z = np.linalg.eigh(Tr)
    U = (ph[:, None] * Z).astype(T.dtype)
    U = apply_q2(U, refl, n, b, group=group)               # Q2 Z
    for r0, V, T in reversed(blocks):                     # Q1 (Q2 Z): block reflectors, last panel first
        U[r0:, :] -= V @ (T @ (V.conj().T @ U[r0:, :]))
    return w, U, band_err, tri_err


def check(verbose=True):
    rng = np.random.default_rng(0)
    worst = 0.0
    for n, b, cplx, order, group in [(40, 4, False, "sweeps", 0), (61, 8, False, "wavefront", 4), (96, 16, True, "sweeps", 16), (75, 6, True, "wavefront", 5),
                                     (130, 32, False, "wavefront", 32), (33, 40, False, "sweeps", 8), (90, 8, False, "sweeps", 3)]:
        M = rng.standard_normal((n, n)) + (1j * rng.standard_normal((n, n)) if cplx else 0.0)
        A = M + M.conj().T
        w, U, be, te = two_stage_eigh(A, b, order, group, panel_qr_wy_gemm if (n + b) % 2 else None)
        nrm = np.linalg.norm(A, 2)
        res = np.linalg.norm(A @ U - U * w[None, :]) / (nrm * n)
        orth = np.linalg.norm(U.conj().T @ U - np.eye(n)) / n
        err = np.abs(w - np.linalg.eigvalsh(A)).max() / nrm
        worst = max(worst, res, orth, err, be / nrm, te / nrm)
        if verbose:
            print(f"n={n:4d} b={b:3d} {'c128' if cplx else 'f64 '} {order:9s} group {group:2d} residual {res:.1e} orthogonality {orth:.1e} eigenvalues {err:.1e} "
                  f"below band {be / nrm:.1e} below sub-diagonal {te / nrm:.1e}")
    return worst


if __name__ == "__main__":
    print("worst", check())
