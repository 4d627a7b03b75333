"""Independent numpy restatement of the hot path, written from SURVEY.md Appendix A (clean
linear algebra, not the operation order of the C++ oracle).  Used to cross-check the oracle,
and -- `kernel_slot_math` -- to validate on the CPU the algebra of the CUDA kernel's
measurement-aligned basis before any GPU run.
"""
import numpy as np

SEL = [9, 10, 0, 1, 12, 13]  # H1 ones, src/my_gmm.cpp:62-67


class NpModel:
    def __init__(self, means, covs, weights, gamma, pca_proj, pca_mean, r=100.0):
        self.means = np.asarray(means, float)
        self.K, self.d = self.means.shape
        self.covs = np.asarray(covs, float).reshape(self.K, self.d, self.d)
        self.weights = np.asarray(weights, float).reshape(-1)
        self.gamma = np.asarray(gamma, float).reshape(-1)[: self.K]
        self.proj = np.asarray(pca_proj, float)
        self.pmean = np.asarray(pca_mean, float).reshape(-1)
        self.D = self.proj.shape[1]
        self.H = self.proj[:, SEL].T.copy()  # H1 * pca_proj^T
        self.BH = self.pmean[SEL].copy()
        self.R = r * np.eye(6)
        self.Q = (1 - self.gamma**2)[:, None, None] * self.covs
        self.B = (1 - self.gamma)[:, None] * self.means


def predict(m, k, x, P):
    g = m.gamma[k]
    return g * x + m.B[k], g * g * P + m.Q[k]


def pseudo_chol(S, mode="cv24"):
    """ParticleFilter::chol (src/pf2DRao.cpp:34-53) -> upper-triangular pseudo factor."""
    n = S.shape[0]
    Lc = np.linalg.cholesky(0.5 * (S + S.T))
    if mode == "exact":
        return Lc.T.copy()
    Rt = np.zeros_like(S)
    for e in range(n):
        lee = Lc[e, e]
        elem = 1.0 / lee if mode == "cv24" else lee
        Rt[e, e] = 1.0 / elem
        Rt[e, e + 1:] = S[e, e + 1:] * elem
    return Rt


def mvnpdf(x, u, S, mode="cv24"):
    Rt = pseudo_chol(S, mode)
    v = np.linalg.solve(Rt.T, x - u)
    return float(np.exp(-0.5 * v @ v - np.log(np.diag(Rt)).sum() - len(x) * np.log(2 * np.pi) / 2))


def kf_update(m, z, x, P):
    y = z - (m.H @ x + m.BH)
    S = m.H @ P @ m.H.T + m.R
    K = P @ m.H.T @ np.linalg.inv(S)
    return x + K @ y, (np.eye(m.d) - K @ m.H) @ P


def resample_closed_form(w, N, u):
    """out[i] = min{k : (u+i)/N <= cumsum_k}, no wrap handling."""
    c = np.cumsum(w)
    t = (u + np.arange(N)) / N
    return np.minimum(np.searchsorted(c, t, side="left"), len(w) - 1).astype(np.int32)


def resample_loop(w, N, u):
    """literal loop of src/pf2DRao.cpp:195-207 in Python floats (slow; small cases)."""
    L = len(w)
    idx = 0
    step = 1.0 / N
    beta = u * step
    out = np.zeros(N, np.int32)
    for i in range(N):
        while beta > w[idx]:
            beta -= w[idx]
            idx = (idx + 1) % L
        beta += step
        out[i] = idx
    return out


def filter_update(m, x, P, meas, ind, mode="cv24"):
    """one frame for one track given the indicators: returns children (x', P'), raw weights."""
    N = x.shape[0]
    xo = np.zeros_like(x)
    Po = np.zeros_like(P)
    w = np.zeros(N)
    for j in range(N):
        z = meas[:, j] if meas.ndim == 2 else meas
        xp, Pp = predict(m, ind[j], x[j], P[j])
        w[j] = mvnpdf(z, m.H @ xp + m.BH, m.H @ Pp @ m.H.T + m.R, mode)
        xo[j], Po[j] = kf_update(m, z, xp, Pp)
    return xo, Po, w


# ---------------------------------------------------------------------------------------------
# CPU model of the CUDA kernel's arithmetic (csrc/mkf_kernels.cuh: slot_math) -- the
# measurement-aligned basis x' = T x with T = [H; N]
# ---------------------------------------------------------------------------------------------
def aligned_basis(H):
    _, _, Vt = np.linalg.svd(H)
    T = np.vstack([H, Vt[H.shape[0]:]])
    return T, np.linalg.inv(T)


def kernel_slot_math(m, T, Ti, k, x, P, z, mode="cv24"):
    r = m.R[0, 0]
    g = m.gamma[k]
    xq = T @ x
    Pq = T @ P @ T.T
    bq = T @ m.B[k]
    Qq = T @ m.Q[k] @ T.T
    xq = g * xq + bq
    Pq = g * g * Pq + Qq
    A, Bm, Cm = Pq[:6, :6], Pq[6:, :6], Pq[6:, 6:]
    y = (z - m.BH) - xq[:6]
    S = A + r * np.eye(6)
    Lc = np.linalg.cholesky(S)
    inv = 1.0 / np.diag(Lc)
    # likelihood
    if mode == "exact":
        v = np.linalg.solve(Lc, y)
        pinv = np.prod(inv)
    else:
        elem = inv if mode == "cv24" else np.diag(Lc)
        v = np.zeros(6)
        ve = np.zeros(6)
        for j in range(6):
            acc = y[j] - sum(ve[e] * A[j, e] for e in range(j))
            v[j] = acc * elem[j]
            ve[j] = v[j] * elem[j]
        pinv = np.prod(elem)
    w = np.exp(-0.5 * v @ v + np.log(pinv) - 3 * np.log(2 * np.pi))
    W = np.linalg.inv(S)
    t = W @ y
    x1 = xq[:6] + y - r * t
    x2 = xq[6:] + Bm @ t
    G = W @ Bm.T
    Cn = Cm - Bm @ G
    Bn = r * G.T
    An = r * np.eye(6) - r * r * W
    Pn = np.block([[An, Bn.T], [Bn, Cn]])
    xn = np.concatenate([x1, x2])
    return Ti @ xn, Ti @ Pn @ Ti.T, float(w)
