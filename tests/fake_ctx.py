"""Test double for the device context: the same call surface as g3py_b200._cabi.Context, computed with
NumPy from the kernel descriptor.  Lives under tests/ only; lets the CPU suite exercise the host-side
assembly (bijection, mean / warping / degree terms, chain rules, sharding) without a GPU."""
import numpy as np
import scipy.linalg as sla

from g3py_b200 import _cabi as cabi


def _eval_desc(desc, th, X1, X2, same, skip_pn=False, grad=False):
    n1, n2 = X1.shape[0], X2.shape[0]
    vals, grads = [], []                                  # grads[node] = {theta index: dK/dtheta}
    for n in range(desc.n_nodes):
        nd = desc.nodes[n]
        op = nd.op
        g = {}
        if op < cabi.K_SUM:
            var = th[nd.var_idx] if nd.var_idx >= 0 else nd.value
            diff = X1[:, None, nd.dim0:nd.dim1] - X2[None, :, nd.dim0:nd.dim1]
            w = nd.dim1 - nd.dim0
            eye = np.eye(n1) if same else np.zeros((n1, n2))
            if op in (cabi.K_SE, cabi.K_MAT32, cabi.K_MAT52, cabi.K_RQ):
                r = th[nd.p0_idx:nd.p0_idx + w]
                d = (diff ** 2 * (0.5 * r ** 2)).sum(-1)
                if op == cabi.K_SE:
                    k, dk = np.exp(-d), -np.exp(-d)
                elif op == cabi.K_MAT32:
                    s = np.sqrt(3 * d); k, dk = (1 + s) * np.exp(-s), -1.5 * np.exp(-s)
                elif op == cabi.K_MAT52:
                    s = np.sqrt(5 * d); k, dk = (1 + s + 5 * d / 3) * np.exp(-s), -(5 / 6) * (1 + s) * np.exp(-s)
                else:
                    al = th[nd.p1_idx]; k = (1 + d / al) ** (-al); dk = -(1 + d / al) ** (-al - 1)
                    g[nd.p1_idx] = var * k * (-np.log1p(d / al) + d / (al + d))
                for j in range(w):
                    g[nd.p0_idx + j] = var * dk * r[j] * diff[:, :, j] ** 2
            elif op == cabi.K_OU:
                r = th[nd.p0_idx:nd.p0_idx + w]
                k = np.exp(-(np.abs(diff) * r).sum(-1))
                for j in range(w):
                    g[nd.p0_idx + j] = -var * k * np.abs(diff[:, :, j])
            elif op == cabi.K_SIN:
                r = th[nd.p0_idx:nd.p0_idx + w]; f = th[nd.p1_idx:nd.p1_idx + w]
                k = np.exp(2 * (np.sin(np.pi * diff * f) ** 2 * r).sum(-1))
                for j in range(w):
                    g[nd.p0_idx + j] = var * k * 2 * np.sin(np.pi * diff[:, :, j] * f[j]) ** 2
                    g[nd.p1_idx + j] = var * k * 2 * r[j] * np.sin(2 * np.pi * diff[:, :, j] * f[j]) * np.pi * diff[:, :, j]
            elif op in (cabi.K_COS, cabi.K_SINC, cabi.K_SM):
                f = th[nd.p1_idx:nd.p1_idx + w]
                pi2 = np.pi ** 2
                if op == cabi.K_SINC:
                    bb = 2 * pi2 * diff * f
                    with np.errstate(invalid="ignore", divide="ignore"):
                        fac = np.where(diff != 0, np.sin(bb) / bb, 1.0)
                        dfac = np.where(diff != 0, (np.cos(bb) - fac) / f, 0.0)
                else:
                    fac = np.cos(2 * np.pi * diff * f)
                    dfac = -np.sin(2 * np.pi * diff * f) * 2 * np.pi * diff
                env = 1.0
                if op == cabi.K_SM:
                    r = th[nd.p0_idx:nd.p0_idx + w]
                    env = np.exp(-2 * pi2 * (diff ** 2 * r ** 2).sum(-1))
                k = env * fac.prod(-1)
                for j in range(w):
                    g[nd.p1_idx + j] = var * env * dfac[:, :, j] * np.delete(fac, j, axis=2).prod(-1)
                    if op == cabi.K_SM:
                        g[nd.p0_idx + j] = var * k * (-4 * pi2 * diff[:, :, j] ** 2 * r[j])
            elif op == cabi.K_DOT:
                r = th[nd.p0_idx:nd.p0_idx + w]
                a, b2 = X1[:, None, nd.dim0:nd.dim1], X2[None, :, nd.dim0:nd.dim1]
                m = (a * b2 * r ** 2).sum(-1) + (th[nd.p1_idx] if nd.p1_idx >= 0 else 0.0)
                pw = ((nd.flags >> 8) & 0xff) or 1
                if nd.flags & cabi.KF_NN:                      # var * arcsin(2m / (1 + 2m)^2)
                    u = 2 * m / (1 + 2 * m) ** 2
                    k = np.arcsin(u)
                    dm = var * (2 * (1 - 2 * m) / (1 + 2 * m) ** 3) / np.sqrt(1 - u * u)
                else:
                    k = m ** pw
                    dm = var * pw * m ** (pw - 1)
                for j in range(w):
                    g[nd.p0_idx + j] = dm * 2 * r[j] * a[:, :, j] * b2[:, :, j]
                if nd.p1_idx >= 0:
                    g[nd.p1_idx] = dm
            elif op == cabi.K_BW:
                k = np.minimum(X1[:, None, nd.dim0:nd.dim1], X2[None, :, nd.dim0:nd.dim1]).prod(-1)
            elif op == cabi.K_VAR:
                k = np.ones((n1, n2))
            elif op == cabi.K_EQ:
                import struct
                e1 = nd.value
                e2 = struct.unpack("<d", struct.pack("<ii", nd.p0_idx, nd.p1_idx))[0] if nd.flags & cabi.KF_EQ2 else e1
                a, b2 = X1[:, None, nd.dim0:nd.dim1], X2[None, :, nd.dim0:nd.dim1]
                k = ((a == e1) * (b2 == e2)).sum(-1).astype(float)
                if nd.flags & cabi.KF_EQ2:
                    k = k + ((a == e2) * (b2 == e1)).sum(-1)
                var = 1.0
            elif op == cabi.K_NOISE:
                k = eye * (0.0 if (skip_pn and nd.flags & cabi.KF_PROCESS_NOISE) else 1.0)
            elif op == cabi.K_WN:
                k = eye if same else (diff == 0).sum(-1).astype(float)
            if nd.var_idx >= 0:
                g[nd.var_idx] = k
            vals.append(var * k)
        elif op == cabi.K_MAX:
            a, b = vals[nd.dim0], vals[nd.dim1]
            m = np.maximum(a, b)
            vals.append(m)
            for i, d in grads[nd.dim0].items():
                g[i] = g.get(i, 0) + d * (m == a)
            for i, d in grads[nd.dim1].items():
                g[i] = g.get(i, 0) + d * (m == b)
        elif op in (cabi.K_SUM, cabi.K_PROD):
            a, b = vals[nd.dim0], vals[nd.dim1]
            vals.append(a + b if op == cabi.K_SUM else a * b)
            for i, d in grads[nd.dim0].items():
                g[i] = g.get(i, 0) + (d if op == cabi.K_SUM else d * b)
            for i, d in grads[nd.dim1].items():
                g[i] = g.get(i, 0) + (d if op == cabi.K_SUM else a * d)
        else:
            c = vals[nd.dim0]
            vals.append(nd.value * c if op == cabi.K_SCALE else nd.value + c)
            for i, d in grads[nd.dim0].items():
                g[i] = d * nd.value if op == cabi.K_SCALE else d
        grads.append(g)
    K = np.where(np.isnan(vals[-1]), 0.0, np.where(np.isinf(vals[-1]), 1e10, vals[-1]))
    return (K, grads[-1]) if grad else K


class FakeContext:
    def __init__(self, jitter=float(np.float32(1e-6))):
        self.jitter = jitter
        self.N = self.D = 0
        self._data_tag = None
        self.calls = 0

    def set_jitter(self, j, tries=20):
        self.jitter = j

    def set_data(self, X):
        self.X = np.array(X, dtype=np.float64)
        if self.X.ndim == 1:
            self.X = self.X[:, None]
        self.N, self.D = self.X.shape
        self._resident = None
        self.uploads = getattr(self, "uploads", 0) + 1

    def set_data_if_changed(self, X):
        X = np.asarray(X, dtype=np.float64)
        if X.ndim == 1:
            X = X[:, None]
        if getattr(self, "X", None) is None or self.X.shape != X.shape or not np.array_equal(self.X, X):
            self.set_data(X)
            self._data_tag = None

    # the resident-factor protocol of _cabi.Context (g3_gp_grad_resume): same call surface, recomputed here
    def eval_key(self, desc, kind, delta, theta, nu):
        return (bytes(desc), int(kind), np.asarray(delta, dtype=np.float64).tobytes(),
                np.atleast_2d(np.asarray(theta, dtype=np.float64)).tobytes(),
                None if nu is None else np.asarray(nu, dtype=np.float64).tobytes(), self.X.tobytes())

    def resident_matches(self, desc, kind, delta, theta, nu):
        r = getattr(self, "_resident", None)
        return r is not None and r[0] == self.eval_key(desc, kind, delta, theta, nu)

    def gp_grad_resume(self):
        key, args = self._resident
        self._resident = None
        self.resumed = getattr(self, "resumed", 0) + 1
        self.calls -= 1                                   # not a new factorisation
        r = self.gp_logp_grad(*args, want_grad=True)
        return r["dtheta"], r["ddelta"]

    def gram(self, desc, X1, X2, theta):
        theta = np.atleast_2d(theta)
        same = X2 is None
        X1 = np.asarray(X1, dtype=np.float64)
        X2 = X1 if same else np.asarray(X2, dtype=np.float64)
        return np.stack([_eval_desc(desc, t, X1, X2, same) for t in theta]), np.zeros(len(theta), dtype=np.int32)

    def gram_vjp(self, desc, X1, X2, theta, W):
        theta = np.atleast_2d(theta)
        same = X2 is None
        X1 = np.asarray(X1, dtype=np.float64)
        X2 = X1 if same else np.asarray(X2, dtype=np.float64)
        W = np.asarray(W, dtype=np.float64).reshape(len(theta), X1.shape[0], X2.shape[0])
        out = np.zeros((len(theta), desc.n_theta))
        for b, t in enumerate(theta):
            _, dK = _eval_desc(desc, t, X1, X2, same, grad=True)
            for i, g in dK.items():
                out[b, i] = np.sum(W[b] * g)
        return out

    def potrf_robust(self, A):
        from oracle import g3_oracle as orc
        L, info = orc.cholesky_robust(np.asarray(A, dtype=np.float64), return_info=True)
        return L, (-1 if info < 0 else 0), 0.0

    def potrf_robust_solve(self, A, rhs):
        L, info, jit = self.potrf_robust(A)
        return L, info, jit, sla.solve_triangular(L, np.asarray(rhs, dtype=np.float64), lower=True)

    def _factor(self, desc, t):
        K, dK = _eval_desc(desc, t, self.X, self.X, True, grad=True)
        m = np.min(np.diag(K))
        if not m > 0:
            K = K + (self.jitter - m) * np.eye(len(K))
        return K, dK, sla.cholesky(K, lower=True)

    def gp_logp_grad(self, desc, kind, delta, theta, nu=None, want_grad=True):
        self.calls += 1
        self._resident = None if want_grad else (self.eval_key(desc, kind, delta, theta, nu), (desc, kind, delta, theta, nu))
        theta = np.atleast_2d(theta)
        B = len(theta)
        delta = np.asarray(delta, dtype=np.float64)
        out = {"beta": np.zeros(B), "logdet": np.zeros(B), "status": np.zeros(B, dtype=np.int32),
               "dtheta": np.zeros((B, desc.n_theta)) if want_grad else None,
               "ddelta": np.zeros((B, self.N)) if want_grad else None}
        for b in range(B):
            d = delta if delta.ndim == 1 else delta[b]
            K, dK, L = self._factor(desc, theta[b])
            u = sla.solve_triangular(L, d, lower=True)
            out["beta"][b] = u @ u
            out["logdet"][b] = np.log(np.diag(L)).sum()
            if want_grad:
                al = sla.solve_triangular(L.T, u, lower=False)
                c = 1.0 if kind == cabi.KIND_GAUSS else (nu[b] + self.N) / (nu[b] - 2 + out["beta"][b])
                W = 0.5 * (c * np.outer(al, al) - sla.cho_solve((L, True), np.eye(self.N)))
                for i, g in dK.items():
                    out["dtheta"][b, i] = np.sum(W * g)
                out["ddelta"][b] = -c * al
        return out

    def gp_posterior(self, desc, Xs, delta, theta, noise=False, cov=False):
        theta = np.asarray(theta, dtype=np.float64).ravel()
        K, _, L = self._factor(desc, theta)
        Xs = np.asarray(Xs, dtype=np.float64)
        Ks = _eval_desc(desc, theta, Xs, self.X, False, skip_pn=not noise)
        Kss = _eval_desc(desc, theta, Xs, Xs, True, skip_pn=not noise)
        V = sla.solve_triangular(L, Ks.T, lower=True)
        u = sla.solve_triangular(L, delta, lower=True)
        C = Kss - V.T @ V
        return {"mean": V.T @ u, "var": np.maximum(np.diag(C), 0.0), "cov": C if cov else None, "beta": float(u @ u), "status": 0}
