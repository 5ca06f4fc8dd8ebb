"""NumPy/SciPy fp64 restatement of g3py's exact-GP hot path (TEST INFRASTRUCTURE).

Every function cites the reference file:line (relative to /root/reference) it
follows.  The reference builds Theano expressions; here the same arithmetic is
executed eagerly with NumPy, in the same order of operations where that matters
(N1 x N2 x D broadcast metric, tt_to_cov, dpotrf + jitter ladder, triangular
solve, guard scans, Murray reverse-mode Cholesky gradient).

A model is described by a neutral nested-dict *spec* (see `build_process`), so
that tests can hand the identical description to the oracle and to the CUDA
front-end without either importing the other.

Status of the pin.  This module is TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and the
CPU-baseline legs of bench.py may import it; the product (g3py_b200/) never does.  PARITY PINNED against outputs
of the reference itself: tests/golden/make_reference_goldens.py imports the unmodified /root/reference/g3py in the
build container through a stand-in for the Theano / PyMC3 API (tests/golden/refshim: lazy graph evaluated with torch
CPU fp64, reverse-mode autodiff, the reference's own CholeskyRobust perform/grad) and records logp, loglike, dlogp,
Gram matrices, posterior location / covariance, Gauss-Hermite moments, quantiles, transports for 30 cases in
tests/golden/reference_g3py.json; tests/test_reference_goldens.py holds this module to them (logp 1e-12, gradient
1e-10, posterior 1e-10).  The stand-in itself reproduces the values a real Theano run printed in
notebooks/07-Student-t-Process.ipynb:206-218 (tests/golden/reference_shim_check.json), which tests/test_oracle.py
also checks directly.  Not reproducible by the stand-in: Theano's graph optimiser, its
BLAS/LAPACK build, the traversal order of dlogp components, the max/min reduction gradient at ties.  Further
cross-checks: finite differences, torch fp64 autograd of an independently written forward, LU-vs-Cholesky posterior.

Constants mode (`strict=True`, default): the float32-rounded literals that stay
in the reference graph when it is run with floatX='float64' are reproduced
(gaussian.py:218, studentT.py:122-128, tensors.py:98,204).  `strict=False` uses
exact doubles.
"""
import math

import numpy as np
import scipy.linalg as sla
from scipy.special import gammaln, digamma

__all__ = [
    "tt_to_num", "tt_to_cov", "cholesky_robust", "murray_cholesky_grad",
    "Consts", "build_kernel", "build_process", "OracleProcess",
    "c2_inputs", "c1_inputs", "c3_inputs", "c4_inputs", "c5_inputs",
]

f32 = np.float32


# --------------------------------------------------------------------------- constants
class Consts:
    """Scalar constants of the reference graph (strict) or exact doubles."""

    def __init__(self, strict=True):
        self.strict = strict
        if strict:
            # gaussian.py:218  tt.log(np.float32(2.0*np.pi)) -> float32 log of a float32 constant, folded by
            # Theano's C thunk (glibc logf == double log rounded to float32 = 1.8378770351409912).  NumPy's
            # SIMD float32 log returns the neighbouring float (1.8378771543502808) on this CPU, i.e. the
            # constant is implementation-defined to 1 float32 ulp (6e-8 * N/2 on logp); the correctly
            # rounded value is used, computed in double so that it does not depend on the host's libm.
            self.log_2pi = float(f32(math.log(float(f32(2.0 * np.pi)))))
            # studentT.py:122-124
            self.pi = float(f32(np.pi))
            # studentT.py:127  np.log(np2 * npi) is a NumPy float32 expression (only used when nu >= 1e6);
            # same remark, same correctly rounded value
            self.log_2pi_student = float(f32(math.log(float(f32(2.0) * f32(np.pi)))))
            # tensors.py:98 and :204
            self.jitter = float(f32(1e-6))
            # gaussian.py:238-241 / studentT.py:143-146
            self.guard = float(f32(-1e30))
            # tensors.py:221
            self.fallback = float(f32(1e-10))
        else:
            self.log_2pi = math.log(2.0 * math.pi)
            self.pi = math.pi
            self.log_2pi_student = math.log(2.0 * math.pi)
            self.jitter = 1e-6
            self.guard = -1e30
            self.fallback = 1e-10


# --------------------------------------------------------------------------- tensors.py
def tt_to_num(r, nan=0.0, inf=1e10):
    """libs/tensors.py:90-92 — NaN -> 0, +inf AND -inf -> +1e10 (sign lost)."""
    r = np.asarray(r, dtype=np.float64)
    return np.where(np.isnan(r), nan, np.where(np.isinf(r), inf, r))


def tt_to_cov(c, consts=None):
    """libs/tensors.py:95-98 — scrub, then shift the diagonal if its minimum is <= 0."""
    consts = consts or Consts()
    r = tt_to_num(c)
    m = np.min(np.diag(r))
    if m > 0.0:
        return r
    return r + (consts.jitter - m) * np.eye(c.shape[0])


def cholesky_robust(K, consts=None, maxtries=20, return_info=False):
    """libs/tensors.py:197-222 — dpotrf, then the jitter ladder, then the 1e-10*I fallback.

    info: 0 = plain success, k>0 = succeeded at ladder try k (1-based), -1 = fallback.
    """
    consts = consts or Consts()
    K = np.asarray(K, dtype=np.float64)
    L, info = sla.lapack.dpotrf(K, lower=1)
    if info == 0:
        L = np.tril(L)
        return (L, 0) if return_info else L
    diagK = np.diag(K)
    n = K.shape[0]
    dK = np.eye(n) * diagK.mean() * consts.jitter                      # :204
    if np.any(diagK <= 0.0):                                            # :205-206
        K = K + np.eye(n) * (diagK.mean() * consts.jitter - diagK.min())
    for tries in range(maxtries):                                       # :208-212
        try:
            L = np.nan_to_num(sla.cholesky(K + dK, lower=True))
            return (L, tries + 1) if return_info else L
        except Exception:
            dK = dK * 10.0
    L = 0 * K + consts.fallback * np.eye(n)                             # :221
    return (L, -1) if return_info else L


def murray_cholesky_grad(L, Lbar):
    """libs/tensors.py:224-260 — reverse-mode of L = chol(K) (Murray 2016, level-3 form).

    Returns the lower-triangular Kbar the reference returns:
    tril(S + S^T) - diag(diag(S)),  S = L^-T tril_half(L^T Lbar) L^-1.
    """
    def tril_and_halve_diagonal(m):
        return np.tril(m) - np.diag(np.diagonal(m) / 2.0)

    def conjugate_solve_triangular(outer, inner):
        # L^-T P L^-1 via two upper-triangular solves with outer.T (:245-248)
        a = sla.solve_triangular(outer.T, tt_to_num(inner).T, lower=False)
        return sla.solve_triangular(outer.T, a.T, lower=False)

    s = conjugate_solve_triangular(L, tril_and_halve_diagonal(tt_to_num(L.T.dot(Lbar))))
    return np.tril(s + s.T) - np.diag(np.diagonal(s))


# --------------------------------------------------------------------------- metrics.py
def _gram_broadcast(x1, x2, dims):
    """hypers/metrics.py:11-13 — (N1,1,D) - (1,N2,D): the materialised difference tensor."""
    a = x1[:, dims[0]:dims[1]][:, None, :]
    b = x2[:, dims[0]:dims[1]][None, :, :]
    return a - b


def ard_l2(x1, x2, dims, rate):
    """hypers/metrics.py:100-102 — dot((x1-x2)**2, 0.5*rate**2)."""
    return np.dot(_gram_broadcast(x1, x2, dims) ** 2, 0.5 * rate ** 2)


def ard_l1(x1, x2, dims, rate):
    """hypers/metrics.py:89-91 — dot(abs(x1-x2), rate)."""
    return np.dot(np.abs(_gram_broadcast(x1, x2, dims)), rate)


def delta_gram(x1, x2, dims):
    """hypers/metrics.py:30-35 — number of equal coordinates (0..D)."""
    return tt_to_num((_gram_broadcast(x1, x2, dims) == 0.0).sum(axis=2).astype(np.float64))


# --------------------------------------------------------------------------- kernels.py
class _Hyper:
    def __init__(self, name, size, positive):
        self.name, self.size, self.positive = name, int(size), bool(positive)


class KernelNode:
    """Base: `hypers` in pymc3 creation order; cov()/dcov() take the natural-space values."""
    hypers = ()

    def layout(self):
        return list(self.hypers)

    def n_theta(self):
        return sum(h.size for h in self.layout())


class _Leaf(KernelNode):
    def __init__(self, spec, D):
        self.kind = spec["type"]
        self.name = spec.get("name", self.kind)
        dims = spec.get("dims")
        self.dims = (0, D) if dims is None else (int(dims[0]), int(dims[1]))
        self.nd = self.dims[1] - self.dims[0]
        self.fixed_var = spec.get("var")                      # KernelProd second var = 1.0 (kernels.py:215-219)
        if self.kind in ("LIN", "POL") and "var" not in spec:
            self.fixed_var = 1.0                               # kernels.py:320,325: LIN / POL default var=1
        if self.kind in ("NIL", "KernelEquals", "KernelEquals2"):
            self.fixed_var = 1.0                               # kernels.py:264,275,312: var=1, and cov() never multiplies by it
        self.eq = (float(spec.get("eq", spec.get("eq1", 0.0))), float(spec.get("eq2", spec.get("eq", 0.0))))
        hy = []
        if self.fixed_var is None:
            hy.append(_Hyper(self.name + "_var", 1, True))     # kernels.py:22-24
        k = self.kind
        if k in ("SE", "MAT32", "MAT52", "RQ", "OU"):
            hy.append(_Hyper(self.name + "_rate", self.nd, True))  # metrics.py:79-83
        if k == "RQ":
            hy.append(_Hyper(self.name + "_alpha", 1, True))   # kernels.py:394-397
        if k in ("SIN", "SM"):                                 # kernels.py:446-454: freq created before rate
            hy.append(_Hyper(self.name + "_freq", self.nd, True))
            hy.append(_Hyper(self.name + "_rate", self.nd, True))
        if k in ("COS", "SINC"):                               # kernels.py:463,476: rate=1.0 constant, freq only
            hy.append(_Hyper(self.name + "_freq", self.nd, True))
        if k in ("KernelDot", "LIN", "POL", "NN"):             # metrics.py:79-83 (ARD.rate), :126-128 (bias after rate)
            hy.append(_Hyper(self.name + "_rate", self.nd, True))
            if k != "KernelDot":
                hy.append(_Hyper(self.name + "_bias", 1, True))
        self.power = int(spec.get("p", 2)) if k == "POL" else 1  # kernels.py:325-327
        self.hypers = tuple(hy)

    def _split(self, th):
        i = 0
        out = {}
        if self.fixed_var is None:
            out["var"] = th[0]
            i = 1
        else:
            out["var"] = float(self.fixed_var)
        for h in self.hypers:
            key = h.name[len(self.name) + 1:]
            if key == "var":
                continue
            out[key] = th[i:i + h.size] if h.size > 1 or key in ("rate", "freq") else th[i]
            i += h.size
        return out

    # value of k(d) and the pieces the gradient needs -------------------------------------
    def cov(self, th, x1, x2, same):
        p = self._split(th)
        k = self.kind
        n1, n2 = x1.shape[0], x2.shape[0]
        if k == "Noise":                                       # kernels.py:367-371
            return p["var"] * np.eye(n1) if same else np.zeros((n1, n2))
        if k == "WN":                                          # kernels.py:381-385
            return p["var"] * np.eye(n1) if same else p["var"] * delta_gram(x1, x2, self.dims)
        if k == "SE":                                          # kernels.py:424-436 exp(-d), ARD_L2
            return p["var"] * np.exp(-ard_l2(x1, x2, self.dims, p["rate"]))
        if k == "OU":                                          # kernels.py:429-431 exp(-d), ARD_L1
            return p["var"] * np.exp(-ard_l1(x1, x2, self.dims, p["rate"]))
        if k == "MAT32":                                       # kernels.py:406-412
            d3 = np.sqrt(3 * ard_l2(x1, x2, self.dims, p["rate"]))
            return p["var"] * ((1 + d3) * np.exp(-d3))
        if k == "MAT52":                                       # kernels.py:415-421
            d = ard_l2(x1, x2, self.dims, p["rate"])
            d5 = np.sqrt(5 * d)
            return p["var"] * ((1 + d5 + 5 * d / 3) * np.exp(-d5))
        if k == "RQ":                                          # kernels.py:388-403
            d = ard_l2(x1, x2, self.dims, p["rate"])
            return p["var"] * np.power(1 + d / p["alpha"], -p["alpha"])
        if k == "SIN":                                         # kernels.py:470-472 (positive exponent, as written)
            diff = _gram_broadcast(x1, x2, self.dims)
            return p["var"] * np.exp(2 * np.dot(np.sin(np.pi * diff * p["freq"]) ** 2, p["rate"]))
        if k == "COS":                                         # kernels.py:466-467
            diff = _gram_broadcast(x1, x2, self.dims)
            return p["var"] * np.prod(np.cos(2 * np.pi * diff * p["freq"]), axis=2)
        if k == "SINC":                                        # kernels.py:479-482 (pi2 = pi**2, as written)
            diff = _gram_broadcast(x1, x2, self.dims)
            pi2 = np.pi ** 2
            with np.errstate(invalid="ignore", divide="ignore"):
                sinc = np.sin(2 * pi2 * diff * p["freq"]) / (2 * pi2 * p["freq"] * diff)
            return p["var"] * np.prod(np.where(diff != 0.0, sinc, 1.0), axis=2)
        if k == "SM":                                          # kernels.py:486-487
            diff = _gram_broadcast(x1, x2, self.dims)
            pi2 = np.pi ** 2
            return p["var"] * (np.exp(-2 * pi2 * np.dot(diff ** 2, p["rate"] ** 2))
                               * np.prod(np.cos(2 * np.pi * diff * p["freq"]), axis=2))
        if k in ("KernelDot", "LIN", "POL"):                   # kernels.py:82-96,319-336; metrics.py:110-137
            return p["var"] * self._dot_metric(p, x1, x2) ** self.power
        if k == "BW":                                          # kernels.py:291-293; metrics.py:49-51 Minimum
            a = x1[:, self.dims[0]:self.dims[1]][:, None, :]
            b = x2[:, self.dims[0]:self.dims[1]][None, :, :]
            return p["var"] * np.prod(np.minimum(a - b * 0, b - a * 0), axis=2)
        if k == "VAR":                                         # kernels.py:296-306
            return p["var"] * np.ones((n1, n2))
        if k == "NIL":                                         # kernels.py:309-320: zeros
            return np.zeros((n1, n2))
        if k in ("KernelEquals", "KernelEquals2"):             # kernels.py:262-288; metrics.py:38-51 DeltaEq / DeltaEq2
            a = x1[:, self.dims[0]:self.dims[1]][:, None, :]
            b = x2[:, self.dims[0]:self.dims[1]][None, :, :]
            e1, e2 = self.eq
            if k == "KernelEquals":
                return tt_to_num(((a == e1) * (b == e1)).sum(axis=2).astype(np.float64))
            return tt_to_num(((a == e1) * (b == e2) + (a == e2) * (b == e1)).sum(axis=2).astype(np.float64))
        if k == "NN":                                          # kernels.py:339-351.  cov(x1) (x2 is None) is elementwise in the Gram
            if not same:                                       # xx_ij; the two-argument form multiplies an N1xN1 by an N2xN2 matrix
                raise NotImplementedError("NN.cov(x1, x2): the reference form only broadcasts when N1 == N2 (kernels.py:351)")
            xx = self._dot_metric(p, x1, x2)
            return p["var"] * np.arcsin(2 * xx / (1 + 2 * xx) ** 2)
        raise ValueError(k)

    def _dot_metric(self, p, x1, x2):
        a = x1[:, self.dims[0]:self.dims[1]][:, None, :]
        b = x2[:, self.dims[0]:self.dims[1]][None, :, :]
        m = np.dot(a * b, p["rate"] ** 2)                      # metrics.py:111-112 ARD_Dot
        return m if self.kind == "KernelDot" else p["bias"] + m    # metrics.py:131-132 ARD_DotBias (LIN, POL, NN)

    def dcov(self, th, x1, x2, same, nan_quirk=False):
        """List of dK/dtheta_p (natural space), one N1xN2 array per scalar hyper, layout order.

        Analytic limits are used at d = 0 for the sqrt-kernels.  With
        `nan_quirk=True` the reference's behaviour is emulated instead: Theano's grad(sqrt) gives
        NaN where d == 0, `tt_to_num` (stochastic.py:309) turns the *whole* rate-gradient
        component NaN -> 0.  The same happens to the `freq` gradient of SINC: the unselected
        branch of its `switch` is sin(0)/0 (kernels.py:479-480) and its gradient is NaN at delta = 0.
        Both are confirmed by executing the reference (tests/golden/reference_g3py.json).
        """
        p = self._split(th)
        k = self.kind
        out = []
        K = self.cov(th, x1, x2, same)
        if self.fixed_var is None:
            out.append(K / p["var"])
        if k in ("Noise", "WN", "BW", "VAR", "NIL", "KernelEquals", "KernelEquals2"):
            return out
        if k in ("KernelDot", "LIN", "POL", "NN"):
            a = x1[:, self.dims[0]:self.dims[1]][:, None, :]
            b = x2[:, self.dims[0]:self.dims[1]][None, :, :]
            m = self._dot_metric(p, x1, x2)
            if k == "NN":                                      # u = 2m / (1 + 2m)^2, d arcsin(u) = du / sqrt(1 - u^2)
                u = 2 * m / (1 + 2 * m) ** 2
                dm = p["var"] * (2 * (1 - 2 * m) / (1 + 2 * m) ** 3) / np.sqrt(1 - u * u)
            else:
                dm = p["var"] * self.power * m ** (self.power - 1)
            for j in range(self.nd):
                out.append(dm * 2 * p["rate"][j] * a[:, :, j] * b[:, :, j])
            if k != "KernelDot":
                out.append(dm)
            return out
        diff = _gram_broadcast(x1, x2, self.dims)
        if k in ("SE", "MAT32", "MAT52", "RQ"):
            r = p["rate"]
            d = np.dot(diff ** 2, 0.5 * r ** 2)
            if k == "SE":
                dkdd = -np.exp(-d)
            elif k == "MAT32":
                dkdd = -1.5 * np.exp(-np.sqrt(3 * d))
            elif k == "MAT52":
                s = np.sqrt(5 * d)
                dkdd = -(5.0 / 6.0) * (1 + s) * np.exp(-s)
            else:
                a = p["alpha"]
                dkdd = -np.power(1 + d / a, -a - 1)
            for j in range(self.nd):
                g = p["var"] * dkdd * (r[j] * diff[:, :, j] ** 2)
                if nan_quirk and k in ("MAT32", "MAT52") and np.any(d == 0.0):
                    g = np.full_like(g, np.nan)
                out.append(g)
            if k == "RQ":
                a = p["alpha"]
                out.append(K * (-np.log1p(d / a) + d / (a + d)))
            return out
        if k == "OU":
            r = p["rate"]
            for j in range(self.nd):
                out.append(-K * np.abs(diff[:, :, j]))
            return out
        if k == "SIN":
            fr, r = p["freq"], p["rate"]
            for j in range(self.nd):                          # freq first (creation order)
                out.append(K * 2 * r[j] * np.sin(2 * np.pi * diff[:, :, j] * fr[j]) * np.pi * diff[:, :, j])
            for j in range(self.nd):
                out.append(K * 2 * np.sin(np.pi * diff[:, :, j] * fr[j]) ** 2)
            return out
        if k in ("COS", "SINC", "SM"):
            fr = p["freq"]
            pi2 = np.pi ** 2
            if k == "SINC":
                b = 2 * pi2 * diff * fr
                with np.errstate(invalid="ignore", divide="ignore"):
                    fac = np.where(diff != 0.0, np.sin(b) / b, 1.0)
                    dfac = np.where(diff != 0.0, (np.cos(b) - fac) / fr, 0.0)
            else:
                a = 2 * np.pi * diff * fr
                fac = np.cos(a)
                dfac = -np.sin(a) * 2 * np.pi * diff
            env = p["var"] * (np.exp(-2 * pi2 * np.dot(diff ** 2, p["rate"] ** 2)) if k == "SM" else 1.0)
            for j in range(self.nd):                          # freq first (creation order)
                others = np.prod(np.delete(fac, j, axis=2), axis=2)
                g = env * dfac[:, :, j] * others
                if nan_quirk and k == "SINC" and np.any(diff[:, :, j] == 0.0):
                    g = np.full_like(g, np.nan)
                out.append(g)
            if k == "SM":
                for j in range(self.nd):
                    out.append(K * (-4 * pi2 * diff[:, :, j] ** 2 * p["rate"][j]))
            return out
        raise ValueError(k)


class _Binary(KernelNode):
    """KernelSum / KernelProd (kernels.py:213-245)."""

    def __init__(self, op, k1, k2):
        self.op, self.k1, self.k2 = op, k1, k2

    def layout(self):
        return self.k1.layout() + self.k2.layout()

    def cov(self, th, x1, x2, same):
        n1 = self.k1.n_theta()
        a = self.k1.cov(th[:n1], x1, x2, same)
        b = self.k2.cov(th[n1:], x1, x2, same)
        if self.op == "max":                                    # kernels.py:247-257
            return np.maximum(a, b)
        return a + b if self.op == "sum" else a * b

    def dcov(self, th, x1, x2, same, nan_quirk=False):
        n1 = self.k1.n_theta()
        da = self.k1.dcov(th[:n1], x1, x2, same, nan_quirk)
        db = self.k2.dcov(th[n1:], x1, x2, same, nan_quirk)
        if self.op == "sum":
            return da + db
        a = self.k1.cov(th[:n1], x1, x2, same)
        b = self.k2.cov(th[n1:], x1, x2, same)
        if self.op == "max":                                    # Theano: grad of maximum = eq(out, x) * g (both on ties)
            m = np.maximum(a, b)
            return [g * (m == a) for g in da] + [g * (m == b) for g in db]
        return [g * b for g in da] + [a * g for g in db]


class _Unary(KernelNode):
    """KernelScale / KernelShift with a constant element (kernels.py:192-210)."""

    def __init__(self, op, c, k):
        self.op, self.c, self.k = op, float(c), k

    def layout(self):
        return self.k.layout()

    def cov(self, th, x1, x2, same):
        a = self.k.cov(th, x1, x2, same)
        return self.c * a if self.op == "scale" else self.c + a

    def dcov(self, th, x1, x2, same, nan_quirk=False):
        d = self.k.dcov(th, x1, x2, same, nan_quirk)
        return [self.c * g for g in d] if self.op == "scale" else d


def build_kernel(spec, D):
    t = spec["type"]
    if t in ("sum", "prod", "max"):
        k2s = dict(spec["k2"])
        leaf = lambda sp: sp["type"] not in ("sum", "prod", "max", "scale", "shift")
        # kernels.py:215-219: a product of two leaf kernels that both have var=None fixes k2.var = 1.0
        if t == "prod" and leaf(spec["k1"]) and leaf(k2s) and spec["k1"].get("var") is None and k2s.get("var") is None:
            k2s["var"] = 1.0
        return _Binary(t, build_kernel(spec["k1"], D), build_kernel(k2s, D))
    if t in ("scale", "shift"):
        return _Unary(t, spec["c"], build_kernel(spec["k"], D))
    return _Leaf(spec, D)


# --------------------------------------------------------------------------- means.py
class _Mean:
    def __init__(self, spec, D):
        self.kind = spec["type"]
        self.name = spec.get("name", self.kind)
        dims = spec.get("dims")
        self.dims = (0, D) if dims is None else tuple(dims)
        nd = self.dims[1] - self.dims[0]
        if self.kind == "Zero":
            self.hypers = ()
        elif self.kind == "Bias":                              # means.py:127-131
            self.hypers = (_Hyper(self.name + "_Bias", 1, False),)
        elif self.kind == "Linear":                            # means.py:147-152
            self.hypers = (_Hyper(self.name + "_Constant", 1, False), _Hyper(self.name + "_Coeff", nd, False))
        elif self.kind == "Power":                             # means.py:162-182: constant + dot(x**n, coeff)
            self.hypers = (_Hyper(self.name + "_Constant", 1, False), _Hyper(self.name + "_Coeff", nd, False))
            self.power = spec.get("n", 2)
        elif self.kind == "BlackBox":                          # means.py:32-41: element[:len(x)], no hypers
            self.hypers = ()
            self.element = np.asarray(spec["element"], dtype=np.float64)
        else:
            raise ValueError(self.kind)

    def layout(self):
        return list(self.hypers)

    def n_theta(self):
        return sum(h.size for h in self.hypers)

    def __call__(self, th, x):
        n = x.shape[0]
        if self.kind == "Zero":                                # means.py:117-119
            return np.zeros(n)
        if self.kind == "Bias":                                # means.py:136-137
            return th[0] * np.ones(n)
        if self.kind == "BlackBox":                            # means.py:37-41
            return self.element[:n].copy()
        xs = x[:, self.dims[0]:self.dims[1]]
        if self.kind == "Power":                               # means.py:181-182
            return th[0] + (xs ** self.power).dot(th[1:])
        return th[0] + xs.dot(th[1:])                          # means.py:158-159

    def jac(self, th, x):
        """d mean / d theta: (P_mean, n)."""
        n = x.shape[0]
        if self.kind == "Zero":
            return np.zeros((0, n))
        if self.kind == "Bias":
            return np.ones((1, n))
        if self.kind == "BlackBox":
            return np.zeros((0, n))
        xs = x[:, self.dims[0]:self.dims[1]]
        if self.kind == "Power":
            return np.vstack([np.ones((1, n)), (xs ** self.power).T])
        return np.vstack([np.ones((1, n)), xs.T])


# --------------------------------------------------------------------------- mappings.py
class _Mapping:
    """Closed-form warpings: inv(y), logdet_dinv(y), forward T(z) and their hyper-derivatives."""

    def __init__(self, spec):
        self.kind = spec["type"]
        self.name = spec.get("name", "BoxShift" if self.kind == "BoxCoxShifted" else self.kind)
        H = lambda s, pos: _Hyper(self.name + "_" + s, 1, pos)
        self.n = int(spec.get("n", 1))
        Hn = lambda s, pos: _Hyper(self.name + "_" + s, self.n, pos)
        self.hypers = {
            "Identity": (),
            "LinearMapping": (H("shift", False), H("scale", True)),        # mappings.py:107-112
            "LogShifted": (H("shift", False),),                            # mappings.py:134-137
            "BoxCoxShifted": (H("shift", False), H("power", True)),        # mappings.py:158-163
            "BoxCoxLinear": (H("shift", False), H("scale", True), H("power", True)),  # :190-197
            "ArcsinhLinear": (H("shift", False), H("scale", True)),        # mappings.py:315-320
            "SinhArcsinh": (H("shift", False), H("scale", True)),          # mappings.py:340-345
            "Logistic": (H("lower", False), H("high", True), H("location", False), H("scale", True)),   # :369-377
            "WarpingTanh": (Hn("a", True), Hn("b", True), Hn("c", False)),                 # mappings.py:262-268
            "WarpingBoxCox": (Hn("shift", True), Hn("power", True), Hn("w", True)),        # mappings.py:290-296
        }[self.kind]

    def layout(self):
        return list(self.hypers)

    def n_theta(self):
        return sum(h.size for h in self.hypers)

    def _dinv_dy(self, th, y):
        """d inv / d y for the Newton-inverse warpings (what tt.grad(tt.sum(func(x) - z), x) evaluates to)."""
        n = self.n
        z = y[:, None]
        if self.kind == "WarpingTanh":
            a, b, c = th[:n], th[n:2 * n], th[2 * n:]
            return 1.0 + np.dot(1.0 / np.cosh(b * (z + c)) ** 2, a * b)
        shift, power, w = th[:n], th[n:2 * n], th[2 * n:]
        return np.dot(np.abs(z + shift) ** (power - 1.0), w)

    def dinv_dy(self, th, y):
        """d inv / d y of every closed form above (composition chains through it)."""
        k = self.kind
        if k in ("WarpingTanh", "WarpingBoxCox"):
            return self._dinv_dy(th, y)
        if k == "Identity":
            return np.ones_like(y)
        if k == "LinearMapping":
            return np.ones_like(y) / th[1]
        if k == "LogShifted":
            return 1.0 / (y - th[0])
        if k == "BoxCoxShifted":
            return np.abs(y + th[0]) ** (th[1] - 1.0)
        if k == "BoxCoxLinear":
            return th[1] * np.abs(th[1] * (y + th[0])) ** (th[2] - 1.0)
        if k == "ArcsinhLinear":
            return th[1] / np.sqrt(1.0 + y * y)
        if k == "SinhArcsinh":
            return np.cosh(th[0] + th[1] * np.arcsinh(y)) * th[1] / np.sqrt(1.0 + y * y)
        if k == "Logistic":
            p = (y - th[0]) / th[1]
            return th[3] / (th[1] * p * (1 - p))
        raise ValueError(k)

    def dlog_dinv_dy(self, th, y):
        """d/dy log|d inv / d y|: the per-point slope of the log-Jacobian (the term a composed map needs when an inner
        map's hypers move the argument of the outer one)."""
        k = self.kind
        if k in ("Identity", "LinearMapping"):
            return np.zeros_like(y)
        if k == "LogShifted":
            return -1.0 / (y - th[0])
        if k == "BoxCoxShifted":
            return (th[1] - 1.0) / (y + th[0])
        if k == "BoxCoxLinear":
            return (th[2] - 1.0) / (y + th[0])
        if k == "ArcsinhLinear":
            return -y / (1.0 + y * y)
        if k == "SinhArcsinh":
            return np.tanh(th[0] + th[1] * np.arcsinh(y)) * th[1] / np.sqrt(1.0 + y * y) - y / (1.0 + y * y)
        if k == "Logistic":
            p = (y - th[0]) / th[1]
            return (1.0 / (1 - p) - 1.0 / p) / th[1]
        m = self.n
        if k == "WarpingTanh":
            a, b, c = th[:m], th[m:2 * m], th[2 * m:]
            u = b * (y[:, None] + c)
            return np.dot(-2.0 * np.tanh(u) / np.cosh(u) ** 2, a * b * b) / self._dinv_dy(th, y)
        shift, power, w = th[:m], th[m:2 * m], th[2 * m:]
        sh = y[:, None] + shift
        return np.dot((power - 1.0) * np.abs(sh) ** (power - 2.0) * np.sign(sh), w) / self._dinv_dy(th, y)

    def _newton_forward(self, th, z, tol=1e-3, n_steps=1024, alpha=0.1):
        """Mapping.__call__ = inverse_function(self.inv, z) (mappings.py:11-12, libs/tensors.py:134-145): damped
        Newton from 0, step alpha = 0.1, slopes below 1 replaced by their sign, stopped when max|inv(x) - z| < 1e-3 on
        the WHOLE vector -- restated literally (the result is only ~1e-3 accurate, by construction)."""
        x = 0.0 * z
        for _ in range(n_steps):
            diff = self.inv(th, x) - z
            dfunc = self._dinv_dy(th, x)
            dfunc = np.where(np.abs(dfunc) < 1.0, np.sign(dfunc), dfunc)
            x = x - alpha * diff / dfunc
            if np.max(np.abs(diff)) < tol:
                break
        return x

    def inv(self, th, y):
        k = self.kind
        if k == "Logistic":                                    # mappings.py:391-393
            with np.errstate(all="ignore"):
                p = np.where(y < th[0], 0.0, np.where(y > th[0] + th[1], 1.0, (y - th[0]) / th[1]))
                return th[2] + th[3] * tt_to_num(np.log(p / (1 - p)))
        if k == "WarpingTanh":                                 # mappings.py:275-278
            n = self.n
            a, b, c = th[:n], th[n:2 * n], th[2 * n:]
            return y + np.dot(np.tanh(b * (y[:, None] + c)), a)
        if k == "WarpingBoxCox":                               # mappings.py:303-306
            n = self.n
            shift, power, w = th[:n], th[n:2 * n], th[2 * n:]
            sh = y[:, None] + shift
            return np.dot((np.sign(sh) * np.abs(sh) ** power - 1.0) / power, w)
        if k == "Identity":
            return y                                           # mappings.py:95-96
        if k == "LinearMapping":
            return y / th[1] + th[0]                           # :121-122
        if k == "LogShifted":
            return np.log(np.maximum(y - th[0], float(f32(1e-32))))   # :145-146
        if k == "BoxCoxShifted":                               # :173-175
            sh = y + th[0]
            if th[1] < 1e-5:
                return np.log(sh)
            return (np.sign(sh) * np.abs(sh) ** th[1] - 1.0) / th[1]
        if k == "BoxCoxLinear":                                # :209-211
            sh = th[1] * (y + th[0])
            if th[2] < float(f32(1e-5)):
                return np.log(sh)
            return (np.sign(sh) * np.abs(sh) ** th[2] - 1.0) / th[2]
        if k == "ArcsinhLinear":
            return np.arcsinh(y) * th[1] + th[0]               # :329-330
        if k == "SinhArcsinh":
            return np.sinh(th[0] + th[1] * np.arcsinh(y))      # :354-355
        raise ValueError(k)

    def logdet_dinv(self, th, y):
        k = self.kind
        n = float(y.shape[0])
        if k == "Logistic":                                    # mappings.py:395-397
            with np.errstate(all="ignore"):
                p = np.where(y < th[0], 0.0, np.where(y > th[0] + th[1], 1.0, (y - th[0]) / th[1]))
                return float(np.sum(tt_to_num(np.log(th[3] / (th[1] * p * (1 - p))))))
        if k in ("WarpingTanh", "WarpingBoxCox"):              # mappings.py:19-23: sum log diag(jacobian(inv))
            with np.errstate(all="ignore"):
                return float(np.sum(np.log(self._dinv_dy(th, y))))
        if k == "Identity":
            return 0.0                                         # :98-99
        if k == "LinearMapping":
            return -n * np.log(th[1])                          # :124-125
        if k == "LogShifted":
            return -np.sum(np.log(y - th[0]))                  # :148-149
        if k == "BoxCoxShifted":
            return (th[1] - 1.0) * np.sum(np.log(np.abs(y + th[0])))   # :177-179
        if k == "BoxCoxLinear":                                # :213-215
            return (th[2] - 1.0) * np.sum(np.log(np.abs(th[1] * (y + th[0])))) + n * np.log(th[1])
        if k == "ArcsinhLinear":
            return n * np.log(th[1]) - 0.5 * np.sum(np.log1p(y ** 2))  # :332-333
        if k == "SinhArcsinh":                                 # :357-358
            return (np.sum(np.log(np.cosh(th[0] + th[1] * np.arcsinh(y)))) + n * np.log(th[1])
                    - 0.5 * np.sum(np.log1p(y ** 2)))
        raise ValueError(k)

    def forward(self, th, z):
        k = self.kind
        if k == "Logistic":                                    # mappings.py:388-389
            return th[0] + th[1] * (0.5 + 0.5 * np.tanh((z - th[2]) / (2 * th[3])))
        if k in ("WarpingTanh", "WarpingBoxCox"):
            return self._newton_forward(th, z)
        if k == "Identity":
            return z
        if k == "LinearMapping":
            return th[1] * (z - th[0])                         # :118-119
        if k == "LogShifted":
            return np.exp(z) + th[0]                           # :142-143
        if k == "BoxCoxShifted":                               # :168-171
            sc = th[1] * z + 1.0
            return np.sign(sc) * np.abs(sc) ** (1.0 / th[1]) - th[0]
        if k == "BoxCoxLinear":                                # :204-207
            sc = th[2] * z + 1.0
            return np.sign(sc) * np.abs(sc) ** (1.0 / th[2]) / th[1] - th[0]
        if k == "ArcsinhLinear":
            return np.sinh((z - th[0]) / th[1])                # :326-327
        if k == "SinhArcsinh":
            return np.sinh((np.arcsinh(z) - th[0]) / th[1])    # :351-352
        raise ValueError(k)

    def grads(self, th, y):
        """Analytic (d inv / d theta [P_map, n], d logdet / d theta [P_map]) of the closed forms above."""
        k = self.kind
        n = float(y.shape[0])
        one = np.ones_like(y)
        if k == "Identity":
            return np.zeros((0, y.shape[0])), np.zeros(0)
        if k == "LinearMapping":
            return np.vstack([one, -y / th[1] ** 2]), np.array([0.0, -n / th[1]])
        if k == "LogShifted":
            live = (y - th[0]) > float(f32(1e-32))
            return np.vstack([np.where(live, -1.0 / (y - th[0]), 0.0)]), np.array([np.sum(1.0 / (y - th[0]))])
        if k in ("BoxCoxShifted", "BoxCoxLinear"):
            if k == "BoxCoxShifted":
                shift, scale, p = th[0], 1.0, th[1]
            else:
                shift, scale, p = th[0], th[1], th[2]
            sh = scale * (y + shift)
            a = np.abs(sh)
            sp = np.sign(sh) * a ** p
            d_shift = a ** (p - 1.0) * scale
            d_scale = a ** (p - 1.0) * (y + shift)
            d_p = (sp * np.log(a) * p - (sp - 1.0)) / p ** 2
            ld_shift = (p - 1.0) * np.sum(1.0 / (y + shift))
            ld_p = np.sum(np.log(a))
            if k == "BoxCoxShifted":
                return np.vstack([d_shift, d_p]), np.array([ld_shift, ld_p])
            return np.vstack([d_shift, d_scale, d_p]), np.array([ld_shift, p * n / scale, ld_p])
        if k == "ArcsinhLinear":
            return np.vstack([one, np.arcsinh(y)]), np.array([0.0, n / th[1]])
        if k == "SinhArcsinh":
            w = th[0] + th[1] * np.arcsinh(y)
            return (np.vstack([np.cosh(w), np.cosh(w) * np.arcsinh(y)]),
                    np.array([np.sum(np.tanh(w)), np.sum(np.tanh(w) * np.arcsinh(y)) + n / th[1]]))
        if k == "Logistic":                                    # interior points (0 < p < 1)
            lower, high, loc, scale = th
            p = (y - lower) / high
            q = np.log(p / (1 - p))
            dq = 1.0 / (p * (1 - p))
            dl = -1.0 / p + 1.0 / (1 - p)                      # d(-log p - log(1-p))/dp
            return (np.vstack([scale * dq * (-1.0 / high), scale * dq * (-p / high), one, q]),
                    np.array([np.sum(dl * (-1.0 / high)), np.sum(-1.0 / high + dl * (-p / high)), 0.0, n / scale]))
        if k == "WarpingTanh":
            m = self.n
            a, b, c = th[:m], th[m:2 * m], th[2 * m:]
            u = b * (y[:, None] + c)
            t, s2 = np.tanh(u), 1.0 / np.cosh(u) ** 2
            Dy = 1.0 + np.dot(s2, a * b)
            dinv = np.vstack([t.T, (a * s2 * (y[:, None] + c)).T, (a * b * s2).T])
            dD = np.vstack([(b * s2).T, (a * s2 - 2.0 * a * b * s2 * t * (y[:, None] + c)).T,
                            (-2.0 * a * b * b * s2 * t).T])
            return dinv, np.sum(dD / Dy, axis=1)
        if k == "WarpingBoxCox":
            m = self.n
            shift, power, w = th[:m], th[m:2 * m], th[2 * m:]
            sh = y[:, None] + shift
            ab = np.abs(sh)
            sp = np.sign(sh) * ab ** power
            Dy = np.dot(ab ** (power - 1.0), w)
            dinv = np.vstack([(w * ab ** (power - 1.0)).T, (w * (sp * np.log(ab) * power - (sp - 1.0)) / power ** 2).T,
                              ((sp - 1.0) / power).T])
            dD = np.vstack([(w * (power - 1.0) * ab ** (power - 2.0) * np.sign(sh)).T,
                            (w * ab ** (power - 1.0) * np.log(ab)).T, (ab ** (power - 1.0)).T])
            return dinv, np.sum(dD / Dy, axis=1)
        raise ValueError(k)

    def grads_fd(self, th, y):
        """4th-order central differences of inv / logdet_dinv (used by tests to check `grads`)."""
        P = self.n_theta()
        dinv = np.zeros((P, y.shape[0]))
        dld = np.zeros(P)
        for i in range(P):
            h = 1e-4 * max(1.0, abs(th[i]))
            def at(s):
                t = np.array(th, dtype=np.float64)
                t[i] += s * h
                return self.inv(t, y), self.logdet_dinv(t, y)
            (a2, b2), (a1, b1), (c1, d1), (c2, d2) = at(2), at(1), at(-1), at(-2)
            dinv[i] = (-a2 + 8 * a1 - 8 * c1 + c2) / (12 * h)
            dld[i] = (-b2 + 8 * b1 - 8 * d1 + d2) / (12 * h)
        return dinv, dld


class _MappingComposed:
    """m1 @ m2 (processes/hypers/mappings.py:57-70): inv(y) = m2.inv(m1.inv(y)), logdet = m2.logdet(m1.inv(y)) +
    m1.logdet(y), forward T(z) = m1(m2(z)); hypers in the order m1's then m2's (MappingOperation.check_hypers :40-43)."""

    def __init__(self, spec):
        self.kind = "composed"
        self.m1, self.m2 = make_mapping(spec["m1"]), make_mapping(spec["m2"])
        self.name = self.m1.name + " " + self.m2.name
        self.hypers = tuple(self.m1.layout()) + tuple(self.m2.layout())
        self.n = 1

    def layout(self):
        return list(self.hypers)

    def n_theta(self):
        return self.m1.n_theta() + self.m2.n_theta()

    def _split(self, th):
        return th[:self.m1.n_theta()], th[self.m1.n_theta():]

    def inv(self, th, y):
        t1, t2 = self._split(th)
        return self.m2.inv(t2, self.m1.inv(t1, y))

    def logdet_dinv(self, th, y):
        t1, t2 = self._split(th)
        return self.m2.logdet_dinv(t2, self.m1.inv(t1, y)) + self.m1.logdet_dinv(t1, y)

    def forward(self, th, z):
        t1, t2 = self._split(th)
        return self.m1.forward(t1, self.m2.forward(t2, z))

    def dinv_dy(self, th, y):
        t1, t2 = self._split(th)
        return self.m2.dinv_dy(t2, self.m1.inv(t1, y)) * self.m1.dinv_dy(t1, y)

    def dlog_dinv_dy(self, th, y):
        t1, t2 = self._split(th)
        return self.m2.dlog_dinv_dy(t2, self.m1.inv(t1, y)) * self.m1.dinv_dy(t1, y) + self.m1.dlog_dinv_dy(t1, y)

    def grads(self, th, y):
        t1, t2 = self._split(th)
        w = self.m1.inv(t1, y)
        di1, dl1 = self.m1.grads(t1, y)
        di2, dl2 = self.m2.grads(t2, w)
        outer = self.m2.dinv_dy(t2, w)                         # d m2.inv / d w
        slope = self.m2.dlog_dinv_dy(t2, w)                    # d log|d m2.inv / d w| / d w
        return (np.vstack([di1 * outer[None, :], di2]),
                np.concatenate([dl1 + np.sum(di1 * slope[None, :], axis=1), dl2]))

    grads_fd = None                                            # bound below (same finite differences as _Mapping)


class _MappingBoxCoxLinear2:
    """BoxCoxLinear2 (processes/hypers/mappings.py:218-251): shifted = scale * y + shift;
    inv = log(shifted) if power < float32(1e-5) else (sgn |shifted|^power - 1) / power;
    logdet_dinv = (power - 1 | -1) * sum(log|shifted|) + N log(scale); forward T(z) = (sgn|power z + 1|^(1/power) - shift) / scale."""

    def __init__(self, spec):
        self.kind = "BoxCoxLinear2"
        self.name = spec.get("name", "BoxCoxLinear2")
        self.hypers = (_Hyper(self.name + "_shift", 1, False), _Hyper(self.name + "_scale", 1, True),
                       _Hyper(self.name + "_power", 1, True))
        self.n = 1
        self.thr = float(np.float32(1e-5))

    def layout(self):
        return list(self.hypers)

    def n_theta(self):
        return 3

    def inv(self, th, y):
        shift, scale, power = th
        sh = scale * y + shift
        with np.errstate(all="ignore"):
            return np.log(sh) if power < self.thr else (np.sign(sh) * np.abs(sh) ** power - 1.0) / power

    def logdet_dinv(self, th, y):
        shift, scale, power = th
        with np.errstate(all="ignore"):
            return (-1.0 if power < self.thr else power - 1.0) * np.sum(np.log(np.abs(scale * y + shift))) + len(y) * np.log(scale)

    def forward(self, th, z):
        shift, scale, power = th
        sc = power * z + 1.0
        return (np.sign(sc) * np.abs(sc) ** (1.0 / power) - shift) / scale

    def dinv_dy(self, th, y):
        shift, scale, power = th
        return np.abs(scale * y + shift) ** (power - 1.0) * scale

    def dlog_dinv_dy(self, th, y):
        shift, scale, power = th
        return (power - 1.0) * scale / (scale * y + shift)

    def grads(self, th, y):
        shift, scale, power = th
        sh = scale * y + shift
        a = np.abs(sh)
        with np.errstate(all="ignore"):
            sp = np.sign(sh) * a ** power
            dinv = np.vstack([a ** (power - 1.0), a ** (power - 1.0) * y, (sp * np.log(a) * power - (sp - 1.0)) / power ** 2])
            dld = np.array([(power - 1.0) * np.sum(1.0 / sh), (power - 1.0) * np.sum(y / sh) + len(y) / scale, np.sum(np.log(a))])
        return dinv, dld


class _MappingInvSum:
    """MappingInvSum (processes/hypers/mappings.py:73-85): inv(y) = m1.inv(y) + m2.inv(y).  Its logdet_dinv is commented
    out in the reference, so the base class' numeric form applies (mappings.py:17-22): sum(log(diag(jacobian(inv)))) =
    sum(log(m1.inv'(y) + m2.inv'(y))); `__call__` is `pass` (the forward map does not exist in the reference: no predict)."""

    def __init__(self, spec):
        self.kind = "invsum"
        self.m1, self.m2 = make_mapping(spec["m1"]), make_mapping(spec["m2"])
        self.name = self.m1.name + " +^ " + self.m2.name
        self.hypers = tuple(self.m1.layout()) + tuple(self.m2.layout())
        self.n = 1

    def layout(self):
        return list(self.hypers)

    def n_theta(self):
        return self.m1.n_theta() + self.m2.n_theta()

    def _split(self, th):
        return th[:self.m1.n_theta()], th[self.m1.n_theta():]

    def inv(self, th, y):
        t1, t2 = self._split(th)
        return self.m1.inv(t1, y) + self.m2.inv(t2, y)

    def dinv_dy(self, th, y):
        t1, t2 = self._split(th)
        return self.m1.dinv_dy(t1, y) + self.m2.dinv_dy(t2, y)

    def logdet_dinv(self, th, y):
        with np.errstate(all="ignore"):
            return float(np.sum(np.log(self.dinv_dy(th, y))))

    def forward(self, th, z):
        raise NotImplementedError("MappingInvSum.__call__ is `pass` in the reference")

    def grads(self, th, y):
        """d inv / d h from the owning map; d logdet / d h = sum(d m.inv'(y) / d h / (m1.inv' + m2.inv')) with the
        hyper-derivative of inv' by 4th-order central differences (closed-form inv')."""
        t1, t2 = self._split(th)
        di1, _ = self.m1.grads(t1, y)
        di2, _ = self.m2.grads(t2, y)
        tot = self.dinv_dy(th, y)
        dld = np.zeros(len(th))
        for k in range(len(th)):
            h = 1e-4 * max(1.0, abs(th[k]))

            def at(s):
                t = np.array(th, dtype=np.float64)
                t[k] += s * h
                return self.dinv_dy(t, y)
            dld[k] = np.sum((-at(2) + 8 * at(1) - 8 * at(-1) + at(-2)) / (12 * h) / tot)
        return np.vstack([di1, di2]), dld


def make_mapping(spec):
    if spec["type"] == "composed":
        return _MappingComposed(spec)
    if spec["type"] == "invsum":
        return _MappingInvSum(spec)
    if spec["type"] == "BoxCoxLinear2":
        return _MappingBoxCoxLinear2(spec)
    return _Mapping(spec)
_MappingComposed.grads_fd = _Mapping.grads_fd


# --------------------------------------------------------------------------- processes
def build_process(spec, D, strict=True):
    if spec.get("kind") == "transport":
        return OracleTransportProcess(spec, D, strict)
    return OracleProcess(spec, D, strict)


class OracleProcess:
    """EllipticalProcess + Gaussian / Student-t distribution (elliptical.py, gaussian.py, studentT.py).

    spec = {"kind": "gauss"|"student", "name": "GP", "location": {...}, "kernel": {...},
            "mapping": {...}, "noisy": True}
    theta (flat, pymc3 bijection order = creation order, elliptical.py:35-52):
      [location..., kernel (+ Noise var)..., mapping..., degree]; positive hypers stored as logs
      (hypers/__init__.py:124-126,190-202: FlatExp = Flat prior + log transform, zero Jacobian,
      -inf when exp(theta) <= 1e-6).
    """

    def __init__(self, spec, D, strict=True):
        self.spec = spec
        self.kind = spec.get("kind", "gauss")
        self.D = D
        self.consts = Consts(strict)
        self.location = _Mean(spec.get("location", {"type": "Zero"}), D)
        self.f_kernel = build_kernel(spec["kernel"], D)
        self.noisy = spec.get("noisy", True)
        if self.noisy:                                         # elliptical.py:26-28
            self.k_noise = _Binary("sum", self.f_kernel, _Leaf({"type": "Noise", "name": "Noise"}, D))
        else:
            self.k_noise = self.f_kernel
        self.mapping = make_mapping(spec.get("mapping", {"type": "Identity"}))
        lay = self.location.layout() + self.k_noise.layout() + self.mapping.layout()
        if self.kind == "student":                             # hypers/__init__.py:151-155
            lay = lay + [_Hyper("Freedom_degree", 1, True)]
        self._layout = lay
        self.n_loc = self.location.n_theta()
        self.n_ker = self.k_noise.n_theta()
        self.n_map = self.mapping.n_theta()
        self.P = sum(h.size for h in lay)

    # ---- theta handling ------------------------------------------------------------------
    def layout(self):
        """[(name, size, positive)] in theta order."""
        return [(h.name, h.size, h.positive) for h in self._layout]

    def positive_mask(self):
        return np.concatenate([np.full(h.size, h.positive) for h in self._layout]) if self._layout else np.zeros(0, bool)

    def natural(self, theta):
        theta = np.asarray(theta, dtype=np.float64)
        m = self.positive_mask()
        return np.where(m, np.exp(np.where(m, theta, 0.0)), theta)

    def split(self, nat):
        a = self.n_loc
        b = a + self.n_ker
        c = b + self.n_map
        return nat[:a], nat[a:b], nat[b:c], nat[c:]

    def logprior(self, theta):
        """Sum of free-RV logp: Flat (0) + NonTransformLog.jacobian_det (hypers/__init__.py:200-201), plus the
        pm.Potential regularisers (stochastic.py:300-306 adds model.potentials for prior and posterior alike)."""
        nat = self.natural(theta)
        m = self.positive_mask()
        return (0.0 if np.all(nat[m] > 1e-6) else -np.inf) + self.potentials(nat)[0]

    def _potential_specs(self):
        """[(substring, reg, c, [hyper names])] from `"potential": [hypers, reg, c]` entries of the spec, with the hyper
        lists `check_potential` iterates over (hypers/__init__.py:94-109): a leaf kernel owns var (+ alpha, periodic
        freq / rate) but not the ARD metric's rate / bias; a composition owns its children's; means and mappings all."""
        out = []

        def kernel_hypers(sp):
            t = sp["type"]
            if t in ("sum", "prod", "max"):
                return kernel_hypers(sp["k1"]) + kernel_hypers(sp["k2"])
            if t in ("scale", "shift"):
                return kernel_hypers(sp["k"])
            name = sp.get("name", t)
            fixed = sp.get("var") is not None or (t in ("LIN", "POL") and "var" not in sp)
            hs = [] if fixed else [name + "_var"]
            if t == "RQ":
                hs.append(name + "_alpha")
            if t in ("SIN", "SM"):
                hs += [name + "_rate", name + "_freq"]
            if t in ("COS", "SINC"):
                hs.append(name + "_freq")
            return hs

        def walk(sp):
            t = sp["type"]
            if "potential" in sp:
                out.append(tuple(sp["potential"]) + (kernel_hypers(sp),))
            if t in ("sum", "prod", "max"):
                walk(sp["k1"]); walk(sp["k2"])
            elif t in ("scale", "shift"):
                walk(sp["k"])

        loc = self.spec.get("location", {"type": "Zero"})
        if "potential" in loc:
            out.append(tuple(loc["potential"]) + ([h.name for h in self.location.layout()],))
        walk(self.spec["kernel"])
        mp = self.spec.get("mapping", {"type": "Identity"})
        if "potential" in mp:
            out.append(tuple(mp["potential"]) + ([h.name for h in self.mapping.layout()],))
        return out

    def potentials(self, nat):
        """(value, d/d natural hypers)."""
        pname = self.spec.get("name", {"gauss": "GP", "student": "TP"}[self.kind])
        if self.spec.get("warped", self.mapping.kind != "Identity"):
            pname = self.spec.get("name", {"gauss": "WGP", "student": "WTP"}[self.kind])
        val, g = 0.0, np.zeros(self.P)
        off = {}
        o = 0
        for h in self._layout:
            off[h.name] = (o, h.size)
            o += h.size
        for sub, reg, c, names in self._potential_specs():
            for nm in names:
                if (pname + "_" + nm).find(sub) > 0:             # k.name.find(hypers) > 0 on the full variable name
                    a, n = off[nm]
                    v = nat[a:a + n]
                    if reg == "L1":
                        val += -c * np.sum(np.abs(v)); g[a:a + n] += -c * np.sign(v)
                    elif reg == "L2":
                        val += -c * np.sum(v ** 2); g[a:a + n] += -2.0 * c * v
        return val, g

    # ---- pieces --------------------------------------------------------------------------
    def cov_inputs(self, nat_k, X):
        """prior_kernel_inputs = tt_to_cov(f_kernel_noise.cov(inputs)) (elliptical.py:71)."""
        return tt_to_cov(self.k_noise.cov(nat_k, X, X, True), self.consts)

    def _core(self, theta, X, y):
        nat = self.natural(theta)
        t_loc, t_ker, t_map, t_deg = self.split(nat)
        K = self.cov_inputs(t_ker, X)
        L, info = cholesky_robust(K, self.consts, return_info=True)
        z = tt_to_num(self.mapping.inv(t_map, y))               # elliptical.py:63 (mapping_outputs)
        delta = self.mapping.inv(t_map, y) - self.location(t_loc, X)   # gaussian.py:208 (raw inv in logp_cho)
        det_m = self.mapping.logdet_dinv(t_map, y)
        return nat, K, L, info, z, delta, det_m

    def logp_terms(self, theta, X, y):
        """Returns dict of the terms of logp_cho (gaussian.py:193-241 / studentT.py:115-146)."""
        c = self.consts
        nat, K, L, info, z, delta, det_m = self._core(theta, X, y)
        n = float(X.shape[0])
        lcho = sla.solve_triangular(L, delta, lower=True)       # gaussian.py:212
        beta = float(lcho.dot(lcho))                            # gaussian.py:215
        logdet = float(np.sum(np.log(np.diag(L))))
        out = {"beta": beta, "logdet": logdet, "det_m": float(det_m), "info": info}
        if self.kind == "gauss":
            npi = -0.5 * n * c.log_2pi                          # gaussian.py:218
            out.update(npi=npi, dot2=-0.5 * beta, det_k=-logdet)
            r = npi + (-0.5 * beta) + (-logdet) + det_m
        else:
            nu = 2.0 + nat[-1]                                  # hypers/__init__.py:159-160 bound + degree
            r1 = -0.5 * (nu + n) * np.log1p(beta / (nu - 2.0))  # studentT.py:126
            if 1e6 <= nu:                                       # studentT.py:127
                r2 = -n * 0.5 * c.log_2pi_student
            else:                                               # studentT.py:128
                r2 = gammaln((nu + n) * 0.5) - gammaln(nu * 0.5) - 0.5 * n * np.log((nu - 2.0) * c.pi)
            out.update(r1=float(r1), r2=float(r2), r3=-logdet, nu=nu)
            r = r1 + r2 + (-logdet) + det_m
        bad = (not np.all(np.isfinite(delta)) or not np.isfinite(det_m)
               or not np.all(np.isfinite(L)) or not np.all(np.isfinite(lcho)))   # gaussian.py:234-241
        out["loglike"] = c.guard if bad else float(r)
        return out

    def loglike(self, theta, X, y):
        return self.logp_terms(theta, X, y)["loglike"]

    def logp(self, theta, X, y):
        """th_logp posterior (stochastic.py:300-306): free-RV terms + observed term."""
        return self.logprior(theta) + self.loglike(theta, X, y)

    # ---- gradient ------------------------------------------------------------------------
    def dlogp(self, theta, X, y, method="analytic", nan_quirk=False):
        """d logp / d theta (theta = transformed/log space), bijection order.

        method="analytic": G = 1/2 (c alpha alpha^T - K^-1)  
        method="murray":   reverse-mode through chol like the reference (tensors.py:224-260)
        Both contract G with the per-kernel dK/dtheta.  Result passed through tt_to_num
        (stochastic.py:308-309).
        """
        nat, K, L, info, z, delta, det_m = self._core(theta, X, y)
        t_loc, t_ker, t_map, t_deg = self.split(nat)
        n = float(X.shape[0])
        u = sla.solve_triangular(L, delta, lower=True)
        beta = float(u.dot(u))
        alpha = sla.solve_triangular(L.T, u, lower=False)
        if self.kind == "gauss":
            cfac = 1.0
        else:
            nu = 2.0 + nat[-1]
            cfac = (nu + n) / (nu - 2.0 + beta)
        if method == "analytic":
            Kinv = sla.cho_solve((L, True), np.eye(L.shape[0]))
            G = 0.5 * (cfac * np.outer(alpha, alpha) - Kinv)
        else:
            # dlogp/dL: quadratic term  d(-c/2 u'u)/dL = c * L^-T u u^T (lower), logdet term -1/L_ii
            Lbar = cfac * np.tril(np.outer(alpha, u)) - np.diag(1.0 / np.diag(L))
            Kbar = murray_cholesky_grad(L, Lbar)
            G = 0.5 * (Kbar + Kbar.T)                           # symmetric form of the tril gradient
        dK = self.k_noise.dcov(t_ker, X, X, True, nan_quirk)
        g_ker = np.array([np.sum(G * d) for d in dK])
        g_delta = -cfac * alpha
        g_loc = -(self.location.jac(t_loc, X) @ g_delta) if self.n_loc else np.zeros(0)
        if self.n_map:
            dinv, dld = self.mapping.grads(t_map, y)
            g_map = dinv @ g_delta + dld
        else:
            g_map = np.zeros(0)
        parts = [g_loc, g_ker, g_map]
        if self.kind == "student":
            bn = beta / (nu - 2.0)
            d_r1 = -0.5 * np.log1p(bn) + 0.5 * (nu + n) * bn / ((nu - 2.0) * (1.0 + bn))
            if 1e6 <= nu:
                d_r2 = 0.0
            else:
                d_r2 = 0.5 * digamma((nu + n) * 0.5) - 0.5 * digamma(nu * 0.5) - 0.5 * n / (nu - 2.0)
            parts.append(np.array([d_r1 + d_r2]))
        g_nat = np.concatenate(parts) + self.potentials(nat)[1]
        m = self.positive_mask()
        g = np.where(m, g_nat * nat, g_nat)                     # chain rule through exp (log-space hypers)
        return tt_to_num(g)

    # ---- posterior -----------------------------------------------------------------------
    def posterior(self, theta, Xs, X, y, noise=False, cov=False, solver="lu"):
        """elliptical.py:78-107.  solver="lu" follows the reference (tsl.solve = general LU);
        solver="chol" is the Cholesky route the CUDA path takes."""
        nat = self.natural(theta)
        t_loc, t_ker, t_map, t_deg = self.split(nat)
        Kxx = self.cov_inputs(t_ker, X)
        kern = self.k_noise if noise else self.f_kernel
        th_k = t_ker if (noise or not self.noisy) else t_ker[:self.f_kernel.n_theta()]
        Ksx = tt_to_num(kern.cov(th_k, Xs, X, False))           # elliptical.py:78-79
        Kss = kern.cov(th_k, Xs, Xs, True)
        if noise:
            Kss = tt_to_cov(Kss, self.consts)                   # elliptical.py:70
        delta = tt_to_num(self.mapping.inv(t_map, y)) - self.location(t_loc, X)
        if solver == "lu":
            a = sla.solve(Kxx, delta)
            B = sla.solve(Kxx, Ksx.T)
        else:
            Lc = cholesky_robust(Kxx, self.consts)
            a = sla.cho_solve((Lc, True), delta)
            B = sla.cho_solve((Lc, True), Ksx.T)
        mean = self.location(t_loc, Xs) + Ksx.dot(a)            # elliptical.py:81-84
        C = Kss - Ksx.dot(B)                                    # elliptical.py:86-92
        var = np.maximum(np.diag(C), 0.0)                       # elliptical.py:94-97 tt_to_bounded(.., 0)
        out = {"location": mean, "kernel_diag": var, "kernel_sd": np.sqrt(var)}
        if cov:
            out["kernel"] = C
        if self.kind == "student":                              # studentT.py:36-49
            Lc = cholesky_robust(Kxx, self.consts)
            al = sla.solve_triangular(Lc, delta, lower=True)
            beta = al.dot(al)
            nu = 2.0 + nat[-1]
            out["scaling"] = (nu + beta - 2.0) / (nu + X.shape[0] - 2.0)
        return out

    def predict(self, theta, Xs, X, y, noise=False, n_gh=10):
        """mean / variance / std / median / quantiles as `predict` assembles them
        (stochastic.py:488-513; gaussian.py:56-73,127-174; studentT.py:45-55)."""
        from scipy import stats
        nat = self.natural(theta)
        t_map = self.split(nat)[2]
        post = self.posterior(theta, Xs, X, y, noise=noise)
        mu, var = post["location"], post["kernel_diag"]
        T = lambda v: self.mapping.forward(t_map, v)
        # gaussian.py:115-174: only the Warped* classes integrate T by Gauss-Hermite; the plain classes use T(mu)
        warped = self.spec.get("warped", self.mapping.kind != "Identity")
        scaling = post.get("scaling", 1.0)
        if warped:                                              # gaussian.py:127-174 Gauss-Hermite n=10
            a, w = np.polynomial.hermite.hermgauss(n_gh)
            sd = np.sqrt(var)
            grille = mu[None, :] + sd[None, :] * np.sqrt(2.0) * a[:, None]
            mean = w.dot(T(grille.ravel()).reshape(grille.shape)) / np.sqrt(np.pi)
            m2 = w.dot((T(grille.ravel()) ** 2).reshape(grille.shape)) / np.sqrt(np.pi)
            variance = m2 - mean ** 2
        else:
            mean = T(mu)                                        # elliptical.py:194-196
            variance = var * scaling                            # studentT.py:45-46
        out = {"mean": mean, "variance": variance, "std": np.sqrt(variance), "median": T(mu)}
        if self.kind == "student":
            q = stats.t.ppf(0.975, df=2.0 + nat[-1] + X.shape[0])   # studentT.py:53 (freedom posterior)
        else:
            q = stats.norm.ppf(0.975)
        sd = np.sqrt(var)
        out["quantile_up"] = T(mu + q * sd)
        out["quantile_down"] = T(mu - q * sd)
        return out


# --------------------------------------------------------------------------- transport.py / transports.py
class OracleTransportProcess:
    """TransportGaussianProcess (processes/transport.py:135-246) over a chain of transports
    (hypers/transports.py), restated literally for chains  [ID | TMapping | TLocation]* @ TKernel.

    spec = {"kind": "transport", "chain": [{"t": "TMapping", "mapping": {...}}, {"t": "TLocation", "location": {...}},
            {"t": "ID"}, {"t": "TKernel", "kernel": {...}, "noisy": True}]}       (outermost first)
    theta follows the chain: `check_hypers` runs t1 then t2 of every composition (transports.py:76-79).
    At most one TMapping and one TLocation (enough for the tests; the product composes several).
    """

    def __init__(self, spec, D, strict=True):
        self.spec = spec
        self.D = D
        self.consts = Consts(strict)
        self.chain = spec["chain"]
        assert self.chain[-1]["t"] == "TKernel"
        self.parts = []                                          # (kind, object, theta offset, n_theta)
        off = 0
        lay = []
        for t in self.chain:
            if t["t"] == "ID":
                obj, hy = None, []
            elif t["t"] == "TMapping":
                obj = make_mapping(t["mapping"])
                hy = obj.layout()
            elif t["t"] == "TLocation":
                obj = _Mean(t["location"], D)
                hy = obj.layout()
            elif t["t"] == "TScale":                             # transports.py:165-181
                obj = _Mean(t["scale"], D)
                hy = obj.layout()
            elif t["t"] == "TKernel":
                k = build_kernel(t["kernel"], D)
                self.f_kernel = k
                self.noisy = bool(t.get("noisy", False))
                if self.noisy:                                   # transports.py:203-206
                    kname = t["kernel"].get("name", t["kernel"]["type"])
                    k = _Binary("sum", k, _Leaf({"type": "Noise", "name": "Noise" + kname}, D))
                self.k_noise = k
                obj, hy = k, k.layout()
            else:
                raise ValueError(t["t"])
            n = sum(h.size for h in hy)
            self.parts.append((t["t"], obj, off, n))
            lay += hy
            off += n
        self._layout = lay
        self.P = off

    def layout(self):
        return [(h.name, h.size, h.positive) for h in self._layout]

    def positive_mask(self):
        return np.concatenate([np.full(h.size, h.positive) for h in self._layout]) if self._layout else np.zeros(0, bool)

    def natural(self, theta):
        theta = np.asarray(theta, dtype=np.float64)
        m = self.positive_mask()
        return np.where(m, np.exp(np.where(m, theta, 0.0)), theta)

    def logprior(self, theta):
        nat = self.natural(theta)
        m = self.positive_mask()
        return 0.0 if np.all(nat[m] > 1e-6) else -np.inf

    # ---- one transport at a time -----------------------------------------------------------
    def _cov(self, th, x1, x2, noise):
        if noise and self.noisy:
            return self.k_noise.cov(th, x1, x1 if x2 is None else x2, x2 is None)
        return self.f_kernel.cov(th[:self.f_kernel.n_theta()], x1, x1 if x2 is None else x2, x2 is None)

    def _call1(self, part, nat, X, v, noise):
        kind, obj, off, n = part
        th = nat[off:off + n]
        if kind == "ID":
            return v                                             # transports.py:123-124
        if kind == "TMapping":
            return obj.forward(th, v)                            # :190-191
        if kind == "TLocation":
            return v + obj(th, X)                                # :152-156
        if kind == "TScale":
            return v * obj(th, X)                                # :171-172
        return cholesky_robust(self._cov(th, X, None, noise), self.consts).dot(v)     # :210-216

    def _inv1(self, part, nat, X, v, noise):
        kind, obj, off, n = part
        th = nat[off:off + n]
        if kind == "ID":
            return v
        if kind == "TMapping":
            return obj.inv(th, v)                                # :193-194
        if kind == "TLocation":
            return v - obj(th, X)                                # :158-159
        if kind == "TScale":
            return v / obj(th, X)                                # :174-175
        return sla.solve_triangular(cholesky_robust(self._cov(th, X, None, noise), self.consts), v, lower=True)  # :227-232

    def _logdet1(self, part, nat, X, v):
        kind, obj, off, n = part
        th = nat[off:off + n]
        if kind == "ID":
            return 1.0                                           # :129-130  tt.ones(()) -- not zero
        if kind == "TMapping":
            return obj.logdet_dinv(th, v)                        # :196-197
        if kind == "TLocation":
            return 0.0                                           # :161-162
        if kind == "TScale":
            return -float(np.sum(np.log(obj(th, X))))            # :177-181
        return -float(np.sum(np.log(np.diag(cholesky_robust(self._cov(th, X, None, True), self.consts)))))   # :234-236

    # ---- chain (TransportComposed, transports.py:93-119) ---------------------------------------
    def call(self, nat, X, v, noise):
        for part in reversed(self.parts):
            v = self._call1(part, nat, X, v, noise)
        return v

    def inv(self, nat, X, v, noise):
        for part in self.parts:
            v = self._inv1(part, nat, X, v, noise)
        return v

    def logdet_dinv(self, nat, X, v):
        tot = 0.0
        for part in self.parts:                                  # t2.logdet(t1.inv(y, noise=True)) + t1.logdet(y)
            tot = tot + self._logdet1(part, nat, X, v)
            v = self._inv1(part, nat, X, v, True)
        return tot

    def loglike(self, theta, X, y):
        """TransportGaussianDistribution.logp_t (transport.py:220-243)."""
        nat = self.natural(theta)
        c = self.consts
        delta = self.inv(nat, X, y, True)
        det_m = self.logdet_dinv(nat, X, y)
        r = -0.5 * float(len(y)) * c.log_2pi + (-0.5) * float(delta.dot(delta)) + det_m
        bad = not np.all(np.isfinite(delta)) or not np.isfinite(det_m)
        return c.guard if bad else float(r)

    def logp(self, theta, X, y):
        return self.logprior(theta) + self.loglike(theta, X, y)

    def dlogp(self, theta, X, y, nan_quirk=False):
        """Gradient through the equivalent warped GP (same density, permuted theta)."""
        if any(t["t"] == "TScale" for t in self.chain):          # no warped-GP equivalent: 4th-order central differences
            theta = np.asarray(theta, dtype=np.float64)
            g = np.zeros(self.P)
            for k in range(self.P):
                h = 1e-4 * max(1.0, abs(theta[k]))
                e = np.zeros(self.P)
                e[k] = h
                f = lambda t: self.logp(t, X, y)
                g[k] = (-f(theta + 2 * e) + 8 * f(theta + e) - 8 * f(theta - e) + f(theta - 2 * e)) / (12 * h)
            return g
        loc = next((t["location"] for t in self.chain if t["t"] == "TLocation"), {"type": "Zero"})
        mp = next((t["mapping"] for t in self.chain if t["t"] == "TMapping"), {"type": "Identity"})
        tk = self.chain[-1]
        wgp = OracleProcess({"kind": "gauss", "warped": True, "location": loc, "kernel": tk["kernel"], "mapping": mp,
                             "noisy": self.noisy}, self.D, self.consts.strict)
        # permutation: chain order -> [location, kernel(+Noise), mapping]
        src = {}
        for kind, obj, off, n in self.parts:
            src[kind] = (off, n)
        order = [src.get("TLocation", (0, 0)), src["TKernel"], src.get("TMapping", (0, 0))]
        idx = np.concatenate([np.arange(o, o + n) for o, n in order]).astype(int)
        g_w = wgp.dlogp(np.asarray(theta)[idx], X, y, nan_quirk=nan_quirk)
        g = np.zeros(self.P)
        g[idx] = g_w
        return g

    # ---- selectors (transport.py:34-100) ------------------------------------------------------
    def transport(self, theta, space, vector, X=None, y=None, prior=False, noise=False):
        nat = self.natural(theta)
        if prior:
            return self.call(nat, space, vector, noise)
        # TransportComposed.posterior: element-wise transports act on `space`, TKernel.posterior does the work
        pre = y
        for part in self.parts[:-1]:
            pre = self._inv1(part, nat, X, pre, True)
        kind, obj, off, n = self.parts[-1]
        th = nat[off:off + n]
        outputs_inv = sla.solve_triangular(cholesky_robust(self._cov(th, X, None, True), self.consts), pre, lower=True)
        cov_space_inputs = self.f_kernel.cov(th[:self.f_kernel.n_theta()], X, space, False)       # :246
        cov = np.block([[self._cov(th, X, None, True), cov_space_inputs],
                        [cov_space_inputs.T, self._cov(th, space, None, noise)]])
        cho = cholesky_robust(cov, self.consts)
        v = cho.dot(np.concatenate([outputs_inv, vector]))[len(y):]
        for part in reversed(self.parts[:-1]):
            v = self._call1(part, nat, space, v, noise)
        return v

    def transport_inv(self, theta, space, vector, X=None, y=None, prior=False, noise=False):
        if not prior:                     # `inv=True` is ignored by TKernel.posterior / TElemwise.posterior
            return self.transport(theta, space, vector, X, y, False, noise)
        return self.inv(self.natural(theta), space, vector, noise)

    def transport_diag(self, theta, space, vector, X=None, y=None, prior=False, noise=False):
        if not prior or len(self.parts) > 1:      # Transport.diag = __call__ except for a bare TKernel (:19-20,218-225)
            return self.transport(theta, space, vector, X, y, prior, noise)
        nat = self.natural(theta)
        return np.sqrt(np.diag(self._cov(nat, space, None, noise))) * vector


# --------------------------------------------------------------------------- synthetic inputs
def c1_inputs():
    rng = np.random.default_rng(0)
    x = np.linspace(0, 10, 200)[:, None]
    y = np.sin(x[:, 0]) + 0.1 * rng.standard_normal(200)
    return x, y


def c2_inputs(N=4096, B=64):
    rng = np.random.default_rng(1)
    X = rng.uniform(0, 10, size=(N, 3))
    f = np.sin(X[:, 0]) + np.cos(X[:, 1] / 2) + 0.1 * X[:, 2]
    y = f + 0.1 * rng.standard_normal(N)
    rng2 = np.random.default_rng(2)
    # theta-bar: Bias 0 ; SE var 1, rates 1,1,1 ; MAT52 var 0.5, rates .5,.5,.5 ; noise 0.05
    tbar = np.array([0.0, 1.0, 1.0, 1.0, 1.0, 0.5, 0.5, 0.5, 0.5, 0.05])
    Theta = np.tile(np.concatenate([[0.0], np.log(tbar[1:])]), (B, 1))
    Theta = Theta + 0.1 * rng2.standard_normal(Theta.shape)
    return X, y, Theta


def c3_inputs(N=2048, M=10000):
    rng = np.random.default_rng(3)
    x = np.sort(rng.uniform(0, 20, size=N))[:, None]
    f = np.sin(2 * np.pi * x[:, 0] / 5.0) * np.exp(-0.02 * x[:, 0]) + 0.05 * rng.standard_normal(N)
    y = np.exp(0.3 * f) + 0.5
    xs = np.linspace(0, 20, M)[:, None]
    return x, y, xs


def c4_inputs(N=16384, M=4096):
    rng = np.random.default_rng(4)
    X = rng.standard_normal((N, 5))
    f = np.sin(X[:, 0]) + 0.5 * X[:, 1] * X[:, 2] + np.cos(X[:, 3]) - 0.3 * X[:, 4]
    y = f + 0.1 * rng.standard_normal(N)
    Xs = rng.standard_normal((M, 5))
    return X, y, Xs


def c5_inputs(N=65536):
    rng = np.random.default_rng(5)
    X = rng.uniform(0, N ** (1.0 / 3.0), size=(N, 3))
    y = np.sin(X[:, 0]) + 0.1 * rng.standard_normal(N)
    return X, y
