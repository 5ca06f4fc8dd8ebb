"""`theano.gof.Op` wrappers of the C ABI — the boundary PyMC3 differentiates through.

Precedent in the reference: `CholeskyRobust(th.gof.Op)` (g3py/libs/tensors.py:174-263): `__props__`,
`make_node` -> `Apply`, `infer_shape`, `perform(node, inputs, output_storage)` on NumPy arrays, and a
symbolic `grad`.  The Ops here follow the same protocol; `perform` is one ctypes call into libg3b.so.

Theano is imported lazily (`build_ops()`); it is not part of this image, so the classes are produced by
a factory that takes the `theano` module (tests pass a minimal stand-in exposing `gof.Op`, `gof.Apply`
and `tensor.as_tensor_variable`).  Ops hold only picklable state (descriptor fields, kind, device): the
device context is looked up per process at `perform` time (stochastic.py:107-119 pickles whole processes).

Ops
---
GramOp(desc)(X1, X2, theta)            -> K                      grad: GramVJPOp
GramVJPOp(desc)(X1, X2, theta, W)      -> dtheta
CholeskyRobustGPU()(K)                 -> L                      (jitter ladder + 1e-10*I fallback; symbolic Murray grad)
GPLogpOp(desc, kind)(X, delta, theta[, nu]) -> (core, beta, logdet)
        core = -1/2 beta - logdet                         (gauss,   gaussian.py:219-224)
        core = -1/2 (nu+N) log1p(beta/(nu-2)) - logdet    (student, studentT.py:126,129)
        grad -> GPLogpGradOp(...) -> (dcore/dtheta, dcore/ddelta, dcore/dnu); X is disconnected
GPPosteriorOp(desc, noise)(X, Xs, delta, theta) -> (mean - m(X*), var)
"""
import numpy as np

from . import _cabi as cabi

_GUARD = float(np.float32(-1e30))
_FALLBACK = float(np.float32(1e-10))


def _desc_key(desc):
    """Hashable, picklable image of a g3_kernel_desc (for __props__ equality / pickling)."""
    return (desc.n_nodes, desc.n_theta) + tuple(
        (n.op, n.dim0, n.dim1, n.var_idx, n.p0_idx, n.p1_idx, n.flags, n.value) for n in list(desc.nodes)[:desc.n_nodes])


def _desc_from_key(key):
    d = cabi.KernelDesc()
    d.n_nodes, d.n_theta = key[0], key[1]
    for i, f in enumerate(key[2:]):
        n = d.nodes[i]
        n.op, n.dim0, n.dim1, n.var_idx, n.p0_idx, n.p1_idx, n.flags, n.value = f
    return d


def _ctx(device):
    from .processes import get_context
    return get_context(device)


def build_ops(theano=None):
    """Return a namespace with the Op classes bound to the given theano module (imported if None)."""
    if theano is None:
        import theano  # noqa: F811  (raises ImportError where Theano is absent)
    tt = theano.tensor
    Op, Apply = theano.gof.Op, theano.gof.Apply

    def as_var(x):
        return tt.as_tensor_variable(x)

    class _DescOp(Op):
        __props__ = ("desc_key", "device")

        def __init__(self, desc, device=0):
            self.desc_key = desc if isinstance(desc, tuple) else _desc_key(desc)
            self.device = int(device)

        @property
        def desc(self):
            return _desc_from_key(self.desc_key)

    class GramVJPOp(_DescOp):
        def make_node(self, x1, x2, theta, w):
            x1, x2, theta, w = as_var(x1), as_var(x2), as_var(theta), as_var(w)
            return Apply(self, [x1, x2, theta, w], [theta.type()])

        def infer_shape(self, node, shapes):
            return [shapes[2]]

        def perform(self, node, inputs, outputs):
            x1, x2, theta, w = inputs
            same = node.inputs[0] is node.inputs[1]           # cov(x1): Noise / WN contribute var * I
            outputs[0][0] = _ctx(self.device).gram_vjp(self.desc, x1, None if same else x2, theta, w)[0].astype(theta.dtype)

    class GramOp(_DescOp):
        """Kernel.cov(x1, x2) (kernels.py:106-110); pass x2 = x1 *as the same variable* for cov(x1)."""

        def make_node(self, x1, x2, theta):
            x1, x2, theta = as_var(x1), as_var(x2), as_var(theta)
            return Apply(self, [x1, x2, theta], [x1.type()])

        def infer_shape(self, node, shapes):
            return [(shapes[0][0], shapes[1][0])]

        def perform(self, node, inputs, outputs):
            x1, x2, theta = inputs
            same = node.inputs[0] is node.inputs[1]
            K, _ = _ctx(self.device).gram(self.desc, x1, None if same else x2, theta)
            outputs[0][0] = K[0].astype(x1.dtype)

        def grad(self, inputs, output_grads):
            x1, x2, theta = inputs
            g = GramVJPOp(self.desc_key, self.device)(x1, x2, theta, output_grads[0])
            return [theano.gradient.grad_undefined(self, 0, x1), theano.gradient.grad_undefined(self, 1, x2), g]

    class CholeskyRobustGPU(Op):
        """Drop-in for `cholesky_robust` (libs/tensors.py:174-263): same ladder, same fallback."""
        __props__ = ("lower", "destructive", "device")

        def __init__(self, device=0):
            self.lower, self.destructive, self.device = True, False, int(device)

        def infer_shape(self, node, shapes):
            return [shapes[0]]

        def make_node(self, x):
            x = as_var(x)
            assert x.ndim == 2
            return Apply(self, [x], [x.type()])

        def perform(self, node, inputs, outputs):
            x = inputs[0]
            L, info, _ = _ctx(self.device).potrf_robust(x)
            if info < 0:
                L = 0 * x + _FALLBACK * np.eye(len(x))
            outputs[0][0] = L.astype(x.dtype)

        def grad(self, inputs, gradients):
            """Reverse mode of the factor (Murray, arXiv:1602.07527), symbolic as in the reference's Op
            (libs/tensors.py:224-261) so that `tt.grad` through the bare factor keeps working when this Op replaces
            `cholesky_robust`:  Kbar = Psi(L^-T Phi(L^T Lbar) L^-1), Phi = lower triangle with the diagonal halved,
            Psi(S) = tril(S + S^T) - diag(S); NaN / inf scrubbed where the reference scrubs them (`tt_to_num`).
            The forward factor inside is this Op again (device); the two triangular solves are Theano's.  The
            differentiated hot path should use GPLogpOp instead: one fused call and no N x N cotangent."""
            import importlib
            tsl = importlib.import_module(theano.__name__ + ".tensor.slinalg")
            solve_upper = getattr(tsl, "solve_upper_triangular", None) or tsl.Solve(A_structure="upper_triangular", lower=False)
            zero, big = np.float32(0), np.float32(1e10)

            def scrub(r):
                return tt.switch(tt.isnan(r), zero, tt.switch(tt.isinf(r), big, r))
            L = self(inputs[0])
            P = scrub(tt.dot(L.T, gradients[0]))
            phi = tt.tril(P) - tt.diag(tt.diagonal(P) / 2.0)
            right = solve_upper(L.T, scrub(phi).T).T          # Phi L^-1
            S = solve_upper(L.T, right)                       # L^-T Phi L^-1
            return [tt.tril(S + S.T) - tt.diag(tt.diagonal(S))]

    class GPLogpGradOp(_DescOp):
        __props__ = ("desc_key", "kind", "device")

        def __init__(self, desc, kind, device=0):
            super().__init__(desc, device)
            self.kind = int(kind)

        def make_node(self, X, delta, theta, nu):
            X, delta, theta, nu = as_var(X), as_var(delta), as_var(theta), as_var(nu)
            return Apply(self, [X, delta, theta, nu], [theta.type(), delta.type(), nu.type()])

        def infer_shape(self, node, shapes):
            return [shapes[2], shapes[1], shapes[3]]

        def perform(self, node, inputs, outputs):
            X, delta, theta, nu = inputs
            ctx = _ctx(self.device)
            ctx.set_data_if_changed(X)                         # X is re-uploaded only when its values changed
            nu_a = np.atleast_1d(np.asarray(nu, dtype=np.float64)) if self.kind == cabi.KIND_STUDENT else None
            desc = self.desc
            hit = getattr(ctx, "_op_last", None)
            if hit is not None and ctx.resident_matches(desc, self.kind, delta, theta, nu_a):
                # GPLogpOp.perform just factored K for these very inputs (th_logp and th_dlogp are evaluated back to
                # back, stochastic.py:300-313): finish the gradient from the resident factor, one N^3 in total
                dth, ddl = ctx.gp_grad_resume()
                r = {"status": hit["status"], "beta": hit["beta"], "dtheta": dth, "ddelta": ddl}
                ctx.op_resumed = getattr(ctx, "op_resumed", 0) + 1
            else:
                r = ctx.gp_logp_grad(desc, self.kind, delta, theta, nu=nu_a, want_grad=True)
            ctx._op_last = None
            bad = bool(r["status"][0] & (cabi.ST_POTRF_FAILED | cabi.ST_NONFINITE_RESULT))
            dth = np.zeros_like(theta) if bad else r["dtheta"][0]
            ddl = np.zeros_like(delta) if bad else r["ddelta"][0]
            dnu = 0.0
            if self.kind == cabi.KIND_STUDENT and not bad:
                n, beta, v = float(len(delta)), float(r["beta"][0]), float(nu_a[0])
                bn = beta / (v - 2.0)
                dnu = -0.5 * np.log1p(bn) + 0.5 * (v + n) * bn / ((v - 2.0) * (1.0 + bn))
            outputs[0][0] = np.asarray(dth, dtype=theta.dtype)
            outputs[1][0] = np.asarray(ddl, dtype=delta.dtype)
            outputs[2][0] = np.asarray(dnu, dtype=np.asarray(nu).dtype)

    class GPLogpOp(_DescOp):
        __props__ = ("desc_key", "kind", "device")

        def __init__(self, desc, kind, device=0):
            super().__init__(desc, device)
            self.kind = int(kind)

        def make_node(self, X, delta, theta, nu=3.0):
            X, delta, theta, nu = as_var(X), as_var(delta), as_var(theta), as_var(nu)
            return Apply(self, [X, delta, theta, nu], [nu.type(), nu.type(), nu.type()])

        def infer_shape(self, node, shapes):
            return [(), (), ()]

        def perform(self, node, inputs, outputs):
            X, delta, theta, nu = inputs
            ctx = _ctx(self.device)
            ctx.set_data_if_changed(X)
            nu_a = np.atleast_1d(np.asarray(nu, dtype=np.float64)) if self.kind == cabi.KIND_STUDENT else None
            # the gradient Op of the same inputs normally follows (NUTS, BFGS): have U = L^-T prepared behind the factorisation
            spec = getattr(ctx, "set_speculate_grad", None)
            if spec is not None:
                spec(1)
            try:
                r = ctx.gp_logp_grad(self.desc, self.kind, delta, theta, nu=nu_a, want_grad=False)
            finally:
                if spec is not None:
                    spec(0)
            ctx._op_last = {"status": r["status"], "beta": r["beta"]}      # GPLogpGradOp may finish from this factor
            beta, logdet, st = float(r["beta"][0]), float(r["logdet"][0]), int(r["status"][0])
            n = float(len(delta))
            if st & cabi.ST_POTRF_FAILED:                      # L = 1e-10 * I (libs/tensors.py:218-222)
                beta, logdet = float(np.sum(np.square(delta))) / _FALLBACK ** 2, n * np.log(_FALLBACK)
            if self.kind == cabi.KIND_STUDENT:
                v = float(nu_a[0])
                core = -0.5 * (v + n) * np.log1p(beta / (v - 2.0)) - logdet
            else:
                core = -0.5 * beta - logdet
            if (st & cabi.ST_NONFINITE_RESULT and not st & cabi.ST_POTRF_FAILED) or not np.isfinite(core):
                core = _GUARD                                   # gaussian.py:234-241
            dt = np.asarray(nu).dtype
            outputs[0][0] = np.asarray(core, dtype=dt)
            outputs[1][0] = np.asarray(beta, dtype=dt)
            outputs[2][0] = np.asarray(logdet, dtype=dt)

        def connection_pattern(self, node):
            # inputs X, delta, theta, nu  x  outputs core, beta, logdet: only `core` carries gradients
            return [[False, False, False], [True, False, False], [True, False, False], [True, False, False]]

        def grad(self, inputs, output_grads):
            X, delta, theta, nu = inputs
            g = output_grads[0]
            dth, ddl, dnu = GPLogpGradOp(self.desc_key, self.kind, self.device)(X, delta, theta, nu)
            return [theano.gradient.DisconnectedType()(), g * ddl, g * dth, g * dnu]

    class GPPosteriorOp(_DescOp):
        __props__ = ("desc_key", "noise", "device")

        def __init__(self, desc, noise=False, device=0):
            super().__init__(desc, device)
            self.noise = bool(noise)

        def make_node(self, X, Xs, delta, theta):
            X, Xs, delta, theta = as_var(X), as_var(Xs), as_var(delta), as_var(theta)
            return Apply(self, [X, Xs, delta, theta], [delta.type(), delta.type()])

        def infer_shape(self, node, shapes):
            return [(shapes[1][0],), (shapes[1][0],)]

        def perform(self, node, inputs, outputs):
            X, Xs, delta, theta = inputs
            ctx = _ctx(self.device)
            ctx.set_data_if_changed(X)
            r = ctx.gp_posterior(self.desc, Xs, delta, theta, noise=self.noise, cov=False)
            outputs[0][0] = r["mean"].astype(delta.dtype)
            outputs[1][0] = r["var"].astype(delta.dtype)

    class _NS:
        pass
    ns = _NS()
    for c in (GramOp, GramVJPOp, CholeskyRobustGPU, GPLogpOp, GPLogpGradOp, GPPosteriorOp):
        # module-level identity so that pickled graphs (stochastic.py:107-119) find the classes again
        c.__qualname__ = c.__name__
        c.__module__ = __name__
        globals()[c.__name__] = c
        setattr(ns, c.__name__, c)
    return ns


_OP_NAMES = ("GramOp", "GramVJPOp", "CholeskyRobustGPU", "GPLogpOp", "GPLogpGradOp", "GPPosteriorOp")


def __getattr__(name):
    # `from g3py_b200.theano_ops import GPLogpOp` (and unpickling) build the classes against the real Theano
    if name in _OP_NAMES:
        build_ops()
        return globals()[name]
    raise AttributeError(name)
