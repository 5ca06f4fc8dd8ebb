"""TransportGaussianProcess — mirror of g3py/processes/transport.py:17-246.

y = t_1(t_2(... t_n(eps))), eps ~ N(0, I).  For chains `[ID | TMapping | TLocation]* @ TKernel` the density
(`TransportGaussianDistribution.logp_t`, transport.py:220-243)

    logp = -N/2 log(2 pi) - 1/2 |L^-1 (T^-1(y) - m(X))|^2 - sum log L_ii + log|dT^-1| (+ 1 per ID)

is the warped-GP density, so logp / dlogp run through the same device pipeline as `WGP` (Gram -> blocked fp64
Cholesky -> triangular solve -> trtri/lauum gradient); only the hyper-parameter ORDER differs (the transport chain
creates hypers outermost first: mapping, location, kernel, `Noise<kernel>`), and the sampling selectors
`transport / transport_inv / transport_diag` push a white-noise vector through the prior or the posterior.  The
posterior needs the Cholesky factor of the (N+M) x (N+M) joint covariance (`transports.py:236-257`), factored on
the device by `g3_potrf_robust`.
"""
import numpy as np

from . import _cabi as cabi
import contextlib

from .hypers.mappings import Identity, Mapping, MappingComposed
from .hypers.means import MeanSum, Zero
from .hypers.transports import ID, TKernel, TLocation, TMapping, TScale, Transport
from .processes import DictObj, EllipticalProcess, StochasticProcess

__all__ = ["TransportGaussianProcess", "TGP"]


class _ScaleWarp(Mapping):
    """TScale seen as a warping (transports.py:165-181): y = z * s(x), inv = y / s(x), log|d inv| = -sum log s(x).  It depends
    on the inputs, which the Mapping interface does not carry: the process sets `.X` to the points the vector lives on (the
    observed inputs for the density, `space` for a transported draw) around every use."""

    elementwise_forward = False          # tied to the points in `.X`: the vector cannot be cut into blocks

    def __init__(self, scale):
        self.scale = scale
        self.hypers = []
        self.name = scale.name
        self.dims = None
        self.shape = None
        self.potential = None
        self.X = None

    def check_hypers(self, parent="", reg=None):
        self.scale.check_hypers(parent, reg)
        self.hypers = list(self.scale.hypers)

    def check_dims(self, x=None):
        self.scale.check_dims(x)

    def default_hypers_dims(self, x=None, y=None):
        return self.scale.default_hypers_dims(x, y)

    def _s(self, p):
        if self.X is None:
            raise RuntimeError("TScale evaluated without its inputs")
        return self.scale(self.X, p)

    def __call__(self, z, p):
        return z * self._s(p)

    def inv(self, y, p):
        return y / self._s(p)

    def logdet_dinv(self, y, p):
        return -float(np.sum(np.log(self._s(p))))

    def dinv_dy(self, y, p):
        return 1.0 / self._s(p)

    def dlog_dinv_dy(self, y, p):
        return np.zeros(len(y))

    def grads(self, y, p):
        s = self._s(p)
        dinv, dld = {}, {}
        for h, J in self.scale.jacobian(self.X, p).items():       # J = d s / d h: (N,) or (size, N)
            J = np.asarray(J, dtype=np.float64)
            dinv[h] = -(y / (s * s)) * J
            dld[h] = -np.sum(J / s, axis=-1)
        return dinv, dld


class TransportGaussianProcess(EllipticalProcess):
    KIND = cabi.KIND_GAUSS
    WARPED = True

    def __init__(self, space=None, transport=None, *args, **kwargs):
        kwargs.setdefault("name", "TGP")
        if not isinstance(transport, Transport):
            raise TypeError("TransportGaussianProcess(space, transport): transport must be a Transport")
        chain = transport.chain()
        if not isinstance(chain[-1], TKernel) or any(isinstance(t, TKernel) for t in chain[:-1]):
            raise NotImplementedError("supported chains: [ID | TMapping | TLocation]* @ TKernel (one kernel, innermost)")
        kinds = [type(t) for t in chain[:-1]]
        if any(k not in (ID, TMapping, TLocation, TScale) for k in kinds):
            raise NotImplementedError("only ID, TMapping, TScale and TLocation may precede the TKernel")
        order = [k for k in kinds if k is not ID]
        if TLocation in order and TMapping in order[order.index(TLocation):]:
            raise NotImplementedError("TMapping inside a TLocation is not supported (put the mappings outermost)")
        if order.count(TScale) > 1 or (TScale in order and (TMapping in order[order.index(TScale):] or
                                                            TLocation in order[:order.index(TScale)])):
            raise NotImplementedError("supported: at most one TScale, inside the TMappings and outside the TLocations")
        self.f_transport = transport
        self.n_id = kinds.count(ID)
        maps = [t.mapping for t in chain[:-1] if isinstance(t, TMapping)]
        self._scale_warp = None
        for t in chain[:-1]:
            if isinstance(t, TScale):
                self._scale_warp = _ScaleWarp(t.scale)
                maps.append(self._scale_warp)                       # innermost warping: inv = scale.inv(maps.inv(y))
        locs = [t.location for t in chain[:-1] if isinstance(t, TLocation)]
        self.f_mapping = Identity()
        if maps:
            self.f_mapping = maps[0]
            for m in maps[1:]:
                self.f_mapping = MappingComposed(self.f_mapping, m)     # inv = m2.inv(m1.inv(y)), transports.py:108-109
        self.f_location = Zero()
        if locs:
            self.f_location = locs[0]
            for m in locs[1:]:
                self.f_location = MeanSum(self.f_location, m)
        tk = chain[-1]
        self.f_degree = None
        self.f_kernel = tk.kernel
        self.noisy = tk.is_noisy
        self.f_kernel_noise = tk.noisy
        kwargs["space"] = space
        StochasticProcess.__init__(self, *args, **kwargs)

    def _check_hypers(self):
        # transport.py:24-27: hypers are created along the chain, outermost transport first
        x = self.inputs
        parent = self.name + "_"
        self.f_transport.check_dims(x)
        self.f_transport.check_hypers(parent, self.registry)
        for comp in (self.f_location, self.f_mapping):        # composites collect the hypers created above
            comp.check_dims(x)
            comp.check_hypers(parent, self.registry)
        # transport.py:27 calls f_transport.check_potential(), which Transport does not forward to its parametrics:
        # potentials set on the pieces of a transport are never registered in the reference
        self._finish_layout()

    @contextlib.contextmanager
    def _scale_at(self, X):
        """The points a TScale is evaluated on while warpings act on a vector living on X."""
        if self._scale_warp is None:
            yield
            return
        keep = self._scale_warp.X
        self._scale_warp.X = X
        try:
            yield
        finally:
            self._scale_warp.X = keep

    def _host_terms(self, nat2d, inputs, outputs, want_grad):
        with self._scale_at(inputs):
            return super()._host_terms(nat2d, inputs, outputs, want_grad)

    def _eval_batch(self, Theta, inputs=None, outputs=None, want_grad=True, nan_quirk=None):
        ll, g, info = super()._eval_batch(Theta, inputs, outputs, want_grad, nan_quirk)
        if self.n_id:                                         # ID.logdet_dinv = tt.ones(()) (transports.py:129-130)
            ll = np.where(ll == self.consts.guard, ll, ll + float(self.n_id))
        return ll, g, info

    # ---- transports of a white-noise vector (transport.py:34-100) ------------------------------
    def _gram(self, X1, X2, nat, noise):
        if noise and self.noisy:
            K, _ = self.ctx.gram(self.desc, X1, X2, self._kernel_theta(nat[None, :]))
        else:
            K, _ = self.ctx.gram(self.desc_f, X1, X2, self._kernel_theta(nat[None, :], self._slots_f, self.desc_f.n_theta))
        return K[0]

    def _chol(self, K):
        L, info, _ = self.ctx.potrf_robust(K)                 # CholeskyRobust incl. ladder and 1e-10*I fallback
        return self.consts.fallback * np.eye(len(K)) if info < 0 else np.tril(L)

    def _chol_solve(self, K, rhs):
        """tsl.solve_lower_triangular(cholesky_robust(K), rhs) (transports.py:227-232), both steps on the device."""
        L, info, _, u = self.ctx.potrf_robust_solve(K, rhs)
        return rhs / self.consts.fallback if info < 0 else u

    def _tk_posterior(self, space, pred, nat, p, noise_pred):
        """TKernel.posterior (transports.py:236-257) with noise_obs=True."""
        X, y = self.inputs, self.outputs
        with np.errstate(all="ignore"), self._scale_at(X):
            pre = self.f_mapping.inv(y, p) - self.f_location(X, p)
        Kxx = self._gram(X, None, nat, True)
        u = self._chol_solve(Kxx, pre)
        Kxs = self._gram(X, space, nat, False)                # kernel.cov(inputs, space): no noise on the cross block
        joint = np.block([[Kxx, Kxs], [Kxs.T, self._gram(space, None, nat, noise_pred)]])
        L = self._chol(joint)
        n = len(y)
        return L[n:, :n] @ u + L[n:, n:] @ pred

    def _transport(self, which, params, space, inputs, outputs, vector, prior, noise, array, _given=False):
        if inputs is not None or outputs is not None:          # per-call substitution (stochastic.py:385-430)
            with self._substituted(inputs, outputs):
                return self._transport(which, params, space, None, None, vector, prior, noise, array, _given=True)
        theta = self._theta(params, array)
        if not self.is_observed and not _given:
            prior = True
        space = self.space if space is None else np.asarray(space, dtype=np.float64)
        if space.ndim < 2:
            space = space.reshape(len(space), 1)
        v = np.asarray(vector, dtype=np.float64).reshape(-1)
        if len(v) != len(space):
            raise ValueError("vector must have one entry per point of space")
        nat = self.natural(theta)
        p = self._accessor(nat)
        elementwise = len(self.f_transport.chain()) > 1
        self.executed["predict"] += 1
        if not prior:
            # TKernel.posterior and TElemwise.posterior ignore their `inv` / `diag` flags: the three posterior
            # selectors coincide in the reference (transports.py:134-136,236-257)
            post = self._tk_posterior(space, v, nat, p, noise)
            with self._scale_at(space):
                return self.f_mapping(self.f_location(space, p) + post, p)
        if which == "inv":                                    # transports.py:227-232 after the element-wise inverses
            with np.errstate(all="ignore"), self._scale_at(space):
                pre = self.f_mapping.inv(v, p) - self.f_location(space, p)
            return self._chol_solve(self._gram(space, None, nat, noise), pre)
        K = self._gram(space, None, nat, noise)
        if which == "diag" and not elementwise:               # TKernel.diag (transports.py:218-225)
            return np.sqrt(np.diag(K)) * v
        with self._scale_at(space):
            return self.f_mapping(self.f_location(space, p) + self._chol(K) @ v, p)

    def transport(self, params=None, space=None, inputs=None, outputs=None, vector=None, prior=False, noise=False,
                  array=False):
        return self._transport("call", params, space, inputs, outputs, vector, prior, noise, array)

    def transport_inv(self, params=None, space=None, inputs=None, outputs=None, vector=None, prior=False, noise=False,
                      array=False):
        return self._transport("inv", params, space, inputs, outputs, vector, prior, noise, array)

    def transport_diag(self, params=None, space=None, inputs=None, outputs=None, vector=None, prior=False,
                       noise=False, array=False):
        return self._transport("diag", params, space, inputs, outputs, vector, prior, noise, array)

    # ---- Monte-Carlo summaries (transport.py:171-211): 30 transported white-noise draws by default -------------
    def sampler(self, params=None, space=None, inputs=None, outputs=None, samples=1, prior=False, noise=False,
                array=False, rng=None):
        rng = rng or np.random.default_rng()
        space = self.space if space is None else np.asarray(space, dtype=np.float64).reshape(len(space), -1)
        rand = rng.standard_normal((len(space), samples))
        return np.array([self.transport(params, space, inputs, outputs, vector=rand[:, i], prior=prior, noise=noise,
                                        array=array) for i in range(samples)]).T

    def _sims(self, simulations, params, space, inputs, outputs, prior, noise, array, rng):
        if simulations is None:
            simulations = 30
        if isinstance(simulations, int):
            return self.sampler(params, space, inputs, outputs, samples=simulations, prior=prior, noise=noise,
                                array=array, rng=rng)
        return np.asarray(simulations)

    def mean(self, params=None, space=None, inputs=None, outputs=None, vector=None, prior=False, noise=False,
             array=False, simulations=None, rng=None):
        return self._sims(simulations, params, space, inputs, outputs, prior, noise, array, rng).mean(axis=1)

    def std(self, params=None, space=None, inputs=None, outputs=None, vector=None, prior=False, noise=False,
            array=False, simulations=None, rng=None):
        return self._sims(simulations, params, space, inputs, outputs, prior, noise, array, rng).std(axis=1)

    def quantiler(self, params=None, space=None, inputs=None, outputs=None, q=0.975, prior=False, noise=False,
                  simulations=None, array=False, rng=None):
        s = self._sims(simulations, params, space, inputs, outputs, prior, noise, array, rng)
        return np.nanpercentile(s, 100 * q, axis=1)

    def predict(self, params=None, space=None, inputs=None, outputs=None, mean=True, std=True, var=False, cov=False,
                median=False, quantiles=False, quantiles_noise=False, samples=0, distribution=False, prior=False,
                noise=False, simulations=None, array=False, rng=None):
        """stochastic.py:444-513 with the Monte-Carlo selectors above (one set of draws shared by all summaries)."""
        if not self.is_observed:
            prior = True
        sims = self._sims(simulations, params, space, inputs, outputs, prior, noise, array, rng)
        values = DictObj()
        if mean:
            values["mean"] = sims.mean(axis=1)
        if std:
            values["std"] = sims.std(axis=1)
        if var:
            values["variance"] = sims.var(axis=1)
        if median:
            values["median"] = np.nanpercentile(sims, 50, axis=1)
        if quantiles:
            values["quantile_up"] = np.nanpercentile(sims, 97.5, axis=1)
            values["quantile_down"] = np.nanpercentile(sims, 2.5, axis=1)
        if samples > 0:
            values["samples"] = self.sampler(params, space, inputs, outputs, samples=samples, prior=prior, noise=noise,
                                             array=array, rng=rng)
        return values


TGP = TransportGaussianProcess
